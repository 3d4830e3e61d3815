"""TheJoker: the orchestrator of thejoker/thejoker.py with the same public methods
(``marginal_ln_likelihood``, ``rejection_sample``, ``iterative_rejection_sample``),
driving per-GPU shards (sharding.DeviceEngine) instead of a schwimmbad pool.

RNG contract: ``self.rng`` is a numpy Generator and is
consumed exactly as the reference consumes it --
  * in_memory=True  (likelihood_helpers.py:91-229): ``rng.uniform(size=n)`` once
    (once per round over all accumulated lls for the iterative sampler), then the
    linear-parameter draws from the same generator;
  * in_memory=False (multiproc_helpers.py:185-427): optional
    ``rng.choice(n_total, size=n, replace=False)``, ``rng.uniform(size=n)``, then
    linear draws from per-task child generators spawned from the generator's seed
    sequence (multiproc_helpers.py:49-54).
On a PCG64 generator the n uniforms are produced on the GPU from the generator's
state, bit-identical to ``rng.uniform(size=n)``, and the host generator is advanced
by n; other bit generators fall back to drawing on the host and uploading.
"""
from __future__ import annotations

import logging
import os
import warnings

import numpy as np

from . import units as u
from .data_helpers import validate_prepare_data
from .helper import CJokerHelper
from .prior import JokerPrior
from .samples import JokerSamples
from .sharding import DeviceEngine, batch_tasks

__all__ = ["TheJoker"]

logger = logging.getLogger("thejoker_b200")


def _is_pcg64(rng):
    return type(rng.bit_generator).__name__ == "PCG64"


class TheJoker:
    """A custom Monte-Carlo sampler for two-body systems (thejoker.py:23-85).

    Parameters
    ----------
    prior : JokerPrior
    pool : optional object with ``map`` and ``close`` (accepted for signature
        compatibility; the work is sharded over ``devices`` instead)
    rng : numpy.random.Generator (optional)
    tempfile_path : str (optional, unused: no temporary cache file is written)
    devices : list of CUDA device indices (default: [LOCAL_RANK or 0])
    jitter_mode : "apply" | "reference" -- see CJokerHelper
    draw : "auto" | "numpy" | "device" -- how linear parameters are drawn for accepted
        samples (CJokerHelper.batch_get_posterior_samples); "auto" reproduces the
        reference's numbers for up to 4096 accepted samples
    group : torch.distributed process group (optional).  SPMD use under torchrun: every
        rank constructs the same TheJoker (same prior samples, same rng seed) and calls
        the same method; each rank evaluates only its contiguous shard of the prior
        cache on its GPU, the max is combined with an integer MAX all-reduce (NCCL) and
        the accepted indices are gathered, so every rank returns identical samples.
    """

    def __init__(self, prior, pool=None, rng=None, tempfile_path=None, devices=None,
                 jitter_mode="apply", group=None, draw="auto"):
        if pool is not None and (not hasattr(pool, "map") or not hasattr(pool, "close")):
            raise TypeError("Input pool object must have .map() and .close() methods.")
        self.pool = pool
        if rng is None:
            rng = np.random.default_rng()
        elif not isinstance(rng, np.random.Generator):
            raise TypeError("The input random number generator must be a "
                            "numpy.random.Generator instance.")
        self.rng = rng
        if not isinstance(prior, JokerPrior):
            raise TypeError("The input prior must be a JokerPrior instance.")
        self.prior = prior
        if tempfile_path is None:
            self._tempfile_path = os.path.expanduser("~/.thejoker/")
        else:
            self._tempfile_path = os.path.abspath(os.path.expanduser(tempfile_path))
        if devices is None:
            devices = [int(os.environ.get("LOCAL_RANK", "0"))]
        self.devices = list(devices)
        self.jitter_mode = jitter_mode
        self.group = group
        self.draw = draw  # see CJokerHelper.batch_get_posterior_samples
        self.last_stats = {}

    @property
    def tempfile_path(self):
        os.makedirs(self._tempfile_path, exist_ok=True)
        return self._tempfile_path

    # -- helpers ----------------------------------------------------------------
    def _make_joker_helper(self, data, device=None):
        """thejoker.py:87-91."""
        all_data, ids, trend_M = validate_prepare_data(data, self.prior.poly_trend,
                                                       self.prior.n_offsets)
        device = self.devices[0] if device is None else device
        return CJokerHelper(all_data, self.prior, trend_M, device=device,
                            jitter_mode=self.jitter_mode)

    def _columns(self, helper, prior_samples):
        """[P, e, omega, M0, s] host columns in internal units + optional ln_prior."""
        if isinstance(prior_samples, str):
            from .cache import PriorCache

            if os.path.isdir(prior_samples):
                prior_samples = PriorCache(prior_samples)
            else:  # .npz container or the reference's HDF5: told apart by the file's bytes
                prior_samples = JokerSamples.read(prior_samples)
        if type(prior_samples).__name__ == "PriorCache":
            cols = prior_samples.columns(rv_unit=helper.internal_units["s"])
            ln_prior = prior_samples.ln_prior() if prior_samples.has_ln_prior else None
            return cols, ln_prior
        if isinstance(prior_samples, JokerSamples):
            names = list(helper.packed_order)
            if prior_samples._uniform_s and len(prior_samples):
                # constant jitter: convert one element, not a 2^28-row column
                cols = prior_samples.columns(units=helper.internal_units, names=names[:4])
                cols.append(float(prior_samples["s"][:1].to_value(helper.internal_units["s"])[0]))
            else:
                cols = prior_samples.columns(units=helper.internal_units, names=names)
            ln_prior = prior_samples["ln_prior"].value if "ln_prior" in prior_samples else None
            return cols, ln_prior
        arr = np.asarray(prior_samples, dtype=np.float64)
        if arr.ndim != 2 or arr.shape[1] != 5:
            raise ValueError("packed prior samples must have shape (n, 5)")
        return [np.ascontiguousarray(arr[:, i]) for i in range(5)], None

    _warned_jitter = False

    def _warn_jitter(self, cols):
        """One warning per process when a non-zero jitter meets jitter_mode="apply": the
        reference's compiled likelihood ignores s (fast_likelihood.pyx:458 stores the inflated
        ivar and never reads it), so results differ from it by design."""
        if TheJoker._warned_jitter or self.jitter_mode != "apply" or len(cols) < 5:
            return
        s = cols[4]
        nonzero = bool(s != 0) if np.ndim(s) == 0 else bool(len(s) and np.any(np.asarray(s[:64]) != 0))
        if nonzero:
            TheJoker._warned_jitter = True
            warnings.warn(
                "prior samples carry a non-zero jitter s and jitter_mode='apply': s is added "
                "to the variance of every epoch (the documented model).  adrn/thejoker's "
                "compiled likelihood ignores s (fast_likelihood.pyx:458), so log-likelihoods "
                "and accepted samples differ from it; pass jitter_mode='reference' to "
                "reproduce the reference bit for bit.", stacklevel=3)

    def _engine(self, data, cols, helper0=None, cyclic=False):
        self._warn_jitter(cols)
        # helper0: the helper the caller already built to read units / columns; reused
        # for its device instead of creating a second handle for the same star
        helpers = {} if helper0 is None else {helper0.device: helper0}

        def make(d, fresh=False):
            if fresh:  # a second shard on the same device: its own handle
                return self._make_joker_helper(data, device=d)
            if d not in helpers:
                helpers[d] = self._make_joker_helper(data, device=d)
            return helpers[d]

        if self.group is None:
            eng = DeviceEngine(make, cols, devices=self.devices)
        else:
            import torch.distributed as dist

            from .sharding import shard_ranges

            n = len(cols[0])
            if cyclic:
                # every range the iterative sampler evaluates is split over all the ranks
                eng = DeviceEngine(make, cols, devices=self.devices[:1], group=self.group,
                                   global_offset=0, global_size=n, cyclic=True)
                return eng, make(self.devices[0])
            rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
            lo, hi = shard_ranges(n, world)[rank]
            local = [c if np.ndim(c) == 0 else c[lo:hi] for c in cols]
            eng = DeviceEngine(make, local, devices=self.devices[:1], group=self.group,
                               global_offset=lo, global_size=n)
        return eng, make(self.devices[0])

    def _engine_device_prior(self, data, n):
        """Engine over n prior samples *drawn on the GPUs* by the library's counter-based
        sampler (csrc/prior_gen.cuh): the likelihood kernel generates them in registers,
        nothing of the prior is stored on the host or in HBM.  None if the prior has a
        distribution family without a device sampler."""
        helpers = {}

        def make(d, fresh=False):
            if fresh:
                return self._make_joker_helper(data, device=d)
            if d not in helpers:
                helpers[d] = self._make_joker_helper(data, device=d)
            return helpers[d]

        helper0 = make(self.devices[0])
        # the seed is taken from self.rng on every rank alike (SPMD contract): sample g is
        # a function of (seed, g), so every sharding of [0, n) sees the same prior
        gen = self.prior.device_generator(int(self.rng.integers(0, 2**62)),
                                          helper0.internal_units["s"])
        if gen is None:
            return None
        from .sharding import shard_ranges

        rank, world, devices = 0, 1, self.devices
        if self.group is not None:
            import torch.distributed as dist

            rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
            devices = self.devices[:1]
        r_lo, r_hi = shard_ranges(n, world)[rank]
        eng = DeviceEngine.from_generator(make, gen, r_hi - r_lo, devices=devices, group=self.group,
                                          global_offset=r_lo, global_size=n)
        return eng, helper0

    @staticmethod
    def _rows(cols, idx):
        out = np.empty((len(idx), 5))
        for j, c in enumerate(cols):
            out[:, j] = c if np.ndim(c) == 0 else c[idx]
        return out

    def _uniform_accept(self, eng, n_accum, max_keep):
        """One accept pass over the first n_accum samples, consuming
        ``rng.uniform(size=n_accum)`` from self.rng."""
        if _is_pcg64(self.rng):
            idx, total, near = eng.accept(self.rng, hi=n_accum, max_keep=max_keep)
            self.rng.bit_generator.advance(int(n_accum))
        else:
            uu = self.rng.uniform(size=n_accum)
            idx, total, near = eng.accept(self.rng, hi=n_accum, max_keep=max_keep, uniforms=uu)
        self.last_stats.update(n_accepted=total, n_near_threshold=near, ll_max=eng.max_value(),
                               n_nonfinite=int(getattr(eng, "last_nonfinite", 0)))
        return idx, total

    def _full_samples(self, helper, rows, rng, n_linear_samples, in_memory, n_batches):
        """make_full_samples_inmem (likelihood_helpers.py:69-88) / make_full_samples
        (multiproc_helpers.py:150-182, with per-task child generators)."""
        if in_memory:
            raw, _ = helper.batch_get_posterior_samples(rows, n_linear_samples, rng, draw=self.draw)
        else:
            n_batches = 1 if n_batches is None else n_batches  # max(1, SerialPool.size)
            tasks = batch_tasks(len(rows), n_batches, arr=rows)
            sg = rng.bit_generator._seed_seq.spawn(len(tasks))
            parts = []
            for task, seq in zip(tasks, sg):
                child = np.random.Generator(np.random.PCG64(seq))
                parts.append(helper.batch_get_posterior_samples(task[0], n_linear_samples, child,
                                                                draw=self.draw)[0])
            raw = np.concatenate(parts) if parts else np.zeros((0, 5 + helper.n_linear))
        return JokerSamples.unpack(raw, helper.internal_units, t_ref=helper.data.t_ref,
                                   poly_trend=self.prior.poly_trend, n_offsets=self.prior.n_offsets)

    # -- public API -------------------------------------------------------------
    def marginal_ln_likelihood(self, data, prior_samples, n_batches=None, in_memory=False):
        """Marginal log-likelihood at each prior sample (thejoker.py:93-138).  Returns
        a numpy array; the computation runs on the configured GPUs."""
        helper0 = self._make_joker_helper(data)
        cols, _ = self._columns(helper0, prior_samples)
        if self.group is None and len(self.devices) == 1:
            # host columns in, host ll out: copies and kernel pipelined in the library
            s = cols[4]
            return helper0.marginal_ln_likelihood_columns(
                *cols[:4], s=None if np.ndim(s) == 0 else s,
                s_const=float(s) if np.ndim(s) == 0 else 0.0)
        eng, _ = self._engine(data, cols, helper0)
        eng.compute_ll()
        return eng.gather_ll()

    def rejection_sample(self, data, prior_samples, n_prior_samples=None,
                         max_posterior_samples=None, n_linear_samples=1, return_logprobs=False,
                         return_all_logprobs=False, n_batches=None, randomize_prior_order=False,
                         in_memory=False):
        """Rejection sampling (thejoker.py:140-257)."""
        if isinstance(prior_samples, (int, np.integer)):
            # thejoker.py:215-218 draws the prior samples first; here they are drawn on
            # the GPUs and never exist on the host
            out = self._rejection_sample_device_prior(data, int(prior_samples),
                                                      max_posterior_samples, n_linear_samples,
                                                      return_logprobs, return_all_logprobs,
                                                      in_memory, n_batches)
            if out is not None:
                return out
            prior_samples = self.prior.sample(size=int(prior_samples),
                                              return_logprobs=return_logprobs, rng=self.rng)
        helper0 = self._make_joker_helper(data)
        cols, ln_prior = self._columns(helper0, prior_samples)
        n_total = len(cols[0])
        if return_logprobs and ln_prior is None:
            raise RuntimeError("return_logprobs=True but ln_prior values not found in prior "
                               "samples: generate them with prior.sample(..., "
                               "return_logprobs=True)")
        sel = None
        if in_memory:
            n_use = n_total  # likelihood_helpers.py:91-127 uses every row it is given
        else:
            if n_prior_samples is None:
                n_use = n_total
            elif n_prior_samples > n_total:
                raise ValueError("Number of prior samples to use is greater than the number of "
                                 f"prior samples passed. n_prior_samples={n_prior_samples} vs. "
                                 f"n_total_samples={n_total}")
            else:
                n_use = int(n_prior_samples)
            if randomize_prior_order:  # multiproc_helpers.py:245-248
                sel = self.rng.choice(n_total, size=n_use, replace=False)
                cols = [c if np.ndim(c) == 0 else c[sel] for c in cols]
            elif n_use < n_total:
                cols = [c if np.ndim(c) == 0 else c[:n_use] for c in cols]
        if max_posterior_samples is None:
            max_posterior_samples = n_use

        eng, helper = self._engine(data, cols, helper0)
        eng.compute_ll()
        good, _ = self._uniform_accept(eng, n_use, max_posterior_samples)
        full_idx = good if sel is None else sel[good]

        samples = self._full_samples(helper, self._rows(cols, good), self.rng, n_linear_samples,
                                     in_memory, n_batches)
        lls = None
        if return_logprobs or return_all_logprobs:
            lls = eng.gather_ll()
        if return_logprobs:
            samples["ln_prior"] = np.repeat(ln_prior[full_idx], n_linear_samples)
            samples["ln_likelihood"] = np.repeat(lls[good], n_linear_samples)
        if return_all_logprobs:
            return samples, lls
        return samples

    def _rejection_sample_device_prior(self, data, n, max_posterior_samples, n_linear_samples,
                                       return_logprobs, return_all_logprobs, in_memory, n_batches):
        got = self._engine_device_prior(data, n)
        if got is None:
            return None
        eng, helper = got
        if max_posterior_samples is None:
            max_posterior_samples = n
        eng.compute_ll()
        good, _ = self._uniform_accept(eng, n, max_posterior_samples)
        # the accepted rows are re-generated from their global indices (any rank can)
        rows = helper.prior_rows(eng.gen, good)
        samples = self._full_samples(helper, rows, self.rng, n_linear_samples, in_memory, n_batches)
        lls = None
        if return_logprobs or return_all_logprobs:
            lls = eng.gather_ll()
        if return_logprobs:
            lp = self.prior.ln_prior_rows(rows, helper.internal_units["s"])
            samples["ln_prior"] = np.repeat(lp, n_linear_samples)
            samples["ln_likelihood"] = np.repeat(lls[good], n_linear_samples)
        if return_all_logprobs:
            return samples, lls
        return samples

    def iterative_rejection_sample(self, data, prior_samples, n_requested_samples,
                                   max_prior_samples=None, n_linear_samples=1,
                                   return_logprobs=False, n_batches=None,
                                   randomize_prior_order=False, init_batch_size=None,
                                   growth_factor=128, in_memory=False):
        """Adaptive rejection sampling (thejoker.py:259-370;
        likelihood_helpers.py:130-229; multiproc_helpers.py:289-427)."""
        helper0 = self._make_joker_helper(data)
        cols, ln_prior = self._columns(helper0, prior_samples)
        n_total = len(cols[0])
        if return_logprobs and ln_prior is None:
            raise RuntimeError("return_logprobs=True but ln_prior values not found in prior "
                               "samples")
        maxiter = 128
        if in_memory:
            safety_factor, n_max = 1, n_total            # likelihood_helpers.py:144-145
        else:
            safety_factor = 4                            # multiproc_helpers.py:333-334
            n_max = n_total if max_prior_samples is None else int(max_prior_samples)
        n_process = growth_factor * n_requested_samples if init_batch_size is None \
            else init_batch_size
        if n_process > n_max:
            raise ValueError("Prior sample library not big enough! For iterative sampling, you "
                             "have to have at least growth_factor * n_requested_samples = "
                             f"{growth_factor * n_requested_samples} samples in the prior samples "
                             f"cache file. You have, or have limited to, {n_max} samples.")
        if not in_memory and randomize_prior_order:      # multiproc_helpers.py:350-354
            all_idx = self.rng.choice(n_total, size=n_max, replace=False)
            cols = [c if np.ndim(c) == 0 else c[all_idx] for c in cols]
        else:
            all_idx = None  # identity (np.arange(n_max) would cost 2 GB at 2^28 rows)
            if n_max < n_total:
                cols = [c if np.ndim(c) == 0 else c[:n_max] for c in cols]

        eng, helper = self._engine(data, cols, helper0, cyclic=True)
        start_idx = 0
        n_accum = 0
        good = np.zeros(0, dtype=np.int64)
        for i in range(maxiter):
            logger.log(1, f"iteration {i}, computing {n_process} likelihoods")
            eng.compute_ll_global(start_idx, start_idx + n_process)
            n_accum = start_idx + n_process
            good, n_good = self._uniform_accept(eng, n_accum, None)
            ll_max = self.last_stats["ll_max"]
            if in_memory and (not np.isfinite(ll_max) or self.last_stats["n_nonfinite"] > 0):
                # likelihood_helpers.py:173-176 *returns* this error object; the test there is
                # np.isfinite over every ll so far, so a -inf (which the max does not show)
                # counts too: the accept kernel counts NaN / +-inf lls (tjb_accept_nonfinite)
                return RuntimeError(f"There are NaN or Inf likelihood values in iteration step {i}!")
            if n_good == 0:
                raise RuntimeError("Failed to find any good samples!")
            logger.log(1, f"{n_good} good samples after rejection sampling")
            if n_good >= n_requested_samples:
                break
            start_idx += n_process
            n_need = n_requested_samples - n_good
            n_process = int(safety_factor * n_need / n_good * n_accum)
            if start_idx + n_process > n_max:
                n_process = n_max - start_idx
            if n_process <= 0:
                break
        else:
            raise RuntimeError("Hit maximum number of iterations!")

        good = good[:n_requested_samples]
        full_idx = good if all_idx is None else all_idx[good]
        samples = self._full_samples(helper, self._rows(cols, good), self.rng, n_linear_samples,
                                     in_memory, n_batches)
        if return_logprobs:
            lls = eng.gather_ll(0, n_accum)
            samples["ln_prior"] = np.repeat(ln_prior[full_idx], n_linear_samples)
            samples["ln_likelihood"] = np.repeat(lls[good], n_linear_samples)
        self.last_stats["n_ll_evaluated"] = n_accum
        return samples
