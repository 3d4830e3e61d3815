"""Multi-star batched driver: many stars against one shared, device-resident prior
cache (BASELINE.json configs[4]: an APOGEE-like batch of 4096
stars x ~20 epochs x 2^22 shared prior samples with multi-survey v0 offsets).

The reference has no API for this: users loop over stars, and every
``TheJoker.rejection_sample`` call re-creates the helper and re-reads the prior cache
(thejoker/thejoker.py:87-91, 213-257).  Here the prior columns are uploaded once per
GPU, one library handle per GPU is re-pointed at each star (``tjb_update_star``), and
stars are sharded over GPUs / ranks with the reference's ``batch_tasks`` rule -- no
collective is needed (each star is an independent rejection-sampling problem).

RNG: star i draws from its own child generator ``Generator(PCG64(seed_seq.spawn(n)[i]))``
(the reference's own device for per-task streams, multiproc_helpers.py:49-54), consumed
as ``rejection_sample(..., in_memory=True)`` consumes it: ``uniform(size=n_prior)`` then
the linear-parameter draws.  Results therefore do not depend on how stars are sharded.
"""
from __future__ import annotations

import numpy as np

from .data_helpers import validate_prepare_data
from .helper import CJokerHelper
from .samples import JokerSamples
from .sharding import shard_ranges

__all__ = ["MultiStarJoker"]


class MultiStarJoker:
    """Rejection-sample many stars against one shared prior cache.

    Parameters
    ----------
    prior : JokerPrior (shared by all stars)
    prior_samples : JokerSamples, or a float64 (n, 5) packed array in internal units
        (the velocity unit of the packed ``s`` column must match the stars' rv unit)
    rng : numpy Generator (PCG64)
    devices : CUDA devices driven by this process
    group : torch.distributed group; stars are sharded over ranks, results gathered
    draw : how the linear parameters of accepted samples are drawn, see
        CJokerHelper.batch_get_posterior_samples ("device" by default: with hundreds of
        accepted samples per star a Python call per row would dominate the star's time)
    streams_per_device : stars in flight per GPU (default 4); results do not depend on it
    """

    def __init__(self, prior, prior_samples, rng=None, devices=(0,), jitter_mode="apply",
                 group=None, draw="device", streams_per_device=4):
        self.prior = prior
        self.rng = np.random.default_rng() if rng is None else rng
        self.devices = list(devices)
        self.jitter_mode = jitter_mode
        self.group = group
        self.draw = draw
        # stars in flight per GPU: each slot has its own library handle, ll buffer, CUDA
        # stream and host thread, so one star's host work (index read-back, row gather,
        # unpacking) overlaps the next star's likelihood kernel
        self.streams_per_device = max(1, int(streams_per_device))
        self._samples = prior_samples
        self._dev = {}      # device -> dict(cols, s, ll, helper)
        self._host_cols = None

    # -- prior residency ---------------------------------------------------------
    def _prepare(self, first_helper_factory):
        import torch

        helper0 = first_helper_factory(self.devices[0])
        ps = self._samples
        if isinstance(ps, JokerSamples):
            cols = ps.columns(units=helper0.internal_units, names=helper0.packed_order)
            uniform_s = ps._uniform_s
        else:
            arr = np.asarray(ps, dtype=np.float64)
            cols = [np.ascontiguousarray(arr[:, i]) for i in range(5)]
            uniform_s = False
        if not uniform_s and len(cols[4]) and np.all(cols[4] == cols[4][0]):
            uniform_s = True
        self._host_cols = cols
        self._s_const = float(cols[4][0]) if uniform_s and len(cols[4]) else 0.0
        n = len(cols[0])
        for d in self.devices:
            with torch.cuda.device(d):
                up = lambda a: torch.from_numpy(a).to(f"cuda:{d}")
                slots = []
                for k in range(self.streams_per_device):
                    first = d == self.devices[0] and k == 0
                    slots.append(dict(
                        helper=helper0 if first else first_helper_factory(d),
                        ll=torch.empty(n, dtype=torch.float64, device=f"cuda:{d}"),
                        stream=torch.cuda.Stream(device=d)))
                self._dev[d] = dict(cols=[up(c) for c in cols[:4]],
                                    s=None if uniform_s else up(cols[4]), slots=slots)
                torch.cuda.synchronize(d)  # the slots' streams read the uploaded columns

    def rejection_sample(self, stars, max_posterior_samples=256, n_linear_samples=1,
                         return_logprobs=False):
        """``stars``: list whose items are what ``TheJoker.rejection_sample`` takes as
        ``data`` (an RVData, or a list / dict of RVData for multi-survey stars).
        Returns a list of JokerSamples (one per star, in order)."""
        import torch

        n_stars = len(stars)
        seqs = self.rng.bit_generator._seed_seq.spawn(n_stars)
        # per-star data preparation happens in the slot threads, overlapped with GPU work
        prepare = lambda i: validate_prepare_data(stars[i], self.prior.poly_trend,
                                                  self.prior.n_offsets)

        def factory(dev):
            all_data, ids, trend_M = prepare(0)
            return CJokerHelper(all_data, self.prior, trend_M, device=dev,
                                jitter_mode=self.jitter_mode)

        if not self._dev:
            self._prepare(factory)
        n_prior = len(self._host_cols[0])
        max_keep = n_prior if max_posterior_samples is None else int(max_posterior_samples)

        # stars -> ranks -> devices, contiguous (batch_tasks rule)
        rank, world = 0, 1
        if self.group is not None:
            import torch.distributed as dist

            rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        r_lo, r_hi = shard_ranges(n_stars, world)[rank]
        dev_ranges = shard_ranges(r_hi - r_lo, len(self.devices))

        results = {}
        stats = {}

        def run_slot(d, slot, star_indices):
            """All the stars of one slot, in order, on the slot's own stream."""
            st = self._dev[d]
            sl = st["slots"][slot]
            helper, ll = sl["helper"], sl["ll"]
            with torch.cuda.device(d), torch.cuda.stream(sl["stream"]):
                for i in star_indices:
                    all_data, ids, trend_M = prepare(i)
                    helper.update_star(all_data, self.prior, trend_M)
                    key = helper.new_llmax_key()
                    helper.marginal_ll_soa(*st["cols"], s=st["s"], s_const=self._s_const, out=ll,
                                           llmax_key=key)
                    child = np.random.Generator(np.random.PCG64(seqs[i]))
                    idx, total, near = helper.accept(ll, key, rng=child, max_keep=max_keep)
                    child.bit_generator.advance(n_prior)
                    good = idx.cpu().numpy()
                    rows = np.empty((len(good), 5))
                    for j, c in enumerate(self._host_cols):
                        rows[:, j] = c[good]
                    raw, lls = helper.batch_get_posterior_samples(rows, n_linear_samples, child,
                                                                  draw=self.draw)
                    smp = JokerSamples.unpack(raw, helper.internal_units, t_ref=all_data.t_ref,
                                              poly_trend=self.prior.poly_trend,
                                              n_offsets=self.prior.n_offsets)
                    if return_logprobs:
                        smp["ln_likelihood"] = lls
                    results[i] = smp
                    stats[i] = dict(n_accepted=total, n_near_threshold=near,
                                    ll_max=helper.llmax_value(key))

        work = []
        for d, (a, b) in zip(self.devices, dev_ranges):
            k = self.streams_per_device
            for slot in range(k):
                mine = list(range(r_lo + a + slot, r_lo + b, k))
                if mine:
                    work.append((d, slot, mine))
        if len(work) == 1:
            run_slot(*work[0])
        elif work:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(len(work)) as ex:
                for f in [ex.submit(run_slot, *w) for w in work]:
                    f.result()
        if self.group is not None:
            import torch.distributed as dist

            parts = [None] * world
            dist.all_gather_object(parts, (results, stats), group=self.group)
            results, stats = {}, {}
            for r, s_ in parts:
                results.update(r)
                stats.update(s_)
        self.last_stats = [stats[i] for i in range(n_stars)]
        return [results[i] for i in range(n_stars)]
