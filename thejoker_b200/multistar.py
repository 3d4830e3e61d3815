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
    """

    def __init__(self, prior, prior_samples, rng=None, devices=(0,), jitter_mode="apply",
                 group=None, draw="device"):
        self.prior = prior
        self.rng = np.random.default_rng() if rng is None else rng
        self.devices = list(devices)
        self.jitter_mode = jitter_mode
        self.group = group
        self.draw = draw
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
        for d in self.devices:
            with torch.cuda.device(d):
                up = lambda a: torch.from_numpy(a).to(f"cuda:{d}")
                self._dev[d] = dict(
                    cols=[up(c) for c in cols[:4]], s=None if uniform_s else up(cols[4]),
                    ll=torch.empty(len(cols[0]), dtype=torch.float64, device=f"cuda:{d}"),
                    helper=helper0 if d == self.devices[0] else first_helper_factory(d))

    def rejection_sample(self, stars, max_posterior_samples=256, n_linear_samples=1,
                         return_logprobs=False):
        """``stars``: list whose items are what ``TheJoker.rejection_sample`` takes as
        ``data`` (an RVData, or a list / dict of RVData for multi-survey stars).
        Returns a list of JokerSamples (one per star, in order)."""
        import torch

        n_stars = len(stars)
        seqs = self.rng.bit_generator._seed_seq.spawn(n_stars)
        prepared = [validate_prepare_data(d, self.prior.poly_trend, self.prior.n_offsets)
                    for d in stars]

        def factory(dev):
            all_data, ids, trend_M = prepared[0]
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
        # one star in flight per device: launch on every device, then collect
        cursors = [r_lo + a for a, b in dev_ranges]
        ends = [r_lo + b for a, b in dev_ranges]
        while any(c < e for c, e in zip(cursors, ends)):
            active = []
            for di, d in enumerate(self.devices):
                if cursors[di] >= ends[di]:
                    continue
                i = cursors[di]
                cursors[di] += 1
                st = self._dev[d]
                all_data, ids, trend_M = prepared[i]
                with torch.cuda.device(d):
                    st["helper"].update_star(all_data, self.prior, trend_M)
                    key = st["helper"].new_llmax_key()
                    st["helper"].marginal_ll_soa(*st["cols"], s=st["s"], s_const=self._s_const,
                                                 out=st["ll"], llmax_key=key)
                active.append((i, d, key))
            for i, d, key in active:
                st = self._dev[d]
                child = np.random.Generator(np.random.PCG64(seqs[i]))
                with torch.cuda.device(d):
                    idx, total, near = st["helper"].accept(st["ll"], key, rng=child,
                                                           max_keep=max_keep)
                    child.bit_generator.advance(n_prior)
                    good = idx.cpu().numpy()
                    rows = np.empty((len(good), 5))
                    for j, c in enumerate(self._host_cols):
                        rows[:, j] = c[good]
                    raw, lls = st["helper"].batch_get_posterior_samples(rows, n_linear_samples, child,
                                                                        draw=self.draw)
                all_data = prepared[i][0]
                s = JokerSamples.unpack(raw, st["helper"].internal_units, t_ref=all_data.t_ref,
                                        poly_trend=self.prior.poly_trend,
                                        n_offsets=self.prior.n_offsets)
                if return_logprobs:
                    s["ln_likelihood"] = lls
                results[i] = s
                stats[i] = dict(n_accepted=total, n_near_threshold=near,
                                ll_max=st["helper"].llmax_value(key))
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
