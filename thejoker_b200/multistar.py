"""Multi-star batched driver: many stars against one shared, device-resident prior
cache (BASELINE.json configs[4]: an APOGEE-like batch of 4096
stars x ~20 epochs x 2^22 shared prior samples with multi-survey v0 offsets).

The reference has no API for this: users loop over stars, and every
``TheJoker.rejection_sample`` call re-creates the helper and re-reads the prior cache
(thejoker/thejoker.py:87-91, 213-257).  Here the prior columns are uploaded once per
GPU, one library handle per GPU is re-pointed at each star (``tjb_update_star``), and
stars are sharded over GPUs / ranks with the reference's ``batch_tasks`` rule -- no
collective is needed (each star is an independent rejection-sampling problem).

Two engines run the star loop.  ``engine="native"`` (default): the loop itself is C++
(``tjb_multistar_rejection``: per-star handle update, ll, accept, row gather and linear
draws on a few host threads with one CUDA stream each, no Python between the kernels);
this process only prepares the stars of the next chunk (``validate_prepare_data``,
``extract_spec``, the child generators) and unpacks the previous chunk's results while the
current chunk runs with the GIL released.  ``engine="python"``: the same steps driven from
Python threads, one library call at a time (needed for ``draw="numpy"`` and for
``max_posterior_samples=None``).  Both give identical results.

RNG: star i draws from its own child generator ``Generator(PCG64(seed_seq.spawn(n)[i]))``
(the reference's own device for per-task streams, multiproc_helpers.py:49-54), consumed
as ``rejection_sample(..., in_memory=True)`` consumes it: ``uniform(size=n_prior)`` then
the linear-parameter draws.  Results therefore do not depend on how stars are sharded.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .data_helpers import validate_prepare_data
from .helper import CJokerHelper, _pcg_struct, extract_spec
from .samples import JokerSamples
from .sharding import shard_ranges

__all__ = ["MultiStarJoker"]


def _pack_for_exchange(ids, packed, stats):
    """This rank's stars as a few flat arrays (cheap to pickle): the packed sample rows of
    all stars back to back, rows per star, t_ref and statistics per star, and the distinct
    unit tables (one, unless the stars' rv units differ)."""
    units_list, unit_of, uidx = [], {}, np.zeros(len(ids), dtype=np.int64)
    for q, i in enumerate(ids):
        units = packed[i][3]
        key = tuple((k, v.dims, v.scale) for k, v in units.items())
        if key not in unit_of:
            unit_of[key] = len(units_list)
            units_list.append(units)
        uidx[q] = unit_of[key]
    raws = [packed[i][0] for i in ids]
    has_ll = bool(ids) and packed[ids[0]][1] is not None
    width = raws[0].shape[1] if raws else 0
    return dict(
        ids=np.asarray(ids, dtype=np.int64), n_rows=np.array([len(r) for r in raws], dtype=np.int64),
        rows=np.concatenate(raws) if raws else np.zeros((0, width)),
        lls=np.concatenate([packed[i][1] for i in ids]) if has_ll else None,
        t_ref=np.array([packed[i][2] for i in ids], dtype=np.float64), uidx=uidx,
        units=units_list,
        stats=np.array([[stats[i]["n_accepted"], stats[i]["n_near_threshold"]] for i in ids],
                       dtype=np.int64).reshape(len(ids), 2),
        ll_max=np.array([stats[i]["ll_max"] for i in ids], dtype=np.float64))


def _unpack_from_exchange(part, poly_trend, n_offsets, results, stats):
    pos = 0
    for q, i in enumerate(part["ids"].tolist()):
        n = int(part["n_rows"][q])
        smp = JokerSamples.unpack(part["rows"][pos:pos + n], part["units"][int(part["uidx"][q])],
                                  t_ref=float(part["t_ref"][q]), poly_trend=poly_trend,
                                  n_offsets=n_offsets)
        if part["lls"] is not None:
            smp["ln_likelihood"] = part["lls"][pos:pos + n]
        pos += n
        results[i] = smp
        stats[i] = dict(n_accepted=int(part["stats"][q, 0]), n_near_threshold=int(part["stats"][q, 1]),
                        ll_max=float(part["ll_max"][q]))


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
    engine : "native" (default; the star loop runs inside the library) or "python"
    """

    # the native engine pre-draws max_posterior_samples x n_linear_samples x L normals per
    # star; beyond this many per star the Python engine (which draws exactly what is
    # needed) is used instead
    _NATIVE_MAX_NORMALS = 1 << 16

    def __init__(self, prior, prior_samples, rng=None, devices=(0,), jitter_mode="apply",
                 group=None, draw="device", streams_per_device=4, engine="native"):
        if engine not in ("native", "python"):
            raise ValueError("engine must be 'native' or 'python'")
        self.engine = engine
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
        self._packed = None  # star -> packed results, kept only for the exchange between ranks

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
                self._dev[d] = dict(cols=[up(c) for c in cols[:4]],
                                    s=None if uniform_s else up(cols[4]), slots=[])
                torch.cuda.synchronize(d)  # the slots' streams read the uploaded columns
        self._helper0 = helper0
        self._factory = first_helper_factory

    def _python_slots(self, d):
        """The Python engine's per-slot helper, ll buffer and stream (created on first use)."""
        import torch

        st = self._dev[d]
        n = len(self._host_cols[0])
        with torch.cuda.device(d):
            while len(st["slots"]) < self.streams_per_device:
                first = d == self.devices[0] and not st["slots"]
                st["slots"].append(dict(
                    helper=self._helper0 if first else self._factory(d),
                    ll=torch.empty(n, dtype=torch.float64, device=f"cuda:{d}"),
                    stream=torch.cuda.Stream(device=d)))
        return st["slots"]

    # -- native engine --------------------------------------------------------------
    def _build_chunk(self, d, star_ids, prepare, seqs, keep, n_per):
        """Host side of one native call: the stars' specs, generator states and
        pre-drawn normals, plus the output arrays."""
        n, L = len(star_ids), self._helper0.n_linear
        n_prior = len(self._host_cols[0])
        specs = (_lib.TjbSpec * n)()
        pcg = (_lib.TjbPcg64 * n)()
        normals = np.empty((n, keep, n_per, L))
        alive, meta = [], []
        for j, i in enumerate(star_ids):
            all_data, ids, trend_M = prepare(i)
            sp = extract_spec(all_data, self.prior, trend_M, self.jitter_mode)
            if sp["n_linear"] != L:
                raise ValueError("all stars must have the same number of linear parameters")
            specs[j] = CJokerHelper._c_spec(sp)
            child = np.random.Generator(np.random.PCG64(seqs[i]))
            pcg[j] = _pcg_struct(child)
            # consumed as rejection_sample(in_memory=True) consumes it: uniform(size=n_prior),
            # then the normals of the accepted rows (a prefix of what is drawn here)
            child.bit_generator.advance(n_prior)
            normals[j] = child.standard_normal((keep, n_per, L))
            alive.append(sp)  # the spec's arrays are read by the native call
            meta.append((sp["internal_units"], all_data.t_ref))
        st = self._dev[d]
        out = dict(idx=np.empty((n, keep), dtype=np.int64), counts=np.zeros((n, 3), dtype=np.int64),
                   llmax=np.empty(n), rows=np.empty((n, keep * n_per, 5 + L)),
                   ll=np.empty((n, keep)))
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        job = _lib.TjbMultiStarJob(
            n_stars=n, specs=specs, pcg=pcg,
            d_P=st["cols"][0].data_ptr(), d_e=st["cols"][1].data_ptr(),
            d_omega=st["cols"][2].data_ptr(), d_M0=st["cols"][3].data_ptr(),
            d_s=st["s"].data_ptr() if st["s"] is not None else None,
            s_const=self._s_const, n_prior=n_prior, max_keep=keep, near_tol=1e-12,
            n_per=n_per, clamp_K=0, n_slots=self.streams_per_device,
            h_normals=vp(normals), h_idx=vp(out["idx"]), h_counts=vp(out["counts"]),
            h_llmax=vp(out["llmax"]), h_rows=vp(out["rows"]), h_ll=vp(out["ll"]))
        return dict(job=job, alive=(alive, specs, pcg, normals), meta=meta, out=out,
                    star_ids=star_ids, device=d)

    def _finish_chunk(self, chunk, n_per, return_logprobs, results, stats):
        out = chunk["out"]
        for j, i in enumerate(chunk["star_ids"]):
            k = int(out["counts"][j, 1])
            units, t_ref = chunk["meta"][j]
            smp = JokerSamples.unpack(out["rows"][j, : k * n_per], units, t_ref=t_ref,
                                      poly_trend=self.prior.poly_trend,
                                      n_offsets=self.prior.n_offsets)
            lls = np.repeat(out["ll"][j, :k], n_per) if return_logprobs else None
            if return_logprobs:
                smp["ln_likelihood"] = lls
            if self._packed is not None:
                self._packed[i] = (out["rows"][j, : k * n_per], lls, t_ref, units)
            results[i] = smp
            stats[i] = dict(n_accepted=int(out["counts"][j, 0]),
                            n_near_threshold=int(out["counts"][j, 2]),
                            ll_max=float(out["llmax"][j]))

    def _run_native(self, prepare, seqs, r_lo, dev_ranges, keep, n_per, return_logprobs,
                    results, stats):
        """Chunks of stars through tjb_multistar_rejection, one call in flight per device;
        this thread prepares the next chunk and unpacks the previous one meanwhile (ctypes
        releases the GIL for the duration of the call)."""
        from concurrent.futures import ThreadPoolExecutor

        lib = _lib.load()

        def call(chunk):
            rc = lib.tjb_multistar_rejection(chunk["device"], ctypes.byref(chunk["job"]))
            _lib.check(rc)  # the error message is per thread: read it here
            return chunk

        # chunk sizes grow so that the GPU starts early and the tail of each call is short
        # against its length
        todo = {}
        for d, (a, b) in zip(self.devices, dev_ranges):
            ids, pos, size = list(range(r_lo + a, r_lo + b)), 0, 4 * self.streams_per_device
            chunks = []
            while pos < len(ids):
                chunks.append(ids[pos:pos + size])
                pos += size
                size = min(4 * size, 512)
            todo[d] = chunks
        with ThreadPoolExecutor(max(1, len(self.devices))) as ex:
            running = {d: None for d in self.devices}
            while any(todo[d] for d in self.devices) or any(running.values()):
                for d in self.devices:
                    nxt = (self._build_chunk(d, todo[d].pop(0), prepare, seqs, keep, n_per)
                           if todo[d] else None)
                    done = running[d].result() if running[d] is not None else None
                    running[d] = ex.submit(call, nxt) if nxt is not None else None
                    if done is not None:
                        self._finish_chunk(done, n_per, return_logprobs, results, stats)

    def rejection_sample(self, stars, max_posterior_samples=256, n_linear_samples=1,
                         return_logprobs=False, gather=True):
        """``stars``: list whose items are what ``TheJoker.rejection_sample`` takes as
        ``data`` (an RVData, or a list / dict of RVData for multi-survey stars).
        Returns a list of JokerSamples (one per star, in order).

        With a process group the stars are sharded over the ranks.  ``gather=True``: the
        ranks exchange their results (as packed arrays) and every rank returns all stars;
        ``gather=False``: no communication, a rank's list holds ``None`` for the stars of
        the other ranks (``last_stats`` likewise)."""
        import time

        import torch

        t_start = time.perf_counter()
        timing = {}
        n_stars = len(stars)
        self._packed = {} if (self.group is not None and gather) else None
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
        n_lin = self._helper0.n_linear
        native = (self.engine == "native" and self.draw == "device" and
                  min(max_keep, n_prior) * int(n_linear_samples) * n_lin <= self._NATIVE_MAX_NORMALS)
        timing["setup_s"] = time.perf_counter() - t_start
        if native:
            t0 = time.perf_counter()
            self._run_native(prepare, seqs, r_lo, dev_ranges, min(max_keep, n_prior),
                             int(n_linear_samples), return_logprobs, results, stats)
            timing["star_loop_s"] = time.perf_counter() - t0
            dev_ranges = [(0, 0)] * len(self.devices)  # nothing left for the Python engine

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
                    if self._packed is not None:
                        self._packed[i] = (raw, lls if return_logprobs else None, all_data.t_ref,
                                           helper.internal_units)
                    results[i] = smp
                    stats[i] = dict(n_accepted=total, n_near_threshold=near,
                                    ll_max=helper.llmax_value(key))

        work = []
        for d, (a, b) in zip(self.devices, dev_ranges):
            if b > a:
                self._python_slots(d)  # created here, not in the slot threads
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
        if self._packed is not None:
            import torch.distributed as dist

            # Packed arrays, not JokerSamples objects: pickling 2048 sample tables costs
            # ~0.25 s and unpickling them ~0.1 s per sender, against ~40 us per star to
            # rebuild a table from its packed rows.
            parts = [None] * world
            t0 = time.perf_counter()
            mine = _pack_for_exchange(list(range(r_lo, r_hi)), self._packed, stats)
            timing["exchange_pack_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            dist.all_gather_object(parts, mine, group=self.group)
            timing["exchange_gather_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            for r, part in enumerate(parts):
                if r != rank:
                    _unpack_from_exchange(part, self.prior.poly_trend, self.prior.n_offsets,
                                          results, stats)
            timing["exchange_unpack_s"] = time.perf_counter() - t0
            self._packed = None
        timing["total_s"] = time.perf_counter() - t_start
        self.last_timing = timing  # where the wall clock of this call went (this rank)
        self.last_stats = [stats.get(i) for i in range(n_stars)]
        return [results.get(i) for i in range(n_stars)]
