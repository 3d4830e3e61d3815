"""Per-GPU shards of the prior cache: the replacement for the reference's
schwimmbad chunk pool (thejoker/multiproc_helpers.py:17-60, utils.py:22-72).

The reference maps contiguous chunks of prior samples over pool workers, each of
which re-reads its rows from a shared HDF5 file and returns a float64[n_i] array that
the master concatenates; max / compare / where then run serially on the master
(multiproc_helpers.py:256-258).  Here a shard is a contiguous index range owned by one
GPU (same split rule, so concatenating shards in order reproduces the global index
order).  Prior columns that live on the host (numpy arrays, memory-mapped cache files)
are streamed through the GPU range by range as they are evaluated -- the counterpart of
the workers' ``read_batch`` slices -- and columns drawn on the device stay resident; in
both cases ll never leaves the device, and the only exchange is the 8-byte max key
(integer MAX all-reduce over NCCL when the shards live in different processes) and the
accepted indices.

Two deployments share this code:
  * one process driving several GPUs (``DeviceEngine(devices=[0, 1, ...])``) -- keys
    are combined on the host;
  * one process per GPU under torchrun (``DeviceEngine(..., group=dist.group.WORLD)``)
    -- every rank holds its own shard; keys are combined with ``dist.all_reduce(MAX)``
    and accepted indices with ``all_gather``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["batch_tasks", "shard_ranges", "DeviceEngine", "merge_accepted", "allreduce_max_key",
           "gather_accepted", "allgather_ragged", "LibComm"]


def batch_tasks(n_tasks, n_batches, arr=None, args=None, start_idx=0):
    """thejoker/utils.py:22-72: equal contiguous split, the first ``n_tasks % n_batches``
    batches get one extra task."""
    args = [] if args is None else list(args)
    tasks = []
    if n_batches > 0 and n_tasks >= n_batches:
        base, rmdr = divmod(n_tasks, n_batches)
        i1 = start_idx
        for i in range(n_batches):
            i2 = i1 + base + (1 if i < rmdr else 0)
            tasks.append([(i1, i2) if arr is None else arr[i1:i2], i1] + args)
            i1 = i2
    else:
        if arr is None:
            tasks.append([(start_idx, n_tasks + start_idx), start_idx] + args)
        else:
            tasks.append([arr[start_idx:n_tasks + start_idx], start_idx] + args)
    return tasks


def bind_to_device_cpus(device):
    """Restrict this process to the host CPUs local to ``device`` (NVML's CPU affinity
    of the GPU, intersected with the CPUs the process may already use), so that the
    page-locked staging buffers of the host-streaming paths are allocated -- first touch
    -- on the GPU's own NUMA node.  For one-process-per-GPU launches (torchrun); call it
    before the first allocation.  Returns the CPU set in effect, or None if nothing was
    changed (no NVML, a single node, or an empty intersection)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        local = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        want = local & allowed
        if not want or want == allowed:
            return None
        os.sched_setaffinity(0, want)
        return sorted(want)
    except Exception:
        return None


def shard_ranges(n, n_shards):
    """[(lo, hi)] per shard; always n_shards entries (empty ranges when n < n_shards)."""
    if n >= n_shards:
        return [t[0] for t in batch_tasks(n, n_shards)]
    return [(min(i, n), min(i + 1, n)) for i in range(n_shards)]


def merge_accepted(per_shard_idx, per_shard_total, max_keep):
    """Concatenate ascending per-shard index lists in shard order and truncate to
    max_keep -- ``good_samples_idx[:max_posterior_samples]`` (likelihood_helpers.py:109)."""
    idx = np.concatenate([np.asarray(a, dtype=np.int64) for a in per_shard_idx]) \
        if per_shard_idx else np.zeros(0, dtype=np.int64)
    total = int(np.sum(per_shard_total))
    if max_keep is not None:
        idx = idx[:max_keep]
    return idx, total


def allreduce_max_key(key, group=None):
    """In-place integer MAX all-reduce of an int64 max-key tensor (the one collective
    of the accept step: multiproc_helpers.py:256-258 does ``lls.max()`` on the master
    after gathering every ll; here only 8 bytes per rank move).  NCCL for CUDA
    tensors; gloo (CPU tests) goes through a host copy when the tensor is on a GPU."""
    import torch.distributed as dist

    if dist.get_backend(group) == "gloo" and key.is_cuda:
        tmp = key.cpu()
        dist.all_reduce(tmp, op=dist.ReduceOp.MAX, group=group)
        key.copy_(tmp)
    else:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
    return key


def gather_accepted(idx, total, near, max_keep, group=None, device=None, nonfinite=None):
    """All ranks get the rank-ordered concatenation of the per-rank ascending index
    lists, truncated to max_keep, plus the global counts.  Two fixed-size tensor
    all-gathers (counts, then indices padded to the longest kept list) -- nothing is
    pickled.  This is the torch.distributed path (gloo in the CPU tests, or several index
    segments per rank); one contiguous shard per rank under NCCL goes through the
    library's own tjb_accept_dist instead.  With ``nonfinite`` (this rank's count of NaN /
    inf lls) the global count is returned as a fourth value."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device("cpu") if dist.get_backend(group) == "gloo" or device is None else device
    idx = np.asarray(idx, dtype=np.int64)
    if max_keep is not None:
        idx = idx[:max_keep]
    mine = torch.tensor([len(idx), int(total), int(near), int(nonfinite or 0)], dtype=torch.int64,
                        device=dev)
    counts = torch.empty(world * 4, dtype=torch.int64, device=dev)  # flat: gloo wants 1-D
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = counts.cpu().numpy().reshape(world, 4)
    m = int(counts[:, 0].max())
    parts = []
    if m > 0:
        send = torch.zeros(m, dtype=torch.int64, device=dev)
        send[: len(idx)] = torch.from_numpy(idx).to(dev)
        recv = torch.empty(world * m, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        recv = recv.cpu().numpy().reshape(world, m)
        parts = [recv[r, : counts[r, 0]] for r in range(world)]
    idx, total = merge_accepted(parts, counts[:, 1], max_keep)
    if nonfinite is not None:
        return idx, total, int(counts[:, 2].sum()), int(counts[:, 3].sum())
    return idx, total, int(counts[:, 2].sum())


class LibComm:
    """The library's own NCCL communicator over the ranks of a torch.distributed group
    (tjb_comm_create).  torch.distributed is used once, to hand rank 0's NCCL unique id to
    the other ranks; the collectives of the accept step then run inside
    libthejoker_b200.so (tjb_accept_dist).  One communicator per (group, device), cached."""

    _cache = {}

    def __init__(self, group, device):
        import ctypes

        import torch
        import torch.distributed as dist

        from . import _lib

        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = torch.zeros(_lib.TJB_COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * _lib.TJB_COMM_ID_BYTES)()
            _lib.check(lib.tjb_comm_unique_id(buf))
            ident = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        src = dist.get_global_rank(group, 0) if group is not None else 0
        if dist.get_backend(group) == "nccl":
            ident = ident.to(f"cuda:{device}")
            dist.broadcast(ident, src=src, group=group)
            ident = ident.cpu()
        else:
            dist.broadcast(ident, src=src, group=group)
        raw = bytes(ident.numpy().tobytes())
        h = ctypes.c_void_p()
        _lib.check(lib.tjb_comm_create(raw, world, rank, int(device), ctypes.byref(h)))
        self.handle, self.rank, self.world, self.device, self._lib = h, rank, world, int(device), lib

    @classmethod
    def get(cls, group, device):
        key = (id(group), int(device))
        if key not in cls._cache:
            cls._cache[key] = cls(group, device)
        return cls._cache[key]

    def close(self):
        if self.handle:
            self._lib.tjb_comm_destroy(self.handle)
            self.handle = None


def allgather_ragged(arr, group, device=None):
    """Rank-ordered concatenation of per-rank float64 arrays of different lengths: a
    length all-gather, then one padded tensor all-gather (no pickling)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device("cpu") if dist.get_backend(group) == "gloo" or device is None else device
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    lens = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens, torch.tensor([len(arr)], dtype=torch.int64, device=dev),
                                group=group)
    lens = lens.cpu().numpy()
    m = int(lens.max())
    if m == 0:
        return np.zeros(0)
    send = torch.zeros(m, dtype=torch.float64, device=dev)
    send[: len(arr)] = torch.from_numpy(arr).to(dev)
    recv = torch.empty(world * m, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, m)
    return np.concatenate([recv[r, : lens[r]] for r in range(world)])


class _Shard:
    __slots__ = ("device", "helper", "lo", "hi", "cols", "s", "ll", "key")


class DeviceEngine:
    """The prior cache, sharded over GPUs, with the hot-path operations on it.

    Parameters
    ----------
    make_helper : callable(device) -> CJokerHelper
    columns : [P, e, omega, M0, s] host float64 arrays in internal units (s may be
        None or a scalar for a constant jitter) -- the *local* part of the cache when
        ``group`` is given, the whole cache otherwise.
    devices : list of CUDA device indices driven by this process.
    group : torch.distributed process group, or None.  With a group, this process's
        columns are the shard of rank ``dist.get_rank(group)`` and ``global_offset``
        is its first global index.
    """

    def __init__(self, make_helper, columns, devices=(0,), group=None, global_offset=0,
                 global_size=None, resident=False, cyclic=False):
        import torch

        self.torch = torch
        self.group = group
        P, e, om, M0, s = columns
        n_local = len(P)
        # cyclic (SPMD ranks, host-resident prior that every rank holds in full): a range
        # [glo, ghi) handed to compute_ll_global is split evenly over ALL ranks, range by
        # range -- the block-cyclic layout the iterative sampler needs, whose early rounds
        # cover a small prefix of the cache that a contiguous layout would leave on rank 0
        # alone (multiproc_helpers.py:355-410 maps every round over the whole pool too).
        # A rank then owns a list of segments (glo, ghi, ll) instead of one contiguous shard.
        self.cyclic = bool(cyclic) and group is not None
        self.segments, self.rounds = [], []
        self.n_local = n_local
        self.global_offset = int(global_offset)
        self.n_global = int(n_local if global_size is None else global_size)
        s_is_scalar = s is None or np.ndim(s) == 0
        self.s_const = 0.0 if s is None else (float(s) if s_is_scalar else 0.0)
        if not s_is_scalar and n_local > 0 and np.all(s == s[0]):
            s_is_scalar, self.s_const = True, float(s[0])
        self.shards = []
        # resident=False (default): the host columns stay where they are (numpy arrays or
        # memory-mapped cache files) and every compute_ll() streams just the requested
        # range through the GPU (tjb_marginal_ll_host_soa_resident) -- nothing of the cache
        # is uploaded that the sampler does not evaluate (the iterative sampler usually
        # stops after a small prefix), and H2D overlaps the kernel.
        # resident=True: upload each shard once, for caches that are evaluated repeatedly.
        self.host_cols = None if resident else [P, e, om, M0, None if s_is_scalar else s]
        if self.cyclic:
            if resident or len(devices) != 1:
                raise ValueError("cyclic sharding streams a host-resident prior, one GPU per rank")
            self._add_shard(make_helper, devices[0], 0, 0, None, None)
            self.peer_max = False
            return
        for d, (lo, hi) in zip(devices, shard_ranges(n_local, len(devices))):
            cols, s_dev = None, None
            if resident:
                with torch.cuda.device(d):
                    torch.cuda.current_stream().synchronize()

                    def up(a):
                        a = np.ascontiguousarray(a[lo:hi], dtype=np.float64)
                        if not a.flags.writeable:  # memory-mapped cache columns are read-only
                            a = a.copy()
                        return torch.from_numpy(a).to(f"cuda:{d}", non_blocking=True)

                    cols = [up(P), up(e), up(om), up(M0)]
                    s_dev = None if s_is_scalar else up(s)
            self._add_shard(make_helper, d, lo, hi, cols, s_dev)
        self._link_peers()

    def _link_peers(self):
        """One process, several GPUs: let every shard's likelihood kernel publish its
        max into all the other shards' keys over NVLink (fused max exchange).  Falls back
        to combining the keys on the host if the GPUs cannot map each other."""
        self.peer_max = False
        if self.group is not None or len(self.shards) < 2:
            return
        try:
            for sh in self.shards:
                sh.helper.set_peer_keys([o.key for o in self.shards if o is not sh])
            self.peer_max = True
        except Exception:
            for sh in self.shards:
                sh.helper.set_peer_keys([])

    def _ctx(self, d):
        """CUDA device context of a shard (a no-op for the CPU stand-in helpers of the
        gloo tests, whose device is the string "cpu")."""
        import contextlib

        return contextlib.nullcontext() if d == "cpu" else self.torch.cuda.device(d)

    @staticmethod
    def _devstr(d):
        return "cpu" if d == "cpu" else f"cuda:{d}"

    def _add_shard(self, make_helper, d, lo, hi, cols, s_dev):
        torch = self.torch
        sh = _Shard()
        sh.device, sh.lo, sh.hi = d, lo, hi
        # a device listed twice gets one handle (stream, max key, peer list) per shard: two
        # shards on one GPU exercise the whole sharded path on a single-GPU box
        dup = any(o.device == d for o in self.shards)
        sh.helper = make_helper(d, True) if dup else make_helper(d)
        sh.cols, sh.s = cols, s_dev
        with self._ctx(d):
            sh.ll = torch.full((hi - lo,), float("nan"), dtype=torch.float64, device=self._devstr(d))
            sh.key = sh.helper.new_llmax_key()
        self.shards.append(sh)

    @classmethod
    def from_device_columns(cls, make_helper, shards, s_const=0.0, group=None, global_offset=0,
                            global_size=None):
        """Engine over prior columns that already live on the GPUs.  ``shards`` is a list
        of ``(device, [P, e, omega, M0] float64 CUDA tensors, s tensor or None)`` in
        index order; nothing is copied."""
        import torch

        self = cls.__new__(cls)
        self.torch, self.group = torch, group
        self.s_const = float(s_const)
        self.shards = []
        lo = 0
        for d, cols, s_dev in shards:
            hi = lo + cols[0].numel()
            self._add_shard(make_helper, d, lo, hi, [c.contiguous() for c in cols], s_dev)
            lo = hi
        self.n_local = lo
        self.global_offset = int(global_offset)
        self.n_global = int(lo if global_size is None else global_size)
        self._link_peers()
        return self

    @classmethod
    def from_generator(cls, make_helper, gen, n_local, devices=(0,), group=None, global_offset=0,
                       global_size=None):
        """Engine over a prior that is *drawn*, not stored: sample g (global index) is a
        pure function of (gen.seed, g) (csrc/prior_gen.cuh), generated in registers by the
        likelihood kernel and re-generated for the accepted indices.  No prior column
        exists on any device; a shard only owns its ll array."""
        import torch

        self = cls.__new__(cls)
        self.torch, self.group, self.gen = torch, group, gen
        self.s_const = 0.0
        self.shards = []
        for d, (lo, hi) in zip(devices, shard_ranges(int(n_local), len(devices))):
            self._add_shard(make_helper, d, lo, hi, None, None)
        self.n_local = int(n_local)
        self.global_offset = int(global_offset)
        self.n_global = int(n_local if global_size is None else global_size)
        self._link_peers()
        return self

    def rows(self, local_idx):
        """Packed (k, 5) host rows [P, e, omega, M0, s] for local sample indices."""
        torch = self.torch
        local_idx = np.asarray(local_idx, dtype=np.int64)
        if getattr(self, "gen", None) is not None:
            return self.shards[0].helper.prior_rows(self.gen, local_idx + self.global_offset)
        out = np.empty((len(local_idx), 5))
        if getattr(self, "host_cols", None) is not None:
            for j, c in enumerate(self.host_cols[:4]):
                out[:, j] = np.asarray(c)[local_idx]
            sc = self.host_cols[4]
            out[:, 4] = self.s_const if sc is None else np.asarray(sc)[local_idx]
            return out
        for sh in self.shards:
            m = (local_idx >= sh.lo) & (local_idx < sh.hi)
            if not m.any():
                continue
            ii = torch.from_numpy(local_idx[m] - sh.lo).to(sh.cols[0].device)
            for j, c in enumerate(sh.cols):
                out[m, j] = c.index_select(0, ii).cpu().numpy()
            out[m, 4] = self.s_const if sh.s is None else sh.s.index_select(0, ii).cpu().numpy()
        return out

    # -- likelihood -----------------------------------------------------------
    def compute_ll(self, lo=0, hi=None):
        """ll[lo:hi) (local indices) on every shard that intersects the range; the
        shard's running-max key is updated.  Asynchronous."""
        hi = self.n_local if hi is None else hi
        torch = self.torch
        if getattr(self, "host_cols", None) is not None:
            return self._compute_ll_streamed(lo, hi)
        for sh in self.shards:
            a, b = max(lo, sh.lo), min(hi, sh.hi)
            if a >= b:
                continue
            with self._ctx(sh.device):
                sl = slice(a - sh.lo, b - sh.lo)
                if getattr(self, "gen", None) is not None:
                    sh.helper.marginal_ll_generated(self.gen, self.global_offset + a, b - a,
                                                    out=sh.ll[sl], llmax_key=sh.key)
                    continue
                sh.helper.marginal_ll_soa(*[c[sl] for c in sh.cols],
                                          s=None if sh.s is None else sh.s[sl],
                                          s_const=self.s_const, out=sh.ll[sl], llmax_key=sh.key)

    def _compute_ll_streamed(self, lo, hi):
        """Host-resident cache: each GPU streams its part of [lo, hi) from host memory;
        one host thread per GPU (the library call blocks, ctypes releases the GIL)."""
        torch = self.torch
        work = []
        for sh in self.shards:
            a, b = max(lo, sh.lo), min(hi, sh.hi)
            if a < b:
                work.append((sh, a, b))

        def run(item):
            sh, a, b = item
            P, e, om, M0, s = self.host_cols
            with self._ctx(sh.device):
                sh.helper.marginal_ll_host_columns(
                    P[a:b], e[a:b], om[a:b], M0[a:b], s=None if s is None else s[a:b],
                    s_const=self.s_const, out=sh.ll[a - sh.lo:b - sh.lo], llmax_key=sh.key)

        if len(work) == 1:
            run(work[0])
        elif work:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(len(work)) as ex:
                list(ex.map(run, work))

    def compute_ll_global(self, glo, ghi):
        """ll for the part of the GLOBAL index range [glo, ghi) this process owns."""
        if getattr(self, "cyclic", False):
            return self._compute_ll_cyclic(int(glo), int(ghi))
        lo = max(glo, self.global_offset) - self.global_offset
        hi = min(ghi, self.global_offset + self.n_local) - self.global_offset
        if lo < hi:
            self.compute_ll(lo, hi)

    def _compute_ll_cyclic(self, glo, ghi):
        """This rank's 1/world slice of the global range, streamed from the host columns
        into a new ll segment."""
        import torch.distributed as dist

        torch = self.torch
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.rounds.append((glo, ghi))
        a, b = shard_ranges(ghi - glo, world)[rank]
        a, b = a + glo, b + glo
        if a >= b:
            return
        sh = self.shards[0]
        P, e, om, M0, s = self.host_cols
        with self._ctx(sh.device):
            ll = torch.empty(b - a, dtype=torch.float64, device=self._devstr(sh.device))
            sh.helper.marginal_ll_host_columns(P[a:b], e[a:b], om[a:b], M0[a:b],
                                               s=None if s is None else s[a:b],
                                               s_const=self.s_const, out=ll, llmax_key=sh.key)
        self.segments.append((a, b, ll))

    def _accept_cyclic(self, rng, n_accum, max_keep, uniforms, near_tol):
        """Accept over this rank's segments (each with its own PCG offset = its first
        global index), then the union over the ranks in ascending global order."""
        torch = self.torch
        sh = self.shards[0]
        self.global_max_key()
        per_idx, total, near, nonfin = [], 0, 0, 0
        with self._ctx(sh.device):
            for a, b, ll in self.segments:
                b = min(b, n_accum)
                if a >= b:
                    continue
                if uniforms is not None:
                    u_dev = torch.from_numpy(np.ascontiguousarray(uniforms[a:b])).to(ll.device)
                    idx, tot, nn = sh.helper.accept(ll[: b - a], sh.key, uniforms=u_dev, index_base=a,
                                                    max_keep=max_keep, near_tol=near_tol)
                else:
                    idx, tot, nn = sh.helper.accept(ll[: b - a], sh.key, rng=rng, rng_offset=a,
                                                    index_base=a, max_keep=max_keep,
                                                    near_tol=near_tol)
                per_idx.append(idx.cpu().numpy())
                total += tot
                near += nn
                nonfin += getattr(sh.helper, "last_nonfinite", 0)
        mine = np.concatenate(per_idx) if per_idx else np.zeros(0, dtype=np.int64)
        # the global first max_keep are among every rank's own first max_keep
        idx, total, near, nonfin = gather_accepted(
            mine, total, near, None if max_keep is None else max_keep, self.group,
            device=sh.key.device, nonfinite=nonfin)
        self.last_nonfinite = nonfin
        idx = np.sort(idx)
        return (idx if max_keep is None else idx[:max_keep]), total, near

    def _gather_ll_cyclic(self, glo, ghi):
        import torch.distributed as dist

        world = dist.get_world_size(self.group)
        mine = [ll.cpu().numpy() for _, _, ll in self.segments]
        flat = allgather_ragged(np.concatenate(mine) if mine else np.zeros(0), self.group,
                                self.shards[0].key.device)
        # every rank knows every rank's segment layout: rank-major, rounds in order
        n_hi = max((b for _, b in self.rounds), default=0)
        out = np.full(n_hi, np.nan)
        pos = 0
        for r in range(world):
            for (lo, hi) in self.rounds:
                a, b = shard_ranges(hi - lo, world)[r]
                out[lo + a:lo + b] = flat[pos:pos + (b - a)]
                pos += b - a
        return out[glo:ghi]

    def reset_max(self):
        self.synchronize()
        for sh in self.shards:
            with self._ctx(sh.device):
                sh.key.copy_(sh.helper.new_llmax_key())

    def synchronize(self):
        for sh in self.shards:
            if sh.device != "cpu":
                self.torch.cuda.synchronize(sh.device)

    def global_max_key(self):
        """Combine the shard keys: host max within the process, then an integer MAX
        all-reduce across processes (NCCL).  Every shard's key tensor is overwritten
        with the global key so the accept kernels read it from device memory."""
        torch = self.torch
        keys = [sh.key for sh in self.shards]
        if len(keys) > 1 and getattr(self, "peer_max", False):
            # every kernel already wrote its max into every GPU's key; what remains is
            # ordering: each GPU's stream waits for the other GPUs' likelihood kernels
            events = []
            for sh in self.shards:
                with self._ctx(sh.device):
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(sh.device))
                    events.append(ev)
            for sh in self.shards:
                st = torch.cuda.current_stream(sh.device)
                for ev in events:
                    st.wait_event(ev)
        elif len(keys) > 1:
            m = max(int(k.item()) for k in keys)
            for k in keys:
                k.fill_(m)
        if self.group is not None:
            if self._lib_collectives():
                sh = self.shards[0]
                with self._ctx(sh.device):
                    sh.helper.allreduce_max_key(LibComm.get(self.group, sh.device), keys[0])
            else:
                allreduce_max_key(keys[0], self.group)
            for k in keys[1:]:
                k.copy_(keys[0].to(k.device))
        return keys[0]

    def max_value(self):
        sh = self.shards[0]
        return sh.helper.llmax_value(sh.key)

    # -- accept ---------------------------------------------------------------
    def accept(self, rng, hi=None, max_keep=None, uniforms=None, near_tol=1e-12):
        """Global ``where(exp(ll - max) > u)[0][:max_keep]`` over local [0, hi).

        ``rng``: numpy Generator.  On PCG64 the uniforms are generated on the device
        from the generator's state (sample with global index g uses the g-th double)
        and the caller advances the generator by the global count afterwards; on any
        other bit generator pass host ``uniforms`` for the local range instead.
        Returns (global indices int64 ascending, n_accepted_total, n_near).
        """
        torch = self.torch
        # hi is a GLOBAL count of accumulated samples; this process owns part of it
        n_accum = self.n_global if hi is None else int(hi)
        if getattr(self, "cyclic", False):
            return self._accept_cyclic(rng, n_accum, max_keep, uniforms, near_tol)
        hi = self.n_local if hi is None else max(0, min(hi - self.global_offset, self.n_local))
        if self._lib_collectives():
            # one rank per GPU over NCCL: max all-reduce, accept and the gathers of counts
            # and indices all run inside the library (tjb_accept_dist)
            sh = self.shards[0]
            with self._ctx(sh.device):
                comm = LibComm.get(self.group, sh.device)
                u_dev = None
                if uniforms is not None:
                    g0 = self.global_offset
                    u_dev = torch.from_numpy(np.ascontiguousarray(uniforms[g0:g0 + hi])).to(sh.ll.device)
                idx, tot, nn = sh.helper.accept_dist(
                    comm, sh.ll[:hi], sh.key, self.global_offset, uniforms=u_dev,
                    rng=None if uniforms is not None else rng, max_keep=max_keep,
                    n_global=n_accum, near_tol=near_tol)
                self.last_nonfinite = sh.helper.last_nonfinite  # summed over the ranks
                return idx.cpu().numpy(), tot, nn
        self.global_max_key()
        per_idx, per_tot, near, nonfin = [], [], 0, 0
        for sh in self.shards:
            a, b = sh.lo, min(hi, sh.hi)
            if a >= b:
                continue
            with self._ctx(sh.device):
                ll = sh.ll[: b - a]
                if uniforms is not None:
                    g0 = self.global_offset
                    u_dev = torch.from_numpy(np.ascontiguousarray(uniforms[g0 + a:g0 + b])).to(ll.device)
                    idx, tot, nn = sh.helper.accept(ll, sh.key, uniforms=u_dev,
                                                    index_base=self.global_offset + a,
                                                    max_keep=max_keep, near_tol=near_tol)
                else:
                    idx, tot, nn = sh.helper.accept(ll, sh.key, rng=rng,
                                                    rng_offset=self.global_offset + a,
                                                    index_base=self.global_offset + a,
                                                    max_keep=max_keep, near_tol=near_tol)
                per_idx.append(idx.cpu().numpy())
                per_tot.append(tot)
                near += nn
                nonfin += getattr(sh.helper, "last_nonfinite", 0)
        idx, total = merge_accepted(per_idx, per_tot, max_keep)
        if self.group is not None:
            idx, total, near, nonfin = gather_accepted(idx, total, near, max_keep, self.group,
                                                       device=self.shards[0].ll.device,
                                                       nonfinite=nonfin)
        self.last_nonfinite = nonfin
        return idx, total, near

    def _lib_collectives(self):
        """True when the cross-rank accept can run inside the library: a torch group on the
        NCCL backend and one shard (GPU) per rank."""
        if self.group is None or len(self.shards) != 1:
            return False
        import torch.distributed as dist

        return dist.get_backend(self.group) == "nccl"

    def gather_ll(self, glo=0, ghi=None):
        """ll over the GLOBAL range [glo, ghi) on every rank (rank-ordered concatenation,
        multiproc_helpers.py:120 ``np.concatenate(results)``)."""
        ghi = self.n_global if ghi is None else ghi
        if getattr(self, "cyclic", False):
            return self._gather_ll_cyclic(glo, ghi)
        lo = max(glo, self.global_offset) - self.global_offset
        hi = min(ghi, self.global_offset + self.n_local) - self.global_offset
        mine = self.download_ll(lo, hi) if lo < hi else np.zeros(0)
        if self.group is None:
            return mine
        import torch.distributed as dist

        return allgather_ragged(mine, self.group, self.shards[0].ll.device)

    # -- host access ------------------------------------------------------------
    def download_ll(self, lo=0, hi=None):
        hi = self.n_local if hi is None else hi
        out = np.empty(hi - lo)
        for sh in self.shards:
            a, b = max(lo, sh.lo), min(hi, sh.hi)
            if a < b:
                out[a - lo:b - lo] = sh.ll[a - sh.lo:b - sh.lo].cpu().numpy()
        return out
