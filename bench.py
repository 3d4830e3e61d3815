#!/usr/bin/env python
"""bench.py -- prior samples / second through the marginal log-likelihood.

Metric (BASELINE.json): prior samples/sec through marginal ll (N=64 epochs, 2^28-sample
default prior, L=2, s=0) at 1/2/4/8 B200.  A "step" is one pass of the hot kernel over
the whole prior cache; under torchrun the 2^28 samples are sharded over the ranks with
the reference's batch_tasks rule (strong scaling, no data-path collective: the shards
are independent; the 8-byte max-key all-reduce of the accept step runs once after the
timed region as a cross-rank check).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU
                                                             # implementation, host cores

One JSON line on stdout (rank 0).  `value`: device-resident SoA prior, CUDA-event time,
max over ranks.  `e2e`: the reference-facing data path on HOST buffers, host->device and
device->host copies inside the timed region -- page-locked prior columns in / ll out
(`e2e.value`), the packed (n, 5) chunk of CJokerHelper.batch_marginal_ln_likelihood
(`e2e.packed_chunk`), and ordinary pageable numpy columns (`e2e.pageable_columns`).
`roofline`: algorithmic FP64 work / kernel time against the FP64
FMA-chain peak measured in this run.  `cpu_baseline` and `--impl reference`: the
reference's own Cython operator compiled from its sources (oracle/_ref, kind "reference";
one process per core like its pool.map) when that build is present, else the C
restatement (oracle/joker_oracle.c, kind "port"), on this box's cores, on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_EPOCHS = 64
LOG2_PRIOR = 28
PRIOR_SEED = 123
METRIC = "prior samples/sec through marginal ll (N=64, 2^28 prior)"
UNIT = "samples/s"

# BASELINE.md section 3 / SURVEY.md section 8(d): reference-algorithm work model
W_EPOCH_L2 = 292.0   # flop per (sample, epoch): Newton k=3 Kepler solve + RV column + Gram sums
W_TAIL = 300.0       # flop per sample: Lambda_K, L x L factorisation, log det, combine
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2


def w_sample(n_epochs):
    return n_epochs * W_EPOCH_L2 + W_TAIL


def make_star():
    """the reference's test fixture with noise, default prior."""
    import thejoker_b200 as tj
    from thejoker_b200 import units as u
    from thejoker_b200.data_helpers import validate_prepare_data
    from thejoker_b200.synthetic import make_noisy_data

    data, _ = make_noisy_data(n_times=N_EPOCHS, seed=42)
    prior = tj.JokerPrior.default(P_min=2 * u.day, P_max=1024 * u.day, sigma_K0=30 * u.km / u.s,
                                  sigma_v=100 * u.km / u.s)
    all_data, ids, trend_M = validate_prepare_data(data, prior.poly_trend, prior.n_offsets)
    return all_data, prior, trend_M


def shard_of(n_total, rank, world):
    from thejoker_b200.sharding import shard_ranges

    return shard_ranges(n_total, world)[rank]


# ---------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            load = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(smax)),
                       reasons=sorted(reasons), samples=len(sm), power_w_max=float(max(power)))
        return out


# ---------------------------------------------------------------------------------
def accept_check(args, helper, prior, gen, ll, key, lo, n, n_total, rank, world, local):
    """Untimed correctness leg: the full accept over all ranks vs rank 0 alone on a window."""
    import torch
    import torch.distributed as dist

    from thejoker_b200 import units as u
    from thejoker_b200.helper import prior_sample_device

    rng = np.random.default_rng(2024)  # PCG64: u of global sample g = its g-th double
    max_keep = 1 << 20
    out = {}
    if world > 1:
        from thejoker_b200.sharding import LibComm

        comm = LibComm.get(dist.group.WORLD, local)
        key_d = key.clone()
        helper.accept_dist(comm, ll, key_d, lo, rng=rng, max_keep=max_keep, n_global=n_total)
        torch.cuda.synchronize()
        dist.barrier()
        key_d = key.clone()
        t0 = time.perf_counter()
        idx, tot, near = helper.accept_dist(comm, ll, key_d, lo, rng=rng, max_keep=max_keep,
                                            n_global=n_total)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["path"] = "tjb_accept_dist (NCCL inside the library)"
    else:
        key_d = key.clone()
        helper.accept(ll, key_d, rng=rng, max_keep=max_keep)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx, tot, near = helper.accept(ll, key_d, rng=rng, max_keep=max_keep)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        out["path"] = "tjb_accept"
    ll_max = helper.llmax_value(key_d)
    assert np.isfinite(ll_max) and ll_max >= ll.max().item()
    out.update(ms=float(dt.item()) * 1e3, n_accepted=int(tot), n_near=int(near), ll_max=ll_max,
               parity=None)
    if rank == 0:
        # rank 0 alone: regenerate a window across the first shard boundary, ll, accept with
        # the global max and the same uniforms; the distributed index set must agree there
        w = min(1 << 22, n_total)
        from thejoker_b200.sharding import shard_ranges

        edge = shard_ranges(n_total, world)[0][1] if world > 1 else n_total // 2
        w_lo = max(0, min(edge - w // 2, n_total - w))
        cols = prior_sample_device(gen, w_lo, w, local, with_s=False)
        ll_w = helper.marginal_ll_soa(*cols, s=None, s_const=0.0)
        idx_w, _, near_w = helper.accept(ll_w, key_d, rng=rng, rng_offset=w_lo, index_base=w_lo,
                                         max_keep=w)
        idx_w = idx_w.cpu().numpy()
        idx_all = idx.cpu().numpy()
        in_w = idx_all[(idx_all >= w_lo) & (idx_all < w_lo + w)]
        truncated = tot > len(idx_all)
        # samples within 1e-12 of the threshold may legitimately differ: counted, excluded
        diff = np.setxor1d(in_w, idx_w)
        if truncated:
            diff = diff[diff <= idx_all[-1]]
        out.update(parity=bool(len(diff) <= near_w), window=[int(w_lo), int(w_lo + w)],
                   n_accepted_in_window=int(len(idx_w)), n_differ=int(len(diff)),
                   n_near_in_window=int(near_w))
    return out


# ---------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import thejoker_b200 as tj

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    numa_cpus = None
    if world > 1 and os.environ.get("TJB_BENCH_NUMA_BIND", "1") == "1":
        from thejoker_b200.sharding import bind_to_device_cpus

        numa_cpus = bind_to_device_cpus(local)  # e2e staging buffers on the GPU's node
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    n_total = 1 << args.log2_prior
    lo, hi = shard_of(n_total, rank, world)
    n = hi - lo
    all_data, prior, trend_M = make_star()
    helper = tj.CJokerHelper(all_data, prior, trend_M, device=local)

    # synthetic default prior: the library's own counter-based sampler (tjb_prior_sample),
    # sample g of the 2^28 a function of (seed, g) -- the shards of any world size are
    # pieces of the same prior, which the accept-parity check below relies on
    from thejoker_b200 import units as u
    from thejoker_b200.helper import prior_sample_device

    gen = prior.device_generator(PRIOR_SEED, u.km / u.s)
    P, e, om, M0 = prior_sample_device(gen, lo, n, local, with_s=False)
    ll = torch.empty(n, dtype=torch.float64, device="cuda")
    key = helper.new_llmax_key()

    def step():
        helper.marginal_ll_soa(P, e, om, M0, s=None, s_const=0.0, out=ll, llmax_key=key)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = ev[0].elapsed_time(ev[-1])
    per_launch = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    ms_per_step = ms_total_max / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- multi-rank accept (untimed): the whole accept step across the ranks -- NCCL MAX
    # all-reduce of the key, numpy-identical PCG64 uniforms at global offsets, compaction,
    # all-gather of counts and indices, all inside the library (tjb_accept_dist) -- checked
    # against a single-GPU recompute by rank 0 of a 2^22-sample window that straddles the
    # boundary between the first two shards (the generator makes any window reproducible)
    accept = accept_check(args, helper, prior, gen, ll, key, lo, n, n_total, rank, world, local)
    ll_max = accept["ll_max"]

    # ---- e2e: host buffers in, host ll out, copies inside the timed region ------------
    # (1) prior samples as separate pinned host columns -- what a JokerSamples holds and
    #     what TheJoker.marginal_ln_likelihood(data, prior_samples) sends
    #     (CJokerHelper.marginal_ln_likelihood_columns -> tjb_marginal_ll_host_soa);
    # (2) the literal reference-facing call CJokerHelper.batch_marginal_ln_likelihood on a
    #     packed (n, 5) pinned chunk (tjb_marginal_ll_host).
    import ctypes

    from thejoker_b200 import _lib

    n_e2e = min(n, 1 << args.log2_e2e)
    e2e_steps = max(1, min(args.steps, 3))
    host_ll = torch.empty(n_e2e, dtype=torch.float64).pin_memory()
    ll_np = host_ll.numpy()

    def time_e2e(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()  # synchronous: returns when ll is back in host memory
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert np.array_equal(ll_np[:1024], ll[:1024].cpu().numpy())
        return n_e2e * world * e2e_steps / float(dt.item())

    cols_host = [t[:n_e2e].cpu().pin_memory() for t in (P, e, om, M0)]
    cols_np = [t.numpy() for t in cols_host]
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)

    def e2e_columns():
        _lib.check(helper._lib.tjb_marginal_ll_host_soa(helper._h, *[vp(c) for c in cols_np], None,
                                                        0.0, n_e2e, vp(ll_np)))

    e2e_value = time_e2e(e2e_columns)

    # (1a) what the platform allows for exactly these bytes: the same four pinned columns
    #      copied to the device and ll copied back with bare cudaMemcpyAsync on two streams,
    #      all ranks at once, no kernel -- the ceiling `e2e.value` is to be read against
    #      (8 ranks share the host's memory and PCIe fabric: tools/h2d_ceiling.py)
    def copy_ceiling():
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dev_cols = [torch.empty(n_e2e, dtype=torch.float64, device="cuda") for _ in range(4)]
        dev_ll = ll[:n_e2e]
        scratch_ll = torch.empty(n_e2e, dtype=torch.float64).pin_memory()

        def once():
            with torch.cuda.stream(s_in):
                for d, h in zip(dev_cols, cols_host):
                    d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s_out):
                scratch_ll.copy_(dev_ll, non_blocking=True)
            s_in.synchronize(); s_out.synchronize()

        once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            once()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return n_e2e * world * e2e_steps / float(dt.item())

    e2e_ceiling = copy_ceiling()

    # (1b) the same call on ordinary (pageable) numpy columns, as a user's JokerSamples
    #      holds them: staged through the library's page-locked ring by host threads
    e2e_pageable = e2e_public = None
    if world == 1:
        n_pg = min(n_e2e, 1 << 26)
        cols_pg = [np.array(c[:n_pg]) for c in cols_np]
        ll_pg = np.empty(n_pg)
        run_pg = lambda: helper.marginal_ln_likelihood_columns(*cols_pg, out=ll_pg)
        run_pg()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            run_pg()
        e2e_pageable = {"value": n_pg * e2e_steps / (time.perf_counter() - t0), "n": int(n_pg),
                        "h2d_bytes_per_step": int(n_pg * 32), "d2h_bytes_per_step": int(n_pg * 8),
                        "call": "CJokerHelper.marginal_ln_likelihood_columns on pageable numpy "
                                "columns, pageable ll out"}
        assert np.array_equal(ll_pg[:1024], ll[:1024].cpu().numpy())
        del ll_pg
        # (1c) the user's call itself: TheJoker.marginal_ln_likelihood(data, JokerSamples),
        #      ordinary numpy columns with units in, a new numpy ll array out
        from thejoker_b200.synthetic import make_noisy_data

        data_pub, _ = make_noisy_data(n_times=N_EPOCHS, seed=42)
        smp = tj.JokerSamples()
        for name, c, unit in zip(("P", "e", "omega", "M0"), cols_pg, (u.day, u.one, u.rad, u.rad)):
            smp[name] = u.Quantity(c, unit)
        smp["s"] = u.Quantity(np.zeros(n_pg), u.km / u.s)
        smp._uniform_s = True
        joker = tj.TheJoker(prior, devices=[local])
        ll_pub = joker.marginal_ln_likelihood(data_pub, smp)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ll_pub = joker.marginal_ln_likelihood(data_pub, smp)
        e2e_public = {"value": n_pg * e2e_steps / (time.perf_counter() - t0), "n": int(n_pg),
                      "h2d_bytes_per_step": int(n_pg * 32), "d2h_bytes_per_step": int(n_pg * 8),
                      "call": "TheJoker.marginal_ln_likelihood(RVData, JokerSamples): pageable numpy "
                              "columns in, new numpy array out (helper construction included)"}
        assert np.array_equal(ll_pub[:1024], ll[:1024].cpu().numpy())
        del cols_pg, smp, ll_pub
    del cols_host, cols_np

    host = torch.empty((n_e2e, 5), dtype=torch.float64).pin_memory()
    host[:, 0].copy_(P[:n_e2e]); host[:, 1].copy_(e[:n_e2e]); host[:, 2].copy_(om[:n_e2e])
    host[:, 3].copy_(M0[:n_e2e]); host[:, 4].zero_()
    chunk_np = host.numpy()

    def e2e_chunk():
        _lib.check(helper._lib.tjb_marginal_ll_host(helper._h, vp(chunk_np), n_e2e, vp(ll_np)))

    e2e_chunk_value = time_e2e(e2e_chunk)
    del host, chunk_np

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline -----------------------------------------------------------------------
    peak_tf, _ = helper.fp64_peak(40000)
    kernel_ms = float(np.mean(per_launch))
    achieved_tf = (n * w_sample(N_EPOCHS)) / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_bytes = n * 40.0  # 4 prior columns read + ll written
    hbm_gbs = hbm_bytes / (kernel_ms * 1e-3) / 1e9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "kernel_counts.json")))
    except Exception:
        pass
    from thejoker_b200 import _lib as tjlib

    src_hash = tjlib.source_hash()
    counts_current = bool(prof) and prof.get("source_sha256") == src_hash
    exec_flop = prof.get("executed_fp64_flop_per_sample")
    executed_tf = None if not exec_flop else n * exec_flop / (kernel_ms * 1e-3) / 1e12
    roofline = {
        # the hardware fraction: FP64 flops the kernel EXECUTES per sample (ncu: 2 DFMA + DMUL
        # + DADD, profiles/kernel_counts.json) x samples / this run's kernel time, over the
        # FP64 FMA-chain peak measured in this run
        "bound": "fp64", "achieved": executed_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": None if executed_tf is None else executed_tf / peak_tf,
        "peak_source": "FP64 FMA-chain microbenchmark (tjb_fp64_peak) measured in this run; "
                       f"nominal {FP64_NOMINAL_TFLOPS:.1f} (MEASURED_PEAKS.json has no FP64 entry)",
        "frac_of_nominal": None if executed_tf is None else executed_tf / FP64_NOMINAL_TFLOPS,
        "executed_fp64_flop_per_sample": exec_flop,
        "fp64_pipe_pct_ncu": prof.get("fp64_pipe_pct"),
        "counts_from": {"file": "profiles/kernel_counts.json", "tag": prof.get("tag"),
                        "source_sha256": prof.get("source_sha256"),
                        "registers_per_thread": prof.get("registers_per_thread"),
                        "matches_this_source_tree": counts_current},
        "source_sha256": src_hash,
        # the reference-algorithm work model (18 988 flop/sample with a
        # 3-iteration Newton solve); the kernel executes ~4x fewer flops, so this is a
        # speed-up-over-the-reference-algorithm figure, not a hardware fraction
        "work_model_flop_per_sample": w_sample(N_EPOCHS),
        "achieved_work_model": achieved_tf,
        "frac_work_model": achieved_tf / peak_tf,
        "kernel": "marginal_ll_kernel<2,false,PriorView,EpochRowsParam>", "kernel_ms": kernel_ms,
        "traffic": prof.get("dram_bytes_per_launch_at_2p28") if n == (1 << 28) else None,
        "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                "algorithmic_bytes_per_sample": 40,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650"},
    }

    cpu = cpu_baseline(args.cpu_seconds) if world == 1 and not args.no_cpu_baseline else None

    rec = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"N={N_EPOCHS} epochs, L=2 (K, v0), s=0, 2^{args.log2_prior} default-prior "
                               "samples sharded over the ranks (configs[1]/metric config)",
                   "l2": "inputs (32 B/sample x 2^28/ranks) are larger than L2",
                   "n_prior": n_total, "n_epochs": N_EPOCHS, "sharding": f"contiguous x{world}"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_e2e * 32),
                "d2h_bytes_per_step": int(n_e2e * 8), "n_per_rank": int(n_e2e), "steps": e2e_steps,
                "rank0_cpu_binding": numa_cpus,
                "copy_ceiling": {"value": e2e_ceiling, "unit": UNIT,
                                 "frac": e2e_value / e2e_ceiling,
                                 "what": "the same pinned columns in and ll out with bare "
                                         "cudaMemcpyAsync on two streams, all ranks at once, no "
                                         "kernel: what the host's memory and PCIe fabric allow "
                                         "for these bytes"},
                "pageable_columns": e2e_pageable,
                "public_api": e2e_public,
                "call": "TheJoker.marginal_ln_likelihood data path: pinned host columns P, e, "
                        "omega, M0 (s constant) in, host ll out (CJokerHelper."
                        "marginal_ln_likelihood_columns -> tjb_marginal_ll_host_soa)",
                "packed_chunk": {"value": e2e_chunk_value, "h2d_bytes_per_step": int(n_e2e * 40),
                                 "d2h_bytes_per_step": int(n_e2e * 8),
                                 "call": "CJokerHelper.batch_marginal_ln_likelihood on a pinned "
                                         "(n,5) chunk -> tjb_marginal_ll_host"}},
        "gpu_launches": args.steps,
        "clocks": clocks,
        "roofline": roofline,
        "ll_max": ll_max,
        "accept_parity": accept["parity"], "n_near_threshold": accept["n_near"],
        "accept_ms": accept["ms"], "accept": accept,
    }
    if cpu is not None:
        rec["cpu_baseline"] = cpu
    print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------
def _oracle_and_chunk(n):
    from oracle.oracle import OracleHelper
    from thejoker_b200.helper import extract_spec
    from thejoker_b200.synthetic import default_prior_columns

    all_data, prior, trend_M = make_star()
    spec = extract_spec(all_data, prior, trend_M)
    chunk = np.ascontiguousarray(np.stack(default_prior_columns(n, seed=123), axis=1))
    return OracleHelper.from_spec(spec), chunk


def host_cores():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1,
    which must not cap the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# -- the reference's own compiled Cython (oracle/_ref), one worker process per core ------
_REF_WORKER = None


def _ref_worker_init(spec, poly_trend, n_offsets):
    global _REF_WORKER
    from oracle.oracle import _limit_blas_threads
    from oracle.ref_cython import RefCythonHelper

    _limit_blas_threads()  # one LAPACK thread per worker process: no oversubscription

    _REF_WORKER = RefCythonHelper(spec, poly_trend, n_offsets)


def _ref_worker_ll(chunk):
    return _REF_WORKER.batch_marginal_ln_likelihood(chunk)


class ReferencePool:
    """thejoker/multiproc_helpers.py:39-58, 96 restated for the timing arm: the prior
    chunk is cut into one contiguous task per worker and every worker calls the
    reference's CJokerHelper.batch_marginal_ln_likelihood -- the real thing, compiled by
    oracle/ref_build/build_ref.py (twobody's Kepler function supplied by the oracle)."""

    kind = "reference"
    what = ("fast_likelihood.pyx compiled unmodified from the reference (oracle/_ref; "
            "twobody's c_rv_from_elements restated), one process per core like its pool.map")

    def __init__(self, cores):
        import multiprocessing as mp

        from thejoker_b200.helper import extract_spec

        from oracle import ref_cython

        all_data, prior, trend_M = make_star()
        spec = extract_spec(all_data, prior, trend_M)
        self.cores = cores
        # the compiled reference operator is loaded in this process too (the forked workers
        # inherit the mapping), so that what ran is visible from the parent's loaded objects
        ref_cython.load()
        self.native_so = ref_cython.ext_path()
        self.pool = mp.get_context("fork").Pool(
            cores, initializer=_ref_worker_init,
            initargs=(spec, int(spec["n_poly"]), int(spec["n_offsets"])))

    def ll(self, chunk):
        parts = [np.ascontiguousarray(c) for c in np.array_split(chunk, self.cores) if len(c)]
        return np.concatenate(self.pool.map(_ref_worker_ll, parts, chunksize=1))

    def close(self):
        self.pool.close()
        self.pool.join()


class PortPool:
    """The C restatement (oracle/joker_oracle.c), OpenMP over samples."""

    kind = "port"
    what = ("oracle/joker_oracle.c (restatement of fast_likelihood.pyx + scipy LAPACK, "
            "OpenMP over samples; oracle/_ref is not built on this box)")

    def __init__(self, cores):
        self.cores = cores
        self.orc, _ = _oracle_and_chunk(1)

    def ll(self, chunk):
        return self.orc.batch_marginal_ln_likelihood(chunk, n_threads=self.cores)

    def close(self):
        pass


def cpu_arm(cores):
    """oracle/_ref when the reference operator was compiled (it travels with the repo),
    else the oracle port."""
    from oracle import ref_cython

    want = os.environ.get("TJB_BENCH_CPU_ARM", "auto")  # "port" forces the restatement
    use_ref = want == "reference" or (want == "auto" and ref_cython.available())
    return ReferencePool(cores) if use_ref else PortPool(cores)


def _prior_chunk(n):
    from thejoker_b200.synthetic import default_prior_columns

    return np.ascontiguousarray(np.stack(default_prior_columns(n, seed=123), axis=1))


def cpu_baseline(seconds=12.0):
    """The reference's CPU implementation of the path on this box's cores, on a bounded
    sample of the same workload."""
    cores = host_cores()
    arm = cpu_arm(cores)
    chunk = _prior_chunk(1 << 12)
    arm.ll(chunk[:256])
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        arm.ll(chunk)
        done += len(chunk)
    dt = time.perf_counter() - t0
    arm.close()
    return {"value": done / dt, "unit": UNIT, "cores": int(cores), "kind": arm.kind,
            "sample": f"{done} default-prior samples at N={N_EPOCHS} in {dt:.1f} s; {arm.what}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = host_cores()
    n_step = 1 << args.log2_ref_step
    arm = cpu_arm(cores)
    chunk = _prior_chunk(n_step)
    for _ in range(max(1, min(args.warmup, 2))):
        arm.ll(chunk[: n_step // 4])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.ll(chunk)
    dt = time.perf_counter() - t0
    arm.close()
    value = n_step * args.steps / dt
    sample = (f"each step = {n_step} samples of the same workload (2^{LOG2_PRIOR} would take "
              f"~{(1 << LOG2_PRIOR) / value / 3600:.1f} h)")
    rec = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"N={N_EPOCHS} epochs, L=2 (K, v0), s=0, default prior; {sample}",
                   "n_prior": 1 << LOG2_PRIOR, "n_epochs": N_EPOCHS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": arm.kind,
                         "sample": sample + "; " + arm.what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "native_so": getattr(arm, "native_so", None),  # the compiled reference operator, loaded here
    }
    print(json.dumps(rec))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-prior", dest="log2_prior", type=int, default=LOG2_PRIOR)
    ap.add_argument("--log2-e2e", dest="log2_e2e", type=int, default=28,
                    help="per-rank samples of the pinned host chunk for the e2e measurement")
    ap.add_argument("--log2-ref-step", dest="log2_ref_step", type=int, default=15)
    ap.add_argument("--cpu-seconds", dest="cpu_seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
