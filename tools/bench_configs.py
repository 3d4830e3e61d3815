"""BASELINE.json configs[2] and configs[3] at N GPUs (torchrun, one rank per GPU):

  cfg3: N=64 epochs, L=3 (v0, v1), per-sample jitter s ~ LogNormal(-2, 1), 2^28 prior samples
        sharded over the ranks: likelihood throughput (CUDA events, max over ranks) and the
        full multi-rank accept (NCCL max-key all-reduce + device PCG64 uniforms + gather).
  cfg4: N=256 epochs, L=2, flat data (K = 1e-4): likelihood throughput over 2^28 sharded
        samples, and the iterative sampler (n_requested=256, growth_factor=128) plus a
        full rejection_sample over a 2^26-row host prior through TheJoker(group=WORLD).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port P tools/bench_configs.py
Rank 0 prints one JSON record and writes gpurun_out/configs_g<N>.json.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import thejoker_b200 as tj  # noqa: E402
from helpers import default_prior  # noqa: E402
from thejoker_b200.data_helpers import validate_prepare_data  # noqa: E402
from thejoker_b200.sharding import allreduce_max_key, shard_ranges  # noqa: E402
from thejoker_b200.synthetic import make_noisy_data  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 28
lo, hi = shard_ranges(1 << LOG2, world)[rank]
n = hi - lo
out = {"n_gpus": world, "n_prior": 1 << LOG2}


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def device_prior(n, seed, jitter):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda: torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    P = torch.exp(r() * (np.log(1024.0) - np.log(2.0)) + np.log(2.0))
    torch.manual_seed(seed)
    ga = torch._standard_gamma(torch.full((n,), 0.867, dtype=torch.float64, device="cuda"))
    gb = torch._standard_gamma(torch.full((n,), 3.03, dtype=torch.float64, device="cuda"))
    e = (ga / (ga + gb)).clamp_(0.0, 1.0 - 1e-12)
    del ga, gb
    om, M0 = (r() * 2 - 1) * np.pi, (r() * 2 - 1) * np.pi
    s = torch.exp(torch.randn(n, dtype=torch.float64, device="cuda", generator=g) - 2.0) \
        if jitter else None
    return P, e, om, M0, s


def time_ll(helper, cols, s, reps=5):
    ll = torch.empty(cols[0].numel(), dtype=torch.float64, device="cuda")
    key = helper.new_llmax_key()
    for _ in range(2):
        helper.marginal_ll_soa(*cols, s=s, out=ll, llmax_key=key)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        helper.marginal_ll_soa(*cols, s=s, out=ll, llmax_key=key)
    e1.record()
    barrier()
    return maxr(e0.elapsed_time(e1) / reps) * 1e-3, ll, key


# ---- cfg3 ---------------------------------------------------------------------------------
prior3 = default_prior(2, sigma_K0=30.0, P_min=2.0, P_max=1024.0)
data3, _ = make_noisy_data(64, seed=42)
all_data, ids, trend_M = validate_prepare_data(data3, prior3.poly_trend, prior3.n_offsets)
h3 = tj.CJokerHelper(all_data, prior3, trend_M, device=local)
P, e, om, M0, s = device_prior(n, 123 + rank, jitter=True)
sec, ll, key = time_ll(h3, [P, e, om, M0], s)
out["cfg3_ll_samples_per_s"] = (1 << LOG2) / sec
out["cfg3_ll_seconds"] = sec
# the multi-rank accept: one integer all-reduce, then per-rank flag / scan / scatter
rng = np.random.default_rng(7)
barrier()
t0 = time.perf_counter()
if world > 1:
    allreduce_max_key(key, dist.group.WORLD)
idx, tot, near = h3.accept(ll, key, rng=rng, rng_offset=lo, index_base=lo, max_keep=256)
counts = torch.tensor([tot, near], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(counts)
torch.cuda.synchronize()
out["cfg3_accept_seconds"] = maxr(time.perf_counter() - t0)
out["cfg3_n_accepted"], out["cfg3_n_near_threshold"] = [int(v) for v in counts.tolist()]
out["cfg3_ll_max"] = h3.llmax_value(key)
del P, e, om, M0, s, ll, h3

# ---- cfg4 ---------------------------------------------------------------------------------
prior4 = default_prior(1, sigma_K0=30.0, P_min=2.0, P_max=1024.0)
flat, _ = make_noisy_data(256, seed=42, K=1e-4)
all_data, ids, trend_M = validate_prepare_data(flat, prior4.poly_trend, prior4.n_offsets)
h4 = tj.CJokerHelper(all_data, prior4, trend_M, device=local)
P, e, om, M0, _ = device_prior(n, 123 + rank, jitter=False)
sec, ll, key = time_ll(h4, [P, e, om, M0], None, reps=3)
out["cfg4_ll_samples_per_s"] = (1 << LOG2) / sec
out["cfg4_ll_seconds"] = sec
del P, e, om, M0, ll, h4
torch.cuda.empty_cache()

n_host = 1 << min(LOG2, 26)
ps = prior4.sample(size=n_host, rng=np.random.default_rng(1))  # same rows on every rank
group = dist.group.WORLD if world > 1 else None


def timed(fn, reps=2):
    fn()
    best = 1e9
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, maxr(time.perf_counter() - t0))
    return best, r


mk = lambda: tj.TheJoker(prior4, rng=np.random.default_rng(42), devices=[local], group=group)
joker = mk()
sec, smp = timed(lambda: joker.iterative_rejection_sample(flat, ps, n_requested_samples=256,
                                                          growth_factor=128))
out["cfg4_iterative_seconds"] = sec
out["cfg4_iterative_n_samples"] = len(smp)
out["cfg4_iterative_n_ll_evaluated"] = int(joker.last_stats["n_ll_evaluated"])
joker = mk()
sec, smp = timed(lambda: joker.rejection_sample(flat, ps, max_posterior_samples=256))
out["cfg4_rejection_host_prior_rows"] = n_host
out["cfg4_rejection_host_prior_seconds"] = sec
out["cfg4_rejection_host_prior_samples_per_s"] = n_host / sec

if rank == 0:
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"configs_g{world}.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
