"""Secondary metric: wall time of TheJoker.rejection_sample end to end --
prior on the host (JokerSamples, SoA upload) and prior drawn on the device -- plus the
accept step alone at 2^28.  Run on the GPU box; writes gpurun_out/e2e_rejection.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import thejoker_b200 as tj  # noqa: E402
from helpers import default_prior  # noqa: E402
from thejoker_b200.synthetic import make_noisy_data  # noqa: E402

out = {}
prior = default_prior(1, sigma_K0=30.0, P_min=2.0, P_max=1024.0)
data, _ = make_noisy_data(64, seed=42)
flat, _ = make_noisy_data(64, seed=42, K=1e-4)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


# (a) host prior, 2^24 samples (configs[1])
n = 1 << 24
ps = prior.sample(size=n, rng=np.random.default_rng(1))
for name, d in (("data", data), ("flat", flat)):
    joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
    t, s = timed(lambda: joker.rejection_sample(d, ps, max_posterior_samples=256, in_memory=True))
    out[f"rejection_2p24_host_prior_{name}"] = dict(seconds=t, n_samples=len(s), samples_per_s=n / t,
                                                  stats=dict(joker.last_stats))
# (a2) host prior, 2^26 samples (2.1 GB of pageable numpy columns): plain + iterative
n = 1 << 26
ps = prior.sample(size=n, rng=np.random.default_rng(2))
joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
t, s = timed(lambda: joker.rejection_sample(flat, ps, max_posterior_samples=256, in_memory=True))
out["rejection_2p26_host_prior_flat"] = dict(seconds=t, n_samples=len(s), samples_per_s=n / t,
                                             stats=dict(joker.last_stats))
t, s = timed(lambda: joker.rejection_sample(data, ps, max_posterior_samples=256, in_memory=True))
out["rejection_2p26_host_prior_data"] = dict(seconds=t, n_samples=len(s), samples_per_s=n / t,
                                             stats=dict(joker.last_stats))
t, s = timed(lambda: joker.iterative_rejection_sample(flat, ps, n_requested_samples=256))
out["iterative_2p26_host_prior_flat_256"] = dict(seconds=t, n_samples=len(s),
                                                 stats=dict(joker.last_stats))
t, s = timed(lambda: joker.iterative_rejection_sample(data, ps, n_requested_samples=8))
out["iterative_2p26_host_prior_data_8"] = dict(seconds=t, n_samples=len(s),
                                               stats=dict(joker.last_stats))
del ps
# (b) device prior, 2^28 samples
n = 1 << 28
joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
t, s = timed(lambda: joker.rejection_sample(flat, n, max_posterior_samples=256), reps=2)
out["rejection_2p28_device_prior_flat"] = dict(seconds=t, n_samples=len(s), samples_per_s=n / t,
                                               stats=dict(joker.last_stats))
# (c) accept alone on 2^28 resident lls, numpy-identical uniforms generated on the device
helper = joker._make_joker_helper(flat)
ll = torch.randn(n, dtype=torch.float64, device="cuda") * 3 - 50
key = helper.new_llmax_key()
helper.llmax_update(ll, key)
rng = np.random.default_rng(7)
t, r = timed(lambda: helper.accept(ll, key, rng=rng, max_keep=256))
out["accept_2p28_pcg64"] = dict(seconds=t, n_accepted=r[1], samples_per_s=n / t)
uu = torch.rand(n, dtype=torch.float64, device="cuda")
t, r = timed(lambda: helper.accept(ll, key, uniforms=uu, max_keep=256))
out["accept_2p28_device_uniform_array"] = dict(seconds=t, n_accepted=r[1], samples_per_s=n / t)
t0 = time.perf_counter()
np.random.default_rng(7).uniform(size=1 << 26)
out["numpy_uniform_2p26_host_seconds"] = time.perf_counter() - t0
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "e2e_rejection.json"), "w"), indent=1)
