"""BASELINE.json configs[0]: 16 epochs, default prior, 2^18 prior samples -- the reference's
own CPU-runnable case.  Times the CPU arms (the compiled reference operator when
oracle/_ref is present, and the C restatement) through a full in-memory rejection pass
(ll + accept + posterior draws) on all host cores, and -- when a GPU is visible -- the same
through TheJoker.rejection_sample.  Writes gpurun_out/cfg1.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from helpers import default_prior  # noqa: E402
from oracle import ref_cython  # noqa: E402
from oracle.oracle import OracleHelper, rejection_accept  # noqa: E402
from thejoker_b200.data_helpers import validate_prepare_data  # noqa: E402
from thejoker_b200.helper import extract_spec  # noqa: E402
from thejoker_b200.synthetic import default_prior_columns, make_noisy_data  # noqa: E402

n = 1 << 18
prior = default_prior(1, sigma_K0=30.0, P_min=2.0, P_max=1024.0)
data, _ = make_noisy_data(16, seed=42)
all_data, ids, trend_M = validate_prepare_data(data, prior.poly_trend, prior.n_offsets)
spec = extract_spec(all_data, prior, trend_M)
chunk = np.ascontiguousarray(np.stack(default_prior_columns(n, seed=123), axis=1))
cores = bench.host_cores()
out = {"n_prior": n, "n_epochs": 16, "cores": cores}

# the reference's operator, one process per core (multiproc_helpers.py:39-58, 96)
bench.make_star = lambda: (all_data, prior, trend_M)
for name, arm in (("reference", bench.ReferencePool(cores) if ref_cython.available() else None),
                  ("port", bench.PortPool(cores))):
    if arm is None:
        continue
    arm.ll(chunk[:4096])
    t0 = time.perf_counter()
    ll = arm.ll(chunk)
    rng = np.random.default_rng(42)
    good = rejection_accept(ll, rng.uniform(size=n))
    OracleHelper.from_spec(spec).batch_get_posterior_samples(chunk[good], 1, rng)
    dt = time.perf_counter() - t0
    arm.close()
    out[f"cpu_{name}_seconds"] = dt
    out[f"cpu_{name}_samples_per_s"] = n / dt
    out["n_accepted"] = int(len(good))

try:
    import torch

    if torch.cuda.is_available():
        import thejoker_b200 as tj

        joker = tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[0])
        joker.rejection_sample(data, chunk, in_memory=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        smp = joker.rejection_sample(data, chunk, in_memory=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.update(gpu_seconds=dt, gpu_samples_per_s=n / dt, gpu_n_accepted=len(smp))
except ImportError:
    pass
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cfg1.json"), "w"), indent=1)
