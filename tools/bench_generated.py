"""Kernel time of the likelihood over a device-drawn prior (tjb_marginal_ll_generated: the
prior is generated in registers inside the kernel) next to the same launch over resident
SoA columns.  usage (GPU box): python tools/bench_generated.py [log2_n]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import thejoker_b200 as tj  # noqa: E402
from helpers import star_spec  # noqa: E402
from thejoker_b200.data_helpers import validate_prepare_data  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
out = {"n": n}
for N, pt in ((64, 1), (20, 2)):
    spec, data, prior = star_spec(N, pt)
    all_data, ids, trend_M = validate_prepare_data(data, prior.poly_trend, prior.n_offsets)
    h = tj.CJokerHelper(all_data, prior, trend_M, device=0)
    gen = prior.device_generator(123, h.internal_units["s"])
    ll = torch.empty(n, dtype=torch.float64, device="cuda")

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5

    ms_gen = timed(lambda: h.marginal_ll_generated(gen, 0, n, out=ll))
    from thejoker_b200.helper import prior_sample_device
    ms_cols = timed(lambda: prior_sample_device(gen, 0, n, 0, with_s=False))
    cols = prior_sample_device(gen, 0, n, 0, with_s=False)
    ms_res = timed(lambda: h.marginal_ll_soa(*cols[:4], s=None, out=ll))
    out[f"N{N}_L{1 + pt}"] = {"generated_ms": ms_gen, "resident_ms": ms_res, "prior_sample_kernel_ms": ms_cols,
                              "generated_samples_per_s": n / ms_gen * 1e3,
                              "resident_samples_per_s": n / ms_res * 1e3}
print(json.dumps(out))
