"""First-contact script for a GPU box: smoke, FP64 peak, kernel timing at a few sizes."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.smoke()
import thejoker_b200 as tj  # noqa: E402
from helpers import prior_chunk, star_spec  # noqa: E402
from oracle.oracle import OracleHelper  # noqa: E402
from thejoker_b200.data_helpers import validate_prepare_data  # noqa: E402

out = {}
for N, pt, sl in [(64, 1, None), (64, 2, (-2, 1)), (256, 1, None), (16, 1, None)]:
    spec, data, prior = star_spec(N, pt)
    all_data, ids, trend_M = validate_prepare_data(data, prior.poly_trend, prior.n_offsets)
    helper = tj.CJokerHelper(all_data, prior, trend_M, device=0)
    if N == 64 and pt == 1:
        out["fp64_peak_tflops"] = helper.fp64_peak(40000)
        print("fp64 peak", out["fp64_peak_tflops"], flush=True)
    n = 1 << 22
    chunk = prior_chunk(n, s_lognormal=sl)
    cols = [torch.from_numpy(np.ascontiguousarray(chunk[:, i])).cuda() for i in range(5)]
    s = cols[4] if sl is not None else None
    ll = torch.empty(n, dtype=torch.float64, device="cuda")
    key = helper.new_llmax_key()
    for _ in range(2):
        helper.marginal_ll_soa(*cols[:4], s=s, out=ll, llmax_key=key)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        helper.marginal_ll_soa(*cols[:4], s=s, out=ll, llmax_key=key)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    m = 1 << 13
    ref = OracleHelper.from_spec(spec).batch_marginal_ln_likelihood(chunk[:m], n_threads=0)
    got = ll[:m].cpu().numpy()
    rel = np.max(np.abs(got - ref) / np.abs(ref))
    rec = dict(N=N, poly_trend=pt, jitter=sl is not None, n=n, ms=ms, samples_per_s=n / ms * 1e3,
               max_rel=float(rel), info=helper.device_info(), llmax=helper.llmax_value(key),
               llmax_ref=float(np.max(ll.cpu().numpy())))
    print(json.dumps(rec), flush=True)
    out[f"N{N}_pt{pt}_{'jit' if sl else 'const'}"] = rec
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sanity.json"), "w"), indent=1)
