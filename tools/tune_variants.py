"""Time every build/variants/*.so (kernel tuning builds) on the GPU box.
Each variant runs in its own process (the library path is fixed at import)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import thejoker_b200 as tj
from helpers import star_spec
from thejoker_b200.data_helpers import validate_prepare_data
out = {"lib": os.path.basename(os.environ["TJB_LIB_PATH"])}
n = 1 << 24
g = torch.Generator(device="cuda").manual_seed(1)
P = torch.exp(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * np.log(512.0) + np.log(2.0))
torch.manual_seed(1)
ga = torch._standard_gamma(torch.full((n,), 0.867, dtype=torch.float64, device="cuda"))
gb = torch._standard_gamma(torch.full((n,), 3.03, dtype=torch.float64, device="cuda"))
e = (ga / (ga + gb)).contiguous()
om = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * np.pi
M0 = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * np.pi
s = torch.exp(torch.randn(n, dtype=torch.float64, device="cuda", generator=g) - 2)
ll = torch.empty(n, dtype=torch.float64, device="cuda")
# span: time baseline in units of the 51.8 d fixture period (3 = the benchmark data, 155 d;
# 77 = 4000 d, where the float phase reduction of the shipped loop starts to send warps
# through the extra-pass path: tests/test_host_logic.py::test_phase_reduction_variants)
shapes = ((64, 1, False, 3.0), (64, 2, True, 3.0), (256, 1, False, 3.0),
          (16, 1, False, 3.0), (64, 1, False, 77.0), (64, 1, False, 193.0))
if os.environ.get("TJB_TUNE_SHAPES") == "L":  # how a CTA shape holds up as n_linear grows
    shapes = ((64, 1, False, 3.0), (64, 2, False, 3.0), (64, 3, False, 3.0), (64, 1, True, 3.0),
              (64, 2, True, 3.0), (64, 3, True, 3.0), (20, 2, True, 3.0))
for N, pt, jit, span in shapes:
    spec, data, prior = star_spec(N, pt, t_span_periods=span)
    all_data, ids, trend_M = validate_prepare_data(data, prior.poly_trend, prior.n_offsets)
    h = tj.CJokerHelper(all_data, prior, trend_M, device=0)
    m = n if N <= 64 else n // 4
    args = [t[:m] for t in (P, e, om, M0)]
    for _ in range(2):
        h.marginal_ll_soa(*args, s=s[:m] if jit else None, out=ll[:m])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        h.marginal_ll_soa(*args, s=s[:m] if jit else None, out=ll[:m])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tag = f"N{N}_L{1+pt}_{'jit' if jit else 'const'}" + ("" if span == 3.0 else f"_span{int(span * 51.8239)}d")
    out[tag] = round(m / ms * 1e3 / 1e9, 4)
    out["ctas_per_sm"] = h.device_info()["ctas_per_sm"]
    out["chk_" + tag] = float(ll[:1000].sum().item())
    out["extra_passes_" + tag] = h.solver_stats(reset=True)["extra_fp64_passes"] // 7  # per launch
print(json.dumps(out))
''' % (ROOT, ROOT)

results = []
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))):
    env = dict(os.environ, TJB_LIB_PATH=lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(line, flush=True)
    results.append(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "tune.jsonl"), "w").write("\n".join(results) + "\n")
