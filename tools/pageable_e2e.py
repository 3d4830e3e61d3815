"""e2e rate of the likelihood on ordinary (pageable) numpy columns with a pageable ll array
out -- what TheJoker.marginal_ln_likelihood(data, JokerSamples) moves -- for a given number
of copy threads (TJB_COPY_THREADS, read when the library's copy pool starts).
usage: [TJB_COPY_THREADS=k] python tools/pageable_e2e.py [log2_n]   (on the GPU box)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import thejoker_b200 as tj  # noqa: E402
from helpers import star_spec  # noqa: E402
from thejoker_b200.synthetic import default_prior_columns  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
spec, data, prior = star_spec(64, 1)
helper = tj.CJokerHelper.from_spec(spec, device=0)
cols = [np.ascontiguousarray(c) for c in default_prior_columns(1 << 22, seed=1)[:4]]
cols = [np.tile(c, n >> 22) for c in cols]
ll = np.empty(n)
helper.marginal_ln_likelihood_columns(*cols, s=None, s_const=0.0, out=ll)
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    helper.marginal_ln_likelihood_columns(*cols, s=None, s_const=0.0, out=ll)
    ts.append(time.perf_counter() - t0)
print(json.dumps({"copy_threads": os.environ.get("TJB_COPY_THREADS", "default"), "n": n,
                  "samples_per_s": n / min(ts), "host_GBs_in": 32 * n / min(ts) / 1e9,
                  "lib": os.path.basename(os.environ.get("TJB_LIB_PATH", "shipped"))}))
