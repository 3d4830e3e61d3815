"""cProfile of the multi-star driver's host side (one slot, so the profile is serial)."""
import cProfile
import os
import pstats
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = ["bench_multistar.py", "8", "22", "1"]
ns = runpy.run_path(os.path.join(ROOT, "tools", "bench_multistar.py"))
ms, stars = ns["ms"], ns["stars"]
import numpy as np  # noqa: E402
import thejoker_b200 as tj  # noqa: E402

rng = np.random.default_rng(0)
more = stars * 32
pr = cProfile.Profile()
pr.enable()
ms.rejection_sample(more, max_posterior_samples=256)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(32)
