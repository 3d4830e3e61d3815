"""cProfile of TheJoker.rejection_sample with a host-resident prior (GPU box)."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import thejoker_b200 as tj  # noqa: E402
from helpers import default_prior  # noqa: E402
from thejoker_b200.synthetic import make_noisy_data  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
prior = default_prior(1, sigma_K0=30.0, P_min=2.0, P_max=1024.0)
flat, _ = make_noisy_data(64, seed=42, K=1e-4)
ps = prior.sample(size=1 << log2n, rng=np.random.default_rng(1))
joker = tj.TheJoker(prior, rng=np.random.default_rng(42))
mode = sys.argv[2] if len(sys.argv) > 2 else "rejection"
if mode == "iterative":
    run = lambda: joker.iterative_rejection_sample(flat, ps, n_requested_samples=256)
else:
    run = lambda: joker.rejection_sample(flat, ps, max_posterior_samples=256, in_memory=True)
run()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
run()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
