"""BASELINE.json configs[4] (APOGEE-like batch): stars x ~20 epochs x 2^22 shared prior
samples with a two-survey v0 offset, through MultiStarJoker on one GPU.  Times a subset of
stars and reports stars/s and prior-sample evaluations/s.  Run on the GPU box."""
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
from thejoker_b200 import units as u  # noqa: E402
from thejoker_b200.prior import Normal  # noqa: E402
from thejoker_b200.synthetic import make_noisy_data  # noqa: E402

n_stars = int(sys.argv[1]) if len(sys.argv) > 1 else 256
log2_prior = int(sys.argv[2]) if len(sys.argv) > 2 else 22
streams = int(sys.argv[3]) if len(sys.argv) > 3 else 4
engine = sys.argv[4] if len(sys.argv) > 4 else "native"  # or "python": the loop driven from Python
gather = os.environ.get("TJB_MS_GATHER", "1") != "0"  # ranks exchange results (every rank returns all stars)
rng = np.random.default_rng(0)
prior = default_prior(1, sigma_K0=30.0, v0_offsets=[Normal("dv0_1", 0.0, 5.0, u.km / u.s)])
ps = prior.sample(size=1 << log2_prior, rng=np.random.default_rng(1))
stars = []
for i in range(n_stars):
    n = int(np.clip(rng.poisson(20), 8, 40))
    full, _ = make_noisy_data(n, seed=1000 + i, K=float(rng.choice([54.0, 5.0, 1e-4])))
    cut = int(rng.integers(3, n - 3))
    off = rng.normal(0, 5.0)
    stars.append([tj.RVData(full._t_bmjd[:cut], full.rv[:cut], full.rv_err[:cut]),
                  tj.RVData(full._t_bmjd[cut:], (full.rv.value[cut:] + off) * u.km / u.s,
                            full.rv_err[cut:])])
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
group = None
if world > 1:  # torchrun: one rank per GPU, stars sharded over the ranks
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    group = dist.group.WORLD
ms = tj.MultiStarJoker(prior, ps, rng=np.random.default_rng(2), devices=[local],
                       streams_per_device=streams, group=group, engine=engine)
ms.rejection_sample(stars[:4], max_posterior_samples=256)  # upload + warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = ms.rejection_sample(stars, max_posterior_samples=256, gather=gather)  # includes the result exchange
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if dist.get_rank() != 0:
        dist.destroy_process_group()
        sys.exit(0)
rec = dict(n_gpus=world, engine=engine, streams_per_device=streams, n_stars=n_stars, n_prior=1 << log2_prior, seconds=dt, stars_per_s=n_stars / dt,
           prior_evaluations_per_s=n_stars * (1 << log2_prior) / dt,
           mean_epochs=float(np.mean([len(s[0]) + len(s[1]) for s in stars])),
           mean_posterior_samples=float(np.mean([len(o) for o in out if o is not None])),
           gather=gather, rank0_timing={k: round(v, 4) for k, v in ms.last_timing.items()},
           extrapolated_4096_stars_s=4096 * dt / n_stars)
print(json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", f"multistar_g{world}_s{streams}_{engine}{'' if gather else '_nogather'}.json"), "w"),
          indent=1)
if world > 1:
    dist.destroy_process_group()
