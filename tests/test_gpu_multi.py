"""Sharded paths: one process driving two GPUs, and two NCCL ranks (one per GPU) in SPMD
mode, must both reproduce the single-GPU result bit for bit -- same ll, same accepted
indices, same posterior samples.  The two-GPU tests are skipped on a single-GPU box; the
one-process sharded path (shard offsets, per-shard PCG64 offsets, the fused max exchange
through the shards' keys, the ordered merge of the accepted indices, the iterative sampler
over shards, a drawn prior split over shards) is also run with two and three shards on ONE
GPU, which a single-GPU box can do."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _setup():
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200.synthetic import make_data

    prior = default_prior(2, sigma_K0=25.0, P_min=5.0, P_max=500.0)
    flat, _ = make_data(12, rng=np.random.default_rng(11), K=1e-4)
    ps = prior.sample(size=100_003, return_logprobs=True, rng=np.random.default_rng(1))
    return tj, prior, flat, ps


def _run(joker, flat, ps):
    out = {}
    out["ll"] = joker.marginal_ln_likelihood(flat, ps)
    s = joker.rejection_sample(flat, ps, max_posterior_samples=200, return_logprobs=True,
                               in_memory=True)
    out["rej_P"], out["rej_K"] = s["P"].value, s["K"].value
    out["rej_ll"] = s["ln_likelihood"].value
    s = joker.iterative_rejection_sample(flat, ps, n_requested_samples=64, in_memory=True)
    out["it_P"], out["it_K"] = s["P"].value, s["K"].value
    out["n_eval"] = joker.last_stats["n_ll_evaluated"]
    # a prior that is drawn (counter-based generator): sample g depends on (seed, g) only,
    # so every sharding evaluates the same prior and accepts the same rows
    s = joker.rejection_sample(flat, 1 << 17, max_posterior_samples=100, return_logprobs=True)
    out["gen_P"], out["gen_K"] = s["P"].value, s["K"].value
    out["gen_lp"], out["gen_ll"] = s["ln_prior"].value, s["ln_likelihood"].value
    return out


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0]])
def test_one_process_several_shards_on_one_gpu(devices):
    """devices=[0, 0]: the same sharded code path as two GPUs (separate handles, streams and
    max keys; the kernels' epilogues max-update the other shards' keys), on one device."""
    if _n_gpus() < 1:
        pytest.skip("needs a GPU")
    tj, prior, flat, ps = _setup()
    a = _run(tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[0]), flat, ps)
    b = _run(tj.TheJoker(prior, rng=np.random.default_rng(42), devices=devices), flat, ps)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    from thejoker_b200.sharding import DeviceEngine
    j2 = tj.TheJoker(prior, rng=np.random.default_rng(42), devices=devices)
    helper0 = j2._make_joker_helper(flat)
    cols, _ = j2._columns(helper0, ps)
    eng, _ = j2._engine(flat, cols)
    assert isinstance(eng, DeviceEngine) and eng.peer_max and len(eng.shards) == len(devices)
    assert len({id(sh.helper) for sh in eng.shards}) == len(devices)
    eng.compute_ll()
    eng.global_max_key()
    eng.synchronize()
    want = float(np.max(a["ll"]))
    assert [sh.helper.llmax_value(sh.key) for sh in eng.shards] == [want] * len(devices)


def test_one_process_two_gpus():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    tj, prior, flat, ps = _setup()
    a = _run(tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[0]), flat, ps)
    b = _run(tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[0, 1]), flat, ps)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    # the max exchange is fused into the kernel epilogue over NVLink peer stores
    from thejoker_b200.sharding import DeviceEngine
    j2 = tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[0, 1])
    helper0 = j2._make_joker_helper(flat)
    cols, _ = j2._columns(helper0, ps)
    eng, _ = j2._engine(flat, cols)
    assert isinstance(eng, DeviceEngine) and eng.peer_max
    eng.compute_ll()
    eng.global_max_key()
    eng.synchronize()
    want = float(np.max(a["ll"]))
    assert [sh.helper.llmax_value(sh.key) for sh in eng.shards] == [want, want]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device(f"cuda:{rank}"))
    tj, prior, flat, ps = _setup()
    joker = tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[rank],
                        group=dist.group.WORLD)
    out = _run(joker, flat, ps)
    # did the accept go through the library's own NCCL communicator?
    helper0 = joker._make_joker_helper(flat)
    eng, _ = joker._engine(flat, joker._columns(helper0, ps)[0])
    out["lib_collectives"] = np.array(eng._lib_collectives())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.destroy_process_group()


def test_one_nccl_rank_library_collectives(tmp_path):
    """The cross-rank accept inside the library (tjb_comm_create, tjb_accept_dist: NCCL bound at
    run time, MAX all-reduce, count / index all-gathers) on a world of ONE rank: the same code
    path as test_two_nccl_ranks_spmd, runnable on a single-GPU box."""
    if _n_gpus() < 1:
        pytest.skip("needs a GPU")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(1, port, str(tmp_path)), nprocs=1, join=True)
    tj, prior, flat, ps = _setup()
    want = _run(tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[0]), flat, ps)
    got = np.load(tmp_path / "rank0.npz")
    for k in want:
        assert np.array_equal(want[k], got[k]), k
    assert bool(got["lib_collectives"])


def test_two_nccl_ranks_spmd(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    tj, prior, flat, ps = _setup()
    want = _run(tj.TheJoker(prior, rng=np.random.default_rng(42), devices=[0]), flat, ps)
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        for k in want:
            assert np.array_equal(want[k], got[k]), (rank, k)
        assert bool(got["lib_collectives"])


def test_multistar_sharded_over_two_gpus():
    """Stars sharded over two GPUs give the same per-star samples as one GPU (per-star
    child RNG streams make the result independent of the sharding)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import thejoker_b200 as tj
    from helpers import default_prior
    from thejoker_b200.synthetic import make_noisy_data

    prior = default_prior(1, sigma_K0=25.0)
    ps = prior.sample(size=1 << 15, rng=np.random.default_rng(1))
    stars = [make_noisy_data(int(n), seed=50 + i, K=1e-4)[0]
             for i, n in enumerate([12, 20, 9, 33, 16])]
    a = tj.MultiStarJoker(prior, ps, rng=np.random.default_rng(5), devices=[0]) \
        .rejection_sample(stars, max_posterior_samples=32)
    b = tj.MultiStarJoker(prior, ps, rng=np.random.default_rng(5), devices=[0, 1]) \
        .rejection_sample(stars, max_posterior_samples=32)
    assert len(a) == len(b) == 5
    for sa, sb in zip(a, b):
        assert len(sa) == len(sb) > 0
        for k in ("P", "K", "v0"):
            assert np.array_equal(sa[k].value, sb[k].value)
