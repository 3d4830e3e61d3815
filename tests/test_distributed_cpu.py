"""The N>1 path on CPU: two gloo ranks each own a contiguous shard of the prior
(batch_tasks rule), compute ll for it (CPU oracle standing in for the kernel), combine
the max through the integer-key all-reduce and the accepted indices through the
rank-ordered gather -- and must reproduce the single-process accept exactly
(thejoker/multiproc_helpers.py:256-263)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, max_keep, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import prior_chunk, star_spec

    from oracle.oracle import OracleHelper
    from thejoker_b200 import _lib
    from thejoker_b200.sharding import allreduce_max_key, gather_accepted, shard_ranges

    spec, _, _ = star_spec(16, 1, K=1e-4)
    chunk = prior_chunk(n)
    lo, hi = shard_ranges(n, world)[rank]
    ll = OracleHelper.from_spec(spec).batch_marginal_ln_likelihood(chunk[lo:hi])
    if rank == 1:
        ll[3] = ll.max() + 1.0  # the global max lives on rank 1
    lib = _lib.load()
    key = torch.tensor([lib.tjb_double_to_key(float(ll.max()))], dtype=torch.int64)
    allreduce_max_key(key)
    gmax = lib.tjb_key_to_double(int(key.item()))
    uu = np.random.default_rng(7).uniform(size=n)[lo:hi]  # sample g uses the g-th uniform
    good = np.where(np.exp(ll - gmax) > uu)[0] + lo
    if max_keep is not None:
        good = good[:max_keep]
    idx, total, near = gather_accepted(good, len(np.where(np.exp(ll - gmax) > uu)[0]), 0, max_keep)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), idx=idx, total=total, gmax=gmax, ll=ll, lo=lo)
    dist.destroy_process_group()


@pytest.mark.parametrize("max_keep", [None, 5])
def test_two_rank_accept_matches_single_process(tmp_path, max_keep):
    n, world = 6001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, max_keep, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(world)]
    ll = np.concatenate([r[0]["ll"], r[1]["ll"]])
    assert len(ll) == n and r[1]["lo"] == 3001  # first shard gets the remainder
    uu = np.random.default_rng(7).uniform(size=n)
    want = np.where(np.exp(ll - ll.max()) > uu)[0]
    assert r[0]["gmax"] == r[1]["gmax"] == ll.max()
    for k in range(world):
        assert r[k]["total"] == len(want)
        assert np.array_equal(r[k]["idx"], want if max_keep is None else want[:max_keep])


def _multistar_worker(rank, world, port, n_stars, out_dir):
    """Two gloo ranks run the multi-star driver over the same star list (the native call
    is replaced by the echoing stand-in of tests/helpers.py: no GPU here)."""
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pickle

    from helpers import fake_multistar_joker, fake_multistar_lib, fake_multistar_stars

    from thejoker_b200 import _lib

    n_prior, keep, n_per, L = 64, 5, 2, 3
    prior, stars = fake_multistar_stars(n_stars)
    lib, calls = fake_multistar_lib(n_prior, keep, n_per, L)
    _lib.load = lambda: lib
    res = {}
    for gather in (True, False):
        ms = fake_multistar_joker(prior, n_prior, L, rng=np.random.default_rng(9),
                                  group=dist.group.WORLD)
        out = ms.rejection_sample(stars, max_posterior_samples=keep, n_linear_samples=n_per,
                                  return_logprobs=True, gather=gather)
        res[gather] = ([None if o is None else
                        {k: np.asarray(getattr(o[k], "value", o[k])) for k in ("P", "K", "dv0_1",
                                                                                "ln_likelihood")}
                        for o in out], ms.last_stats, [None if o is None else o.t_ref for o in out])
    res["n_computed"] = sum(calls)
    with open(os.path.join(out_dir, f"ms{rank}.pkl"), "wb") as f:
        pickle.dump(res, f)
    dist.destroy_process_group()


def test_two_rank_multistar_sharding_and_exchange(tmp_path, monkeypatch):
    """Stars are sharded over the ranks (batch_tasks rule), each star is computed exactly
    once, and after the packed exchange every rank holds what a single process computes."""
    import pickle

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import fake_multistar_joker, fake_multistar_lib, fake_multistar_stars

    from thejoker_b200 import _lib
    from thejoker_b200.sharding import shard_ranges

    n_stars, world = 23, 2
    mp.spawn(_multistar_worker, args=(world, _free_port(), n_stars, str(tmp_path)), nprocs=world,
             join=True)
    prior, stars = fake_multistar_stars(n_stars)
    lib, _ = fake_multistar_lib(64, 5, 2, 3)
    monkeypatch.setattr(_lib, "load", lambda: lib)
    ms = fake_multistar_joker(prior, 64, 3, rng=np.random.default_rng(9))
    want = ms.rejection_sample(stars, max_posterior_samples=5, n_linear_samples=2,
                               return_logprobs=True)
    r = [pickle.load(open(tmp_path / f"ms{k}.pkl", "rb")) for k in range(world)]
    assert r[0]["n_computed"] + r[1]["n_computed"] == 2 * n_stars  # two passes, each star once
    ranges = shard_ranges(n_stars, world)
    for k in range(world):
        full, stats, t_ref = r[k][True]
        assert stats == ms.last_stats
        for i in range(n_stars):
            assert t_ref[i] == want[i].t_ref
            for name in ("P", "K", "dv0_1", "ln_likelihood"):
                assert np.array_equal(full[i][name],
                                      np.asarray(getattr(want[i][name], "value", want[i][name])))
        own, own_stats, _ = r[k][False]
        lo, hi = ranges[k]
        for i in range(n_stars):
            assert (own[i] is not None) == (lo <= i < hi) == (own_stats[i] is not None)
            if own[i] is not None:
                assert np.array_equal(own[i]["K"], full[i]["K"])


# ---- iterative sampler under SPMD ranks: block-cyclic rounds ---------------------------------
class _CpuStandInHelper:
    """What sharding.DeviceEngine needs from a CJokerHelper, on CPU tensors, with the
    oracle standing in for the kernel (test infrastructure: no GPU in this container)."""

    def __init__(self, spec, data):
        from oracle.oracle import OracleHelper
        from thejoker_b200 import _lib

        self.spec, self.device = spec, "cpu"
        self.internal_units = spec["internal_units"]
        self.packed_order = ["P", "e", "omega", "M0", "s"]
        self.n_linear = spec["n_linear"]
        self.data = data
        self._orc = OracleHelper.from_spec(spec)
        self._lib = _lib.load()
        self.n_ll = 0

    def new_llmax_key(self):
        return torch.tensor([self._lib.tjb_double_to_key(float("-inf"))], dtype=torch.int64)

    def llmax_value(self, key):
        return self._lib.tjb_key_to_double(int(key.item()))

    def marginal_ll_host_columns(self, P, e, omega, M0, s=None, s_const=0.0, out=None,
                                 llmax_key=None):
        n = len(P)
        chunk = np.stack([P, e, omega, M0, np.full(n, s_const) if s is None else s], axis=1)
        ll = self._orc.batch_marginal_ln_likelihood(np.ascontiguousarray(chunk))
        self.n_ll += n
        out.copy_(torch.from_numpy(ll))
        if llmax_key is not None and n:
            llmax_key[0] = max(int(llmax_key.item()), self._lib.tjb_double_to_key(float(ll.max())))
        return out

    def accept(self, ll, llmax_key, uniforms=None, rng=None, rng_offset=0, index_base=0,
               max_keep=None, near_tol=1e-12):
        n = ll.numel()
        if uniforms is None:  # the (rng_offset + i)-th double of the generator, not advanced
            bg = np.random.PCG64()
            bg.state = rng.bit_generator.state
            bg.advance(int(rng_offset))
            uu = np.random.Generator(bg).random(n)
        else:
            uu = uniforms.numpy()
        a = np.exp(ll.numpy() - self.llmax_value(llmax_key))
        good = np.where(a > uu)[0]
        keep = good if max_keep is None else good[:max_keep]
        return torch.from_numpy(keep + index_base), len(good), int(np.sum(np.abs(a - uu) <= near_tol))

    def batch_get_posterior_samples(self, rows, n_linear_samples_per, rng, draw="auto"):
        rows = np.asarray(rows).reshape(-1, 5)
        out = np.zeros((len(rows) * n_linear_samples_per, 5 + self.n_linear))
        out[:, :5] = np.repeat(rows, n_linear_samples_per, axis=0)
        return out, np.zeros(len(out))


def _iterative_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    group = None
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        group = dist.group.WORLD
    from helpers import prior_chunk, star_spec

    import thejoker_b200 as tj

    spec, data, prior = star_spec(12, 1, K=2.0, sigma=1.0)   # informative: several rounds
    helper = _CpuStandInHelper(spec, data)
    chunk = prior_chunk(60_000, seed=5)
    joker = tj.TheJoker(prior, rng=np.random.default_rng(11), devices=["cpu"], group=group)
    joker._make_joker_helper = lambda data, device=None: helper
    smp = joker.iterative_rejection_sample(data, chunk, n_requested_samples=24, init_batch_size=512,
                                           growth_factor=8, in_memory=True, return_logprobs=False)
    np.savez(os.path.join(out_dir, f"it{world}_{rank}.npz"), P=smp["P"].value, e=smp["e"].value,
             n_ll=helper.n_ll, n_eval=joker.last_stats["n_ll_evaluated"],
             n_acc=joker.last_stats["n_accepted"])
    if world > 1:
        dist.destroy_process_group()


def test_iterative_sampler_block_cyclic_matches_single_process(tmp_path):
    """iterative_rejection_sample under two SPMD ranks: every round's range is split over
    both ranks (block-cyclic; multiproc_helpers.py:355-410 maps each round over the whole
    pool), each segment is accepted with its own PCG offset, and both ranks return exactly
    what one process returns -- while each evaluates about half of the rows."""
    _iterative_worker(0, 1, 0, str(tmp_path))
    mp.spawn(_iterative_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    one = np.load(tmp_path / "it1_0.npz")
    two = [np.load(tmp_path / f"it2_{k}.npz") for k in range(2)]
    assert len(one["P"]) == 24 and one["n_eval"] > 512  # more than one round was needed
    for r in two:
        assert np.array_equal(r["P"], one["P"]) and np.array_equal(r["e"], one["e"])
        assert r["n_eval"] == one["n_eval"] and r["n_acc"] == one["n_acc"]
    assert two[0]["n_ll"] + two[1]["n_ll"] == one["n_ll"]
    assert abs(int(two[0]["n_ll"]) - int(two[1]["n_ll"])) <= 8  # balanced in every round
