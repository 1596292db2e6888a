"""world_size-2 CPU (gloo) test of the N>1 host logic: sharding, the value all-gather, the reference's top-20 selection
order and the deterministic arg-max must give the same answer as the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _objective(x):
    x = np.asarray(x).reshape(len(x), -1)
    return -((x - 0.3) ** 2).sum(axis=1) + 0.1 * np.sin(7 * x).sum(axis=1)


def _descend(x):
    x = np.asarray(x, dtype=np.float64).copy()
    flat = x.reshape(len(x), -1)
    for _ in range(25):  # a deterministic stand-in for the device gradient descent
        g = -2 * (flat - 0.3) + 0.7 * np.cos(7 * flat)
        flat += 0.05 * g
    return _objective(flat), x


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, starts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cornell_moe_b200 import multigpu
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = multigpu.sharded_multistart(_objective, _descend, starts)
        q.put((rank, res[0], res[1], res[2], res[3]))
    finally:
        dist.destroy_process_group()


def test_top_k_matches_reference_priority_queue():
    from cornell_moe_b200 import multigpu
    import heapq  # noqa: F401
    rng = np.random.default_rng(0)
    for n in (5, 20, 21, 100):
        v = rng.standard_normal(n)
        v[rng.integers(0, n, 3)] = v[0]  # ties
        top = multigpu.top_k_indices(v)
        # C++ semantics: std::priority_queue<std::pair<double,int>> of (-value, idx), keep k, pop all
        pq = []
        k = min(20, n)
        for i in range(n):
            item = (-v[i], i)
            if i < k:
                pq.append(item)
            elif max(pq) > item:
                pq.remove(max(pq))
                pq.append(item)
        expect = [i for _, i in sorted(pq, reverse=True)]
        assert top == expect
        assert set(top) == set(np.argsort(-v, kind="stable")[:k]) or len(set(v)) < n


def test_sharded_multistart_world2_matches_single_process():
    import torch.multiprocessing as mp
    from cornell_moe_b200 import multigpu
    rng = np.random.default_rng(3)
    starts = rng.uniform(size=(53, 2, 3))
    single = multigpu.sharded_multistart(_objective, _descend, starts)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, starts, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, pt, val, found, values in results:
        np.testing.assert_array_equal(values, single[3])
        np.testing.assert_array_equal(pt, single[0])
        assert val == single[1] and found == single[2]
    assert single[2] is True


def test_gather_values_single_process_identity():
    from cornell_moe_b200 import multigpu
    v = np.arange(7.0)
    np.testing.assert_array_equal(multigpu.gather_values(v, 7), v)
    np.testing.assert_array_equal(multigpu.shard_indices(10, 1, 4), [1, 5, 9])


class _FakeEnsemble:
    """Duck-typed stand-in for capi.GaussianProcessEnsemble: deterministic host functions with the same method
    signatures, so the wiring of the sharded MCMC drivers can be checked on CPU."""

    def kg(self, sub, Xp, num_mc, best_so_far, inner, inner_bounds, discrete_pts, num_fidelity=0, seed=0):
        assert Xp is None and num_mc == 64 and inner == "inner" and inner_bounds == "ib" and discrete_pts == "disc"
        return _objective(sub) + 1e-3 * seed + float(np.sum(best_so_far))

    def kg_gradient_descent(self, sub, Xp, num_mc, best_so_far, outer, inner, domain_bounds, inner_bounds, discrete_pts,
                            num_fidelity=0, seed=0):
        assert outer == "outer" and domain_bounds == "db" and inner_bounds == "ib"
        v, p = _descend(sub)
        return v + 1e-3 * seed + float(np.sum(best_so_far)), p

    def ei(self, sub, Xp, num_mc, best_so_far, seed=0, analytic_single=False):
        assert analytic_single is True
        return np.maximum(0.0, _objective(sub) + 0.5)

    def ei_gradient_descent(self, sub, Xp, num_mc, best_so_far, outer, domain_bounds, seed=0):
        v, p = _descend(sub)
        return np.maximum(0.0, v + 0.5), p


def _worker_mcmc(rank, world, port, starts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from cornell_moe_b200 import multigpu
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ens = _FakeEnsemble()
        a = multigpu.multistart_kg_mcmc(ens, starts, None, 64, np.array([0.25, 0.5]), "outer", "inner", "db", "ib", "disc",
                                        seed=7)
        b = multigpu.multistart_ei_mcmc(ens, starts, None, 64, np.array([0.25, 0.5]), "outer", "db", seed=7)
        q.put((rank, a, b))
    finally:
        dist.destroy_process_group()


def test_sharded_mcmc_drivers_world2_match_single_process():
    import torch.multiprocessing as mp
    from cornell_moe_b200 import multigpu
    rng = np.random.default_rng(5)
    starts = rng.uniform(size=(41, 2, 3))
    ens = _FakeEnsemble()
    single_kg = multigpu.multistart_kg_mcmc(ens, starts, None, 64, np.array([0.25, 0.5]), "outer", "inner", "db", "ib",
                                            "disc", seed=7)
    single_ei = multigpu.multistart_ei_mcmc(ens, starts, None, 64, np.array([0.25, 0.5]), "outer", "db", seed=7)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_mcmc, args=(r, 2, port, starts, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b in results:
        for got, want in ((a, single_kg), (b, single_ei)):
            np.testing.assert_array_equal(got[0], want[0])
            assert got[1] == want[1] and got[2] == want[2]
            np.testing.assert_array_equal(got[3], want[3])
    # the EI driver starts from 0.0 (gpp_expected_improvement_mcmc_optimization.hpp:914), the KG driver from -inf
    assert single_kg[2] is True and single_ei[1] >= 0.0
