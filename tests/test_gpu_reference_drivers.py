"""GPU parity against the reference's own multistart DRIVERS (not only its evaluators), multi-device sharding inside the
C ABI, the incremental append and the batched posterior-mean screening call.

The reference drivers (ComputeKGOptimalPointsToSampleViaMultistartGradientDescent, gpp_knowledge_gradient_optimization.hpp:
859-935, and ComputeOptimalPointsToSampleViaMultistartGradientDescent, gpp_math.hpp:1683-1802) run UNMODIFIED through
oracle/ref_driver.cpp with one thread and NormalRNG(seed).  Every evaluation inside a driver call rewinds that generator,
so the whole call replays the same first draws; `oracle.normal_draws(seed, n)` returns them and the device path gets them
as `normals_table` — both sides consume identical normals (checked bit-for-bit on the CPU in
tests/test_oracle_vs_reference.py::test_normal_rng_table_replay)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as orc
from synth import EXAMPLE_INNER_GD, make_problem, unit_bounds

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not orc.have_reference(), reason="compiled reference (oracle/_ref) did not travel")


@pytest.fixture(scope="module")
def capi():
    from cornell_moe_b200 import capi as c
    assert c.device_count() > 0
    return c


def _kg_setup(capi, kernel=0, N=30, dim=3, q=2, ns=40, seed=3):
    prob = make_problem(N, dim, seed=seed, noise=0.05)
    gp = capi.GaussianProcess(kernel, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    rng = np.random.default_rng(seed + 1)
    starts = rng.uniform(0.05, 0.95, size=(ns, q, dim))
    disc = rng.uniform(size=(8, dim))
    return prob, gp, starts, disc


def _agreeing(pairs, tol):
    """Number of (ours, reference) point pairs that agree to `tol`."""
    return sum(1 for a, b in pairs if np.abs(a - b).max() <= tol)


@needs_ref
@pytest.mark.parametrize("kernel", [0, 1])
def test_multistart_kg_matches_reference_driver(capi, kernel):
    """Whole pipeline (screen 40 starts -> keep 20 -> restarted gradient descent -> strict arg-max) against the
    reference's own driver on identical normals.  With ONE descent step per start the comparison is exact (measured
    agreement 1e-15); longer descents agree to 1e-15 as well except where a trajectory crosses a kink of the MC objective
    (inner arg-min switch / Armijo decision) and a last-bit difference picks the other branch — so those are compared
    start by start and a majority must coincide."""
    prob, gp, starts, disc = _kg_setup(capi, kernel)
    ref, lm = orc.load_reference().gp(kernel, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    assert lm == 0
    q, mc, seed = starts.shape[1], 64, 4242
    best = float(ref.mean_additional(disc).min())
    b3 = unit_bounds(3)
    table = orc.normal_draws(seed, (mc // 2) * q)
    outer1 = [40, 1, 1, 0, 0.7, 0.4, 0.2, 1e-7]
    bp_ref, found_ref = orc.ref_multistart_kg(ref, starts, None, mc, best, outer1, EXAMPLE_INNER_GD, b3, b3, disc, seed)
    bp, bv, found, sv = capi.multistart_kg(gp, starts, None, mc, best, outer1, EXAMPLE_INNER_GD, b3, b3, disc, seed=1,
                                           table=table)
    for i in (0, 7, 19, 39):  # the screening values the driver ranks the starts by: ONE state, built with start 0
        v = ref.kg_reused_state(starts[0], starts[i], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc)
        np.testing.assert_allclose(sv[i], v, rtol=1e-6, atol=1e-9)
    # ... which is not what a freshly constructed state gives (the reference quirk the drivers reproduce)
    fresh = capi.multistart_kg(gp, starts, None, mc, best, outer1, EXAMPLE_INNER_GD, b3, b3, disc, seed=1, table=table,
                               fresh_discretisation=True)[3]
    np.testing.assert_allclose(fresh[7], ref.kg(starts[7], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc), rtol=1e-6,
                               atol=1e-9)
    assert found == found_ref
    np.testing.assert_allclose(bp, bp_ref, rtol=0, atol=1e-9)
    v_ref = ref.kg_reused_state(starts[0], bp_ref, None, mc, best, table, EXAMPLE_INNER_GD, b3, disc)
    np.testing.assert_allclose(bv, v_ref, rtol=1e-7, atol=1e-10)
    # restarted descent, one start per driver call on both sides
    outer = [1, 3, 1, 0, 0.7, 0.4, 0.2, 1e-7]
    pairs = []
    for i in range(10):
        r_pt, _ = orc.ref_multistart_kg(ref, starts[i:i + 1], None, mc, best, outer, EXAMPLE_INNER_GD, b3, b3, disc, seed)
        o_pt = capi.multistart_kg(gp, starts[i:i + 1], None, mc, best, outer, EXAMPLE_INNER_GD, b3, b3, disc, seed=1,
                                  table=table)[0]
        pairs.append((o_pt, r_pt))
    assert _agreeing(pairs, 1e-9) >= 6, [float(np.abs(a - b).max()) for a, b in pairs]


@needs_ref
@pytest.mark.parametrize("q", [1, 2])
def test_multistart_ei_matches_reference_driver(capi, q):
    prob = make_problem(25, 3, seed=8, noise=0.05)
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    ref, lm = orc.load_reference().gp(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    assert lm == 0
    rng = np.random.default_rng(9)
    starts = rng.uniform(0.05, 0.95, size=(35, q, 3))
    mc, seed = 256, 99
    best = float(prob["y"].min()) + 0.2
    table = orc.normal_draws(seed, mc * q)
    outer1 = [35, 1, 1, 0, 0.7, 0.5, 0.2, 1e-7]
    bp_ref = orc.ref_multistart_ei(ref, starts, None, mc, best, outer1, unit_bounds(3), seed)
    bp, bv, found, sv = capi.multistart_ei(gp, starts, None, mc, best, outer1, unit_bounds(3), seed=1, table=table)
    np.testing.assert_allclose(bp, bp_ref, rtol=0, atol=1e-9)
    assert found
    outer = [1, 8, 2, 0, 0.7, 0.5, 0.2, 1e-7]
    pairs = []
    for i in range(10):
        r_pt = orc.ref_multistart_ei(ref, starts[i:i + 1], None, mc, best, outer, unit_bounds(3), seed)
        o_pt = capi.multistart_ei(gp, starts[i:i + 1], None, mc, best, outer, unit_bounds(3), seed=1, table=table)[0]
        pairs.append((o_pt, r_pt))
    assert _agreeing(pairs, 1e-8) >= 7, [float(np.abs(a - b).max()) for a, b in pairs]


def test_multistart_kg_multi_device_is_bit_identical(capi):
    if capi.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _, gp, starts, disc = _kg_setup(capi, 0, N=40, dim=3, q=2, ns=50, seed=11)
    best = float(gp.posterior(disc[:, None, :], (), ("mean",))["mean"].min())
    outer = [50, 4, 1, 0, 0.7, 0.4, 0.2, 1e-7]
    b3 = unit_bounds(3)
    one = capi.multistart_kg(gp, starts, None, 128, best, outer, EXAMPLE_INNER_GD, b3, b3, disc, seed=5)
    devs = list(range(min(4, capi.device_count())))
    many = capi.multistart_kg(gp, starts, None, 128, best, outer, EXAMPLE_INNER_GD, b3, b3, disc, seed=5, devices=devs)
    np.testing.assert_array_equal(one[3], many[3])
    np.testing.assert_array_equal(one[0], many[0])
    assert one[1] == many[1] and one[2] == many[2]
    e1 = capi.multistart_ei(gp, starts, None, 512, 0.0, outer, b3, seed=5)
    e2 = capi.multistart_ei(gp, starts, None, 512, 0.0, outer, b3, seed=5, devices=devs)
    np.testing.assert_array_equal(e1[3], e2[3])
    np.testing.assert_array_equal(e1[0], e2[0])


def test_sharded_multistart_under_torchrun_equals_single_process(capi, tmp_path):
    """multigpu.multistart_kg with one process per GPU (NCCL) returns exactly what the single-process driver returns."""
    if capi.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    out = tmp_path / "sharded.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "mgpu_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = np.load(out)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mgpu_worker
    gp, args = mgpu_worker.problem(capi, 0)
    bp, bv, found, sv = capi.multistart_kg(gp, *args)
    np.testing.assert_array_equal(got["start_values"], sv)
    np.testing.assert_array_equal(got["best_point"], bp)
    assert float(got["best_value"]) == bv and bool(got["found"]) == found


def test_add_sampled_points_incremental_matches_refit(capi):
    """O(N^2) append (bordered Cholesky) against a fit of the enlarged training set, with and without derivative
    observations; a failing append (duplicate point, zero noise) leaves the handle usable."""
    for g_idx, N, m in [((), 300, 3), ((0, 2), 120, 2), ((), 1100, 5)]:
        prob = make_problem(N + m, 3, g_idx=g_idx, seed=31)
        b = 1 + len(g_idx)
        X, y = prob["X"], prob["y"].reshape(N + m, b)
        gp = capi.GaussianProcess(1, 1.3, prob["lengths"], X[:N], y[:N].ravel(), prob["noise"], prob["derivs"])
        gp.add_sampled_points(X[N:], y[N:].ravel())
        full = capi.GaussianProcess(1, 1.3, prob["lengths"], X, y.ravel(), prob["noise"], prob["derivs"])
        K1, a1, m1 = gp.state()
        K2, a2, m2 = full.state()
        assert m1 == m2
        np.testing.assert_allclose(np.tril(K1), np.tril(K2), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(a1, a2, rtol=1e-6, atol=1e-7)
        pts = np.random.default_rng(1).uniform(size=(4, 1, 3))
        np.testing.assert_allclose(gp.posterior(pts)["var"], full.posterior(pts)["var"], rtol=1e-7, atol=1e-10)
    prob = make_problem(50, 2, seed=2)
    prob["noise"][:] = 0.0
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    before = gp.state()
    with pytest.raises(capi.SingularMatrixError) as e:
        gp.add_sampled_points(prob["X"][7:8], prob["y"][7:8])
    assert e.value.info == 51
    after = gp.state()
    np.testing.assert_array_equal(before[0], after[0])
    assert gp.N == 50
    plan = capi.KGPlan(gp, 32, 0.0, EXAMPLE_INNER_GD, unit_bounds(2), prob["X"][:4], 2, 1)
    gp.add_sampled_points(np.array([[0.123, 0.456]]), np.array([0.3]))
    with pytest.raises(capi.InvalidValueError):  # the plan was sized for the previous fit
        plan.upload(np.zeros((1, 1, 2)))


@needs_ref
def test_posterior_mean_screening_batch(capi):
    """§8f rank 3: the 1e4-point posterior-mean screen the examples do point by point, as ONE device call."""
    import time
    prob = make_problem(500, 8, seed=17)
    gp = capi.GaussianProcess(1, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    ref, _ = orc.load_reference().gp(1, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    pts = np.random.default_rng(3).uniform(size=(10000, 8))
    gp.posterior(pts[:8, None, :], (), ("mean",))
    t0 = time.perf_counter()
    res = gp.posterior(pts[:, None, :], (), ("mean", "grad_mean"))
    dt = time.perf_counter() - t0
    mu = res["mean"].ravel()
    np.testing.assert_allclose(mu[:400], ref.mean_additional(pts[:400]), rtol=1e-9, atol=1e-10)
    one = gp.posterior(pts[None, :50, :], (), ("mean", "grad_mean"))  # 50 points as ONE set (> old 96-row cap at 200)
    np.testing.assert_allclose(one["mean"].ravel(), mu[:50], rtol=1e-13)
    np.testing.assert_allclose(one["grad_mean"].ravel(), res["grad_mean"][:50].ravel(), rtol=1e-12, atol=1e-14)
    big = gp.posterior(pts[None, :200, :], (), ("mean",))
    np.testing.assert_allclose(big["mean"].ravel(), mu[:200], rtol=1e-13)
    print(f"10000-point posterior mean + gradient screen: {dt * 1e3:.2f} ms")
    assert dt < 0.5
    sys.path.insert(0, os.path.join(ROOT, "cornell-moe_b200"))
    import GPP
    g = GPP.GaussianProcess([1.0, list(prob["lengths"])], list(prob["X"].ravel()), list(prob["y"]), [float(prob["noise"][0])],
                            [], 0, 8, 500)
    lst = GPP.compute_posterior_mean_of_points(g, 0, list(pts[:300].ravel()), 300)
    np.testing.assert_allclose(-np.array(lst), mu[:300], rtol=1e-12)
    assert abs(GPP.compute_posterior_mean(g, 0, list(pts[5])) - lst[5]) < 1e-12


@needs_ref
@pytest.mark.parametrize("kernel,g_idx,N,dim", [(0, (), 60, 3), (1, (), 60, 3), (0, (0, 2), 25, 3), (1, (0,), 20, 2),
                                               (0, (), 400, 6), (1, (), 700, 5)])
def test_grad_log_marginal_likelihood_matches_reference(capi, kernel, g_idx, N, dim):
    """§8f rank 2: hyper-parameter gradient of log p(y | X, theta) on the device vs
    LogMarginalLikelihoodEvaluator::ComputeGradLogLikelihood (incl. the Matern routine's value-entry-only quirk)."""
    prob = make_problem(N, dim, g_idx=g_idx, seed=5 + N)
    args = (kernel, 1.3, prob["lengths"] * np.linspace(1.0, 1.5, dim), prob["X"], prob["y"], prob["noise"],
            prob["derivs"])
    got = capi.grad_log_marginal_likelihood(*args)
    want = orc.load_reference().grad_log_marginal_likelihood(*args)
    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-9 * np.abs(want).max())


@needs_ref
def test_multistart_ei_simplex_domain_matches_reference_driver(capi):
    """§8f rank 4: the q-EI multistart driver over SimplexIntersectTensorProductDomain (gpp_domain.cpp:107-289) vs the
    reference's own driver instantiated with that domain."""
    prob = make_problem(25, 3, seed=8, noise=0.05)
    prob["X"] *= 0.33  # training data inside the simplex
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    ref, lm = orc.load_reference().gp(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    assert lm == 0
    rng = np.random.default_rng(19)
    q, mc, seed = 2, 256, 7
    starts = rng.dirichlet(np.ones(4), size=(30, q))[:, :, :3] * 0.9
    best = float(prob["y"].min()) + 0.2
    outer = [30, 10, 2, 0, 0.7, 0.8, 1.0, 1e-7]  # max_relative_change = 1.0 exercises the epsilon tweak
    bounds = np.tile([0.0, 1.0], 3)
    table = orc.normal_draws(seed, mc * q)
    outer1 = [30, 1, 1, 0, 0.7, 0.8, 1.0, 1e-7]
    bp_ref = orc.ref_multistart_ei_simplex(ref, starts, None, mc, best, outer1, bounds, seed)
    bp, bv, found, sv = capi.multistart_ei(gp, starts, None, mc, best, outer1, bounds, seed=1, table=table,
                                           domain_type=capi.SIMPLEX)
    assert np.all(bp >= 0.0) and np.all(bp.sum(axis=1) <= 1.0 + 1e-12)
    np.testing.assert_allclose(bp, bp_ref, rtol=0, atol=1e-9)
    pairs = []
    for i in range(10):
        r_pt = orc.ref_multistart_ei_simplex(ref, starts[i:i + 1], None, mc, best, outer, bounds, seed)
        o_pt = capi.multistart_ei(gp, starts[i:i + 1], None, mc, best, outer, bounds, seed=1, table=table,
                                  domain_type=capi.SIMPLEX)[0]
        assert np.all(o_pt >= 0.0) and np.all(o_pt.sum(axis=1) <= 1.0 + 1e-12)
        pairs.append((o_pt, r_pt))
    assert _agreeing(pairs, 1e-8) >= 7, [float(np.abs(a - b).max()) for a, b in pairs]
