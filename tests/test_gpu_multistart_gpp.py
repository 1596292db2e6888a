"""GPU tests of the multistart drivers and of the pybind11 `GPP` module (the reference's moe.build.GPP surface)."""
import os
import sys

import numpy as np
import pytest

import oracle as orc
from gpu_util import checker
from synth import EXAMPLE_INNER_GD, make_problem, unit_bounds

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from cornell_moe_b200 import capi as c
    assert c.device_count() > 0
    return c


def _ref_restarted_gd(backend, evaluate, x0, gd, bounds, dim):
    """The reference's GradientDescentOptimizer (gpp_optimization.hpp:620-705, 1144-1185), restated on top of the
    checker's value/gradient evaluation — an independent trajectory to compare the device-driven one with."""
    _, max_steps, max_restarts, _, gamma, pre_mult, mrc, tol = gd
    x = np.array(x0, dtype=np.float64)
    for _ in range(int(max_restarts)):
        start = x.copy()
        for i in range(int(max_steps)):
            _, g = evaluate(x, True)
            alpha = pre_mult * (i + 1.0) ** (-gamma)
            step = alpha * g
            for k in range(x.shape[0]):
                step[k] = backend.limit_update(bounds, mrc, x[k], step[k])
            x = x + step
            if np.linalg.norm(step) < tol / max_steps:
                break
        if np.linalg.norm(start - x) <= tol:
            break
    return evaluate(x, False), x


def test_kg_gradient_descent_matches_restated_reference(capi):
    prob = make_problem(20, 3, seed=13, noise=0.05)
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    backend = checker()
    ref, _ = backend.gp(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    rng = np.random.default_rng(2)
    q, mc, seed = 2, 32, 77
    starts = rng.uniform(0.2, 0.8, size=(3, q, 3))
    disc = rng.uniform(size=(6, 3))
    best = float(ref.mean_additional(disc).min())
    outer = [1, 4, 2, 0, 0.7, 0.5, 0.2, 1e-7]
    table = orc.philox_normals(seed, 0, mc // 2, q)

    def evaluate(x, want_grad):
        if want_grad:
            return ref.kg(x, None, mc, best, table, EXAMPLE_INNER_GD, unit_bounds(3), disc, grad=True)
        return ref.kg(x, None, mc, best, table, EXAMPLE_INNER_GD, unit_bounds(3), disc)

    vals, pts = capi.kg_gradient_descent(gp, starts, None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(3),
                                         unit_bounds(3), disc, seed=seed)
    for s in range(3):
        v, x = _ref_restarted_gd(backend, evaluate, starts[s], outer, unit_bounds(3), 3)
        np.testing.assert_allclose(pts[s], x, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(vals[s], v, rtol=1e-5, atol=1e-8)


def test_multistart_kg_selects_strict_argmax_of_top20(capi):
    prob = make_problem(30, 2, seed=3, noise=0.05)
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    rng = np.random.default_rng(4)
    q, mc, seed = 2, 64, 5
    starts = rng.uniform(size=(40, q, 2))
    disc = rng.uniform(size=(8, 2))
    best = float(gp.posterior(disc[:, None, :], (), ("mean",))["mean"].min())
    outer = [40, 3, 1, 0, 0.7, 0.3, 0.2, 1e-7]
    bp, bv, found, sv = capi.multistart_kg(gp, starts, None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(2),
                                           unit_bounds(2), disc, seed=seed)
    assert found and np.isfinite(bv)
    # the driver evaluates through the reference's reused state: the discretisation set keeps the FIRST start's points
    plan = capi.KGPlan(gp, mc, best, EXAMPLE_INNER_GD, unit_bounds(2), disc, len(starts), q, seed=seed, want_grad=False)
    plan.set_stale_union(starts[0])
    plan.upload(starts)
    plan.run()
    plan.sync()
    np.testing.assert_array_equal(sv, plan.download()[0])
    fresh = capi.multistart_kg(gp, starts, None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(2), unit_bounds(2), disc,
                               seed=seed, fresh_discretisation=True)[3]
    np.testing.assert_array_equal(fresh, gp.kg(starts, None, mc, best, EXAMPLE_INNER_GD, unit_bounds(2), disc, seed=seed))
    from cornell_moe_b200 import multigpu
    top = multigpu.top_k_indices(sv)
    vals, pts = capi.kg_gradient_descent(gp, starts[top], None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(2),
                                         unit_bounds(2), disc, seed=seed, stale_union=starts[0])
    k = int(np.argmax(vals))  # numpy argmax = first maximiser = the strict-< update in slot order
    assert bv == vals[k]
    np.testing.assert_array_equal(bp, pts[k])
    assert np.all(bp >= 0.0) and np.all(bp <= 1.0)
    # the sharded driver run in a single process gives the same answer
    bp2, bv2, found2, sv2 = multigpu.multistart_kg(gp, starts, None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(2),
                                                   unit_bounds(2), disc, seed=seed)
    assert bv2 == bv and found2
    np.testing.assert_array_equal(bp2, bp)


def test_multistart_ei_mc_and_analytic(capi):
    prob = make_problem(30, 2, seed=8, noise=0.02)
    gp = capi.GaussianProcess(1, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    rng = np.random.default_rng(1)
    best = float(prob["y"].min())
    outer = [30, 5, 2, 0, 0.7, 0.2, 0.2, 1e-8]
    for q in (1, 3):
        starts = rng.uniform(size=(30, q, 2))
        bp, bv, found, sv = capi.multistart_ei(gp, starts, None, 512, best, outer, unit_bounds(2), seed=9)
        assert found and bv >= sv.max() - 1e-12 and bv >= 0.0
        assert np.all(bp >= 0.0) and np.all(bp <= 1.0)
    # analytic 1-EI (gpp_math.cpp:2196-2253) against the closed form evaluated from the checker's posterior
    ref, _ = checker().gp(1, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    starts = rng.uniform(size=(5, 1, 2))
    null = [5, 0, 0, 0, 0.7, 0.2, 0.2, 1e-8]  # max_num_restarts = 0: objective at the starts only
    vals, _ = capi.ei_gradient_descent(gp, starts, None, 64, best, null, unit_bounds(2), seed=1)
    from math import erfc, exp, pi, sqrt
    for s in range(5):
        post = ref.posterior(starts[s], (), ("mean", "var"))
        mu, sig = post["mean"][0], sqrt(post["var"][0])
        z = (best - mu) / sig
        ei = (best - mu) * 0.5 * erfc(-z / sqrt(2)) + sig * exp(-0.5 * z * z) / sqrt(2 * pi)
        np.testing.assert_allclose(vals[s], max(0.0, ei), rtol=1e-9, atol=1e-13)


def test_gpp_module_mirrors_reference_surface(capi):
    sys.path.insert(0, os.path.join(ROOT, "cornell-moe_b200"))
    import GPP as C_GP
    prob = make_problem(25, 3, seed=31, noise=0.05)
    # exactly what py/cpp_wrappers/gaussian_process.py:56-86 passes: flat python lists
    gp = C_GP.GaussianProcess([1.0, list(prob["lengths"])], list(prob["X"].ravel()), list(prob["y"]),
                              list(prob["noise"]), [], 0, 3, 25)
    assert gp.dim == 3 and gp.num_sampled == 25
    ref, _ = checker().gp(1, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])  # Python boundary = Matern-5/2
    pts = np.random.default_rng(0).uniform(size=(4, 3))
    r = ref.posterior(pts, (), ("mean", "var", "chol_var", "grad_mean"))
    np.testing.assert_allclose(gp.compute_mean_of_points(list(pts.ravel()), 4), r["mean"], rtol=1e-9)
    V = np.array(gp.compute_variance_of_points(list(pts.ravel()), 4)).reshape(4, 4)
    np.testing.assert_allclose(np.tril(V), np.tril(r["var"].reshape(4, 4).T), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(V, V.T, rtol=1e-9)
    np.testing.assert_allclose(gp.compute_grad_mean_of_points(list(pts.ravel()), 4), r["grad_mean"], rtol=1e-8, atol=1e-10)
    rnd = C_GP.RandomnessSourceContainer(4)
    rnd.SetExplicitNormalRNGSeed(123)
    q, mc = 2, 64
    Xq = pts[:q]
    disc = np.random.default_rng(1).uniform(size=(5, 3))
    best = float(ref.mean_additional(disc).min())

    class Params:  # stands in for py/cpp_wrappers/optimization._CppOptimizerParameters
        domain_type = C_GP.DomainTypes.tensor_product
        optimizer_type = C_GP.OptimizerTypes.gradient_descent
        num_random_samples = 16
        optimizer_parameters = C_GP.GradientDescentParameters(*EXAMPLE_INNER_GD)

    kg = C_GP.compute_knowledge_gradient(gp, 0, Params, list(unit_bounds(3)), list(disc.ravel()), list(Xq.ravel()), [], 5,
                                         q, 0, mc, best, rnd)
    gkg = C_GP.compute_grad_knowledge_gradient(gp, 0, Params, list(unit_bounds(3)), list(disc.ravel()), list(Xq.ravel()),
                                               [], 5, q, 0, mc, best, rnd)
    table = orc.philox_normals(123, 0, mc // 2, q)
    v, g = ref.kg(Xq, None, mc, best, table, EXAMPLE_INNER_GD, unit_bounds(3), disc, grad=True)
    np.testing.assert_allclose(kg, v, rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(np.array(gkg).reshape(q, 3), g, rtol=1e-5, atol=1e-8)
    ei = C_GP.compute_expected_improvement(gp, list(Xq.ravel()), [], q, 0, mc, float(prob["y"].min()) + 0.2, False, rnd)
    t2 = orc.philox_normals(123, 0, mc, q)
    np.testing.assert_allclose(ei, ref.ei(Xq, None, mc, float(prob["y"].min()) + 0.2, t2), rtol=1e-9, atol=1e-12)
    status = {}
    outer = type("P", (), dict(domain_type=C_GP.DomainTypes.tensor_product,
                               optimizer_type=C_GP.OptimizerTypes.gradient_descent, num_random_samples=16,
                               optimizer_parameters=C_GP.GradientDescentParameters(24, 2, 1, 0, 0.7, 0.3, 0.2, 1e-7)))
    best_pts = C_GP.multistart_knowledge_gradient_optimization(outer, Params, gp, 0, list(unit_bounds(3)),
                                                               list(disc.ravel()), [], 5, q, 0, best, mc, 4, rnd, status)
    assert status["gradient_descent_tensor_product_domain_found_update"] is True
    assert len(best_pts) == q * 3 and all(0.0 <= x <= 1.0 for x in best_pts)
    with pytest.raises(C_GP.BoundsException):
        C_GP.multistart_knowledge_gradient_optimization(outer, Params, gp, 0, list(unit_bounds(3)), list(disc.ravel()), [],
                                                        5, q, 0, best, mc, 5, rnd, status)  # more threads than RNGs
    with pytest.raises(C_GP.SingularMatrixException):
        X2 = prob["X"].copy()
        X2[1] = X2[0]
        C_GP.GaussianProcess([1.0, list(prob["lengths"])], list(X2.ravel()), list(prob["y"]), [0.0], [], 0, 3, 25)
    vals = C_GP.evaluate_KG_at_point_list(gp, 0, Params, list(unit_bounds(3)), list(disc.ravel()), list(pts.ravel()), 2, 5, q,
                                          0, best, mc, 4, rnd, status)
    assert len(vals) == 2 and status["evaluate_KG_at_point_list"] is True
    np.testing.assert_allclose(vals[0], kg, rtol=0, atol=0)
