"""GPU: the sharded (multi-GPU) drivers of the MCMC-averaged acquisition functions, run in a single process, against the
C-ABI drivers: "screen every start, keep the top 20 in the reference's priority-queue order, restarted gradient descent,
strict-> arg-max in slot order".  (The world_size-2 host logic itself is covered on CPU by tests/test_multigpu_gloo.py.)"""
import numpy as np
import pytest

from synth import EXAMPLE_INNER_GD, make_problem, unit_bounds

pytestmark = pytest.mark.gpu


def test_sharded_mcmc_drivers_match_c_drivers():
    from cornell_moe_b200 import capi, multigpu
    assert capi.device_count() > 0
    M, dim, q, mc = 3, 2, 2, 32
    prob = make_problem(25, dim, seed=12, noise=0.05)
    rng = np.random.default_rng(3)
    hypers = np.concatenate([rng.uniform(0.8, 1.5, size=(M, 1)), rng.uniform(0.4, 0.9, size=(M, dim))], axis=1)
    noises = rng.uniform(0.05, 0.15, size=(M, 1))
    ens = capi.GaussianProcessEnsemble(hypers, noises, prob["X"], prob["y"])
    rng = np.random.default_rng(14)
    starts = rng.uniform(size=(30, q, dim))
    disc = rng.uniform(size=(M, 6, dim))
    best = np.array([float(m.posterior(disc[i][:, None, :], (), ("mean",))["mean"].min())
                     for i, m in enumerate(ens.members)])
    outer = [30, 3, 1, 0, 0.7, 0.3, 0.2, 1e-7]
    args = (starts, None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(dim), unit_bounds(dim), disc)
    bp, bv, found, sv = ens.multistart_kg(*args, seed=5)
    bp_s, bv_s, found_s, sv_s = multigpu.multistart_kg_mcmc(ens, *args, seed=5)
    np.testing.assert_allclose(sv_s, sv, rtol=1e-13, atol=1e-15)
    assert found and found_s
    np.testing.assert_allclose(bv_s, bv, rtol=1e-12)
    np.testing.assert_allclose(bp_s, bp, rtol=1e-12, atol=1e-14)
    top = multigpu.top_k_indices(sv)
    gv, gpts = ens.kg_gradient_descent(starts[top], None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(dim),
                                       unit_bounds(dim), disc, seed=5)
    k = int(np.argmax(gv))  # numpy argmax = first maximiser = the strict-< update in slot order
    np.testing.assert_allclose(bv, gv[k], rtol=1e-12)
    np.testing.assert_allclose(bp, gpts[k], rtol=1e-12, atol=1e-14)
    ybest = np.full(M, float(prob["y"].min()))
    gd = [25, 4, 2, 0, 0.7, 0.2, 0.2, 1e-8]
    for qq in (1, 2):
        st = rng.uniform(size=(25, qq, dim))
        bp, bv, found, sv = ens.multistart_ei(st, None, 256, ybest, gd, unit_bounds(dim), seed=3)
        bp2, bv2, found2, sv2 = multigpu.multistart_ei_mcmc(ens, st, None, 256, ybest, gd, unit_bounds(dim), seed=3)
        np.testing.assert_allclose(sv2, sv, rtol=1e-13, atol=1e-16)
        assert found == found2
        np.testing.assert_allclose(bv2, bv, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(bp2, bp, rtol=1e-12, atol=1e-14)


def test_kg_se_fidelity_dimension_batched_line_search():
    """SquareExponential kernel with a fidelity coordinate (gradient pinned to 0 there) through the batched backtracking
    path — the Matern twin of this case is tests/test_gpu_kg.py::test_kg_fidelity_dimension."""
    from cornell_moe_b200 import capi
    from gpu_util import checker
    prob = make_problem(14, 3, seed=21, noise=0.1)
    gp = capi.GaussianProcess(0, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    ref, lm = checker().gp(0, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    assert lm == 0
    rng = np.random.default_rng(8)
    cands = rng.uniform(size=(2, 2, 3))
    disc = rng.uniform(size=(6, 2))
    table = rng.standard_normal(8 * 2)
    kg, grad = gp.kg(cands, None, 16, 0.1, EXAMPLE_INNER_GD, unit_bounds(2), disc, num_fidelity=1, table=table, grad=True)
    for c in range(2):
        v, g = ref.kg(cands[c], None, 16, 0.1, table, EXAMPLE_INNER_GD, unit_bounds(2), disc, num_fidelity=1, grad=True)
        np.testing.assert_allclose(kg[c], v, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(grad[c], g, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("kernel,g_idx,N,dim", [(1, (), 20, 3), (0, (), 30, 2), (1, (0, 2), 12, 3), (0, (), 300, 4)])
def test_log_marginal_likelihood_matches_checker(kernel, g_idx, N, dim):
    """cmoe_log_marginal_likelihood (value only) = device fit + host reductions, against
    LogMarginalLikelihoodEvaluator::ComputeLogLikelihood of the compiled reference (or its pinned restatement)."""
    from cornell_moe_b200 import capi
    from gpu_util import checker
    prob = make_problem(N, dim, g_idx=g_idx, seed=5 + N)
    args = (kernel, 1.3, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    got = capi.log_marginal_likelihood(*args)
    want = checker().log_marginal_likelihood(*args)
    # 1e-8, not 1e-9: -1/2 y^T K^-1 y and -sum log L_ii nearly cancel at N = 300 (the value is ~1e-2 of either term), so
    # the 1e-11-level differences of two correct factorisations show up two digits higher in the sum
    np.testing.assert_allclose(got, want, rtol=1e-8)
