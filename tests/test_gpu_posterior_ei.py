"""GPU parity: batched posterior (mean / variance / Cholesky and their gradients) and the q-EI Monte-Carlo estimator,
CUDA path through the C ABI vs the CPU checker on identical inputs (table-fed normals for the MC part)."""
import numpy as np
import pytest

import oracle as orc
from gpu_util import checker
from synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from cornell_moe_b200 import capi as c
    assert c.device_count() > 0
    return c


def _pair(capi, kernel, prob):
    gp = capi.GaussianProcess(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"],
                              prob["derivs"])
    ref, lm = checker().gp(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    assert lm == 0
    return gp, ref


# tolerances follow the reference's own Python-vs-C++ bars (tests/cpp_wrappers/gaussian_process_test.py:93-163):
# mean 3e-13, grad mean 3e-12, var 3e-13, grad var 3e-12, chol 3e-12, grad chol 3e-10 (relative to the scale of the
# quantity); the solves here are blocked, so a factor ~cond(K)*eps is allowed on top.
@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("N,dim,g_idx,num", [(50, 2, (), 1), (7, 3, (0, 1, 2), 5), (60, 3, (), 5), (40, 4, (1, 3), 3),
                                             (200, 6, (), 4)])
def test_posterior_matches_checker(capi, kernel, N, dim, g_idx, num):
    prob = make_problem(N, dim, g_idx=g_idx, seed=3 * N + dim, noise=1e-2)
    gp, ref = _pair(capi, kernel, prob)
    rng = np.random.default_rng(17)
    sets = rng.uniform(size=(3, num, dim))
    want = ("mean", "grad_mean", "var", "chol_var", "grad_var", "grad_chol")
    for ds in ((), g_idx):
        out = gp.posterior(sets, ds, want)
        Q = num * (1 + len(ds))
        for s in range(3):
            r = ref.posterior(sets[s], ds, want)
            assert r["rc"] == 0
            np.testing.assert_allclose(out["mean"][s], r["mean"], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(out["grad_mean"][s], r["grad_mean"], rtol=1e-8, atol=1e-9)
            V = out["var"][s].reshape(Q, Q).T
            Vr = r["var"].reshape(Q, Q).T
            np.testing.assert_allclose(np.tril(V), np.tril(Vr), rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(V, V.T, rtol=1e-9, atol=1e-11)
            L = out["chol_var"][s].reshape(Q, Q).T
            np.testing.assert_allclose(np.tril(L), np.tril(r["chol_var"].reshape(Q, Q).T), rtol=1e-7, atol=1e-9)
            assert np.all(np.triu(L, 1) == 0.0)
            np.testing.assert_allclose(out["grad_var"][s], r["grad_var"], rtol=1e-7, atol=1e-8)
            np.testing.assert_allclose(out["grad_chol"][s], r["grad_chol"], rtol=1e-6, atol=1e-7)


def test_posterior_singular_variance(capi):
    prob = make_problem(20, 2, seed=2)
    gp, _ = _pair(capi, 0, prob)
    pts = np.array([[0.3, 0.4], [0.3, 0.4], [0.7, 0.1]])  # duplicate point -> second... third pivot fails
    with pytest.raises(capi.SingularMatrixError) as e:
        gp.posterior(pts[None], (), ("chol_var",))
    assert e.value.info == 2


@pytest.mark.parametrize("q,p", [(1, 0), (1, 5), (3, 2), (10, 0)])  # the reference's ping-test shapes, gpp_math_test.cpp:1690-1700
def test_ei_table_fed(capi, q, p):
    prob = make_problem(30, 3, seed=4)
    gp, ref = _pair(capi, 0, prob)
    rng = np.random.default_rng(3141)
    cands = rng.uniform(size=(4, q, 3))
    Xp = rng.uniform(size=(p, 3))
    mc = 64
    table = rng.standard_normal(mc * (q + p))
    best = float(prob["y"].min()) + 0.3
    ei, grad = gp.ei(cands, Xp, mc, best, table=table, grad=True)
    for c in range(4):
        v, g = ref.ei(cands[c], Xp, mc, best, table, grad=True)
        np.testing.assert_allclose(ei[c], v, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(grad[c], g, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(gp.ei(cands, Xp, mc, best, table=table), ei, rtol=0, atol=0)


def test_ei_with_derivative_observations_in_gp(capi):
    prob = make_problem(25, 3, g_idx=(0, 2), seed=6)
    gp, ref = _pair(capi, 1, prob)
    rng = np.random.default_rng(5)
    cands = rng.uniform(size=(3, 2, 3))
    table = rng.standard_normal(128 * 2)
    best = float(prob["y"][::3].min()) + 0.2
    ei, grad = gp.ei(cands, None, 128, best, table=table, grad=True)
    for c in range(3):
        v, g = ref.ei(cands[c], None, 128, best, table, grad=True)
        np.testing.assert_allclose(ei[c], v, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(grad[c], g, rtol=1e-6, atol=1e-9)


def test_ei_philox_equals_host_stream_and_3sigma(capi):
    """Native Philox draws == the host restatement fed through the checker (tight), and a different seed agrees within
    3 standard errors of the MC estimator (the north-star tolerance)."""
    prob = make_problem(200, 6, seed=12)
    gp, ref = _pair(capi, 0, prob)
    rng = np.random.default_rng(7)
    q, mc = 4, 10000
    cands = rng.uniform(size=(6, q, 6))
    best = float(prob["y"].min()) + 0.5
    ei, grad = gp.ei(cands, None, mc, best, seed=0xC0FFEE, grad=True)
    table = orc.philox_normals(0xC0FFEE, 0, mc, q)
    for c in range(2):
        v, g = ref.ei(cands[c], None, mc, best, table, grad=True)
        np.testing.assert_allclose(ei[c], v, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(grad[c], g, rtol=1e-5, atol=1e-8)
    ei2 = gp.ei(cands, None, mc, best, seed=12345)
    # per-sample improvement has std <= ~ (ei + a few sigma); bound the standard error generously from the values
    se = np.sqrt(2.0) * (np.abs(ei) + 1.0) / np.sqrt(mc)
    assert np.all(np.abs(ei - ei2) < 3.0 * se + 1e-12)
    assert np.any(ei > 0)
