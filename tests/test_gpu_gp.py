"""GPU parity: covariance build, blocked Cholesky, triangular solves, GP fit — CUDA path (through the C ABI) vs the
CPU checker on identical inputs."""
import numpy as np
import pytest

import oracle as orc
from gpu_util import checker, tril_close
from synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from cornell_moe_b200 import capi as c
    assert c.device_count() > 0, "needs a CUDA device"
    return c


def test_cholesky_known_answers(capi):
    # reference KATs: gpp_linear_algebra_test.cpp:238-262 (integer matrices, exact factors) and :358-368
    A = np.array([[81.0, 27, 0, 90], [27, 13, 8, 44], [0, 8, 52, 40], [90, 44, 40, 217]])
    L = capi.cholesky(A)
    np.testing.assert_array_equal(np.tril(L), np.array([[9.0, 0, 0, 0], [3, 2, 0, 0], [0, 4, 6, 0], [10, 7, 2, 8]]))
    B = np.array([[25.0, 15, -5], [15, 18, 0], [-5, 0, 11]])
    np.testing.assert_array_equal(np.tril(capi.cholesky(B)), np.array([[5.0, 0, 0], [3, 3, 0], [-1, 1, 3]]))
    W = np.array([[4.0, 12, -16], [12, 37, -43], [-16, -43, 98]])
    Lw = capi.cholesky(W)
    np.testing.assert_array_equal(np.tril(Lw), np.array([[2.0, 0, 0], [6, 1, 0], [-8, 5, 3]]))
    x = capi.potrs(np.tril(Lw), np.array([-20.0, -43.0, 192.0]))
    np.testing.assert_allclose(x, [1.0, 2.0, 3.0], rtol=1e-14)


@pytest.mark.parametrize("n", [1, 5, 11, 20, 63, 64, 65, 130, 257, 500])
def test_cholesky_random_spd(capi, n):
    rng = np.random.default_rng(34187 + n)
    A = rng.standard_normal((n, n))
    A = A @ A.T + n * np.eye(n)
    L = np.tril(capi.cholesky(A))
    rc, Lref = checker().cholesky(A)
    assert rc == 0
    tril_close(L, Lref, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(L @ L.T, A, rtol=1e-12, atol=1e-11 * n)
    B = rng.standard_normal((n, 37))
    X = capi.potrs(L, B)
    np.testing.assert_allclose(A @ X, B, rtol=1e-9, atol=1e-9)


def test_cholesky_failure_index(capi):
    # exactly singular in integer arithmetic: pivot 3 is 0 -> k+1 = 3, as gpp_linear_algebra.cpp:141-142
    L0 = np.array([[2.0, 0, 0, 0], [1, 3, 0, 0], [4, 1, 0, 0], [2, 2, 1, 5]])
    with pytest.raises(capi.SingularMatrixError) as e:
        capi.cholesky(L0 @ L0.T)
    assert e.value.info == 3
    # failure in a later block of the blocked algorithm
    n = 150
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n, n))
    A = A @ A.T + n * np.eye(n)
    A[100, :] = 0.0
    A[:, 100] = 0.0
    with pytest.raises(capi.SingularMatrixError) as e:
        capi.cholesky(A)
    assert e.value.info == 101


@pytest.mark.parametrize("n", [1024, 1090, 1536, 2306])
def test_cholesky_cooperative_path(capi, n):
    """n >= 1024 (even) takes the one-launch-per-panel path (potrf_coop.cu): flag-chained panel steps, TMA-fed DMMA
    trailing update.  Compared with the reference's ComputeCholeskyFactorL on the same matrix."""
    rng = np.random.default_rng(977 + n)
    G = rng.standard_normal((n, n // 2))
    A = G @ G.T + 0.5 * n * np.eye(n)
    L = np.tril(capi.cholesky(A))
    rc, Lref = checker().cholesky(A)
    assert rc == 0
    tril_close(L, Lref, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(L @ L.T, A, rtol=1e-12, atol=1e-11 * n)
    # one right-hand side on a large factor: the cooperative owner/helper triangular solve (trsv_coop.cu)
    b = rng.standard_normal(n)
    x = capi.potrs(L, b)
    np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    # odd sizes keep the launch-per-step path; both must agree with the reference
    Lo = np.tril(capi.cholesky(A[: n - 1, : n - 1]))
    tril_close(Lo, Lref[: n - 1, : n - 1], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("n", [1536, 2306])
@pytest.mark.parametrize("block", [0, 2, 3])
def test_cholesky_cooperative_gated_rows(capi, n, block, monkeypatch):
    """Update-bound panels let the CTAs of the rows below the diagonal block keep working on trailing tiles until a
    given diagonal block of the panel is factored.  That schedule only engages for n >= 3328 by default;
    CMOE_CHOL_GATE="1,<block>" forces it on every panel with trailing columns, so the small sizes cover it too.  Same
    matrix, both schedules, bit-identical factors expected (the arithmetic of a row block does not depend on when it
    runs)."""
    rng = np.random.default_rng(1234 + n)
    G = rng.standard_normal((n, n // 2))
    A = G @ G.T + 0.5 * n * np.eye(n)
    monkeypatch.setenv("CMOE_CHOL_GATE", "100000,0")
    L0 = np.tril(capi.cholesky(A))
    monkeypatch.setenv("CMOE_CHOL_GATE", "1,%d" % block)
    L1 = np.tril(capi.cholesky(A))
    np.testing.assert_array_equal(L0, L1)
    rc, Lref = checker().cholesky(A)
    assert rc == 0
    tril_close(L1, Lref, rtol=1e-10, atol=1e-11)
    # a failed pivot inside the chain while row blocks are still held back
    bad = 700
    A[bad, :] = 0.0
    A[:, bad] = 0.0
    with pytest.raises(capi.SingularMatrixError) as e:
        capi.cholesky(A)
    assert e.value.info == bad + 1


@pytest.mark.parametrize("bad", [0, 70, 300, 1279, 1535])
def test_cholesky_cooperative_failure_index(capi, bad):
    n = 1536
    rng = np.random.default_rng(3)
    G = rng.standard_normal((n, 64))
    A = G @ G.T + n * np.eye(n)
    A[bad, :] = 0.0
    A[:, bad] = 0.0
    with pytest.raises(capi.SingularMatrixError) as e:
        capi.cholesky(A)
    assert e.value.info == bad + 1


def test_gp_fit_n2000_matches_reference_factor(capi):
    """Large fit against the reference's own factor (not only residual properties): N = 2000, d = 10."""
    prob = make_problem(2000, 10, seed=77, length=0.5)
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    ref, lm = checker().gp(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    assert lm == 0
    K, kinvy, mean = gp.state()
    Kr, kr, mr = ref.state()
    assert mean == mr
    tril_close(K, Kr, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(kinvy, kr, rtol=1e-6, atol=1e-6)


def test_philox_stream_matches_host(capi):
    dev = capi.philox_normals(0xC0FFEE, 5, 300, 7)
    host = orc.philox_normals(0xC0FFEE, 5, 300, 7)
    np.testing.assert_allclose(dev, host, rtol=1e-12, atol=1e-14)
    assert abs(dev.mean()) < 0.1 and abs(dev.std() - 1.0) < 0.1


@pytest.mark.parametrize("kernel", [0, 1])
def test_gp_fit_offset_domain(capi, kernel):
    """Training points far from the origin relative to the length scale ([1000, 1002]^4, l = 0.5: |x| ~ 4000 length
    scales).  The fast covariance build forms -r^2/2 as x.y - |x|^2/2 - |y|^2/2; on uncentred coordinates every entry
    would carry ~eps (|x|^2 + |y|^2) ~ 4e-9 of absolute error in the exponent.  The coordinates are centred before
    scaling (the kernels are translation invariant), so the factor matches the reference — which differences first —
    to the same tolerance as a unit-cube problem."""
    prob = make_problem(200, 4, seed=91)
    X0 = prob["X"].copy()
    prob["X"] = 1000.0 + 2.0 * X0
    prob["lengths"] = np.full(4, 0.5)
    gp = capi.GaussianProcess(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"],
                              prob["derivs"])
    ref, lm = checker().gp(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    assert lm == 0
    K, kinvy, mean = gp.state()
    Kr, kr, mr = ref.state()
    tril_close(K, Kr, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(kinvy, kr, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("N,dim,g_idx", [(50, 2, ()), (24, 3, (0, 2)), (200, 6, ()), (300, 4, (0, 1, 2, 3)), (500, 8, ())])
def test_gp_fit_matches_checker(capi, kernel, N, dim, g_idx):
    prob = make_problem(N, dim, g_idx=g_idx, seed=N + dim)
    gp = capi.GaussianProcess(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"],
                              prob["derivs"])
    ref, lm = checker().gp(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    assert lm == 0
    K, kinvy, mean = gp.state()
    Kr, kr, mr = ref.state()
    assert mean == mr
    tril_close(K, Kr, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(kinvy, kr, rtol=1e-7, atol=1e-8)


def test_gp_singular_reports_leading_minor(capi):
    prob = make_problem(10, 2, seed=1)
    prob["X"][1] = prob["X"][0]
    prob["noise"][:] = 0.0
    with pytest.raises(capi.SingularMatrixError) as e:
        capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    assert e.value.info == 2


@pytest.mark.parametrize("kernel", [0, 1])
def test_gp_singular_on_tma_cov_path(capi, kernel):
    """N >= 256 (even) builds K with the TMA / DMMA kernel: coincident points must still give k = alpha bit-for-bit, so
    a duplicate with zero noise fails the factorisation at the reference's leading minor."""
    prob = make_problem(300, 5, seed=21)
    prob["X"][217] = prob["X"][40]
    prob["noise"][:] = 0.0
    _, lm = checker().gp(kernel, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    assert lm == 218
    capi.set_option("cov_tma", 1)
    try:
        with pytest.raises(capi.SingularMatrixError) as e:
            capi.GaussianProcess(kernel, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    finally:
        capi.set_option("cov_tma", 0)
    assert e.value.info == lm


def test_gp_cov_build_tma_matches_checker(capi):
    # K itself (not only its factor): a huge noise term keeps the factor's first column ~ K[:, 0] / sqrt(K00)
    for N, dim, tma in [(256, 3, 1), (386, 10, 1), (1000, 7, 1), (1000, 7, 0)]:
        prob = make_problem(N, dim, seed=N)
        capi.set_option("cov_tma", tma)
        try:
            gp = capi.GaussianProcess(0, 1.7, prob["lengths"], prob["X"], prob["y"], prob["noise"])
        finally:
            capi.set_option("cov_tma", 0)
        ref, lm = checker().gp(0, 1.7, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
        assert lm == 0
        tril_close(gp.state()[0], ref.state()[0], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("legacy,tma", [(1, 0), (0, 1)])
def test_gp_large_fit_other_kernel_generations(capi, legacy, tma):
    """The round-1 launch-per-step factorisation / chained solve and the TMA covariance build stay selectable at run
    time (cmoe_set_option); both must give the same fit as the defaults to rounding."""
    prob = make_problem(1500, 6, seed=15)
    ref_fit = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"]).state()
    capi.set_option("legacy_linalg", legacy)
    capi.set_option("cov_tma", tma)
    try:
        other = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"]).state()
    finally:
        capi.set_option("legacy_linalg", 0)
        capi.set_option("cov_tma", 0)
    tril_close(other[0], ref_fit[0], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(other[1], ref_fit[1], rtol=1e-6, atol=1e-7)


def test_gp_large_fit_residual(capi):
    # size-independent property at a config-5-like size: L L^T reproduces K and K (K^-1 y) reproduces y - mean
    prob = make_problem(2000, 10, seed=5, length=0.5)
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    L, kinvy, mean = gp.state()
    L = np.tril(L)
    X = prob["X"] / prob["lengths"]
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    K = np.exp(-0.5 * d2) + prob["noise"][0] * np.eye(2000)
    np.testing.assert_allclose(L @ L.T, K, rtol=0, atol=1e-11)
    np.testing.assert_allclose(K @ kinvy, prob["y"] - mean, rtol=0, atol=1e-8)


def test_gp_config5_fit_properties(capi):
    """BASELINE.json configs[4] at full size (N = 5000, d = 10): look-ahead blocked Cholesky + chained triangular solves.
    Size-independent properties on random samples: (L L^T)_ij = K_ij and (K K^-1 (y - m))_i = (y - m)_i."""
    N, d = 5000, 10
    prob = make_problem(N, d, seed=55, length=0.5)
    gp = capi.GaussianProcess(0, 1.3, prob["lengths"], prob["X"], prob["y"], prob["noise"])
    L, kinvy, mean = gp.state()
    L = np.tril(L)
    rng = np.random.default_rng(1)
    idx = rng.integers(0, N, size=(3000, 2))
    i, j = np.maximum(idx[:, 0], idx[:, 1]), np.minimum(idx[:, 0], idx[:, 1])
    Xs = prob["X"] / prob["lengths"]
    want = 1.3 * np.exp(-0.5 * ((Xs[i] - Xs[j]) ** 2).sum(axis=1)) + (i == j) * prob["noise"][0]
    got = np.einsum("ij,ij->i", L[i], L[j])
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-12)
    rows = rng.integers(0, N, size=40)
    Krows = 1.3 * np.exp(-0.5 * ((Xs[rows][:, None, :] - Xs[None, :, :]) ** 2).sum(-1))
    Krows[np.arange(40), rows] += prob["noise"][0]
    np.testing.assert_allclose(Krows @ kinvy, (prob["y"] - mean)[rows], rtol=0, atol=2e-8)
    np.testing.assert_allclose(mean, prob["y"].mean(), rtol=1e-12)
