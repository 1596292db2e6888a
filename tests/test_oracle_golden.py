"""CPU: the plain-C oracle against (a) the known-answer cases of the reference's own tests and (b) golden vectors
generated from the unmodified reference (tests/golden/make_golden.py).  Needs neither /root/reference nor a GPU."""
import os

import numpy as np
import pytest

import oracle as orc
from synth import DISCRETE_ONLY_GD, EXAMPLE_INNER_GD, unit_bounds

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files})


@pytest.fixture(scope="module")
def o():
    return orc.load_oracle()


def test_cholesky_known_answers(o):
    # gpp_linear_algebra_test.cpp:238-262 (exact integer factors) and :358-368 (factor + solve -> {1,2,3})
    A = np.array([[81.0, 27, 0, 90], [27, 13, 8, 44], [0, 8, 52, 40], [90, 44, 40, 217]])
    rc, L = o.cholesky(A)
    assert rc == 0
    np.testing.assert_array_equal(np.tril(L), np.array([[9.0, 0, 0, 0], [3, 2, 0, 0], [0, 4, 6, 0], [10, 7, 2, 8]]))
    B = np.array([[25.0, 15, -5], [15, 18, 0], [-5, 0, 11]])
    np.testing.assert_array_equal(np.tril(o.cholesky(B)[1]), np.array([[5.0, 0, 0], [3, 3, 0], [-1, 1, 3]]))
    W = np.array([[4.0, 12, -16], [12, 37, -43], [-16, -43, 98]])
    rc, Lw = o.cholesky(W)
    np.testing.assert_array_equal(np.tril(Lw), np.array([[2.0, 0, 0], [6, 1, 0], [-8, 5, 3]]))
    np.testing.assert_array_equal(o.potrs(np.tril(Lw), np.array([-20.0, -43.0, 192.0])), [1.0, 2.0, 3.0])


@pytest.mark.parametrize("n", [5, 11, 20])
def test_cholesky_random_spd_residual(o, n):
    # gpp_linear_algebra_test.cpp:289-333: ||L L^T - A|| within a few eps per entry
    rng = np.random.default_rng(34187 + n)
    A = rng.standard_normal((n, n))
    A = A @ A.T + n * np.eye(n)
    rc, L = o.cholesky(A)
    L = np.tril(L)
    assert rc == 0
    np.testing.assert_allclose(L @ L.T, A, rtol=0, atol=3 * np.finfo(float).eps * n * np.abs(A).max())


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(o, case):
    g = lambda k: GOLD[f"{case}/{k}"]
    kernel, g_idx = int(g("kernel")), tuple(int(v) for v in g("g_idx"))
    gp, lm = o.gp(kernel, 1.0, g("lengths"), g("X"), g("y"), g("noise"), g_idx)
    assert lm == 0
    K, kinvy, mean = gp.state()
    np.testing.assert_allclose(np.tril(K), g("K_chol_lower"), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(kinvy, g("K_inv_y"), rtol=1e-9, atol=1e-11)
    assert mean == float(g("mean"))
    want = ("mean", "grad_mean", "var", "chol_var", "grad_var", "grad_chol")
    post = gp.posterior(g("pts"), g_idx, want)
    Q = 3 * (1 + len(g_idx))
    for k in want:
        a, b = post[k], g(f"post_{k}")
        if k in ("var", "chol_var"):
            a, b = np.tril(a.reshape(Q, Q).T), np.tril(b.reshape(Q, Q).T)
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-11, err_msg=k)
    ei, gei = gp.ei(g("mc_Xq"), g("mc_Xp"), 32, float(g("mc_best_ei")), g("ei_table"), grad=True)
    np.testing.assert_allclose(ei, g("ei"), rtol=1e-11)
    np.testing.assert_allclose(gei, g("ei_grad"), rtol=1e-8, atol=1e-11)
    dim = g("X").shape[1]
    for tag, gd in (("discrete", DISCRETE_ONLY_GD), ("linesearch", EXAMPLE_INNER_GD)):
        kg, gkg, bp = gp.kg(g("mc_Xq"), g("mc_Xp"), 32, float(g("mc_best_kg")), g("kg_table"), gd, unit_bounds(dim),
                            g("mc_disc"), grad=True, want_best_points=True)
        np.testing.assert_allclose(bp, g(f"kg_{tag}_xstar"), rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(kg, g(f"kg_{tag}"), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(gkg, g(f"kg_{tag}_grad"), rtol=1e-6, atol=1e-9)


def test_philox_host_stream_statistics():
    z = orc.philox_normals(0xC0FFEE, 0, 20000, 8)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    # counter-based: any sub-range equals the same range drawn on its own
    np.testing.assert_array_equal(orc.philox_normals(0xC0FFEE, 100, 5, 8), z[100:105])
    # Philox4x32-10 known-answer test (Random123 kat_vectors: counter = key = 0)
    import ctypes
    lib = ctypes.CDLL(orc.oracle_path())
    assert hasattr(lib, "oracle_philox_normals")


WIDE = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_widened.npz"))


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in WIDE.files if k.startswith("mcmc")}))
def test_oracle_mcmc_matches_golden(o, case):
    g = lambda k: WIDE[f"{case}/{k}"]
    dim, nf = g("X").shape[1], int(g("nf"))
    args = (g("hypers"), g("noises"), g("X"), g("y"), g("g_idx"), g("Xq"), g("Xp"))
    kg, gkg = o.kg_mcmc(*args, 16, g("best"), g("kg_table"), EXAMPLE_INNER_GD, unit_bounds(dim - nf), g("disc"),
                        num_fidelity=nf, grad=True)
    np.testing.assert_allclose(kg, g("kg"), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gkg, g("kg_grad"), rtol=1e-6, atol=1e-9)
    ei, gei = o.ei_mcmc(*args, 16, g("best") + 1.0, g("ei_table"), grad=True)
    np.testing.assert_allclose(ei, g("ei"), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(gei, g("ei_grad"), rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in WIDE.files if k.startswith("pmopt")}))
def test_oracle_posterior_mean_optimization_matches_golden(o, case):
    g = lambda k: WIDE[f"{case}/{k}"]
    nf = int(g("nf"))
    gp, lm = o.gp(int(g("kernel")), 1.0, g("lengths"), g("X"), g("y"), g("noise"), g("g_idx"))
    assert lm == 0
    bp, val = gp.posterior_mean_optimization(g("x0"), g("gd"), unit_bounds(3 - nf), nf)
    np.testing.assert_allclose(bp, g("best_point"), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(val, float(g("best_value")), rtol=1e-12)
