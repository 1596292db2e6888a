"""GPU: the CUDA path (through the C ABI) against the committed golden vectors generated from the unmodified reference
(tests/golden/make_golden.py) — this parity check needs no CPU checker at run time."""
import os

import numpy as np
import pytest

from synth import DISCRETE_ONLY_GD, EXAMPLE_INNER_GD, unit_bounds

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files})


@pytest.mark.parametrize("case", CASES)
def test_cuda_matches_golden(case):
    from cornell_moe_b200 import capi
    g = lambda k: GOLD[f"{case}/{k}"]
    kernel, g_idx = int(g("kernel")), tuple(int(v) for v in g("g_idx"))
    gp = capi.GaussianProcess(kernel, 1.0, g("lengths"), g("X"), g("y"), g("noise"), g_idx)
    K, kinvy, mean = gp.state()
    np.testing.assert_allclose(np.tril(K), g("K_chol_lower"), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(kinvy, g("K_inv_y"), rtol=1e-8, atol=1e-10)
    assert mean == float(g("mean"))
    want = ("mean", "grad_mean", "var", "chol_var", "grad_var", "grad_chol")
    post = gp.posterior(g("pts"), g_idx, want)
    Q = 3 * (1 + len(g_idx))
    for k in want:
        a, b = post[k][0], g(f"post_{k}")
        if k in ("var", "chol_var"):
            a, b = np.tril(a.reshape(Q, Q).T), np.tril(b.reshape(Q, Q).T)
        np.testing.assert_allclose(a, b, rtol=1e-7, atol=1e-9, err_msg=k)
    ei, gei = gp.ei(g("mc_Xq"), g("mc_Xp"), 32, float(g("mc_best_ei")), table=g("ei_table"), grad=True)
    np.testing.assert_allclose(ei[0], g("ei"), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gei[0], g("ei_grad"), rtol=1e-6, atol=1e-9)
    dim = g("X").shape[1]  # q-KG and d-KG (derivative observations) alike
    for tag, gd in (("discrete", DISCRETE_ONLY_GD), ("linesearch", EXAMPLE_INNER_GD)):
        kg, gkg = gp.kg(g("mc_Xq"), g("mc_Xp"), 32, float(g("mc_best_kg")), gd, unit_bounds(dim), g("mc_disc"),
                        table=g("kg_table"), grad=True)
        np.testing.assert_allclose(kg[0], g(f"kg_{tag}"), rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(gkg[0], g(f"kg_{tag}_grad"), rtol=1e-5, atol=1e-8)


WIDE = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_widened.npz"))


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in WIDE.files if k.startswith("mcmc")}))
def test_cuda_mcmc_matches_golden(case):
    from cornell_moe_b200 import capi
    g = lambda k: WIDE[f"{case}/{k}"]
    dim, nf = g("X").shape[1], int(g("nf"))
    ens = capi.GaussianProcessEnsemble(g("hypers"), g("noises"), g("X"), g("y"), g("g_idx"))
    kg, gkg = ens.kg(g("Xq"), g("Xp"), 16, g("best"), EXAMPLE_INNER_GD, unit_bounds(dim - nf), g("disc"),
                     num_fidelity=nf, table=g("kg_table"), grad=True)
    np.testing.assert_allclose(kg[0], g("kg"), rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(gkg[0], g("kg_grad"), rtol=1e-5, atol=1e-8)
    ei, gei = ens.ei(g("Xq"), g("Xp"), 16, g("best") + 1.0, table=g("ei_table"), grad=True)
    np.testing.assert_allclose(ei[0], g("ei"), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gei[0], g("ei_grad"), rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in WIDE.files if k.startswith("pmopt")}))
def test_cuda_posterior_mean_optimization_matches_golden(case):
    from cornell_moe_b200 import capi
    g = lambda k: WIDE[f"{case}/{k}"]
    nf = int(g("nf"))
    gp = capi.GaussianProcess(int(g("kernel")), 1.0, g("lengths"), g("X"), g("y"), g("noise"), g("g_idx"))
    bp, val, found = capi.posterior_mean_optimization(gp, g("x0"), g("gd"), unit_bounds(3 - nf), nf)
    assert found
    np.testing.assert_allclose(bp, g("best_point"), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(val, float(g("best_value")), rtol=1e-9)
