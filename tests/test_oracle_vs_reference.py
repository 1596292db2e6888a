"""Pins the plain-C oracle (oracle/moe_oracle.c) against the UNMODIFIED reference compiled into
oracle/_ref/libmoe_ref.so.  Skipped where the reference .so is absent (it is built from /root/reference by
`make -C oracle ref`, in the build container, and travels with the snapshot)."""
import numpy as np
import pytest

import oracle as orc
from synth import DISCRETE_ONLY_GD, EXAMPLE_INNER_GD, make_problem, unit_bounds

pytestmark = pytest.mark.skipif(not orc.have_reference(), reason="oracle/_ref/libmoe_ref.so not built")


@pytest.fixture(scope="module")
def libs():
    return orc.load_oracle(), orc.load_reference()


def _gp_pair(libs, kernel, prob):
    out = []
    for b in libs:
        gp, lm = b.gp(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
        assert lm == 0
        out.append(gp)
    return out


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("g1,g2", [((), ()), ((0, 2), ()), ((), (1,)), ((0, 1, 2), (0, 2)), ((1,), (1,))])
def test_covariance_blocks(libs, kernel, g1, g2):
    rng = np.random.default_rng(5)
    o, r = libs
    for _ in range(5):
        p1, p2 = rng.uniform(size=3), rng.uniform(size=3)
        ls = rng.uniform(0.5, 2.5, size=3)
        for grad in (False, True):
            a = o.covariance(kernel, 2.80723, ls, p1, g1, p2, g2, grad=grad)
            b = r.covariance(kernel, 2.80723, ls, p1, g1, p2, g2, grad=grad)
            np.testing.assert_allclose(a, b, rtol=1e-14, atol=1e-15)
    # coincident points (Matern grad of the Hessian block is zeroed at r = 0)
    p = rng.uniform(size=3)
    a = o.covariance(kernel, 1.3, [1.0, 0.7, 2.0], p, g1, p, g2, grad=True)
    b = r.covariance(kernel, 1.3, [1.0, 0.7, 2.0], p, g1, p, g2, grad=True)
    np.testing.assert_allclose(a, b, rtol=1e-14, atol=1e-15)


def test_cholesky_and_solve(libs):
    o, r = libs
    rng = np.random.default_rng(34187)
    for n in (1, 5, 11, 20, 63):
        A = rng.standard_normal((n, n))
        A = A @ A.T + n * np.eye(n)
        rc_o, Lo = o.cholesky(A)
        rc_r, Lr = r.cholesky(A)
        assert rc_o == rc_r == 0
        np.testing.assert_allclose(np.tril(Lo), np.tril(Lr), rtol=1e-13, atol=1e-14)  # FMA contraction differs
        B = rng.standard_normal((n, 3))
        np.testing.assert_allclose(o.potrs(Lo, B), r.potrs(Lr, B), rtol=1e-11, atol=1e-13)
    # exactly singular (integer arithmetic stays exact): pivot 3 is 0 -> both return 3
    L0 = np.array([[2.0, 0, 0, 0], [1, 3, 0, 0], [4, 1, 0, 0], [2, 2, 1, 5]])
    A = L0 @ L0.T
    assert o.cholesky(A)[0] == r.cholesky(A)[0] == 3
    # known-answer cases from the reference's own tests (gpp_linear_algebra_test.cpp:238-262, 358-368)
    W = np.array([[4.0, 12, -16], [12, 37, -43], [-16, -43, 98]])
    for b in (o, r):
        rc, L = b.cholesky(W)
        assert rc == 0
        np.testing.assert_array_equal(np.tril(L), np.array([[2.0, 0, 0], [6, 1, 0], [-8, 5, 3]]))


def test_limit_update(libs):
    o, r = libs
    rng = np.random.default_rng(3)
    b = np.array([0.0, 1.0, -2.0, 3.0, 0.5, 0.6])
    for _ in range(200):
        x = np.array([rng.uniform(0, 1), rng.uniform(-2, 3), rng.uniform(0.5, 0.6)])
        u = rng.standard_normal(3) * rng.choice([1e-3, 0.1, 1.0, 10.0])
        np.testing.assert_allclose(o.limit_update(b, 0.3, x, u), r.limit_update(b, 0.3, x, u), rtol=1e-15, atol=0)


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("g_idx", [(), (0, 2)])
def test_gp_fit_and_posterior(libs, kernel, g_idx):
    prob = make_problem(24, 3, g_idx=g_idx, seed=11)
    go, gr = _gp_pair(libs, kernel, prob)
    Ko, bo, mo = go.state()
    Kr, br, mr = gr.state()
    np.testing.assert_allclose(np.tril(Ko), np.tril(Kr), rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(bo, br, rtol=1e-10, atol=1e-12)
    assert mo == mr
    pts = np.random.default_rng(2).uniform(size=(4, 3))
    want = ("mean", "grad_mean", "var", "chol_var", "grad_var", "grad_chol")
    for ds in ((), g_idx):
        po = go.posterior(pts, ds, want)
        pr = gr.posterior(pts, ds, want)
        assert po["rc"] == pr["rc"] == 0
        Q = 4 * (1 + len(ds))
        for k in want:
            a, b = po[k], pr[k]
            if k in ("var", "chol_var"):
                # only the lower triangle is defined by the reference
                a = np.tril(a.reshape(Q, Q).T)
                b = np.tril(b.reshape(Q, Q).T)
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-11, err_msg=k)


def test_gp_singular(libs):
    prob = make_problem(10, 2, seed=1)
    prob["X"][1] = prob["X"][0]  # K[:2,:2] = [[1,1],[1,1]] -> second pivot is exactly 0
    prob["noise"][:] = 0.0
    res = [b.gp(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"]) for b in libs]
    assert res[0][0] is None and res[1][0] is None
    assert res[0][1] == res[1][1] == 2


@pytest.mark.parametrize("q,p", [(1, 0), (1, 5), (3, 2), (10, 0)])
def test_ei_table_fed(libs, q, p):
    prob = make_problem(30, 3, seed=4)
    go, gr = _gp_pair(libs, 0, prob)
    rng = np.random.default_rng(3141)
    Xq, Xp = rng.uniform(size=(q, 3)), rng.uniform(size=(p, 3))
    mc = 64
    table = rng.standard_normal(mc * (q + p))
    best = float(prob["y"].min()) + 0.3
    vo, gradо = go.ei(Xq, Xp, mc, best, table, grad=True)
    vr, gradr = gr.ei(Xq, Xp, mc, best, table, grad=True)
    assert vr > 0
    np.testing.assert_allclose(vo, vr, rtol=1e-11)
    np.testing.assert_allclose(gradо, gradr, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("gd", [DISCRETE_ONLY_GD, EXAMPLE_INNER_GD, [1, 20, 3, 3, 0.7, 1.0, 0.2, 1e-7]])
@pytest.mark.parametrize("q,p,g_idx", [(1, 0, ()), (2, 0, ()), (1, 2, ()), (3, 2, ()), (2, 1, (0, 1))])
def test_kg_table_fed(libs, gd, q, p, g_idx):
    prob = make_problem(16, 3, g_idx=g_idx, seed=9, noise=0.1)
    go, gr = _gp_pair(libs, 0, prob)
    rng = np.random.default_rng(27)
    Xq, Xp = rng.uniform(size=(q, 3)), rng.uniform(size=(p, 3))
    disc = rng.uniform(size=(5, 3))
    mc = 16
    Q = (q + p) * (1 + len(g_idx))
    table = rng.standard_normal((mc // 2) * Q)
    best = float(go.mean_additional(disc).min())
    vo, gradо, bpo = go.kg(Xq, Xp, mc, best, table, gd, unit_bounds(3), disc, grad=True, want_best_points=True)
    vr, gradr, bpr = gr.kg(Xq, Xp, mc, best, table, gd, unit_bounds(3), disc, grad=True, want_best_points=True)
    np.testing.assert_allclose(bpo, bpr, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(vo, vr, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gradо, gradr, rtol=1e-6, atol=1e-9)
    # value-only entry point agrees with the value returned by the gradient entry point
    np.testing.assert_allclose(gr.kg(Xq, Xp, mc, best, table, gd, unit_bounds(3), disc), vr, rtol=1e-12)
    np.testing.assert_allclose(go.kg(Xq, Xp, mc, best, table, gd, unit_bounds(3), disc), vo, rtol=1e-12)


def test_kg_fidelity_dim(libs):
    prob = make_problem(14, 3, seed=21, noise=0.1)
    go, gr = _gp_pair(libs, 1, prob)
    rng = np.random.default_rng(8)
    Xq = rng.uniform(size=(2, 3))
    disc = rng.uniform(size=(6, 2))
    table = rng.standard_normal(8 * 2)
    args = (Xq, None, 16, 0.1, table, EXAMPLE_INNER_GD, unit_bounds(2), disc)
    vo, go_ = go.kg(*args, num_fidelity=1, grad=True)
    vr, gr_ = gr.kg(*args, num_fidelity=1, grad=True)
    np.testing.assert_allclose(vo, vr, rtol=1e-9)
    np.testing.assert_allclose(go_, gr_, rtol=1e-6, atol=1e-9)


def _ensemble_inputs(M, dim, g, seed):
    rng = np.random.default_rng(seed)
    hypers = np.concatenate([rng.uniform(0.8, 1.5, size=(M, 1)), rng.uniform(0.4, 0.9, size=(M, dim))], axis=1)
    noises = rng.uniform(0.05, 0.15, size=(M, 1 + g))
    return hypers, noises


@pytest.mark.parametrize("q,p,g_idx,nf", [(1, 0, (), 0), (2, 1, (), 0), (2, 0, (0,), 0), (1, 0, (), 1), (2, 1, (), 1)])
def test_kg_mcmc_table_fed(libs, q, p, g_idx, nf):
    """MCMC-averaged q-KG incl. the fidelity-cost quotient rule (gpp_knowledge_gradient_mcmc_optimization.cpp:87-180)."""
    o, r = libs
    M, dim, mc, num_pts = 3, 3, 8, 4
    prob = make_problem(12, dim, g_idx=g_idx, seed=31, noise=0.1)
    hypers, noises = _ensemble_inputs(M, dim, len(g_idx), 5)
    rng = np.random.default_rng(6)
    Xq, Xp = rng.uniform(0.2, 0.9, size=(q, dim)), rng.uniform(size=(p, dim))
    disc = rng.uniform(size=(M, num_pts, dim - nf))
    best = rng.uniform(-0.5, 0.5, size=M)
    table = rng.standard_normal((mc // 2) * (q + p) * (1 + len(g_idx)))
    args = (hypers, noises, prob["X"], prob["y"], prob["derivs"], Xq, Xp, mc, best, table, EXAMPLE_INNER_GD,
            unit_bounds(dim - nf), disc)
    vo, go_ = o.kg_mcmc(*args, num_fidelity=nf, grad=True)
    vr, gr_ = r.kg_mcmc(*args, num_fidelity=nf, grad=True)
    np.testing.assert_allclose(vo, vr, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(go_, gr_, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(o.kg_mcmc(*args, num_fidelity=nf), vo, rtol=1e-12)


@pytest.mark.parametrize("q,p", [(1, 0), (3, 2)])
def test_ei_mcmc_table_fed(libs, q, p):
    o, r = libs
    M, dim, mc = 4, 3, 16
    prob = make_problem(15, dim, seed=32, noise=0.05)
    hypers, noises = _ensemble_inputs(M, dim, 0, 7)
    rng = np.random.default_rng(8)
    Xq, Xp = rng.uniform(size=(q, dim)), rng.uniform(size=(p, dim))
    best = rng.uniform(0.5, 1.5, size=M)
    table = rng.standard_normal(mc * (q + p))
    args = (hypers, noises, prob["X"], prob["y"], prob["derivs"], Xq, Xp, mc, best, table)
    vo, go_ = o.ei_mcmc(*args, grad=True)
    vr, gr_ = r.ei_mcmc(*args, grad=True)
    np.testing.assert_allclose(vo, vr, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(go_, gr_, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("kernel,g_idx,nf", [(0, (), 0), (1, (0, 2), 0), (1, (), 1)])
def test_posterior_mean_optimization(libs, kernel, g_idx, nf):
    """ComputeOptimalPosteriorMean from one start (gpp_knowledge_gradient_optimization.cpp:420-472)."""
    prob = make_problem(20, 3, g_idx=g_idx, seed=3)
    go, gr = _gp_pair(libs, kernel, prob)
    gd = [1, 50, 3, 0, 0.7, 1.0, 0.2, 1e-8]
    x0 = np.array([0.3, 0.6, 0.5])[: 3 - nf]
    bo, vo = go.posterior_mean_optimization(x0, gd, unit_bounds(3 - nf), nf)
    br, vr = gr.posterior_mean_optimization(x0, gd, unit_bounds(3 - nf), nf)
    np.testing.assert_allclose(bo, br, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(vo, vr, rtol=1e-12)


@pytest.mark.parametrize("kernel,g_idx,N,dim", [(1, (), 20, 3), (0, (), 30, 2), (1, (0, 2), 12, 3), (0, (1,), 15, 2)])
def test_log_marginal_likelihood(libs, kernel, g_idx, N, dim):
    """gpp_model_selection.cpp:540-612 (1e-6 jitter on top of the noise, y centred by the mean of the function values)."""
    o, r = libs
    prob = make_problem(N, dim, g_idx=g_idx, seed=5 + N)
    args = (kernel, 1.3, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    np.testing.assert_allclose(o.log_marginal_likelihood(*args), r.log_marginal_likelihood(*args), rtol=1e-12)


@pytest.mark.parametrize("seed", range(12))
def test_randomized_sweep(libs, seed):
    """Random small configurations (kernel, dimension, derivative observations, q, p, fidelity dims, inner optimiser):
    posterior with all gradients, q-EI and q-KG value + gradient of the restatement against the compiled reference."""
    rng = np.random.default_rng(1000 + seed)
    kernel = int(rng.integers(0, 2))
    dim = int(rng.integers(2, 5))
    g_idx = tuple(sorted(rng.choice(dim, size=int(rng.integers(0, min(3, dim) + 1)), replace=False).tolist()))
    N = int(rng.integers(6, 16))
    q, p = int(rng.integers(1, 4)), int(rng.integers(0, 3))
    nf = int(rng.integers(0, 2)) if dim >= 3 else 0
    prob = make_problem(N, dim, g_idx=g_idx, seed=2000 + seed, noise=float(rng.uniform(0.02, 0.2)))
    go, gr = _gp_pair(libs, kernel, prob)
    pts = rng.uniform(size=(q + p, dim))
    want = ("mean", "grad_mean", "var", "chol_var", "grad_var", "grad_chol")
    po, pr = go.posterior(pts, g_idx, want), gr.posterior(pts, g_idx, want)
    for k in want:
        np.testing.assert_allclose(po[k], pr[k], rtol=1e-8, atol=1e-10, err_msg=k)
    Xq, Xp = pts[:q], pts[q:]
    mc = 8
    t_ei = rng.standard_normal(mc * (q + p))
    best_ei = float(prob["y"][:: 1 + len(g_idx)].min()) + 0.3
    vo, go_ = go.ei(Xq, Xp, mc, best_ei, t_ei, grad=True)
    vr, gr_ = gr.ei(Xq, Xp, mc, best_ei, t_ei, grad=True)
    np.testing.assert_allclose(vo, vr, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(go_, gr_, rtol=1e-7, atol=1e-10)
    disc = rng.uniform(size=(int(rng.integers(1, 6)), dim - nf))
    gd = [1, int(rng.integers(0, 8)), int(rng.integers(1, 3)), 3, float(rng.choice([0.0, 0.7])), float(rng.choice([0.5, 1.0, 2.0])),
          float(rng.choice([0.1, 0.5])), 1e-9]
    t_kg = rng.standard_normal((mc // 2) * (q + p) * (1 + len(g_idx)))
    Xq_f, Xp_f = Xq.copy(), Xp.copy()
    best_kg = float(rng.uniform(-0.5, 0.5))
    ko, kgo, bpo = go.kg(Xq_f, Xp_f, mc, best_kg, t_kg, gd, unit_bounds(dim - nf), disc, num_fidelity=nf, grad=True,
                         want_best_points=True)
    kr, kgr, bpr = gr.kg(Xq_f, Xp_f, mc, best_kg, t_kg, gd, unit_bounds(dim - nf), disc, num_fidelity=nf, grad=True,
                         want_best_points=True)
    np.testing.assert_allclose(bpo, bpr, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(ko, kr, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(kgo, kgr, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("kernel,g_idx,N,dim", [(0, (), 20, 3), (1, (), 20, 3), (0, (0, 2), 10, 3), (0, (1,), 12, 2),
                                               (1, (0,), 10, 2)])
def test_grad_log_marginal_likelihood(libs, kernel, g_idx, N, dim):
    """d log p / d (alpha, lengths, noise by type), gpp_model_selection.cpp:629-677 — incl. the Matern routine's quirk of
    filling only the value-value entry of its block.  Oracle groundwork for SURVEY.md 8f rank 2 (no device path yet)."""
    o, r = libs
    prob = make_problem(N, dim, g_idx=g_idx, seed=5 + N)
    args = (kernel, 1.3, prob["lengths"] * np.linspace(1.0, 1.5, dim), prob["X"], prob["y"], prob["noise"], prob["derivs"])
    a, b = o.grad_log_marginal_likelihood(*args), r.grad_log_marginal_likelihood(*args)
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-11 * np.abs(b).max())
    if kernel == 0:  # and it is the derivative of the (pinned) value: central differences on alpha and the lengths
        base = list(args)
        for k in range(1 + dim):
            def val(h):
                al = 1.3 + (h if k == 0 else 0.0)
                ls = np.array(base[2], dtype=np.float64)
                if k > 0:
                    ls[k - 1] += h
                return o.log_marginal_likelihood(kernel, al, ls, *base[3:])
            fd = (val(1e-5) - val(-1e-5)) / 2e-5
            np.testing.assert_allclose(a[k], fd, rtol=2e-5, atol=1e-6)


@pytest.mark.skipif(not orc.have_reference(), reason="compiled reference not built")
def test_normal_rng_table_replay():
    """The device path is compared with the reference's UNMODIFIED multistart drivers by handing it the normals those
    drivers consume.  Basis: every evaluation rewinds its NormalRNG, so an evaluation driven by NormalRNG(seed) equals —
    bit for bit — one driven by NormalRNGSimulator over the first draws of NormalRNG(seed); and the drivers run."""
    from synth import EXAMPLE_INNER_GD, make_problem, unit_bounds
    ref = orc.load_reference()
    prob = make_problem(40, 3, seed=3)
    gp, lm = ref.gp(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    assert lm == 0
    rng = np.random.default_rng(0)
    q, mc, seed = 2, 32, 1234
    cands = rng.uniform(size=(3, q, 3))
    disc = rng.uniform(size=(6, 3))
    best = float(gp.mean_additional(disc).min())
    vals, grads = orc.kg_grad_at_point_list(ref, gp, cands, None, mc, best, EXAMPLE_INNER_GD, unit_bounds(3), disc, 1,
                                            seed=seed)
    table = orc.normal_draws(seed, (mc // 2) * q)
    for c in range(3):
        v, g = gp.kg(cands[c], None, mc, best, table, EXAMPLE_INNER_GD, unit_bounds(3), disc, grad=True)
        assert v == vals[c]
        np.testing.assert_array_equal(g, grads[c])
    starts = rng.uniform(size=(25, q, 3))
    outer = [1, 4, 1, 0, 0.7, 0.5, 0.3, 1e-7]
    bp, found = orc.ref_multistart_kg(gp, starts, None, mc, best, outer, EXAMPLE_INNER_GD, unit_bounds(3),
                                      unit_bounds(3), disc, seed)
    assert found and np.all((bp >= 0.0) & (bp <= 1.0))
    # the driver's winner is at least as good as the best start it screened
    v_best = gp.kg(bp, None, mc, best, table, EXAMPLE_INNER_GD, unit_bounds(3), disc)
    v_starts = [gp.kg(s, None, mc, best, table, EXAMPLE_INNER_GD, unit_bounds(3), disc) for s in starts]
    assert v_best >= max(v_starts) - 1e-12


@pytest.mark.skipif(not orc.have_reference(), reason="compiled reference not built")
def test_simplex_limit_update_host_logic():
    """SimplexIntersectTensorProductDomain::LimitUpdate (gpp_domain.cpp:234-289) as the multistart drivers' host code
    restates it (cmoe_limit_update needs no device): random points in the simplex, random steps, all three regimes
    (inside, clipped by the box, clipped by the diagonal face), max_relative_change = 1 included."""
    from cornell_moe_b200 import capi
    rng = np.random.default_rng(12)
    hit_face = 0
    for trial in range(400):
        dim = int(rng.integers(2, 7))
        x = rng.dirichlet(np.ones(dim + 1))[:dim] * rng.uniform(0.6, 1.0)
        lo = np.minimum(x, rng.uniform(0.0, 0.2, dim)) * rng.integers(0, 2, dim)
        hi = np.maximum(x, rng.uniform(0.5, 1.3, dim))
        bounds = np.stack([lo, hi], axis=1).ravel()
        upd = (rng.standard_normal(dim) + rng.choice([0.0, 1.0])) * rng.choice([0.01, 0.3, 2.0])
        mrc = float(rng.choice([0.3, 0.8, 1.0]))
        want = orc.ref_limit_update_simplex(bounds, mrc, x, upd)
        got = capi.limit_update(capi.SIMPLEX, bounds, mrc, x, upd)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-300)
        box = capi.limit_update(capi.TENSOR_PRODUCT, bounds, mrc, x, upd)
        ref_box = orc.load_reference().limit_update(bounds, mrc, x, upd.copy())
        np.testing.assert_allclose(box, ref_box if ref_box is not None else box, rtol=1e-14, atol=0)
        hit_face += (x + box).sum() > 1.0 + 1e-9  # the box-limited step would leave the simplex: the face clip fires
    assert hit_face > 20


@pytest.mark.skipif(not orc.have_reference(), reason="compiled reference not built")
def test_reference_driver_state_reuse_quirk():
    """KnowledgeGradientState::SetCurrentPoint does not refresh discretized_set (…optimization.cpp:233-243 vs :259-261):
    an evaluation through a state constructed with other points differs from a fresh one, and the multistart driver —
    which reuses one state built with the first start — ranks its starts by those values.  The device drivers
    reproduce this (cmoe_kg_plan_set_stale_union); this pins the reference side of it."""
    from synth import EXAMPLE_INNER_GD, make_problem, unit_bounds
    ref = orc.load_reference()
    prob = make_problem(30, 3, seed=3, noise=0.05)
    gp, lm = ref.gp(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
    rng = np.random.default_rng(4)
    starts = rng.uniform(0.05, 0.95, size=(40, 2, 3))
    disc = rng.uniform(size=(8, 3))
    q, mc, seed = 2, 64, 4242
    best = float(gp.mean_additional(disc).min())
    b3 = unit_bounds(3)
    table = orc.normal_draws(seed, (mc // 2) * q)
    same = gp.kg_reused_state(starts[2], starts[2], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc)
    assert same == gp.kg(starts[2], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc)
    pick = [0, 5, 18, 34, 26, 25, 31]
    reused = [gp.kg_reused_state(starts[0], starts[i], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc) for i in pick]
    fresh = [gp.kg(starts[i], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc) for i in pick]
    assert reused[0] == fresh[0]
    assert any(abs(a - b) > 1e-6 for a, b in zip(reused[1:], fresh[1:]))
    # one descent step from a single start: the driver's result is start + limited step computed with the start's own
    # (then still fresh) state — and the final value is evaluated through the reused state
    outer1 = [1, 1, 1, 0, 0.7, 0.4, 0.2, 1e-7]
    bp, found = orc.ref_multistart_kg(gp, starts[3:4], None, mc, best, outer1, EXAMPLE_INNER_GD, b3, b3, disc, seed)
    v, g = gp.kg(starts[3], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc, grad=True)
    step = 0.4 * g
    for k in range(q):
        step[k] = ref.limit_update(b3, 0.2, starts[3][k], step[k])
    np.testing.assert_allclose(bp, starts[3] + step, rtol=0, atol=1e-14)
