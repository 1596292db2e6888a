"""Generates tests/golden/reference_vectors.npz from the UNMODIFIED reference (oracle/_ref/libmoe_ref.so, built from
/root/reference by `make -C oracle ref`).  Run from the repo root in the build container:

    python tests/golden/make_golden.py

The vectors pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) on boxes
where neither /root/reference nor the compiled reference is available."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as orc  # noqa: E402
from synth import EXAMPLE_INNER_GD, DISCRETE_ONLY_GD, make_problem, unit_bounds  # noqa: E402


def main():
    ref = orc.load_reference()
    out = {}
    cases = [("se_plain", 0, 18, 3, ()), ("matern_plain", 1, 18, 3, ()), ("se_deriv", 0, 12, 3, (0, 2)),
             ("matern_deriv", 1, 12, 3, (1,))]
    rng = np.random.default_rng(2026)
    for name, kernel, N, dim, g_idx in cases:
        prob = make_problem(N, dim, g_idx=g_idx, seed=100 + N + kernel, noise=0.05)
        gp, lm = ref.gp(kernel, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
        assert lm == 0
        K, kinvy, mean = gp.state()
        pts = rng.uniform(size=(3, dim))
        post = gp.posterior(pts, g_idx, ("mean", "grad_mean", "var", "chol_var", "grad_var", "grad_chol"))
        out[f"{name}/kernel"] = kernel
        out[f"{name}/g_idx"] = np.array(g_idx, dtype=np.int32)
        for k in ("X", "y", "lengths", "noise"):
            out[f"{name}/{k}"] = prob[k]
        out[f"{name}/K_chol_lower"] = np.tril(K)
        out[f"{name}/K_inv_y"] = kinvy
        out[f"{name}/mean"] = mean
        out[f"{name}/pts"] = pts
        for k, v in post.items():
            if k != "rc":
                out[f"{name}/post_{k}"] = v
        # MC estimators with table-fed normals
        q, p, mc = 2, 1, 32
        Xq, Xp = rng.uniform(size=(q, dim)), rng.uniform(size=(p, dim))
        disc = rng.uniform(size=(5, dim))
        best = float(gp.mean_additional(disc).min())
        t_ei = rng.standard_normal(mc * (q + p))
        ei, gei = gp.ei(Xq, Xp, mc, float(prob["y"][:: 1 + len(g_idx)].min()) + 0.25, t_ei, grad=True)
        out[f"{name}/mc_Xq"], out[f"{name}/mc_Xp"], out[f"{name}/mc_disc"] = Xq, Xp, disc
        out[f"{name}/mc_best_kg"] = best
        out[f"{name}/mc_best_ei"] = float(prob["y"][:: 1 + len(g_idx)].min()) + 0.25
        out[f"{name}/ei_table"], out[f"{name}/ei"], out[f"{name}/ei_grad"] = t_ei, ei, gei
        t_kg = rng.standard_normal((mc // 2) * (q + p) * (1 + len(g_idx)))
        out[f"{name}/kg_table"] = t_kg
        for tag, gd in (("discrete", DISCRETE_ONLY_GD), ("linesearch", EXAMPLE_INNER_GD)):
            kg, gkg, bp = gp.kg(Xq, Xp, mc, best, t_kg, gd, unit_bounds(dim), disc, grad=True, want_best_points=True)
            out[f"{name}/kg_{tag}"], out[f"{name}/kg_{tag}_grad"], out[f"{name}/kg_{tag}_xstar"] = kg, gkg, bp
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


def main_widened():
    """Second fixture file: the rows added from SURVEY.md 8f (MCMC-averaged q-KG / q-EI, posterior-mean optimisation)."""
    ref = orc.load_reference()
    out = {}
    rng = np.random.default_rng(2027)
    for name, dim, g_idx, nf, q, p in (("mcmc_plain", 3, (), 0, 2, 1), ("mcmc_fidelity", 3, (), 1, 2, 0),
                                       ("mcmc_deriv", 3, (0,), 0, 2, 0)):
        M, mc, num_pts = 3, 16, 4
        prob = make_problem(14, dim, g_idx=g_idx, seed=77 + dim + nf, noise=0.1)
        hypers = np.concatenate([rng.uniform(0.8, 1.5, size=(M, 1)), rng.uniform(0.4, 0.9, size=(M, dim))], axis=1)
        noises = rng.uniform(0.05, 0.15, size=(M, 1 + len(g_idx)))
        Xq, Xp = rng.uniform(0.2, 0.9, size=(q, dim)), rng.uniform(size=(p, dim))
        disc = rng.uniform(size=(M, num_pts, dim - nf))
        best = rng.uniform(-0.5, 0.5, size=M)
        t_kg = rng.standard_normal((mc // 2) * (q + p) * (1 + len(g_idx)))
        t_ei = rng.standard_normal(mc * (q + p))
        kg, gkg = ref.kg_mcmc(hypers, noises, prob["X"], prob["y"], prob["derivs"], Xq, Xp, mc, best, t_kg,
                              EXAMPLE_INNER_GD, unit_bounds(dim - nf), disc, num_fidelity=nf, grad=True)
        ei, gei = ref.ei_mcmc(hypers, noises, prob["X"], prob["y"], prob["derivs"], Xq, Xp, mc, best + 1.0, t_ei,
                              grad=True)
        for k, v in dict(hypers=hypers, noises=noises, X=prob["X"], y=prob["y"], g_idx=np.array(g_idx, dtype=np.int32),
                         nf=nf, Xq=Xq, Xp=Xp, disc=disc, best=best, kg_table=t_kg, ei_table=t_ei, kg=kg, kg_grad=gkg,
                         ei=ei, ei_grad=gei).items():
            out[f"{name}/{k}"] = v
    for name, kernel, g_idx, nf in (("pmopt_se", 0, (), 0), ("pmopt_matern_deriv", 1, (0, 2), 0), ("pmopt_fid", 1, (), 1)):
        prob = make_problem(20, 3, g_idx=g_idx, seed=3)
        gp, lm = ref.gp(kernel, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
        gd = np.array([1, 50, 3, 0, 0.7, 1.0, 0.2, 1e-8])
        x0 = np.array([0.3, 0.6, 0.5])[: 3 - nf]
        bp, val = gp.posterior_mean_optimization(x0, gd, unit_bounds(3 - nf), nf)
        for k, v in dict(kernel=kernel, g_idx=np.array(g_idx, dtype=np.int32), nf=nf, X=prob["X"], y=prob["y"],
                         lengths=prob["lengths"], noise=prob["noise"], gd=gd, x0=x0, best_point=bp, best_value=val).items():
            out[f"{name}/{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors_widened.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    if "--widened-only" not in sys.argv:
        main()
    main_widened()
