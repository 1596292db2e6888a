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


if __name__ == "__main__":
    main()
