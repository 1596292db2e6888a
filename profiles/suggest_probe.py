"""End-to-end timing of ONE Bayesian-optimisation suggestion as `examples/main.py` of the reference drives it
(README "KG takes ... seconds": Branin, d = 2, q = 4, 16 hyper-parameter samples, num_mc = 2^7, 200 multistarts):
  1. discretisation: MCMC-averaged q-EI multistart (q = 10, num_mc = 2^10)        main.py:170-171
  2. per member: posterior-mean screen of 1000 + N points + posterior_mean_optimization   main.py:172-196
  3. MCMC-averaged q-KG multistart (q = 4, num_mc = 2^7, 200 starts, GD 50 steps x 2 restarts)   main.py:201-204
Everything goes through the C ABI (cornell_moe_b200.capi); Matern-5/2 like the reference's Python surface.
Not the bench metric — a usage-level measurement kept under profiles/."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cornell_moe_b200 import capi  # noqa: E402


def branin(x):
    a, b, c, r, s, t = 1.0, 5.1 / (4 * np.pi ** 2), 5.0 / np.pi, 6.0, 10.0, 1.0 / (8 * np.pi)
    return a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s


rng = np.random.default_rng(2024)
lo, hi = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
bounds = np.stack([lo, hi], axis=1).ravel()
N, M, dim, q = 20, 16, 2, 4
X = lo + (hi - lo) * rng.uniform(size=(N, dim))
y = branin(X)
y = (y - y.mean()) / y.std()
hypers = np.concatenate([rng.uniform(0.8, 1.5, size=(M, 1)), rng.uniform(3.0, 7.0, size=(M, dim))], axis=1)
noises = np.full((M, 1), 1e-4)
outer = [200, 50, 2, 4, 0.7, 1.0, 0.5, 1e-10]      # cpp_sgd_params_kg (main.py:132-139)
inner = [1, 6, 1, 3, 0.0, 1.0, 0.1, 1e-10]         # cpp_sgd_params_ps (main.py:123-130)
ps_gd = [1, 6, 1, 3, 0.0, 1.0, 0.1, 1e-10]


def suggestion(seed):
    t = {}
    t0 = time.perf_counter()
    ens = capi.GaussianProcessEnsemble(hypers, noises, X, y)
    t["fit_16_gps"] = time.perf_counter() - t0
    r = np.random.default_rng(seed)
    t0 = time.perf_counter()
    starts = lo + (hi - lo) * r.uniform(size=(200, 10, dim))
    best = np.full(M, float(y.min()))
    disc_shared, _, _, _ = ens.multistart_ei(starts, None, 1024, best, outer, bounds, seed=seed)
    t["stage1_qEI_mcmc_q10"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    disc = np.zeros((M, 11, dim))
    best_mu = np.zeros(M)
    for m, gp in enumerate(ens.members):
        pts = np.concatenate([lo + (hi - lo) * r.uniform(size=(1000, dim)), X])
        mu = gp.posterior(pts[:, None, :], (), ("mean",))["mean"].ravel()
        x0 = pts[np.argmin(mu)]
        xr, val, _ = capi.posterior_mean_optimization(gp, x0, ps_gd, bounds)
        if -val > mu.min():
            xr = x0
        disc[m, :10] = disc_shared
        disc[m, 10] = xr
        best_mu[m] = gp.posterior(disc[m][:, None, :], (), ("mean",))["mean"].min()
    t["stage2_posterior_mean_screen_and_opt"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    starts = lo + (hi - lo) * r.uniform(size=(200, q, dim))
    pts, val, found, _ = ens.multistart_kg(starts, None, 128, best_mu, outer, inner, bounds, bounds, disc, seed=seed)
    t["stage3_qKG_mcmc_q4"] = time.perf_counter() - t0
    t["total"] = sum(t.values())
    return t, pts, val


suggestion(1)  # warm-up: kernel load, workspace allocation
res, pts, val = suggestion(2)
print(json.dumps({"scenario": "Branin d=2, N=20, 16 hyper samples, q=4, num_mc=128 (KG) / 1024 (EI), 200 multistarts",
                  "reference_readme_seconds": 100.08, "seconds": res, "kg_value": val, "points": pts.tolist()}))
