"""Throughput of the BASELINE.json parity configurations on one GPU (device time, steady state).  Not the bench metric —
bench.py measures configs[2]; this records the other shapes for DESIGN.md / profiles."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cornell_moe_b200 import capi  # noqa: E402

INNER = [1, 6, 1, 3, 0.0, 1.0, 0.1, 1e-10]
out = {}


def problem(N, d, g_idx=(), seed=0, noise=1e-2):
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(N, d))
    cols = [np.sin(3 * X).sum(axis=1)] + [3 * np.cos(3 * X[:, a]) for a in g_idx]
    y = np.stack(cols, axis=1) + np.sqrt(noise) * rng.standard_normal((N, 1 + len(g_idx)))
    return X, y.ravel()


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


# config 2: q-EI MC + grad, d=6, N=200, q=4, num_mc=10000, multistart=256
X, y = problem(200, 6, seed=2)
gp = capi.GaussianProcess(0, 1.0, np.full(6, 0.5), X, y, [1e-2])
c2 = np.random.default_rng(7).uniform(size=(256, 4, 6))
t = timed(lambda: gp.ei(c2, None, 10000, float(y.min()), seed=1, grad=True))
out["config2_qEI_N200_d6_q4_mc10000_ms256"] = {"sec_per_eval": t, "mc_samples_per_s": 256 * 10000 / t}

# config 4: d-KG, d=4, N=300, 4 derivative observations, q=4, num_mc=8192 (64 candidates)
X, y = problem(300, 4, g_idx=(0, 1, 2, 3), seed=4)
gp4 = capi.GaussianProcess(0, 1.0, np.full(4, 0.5), X, y, [1e-2] * 5, derivs=(0, 1, 2, 3))
c4 = np.random.default_rng(8).uniform(size=(64, 4, 4))
disc = np.random.default_rng(9).uniform(size=(10, 4))
best = float(gp4.posterior(disc[:, None, :], (), ("mean",))["mean"].min())
t = timed(lambda: gp4.kg(c4, None, 8192, best, INNER, np.tile([0.0, 1.0], 4), disc, seed=1, grad=True), reps=2)
out["config4_dKG_N300_d4_g4_q4_mc8192_ms64"] = {"sec_per_eval": t, "mc_samples_per_s": 64 * 8192 / t,
                                                  "fit_usec": [float(v) for v in gp4.fit_timings_usec()]}

# config 5: large-N fit
X, y = problem(5000, 10, seed=5)
gp5 = capi.GaussianProcess(0, 1.0, np.full(10, 0.5), X, y, [1e-2])
out["config5_fit_N5000_d10"] = {"fit_usec_cov_chol_solve": [float(v) for v in gp5.fit_timings_usec()],
                                "cov_build_usec": gp5.bench_cov_build(30), "cholesky_usec": gp5.bench_cholesky(3)}
print(json.dumps(out))
