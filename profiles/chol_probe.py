"""Timing + residual probe of the large-N fit (config 5) pieces: covariance build, Cholesky, K^-1 y."""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cornell_moe_b200 import capi  # noqa: E402

for N in [int(a) for a in sys.argv[1:]] or [5000]:
    d = 10
    rng = np.random.default_rng(5)
    X = rng.uniform(size=(N, d))
    y = np.sin(3 * X).sum(axis=1) + 0.1 * rng.standard_normal(N)
    t0 = time.time()
    gp = capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, 1.0, np.full(d, 0.5), X, y, [1e-2])
    print(N, "first fit usec (cov, chol, solve):", gp.fit_timings_usec(), "wall", time.time() - t0, flush=True)
    gp2 = capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, 1.0, np.full(d, 0.5), X, y, [1e-2])
    print(N, "second fit usec (cov, chol, solve):", gp2.fit_timings_usec(), flush=True)
    chol = gp.bench_cholesky(10)
    print(N, "warm: cov usec", gp.bench_cov_build(20), " chol usec", chol, " TFLOP/s", N ** 3 / 3 / chol * 1e-6, flush=True)
    L, kinvy, mean = gp2.state()
    L = np.tril(L)
    Xs = X / 0.5
    idx = rng.integers(0, N, size=(4000, 2))
    i, j = np.maximum(idx[:, 0], idx[:, 1]), np.minimum(idx[:, 0], idx[:, 1])
    want = np.exp(-0.5 * ((Xs[i] - Xs[j]) ** 2).sum(axis=1)) + (i == j) * 1e-2
    got = np.einsum("ij,ij->i", L[i], L[j])
    print(N, "max |LL^T - K| on 4000 samples:", np.abs(got - want).max(), flush=True)
