"""Probe for the covariance-build kernel (config 5: N=5000, d=10): time per launch + entry-wise check against numpy.
Usage: [CMOE_B200_LIB=variants/libvar_x.so] python profiles/cov_probe.py [N] [kernel]"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cornell_moe_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else capi.SQUARE_EXPONENTIAL
d = 10
rng = np.random.default_rng(5)
X = rng.uniform(size=(N, d))
y = np.sin(3 * X).sum(axis=1) + 0.1 * rng.standard_normal(N)
ls = np.full(d, 0.5)
gp = capi.GaussianProcess(kernel, 1.3, ls, X, y, [1e-2])
us = gp.bench_cov_build(30)
L = gp.state()[0]  # [n, n] lower factor, row-major view
# the factor of the freshly built matrix reproduces it: check L L^T against the closed form on random entries
idx = rng.integers(0, N, size=(2000, 2))
i, j = np.maximum(idx[:, 0], idx[:, 1]), np.minimum(idx[:, 0], idx[:, 1])
r2 = (((X[i] - X[j]) / ls) ** 2).sum(axis=1)
if kernel == capi.SQUARE_EXPONENTIAL:
    want = 1.3 * np.exp(-0.5 * r2)
else:
    a = np.sqrt(5.0 * r2)
    want = 1.3 * np.exp(-a) * (1 + a + 5.0 / 3.0 * r2)
want = want + (i == j) * 1e-2
Lm = np.tril(L)
got = np.einsum("ij,ij->i", Lm[i], Lm[j])
algo = 4.0 * N * (N + 1) + 8.0 * N * d
print(f"cov_build N={N} kernel={kernel}: {us:.2f} us/launch, {algo / us * 1e-3:.1f} GB/s (lower triangle), "
      f"max |LL^T - K| on 2000 entries = {np.abs(got - want).max():.2e}")
