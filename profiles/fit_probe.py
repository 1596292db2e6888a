"""Probe used for profiling the large-N GP fit (config 5): builds a GP at N=5000, d=10 through the C ABI."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cornell_moe_b200 import capi  # noqa: E402

N, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 5000), 10
rng = np.random.default_rng(5)
X = rng.uniform(size=(N, d))
y = np.sin(3 * X).sum(axis=1) + 0.1 * rng.standard_normal(N)
gp = capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, 1.0, np.full(d, 0.5), X, y, [1e-2])
print("fit usec (cov, chol, solve):", gp.fit_timings_usec())
gp2 = capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, 1.0, np.full(d, 0.5), X, y, [1e-2])
print("second fit in the same process (kernels loaded) usec (cov, chol, solve):", gp2.fit_timings_usec())
print("warm: cov usec", gp.bench_cov_build(20), " chol usec", gp.bench_cholesky(5))
