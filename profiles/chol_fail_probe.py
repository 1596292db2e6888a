import sys, numpy as np
sys.path.insert(0, '/root/repo')
from cornell_moe_b200 import capi
n = 1536
rng = np.random.default_rng(3)
G = rng.standard_normal((n, 64))
A = G @ G.T + n * np.eye(n)
bad = int(sys.argv[1])
A[bad, :] = 0.0
A[:, bad] = 0.0
try:
    capi.cholesky(A)
    print("no error?!")
except capi.SingularMatrixError as e:
    print("singular info", e.info)
