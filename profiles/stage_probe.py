import os, sys, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/cornell-moe_b200')
from synth import make_problem
import importlib
sys.path.insert(0, '/root/repo')
import importlib.util
spec = importlib.util.spec_from_file_location("capi", "/root/repo/cornell-moe_b200/capi.py")
capi = importlib.util.module_from_spec(spec); spec.loader.exec_module(capi)
prob = make_problem(300, 4, g_idx=(0, 1, 2, 3), seed=44, noise=1e-2)
gp = capi.GaussianProcess(0, prob["alpha"], prob["lengths"], prob["X"], prob["y"], prob["noise"], prob["derivs"])
rng = np.random.default_rng(5)
cands = rng.uniform(size=(8, 4, 4)); disc = rng.uniform(size=(10, 4))
from synth import EXAMPLE_INNER_GD, unit_bounds
res = {}
for st in ("0", "1"):
    os.environ["CMOE_GEN_STAGE"] = st
    kg, grad = gp.kg(cands, None, 2048, 0.1, EXAMPLE_INNER_GD, unit_bounds(4), disc, seed=7, grad=True)
    res[st] = (kg.copy(), grad.copy())
print("kg diff", np.abs(res["0"][0] - res["1"][0]).max(), "rel", (np.abs(res["0"][0] - res["1"][0]) / np.abs(res["0"][0])).max())
print("grad diff", np.abs(res["0"][1] - res["1"][1]).max())
print(res["0"][0][:3], res["1"][0][:3])
