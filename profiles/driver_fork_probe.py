"""Per-start comparison of the device multistart driver with the reference's own driver (same normals)."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0]); sys.path.insert(0, __file__.rsplit("/", 2)[0] + "/tests")
import oracle as orc
from cornell_moe_b200 import capi
from synth import EXAMPLE_INNER_GD, make_problem, unit_bounds

kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 0
prob = make_problem(30, 3, seed=3, noise=0.05)
gp = capi.GaussianProcess(kernel, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
ref, lm = orc.load_reference().gp(kernel, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"])
rng = np.random.default_rng(4)
starts = rng.uniform(0.05, 0.95, size=(40, 2, 3))
disc = rng.uniform(size=(8, 3))
q, mc, seed = 2, 64, 4242
best = float(ref.mean_additional(disc).min())
b3 = unit_bounds(3)
table = orc.normal_draws(seed, (mc // 2) * q)
for outer in ([40, 6, 2, 0, 0.7, 0.4, 0.2, 1e-7], [40, 1, 1, 0, 0.7, 0.4, 0.2, 1e-7]):
    print("outer", outer)
    for i in range(12):
        bp_ref, f_ref = orc.ref_multistart_kg(ref, starts[i:i + 1], None, mc, best, outer, EXAMPLE_INNER_GD, b3, b3, disc, seed)
        bp, bv, f, sv = capi.multistart_kg(gp, starts[i:i + 1], None, mc, best, outer, EXAMPLE_INNER_GD, b3, b3, disc,
                                           seed=1, table=table)
        v_ref = ref.kg(bp_ref, None, mc, best, table, EXAMPLE_INNER_GD, b3, disc)
        # one gradient evaluation at the start, both sides
        v0, g0 = ref.kg(starts[i], None, mc, best, table, EXAMPLE_INNER_GD, b3, disc, grad=True)
        kg0, gg0 = gp.kg(starts[i:i + 1], None, mc, best, EXAMPLE_INNER_GD, b3, disc, table=table, grad=True)
        print(i, "max|dx|=%.2e" % np.abs(bp - bp_ref).max(), "v ours %.6e ref %.6e" % (bv, v_ref),
              "grad0 maxdiff %.1e (|g| %.1e)" % (np.abs(gg0[0] - g0).max(), np.abs(g0).max()), f, f_ref)
