"""Worker for test_sharded_multistart_under_torchrun_equals_single_process: one process per GPU (torchrun, NCCL)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from synth import EXAMPLE_INNER_GD, make_problem, unit_bounds  # noqa: E402


def problem(capi, device):
    prob = make_problem(40, 3, seed=21, noise=0.05)
    gp = capi.GaussianProcess(0, 1.0, prob["lengths"], prob["X"], prob["y"], prob["noise"], device=device)
    rng = np.random.default_rng(22)
    starts = rng.uniform(0.05, 0.95, size=(45, 2, 3))
    disc = rng.uniform(size=(8, 3))
    best = float(gp.posterior(disc[:, None, :], (), ("mean",))["mean"].min())
    outer = [45, 4, 1, 0, 0.7, 0.4, 0.2, 1e-7]
    b3 = unit_bounds(3)
    return gp, (starts, None, 128, best, outer, EXAMPLE_INNER_GD, b3, b3, disc)


def main():
    import torch
    import torch.distributed as dist
    from cornell_moe_b200 import capi, multigpu
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gp, args = problem(capi, local)
    bp, bv, found, sv = multigpu.multistart_kg(gp, *args, device=f"cuda:{local}")
    if rank == 0:
        np.savez(sys.argv[1], best_point=bp, best_value=bv, found=found, start_values=sv)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
