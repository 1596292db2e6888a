import ctypes, sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cornell_moe_b200 import capi
out = np.zeros(13)
rc = capi.lib().cmoe_bench_chain_latencies(0, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
names = ["dependent DFMA", "RSQ64H seed + 1 op", "shuffle + 1 op", "STS/syncwarp/LDS round trip + op", "block fence by lane 0 + shuffle",
         "library rsqrt + 1 op", "1/sqrt (sqrt + divide) + 1 op", "chol32_warp<fast> per column", "chol32_warp_pair per column", "chol32_warp_pair<lean> per column", "chol32_warp_pipe (software-pipelined) per column", "  same, warp 0 of a 512-thread CTA (128-reg cap)", "  same + two follower warps"]
print("rc", rc)
for n, v in zip(names, out):
    print(f"{n:45s} {v:8.1f} cycles")
