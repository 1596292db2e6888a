"""SM clock while the N = 5000 factorisation runs back to back (is the latency-bound panel chain running at max clock?)."""
import subprocess, sys, threading, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from cornell_moe_b200 import capi
N, d = 5000, 10
rng = np.random.default_rng(5)
X = rng.uniform(size=(N, d)); y = np.sin(3 * X).sum(axis=1)
gp = capi.GaussianProcess(0, 1.0, np.full(d, 0.5), X, y, [1e-2])
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown",
                      "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
rows = []
threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True).start()
time.sleep(0.5)
t0 = time.time()
us = gp.bench_cholesky(400)
print("chol usec (400 back-to-back):", us, "wall", time.time() - t0)
time.sleep(0.3)
p.terminate()
print("\n".join(rows))
