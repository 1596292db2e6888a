#!/usr/bin/env python
"""bench.py — throughput of the GP-posterior + Monte-Carlo acquisition hot path on B200 (BASELINE.json).

Default (`--config c3`): q-KG Monte-Carlo sample-evaluations per second at the north-star shape.  A "step" = one batched
q-KG value+gradient evaluation (what one outer optimiser step / the multistart pre-screen consumes) of `multistart`
candidates x `num_mc` samples at N=500, d=8, q=8 on synthetic data.  With --gpus N > 1 the candidates are sharded
over ranks (one process per GPU, launched by torchrun) with no data-path collective; the path's only exchange — one
all-gather of the per-candidate values followed by the arg-max every rank computes identically — runs INSIDE the timed
step (strong scaling: total work fixed).

Other BASELINE.json configurations, same JSON contract:
  --config c2   q-EI MC value+gradient, d=6, N=200, q=4, num_mc=10000, multistart=256           (configs[1])
  --config c4   d-KG (4 derivative observations per point), d=4, N=300, q=4, num_mc=8192         (configs[3]; --gpus 4)
  --config c5   large-N GP fit, N=5000, d=10: covariance build + blocked Cholesky + K^-1 y        (configs[4]; replicas)
  --kernel matern52 switches the covariance kernel (the reference's Python front end only ever builds Matern-5/2).

Output: ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM before the timed region, CUDA
events on the launching stream, max over ranks); `e2e` is the same metric through the host-buffer C-ABI call (plan
creation + H2D + kernels + D2H inside the timed region).  `roofline` describes the dominant kernel; `cpu_baseline` is
the reference's own C++ path (oracle/_ref) timed on this host's cores on a bounded sample.
--impl reference times the CPU reference alone (rank 0 only; thread count from the CPU affinity mask, NOT from
OMP_NUM_THREADS, which torchrun forces to 1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: the output is ONE JSON line

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

INNER_GD = [1, 6, 1, 3, 0.0, 1.0, 0.1, 1e-10]  # the reference examples' inner optimiser (examples/main.py:123-130)
SEED_PHILOX = 0xC0FFEE
KERNELS = {"se": 0, "matern52": 1}

CONFIGS = {
    "c3": dict(kind="kg", N=500, dim=8, q=8, num_mc=16384, multistart=1024, g=(), num_pts=10, noise=1e-2, length=0.5,
               alpha=1.0, baseline="BASELINE.json configs[2]",
               metric="q-KG MC sample-evals/sec at (N=500,d=8,q=8,mc=16384)"),
    "c4": dict(kind="kg", N=300, dim=4, q=4, num_mc=8192, multistart=256, g=(0, 1, 2, 3), num_pts=10, noise=1e-2,
               length=0.5, alpha=1.0, baseline="BASELINE.json configs[3]",
               metric="d-KG MC sample-evals/sec at (N=300,d=4,4 derivative obs,q=4,mc=8192)"),
    "c2": dict(kind="ei", N=200, dim=6, q=4, num_mc=10000, multistart=256, g=(), num_pts=0, noise=1e-2, length=0.5,
               alpha=1.0, baseline="BASELINE.json configs[1]",
               metric="q-EI MC sample-evals/sec at (N=200,d=6,q=4,mc=10000)"),
    "c5": dict(kind="fit", N=5000, dim=10, q=0, num_mc=0, multistart=0, g=(), num_pts=0, noise=1e-2, length=0.5,
               alpha=1.0, baseline="BASELINE.json configs[4]",
               metric="large-N GP fit (cov build + Cholesky + K^-1 y) TFLOP/s on n^3/3 at (N=5000,d=10)"),
}


def make_workload(w):
    rng = np.random.default_rng(20260924)
    X = rng.uniform(size=(w["N"], w["dim"]))
    g = tuple(w["g"])
    cols = [np.sin(3.0 * X).sum(axis=1)] + [3.0 * np.cos(3.0 * X[:, a]) for a in g]
    y = np.stack(cols, axis=1) + np.sqrt(w["noise"]) * rng.standard_normal((w["N"], 1 + len(g)))
    cands = np.random.default_rng(7).uniform(size=(max(1, w["multistart"]), max(1, w["q"]), w["dim"]))
    disc = np.random.default_rng(11).uniform(size=(max(1, w["num_pts"]), w["dim"]))
    return dict(X=X, y=y.ravel(), lengths=np.full(w["dim"], w["length"]), noise=np.full(1 + len(g), w["noise"]),
                derivs=g, cands=cands, disc=disc, bounds=np.tile([0.0, 1.0], w["dim"]))


def host_cpus():
    """(threads usable by this process, physical cores among them).  The affinity mask is what the process may use;
    torchrun's OMP_NUM_THREADS=1 is deliberately ignored (the reference arm passes its thread count explicitly)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except Exception:
        cpus = list(range(os.cpu_count() or 1))
    cores = set()
    try:
        cur = {}
        for line in open("/proc/cpuinfo"):
            if ":" in line:
                k, v = [t.strip() for t in line.split(":", 1)]
                cur[k] = v
            elif cur:
                if int(cur.get("processor", -1)) in cpus:
                    cores.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
                cur = {}
        if cur and int(cur.get("processor", -1)) in cpus:
            cores.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
    except Exception:
        pass
    return len(cpus), (len(cores) if cores else max(1, len(cpus) // 2))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref = the unmodified reference C++ compiled from /root/reference; else the C port)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_backend():
    import oracle as orc
    return orc.load_reference() if orc.have_reference() else orc.load_oracle()


def cpu_gp(backend, w, wl, kernel):
    gp, lm = backend.gp(kernel, w["alpha"], wl["lengths"], wl["X"], wl["y"], wl["noise"], wl["derivs"])
    assert lm == 0
    return gp


def cpu_step(backend, gp, w, wl, cands, num_mc, threads, best):
    """One bounded sample of the workload on the host cores; returns (units processed, seconds)."""
    import oracle as orc
    t0 = time.perf_counter()
    if w["kind"] == "kg":
        orc.kg_grad_at_point_list(backend, gp, cands, None, num_mc, best, INNER_GD, wl["bounds"], wl["disc"], threads,
                                  want_grad=True)
    else:
        orc.ei_grad_at_point_list(backend, gp, cands, None, num_mc, best, threads)
    return cands.shape[0] * num_mc, time.perf_counter() - t0


def cpu_fit_sample(backend, w, kernel, N):
    """The reference's GaussianProcess constructor (single-threaded by construction) on a bounded N; TFLOP/s on n^3/3."""
    ws = dict(w, N=N)
    wls = make_workload(ws)
    t0 = time.perf_counter()
    cpu_gp(backend, ws, wls, kernel)
    dt = time.perf_counter() - t0
    return N ** 3 / 3.0 / dt * 1e-12, dt


def best_so_far_cpu(backend, gp, w, wl):
    if w["kind"] == "kg":
        return float(gp.mean_additional(wl["disc"]).min())  # py/cpp_wrappers/knowledge_gradient.py:361-368
    return float(wl["y"].reshape(w["N"], -1)[:, 0].min())


def cpu_sample_shape(w, threads):
    # one candidate per thread (OpenMP static over candidates), enough MC samples that the per-candidate set-up
    # (posterior state, Cholesky of the q x q variance) is amortised as it is for the GPU arm
    # (derivative observations make a CPU sample ~7x dearer — an n = N(1+g) system behind every evaluation; 512 samples per
    # candidate: ~40 s on 128 threads.  Fewer would leave the per-candidate set-up unamortised — 256 samples measured
    # 1.0e3 sample-evals/s against 1.9e3 at 1024 — and understate the CPU.)
    if w["kind"] == "kg" and w["g"]:
        return threads, 512
    return threads, {"kg": 1024, "ei": 1 << 20}[w["kind"]]


def reference_arm(args, w, wl, kernel, config):
    backend = cpu_backend()
    threads, cores = host_cpus()
    base = {"impl": "reference", "metric": w["metric"], "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "gpu_launches": 0}
    if w["kind"] == "fit":
        Ns = 2000
        cpu_fit_sample(backend, w, kernel, 500)  # warm-up
        vals = [cpu_fit_sample(backend, w, kernel, Ns) for _ in range(max(1, args.steps))]
        tf = float(np.mean([v[0] for v in vals]))
        ms = 1e3 * float(np.mean([v[1] for v in vals]))
        base.update({"value": tf, "unit": "TFLOP/s", "ms_per_step": ms,
                     "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": 1, "threads": 1, "kind": backend.name,
                                      "sample": f"GaussianProcess constructor at N={Ns}, d={w['dim']} (single-threaded "
                                                f"by construction), {ms:.0f} ms per fit"},
                     "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return
    gp = cpu_gp(backend, w, wl, kernel)
    best = best_so_far_cpu(backend, gp, w, wl)
    # give the reference its best thread count: all hardware threads or one per physical core (SMT often hurts this
    # FP64-heavy loop); calibrated on a short run, which doubles as the warm-up
    trial = {}
    for tt in sorted({threads, cores}):
        cc = np.resize(wl["cands"], (tt,) + wl["cands"].shape[1:])
        for _ in range(max(1, args.warmup // 2)):
            units, dt = cpu_step(backend, gp, w, wl, cc, 64 if w["kind"] == "kg" else 4096, tt, best)
        trial[tt] = units / dt
    use = max(trial, key=trial.get)
    sc, smc = cpu_sample_shape(w, use)
    cands = np.resize(wl["cands"], (sc,) + wl["cands"].shape[1:])
    runs = [cpu_step(backend, gp, w, wl, cands, smc, use, best) for _ in range(args.steps)]
    total_t = sum(r[1] for r in runs)
    value = sum(r[0] for r in runs) / total_t
    sample = (f"{sc} candidates x {smc} MC samples per step (value+gradient), OpenMP static over candidates, "
              f"{use} threads on {cores} physical cores ({threads} hardware threads available)")
    base.update({"value": value, "unit": "sample-evals/s", "ms_per_step": 1e3 * total_t / args.steps,
                 "cpu_baseline": {"value": value, "unit": "sample-evals/s", "cores": min(use, cores), "threads": use,
                                  "kind": backend.name, "sample": sample},
                 "e2e": {"value": value, "unit": "sample-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return peaks.get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def fit_extra(capi, device, fp64_dmma, with_cpu):
    """Config-5 kernels: covariance build vs HBM, blocked Cholesky vs the DMMA peak, K^-1 y; reference ctor beside it."""
    hbm, src = hbm_peak()
    Nl, dl = 5000, 10
    w5 = CONFIGS["c5"]
    wl5 = make_workload(w5)
    mk = lambda: capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, 1.0, wl5["lengths"], wl5["X"], wl5["y"], wl5["noise"],  # noqa: E731
                                      device=device)
    gpl = mk()
    del gpl  # first fit of the process = kernel loading; the reported fit is the second one
    gpl = mk()
    t_cov = gpl.bench_cov_build(20)
    capi.set_option("cov_tma", 1)  # the TMA / DMMA variant (tensor-map tile stores), off by default: slower at d = 10
    try:
        t_cov_tma = gpl.bench_cov_build(20)
    finally:
        capi.set_option("cov_tma", 0)
    t_chol = gpl.bench_cholesky(5)
    fit = [float(x) for x in gpl.fit_timings_usec()]
    cov_bytes = 4.0 * Nl * (Nl + 1) + 8.0 * Nl * dl
    out = {
        "cov_build_N5000_d10": {"bound": "hbm", "usec": t_cov, "achieved": cov_bytes / t_cov * 1e-3, "peak": hbm,
                                "unit": "GB/s", "frac": cov_bytes / t_cov * 1e-3 / hbm, "peak_source": src,
                                "kernel": "cov_build_g0_kernel (LDGSTS, default)",
                                "tma_variant_usec": t_cov_tma, "tma_variant_frac": cov_bytes / t_cov_tma * 1e-3 / hbm},
        "cholesky_N5000": {"bound": "tensor (FP64 DMMA)", "usec": t_chol, "achieved": Nl ** 3 / 3.0 / t_chol * 1e-6,
                           "peak": fp64_dmma, "unit": "TFLOP/s", "frac": Nl ** 3 / 3.0 / t_chol * 1e-6 / fp64_dmma,
                           "frac_of_nominal_40": Nl ** 3 / 3.0 / t_chol * 1e-6 / 40.0,
                           "peak_source": "measured live: DMMA m8n8k4 microbenchmark"},
        "kinv_y_N5000": {"bound": "hbm (two passes over the factor)", "usec": fit[2],
                         "achieved": 2 * 4.0 * Nl * (Nl + 1) / fit[2] * 1e-3, "peak": hbm, "unit": "GB/s",
                         "frac": 2 * 4.0 * Nl * (Nl + 1) / fit[2] * 1e-3 / hbm},
        "gp_fit_N5000_usec_cov_chol_solve": fit}
    # the round-1 kernels on the same box (CMOE_LEGACY_LINALG=1 is read once per process, hence the subprocess)
    try:
        env = dict(os.environ, CMOE_LEGACY_LINALG="1")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "chol_probe.py"), "5000"], env=env,
                           capture_output=True, text=True, timeout=240)
        for line in r.stdout.splitlines():
            if "warm:" in line:
                tok = line.split()
                out["legacy_round1_kernels"] = {"cov_build_usec": float(tok[tok.index("cov") + 2]),
                                                "cholesky_usec": float(tok[tok.index("chol") + 2])}
            if "second fit" in line:
                out.setdefault("legacy_round1_kernels", {})["fit_line"] = line.strip()
    except Exception as e:
        out["legacy_round1_kernels"] = {"error": str(e)}
    if with_cpu:
        try:
            backend = cpu_backend()
            cpu_fit_sample(backend, w5, 0, 300)
            tf, dt = cpu_fit_sample(backend, w5, 0, 2000)
            out["gp_fit_cpu_baseline"] = {"value": tf, "unit": "TFLOP/s on n^3/3", "cores": 1, "kind": backend.name,
                                          "sample": f"reference GaussianProcess constructor at N=2000, d=10: {dt:.2f} s",
                                          "gpu_value": Nl ** 3 / 3.0 / sum(fit) * 1e-6}
        except Exception as e:
            out["gp_fit_cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": str(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--kernel", default="se", choices=sorted(KERNELS))
    ap.add_argument("--multistart", type=int, default=None)
    ap.add_argument("--num-mc", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = dict(CONFIGS[args.config])
    if args.multistart:
        w["multistart"] = args.multistart
    if args.num_mc:
        w["num_mc"] = args.num_mc
    kernel = KERNELS[args.kernel]
    wl = make_workload(w)
    if w["kind"] == "fit":
        desc = (f"GP fit, synthetic {args.kernel} GP, N={w['N']}, d={w['dim']}: covariance build + blocked Cholesky + "
                f"K^-1 y ({w['baseline']})")
        sharding = f"replicas only: {world} independent fit(s), no collective"
    else:
        nd = len(w["g"])
        desc = (f"{'q-KG' if w['kind'] == 'kg' else 'q-EI'} value+gradient, synthetic {args.kernel} GP, N={w['N']}, "
                f"d={w['dim']}, q={w['q']}, num_mc={w['num_mc']}, multistart={w['multistart']}"
                + (f", {nd} derivative observations per point (n={w['N'] * (1 + nd)})" if nd else "")
                + (f", discrete_pts={w['num_pts']}, inner GD steps=6 restarts=1" if w["kind"] == "kg" else "")
                + f" ({w['baseline']})")
        sharding = f"candidates strided over {world} rank(s); value all-gather + arg-max inside the timed step"
    config = {"workload": desc, "sharding": sharding, "l2": "flushed between timed steps (256 MiB memset)",
              "y": "sum_k sin(3 x_k) (+ its partial derivatives for derivative observations) + N(0, noise); "
                   "SURVEY 8(d) allows this generator in place of a prior draw"}

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, w, wl, kernel, config)
        return

    import torch
    from cornell_moe_b200 import capi
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank if world > 1 else 0
    assert capi.device_count() > device, "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(device)
    tdev = f"cuda:{device}"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=tdev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=tdev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    fp64_fma, fp64_dmma = capi.fp64_peaks(device)
    sampler = ClockSampler(device)

    # ------------------------------------------------------------------------------------------------------------
    # config 5: the fit itself is the step (replicas only)
    # ------------------------------------------------------------------------------------------------------------
    if w["kind"] == "fit":
        mk = lambda: capi.GaussianProcess(kernel, w["alpha"], wl["lengths"], wl["X"], wl["y"], wl["noise"], device=device)  # noqa: E731
        for _ in range(args.warmup):
            mk()
        barrier()
        if rank == 0:
            sampler.start()
        dev_us, e2e_s, parts = 0.0, 0.0, np.zeros(3)
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g5 = mk()
            e2e_s += time.perf_counter() - t0
            parts += g5.fit_timings_usec()
            dev_us += float(np.sum(g5.fit_timings_usec()))
            del g5
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        dev_us, e2e_s = reduce_max([dev_us, e2e_s])
        if rank != 0:
            if world > 1:
                dist.destroy_process_group()
            return
        flops = w["N"] ** 3 / 3.0
        ms = dev_us / args.steps * 1e-3
        parts /= args.steps
        out = {"metric": w["metric"], "value": world * flops / (ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "e2e": {"value": world * flops / (e2e_s / args.steps) * 1e-12, "unit": "TFLOP/s",
                       "h2d_bytes_per_step": int(wl["X"].nbytes + wl["y"].nbytes), "d2h_bytes_per_step": 4},
               "gpu_launches": None, "clocks": clocks,
               "roofline": {"bound": "tensor", "kernel": "chol_step_kernel (potrf_coop.cu)",
                            "achieved": flops / (parts[1] * 1e-6) * 1e-12, "peak": fp64_dmma, "unit": "TFLOP/s",
                            "frac": flops / (parts[1] * 1e-6) * 1e-12 / fp64_dmma, "traffic": None,
                            "peak_source": "measured live: DMMA m8n8k4 microbenchmark (MEASURED_PEAKS.json has no FP64 "
                                           "entry); nominal B200 FP64 is 40 TFLOP/s",
                            "frac_of_nominal_40": flops / (parts[1] * 1e-6) * 1e-12 / 40.0,
                            "kernel_share_of_step": float(parts[1] / parts.sum())},
               "fit_usec_cov_chol_solve": [float(x) for x in parts]}
        if not args.no_extra:
            out["extra"] = fit_extra(capi, device, fp64_dmma, with_cpu=not args.no_cpu_baseline)
        print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------------------------------------------------
    # MC acquisition configs (c3 / c4: q-KG, c2: q-EI)
    # ------------------------------------------------------------------------------------------------------------
    gp = capi.GaussianProcess(kernel, w["alpha"], wl["lengths"], wl["X"], wl["y"], wl["noise"], wl["derivs"],
                              device=device)
    ncand = len(wl["cands"])
    my = wl["cands"][rank::world]
    slots = int(np.ceil(ncand / world))

    def exchange(vals):
        """The path's one exchange: all-gather of the per-candidate values, identical arg-max on every rank."""
        if world == 1:
            return vals, int(np.argmax(vals))
        loc = torch.full((slots,), float("-inf"), dtype=torch.float64, device=tdev)
        loc[: len(vals)] = torch.from_numpy(vals).to(tdev)
        allv = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(allv, loc)
        full = torch.stack(allv, dim=1).reshape(-1)[:ncand]  # index = local * world + rank
        return full.cpu().numpy(), int(torch.argmax(full).item())

    stats = None
    if w["kind"] == "kg":
        # best_so_far = min posterior mean over the discrete set, as the reference's Python wrapper computes it
        best = float(gp.posterior(wl["disc"][:, None, :], (), ("mean",))["mean"].min())
        plan = capi.KGPlan(gp, w["num_mc"], best, INNER_GD, wl["bounds"], wl["disc"], len(my), w["q"], seed=SEED_PHILOX,
                           want_grad=True)
        plan.upload(my)

        def device_step():
            flush.zero_()
            torch.cuda.synchronize()
            plan.run()
            plan.sync()
            tot, mc, nl = plan.timings()
            t0 = time.perf_counter()
            vals = plan.download()[0]
            full, am = exchange(vals)
            torch.cuda.synchronize()
            return tot, mc, nl, 1e3 * (time.perf_counter() - t0), full, am

        def host_step():
            vals, grad = gp.kg(my, None, w["num_mc"], best, INNER_GD, wl["bounds"], wl["disc"], seed=SEED_PHILOX,
                               grad=True)
            full, am = exchange(vals)
            return vals, grad, full, am
    else:
        best = float(wl["y"].min())
        plan = None

        def device_step():
            # q-EI has no device-resident plan API: the device-side figure is the host call minus nothing — `value`
            # and `e2e` are both measured through cmoe_ei_eval (host buffers)
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            vals, _ = gp.ei(my, None, w["num_mc"], best, seed=SEED_PHILOX, grad=True)
            tot = 1e3 * (time.perf_counter() - t0)
            t1 = time.perf_counter()
            full, am = exchange(vals)
            torch.cuda.synchronize()
            return tot, tot, 0, 1e3 * (time.perf_counter() - t1), full, am

        def host_step():
            vals, grad = gp.ei(my, None, w["num_mc"], best, seed=SEED_PHILOX, grad=True)
            full, am = exchange(vals)
            return vals, grad, full, am

    for _ in range(args.warmup):
        device_step()
    barrier()
    if rank == 0:
        sampler.start()
    t_host0 = time.perf_counter()
    dev_ms, mc_ms, exch_ms, launches = 0.0, 0.0, 0.0, 0
    for _ in range(args.steps):
        tot, mc, nl, ex, full, argmax = device_step()
        dev_ms += tot
        mc_ms += mc
        exch_ms += ex
        launches += nl
    barrier()
    t_host = time.perf_counter() - t_host0
    clocks = sampler.stop() if rank == 0 else None
    if plan is not None:
        kg, grad, stats = plan.download()
        del plan  # free the device-resident plan's workspace before the host-API runs allocate theirs
    else:
        kg, grad = None, None

    # end-to-end through the host-buffer API (plan creation, H2D, kernels, D2H, exchange inside the timed region)
    e2e_steps = []
    for it in range(1 + args.steps):  # first call = warm-up (allocates the handle's cached workspace)
        barrier()
        t0 = time.perf_counter()
        kg_e, grad_e, full_e, argmax_e = host_step()
        barrier()
        if it > 0:
            e2e_steps.append(time.perf_counter() - t0)
    e2e_s = float(np.sum(e2e_steps))
    if kg is not None:
        assert np.array_equal(kg_e, kg), "host-API result differs from the device-resident plan"
    assert argmax_e == argmax
    dev_ms, mc_ms, exch_ms, e2e_s, t_host = reduce_max([dev_ms, mc_ms, exch_ms, e2e_s, t_host])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_samples = ncand * w["num_mc"]
    ms_per_step = (dev_ms + exch_ms) / args.steps
    value = total_samples / (ms_per_step * 1e-3)
    e2e_value = total_samples / (e2e_s / args.steps)
    out = {"metric": w["metric"], "value": value, "unit": "sample-evals/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "e2e": {"value": e2e_value, "unit": "sample-evals/s", "h2d_bytes_per_step": int(wl["cands"].nbytes),
                   "d2h_bytes_per_step": int(8 * ncand + grad_e.nbytes * world)},
           "gpu_launches": launches, "clocks": clocks, "exchange_us_per_step": 1e3 * exch_ms / args.steps,
           "device_ms_per_step_without_exchange": dev_ms / args.steps,
           "host_wall_ms_per_step": 1e3 * t_host / args.steps, "argmax_index": argmax,
           "kg_checksum": float(np.sum(full))}

    if w["kind"] == "kg":
        # roofline of the dominant kernel: FP64 flops of the work it EXECUTED (from its own counters) / its time.
        # Conventions of SURVEY.md 8(d): FMA = 2 flops, exp = 1 flop.  Per training/union row (b = 1 + #derivative obs):
        #   point evaluation (value + gradient at one query point): dot 2d, weights 2qb, exp 1, value 2, scale 1, grad 2d
        #   line batch (all KB = 8 backtracking trials of a step):  two dots 4d, weights 2qb, two exps 2, scale 1,
        #                                                            KB fma 2KB, KB-1 squarings
        #   Matern-5/2 line batch (KB = 6 value-only trials sharing dots and weights): 4d + 2q + 3, then per trial
        #                                                            distance 3, sqrt 1, exp 1, polynomial + sum 6
        #   derivative observations (g > 0): the (1+g) Q weight FMAs per training point are executed ONCE per sample
        #   (per-lane weight columns), an evaluation then costs per POINT: dot 2d, exp 1, 5 per derivative row, 9 for
        #   value / gradient weights, grad 2d; a line batch per POINT: 4d, 5 per derivative row, two exps, 3, 5 KB - 1
        b = 1 + len(w["g"])
        rows = w["N"] * b + w["q"] * b
        flops_sample = 0
        if w["g"]:
            KB = 8
            pts = w["N"] + w["q"]
            flops_point = pts * (4 * w["dim"] + 5 * len(w["g"]) + 10)
            flops_line = pts * (4 * w["dim"] + 5 * len(w["g"]) + 5 + 5 * KB - 1)
            flops_sample = 2 * w["N"] * b * (w["q"] * b)
        elif kernel == 1:
            KB = 6
            flops_point = rows * (2 * w["dim"] + 2 * w["q"] + 1 + 1 + 8 + 2 * w["dim"])
            flops_line = rows * (4 * w["dim"] + 2 * w["q"] + 3 + KB * 11)
        else:
            KB = 8
            flops_point = rows * (2 * w["dim"] + 2 * w["q"] * b + 1 + 2 + 1 + 2 * w["dim"])
            flops_line = rows * (4 * w["dim"] + 2 * w["q"] * b + 2 + 1 + 2 * KB + (KB - 1))
        executed = (stats["point_evals"] * flops_point + stats["line_batches"] * flops_line +
                    stats["mc_samples"] * flops_sample)
        achieved = executed / (mc_ms / args.steps * 1e-3) * 1e-12
        # DRAM traffic of the same kernel from the committed `ncu --set full` capture (dram__bytes_read + write); the
        # capture ran 256 candidates per launch and the traffic is per-sample records: it scales with the candidates
        traffic, traffic_src = None, None
        cap_file = {"c3": "r1_kg_mc_kernel_ncu.json", "c4": "r2_kg_mc_gen_ncu.json"}.get(args.config)
        if cap_file and kernel == 0:
            try:
                with open(os.path.join(ROOT, "profiles", cap_file)) as f:
                    cap = json.load(f)
                traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * len(my) / cap["candidates_per_launch"]
                traffic_src = f"profiles/{cap_file} (ncu --set full, scaled by candidates per launch)"
            except Exception:
                pass
        out["roofline"] = {
            "bound": "fp64-fma (vector pipe; neither hbm nor tensor)", "achieved": achieved, "peak": fp64_fma,
            "unit": "TFLOP/s", "frac": achieved / fp64_fma if fp64_fma else None,
            "frac_of_nominal_40": achieved / 40.0, "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "kg_mc_kernel" if not w["g"] else "kg_mc_gen_kernel", "kernel_share_of_step": mc_ms / dev_ms,
            "peak_source": "measured live: DFMA chain microbenchmark (cmoe_bench_fp64_peaks); MEASURED_PEAKS.json has "
                           "no FP64 entry, nominal B200 FP64 is 40 TFLOP/s",
            "reference_evals_per_sample": stats["posterior_evals"] / max(1, stats["mc_samples"]),
            "point_evals_per_sample": stats["point_evals"] / max(1, stats["mc_samples"]),
            "line_batches_per_sample": stats["line_batches"] / max(1, stats["mc_samples"]),
            "flops_per_point_eval": flops_point, "flops_per_line_batch": flops_line,
            "flops_per_sample_setup": flops_sample,
            "executed_flops_per_sample": executed / max(1, stats["mc_samples"])}
        out["fp64_dmma_peak_tflops"] = fp64_dmma
    else:
        out["roofline"] = {"bound": "issue/RNG (SURVEY 8d: no bandwidth roofline for q-EI)", "achieved": value,
                           "peak": None, "unit": "sample-evals/s", "frac": None, "traffic": None,
                           "kernel": "ei_sample_kernel"}

    if not args.no_extra and world == 1 and args.config == "c3":
        extra = fit_extra(capi, device, fp64_dmma, with_cpu=not args.no_cpu_baseline)
        if kernel == 0:
            # the kernel the reference's Python front end actually builds (gpp_python_gaussian_process.cpp:53)
            try:
                gpm = capi.GaussianProcess(capi.MATERN_NU_2P5, w["alpha"], wl["lengths"], wl["X"], wl["y"], wl["noise"],
                                           device=device)
                bm = float(gpm.posterior(wl["disc"][:, None, :], (), ("mean",))["mean"].min())
                pm = capi.KGPlan(gpm, w["num_mc"], bm, INNER_GD, wl["bounds"], wl["disc"], ncand, w["q"],
                                 seed=SEED_PHILOX, want_grad=True)
                pm.upload(wl["cands"])
                tm = []
                for it in range(3):
                    flush.zero_()
                    torch.cuda.synchronize()
                    pm.run()
                    pm.sync()
                    if it > 0:
                        tm.append(pm.timings()[0])
                extra["c3_matern52"] = {"metric": "q-KG MC sample-evals/sec, Matern-5/2 kernel, same shape",
                                        "value": total_samples / (float(np.mean(tm)) * 1e-3),
                                        "ms_per_step": float(np.mean(tm)), "steps": len(tm)}
                del pm, gpm
            except Exception as e:
                extra["c3_matern52"] = {"value": None, "error": str(e)}
        out["extra"] = extra

    if not args.no_cpu_baseline and world == 1:
        try:
            backend = cpu_backend()
            threads, cores = host_cpus()
            cgp = cpu_gp(backend, w, wl, kernel)
            cbest = best_so_far_cpu(backend, cgp, w, wl)
            sc, smc = cpu_sample_shape(w, threads)
            cands = np.resize(wl["cands"], (sc,) + wl["cands"].shape[1:])
            cpu_step(backend, cgp, w, wl, cands, max(16, smc // 16), threads, cbest)  # warm-up
            units, dt = cpu_step(backend, cgp, w, wl, cands, smc, threads, cbest)
            out["cpu_baseline"] = {"value": units / dt, "unit": "sample-evals/s", "cores": cores, "threads": threads,
                                   "kind": backend.name,
                                   "sample": f"{sc} candidates x {smc} MC samples (value+gradient), {dt:.1f} s"}
        except Exception as e:  # the checker .so did not travel
            out["cpu_baseline"] = {"value": None, "unit": "sample-evals/s", "cores": 0, "kind": "unavailable",
                                   "sample": str(e)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
