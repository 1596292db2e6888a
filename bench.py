#!/usr/bin/env python
"""bench.py — q-KG Monte-Carlo sample-evaluations per second at the north-star shape (BASELINE.json).

A "step" = one batched q-KG value+gradient evaluation (what one outer optimiser step / the multistart pre-screen
consumes) of `multistart` candidates x `num_mc` samples at N=500, d=8, q=8 on synthetic data.  With --gpus N > 1 the
candidates are sharded over ranks (one process per GPU, launched by torchrun) with no data-path collective; the only
exchange is one all-gather of the per-candidate values for the global arg-max (strong scaling: total work fixed).

Output: ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM before the timed region, CUDA
events on the launching stream, max over ranks); `e2e` is the same metric through the host-buffer C-ABI call
(plan creation + H2D + kernels + D2H inside the timed region).  `roofline` describes the dominant kernel
(kg_mc_kernel, bound by the FP64 vector pipe — it is neither HBM- nor tensor-bound); `cpu_baseline` is
the reference's own C++ path (oracle/_ref) timed on this host's cores on a bounded sample.
--impl reference times the CPU reference alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(N=500, dim=8, q=8, num_mc=16384, multistart=1024, num_pts=10, noise=1e-2, length=0.5, alpha=1.0)
INNER_GD = [1, 6, 1, 3, 0.0, 1.0, 0.1, 1e-10]  # the reference examples' inner optimiser (examples/main.py:123-130)
SEED_PHILOX = 0xC0FFEE
METRIC = "q-KG MC sample-evals/sec at (N=500,d=8,q=8,mc=16384)"


def make_workload(w=WORKLOAD):
    rng = np.random.default_rng(20260924)
    X = rng.uniform(size=(w["N"], w["dim"]))
    y = np.sin(3.0 * X).sum(axis=1) + np.sqrt(w["noise"]) * rng.standard_normal(w["N"])
    cands = np.random.default_rng(7).uniform(size=(w["multistart"], w["q"], w["dim"]))
    disc = np.random.default_rng(11).uniform(size=(w["num_pts"], w["dim"]))
    return dict(X=X, y=y, lengths=np.full(w["dim"], w["length"]), noise=np.array([w["noise"]]), cands=cands,
                disc=disc, bounds=np.tile([0.0, 1.0], w["dim"]))


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
        # samples under load: keep the upper half (the idle samples before/after the region read low)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_step(backend_name_only=False):
    import oracle as orc
    backend = orc.load_reference() if orc.have_reference() else orc.load_oracle()
    return backend


def run_cpu(backend, wl, cands, num_mc, threads, best, want_grad=True):
    import oracle as orc
    gp, lm = backend.gp(0, WORKLOAD["alpha"], wl["lengths"], wl["X"], wl["y"], wl["noise"])
    assert lm == 0
    t0 = time.perf_counter()
    orc.kg_grad_at_point_list(backend, gp, cands, None, num_mc, best, INNER_GD, wl["bounds"], wl["disc"], threads,
                              want_grad=want_grad)
    dt = time.perf_counter() - t0
    return cands.shape[0] * num_mc / dt, dt


def best_so_far_from_cpu(backend, wl):
    gp, lm = backend.gp(0, WORKLOAD["alpha"], wl["lengths"], wl["X"], wl["y"], wl["noise"])
    return float(gp.mean_additional(wl["disc"]).min())  # py/cpp_wrappers/knowledge_gradient.py:361-368


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--multistart", type=int, default=WORKLOAD["multistart"])
    ap.add_argument("--num-mc", type=int, default=WORKLOAD["num_mc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = dict(WORKLOAD, multistart=args.multistart, num_mc=args.num_mc)
    wl = make_workload(w)
    config = {"workload": f"q-KG value+gradient, synthetic SE GP, N={w['N']}, d={w['dim']}, q={w['q']}, "
                          f"num_mc={w['num_mc']}, multistart={w['multistart']}, discrete_pts={w['num_pts']}, "
                          f"inner GD steps=6 restarts=1 (BASELINE.json configs[2])",
              "sharding": f"candidates strided over {world} rank(s)", "l2": "flushed between timed steps (256 MiB memset)"}

    if args.impl == "reference":
        if rank != 0:
            return
        backend = cpu_reference_step()
        best = best_so_far_from_cpu(backend, wl)
        # give the reference its best thread count: all hardware threads or one per physical core (SMT often hurts
        # this FP64-heavy loop); calibrated on a short run, which doubles as the warm-up
        tmax = backend.max_threads()
        trial = {}
        for tt in sorted({tmax, max(1, tmax // 2)}):
            cc = np.resize(wl["cands"], (tt,) + wl["cands"].shape[1:])
            trial[tt] = run_cpu(backend, wl, cc, 64, tt, best)[0]
        threads = max(trial, key=trial.get)
        sample_c, sample_mc = threads, 256
        cands = wl["cands"][:sample_c] if sample_c <= len(wl["cands"]) else np.resize(wl["cands"], (sample_c,) + wl["cands"].shape[1:])
        vals = [run_cpu(backend, wl, cands, sample_mc, threads, best) for _ in range(args.steps)]
        total_t = sum(v[1] for v in vals)
        value = sample_c * sample_mc * args.steps / total_t
        sample = f"{sample_c} candidates x {sample_mc} MC samples per step (value+gradient), OpenMP static over candidates"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": "sample-evals/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": value, "unit": "sample-evals/s", "cores": threads,
                                           "kind": backend.name, "sample": sample},
                          "e2e": {"value": value, "unit": "sample-evals/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import torch
    from cornell_moe_b200 import capi
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank if world > 1 else 0
    assert capi.device_count() > device, "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(device)

    gp = capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, w["alpha"], wl["lengths"], wl["X"], wl["y"], wl["noise"],
                              device=device)
    # best_so_far = min posterior mean over the discrete set, as the reference's Python wrapper computes it
    best = float(gp.posterior(wl["disc"][:, None, :], (), ("mean",))["mean"].min())
    my = wl["cands"][rank::world]
    plan = capi.KGPlan(gp, w["num_mc"], best, INNER_GD, wl["bounds"], wl["disc"], len(my), w["q"], seed=SEED_PHILOX,
                       want_grad=True)
    plan.upload(my)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{device}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step():
        flush.zero_()
        torch.cuda.synchronize()
        plan.run()
        plan.sync()
        return plan.timings()

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(device)
    barrier()
    if rank == 0:
        sampler.start()
    t_host0 = time.perf_counter()
    dev_ms, mc_ms, launches = 0.0, 0.0, 0
    for _ in range(args.steps):
        tot, mc, nl = one_step()
        dev_ms += tot
        mc_ms += mc
        launches += nl
    barrier()
    t_host = time.perf_counter() - t_host0
    clocks = sampler.stop() if rank == 0 else None
    kg, grad, stats = plan.download()

    # end-to-end through the host-buffer API (plan creation, H2D, kernels, D2H inside the timed region)
    del plan  # free the device-resident plan's workspace before the host-API runs allocate theirs
    e2e_steps = []
    for it in range(1 + args.steps):  # first call = warm-up (allocates the handle's cached workspace)
        barrier()
        t0 = time.perf_counter()
        kg_e, grad_e = gp.kg(my, None, w["num_mc"], best, INNER_GD, wl["bounds"], wl["disc"], seed=SEED_PHILOX, grad=True)
        barrier()
        if it > 0:
            e2e_steps.append(time.perf_counter() - t0)
    e2e_s = float(np.sum(e2e_steps))
    assert np.array_equal(kg_e, kg), "host-API result differs from the device-resident plan"


    times = torch.tensor([dev_ms, mc_ms, e2e_s, t_host], dtype=torch.float64, device=f"cuda:{device}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        # the one exchange of the path: all-gather per-candidate values -> identical global arg-max on every rank
        loc = torch.full((int(np.ceil(len(wl["cands"]) / world)),), float("-inf"), dtype=torch.float64, device=f"cuda:{device}")
        loc[: len(kg)] = torch.from_numpy(kg).to(loc.device)
        allv = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(allv, loc)
        full = torch.stack(allv, dim=1).reshape(-1)[: len(wl["cands"])].cpu().numpy()  # index = local*world + rank
    else:
        full = kg
    argmax = int(np.argmax(full))
    dev_ms, mc_ms, e2e_s, t_host = [float(x) for x in times.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_samples = len(wl["cands"]) * w["num_mc"]
    ms_per_step = dev_ms / args.steps
    value = total_samples / (ms_per_step * 1e-3)
    e2e_value = total_samples / (e2e_s / args.steps)
    # roofline of the dominant kernel: FP64 flops of the work it EXECUTED (from its own counters) / its time.
    # Conventions of SURVEY.md 8(d): FMA = 2 flops, exp = 1 flop.  Per training/union row:
    #   point evaluation (value + gradient at one query point): dot 2d, weight a_j 2q, exp 1, value 2, scale 1, gradient 2d
    #   line batch (all KB = 8 backtracking trials of a step):  two dots 4d, weight 2q, two exps 2, scale 1,
    #                                                            KB fma 2KB, KB-1 squarings
    rows = w["N"] + w["q"]
    KB = 8
    flops_point = rows * (2 * w["dim"] + 2 * w["q"] + 1 + 2 + 1 + 2 * w["dim"])
    flops_line = rows * (4 * w["dim"] + 2 * w["q"] + 2 + 1 + 2 * KB + (KB - 1))
    fp64_fma, fp64_dmma = capi.fp64_peaks(device)
    executed = stats["point_evals"] * flops_point + stats["line_batches"] * flops_line
    achieved = executed / (mc_ms / args.steps * 1e-3) * 1e-12
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (dram__bytes_read + write); the capture
    # ran 256 candidates per launch and the traffic is per-sample records, so it scales with the candidates of a launch
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_kg_mc_kernel_ncu.json")) as f:
            cap = json.load(f)
        traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * len(my) / cap["candidates_per_launch"]
        traffic_src = "profiles/r1_kg_mc_kernel_ncu.json (ncu --set full, scaled by candidates per launch)"
    except Exception:
        pass
    roofline = {"bound": "fp64-fma (vector pipe; neither hbm nor tensor)", "achieved": achieved, "peak": fp64_fma,
                "unit": "TFLOP/s", "frac": achieved / fp64_fma if fp64_fma else None, "traffic": traffic,
                "traffic_source": traffic_src, "kernel": "kg_mc_kernel", "kernel_share_of_step": mc_ms / dev_ms,
                "peak_source": "measured live: DFMA chain microbenchmark (cmoe_bench_fp64_peaks)",
                "reference_evals_per_sample": stats["posterior_evals"] / max(1, stats["mc_samples"]),
                "point_evals_per_sample": stats["point_evals"] / max(1, stats["mc_samples"]),
                "line_batches_per_sample": stats["line_batches"] / max(1, stats["mc_samples"]),
                "flops_per_point_eval": flops_point, "flops_per_line_batch": flops_line,
                "executed_flops_per_sample": executed / max(1, stats["mc_samples"])}
    out = {"metric": METRIC, "value": value, "unit": "sample-evals/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "e2e": {"value": e2e_value, "unit": "sample-evals/s",
                   "h2d_bytes_per_step": int(wl["cands"].nbytes), "d2h_bytes_per_step": int(kg.nbytes * world + (grad.nbytes if grad is not None else 0) * world)},
           "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
           "host_wall_ms_per_step": 1e3 * t_host / args.steps, "argmax_index": argmax,
           "kg_checksum": float(np.sum(full)), "fp64_dmma_peak_tflops": fp64_dmma}

    if not args.no_extra:
        # secondary kernels of the path (config 5 shape): covariance build vs HBM, blocked Cholesky vs DMMA peak
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm = peaks.get("hbm_gbs", 6650.0)
            src = "MEASURED_PEAKS.json"
        except Exception:
            hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
        Nl, dl = 5000, 10
        rng = np.random.default_rng(5)
        Xl = rng.uniform(size=(Nl, dl))
        yl = np.sin(3 * Xl).sum(axis=1) + 0.1 * rng.standard_normal(Nl)
        gpl = capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, 1.0, np.full(dl, 0.5), Xl, yl, [1e-2], device=device)
        del gpl  # first fit of the process = kernel loading; the reported fit is the second one
        gpl = capi.GaussianProcess(capi.SQUARE_EXPONENTIAL, 1.0, np.full(dl, 0.5), Xl, yl, [1e-2], device=device)
        t_cov = gpl.bench_cov_build(20)
        t_chol = gpl.bench_cholesky(3)
        cov_bytes = 4.0 * Nl * (Nl + 1) + 8.0 * Nl * dl
        out["extra"] = {
            "cov_build_N5000_d10": {"bound": "hbm", "usec": t_cov, "achieved": cov_bytes / t_cov * 1e-3, "peak": hbm,
                                    "unit": "GB/s", "frac": cov_bytes / t_cov * 1e-3 / hbm, "peak_source": src},
            "cholesky_N5000": {"bound": "tensor (FP64 DMMA)", "usec": t_chol, "achieved": Nl ** 3 / 3.0 / t_chol * 1e-6,
                               "peak": fp64_dmma, "unit": "TFLOP/s", "frac": Nl ** 3 / 3.0 / t_chol * 1e-6 / fp64_dmma,
                               "peak_source": "measured live: DMMA m8n8k4 microbenchmark"},
            "gp_fit_N5000_usec_cov_chol_solve": [float(x) for x in gpl.fit_timings_usec()]}

    if not args.no_cpu_baseline and world == 1:
        try:
            backend = cpu_reference_step()
            threads = backend.max_threads()
            sc, smc = threads, 256
            cands = wl["cands"][:sc] if sc <= len(wl["cands"]) else np.resize(wl["cands"], (sc,) + wl["cands"].shape[1:])
            v, dt = run_cpu(backend, wl, cands, smc, threads, best)
            out["cpu_baseline"] = {"value": v, "unit": "sample-evals/s", "cores": threads, "kind": backend.name,
                                   "sample": f"{sc} candidates x {smc} MC samples (value+gradient), {dt:.1f} s"}
        except Exception as e:  # the checker .so did not travel
            out["cpu_baseline"] = {"value": None, "unit": "sample-evals/s", "cores": 0, "kind": "unavailable",
                                   "sample": str(e)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
