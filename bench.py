#!/usr/bin/env python3
"""bench.py — ABneutral fits/sec on the C4 workload (BASELINE.json configs[3]).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    (CPU arm: the oracle port of the reference)

A *step* is one pass of the hot path over one batch of synthetic windows: for every window,
`n_starts` Nelder-Mead fits from random simplices (ab_neutral::run), best-of-starts, then `n_boot`
bootstrap refits (boot_model::run).  One fit = one NM run to termination.  Windows are independent, so
ranks shard them with no data-path collective (weak scaling: every rank processes `--windows` windows).

Printed JSON (one line, rank 0):
  value      fits/s with all inputs already resident in HBM (kernels only, CUDA events on the library stream)
  e2e        fits/s through the host-buffer C-ABI call abfit_alphabeta_batch (= alphabeta::run per window): pinned
             host inputs copied H2D and results D2H inside the timed region, incl. compiling the pedigrees, the host
             draw of the bootstrap vary-vertices that depends on the fit result, and the bootstrap statistics
  roofline   dominant kernel k_fit_starts against the FP64 (DFMA) peak measured in the same run;
             achieved = executed objective evaluations x algorithmic FLOPs per evaluation / kernel time
  cpu_baseline  oracle port of the reference ("literal work": per-pair matrix_power, no early exit) on all
             host cores, bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SEED = 0xAB0B200
GOLDEN_PED = os.path.join(ROOT, "tests", "golden", "pedigree.txt")


# ---------------------------------------------------------------------------------------------
# synthetic C4 windows: the 351-row time structure of data/pedigree.txt,
# D = c + dt(alpha,beta,w) + N(0, 5e-4), clipped at 0
# ---------------------------------------------------------------------------------------------
def load_shape():
    rows = []
    for line in open(GOLDEN_PED).read().split("\n")[1:]:
        if line:
            rows.append([float(x) for x in line.split(" ")[:4]])
    return np.array(rows)


def synth_windows(W, first_window, shape):
    """returns pedigrees [W, N, 4] and p0uu [W]; generator only (numpy), not the product path"""
    rng = np.random.default_rng([SEED, first_window])
    N = len(shape)
    a = 10 ** rng.uniform(-5, -3.3, W)
    b = 10 ** rng.uniform(-4, -2.3, W)
    w = rng.uniform(0, 0.1, W)
    c = rng.uniform(0, 0.005, W)
    p0uu = rng.uniform(0.6, 0.95, W)
    G = np.empty((W, 3, 3))
    G[:, 0, 0] = (1 - a) ** 2; G[:, 0, 1] = 2 * (1 - a) * a; G[:, 0, 2] = a ** 2
    G[:, 1, 0] = 0.25 * (b + 1 - a) ** 2; G[:, 1, 1] = 0.5 * (b + 1 - a) * (a + 1 - b); G[:, 1, 2] = 0.25 * (a + 1 - b) ** 2
    G[:, 2, 0] = b ** 2; G[:, 2, 1] = 2 * (1 - b) * b; G[:, 2, 2] = (1 - b) ** 2
    tmax = int(shape[:, :3].max())
    P = [np.broadcast_to(np.eye(3), (W, 3, 3)).copy()]
    for _ in range(tmax):
        P.append(P[-1] @ G)
    sv0 = np.stack([p0uu, w * (1 - p0uu), (1 - w) * (1 - p0uu)], axis=1)
    tri = shape[:, :3].astype(int)
    uniq, inv = np.unique(tri, axis=0, return_inverse=True)
    dt_u = np.empty((W, len(uniq)))
    for u, (t0, t1, t2) in enumerate(uniq):
        s = np.einsum("wi,wij->wj", sv0, P[t0])
        A, B = P[t1 - t0], P[t2 - t0]
        d = 0.5 * (A[:, :, 0] * B[:, :, 1] + A[:, :, 1] * B[:, :, 0] + A[:, :, 1] * B[:, :, 2] + A[:, :, 2] * B[:, :, 1]) \
            + (A[:, :, 0] * B[:, :, 2] + A[:, :, 2] * B[:, :, 0])
        dt_u[:, u] = np.einsum("wk,wk->w", s, d)
    D = np.maximum(c[:, None] + dt_u[:, inv.ravel()] + rng.normal(0, 5e-4, (W, N)), 0.0)
    peds = np.empty((W, N, 4))
    peds[:, :, :3] = shape[None, :, :3]
    peds[:, :, 3] = D
    return peds, p0uu


# ---------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, all host threads, bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_reference_sample(shape, n_starts_sample, n_boot_sample, literal=True, threads=None):
    """times ab_neutral::run + boot_model::run of ONE synthetic window on the host cores.
    literal=True: the reference's own work (per-pair matrix_power, all max_iters on stalled starts)."""
    from oracle import abref_py as o
    from __graft_entry__ import _load_product

    ab = _load_product()  # host-side input generators only (no device needed)
    o.build()
    threads = threads or o.hw_threads()
    peds, p0uu = synth_windows(1, 0, shape)
    ped, u = peds[0], float(p0uu[0])
    pb = o.Problem(ped, u, u, 1.0)
    sx = ab.gen_start_simplices(SEED, 0, n_starts_sample, float(ped[:, 3].max()))
    flags = 0 if literal else (o.FAST_DIVERGENCE | o.EARLY_EXIT_ON_STALL)
    t0 = time.perf_counter()
    rc, best, allr, pred, resid = o.ab_neutral(pb, sx, max_iters=10000, flags=flags, n_threads=threads)
    idx = ab.gen_resample_idx(SEED, 0, n_boot_sample, len(ped))
    vary = ab.gen_vary_vertices(SEED, 0, n_boot_sample, best["theta"])
    rc2, rows, fits = o.boot_model(pb, best["theta"], pred, resid, idx, vary, max_iters=1000, flags=flags,
                                   n_threads=threads)
    dt = time.perf_counter() - t0
    assert rc == 0 and rc2 == 0
    return (n_starts_sample + n_boot_sample) / dt, dt, threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = load_shape()
    from oracle import abref_py as o

    o.build()
    threads = o.hw_threads()
    # bounded sample: ~ 8 fits per thread and step (literal reference work is ~1.3 s per fit and core)
    ns = max(8, 8 * threads)
    nb = max(2, ns // 10)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, _ = cpu_reference_sample(shape, ns, nb, literal=True, threads=threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = f"1 synthetic C4 window x ({ns} starts + {nb} bootstrap replicates) per step, literal reference work"
    line = {
        "impl": "reference", "metric": "ABneutral fits/sec (window x start x boot, f64)", "value": value,
        "unit": "fits/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, shape),
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of alphabeta-rs (oracle/abref.c), not the Rust binary: no cargo/rustc in this image",
    }
    print(json.dumps(line))


def workload_config(args, shape):
    return {"workload": "C4 synthetic gbM metaprofile (BASELINE.json configs[3])", "windows_per_gpu": args.windows,
            "n_starts": args.starts, "n_boot": args.boots, "pairs_per_window": int(len(shape)),
            "distinct_triples": 10, "tmax": 32, "max_iters_fit": 10000, "max_iters_boot": 1000,
            "fits_per_window": args.starts + args.boots,
            "l2": "inputs (start simplices + resample indices, GBs) are far larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def small_case_c2(ab, ctx, n_starts=1000, n_boot=1000):
    """BASELINE.json configs[0..1] on the GPU: the repo's example pedigree (data/nodelist.txt + edgelist.txt,
    6 pairs, tests/golden/pedigree_generated.txt) fitted from 1000 starts, then 1000 bootstrap replicates —
    one abfit_alphabeta_batch call from host buffers.  Latency-bound by construction (2000 fits), so it is
    reported as wall time, not as a roofline fraction."""
    ped = np.loadtxt(os.path.join(ROOT, "tests", "golden", "pedigree_generated.txt"), skiprows=1)
    p0uu = 0.6554051647850442
    prob = [ab.Problem(ped, p0uu, p0uu, 1.0)]
    sx = ab.gen_start_simplices(SEED, 0, n_starts, float(ped[:, 3].max()))[None]
    idx = ab.gen_resample_idx(SEED, 0, n_boot, len(ped)).ravel()
    ctx.alphabeta_batch(prob, sx, idx, SEED)  # warm
    t0 = time.perf_counter()
    out = ctx.alphabeta_batch(prob, sx, idx, SEED)
    dt = time.perf_counter() - t0
    b = out["best"][0]
    return {"c2": {"workload": "data/nodelist.txt pedigree (6 pairs): 1000 starts + 1000 bootstrap replicates, one call",
                   "wall_ms": 1e3 * dt, "fits_per_s": (n_starts + n_boot) / dt,
                   "best": {"alpha": float(b["theta"][0]), "beta": float(b["theta"][1]), "lse": float(b["lse"])}}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=10000, help="windows per GPU and step")
    ap.add_argument("--starts", type=int, default=1000)
    ap.add_argument("--boots", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import _load_product

    ab = _load_product()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    assert world == args.gpus or world == 1

    W, NS, NB = args.windows, args.starts, args.boots
    shape = load_shape()
    N = len(shape)
    from alphabeta_rs_b200 import multi

    first, _cnt = multi.window_shard(world * W, rank, world)  # every rank fits its own windows (weak scaling)
    assert _cnt == W
    ctx = ab.Context(local)
    fp64_peak = ctx.measure_fp64_peak()  # TFLOP/s, DFMA = 2 FLOP

    # ---- inputs: host-generated from seeds (north_star), staged in PINNED host memory ----------
    peds, p0uu = synth_windows(W, first, shape)
    probs = [ab.Problem(peds[i], float(p0uu[i]), float(p0uu[i]), 1.0) for i in range(W)]
    sx_t = torch.empty((W, NS, 5, 4), dtype=torch.float64, pin_memory=True)
    idx_t = torch.empty((W, NB, N), dtype=torch.int32, pin_memory=True)
    vary_t = torch.empty((W, NB, 4, 4), dtype=torch.float64, pin_memory=True)
    sx, idx, vary = sx_t.numpy(), idx_t.numpy(), vary_t.numpy()
    pool = ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 8))
    maxd = peds[:, :, 3].max(axis=1)

    def gen_w(i):
        sx[i] = ab.gen_start_simplices(SEED, first + i, NS, float(maxd[i]))
        idx[i] = ab.gen_resample_idx(SEED, first + i, NB, N)

    list(pool.map(gen_w, range(W)))

    def gen_vary(best_theta):
        def one(i):
            vary[i] = ab.gen_vary_vertices(SEED, first + i, NB, best_theta[i])
        list(pool.map(one, range(W)))

    fits_per_step = W * (NS + NB)
    batch = ctx.batch(probs)
    fl = batch.flops_per_eval(0)

    # ---- resident run: everything in HBM before the timed region ------------------------------
    batch.upload_starts(sx)
    batch.run_fit()
    res = batch.download_fit()
    gen_vary(res.best["theta"])
    batch.upload_boot(idx, vary)
    batch.run_boot()
    batch.sync()

    def resident_step():
        batch.run_fit()
        batch.run_boot()

    for _ in range(max(0, args.warmup - 1)):
        resident_step()
    ctx.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ctx.timer_start()
    fit_ms = sel_ms = boot_ms = 0.0
    evals_fit = evals_boot = 0
    launches = 0
    for _ in range(args.steps):
        resident_step()
        # per-kernel CUDA events recorded by the library on its own stream (no host sync inside the kernels'
        # critical path: timing() waits for the events of this step only)
        t = batch.timing()
        fit_ms += t["fit_ms"]; sel_ms += t["select_ms"]; boot_ms += t["boot_ms"]
        evals_fit += t["evals_fit"]; evals_boot += t["evals_boot"]; launches += t["launches"]
    total_ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    t_max = total_ms
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max = float(tt.item())
    value = world * fits_per_step * args.steps / (t_max * 1e-3)

    # ---- end-to-end: host buffers through the one-shot C-ABI calls ------------------------------
    best_t = torch.zeros((W, 64), dtype=torch.uint8, pin_memory=True)
    best_h = best_t.numpy().view(ab.FIT_DTYPE).reshape(W)
    pred_t = torch.empty(W * N, dtype=torch.float64, pin_memory=True)
    resid_t = torch.empty(W * N, dtype=torch.float64, pin_memory=True)
    rows_t = torch.empty((W, NB, 7), dtype=torch.float64, pin_memory=True)
    status_h = np.zeros(W, dtype=np.int32)

    analysis_h = np.empty((W, 32))
    packed = ab._pack_problems(probs)

    def e2e_step():
        # abfit_alphabeta_batch = alphabeta::run for every window: fit, host draw of the vary vertices from the
        # best-of-starts, bootstrap, statistics; host buffers in, host buffers out
        ctx.alphabeta_batch(probs, sx, idx, SEED, first_problem_id=first, best=best_h, pred=pred_t.numpy(),
                            resid=resid_t.numpy(), status=status_h, rows=rows_t.numpy(), analysis=analysis_h, packed=packed)

    h2d = sx.nbytes + idx.nbytes + vary.nbytes + peds.nbytes
    d2h = best_h.nbytes + 2 * pred_t.numpy().nbytes + rows_t.numpy().nbytes + status_h.nbytes
    e2e_step()  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * fits_per_step * args.steps / e2e_s

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    achieved = evals_fit * fl["flops"] / (fit_ms * 1e-3) / 1e12  # TFLOP/s, algorithmic
    # DRAM bytes of one k_fit_starts launch at the default configuration, from an ncu capture of this command
    # (profiles/r01_dram_traffic_bench_default.csv: 1.784 GB read + 0.780 GB written = the start simplices in and the
    # fit records out; the kernel never re-reads HBM).  Other configurations were not captured.
    traffic = 1784302336 + 779979008 if (W, NS, NB) == (10000, 1000, 100) else None
    roofline = {"bound": "fp64", "kernel": "k_fit_starts", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak, "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, ncu)",
                "bound_note": "FP64 vector pipe (BASELINE.json: % FP64 peak); no tensor cores, HBM traffic is 2.6 GB per 2.3 s launch",
                "peak_source": "DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 figure; nominal 37.2)",
                "flops_per_eval": fl["flops"], "evals_per_launch": evals_fit / args.steps,
                "kernel_ms": fit_ms / args.steps, "kernel_share_of_step": fit_ms / total_ms,
                "boot_kernel_ms": boot_ms / args.steps,
                "boot_achieved": evals_boot * fl["flops"] / (boot_ms * 1e-3) / 1e12 if boot_ms > 0 else None}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import abref_py as o

        threads = o.hw_threads()
        ns = max(8, 8 * threads)
        nb = max(2, ns // 10)
        v, dt, _ = cpu_reference_sample(shape, ns, nb, literal=True, threads=threads)
        v2, dt2, _ = cpu_reference_sample(shape, 16 * ns, 16 * nb, literal=False, threads=threads)
        cpu = {"value": v, "unit": "fits/s", "cores": threads, "kind": "port",
               "sample": f"1 C4 window x ({ns} starts + {nb} boots), literal reference work, {dt:.1f} s",
               "minimal_work_value": v2,
               "minimal_work_note": "same port with the power table + stall early-exit the GPU path uses "
                                    f"({16 * ns} starts + {16 * nb} boots, {dt2:.1f} s): separates algorithmic "
                                    "from hardware speed-up"}

    # ---- other BASELINE configs, reported beside the headline (rank 0, outside every timed region) -------------
    aux = None
    if rank == 0:
        try:
            aux = small_case_c2(ab, ctx)
        except Exception as e:  # never let a side measurement break the contract line
            aux = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "ABneutral fits/sec (window x start x boot, f64)", "value": value, "unit": "fits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, shape),
            "e2e": {"value": e2e_value, "unit": "fits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "evals_per_fit": {"starts": evals_fit / (args.steps * W * NS), "boots": evals_boot / (args.steps * W * NB)},
            "other_configs": aux,
        }
        print(json.dumps(line))
    batch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
