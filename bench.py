#!/usr/bin/env python
"""bench.py -- PDHG iterations/s on ROF-TV 4096^2 (BASELINE.json metric) with roofline evidence.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one PDHG iteration (BackendPDHG::PerformIteration, backend_pdhg.cu:311-381) on the
metric config of SURVEY.md section 8(d): ROF 4096 x 4096 gray, BlockGradient2D + 1d:square +
norm2:ind_leq0, alpha = 1 preconditioning, Alg1 steps, residuals every 10 iterations.
Rank 0 prints ONE JSON line (contract in the task statement):

  value      iterations/s with all inputs resident in HBM, CUDA-event timed over K iterations
  e2e        iterations/s of a whole solve through the public Solver API with pinned HOST
             buffers: problem upload (H2D), K iterations, solution download (D2H)
  roofline   dominant kernel (fused dual pass): algorithmic bytes / event-timed duration vs the
             measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the OpenMP oracle port of the same iteration on the host cores (bounded sample)

--impl reference times the reference algorithm's CPU restatement (prost has no CPU path of its
own; oracle/prost_oracle.cpp, kind "port") on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = NY = 4096
LAM = 10.0
RESIDUAL_ITER = 10
# primal pass reads y (2N) + x (N) + f (N), writes x+ (N); dual pass reads x+, x (2N),
# y (2N), writes y+ (2N): 11 floats = 44 B per pixel per iteration, 20 B primal + 24 B dual.
BYTES_PER_PX_ITER = 44
BYTES_PER_PX_PRIMAL = 20
BYTES_PER_PX_DUAL = 24
# one-pass tiled iteration (pb_tile.cu): read y (2N), x (N), f (N); write x+ (N), y+ (2N) = 7 floats
BYTES_PER_PX_TILE = 28
# the same pass on residual-refresh iterations also reads the previous dual iterate (2N): 9 floats
BYTES_PER_PX_TILE_CHECK = 36


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(pb, syn, ctx, f):
    desc = syn.rof(NX, NY, LAM, f=f)
    prob = pb.create_problem(ctx, desc)
    return prob


def time_to_residual(pb, ctx, make_problem, comm=None, max_iters=20000):
    """BASELINE metric, third part: wall time of Solver.Solve until r_p < eps_p and r_d < eps_d with all four
    tolerances 1e-4 (backend.hpp:71-74), Alg2 with gamma = 0.05 lambda like the reference's ROF example
    (matlab/examples/example_rof_primaldual.m:36-38).  Includes the copy-back of the solution."""
    popts = pb.pdhg_options(scale_steps_operator=0, stepsize="alg2", alg2_gamma=0.05 * LAM,
                            residual_iter=RESIDUAL_ITER)
    sopts = pb.solver_options(verbose=0, max_iters=max_iters, tol_rel_primal=1e-4, tol_rel_dual=1e-4,
                              tol_abs_primal=1e-4, tol_abs_dual=1e-4, num_cback_calls=0)
    prob = make_problem()
    be = pb.BackendPDHG(ctx, prob, popts, sopts, comm=comm) if comm else pb.BackendPDHG(ctx, prob, popts, sopts)
    solver = pb.Solver(prob, be)
    solver.SetOptions(sopts)
    solver.Initialize()
    ctx.synchronize()
    t0 = time.perf_counter()
    result = solver.Solve()
    ctx.synchronize()
    dt = time.perf_counter() - t0
    res = be.residuals()
    return {"seconds": dt, "iterations": int(solver.iterations), "converged": result == pb.Solver.CONVERGED,
            "stepsize": "alg2, gamma = 0.05 lambda", "tolerances": 1e-4, "max_iters": max_iters,
            "primal_residual": res["primal_residual"], "dual_residual": res["dual_residual"],
            "eps_primal": res["eps_primal"], "eps_dual": res["eps_dual"],
            "what": "Solver.Solve wall time (residual check every 10 iterations) incl. D2H of x, z, y, w"}


def run_reference_arm(args, rank, world):
    """Reference arm for this tier: the CPU restatement of prost's PDHG iteration on host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import OracleProblem, OraclePDHG, num_threads
    from prost_b200 import synthetic as syn
    threads = num_threads()
    # bounded sample: a slab of columns sized so that K+W iterations take about two minutes;
    # the iteration is bandwidth-bound on the host, so iterations/s scales with 1/columns.
    probe_cols = 256
    desc = syn.rof(probe_cols, NY, LAM)
    o = OraclePDHG(OracleProblem(desc), stepsize="alg1", residual_iter=RESIDUAL_ITER,
                   tol_rel_primal=0, tol_rel_dual=0, tol_abs_primal=0, tol_abs_dual=0)
    o.initialize()
    o.iterate(2)
    t0 = time.perf_counter()
    o.iterate(3)
    t_iter_probe = (time.perf_counter() - t0) / 3
    budget = 120.0
    cols = int(min(NX, max(64, probe_cols * budget / max(t_iter_probe * (args.steps + args.warmup), 1e-9))))
    cols = max(64, (cols // 64) * 64)
    if cols != probe_cols:
        desc = syn.rof(cols, NY, LAM)
        o = OraclePDHG(OracleProblem(desc), stepsize="alg1", residual_iter=RESIDUAL_ITER,
                       tol_rel_primal=0, tol_rel_dual=0, tol_abs_primal=0, tol_abs_dual=0)
        o.initialize()
    o.iterate(args.warmup)
    t0 = time.perf_counter()
    o.iterate(args.steps)
    dt = time.perf_counter() - t0
    frac = cols / NX
    value = args.steps / dt * frac            # iterations/s of the FULL 4096^2 image
    sample = f"{args.steps} iterations on a {cols}x{NY} column slab ({frac:.4f} of the image), scaled by area"
    line = {
        "impl": "reference", "metric": "pdhg_iterations_per_second", "value": value, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps / frac * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, scaling="strong"):
    cols = NX * n_gpus if scaling == "weak" else NX
    per_gpu_mb = 28 * (cols // n_gpus) * NY / 1e6
    return {"workload": f"ROF-TV {cols}x{NY} gray, PDHG Alg1, BlockGradient2D + 1d:square(lambda={LAM:g}) + "
                        f"norm2:ind_leq0, alpha=1 preconditioning, residual_iter={RESIDUAL_ITER}",
            "n_pixels": cols * NY, "bytes_per_iteration_algorithmic": BYTES_PER_PX_ITER * cols * NY,
            "cache": (f"per-GPU state (x,x_prev,y,y_prev,f) = {per_gpu_mb:.0f} MB; "
                      + ("exceeds the 126 MB L2, no flush needed" if per_gpu_mb > 126 else
                         "FITS the 126 MB L2: the strong-scaling slabs run L2-resident by construction "
                         "(that is the workload, not a cached repeat: every iteration reads the previous "
                         "iteration's output)")),
            "parallelism": (f"slab{n_gpus}: column slabs, one-pass ring kernel per slab with peer-to-peer halo "
                            f"stores over NVLink from its edge tiles, "
                            f"NCCL all-reduce of 4 residual sums every {RESIDUAL_ITER} iterations")
            if n_gpus > 1 else "single"}


def run_slab_arm(args, rank, local_rank, world):
    """N > 1: one process per GPU, each owning a block of image columns."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import prost_b200 as pb
    from prost_b200 import synthetic as syn
    from prost_b200 import distributed as pbd

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pb.Context(local_rank, stream.cuda_stream)
    comm = pbd.init_comm(ctx)

    weak = args.scaling == "weak"
    cols_total = NX * world if weak else NX
    part = pbd.SlabPartition(cols_total, world, align=4)
    x0, x1 = part.range(rank)
    w = x1 - x0
    f = syn.image(cols_total, NY, x0=x0, x1=x1)        # this rank's columns only (counter-based RNG)
    n, m = w * NY, 2 * w * NY
    norm = cols_total / NX                               # iterations of a 4096^2 image per iteration

    def max_over_ranks(v):
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    popts = pb.pdhg_options(scale_steps_operator=0, stepsize="alg1", residual_iter=RESIDUAL_ITER)
    sopts = pb.solver_options(verbose=0, max_iters=args.steps, tol_rel_primal=0, tol_rel_dual=0,
                              tol_abs_primal=0, tol_abs_dual=0, num_cback_calls=0)
    prob = pb.create_problem(ctx, syn.rof(w, NY, LAM, f=f))
    prob.Initialize()
    be = pb.BackendPDHG(ctx, prob, popts, sopts, comm=comm)
    be.Initialize()
    assert be.is_fused
    be.PerformIteration(args.warmup)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = be.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    be.PerformIteration(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = be.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps / (ms * 1e-3) * norm

    prof_iters = min(200, max(20, args.steps // 10))
    d = be.profile_detail(prof_iters)
    peak, peak_src = measured_peak()
    wmax = max(part.width(r) for r in range(world))
    one_pass = d["n_tile"] > 0
    fits_l2 = 28 * wmax * NY / 1e6 < 126
    l2_note = ("per-GPU slab state fits the 126 MB L2 at this N, so the algorithmic GB/s can exceed the "
               "HBM copy peak" if fits_l2 else "slab exceeds L2")
    if one_pass:
        # every iteration but the first is ONE launch of the ring kernel on the slab (pb_tile.cu, SLAB):
        # the left-/right-edge tiles store their edge columns into the neighbours' memory over NVLink
        t_tile, t_chk = max_over_ranks(d["tile_ms"]), max_over_ranks(d["tile_check_ms"])
        tile_bytes = BYTES_PER_PX_TILE * wmax * NY
        n_all = max(d["n_tile"] + d["n_two_pass"] + d["n_tile_check"], 1.0)
        avg_px = (d["n_tile"] * BYTES_PER_PX_TILE + d["n_tile_check"] * BYTES_PER_PX_TILE_CHECK +
                  d["n_two_pass"] * BYTES_PER_PX_ITER) / n_all
        ach = tile_bytes / (t_tile * 1e-3) / 1e9
        ach_iter = avg_px * wmax * NY * (args.steps / (ms * 1e-3)) / 1e9
        roofline = {
            "bound": "hbm", "kernel": "grad2d_iteration_ring_kernel<SQUARE, IND_LEQ0, CHECK=false, SLAB=true> (whole "
                                      "PDHG iteration in one pass, per GPU on its slab)",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": tile_bytes, "ms_per_launch": t_tile,
            "launch_share": d["n_tile"] / n_all,
            "residual_refresh_iterations": {"share": d["n_tile_check"] / n_all, "ms_per_launch": t_chk,
                                            "algorithmic_bytes_per_launch": BYTES_PER_PX_TILE_CHECK * wmax * NY},
            "whole_iteration_per_gpu": {"achieved": ach_iter, "frac": ach_iter / peak},
            "note": l2_note + "; ms_per_launch is measured with a host synchronisation after every iteration "
                              "(max over ranks), so it includes the launch skew between ranks that the "
                              "back-to-back timed region hides",
        }
    else:
        t_primal, t_dual = max_over_ranks(d["primal_ms"]), max_over_ranks(d["dual_ms"])
        dual_bytes = BYTES_PER_PX_DUAL * wmax * NY
        primal_bytes = BYTES_PER_PX_PRIMAL * wmax * NY
        ach_dual = dual_bytes / (t_dual * 1e-3) / 1e9
        ach_primal = primal_bytes / (t_primal * 1e-3) / 1e9
        ach_iter = BYTES_PER_PX_ITER * wmax * NY * (args.steps / (ms * 1e-3)) / 1e9
        roofline = {
            "bound": "hbm", "kernel": "grad_dual_norm2_kernel (fused dual pass), per GPU on its slab",
            "achieved": ach_dual, "peak": peak, "unit": "GB/s", "frac": ach_dual / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": dual_bytes, "ms_per_launch": t_dual,
            "primal_pass": {"achieved": ach_primal, "frac": ach_primal / peak,
                            "algorithmic_bytes_per_launch": primal_bytes, "ms_per_launch": t_primal},
            "whole_iteration_per_gpu": {"achieved": ach_iter, "frac": ach_iter / peak},
            "note": l2_note,
        }
    res = be.residuals()
    p2p = comm.peer_to_peer
    del be

    # end to end through the public Solver API with pinned host buffers, collectively on all ranks
    f_pin = torch.from_numpy(f).pin_memory()
    x0_pin = torch.zeros(n, dtype=torch.float32).pin_memory()
    y0_pin = torch.zeros(m, dtype=torch.float32).pin_memory()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prob2 = pb.create_problem(ctx, syn.rof(w, NY, LAM, f=f_pin.numpy()))
    be2 = pb.BackendPDHG(ctx, prob2, popts, sopts, comm=comm)
    solver = pb.Solver(prob2, be2)
    solver.SetOptions(sopts, x0=x0_pin.numpy(), y0=y0_pin.numpy())
    solver.Initialize()
    solver.Solve()
    ctx.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    h2d = 4 * (n + n + m) + 4 * (n + m)
    d2h = 4 * (2 * n + 2 * m)
    e2e = {"value": args.steps / t_e2e * norm, "unit": "iter/s", "h2d_bytes_per_step": h2d * world / args.steps,
           "d2h_bytes_per_step": d2h * world / args.steps, "seconds_total": t_e2e,
           "what": "per rank: Problem build + Solver.Initialize + Solver.Solve(max_iters=K) + solution copy-back; "
                   "max over ranks"}
    del be2, solver
    ttr = time_to_residual(pb, ctx, lambda: pb.create_problem(ctx, syn.rof(w, NY, LAM, f=f_pin.numpy())), comm=comm)
    if rank == 0:
        line = {
            "metric": "pdhg_iterations_per_second", "value": value, "unit": "iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world, args.scaling), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
            "halo_mode": "peer-to-peer stores over NVLink (CUDA IPC)" if p2p else "NCCL send/recv staging",
            "one_pass_iterations": bool(one_pass), "time_to_residual_1e-4": ttr,
            "residuals_after_run": res,
        }
        print(json.dumps(line), flush=True)
    comm.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N>1: strong = the 4096^2 image split into N column slabs (the BASELINE metric); "
                         "weak = 4096 columns per GPU (a 4096N x 4096 image), value normalised to 4096^2 iterations")
    ap.add_argument("--nx", type=int, default=NX, help="image columns (default 4096 = the BASELINE metric config; "
                                                       "other values are for scaling experiments only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    globals()["NX"] = args.nx

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import prost_b200 as pb
    from prost_b200 import synthetic as syn

    torch.cuda.set_device(local_rank)
    if world > 1:
        run_slab_arm(args, rank, local_rank, world)
        return

    # a dedicated (non-default) stream: the context launches on it and the CUDA events below are
    # recorded on it, so the events bracket exactly the kernels of the timed region
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = pb.Context(local_rank, stream.cuda_stream)

    f = syn.image(NX, NY)
    n, m = NX * NY, 2 * NX * NY

    # ---------------- device-resident throughput (value) ----------------------------------------
    prob = build_problem(pb, syn, ctx, f)
    prob.Initialize()
    popts = pb.pdhg_options(scale_steps_operator=0, stepsize="alg1", residual_iter=RESIDUAL_ITER)
    # num_cback_calls = 0: no intermediate callbacks, i.e. the solution is copied back once at the end
    # (each intermediate callback of Solver::Solve is a full D2H copy of x, z, y, w: solver.cu:152-167)
    sopts = pb.solver_options(verbose=0, max_iters=args.steps, tol_rel_primal=0, tol_rel_dual=0,
                              tol_abs_primal=0, tol_abs_dual=0, num_cback_calls=0)
    be = pb.BackendPDHG(ctx, prob, popts, sopts)
    be.Initialize()
    assert be.is_fused, "fused PDHG path not selected"
    be.PerformIteration(args.warmup)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = be.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    be.PerformIteration(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = be.launch_count - launches0
    clocks = sampler.stop()
    value = args.steps / (ms * 1e-3)

    # ---------------- per-kernel timing for the roofline (events around each kernel) ----------------
    # Iterations that refresh the residuals (1 in RESIDUAL_ITER) run as two passes (44 B/px + the
    # previous dual iterate), all others as ONE tiled kernel that moves 28 B/px (pb_tile.cu): read
    # y (2 floats), x, f; write x+, y+ (2 floats).  The dominant kernel is the tiled one.
    prof_iters = min(400, max(40, args.steps // 5))
    d = be.profile_detail(prof_iters)
    peak, peak_src = measured_peak()
    tiled = d["n_tile"] > 0
    if tiled:
        kernel_bytes = BYTES_PER_PX_TILE * n
        kernel_ms = d["tile_ms"]
        kernel_name = ("grad2d_iteration_ring_kernel<SQUARE, IND_LEQ0, CHECK=false> (whole PDHG iteration in one "
                       "pass, persistent TMA ring)")
    else:
        kernel_bytes = BYTES_PER_PX_DUAL * n
        kernel_ms = d["dual_ms"]
        kernel_name = "grad_dual_norm2_kernel (fused dual pass)"
    achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9
    n_all = max(d["n_tile"] + d["n_two_pass"] + d["n_tile_check"], 1.0)
    frac_tile, frac_chk, frac_two = d["n_tile"] / n_all, d["n_tile_check"] / n_all, d["n_two_pass"] / n_all
    iter_bytes = (frac_tile * BYTES_PER_PX_TILE + frac_chk * BYTES_PER_PX_TILE_CHECK + frac_two * BYTES_PER_PX_ITER) * n
    achieved_iter = iter_bytes * value / 1e9
    roofline = {
        "bound": "hbm", "kernel": kernel_name,
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel, ncu --set full capture
        # committed as profiles/r01_tile_ring_full_raw.csv (268.5 + 160.5 MB: part of the previous
        # iteration's output is still in the 126 MB L2, so less than the algorithmic 469.8 MB)
        "traffic": 428923648 if tiled else None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": kernel_bytes, "ms_per_launch": kernel_ms,
        "launch_share": frac_tile if tiled else frac_two,
        "residual_refresh_iterations": {
            "share": frac_chk, "kernel": "grad2d_iteration_ring_kernel<..., CHECK=true>",
            "ms_per_launch": d["tile_check_ms"], "algorithmic_bytes_per_launch": BYTES_PER_PX_TILE_CHECK * n,
            "achieved": (BYTES_PER_PX_TILE_CHECK * n / (d["tile_check_ms"] * 1e-3) / 1e9) if d["tile_check_ms"] else None,
            "note": "1 in residual_iter iterations: the same pass also reads the previous dual iterate and "
                    "accumulates the four residual sums; included in value"},
        "two_pass_iterations": {
            "share": frac_two,
            "primal_pass": {"ms_per_launch": d["primal_ms"], "algorithmic_bytes_per_launch": BYTES_PER_PX_PRIMAL * n},
            "dual_pass": {"ms_per_launch": d["dual_ms"], "algorithmic_bytes_per_launch": BYTES_PER_PX_DUAL * n},
            "note": "iteration 0 only (K^T y := 0, K x_prev := 0 special cases)"},
        "whole_iteration": {"algorithmic_bytes": iter_bytes, "achieved": achieved_iter, "frac": achieved_iter / peak,
                            "frac_of_8TBs_nominal": achieved_iter / 8000.0},
        # the same iterations/s expressed against SURVEY.md 8(d)'s two-pass minimum of 44 B/px
        "equivalent_at_44B_per_px": {"achieved": BYTES_PER_PX_ITER * n * value / 1e9,
                                     "frac": BYTES_PER_PX_ITER * n * value / 1e9 / peak},
        "finalize_ms": d["finalize_ms"],
    }
    res = be.residuals()
    del be

    # ---------------- end to end through the public API with host buffers (e2e) --------------------
    f_pin = torch.from_numpy(f).pin_memory()
    x0_pin = torch.zeros(n, dtype=torch.float32).pin_memory()
    y0_pin = torch.zeros(m, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prob2 = build_problem(pb, syn, ctx, f_pin.numpy())           # H2D: f (prox coefficient b)
    be2 = pb.BackendPDHG(ctx, prob2, popts, sopts)
    solver = pb.Solver(prob2, be2)
    solver.SetOptions(sopts, x0=x0_pin.numpy(), y0=y0_pin.numpy())
    t1 = time.perf_counter()
    solver.Initialize()                                          # scaling upload, x0 / y0 H2D
    ctx.synchronize()
    t2 = time.perf_counter()
    solver.Solve()                                               # K iterations + D2H of x, z, y, w
    ctx.synchronize()
    t3 = time.perf_counter()
    t_e2e = t3 - t0
    h2d = 4 * (n + n + m) + 4 * (n + m)          # f, x0 (x2 buffers share one upload each), y0, scaling
    d2h = 4 * (2 * n + 2 * m)                    # x, w, y, z
    e2e = {"value": args.steps / t_e2e, "unit": "iter/s", "h2d_bytes_per_step": h2d / args.steps,
           "d2h_bytes_per_step": d2h / args.steps, "seconds_total": t_e2e,
           "seconds": {"problem_build_h2d": t1 - t0, "solver_initialize": t2 - t1, "solve_and_d2h": t3 - t2},
           "what": "Problem build + Solver.Initialize + Solver.Solve(max_iters=K, num_cback_calls=0) + solution "
                   "copy-back of x, z, y, w"}
    del be2, solver

    # ---------------- time to residual 1e-4 (third part of the BASELINE metric) ---------------------
    ttr = time_to_residual(pb, ctx, lambda: build_problem(pb, syn, ctx, f_pin.numpy()))

    # ---------------- CPU baseline (oracle port, bounded sample) ------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_binding import OracleProblem, OraclePDHG, num_threads
        cols = 1024
        o = OraclePDHG(OracleProblem(syn.rof(cols, NY, LAM)), stepsize="alg1", residual_iter=RESIDUAL_ITER,
                       tol_rel_primal=0, tol_rel_dual=0, tol_abs_primal=0, tol_abs_dual=0)
        o.initialize()
        o.iterate(2)
        t0 = time.perf_counter()
        its = 0
        while time.perf_counter() - t0 < 12.0:
            o.iterate(5)
            its += 5
        dt = time.perf_counter() - t0
        cpu = {"value": its / dt * cols / NX, "unit": "iter/s", "cores": num_threads(), "kind": "port",
               "sample": f"{its} iterations of the OpenMP oracle on a {cols}x{NY} slab (1/{NX // cols} of the "
                         f"image), scaled by area"}

    line = {
        "metric": "pdhg_iterations_per_second", "value": value, "unit": "iter/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(1), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "time_to_residual_1e-4": ttr, "residuals_after_run": res,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
