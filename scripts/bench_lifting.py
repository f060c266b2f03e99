#!/usr/bin/env python
"""BASELINE config 3: lifted multilabel problem 2048 x 2048 x 32 labels (in-tree operators: BlockGradient2D over
the label planes + identity block, ProxElemOperation<IndSimplex>, Norm2 ball on the gradient rows, ProxIndEpiQuad on
the identity rows), PDHG 'boyd' step sizes, slab-sharded over the GPUs of one box.

    python scripts/bench_lifting.py [--steps 100 --warmup 10 --nx 2048 --ny 2048 --labels 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_lifting.py

Prints ONE JSON line (rank 0): iterations/s of the whole problem (strong scaling: the same 2048^2 x 32 problem
on N GPUs), device-timed with CUDA events, max over ranks.  Algorithmic bytes per iteration: 56 B per
pixel-label (primal pass reads y (3), x, writes x+; dual passes read x+, x, y (3), b/c coefficients (1), write
y+ (3) = 14 floats)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BYTES_PER_PIXEL_LABEL = 56


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--labels", type=int, default=32)
    args = ap.parse_args()
    import torch
    import prost_b200 as pb
    from prost_b200 import synthetic as syn
    from prost_b200 import distributed as pbd
    from bench import measured_peak

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pb.Context(local, stream.cuda_stream)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = pbd.init_comm(ctx)

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    part = pbd.SlabPartition(args.nx, world, align=4)
    x0, x1 = part.range(rank)
    t0 = time.perf_counter()
    desc = syn.lifting(args.nx, args.ny, args.labels, x0=x0, x1=x1)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    prob = pb.create_problem(ctx, desc)
    prob.Initialize()
    popts = pb.pdhg_options(scale_steps_operator=0, stepsize="boyd", residual_iter=10)
    sopts = pb.solver_options(verbose=0, max_iters=args.steps, tol_rel_primal=0, tol_rel_dual=0, tol_abs_primal=0,
                              tol_abs_dual=0, num_cback_calls=0)
    be = pb.BackendPDHG(ctx, prob, popts, sopts, comm=comm)
    be.Initialize()
    ctx.synchronize()
    t_setup = time.perf_counter() - t0
    assert be.is_fused, "fused PDHG path not selected"
    be.PerformIteration(args.warmup)
    torch.cuda.synchronize()
    launches0 = be.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    be.PerformIteration(args.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = be.launch_count - launches0
    res = be.residuals()
    value = args.steps / (ms * 1e-3)
    n_pl = args.nx * args.ny * args.labels
    peak, peak_src = measured_peak()
    wmax = max(part.width(r) for r in range(world))
    per_gpu_bytes = BYTES_PER_PIXEL_LABEL * wmax * args.ny * args.labels
    ach = per_gpu_bytes * value / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "pdhg_iterations_per_second", "value": value, "unit": "iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"lifted multilabel {args.nx}x{args.ny}x{args.labels} labels (BlockGradient2D + identity "
                                   f"block, ind_simplex, norm2:ind_leq0 dim {2 * args.labels}, ind_epi_quad), PDHG boyd, "
                                   f"residual_iter=10, alpha=1 preconditioning",
                       "n_pixel_labels": n_pl, "bytes_per_iteration_algorithmic": BYTES_PER_PIXEL_LABEL * n_pl,
                       "parallelism": f"slab{world}" if world > 1 else "single"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole iteration per GPU (specialised two-pass stencil kernels)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_iteration_per_gpu": per_gpu_bytes},
            "halo_mode": (("peer-to-peer stores over NVLink (CUDA IPC)" if comm.peer_to_peer else "NCCL send/recv staging")
                          if comm else None),
            "host_seconds": {"generate_description": t_gen, "create_initialize": t_setup},
            "residuals_after_run": res}), flush=True)
    del be
    if comm:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
