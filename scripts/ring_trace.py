#!/usr/bin/env python
"""Scaling experiment: timeline of the one-pass ring kernel's launches on this rank's slab (PB_RING_TRACE=1).

    PB_RING_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/ring_trace.py

Prints per rank: mean kernel duration (first CTA start -> last CTA end), mean gap between consecutive launches
(end -> next start), mean / max of the longest halo waits of the left- and right-edge tiles, all in microseconds."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import prost_b200 as pb
    from prost_b200 import synthetic as syn
    from prost_b200 import distributed as pbd
    from prost_b200._capi import lib
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nx = int(os.environ.get("PB_TRACE_NX", "4096"))
    torch.cuda.set_device(local)
    ctx = pb.Context(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        comm = pbd.init_comm(ctx)
    part = pbd.SlabPartition(nx, world, align=4)
    x0, x1 = part.range(rank)
    prob = pb.create_problem(ctx, syn.rof(x1 - x0, 4096, 10.0, f=syn.image(nx, 4096, x0=x0, x1=x1)))
    popts = pb.pdhg_options(scale_steps_operator=0, stepsize="alg1", residual_iter=10)
    sopts = pb.solver_options(verbose=0, max_iters=1000, num_cback_calls=0, tol_rel_primal=0, tol_rel_dual=0,
                              tol_abs_primal=0, tol_abs_dual=0)
    be = pb.BackendPDHG(ctx, prob, popts, sopts, comm=comm)
    prob.Initialize()
    be.Initialize()
    be.PerformIteration(1001)
    ctx.synchronize()
    n = 1000
    buf = np.zeros(4 * n, dtype=np.uint64)
    got = lib.pb_ring_trace_read(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), n)
    t = buf.reshape(n, 4).astype(np.float64)[200:min(got, n)]
    dur = (t[:, 1] - t[:, 0]) / 1e3
    gap = (t[1:, 0] - t[:-1, 1]) / 1e3
    period = (t[1:, 0] - t[:-1, 0]) / 1e3
    idx = np.arange(200, 200 + len(t))
    chk = (idx % 10) == 9          # launch k is iteration k+1 (iteration 0 runs the two-pass kernels)
    out = {"rank": rank, "world": world, "launches": int(got), "slab_cols": x1 - x0,
           "kernel_us_plain": float(dur[~chk].mean()), "kernel_us_refresh": float(dur[chk].mean()),
           "gap_us_mean": float(gap.mean()), "gap_us_p90": float(np.percentile(gap, 90)),
           "period_us_mean": float(period.mean()),
           "wait_left_us_mean": float(t[:, 2].mean() / 1e3), "wait_left_us_max": float(t[:, 2].max() / 1e3),
           "wait_right_us_mean": float(t[:, 3].mean() / 1e3), "wait_right_us_max": float(t[:, 3].max() / 1e3)}
    print("RING_TRACE " + json.dumps(out), flush=True)
    if comm:
        comm.barrier()
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
