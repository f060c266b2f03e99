#!/usr/bin/env python
"""Round-2 bring-up check for the EXPERIMENTAL multi-iteration ring launches (PB_RING_ITERS > 1, pb_tile.cu
RingMulti): several non-refresh PDHG iterations per launch with per-tile dependencies instead of the kernel
boundary.  tests/test_gpu_ring_multi.py runs it on one GPU; on column slabs the switch is ignored.

    python scripts/check_ring_multi.py                 # one GPU: bitwise comparison against single launches
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_ring_multi.py

The script re-executes itself with PB_RING_ITERS=1 (reference) and PB_RING_ITERS=9 and compares x, y, z, w and
the residuals bit for bit (the arithmetic is identical; only the launch structure differs), then reports the
iteration rate of both settings on ROF 4096^2 (or this rank's slab of it)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [("rof_100x124", (100, 124), "alg1", 10, 57), ("rof_200x380_boyd", (200, 380), "boyd", 7, 64),
         ("rof_63x128_goldstein", (63, 128), "goldstein", 4, 41), ("rof_1024x512", (1024, 512), "alg1", 10, 200)]


def worker():
    import numpy as np
    import torch
    import prost_b200 as pb
    from prost_b200 import synthetic as syn
    from prost_b200 import distributed as pbd
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    ctx = pb.Context(local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        comm = pbd.init_comm(ctx)
    out = {}
    for name, (nx, ny), step, res_iter, iters in ([] if os.environ.get("PB_RING_DEBUG") else CASES):
        part = pbd.SlabPartition(nx, world)
        x0, x1 = part.range(rank)
        desc = syn.rof(x1 - x0, ny, 10.0, f=syn.image(nx, ny, x0=x0, x1=x1))
        prob = pb.create_problem(ctx, desc)
        popts = pb.pdhg_options(scale_steps_operator=0, stepsize=step, residual_iter=res_iter)
        sopts = pb.solver_options(verbose=0, max_iters=iters, num_cback_calls=0, tol_rel_primal=0, tol_rel_dual=0,
                                  tol_abs_primal=0, tol_abs_dual=0)
        be = pb.BackendPDHG(ctx, prob, popts, sopts, comm=comm)
        prob.Initialize()
        be.Initialize()
        be.PerformIteration(iters // 3)
        be.PerformIteration(iters - iters // 3)
        x, z, y, w = be.current_solution()
        np.savez(os.path.join(os.environ["PB_CHECK_DIR"], f"{name}_r{rank}_{os.environ['PB_CHECK_TAG']}.npz"),
                 x=x, y=y, z=z, w=w, res=np.array(list(be.residuals().values())))
        out[name] = int(be.one_pass_iterations)
        del be, prob
    # rate on the metric config
    part = pbd.SlabPartition(4096, world, align=4)
    x0, x1 = part.range(rank)
    prob = pb.create_problem(ctx, syn.rof(x1 - x0, 4096, 10.0, f=syn.image(4096, 4096, x0=x0, x1=x1)))
    popts = pb.pdhg_options(scale_steps_operator=0, stepsize="alg1", residual_iter=10)
    sopts = pb.solver_options(verbose=0, max_iters=2000, num_cback_calls=0, tol_rel_primal=0, tol_rel_dual=0,
                              tol_abs_primal=0, tol_abs_dual=0)
    be = pb.BackendPDHG(ctx, prob, popts, sopts, comm=comm)
    prob.Initialize()
    be.Initialize()
    be.PerformIteration(50)
    ctx.synchronize()
    t0 = time.perf_counter()
    be.PerformIteration(2000)
    ctx.synchronize()
    out["iter_per_s"] = 2000 / (time.perf_counter() - t0)
    if rank == 0:
        print("RING_MULTI " + json.dumps(out), flush=True)
    if comm:
        comm.close()
        dist.destroy_process_group()


def main():
    if os.environ.get("PB_CHECK_DIR"):
        return worker()
    import tempfile
    import numpy as np
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    tmp = os.environ.get("PB_CHECK_SHARED") or tempfile.mkdtemp(prefix="ring_multi_")
    reports = {}
    # tag -> environment: single launches (the reference), 9 iterations per launch with per-tile and with per-CTA
    # dependencies; the PB_RING_DEBUG variants skip synchronisation steps (timing experiments, results unusable)
    settings = [("1", dict(PB_RING_ITERS="1")), ("9", dict(PB_RING_ITERS="9")),
                ("9coarse", dict(PB_RING_ITERS="9", PB_RING_COARSE="1")),
                ("16coarse", dict(PB_RING_ITERS="16", PB_RING_COARSE="1"))]
    if os.environ.get("PB_CHECK_DEBUG_VARIANTS"):
        settings += [("9_noprobe", dict(PB_RING_ITERS="9", PB_RING_DEBUG="2")),
                     ("9_noprobe_norelease", dict(PB_RING_ITERS="9", PB_RING_DEBUG="3"))]
    for k, (tag, extra) in enumerate(settings):
        env = dict(os.environ, PB_CHECK_DIR=tmp, PB_CHECK_TAG=tag, **extra)
        if world > 1:
            env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 1 + k)
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=900)
        lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RING_MULTI ")]
        if p.returncode != 0:
            print(tag, p.stdout[-2000:], p.stderr[-2000:])
            sys.exit(1)
        if lines:
            reports[tag] = json.loads(lines[-1][len("RING_MULTI "):])
    ok = True
    for name, *_ in CASES:
        a = np.load(os.path.join(tmp, f"{name}_r{rank}_1.npz"))
        for tag in ("9", "9coarse", "16coarse"):
            b = np.load(os.path.join(tmp, f"{name}_r{rank}_{tag}.npz"))
            same = all(np.array_equal(a[k], b[k]) for k in ("x", "y", "z", "w", "res"))
            ok &= same
            print(f"rank {rank} {name} [{tag}]: {'bit-identical' if same else 'DIFFERENT'}")
    if rank == 0:
        print(json.dumps(reports))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
