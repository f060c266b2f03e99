"""One rank of the multi-GPU slab parity check (launched by tests/test_gpu_slab.py through
``python -m torch.distributed.run``).  Every rank runs its column slab of the same global problem
with halo exchange + residual all-reduce; rank 0 additionally runs the whole problem on one GPU
and compares: the slab iterates must be the single-GPU iterates restricted to the slab."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    from prost_b200 import synthetic as syn
    return {
        "rof_vec4": (lambda: syn.rof(40, 36), dict(stepsize="alg1", residual_iter=3), 60),
        "rof_scalar": (lambda: syn.rof(37, 35), dict(stepsize="boyd", residual_iter=2), 60),
        "rof_alg2": (lambda: syn.rof(64, 128), dict(stepsize="alg2", residual_iter=5, alg2_gamma=0.5), 50),
        "rof_goldstein": (lambda: syn.rof(33, 64), dict(stepsize="goldstein", residual_iter=4), 50),
        "tvl1": (lambda: syn.tvl1(40, 32, nc=3), dict(stepsize="boyd", residual_iter=5), 60),
        "tv3d": (lambda: syn.tv3d(14, 16, 6), dict(stepsize="alg1", residual_iter=5), 40),
        "lifting": (lambda: syn.lifting(18, 12, 6), dict(stepsize="boyd", residual_iter=5), 40),
        # columns of 64 / 128 pixels: the staged passes copy their operands cooperatively, halo columns included
        "lifting_coop": (lambda: syn.lifting(24, 64, 8), dict(stepsize="boyd", residual_iter=5), 40),
        "lifting_coop128": (lambda: syn.lifting(16, 128, 6), dict(stepsize="alg1", residual_iter=6), 30),
        "rof_big": (lambda: syn.rof(1024, 512), dict(stepsize="alg1", residual_iter=10), 200),
        # one-pass ring kernel on slabs (pb_tile.cu, SLAB): several tile columns / rows per slab; slab width 63
        # puts the right edge column in a one-column tile (62 = 2 * 31), width 65 in the middle of one
        "rof_tiles65": (lambda: syn.rof(130, 252), dict(stepsize="boyd", residual_iter=4), 60),
        "rof_tiles63": (lambda: syn.rof(126, 128), dict(stepsize="goldstein", residual_iter=3), 60),
        "rof_tiles_alg2": (lambda: syn.rof(200, 380), dict(stepsize="alg2", residual_iter=7, alg2_gamma=0.5), 45),
    }


def main():
    import torch
    import torch.distributed as dist
    import prost_b200 as pb
    from prost_b200 import distributed as pbd
    from pdhg_util import TOL

    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = pb.Context(local)
    comm = pbd.init_comm(ctx)
    report = {"world": world, "p2p": comm.peer_to_peer, "cases": {}}
    only = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] else None

    for name, (make, opts, iters) in cases().items():
        if only and name not in only:
            continue
        desc = make()
        nx, ny, L = pbd._grid_of(desc)
        part = pbd.SlabPartition(nx, world)
        local_desc = pbd.shard_description(desc, part, rank)
        tol = dict(tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4)

        def run(d, c):
            prob = pb.create_problem(ctx, d)
            popts = pb.pdhg_options(scale_steps_operator=0, **opts)
            sopts = pb.solver_options(verbose=0, max_iters=iters, num_cback_calls=0, **tol)
            be = pb.BackendPDHG(ctx, prob, popts, sopts, comm=c)
            prob.Initialize()
            be.Initialize()
            # two chunks with a solution read in between: exercises the collective
            # current_solution and the sequence numbering across calls
            be.PerformIteration(iters // 2)
            be.current_solution()
            be.PerformIteration(iters - iters // 2)
            x, z, y, w = be.current_solution()
            return dict(x=x, z=z, y=y, w=w, res=be.residuals(), steps=be.stepsizes(), fused=be.is_fused,
                        one_pass=int(be.one_pass_iterations))

        mine = run(local_desc, comm)
        gathered = [None] * world
        dist.gather_object({k: mine[k] for k in ("x", "z", "y", "w")}, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            ref = run(desc, None)
            out = {"res": mine["res"], "res_single": ref["res"], "steps": mine["steps"],
                   "steps_single": ref["steps"], "err": {}, "iters": iters,
                   "one_pass": mine["one_pass"], "one_pass_single": ref["one_pass"]}
            for k in ("x", "z", "y", "w"):
                glob = pbd.gather_planar([g[k] for g in gathered], part, ny)
                denom = max(float(np.abs(ref[k]).max()), 1e-30)
                out["err"][k] = float(np.abs(glob.astype(np.float64) - ref[k]).max() / denom)
            report["cases"][name] = out
        comm.barrier()

    if rank == 0:
        print("SLAB_REPORT " + json.dumps(report), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
