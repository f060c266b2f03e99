"""One rank of the row-sharded ADMM parity check (launched by tests/test_gpu_admm_sharded.py through
``python -m torch.distributed.run``): every rank solves its block of rows of the same LASSO-type problem; rank 0
also solves the whole problem on one GPU and compares x (replicated), the gathered z / y, w and the residuals."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import prost_b200 as pb
    from prost_b200 import distributed as pbd
    from prost_b200 import synthetic as syn
    import admm_cases

    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = pb.Context(local)
    comm = pbd.init_comm(ctx)
    report = {"world": world, "p2p": comm.peer_to_peer, "cases": {}}
    cases = {
        "lasso_sparse": (lambda: syn.lasso(3000, 700, nnz_per_row=5), 25, dict(residual_iter=1)),
        "lasso_sparse_dense": (lambda: syn.lasso(2500, 640, nnz_per_row=6, dense=96), 25, dict(residual_iter=2)),
        "lasso_tol": (lambda: syn.lasso(2048, 512, nnz_per_row=8), 60, dict(residual_iter=1, cg_max_iter=20)),
    }
    for name, (make, iters, opts) in cases.items():
        desc = make()
        part = pbd.RowPartition(desc["nrows"], world)
        local_desc = pbd.shard_rows(desc, part, rank)
        tol = dict(tol_rel_primal=1e-3, tol_rel_dual=1e-3, tol_abs_primal=1e-3, tol_abs_dual=1e-3) \
            if name == "lasso_tol" else dict(tol_rel_primal=0, tol_rel_dual=0, tol_abs_primal=0, tol_abs_dual=0)

        def run(d, c):
            prob = pb.create_problem(ctx, d)
            sopts = pb.solver_options(verbose=0, max_iters=iters, num_cback_calls=0, **tol)
            be = pb.BackendADMM(ctx, prob, pb.admm_options(**opts), sopts, comm=c)
            solver = pb.Solver(prob, be)
            solver.SetOptions(sopts)
            solver.Initialize()
            solver.Solve()
            return dict(x=np.array(solver.cur_primal_sol), z=np.array(solver.cur_primal_constr_sol),
                        y=np.array(solver.cur_dual_sol), w=np.array(solver.cur_dual_constr_sol), res=be.residuals(),
                        iterations=int(solver.iterations), cg=be.stepsizes()[2])

        mine = run(local_desc, comm)
        gathered = [None] * world
        dist.gather_object({k: mine[k] for k in ("x", "z", "y", "w")}, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            ref = run(desc, None)
            out = {"res": mine["res"], "res_single": ref["res"], "iterations": mine["iterations"],
                   "iterations_single": ref["iterations"], "cg": mine["cg"], "cg_single": ref["cg"], "err": {}}
            glob = {"x": gathered[0]["x"], "w": gathered[0]["w"],
                    "z": np.concatenate([g["z"] for g in gathered]), "y": np.concatenate([g["y"] for g in gathered])}
            for k in ("x", "z", "y", "w"):
                denom = max(float(np.abs(ref[k]).max()), 1e-30)
                out["err"][k] = float(np.abs(glob[k].astype(np.float64) - ref[k]).max() / denom)
            # the replicated vectors must be IDENTICAL on all ranks (same all-reduced bits, same recurrences)
            out["replicas_identical"] = bool(all(np.array_equal(g["x"], gathered[0]["x"]) and
                                                 np.array_equal(g["w"], gathered[0]["w"]) for g in gathered))
            report["cases"][name] = out
        comm.barrier()
    if rank == 0:
        print("ADMM_SHARD_REPORT " + json.dumps(report), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
