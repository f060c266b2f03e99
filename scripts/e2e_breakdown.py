#!/usr/bin/env python
"""Where the fixed cost of an end-to-end solve goes (ROF 4096^2 through the public API, host buffers):
times Problem build, Problem.Initialize, Backend.Initialize and Solver.Solve for 1 and for K iterations."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import prost_b200 as pb
    from prost_b200 import synthetic as syn
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    NX = NY = 4096
    ctx = pb.Context(0)
    f = syn.image(NX, NY)
    n, m = NX * NY, 2 * NX * NY
    f_pin = torch.from_numpy(f).pin_memory().numpy()
    x0 = torch.zeros(n).pin_memory().numpy()
    y0 = torch.zeros(m).pin_memory().numpy()
    out = {}
    for rep, iters in enumerate([1, 1, K]):
        popts = pb.pdhg_options(scale_steps_operator=0, stepsize="alg1", residual_iter=10)
        sopts = pb.solver_options(verbose=0, max_iters=iters, tol_rel_primal=0, tol_rel_dual=0, tol_abs_primal=0,
                                  tol_abs_dual=0, num_cback_calls=0)
        ctx.synchronize()
        t = [time.perf_counter()]
        desc = syn.rof(NX, NY, 10.0, f=f_pin)
        t.append(time.perf_counter())
        prob = pb.create_problem(ctx, desc)
        ctx.synchronize(); t.append(time.perf_counter())
        be = pb.BackendPDHG(ctx, prob, popts, sopts)
        solver = pb.Solver(prob, be)
        solver.SetOptions(sopts, x0=x0, y0=y0)
        prob.Initialize()
        ctx.synchronize(); t.append(time.perf_counter())
        be.Initialize(x0, y0)
        ctx.synchronize(); t.append(time.perf_counter())
        solver.Solve()
        ctx.synchronize(); t.append(time.perf_counter())
        names = ["describe", "create_problem(h2d f)", "Problem.Initialize", "Backend.Initialize(h2d x0,y0)",
                 f"Solve({iters})+d2h"]
        out[f"run{rep}_iters{iters}"] = {k: round(b - a, 4) for k, a, b in zip(names, t[:-1], t[1:])}
        out[f"run{rep}_iters{iters}"]["total"] = round(t[-1] - t[0], 4)
        del solver, be, prob
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
