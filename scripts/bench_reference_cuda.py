#!/usr/bin/env python
"""BASELINE.md line A: the UNMODIFIED reference CUDA solver (tum-vision/prost compiled for sm_100 into
oracle/_ref/, see oracle/ref_build/Makefile) and this repository's library timed by the SAME C++ program
(oracle/driver/prost_driver.cu, written against prost's public C++ API; prost_b200/lib/prost_b200_driver is
that source compiled against include/prost/*.hpp + the C ABI) on the metric config: ROF-TV nx x ny, PDHG Alg1,
residual_iter = 10, K iterations, wall clock around Solver::Solve() as reported by the program (`solve_ms`:
iterations + the final copy-back of x, z, y, w).

    python scripts/bench_reference_cuda.py [--nx 4096 --ny 4096 --iters 500]

Needs a GPU and oracle/_ref (built where /root/reference is mounted; travels to the GPU box).  Prints one JSON
line with both rates and the maximum relative difference of the iterates."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=500)
    args = ap.parse_args()
    import numpy as np
    import ref_driver
    from prost_b200 import synthetic as syn
    desc = syn.rof(args.nx, args.ny, 10.0)
    opts = dict(stepsize="alg1", residual_iter=10, timeout=3000)
    out = {"workload": f"ROF-TV {args.nx}x{args.ny}, PDHG Alg1, residual_iter=10, {args.iters} iterations, "
                       f"same C++ driver program for both libraries"}
    runs = {}
    for tag, binary in (("reference_sm100", ref_driver.REF_DRIVER), ("prost_b200", ref_driver.OUR_DRIVER)):
        if not ref_driver.available(binary):
            out[tag] = {"unavailable": os.path.relpath(binary, ROOT)}
            continue
        r = ref_driver.run_solve(desc, args.iters, binary=binary, **opts)
        ms = float(r["info"]["solve_ms"])
        runs[tag] = r
        out[tag] = {"solve_ms": ms, "iter_per_s": args.iters / (ms * 1e-3), "residuals": r["res"]}
    if len(runs) == 2:
        a, b = runs["reference_sm100"], runs["prost_b200"]
        out["max_rel_diff"] = {k: float(np.abs(a[k] - b[k]).max() / max(float(np.abs(a[k]).max()), 1e-30))
                               for k in ("x", "y", "z", "w")}
        out["speedup"] = out["prost_b200"]["iter_per_s"] / out["reference_sm100"]["iter_per_s"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
