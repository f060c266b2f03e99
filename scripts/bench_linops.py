#!/usr/bin/env python
"""Operator-apply timings outside the fused PDHG passes: CSR SpMV and
dense GEMV at BASELINE config 5's sizes, the four Kronecker blocks at example_multilabel_tight.m:76-87's shapes.

    python scripts/bench_linops.py [--reps 20] [--only sparse,dense,kron]

One JSON line: per operator and direction the CUDA-event time of one LinearOperator::Eval / EvalAdjoint on device
vectors (torch tensors, pb_linop_eval), the algorithmic bytes (matrix entries once + both vectors) and the
fraction of the measured HBM copy bandwidth.  Under ncu the same script gives the launch list / full captures."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="sparse,dense,kron")
    args = ap.parse_args()
    import numpy as np
    import scipy.sparse as sp
    import torch
    import prost_b200 as pb
    from prost_b200 import synthetic as syn
    from prost_b200._capi import lib, check
    from bench import measured_peak

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pb.Context(0, stream.cuda_stream)
    peak, _ = measured_peak()
    fp = C.POINTER(C.c_float)
    out = {}

    def time_op(name, blocks, matrix_bytes, flush=None):
        op = pb.create_linop(ctx, blocks)
        m, n = op.nrows, op.ncols
        x = torch.rand(n, device="cuda")
        y = torch.rand(m, device="cuda")
        rx, ry = torch.empty(m, device="cuda"), torch.empty(n, device="cuda")
        for tag, res, rhs, tr in (("forward", rx, x, 0), ("adjoint", ry, y, 1)):
            def run():
                check(lib.pb_linop_eval(op._h, C.cast(res.data_ptr(), fp), C.cast(rhs.data_ptr(), fp), 0.0, tr))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                run()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            nbytes = matrix_bytes + 4 * (m + n)
            out[f"{name}:{tag}"] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / (ms * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak, "rows": m, "cols": n}
        del op

    only = set(args.only.split(","))
    if "sparse" in only:
        desc = syn.lasso(4194304, 1048576, nnz_per_row=12, dense=0)
        K = desc["blocks"][0][3][0]
        time_op("sparse_4194304x1048576_12nnz", [("sparse", 0, 0, [K])], 8 * K.nnz)
    if "dense" in only:
        D = (syn.normal(46, np.arange(4096 * 4096, dtype=np.uint64)) / 64.0).reshape(4096, 4096)
        time_op("dense_4096x4096", [("dense", 0, 0, [D])], 4 * 4096 * 4096)
    if "kron" in only:
        L, npix = 16, 1024 * 1024
        r = np.random.default_rng(0)
        ones = sp.csc_matrix(np.ones((1, L), np.float32))                       # sum over labels
        pairs = [(i, j) for i in range(L) for j in range(i + 1, L)]
        P = sp.lil_matrix((L, len(pairs)), dtype=np.float32)                     # pair_local'
        for k, (i, j) in enumerate(pairs):
            P[i, k], P[j, k] = -1.0, 1.0
        P = sp.csc_matrix(P)
        Kd = r.standard_normal((16, 32)).astype(np.float32)
        K64 = r.standard_normal((64, 64)).astype(np.float32)
        time_op(f"sparse_kron_id_ones1x{L}_d{npix}", [("sparse_kron_id", 0, 0, [ones, npix])], 0)
        time_op(f"sparse_kron_id_pairs{L}x{len(pairs)}_d{npix // 4}", [("sparse_kron_id", 0, 0, [P, npix // 4])], 0)
        time_op(f"id_kron_sparse_pairs{L}x{len(pairs)}_d{npix // 4}", [("id_kron_sparse", 0, 0, [P, npix // 4])], 0)
        time_op(f"dense_kron_id_16x32_d{npix}", [("dense_kron_id", 0, 0, [Kd, npix])], 0)
        time_op(f"id_kron_dense_16x32_d{npix}", [("id_kron_dense", 0, 0, [Kd, npix])], 0)
        time_op(f"dense_kron_id_64x64_d{npix // 4}", [("dense_kron_id", 0, 0, [K64, npix // 4])], 0)
        time_op(f"dense_kron_id_64x64_d{npix}", [("dense_kron_id", 0, 0, [K64, npix])], 0)
        time_op(f"id_kron_dense_64x64_d{npix}", [("id_kron_dense", 0, 0, [K64, npix])], 0)
    print(json.dumps({"peak_GBps": peak, "ops": out}), flush=True)


if __name__ == "__main__":
    main()
