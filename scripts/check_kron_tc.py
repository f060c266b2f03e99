#!/usr/bin/env python
"""Bring-up check of the tcgen05 Kronecker path (pb_kron_tc.cu): max-norm error of every tensor-core case against a
float64 product, forward and adjoint.."""
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import prost_b200 as pb
import cases

ctx = pb.Context(0)
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, blocks in sorted(cases.linop_kron_tensor_core_cases().items()):
    if only and only not in name:
        continue
    op = pb.create_linop(ctx, blocks)
    r = np.random.default_rng(zlib.crc32(name.encode()))
    x, y = r.standard_normal(op.ncols).astype(np.float32), r.standard_normal(op.nrows).astype(np.float32)
    fwd, adj = op.Eval(x), op.EvalAdjoint(y)
    wf, wa = cases.kron_apply_f64(blocks, x, False), cases.kron_apply_f64(blocks, y, True)
    ef, ea = np.abs(fwd - wf).max() / np.abs(wf).max(), np.abs(adj - wa).max() / np.abs(wa).max()
    print(f"{name:34s} forward {ef:.2e}  adjoint {ea:.2e}  {'ok' if max(ef, ea) <= 1e-5 else 'BAD'}", flush=True)
