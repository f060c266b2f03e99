"""GPU suite, multi-GPU part: row-sharded BackendADMM (BASELINE config 5 "row-sharded") must
reproduce the single-GPU solve.  K x is local to a rank's rows, K^T r is summed with ncclAllReduce, sums over rows
inside the reduction kernels through peer-mapped slots.  Needs >= 2 GPUs (``gpurun --gpus 2``)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29811 + world),
           os.path.join(ROOT, "tests", "admm_shard_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("ADMM_SHARD_REPORT ")]
    assert p.returncode == 0 and lines, f"worker failed rc={p.returncode}\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return json.loads(lines[-1][len("ADMM_SHARD_REPORT "):])


@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_admm_matches_single_gpu(world):
    import prost_b200 as pb
    if pb.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    rep = _run(world)
    assert rep["cases"]
    for name, c in rep["cases"].items():
        assert c["replicas_identical"], name
        # the partial K^T r vectors are summed in a different order than one GPU sums its rows: iterates agree to
        # float rounding amplified by the CG recurrences (north star: 1e-5 on iterates, 1e-4 on residuals)
        for k, e in c["err"].items():
            assert e <= 2e-5, f"{name}: {k} differs from the single-GPU run by {e:.3e}"
        assert c["iterations"] == c["iterations_single"], name
        assert abs(c["cg"] - c["cg_single"]) <= max(2, 0.02 * c["cg_single"]), (name, c["cg"], c["cg_single"])
        for k in ("primal_var_norm", "dual_var_norm", "eps_primal", "eps_dual"):
            assert abs(c["res"][k] - c["res_single"][k]) <= 1e-4 * max(abs(c["res_single"][k]), 1e-6), (name, k)
        for k, scale in (("primal_residual", "primal_var_norm"), ("dual_residual", "dual_var_norm")):
            assert abs(c["res"][k] - c["res_single"][k]) <= 1e-4 * max(abs(c["res_single"][scale]), 1e-6), (name, k)
