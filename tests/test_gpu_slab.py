"""GPU suite, multi-GPU part: the slab decomposition must reproduce the single-GPU
iterates.  Needs >= 2 GPUs (``gpurun --gpus 2``); with one GPU only the world-size-1 communicator
path is exercised.  Each configuration runs through ``torch.distributed.run`` (one process per
GPU) in both halo modes: peer-to-peer stores from inside the fused passes (CUDA IPC over NVLink)
and NCCL send/recv staging."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import prost_b200 as pb
    return pb.device_count()


def _run(world, halo, only=""):
    env = dict(os.environ, PB_HALO=halo)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29711 + world + (7 if halo == "nccl" else 0)),
           os.path.join(ROOT, "tests", "slab_worker.py"), only]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("SLAB_REPORT ")]
    assert p.returncode == 0 and lines, f"worker failed rc={p.returncode}\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return json.loads(lines[-1][len("SLAB_REPORT "):])


def _check(rep, expect_p2p):
    assert rep["cases"], "no cases ran"
    if expect_p2p is not None:
        assert rep["p2p"] == expect_p2p
    for name, c in rep["cases"].items():
        # same float operations on the same data: the iterates agree far below the 1e-5 bar (the
        # only difference is the fold order of the double residual sums, which can flip the last
        # bit of a float residual and, through adaptive steps, perturb iterates at 1e-7)
        for k, e in c["err"].items():
            assert e <= 2e-6, f"{name}: {k} differs from the single-GPU run by {e:.3e}"
        for k, v in c["res"].items():
            ref = c["res_single"][k]
            assert abs(v - ref) <= 1e-5 * max(abs(ref), 1e-6), f"{name}: {k} {v} vs {ref}"
        for a, b in zip(c["steps"], c["steps_single"]):
            assert abs(a - b) <= 1e-6 * max(abs(b), 1e-12), f"{name}: step sizes"
        # ROF-shaped problems run every iteration but the first as ONE pass (pb_tile.cu ring kernel); on
        # slabs that needs the peer-to-peer halo blocks (or a world of one rank)
        if name.startswith("rof") and name != "rof_scalar":
            assert c["one_pass_single"] == c["iters"] - 1, f"{name}: single-GPU run did not use the ring kernel"
            if rep["p2p"] or rep["world"] == 1:
                assert c["one_pass"] == c["iters"] - 1, f"{name}: slabs did not use the one-pass ring kernel"
            else:
                assert c["one_pass"] == 0


def test_world1_communicator():
    """A communicator of one rank: no neighbours, all-reduce is the identity."""
    if _gpus() < 1:
        pytest.skip("no GPU")
    _check(_run(1, "p2p", "rof_vec4,lifting,lifting_coop,rof_tiles65,rof_tiles_alg2"), None)


@pytest.mark.parametrize("halo", ["p2p", "nccl"])
def test_two_slabs_match_single_gpu(halo):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _check(_run(2, halo), halo == "p2p")


def test_all_gpus_match_single_gpu():
    n = _gpus()
    if n < 3:
        pytest.skip("needs more than 2 GPUs")
    _check(_run(min(n, 8), "p2p"), True)
