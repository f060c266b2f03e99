"""GPU test of the optional multi-iteration ring launches (PB_RING_ITERS > 1, RingMulti in pb_tile.cu): several
non-refresh PDHG iterations per launch with per-tile (or per-CTA) dependencies instead of the kernel boundary.  The
arithmetic is the same, only the launch structure differs, so x, y, z, w and the residuals must be bit-identical to
single launches -- odd and even batch sizes, refresh iterations in between, PerformIteration split in two calls
(scripts/check_ring_multi.py runs each setting in its own process because the switch is read once per process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_iteration_launches_are_bit_identical():
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_ring_multi.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("rank 0 ")]
    assert len(lines) >= 12 and all(ln.endswith("bit-identical") for ln in lines), p.stdout[-3000:]
