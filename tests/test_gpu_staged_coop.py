"""GPU test: cooperative staging of the lifting passes (16-byte row copies shared by a CTA, PB_STAGED_COOP=1, the
default when a column is a multiple of 64 pixels) is bit-identical to the per-thread staging (PB_STAGED_COOP=0) on
problems whose columns span one, two and three CTAs, 4 - 32 labels, three step-size rules, refresh iterations in
between (scripts/check_staged_coop.py runs both settings in separate processes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cooperative_staging_is_bit_identical():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_staged_coop.py")], capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    lines = [ln for ln in p.stdout.splitlines() if ln.endswith("bit-identical")]
    assert len(lines) == 4, p.stdout[-3000:]
