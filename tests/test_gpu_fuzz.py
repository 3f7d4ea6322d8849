"""Randomised shapes through every mode (BSVD-64 fp16 / bf16 / fp32x3, blind c32) against the fp32 CPU oracle,
with the streaming schedule required to reproduce the clip schedule bit for bit (tools/fuzz_sizes.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_shapes_all_modes():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_sizes.py"), "20", "11"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, (r.stdout + r.stderr)[-3000:]
    res = json.loads(line[-1])
    assert r.returncode == 0 and not res["failures"], res
    assert res["worst_max_abs"]["fp32x3"] < 1e-4 and res["worst_max_abs"]["fp16"] < 1e-3
