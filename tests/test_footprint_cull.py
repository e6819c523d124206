"""The conservative footprint tests of the composite kernels (bounding box of the alpha >= 1/255 region, and the opt-in
exact ellipse-vs-rectangle test, composite_common.cuh: footprint_hits_rect) may only drop (rectangle, instance) pairs in
which NO pixel passes the reference's blend conditions (forward.cu:344-352).  tools/check_exact_cull.py restates both
tests in float32 numpy and checks that against the CPU oracle's per-Gaussian intermediates on a tile sample; it exits
non-zero on the first violated pair.  CPU only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config,tiles", [("tiny", 16), ("small", 48)])
def test_footprint_tests_never_drop_a_blending_pair(config, tiles, built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_exact_cull.py"), config, str(tiles)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok: no needed pair dropped" in r.stdout
