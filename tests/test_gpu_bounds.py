"""The bounds-checked build of the library (-DPPP_CHECK_BOUNDS: device-side assertions on the index arithmetic of
every kernel that reads or writes without clamping -- unclamped candidate loads behind the sorted array, key stores
through shared-memory cursors, positions rebuilt from ordinals, distributed-shared-memory addresses of the cluster
sort, counting-sort slots, scan tiles) runs a small pass over the hot path without tripping one.
compute-sanitizer is not available on the measurement pool; this is the stand-in."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_LIB = os.path.join(ROOT, "polishpathplanning_b200", "libppp_gpu_check.so")


def test_small_pass_with_device_side_bounds_assertions():
    if not os.path.exists(CHECK_LIB):
        pytest.skip("libppp_gpu_check.so not built (python -m polishpathplanning_b200.build --check)")
    env = dict(os.environ, PPP_GPU_LIB=CHECK_LIB)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "memcheck_small.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    out = r.stdout + r.stderr
    assert "PPP_DEV_ASSERT failed" not in out, out[-2000:]
    assert r.returncode == 0 and "done" in r.stdout, out[-2000:]
