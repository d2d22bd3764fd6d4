"""Multi-GPU parity across PROCESSES (needs >= 2 GPUs on the box; skipped otherwise): tools/multi_gpu_check.py
under torchrun.  The same exchange kernels, flags and result delivery run on ONE GPU in
tests/test_gpu_exchange.py, which every box executes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_exchange_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PPP_CHECK_N="200000")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("normals_bitexact=True knn_ids=True contours=True") == 2
