"""-m gpu (needs >= 2 GPUs, skipped otherwise): one scene shared by two ranks - every rank integrates the voxel blocks it
owns and casts its share of the raycast tiles, results travel as NVLink peer stores - must stay BITWISE equal to a
single-GPU engine on the same frames (pose, hash table, voxel blocks, visible list, raycast image, ICP maps), frame after
frame, on every rank.  tools/sharded_run.py --check does the comparison and exits non-zero on any difference."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs with peer access")
def test_two_rank_sharded_run_is_bitwise_equal_to_single_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "sharded_run.py"), "--check", "--frames", "6", "--size", "320x240",
           "--voxel", "0.005", "--pool", "0x10000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("BITWISE EQUAL") == 12
