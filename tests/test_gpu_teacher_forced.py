"""-m gpu: every CUDA stage against the REAL reference CPU engines (oracle/_ref), teacher forced.

The CUDA side is driven through the C ABI (libitm_b200.so) only.
"""
import numpy as np
import pytest

import parity
from infinitam_b200 import synth
from oracle import ref

pytestmark = pytest.mark.gpu


def _run(width, height, n_frames, noise=False, **kw):
    if not ref.available("parity"):
        pytest.skip("oracle/_ref/libitm_ref.so not built (needs /root/reference at build time)")
    oracle = ref.RefEngine(width, height, **kw)
    eng = parity.make_cuda_engine(oracle)
    seq = synth.sequence(n_frames, width, height, noise=noise)
    rows = []
    try:
        for k in range(n_frames):
            rows.append(parity.compare_frame(oracle, eng, seq[k], k, strict=True))
    finally:
        eng.close()
        oracle.close()
    return rows


def test_vga_5mm_teacher_forced():
    """BASELINE config 1/2: 640x480, 5 mm voxels, mu = 0.02."""
    rows = _run(640, 480, 5)
    assert all(r["hash_equal"] and r["visible_equal"] for r in rows)
    assert max(r["voxel_max_dsdf"] for r in rows) <= 1
    assert max(r["raycast_max_diff_m"] for r in rows) <= 1e-4
    assert max(r["pose_rot_rad"] for r in rows) <= 1e-4 and max(r["pose_trans_m"] for r in rows) <= 1e-4


def test_qvga_noisy_teacher_forced():
    """noisy depth with 1% dropped pixels (holes exercise the pyramid / bilinear hole logic)."""
    rows = _run(320, 240, 6, noise=True)
    assert all(r["hash_equal"] and r["visible_equal"] for r in rows)


def test_small_voxels_teacher_forced():
    """2.5 mm voxels: ray segments span several blocks (more steps per pixel, more hash collisions)."""
    rows = _run(320, 240, 3, voxel_size=0.0025)
    assert all(r["hash_equal"] and r["visible_equal"] for r in rows)


def test_vga_100_frames_teacher_forced():
    """the whole BASELINE sequence (100 frames, 640x480, 5 mm): late-sequence states - longer excess chains, re-visited blocks
    at w = maxW, 8 k visible blocks - are compared at full size, every stage of every frame against the real reference"""
    rows = _run(640, 480, 100)
    assert all(r["hash_equal"] and r["visible_equal"] and r["view_equal"] and r["minmax_equal"] for r in rows)
    assert max(r["voxel_max_dsdf"] for r in rows) <= 1 and max(r["voxel_max_dw"] for r in rows) <= 1
    assert max(r["raycast_hit_mismatch"] for r in rows) == 0 and max(r["raycast_max_diff_m"] for r in rows) <= 1e-4
    assert max(r["pose_rot_rad"] for r in rows) <= 1e-4 and max(r["pose_trans_m"] for r in rows) <= 1e-4
    assert rows[-1]["counters_ref"][0] > 7500  # the late frames really are the heavy ones
