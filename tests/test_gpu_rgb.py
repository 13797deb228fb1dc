"""-m gpu: ITMVoxel_s_rgb (BASELINE configs[4]): 8-byte voxels, colour integration with bilinear view->rgb fetch
(computeUpdatedVoxelColorInfo), raycasting over the wider voxel - every stage teacher-forced against the reference CPU
engines instantiated for ITMVoxel_s_rgb (oracle/_ref/libitm_ref_rgb.so)."""
import numpy as np
import pytest

import parity
from infinitam_b200 import capi, synth
from oracle import ref

pytestmark = pytest.mark.gpu


def _rgb_frame(k, w, h):
    yy, xx = np.mgrid[0:h, 0:w]
    return np.stack([(xx * 7 + 13 * k) & 255, (yy * 5 + 3 * k) & 255, ((xx ^ yy) * 3) & 255, np.full_like(xx, 255)], -1).astype(np.uint8)


@pytest.mark.skipif(not ref.available("rgb"), reason="oracle/_ref/libitm_ref_rgb.so not built (needs /root/reference at build time)")
def test_colour_voxels_match_reference():
    w, h, n = 320, 240, 4
    o = ref.RefEngine(w, h, flavour="rgb")
    assert o.const("sizeof_voxel") == 8 and o.const("has_color") == 1
    eng = parity.make_cuda_engine(o)
    seq = synth.sequence(n, w, h)
    for k in range(n):
        rgb = _rgb_frame(k, w, h)
        o.set_rgb(rgb)
        eng.write(capi.BUF_RGB, rgb)
        r = parity.compare_frame(o, eng, seq[k], k, strict=True)
        assert r["voxel_max_dsdf"] <= 1 and r["voxel_max_dclr"] <= 1 and r["voxel_max_dwcolor"] == 0
    v = eng.read(capi.BUF_VOXELS)
    assert int(np.count_nonzero((v >> np.uint64(48)) & np.uint64(0xFF))) > 100000  # colour really was integrated
    eng.close(); o.close()


@pytest.mark.skipif(not ref.available("rgb"), reason="oracle/_ref/libitm_ref_rgb.so not built")
def test_colour_process_frame_host_api():
    """ProcessFrame(rgb, depth) end to end (rgb upload ordered before integration), free running, 3 frames"""
    w, h = 320, 240
    o = ref.RefEngine(w, h, flavour="rgb")
    eng = parity.make_cuda_engine(o)
    seq = synth.sequence(3, w, h)
    for k in range(3):
        rgb = _rgb_frame(k, w, h)
        o.set_rgb(rgb)
        o.process_frame(seq[k])
        pose = eng.ProcessFrame(rgb, seq[k])
        rot, trans = parity.pose_diff(pose, o.pose_M)
        assert rot <= 1e-4 and trans <= 1e-4
    # free running: the two poses agree to ~1e-6, so a voxel sitting exactly on the colour gate (|eta / mu| <= 0.25) or on the
    # image border may be coloured on one side only - everywhere else the colours must agree
    a, b = eng.read(capi.BUF_VOXELS), o.voxels
    touched = ((b >> np.uint64(48)) & np.uint64(0xFF)) > 0
    assert touched.sum() > 100000
    bad = np.zeros(a.shape, bool)
    for shift in (24, 32, 40):
        ca = ((a >> np.uint64(shift)) & np.uint64(0xFF)).astype(np.int32)
        cb = ((b >> np.uint64(shift)) & np.uint64(0xFF)).astype(np.int32)
        bad |= np.abs(ca - cb) > 2
    assert bad.sum() <= 2e-3 * touched.sum()
    eng.close(); o.close()
