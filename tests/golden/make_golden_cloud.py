#!/usr/bin/env python
"""Generates tests/golden/ref_cloud_lowlevel_qqvga.npz from the REAL reference CPU engines (oracle/_ref/libitm_ref.so): golden
vectors for IITMVisualisationEngine::CreatePointCloud (as ITMTrackingController::Prepare calls it for TRACKER_COLOR) and for
the ITMLowLevelEngine helpers only the colour / Ren trackers use.

Run in the development container (the GPU box has no /root/reference):
    python tests/golden/make_golden_cloud.py

Workload: frame 0 of the 160x120 golden sequence, fused at the identity pose (bit-identical scene everywhere).
  cloud     for the identity and for a small rgb-to-depth calibration, with and without skipPoints: noTotalPoints, CRC32 of
            the locations, of the colours and of the shaded raycast image, the first 16 locations
  lowlevel  CopyImage, FilterSubsample, FilterSubsampleWithHoles(Vector4f), GradientX, GradientY on seeded random 162x122
            images, the output image pre-filled with 0x5A bytes (the gradient drivers clear only part of it): CRC32 per output
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from infinitam_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

W, H = 160, 120
LW, LH = 162, 122
PREFILL = 0x5A


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def trafo():
    a = np.float32(np.deg2rad(2.0))
    T = np.eye(4, dtype=np.float32)
    T[0, 0], T[0, 2], T[2, 0], T[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    T[:3, 3] = [0.025, -0.01, 0.005]
    return np.ascontiguousarray(T.T).reshape(16).astype(np.float32)  # ITMExtrinsics::calib, column-major


def lowlevel_inputs():
    rng = np.random.default_rng(11)
    rgba = rng.integers(0, 256, size=(LH, LW, 4), dtype=np.uint8)
    f4 = rng.normal(size=(LH, LW, 4)).astype(np.float32)
    f4[..., 3] = np.where(rng.random((LH, LW)) < 0.3, -1.0, 1.0).astype(np.float32)
    f4[:20, :20, 3] = -1.0
    return rgba, f4


def main():
    seq = synth.sequence(1, W, H, noise=True)
    e = ref.RefEngine(W, H)
    e.process_frame(seq[0])
    out = {"W": W, "H": H, "depth_crc": np.uint64(crc(seq[0])), "trafo": trafo(), "LW": LW, "LH": LH, "prefill": PREFILL}
    k = 0
    for T in (None, trafo()):
        for skip in (False, True):
            loc, clr = e.create_point_cloud(T, skip_points=skip)
            out["cloud%d_use_trafo" % k], out["cloud%d_skip" % k] = np.int64(T is not None), np.int64(skip)
            out["cloud%d_n" % k] = np.int64(len(loc))
            out["cloud%d_crc" % k] = np.array([crc(loc), crc(clr), crc(e.raycast_image)], dtype=np.uint64)
            out["cloud%d_head" % k] = loc[:16].copy()
            k += 1
    rgba, f4 = lowlevel_inputs()
    out["lowlevel_in_crc"] = np.array([crc(rgba), crc(f4)], dtype=np.uint64)
    out["lowlevel_crc"] = np.array([crc(e.low_level(op, f4 if op == 2 else rgba, prefill=PREFILL)) for op in range(5)], dtype=np.uint64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cloud_lowlevel_qqvga.npz"), **out)
    print({k: (v.shape if getattr(v, "ndim", 0) else v) for k, v in out.items()})
    e.close()


if __name__ == "__main__":
    main()
