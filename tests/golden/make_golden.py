#!/usr/bin/env python
"""Generates tests/golden/ref_qqvga.npz from the REAL reference CPU engines (oracle/_ref/libitm_ref.so,
built by oracle/build_ref.py from the unmodified sources under /root/reference).

Run in the development container (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The fixture pins the oracle port (tests/test_oracle_port.py) and the CUDA path (tests/test_gpu_golden.py)
where the reference library itself is not available.

Workload: 160x120 synthetic sequence (noise + dropped pixels), 4 frames, default scene parameters.
Stored per frame: pose, counters, the non-empty hash entries, the visible list, CRC32s of the voxel array
(padding byte masked), of the expected-depth image and of the raycast / ICP maps, plus a sparse sample of
raycast pixels, and one single ICP evaluation (ComputeGandH) per pyramid level on frame 1.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from infinitam_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

W, H, N = 160, 120, 4


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def main():
    seq = synth.sequence(N, W, H, noise=True)
    e = ref.RefEngine(W, H)
    out = {"W": W, "H": H, "N": N, "depth_crc": np.array([crc(seq[k]) for k in range(N)], dtype=np.uint64)}
    for k in range(N):
        e.update_view(seq[k])
        if k == 1:
            # single evaluations at the pre-tracking pose, every level
            e.icp_prepare()
            inv = e.mat_inv(e.pose_M)
            for lvl in range(5):
                n, o = e.icp_gandh(lvl, inv)
                out["f1_gandh_l%d" % lvl] = o.copy()
        e.track()
        e.allocate()
        e.integrate()
        e.expected_depths()
        e.icp_maps()
        h = e.hash_entries
        live = np.nonzero(h["ptr"] >= -1)[0].astype(np.int32)
        out["f%d_pose" % k] = e.pose_M.copy()
        out["f%d_counters" % k] = e.counters.copy()
        out["f%d_hash_slots" % k] = live
        out["f%d_hash_pos" % k] = h["pos"][live].copy()
        out["f%d_hash_ptr" % k] = h["ptr"][live].copy()
        out["f%d_hash_offset" % k] = h["offset"][live].copy()
        out["f%d_visible" % k] = e.visible_ids[: e.counters[0]].copy()
        out["f%d_crc" % k] = np.array([crc(e.voxels & 0x00FFFFFF), crc(e.minmax), crc(e.raycast_result), crc(e.points), crc(e.normals),
                                         crc(e.raycast_image), crc(e.visible_types), crc(e.depth)], dtype=np.uint64)
        out["f%d_raycast_sample" % k] = e.raycast_result[::7, ::9].copy()
        out["f%d_points_sample" % k] = e.points[::7, ::9].copy()
        out["f%d_normals_sample" % k] = e.normals[::7, ::9].copy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_qqvga.npz"), **out)
    print("written", {k: getattr(v, "shape", v) for k, v in list(out.items())[:12]})


if __name__ == "__main__":
    main()
