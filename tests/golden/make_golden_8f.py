#!/usr/bin/env python
"""Generates tests/golden/ref_rows8f_qqvga.npz from the REAL reference CPU engines (oracle/_ref/libitm_ref.so): golden
vectors for the SURVEY.md 8f rows - MeshScene, the free-view GetImage chain (FindVisibleBlocks + CreateExpectedDepths +
RenderImage) and ForwardRender.

Run in the development container (the GPU box has no /root/reference):
    python tests/golden/make_golden_8f.py

Workload: frame 0 of the 160x120 golden sequence (tests/golden/make_golden.py).  Frame 0 is fused at the identity pose on
every implementation, so the scene the three rows read is bit-identical everywhere and the vectors can be exact:
  mesh      noTotalTriangles, CRC32 of the triangle array, the first and last 32 triangles
  free view per render type (grey, normal): the camera, visible-list CRC + length, min/max CRC, raycast CRC, image CRC,
            a sparse sample of the image
  forward   camera moved by a few millimetres without a new raycast: CRC of forwardProjection, CRC of the sorted
            missing-point list + its length, CRC of the forward-rendered raycastImage
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


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def free_pose(k):
    M = np.eye(4, dtype=np.float32)
    a = np.float32(np.deg2rad(6.0 + 4.0 * k))
    M[0, 0], M[0, 2], M[2, 0], M[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    M[:3, 3] = [0.1, -0.04 + 0.02 * k, 0.06]
    return M.T.reshape(16).astype(np.float32)   # column-major


def moved_pose(M16):
    m = np.array(M16, dtype=np.float32, copy=True)
    m[12] += np.float32(0.006)
    m[13] -= np.float32(0.003)
    m[14] += np.float32(0.002)
    return m


def main():
    seq = synth.sequence(1, W, H, noise=True)
    e = ref.RefEngine(W, H)
    e.process_frame(seq[0])
    out = {"W": W, "H": H, "depth_crc": np.uint64(crc(seq[0]))}
    tri = np.array(e.mesh_scene(), copy=True)
    out["mesh_n"] = np.int64(len(tri))
    out["mesh_crc"] = np.uint64(crc(tri))
    out["mesh_head"], out["mesh_tail"] = tri[:32].copy(), tri[-32:].copy()
    K = np.array(synth.intrinsics_for(W, H), dtype=np.float32)
    for k, t in enumerate((3, 5)):   # InfiniTAM_IMAGE_FREECAMERA_SHADED, ..._COLOUR_FROM_NORMAL
        M = free_pose(k)
        img = e.get_image(t, M, K, W, H)
        out["free%d_type" % k], out["free%d_pose" % k], out["free%d_intr" % k] = np.int64(t), M, K
        out["free%d_nvis" % k] = np.int64(len(e.free_visible_ids))
        out["free%d_crc" % k] = np.array([crc(e.free_visible_ids), crc(e.free_minmax), crc(e.free_raycast_result), crc(img)], dtype=np.uint64)
        out["free%d_sample" % k] = img[::5, ::7].copy()
    pose = moved_pose(e.pose_M)
    e.pose_M = pose
    e.expected_depths()
    e.forward_render()
    out["fwd_pose"] = pose
    out["fwd_crc"] = np.array([crc(e.forward_projection), crc(np.sort(e.fwd_missing_points)), crc(e.raycast_image), crc(e.minmax)], dtype=np.uint64)
    out["fwd_nmissing"] = np.int64(len(e.fwd_missing_points))
    out["fwd_nvalid"] = np.int64(np.count_nonzero(e.forward_projection[..., 3] > 0))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_rows8f_qqvga.npz"), **out)
    print({k: (v.shape if getattr(v, "ndim", 0) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
