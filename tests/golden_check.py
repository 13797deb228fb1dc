"""Compares an engine's state with the committed golden vectors (tests/golden/ref_qqvga.npz, produced from the real
reference CPU engines by tests/golden/make_golden.py)."""
import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_qqvga.npz")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def load():
    return np.load(GOLDEN)


def golden_sequence(g):
    from infinitam_b200 import synth
    W, H, N = int(g["W"]), int(g["H"]), int(g["N"])
    seq = synth.sequence(N, W, H, noise=True)
    assert [crc(seq[k]) for k in range(N)] == [int(x) for x in g["depth_crc"]], "synthetic generator no longer reproduces the golden input"
    return seq


def check_frame(g, k, *, pose, counters, hash_entries, visible_ids, voxels_u32, minmax, raycast, points, normals, image,
                visible_types, depth, exact_pose=True, exact_maps=True):
    """All arrays are host numpy arrays in the reference's layouts."""
    assert list(counters[:3]) == list(g["f%d_counters" % k]), "counters: %s vs golden %s" % (counters[:3], g["f%d_counters" % k])
    live = np.nonzero(hash_entries["ptr"] >= -1)[0].astype(np.int32)
    assert np.array_equal(live, g["f%d_hash_slots" % k]), "set of occupied hash slots differs"
    assert np.array_equal(hash_entries["pos"][live], g["f%d_hash_pos" % k])
    assert np.array_equal(hash_entries["ptr"][live], g["f%d_hash_ptr" % k])
    assert np.array_equal(hash_entries["offset"][live], g["f%d_hash_offset" % k])
    n = int(counters[0])
    assert np.array_equal(np.sort(visible_ids[:n]), np.sort(g["f%d_visible" % k])), "visible set differs"
    c = [int(x) for x in g["f%d_crc" % k]]
    assert crc(depth) == c[7], "depth image differs"
    assert crc(visible_types) == c[6], "entriesVisibleType differs"
    assert crc(voxels_u32 & 0x00FFFFFF) == c[0], "voxel array differs"
    assert crc(minmax) == c[1], "expected-depth image differs"
    if exact_pose:
        assert np.array_equal(pose, g["f%d_pose" % k]), "pose differs"
    else:
        assert np.abs(pose - g["f%d_pose" % k]).max() <= 1e-4
    if exact_maps:
        assert crc(raycast) == c[2] and crc(points) == c[3] and crc(normals) == c[4] and crc(image) == c[5], "raycast / ICP maps differ"
    else:
        rs = g["f%d_raycast_sample" % k]
        mine = raycast[::7, ::9]
        assert np.array_equal(mine[..., 3], rs[..., 3]), "raycast hit mask differs"
        hit = rs[..., 3] > 0
        assert np.abs(mine[..., :3] - rs[..., :3])[hit].max() * 0.005 <= 1e-4
        ps = g["f%d_points_sample" % k]
        assert np.array_equal(points[::7, ::9][..., 3], ps[..., 3])
        assert np.abs(points[::7, ::9] - ps).max() <= 1e-4
