"""CPU (-m "not gpu"): pins the oracle.

1. the C restatement (oracle/itm_oracle.c) against the golden vectors generated from the real reference;
2. when oracle/_ref/libitm_ref.so exists (built from /root/reference in the dev container), the restatement against the
   real reference engines, stage by stage, bit for bit.
"""
import numpy as np
import pytest

import golden_check
from infinitam_b200 import synth
from oracle import port, ref


def _state(e):
    return dict(pose=e.pose_M, counters=e.counters, hash_entries=e.hash_entries, visible_ids=e.visible_ids, voxels_u32=e.voxels,
                minmax=e.minmax, raycast=e.raycast_result, points=e.points, normals=e.normals, image=e.raycast_image,
                visible_types=e.visible_types, depth=e.depth)


def test_port_reproduces_golden_vectors():
    g = golden_check.load()
    seq = golden_check.golden_sequence(g)
    e = port.PortEngine(int(g["W"]), int(g["H"]))
    for k in range(int(g["N"])):
        e.update_view(seq[k])
        if k == 1:
            e.icp_prepare()
            inv = e.mat_inv(e.pose_M)
            for lvl in range(5):
                _, o = e.icp_gandh(lvl, inv)
                assert np.array_equal(o, g["f1_gandh_l%d" % lvl]), "ComputeGandH level %d differs from golden" % lvl
        e.track()
        e.allocate()
        e.integrate()
        e.expected_depths()
        e.icp_maps()
        golden_check.check_frame(g, k, **_state(e))
    e.close()


def _eq_engines(a, b, what=""):
    ha, hb = a.hash_entries, b.hash_entries
    assert np.array_equal(ha["pos"], hb["pos"]) and np.array_equal(ha["ptr"], hb["ptr"]) and np.array_equal(ha["offset"], hb["offset"]), what + " hash"
    assert np.array_equal(a.counters, b.counters), what + " counters"
    assert np.array_equal(a.visible_ids[: a.counters[0]], b.visible_ids[: b.counters[0]]), what + " visible ids"
    assert np.array_equal(a.visible_types, b.visible_types), what + " visible types"
    assert np.array_equal(a.voxels & 0x00FFFFFF, b.voxels & 0x00FFFFFF), what + " voxels"
    assert np.array_equal(a.pose_M, b.pose_M), what + " pose"


@pytest.mark.parametrize("size,voxel,noise", [((320, 240), 0.005, True), ((160, 120), 0.0025, False)])
def test_port_matches_real_reference_stage_by_stage(size, voxel, noise):
    if not ref.available("parity"):
        pytest.skip("oracle/_ref/libitm_ref.so not present (needs /root/reference to build)")
    W, H = size
    a = ref.RefEngine(W, H, voxel_size=voxel)
    b = port.PortEngine(W, H, voxel_size=voxel)
    seq = synth.sequence(4, W, H, noise=noise)
    for k in range(4):
        a.update_view(seq[k]); b.update_view(seq[k])
        assert np.array_equal(a.depth, b.depth)
        a.icp_prepare(); b.icp_prepare()
        for lvl in range(1, 5):
            assert np.array_equal(a.pyramid_level(lvl)[0], b.pyramid_level(lvl)[0]), "pyramid level %d" % lvl
            assert np.array_equal(a.pyramid_level(lvl)[1], b.pyramid_level(lvl)[1])
        if k > 0:
            inv = a.mat_inv(a.pose_M)
            assert np.array_equal(inv, b.mat_inv(b.pose_M))
            for lvl in range(5):
                na, oa = a.icp_gandh(lvl, inv)
                nb, ob = b.icp_gandh(lvl, inv)
                assert na == nb and np.array_equal(oa, ob), "ComputeGandH level %d" % lvl
        a.track(); b.track()
        assert np.array_equal(a.pose_M, b.pose_M) and np.array_equal(a.pose_params, b.pose_params), "frame %d pose" % k
        a.allocate(); b.allocate()
        _eq_engines(a, b, "frame %d after allocate:" % k)
        a.integrate(); b.integrate()
        _eq_engines(a, b, "frame %d after integrate:" % k)
        a.expected_depths(); b.expected_depths()
        assert np.array_equal(a.minmax, b.minmax)
        a.icp_maps(); b.icp_maps()
        assert np.array_equal(a.raycast_result, b.raycast_result)
        assert np.array_equal(a.points, b.points) and np.array_equal(a.normals, b.normals)
        assert np.array_equal(a.raycast_image, b.raycast_image)
        assert np.array_equal(a.pose_pointcloud_M, b.pose_pointcloud_M) and a.age == b.age
    a.close(); b.close()


def test_icp_config_matches_reference_constructor():
    """ITMDepthTracker constructor: iterations {2,4,6,8,10}, distThresh 0.002 .. 0.01 (ITMDepthTracker.cpp:19-28)"""
    b = port.PortEngine(64, 48)
    n, iters, thr, typ = b.icp_config()
    assert n == 5 and list(iters) == [2, 4, 6, 8, 10] and list(typ) == [3, 3, 1, 1, 1]
    if ref.available("parity"):
        a = ref.RefEngine(64, 48)
        n2, iters2, thr2, typ2 = a.icp_config()
        assert n2 == n and np.array_equal(iters, iters2) and np.array_equal(thr, thr2) and np.array_equal(typ, typ2)
        a.close()
    b.close()


def test_pose_math_matches_reference():
    rng = np.random.default_rng(7)
    if not ref.available("parity"):
        pytest.skip("needs oracle/_ref")
    a = ref.RefEngine(64, 48)
    b = port.PortEngine(64, 48)
    for _ in range(200):
        p = np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.4, 3)]).astype(np.float32)
        M = np.zeros(16, np.float32)
        a.lib.ref_pose_from_params(p.ctypes.data_as(ref._f32p), M.ctypes.data_as(ref._f32p))
        M2 = np.zeros(16, np.float32)
        b.lib.ref_pose_from_params(p.ctypes.data_as(ref._f32p), M2.ctypes.data_as(ref._f32p))
        assert np.array_equal(M, M2)
        inv = a.mat_inv(M)
        assert np.array_equal(inv, b.mat_inv(M))
        ra, rb = a.pose_from_invm_coerced(inv), b.pose_from_invm_coerced(inv)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    a.close(); b.close()


# ---- SURVEY 8f rows of the restatement: ForwardRender / useApproximateRaycast, free-view rendering, meshing -----------------

def _free_pose(k):
    M = np.eye(4, dtype=np.float32)
    a = np.float32(np.deg2rad(6.0 + 4.0 * k))
    M[0, 0], M[0, 2], M[2, 0], M[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    M[:3, 3] = [0.1, -0.04 + 0.02 * k, 0.06]
    return M.T.reshape(16).astype(np.float32)


@pytest.mark.skipif(not ref.available("parity"), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_port_rows_8f_equal_the_reference():
    """approximate-raycast sequence (the full / approximate decision, forward projection, missing-point list, forward-rendered
    image), free-view images with their visible lists / ranges / raycasts, and the mesh: bit for bit"""
    w, h, n = 160, 120, 8
    seq = synth.sequence(n, w, h, noise=True)
    r, p = ref.RefEngine(w, h), port.PortEngine(w, h)
    r.set_use_approximate_raycast(True)
    p.set_use_approximate_raycast(True)
    n_fwd = 0
    for k in range(n):
        r.process_frame(seq[k])
        p.process_frame(seq[k])
        assert np.array_equal(p.pose_M, r.pose_M), "frame %d pose" % k
        assert p.requires_full_rendering == r.requires_full_rendering and p.age == r.age
        assert np.array_equal(p.raycast_image, r.raycast_image), "frame %d raycast image" % k
        if not r.requires_full_rendering:
            n_fwd += 1
            assert np.array_equal(p.fwd_missing_points, r.fwd_missing_points)
            assert np.array_equal(p.forward_projection, r.forward_projection)
    assert n_fwd >= 2
    K = np.array(synth.intrinsics_for(w, h), np.float32)
    for k, t in enumerate((3, 5)):
        img_r, img_p = r.get_image(t, _free_pose(k), K, w, h), p.get_image(t, _free_pose(k), K, w, h)
        assert np.array_equal(p.free_visible_ids, r.free_visible_ids)
        assert np.array_equal(p.free_minmax, r.free_minmax)
        assert np.array_equal(p.free_raycast_result, r.free_raycast_result)
        assert img_r.any() and np.array_equal(img_p, img_r), "free-view image type %d" % t
    tri_r, tri_p = r.mesh_scene(), p.mesh_scene()
    assert len(tri_r) > 10000 and tri_p.shape == tri_r.shape
    assert np.array_equal(tri_p.view(np.uint32), tri_r.view(np.uint32))
    r.close(); p.close()


def test_port_rows_8f_reproduce_golden_vectors():
    """the restatement against tests/golden/ref_rows8f_qqvga.npz (made from the real reference): runs everywhere"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_rows8f_qqvga.npz"))
    w, h = int(g["W"]), int(g["H"])
    seq = synth.sequence(1, w, h, noise=True)
    assert golden_check.crc(seq[0]) == int(g["depth_crc"])
    p = port.PortEngine(w, h)
    p.process_frame(seq[0])
    tri = p.mesh_scene()
    assert len(tri) == int(g["mesh_n"]) and golden_check.crc(tri) == int(g["mesh_crc"])
    for k in range(2):
        img = p.get_image(int(g["free%d_type" % k]), g["free%d_pose" % k], g["free%d_intr" % k], w, h)
        c = [int(x) for x in g["free%d_crc" % k]]
        assert len(p.free_visible_ids) == int(g["free%d_nvis" % k])
        assert [golden_check.crc(p.free_visible_ids), golden_check.crc(p.free_minmax), golden_check.crc(p.free_raycast_result),
                golden_check.crc(img)] == c
    p.pose_M = g["fwd_pose"]
    p.expected_depths()
    p.forward_render()
    c = [int(x) for x in g["fwd_crc"]]
    assert len(p.fwd_missing_points) == int(g["fwd_nmissing"])
    assert [golden_check.crc(p.forward_projection), golden_check.crc(np.sort(p.fwd_missing_points)), golden_check.crc(p.raycast_image),
            golden_check.crc(p.minmax)] == c
    p.close()


@pytest.mark.skipif(not ref.available("parity"), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("bilateral", [False, True], ids=["sensor-noise-model", "bilateral-filter+sensor-noise-model"])
def test_port_weighted_icp_equals_the_reference(bilateral):
    """filterDepth (5 passes), computeNormalAndWeight, the weight pyramid, single weighted evaluations and
    ITMWeightedICPTracker's Gauss-Newton loop - bit for bit (both sides call glibc's expf / acosf)"""
    w, h, n = 160, 120, 6
    seq = synth.sequence(n, w, h, noise=True)
    r, p = ref.RefEngine(w, h, wicp=True, bilateral=bilateral), port.PortEngine(w, h, wicp=True, bilateral=bilateral)
    for k in range(n):
        r.update_view(seq[k])
        p.update_view(seq[k])
        assert np.array_equal(p.depth, r.depth), "frame %d filtered depth" % k
        assert np.array_equal(p.depth_uncertainty, r.depth_uncertainty, equal_nan=True)
        assert np.array_equal(p.depth_normal, r.depth_normal)
        if k == 2:
            r.wicp_prepare(); p.wicp_prepare()
            inv = r.mat_inv(r.pose_M)
            for level in range(5):
                n_r, g_r = r.wicp_gandh(level, inv)
                n_p, g_p = p.wicp_gandh(level, inv)
                assert n_r == n_p and np.array_equal(g_r, g_p), "level %d" % level
        for eng in (r, p):
            eng.track(); eng.allocate(); eng.integrate(); eng.expected_depths(); eng.icp_maps()
        assert np.array_equal(p.pose_M, r.pose_M), "frame %d pose" % k
    assert np.abs(r.pose_M - np.eye(4, dtype=np.float32).reshape(16)).max() > 0.01
    r.close(); p.close()


def test_port_point_cloud_and_low_level_helpers_reproduce_golden_vectors():
    """the restatement's CreatePointCloud (TRACKER_COLOR branch of Prepare) and ITMLowLevelEngine helpers against
    tests/golden/ref_cloud_lowlevel_qqvga.npz (made from the real reference by tests/golden/make_golden_cloud.py): runs everywhere"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_cloud as mk
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cloud_lowlevel_qqvga.npz"))
    w, h = int(g["W"]), int(g["H"])
    seq = synth.sequence(1, w, h, noise=True)
    assert golden_check.crc(seq[0]) == int(g["depth_crc"])
    p = port.PortEngine(w, h)
    p.process_frame(seq[0])
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for k in range(4):
        T = g["trafo"] if int(g["cloud%d_use_trafo" % k]) else ident
        loc, clr = p.create_point_cloud(T, skip_points=bool(int(g["cloud%d_skip" % k])))
        assert len(loc) == int(g["cloud%d_n" % k])
        assert np.array_equal(loc[:16], g["cloud%d_head" % k])
        assert [golden_check.crc(loc), golden_check.crc(clr), golden_check.crc(p.raycast_image)] == [int(x) for x in g["cloud%d_crc" % k]]
    rgba, f4 = mk.lowlevel_inputs()
    assert [golden_check.crc(rgba), golden_check.crc(f4)] == [int(x) for x in g["lowlevel_in_crc"]]
    got = [golden_check.crc(p.low_level(op, f4 if op == 2 else rgba, prefill=int(g["prefill"]))) for op in range(5)]
    assert got == [int(x) for x in g["lowlevel_crc"]]
    p.close()
