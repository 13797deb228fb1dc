"""-m gpu: the drop-in boundary end to end.  The reference's own host objects (ITMScene, ITMRenderState_VH,
ITMTrackingState, ITMView, ITMTrackingController, ITMPose - compiled from the reference sources with CUDA memory
placement) run ITMMainEngine::ProcessFrame through include/itm_b200_adapter.hpp -> libitm_b200.so, and are compared
with the reference's CPU engines on the same frames (oracle/adapter_harness.cpp vs oracle/ref_harness.cpp)."""
import numpy as np
import pytest

import parity
from infinitam_b200 import synth
from oracle import adapter, ref

pytestmark = pytest.mark.gpu

needs_libs = pytest.mark.skipif(not (adapter.available() and ref.available("parity")),
                                reason="oracle/_ref not built (needs /root/reference at build time)")


def _exact_scene_checks(o, a):
    """hash table and visible list bit exact, voxels within 1 LSB, raycast / ICP maps within 1e-4 m"""
    assert [int(x) for x in a.counters] == [int(x) for x in o.counters] + [int(o.age)]
    assert parity.hash_equal(a.read(adapter.READ_HASH), o.hash_entries)
    n_vis = int(o.counters[0])
    assert np.array_equal(np.sort(a.read(adapter.READ_VISIBLE_IDS)[:n_vis]), np.sort(o.visible_ids[:n_vis]))
    assert np.array_equal(a.read(adapter.READ_VISIBLE_TYPES), o.visible_types)
    ds, dw, _, _ = parity.voxel_diff(a.read(adapter.READ_VOXELS), o.voxels)
    assert ds <= 1 and dw <= 1
    ray_a, ray_o = a.read(adapter.READ_RAYCAST).reshape(o.H, o.W, 4), o.raycast_result
    assert np.array_equal(ray_a[..., 3] > 0, ray_o[..., 3] > 0)
    assert np.abs(ray_a[..., :3] - ray_o[..., :3]).max() * o.voxel_size <= 1e-4
    pts_a, nrm_a = a.read(adapter.READ_POINTS).reshape(o.H, o.W, 4), a.read(adapter.READ_NORMALS).reshape(o.H, o.W, 4)
    assert np.array_equal(pts_a[..., 3], o.points[..., 3])
    assert np.abs(pts_a - o.points).max() <= 1e-4 and np.abs(nrm_a - o.normals).max() <= 1e-3


@needs_libs
@pytest.mark.parametrize("device_loop", [True, False], ids=["device-LM-loop", "reference-host-LM-loop+ComputeGandH"])
def test_reference_host_objects_with_b200_engines(device_loop):
    w, h, n = 320, 240, 6
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h)
    a = adapter.AdapterEngine(w, h, intr=o.intr, device_loop=device_loop)
    # frame 0 runs at the identity pose on both sides: everything downstream of it must agree exactly
    o.process_frame(seq[0])
    a.process_frame(seq[0])
    assert np.array_equal(a.pose_M, o.pose_M)
    _exact_scene_checks(o, a)
    # free running from here on: each side tracks against its own maps
    for k in range(1, n):
        o.process_frame(seq[k])
        a.process_frame(seq[k])
        rot, trans = parity.pose_diff(a.pose_M, o.pose_M)
        assert rot <= 1e-4 and trans <= 1e-4, "frame %d pose differs: %g rad %g m" % (k, rot, trans)
        ca, co = a.counters, o.counters
        assert abs(int(ca[0]) - int(co[0])) <= 0.01 * co[0] and abs(int(ca[1]) - int(co[1])) <= 0.01 * (o.n_local - co[1])
        assert int(ca[3]) == int(o.age)
    a.close(); o.close()


@needs_libs
def test_adapter_approximate_raycast_and_free_view():
    """settings.useApproximateRaycast through the reference's own ITMTrackingController (Track decides on the host, Prepare
    calls the adapter's CreateICPMaps or ForwardRender), then a free-view rendering through FindVisibleBlocks +
    CreateExpectedDepths + RenderImage of the adapter - against the reference CPU engines doing the same."""
    w, h, n = 320, 240, 12
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h)
    o.set_use_approximate_raycast(True)
    a = adapter.AdapterEngine(w, h, intr=o.intr)
    a.set_use_approximate_raycast(True)
    n_fwd = 0
    for k in range(n):
        o.process_frame(seq[k])
        a.process_frame(seq[k])
        rot, trans = parity.pose_diff(a.pose_M, o.pose_M)
        assert rot <= 1e-4 and trans <= 1e-4, "frame %d pose differs: %g rad %g m" % (k, rot, trans)
        assert a.requires_full_rendering == o.requires_full_rendering
        assert int(a.counters[3]) == int(o.age)
        n_fwd += 0 if o.requires_full_rendering else 1
        img_a, img_o = a.read(adapter.READ_RAYCAST_IMAGE).reshape(h, w, 4), o.raycast_image
        bad = np.count_nonzero(np.abs(img_a.astype(np.int32) - img_o.astype(np.int32)).max(axis=2) > 1)
        assert bad <= 0.002 * w * h, "frame %d: %d raycast-image pixels differ" % (k, bad)
    assert n_fwd > 0
    M = np.eye(4, dtype=np.float32)
    M[:3, 3] = [0.1, -0.04, 0.06]
    M = M.T.reshape(16)
    for rt in (0, 2):
        img_a = a.get_free_image(rt, M, o.intr, w, h)
        img_o = o.get_image(3 if rt == 0 else 5, M, o.intr, w, h)
        bad = np.count_nonzero(np.abs(img_a.astype(np.int32) - img_o.astype(np.int32)).max(axis=2) > 1)
        assert img_o.any() and bad <= 0.002 * w * h, "free-view type %d: %d pixels differ" % (rt, bad)
    # the colour tracker's Prepare branch (CreateExpectedDepths at the colour camera + CreatePointCloud) through the adapter;
    # the two scenes agree to the free-running tolerance only, so compare counts and the clouds as sets
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.025, -0.01, 0.005]  # ITMExtrinsics::calib of the colour camera, column-major below
    T = np.ascontiguousarray(T.T).reshape(16)
    for skip in (False, True):
        loc_a, clr_a = a.create_point_cloud(T, skip)
        loc_o, clr_o = o.create_point_cloud(T, skip)
        assert len(loc_o) > 2000 and abs(len(loc_a) - len(loc_o)) <= 0.002 * len(loc_o), "point cloud sizes: %d vs %d" % (len(loc_a), len(loc_o))
        assert np.all(loc_a[:, 3] == 1.0) and not clr_a.any()
        assert np.abs(loc_a[:, :3].mean(axis=0) - loc_o[:, :3].mean(axis=0)).max() <= 1e-3
        assert np.abs(loc_a[:, :3].min(axis=0) - loc_o[:, :3].min(axis=0)).max() <= 2e-2
    a.close(); o.close()


@needs_libs
def test_adapter_swapping_engine():
    """settings.useSwapping through the reference's own ITMScene / ITMGlobalCache (CUDA memory placement) and the adapter's
    ITMSwappingEngine_B200 + swap-aware AllocateSceneFromDepth, against the reference CPU engines with swapping: a camera
    that leaves and re-enters its first view, so blocks are parked on the host and merged back"""
    w, h = 320, 240
    o = ref.RefEngine(w, h, use_swapping=True)
    a = adapter.AdapterEngine(w, h, intr=o.intr, use_swapping=True)
    frames = list(range(0, 60, 2)) + list(range(58, -1, -2))   # the same inter-frame motion as the Layer B swapping test
    for i, k in enumerate(frames):
        depth = synth.render_depth(k, w, h)
        o.process_frame(depth)
        a.process_frame(depth)
        rot, trans = parity.pose_diff(a.pose_M, o.pose_M)
        assert rot <= 1e-4 and trans <= 1e-4, "frame %d (%d) pose differs: %g rad %g m" % (i, k, rot, trans)
        if i == 0:   # identity pose on both sides: exact
            assert [int(x) for x in a.counters[:3]] == [int(x) for x in o.counters]
            assert parity.hash_equal(a.read(adapter.READ_HASH), o.hash_entries)
    n_ref, n_adp = int(o.has_stored_data.sum()), a.stored_blocks
    assert n_ref > 100, "the trajectory should swap blocks out (reference parked %d)" % n_ref
    assert abs(n_adp - n_ref) <= 0.02 * n_ref + 2, "blocks parked on the host: adapter %d reference %d" % (n_adp, n_ref)
    ca, co = a.counters, o.counters
    assert abs(int(ca[1]) - int(co[1])) <= 0.02 * (o.n_local - co[1]) + 2
    a.close(); o.close()


@needs_libs
def test_adapter_low_level_helpers_and_view_variants():
    """The rest of ITMLowLevelEngine (CopyImage, FilterSubsample, FilterSubsampleWithHoles(Vector4f), GradientX / GradientY:
    ITMLowLevelEngine_CPU.cpp:12-108) and of ITMViewBuilder (float-depth and IMU UpdateView) through the adapter, bit for bit
    against the reference CPU engine - including the gradient drivers' partial clear of the output image."""
    w, h = 322, 242  # newDims = 161 x 121: odd sizes inside the 2x2 quads' reach
    rng = np.random.default_rng(7)
    o = ref.RefEngine(w, h)
    a = adapter.AdapterEngine(w, h, intr=o.intr)
    rgba = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    f4 = rng.normal(size=(h, w, 4)).astype(np.float32)
    f4[..., 3] = np.where(rng.random((h, w)) < 0.3, -1.0, 1.0)  # holes
    f4[:40, :40, 3] = -1.0  # whole quads without a valid tap
    for op, img in ((0, rgba), (1, rgba), (2, f4), (3, rgba), (4, rgba)):
        got, want = a.low_level(op, img, prefill=0x5A), o.low_level(op, img, prefill=0x5A)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint8), want.view(np.uint8)), "ITMLowLevelEngine helper %d differs" % op
    g = o.low_level(3, rgba, prefill=0x5A)
    assert np.all(g[1:-1, 1:-1, 3] == 255) and np.all(g[-1, :, 0] == 0x5A5A) and np.all(g[0, :, 0] == 0)  # the reference's partial clear
    depth = np.where(rng.random((h, w)) < 0.1, -1.0, rng.uniform(0.5, 3.0, (h, w))).astype(np.float32)
    raw = rng.integers(0, 4000, size=(h, w)).astype(np.int16)
    d_float, d_imu = a.update_view_variants(depth, raw)
    assert np.array_equal(d_float, depth)
    o.update_view(raw)
    assert np.array_equal(d_imu, o.depth)
    a.close(); o.close()
