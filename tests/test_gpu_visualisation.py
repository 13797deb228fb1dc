"""-m gpu: SURVEY.md 8f rows 1-2 against the reference CPU engines.

* ForwardRender / useApproximateRaycast (ITMVisualisationEngine_CPU.cpp:289-354, ITMTrackingController.cpp:11-46):
  teacher forced per call, and free running over a sequence (the full / approximate decision is taken on the device).
* ITMMainEngine::GetImage (ITMMainEngine.cpp:134-192): FindVisibleBlocks + CreateExpectedDepths + RenderImage from a free
  camera (grey, normal and - for ITMVoxel_s_rgb - colour rendering), DepthToUchar4, raycast image, rgb.

Tolerances: visible list, min/max image, forward projection, missing-point set: bit exact; rendered images within 1 grey
level (observed: equal); raycast points within 1e-4 m."""
import numpy as np
import pytest

import parity
from infinitam_b200 import capi, synth
from oracle import ref

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not ref.available("parity"), reason="oracle/_ref not built (needs /root/reference at build time)")


def _img_diff(a, b):
    return int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max())


def _free_pose(k=0):
    """a camera a little to the side of and above the trajectory, looking slightly down: column-major M (world -> camera)"""
    ang_y, ang_x = np.deg2rad(8.0 + 3.0 * k), np.deg2rad(-5.0)
    Ry = np.array([[np.cos(ang_y), 0, np.sin(ang_y)], [0, 1, 0], [-np.sin(ang_y), 0, np.cos(ang_y)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(ang_x), -np.sin(ang_x)], [0, np.sin(ang_x), np.cos(ang_x)]])
    M = np.eye(4)
    M[:3, :3] = Rx @ Ry
    M[:3, 3] = [0.12, -0.05, 0.08 + 0.02 * k]
    return M.T.astype(np.float32).reshape(16)  # column-major


@needs_ref
def test_forward_render_teacher_forced():
    w, h, n = 320, 240, 8
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h)
    eng = parity.make_cuda_engine(o)
    for k in range(n):
        r = parity.compare_frame(o, eng, seq[k], k, strict=True)
        if k < 2:
            continue
        # the oracle is now at frame k's pose with a full raycast; move the camera by the next frame's motion without a new
        # raycast and forward-project (what Prepare does when !requiresFullRendering)
        pose_next = np.array(o.pose_M, copy=True)
        pose_next[12] += 0.004 * (1 + (k % 3))   # a few millimetres of translation
        pose_next[13] -= 0.003
        saved_pose, saved_age = np.array(o.pose_M, copy=True), o.age
        o.pose_M = pose_next
        eng.set_state(o.pose_M, o.pose_pointcloud_M, [*o.counters, o.age, 0, 0])
        o.expected_depths()
        eng.RunStage(capi.STAGE_EXPECTED_DEPTHS)
        assert np.array_equal(eng.read_image(capi.BUF_MINMAX, 2), o.minmax)
        eng.write(capi.BUF_RAYCAST_RESULT, o.raycast_result)   # teacher forcing (observed: already bit equal)
        o.forward_render()
        eng.RunStage(capi.STAGE_FORWARD_RENDER)
        fp_gpu, fp_ref = eng.read_image(capi.BUF_FORWARD_PROJECTION, 4), o.forward_projection
        _, _, st = eng.get_state()
        miss_ref = o.fwd_missing_points
        assert int(st[5]) == len(miss_ref), "noFwdProjMissingPoints: gpu %d ref %d" % (st[5], len(miss_ref))
        miss_gpu = eng.read(capi.BUF_FWD_MISSING_POINTS)[:int(st[5])]
        assert np.array_equal(np.sort(miss_gpu), miss_ref), "missing-point set differs"   # the reference's list is in raster order
        assert np.array_equal(fp_gpu[..., 3] > 0, fp_ref[..., 3] > 0), "forward projection validity differs"
        d = np.abs(fp_gpu[..., :3] - fp_ref[..., :3]).max() * o.voxel_size
        assert d <= parity.TOL_RAYCAST_M, "forward projection differs by %g m" % d
        assert _img_diff(eng.read_image(capi.BUF_RAYCAST_IMAGE, 4), o.raycast_image) <= 1
        # restore: the next compare_frame continues from the oracle's real state
        o.pose_M = saved_pose
        o.age = saved_age
    eng.close(); o.close()


@needs_ref
def test_approximate_raycast_free_running():
    """useApproximateRaycast on both sides: same full / approximate decisions, poses within tolerance, and the raycast
    image of every frame equal (forward rendered frames included)"""
    w, h, n = 320, 240, 24
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h)
    o.set_use_approximate_raycast(True)
    p = capi.default_params(w, h)
    p.fx, p.fy, p.cx, p.cy = o.intr
    p.use_approximate_raycast = 1
    from infinitam_b200.engines import ITMMainEngine
    eng = ITMMainEngine(p)
    n_fwd = 0
    for k in range(n):
        o.process_frame(seq[k])
        pose = eng.ProcessFrame(None, seq[k])
        rot, trans = parity.pose_diff(pose, o.pose_M)
        assert rot <= 1e-4 and trans <= 1e-4, "frame %d pose differs: %g rad %g m" % (k, rot, trans)
        _, pc, st = eng.get_state()
        assert bool(st[4]) == o.requires_full_rendering, "frame %d: full/approximate decision differs" % k
        assert int(st[3]) == o.age, "frame %d: age_pointCloud gpu %d ref %d" % (k, st[3], o.age)
        n_fwd += 0 if st[4] else 1
        img_g, img_r = eng.read_image(capi.BUF_RAYCAST_IMAGE, 4), o.raycast_image
        # free running: tiny pose differences move a few silhouette pixels; the bulk of the image must agree
        bad = np.count_nonzero(np.abs(img_g.astype(np.int32) - img_r.astype(np.int32)).max(axis=2) > 1)
        assert bad <= 0.002 * w * h, "frame %d: %d raycast-image pixels differ" % (k, bad)
    assert n_fwd >= n // 3, "the sequence should exercise ForwardRender (got %d of %d frames)" % (n_fwd, n)
    eng.close(); o.close()


def _check_get_image(o, eng, w, h, colour):
    K = synth.intrinsics_for(w, h)
    for k, t in enumerate((capi.IMAGE_FREECAMERA_SHADED, capi.IMAGE_FREECAMERA_COLOUR_FROM_NORMAL, capi.IMAGE_FREECAMERA_COLOUR_FROM_VOLUME)):
        M = _free_pose(k)
        img_r = o.get_image(t, M, K, w, h)
        img_g = eng.GetImage(t, M, K, w, h)
        vis_r = o.free_visible_ids
        vis_g = eng.read(capi.BUF_FREEVIEW_VISIBLE_IDS)[:len(vis_r)]
        assert np.array_equal(vis_g, vis_r), "FindVisibleBlocks differs (type %d)" % t   # same order as the serial loop
        assert np.array_equal(eng.read(capi.BUF_FREEVIEW_MINMAX).reshape(h, w, 2), o.free_minmax)
        rc_g, rc_r = eng.read(capi.BUF_FREEVIEW_RAYCAST_RESULT).reshape(h, w, 4), o.free_raycast_result
        assert np.array_equal(rc_g[..., 3] > 0, rc_r[..., 3] > 0)
        assert np.abs(rc_g[..., :3] - rc_r[..., :3]).max() * o.voxel_size <= parity.TOL_RAYCAST_M
        assert img_r[..., :3].any(), "empty reference rendering - the test camera sees nothing"
        assert _img_diff(img_g, img_r) <= 1, "free-view image type %d differs by %d" % (t, _img_diff(img_g, img_r))
        if t == capi.IMAGE_FREECAMERA_COLOUR_FROM_VOLUME and colour:
            assert len(np.unique(img_r[..., :3].reshape(-1, 3), axis=0)) > 4, "colour rendering should not be flat"
    for t in (capi.IMAGE_ORIGINAL_DEPTH, capi.IMAGE_SCENERAYCAST, capi.IMAGE_ORIGINAL_RGB, capi.IMAGE_UNKNOWN):
        assert _img_diff(eng.GetImage(t), o.get_image(t)) <= (1 if t == capi.IMAGE_SCENERAYCAST else 0), "GetImage type %d" % t


@needs_ref
def test_get_image_free_view():
    w, h, n = 320, 240, 5
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h)
    eng = parity.make_cuda_engine(o)
    assert not eng.GetImage(capi.IMAGE_SCENERAYCAST).any()   # no view yet: GetImage returns without touching the image
    for k in range(n):
        parity.compare_frame(o, eng, seq[k], k, strict=True)   # teacher forced: both scenes stay identical
    parity.push_scene(o, eng)
    eng.write(capi.BUF_RGB, np.full((h, w, 4), 128, np.uint8))   # the oracle's default view->rgb
    _check_get_image(o, eng, w, h, colour=False)
    # another image size re-creates renderState_freeview
    M, K = _free_pose(1), synth.intrinsics_for(200, 152)
    assert _img_diff(eng.GetImage(capi.IMAGE_FREECAMERA_SHADED, M, K, 200, 152), o.get_image(capi.IMAGE_FREECAMERA_SHADED, M, K, 200, 152)) <= 1
    eng.close(); o.close()


@pytest.mark.skipif(not ref.available("rgb"), reason="oracle/_ref/libitm_ref_rgb.so not built")
def test_get_image_colour_from_volume():
    w, h, n = 320, 240, 4
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h, flavour="rgb")
    eng = parity.make_cuda_engine(o)
    rgb = synth.checker_rgb(w, h) if hasattr(synth, "checker_rgb") else None
    if rgb is None:
        yy, xx = np.mgrid[0:h, 0:w]
        rgb = np.stack([(xx * 3) % 256, (yy * 5) % 256, ((xx // 16 + yy // 16) % 2) * 200 + 20, np.full_like(xx, 255)], axis=-1).astype(np.uint8)
    o.set_rgb(rgb)
    eng.write(capi.BUF_RGB, rgb)
    for k in range(n):
        parity.compare_frame(o, eng, seq[k], k, strict=True)
    parity.push_scene(o, eng)
    _check_get_image(o, eng, w, h, colour=True)
    eng.close(); o.close()


def _rgb_to_depth_trafo():
    """a small colour-to-depth extrinsic calibration (ITMExtrinsics::calib, column-major): 2 degrees about y, a few cm"""
    a = np.deg2rad(2.0)
    T = np.eye(4, dtype=np.float32)
    T[0, 0], T[0, 2], T[2, 0], T[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    T[:3, 3] = [0.025, -0.01, 0.005]
    return np.ascontiguousarray(T.T).reshape(16)


@pytest.mark.parametrize("flavour", ["parity", "rgb"])
def test_create_point_cloud_equals_the_reference(flavour):
    """IITMVisualisationEngine::CreatePointCloud as ITMTrackingController::Prepare calls it for TRACKER_COLOR
    (ITMTrackingController.cpp:22-28): expected depths at the colour camera's pose, raycast through
    invM_d * trafo_rgb_to_depth, shaded image, points and interpolated voxel colours compacted in raster order - with
    and without skipPoints, for plain and colour voxels, on the reference's own scene (teacher forced)."""
    if not ref.available(flavour):
        pytest.skip("oracle/_ref library for flavour %r not built" % flavour)
    w, h, n = 320, 240, 4
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h, flavour=flavour)
    eng = parity.make_cuda_engine(o)
    if flavour == "rgb":
        yy, xx = np.mgrid[0:h, 0:w]
        rgb = np.stack([(xx * 3) % 256, (yy * 5) % 256, ((xx // 16 + yy // 16) % 2) * 200 + 20, np.full_like(xx, 255)], axis=-1).astype(np.uint8)
        o.set_rgb(rgb)
        eng.write(capi.BUF_RGB, rgb)
    for k in range(n):
        parity.compare_frame(o, eng, seq[k], k, strict=True)
    parity.push_scene(o, eng)
    maps_before = (eng.read(capi.BUF_POINTS).copy(), eng.read(capi.BUF_NORMALS).copy(), eng.read(capi.BUF_RAYCAST_RESULT).copy())
    for T in (None, _rgb_to_depth_trafo()):
        for skip in (False, True):
            loc_a, clr_a, img_a = eng.CreatePointCloud(T, None, skip, with_image=True)
            loc_o, clr_o = o.create_point_cloud(T, skip_points=skip)
            assert len(loc_o) > (2000 if skip else 10000), "the reference found too few points for a meaningful check"
            assert len(loc_a) == len(loc_o), "noTotalPoints: %d vs %d" % (len(loc_a), len(loc_o))
            assert np.array_equal(img_a, o.raycast_image), "shaded raycast differs"
            assert np.array_equal(loc_a, loc_o), "locations differ"
            assert np.array_equal(clr_a, clr_o), "colours differ (max %g)" % np.abs(clr_a - clr_o).max()
            if flavour == "rgb":
                assert clr_o[:, :3].max() > 0.1 and np.all(clr_o[:, 3] == 1.0)
            else:
                assert not clr_o.any()
    # a query: the depth tracker's maps and the live raycast are untouched
    assert np.array_equal(maps_before[0], eng.read(capi.BUF_POINTS)) and np.array_equal(maps_before[1], eng.read(capi.BUF_NORMALS))
    assert np.array_equal(maps_before[2], eng.read(capi.BUF_RAYCAST_RESULT))
    eng.close(); o.close()
