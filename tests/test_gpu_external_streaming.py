"""-m gpu: the fork's TRACKER_EXTERNAL mode (pose from outside, fusion without ICP), the streaming submit / wait API,
the Kinect disparity conversion and the two integration kernels against each other.

Reference behaviour checked against: ITMExternalTracker::TrackCamera is empty (Engine/ITMExternalTracker.cpp:27-30) and the
pose source writes trackingState->pose_d before the frame (Engine/RosPoseSourceEngine.cpp:112-118), so one frame of that mode is
UpdateView, SetM(pose), AllocateSceneFromDepth, IntegrateIntoScene, CreateExpectedDepths, CreateICPMaps on the reference side.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import parity
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine

pytestmark = pytest.mark.gpu


def _gt_pose(k):
    """ground-truth camera-from-world of frame k, column-major float32[16]"""
    return np.ascontiguousarray(synth.ground_truth_pose(k).astype(np.float32).T).reshape(16)


def _oracle(w, h):
    from oracle import ref
    if ref.available("parity"):
        return ref.RefEngine(w, h)
    from oracle import port
    return port.PortEngine(w, h)


def test_external_pose_mode_matches_the_reference():
    """8 frames fused with externally supplied poses: hash table, free lists, visible list and every voxel equal the
    reference engines driven the same way; the tracker never runs (0 ICP evaluations)"""
    W, H = 320, 240
    o = _oracle(W, H)
    p = parity.cuda_params(o)
    p.tracker_type = capi.TRACKER_EXTERNAL
    eng = ITMMainEngine(p)
    seq = synth.sequence(8, W, H)
    for k in range(8):
        M = _gt_pose(k)
        o.update_view(seq[k])
        o.pose_M = M
        o.allocate()
        o.integrate()
        o.expected_depths()
        o.icp_maps()
        pose = eng.ProcessFrameWithPose(None, seq[k], M)
        assert np.array_equal(pose, np.asarray(o.pose_M, np.float32).reshape(16))
        assert int(eng.icp_stats().sum()) == 0
    parity.assert_scene_equal(o, eng)
    assert np.array_equal(eng.read(capi.BUF_RAYCAST_RESULT).reshape(H, W, 4)[..., 3] > 0, o.raycast_result[..., 3] > 0)
    eng.close()
    o.close()


def test_process_frame_with_pose_on_an_icp_engine_replays_its_own_trajectory():
    """feeding an ICP engine's own poses back through process_frame_with_pose reproduces its scene bit for bit"""
    W, H = 640, 480
    seq = synth.sequence(6, W, H)
    a = ITMMainEngine(width=W, height=H)
    poses = [a.ProcessFrame(None, seq[k]).copy() for k in range(6)]
    b = ITMMainEngine(width=W, height=H)
    for k in range(6):
        b.ProcessFrameWithPose(None, seq[k], poses[k])
    assert a.read(capi.BUF_HASH).tobytes() == b.read(capi.BUF_HASH).tobytes()
    assert a.read(capi.BUF_VOXELS).tobytes() == b.read(capi.BUF_VOXELS).tobytes()
    assert a.read(capi.BUF_POINTS).tobytes() == b.read(capi.BUF_POINTS).tobytes()
    a.close()
    b.close()


def test_streaming_submit_wait_equals_blocking_process_frame():
    """submit_frame / wait_frame with up to MAX_IN_FLIGHT frames queued: per-frame poses and counters, final scene and maps
    identical to the blocking ProcessFrame; tickets count from 1; a stale ticket is refused"""
    import torch
    W, H = 640, 480
    n = 14
    seq = torch.from_numpy(synth.sequence(n, W, H)).pin_memory()
    rgb = torch.full((H, W, 4), 128, dtype=torch.uint8).pin_memory()
    a = ITMMainEngine(width=W, height=H)
    ref_poses, ref_counters = [], []
    for k in range(n):
        ref_poses.append(a.ProcessFrame(rgb, seq[k]).copy())
        ref_counters.append(a.Sync()[1].copy())
    b = ITMMainEngine(width=W, height=H)
    tickets, got = [], {}
    for k in range(n):
        tickets.append(b.SubmitFrame(rgb, seq[k]))
        assert tickets[-1] == k + 1
        if k >= 2:  # keep three frames in flight
            t = tickets[k - 2]
            got[t] = b.WaitFrame(t)
    for t in tickets[-2:]:
        got[t] = b.WaitFrame(t)
    for k in range(n):
        pose, counters = got[k + 1]
        assert np.array_equal(pose, ref_poses[k]), k
        assert np.array_equal(counters, ref_counters[k]), k
    with pytest.raises(capi.ItmError):
        b.WaitFrame(n + 5)
    b.Sync()
    assert a.read(capi.BUF_HASH).tobytes() == b.read(capi.BUF_HASH).tobytes()
    assert a.read(capi.BUF_VOXELS).tobytes() == b.read(capi.BUF_VOXELS).tobytes()
    assert a.read(capi.BUF_NORMALS).tobytes() == b.read(capi.BUF_NORMALS).tobytes()
    assert a.read(capi.BUF_RGB).tobytes() == b.read(capi.BUF_RGB).tobytes()
    a.close()
    b.close()


def test_streaming_without_waiting_applies_back_pressure_and_keeps_results():
    """more frames submitted than MAX_IN_FLIGHT before the first wait: submit collects the oldest results itself"""
    import torch
    W, H = 320, 240
    n = 10
    seq = torch.from_numpy(synth.sequence(n, W, H)).pin_memory()
    a = ITMMainEngine(width=W, height=H)
    ref_poses = [a.ProcessFrame(None, seq[k]).copy() for k in range(n)]
    b = ITMMainEngine(width=W, height=H)
    tickets = [b.SubmitFrame(None, seq[k]) for k in range(n)]
    # the newest MAX_IN_FLIGHT results are in the ring, older ones that were collected early are kept for one ring period
    for k in range(n - capi.MAX_IN_FLIGHT, n):
        pose, _ = b.WaitFrame(tickets[k])
        assert np.array_equal(pose, ref_poses[k])
    a.close()
    b.close()


def test_kinect_disparity_conversion_layer_a_and_layer_b():
    """convertDisparityToDepth (DeviceAgnostic/ITMViewBuilder.h:7-20): bit equal to the formula evaluated in fp32 in the
    reference's operation order, including the zero-denominator and non-positive cases"""
    import torch
    W, H = 64, 48
    rng = np.random.default_rng(7)
    raw = rng.integers(300, 1090, size=(H, W)).astype(np.int16)
    c1, c2, fx = np.float32(1090.0), np.float32(0.075), np.float32(573.71)
    raw[0, 0] = 1090  # disparity_tmp == 0 -> depth 0 -> -1
    raw[0, 1] = 1200  # negative depth -> -1
    tmp = c1 - raw.astype(np.float32)
    with np.errstate(divide="ignore"):
        depth = np.where(tmp == 0, np.float32(0), (np.float32(8.0) * c2 * fx) / tmp).astype(np.float32)
    want = np.where(depth > 0, depth, np.float32(-1.0)).astype(np.float32)
    lib = capi.load()
    p = capi.default_params(W, H)
    ctx = C.c_void_p()
    capi.check(lib.itm_b200_ctx_create(C.byref(p), None, C.byref(ctx)))
    d_in = torch.from_numpy(raw).cuda()
    d_out = torch.zeros((H, W), dtype=torch.float32, device="cuda")
    capi.check(lib.itm_b200_convert_disparity_to_depth(ctx, d_out.data_ptr(), d_in.data_ptr(), W, H, float(c1), float(c2), float(fx)))
    assert np.array_equal(d_out.cpu().numpy(), want)
    lib.itm_b200_ctx_destroy(ctx)
    # Layer B: an engine created with depth_source = KINECT_DISPARITY converts the same way in its view stage
    p2 = capi.default_params(W, H)
    p2.depth_source = capi.DEPTH_KINECT_DISPARITY
    p2.depth_calib_a, p2.depth_calib_b, p2.fx = float(c1), float(c2), float(fx)
    eng = ITMMainEngine(p2)
    eng.UploadDepth(raw)
    eng.RunStage(capi.STAGE_VIEW)
    assert np.array_equal(eng.read_image(capi.BUF_DEPTH), want)
    eng.close()


def test_invalid_params_are_rejected():
    """no_icp_run_till_level / tracking_regime / tiny images used to reach the kernels unchecked"""
    lib = capi.load()
    for mutate in (lambda p: setattr(p, "no_icp_run_till_level", -1), lambda p: setattr(p, "no_icp_run_till_level", 5),
                   lambda p: p.tracking_regime.__setitem__(2, 0), lambda p: setattr(p, "tracker_type", 9),
                   lambda p: setattr(p, "depth_source", 3), lambda p: setattr(p, "device", 99)):
        p = capi.default_params(320, 240)
        mutate(p)
        h = C.c_void_p()
        assert lib.itm_b200_engine_create(C.byref(p), C.byref(h)) != 0
    p = capi.default_params(4, 4)
    h = C.c_void_p()
    assert lib.itm_b200_engine_create(C.byref(p), C.byref(h)) == capi.EINVAL


def test_both_integration_kernels_agree_bitwise():
    """the round-1 row kernel (ITM_B200_INTEGRATE=rows) and the packed column kernel leave identical scenes after 12
    free-running 640x480 frames (run in a child process: the variant is read once per process)"""
    code = (
        "import sys, hashlib; sys.path.insert(0, %r)\n"
        "from infinitam_b200 import synth, capi\n"
        "from infinitam_b200.engines import ITMMainEngine\n"
        "seq = synth.sequence(12, 640, 480)\n"
        "e = ITMMainEngine(width=640, height=480)\n"
        "for k in range(12): e.ProcessFrame(None, seq[k])\n"
        "print(hashlib.sha256(e.read(capi.BUF_VOXELS).tobytes()).hexdigest(), hashlib.sha256(e.read(capi.BUF_HASH).tobytes()).hexdigest())\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = []
    for variant in ("rows", "cols"):
        env = dict(os.environ, ITM_B200_INTEGRATE=variant)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        out.append(r.stdout.strip().splitlines()[-1])
    assert out[0] == out[1]
