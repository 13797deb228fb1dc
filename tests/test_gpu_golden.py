"""-m gpu: the CUDA path (through the C ABI) against the committed golden vectors from the real reference, and against
the C restatement on configurations the reference cannot build (run-time pool sizes, 2 mm voxels, 1280x720)."""
import numpy as np
import pytest

import golden_check
import parity
from infinitam_b200 import capi, synth
from infinitam_b200.engines import ITMMainEngine
from oracle import port

pytestmark = pytest.mark.gpu


def _cuda_state(eng):
    pose, _, st = eng.get_state()
    return dict(pose=pose, counters=st[:3], hash_entries=eng.read(capi.BUF_HASH), visible_ids=eng.read(capi.BUF_VISIBLE_IDS),
                voxels_u32=eng.read(capi.BUF_VOXELS), minmax=eng.read_image(capi.BUF_MINMAX, 2),
                raycast=eng.read_image(capi.BUF_RAYCAST_RESULT, 4), points=eng.read_image(capi.BUF_POINTS, 4),
                normals=eng.read_image(capi.BUF_NORMALS, 4), image=eng.read_image(capi.BUF_RAYCAST_IMAGE, 4),
                visible_types=eng.read(capi.BUF_VISIBLE_TYPES), depth=eng.read_image(capi.BUF_DEPTH))


def test_cuda_free_running_reproduces_golden_vectors():
    """ProcessFrame x4 on the golden input.  Everything that is integer / byte / index work must be bit exact; the pose
    (device libm + tree-ordered sums in the ICP reduction) must stay within 1e-4, and it does so closely enough that the
    scene stays identical on this sequence."""
    g = golden_check.load()
    seq = golden_check.golden_sequence(g)
    eng = ITMMainEngine(width=int(g["W"]), height=int(g["H"]))
    for k in range(int(g["N"])):
        pose = eng.ProcessFrame(None, seq[k])
        rot, trans = parity.pose_diff(pose, g["f%d_pose" % k])
        assert rot <= 1e-4 and trans <= 1e-4
        if k == 0:  # no tracking yet: the whole frame is bit exact
            golden_check.check_frame(g, k, exact_pose=True, exact_maps=True, **_cuda_state(eng))
    eng.close()


@pytest.fixture(params=[1, 2], ids=["ordered-scans", "compact-lists"])
def alloc_mode(request):
    """both implementations of AllocateSceneFromDepth (itm_b200_set_alloc_mode) must give the reference's table bit for bit"""
    prev = capi.set_alloc_mode(request.param)
    yield request.param
    capi.set_alloc_mode(prev)


def test_cuda_teacher_forced_against_port_with_runtime_pools(alloc_mode):
    """small pools (4096 blocks, 2^15 buckets): many bucket collisions -> excess-list allocation is exercised heavily"""
    o = port.PortEngine(320, 240, n_local=0x2000, n_bucket=0x4000, n_excess=0x2000)
    eng = parity.make_cuda_engine(o)
    seq = synth.sequence(4, 320, 240)
    rows = [parity.compare_frame(o, eng, seq[k], k, strict=True) for k in range(4)]
    n_excess_used = 0x2000 - 1 - rows[-1]["counters_ref"][2]
    assert n_excess_used > 300, "test is meant to exercise the excess list (used %d)" % n_excess_used
    eng.close(); o.close()


def test_cuda_pool_exhaustion_matches_oracle(alloc_mode):
    """the voxel-block pool and the excess list both run dry while the camera moves; the reference keeps decrementing its
    counters and silently skips the blocks (ITMSceneReconstructionEngine_CPU.cpp:187-189, :206) - so must we"""
    o = port.PortEngine(320, 240, n_local=7168, n_bucket=0x4000, n_excess=1600)
    eng = parity.make_cuda_engine(o)
    seq = synth.sequence(8, 320, 240)
    rows = [parity.compare_frame(o, eng, seq[k], k, strict=True) for k in range(8)]
    assert rows[-1]["counters_ref"][1] < 0 and rows[-1]["counters_ref"][2] < 0, "both pools were supposed to be exhausted"
    assert max(r["counters_ref"][0] for r in rows) <= 7168
    eng.close(); o.close()


def test_cuda_allocation_lists_equal_scans_free_running():
    """40 free-running 640x480 frames, once with each allocation implementation: identical hash table, free-list heads,
    entriesVisibleType, visible list and pose after every tenth frame (the camera sweeps, so entries leave and re-enter
    the list and the excess list grows)"""
    seq = synth.sequence(40, 640, 480)
    snaps = []
    for mode in (1, 2):
        prev = capi.set_alloc_mode(mode)
        try:
            eng = ITMMainEngine(width=640, height=480)
            got = []
            for k in range(40):
                eng.ProcessFrame(None, seq[k])
                if k % 10 == 9:
                    pose, _, st = eng.get_state()
                    n = int(st[0])
                    got.append((pose.copy(), st[:3].copy(), eng.read(capi.BUF_HASH).copy(), eng.read(capi.BUF_VISIBLE_IDS)[:n].copy(),
                                eng.read(capi.BUF_VISIBLE_TYPES).copy()))
            eng.close()
        finally:
            capi.set_alloc_mode(prev)
        snaps.append(got)
    for a, b in zip(*snaps):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_cuda_hd_2mm_teacher_forced_against_port():
    """BASELINE configs[2] shape: 1280x720, 2 mm voxels, mu = 0.02 (band of 2.5 blocks), enlarged pools"""
    o = port.PortEngine(1280, 720, voxel_size=0.002, n_local=0x40000, n_bucket=0x200000, n_excess=0x40000, fast=False)
    eng = parity.make_cuda_engine(o)
    seq = synth.sequence(2, 1280, 720)
    rows = [parity.compare_frame(o, eng, seq[k], k, strict=True) for k in range(2)]
    assert rows[0]["counters_ref"][0] > 30000
    eng.close(); o.close()


def test_cuda_rows_8f_reproduce_golden_vectors():
    """MeshScene, the free-view GetImage chain and ForwardRender against tests/golden/ref_rows8f_qqvga.npz (made from the real
    reference by tests/golden/make_golden_8f.py).  Frame 0 is fused at the identity pose, so the scene is bit-identical to
    the reference's and every vector is compared exactly."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_rows8f_qqvga.npz"))
    W, H = int(g["W"]), int(g["H"])
    seq = synth.sequence(1, W, H, noise=True)
    assert golden_check.crc(seq[0]) == int(g["depth_crc"]), "synthetic generator no longer reproduces the golden input"
    eng = ITMMainEngine(width=W, height=H)
    eng.ProcessFrame(None, seq[0])
    # ---- meshing
    tri = eng.UpdateMesh()
    assert len(tri) == int(g["mesh_n"])
    assert np.array_equal(tri[:32], g["mesh_head"]) and np.array_equal(tri[-32:], g["mesh_tail"])
    assert golden_check.crc(tri) == int(g["mesh_crc"]), "triangle array differs"
    # ---- free-view rendering
    for k in range(2):
        img = eng.GetImage(int(g["free%d_type" % k]), g["free%d_pose" % k], g["free%d_intr" % k], W, H)
        n = int(g["free%d_nvis" % k])
        c = [int(x) for x in g["free%d_crc" % k]]
        assert golden_check.crc(eng.read(capi.BUF_FREEVIEW_VISIBLE_IDS)[:n]) == c[0], "FindVisibleBlocks differs"
        assert golden_check.crc(eng.read(capi.BUF_FREEVIEW_MINMAX)) == c[1], "free-view expected depths differ"
        assert golden_check.crc(eng.read(capi.BUF_FREEVIEW_RAYCAST_RESULT)) == c[2], "free-view raycast differs"
        d = np.abs(img[::5, ::7].astype(np.int32) - g["free%d_sample" % k].astype(np.int32)).max()
        assert d <= 1, "free-view image differs by %d levels" % d
        assert golden_check.crc(img) == c[3], "free-view image differs"
    # ---- forward rendering at a pose a few millimetres away, without a new raycast
    pose, pc, st = eng.get_state()
    eng.set_state(g["fwd_pose"], pc, st)
    eng.RunStage(capi.STAGE_EXPECTED_DEPTHS)
    eng.RunStage(capi.STAGE_FORWARD_RENDER)
    c = [int(x) for x in g["fwd_crc"]]
    _, _, st = eng.get_state()
    assert golden_check.crc(eng.read_image(capi.BUF_MINMAX, 2)) == c[3]
    assert int(st[5]) == int(g["fwd_nmissing"])
    assert golden_check.crc(np.sort(eng.read(capi.BUF_FWD_MISSING_POINTS)[:int(st[5])])) == c[1], "missing-point set differs"
    fp = eng.read_image(capi.BUF_FORWARD_PROJECTION, 4)
    assert int(np.count_nonzero(fp[..., 3] > 0)) == int(g["fwd_nvalid"])
    assert golden_check.crc(fp) == c[0], "forward projection differs"
    assert golden_check.crc(eng.read_image(capi.BUF_RAYCAST_IMAGE, 4)) == c[2], "forward-rendered image differs"
    eng.close()


def test_cuda_point_cloud_and_low_level_helpers_reproduce_golden_vectors():
    """CreatePointCloud (the TRACKER_COLOR branch of ITMTrackingController::Prepare) and the colour-tracker helpers of
    ITMLowLevelEngine against tests/golden/ref_cloud_lowlevel_qqvga.npz (made from the real reference by
    tests/golden/make_golden_cloud.py) - exact, and independent of oracle/_ref being present."""
    import ctypes as C
    import os
    import sys
    torch = pytest.importorskip("torch")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_cloud as mk
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cloud_lowlevel_qqvga.npz"))
    W, H = int(g["W"]), int(g["H"])
    seq = synth.sequence(1, W, H, noise=True)
    assert golden_check.crc(seq[0]) == int(g["depth_crc"]), "synthetic generator no longer reproduces the golden input"
    eng = ITMMainEngine(width=W, height=H)
    eng.ProcessFrame(None, seq[0])
    for k in range(4):
        T = g["trafo"] if int(g["cloud%d_use_trafo" % k]) else None
        loc, clr, img = eng.CreatePointCloud(T, None, bool(int(g["cloud%d_skip" % k])), with_image=True)
        assert len(loc) == int(g["cloud%d_n" % k]), "noTotalPoints %d, golden %d" % (len(loc), int(g["cloud%d_n" % k]))
        assert np.array_equal(loc[:16], g["cloud%d_head" % k])
        assert [golden_check.crc(loc), golden_check.crc(clr), golden_check.crc(img)] == [int(x) for x in g["cloud%d_crc" % k]]
    eng.close()
    # ---- ITMLowLevelEngine helpers through Layer A on caller-owned buffers
    LW, LH, prefill = int(g["LW"]), int(g["LH"]), int(g["prefill"])
    rgba, f4 = mk.lowlevel_inputs()
    assert [golden_check.crc(rgba), golden_check.crc(f4)] == [int(x) for x in g["lowlevel_in_crc"]], "numpy no longer reproduces the golden input"
    lib = capi.load()
    p = capi.default_params(W, H)
    ctx = C.c_void_p()
    capi.check(lib.itm_b200_ctx_create(C.byref(p), None, C.byref(ctx)))
    d_rgba = torch.from_numpy(rgba.reshape(-1)).cuda()
    d_f4 = torch.from_numpy(f4.reshape(-1)).cuda()
    half = (LH // 2) * (LW // 2)
    outs = [torch.full((LH * LW * 4,), prefill, dtype=torch.uint8, device="cuda"),   # CopyImage
            torch.full((half * 4,), prefill, dtype=torch.uint8, device="cuda"),      # FilterSubsample
            torch.full((half * 16,), prefill, dtype=torch.uint8, device="cuda"),     # FilterSubsampleWithHoles(Vector4f)
            torch.full((LH * LW * 8,), prefill, dtype=torch.uint8, device="cuda"),   # GradientX (Vector4s)
            torch.full((LH * LW * 8,), prefill, dtype=torch.uint8, device="cuda")]   # GradientY
    torch.cuda.synchronize()
    capi.check(lib.itm_b200_copy_image(ctx, outs[0].data_ptr(), d_rgba.data_ptr(), LH * LW * 4))
    capi.check(lib.itm_b200_filter_subsample_rgba(ctx, outs[1].data_ptr(), d_rgba.data_ptr(), LW, LH))
    capi.check(lib.itm_b200_filter_subsample_with_holes_float4(ctx, outs[2].data_ptr(), d_f4.data_ptr(), LW, LH))
    capi.check(lib.itm_b200_gradient_x(ctx, outs[3].data_ptr(), d_rgba.data_ptr(), LW, LH))
    capi.check(lib.itm_b200_gradient_y(ctx, outs[4].data_ptr(), d_rgba.data_ptr(), LW, LH))
    got = [golden_check.crc(o.cpu().numpy()) for o in outs]
    assert got == [int(x) for x in g["lowlevel_crc"]], "low-level helper outputs differ from the golden vectors: %s" % (got,)
    lib.itm_b200_ctx_destroy(ctx)
