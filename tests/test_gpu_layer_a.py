"""-m gpu: the stage-level C ABI (Layer A) on caller-owned device buffers - exactly what the ITMLib adapter classes
forward to (include/itm_b200_adapter.hpp) - checked against the oracle.  torch only provides the device memory."""
import ctypes as C

import numpy as np
import pytest

import parity
from infinitam_b200 import capi, synth
from oracle import port

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HASH_DT = np.dtype({"names": ["pos", "offset", "ptr"], "formats": [("<i2", 3), "<i4", "<i4"], "offsets": [0, 8, 12], "itemsize": 16})


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dev(arr):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).cuda()


def _host(t, dtype, shape=None):
    a = t.cpu().numpy().view(dtype)
    return a.reshape(shape) if shape is not None else a


def test_stage_functions_on_caller_buffers():
    W, H = 320, 240
    o = port.PortEngine(W, H)
    lib = capi.load()
    p = capi.default_params(W, H)
    p.fx, p.fy, p.cx, p.cy = o.intr
    ctx = C.c_void_p()
    capi.check(lib.itm_b200_ctx_create(C.byref(p), None, C.byref(ctx)))
    P = W * H
    # caller-owned "MemoryBlocks"
    voxels = torch.empty(o.n_local * 512 * 4, dtype=torch.uint8, device="cuda")
    hash_t = torch.empty(o.n_entries * 16, dtype=torch.uint8, device="cuda")
    vba = torch.empty(o.n_local * 4, dtype=torch.uint8, device="cuda")
    exl = torch.empty(o.n_excess * 4, dtype=torch.uint8, device="cuda")
    vis_ids = torch.zeros(o.n_local * 4, dtype=torch.uint8, device="cuda")
    vis_type = torch.zeros(o.n_entries + 1024, dtype=torch.uint8, device="cuda")
    minmax = torch.zeros(P * 8, dtype=torch.uint8, device="cuda")
    ray = torch.zeros(P * 16, dtype=torch.uint8, device="cuda")
    img = torch.zeros(P * 4, dtype=torch.uint8, device="cuda")
    pts = torch.zeros(P * 16, dtype=torch.uint8, device="cuda")
    nrm = torch.zeros(P * 16, dtype=torch.uint8, device="cuda")
    raw = torch.zeros(P * 2, dtype=torch.uint8, device="cuda")
    depth = torch.zeros(P * 4, dtype=torch.uint8, device="cuda")
    lvl1 = torch.zeros((P // 4) * 4, dtype=torch.uint8, device="cuda")

    scene = capi.Scene(voxels.data_ptr(), hash_t.data_ptr(), vba.data_ptr(), exl.data_ptr(), 0, 0)
    rs = capi.RenderState(vis_ids.data_ptr(), vis_type.data_ptr(), 0, minmax.data_ptr(), ray.data_ptr(), img.data_ptr())
    ts = capi.TrackingState(pts.data_ptr(), nrm.data_ptr())
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for i in range(16):
        ts.pose_d[i] = ident[i]
        ts.pose_point_cloud[i] = ident[i]
    ts.age_point_cloud = -1

    torch.cuda.synchronize()
    capi.check(lib.itm_b200_reset_scene(ctx, C.byref(scene)))
    assert (scene.last_free_block_id, scene.last_free_excess_list_id) == (o.n_local - 1, o.n_excess - 1)
    assert np.array_equal(_host(vba, np.int32), o.vba_alloc_list)
    assert np.array_equal(_host(voxels, np.uint32) & 0xFFFFFF, o.voxels & 0xFFFFFF)

    seq = synth.sequence(3, W, H)
    for k in range(3):
        raw.copy_(_dev(seq[k]))
        torch.cuda.synchronize()  # the context has a stream of its own (include/itm_b200.h: itm_b200_ctx_create)
        o.update_view(seq[k])
        capi.check(lib.itm_b200_convert_depth_affine_to_float(ctx, depth.data_ptr(), raw.data_ptr(), W, H, 0.001, 0.0))
        assert np.array_equal(_host(depth, np.float32, (H, W)), o.depth)
        capi.check(lib.itm_b200_filter_subsample_with_holes(ctx, lvl1.data_ptr(), depth.data_ptr(), W, H))
        o.icp_prepare()
        assert np.array_equal(_host(lvl1, np.float32, (H // 2, W // 2)), o.pyramid_level(1)[0])

        if o.age != -1:
            # one ComputeGandH per level at the current pose, then the whole TrackCamera
            inv = o.mat_inv(o.pose_M)
            n_ref, out_ref = o.icp_gandh(0, inv)
            f = C.c_float(); nv = C.c_int()
            nabla, hess = np.zeros(6, np.float32), np.zeros(36, np.float32)
            intr = np.array(o.intr, np.float32)
            _, _, thr, typ = o.icp_config()
            pc = o.pose_pointcloud_M
            capi.check(lib.itm_b200_compute_g_and_h(ctx, depth.data_ptr(), W, H, _f(intr), pts.data_ptr(), nrm.data_ptr(), W, H, _f(intr),
                                                    _f(inv), _f(pc), float(thr[0]), int(typ[0]), C.byref(f), _f(nabla), _f(hess), C.byref(nv)))
            assert nv.value == n_ref
            assert abs(f.value - out_ref[1]) <= 5e-4 * abs(out_ref[1])  # the reference sums ~77k terms serially in fp32
            assert np.allclose(nabla, out_ref[2:8], rtol=2e-4, atol=1e-3 * np.abs(out_ref[2:8]).max())
            assert np.allclose(hess, out_ref[8:44], rtol=2e-4, atol=1e-4 * np.abs(out_ref[8:44]).max())
            o.track()
            capi.check(lib.itm_b200_track_camera(ctx, depth.data_ptr(), C.byref(ts)))
            rot, trans = parity.pose_diff(np.array(ts.pose_d[:], np.float32), o.pose_M)
            assert rot <= 1e-4 and trans <= 1e-4
        else:
            o.track()
        pose = o.pose_M  # teacher forcing
        for i in range(16):
            ts.pose_d[i] = float(pose[i])

        o.allocate()
        capi.check(lib.itm_b200_allocate_scene_from_depth(ctx, C.byref(scene), C.byref(rs), depth.data_ptr(), _f(pose), 0))
        c = o.counters
        assert (rs.no_visible_entries, scene.last_free_block_id, scene.last_free_excess_list_id) == tuple(int(x) for x in c)
        assert parity.hash_equal(_host(hash_t, HASH_DT), o.hash_entries)
        assert np.array_equal(_host(vis_ids, np.int32)[: c[0]], o.visible_ids[: c[0]])

        o.integrate()
        capi.check(lib.itm_b200_integrate_into_scene(ctx, C.byref(scene), C.byref(rs), depth.data_ptr(), _f(pose)))
        assert np.array_equal(_host(voxels, np.uint32) & 0xFFFFFF, o.voxels & 0xFFFFFF)

        o.expected_depths()
        intr = np.array(o.intr, np.float32)
        capi.check(lib.itm_b200_create_expected_depths(ctx, C.byref(scene), C.byref(rs), _f(pose), _f(intr)))
        assert np.array_equal(_host(minmax, np.float32, (H, W, 2)), o.minmax)

        o.icp_maps()
        capi.check(lib.itm_b200_create_icp_maps(ctx, C.byref(scene), C.byref(rs), C.byref(ts)))
        assert np.array_equal(_host(ray, np.float32, (H, W, 4)), o.raycast_result)
        assert np.array_equal(_host(pts, np.float32, (H, W, 4)), o.points)
        assert np.array_equal(_host(nrm, np.float32, (H, W, 4)), o.normals)
        assert np.array_equal(_host(img, np.uint8, (H, W, 4)), o.raycast_image)
        assert np.array_equal(np.array(ts.pose_point_cloud[:], np.float32), o.pose_pointcloud_M)
    lib.itm_b200_ctx_destroy(ctx)
    o.close()
