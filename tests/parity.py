"""Stage-by-stage ("teacher forced") comparison of the CUDA engines with an oracle engine.

The oracle object must look like oracle.ref.RefEngine (the real reference CPU engines) or
oracle.port.PortEngine (the C restatement) - both expose the same attributes.  Before each
candidate stage the CUDA engine is given the oracle's pre-stage state, so a difference in one
stage cannot leak into the next one.

Tolerances are the ones BASELINE.json states: hash table / visible list bit exact, voxel sdf
and weight within 1 LSB, raycast points within 1e-4 m, pose within 1e-4.
"""
from __future__ import annotations

import numpy as np

from infinitam_b200 import capi
from infinitam_b200.engines import ITMMainEngine

TOL_RAYCAST_M = 1e-4
TOL_POSE = 1e-4


def make_cuda_engine(oracle) -> ITMMainEngine:
    return ITMMainEngine(cuda_params(oracle))


def cuda_params(oracle):
    """itm_b200_params matching the oracle engine's configuration"""
    p = capi.default_params(oracle.W, oracle.H)
    p.fx, p.fy, p.cx, p.cy = oracle.intr
    p.voxel_size, p.mu, p.max_w = oracle.voxel_size, oracle.mu, oracle.max_w
    p.view_frustum_min, p.view_frustum_max = oracle.vf_min, oracle.vf_max
    p.sdf_local_block_num, p.sdf_bucket_num, p.sdf_excess_list_size = oracle.n_local, oracle.n_bucket, oracle.n_excess
    p.rgb_fx, p.rgb_fy, p.rgb_cx, p.rgb_cy = oracle.intr
    if oracle.const("sizeof_voxel") == 8:
        p.voxel_type = capi.VOXEL_S_RGB
    return p


def assert_scene_equal(oracle, eng: ITMMainEngine):
    """free-running comparison of the persistent state: counters, hash table, free lists, visible list bit exact, voxels
    within 1 LSB (returns the number of voxels whose sdf / weight differ at all)"""
    _, _, st = eng.get_state()
    c = oracle.counters
    assert list(st[:3]) == [int(x) for x in c], "counters differ: gpu %s ref %s" % (st[:3], c)
    assert hash_equal(eng.read(capi.BUF_HASH), oracle.hash_entries), "hash table differs"
    n_vis = int(c[0])
    assert np.array_equal(eng.read(capi.BUF_VISIBLE_IDS)[:n_vis], oracle.visible_ids[:n_vis]), "visible list differs"
    assert np.array_equal(eng.read(capi.BUF_VISIBLE_TYPES), oracle.visible_types), "entriesVisibleType differs"
    assert np.array_equal(eng.read(capi.BUF_VBA_ALLOC_LIST), oracle.vba_alloc_list)
    assert np.array_equal(eng.read(capi.BUF_EXCESS_ALLOC_LIST), oracle.excess_alloc_list)
    ds, dw, ns, nw = voxel_diff(eng.read(capi.BUF_VOXELS), oracle.voxels)
    assert ds <= 1 and dw <= 1, "voxels differ by more than 1 LSB: sdf %d w %d" % (ds, dw)
    return ns, nw


def push_scene(oracle, eng: ITMMainEngine):
    """oracle scene + render state + tracking state -> CUDA engine"""
    eng.write(capi.BUF_HASH, oracle.hash_entries)
    eng.write(capi.BUF_VOXELS, oracle.voxels)
    eng.write(capi.BUF_VBA_ALLOC_LIST, oracle.vba_alloc_list)
    eng.write(capi.BUF_EXCESS_ALLOC_LIST, oracle.excess_alloc_list)
    eng.write(capi.BUF_VISIBLE_IDS, oracle.visible_ids)
    eng.write(capi.BUF_VISIBLE_TYPES, oracle.visible_types)
    push_counters_pose(oracle, eng)


def push_counters_pose(oracle, eng):
    c = oracle.counters
    eng.set_state(oracle.pose_M, oracle.pose_pointcloud_M, [c[0], c[1], c[2], oracle.age, 0, 0])


def push_maps(oracle, eng):
    eng.write(capi.BUF_POINTS, oracle.points)
    eng.write(capi.BUF_NORMALS, oracle.normals)


def hash_equal(a, b):
    """compare ITMHashEntry arrays ignoring the 2 padding bytes"""
    return (np.array_equal(a["pos"], b["pos"]) and np.array_equal(a["offset"], b["offset"]) and np.array_equal(a["ptr"], b["ptr"]))


def voxel_diff(a_u32, b_u32):
    """max |sdf| and |w| difference between two ITMVoxel_s arrays given as uint32 words"""
    a_u32, b_u32 = a_u32.astype(np.uint32), b_u32.astype(np.uint32)  # uint64 words of ITMVoxel_s_rgb: the low half
    sa = (a_u32 & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int32)
    sb = (b_u32 & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int32)
    wa = ((a_u32 >> 16) & 0xFF).astype(np.int32)
    wb = ((b_u32 >> 16) & 0xFF).astype(np.int32)
    ds, dw = np.abs(sa - sb), np.abs(wa - wb)
    return int(ds.max()), int(dw.max()), int(np.count_nonzero(ds)), int(np.count_nonzero(dw))


def voxel_colour_diff(a_u64, b_u64):
    """max difference of clr.r/g/b and of w_color between two ITMVoxel_s_rgb arrays given as uint64 words"""
    d = 0
    for shift in (24, 32, 40):
        ca = ((a_u64 >> np.uint64(shift)) & np.uint64(0xFF)).astype(np.int32)
        cb = ((b_u64 >> np.uint64(shift)) & np.uint64(0xFF)).astype(np.int32)
        d = max(d, int(np.abs(ca - cb).max()))
    wa = ((a_u64 >> np.uint64(48)) & np.uint64(0xFF)).astype(np.int32)
    wb = ((b_u64 >> np.uint64(48)) & np.uint64(0xFF)).astype(np.int32)
    return d, int(np.abs(wa - wb).max())


def pose_diff(Ma, Mb):
    """(rotation difference in rad, translation difference in m) between two column-major 4x4 poses"""
    A = np.asarray(Ma, np.float64).reshape(4, 4).T
    B = np.asarray(Mb, np.float64).reshape(4, 4).T
    R = A[:3, :3] @ B[:3, :3].T
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    s = np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    ang = float(np.arctan2(s, c))
    return ang, float(np.abs(A[:3, 3] - B[:3, 3]).max())


def compare_frame(oracle, eng: ITMMainEngine, depth_i16, frame_no, report=None, strict=True):
    """Runs one ProcessFrame on the oracle stage by stage and checks every CUDA stage against it.
    Returns a dict of measured differences; raises AssertionError on a violated tolerance."""
    r = {"frame": frame_no, "failures": []}

    def check(cond, msg="check failed"):
        if not cond:
            if strict:
                raise AssertionError("frame %d: %s" % (frame_no, msg))
            r["failures"].append(msg)

    depth_i16 = np.ascontiguousarray(depth_i16, dtype=np.int16)

    # ---- view -----------------------------------------------------------------------------
    oracle.update_view(depth_i16)
    eng.UploadDepth(depth_i16)
    eng.RunStage(capi.STAGE_VIEW)
    d_ref = oracle.depth
    d_gpu = eng.read_image(capi.BUF_DEPTH)
    r["view_equal"] = bool(np.array_equal(d_ref, d_gpu))
    check(r["view_equal"], "depth conversion differs")
    oracle.icp_prepare()
    for lvl, buf in ((1, capi.BUF_PYRAMID_1), (2, capi.BUF_PYRAMID_2), (3, capi.BUF_PYRAMID_3), (4, capi.BUF_PYRAMID_4)):
        ref_l, _ = oracle.pyramid_level(lvl)
        gpu_l = eng.read(buf).reshape(ref_l.shape)
        check(np.array_equal(ref_l, gpu_l), "pyramid level %d differs" % lvl)
    r["pyramid_equal"] = True

    # ---- track ----------------------------------------------------------------------------
    push_counters_pose(oracle, eng)
    if oracle.age != -1:
        push_maps(oracle, eng)
    oracle.track()
    eng.RunStage(capi.STAGE_TRACK)
    pose_gpu, _, _ = eng.get_state()
    rot, trans = pose_diff(pose_gpu, oracle.pose_M)
    r["pose_rot_rad"], r["pose_trans_m"] = rot, trans
    check(rot <= TOL_POSE and trans <= TOL_POSE, "pose differs: %g rad, %g m" % (rot, trans))

    # ---- allocate ---------------------------------------------------------------------------
    push_counters_pose(oracle, eng)  # teacher forcing: continue from the oracle's pose
    oracle.allocate()
    eng.RunStage(capi.STAGE_ALLOCATE)
    _, _, st = eng.get_state()
    c = oracle.counters
    r["counters_ref"], r["counters_gpu"] = [int(x) for x in c], [int(x) for x in st[:3]]
    check(list(st[:3]) == list(c), "allocation counters differ: gpu %s ref %s" % (st[:3], c))
    h_gpu = eng.read(capi.BUF_HASH)
    h_ref = oracle.hash_entries
    r["hash_equal"] = bool(hash_equal(h_gpu, h_ref))
    if not r["hash_equal"]:
        bad = np.nonzero((h_gpu["ptr"] != h_ref["ptr"]) | (h_gpu["offset"] != h_ref["offset"]) | np.any(h_gpu["pos"] != h_ref["pos"], axis=1))[0]
        check(False, "hash table differs in %d slots, first %s: gpu %s ref %s" % (len(bad), bad[:5], h_gpu[bad[:5]], h_ref[bad[:5]]))
    n_vis = int(c[0])
    vis_gpu = eng.read(capi.BUF_VISIBLE_IDS)[:n_vis]
    vis_ref = oracle.visible_ids[:n_vis]
    r["visible_equal"] = bool(np.array_equal(np.sort(vis_gpu), np.sort(vis_ref)))
    r["visible_same_order"] = bool(np.array_equal(vis_gpu, vis_ref))
    check(r["visible_equal"], "visible list differs")
    check(np.array_equal(eng.read(capi.BUF_VISIBLE_TYPES), oracle.visible_types), "entriesVisibleType differs")
    check(np.array_equal(eng.read(capi.BUF_VBA_ALLOC_LIST), oracle.vba_alloc_list))
    if r["failures"]:
        push_scene(oracle, eng)  # non-strict mode: keep the following stages teacher forced

    # ---- integrate --------------------------------------------------------------------------
    oracle.integrate()
    eng.RunStage(capi.STAGE_INTEGRATE)
    v_gpu = eng.read(capi.BUF_VOXELS)
    v_ref = oracle.voxels
    if v_gpu.dtype == np.uint32 and np.array_equal(v_gpu & np.uint32(0x00FFFFFF), v_ref & np.uint32(0x00FFFFFF)):
        ds, dw, ns, nw = 0, 0, 0, 0  # identical up to the padding byte: skip the element-wise differences (33 M voxels)
    else:
        ds, dw, ns, nw = voxel_diff(v_gpu, v_ref)
    r["voxel_max_dsdf"], r["voxel_max_dw"], r["voxel_n_dsdf"], r["voxel_n_dw"] = ds, dw, ns, nw
    check(ds <= 1 and dw <= 1, "voxels differ by more than 1 LSB: sdf %d w %d" % (ds, dw))
    if v_gpu.dtype == np.uint64:
        dc, dwc = voxel_colour_diff(v_gpu, oracle.voxels)
        r["voxel_max_dclr"], r["voxel_max_dwcolor"] = dc, dwc
        ns += int(np.count_nonzero((v_gpu ^ oracle.voxels) & np.uint64(0x00FFFFFFFFFFFFFF)))
        check(dc <= 1 and dwc <= 1, "voxel colours differ by more than 1 LSB: clr %d w_color %d" % (dc, dwc))
    if ns or nw:
        eng.write(capi.BUF_VOXELS, oracle.voxels)  # keep teacher forcing exact

    # ---- expected depths --------------------------------------------------------------------
    oracle.expected_depths()
    eng.RunStage(capi.STAGE_EXPECTED_DEPTHS)
    mm_gpu = eng.read_image(capi.BUF_MINMAX, 2)
    mm_ref = oracle.minmax
    r["minmax_equal"] = bool(np.array_equal(mm_gpu, mm_ref))
    check(r["minmax_equal"], "expected-depth image differs in %d px" % int(np.count_nonzero(np.any(mm_gpu != mm_ref, axis=2))))

    if not r["minmax_equal"]:
        eng.write(capi.BUF_MINMAX, mm_ref)

    # ---- raycast + ICP maps -------------------------------------------------------------------
    oracle.icp_maps()
    eng.RunStage(capi.STAGE_ICP_MAPS)
    rc_gpu = eng.read_image(capi.BUF_RAYCAST_RESULT, 4)
    rc_ref = oracle.raycast_result
    hit_gpu, hit_ref = rc_gpu[..., 3] > 0, rc_ref[..., 3] > 0
    r["raycast_hit_mismatch"] = int(np.count_nonzero(hit_gpu != hit_ref))
    both = hit_gpu & hit_ref
    dpt = np.abs(rc_gpu[..., :3] - rc_ref[..., :3])[both] * oracle.voxel_size if both.any() else np.zeros(1)
    r["raycast_max_diff_m"] = float(dpt.max())
    r["raycast_bit_equal"] = bool(np.array_equal(rc_gpu, rc_ref))
    check(r["raycast_hit_mismatch"] == 0, "raycast hit mask differs in %d px" % r["raycast_hit_mismatch"])
    check(r["raycast_max_diff_m"] <= TOL_RAYCAST_M)
    p_gpu, p_ref = eng.read_image(capi.BUF_POINTS, 4), oracle.points
    n_gpu, n_ref = eng.read_image(capi.BUF_NORMALS, 4), oracle.normals
    check(np.array_equal(p_gpu[..., 3], p_ref[..., 3]), "ICP point validity differs")
    r["points_max_diff_m"] = float(np.abs(p_gpu - p_ref).max())
    r["normals_max_diff"] = float(np.abs(n_gpu - n_ref).max())
    check(r["points_max_diff_m"] <= TOL_RAYCAST_M)
    check(r["normals_max_diff"] <= 1e-3)
    img_gpu, img_ref = eng.read_image(capi.BUF_RAYCAST_IMAGE, 4), oracle.raycast_image
    r["image_max_diff"] = int(np.abs(img_gpu.astype(np.int32) - img_ref.astype(np.int32)).max())
    check(r["image_max_diff"] <= 1)
    _, pc_gpu, st = eng.get_state()
    check(np.array_equal(pc_gpu, oracle.pose_pointcloud_M), "pose_pointCloud differs")
    check(int(st[3]) == oracle.age)
    if report is not None:
        report.append(r)
    return r


def compare_free_running(oracle, eng: ITMMainEngine, frames):
    """Both engines process the sequence on their own (no teacher forcing); returns per-frame pose differences."""
    out = []
    for k, d in enumerate(frames):
        d = np.ascontiguousarray(d, dtype=np.int16)
        oracle.process_frame(d)
        pose = eng.ProcessFrame(None, d)
        rot, trans = pose_diff(pose, oracle.pose_M)
        _, _, st = eng.get_state()
        out.append({"frame": k, "rot": rot, "trans": trans, "counters_gpu": [int(x) for x in st[:3]],
                    "counters_ref": [int(x) for x in oracle.counters]})
    return out
