"""-m gpu: host swapping (settings.useSwapping, BASELINE configs[4]) teacher-forced against the reference CPU engines:
allocation with the enlarged frustum, swap-state marking and re-allocation of swapped-out entries; IntegrateGlobalIntoLocal
(host copy merged into the fresh block) and SaveToGlobalMemory (invisible blocks leave active memory, their slots go back
to the free list in slot order).  The camera sweeps away from the first view and back, so blocks travel both ways."""
import numpy as np
import pytest

import parity
from infinitam_b200 import capi, synth
from oracle import ref

pytestmark = pytest.mark.gpu


def _frames():
    return list(range(0, 100, 4)) + list(range(96, -1, -4))


def _run(flavour):
    w, h = 320, 240
    o = ref.RefEngine(w, h, flavour=flavour, use_swapping=True)
    p = capi.default_params(w, h)
    p.fx, p.fy, p.cx, p.cy = o.intr
    p.rgb_fx, p.rgb_fy, p.rgb_cx, p.rgb_cy = o.intr
    p.use_swapping = 1
    if flavour == "rgb":
        p.voxel_type = capi.VOXEL_S_RGB
    from infinitam_b200.engines import ITMMainEngine
    eng = ITMMainEngine(p)
    vmask = np.uint64(0x00FFFFFFFFFFFFFF) if flavour == "rgb" else np.uint32(0x00FFFFFF)  # the padding byte is not compared
    total_in = total_out = 0
    yy, xx = np.mgrid[0:h, 0:w]
    for i, k in enumerate(_frames()):
        depth = synth.render_depth(k, w, h)
        if flavour == "rgb":
            rgb = np.stack([(xx * 7 + 13 * i) & 255, (yy * 5 + 3 * i) & 255, ((xx ^ yy) * 3) & 255, np.full_like(xx, 255)], -1).astype(np.uint8)
            o.set_rgb(rgb)
            eng.write(capi.BUF_RGB, rgb)
        # ---- view + track on the oracle, pose handed over (tracking parity is covered elsewhere)
        o.update_view(depth)
        eng.UploadDepth(depth)
        eng.RunStage(capi.STAGE_VIEW)
        o.track()
        parity.push_counters_pose(o, eng)
        # ---- allocate
        o.allocate()
        eng.RunStage(capi.STAGE_ALLOCATE)
        _, _, st = eng.get_state()
        assert list(st[:3]) == [int(x) for x in o.counters], "frame %d: counters after allocate" % i
        assert parity.hash_equal(eng.read(capi.BUF_HASH), o.hash_entries), "frame %d: hash after allocate" % i
        n_vis = int(o.counters[0])
        assert np.array_equal(eng.read(capi.BUF_VISIBLE_IDS)[:n_vis], o.visible_ids[:n_vis])
        assert np.array_equal(eng.read(capi.BUF_VISIBLE_TYPES), o.visible_types)
        assert np.array_equal(eng.read(capi.BUF_SWAP_STATES), o.swap_states), "frame %d: swap states after allocate" % i
        # ---- integrate
        o.integrate()
        eng.RunStage(capi.STAGE_INTEGRATE)
        assert np.array_equal(eng.read(capi.BUF_VOXELS) & vmask, o.voxels & vmask), "frame %d: voxels after integrate" % i
        # ---- swap in / out
        stored_before = o.has_stored_data.copy()
        o.swap()
        eng.RunStage(capi.STAGE_SWAP)
        has, blocks, n_in, n_out = eng.global_cache()
        total_in += n_in
        total_out += n_out
        _, _, st = eng.get_state()
        assert list(st[:3]) == [int(x) for x in o.counters], "frame %d: counters after swap" % i
        assert parity.hash_equal(eng.read(capi.BUF_HASH), o.hash_entries), "frame %d: hash after swap" % i
        assert np.array_equal(eng.read(capi.BUF_SWAP_STATES), o.swap_states), "frame %d: swap states after swap" % i
        free_head = int(o.counters[1]) + 1
        assert np.array_equal(eng.read(capi.BUF_VBA_ALLOC_LIST)[:free_head], o.vba_alloc_list[:free_head]), "frame %d: free list" % i
        assert np.array_equal(eng.read(capi.BUF_VOXELS) & vmask, o.voxels & vmask), "frame %d: voxels after swap" % i
        assert np.array_equal(has, o.has_stored_data), "frame %d: hasStoredData" % i
        for entry in np.nonzero(o.has_stored_data != stored_before)[0][:64]:
            assert np.array_equal(blocks[entry] & vmask, o.stored_voxel_block(entry) & vmask), "frame %d: stored block %d" % (i, entry)
        # ---- raycast for the next frame's tracking
        o.expected_depths()
        eng.RunStage(capi.STAGE_EXPECTED_DEPTHS)
        o.icp_maps()
        eng.RunStage(capi.STAGE_ICP_MAPS)
        assert np.array_equal(eng.read_image(capi.BUF_RAYCAST_RESULT, 4), o.raycast_result), "frame %d: raycast" % i
    eng.close(); o.close()
    return total_in, total_out


@pytest.mark.skipif(not ref.available("parity"), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_swapping_matches_reference():
    total_in, total_out = _run("parity")
    assert total_out > 5000 and total_in > 5000  # blocks really went both ways


@pytest.mark.skipif(not ref.available("rgb"), reason="oracle/_ref/libitm_ref_rgb.so not built")
def test_swapping_with_colour_voxels_matches_reference():
    total_in, total_out = _run("rgb")
    assert total_out > 5000 and total_in > 5000


@pytest.mark.skipif(not ref.available("parity"), reason="oracle/_ref not built")
def test_process_frame_with_swapping_free_running():
    """ITMMainEngine::ProcessFrame with useSwapping through the host API (swap stage inside the frame): poses follow the
    reference and the same number of blocks is parked on the host"""
    w, h = 320, 240
    o = ref.RefEngine(w, h, use_swapping=True)
    p = capi.default_params(w, h)
    p.fx, p.fy, p.cx, p.cy = o.intr
    p.use_swapping = 1
    from infinitam_b200.engines import ITMMainEngine
    eng = ITMMainEngine(p)
    for k in range(0, 40, 2):
        depth = synth.render_depth(k, w, h)
        o.process_frame(depth)
        pose = eng.ProcessFrame(None, depth)
        rot, trans = parity.pose_diff(pose, o.pose_M)
        assert rot <= 1e-4 and trans <= 1e-4, "frame %d: %g rad %g m" % (k, rot, trans)
    has, _, _, _ = eng.global_cache()
    assert abs(int(has.sum()) - int(o.has_stored_data.sum())) <= 0.02 * max(1, int(o.has_stored_data.sum())) + 2
    eng.close(); o.close()
