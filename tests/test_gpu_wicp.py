"""-m gpu: SURVEY.md 8f row 4 - the weighted ICP tracker (TRACKER_WICP) and the view builder's sensor-noise model and
bilateral depth filter, through the drop-in boundary: the reference's own ITMWeightedICPTracker host loop, image hierarchies
and ITMView (CUDA memory placement) over the adapter's ComputeGandH / DepthFiltering / ComputeNormalAndWeights, against
ITMWeightedICPTracker_CPU / ITMViewBuilder_CPU.

Tolerances.  filterDepth and computeNormalAndWeight go through exp / acos, where CUDA's and glibc's libm differ by an ulp or
two, so these rows are compared with tolerances instead of bit for bit: filtered depth 1e-5 m after the five passes (1 % of
the sensor's 1 mm quantum), sigma_z 1e-7 absolute, normals 1e-5 (1e-3 behind the filter); one evaluation: same number of
valid points, f / gradient / Hessian to 1e-3 relative (fp32 sums in another order); tracked pose 1e-4 rad / 1e-4 m per frame
(BASELINE.json) - 5e-4 behind the bilateral filter, whose ulp-level input differences the undamped Gauss-Newton loop of this
tracker amplifies."""
import numpy as np
import pytest

import parity
from infinitam_b200 import synth
from oracle import adapter, ref

pytestmark = pytest.mark.gpu

needs_libs = pytest.mark.skipif(not (adapter.available() and ref.available("parity")),
                                reason="oracle/_ref not built (needs /root/reference at build time)")


@needs_libs
@pytest.mark.parametrize("bilateral", [False, True], ids=["sensor-noise-model", "bilateral-filter+sensor-noise-model"])
def test_view_builder_filters(bilateral):
    w, h = 320, 240
    depth = synth.sequence(2, w, h, noise=True)[1]   # noise and dropped pixels: holes inside the filter windows
    o = ref.RefEngine(w, h, wicp=True, bilateral=bilateral)
    a = adapter.AdapterEngine(w, h, intr=o.intr, wicp=True, bilateral=bilateral)
    o.update_view(depth)
    a.update_view(depth)
    d_o, d_a = o.depth, a.read(adapter.READ_DEPTH).reshape(h, w)
    assert np.array_equal(d_a < 0, d_o < 0) and np.array_equal(d_a == 0, d_o == 0), "hole / border pattern of the depth image differs"
    assert np.abs(d_a - d_o).max() <= (1e-5 if bilateral else 0.0), "depth differs by %g m" % np.abs(d_a - d_o).max()
    s_o, s_a = o.depth_uncertainty, a.read(adapter.READ_DEPTH_UNCERTAINTY).reshape(h, w)
    n_o, n_a = o.depth_normal, a.read(adapter.READ_DEPTH_NORMAL).reshape(h, w, 4)
    assert np.array_equal(s_a < 0, s_o < 0), "validity of sigma_z differs"
    assert np.array_equal(n_a[..., 3], n_o[..., 3]), "validity of the depth normals differs"
    # sigma_z = 0.0012 + 0.0019 (z - 0.4)^2 + 1e-4 / sqrt(z) * (theta / (pi/2 - theta))^2 is ill conditioned by construction: NaN when
    # rounding leaves the normalised z component a hair above 1 (acos), +inf / huge at grazing angles (theta -> pi/2).  The
    # reference has the same singularities; what the tracker consumes is the weight 0.0012 / sigma_z * 0.5 + 0.5 (0 when sigma_z
    # is not positive, ITMWeightedICPTracker_CPU.cpp:46), which is well conditioned - compare that, and sigma_z itself where it
    # is in the sensor's normal range.
    if not bilateral:
        assert np.array_equal(np.isnan(s_a), np.isnan(s_o)) and np.array_equal(np.isinf(s_a), np.isinf(s_o)), "NaN / inf pattern differs"
    assert (~np.isfinite(s_o)).mean() < 0.01

    def weight(sz):
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(sz > 0, np.float32(0.0012) / sz * np.float32(0.5) + np.float32(0.5), np.float32(0.0))

    dw = np.abs(weight(s_a) - weight(s_o))
    ok = np.isfinite(s_o) & np.isfinite(s_a) & (s_o > 0) & (s_o < 0.05)
    assert ok.sum() > 0.5 * w * h
    rel = np.abs(s_a - s_o)[ok] / s_o[ok]
    dn = np.abs(n_a - n_o)[ok].max(axis=-1)
    if not bilateral:   # identical input depth: only acosf separates the two sides
        assert dw.max() <= 1e-6, "ICP weights differ by %g" % dw.max()
        assert rel.max() <= 1e-5, "sigma_z differs by %g relative" % rel.max()
        assert dn.max() <= 1e-5, "normals differ by %g" % dn.max()
    else:
        # the filtered depth differs by a few ulp, and the reference's normal (x and y are "unprojected" by MULTIPLYING with
        # the focal length, ITMViewBuilder.h:87-90) cancels catastrophically where the surface is flat along an image axis: a
        # handful of pixels move visibly.  Bound the bulk tightly and the outliers loosely.
        assert np.percentile(dw, 99.9) <= 1e-3 and dw.max() <= 0.25, "ICP weights: p99.9 %g max %g" % (np.percentile(dw, 99.9), dw.max())
        assert np.percentile(rel, 99.9) <= 2e-2, "sigma_z: p99.9 relative difference %g" % np.percentile(rel, 99.9)
        assert np.percentile(dn, 99.9) <= 2e-2, "normals: p99.9 difference %g" % np.percentile(dn, 99.9)
    a.close(); o.close()


@needs_libs
def test_weighted_single_evaluations():
    """frame 0 is fused at the identity pose on both sides (identical scenes and ICP maps); then every pyramid level of
    frame 1 is evaluated once at the same pose"""
    w, h = 320, 240
    seq = synth.sequence(2, w, h)
    o = ref.RefEngine(w, h, wicp=True)
    a = adapter.AdapterEngine(w, h, intr=o.intr, wicp=True)
    o.process_frame(seq[0])
    a.process_frame(seq[0])
    assert np.array_equal(a.read(adapter.READ_POINTS).reshape(h, w, 4), o.points)
    o.update_view(seq[1])
    a.update_view(seq[1])
    o.wicp_prepare()
    inv = o.mat_inv(o.pose_M)
    for level in range(5):
        n_o, g_o = o.wicp_gandh(level, inv)
        n_a, g_a = a.wicp_gandh(level, inv)
        assert n_a == n_o and n_o > 100, "level %d: valid points %d vs %d" % (level, n_a, n_o)
        assert abs(g_a[1] - g_o[1]) <= 1e-3 * abs(g_o[1]), "level %d: f %g vs %g" % (level, g_a[1], g_o[1])
        scale = np.abs(g_o[2:]).max()
        assert np.abs(g_a[2:] - g_o[2:]).max() <= 1e-3 * scale, "level %d: gradient / Hessian differ by %g (scale %g)" % (
            level, np.abs(g_a[2:] - g_o[2:]).max(), scale)
    a.close(); o.close()


@needs_libs
@pytest.mark.parametrize("bilateral", [False, True], ids=["wicp", "wicp+bilateral-filter"])
def test_weighted_icp_free_running(bilateral):
    w, h, n = 320, 240, 8
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h, wicp=True, bilateral=bilateral)
    a = adapter.AdapterEngine(w, h, intr=o.intr, wicp=True, bilateral=bilateral)
    for k in range(n):
        o.process_frame(seq[k])
        a.process_frame(seq[k])
        rot, trans = parity.pose_diff(a.pose_M, o.pose_M)
        tol = 5e-4 if bilateral else 1e-4
        assert rot <= tol and trans <= tol, "frame %d pose differs: %g rad %g m" % (k, rot, trans)
    ca, co = a.counters, o.counters
    assert abs(int(ca[0]) - int(co[0])) <= 0.01 * co[0] + 2
    moved = np.abs(o.pose_M - np.eye(4, dtype=np.float32).T.reshape(16)).max()
    assert moved > 0.02, "the tracker was supposed to follow the camera (moved %g)" % moved
    a.close(); o.close()


@needs_libs
@pytest.mark.parametrize("bilateral", [False, True], ids=["wicp", "wicp+bilateral-filter"])
def test_weighted_icp_in_the_device_loop_layer_b(bilateral):
    """TRACKER_WICP inside Layer B (no host round trip per evaluation: bilateral filter, sigma_z, both pyramids and the whole
    Gauss-Newton loop are enqueued with the frame) against the reference CPU engines, free running"""
    from infinitam_b200 import capi
    w, h, n = 320, 240, 8
    seq = synth.sequence(n, w, h)
    o = ref.RefEngine(w, h, wicp=True, bilateral=bilateral)
    p = parity.cuda_params(o)
    p.tracker_type = capi.TRACKER_WICP
    p.use_bilateral_filter = 1 if bilateral else 0
    from infinitam_b200.engines import ITMMainEngine
    eng = ITMMainEngine(p)
    tol = 5e-4 if bilateral else 1e-4
    for k in range(n):
        o.process_frame(seq[k])
        pose = eng.ProcessFrame(None, seq[k])
        rot, trans = parity.pose_diff(pose, o.pose_M)
        assert rot <= tol and trans <= tol, "frame %d pose differs: %g rad %g m" % (k, rot, trans)
        if k > 0:
            assert int(eng.icp_stats().sum()) >= 5  # every level evaluated at least once
    _, _, st = eng.get_state()
    assert abs(int(st[0]) - int(o.counters[0])) <= 0.01 * o.counters[0] + 2
    eng.close()
    o.close()
