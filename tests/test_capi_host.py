"""CPU (-m "not gpu"): the C-ABI library loads, exports every declared symbol, refuses to run without a GPU, and its
host-side pose helpers agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from infinitam_b200 import capi
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "itm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(itm_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, "declared in include/itm_b200.h but not exported: %s" % missing
    assert sorted(capi.SYMBOLS) == declared, "capi.SYMBOLS out of sync with the header"


def test_default_params_are_the_reference_defaults():
    p = capi.default_params(640, 480)
    assert (p.width, p.height) == (640, 480)
    assert (p.fx, p.fy, p.cx, p.cy) == (580.0, 580.0, 320.0, 240.0)  # ITMIntrinsics.h:49
    assert abs(p.voxel_size - 0.005) < 1e-9 and abs(p.mu - 0.02) < 1e-9 and p.max_w == 100  # ITMLibSettings.cpp:10
    assert (p.sdf_local_block_num, p.sdf_bucket_num, p.sdf_excess_list_size) == (0x10000, 0x100000, 0x20000)
    assert list(p.tracking_regime)[:5] == [3, 3, 1, 1, 1] and p.no_hierarchy_levels == 5


def test_no_cpu_fallback():
    lib = capi.load()
    if lib.itm_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    p = capi.default_params(64, 48)
    rc = lib.itm_b200_engine_create(C.byref(p), C.byref(h))
    assert rc == capi.ENODEVICE and not h.value
    assert b"no CPU fallback" in lib.itm_b200_last_error()
    rc = lib.itm_b200_ctx_create(C.byref(p), None, C.byref(h))
    assert rc == capi.ENODEVICE
    from infinitam_b200.engines import ITMMainEngine
    with pytest.raises(capi.ItmError):
        ITMMainEngine(p)


def test_invalid_params_rejected():
    lib = capi.load()
    h = C.c_void_p()
    p = capi.default_params(64, 48)
    p.sdf_bucket_num = 1000  # not a power of two
    assert lib.itm_b200_engine_create(C.byref(p), C.byref(h)) == capi.EINVAL
    p = capi.default_params(0, 48)
    assert lib.itm_b200_engine_create(C.byref(p), C.byref(h)) == capi.EINVAL
    assert lib.itm_b200_engine_create(None, C.byref(h)) == capi.EINVAL


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def test_host_pose_helpers_match_oracle():
    lib = capi.load()
    o = port.PortEngine(64, 48)
    rng = np.random.default_rng(3)
    for _ in range(2000):
        p6 = np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.5, 3)]).astype(np.float32)
        M = np.zeros(16, np.float32)
        o.lib.ref_pose_from_params(_f(p6), _f(M))
        inv_o = o.mat_inv(M)
        inv_c = np.zeros(16, np.float32)
        assert lib.itm_b200_mat4_inv(_f(M), _f(inv_c)) == 0
        assert np.array_equal(inv_o, inv_c)
        Mo, io, po = o.pose_from_invm_coerced(inv_o)
        Mc, ic, pc = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(6, np.float32)
        lib.itm_b200_pose_from_inv_m_coerced(_f(inv_o), _f(Mc), _f(ic), _f(pc))
        assert np.array_equal(Mo, Mc) and np.array_equal(io, ic) and np.array_equal(po, pc)
        # Cholesky step on a random SPD system
        A = rng.normal(size=(6, 6)).astype(np.float32)
        Hm = (A @ A.T + 6 * np.eye(6, dtype=np.float32)).astype(np.float32)
        g = rng.normal(size=6).astype(np.float32)
        for short in (0, 1):
            so = o.compute_delta(g, Hm.reshape(36), short)
            sc = np.zeros(6, np.float32)
            lib.itm_b200_compute_delta(_f(g), _f(np.ascontiguousarray(Hm.reshape(36))), short, _f(sc))
            assert np.array_equal(so, sc)
    o.close()


def test_write_stl_obj_match_the_reference(tmp_path):
    """itm_b200_write_stl / _write_obj are host-only entry points: on the triangle array of the reference's own MeshScene they
    must write the very bytes ITMMesh::WriteSTL / WriteOBJ write (Objects/ITMMesh.h:34-118)"""
    import filecmp

    import numpy as np
    import pytest

    from infinitam_b200 import capi, synth
    from oracle import ref
    if not ref.available("parity"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    o = ref.RefEngine(160, 120)
    o.process_frame(synth.sequence(1, 160, 120)[0])
    tri = np.ascontiguousarray(o.mesh_scene()[:20000], dtype=np.float32)
    o.lib.ref_write_stl(o.h, str(tmp_path / "ref_full.stl").encode())
    lib = capi.load()
    full = np.ascontiguousarray(o.mesh_scene(), dtype=np.float32)
    capi.check(lib.itm_b200_write_stl(str(tmp_path / "ours.stl").encode(), full.ctypes.data, len(full)))
    assert filecmp.cmp(tmp_path / "ref_full.stl", tmp_path / "ours.stl", shallow=False)
    o.write_obj(tmp_path / "ref.obj")
    capi.check(lib.itm_b200_write_obj(str(tmp_path / "ours.obj").encode(), full.ctypes.data, len(full)))
    assert filecmp.cmp(tmp_path / "ref.obj", tmp_path / "ours.obj", shallow=False)
    assert len(tri) > 1000
    o.close()
