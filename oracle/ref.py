"""ctypes view of oracle/_ref/libitm_ref*.so (the REAL reference CPU engines).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this module; the product (infinitam_b200/) never does.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)

HASH_ENTRY_DTYPE = np.dtype(
    {"names": ["pos", "offset", "ptr"], "formats": [("<i2", 3), "<i4", "<i4"], "offsets": [0, 8, 12], "itemsize": 16}
)
VOXEL_S_DTYPE = np.dtype({"names": ["sdf", "w_depth"], "formats": ["<i2", "u1"], "offsets": [0, 2], "itemsize": 4})


def lib_path(flavour: str = "parity") -> str:
    name = {"parity": "libitm_ref.so", "fast": "libitm_ref_fast.so", "fast1": "libitm_ref_fast1.so", "rgb": "libitm_ref_rgb.so"}[flavour]
    return os.path.join(REF_DIR, name)


def available(flavour: str = "parity") -> bool:
    return os.path.exists(lib_path(flavour))


_libs = {}


class _Prefixed:
    """view of a CDLL whose exported functions are named <prefix>_xxx, exposed as ref_xxx"""

    def __init__(self, lib, prefix):
        self._lib, self._prefix = lib, prefix

    def __getattr__(self, name):
        if name.startswith("ref_"):
            return getattr(self._lib, self._prefix + name[3:])
        return getattr(self._lib, name)


def load(flavour: str = "parity"):
    if flavour in _libs:
        return _libs[flavour]
    lib = _Prefixed(C.CDLL(lib_path(flavour)), "ref")
    lib.ref_const.restype = C.c_int
    lib.ref_const.argtypes = [C.c_char_p]
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_int, C.c_float, C.c_float]
    declare_common(lib)
    _libs[flavour] = lib
    return lib


def _has(lib, name):
    if isinstance(lib, _Prefixed):
        return hasattr(lib._lib, lib._prefix + name[3:])
    return hasattr(lib, name)


def declare_common(lib):
    for name in ("ref_destroy", "ref_track", "ref_integrate", "ref_expected_depths", "ref_icp_maps", "ref_prepare",
                 "ref_icp_prepare"):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = [C.c_void_p]
    if hasattr(lib._lib if isinstance(lib, _Prefixed) else lib, (lib._prefix if isinstance(lib, _Prefixed) else "ref") + "_set_rgb"):
        lib.ref_set_rgb.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_set_rgb.restype = None
        lib.ref_set_use_swapping.argtypes = [C.c_int]
        lib.ref_set_use_swapping.restype = None
        lib.ref_swap.argtypes = [C.c_void_p]
        lib.ref_swap.restype = None
        for name in ("ref_swap_states", "ref_has_stored_data", "ref_stored_voxel_blocks"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = C.c_void_p
    if _has(lib, "ref_forward_render"):
        lib.ref_set_use_approximate_raycast.argtypes = [C.c_void_p, C.c_int]
        lib.ref_set_use_approximate_raycast.restype = None
        lib.ref_forward_render.argtypes = [C.c_void_p]
        lib.ref_forward_render.restype = None
        for name in ("ref_requires_full_rendering", "ref_no_fwd_missing_points", "ref_free_no_visible"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = C.c_int
        for name in ("ref_forward_projection", "ref_fwd_missing_points", "ref_free_visible_ids", "ref_free_minmax",
                     "ref_free_raycast_result"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = C.c_void_p
        lib.ref_get_image.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_void_p]
        lib.ref_get_image.restype = C.c_int
    if _has(lib, "ref_create_point_cloud"):
        lib.ref_create_point_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_create_point_cloud.restype = C.c_int
    if _has(lib, "ref_low_level"):
        lib.ref_low_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.ref_low_level.restype = C.c_longlong
    if _has(lib, "ref_mesh_scene"):
        lib.ref_mesh_scene.argtypes = [C.c_void_p, C.POINTER(_f32p), _i32p]
        lib.ref_mesh_scene.restype = C.c_int
        if _has(lib, "ref_write_stl"):  # the C restatement has no file writers
            lib.ref_write_stl.argtypes = [C.c_void_p, C.c_char_p]
            lib.ref_write_stl.restype = None
            lib.ref_write_obj.argtypes = [C.c_void_p, C.c_char_p]
            lib.ref_write_obj.restype = None
    if _has(lib, "ref_set_tracker_wicp"):
        lib.ref_set_tracker_wicp.argtypes = [C.c_int, C.c_int]
        lib.ref_set_tracker_wicp.restype = None
        lib.ref_wicp_prepare.argtypes = [C.c_void_p]
        lib.ref_wicp_prepare.restype = None
        lib.ref_wicp_gandh.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
        lib.ref_wicp_gandh.restype = C.c_int
        for name in ("ref_depth_uncertainty", "ref_depth_normal"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = C.c_void_p
    lib.ref_update_view.argtypes = [C.c_void_p, C.c_void_p]
    lib.ref_update_view.restype = None
    lib.ref_process_frame.argtypes = [C.c_void_p, C.c_void_p]
    lib.ref_process_frame.restype = None
    lib.ref_process_frame_timed.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib.ref_process_frame_timed.restype = None
    lib.ref_allocate.argtypes = [C.c_void_p, C.c_int]
    lib.ref_allocate.restype = None
    lib.ref_icp_gandh.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
    lib.ref_icp_gandh.restype = C.c_int
    lib.ref_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.POINTER(_f32p), _i32p, _i32p, _f32p]
    lib.ref_icp_config.argtypes = [C.c_void_p, _i32p, _i32p, _f32p, _i32p]
    for name in ("ref_get_pose", "ref_set_pose", "ref_get_pose_pointcloud", "ref_set_pose_pointcloud",
                 "ref_get_pose_params"):
        getattr(lib, name).argtypes = [C.c_void_p, _f32p]
        getattr(lib, name).restype = None
    lib.ref_get_age.argtypes = [C.c_void_p]
    lib.ref_get_age.restype = C.c_int
    lib.ref_set_age.argtypes = [C.c_void_p, C.c_int]
    lib.ref_mat_inv.argtypes = [_f32p, _f32p]
    lib.ref_pose_from_invm_coerced.argtypes = [_f32p, _f32p, _f32p, _f32p]
    lib.ref_pose_from_params.argtypes = [_f32p, _f32p]
    lib.ref_compute_delta.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, _f32p]
    for name in ("ref_hash_entries", "ref_voxels", "ref_vba_alloc_list", "ref_excess_alloc_list", "ref_visible_ids",
                 "ref_visible_types", "ref_depth", "ref_minmax", "ref_raycast_result", "ref_raycast_image",
                 "ref_points", "ref_normals"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = C.c_void_p
    lib.ref_get_counters.argtypes = [C.c_void_p, _i32p]
    lib.ref_set_counters.argtypes = [C.c_void_p, _i32p]
    return lib


def _low_level(fn, handle, op, image, prefill):
    h, w = image.shape[:2]
    src = np.ascontiguousarray(image, dtype=np.float32 if op == 2 else np.uint8)
    assert src.shape == (h, w, 4)
    if op in (1, 2):
        out = np.zeros((h // 2, w // 2, 4), dtype=src.dtype)
    elif op == 0:
        out = np.zeros((h, w, 4), dtype=np.uint8)
    else:
        out = np.zeros((h, w, 4), dtype=np.int16)
    n = fn(handle, int(op), src.ctypes.data, w, h, out.ctypes.data, int(prefill))
    if n != out.nbytes:
        raise RuntimeError("low_level(%d) returned %d, expected %d bytes" % (op, n, out.nbytes))
    return out


def _fp(a):
    return a.ctypes.data_as(_f32p)


@contextlib.contextmanager
def _stdout_to_stderr():
    """C-level stdout of the reference goes to stderr while inside (bench.py must print exactly one JSON line on stdout)"""
    sys.stdout.flush()
    saved = os.dup(1)
    try:
        os.dup2(2, 1)
        yield
    finally:
        try:
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(saved, 1)
        os.close(saved)


def _view(ptr, dtype, count):
    dtype = np.dtype(dtype)
    buf = (C.c_char * (count * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class RefEngine:
    """The reference CPU engines composed like ITMMainEngine (ITMLib/Engine/ITMMainEngine.cpp:17-127)."""

    def __init__(self, width=640, height=480, intr=None, voxel_size=0.005, mu=0.02, max_w=100, vf_min=0.35,
                 vf_max=3.0, flavour="parity", use_swapping=False, wicp=False, bilateral=False):
        from infinitam_b200 import synth  # numpy-only helper

        self.use_swapping = use_swapping
        self.wicp, self.bilateral = wicp, bilateral
        self.W, self.H = width, height
        self.intr = tuple(float(x) for x in (intr or synth.intrinsics_for(width, height)))
        self.voxel_size, self.mu, self.max_w, self.vf_min, self.vf_max = voxel_size, mu, max_w, vf_min, vf_max
        self._create(flavour)
        self.n_entries = self.n_bucket + self.n_excess

    def _create(self, flavour):
        self.lib = load(flavour)
        if self.wicp or self.bilateral:
            self.lib.ref_set_tracker_wicp(int(self.wicp), int(self.bilateral))
        if self.use_swapping:
            self.lib.ref_set_use_swapping(1)
        with _stdout_to_stderr():  # ITMLibSettings() prints its tracker type on std::cout (ITMLibSettings.cpp:85-86)
            self.h = self.lib.ref_create(self.W, self.H, *self.intr, self.voxel_size, self.mu, self.max_w, self.vf_min, self.vf_max)
        if self.wicp or self.bilateral:
            self.lib.ref_set_tracker_wicp(0, 0)
        if self.use_swapping:
            self.lib.ref_set_use_swapping(0)
        c = self.const
        self.n_local = c("SDF_LOCAL_BLOCK_NUM")
        self.n_bucket = c("SDF_BUCKET_NUM")
        self.n_excess = c("SDF_EXCESS_LIST_SIZE")

    def const(self, name):
        return self.lib.ref_const(name.encode())

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stages -------------------------------------------------------------
    def set_rgb(self, rgba_u8):
        """view->rgb of the following frames, (H, W, 4) uint8"""
        a = np.ascontiguousarray(rgba_u8, dtype=np.uint8)
        assert a.size == self.W * self.H * 4
        self.lib.ref_set_rgb(self.h, a.ctypes.data)

    def update_view(self, depth_i16):
        d = np.ascontiguousarray(depth_i16, dtype=np.int16)
        self.lib.ref_update_view(self.h, d.ctypes.data)

    def track(self):
        self.lib.ref_track(self.h)

    def allocate(self, only_visible=False):
        self.lib.ref_allocate(self.h, int(only_visible))

    def integrate(self):
        self.lib.ref_integrate(self.h)

    def swap(self):
        """ITMSwappingEngine::IntegrateGlobalIntoLocal + SaveToGlobalMemory"""
        self.lib.ref_swap(self.h)

    @property
    def swap_states(self):
        return _view(self.lib.ref_swap_states(self.h), np.uint8, self.n_entries)

    @property
    def has_stored_data(self):
        return _view(self.lib.ref_has_stored_data(self.h), np.uint8, self.n_entries)

    def stored_voxel_block(self, entry):
        """one block of the global cache (512 voxel words)"""
        dt = np.dtype(np.uint64 if self.const("sizeof_voxel") == 8 else np.uint32)
        return _view(self.lib.ref_stored_voxel_blocks(self.h) + int(entry) * 512 * dt.itemsize, dt, 512)

    def expected_depths(self):
        self.lib.ref_expected_depths(self.h)

    def icp_maps(self):
        self.lib.ref_icp_maps(self.h)

    # ---- useApproximateRaycast / ForwardRender / GetImage ---------------------------------
    def set_use_approximate_raycast(self, on=True):
        self.lib.ref_set_use_approximate_raycast(self.h, int(on))

    @property
    def requires_full_rendering(self):
        return bool(self.lib.ref_requires_full_rendering(self.h))

    def forward_render(self):
        """IITMVisualisationEngine::ForwardRender + age_pointCloud++ (ITMTrackingController.cpp:40-44)"""
        self.lib.ref_forward_render(self.h)

    def prepare(self):
        """ITMTrackingController::Prepare"""
        self.lib.ref_prepare(self.h)

    @property
    def forward_projection(self):
        return self._img(self.lib.ref_forward_projection, np.float32, 4)

    @property
    def fwd_missing_points(self):
        n = self.lib.ref_no_fwd_missing_points(self.h)
        return _view(self.lib.ref_fwd_missing_points(self.h), np.int32, self.W * self.H)[:n]

    def create_point_cloud(self, trafo_rgb_to_depth=None, skip_points=False):
        """Prepare()'s TRACKER_COLOR branch: (locations[n, 4], colours[n, 4]); overwrites the ICP maps and the render state"""
        T = None if trafo_rgb_to_depth is None else np.ascontiguousarray(trafo_rgb_to_depth, dtype=np.float32).reshape(16)
        n = self.lib.ref_create_point_cloud(self.h, None if T is None else T.ctypes.data, 1 if skip_points else 0)
        if n < 0:
            raise RuntimeError("ref_create_point_cloud: no view yet")
        return self.points.reshape(-1, 4)[:n].copy(), self.normals.reshape(-1, 4)[:n].copy()

    def low_level(self, op, image, prefill=0):
        """ITMLowLevelEngine_CPU helper `op` (0 CopyImage, 1 FilterSubsample, 2 FilterSubsampleWithHoles(Vector4f), 3 GradientX,
        4 GradientY) on an (h, w, 4) image; returns the output image (uint8 / float32 / int16, 4 channels)"""
        return _low_level(self.lib.ref_low_level, self.h, op, image, prefill)

    def get_image(self, image_type, pose_M=None, intr=None, width=None, height=None):
        """ITMMainEngine::GetImage; returns (h, w, 4) uint8 or None when there is no view yet"""
        w, h = width or self.W, height or self.H
        out = np.zeros((h, w, 4), np.uint8)
        M = np.ascontiguousarray(pose_M if pose_M is not None else np.eye(4).T.reshape(16), dtype=np.float32).reshape(16)
        k = np.ascontiguousarray(intr if intr is not None else self.intr, dtype=np.float32).reshape(4)
        rc = self.lib.ref_get_image(self.h, int(image_type), _fp(M), _fp(k), w, h, out.ctypes.data)
        self._free_dims = (w, h)
        return out if rc == 0 else None

    @property
    def free_visible_ids(self):
        n = self.lib.ref_free_no_visible(self.h)
        return _view(self.lib.ref_free_visible_ids(self.h), np.int32, self.n_local)[:n]

    @property
    def free_minmax(self):
        w, h = self._free_dims
        return _view(self.lib.ref_free_minmax(self.h), np.float32, w * h * 2).reshape(h, w, 2)

    @property
    def free_raycast_result(self):
        w, h = self._free_dims
        return _view(self.lib.ref_free_raycast_result(self.h), np.float32, w * h * 4).reshape(h, w, 4)

    # ---- weighted ICP (wicp=True engines) ----------------------------------------------------
    def wicp_prepare(self):
        self.lib.ref_wicp_prepare(self.h)

    def wicp_gandh(self, level, approx_inv_pose16):
        inv = np.ascontiguousarray(approx_inv_pose16, dtype=np.float32).reshape(16)
        out = np.zeros(44, dtype=np.float32)
        n = self.lib.ref_wicp_gandh(self.h, level, _fp(inv), _fp(out))
        return n, out

    @property
    def depth_uncertainty(self):
        return self._img(self.lib.ref_depth_uncertainty, np.float32, 1)

    @property
    def depth_normal(self):
        return self._img(self.lib.ref_depth_normal, np.float32, 4)

    # ---- meshing -----------------------------------------------------------------------
    def mesh_scene(self, whole_buffer=False):
        """ITMMeshingEngine::MeshScene; returns the (n, 9) float32 triangle array (a view onto the reference's mesh; the
        whole noMaxTriangles buffer when whole_buffer)"""
        p, nmax = _f32p(), C.c_int()
        n = self.lib.ref_mesh_scene(self.h, C.byref(p), C.byref(nmax))
        self.no_max_triangles = nmax.value
        rows = nmax.value if whole_buffer else n
        arr = np.ctypeslib.as_array(p, shape=(nmax.value, 9))
        self.no_total_triangles = n
        return arr[:rows]

    def write_stl(self, path):
        self.lib.ref_write_stl(self.h, str(path).encode())

    def write_obj(self, path):
        self.lib.ref_write_obj(self.h, str(path).encode())

    def process_frame(self, depth_i16):
        d = np.ascontiguousarray(depth_i16, dtype=np.int16)
        self.lib.ref_process_frame(self.h, d.ctypes.data)

    def process_frame_timed(self, depth_i16):
        d = np.ascontiguousarray(depth_i16, dtype=np.int16)
        ms = (C.c_double * 6)()
        self.lib.ref_process_frame_timed(self.h, d.ctypes.data, ms)
        return list(ms)

    # ---- ICP ----------------------------------------------------------------
    def icp_prepare(self):
        self.lib.ref_icp_prepare(self.h)

    def icp_gandh(self, level, approx_inv_pose16):
        inv = np.ascontiguousarray(approx_inv_pose16, dtype=np.float32).reshape(16)
        out = np.zeros(44, dtype=np.float32)
        n = self.lib.ref_icp_gandh(self.h, level, _fp(inv), _fp(out))
        return n, out

    def pyramid_level(self, level):
        p = _f32p()
        w, h = C.c_int(), C.c_int()
        intr = np.zeros(4, dtype=np.float32)
        self.lib.ref_pyramid_level(self.h, level, C.byref(p), C.byref(w), C.byref(h), _fp(intr))
        arr = np.ctypeslib.as_array(p, shape=(h.value, w.value))
        return arr, intr

    def icp_config(self):
        n = C.c_int()
        iters = np.zeros(8, dtype=np.int32)
        thr = np.zeros(8, dtype=np.float32)
        typ = np.zeros(8, dtype=np.int32)
        self.lib.ref_icp_config(self.h, C.byref(n), iters.ctypes.data_as(_i32p), _fp(thr), typ.ctypes.data_as(_i32p))
        k = n.value
        return k, iters[:k].copy(), thr[:k].copy(), typ[:k].copy()

    # ---- pose ---------------------------------------------------------------
    def _get16(self, fn):
        m = np.zeros(16, dtype=np.float32)
        fn(self.h, _fp(m))
        return m

    @property
    def pose_M(self):
        """column-major 16 floats (ORUtils/Matrix.h:8-33)"""
        return self._get16(self.lib.ref_get_pose)

    @pose_M.setter
    def pose_M(self, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
        self.lib.ref_set_pose(self.h, _fp(m))

    @property
    def pose_pointcloud_M(self):
        return self._get16(self.lib.ref_get_pose_pointcloud)

    @pose_pointcloud_M.setter
    def pose_pointcloud_M(self, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
        self.lib.ref_set_pose_pointcloud(self.h, _fp(m))

    @property
    def pose_params(self):
        p = np.zeros(6, dtype=np.float32)
        self.lib.ref_get_pose_params(self.h, _fp(p))
        return p

    @property
    def age(self):
        return self.lib.ref_get_age(self.h)

    @age.setter
    def age(self, a):
        self.lib.ref_set_age(self.h, int(a))

    def mat_inv(self, m16):
        a = np.ascontiguousarray(m16, dtype=np.float32).reshape(16)
        o = np.zeros(16, dtype=np.float32)
        self.lib.ref_mat_inv(_fp(a), _fp(o))
        return o

    def pose_from_invm_coerced(self, inv16):
        a = np.ascontiguousarray(inv16, dtype=np.float32).reshape(16)
        M, inv, p = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(6, np.float32)
        self.lib.ref_pose_from_invm_coerced(_fp(a), _fp(M), _fp(inv), _fp(p))
        return M, inv, p

    def compute_delta(self, nabla6, hess36, short_iteration):
        n = np.ascontiguousarray(nabla6, dtype=np.float32).reshape(6)
        h = np.ascontiguousarray(hess36, dtype=np.float32).reshape(36)
        s = np.zeros(6, np.float32)
        self.lib.ref_compute_delta(self.h, _fp(n), _fp(h), int(short_iteration), _fp(s))
        return s

    # ---- raw state (numpy views onto the reference's own buffers) -------------
    @property
    def hash_entries(self):
        return _view(self.lib.ref_hash_entries(self.h), HASH_ENTRY_DTYPE, self.n_entries)

    @property
    def voxels(self):
        """raw view, one word per voxel: uint32 for ITMVoxel_s (sdf | w_depth<<16 | pad<<24), uint64 for ITMVoxel_s_rgb
        (sdf | w_depth<<16 | clr.r<<24 | clr.g<<32 | clr.b<<40 | w_color<<48 | pad<<56)"""
        dt = np.uint64 if self.const("sizeof_voxel") == 8 else np.uint32
        return _view(self.lib.ref_voxels(self.h), dt, self.n_local * 512)

    @property
    def vba_alloc_list(self):
        return _view(self.lib.ref_vba_alloc_list(self.h), np.int32, self.n_local)

    @property
    def excess_alloc_list(self):
        return _view(self.lib.ref_excess_alloc_list(self.h), np.int32, self.n_excess)

    @property
    def visible_ids(self):
        return _view(self.lib.ref_visible_ids(self.h), np.int32, self.n_local)

    @property
    def visible_types(self):
        return _view(self.lib.ref_visible_types(self.h), np.uint8, self.n_entries)

    @property
    def counters(self):
        c = np.zeros(3, dtype=np.int32)
        self.lib.ref_get_counters(self.h, c.ctypes.data_as(_i32p))
        return c

    @counters.setter
    def counters(self, c):
        c = np.ascontiguousarray(c, dtype=np.int32)
        self.lib.ref_set_counters(self.h, c.ctypes.data_as(_i32p))

    def _img(self, fn, dtype, ch):
        p = fn(self.h)
        if not p:
            return None
        a = _view(p, dtype, self.W * self.H * ch)
        return a.reshape(self.H, self.W, ch) if ch > 1 else a.reshape(self.H, self.W)

    @property
    def depth(self):
        return self._img(self.lib.ref_depth, np.float32, 1)

    @property
    def minmax(self):
        return self._img(self.lib.ref_minmax, np.float32, 2)

    @property
    def raycast_result(self):
        return self._img(self.lib.ref_raycast_result, np.float32, 4)

    @property
    def raycast_image(self):
        return self._img(self.lib.ref_raycast_image, np.uint8, 4)

    @property
    def points(self):
        return self._img(self.lib.ref_points, np.float32, 4)

    @property
    def normals(self):
        return self._img(self.lib.ref_normals, np.float32, 4)
