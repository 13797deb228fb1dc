"""ctypes view of oracle/_ref/libitm_adapter.so: the reference's own host objects (CUDA memory) driven through
include/itm_b200_adapter.hpp and the C ABI of libitm_b200.so (oracle/adapter_harness.cpp).

TEST INFRASTRUCTURE.  Only tests/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .ref import HASH_ENTRY_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libitm_adapter.so")

(READ_HASH, READ_VOXELS, READ_VISIBLE_IDS, READ_RAYCAST, READ_POINTS, READ_NORMALS, READ_VISIBLE_TYPES, READ_RAYCAST_IMAGE,
 READ_DEPTH, READ_DEPTH_UNCERTAINTY, READ_DEPTH_NORMAL) = range(11)
_DTYPES = {READ_HASH: HASH_ENTRY_DTYPE, READ_VOXELS: np.uint32, READ_VISIBLE_IDS: np.int32, READ_RAYCAST: np.float32,
           READ_POINTS: np.float32, READ_NORMALS: np.float32, READ_VISIBLE_TYPES: np.uint8, READ_RAYCAST_IMAGE: np.uint8,
           READ_DEPTH: np.float32, READ_DEPTH_UNCERTAINTY: np.float32, READ_DEPTH_NORMAL: np.float32}


def available() -> bool:
    return os.path.exists(LIB)


class AdapterEngine:
    """ITMMainEngine::ProcessFrame composed from reference host objects + the B200 adapter engines."""

    def __init__(self, w, h, intr=None, voxel_size=0.005, mu=0.02, max_w=100, vf_min=0.35, vf_max=3.0, device_loop=True,
                 use_swapping=False, wicp=False, bilateral=False):
        lib = C.CDLL(LIB)
        lib.adp_set_tracker_wicp.argtypes = [C.c_int, C.c_int]
        lib.adp_update_view.argtypes = [C.c_void_p, C.c_void_p]
        lib.adp_wicp_gandh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.adp_set_use_swapping.argtypes = [C.c_int]
        lib.adp_count_stored.argtypes = [C.c_void_p]
        lib.adp_create.restype = C.c_void_p
        lib.adp_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_int, C.c_float, C.c_float, C.c_int]
        lib.adp_destroy.argtypes = [C.c_void_p]
        lib.adp_process_frame.argtypes = [C.c_void_p, C.c_void_p]
        lib.adp_get_pose.argtypes = [C.c_void_p, C.c_void_p]
        lib.adp_counters.argtypes = [C.c_void_p, C.c_void_p]
        lib.adp_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
        lib.adp_read.restype = C.c_longlong
        lib.adp_last_error.restype = C.c_char_p
        lib.adp_set_use_approximate_raycast.argtypes = [C.c_void_p, C.c_int]
        lib.adp_requires_full_rendering.argtypes = [C.c_void_p]
        lib.adp_get_free_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.lib, self.W, self.H = lib, w, h
        s = w / 640.0
        fx, fy, cx, cy = intr if intr is not None else (580.0 * s, 580.0 * s, w / 2.0, h / 2.0)
        lib.adp_set_use_swapping(int(use_swapping))
        lib.adp_set_tracker_wicp(int(wicp), int(bilateral))
        self.h = lib.adp_create(w, h, fx, fy, cx, cy, voxel_size, mu, max_w, vf_min, vf_max, int(device_loop))
        lib.adp_set_use_swapping(0)
        lib.adp_set_tracker_wicp(0, 0)
        if not self.h:
            raise RuntimeError("adp_create failed: %s" % lib.adp_last_error().decode())

    def close(self):
        if self.h:
            self.lib.adp_destroy(self.h)
            self.h = None

    def process_frame(self, depth_i16):
        d = np.ascontiguousarray(depth_i16, np.int16)
        if self.lib.adp_process_frame(self.h, d.ctypes.data) != 0:
            raise RuntimeError("adp_process_frame: %s" % self.lib.adp_last_error().decode())

    def set_use_approximate_raycast(self, on=True):
        self.lib.adp_set_use_approximate_raycast(self.h, int(on))

    def update_view(self, depth_i16):
        d = np.ascontiguousarray(depth_i16, np.int16)
        if self.lib.adp_update_view(self.h, d.ctypes.data) != 0:
            raise RuntimeError("adp_update_view: %s" % self.lib.adp_last_error().decode())

    def wicp_gandh(self, level, approx_inv_pose16):
        inv = np.ascontiguousarray(approx_inv_pose16, np.float32).reshape(16)
        out = np.zeros(44, np.float32)
        n = self.lib.adp_wicp_gandh(self.h, level, inv.ctypes.data, out.ctypes.data)
        if n < 0:
            raise RuntimeError("adp_wicp_gandh: %s" % self.lib.adp_last_error().decode())
        return n, out

    @property
    def stored_blocks(self):
        """number of voxel blocks parked in the reference's host-side ITMGlobalCache"""
        return self.lib.adp_count_stored(self.h)

    def create_point_cloud(self, trafo_rgb_to_depth=None, skip_points=False):
        """Prepare()'s TRACKER_COLOR branch through the adapter: (locations[n, 4], colours[n, 4])"""
        self.lib.adp_create_point_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        T = None if trafo_rgb_to_depth is None else np.ascontiguousarray(trafo_rgb_to_depth, dtype=np.float32).reshape(16)
        n = self.lib.adp_create_point_cloud(self.h, None if T is None else T.ctypes.data, 1 if skip_points else 0)
        if n < 0:
            raise RuntimeError("adp_create_point_cloud: %s" % self.lib.adp_last_error().decode())
        return self.read(READ_POINTS).reshape(-1, 4)[:n].copy(), self.read(READ_NORMALS).reshape(-1, 4)[:n].copy()

    def low_level(self, op, image, prefill=0):
        """the same helper through ITMLowLevelEngine_B200 (see oracle.ref.RefEngine.low_level)"""
        from oracle.ref import _low_level
        self.lib.adp_low_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        self.lib.adp_low_level.restype = C.c_longlong
        try:
            return _low_level(self.lib.adp_low_level, self.h, op, image, prefill)
        except RuntimeError as ex:
            raise RuntimeError("%s: %s" % (ex, self.lib.adp_last_error().decode()))

    def update_view_variants(self, depth_f32, raw_depth_i16):
        """ITMViewBuilder_B200::UpdateView(rgb, float depth) and UpdateView(rgb, short depth, filter, imu) on fresh views:
        returns (device depth after the float variant, device depth after the IMU variant)"""
        self.lib.adp_update_view_variants.argtypes = [C.c_void_p] * 5
        d = np.ascontiguousarray(depth_f32, np.float32)
        r = np.ascontiguousarray(raw_depth_i16, np.int16)
        o1, o2 = np.zeros_like(d), np.zeros_like(d)
        rc = self.lib.adp_update_view_variants(self.h, d.ctypes.data, o1.ctypes.data, r.ctypes.data, o2.ctypes.data)
        if rc != 0:
            raise RuntimeError("adp_update_view_variants: rc %d %s" % (rc, self.lib.adp_last_error().decode()))
        return o1, o2

    def save_scene_to_mesh(self, path):
        self.lib.adp_save_scene_to_mesh.argtypes = [C.c_void_p, C.c_char_p]
        n = self.lib.adp_save_scene_to_mesh(self.h, str(path).encode())
        if n < 0:
            raise RuntimeError("adp_save_scene_to_mesh: %s" % self.lib.adp_last_error().decode())
        return n

    @property
    def requires_full_rendering(self):
        return bool(self.lib.adp_requires_full_rendering(self.h))

    def get_free_image(self, render_type, pose_M, intr, w, h):
        """FindVisibleBlocks + CreateExpectedDepths + RenderImage on a free-view render state (ITMMainEngine.cpp:167-186)"""
        M = np.ascontiguousarray(pose_M, np.float32).reshape(16)
        k = np.ascontiguousarray(intr, np.float32).reshape(4)
        out = np.zeros((h, w, 4), np.uint8)
        if self.lib.adp_get_free_image(self.h, int(render_type), M.ctypes.data, k.ctypes.data, w, h, out.ctypes.data) != 0:
            raise RuntimeError("adp_get_free_image: %s" % self.lib.adp_last_error().decode())
        return out

    @property
    def pose_M(self):
        m = np.zeros(16, np.float32)
        self.lib.adp_get_pose(self.h, m.ctypes.data)
        return m

    @property
    def counters(self):
        c = np.zeros(4, np.int32)
        self.lib.adp_counters(self.h, c.ctypes.data)
        return c

    def read(self, which):
        need = -self.lib.adp_read(self.h, which, None, 0)
        out = np.empty(need // np.dtype(_DTYPES[which]).itemsize, dtype=_DTYPES[which])
        got = self.lib.adp_read(self.h, which, out.ctypes.data, need)
        if got != need:
            raise RuntimeError("adp_read(%d) failed" % which)
        return out
