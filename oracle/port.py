"""ctypes view of oracle/libitm_oracle.so (the plain-C restatement, oracle/itm_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
Same Python interface as oracle.ref.RefEngine, plus run-time pool sizes.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build_port
from .ref import RefEngine, _Prefixed, declare_common

MAX_LEVELS = 8


class PortParams(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("voxel_size", C.c_float), ("mu", C.c_float), ("max_w", C.c_int),
        ("vf_min", C.c_float), ("vf_max", C.c_float), ("stop_at_max_w", C.c_int),
        ("calib_a", C.c_float), ("calib_b", C.c_float),
        ("n_local", C.c_int), ("n_bucket", C.c_int), ("n_excess", C.c_int),
        ("n_levels", C.c_int), ("regime", C.c_int * MAX_LEVELS), ("no_icp_run_till_level", C.c_int),
        ("icp_dist_thresh", C.c_float), ("icp_termination", C.c_float),
    ]


_libs = {}


def load(fast=False):
    if fast in _libs:
        return _libs[fast]
    path = build_port.LIB_FAST if fast else build_port.LIB
    if not os.path.exists(path):
        build_port.build()
    lib = _Prefixed(C.CDLL(path), "port")
    lib.port_default_params.argtypes = [C.POINTER(PortParams), C.c_int, C.c_int]
    lib.port_default_params.restype = None
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.POINTER(PortParams)]
    declare_common(lib)
    _libs[fast] = lib
    return lib


class PortEngine(RefEngine):
    def __init__(self, width=640, height=480, intr=None, voxel_size=0.005, mu=0.02, max_w=100, vf_min=0.35, vf_max=3.0,
                 n_local=0x10000, n_bucket=0x100000, n_excess=0x20000, fast=False, wicp=False, bilateral=False):
        self.n_local, self.n_bucket, self.n_excess = n_local, n_bucket, n_excess
        self._fast = fast
        super().__init__(width, height, intr, voxel_size, mu, max_w, vf_min, vf_max, wicp=wicp, bilateral=bilateral)

    def _create(self, flavour):
        self.lib = load(self._fast)
        p = PortParams()
        self.lib.port_default_params(C.byref(p), self.W, self.H)
        p.fx, p.fy, p.cx, p.cy = self.intr
        p.voxel_size, p.mu, p.max_w, p.vf_min, p.vf_max = self.voxel_size, self.mu, self.max_w, self.vf_min, self.vf_max
        p.n_local, p.n_bucket, p.n_excess = self.n_local, self.n_bucket, self.n_excess
        self.params = p
        self.lib.ref_set_tracker_wicp(int(self.wicp), int(self.bilateral))
        self.h = self.lib.ref_create(C.byref(p))
        self.lib.ref_set_tracker_wicp(0, 0)

    def const(self, name):
        return {"SDF_LOCAL_BLOCK_NUM": self.n_local, "SDF_BUCKET_NUM": self.n_bucket, "SDF_EXCESS_LIST_SIZE": self.n_excess,
                "sizeof_voxel": 4, "sizeof_hash_entry": 16, "has_color": 0, "openmp": 0}.get(name, -1)
