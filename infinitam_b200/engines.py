"""Host-side mirror of the reference's engine interface for the fusion hot path.

Class and method names follow ITMLib so that parity tests read like the reference's own call
sites (ITMLib/Engine/ITMMainEngine.cpp:111-127):

    ITMMainEngine.ProcessFrame(rgb, rawDepth)
    ITMMainEngine.GetTrackingState() -> pose_d

Everything here forwards to the C ABI in libitm_b200.so; no arithmetic of the path is done in
Python and there is no fallback when the library or the GPU is missing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

HASH_ENTRY_DTYPE = np.dtype(
    {"names": ["pos", "offset", "ptr"], "formats": [("<i2", 3), "<i4", "<i4"], "offsets": [0, 8, 12], "itemsize": 16}
)

_BUF_DTYPES = {
    capi.BUF_VOXELS: np.uint32, capi.BUF_HASH: HASH_ENTRY_DTYPE, capi.BUF_VBA_ALLOC_LIST: np.int32,
    capi.BUF_EXCESS_ALLOC_LIST: np.int32, capi.BUF_VISIBLE_IDS: np.int32, capi.BUF_VISIBLE_TYPES: np.uint8,
    capi.BUF_DEPTH: np.float32, capi.BUF_MINMAX: np.float32, capi.BUF_RAYCAST_RESULT: np.float32,
    capi.BUF_RAYCAST_IMAGE: np.uint8, capi.BUF_POINTS: np.float32, capi.BUF_NORMALS: np.float32,
    capi.BUF_RAW_DEPTH: np.int16, capi.BUF_PYRAMID_1: np.float32, capi.BUF_PYRAMID_2: np.float32,
    capi.BUF_PYRAMID_3: np.float32, capi.BUF_PYRAMID_4: np.float32, capi.BUF_RGB: np.uint8, capi.BUF_SWAP_STATES: np.uint8,
    capi.BUF_FORWARD_PROJECTION: np.float32, capi.BUF_FWD_MISSING_POINTS: np.int32, capi.BUF_FREEVIEW_VISIBLE_IDS: np.int32,
    capi.BUF_FREEVIEW_MINMAX: np.float32, capi.BUF_FREEVIEW_RAYCAST_RESULT: np.float32, capi.BUF_FREEVIEW_IMAGE: np.uint8,
}


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class ITMMainEngine:
    """ITMLib/Engine/ITMMainEngine.h:50-130 - owns scene, tracking state, render state and view (all in HBM)."""

    def __init__(self, params: capi.Params | None = None, width: int = 640, height: int = 480):
        self.lib = capi.load()
        self.params = params if params is not None else capi.default_params(width, height)
        self.W, self.H = self.params.width, self.params.height
        h = C.c_void_p()
        capi.check(self.lib.itm_b200_engine_create(C.byref(self.params), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.itm_b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ITMMainEngine API ---------------------------------------------------------------
    def ProcessFrame(self, rgbImage, rawDepthImage):
        """ITMMainEngine::ProcessFrame (ITMMainEngine.cpp:111).  rgbImage may be None; rawDepthImage is an
        int16 host array (or a pinned torch tensor / raw address).  Returns pose_d->GetM() (column-major 16)."""
        pose = np.zeros(16, dtype=np.float32)
        capi.check(self.lib.itm_b200_engine_process_frame(self.h, _addr(rgbImage), _addr(rawDepthImage), _f32p(pose)))
        return pose

    def ProcessFrameWithPose(self, rgbImage, rawDepthImage, pose_M):
        """The fork's TRACKER_EXTERNAL mode: trackingState->pose_d is set from outside (RosPoseSourceEngine.cpp:112-118,
        pose_d->SetM semantics) and the frame is fused without ICP (ITMExternalTracker.cpp:27-30).  pose_M: column-major 16
        floats or None (keep the current pose)."""
        pose = np.zeros(16, dtype=np.float32)
        m = None if pose_M is None else np.ascontiguousarray(pose_M, np.float32).reshape(16)
        capi.check(self.lib.itm_b200_engine_process_frame_with_pose(self.h, _addr(rgbImage), _addr(rawDepthImage),
                                                                    None if m is None else _f32p(m), _f32p(pose)))
        return pose

    def SubmitFrame(self, rgbImage, rawDepthImage, pose_M=None) -> int:
        """Streaming ProcessFrame, first half: upload (copy stream) + frame enqueued; returns the frame's ticket.  The host
        images must be pinned and stay untouched until WaitFrame(ticket) returned."""
        t = C.c_ulonglong()
        m = None if pose_M is None else np.ascontiguousarray(pose_M, np.float32).reshape(16)
        capi.check(self.lib.itm_b200_engine_submit_frame(self.h, _addr(rgbImage), _addr(rawDepthImage),
                                                         None if m is None else _f32p(m), C.byref(t)))
        return t.value

    def WaitFrame(self, ticket: int):
        """second half: (pose_d->GetM(), counters) of that frame, polled from host-mapped memory"""
        pose = np.zeros(16, dtype=np.float32)
        counters = np.zeros(6, dtype=np.int32)
        capi.check(self.lib.itm_b200_engine_wait_frame(self.h, C.c_ulonglong(ticket), _f32p(pose), _i32p(counters)))
        return pose, counters

    def GetImage(self, getImageType: int, pose=None, intrinsics=None, width: int | None = None, height: int | None = None):
        """ITMMainEngine::GetImage (ITMMainEngine.cpp:134-192): returns the (h, w, 4) uint8 image.  pose (column-major
        16 floats, pose->GetM()) and intrinsics (fx, fy, cx, cy) are needed by the FREECAMERA types only."""
        w, h = width or self.W, height or self.H
        out = np.zeros((h, w, 4), dtype=np.uint8)
        p = None if pose is None else np.ascontiguousarray(pose, np.float32).reshape(16)
        k = None if intrinsics is None else np.ascontiguousarray(intrinsics, np.float32).reshape(4)
        capi.check(self.lib.itm_b200_engine_get_image(self.h, int(getImageType), None if p is None else _f32p(p),
                                                      None if k is None else _f32p(k), out.ctypes.data, w, h))
        return out

    def CreatePointCloud(self, trafo_rgb_to_depth=None, intrinsics_rgb=None, skipPoints: bool = False, with_image: bool = False):
        """The TRACKER_COLOR branch of ITMTrackingController::Prepare (ITMTrackingController.cpp:22-28) as a query:
        CreateExpectedDepths at the colour camera's pose + IITMVisualisationEngine::CreatePointCloud.  Returns
        (locations[n, 4], colours[n, 4]) (+ the (h, w, 4) shaded raycast with with_image); the live maps stay untouched."""
        T = None if trafo_rgb_to_depth is None else np.ascontiguousarray(trafo_rgb_to_depth, np.float32).reshape(16)
        k = None if intrinsics_rgb is None else np.ascontiguousarray(intrinsics_rgb, np.float32).reshape(4)
        n = C.c_int()
        cap = self.W * self.H
        loc = np.zeros((cap, 4), dtype=np.float32)
        clr = np.zeros((cap, 4), dtype=np.float32)
        img = np.zeros((self.H, self.W, 4), dtype=np.uint8) if with_image else None
        capi.check(self.lib.itm_b200_engine_create_point_cloud(self.h, None if T is None else _f32p(T), None if k is None else _f32p(k),
                                                               1 if skipPoints else 0, loc.ctypes.data, clr.ctypes.data, cap,
                                                               None if img is None else img.ctypes.data, C.byref(n)))
        out = (loc[: n.value].copy(), clr[: n.value].copy())
        return out + (img,) if with_image else out

    def UpdateMesh(self):
        """ITMMainEngine::UpdateMesh (ITMMainEngine.cpp:97-101): marching cubes over the scene; returns the (n, 9) float32
        triangle array (p0, p1, p2 per row) in the reference's order"""
        n = C.c_uint()
        capi.check(self.lib.itm_b200_engine_mesh_scene(self.h, None, 0, C.byref(n)))
        tri = np.zeros((n.value, 9), dtype=np.float32)
        if n.value:
            capi.check(self.lib.itm_b200_engine_mesh_scene(self.h, tri.ctypes.data, n.value, C.byref(n)))
        return tri

    def SaveSceneToMesh(self, objFileName: str):
        """ITMMainEngine::SaveSceneToMesh (ITMMainEngine.cpp:103-109): MeshScene + ITMMesh::WriteSTL"""
        capi.check(self.lib.itm_b200_engine_save_scene_to_mesh(self.h, str(objFileName).encode()))

    def EnqueueFrameDevice(self, raw_depth_dev_ptr: int):
        capi.check(self.lib.itm_b200_engine_enqueue_frame_dev(self.h, C.c_void_p(raw_depth_dev_ptr)))

    def PlaceDepthDevice(self, raw_depth_dev_ptr: int):
        """asynchronous D2D copy of a raw depth frame into the engine's own buffer; returns that buffer's device address (pass
        it to EnqueueFrameDevice: the frame then starts without a copy)"""
        capi.check(self.lib.itm_b200_engine_copy_to_buffer_dev(self.h, capi.BUF_RAW_DEPTH, C.c_void_p(raw_depth_dev_ptr), self.W * self.H * 2))
        return self.buffer_info(capi.BUF_RAW_DEPTH)[0]

    def stream(self) -> int:
        """the cudaStream_t (as an integer) the engine enqueues its frames on, e.g. for torch.cuda.ExternalStream"""
        p = C.c_void_p()
        capi.check(self.lib.itm_b200_engine_get_stream(self.h, C.byref(p)))
        return p.value or 0

    def Sync(self):
        pose = np.zeros(16, dtype=np.float32)
        counters = np.zeros(6, dtype=np.int32)
        capi.check(self.lib.itm_b200_engine_sync(self.h, _f32p(pose), _i32p(counters)))
        return pose, counters

    def ResetScene(self):
        capi.check(self.lib.itm_b200_engine_reset(self.h))

    # ---- stage-level access for teacher-forced parity tests ----------------------------------
    def UploadDepth(self, rawDepthImage):
        capi.check(self.lib.itm_b200_engine_upload_depth(self.h, _addr(rawDepthImage)))

    def RunStage(self, stage: int):
        capi.check(self.lib.itm_b200_engine_run_stage(self.h, stage))

    def buffer_info(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        capi.check(self.lib.itm_b200_engine_get_buffer(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def read(self, which, count=None):
        _, nbytes = self.buffer_info(which)
        dt = np.dtype(_BUF_DTYPES[which])
        if which == capi.BUF_VOXELS and self.params.voxel_type == capi.VOXEL_S_RGB:
            dt = np.dtype(np.uint64)  # one word per ITMVoxel_s_rgb
        n = nbytes // dt.itemsize if count is None else count
        out = np.empty(n, dtype=dt)
        capi.check(self.lib.itm_b200_engine_read_buffer(self.h, which, out.ctypes.data, n * dt.itemsize, 0))
        return out

    def write(self, which, arr):
        a = np.ascontiguousarray(arr)
        capi.check(self.lib.itm_b200_engine_write_buffer(self.h, which, a.ctypes.data, a.nbytes, 0))

    def read_image(self, which, channels=1):
        a = self.read(which)
        return a.reshape(self.H, self.W, channels) if channels > 1 else a.reshape(self.H, self.W)

    def get_state(self):
        pose, pc, st = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(6, np.int32)
        capi.check(self.lib.itm_b200_engine_get_state(self.h, _f32p(pose), _f32p(pc), _i32p(st)))
        return pose, pc, st

    def set_state(self, pose_d=None, pose_point_cloud=None, state6=None):
        p = None if pose_d is None else np.ascontiguousarray(pose_d, np.float32).reshape(16)
        q = None if pose_point_cloud is None else np.ascontiguousarray(pose_point_cloud, np.float32).reshape(16)
        s = None if state6 is None else np.ascontiguousarray(state6, np.int32).reshape(6)
        capi.check(self.lib.itm_b200_engine_set_state(
            self.h, None if p is None else _f32p(p), None if q is None else _f32p(q), None if s is None else _i32p(s)))

    def swap_counts(self):
        """(swapped_in, swapped_out) of the last frame of a swapping engine"""
        n_in, n_out = C.c_int(0), C.c_int(0)
        capi.check(self.lib.itm_b200_engine_global_cache(self.h, None, None, C.byref(n_in), C.byref(n_out)))
        return n_in.value, n_out.value

    def global_cache(self):
        """(hasStoredData uint8[entries], storedVoxelBlocks words[entries, 512], swapped_in, swapped_out) of a swapping engine -
        numpy views onto the engine's host memory"""
        has, blocks = C.c_void_p(), C.c_void_p()
        n_in, n_out = C.c_int(), C.c_int()
        capi.check(self.lib.itm_b200_engine_global_cache(self.h, C.byref(has), C.byref(blocks), C.byref(n_in), C.byref(n_out)))
        n = self.params.sdf_bucket_num + self.params.sdf_excess_list_size
        wdt = np.uint64 if self.params.voxel_type == capi.VOXEL_S_RGB else np.uint32
        h = np.frombuffer((C.c_char * n).from_address(has.value), dtype=np.uint8)
        b = np.frombuffer((C.c_char * (n * 512 * np.dtype(wdt).itemsize)).from_address(blocks.value), dtype=wdt).reshape(n, 512)
        return h, b, n_in.value, n_out.value

    def icp_stats(self):
        """ComputeGandH evaluations per pyramid level in the last synced frame (level 0 = full resolution)"""
        n = np.zeros(capi.MAX_LEVELS, np.int32)
        capi.check(self.lib.itm_b200_engine_icp_stats(self.h, _i32p(n)))
        return n

    def set_profiling(self, on=True):
        """True / 1: a time stamp at every stage boundary; 2: frame start and end only (no event nodes between the kernels)"""
        capi.check(self.lib.itm_b200_engine_set_profiling(self.h, int(on)))

    def shard_times(self):
        """sharded engine, profiling 1: (partial ray cast, barrier wait, composition) of the last frame in ms"""
        ms = np.zeros(3, np.float32)
        capi.check(self.lib.itm_b200_engine_shard_times(self.h, _f32p(ms)))
        return ms

    def shard_unresolved(self):
        """sharded engine: pixels of the last composed raycast that no rank could march completely (reported as misses)"""
        n = C.c_int(0)
        capi.check(self.lib.itm_b200_engine_shard_unresolved(self.h, C.byref(n)))
        return n.value

    def stage_times(self):
        ms = np.zeros(8, np.float32)
        capi.check(self.lib.itm_b200_engine_stage_times(self.h, _f32p(ms)))
        return ms


def _addr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch tensor (pinned host memory)
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))
