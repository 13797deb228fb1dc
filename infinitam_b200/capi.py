"""ctypes binding of libitm_b200.so (include/itm_b200.h).

The CUDA library is the product; this module only loads it.  There is no fallback of any
kind: if the shared library is missing or no CUDA device is present, creating an engine
raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ITM_B200_LIB: an experimental variant built by ITM_B200_VARIANT=... python -m infinitam_b200.build (A/B measurements only)
LIB_PATH = os.environ.get("ITM_B200_LIB") or os.path.join(HERE, "libitm_b200.so")

MAX_LEVELS = 8
OK, EINVAL, ECUDA, ENODEVICE, EUNSUPPORTED = 0, -1, -2, -3, -4
ITER_ROTATION, ITER_TRANSLATION, ITER_BOTH, ITER_NONE = 1, 2, 3, 4

(BUF_VOXELS, BUF_HASH, BUF_VBA_ALLOC_LIST, BUF_EXCESS_ALLOC_LIST, BUF_VISIBLE_IDS, BUF_VISIBLE_TYPES, BUF_DEPTH,
 BUF_MINMAX, BUF_RAYCAST_RESULT, BUF_RAYCAST_IMAGE, BUF_POINTS, BUF_NORMALS, BUF_RAW_DEPTH, BUF_PYRAMID_1,
 BUF_PYRAMID_2, BUF_PYRAMID_3, BUF_PYRAMID_4, BUF_RGB, BUF_SWAP_STATES, BUF_FORWARD_PROJECTION, BUF_FWD_MISSING_POINTS,
 BUF_FREEVIEW_VISIBLE_IDS, BUF_FREEVIEW_MINMAX, BUF_FREEVIEW_RAYCAST_RESULT, BUF_FREEVIEW_IMAGE, BUF_COUNT) = range(26)
VOXEL_S, VOXEL_S_RGB = 0, 1
TRACKER_ICP, TRACKER_EXTERNAL, TRACKER_WICP = 0, 1, 2
DEPTH_AFFINE, DEPTH_KINECT_DISPARITY = 0, 1
MAX_IN_FLIGHT = 4

(STAGE_VIEW, STAGE_TRACK, STAGE_ALLOCATE, STAGE_INTEGRATE, STAGE_EXPECTED_DEPTHS, STAGE_ICP_MAPS, STAGE_SWAP,
 STAGE_FORWARD_RENDER, STAGE_TRACK_DECIDE) = range(9)
# ITMMainEngine::GetImageType (Engine/ITMMainEngine.h:78-87), IITMVisualisationEngine::RenderImageType
(IMAGE_ORIGINAL_RGB, IMAGE_ORIGINAL_DEPTH, IMAGE_SCENERAYCAST, IMAGE_FREECAMERA_SHADED, IMAGE_FREECAMERA_COLOUR_FROM_VOLUME,
 IMAGE_FREECAMERA_COLOUR_FROM_NORMAL, IMAGE_UNKNOWN) = range(7)
RENDER_SHADED_GREYSCALE, RENDER_COLOUR_FROM_VOLUME, RENDER_COLOUR_FROM_NORMAL = range(3)


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("voxel_size", C.c_float), ("mu", C.c_float), ("max_w", C.c_int),
        ("view_frustum_min", C.c_float), ("view_frustum_max", C.c_float),
        ("stop_integrating_at_max_w", C.c_int),
        ("depth_calib_a", C.c_float), ("depth_calib_b", C.c_float),
        ("sdf_local_block_num", C.c_int), ("sdf_bucket_num", C.c_int), ("sdf_excess_list_size", C.c_int),
        ("no_hierarchy_levels", C.c_int), ("tracking_regime", C.c_int * MAX_LEVELS),
        ("no_icp_run_till_level", C.c_int),
        ("depth_tracker_icp_threshold", C.c_float), ("depth_tracker_termination_threshold", C.c_float),
        ("device", C.c_int),
        ("voxel_type", C.c_int),
        ("rgb_fx", C.c_float), ("rgb_fy", C.c_float), ("rgb_cx", C.c_float), ("rgb_cy", C.c_float),
        ("trafo_rgb_to_depth_inv", C.c_float * 16),
        ("use_swapping", C.c_int),
        ("use_approximate_raycast", C.c_int),
        ("icp_max_ctas", C.c_int),
        ("tracker_type", C.c_int),
        ("depth_source", C.c_int),
        ("use_bilateral_filter", C.c_int),
        ("swap_cache_blocks", C.c_int),
    ]


class Scene(C.Structure):
    _fields_ = [
        ("voxel_blocks_dev", C.c_void_p), ("hash_entries_dev", C.c_void_p),
        ("vba_allocation_list_dev", C.c_void_p), ("excess_allocation_list_dev", C.c_void_p),
        ("last_free_block_id", C.c_int), ("last_free_excess_list_id", C.c_int),
        ("swap_states_dev", C.c_void_p),
    ]


class SwapBuffers(C.Structure):
    _fields_ = [("needed_entry_ids_dev", C.c_void_p), ("synced_voxel_blocks_dev", C.c_void_p), ("has_synced_data_dev", C.c_void_p)]


class RenderState(C.Structure):
    _fields_ = [
        ("visible_entry_ids_dev", C.c_void_p), ("entries_visible_type_dev", C.c_void_p),
        ("no_visible_entries", C.c_int),
        ("rendering_range_image_dev", C.c_void_p), ("raycast_result_dev", C.c_void_p),
        ("raycast_image_dev", C.c_void_p),
        ("forward_projection_dev", C.c_void_p), ("fwd_proj_missing_points_dev", C.c_void_p),
        ("no_fwd_proj_missing_points", C.c_int), ("img_width", C.c_int), ("img_height", C.c_int),
    ]


class TrackingState(C.Structure):
    _fields_ = [
        ("points_map_dev", C.c_void_p), ("normals_map_dev", C.c_void_p),
        ("pose_d", C.c_float * 16), ("pose_point_cloud", C.c_float * 16),
        ("age_point_cloud", C.c_int),
    ]


MAX_SHARDS = 8
IPC_HANDLE_BYTES = 64


class Shard(C.Structure):
    _fields_ = [
        ("rank", C.c_int), ("world", C.c_int),
        ("axis", C.c_int), ("origin_block", C.c_int), ("thickness_blocks", C.c_int),
        ("partial_raycast_dev", (C.c_void_p * MAX_SHARDS) * 2), ("tile_hit_dev", (C.c_void_p * MAX_SHARDS) * 2),
        ("barrier_flags_dev", C.c_void_p * MAX_SHARDS), ("stream", C.c_void_p),
        ("halo_blocks", C.c_int),
    ]


# every symbol include/itm_b200.h declares
SYMBOLS = [
    "itm_b200_default_params", "itm_b200_last_error", "itm_b200_device_count", "itm_b200_launch_count",
    "itm_b200_ctx_create", "itm_b200_ctx_destroy", "itm_b200_reset_scene", "itm_b200_allocate_scene_from_depth",
    "itm_b200_integrate_into_scene", "itm_b200_integrate_into_scene_rgb", "itm_b200_create_expected_depths", "itm_b200_create_icp_maps",
    "itm_b200_convert_depth_affine_to_float", "itm_b200_filter_subsample_with_holes", "itm_b200_compute_g_and_h",
    "itm_b200_track_camera", "itm_b200_engine_create", "itm_b200_engine_destroy", "itm_b200_engine_reset",
    "itm_b200_engine_create_sharded", "itm_b200_ipc_alloc", "itm_b200_ipc_open", "itm_b200_ipc_close", "itm_b200_ipc_free",
    "itm_b200_shard_owner_of_block", "itm_b200_engine_process_frame", "itm_b200_engine_enqueue_frame_dev", "itm_b200_engine_sync",
    "itm_b200_engine_upload_depth", "itm_b200_engine_run_stage", "itm_b200_engine_get_buffer",
    "itm_b200_engine_read_buffer", "itm_b200_engine_write_buffer", "itm_b200_engine_global_cache", "itm_b200_engine_get_state",
    "itm_b200_engine_set_state", "itm_b200_engine_icp_stats", "itm_b200_engine_set_profiling", "itm_b200_engine_stage_times",
    "itm_b200_mat4_inv", "itm_b200_pose_from_inv_m_coerced", "itm_b200_compute_delta",
    "itm_b200_forward_render", "itm_b200_find_visible_blocks", "itm_b200_find_surface", "itm_b200_render_image",
    "itm_b200_engine_get_image", "itm_b200_create_point_cloud", "itm_b200_engine_create_point_cloud", "itm_b200_copy_image",
    "itm_b200_filter_subsample_rgba", "itm_b200_filter_subsample_with_holes_float4", "itm_b200_gradient_x", "itm_b200_gradient_y", "itm_b200_mesh_scene", "itm_b200_write_stl", "itm_b200_write_obj",
    "itm_b200_engine_mesh_scene", "itm_b200_engine_save_scene_to_mesh",
    "itm_b200_take_cuda_error", "itm_b200_engine_get_stream", "itm_b200_engine_copy_to_buffer_dev", "itm_b200_compute_g_and_h_weighted", "itm_b200_depth_filtering", "itm_b200_compute_normal_and_weights", "itm_b200_swap_in_select", "itm_b200_swap_in_apply", "itm_b200_swap_out",
    "itm_b200_convert_disparity_to_depth", "itm_b200_engine_process_frame_with_pose", "itm_b200_engine_submit_frame",
    "itm_b200_engine_wait_frame", "itm_b200_shard_block_resident", "itm_b200_engine_shard_times", "itm_b200_track_camera_weighted",
    "itm_b200_set_alloc_mode", "itm_b200_engine_shard_unresolved", "itm_b200_shard_block_resident_halo",
    "itm_b200_engine_shard_export", "itm_b200_engine_shard_attach",
]

_lib = None


class ItmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("itm_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Loads libitm_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ItmError(ENODEVICE, "libitm_b200.so not built (python -m infinitam_b200.build); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    f32p, i32p, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
    lib.itm_b200_default_params.argtypes = [C.POINTER(Params), C.c_int, C.c_int]
    lib.itm_b200_default_params.restype = None
    lib.itm_b200_last_error.restype = C.c_char_p
    lib.itm_b200_device_count.restype = C.c_int
    lib.itm_b200_launch_count.restype = C.c_ulonglong
    lib.itm_b200_ctx_create.argtypes = [C.POINTER(Params), vp, C.POINTER(vp)]
    lib.itm_b200_ctx_destroy.argtypes = [vp]
    lib.itm_b200_ctx_destroy.restype = None
    lib.itm_b200_reset_scene.argtypes = [vp, C.POINTER(Scene)]
    lib.itm_b200_allocate_scene_from_depth.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), vp, f32p, C.c_int]
    lib.itm_b200_integrate_into_scene.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), vp, f32p]
    lib.itm_b200_integrate_into_scene_rgb.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), vp, vp, f32p]
    lib.itm_b200_create_expected_depths.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), f32p, f32p]
    lib.itm_b200_create_icp_maps.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), C.POINTER(TrackingState)]
    lib.itm_b200_forward_render.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), vp, C.POINTER(TrackingState)]
    lib.itm_b200_find_visible_blocks.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), f32p, f32p]
    lib.itm_b200_find_surface.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), f32p, f32p]
    lib.itm_b200_render_image.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), f32p, f32p, vp, C.c_int]
    lib.itm_b200_engine_get_image.argtypes = [vp, C.c_int, f32p, f32p, vp, C.c_int, C.c_int]
    lib.itm_b200_copy_image.argtypes = [vp, vp, vp, C.c_size_t]
    for name in ("itm_b200_filter_subsample_rgba", "itm_b200_filter_subsample_with_holes_float4", "itm_b200_gradient_x", "itm_b200_gradient_y"):
        getattr(lib, name).argtypes = [vp, vp, vp, C.c_int, C.c_int]
    lib.itm_b200_create_point_cloud.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), C.POINTER(TrackingState), f32p, f32p,
                                                C.c_int, C.POINTER(C.c_int)]
    lib.itm_b200_engine_create_point_cloud.argtypes = [vp, f32p, f32p, C.c_int, vp, vp, C.c_int, vp, C.POINTER(C.c_int)]
    lib.itm_b200_swap_in_select.argtypes = [vp, C.POINTER(Scene), C.POINTER(SwapBuffers), i32p]
    lib.itm_b200_swap_in_apply.argtypes = [vp, C.POINTER(Scene), C.POINTER(SwapBuffers), C.c_int]
    lib.itm_b200_swap_out.argtypes = [vp, C.POINTER(Scene), C.POINTER(RenderState), C.POINTER(SwapBuffers), i32p]
    lib.itm_b200_mesh_scene.argtypes = [vp, C.POINTER(Scene), vp, C.c_uint, C.POINTER(C.c_uint)]
    lib.itm_b200_write_stl.argtypes = [C.c_char_p, vp, C.c_uint]
    lib.itm_b200_write_obj.argtypes = [C.c_char_p, vp, C.c_uint]
    lib.itm_b200_engine_mesh_scene.argtypes = [vp, vp, C.c_uint, C.POINTER(C.c_uint)]
    lib.itm_b200_engine_save_scene_to_mesh.argtypes = [vp, C.c_char_p]
    lib.itm_b200_convert_depth_affine_to_float.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float]
    lib.itm_b200_filter_subsample_with_holes.argtypes = [vp, vp, vp, C.c_int, C.c_int]
    lib.itm_b200_compute_g_and_h.argtypes = [vp, vp, C.c_int, C.c_int, f32p, vp, vp, C.c_int, C.c_int, f32p, f32p, f32p,
                                             C.c_float, C.c_int, f32p, f32p, f32p, i32p]
    lib.itm_b200_compute_g_and_h_weighted.argtypes = [vp, vp, vp, C.c_int, C.c_int, f32p, vp, vp, C.c_int, C.c_int, f32p, f32p, f32p,
                                                      C.c_float, C.c_int, f32p, f32p, f32p, i32p]
    lib.itm_b200_depth_filtering.argtypes = [vp, vp, vp, C.c_int, C.c_int]
    lib.itm_b200_compute_normal_and_weights.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, f32p]
    lib.itm_b200_track_camera.argtypes = [vp, vp, C.POINTER(TrackingState)]
    lib.itm_b200_track_camera_weighted.argtypes = [vp, vp, vp, C.POINTER(TrackingState)]
    lib.itm_b200_engine_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.itm_b200_engine_create_sharded.argtypes = [C.POINTER(Params), C.POINTER(Shard), C.POINTER(vp)]
    lib.itm_b200_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(vp), C.c_char_p]
    lib.itm_b200_ipc_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.itm_b200_ipc_close.argtypes = [vp]
    lib.itm_b200_ipc_free.argtypes = [vp]
    lib.itm_b200_shard_owner_of_block.argtypes = [C.c_int] * 7
    lib.itm_b200_shard_block_resident.argtypes = [C.c_int] * 8
    lib.itm_b200_engine_destroy.argtypes = [vp]
    lib.itm_b200_engine_destroy.restype = None
    lib.itm_b200_engine_reset.argtypes = [vp]
    lib.itm_b200_engine_process_frame.argtypes = [vp, vp, vp, f32p]
    lib.itm_b200_engine_process_frame_with_pose.argtypes = [vp, vp, vp, f32p, f32p]
    lib.itm_b200_engine_submit_frame.argtypes = [vp, vp, vp, f32p, C.POINTER(C.c_ulonglong)]
    lib.itm_b200_engine_wait_frame.argtypes = [vp, C.c_ulonglong, f32p, i32p]
    lib.itm_b200_convert_disparity_to_depth.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    lib.itm_b200_engine_enqueue_frame_dev.argtypes = [vp, vp]
    lib.itm_b200_engine_get_stream.argtypes = [vp, C.POINTER(vp)]
    lib.itm_b200_engine_copy_to_buffer_dev.argtypes = [vp, C.c_int, vp, C.c_size_t]
    lib.itm_b200_engine_sync.argtypes = [vp, f32p, i32p]
    lib.itm_b200_engine_upload_depth.argtypes = [vp, vp]
    lib.itm_b200_engine_run_stage.argtypes = [vp, C.c_int]
    lib.itm_b200_engine_get_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.itm_b200_engine_read_buffer.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_size_t]
    lib.itm_b200_engine_write_buffer.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_size_t]
    lib.itm_b200_engine_global_cache.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), i32p, i32p]
    lib.itm_b200_engine_get_state.argtypes = [vp, f32p, f32p, i32p]
    lib.itm_b200_engine_set_state.argtypes = [vp, f32p, f32p, i32p]
    lib.itm_b200_engine_icp_stats.argtypes = [vp, i32p]
    lib.itm_b200_engine_set_profiling.argtypes = [vp, C.c_int]
    lib.itm_b200_engine_stage_times.argtypes = [vp, f32p]
    lib.itm_b200_engine_shard_times.argtypes = [vp, f32p]
    lib.itm_b200_mat4_inv.argtypes = [f32p, f32p]
    lib.itm_b200_pose_from_inv_m_coerced.argtypes = [f32p, f32p, f32p, f32p]
    lib.itm_b200_compute_delta.argtypes = [f32p, f32p, C.c_int, f32p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise ItmError(rc, load().itm_b200_last_error().decode(errors="replace"))


def default_params(width=640, height=480) -> Params:
    p = Params()
    load().itm_b200_default_params(C.byref(p), width, height)
    return p


def set_alloc_mode(mode):
    """itm_b200_set_alloc_mode: 0 automatic, 1 ordered scans, 2 compact lists; returns the previous mode."""
    lib = load()
    lib.itm_b200_set_alloc_mode.argtypes = [C.c_int]
    lib.itm_b200_set_alloc_mode.restype = C.c_int
    prev = lib.itm_b200_set_alloc_mode(int(mode))
    if prev < 0:
        raise ItmError(prev, lib.itm_b200_last_error().decode())
    return prev
