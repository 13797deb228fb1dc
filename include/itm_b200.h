/* itm_b200.h - C ABI of the B200-native InfiniTAM dense-fusion engines (libitm_b200.so).
 *
 * This is the drop-in boundary.  The reference selects its engines with
 * `switch (settings->deviceType)` in three places (ITMLib/Engine/ITMMainEngine.cpp:21-45,
 * ITMLib/Engine/ITMDenseMapper.cpp:16-34, ITMLib/Engine/ITMTrackerFactory.h:182-235); each
 * function below is what one virtual method of those engine interfaces forwards to - the
 * adapter classes that do the forwarding are in include/itm_b200_adapter.hpp and the wiring is
 * shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no CUDA or torch types in any signature
 *     (a cudaStream_t is passed as void*).
 *   - every pointer named *_dev is DEVICE memory owned by the caller (in the reference: by
 *     ORUtils::MemoryBlock<T>, ORUtils/MemoryBlock.h:173-263); the library only borrows it for
 *     the duration of the call.  The handle owns scratch memory only.
 *   - matrices are the reference's column-major float[16] (ORUtils/Matrix.h:8-33).
 *   - layouts are the reference's: ITMHashEntry 16 B, ITMVoxel_s 4 B, Vector4f / Vector2f /
 *     Vector4u images (ITMLib/Utils/ITMLibDefines.h:71-82, 157-179).
 *   - every function returns 0 on success, a negative ITM_B200_E* code otherwise;
 *     itm_b200_last_error() gives the message (the adapter turns it into DIEWITHEXCEPTION,
 *     ORUtils/PlatformIndependence.h:34-38).  Pool exhaustion is not an error (the reference
 *     silently skips the block, ITMSceneReconstructionEngine_CPU.cpp:187-189); it is counted.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     ITM_B200_ENODEVICE.
 */
#ifndef ITM_B200_H
#define ITM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ITM_B200_OK 0
#define ITM_B200_EINVAL (-1)
#define ITM_B200_ECUDA (-2)
#define ITM_B200_ENODEVICE (-3)
#define ITM_B200_EUNSUPPORTED (-4)

#define ITM_B200_MAX_LEVELS 8

/* TrackerIterationType, ITMLib/Utils/ITMLibDefines.h:278-283 */
#define ITM_B200_ITER_ROTATION 1
#define ITM_B200_ITER_TRANSLATION 2
#define ITM_B200_ITER_BOTH 3
#define ITM_B200_ITER_NONE 4

/* What ITMLibSettings + ITMSceneParams + ITMRGBDCalib + the hash #defines carry
 * (ITMLib/Utils/ITMLibSettings.cpp:9-88, ITMLib/Objects/ITMSceneParams.h:14-70,
 *  ITMLib/Utils/ITMLibDefines.h:37-62).  Pool sizes are run-time here. */
typedef struct itm_b200_params {
  int width, height;                 /* depth image size */
  float fx, fy, cx, cy;              /* intrinsics_d.projectionParamsSimple.all */
  float voxel_size;                  /* sceneParams.voxelSize      (0.005) */
  float mu;                          /* sceneParams.mu             (0.02)  */
  int max_w;                         /* sceneParams.maxW           (100)   */
  float view_frustum_min;            /* sceneParams.viewFrustum_min (0.35) */
  float view_frustum_max;            /* sceneParams.viewFrustum_max (3.0)  */
  int stop_integrating_at_max_w;     /* sceneParams.stopIntegratingAtMaxW (0) */
  float depth_calib_a, depth_calib_b;/* ITMDisparityCalib TRAFO_AFFINE params (1/1000, 0) */
  int sdf_local_block_num;           /* SDF_LOCAL_BLOCK_NUM   (0x10000)  */
  int sdf_bucket_num;                /* SDF_BUCKET_NUM        (0x100000), power of two */
  int sdf_excess_list_size;          /* SDF_EXCESS_LIST_SIZE  (0x20000)  */
  int no_hierarchy_levels;           /* settings.noHierarchyLevels (5) */
  int tracking_regime[ITM_B200_MAX_LEVELS]; /* settings.trackingRegime, level 0 = full res */
  int no_icp_run_till_level;         /* settings.noICPRunTillLevel (0) */
  float depth_tracker_icp_threshold; /* settings.depthTrackerICPThreshold (0.1*0.1) */
  float depth_tracker_termination_threshold; /* (1e-3) */
  int device;                        /* CUDA device ordinal */
  /* voxel type (ITMLib/Utils/ITMLibDefines.h:205 picks it at compile time; here it is a run-time choice):
   * ITM_B200_VOXEL_S = ITMVoxel_s (4 B: short sdf, uchar w_depth), ITM_B200_VOXEL_S_RGB = ITMVoxel_s_rgb (8 B: + uchar
   * clr[3], uchar w_color; ITMLibDefines.h:127-155) with colour integration from view->rgb */
  int voxel_type;
  float rgb_fx, rgb_fy, rgb_cx, rgb_cy;      /* calib.intrinsics_rgb.projectionParamsSimple.all (rgb image = depth image size) */
  float trafo_rgb_to_depth_inv[16];  /* calib.trafo_rgb_to_depth.calib_inv, column-major (identity by default) */
  /* settings.useSwapping (ITMLibSettings.cpp:35, off by default): Layer B keeps an ITMGlobalCache in host memory and runs
   * ITMSwappingEngine::IntegrateGlobalIntoLocal / SaveToGlobalMemory after every integration (ITMDenseMapper.cpp:59-64) */
  int use_swapping;
  /* settings.useApproximateRaycast (ITMLibSettings.cpp:32, off by default): ITMTrackingController::Track/Prepare
   * (ITMTrackingController.cpp:11-46) skip the full raycast while the camera stays close to the pose of the last one
   * (ITMTrackingState::TrackerFarFromPointCloud, Objects/ITMTrackingState.h:41-59) and forward-project it instead */
  int use_approximate_raycast;
  /* 0: the tracker's persistent kernel takes one CTA on every SM (lowest latency for one scene).  n > 0: at most n CTAs,
   * so that the trackers of several scenes sharing the GPU (BASELINE configs[3]: 8 scenes per GPU, one engine and stream
   * each) run side by side instead of one after the other - most of a TrackCamera is spent on pyramid levels that occupy
   * fewer than 40 CTAs anyway.  Changes the summation order of the ICP sums (poses agree to ~1e-6, not bit for bit). */
  int icp_max_ctas;
  /* settings.trackerType (ITMLib/Utils/ITMLibSettings.h:40-54).  ITM_B200_TRACKER_ICP: ITMDepthTracker (the default here).
   * ITM_B200_TRACKER_EXTERNAL: the fork's default (ITMLibSettings.cpp:44) - ITMExternalTracker::TrackCamera is a no-op
   * (Engine/ITMExternalTracker.cpp:27-30) and the pose comes from outside (Engine/RosPoseSourceEngine.cpp:112-118 writes
   * trackingState->pose_d before every frame): use itm_b200_engine_process_frame_with_pose / _submit_frame with a pose.
   * ITM_B200_TRACKER_WICP: ITMWeightedICPTracker (Engine/ITMWeightedICPTracker.cpp:58-192) - bilateral-filtered depth,
   * sensor-noise weights, plain Gauss-Newton - inside the same device-side loop. */
  int tracker_type;
  /* calib.disparityCalib.type (ITMLib/Objects/ITMDisparityCalib.h:24-30): how raw shorts become metres in UpdateView
   * (ITMViewBuilder_CPU.cpp:38-46).  AFFINE: d * depth_calib_a + depth_calib_b.  KINECT_DISPARITY:
   * 8 * depth_calib_b * fx / (depth_calib_a - d)  (convertDisparityToDepth, DeviceAgnostic/ITMViewBuilder.h:7-20). */
  int depth_source;
  /* settings.useBilateralFilter (ITMLibSettings.cpp:41, off by default): UpdateView runs five passes of the 5x5 bilateral depth
   * filter (ITMViewBuilder_CPU.cpp:50-59).  settings.modelSensorNoise is implied by ITM_B200_TRACKER_WICP (ITMLibSettings.cpp:51-53). */
  int use_bilateral_filter;
  /* Layer B with use_swapping: the global cache (ITMGlobalCache, ITMLib/Objects/ITMGlobalCache.h:21-36 - room for every hash
   * entry's block in the reference, 2.4 / 4.8 GB) is a pool of this many voxel blocks in host-mapped pinned memory that the
   * swapping kernels read and write directly over PCIe, with a per-entry slot index on the device; a block gets its slot the
   * first time it is swapped out.  0 = 4 x sdf_local_block_num (at most one per hash entry).  A full pool sets error flag 4
   * (the block's data is then lost, the reference never runs out). */
  int swap_cache_blocks;
} itm_b200_params;
#define ITM_B200_VOXEL_S 0
#define ITM_B200_VOXEL_S_RGB 1
#define ITM_B200_TRACKER_ICP 0
#define ITM_B200_TRACKER_EXTERNAL 1
#define ITM_B200_TRACKER_WICP 2
#define ITM_B200_DEPTH_AFFINE 0
#define ITM_B200_DEPTH_KINECT_DISPARITY 1

/* Fills *p with the reference's defaults (ICP tracker regime) for a width x height sensor;
 * intrinsics default to ITMIntrinsics() = (580, 580, 320, 240) scaled by width/640. */
void itm_b200_default_params(itm_b200_params *p, int width, int height);

const char *itm_b200_last_error(void);
/* number of CUDA devices visible, or a negative error */
int itm_b200_device_count(void);
/* total kernels launched by this library in the calling process so far */
unsigned long long itm_b200_launch_count(void);
/* the CUDA runtime's pending (non-sticky) error code of the calling thread, cleared by this call; 0 = none.  No entry point
 * of the library leaves one behind (hosts such as PyTorch that share the CUDA runtime would trip over it). */
int itm_b200_take_cuda_error(void);
/* Which kernels AllocateSceneFromDepth uses (ITMSceneReconstructionEngine_CPU.cpp:117-291; both give the identical hash table,
 * free lists, entriesVisibleType and visibleEntryIDs): 0 = automatic (compact per-bin lists where a frame touches fewer hash
 * slots than the table has, else two ordered whole-table scans), 1 = always the scans, 2 = the lists wherever they apply
 * (not for swapping / sharded engines or onlyUpdateVisibleList).  Process-wide; takes effect for engines created afterwards.
 * Returns the previous mode, or a negative error.  For A/B measurements and tests. */
int itm_b200_set_alloc_mode(int mode);

/* ===================================================================================== *
 *  Layer A - one function per engine method, on caller-owned device buffers.            *
 *  All of these return after the work has completed (stream-synchronised), like the     *
 *  reference's own device engines do at their host-visible outputs.                     *
 * ===================================================================================== */

typedef struct itm_b200_ctx itm_b200_ctx;

/* stream: the cudaStream_t every call of this context runs on, or NULL for a private NON-BLOCKING stream.  A private stream
 * does not wait for work the caller still has in flight elsewhere - not even on the legacy default stream (an asynchronous
 * cudaMemset of an input buffer, a kernel that produces the depth image): synchronise before calling, or pass the stream
 * that work is on.  Code written against the legacy default stream, like ITMLib's host objects, passes cudaStreamLegacy
 * ((void *)0x1), which is what include/itm_b200_adapter.hpp does. */
int itm_b200_ctx_create(const itm_b200_params *params, void *stream, itm_b200_ctx **out);
void itm_b200_ctx_destroy(itm_b200_ctx *ctx);

/* ITMScene<ITMVoxel_s, ITMVoxelBlockHash>: index + localVBA (ITMLib/Objects/ITMScene.h:20-51) */
typedef struct itm_b200_scene {
  void *voxel_blocks_dev;            /* localVBA.GetVoxelBlocks():  ITMVoxel_s / _s_rgb [local*512] */
  void *hash_entries_dev;            /* index.GetEntries():         ITMHashEntry[bucket+excess] */
  int *vba_allocation_list_dev;      /* localVBA.GetAllocationList(): int[local]               */
  int *excess_allocation_list_dev;   /* index.GetExcessAllocationList(): int[excess]           */
  int last_free_block_id;            /* localVBA.lastFreeBlockId            (host, in/out)      */
  int last_free_excess_list_id;      /* index.Get/SetLastFreeExcessListId() (host, in/out)      */
  /* scene->useSwapping: globalCache->GetSwapStates(true), ITMHashSwapState[bucket+excess] (1 byte each); NULL otherwise.
   * AllocateSceneFromDepth then flags visible entries for swap-in and re-allocates swapped-out ones
   * (ITMSceneReconstructionEngine_CPU.cpp:250-253, 272-285) */
  unsigned char *swap_states_dev;
} itm_b200_scene;

/* ITMRenderState_VH (ITMLib/Objects/ITMRenderState_VH.h:18-70, ITMRenderState.h:20-85) */
typedef struct itm_b200_render_state {
  int *visible_entry_ids_dev;            /* int[local] */
  unsigned char *entries_visible_type_dev; /* uchar[bucket+excess] */
  int no_visible_entries;                /* host, in/out */
  float *rendering_range_image_dev;      /* Vector2f[w*h] */
  float *raycast_result_dev;             /* Vector4f[w*h] */
  unsigned char *raycast_image_dev;      /* Vector4u[w*h] */
  float *forward_projection_dev;         /* Vector4f[w*h]; only ForwardRender needs it (may be NULL otherwise) */
  int *fwd_proj_missing_points_dev;      /* int[w*h];      only ForwardRender needs it */
  int no_fwd_proj_missing_points;        /* host, out */
  int img_width, img_height;             /* renderingRangeImage->noDims; 0 = the context's depth image size.  Free-view
                                          * render states (ITMMainEngine::GetImage) may have another size */
} itm_b200_render_state;

/* ITMTrackingState (ITMLib/Objects/ITMTrackingState.h:19-85) */
typedef struct itm_b200_tracking_state {
  float *points_map_dev;             /* pointCloud->locations: Vector4f[w*h] */
  float *normals_map_dev;            /* pointCloud->colours:   Vector4f[w*h] */
  float pose_d[16];                  /* pose_d->GetM()           (host, in/out) */
  float pose_point_cloud[16];        /* pose_pointCloud->GetM()  (host, in/out) */
  int age_point_cloud;               /* host, in/out */
} itm_b200_tracking_state;

/* ITMSceneReconstructionEngine::ResetScene (Engine/ITMSceneReconstructionEngine.h:35) */
int itm_b200_reset_scene(itm_b200_ctx *ctx, itm_b200_scene *scene);

/* ITMSceneReconstructionEngine::AllocateSceneFromDepth (Engine/ITMSceneReconstructionEngine.h:41-42)
 * depth_dev = view->depth (float metres); pose_M = trackingState->pose_d->GetM(). */
int itm_b200_allocate_scene_from_depth(itm_b200_ctx *ctx, itm_b200_scene *scene, itm_b200_render_state *rs,
                                       const float *depth_dev, const float pose_M[16], int only_update_visible_list);

/* ITMSceneReconstructionEngine::IntegrateIntoScene (Engine/ITMSceneReconstructionEngine.h:47-48).  The _rgb variant is
 * the one for ITM_B200_VOXEL_S_RGB contexts: rgb_dev = view->rgb (Vector4u[w*h]). */
int itm_b200_integrate_into_scene(itm_b200_ctx *ctx, itm_b200_scene *scene, const itm_b200_render_state *rs,
                                  const float *depth_dev, const float pose_M[16]);
int itm_b200_integrate_into_scene_rgb(itm_b200_ctx *ctx, itm_b200_scene *scene, const itm_b200_render_state *rs,
                                      const float *depth_dev, const unsigned char *rgb_dev, const float pose_M[16]);

/* IITMVisualisationEngine::CreateExpectedDepths (Engine/ITMVisualisationEngine.h:46-47) */
int itm_b200_create_expected_depths(itm_b200_ctx *ctx, const itm_b200_scene *scene, itm_b200_render_state *rs,
                                    const float pose_M[16], const float intrinsics[4]);

/* IITMVisualisationEngine::CreateICPMaps (Engine/ITMVisualisationEngine.h:66-67): raycast at
 * ts->pose_d, fill points/normals/grey maps, set ts->pose_point_cloud = ts->pose_d. */
int itm_b200_create_icp_maps(itm_b200_ctx *ctx, const itm_b200_scene *scene, itm_b200_render_state *rs,
                             itm_b200_tracking_state *ts);

/* IITMVisualisationEngine::CreatePointCloud (Engine/ITMVisualisationEngine.h:62-63; CPU reference
 * ITMVisualisationEngine_CPU.cpp:242-264 and RenderPointCloud :424-462), the colour tracker's model of the scene: rays are
 * cast with inv_M = ts->pose_d->GetInvM() * view->calib->trafo_rgb_to_depth.calib (the caller's product, used as it is) and
 * the colour camera's intrinsics into rs->raycast_result, rs->raycast_image is shaded, and every pixel that kept its point
 * (with skip_points: of the odd columns of the odd rows) appends, in raster order, location = point in metres (w = 1) to
 * ts->points_map_dev and the interpolated voxel colour (w = 1; all zero for voxels without colour) to ts->normals_map_dev.
 * *no_total_points = pointCloud->noTotalPoints; ts->pose_point_cloud = ts->pose_d. */
int itm_b200_create_point_cloud(itm_b200_ctx *ctx, const itm_b200_scene *scene, itm_b200_render_state *rs, itm_b200_tracking_state *ts,
                                const float inv_M[16], const float intrinsics_rgb[4], int skip_points, int *no_total_points);

/* IITMVisualisationEngine::ForwardRender (Engine/ITMVisualisationEngine.h:73-74; CPU reference
 * ITMVisualisationEngine_CPU.cpp:289-354): rs->raycast_result (the last full raycast) is projected to ts->pose_d into
 * rs->forward_projection, pixels left without a point are ray cast, rs->raycast_image is shaded from the result.
 * depth_dev = view->depth.  rs->no_fwd_proj_missing_points is set; the list itself comes out in no particular order
 * (the reference's is raster order; nothing reads it). */
int itm_b200_forward_render(itm_b200_ctx *ctx, const itm_b200_scene *scene, itm_b200_render_state *rs, const float *depth_dev,
                            const itm_b200_tracking_state *ts);

/* IITMVisualisationEngine::FindVisibleBlocks (Engine/ITMVisualisationEngine.h:42-43): every allocated block with a corner
 * inside the image of (pose_M, intrinsics); rs->visible_entry_ids in ascending slot order, rs->no_visible_entries set. */
int itm_b200_find_visible_blocks(itm_b200_ctx *ctx, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                                 const float intrinsics[4]);

/* IITMVisualisationEngine::FindSurface (Engine/ITMVisualisationEngine.h:57-58): raycast from (pose_M, intrinsics) into
 * rs->raycast_result using rs->rendering_range_image. */
int itm_b200_find_surface(itm_b200_ctx *ctx, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                          const float intrinsics[4]);

/* IITMVisualisationEngine::RenderImage (Engine/ITMVisualisationEngine.h:51-53): raycast into rs->raycast_result and shade
 * into out_image_dev (Vector4u, rs image size).  type: IITMVisualisationEngine::RenderImageType. */
#define ITM_B200_RENDER_SHADED_GREYSCALE 0
#define ITM_B200_RENDER_COLOUR_FROM_VOLUME 1
#define ITM_B200_RENDER_COLOUR_FROM_NORMAL 2
int itm_b200_render_image(itm_b200_ctx *ctx, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                          const float intrinsics[4], unsigned char *out_image_dev, int type);

/* ITMSwappingEngine (Engine/ITMSwappingEngine.h:22-31; CPU reference ITMSwappingEngine_CPU.cpp).  The global cache lives in
 * HOST memory and stays with the caller (ITMGlobalCache, Objects/ITMGlobalCache.h); these are the device halves of the two
 * methods, split where the reference's own CUDA engine crosses the bus:
 *   IntegrateGlobalIntoLocal = swap_in_select -> [host: LoadFromGlobalMemory's copy loop + H2D] -> swap_in_apply
 *   SaveToGlobalMemory       = swap_out      -> [host: D2H + SetStoredData loop]
 * Entries are selected in ascending slot order, at most SDF_TRANSFER_BLOCK_NUM (0x1000) per call, like the serial loops. */
typedef struct itm_b200_swap_buffers {
  int *needed_entry_ids_dev;           /* globalCache->GetNeededEntryIDs(true):    int[0x1000]          */
  void *synced_voxel_blocks_dev;       /* globalCache->GetSyncedVoxelBlocks(true): TVoxel[0x1000 * 512] */
  unsigned char *has_synced_data_dev;  /* globalCache->GetHasSyncedData(true):     bool[0x1000]         */
} itm_b200_swap_buffers;
/* entries whose swap state is 1 -> needed_entry_ids_dev[0..*no_needed) */
int itm_b200_swap_in_select(itm_b200_ctx *ctx, const itm_b200_scene *scene, const itm_b200_swap_buffers *sw, int *no_needed);
/* combine synced_voxel_blocks_dev (where has_synced_data_dev) into the active blocks, swap state -> 2 */
int itm_b200_swap_in_apply(itm_b200_ctx *ctx, itm_b200_scene *scene, const itm_b200_swap_buffers *sw, int no_needed);
/* entries with swap state 2 that are allocated and not visible: ids -> needed_entry_ids_dev, blocks -> synced_voxel_blocks_dev,
 * the blocks are reset and returned to the free list (scene->last_free_block_id is updated), entries get ptr = -1, state 0 */
int itm_b200_swap_out(itm_b200_ctx *ctx, itm_b200_scene *scene, const itm_b200_render_state *rs, const itm_b200_swap_buffers *sw,
                      int *no_needed);

/* ITMMeshingEngine::MeshScene (Engine/ITMMeshingEngine.h:22; CPU reference ITMMeshingEngine_CPU.cpp:19-58): marching
 * cubes over every allocated voxel block.  triangles_dev = mesh->triangles (ITMMesh::Triangle = 9 floats, Objects/ITMMesh.h:17)
 * with room for no_max_triangles (ITMMesh::noMaxTriangles = SDF_LOCAL_BLOCK_NUM * 32); it is cleared and filled in the
 * reference's serial order (entry id, z, y, x, case-table order), *no_total_triangles = mesh->noTotalTriangles. */
int itm_b200_mesh_scene(itm_b200_ctx *ctx, const itm_b200_scene *scene, float *triangles_dev, unsigned no_max_triangles,
                        unsigned *no_total_triangles);

/* ITMMesh::WriteSTL / WriteOBJ (Objects/ITMMesh.h:34-118) for a HOST triangle array (9 floats per triangle): same bytes as
 * the reference writes.  No GPU needed. */
int itm_b200_write_stl(const char *file_name, const float *triangles_host, unsigned no_triangles);
int itm_b200_write_obj(const char *file_name, const float *triangles_host, unsigned no_triangles);

/* ITMViewBuilder::ConvertDepthAffineToFloat (Engine/ITMViewBuilder.h) */
int itm_b200_convert_depth_affine_to_float(itm_b200_ctx *ctx, float *depth_out_dev, const short *depth_in_dev, int w, int h,
                                           float a, float b);

/* ITMViewBuilder::ConvertDisparityToDepth (Engine/ITMViewBuilder.h; convertDisparityToDepth, DeviceAgnostic/ITMViewBuilder.h:7-20):
 * Kinect raw disparities to metres, depth = 8 * c2 * fx_depth / (c1 - d), non-positive results become -1. */
int itm_b200_convert_disparity_to_depth(itm_b200_ctx *ctx, float *depth_out_dev, const short *disparity_in_dev, int w, int h,
                                        float c1, float c2, float fx_depth);

/* ITMLowLevelEngine::FilterSubsampleWithHoles(float) (Engine/ITMLowLevelEngine.h:22) */
int itm_b200_filter_subsample_with_holes(itm_b200_ctx *ctx, float *out_dev, const float *in_dev, int w_in, int h_in);

/* The rest of ITMLowLevelEngine (Engine/ITMLowLevelEngine.h:14-27; CPU reference DeviceSpecific/CPU/ITMLowLevelEngine_CPU.cpp:12-108,
 * DeviceAgnostic/ITMLowLevelEngine.h): image helpers only the colour / Ren trackers call.  CopyImage copies `bytes` =
 * image_in->dataSize * sizeof(T); the subsamplers write (w_in / 2) x (h_in / 2) pixels; GradientX / GradientY write Vector4s
 * (short4) - interior pixels get the 3x3 Sobel / 8 of the colour channels and w = 255, and like the reference only the
 * first w*h*sizeof(Vector3s) bytes of the output are cleared beforehand. */
int itm_b200_copy_image(itm_b200_ctx *ctx, void *out_dev, const void *in_dev, size_t bytes);
int itm_b200_filter_subsample_rgba(itm_b200_ctx *ctx, unsigned char *out_dev, const unsigned char *in_dev, int w_in, int h_in);
int itm_b200_filter_subsample_with_holes_float4(itm_b200_ctx *ctx, float *out_dev, const float *in_dev, int w_in, int h_in);
int itm_b200_gradient_x(itm_b200_ctx *ctx, short *grad_dev, const unsigned char *image_dev, int w, int h);
int itm_b200_gradient_y(itm_b200_ctx *ctx, short *grad_dev, const unsigned char *image_dev, int w, int h);

/* ITMDepthTracker::ComputeGandH (Engine/ITMDepthTracker.h:56): one evaluation of the
 * point-to-plane error at approx_inv_pose.  hessian is the full 6x6 (column-major, r + c*6) as
 * the reference returns it; returns noValidPoints through *no_valid_points. */
int itm_b200_compute_g_and_h(itm_b200_ctx *ctx, const float *level_depth_dev, int w, int h, const float view_intrinsics[4],
                             const float *points_map_dev, const float *normals_map_dev, int scene_w, int scene_h,
                             const float scene_intrinsics[4], const float approx_inv_pose[16], const float scene_pose[16],
                             float dist_thresh, int iteration_type, float *f, float nabla[6], float hessian[36],
                             int *no_valid_points);

/* ITMWeightedICPTracker::ComputeGandH (Engine/ITMWeightedICPTracker.h:58; CPU reference
 * ITMWeightedICPTracker_CPU.cpp:14-85): the same evaluation with the per-pixel weight 0.0012 / sigma_z * 0.5 + 0.5 taken from
 * level_weight_dev (the level of the weight hierarchy built from view->depthUncertainty).  ITMWeightedICPTracker's own host
 * loop (Engine/ITMWeightedICPTracker.cpp:164-192, plain Gauss-Newton) runs on top of it unchanged. */
int itm_b200_compute_g_and_h_weighted(itm_b200_ctx *ctx, const float *level_depth_dev, const float *level_weight_dev, int w, int h,
                                      const float view_intrinsics[4], const float *points_map_dev, const float *normals_map_dev,
                                      int scene_w, int scene_h, const float scene_intrinsics[4], const float approx_inv_pose[16],
                                      const float scene_pose[16], float dist_thresh, int iteration_type, float *f, float nabla[6],
                                      float hessian[36], int *no_valid_points);

/* ITMViewBuilder::DepthFiltering (Engine/ITMViewBuilder.h:29; filterDepth, DeviceAgnostic/ITMViewBuilder.h:31-56): one pass of
 * the 5x5 bilateral depth filter (settings.useBilateralFilter runs five, ITMViewBuilder_CPU.cpp:50-59).  Uses expf: results
 * agree with the CPU reference to a few ulp, not bit for bit. */
int itm_b200_depth_filtering(itm_b200_ctx *ctx, float *out_dev, const float *in_dev, int w, int h);
/* ITMViewBuilder::ComputeNormalAndWeights (Engine/ITMViewBuilder.h:30; computeNormalAndWeight :59-114): view->depthNormal
 * (Vector4f) and view->depthUncertainty (sigma_z) of settings.modelSensorNoise.  Interior pixels only, like the reference. */
int itm_b200_compute_normal_and_weights(itm_b200_ctx *ctx, float *normal_out_dev, float *sigma_z_out_dev, const float *depth_dev, int w,
                                        int h, const float intrinsics[4]);

/* ITMTracker::TrackCamera (Engine/ITMTracker.h:26) for the depth ICP tracker: builds the depth
 * pyramid from depth_dev and runs the whole Levenberg-Marquardt loop on the device; ts->pose_d
 * is updated.  The map/pose fields of *ts are the tracker's inputs. */
int itm_b200_track_camera(itm_b200_ctx *ctx, const float *depth_dev, itm_b200_tracking_state *ts);

/* ITMWeightedICPTracker::TrackCamera (Engine/ITMWeightedICPTracker.cpp:164-192) with the whole Gauss-Newton loop on the device:
 * depth pyramid from depth_dev, weight pyramid from depth_uncertainty_dev (view->depthUncertainty), per-pixel weights, no host
 * round trip per evaluation.  ts->pose_d is updated. */
int itm_b200_track_camera_weighted(itm_b200_ctx *ctx, const float *depth_dev, const float *depth_uncertainty_dev, itm_b200_tracking_state *ts);

/* ===================================================================================== *
 *  Layer B - ITMMainEngine: owns scene, render state, tracking state and view in HBM    *
 *  and runs ProcessFrame (ITMLib/Engine/ITMMainEngine.cpp:111-127) without host syncs.  *
 * ===================================================================================== */

typedef struct itm_b200_engine itm_b200_engine;

int itm_b200_engine_create(const itm_b200_params *params, itm_b200_engine **out);
void itm_b200_engine_destroy(itm_b200_engine *e);
/* denseMapper->ResetScene + fresh tracking state */
int itm_b200_engine_reset(itm_b200_engine *e);

/* ITMMainEngine::ProcessFrame(rgbImage, rawDepthImage).  rgb_host (Vector4u[w*h]) may be NULL;
 * raw_depth_host is short[w*h] in host memory (pinned memory avoids a staging copy).  Returns
 * after the frame is complete; pose_out (may be NULL) receives trackingState->pose_d->GetM(). */
int itm_b200_engine_process_frame(itm_b200_engine *e, const unsigned char *rgb_host, const short *raw_depth_host,
                                  float pose_out[16]);

/* The fork's deployment mode (settings.trackerType == TRACKER_EXTERNAL): the pose of this frame comes from outside
 * (RosPoseSourceEngine.cpp:112-118 does pose_d->SetT / SetR before ProcessFrame) and the frame is fused without ICP
 * (ITMExternalTracker::TrackCamera is empty, Engine/ITMExternalTracker.cpp:27-30).  pose_M_in = the camera-from-world matrix
 * pose_d->GetM() (column-major); works with any tracker_type - the tracker is skipped for this frame.  pose_M_in == NULL
 * keeps the current pose_d (a TRACKER_EXTERNAL engine then simply does not move). */
int itm_b200_engine_process_frame_with_pose(itm_b200_engine *e, const unsigned char *rgb_host, const short *raw_depth_host,
                                            const float pose_M_in[16], float pose_out[16]);

/* ---- streaming ProcessFrame ------------------------------------------------------------------------------------
 * ITMMainEngine::ProcessFrame split into submit and wait so that the upload of frame k+1 and the read-back of frame k's pose
 * overlap the fusion of frame k: submit copies the (pinned) host images into one of the engine's staging buffers on a copy
 * stream and enqueues the frame behind it; the frame's last kernel publishes pose + counters into host-mapped memory, which
 * wait_frame polls - no stream synchronisation, no D2H copy call on the critical path.  Up to ITM_B200_MAX_IN_FLIGHT
 * frames may be submitted ahead of the oldest one not yet waited for (submit blocks beyond that); the host buffers of a
 * frame may be reused once its wait_frame returned (or after itm_b200_engine_sync).  Tickets count from 1.
 * pose_M_in: optional external pose for this frame (see process_frame_with_pose); NULL = track / keep.
 * counters = {noVisibleEntries, lastFreeBlockId, lastFreeExcessListId, allocFailures, errorFlags, icpEvaluations}. */
#define ITM_B200_MAX_IN_FLIGHT 4
int itm_b200_engine_submit_frame(itm_b200_engine *e, const unsigned char *rgb_host, const short *raw_depth_host,
                                 const float pose_M_in[16], unsigned long long *ticket);
int itm_b200_engine_wait_frame(itm_b200_engine *e, unsigned long long ticket, float pose_out[16], int counters[6]);

/* ---- spatial sharding of ONE scene across the GPUs of one NVLink domain (BASELINE configs[2]) ----------
 * One process per GPU; rank 0's raw depth frame is broadcast to all (NCCL, by the host: infinitam_b200/multi.py).
 *   - The index - hash-table positions and chain links, excess list, visible list, pose, images - is replicated: every rank runs
 *     the same deterministic allocation on the same frame, so it is identical on every rank and identical to a single GPU.
 *   - The voxel payload is partitioned by block coordinate: slabs of `thickness_blocks` blocks along `axis`, rank r owning
 *     [origin_block + r * thickness, origin_block + (r + 1) * thickness) (the first / last rank open-ended).  A block's voxels
 *     exist only where it is RESIDENT: on its owner and, as a one-block halo for the trilinear taps of the ray cast, on the
 *     neighbouring slab's rank.  Elsewhere its hash entry carries ptr = -1.  params.sdf_local_block_num is the pool PER RANK,
 *     so the scene a box can hold grows with the number of GPUs.
 *   - Every rank integrates its resident blocks, renders the expected depths from all visible blocks (their positions are in
 *     the replicated index) and marches every ray over the range a single GPU would, noting whether a sample or trilinear tap
 *     fell into a block held elsewhere.  A ray that never met one has seen exactly what a single GPU holds: its result - hit
 *     or miss - is the single-GPU result bit for bit.  After one cross-GPU barrier every rank composes the full image from
 *     such complete results, pulling through NVLink peer pointers the tiles it could not complete itself; rays no rank could
 *     complete are marched once more with peer reads of the voxels held elsewhere (shard_export / shard_attach below).  ICP maps
 *     and the tracker run replicated on the composed image - which is the single GPU's - so all ranks compute the identical
 *     pose and neither a pose broadcast nor a G/H all-reduce is needed.
 * The peer-visible buffers (two partial images of width*height*16 bytes, two tile-flag arrays of ceil(w/16)*ceil(h/8) bytes,
 * ITM_B200_MAX_SHARDS barrier words) are allocated with itm_b200_ipc_alloc, their handles exchanged by the host
 * (torch.distributed all_gather in infinitam_b200/multi.py) and opened with itm_b200_ipc_open. */
#define ITM_B200_MAX_SHARDS 8
#define ITM_B200_IPC_HANDLE_BYTES 64
typedef struct itm_b200_shard {
  int rank, world;
  int axis;                                           /* 0 = x, 1 = y, 2 = z (block coordinates) */
  int origin_block, thickness_blocks;
  void *partial_raycast_dev[2][ITM_B200_MAX_SHARDS];  /* Vector4f[w*h] per frame parity and rank ([..][rank] = the local one) */
  void *tile_hit_dev[2][ITM_B200_MAX_SHARDS];         /* unsigned char[tiles] per frame parity and rank */
  void *barrier_flags_dev[ITM_B200_MAX_SHARDS];       /* unsigned[ITM_B200_MAX_SHARDS] of every rank, zero-initialised */
  void *stream;                                       /* cudaStream_t the frames are enqueued on (NULL: a private one) */
  int halo_blocks;                                    /* blocks beyond its slab a rank keeps resident (and integrates redundantly); 0 = 1.
                                                       * A ray can be marched by a rank as long as every allocated block it samples is
                                                       * resident there: a wider halo leaves fewer pixels that no rank can complete
                                                       * (itm_b200_engine_shard_unresolved) at the price of more redundant integration */
} itm_b200_shard;
int itm_b200_engine_create_sharded(const itm_b200_params *params, const itm_b200_shard *shard, itm_b200_engine **out);
/* Optional second step (all ranks, before the first frame): IPC handles of this rank's voxel pool and hash table, to be opened by
 * the peers (itm_b200_ipc_open) and handed to their engines with shard_attach ([rank] is ignored).  With peers attached, the
 * rays no rank can march completely on its own voxels are marched once more after the composition, reading the blocks held
 * elsewhere from their owners over NVLink - the composed raycast image is then the single GPU's in every pixel.  Without,
 * those pixels are reported as misses (itm_b200_engine_shard_unresolved counts them either way). */
int itm_b200_engine_shard_export(itm_b200_engine *e, unsigned char voxels_handle[ITM_B200_IPC_HANDLE_BYTES],
                                 unsigned char hash_handle[ITM_B200_IPC_HANDLE_BYTES]);
int itm_b200_engine_shard_attach(itm_b200_engine *e, void *const peer_voxels_dev[ITM_B200_MAX_SHARDS], void *const peer_hash_dev[ITM_B200_MAX_SHARDS]);
/* cudaMalloc + zero fill + cudaIpcGetMemHandle / cudaIpcOpenMemHandle (peer access enabled lazily) / close / free */
int itm_b200_ipc_alloc(size_t bytes, void **dev_ptr, unsigned char handle[ITM_B200_IPC_HANDLE_BYTES]);
int itm_b200_ipc_open(const unsigned char handle[ITM_B200_IPC_HANDLE_BYTES], void **dev_ptr);
int itm_b200_ipc_close(void *dev_ptr);
int itm_b200_ipc_free(void *dev_ptr);
/* rank that owns the voxel block at block coordinate (x, y, z); itm_b200_shard_block_resident: 1 if `rank` keeps its payload */
int itm_b200_shard_owner_of_block(int x, int y, int z, int world, int axis, int origin_block, int thickness_blocks);
int itm_b200_shard_block_resident(int x, int y, int z, int rank, int world, int axis, int origin_block, int thickness_blocks);  /* halo 1 */
int itm_b200_shard_block_resident_halo(int x, int y, int z, int rank, int world, int axis, int origin_block, int thickness_blocks, int halo_blocks);

/* The cudaStream_t every frame of this engine is enqueued on (borrowed; valid until destroy): lets the host order its own
 * work - e.g. producing the next depth frame on the device - before or after frames without a host synchronisation. */
int itm_b200_engine_get_stream(itm_b200_engine *e, void **stream);

/* Asynchronous device-to-device copy into one of the engine's buffers (ITM_B200_BUF_*), ordered on the engine's stream: e.g. a
 * depth frame produced on the device placed into ITM_B200_BUF_RAW_DEPTH ahead of itm_b200_engine_enqueue_frame_dev(e, that buffer). */
int itm_b200_engine_copy_to_buffer_dev(itm_b200_engine *e, int which, const void *src_dev, size_t bytes);

/* Same frame, input already resident in HBM, enqueued asynchronously on the engine's stream.  raw_depth_dev may be the engine's
 * own ITM_B200_BUF_RAW_DEPTH buffer (no copy then). */
int itm_b200_engine_enqueue_frame_dev(itm_b200_engine *e, const short *raw_depth_dev);
/* Wait for everything enqueued; counters = {noVisibleEntries, lastFreeBlockId,
 * lastFreeExcessListId, allocFailures, errorFlags, icpEvaluations of last frame}. */
int itm_b200_engine_sync(itm_b200_engine *e, float pose_out[16], int counters[6]);

/* Single stages on the engine's own state, for stage-by-stage parity tests ("teacher forcing"):
 * 0 view (needs a frame uploaded with itm_b200_engine_upload_depth), 1 track, 2 allocate,
 * 3 integrate, 4 expected depths, 5 raycast + ICP maps, 6 swap in / out (use_swapping engines; part of stage 3's slot in
 * a whole frame, ITMDenseMapper.cpp:59-64), 7 ForwardRender (what Prepare runs instead of stage 5 when
 * !requiresFullRendering), 8 the requiresFullRendering decision of ITMTrackingController::Track. */
int itm_b200_engine_upload_depth(itm_b200_engine *e, const short *raw_depth_host);
int itm_b200_engine_run_stage(itm_b200_engine *e, int stage);

/* Device buffers of the engine's state (borrowed pointers, valid until destroy). */
enum {
  ITM_B200_BUF_VOXELS = 0, ITM_B200_BUF_HASH, ITM_B200_BUF_VBA_ALLOC_LIST, ITM_B200_BUF_EXCESS_ALLOC_LIST,
  ITM_B200_BUF_VISIBLE_IDS, ITM_B200_BUF_VISIBLE_TYPES, ITM_B200_BUF_DEPTH, ITM_B200_BUF_MINMAX, ITM_B200_BUF_RAYCAST_RESULT,
  ITM_B200_BUF_RAYCAST_IMAGE, ITM_B200_BUF_POINTS, ITM_B200_BUF_NORMALS, ITM_B200_BUF_RAW_DEPTH, ITM_B200_BUF_PYRAMID_1,
  ITM_B200_BUF_PYRAMID_2, ITM_B200_BUF_PYRAMID_3, ITM_B200_BUF_PYRAMID_4, ITM_B200_BUF_RGB, ITM_B200_BUF_SWAP_STATES,
  ITM_B200_BUF_FORWARD_PROJECTION, ITM_B200_BUF_FWD_MISSING_POINTS,
  /* renderState_freeview of the last GetImage call (ITMMainEngine.cpp:176-180) */
  ITM_B200_BUF_FREEVIEW_VISIBLE_IDS, ITM_B200_BUF_FREEVIEW_MINMAX, ITM_B200_BUF_FREEVIEW_RAYCAST_RESULT, ITM_B200_BUF_FREEVIEW_IMAGE,
  ITM_B200_BUF_COUNT
};
int itm_b200_engine_get_buffer(itm_b200_engine *e, int which, void **dev_ptr, size_t *bytes);
/* Blocking copies between one of those buffers and host memory (the reference's
 * MemoryBlock::UpdateHostFromDevice / UpdateDeviceFromHost, ORUtils/MemoryBlock.h:112-121). */
int itm_b200_engine_read_buffer(itm_b200_engine *e, int which, void *host_dst, size_t bytes, size_t offset);
int itm_b200_engine_write_buffer(itm_b200_engine *e, int which, const void *host_src, size_t bytes, size_t offset);

/* ITMGlobalCache of a swapping engine (host memory, borrowed): hasStoredData[bucket+excess] and the stored voxel
 * blocks (ITMLib/Objects/ITMGlobalCache.h:21-36).  *swapped_in / *swapped_out: entries moved by the last frame.
 * The engine keeps the cache as a pool in host-mapped memory (params.swap_cache_blocks); this call waits for the pending
 * frames and materialises the reference's dense layout from it (for inspection and tests - not a per-frame call).  With
 * has_stored_data == stored_voxel_blocks == NULL only the two counts are fetched. */
int itm_b200_engine_global_cache(itm_b200_engine *e, const unsigned char **has_stored_data, const void **stored_voxel_blocks,
                                 int *swapped_in, int *swapped_out);

/* Host-visible tracking / scene state: pose_d, pose_pointCloud (column-major), and
 * state6 = {noVisibleEntries, lastFreeBlockId, lastFreeExcessListId, age_pointCloud, requiresFullRendering,
 * noFwdProjMissingPoints} (set_state ignores the last two). */
int itm_b200_engine_get_state(itm_b200_engine *e, float pose_d[16], float pose_point_cloud[16], int state6[6]);
int itm_b200_engine_set_state(itm_b200_engine *e, const float pose_d[16], const float pose_point_cloud[16], const int state6[6]);

/* What ITMTrackingController::Prepare does for TRACKER_COLOR (Engine/ITMTrackingController.cpp:22-28): CreateExpectedDepths
 * at pose_rgb = trafo_rgb_to_depth.calib_inv * pose_d with the colour camera's intrinsics, then CreatePointCloud (see
 * itm_b200_create_point_cloud).  Here it is a query: it renders into buffers of its own and leaves the live render and
 * tracking state - the depth tracker's maps - untouched.  trafo_rgb_to_depth: ITMExtrinsics::calib (column-major 4x4), NULL =
 * identity; intrinsics_rgb: NULL = the depth camera's.  locations_host / colours_host: Vector4f[capacity_points] or NULL;
 * image_host: Vector4u[w*h] (the shaded raycast) or NULL.  *no_total_points is the full count even if fewer were copied. */
int itm_b200_engine_create_point_cloud(itm_b200_engine *e, const float trafo_rgb_to_depth[16], const float intrinsics_rgb[4], int skip_points,
                                       float *locations_host, float *colours_host, int capacity_points, unsigned char *image_host,
                                       int *no_total_points);

/* ITMMainEngine::GetImage (ITMLib/Engine/ITMMainEngine.cpp:134-192).  image_type is ITMMainEngine::GetImageType
 * (Engine/ITMMainEngine.h:78-87).  out_host receives Vector4u[out_w * out_h]; for the ORIGINAL_* / SCENERAYCAST types
 * out_w x out_h must be the sensor size; the FREECAMERA types render a view of any size from (pose_M, intrinsics) through
 * FindVisibleBlocks + CreateExpectedDepths + RenderImage on the engine's renderState_freeview. */
#define ITM_B200_IMAGE_ORIGINAL_RGB 0
#define ITM_B200_IMAGE_ORIGINAL_DEPTH 1
#define ITM_B200_IMAGE_SCENERAYCAST 2
#define ITM_B200_IMAGE_FREECAMERA_SHADED 3
#define ITM_B200_IMAGE_FREECAMERA_COLOUR_FROM_VOLUME 4
#define ITM_B200_IMAGE_FREECAMERA_COLOUR_FROM_NORMAL 5
int itm_b200_engine_get_image(itm_b200_engine *e, int image_type, const float pose_M[16], const float intrinsics[4],
                              unsigned char *out_host, int out_w, int out_h);

/* ITMMainEngine::UpdateMesh / SaveSceneToMesh (ITMMainEngine.cpp:97-109): meshes the engine's scene into a device mesh of
 * ITMMesh::noMaxTriangles and copies the first min(noTotalTriangles, capacity_triangles) triangles to triangles_host
 * (9 floats each; may be NULL to only count).  *no_total_triangles = mesh->noTotalTriangles. */
int itm_b200_engine_mesh_scene(itm_b200_engine *e, float *triangles_host, unsigned capacity_triangles, unsigned *no_total_triangles);
/* SaveSceneToMesh: MeshScene + ITMMesh::WriteSTL */
int itm_b200_engine_save_scene_to_mesh(itm_b200_engine *e, const char *file_name);

/* ICP evaluations (ComputeGandH calls) per pyramid level during the last frame fetched by
 * itm_b200_engine_sync / _process_frame (level 0 = full resolution). */
int itm_b200_engine_icp_stats(itm_b200_engine *e, int evals_per_level[ITM_B200_MAX_LEVELS]);

/* Per-stage device times (CUDA events) of the last processed frame in milliseconds:
 * {h2d+view, track, allocate, integrate, expected depths, raycast, icp maps, total}.
 * Enabled by itm_b200_engine_set_profiling(e, 1) (an event record at every stage boundary: a node between two kernels of the
 * frame graph each, ~1.5 us apiece).  set_profiling(e, 2) records the frame's start and end only: ms8[7] is then the
 * frame's device time without those gaps, ms8[0..6] are 0. */
int itm_b200_engine_set_profiling(itm_b200_engine *e, int on);
int itm_b200_engine_stage_times(itm_b200_engine *e, float ms8[8]);

/* Sharded engines with set_profiling(1): device time of the last frame's partial ray cast, of the wait at the cross-GPU barrier
 * (= how far behind the slowest rank was) and of the nearest-hit composition (NVLink peer reads), in milliseconds. */
int itm_b200_engine_shard_times(itm_b200_engine *e, float ms3[3]);
/* Sharded engines: pixels of the last frame's composed raycast whose ray no rank could march completely on its own voxels
 * (it passes through allocated blocks of two slabs beyond the one-block halo); they are reported as misses.  Every other
 * pixel of the composed image - hit or miss - is bit-identical to a single GPU's.  Waits for the pending frames. */
int itm_b200_engine_shard_unresolved(itm_b200_engine *e, int *pixels);

/* ---- host-side pose arithmetic (no GPU needed; used by the adapter and by tests) ---------- */
/* Matrix4f::inv (ORUtils/Matrix.h:162-218) */
int itm_b200_mat4_inv(const float m[16], float out[16]);
/* ITMPose::SetInvM + Coerce + GetM/GetInvM (ITMLib/Objects/ITMPose.cpp:309-326) */
int itm_b200_pose_from_inv_m_coerced(const float inv_m[16], float m_out[16], float inv_out[16], float params_out[6]);
/* ITMDepthTracker::ComputeDelta (ITMLib/Engine/ITMDepthTracker.cpp:85-102) */
int itm_b200_compute_delta(const float nabla[6], const float hessian[36], int short_iteration, float step_out[6]);

#ifdef __cplusplus
}
#endif
#endif /* ITM_B200_H */
