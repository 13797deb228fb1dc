// itm_b200_adapter.hpp - the B200 engines behind ITMLib's own engine interfaces.
//
// This header is compiled INSIDE the reference tree (it includes ITMLib headers); it is the code a
// maintainer adds next to ITMLib/Engine/DeviceSpecific/{CPU,CUDA}/ to get a DEVICE_B200-style
// engine set.  Every method forwards 1:1 to a function of the C ABI (include/itm_b200.h); nothing
// of the path is computed here and there is no CPU fallback - a failing C call becomes
// DIEWITHEXCEPTION (ORUtils/PlatformIndependence.h:34-38), like the reference's own error path.
//
//   reference interface (file:line)                                    -> adapter class below
//   ITMSceneReconstructionEngine<TVoxel,TIndex>                        -> ITMSceneReconstructionEngine_B200
//       (ITMLib/Engine/ITMSceneReconstructionEngine.h:29-52)
//   ITMVisualisationEngine<TVoxel,TIndex> / IITMVisualisationEngine    -> ITMVisualisationEngine_B200
//       (ITMLib/Engine/ITMVisualisationEngine.h:18-110)
//   ITMSwappingEngine<TVoxel,TIndex>                                   -> ITMSwappingEngine_B200
//       (ITMLib/Engine/ITMSwappingEngine.h:22-31)
//   ITMMeshingEngine<TVoxel,TIndex>::MeshScene                         -> ITMMeshingEngine_B200
//       (ITMLib/Engine/ITMMeshingEngine.h:15-30)
//   ITMDepthTracker (TrackCamera + ComputeGandH)                       -> ITMDepthTracker_B200
//       (ITMLib/Engine/ITMDepthTracker.h:24-66)
//   ITMWeightedICPTracker (ComputeGandH under the reference's own loop) -> ITMWeightedICPTracker_B200
//       (ITMLib/Engine/ITMWeightedICPTracker.h:23-68)
//   ITMLowLevelEngine::FilterSubsampleWithHoles(float)                 -> ITMLowLevelEngine_B200
//       (ITMLib/Engine/ITMLowLevelEngine.h:16-34)
//   ITMViewBuilder::UpdateView / ConvertDepthAffineToFloat             -> ITMViewBuilder_B200
//       (ITMLib/Engine/ITMViewBuilder.h:19-60)
//
// All reference objects handed to these engines (ITMScene, ITMRenderState_VH, ITMTrackingState,
// ITMView) must have been constructed with MEMORYDEVICE_CUDA / useGPU = true, exactly as the
// reference's DEVICE_CUDA branch does (ITMMainEngine.cpp:17-18, ITMTrackingController.h:44): the
// library borrows their device pointers for the duration of a call and never frees them.
//
// The voxel-block-hash index with ITMVoxel_s or ITMVoxel_s_rgb is supported (pass the voxel type to ITMB200Context);
// other instantiations fail at compile time.
#pragma once

#include <string>

#include "itm_b200.h"

#include "ITMLib/Engine/ITMDepthTracker.h"
#include "ITMLib/Engine/ITMLowLevelEngine.h"
#include "ITMLib/Engine/ITMMeshingEngine.h"
#include "ITMLib/Engine/ITMSceneReconstructionEngine.h"
#include "ITMLib/Engine/ITMSwappingEngine.h"
#include "ITMLib/Engine/ITMViewBuilder.h"
#include "ITMLib/Engine/ITMWeightedICPTracker.h"
#include "ITMLib/Engine/ITMVisualisationEngine.h"
#include "ITMLib/Objects/ITMRenderState_VH.h"
#include "ITMLib/Utils/ITMLibSettings.h"

namespace ITMLib {
namespace Engine {

inline void itm_b200_check(int rc, const char *what) {
  if (rc != ITM_B200_OK) {
    static std::string msg;  // DIEWITHEXCEPTION keeps the pointer only until the throw
    msg = std::string(what) + ": " + itm_b200_last_error();
    DIEWITHEXCEPTION(msg.c_str());
  }
}

/// One per engine set (= one scene): owns the library context, i.e. scratch memory and the stream.
class ITMB200Context {
 public:
  itm_b200_ctx *ctx;
  itm_b200_params params;

  /// voxelType: ITM_B200_VOXEL_S for ITMVoxel_s, ITM_B200_VOXEL_S_RGB for ITMVoxel_s_rgb (ITMLibDefines.h:205 picks one)
  ITMB200Context(const ITMLibSettings *settings, const ITMRGBDCalib *calib, Vector2i imgSize_d, int device = 0,
                 int voxelType = ITM_B200_VOXEL_S) : ctx(NULL) {
    itm_b200_default_params(&params, imgSize_d.x, imgSize_d.y);
    params.voxel_type = voxelType;
    const Vector4f &kr = calib->intrinsics_rgb.projectionParamsSimple.all;
    params.rgb_fx = kr.x; params.rgb_fy = kr.y; params.rgb_cx = kr.z; params.rgb_cy = kr.w;
    for (int i = 0; i < 16; ++i) params.trafo_rgb_to_depth_inv[i] = calib->trafo_rgb_to_depth.calib_inv.m[i];
    params.use_swapping = settings->useSwapping ? 1 : 0;
    params.use_approximate_raycast = settings->useApproximateRaycast ? 1 : 0;
    const Vector4f &k = calib->intrinsics_d.projectionParamsSimple.all;
    params.fx = k.x; params.fy = k.y; params.cx = k.z; params.cy = k.w;
    const ITMSceneParams &sp = settings->sceneParams;
    params.voxel_size = sp.voxelSize;
    params.mu = sp.mu;
    params.max_w = sp.maxW;
    params.view_frustum_min = sp.viewFrustum_min;
    params.view_frustum_max = sp.viewFrustum_max;
    params.stop_integrating_at_max_w = sp.stopIntegratingAtMaxW ? 1 : 0;
    params.depth_calib_a = calib->disparityCalib.params.x;
    params.depth_calib_b = calib->disparityCalib.params.y;
    params.sdf_local_block_num = SDF_LOCAL_BLOCK_NUM;
    params.sdf_bucket_num = SDF_BUCKET_NUM;
    params.sdf_excess_list_size = SDF_EXCESS_LIST_SIZE;
    params.no_hierarchy_levels = settings->noHierarchyLevels;
    for (int l = 0; l < settings->noHierarchyLevels && l < ITM_B200_MAX_LEVELS; ++l)
      params.tracking_regime[l] = (int)settings->trackingRegime[l];  // same numeric values, ITMLibDefines.h:278-283
    params.no_icp_run_till_level = settings->noICPRunTillLevel;
    params.depth_tracker_icp_threshold = settings->depthTrackerICPThreshold;
    params.depth_tracker_termination_threshold = settings->depthTrackerTerminationThreshold;
    params.device = device;
    // ITMLib's host objects work on the legacy default stream (cudaMemset in MemoryBlock::Clear, blocking cudaMemcpy); the
    // engines run on that stream too, so that they are ordered after whatever the host code still has in flight there
    itm_b200_check(itm_b200_ctx_create(&params, (void *)cudaStreamLegacy, &ctx), "itm_b200_ctx_create");
  }
  ~ITMB200Context() { itm_b200_ctx_destroy(ctx); }

 private:
  ITMB200Context(const ITMB200Context &);
  ITMB200Context &operator=(const ITMB200Context &);
};

namespace b200_detail {

template <class TVoxel>
inline itm_b200_scene scene_view(ITMScene<TVoxel, ITMVoxelBlockHash> *scene) {
  itm_b200_scene s;
  s.voxel_blocks_dev = scene->localVBA.GetVoxelBlocks();
  s.hash_entries_dev = scene->index.GetEntries();
  s.vba_allocation_list_dev = scene->localVBA.GetAllocationList();
  s.excess_allocation_list_dev = scene->index.GetExcessAllocationList();
  s.last_free_block_id = scene->localVBA.lastFreeBlockId;
  s.last_free_excess_list_id = scene->index.GetLastFreeExcessListId();
  s.swap_states_dev = scene->useSwapping ? (unsigned char *)scene->globalCache->GetSwapStates(true) : NULL;
  return s;
}

inline itm_b200_render_state render_state_view(ITMRenderState_VH *rs) {
  itm_b200_render_state r;
  r.visible_entry_ids_dev = rs->GetVisibleEntryIDs();
  r.entries_visible_type_dev = rs->GetEntriesVisibleType();
  r.no_visible_entries = rs->noVisibleEntries;
  r.rendering_range_image_dev = (float *)rs->renderingRangeImage->GetData(MEMORYDEVICE_CUDA);
  r.raycast_result_dev = (float *)rs->raycastResult->GetData(MEMORYDEVICE_CUDA);
  r.raycast_image_dev = (unsigned char *)rs->raycastImage->GetData(MEMORYDEVICE_CUDA);
  r.forward_projection_dev = (float *)rs->forwardProjection->GetData(MEMORYDEVICE_CUDA);
  r.fwd_proj_missing_points_dev = rs->fwdProjMissingPoints->GetData(MEMORYDEVICE_CUDA);
  r.no_fwd_proj_missing_points = rs->noFwdProjMissingPoints;
  r.img_width = rs->renderingRangeImage->noDims.x;   // free-view render states have their own size (ITMMainEngine.cpp:176)
  r.img_height = rs->renderingRangeImage->noDims.y;
  return r;
}

inline itm_b200_tracking_state tracking_state_view(ITMTrackingState *ts) {
  itm_b200_tracking_state t;
  t.points_map_dev = (float *)ts->pointCloud->locations->GetData(MEMORYDEVICE_CUDA);
  t.normals_map_dev = (float *)ts->pointCloud->colours->GetData(MEMORYDEVICE_CUDA);
  for (int i = 0; i < 16; ++i) {
    t.pose_d[i] = ts->pose_d->GetM().m[i];
    t.pose_point_cloud[i] = ts->pose_pointCloud->GetM().m[i];
  }
  t.age_point_cloud = ts->age_pointCloud;
  return t;
}

}  // namespace b200_detail

// ---------------------------------------------------------------------------------------------
template <class TVoxel, class TIndex>
class ITMSceneReconstructionEngine_B200;  // only the voxel-block-hash specialisation exists

template <class TVoxel>
class ITMSceneReconstructionEngine_B200<TVoxel, ITMVoxelBlockHash> : public ITMSceneReconstructionEngine<TVoxel, ITMVoxelBlockHash> {
  ITMB200Context *c;

 public:
  explicit ITMSceneReconstructionEngine_B200(ITMB200Context *context) : c(context) {
    // the library reads and writes the reference's packed voxels directly
    static_assert((sizeof(TVoxel) == 4 && !TVoxel::hasColorInformation) || (sizeof(TVoxel) == 8 && TVoxel::hasColorInformation),
                  "libitm_b200: ITMVoxel_s or ITMVoxel_s_rgb");
    if ((c->params.voxel_type == ITM_B200_VOXEL_S_RGB) != (bool)TVoxel::hasColorInformation)
      DIEWITHEXCEPTION("libitm_b200: the context was created for another voxel type");
  }

  void ResetScene(ITMScene<TVoxel, ITMVoxelBlockHash> *scene) {
    itm_b200_scene s = b200_detail::scene_view(scene);
    itm_b200_check(itm_b200_reset_scene(c->ctx, &s), "ResetScene");
    scene->localVBA.lastFreeBlockId = s.last_free_block_id;
    scene->index.SetLastFreeExcessListId(s.last_free_excess_list_id);
  }

  void AllocateSceneFromDepth(ITMScene<TVoxel, ITMVoxelBlockHash> *scene, const ITMView *view, const ITMTrackingState *trackingState,
                              const ITMRenderState *renderState, bool onlyUpdateVisibleList = false) {
    ITMRenderState_VH *rsVH = (ITMRenderState_VH *)renderState;
    itm_b200_scene s = b200_detail::scene_view(scene);
    itm_b200_render_state r = b200_detail::render_state_view(rsVH);
    itm_b200_check(itm_b200_allocate_scene_from_depth(c->ctx, &s, &r, view->depth->GetData(MEMORYDEVICE_CUDA),
                                                      trackingState->pose_d->GetM().m, onlyUpdateVisibleList ? 1 : 0),
                   "AllocateSceneFromDepth");
    // host-visible counters other ITMLib code reads (ITMRenderState_VH.h:36, ITMLocalVBA.h:36, ITMVoxelBlockHash.h:83-84)
    rsVH->noVisibleEntries = r.no_visible_entries;
    scene->localVBA.lastFreeBlockId = s.last_free_block_id;
    scene->index.SetLastFreeExcessListId(s.last_free_excess_list_id);
  }

  void IntegrateIntoScene(ITMScene<TVoxel, ITMVoxelBlockHash> *scene, const ITMView *view, const ITMTrackingState *trackingState,
                          const ITMRenderState *renderState) {
    itm_b200_scene s = b200_detail::scene_view(scene);
    const itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)renderState);
    if (TVoxel::hasColorInformation)
      itm_b200_check(itm_b200_integrate_into_scene_rgb(c->ctx, &s, &r, view->depth->GetData(MEMORYDEVICE_CUDA),
                                                       (const unsigned char *)view->rgb->GetData(MEMORYDEVICE_CUDA), trackingState->pose_d->GetM().m),
                     "IntegrateIntoScene");
    else
      itm_b200_check(itm_b200_integrate_into_scene(c->ctx, &s, &r, view->depth->GetData(MEMORYDEVICE_CUDA), trackingState->pose_d->GetM().m),
                     "IntegrateIntoScene");
  }
};

// ---------------------------------------------------------------------------------------------
/// Host swapping.  The global cache (host memory) and its transfer buffers are the reference's own ITMGlobalCache; the
/// host halves below are what ITMSwappingEngine_CUDA does between its kernels (ITMSwappingEngine_CUDA.cu:46-93, 121-178).
template <class TVoxel, class TIndex>
class ITMSwappingEngine_B200;

template <class TVoxel>
class ITMSwappingEngine_B200<TVoxel, ITMVoxelBlockHash> : public ITMSwappingEngine<TVoxel, ITMVoxelBlockHash> {
  ITMB200Context *c;

  static itm_b200_swap_buffers buffers(ITMGlobalCache<TVoxel> *gc) {
    itm_b200_swap_buffers b;
    b.needed_entry_ids_dev = gc->GetNeededEntryIDs(true);
    b.synced_voxel_blocks_dev = gc->GetSyncedVoxelBlocks(true);
    b.has_synced_data_dev = (unsigned char *)gc->GetHasSyncedData(true);
    return b;
  }

 public:
  explicit ITMSwappingEngine_B200(ITMB200Context *context) : c(context) {}

  void IntegrateGlobalIntoLocal(ITMScene<TVoxel, ITMVoxelBlockHash> *scene, ITMRenderState *renderState) {
    (void)renderState;
    ITMGlobalCache<TVoxel> *gc = scene->globalCache;
    itm_b200_scene s = b200_detail::scene_view(scene);
    const itm_b200_swap_buffers b = buffers(gc);
    int n = 0;
    itm_b200_check(itm_b200_swap_in_select(c->ctx, &s, &b, &n), "IntegrateGlobalIntoLocal");
    if (n <= 0) return;
    TVoxel *blocksHost = gc->GetSyncedVoxelBlocks(false);
    bool *hasHost = gc->GetHasSyncedData(false);
    int *idsHost = gc->GetNeededEntryIDs(false);
    ITMSafeCall(cudaMemcpy(idsHost, b.needed_entry_ids_dev, sizeof(int) * n, cudaMemcpyDeviceToHost));
    memset(blocksHost, 0, (size_t)n * SDF_BLOCK_SIZE3 * sizeof(TVoxel));
    memset(hasHost, 0, (size_t)n * sizeof(bool));
    for (int i = 0; i < n; i++) {
      const int entryId = idsHost[i];
      if (gc->HasStoredData(entryId)) {
        hasHost[i] = true;
        memcpy(blocksHost + (size_t)i * SDF_BLOCK_SIZE3, gc->GetStoredVoxelBlock(entryId), SDF_BLOCK_SIZE3 * sizeof(TVoxel));
      }
    }
    ITMSafeCall(cudaMemcpy(b.has_synced_data_dev, hasHost, sizeof(bool) * n, cudaMemcpyHostToDevice));
    ITMSafeCall(cudaMemcpy(b.synced_voxel_blocks_dev, blocksHost, sizeof(TVoxel) * SDF_BLOCK_SIZE3 * n, cudaMemcpyHostToDevice));
    itm_b200_check(itm_b200_swap_in_apply(c->ctx, &s, &b, n), "IntegrateGlobalIntoLocal");
  }

  void SaveToGlobalMemory(ITMScene<TVoxel, ITMVoxelBlockHash> *scene, ITMRenderState *renderState) {
    ITMGlobalCache<TVoxel> *gc = scene->globalCache;
    itm_b200_scene s = b200_detail::scene_view(scene);
    const itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)renderState);
    const itm_b200_swap_buffers b = buffers(gc);
    int n = 0;
    itm_b200_check(itm_b200_swap_out(c->ctx, &s, &r, &b, &n), "SaveToGlobalMemory");
    scene->localVBA.lastFreeBlockId = s.last_free_block_id;
    if (n <= 0) return;
    TVoxel *blocksHost = gc->GetSyncedVoxelBlocks(false);
    int *idsHost = gc->GetNeededEntryIDs(false);
    ITMSafeCall(cudaMemcpy(idsHost, b.needed_entry_ids_dev, sizeof(int) * n, cudaMemcpyDeviceToHost));
    ITMSafeCall(cudaMemcpy(blocksHost, b.synced_voxel_blocks_dev, sizeof(TVoxel) * SDF_BLOCK_SIZE3 * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) gc->SetStoredData(idsHost[i], blocksHost + (size_t)i * SDF_BLOCK_SIZE3);
  }
};

// ---------------------------------------------------------------------------------------------
template <class TVoxel, class TIndex>
class ITMVisualisationEngine_B200;

template <class TVoxel>
class ITMVisualisationEngine_B200<TVoxel, ITMVoxelBlockHash> : public ITMVisualisationEngine<TVoxel, ITMVoxelBlockHash> {
  ITMB200Context *c;
  typedef ITMScene<TVoxel, ITMVoxelBlockHash> Scene;

 public:
  ITMVisualisationEngine_B200(const Scene *scene, ITMB200Context *context) : ITMVisualisationEngine<TVoxel, ITMVoxelBlockHash>(scene), c(context) {}

  ITMRenderState_VH *CreateRenderState(const Vector2i &imgSize) const {
    return new ITMRenderState_VH(ITMVoxelBlockHash::noTotalEntries, imgSize, this->scene->sceneParams->viewFrustum_min,
                                 this->scene->sceneParams->viewFrustum_max, MEMORYDEVICE_CUDA);
  }

  void CreateExpectedDepths(const ITMPose *pose, const ITMIntrinsics *intrinsics, ITMRenderState *renderState) const {
    itm_b200_scene s = b200_detail::scene_view(const_cast<Scene *>(this->scene));
    itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)renderState);
    const Vector4f &k = intrinsics->projectionParamsSimple.all;
    const float intr[4] = {k.x, k.y, k.z, k.w};
    itm_b200_check(itm_b200_create_expected_depths(c->ctx, &s, &r, pose->GetM().m, intr), "CreateExpectedDepths");
  }

  void CreateICPMaps(const ITMView *view, ITMTrackingState *trackingState, ITMRenderState *renderState) const {
    (void)view;
    itm_b200_scene s = b200_detail::scene_view(const_cast<Scene *>(this->scene));
    itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)renderState);
    itm_b200_tracking_state t = b200_detail::tracking_state_view(trackingState);
    itm_b200_check(itm_b200_create_icp_maps(c->ctx, &s, &r, &t), "CreateICPMaps");
    trackingState->pose_pointCloud->SetFrom(trackingState->pose_d);  // ITMVisualisationEngine_CPU.cpp:273
  }

  void ForwardRender(const ITMView *view, ITMTrackingState *trackingState, ITMRenderState *renderState) const {
    itm_b200_scene s = b200_detail::scene_view(const_cast<Scene *>(this->scene));
    itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)renderState);
    const itm_b200_tracking_state t = b200_detail::tracking_state_view(trackingState);
    itm_b200_check(itm_b200_forward_render(c->ctx, &s, &r, view->depth->GetData(MEMORYDEVICE_CUDA), &t), "ForwardRender");
    renderState->noFwdProjMissingPoints = r.no_fwd_proj_missing_points;
  }

  void FindVisibleBlocks(const ITMPose *pose, const ITMIntrinsics *intrinsics, ITMRenderState *renderState) const {
    itm_b200_scene s = b200_detail::scene_view(const_cast<Scene *>(this->scene));
    ITMRenderState_VH *rsVH = (ITMRenderState_VH *)renderState;
    itm_b200_render_state r = b200_detail::render_state_view(rsVH);
    const Vector4f &k = intrinsics->projectionParamsSimple.all;
    const float intr[4] = {k.x, k.y, k.z, k.w};
    itm_b200_check(itm_b200_find_visible_blocks(c->ctx, &s, &r, pose->GetM().m, intr), "FindVisibleBlocks");
    rsVH->noVisibleEntries = r.no_visible_entries;
  }

  void RenderImage(const ITMPose *pose, const ITMIntrinsics *intrinsics, const ITMRenderState *renderState, ITMUChar4Image *outputImage,
                   IITMVisualisationEngine::RenderImageType type) const {
    itm_b200_scene s = b200_detail::scene_view(const_cast<Scene *>(this->scene));
    itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)const_cast<ITMRenderState *>(renderState));
    r.img_width = outputImage->noDims.x;  // RenderImage_common: imgSize = outputImage->noDims
    r.img_height = outputImage->noDims.y;
    const Vector4f &k = intrinsics->projectionParamsSimple.all;
    const float intr[4] = {k.x, k.y, k.z, k.w};
    itm_b200_check(itm_b200_render_image(c->ctx, &s, &r, pose->GetM().m, intr, (unsigned char *)outputImage->GetData(MEMORYDEVICE_CUDA), (int)type),
                   "RenderImage");
  }

  void FindSurface(const ITMPose *pose, const ITMIntrinsics *intrinsics, const ITMRenderState *renderState) const {
    itm_b200_scene s = b200_detail::scene_view(const_cast<Scene *>(this->scene));
    itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)const_cast<ITMRenderState *>(renderState));
    r.img_width = renderState->raycastResult->noDims.x;
    r.img_height = renderState->raycastResult->noDims.y;
    const Vector4f &k = intrinsics->projectionParamsSimple.all;
    const float intr[4] = {k.x, k.y, k.z, k.w};
    itm_b200_check(itm_b200_find_surface(c->ctx, &s, &r, pose->GetM().m, intr), "FindSurface");
  }

  // the colour tracker's model of the scene (ITMTrackingController.cpp:22-28 calls it for TRACKER_COLOR)
  void CreatePointCloud(const ITMView *view, ITMTrackingState *trackingState, ITMRenderState *renderState, bool skipPoints) const {
    itm_b200_scene s = b200_detail::scene_view(const_cast<Scene *>(this->scene));
    itm_b200_render_state r = b200_detail::render_state_view((ITMRenderState_VH *)renderState);
    r.img_width = renderState->raycastResult->noDims.x;  // CreatePointCloud_common: imgSize = raycastResult->noDims
    r.img_height = renderState->raycastResult->noDims.y;
    itm_b200_tracking_state t = b200_detail::tracking_state_view(trackingState);
    const Matrix4f invM = trackingState->pose_d->GetInvM() * view->calib->trafo_rgb_to_depth.calib;  // ITMVisualisationEngine_CPU.cpp:247
    const Vector4f &k = view->calib->intrinsics_rgb.projectionParamsSimple.all;
    const float intr[4] = {k.x, k.y, k.z, k.w};
    int noTotalPoints = 0;
    itm_b200_check(itm_b200_create_point_cloud(c->ctx, &s, &r, &t, invM.m, intr, skipPoints ? 1 : 0, &noTotalPoints), "CreatePointCloud");
    trackingState->pose_pointCloud->SetFrom(trackingState->pose_d);
    trackingState->pointCloud->noTotalPoints = noTotalPoints;
  }
};

// ---------------------------------------------------------------------------------------------
template <class TVoxel, class TIndex>
class ITMMeshingEngine_B200;

template <class TVoxel>
class ITMMeshingEngine_B200<TVoxel, ITMVoxelBlockHash> : public ITMMeshingEngine<TVoxel, ITMVoxelBlockHash> {
  ITMB200Context *c;

 public:
  explicit ITMMeshingEngine_B200(ITMB200Context *context) : c(context) {}

  /// mesh must have been created with MEMORYDEVICE_CUDA (ITMMainEngine.cpp:48 does so for DEVICE_CUDA)
  void MeshScene(ITMMesh *mesh, const ITMScene<TVoxel, ITMVoxelBlockHash> *scene) {
    itm_b200_scene s = b200_detail::scene_view(const_cast<ITMScene<TVoxel, ITMVoxelBlockHash> *>(scene));
    unsigned n = 0;
    itm_b200_check(itm_b200_mesh_scene(c->ctx, &s, (float *)mesh->triangles->GetData(MEMORYDEVICE_CUDA), ITMMesh::noMaxTriangles, &n), "MeshScene");
    mesh->noTotalTriangles = n;
  }
};

// ---------------------------------------------------------------------------------------------
class ITMLowLevelEngine_B200 : public ITMLowLevelEngine {
  ITMB200Context *c;

 public:
  explicit ITMLowLevelEngine_B200(ITMB200Context *context) : c(context) {}

  void FilterSubsampleWithHoles(ITMFloatImage *image_out, const ITMFloatImage *image_in) const {
    const Vector2i in = image_in->noDims;
    image_out->ChangeDims(Vector2i(in.x / 2, in.y / 2));  // ITMLowLevelEngine_CPU.cpp:52-54
    itm_b200_check(itm_b200_filter_subsample_with_holes(c->ctx, image_out->GetData(MEMORYDEVICE_CUDA), image_in->GetData(MEMORYDEVICE_CUDA), in.x, in.y),
                   "FilterSubsampleWithHoles");
  }

  // the helpers only the colour / Ren trackers call (ITMLowLevelEngine_CPU.cpp:12-108), images in HBM
  void CopyImage(ITMUChar4Image *image_out, const ITMUChar4Image *image_in) const {
    itm_b200_check(itm_b200_copy_image(c->ctx, image_out->GetData(MEMORYDEVICE_CUDA), image_in->GetData(MEMORYDEVICE_CUDA),
                                       image_in->dataSize * sizeof(Vector4u)), "CopyImage");
  }
  void CopyImage(ITMFloatImage *image_out, const ITMFloatImage *image_in) const {
    itm_b200_check(itm_b200_copy_image(c->ctx, image_out->GetData(MEMORYDEVICE_CUDA), image_in->GetData(MEMORYDEVICE_CUDA),
                                       image_in->dataSize * sizeof(float)), "CopyImage");
  }
  void CopyImage(ITMFloat4Image *image_out, const ITMFloat4Image *image_in) const {
    itm_b200_check(itm_b200_copy_image(c->ctx, image_out->GetData(MEMORYDEVICE_CUDA), image_in->GetData(MEMORYDEVICE_CUDA),
                                       image_in->dataSize * sizeof(Vector4f)), "CopyImage");
  }
  void FilterSubsample(ITMUChar4Image *image_out, const ITMUChar4Image *image_in) const {
    const Vector2i in = image_in->noDims;
    image_out->ChangeDims(Vector2i(in.x / 2, in.y / 2));
    itm_b200_check(itm_b200_filter_subsample_rgba(c->ctx, (unsigned char *)image_out->GetData(MEMORYDEVICE_CUDA),
                                                  (const unsigned char *)image_in->GetData(MEMORYDEVICE_CUDA), in.x, in.y), "FilterSubsample");
  }
  void FilterSubsampleWithHoles(ITMFloat4Image *image_out, const ITMFloat4Image *image_in) const {
    const Vector2i in = image_in->noDims;
    image_out->ChangeDims(Vector2i(in.x / 2, in.y / 2));
    itm_b200_check(itm_b200_filter_subsample_with_holes_float4(c->ctx, (float *)image_out->GetData(MEMORYDEVICE_CUDA),
                                                               (const float *)image_in->GetData(MEMORYDEVICE_CUDA), in.x, in.y),
                   "FilterSubsampleWithHoles");
  }
  void GradientX(ITMShort4Image *grad_out, const ITMUChar4Image *image_in) const {
    grad_out->ChangeDims(image_in->noDims);
    itm_b200_check(itm_b200_gradient_x(c->ctx, (short *)grad_out->GetData(MEMORYDEVICE_CUDA), (const unsigned char *)image_in->GetData(MEMORYDEVICE_CUDA),
                                       image_in->noDims.x, image_in->noDims.y), "GradientX");
  }
  void GradientY(ITMShort4Image *grad_out, const ITMUChar4Image *image_in) const {
    grad_out->ChangeDims(image_in->noDims);
    itm_b200_check(itm_b200_gradient_y(c->ctx, (short *)grad_out->GetData(MEMORYDEVICE_CUDA), (const unsigned char *)image_in->GetData(MEMORYDEVICE_CUDA),
                                       image_in->noDims.x, image_in->noDims.y), "GradientY");
  }
};

// ---------------------------------------------------------------------------------------------
class ITMViewBuilder_B200 : public ITMViewBuilder {
  ITMB200Context *c;

 public:
  ITMViewBuilder_B200(const ITMRGBDCalib *calib, ITMB200Context *context) : ITMViewBuilder(calib), c(context) {}

  void ConvertDepthAffineToFloat(ITMFloatImage *depth_out, const ITMShortImage *depth_in, Vector2f depthCalibParams) {
    itm_b200_check(itm_b200_convert_depth_affine_to_float(c->ctx, depth_out->GetData(MEMORYDEVICE_CUDA), depth_in->GetData(MEMORYDEVICE_CUDA),
                                                          depth_in->noDims.x, depth_in->noDims.y, depthCalibParams.x, depthCalibParams.y),
                   "ConvertDepthAffineToFloat");
  }

  void DepthFiltering(ITMFloatImage *image_out, const ITMFloatImage *image_in) {
    itm_b200_check(itm_b200_depth_filtering(c->ctx, image_out->GetData(MEMORYDEVICE_CUDA), image_in->GetData(MEMORYDEVICE_CUDA), image_in->noDims.x,
                                            image_in->noDims.y),
                   "DepthFiltering");
  }

  void ComputeNormalAndWeights(ITMFloat4Image *normal_out, ITMFloatImage *sigmaZ_out, const ITMFloatImage *depth_in, Vector4f intrinsic) {
    const float intr[4] = {intrinsic.x, intrinsic.y, intrinsic.z, intrinsic.w};
    itm_b200_check(itm_b200_compute_normal_and_weights(c->ctx, (float *)normal_out->GetData(MEMORYDEVICE_CUDA), sigmaZ_out->GetData(MEMORYDEVICE_CUDA),
                                                       depth_in->GetData(MEMORYDEVICE_CUDA), depth_in->noDims.x, depth_in->noDims.y, intr),
                   "ComputeNormalAndWeights");
  }

  // same sequence as ITMViewBuilder_CPU::UpdateView (ITMViewBuilder_CPU.cpp:14-64) with the images in HBM
  void UpdateView(ITMView **view_ptr, ITMUChar4Image *rgbImage, ITMShortImage *rawDepthImage, bool useBilateralFilter, bool modelSensorNoise = false) {
    if (*view_ptr == NULL) {
      *view_ptr = new ITMView(calib, rgbImage->noDims, rawDepthImage->noDims, true);
      if (this->shortImage != NULL) delete this->shortImage;
      this->shortImage = new ITMShortImage(rawDepthImage->noDims, true, true);
      if (this->floatImage != NULL) delete this->floatImage;
      this->floatImage = new ITMFloatImage(rawDepthImage->noDims, true, true);
      if (modelSensorNoise) {
        (*view_ptr)->depthNormal = new ITMFloat4Image(rawDepthImage->noDims, true, true);
        (*view_ptr)->depthUncertainty = new ITMFloatImage(rawDepthImage->noDims, true, true);
      }
    }
    ITMView *view = *view_ptr;
    view->rgb->SetFrom(rgbImage, ORUtils::MemoryBlock<Vector4u>::CPU_TO_CUDA);
    this->shortImage->SetFrom(rawDepthImage, ORUtils::MemoryBlock<short>::CPU_TO_CUDA);
    switch (view->calib->disparityCalib.type) {  // ITMViewBuilder_CPU.cpp:38-48
      case ITMDisparityCalib::TRAFO_KINECT:
        this->ConvertDisparityToDepth(view->depth, this->shortImage, &(view->calib->intrinsics_d), view->calib->disparityCalib.params);
        break;
      case ITMDisparityCalib::TRAFO_AFFINE:
        this->ConvertDepthAffineToFloat(view->depth, this->shortImage, view->calib->disparityCalib.params);
        break;
      default:
        break;
    }
    if (useBilateralFilter) {
      // 5 steps of bilateral filtering
      this->DepthFiltering(this->floatImage, view->depth);
      this->DepthFiltering(view->depth, this->floatImage);
      this->DepthFiltering(this->floatImage, view->depth);
      this->DepthFiltering(view->depth, this->floatImage);
      this->DepthFiltering(this->floatImage, view->depth);
      view->depth->SetFrom(this->floatImage, ORUtils::MemoryBlock<float>::CUDA_TO_CUDA);
    }
    if (modelSensorNoise)
      this->ComputeNormalAndWeights(view->depthNormal, view->depthUncertainty, view->depth, view->calib->intrinsics_d.projectionParamsSimple.all);
  }

  // ITMViewBuilder_CPU.cpp:78-90: fx_depth = depthIntrinsics->projectionParamsSimple.fx
  void ConvertDisparityToDepth(ITMFloatImage *depth_out, const ITMShortImage *depth_in, const ITMIntrinsics *depthIntrinsics, Vector2f disparityCalibParams) {
    itm_b200_check(itm_b200_convert_disparity_to_depth(c->ctx, depth_out->GetData(MEMORYDEVICE_CUDA), depth_in->GetData(MEMORYDEVICE_CUDA),
                                                       depth_in->noDims.x, depth_in->noDims.y, disparityCalibParams.x, disparityCalibParams.y,
                                                       depthIntrinsics->projectionParamsSimple.fx),
                   "ConvertDisparityToDepth");
  }
  // ITMViewBuilder_CPU.cpp:66-75: the caller has already filled the view's own host images; only the upload happens here
  void UpdateView(ITMView **view_ptr, ITMUChar4Image *rgbImage, ITMFloatImage *depthImage) {
    if (*view_ptr == NULL) *view_ptr = new ITMView(calib, rgbImage->noDims, depthImage->noDims, true);
    (*view_ptr)->rgb->UpdateDeviceFromHost();
    (*view_ptr)->depth->UpdateDeviceFromHost();
  }
  // ITMViewBuilder_CPU.cpp:77-92: a view that also carries the IMU measurement, then the plain update
  void UpdateView(ITMView **view_ptr, ITMUChar4Image *rgbImage, ITMShortImage *depthImage, bool useBilateralFilter, ITMIMUMeasurement *imuMeasurement) {
    if (*view_ptr == NULL) {
      *view_ptr = new ITMViewIMU(calib, rgbImage->noDims, depthImage->noDims, true);
      if (this->shortImage != NULL) delete this->shortImage;
      this->shortImage = new ITMShortImage(depthImage->noDims, true, true);
      if (this->floatImage != NULL) delete this->floatImage;
      this->floatImage = new ITMFloatImage(depthImage->noDims, true, true);
    }
    ((ITMViewIMU *)(*view_ptr))->imu->SetFrom(imuMeasurement);
    this->UpdateView(view_ptr, rgbImage, depthImage, useBilateralFilter);
  }
};

// ---------------------------------------------------------------------------------------------
/// ITMDepthTracker with (a) the whole Levenberg-Marquardt loop of TrackCamera on the device and
/// (b) ComputeGandH as a single-evaluation entry point, so the base class' own host loop
/// (ITMDepthTracker.cpp:145-199) also works on top of it (set useDeviceLoop = false).
class ITMDepthTracker_B200 : public ITMDepthTracker {
  ITMB200Context *c;

 public:
  bool useDeviceLoop;

  ITMDepthTracker_B200(Vector2i imgSize, TrackerIterationType *trackingRegime, int noHierarchyLevels, int noICPRunTillLevel, float distThresh,
                       float terminationThreshold, const ITMLowLevelEngine *lowLevelEngine, ITMB200Context *context)
      : ITMDepthTracker(imgSize, trackingRegime, noHierarchyLevels, noICPRunTillLevel, distThresh, terminationThreshold, lowLevelEngine, MEMORYDEVICE_CUDA),
        c(context), useDeviceLoop(true) {}

  void TrackCamera(ITMTrackingState *trackingState, const ITMView *view) {
    if (!useDeviceLoop) {
      ITMDepthTracker::TrackCamera(trackingState, view);
      return;
    }
    itm_b200_tracking_state t = b200_detail::tracking_state_view(trackingState);
    itm_b200_check(itm_b200_track_camera(c->ctx, view->depth->GetData(MEMORYDEVICE_CUDA), &t), "TrackCamera");
    Matrix4f M;
    for (int i = 0; i < 16; ++i) M.m[i] = t.pose_d[i];
    trackingState->pose_d->SetM(M);  // already coerced on the device (ITMDepthTracker.cpp:194-196)
  }

 protected:
  int ComputeGandH(float &f, float *nabla, float *hessian, Matrix4f approxInvPose) {
    const Vector4f &vk = viewHierarchyLevel->intrinsics, &sk = sceneHierarchyLevel->intrinsics;
    const float viewIntr[4] = {vk.x, vk.y, vk.z, vk.w}, sceneIntr[4] = {sk.x, sk.y, sk.z, sk.w};
    const Vector2i viewSize = viewHierarchyLevel->depth->noDims, sceneSize = sceneHierarchyLevel->pointsMap->noDims;
    int noValid = 0;
    float h36[36];
    itm_b200_check(itm_b200_compute_g_and_h(c->ctx, viewHierarchyLevel->depth->GetData(MEMORYDEVICE_CUDA), viewSize.x, viewSize.y, viewIntr,
                                            (const float *)sceneHierarchyLevel->pointsMap->GetData(MEMORYDEVICE_CUDA),
                                            (const float *)sceneHierarchyLevel->normalsMap->GetData(MEMORYDEVICE_CUDA), sceneSize.x, sceneSize.y,
                                            sceneIntr, approxInvPose.m, scenePose.m, distThresh[levelId], (int)iterationType, &f, nabla, h36,
                                            &noValid),
                   "ComputeGandH");
    // stride 6 for 3- and 6-parameter iterations alike (ITMDepthTracker_CPU.cpp:72-73)
    for (int i = 0; i < 36; ++i) hessian[i] = h36[i];
    return noValid;
  }
};

// ---------------------------------------------------------------------------------------------
/// ITMWeightedICPTracker (TRACKER_WICP): the reference's own Gauss-Newton loop, pyramid and weight pyramid
/// (ITMWeightedICPTracker.cpp:58-192, over ITMLowLevelEngine_B200::FilterSubsampleWithHoles) with every evaluation on the device.
class ITMWeightedICPTracker_B200 : public ITMWeightedICPTracker {
  ITMB200Context *c;

 public:
  ITMWeightedICPTracker_B200(Vector2i imgSize, TrackerIterationType *trackingRegime, int noHierarchyLevels, int noICPRunTillLevel, float distThresh,
                             float terminationThreshold, const ITMLowLevelEngine *lowLevelEngine, ITMB200Context *context)
      : ITMWeightedICPTracker(imgSize, trackingRegime, noHierarchyLevels, noICPRunTillLevel, distThresh, terminationThreshold, lowLevelEngine,
                              MEMORYDEVICE_CUDA),
        c(context) {}

  /// true (default): the whole Gauss-Newton loop runs on the device (itm_b200_track_camera_weighted, no host round trip per
  /// evaluation); false: the reference's own host loop (ITMWeightedICPTracker.cpp:164-192) over ComputeGandH below
  bool useDeviceLoop = true;

  void TrackCamera(ITMTrackingState *trackingState, const ITMView *view) {
    if (!useDeviceLoop) {
      ITMWeightedICPTracker::TrackCamera(trackingState, view);
      return;
    }
    itm_b200_tracking_state t = b200_detail::tracking_state_view(trackingState);
    itm_b200_check(itm_b200_track_camera_weighted(c->ctx, view->depth->GetData(MEMORYDEVICE_CUDA), view->depthUncertainty->GetData(MEMORYDEVICE_CUDA), &t),
                   "TrackCamera (weighted)");
    Matrix4f M;
    for (int i = 0; i < 16; ++i) M.m[i] = t.pose_d[i];
    trackingState->pose_d->SetM(M);
  }

 protected:
  int ComputeGandH(float &f, float *nabla, float *hessian, Matrix4f approxInvPose) {
    const Vector4f &vk = viewHierarchyLevel->intrinsics, &sk = sceneHierarchyLevel->intrinsics;
    const float viewIntr[4] = {vk.x, vk.y, vk.z, vk.w}, sceneIntr[4] = {sk.x, sk.y, sk.z, sk.w};
    const Vector2i viewSize = viewHierarchyLevel->depth->noDims, sceneSize = sceneHierarchyLevel->pointsMap->noDims;
    int noValid = 0;
    float h36[36];
    itm_b200_check(itm_b200_compute_g_and_h_weighted(c->ctx, viewHierarchyLevel->depth->GetData(MEMORYDEVICE_CUDA),
                                                     weightHierarchyLevel->depth->GetData(MEMORYDEVICE_CUDA), viewSize.x, viewSize.y, viewIntr,
                                                     (const float *)sceneHierarchyLevel->pointsMap->GetData(MEMORYDEVICE_CUDA),
                                                     (const float *)sceneHierarchyLevel->normalsMap->GetData(MEMORYDEVICE_CUDA), sceneSize.x,
                                                     sceneSize.y, sceneIntr, approxInvPose.m, scenePose.m, distThresh[levelId], (int)iterationType,
                                                     &f, nabla, h36, &noValid),
                   "ComputeGandH (weighted)");
    for (int i = 0; i < 36; ++i) hessian[i] = h36[i];
    return noValid;
  }
};

}  // namespace Engine
}  // namespace ITMLib
