// oracle/adapter_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drop-in proof: the reference's OWN host objects (ITMScene, ITMRenderState_VH, ITMTrackingState,
// ITMView, ITMTrackingController, ITMDepthTracker's host LM loop, ITMPose - compiled from
// /root/reference/InfiniTAM where they lie, with CUDA memory exactly as the reference's
// DEVICE_CUDA branch allocates it) driven through the adapter classes of
// include/itm_b200_adapter.hpp, i.e. through the C ABI of libitm_b200.so.  The composition mirrors
// ITMMainEngine's constructor and ProcessFrame (ITMLib/Engine/ITMMainEngine.cpp:17-68, 111-127)
// and ITMDenseMapper::ProcessFrame (ITMLib/Engine/ITMDenseMapper.cpp:51-65); the only change a
// maintainer makes is which engine classes are constructed.
//
// Built by oracle/build_ref.py into oracle/_ref/libitm_adapter.so; loaded only by tests/.
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

// single evaluations of the weighted tracker are driven from outside (SetEvaluationData & co. are private): open the
// classes up, the layout is unchanged
#define private public
#define protected public
#include "ITMLib/Engine/ITMTrackingController.h"
#include "itm_b200_adapter.hpp"
#undef private
#undef protected

using namespace ITMLib::Engine;
using namespace ITMLib::Objects;

typedef ITMVoxel TV;
typedef ITMVoxelIndex TI;

struct adp_engine {
  ITMLibSettings *settings;
  ITMRGBDCalib calib;
  ITMB200Context *ctx;
  ITMScene<TV, TI> *scene;
  ITMLowLevelEngine_B200 *lowLevel;
  ITMViewBuilder_B200 *viewBuilder;
  ITMVisualisationEngine_B200<TV, TI> *vis;
  ITMSceneReconstructionEngine_B200<TV, TI> *reco;
  ITMSwappingEngine_B200<TV, TI> *swapper;  // NULL unless created after adp_set_use_swapping(1)
  ITMDepthTracker_B200 *tracker;
  ITMWeightedICPTracker_B200 *wtracker;  // instead of tracker when created after adp_set_tracker_wicp(1, ..)
  ITMTrackingController *controller;
  ITMTrackingState *trackingState;
  ITMRenderState *renderState;
  ITMRenderState *renderStateFree;
  ITMUChar4Image *freeOut;
  ITMMesh *mesh;
  ITMMeshingEngine_B200<TV, TI> *meshing;
  ITMView *view;
  ITMUChar4Image *rgb;
  ITMShortImage *rawDepth;
  Vector2i imgSize;
};

static std::string g_err;
static int g_nextUseSwapping = 0;
static int g_nextWicp = 0, g_nextBilateral = 0;

extern "C" {

const char *adp_last_error() { return g_err.c_str(); }
// settings.useSwapping of the engines created from now on (ITMDenseMapper.cpp:16-34, 59-64)
void adp_set_use_swapping(int on) { g_nextUseSwapping = on; }
// TRACKER_WICP (+ settings.modelSensorNoise) and settings.useBilateralFilter of the engines created from now on
void adp_set_tracker_wicp(int on, int bilateral) { g_nextWicp = on; g_nextBilateral = bilateral; }

adp_engine *adp_create(int W, int H, float fx, float fy, float cx, float cy, float voxelSize, float mu, int maxW, float vfMin, float vfMax,
                       int deviceLoop) {
  try {
    adp_engine *e = new adp_engine();
    e->settings = new ITMLibSettings();
    e->settings->deviceType = ITMLibSettings::DEVICE_CUDA;  // memory placement of every reference object below
    e->settings->trackerType = ITMLibSettings::TRACKER_ICP;
    e->settings->useSwapping = g_nextUseSwapping != 0;
    e->settings->useApproximateRaycast = false;
    e->settings->useBilateralFilter = g_nextBilateral != 0;
    e->settings->modelSensorNoise = g_nextWicp != 0;
    if (g_nextWicp) e->settings->trackerType = ITMLibSettings::TRACKER_WICP;
    e->settings->sceneParams.voxelSize = voxelSize;
    e->settings->sceneParams.mu = mu;
    e->settings->sceneParams.maxW = maxW;
    e->settings->sceneParams.viewFrustum_min = vfMin;
    e->settings->sceneParams.viewFrustum_max = vfMax;
    e->imgSize = Vector2i(W, H);
    e->calib.intrinsics_d.SetFrom(fx, fy, cx, cy, (float)W, (float)H);
    e->calib.intrinsics_rgb.SetFrom(fx, fy, cx, cy, (float)W, (float)H);
    e->calib.disparityCalib.SetFrom(1.0f / 1000.0f, 0.0f, ITMDisparityCalib::TRAFO_AFFINE);

    // ITMMainEngine::ITMMainEngine (ITMMainEngine.cpp:17-68) with the B200 engine set
    e->ctx = new ITMB200Context(e->settings, &e->calib, e->imgSize);
    e->scene = new ITMScene<TV, TI>(&e->settings->sceneParams, e->settings->useSwapping, MEMORYDEVICE_CUDA);
    e->swapper = e->settings->useSwapping ? new ITMSwappingEngine_B200<TV, TI>(e->ctx) : NULL;
    e->lowLevel = new ITMLowLevelEngine_B200(e->ctx);
    e->viewBuilder = new ITMViewBuilder_B200(&e->calib, e->ctx);
    e->vis = new ITMVisualisationEngine_B200<TV, TI>(e->scene, e->ctx);
    e->reco = new ITMSceneReconstructionEngine_B200<TV, TI>(e->ctx);
    e->renderState = e->vis->CreateRenderState(e->imgSize);
    e->reco->ResetScene(e->scene);
    e->tracker = NULL;
    e->wtracker = NULL;
    if (g_nextWicp) {
      e->wtracker = new ITMWeightedICPTracker_B200(e->imgSize, e->settings->trackingRegime, e->settings->noHierarchyLevels,
                                                   e->settings->noICPRunTillLevel, e->settings->depthTrackerICPThreshold,
                                                   e->settings->depthTrackerTerminationThreshold, e->lowLevel, e->ctx);
    } else {
      e->tracker = new ITMDepthTracker_B200(e->imgSize, e->settings->trackingRegime, e->settings->noHierarchyLevels, e->settings->noICPRunTillLevel,
                                            e->settings->depthTrackerICPThreshold, e->settings->depthTrackerTerminationThreshold, e->lowLevel, e->ctx);
      e->tracker->useDeviceLoop = deviceLoop != 0;
    }
    ITMTracker *anyTracker = e->wtracker ? (ITMTracker *)e->wtracker : (ITMTracker *)e->tracker;
    e->controller = new ITMTrackingController(anyTracker, e->vis, e->lowLevel, e->settings);
    e->trackingState = e->controller->BuildTrackingState(e->imgSize);
    anyTracker->UpdateInitialPose(e->trackingState);
    e->view = NULL;
    e->renderStateFree = NULL;
    e->freeOut = NULL;
    e->mesh = NULL;
    e->meshing = new ITMMeshingEngine_B200<TV, TI>(e->ctx);
    e->rgb = new ITMUChar4Image(e->imgSize, true, false);
    e->rawDepth = new ITMShortImage(e->imgSize, true, false);
    memset(e->rgb->GetData(MEMORYDEVICE_CPU), 128, (size_t)W * H * 4);
    return e;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return NULL;
  }
}

void adp_destroy(adp_engine *e) {
  if (!e) return;
  if (e->swapper) delete e->swapper;
  delete e->renderState;
  if (e->renderStateFree) delete e->renderStateFree;
  if (e->freeOut) delete e->freeOut;
  if (e->mesh) delete e->mesh;
  delete e->meshing;
  delete e->scene;
  delete e->controller;
  if (e->tracker) delete e->tracker;
  if (e->wtracker) delete e->wtracker;
  delete e->lowLevel;
  delete e->viewBuilder;
  delete e->trackingState;
  if (e->view) delete e->view;
  delete e->vis;
  delete e->reco;
  delete e->rgb;
  delete e->rawDepth;
  delete e->ctx;
  delete e->settings;
  delete e;
}

// ITMMainEngine::ProcessFrame (ITMMainEngine.cpp:111-127)
int adp_process_frame(adp_engine *e, const short *depth) {
  try {
    memcpy(e->rawDepth->GetData(MEMORYDEVICE_CPU), depth, (size_t)e->imgSize.x * e->imgSize.y * sizeof(short));
    e->viewBuilder->UpdateView(&e->view, e->rgb, e->rawDepth, e->settings->useBilateralFilter, e->settings->modelSensorNoise);
    e->controller->Track(e->trackingState, e->view);
    // ITMDenseMapper::ProcessFrame (ITMDenseMapper.cpp:51-65)
    e->reco->AllocateSceneFromDepth(e->scene, e->view, e->trackingState, e->renderState);
    e->reco->IntegrateIntoScene(e->scene, e->view, e->trackingState, e->renderState);
    if (e->swapper) {
      e->swapper->IntegrateGlobalIntoLocal(e->scene, e->renderState);
      e->swapper->SaveToGlobalMemory(e->scene, e->renderState);
    }
    e->controller->Prepare(e->trackingState, e->view, e->renderState);
    return 0;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// settings.useApproximateRaycast: ITMTrackingController::Track / Prepare then alternate CreateICPMaps and ForwardRender
void adp_set_use_approximate_raycast(adp_engine *e, int on) { e->settings->useApproximateRaycast = on != 0; }
int adp_requires_full_rendering(adp_engine *e) { return e->trackingState->requiresFullRendering ? 1 : 0; }

// the free-view branch of ITMMainEngine::GetImage (ITMMainEngine.cpp:167-186) through the adapter's visualisation engine
int adp_get_free_image(adp_engine *e, int renderType, const float *poseM16, const float *intr4, int w, int h, unsigned char *out) {
  try {
    if (e->freeOut == NULL || e->freeOut->noDims.x != w || e->freeOut->noDims.y != h) {
      if (e->freeOut) delete e->freeOut;
      if (e->renderStateFree) delete e->renderStateFree;
      e->freeOut = new ITMUChar4Image(Vector2i(w, h), true, false);
      e->renderStateFree = e->vis->CreateRenderState(e->freeOut->noDims);
    }
    Matrix4f M(poseM16);
    ITMPose pose; pose.SetM(M);
    ITMIntrinsics intr; intr.SetFrom(intr4[0], intr4[1], intr4[2], intr4[3], (float)w, (float)h);
    e->vis->FindVisibleBlocks(&pose, &intr, e->renderStateFree);
    e->vis->CreateExpectedDepths(&pose, &intr, e->renderStateFree);
    e->vis->RenderImage(&pose, &intr, e->renderStateFree, e->renderStateFree->raycastImage, (IITMVisualisationEngine::RenderImageType)renderType);
    e->freeOut->SetFrom(e->renderStateFree->raycastImage, ORUtils::MemoryBlock<Vector4u>::CUDA_TO_CPU);
    memcpy(out, e->freeOut->GetData(MEMORYDEVICE_CPU), (size_t)w * h * 4);
    return 0;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// the TRACKER_COLOR branch of ITMTrackingController::Prepare (ITMTrackingController.cpp:22-28) through the adapter's
// visualisation engine; locations / colours come back with adp_read(4 / 5).  Returns noTotalPoints.
int adp_create_point_cloud(adp_engine *e, const float *trafo16, int skipPoints) {
  try {
    if (e->view == NULL) return -1;
    if (trafo16) { Matrix4f T(trafo16); e->calib.trafo_rgb_to_depth.SetFrom(T); e->view->calib->trafo_rgb_to_depth.SetFrom(T); }  // the view holds a copy
    ITMPose pose_rgb(e->view->calib->trafo_rgb_to_depth.calib_inv * e->trackingState->pose_d->GetM());
    e->vis->CreateExpectedDepths(&pose_rgb, &(e->view->calib->intrinsics_rgb), e->renderState);
    e->vis->CreatePointCloud(e->view, e->trackingState, e->renderState, skipPoints != 0);
    return e->trackingState->pointCloud->noTotalPoints;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// the same helpers through ITMLowLevelEngine_B200 (images in HBM); see ref_low_level in ref_harness.cpp for the contract
long long adp_low_level(adp_engine *e, int op, const void *in, int w, int h, void *out, int prefillByte) {
  try {
    const Vector2i dims(w, h), half(w / 2, h / 2);
    const size_t P = (size_t)w * h, Q = (size_t)half.x * half.y;
    if (op == 2) {
      ITMFloat4Image src(dims, true, true), dst(half, true, true);
      memcpy(src.GetData(MEMORYDEVICE_CPU), in, P * 16);
      src.UpdateDeviceFromHost();
      dst.Clear((unsigned char)prefillByte);
      e->lowLevel->FilterSubsampleWithHoles(&dst, &src);
      dst.UpdateHostFromDevice();
      memcpy(out, dst.GetData(MEMORYDEVICE_CPU), Q * 16);
      return (long long)(Q * 16);
    }
    ITMUChar4Image src(dims, true, true);
    memcpy(src.GetData(MEMORYDEVICE_CPU), in, P * 4);
    src.UpdateDeviceFromHost();
    if (op == 0 || op == 1) {
      ITMUChar4Image dst(op == 0 ? dims : half, true, true);
      dst.Clear((unsigned char)prefillByte);
      if (op == 0) e->lowLevel->CopyImage(&dst, &src); else e->lowLevel->FilterSubsample(&dst, &src);
      dst.UpdateHostFromDevice();
      const size_t bytes = (op == 0 ? P : Q) * 4;
      memcpy(out, dst.GetData(MEMORYDEVICE_CPU), bytes);
      return (long long)bytes;
    }
    ITMShort4Image grad(dims, true, true);
    grad.Clear((unsigned char)prefillByte);
    if (op == 3) e->lowLevel->GradientX(&grad, &src); else e->lowLevel->GradientY(&grad, &src);
    grad.UpdateHostFromDevice();
    memcpy(out, grad.GetData(MEMORYDEVICE_CPU), P * 8);
    return (long long)(P * 8);
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// ITMViewBuilder::UpdateView(view, rgb, float depth): the view's own host images are filled first, the call uploads them;
// returns the device depth back in depthOut.  Then the IMU variant on a fresh view: returns 0 if the measurement arrived.
int adp_update_view_variants(adp_engine *e, const float *depth, float *depthOut, const short *rawDepth, float *imuDepthOut) {
  try {
    const size_t P = (size_t)e->imgSize.x * e->imgSize.y;
    ITMFloatImage fimg(e->imgSize, true, false);
    ITMView *v = NULL;
    ITMViewBuilder_B200 vb(&e->calib, e->ctx);
    vb.UpdateView(&v, e->rgb, &fimg);  // creates the view
    memcpy(v->depth->GetData(MEMORYDEVICE_CPU), depth, P * 4);
    vb.UpdateView(&v, e->rgb, &fimg);
    if (cudaMemcpy(depthOut, v->depth->GetData(MEMORYDEVICE_CUDA), P * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    delete v;
    ITMView *vi = NULL;
    ITMIMUMeasurement imu;
    imu.R.setIdentity();
    imu.R.m[1] = 0.25f;
    memcpy(e->rawDepth->GetData(MEMORYDEVICE_CPU), rawDepth, P * sizeof(short));
    vb.UpdateView(&vi, e->rgb, e->rawDepth, false, &imu);
    const bool ok = ((ITMViewIMU *)vi)->imu->R.m[1] == 0.25f;
    if (cudaMemcpy(imuDepthOut, vi->depth->GetData(MEMORYDEVICE_CUDA), P * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    delete vi;
    return ok ? 0 : 1;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// ITMMainEngine::SaveSceneToMesh (ITMMainEngine.cpp:103-109): MeshScene into a CUDA ITMMesh, then the reference's own WriteSTL
int adp_save_scene_to_mesh(adp_engine *e, const char *fileName) {
  try {
    if (!e->mesh) e->mesh = new ITMMesh(MEMORYDEVICE_CUDA);
    e->meshing->MeshScene(e->mesh, e->scene);
    e->mesh->WriteSTL(fileName);
    return (int)e->mesh->noTotalTriangles;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// blocks parked in the host-side global cache (ITMGlobalCache::HasStoredData)
int adp_count_stored(adp_engine *e) {
  if (!e->swapper) return -1;
  int n = 0;
  for (int i = 0; i < e->scene->globalCache->noTotalEntries; ++i) n += e->scene->globalCache->HasStoredData(i) ? 1 : 0;
  return n;
}

// ITMViewBuilder::UpdateView alone (with the engine's filter settings)
int adp_update_view(adp_engine *e, const short *depth) {
  try {
    memcpy(e->rawDepth->GetData(MEMORYDEVICE_CPU), depth, (size_t)e->imgSize.x * e->imgSize.y * sizeof(short));
    e->viewBuilder->UpdateView(&e->view, e->rgb, e->rawDepth, e->settings->useBilateralFilter, e->settings->modelSensorNoise);
    return 0;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

// one ITMWeightedICPTracker::ComputeGandH at approxInvPose on pyramid level `level` of the current view;
// out: [0] = noValidPoints, [1] = f, [2..7] = nabla, [8..43] = hessian
int adp_wicp_gandh(adp_engine *e, int level, const float *approxInvPose, float *out) {
  try {
    e->wtracker->SetEvaluationData(e->trackingState, e->view);
    e->wtracker->PrepareForEvaluation();
    e->wtracker->SetEvaluationParams(level);
    Matrix4f inv(approxInvPose);
    float f = 0.f, nabla[6] = {0, 0, 0, 0, 0, 0}, hess[36];
    for (int i = 0; i < 36; ++i) hess[i] = 0.f;
    const int n = e->wtracker->ComputeGandH(f, nabla, hess, inv);
    out[0] = (float)n; out[1] = f;
    for (int i = 0; i < 6; ++i) out[2 + i] = nabla[i];
    for (int i = 0; i < 36; ++i) out[8 + i] = hess[i];
    return n;
  } catch (std::exception &ex) {
    g_err = ex.what();
    return -1;
  }
}

void adp_get_pose(adp_engine *e, float *M16) { memcpy(M16, e->trackingState->pose_d->GetM().m, 64); }

// counters = {noVisibleEntries, lastFreeBlockId, lastFreeExcessListId, age_pointCloud}
void adp_counters(adp_engine *e, int *c4) {
  c4[0] = ((ITMRenderState_VH *)e->renderState)->noVisibleEntries;
  c4[1] = e->scene->localVBA.lastFreeBlockId;
  c4[2] = e->scene->index.GetLastFreeExcessListId();
  c4[3] = e->trackingState->age_pointCloud;
}

// which: 0 hash entries, 1 voxel blocks, 2 visible ids, 3 raycast result, 4 points map, 5 normals map, 6 entriesVisibleType,
// 7 raycastImage, 8 view->depth, 9 view->depthUncertainty, 10 view->depthNormal
long long adp_read(adp_engine *e, int which, void *dst, long long capacity) {
  const size_t P = (size_t)e->imgSize.x * e->imgSize.y;
  const void *src = NULL;
  size_t bytes = 0;
  ITMRenderState_VH *rs = (ITMRenderState_VH *)e->renderState;
  switch (which) {
    case 0: src = e->scene->index.GetEntries(); bytes = (size_t)ITMVoxelBlockHash::noTotalEntries * sizeof(ITMHashEntry); break;
    case 1: src = e->scene->localVBA.GetVoxelBlocks(); bytes = (size_t)e->scene->localVBA.allocatedSize * sizeof(TV); break;
    case 2: src = rs->GetVisibleEntryIDs(); bytes = (size_t)SDF_LOCAL_BLOCK_NUM * sizeof(int); break;
    case 3: src = rs->raycastResult->GetData(MEMORYDEVICE_CUDA); bytes = P * 16; break;
    case 4: src = e->trackingState->pointCloud->locations->GetData(MEMORYDEVICE_CUDA); bytes = P * 16; break;
    case 5: src = e->trackingState->pointCloud->colours->GetData(MEMORYDEVICE_CUDA); bytes = P * 16; break;
    case 6: src = rs->GetEntriesVisibleType(); bytes = (size_t)ITMVoxelBlockHash::noTotalEntries; break;
    case 7: src = rs->raycastImage->GetData(MEMORYDEVICE_CUDA); bytes = P * 4; break;
    case 8: src = e->view->depth->GetData(MEMORYDEVICE_CUDA); bytes = P * 4; break;
    case 9: if (!e->view->depthUncertainty) return -1; src = e->view->depthUncertainty->GetData(MEMORYDEVICE_CUDA); bytes = P * 4; break;
    case 10: if (!e->view->depthNormal) return -1; src = e->view->depthNormal->GetData(MEMORYDEVICE_CUDA); bytes = P * 16; break;
    default: return -1;
  }
  if ((long long)bytes > capacity) return -(long long)bytes;
  if (cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (long long)bytes;
}

}  // extern "C"
