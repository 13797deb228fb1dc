// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C API around the UNMODIFIED reference CPU engines (compiled from
// /root/reference/InfiniTAM where they lie, see oracle/build_ref.py; output
// oracle/_ref/libitm_ref*.so).  It composes the public engine classes exactly
// the way ITMMainEngine's constructor and ProcessFrame do
// (ITMLib/Engine/ITMMainEngine.cpp:17-68, 111-127), but keeps every stage
// callable on its own and every piece of cross-frame state reachable as a raw
// pointer, so tests can run "teacher forced" comparisons stage by stage.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <map>
#include <string>
#include <chrono>
#include <iostream>
#include <stdexcept>

// ITMDepthTracker keeps SetEvaluationData / PrepareForEvaluation /
// SetEvaluationParams private and ComputeGandH protected
// (ITMLib/Engine/ITMDepthTracker.h:25-56).  The harness needs to drive single
// evaluations, so it opens the class up; layout is unchanged.
#define private public
#define protected public
#include "ITMLib/ITMLib.h"
#undef private
#undef protected

using namespace ITMLib::Engine;
using namespace ITMLib::Objects;

typedef ITMVoxelIndex TI;
#ifdef REF_VOXEL_RGB
// Colour flavour (BASELINE configs[4]).  The reference fixes its voxel type with a typedef (ITMLib/Utils/ITMLibDefines.h:205)
// and instantiates its engine templates for that type only, at the end of each .cpp.  Instead of patching a copy of the
// header, the two voxel-dependent engine sources are compiled HERE, where they lie, and instantiated for ITMVoxel_s_rgb.
typedef ITMVoxel_s_rgb TV;
#include "ITMLib/Engine/DeviceSpecific/CPU/ITMSceneReconstructionEngine_CPU.cpp"
#include "ITMLib/Engine/DeviceSpecific/CPU/ITMVisualisationEngine_CPU.cpp"
#include "ITMLib/Engine/DeviceSpecific/CPU/ITMSwappingEngine_CPU.cpp"
template class ITMLib::Engine::ITMSceneReconstructionEngine_CPU<ITMVoxel_s_rgb, ITMVoxelBlockHash>;
template class ITMLib::Engine::ITMVisualisationEngine_CPU<ITMVoxel_s_rgb, ITMVoxelBlockHash>;
template class ITMLib::Engine::ITMSwappingEngine_CPU<ITMVoxel_s_rgb, ITMVoxelBlockHash>;
#include "ITMLib/Engine/DeviceSpecific/CPU/ITMMeshingEngine_CPU.cpp"
template class ITMLib::Engine::ITMMeshingEngine_CPU<ITMVoxel_s_rgb, ITMVoxelBlockHash>;
#else
typedef ITMVoxel TV;
#endif

struct ref_engine {
  ITMLibSettings *settings;
  ITMRGBDCalib calib;
  ITMScene<TV, TI> *scene;
  ITMLowLevelEngine_CPU *lowLevel;
  ITMViewBuilder_CPU *viewBuilder;
  ITMVisualisationEngine_CPU<TV, TI> *vis;
  ITMSceneReconstructionEngine_CPU<TV, TI> *reco;
  ITMSwappingEngine_CPU<TV, TI> *swapper;  // NULL unless created after ref_set_use_swapping(1)
  ITMDepthTracker_CPU *tracker;
  ITMWeightedICPTracker_CPU *wtracker;  // instead of tracker when created after ref_set_tracker_wicp(1, ..)
  ITMTrackingController *controller;
  ITMTrackingState *trackingState;
  ITMRenderState *renderState;
  ITMRenderState *renderStateFree;  // ITMMainEngine::renderState_freeview
  ITMUChar4Image *freeOut;
  ITMMesh *mesh;
  ITMMeshingEngine_CPU<TV, TI> *meshing;
  ITMView *view;
  ITMUChar4Image *rgb;
  ITMShortImage *rawDepth;
  Vector2i imgSize;
};

static double now_ms() {
  return std::chrono::duration<double, std::milli>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int g_nextUseSwapping = 0;
static int g_nextWicp = 0, g_nextBilateral = 0;

extern "C" {

// settings.useSwapping of the engines created from now on (ITMLibSettings.cpp:35; ITMDenseMapper.cpp:20,59-64)
void ref_set_use_swapping(int on) { g_nextUseSwapping = on; }
// engines created from now on use TRACKER_WICP (ITMWeightedICPTracker_CPU + settings.modelSensorNoise, what
// ITMLibSettings.cpp:52-57 selects for that tracker type) and, optionally, settings.useBilateralFilter
void ref_set_tracker_wicp(int on, int bilateral) { g_nextWicp = on; g_nextBilateral = bilateral; }

int ref_const(const char *name) {
  std::string n(name);
  if (n == "SDF_BLOCK_SIZE") return SDF_BLOCK_SIZE;
  if (n == "SDF_LOCAL_BLOCK_NUM") return SDF_LOCAL_BLOCK_NUM;
  if (n == "SDF_BUCKET_NUM") return SDF_BUCKET_NUM;
  if (n == "SDF_EXCESS_LIST_SIZE") return SDF_EXCESS_LIST_SIZE;
  if (n == "SDF_HASH_MASK") return SDF_HASH_MASK;
  if (n == "sizeof_voxel") return (int)sizeof(TV);
  if (n == "sizeof_hash_entry") return (int)sizeof(ITMHashEntry);
  if (n == "has_color") return TV::hasColorInformation ? 1 : 0;
#ifdef WITH_OPENMP
  if (n == "openmp") return 1;
#else
  if (n == "openmp") return 0;
#endif
  return -1;
}

ref_engine *ref_create(int W, int H, float fx, float fy, float cx, float cy,
                       float voxelSize, float mu, int maxW, float vfMin, float vfMax) {
  ref_engine *e = new ref_engine();
  e->settings = new ITMLibSettings();
  // fork defaults differ from upstream (ITMLib/Utils/ITMLibSettings.cpp:44)
  e->settings->deviceType = ITMLibSettings::DEVICE_CPU;
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

  e->scene = new ITMScene<TV, TI>(&e->settings->sceneParams, e->settings->useSwapping, MEMORYDEVICE_CPU);
  e->swapper = e->settings->useSwapping ? new ITMSwappingEngine_CPU<TV, TI>() : NULL;
  e->lowLevel = new ITMLowLevelEngine_CPU();
  e->viewBuilder = new ITMViewBuilder_CPU(&e->calib);
  e->vis = new ITMVisualisationEngine_CPU<TV, TI>(e->scene);
  e->reco = new ITMSceneReconstructionEngine_CPU<TV, TI>();
  e->renderState = e->vis->CreateRenderState(e->imgSize);
  e->reco->ResetScene(e->scene);
  e->tracker = NULL;
  e->wtracker = NULL;
  if (g_nextWicp)
    e->wtracker = new ITMWeightedICPTracker_CPU(
        e->imgSize, e->settings->trackingRegime, e->settings->noHierarchyLevels,
        e->settings->noICPRunTillLevel, e->settings->depthTrackerICPThreshold,
        e->settings->depthTrackerTerminationThreshold, e->lowLevel);
  else
    e->tracker = new ITMDepthTracker_CPU(
        e->imgSize, e->settings->trackingRegime, e->settings->noHierarchyLevels,
        e->settings->noICPRunTillLevel, e->settings->depthTrackerICPThreshold,
        e->settings->depthTrackerTerminationThreshold, e->lowLevel);
  ITMTracker *anyTracker = e->wtracker ? (ITMTracker *)e->wtracker : (ITMTracker *)e->tracker;
  e->controller = new ITMTrackingController(anyTracker, e->vis, e->lowLevel, e->settings);
  e->trackingState = e->controller->BuildTrackingState(e->imgSize);
  anyTracker->UpdateInitialPose(e->trackingState);
  e->view = NULL;
  e->renderStateFree = NULL;
  e->freeOut = NULL;
  e->mesh = NULL;
  e->meshing = NULL;
  e->rgb = new ITMUChar4Image(e->imgSize, true, false);
  e->rawDepth = new ITMShortImage(e->imgSize, true, false);
  memset(e->rgb->GetData(MEMORYDEVICE_CPU), 128, (size_t)W * H * 4);
  return e;
}

void ref_destroy(ref_engine *e) {
  if (e->swapper) delete e->swapper;
  delete e->renderState;
  if (e->renderStateFree) delete e->renderStateFree;
  if (e->freeOut) delete e->freeOut;
  if (e->mesh) delete e->mesh;
  if (e->meshing) delete e->meshing;
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
  delete e->settings;
  delete e;
}

// ---- stages, in ProcessFrame order (ITMMainEngine.cpp:111-127) -------------
// colour image of the next frames (Vector4u[w*h]); the default is constant grey
void ref_set_rgb(ref_engine *e, const unsigned char *rgba) {
  memcpy(e->rgb->GetData(MEMORYDEVICE_CPU), rgba, (size_t)e->imgSize.x * e->imgSize.y * 4);
}

void ref_update_view(ref_engine *e, const short *depth) {
  memcpy(e->rawDepth->GetData(MEMORYDEVICE_CPU), depth, (size_t)e->imgSize.x * e->imgSize.y * sizeof(short));
  e->viewBuilder->UpdateView(&e->view, e->rgb, e->rawDepth, e->settings->useBilateralFilter, e->settings->modelSensorNoise);
}
void ref_track(ref_engine *e) { e->controller->Track(e->trackingState, e->view); }
void ref_allocate(ref_engine *e, int onlyVisible) {
  e->reco->AllocateSceneFromDepth(e->scene, e->view, e->trackingState, e->renderState, onlyVisible != 0);
}
void ref_integrate(ref_engine *e) {
  e->reco->IntegrateIntoScene(e->scene, e->view, e->trackingState, e->renderState);
}
// ITMDenseMapper::ProcessFrame's swapping part (ITMDenseMapper.cpp:59-64)
void ref_swap(ref_engine *e) {
  if (!e->swapper) return;
  e->swapper->IntegrateGlobalIntoLocal(e->scene, e->renderState);
  e->swapper->SaveToGlobalMemory(e->scene, e->renderState);
}
unsigned char *ref_swap_states(ref_engine *e) { return e->swapper ? (unsigned char *)e->scene->globalCache->GetSwapStates(false) : NULL; }
unsigned char *ref_has_stored_data(ref_engine *e) {
  static_assert(sizeof(bool) == 1, "bool");
  return e->swapper ? (unsigned char *)e->scene->globalCache->hasStoredData : NULL;  // private member, opened up above
}
void *ref_stored_voxel_blocks(ref_engine *e) { return e->swapper ? (void *)e->scene->globalCache->GetStoredVoxelBlock(0) : NULL; }
void ref_expected_depths(ref_engine *e) {
  e->vis->CreateExpectedDepths(e->trackingState->pose_d, &(e->view->calib->intrinsics_d), e->renderState);
}
void ref_icp_maps(ref_engine *e) {
  // what ITMTrackingController::Prepare does after CreateExpectedDepths
  // (ITMTrackingController.cpp:33-39)
  e->vis->CreateICPMaps(e->view, e->trackingState, e->renderState);
  e->trackingState->pose_pointCloud->SetFrom(e->trackingState->pose_d);
  if (e->trackingState->age_pointCloud == -1) e->trackingState->age_pointCloud = -2;
  else e->trackingState->age_pointCloud = 0;
}
// settings.useApproximateRaycast (ITMLibSettings.cpp:32); read by ITMTrackingController::Track (:15)
void ref_set_use_approximate_raycast(ref_engine *e, int on) { e->settings->useApproximateRaycast = on != 0; }
int ref_requires_full_rendering(ref_engine *e) { return e->trackingState->requiresFullRendering ? 1 : 0; }
// what ITMTrackingController::Prepare does when !requiresFullRendering (ITMTrackingController.cpp:40-44)
void ref_forward_render(ref_engine *e) {
  e->vis->ForwardRender(e->view, e->trackingState, e->renderState);
  e->trackingState->age_pointCloud++;
}
float *ref_forward_projection(ref_engine *e) { return (float *)e->renderState->forwardProjection->GetData(MEMORYDEVICE_CPU); }
int *ref_fwd_missing_points(ref_engine *e) { return e->renderState->fwdProjMissingPoints->GetData(MEMORYDEVICE_CPU); }
int ref_no_fwd_missing_points(ref_engine *e) { return e->renderState->noFwdProjMissingPoints; }

// ITMMainEngine::GetImage (ITMMainEngine.cpp:134-192) composed from the same public calls.  type = GetImageType.
// out: Vector4u[w*h].  Returns 0, or -1 when there is no view yet.
int ref_get_image(ref_engine *e, int type, const float *poseM16, const float *intr4, int w, int h, unsigned char *out) {
  if (e->view == NULL) return -1;
  if (e->freeOut == NULL || e->freeOut->noDims.x != w || e->freeOut->noDims.y != h) {
    if (e->freeOut) delete e->freeOut;
    e->freeOut = new ITMUChar4Image(Vector2i(w, h), true, false);
  }
  ITMUChar4Image *o = e->freeOut;
  o->Clear();
  switch (type) {
    case ITMMainEngine::InfiniTAM_IMAGE_ORIGINAL_RGB:
      o->SetFrom(e->view->rgb, ORUtils::MemoryBlock<Vector4u>::CPU_TO_CPU);
      break;
    case ITMMainEngine::InfiniTAM_IMAGE_ORIGINAL_DEPTH:
      IITMVisualisationEngine::DepthToUchar4(o, e->view->depth);
      break;
    case ITMMainEngine::InfiniTAM_IMAGE_SCENERAYCAST:
      o->SetFrom(e->renderState->raycastImage, ORUtils::MemoryBlock<Vector4u>::CPU_TO_CPU);
      break;
    case ITMMainEngine::InfiniTAM_IMAGE_FREECAMERA_SHADED:
    case ITMMainEngine::InfiniTAM_IMAGE_FREECAMERA_COLOUR_FROM_VOLUME:
    case ITMMainEngine::InfiniTAM_IMAGE_FREECAMERA_COLOUR_FROM_NORMAL: {
      IITMVisualisationEngine::RenderImageType rt = IITMVisualisationEngine::RENDER_SHADED_GREYSCALE;
      if (type == ITMMainEngine::InfiniTAM_IMAGE_FREECAMERA_COLOUR_FROM_VOLUME) rt = IITMVisualisationEngine::RENDER_COLOUR_FROM_VOLUME;
      else if (type == ITMMainEngine::InfiniTAM_IMAGE_FREECAMERA_COLOUR_FROM_NORMAL) rt = IITMVisualisationEngine::RENDER_COLOUR_FROM_NORMAL;
      if (e->renderStateFree == NULL || e->renderStateFree->raycastImage->noDims.x != w || e->renderStateFree->raycastImage->noDims.y != h) {
        if (e->renderStateFree) delete e->renderStateFree;
        e->renderStateFree = e->vis->CreateRenderState(o->noDims);
      }
      Matrix4f M(poseM16);
      ITMPose pose; pose.SetM(M);
      ITMIntrinsics intr; intr.SetFrom(intr4[0], intr4[1], intr4[2], intr4[3], (float)w, (float)h);
      e->vis->FindVisibleBlocks(&pose, &intr, e->renderStateFree);
      e->vis->CreateExpectedDepths(&pose, &intr, e->renderStateFree);
      e->vis->RenderImage(&pose, &intr, e->renderStateFree, e->renderStateFree->raycastImage, rt);
      o->SetFrom(e->renderStateFree->raycastImage, ORUtils::MemoryBlock<Vector4u>::CPU_TO_CPU);
      break;
    }
    default: break;
  }
  memcpy(out, o->GetData(MEMORYDEVICE_CPU), (size_t)w * h * 4);
  return 0;
}
// renderState_freeview after the last free-view ref_get_image
int *ref_free_visible_ids(ref_engine *e) { return e->renderStateFree ? ((ITMRenderState_VH *)e->renderStateFree)->GetVisibleEntryIDs() : NULL; }
int ref_free_no_visible(ref_engine *e) { return e->renderStateFree ? ((ITMRenderState_VH *)e->renderStateFree)->noVisibleEntries : -1; }
float *ref_free_minmax(ref_engine *e) { return e->renderStateFree ? (float *)e->renderStateFree->renderingRangeImage->GetData(MEMORYDEVICE_CPU) : NULL; }
float *ref_free_raycast_result(ref_engine *e) { return e->renderStateFree ? (float *)e->renderStateFree->raycastResult->GetData(MEMORYDEVICE_CPU) : NULL; }

// What ITMTrackingController::Prepare does for TRACKER_COLOR (ITMTrackingController.cpp:22-28): expected depths at the colour
// camera's pose, then CreatePointCloud into trackingState->pointCloud (read it back with ref_points / ref_normals).
// trafo16 = ITMExtrinsics::calib to install first (NULL keeps the current one).  Returns noTotalPoints.
int ref_create_point_cloud(ref_engine *e, const float *trafo16, int skipPoints) {
  if (e->view == NULL) return -1;
  if (trafo16) { Matrix4f T(trafo16); e->calib.trafo_rgb_to_depth.SetFrom(T); e->view->calib->trafo_rgb_to_depth.SetFrom(T); }  // the view holds a copy
  ITMPose pose_rgb(e->view->calib->trafo_rgb_to_depth.calib_inv * e->trackingState->pose_d->GetM());
  e->vis->CreateExpectedDepths(&pose_rgb, &(e->view->calib->intrinsics_rgb), e->renderState);
  e->vis->CreatePointCloud(e->view, e->trackingState, e->renderState, skipPoints != 0);
  return e->trackingState->pointCloud->noTotalPoints;
}

// ITMLowLevelEngine's colour-tracker helpers on caller data (ITMLowLevelEngine_CPU.cpp:12-108).
// op: 0 CopyImage(uchar4), 1 FilterSubsample(uchar4), 2 FilterSubsampleWithHoles(Vector4f), 3 GradientX, 4 GradientY.
// in: w*h input pixels; out: the output image's pixels (ops 1, 2: (w/2)*(h/2)); the output image is filled with
// prefillByte first (the gradient drivers clear only part of it).  Returns the number of output bytes.
long long ref_low_level(ref_engine *e, int op, const void *in, int w, int h, void *out, int prefillByte) {
  const Vector2i dims(w, h), half(w / 2, h / 2);
  const size_t P = (size_t)w * h, Q = (size_t)half.x * half.y;
  if (op == 2) {
    ITMFloat4Image src(dims, true, false), dst(half, true, false);
    memcpy(src.GetData(MEMORYDEVICE_CPU), in, P * 16);
    dst.Clear((unsigned char)prefillByte);
    e->lowLevel->FilterSubsampleWithHoles(&dst, &src);
    memcpy(out, dst.GetData(MEMORYDEVICE_CPU), Q * 16);
    return (long long)(Q * 16);
  }
  ITMUChar4Image src(dims, true, false);
  memcpy(src.GetData(MEMORYDEVICE_CPU), in, P * 4);
  if (op == 0 || op == 1) {
    ITMUChar4Image dst(op == 0 ? dims : half, true, false);
    dst.Clear((unsigned char)prefillByte);
    if (op == 0) e->lowLevel->CopyImage(&dst, &src); else e->lowLevel->FilterSubsample(&dst, &src);
    const size_t bytes = (op == 0 ? P : Q) * 4;
    memcpy(out, dst.GetData(MEMORYDEVICE_CPU), bytes);
    return (long long)bytes;
  }
  ITMShort4Image grad(dims, true, false);
  grad.Clear((unsigned char)prefillByte);
  if (op == 3) e->lowLevel->GradientX(&grad, &src); else e->lowLevel->GradientY(&grad, &src);
  memcpy(out, grad.GetData(MEMORYDEVICE_CPU), P * 8);
  return (long long)(P * 8);
}

// ITMMainEngine::UpdateMesh (ITMMainEngine.cpp:97-101): returns noTotalTriangles; *triangles = ITMMesh::Triangle array
int ref_mesh_scene(ref_engine *e, float **triangles, int *noMaxTriangles) {
  if (!e->mesh) { e->mesh = new ITMMesh(MEMORYDEVICE_CPU); e->meshing = new ITMMeshingEngine_CPU<TV, TI>(); }
  e->meshing->MeshScene(e->mesh, e->scene);
  *triangles = (float *)e->mesh->triangles->GetData(MEMORYDEVICE_CPU);
  *noMaxTriangles = (int)ITMMesh::noMaxTriangles;
  return (int)e->mesh->noTotalTriangles;
}
void ref_write_stl(ref_engine *e, const char *fileName) { if (e->mesh) e->mesh->WriteSTL(fileName); }
void ref_write_obj(ref_engine *e, const char *fileName) { if (e->mesh) e->mesh->WriteOBJ(fileName); }

void ref_prepare(ref_engine *e) { e->controller->Prepare(e->trackingState, e->view, e->renderState); }

void ref_process_frame(ref_engine *e, const short *depth) {
  ref_update_view(e, depth);
  e->controller->Track(e->trackingState, e->view);
  e->reco->AllocateSceneFromDepth(e->scene, e->view, e->trackingState, e->renderState);
  e->reco->IntegrateIntoScene(e->scene, e->view, e->trackingState, e->renderState);
  ref_swap(e);
  e->controller->Prepare(e->trackingState, e->view, e->renderState);
}

// same, with wall-clock per stage: view, track, allocate, integrate, expected depths, raycast+ICP maps
void ref_process_frame_timed(ref_engine *e, const short *depth, double *ms6) {
  double t0 = now_ms();
  ref_update_view(e, depth);
  double t1 = now_ms();
  e->controller->Track(e->trackingState, e->view);
  double t2 = now_ms();
  e->reco->AllocateSceneFromDepth(e->scene, e->view, e->trackingState, e->renderState);
  double t3 = now_ms();
  e->reco->IntegrateIntoScene(e->scene, e->view, e->trackingState, e->renderState);
  ref_swap(e);
  double t4 = now_ms();
  ref_expected_depths(e);
  double t5 = now_ms();
  ref_icp_maps(e);
  double t6 = now_ms();
  ms6[0] = t1 - t0; ms6[1] = t2 - t1; ms6[2] = t3 - t2; ms6[3] = t4 - t3; ms6[4] = t5 - t4; ms6[5] = t6 - t5;
}

// ---- single ICP evaluations ------------------------------------------------
// Runs SetEvaluationData + PrepareForEvaluation (builds the depth pyramid).
void ref_icp_prepare(ref_engine *e) {
  e->tracker->SetEvaluationData(e->trackingState, e->view);
  e->tracker->PrepareForEvaluation();
}
// out: [0]=noValidPoints (as float), [1]=f, [2..7]=nabla, [8..43]=hessian 6x6 as returned
int ref_icp_gandh(ref_engine *e, int level, const float *approxInvPose, float *out) {
  e->tracker->SetEvaluationParams(level);
  Matrix4f inv(approxInvPose);
  float f = 0.f, nabla[6] = {0, 0, 0, 0, 0, 0}, hess[36];
  for (int i = 0; i < 36; ++i) hess[i] = 0.f;
  int n = e->tracker->ComputeGandH(f, nabla, hess, inv);
  out[0] = (float)n; out[1] = f;
  for (int i = 0; i < 6; ++i) out[2 + i] = nabla[i];
  for (int i = 0; i < 36; ++i) out[8 + i] = hess[i];
  return n;
}
// ---- weighted ICP (TRACKER_WICP engines) ---------------------------------------
void ref_wicp_prepare(ref_engine *e) {
  e->wtracker->SetEvaluationData(e->trackingState, e->view);
  e->wtracker->PrepareForEvaluation();
}
int ref_wicp_gandh(ref_engine *e, int level, const float *approxInvPose, float *out) {
  e->wtracker->SetEvaluationParams(level);
  Matrix4f inv(approxInvPose);
  float f = 0.f, nabla[6] = {0, 0, 0, 0, 0, 0}, hess[36];
  for (int i = 0; i < 36; ++i) hess[i] = 0.f;
  int n = e->wtracker->ComputeGandH(f, nabla, hess, inv);
  out[0] = (float)n; out[1] = f;
  for (int i = 0; i < 6; ++i) out[2 + i] = nabla[i];
  for (int i = 0; i < 36; ++i) out[8 + i] = hess[i];
  return n;
}
float *ref_depth_uncertainty(ref_engine *e) { return (e->view && e->view->depthUncertainty) ? e->view->depthUncertainty->GetData(MEMORYDEVICE_CPU) : NULL; }
float *ref_depth_normal(ref_engine *e) { return (e->view && e->view->depthNormal) ? (float *)e->view->depthNormal->GetData(MEMORYDEVICE_CPU) : NULL; }

int ref_pyramid_level(ref_engine *e, int level, float **data, int *w, int *h, float *intrinsics4) {
  ITMTemplatedHierarchyLevel<ITMFloatImage> *l = e->tracker->viewHierarchy->levels[level];
  *data = l->depth->GetData(MEMORYDEVICE_CPU);
  *w = l->depth->noDims.x; *h = l->depth->noDims.y;
  intrinsics4[0] = l->intrinsics.x; intrinsics4[1] = l->intrinsics.y;
  intrinsics4[2] = l->intrinsics.z; intrinsics4[3] = l->intrinsics.w;
  return 0;
}
void ref_icp_config(ref_engine *e, int *noLevels, int *itersPerLevel, float *distThresh, int *iterType) {
  *noLevels = e->tracker->viewHierarchy->noLevels;
  for (int i = 0; i < *noLevels; ++i) {
    itersPerLevel[i] = e->tracker->noIterationsPerLevel[i];
    distThresh[i] = e->tracker->distThresh[i];
    iterType[i] = (int)e->tracker->viewHierarchy->levels[i]->iterationType;
  }
}

// ---- pose helpers (ITMLib/Objects/ITMPose.cpp) -----------------------------
void ref_get_pose(ref_engine *e, float *M16) { memcpy(M16, e->trackingState->pose_d->GetM().m, 64); }
void ref_set_pose(ref_engine *e, const float *M16) { Matrix4f M(M16); e->trackingState->pose_d->SetM(M); }
void ref_get_pose_pointcloud(ref_engine *e, float *M16) { memcpy(M16, e->trackingState->pose_pointCloud->GetM().m, 64); }
void ref_set_pose_pointcloud(ref_engine *e, const float *M16) { Matrix4f M(M16); e->trackingState->pose_pointCloud->SetM(M); }
void ref_get_pose_params(ref_engine *e, float *p6) { for (int i = 0; i < 6; ++i) p6[i] = e->trackingState->pose_d->params.all[i]; }
int ref_get_age(ref_engine *e) { return e->trackingState->age_pointCloud; }
void ref_set_age(ref_engine *e, int a) { e->trackingState->age_pointCloud = a; }
void ref_mat_inv(const float *in16, float *out16) { Matrix4f a(in16), b; a.inv(b); memcpy(out16, b.m, 64); }
// SetInvM + Coerce + GetM/GetInvM round trip used by the LM loop (ITMDepthTracker.cpp:190-193)
void ref_pose_from_invm_coerced(const float *invM16, float *M16, float *invOut16, float *params6) {
  ITMPose p; Matrix4f inv(invM16);
  p.SetInvM(inv); p.Coerce();
  memcpy(M16, p.GetM().m, 64);
  Matrix4f i2 = p.GetInvM(); memcpy(invOut16, i2.m, 64);
  for (int k = 0; k < 6; ++k) params6[k] = p.params.all[k];
}
void ref_pose_from_params(const float *params6, float *M16) {
  ITMPose p(params6[0], params6[1], params6[2], params6[3], params6[4], params6[5]);
  memcpy(M16, p.GetM().m, 64);
}
// Cholesky solve exactly as ComputeDelta does (ITMDepthTracker.cpp:85-102)
void ref_compute_delta(ref_engine *e, const float *nabla, const float *hessian36, int shortIteration, float *step6) {
  float n[6], h[36];
  memcpy(n, nabla, sizeof(n)); memcpy(h, hessian36, sizeof(h));
  e->tracker->ComputeDelta(step6, n, h, shortIteration != 0);
}

// ---- raw state -------------------------------------------------------------
void *ref_hash_entries(ref_engine *e) { return e->scene->index.GetEntries(); }
void *ref_voxels(ref_engine *e) { return e->scene->localVBA.GetVoxelBlocks(); }
int *ref_vba_alloc_list(ref_engine *e) { return e->scene->localVBA.GetAllocationList(); }
int *ref_excess_alloc_list(ref_engine *e) { return e->scene->index.GetExcessAllocationList(); }
int *ref_visible_ids(ref_engine *e) { return ((ITMRenderState_VH *)e->renderState)->GetVisibleEntryIDs(); }
unsigned char *ref_visible_types(ref_engine *e) { return ((ITMRenderState_VH *)e->renderState)->GetEntriesVisibleType(); }
// counters: [0]=noVisibleEntries [1]=lastFreeBlockId [2]=lastFreeExcessListId
void ref_get_counters(ref_engine *e, int *c3) {
  c3[0] = ((ITMRenderState_VH *)e->renderState)->noVisibleEntries;
  c3[1] = e->scene->localVBA.lastFreeBlockId;
  c3[2] = e->scene->index.GetLastFreeExcessListId();
}
void ref_set_counters(ref_engine *e, const int *c3) {
  ((ITMRenderState_VH *)e->renderState)->noVisibleEntries = c3[0];
  e->scene->localVBA.lastFreeBlockId = c3[1];
  e->scene->index.SetLastFreeExcessListId(c3[2]);
}
float *ref_depth(ref_engine *e) { return e->view ? e->view->depth->GetData(MEMORYDEVICE_CPU) : NULL; }
float *ref_minmax(ref_engine *e) { return (float *)e->renderState->renderingRangeImage->GetData(MEMORYDEVICE_CPU); }
float *ref_raycast_result(ref_engine *e) { return (float *)e->renderState->raycastResult->GetData(MEMORYDEVICE_CPU); }
unsigned char *ref_raycast_image(ref_engine *e) { return (unsigned char *)e->renderState->raycastImage->GetData(MEMORYDEVICE_CPU); }
float *ref_points(ref_engine *e) { return (float *)e->trackingState->pointCloud->locations->GetData(MEMORYDEVICE_CPU); }
float *ref_normals(ref_engine *e) { return (float *)e->trackingState->pointCloud->colours->GetData(MEMORYDEVICE_CPU); }

}  // extern "C"
