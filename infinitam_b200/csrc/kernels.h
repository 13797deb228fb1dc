// Host-side launch interface of the CUDA kernels (internal to libitm_b200.so).
#pragma once
#include <cuda_runtime.h>
#include "itm_common.cuh"

namespace itm {

struct AllocArgs {
  const float *depth;            // view->depth, float metres
  void *hashTable;               // ITMHashEntry[nEntries]
  const int *vbaAllocList;       // ITMLocalVBA allocation list
  const int *excessAllocList;    // ITMVoxelBlockHash excess allocation list
  int *visibleIds;               // ITMRenderState_VH::visibleEntryIDs
  unsigned char *visType;        // ITMRenderState_VH::entriesVisibleType
  int visibleCapacity;           // SDF_LOCAL_BLOCK_NUM
  unsigned *allocKey;            // scratch: one key per slot (padded to 8192), all zero between frames
  unsigned long long *scanTickets;     // scratch: [0] alloc scan, [1] visible scan
  unsigned long long *allocTileState;  // scratch: one word per 8192-slot tile
  unsigned long long *visTileState;
  FrameState *st;
  ViewParams vp;
  SceneParams sp;
  int onlyUpdateVisibleList;
};

struct IntegrateArgs {
  const float *depth;
  void *voxels;
  const void *hashTable;
  const int *visibleIds;
  FrameState *st;
  ViewParams vp;
  SceneParams sp;
};

struct RenderArgs {
  const void *voxels;
  const void *hashTable;
  const int *visibleIds;
  float *minmax;        // Vector2f[W*H] renderingRangeImage
  float *raycastResult; // Vector4f[W*H]
  float *pointsMap;     // Vector4f[W*H]
  float *normalsMap;    // Vector4f[W*H]
  unsigned char *raycastImage;  // Vector4u[W*H]
  FrameState *st;
  ViewParams vp;
  SceneParams sp;
};

struct IcpLevelArgs {
  const float *depth;     // level depth image
  int w, h;
  float fx, fy, cx, cy;   // level intrinsics
  float distThresh;
  int iterationType;
};

struct IcpArgs {
  const float *pointsMap;
  const float *normalsMap;
  ViewParams sceneVp;     // full-resolution maps + level-0 intrinsics
  FrameState *st;
  double *partials;       // scratch: [maxCtas][32]
  unsigned *ctaCounter;   // scratch
  float terminationThreshold;
};

int alloc_step_bound(const SceneParams &sp);
void launch_reset_scene(void *voxels, int *vbaAllocList, void *table, int *excessAllocList, const SceneParams &sp, cudaStream_t s);
void launch_allocate(const AllocArgs &a, cudaStream_t s);
void launch_integrate(const IntegrateArgs &a, cudaStream_t s);
void launch_expected_depths(const RenderArgs &a, cudaStream_t s);
void launch_raycast(const RenderArgs &a, cudaStream_t s);
void launch_icp_maps(const RenderArgs &a, cudaStream_t s);

void launch_convert_depth(const short *raw, float *out, int n, float a, float b, cudaStream_t s);
void launch_subsample_holes(float *out, const float *in, int wIn, int hIn, cudaStream_t s);
void launch_view_pyramid(const short *raw, float a, float b, float *const *levels, int W, int H, int nLevels, cudaStream_t s);

// The whole ITMDepthTracker::TrackCamera LM loop as ONE persistent cooperative kernel (levels[l], iters[l] for
// l < nLevels; barrier = 2 zero-initialised words of scratch).
cudaError_t launch_icp_track(const IcpArgs &a, const IcpLevelArgs *levels, const int *iters, int nLevels, int noIcpLevel,
                             unsigned *barrier, cudaStream_t s);
// One stand-alone evaluation at poseIn (16 floats, device); [n, f, nabla6, hessian36] left in out44 (device).
void launch_icp_eval_single(const IcpArgs &a, const IcpLevelArgs &lv, float *out44, const float *poseIn, cudaStream_t s);
int icp_max_ctas();
int icp_track_grid();

void launch_set_pose(FrameState *st, cudaStream_t s);  // recompute invM_d from M_d on device

}  // namespace itm
