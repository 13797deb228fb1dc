// Host side of libitm_b200.so: the C ABI declared in include/itm_b200.h.
//
// Layer A forwards one reference engine method per call onto caller-owned device buffers;
// Layer B composes them the way ITMMainEngine does (ITMLib/Engine/ITMMainEngine.cpp:17-127,
// ITMLib/Engine/ITMDenseMapper.cpp:51-65, ITMLib/Engine/ITMTrackingController.cpp:11-46) with
// all cross-frame state resident in HBM and no host synchronisation inside a frame.
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/itm_b200.h"
#include "itm_common.cuh"
#include "kernels.h"
#include "pose_math.cuh"

using namespace itm;

namespace {

thread_local std::string g_lastError;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const std::string &msg) {
  g_lastError = msg;
  return code;
}

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ? ITM_B200_ENODEVICE : ITM_B200_ECUDA, \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                             \
  } while (0)

// releases report failures instead of leaving a pending error behind for whoever shares the CUDA runtime
#define RELEASE(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      fprintf(stderr, "libitm_b200: %s: %s\n", #expr, cudaGetErrorString(_e));                     \
      cudaGetLastError();                                                                          \
    }                                                                                              \
  } while (0)

// Every entry point runs on the context's device and leaves the caller's current device as it found it (hosts such as
// PyTorch switch devices between calls; two engines on different devices may live in one process).
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  explicit DeviceScope(int dev) {
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceScope() {
    if (switched) cudaSetDevice(prev);
  }
};
#define ON_DEVICE_OF_CTX(c) DeviceScope _deviceScope((c) ? (c)->p.device : -1)
#define ON_DEVICE_OF_ENGINE(e) DeviceScope _deviceScope(((e) && (e)->c) ? (e)->c->p.device : -1)

struct LevelCfg {
  int w, h;
  float fx, fy, cx, cy;
  float distThresh;
  int iterationType;
  int noIterations;
};

}  // namespace

struct itm_b200_ctx {
  itm_b200_params p;
  SceneParams sp;
  ViewParams vp;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  // scratch
  unsigned *allocKey = nullptr;
  AllocLists allocLists = {};
  unsigned long long *scanTickets = nullptr;
  unsigned long long *allocTileState = nullptr;
  unsigned long long *visTileState = nullptr;
  unsigned long long *otherTileState = nullptr;  // FindVisibleBlocks / swap selection / meshing: 8192-slot tiles, ticket [2]
  unsigned long long *pcTileState = nullptr;     // CreatePointCloud: pcTiles 8192-pixel tiles + their ticket, made on first use
  int pcTiles = 0;
  double *icpPartials = nullptr;
  unsigned *icpCounter = nullptr;
  unsigned long long *icpRows = nullptr;   // tagged CTA partial sums of k_icp_track
  unsigned long long *icpBcast = nullptr;  // tagged pose broadcast ring
  unsigned *icpEpochDev = nullptr;         // launch number of k_icp_track, advanced on the device, never reset
  // MeshScene scratch (allocated on first use)
  int *meshBlockList = nullptr;
  unsigned *meshCounts = nullptr;
  unsigned long long *meshOffsets = nullptr;
  FrameState *meshSt = nullptr;       // device
  FrameState *meshHst = nullptr;      // pinned
  int *fwdKey = nullptr;       // ForwardRender: winning source pixel + 1 per destination pixel, all zero between calls
  float *icpOut = nullptr;     // 44 floats
  float *icpPoseIn = nullptr;  // 16 floats
  FrameState *st = nullptr;    // device
  FrameState *hst = nullptr;   // pinned host mirror
  float *pyramid[ITM_MAX_LEVELS] = {nullptr};  // levels 1.. (level 0 is the caller's depth image)
  float *weightPyramid[ITM_MAX_LEVELS] = {nullptr};  // TRACKER_WICP: levels 1.. of the depth-uncertainty hierarchy (allocated on first use)
  LevelCfg levels[ITM_MAX_LEVELS];
  int nLevels = 0;
};

namespace {

int validate_params(const itm_b200_params *p) {
  if (!p) return fail(ITM_B200_EINVAL, "params is NULL");
  if (p->width <= 0 || p->height <= 0) return fail(ITM_B200_EINVAL, "image size must be positive");
  if (p->sdf_bucket_num <= 0 || (p->sdf_bucket_num & (p->sdf_bucket_num - 1))) return fail(ITM_B200_EINVAL, "sdf_bucket_num must be a power of two");
  if (p->sdf_local_block_num <= 0 || p->sdf_excess_list_size <= 0) return fail(ITM_B200_EINVAL, "pool sizes must be positive");
  if ((long long)p->sdf_bucket_num + p->sdf_excess_list_size >= (1 << 21) * 1024LL) return fail(ITM_B200_EUNSUPPORTED, "hash table too large");
  if (p->no_hierarchy_levels < 1 || p->no_hierarchy_levels > ITM_MAX_LEVELS) return fail(ITM_B200_EINVAL, "no_hierarchy_levels out of range");
  if (!(p->voxel_size > 0) || !(p->mu > 0)) return fail(ITM_B200_EINVAL, "voxel_size and mu must be positive");
  if (p->voxel_type != ITM_B200_VOXEL_S && p->voxel_type != ITM_B200_VOXEL_S_RGB) return fail(ITM_B200_EINVAL, "unknown voxel_type");
  if (p->no_icp_run_till_level < 0 || p->no_icp_run_till_level >= p->no_hierarchy_levels)
    return fail(ITM_B200_EINVAL, "no_icp_run_till_level must lie in [0, no_hierarchy_levels)");
  for (int l = 0; l < p->no_hierarchy_levels; ++l)
    if (p->tracking_regime[l] < ITM_B200_ITER_ROTATION || p->tracking_regime[l] > ITM_B200_ITER_NONE)
      return fail(ITM_B200_EINVAL, "tracking_regime[] entries must be ITM_B200_ITER_ROTATION .. ITM_B200_ITER_NONE");
  // the coarsest pyramid level must still be an image, and the stencils (ICP maps +-2 px, integration [1, W-2]) need room
  if ((p->width >> (p->no_hierarchy_levels - 1)) < 1 || (p->height >> (p->no_hierarchy_levels - 1)) < 1 || p->width < 8 || p->height < 8)
    return fail(ITM_B200_EINVAL, "image too small for the pyramid / stencils (at least 8x8 and 1 pixel on the coarsest level)");
  if (p->tracker_type != ITM_B200_TRACKER_ICP && p->tracker_type != ITM_B200_TRACKER_EXTERNAL && p->tracker_type != ITM_B200_TRACKER_WICP)
    return fail(ITM_B200_EINVAL, "unknown tracker_type");
  if (p->depth_source != ITM_B200_DEPTH_AFFINE && p->depth_source != ITM_B200_DEPTH_KINECT_DISPARITY)
    return fail(ITM_B200_EINVAL, "unknown depth_source");
  if (p->icp_max_ctas < 0) return fail(ITM_B200_EINVAL, "icp_max_ctas must be >= 0");
  if (p->swap_cache_blocks < 0) return fail(ITM_B200_EINVAL, "swap_cache_blocks must be >= 0");
  return ITM_B200_OK;
}

void derive(itm_b200_ctx *c) {
  const itm_b200_params &p = c->p;
  c->sp.voxelSize = p.voxel_size;
  c->sp.mu = p.mu;
  c->sp.maxW = p.max_w;
  c->sp.vfMin = p.view_frustum_min;
  c->sp.vfMax = p.view_frustum_max;
  c->sp.stopAtMaxW = p.stop_integrating_at_max_w;
  c->sp.nLocal = p.sdf_local_block_num;
  c->sp.nBuckets = p.sdf_bucket_num;
  c->sp.nExcess = p.sdf_excess_list_size;
  c->sp.nEntries = p.sdf_bucket_num + p.sdf_excess_list_size;
  c->sp.hashMask = (unsigned)p.sdf_bucket_num - 1u;
  c->sp.voxelWords = p.voxel_type == ITM_B200_VOXEL_S_RGB ? 2 : 1;
  c->vp.W = p.width;
  c->vp.H = p.height;
  c->vp.fx = p.fx;
  c->vp.fy = p.fy;
  c->vp.cx = p.cx;
  c->vp.cy = p.cy;
  // ITMDepthTracker constructor (ITMLib/Engine/ITMDepthTracker.cpp:11-36) and
  // PrepareForEvaluation's intrinsics halving (:62-75)
  c->nLevels = p.no_hierarchy_levels;
  int w = p.width, h = p.height;
  float fx = p.fx, fy = p.fy, cx = p.cx, cy = p.cy;
  for (int l = 0; l < c->nLevels; ++l) {
    LevelCfg &L = c->levels[l];
    L.w = w; L.h = h; L.fx = fx; L.fy = fy; L.cx = cx; L.cy = cy;
    L.iterationType = p.tracking_regime[l];
    L.noIterations = 2 + 2 * l;
    w /= 2; h /= 2;
    fx = fx * 0.5f; fy = fy * 0.5f; cx = cx * 0.5f; cy = cy * 0.5f;
  }
  const float distThresh = p.depth_tracker_icp_threshold;
  const float step = distThresh / c->nLevels;
  c->levels[c->nLevels - 1].distThresh = distThresh;
  for (int l = c->nLevels - 2; l >= 0; --l) c->levels[l].distThresh = c->levels[l + 1].distThresh - step;
}

int ctx_alloc(itm_b200_ctx *c, void *stream) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(ITM_B200_ENODEVICE, "no CUDA device available: this library has no CPU fallback");
  if (c->p.device < 0 || c->p.device >= ndev) return fail(ITM_B200_EINVAL, "params.device is not a visible CUDA device");
  CU(cudaSetDevice(c->p.device));  // the caller's device is restored by the entry point's DeviceScope
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->ownStream = true;
  }
  const int numTiles = (c->sp.nEntries + 8191) / 8192;
  const long long maxKey = (long long)c->p.width * c->p.height * alloc_step_bound(c->sp);
  if (maxKey >= 0xFFFFFFFFLL) return fail(ITM_B200_EUNSUPPORTED, "image size x ray-segment steps exceeds the 32-bit allocation key");
  CU(cudaMalloc(&c->allocKey, (size_t)numTiles * 8192 * sizeof(unsigned)));
  CU(cudaMemsetAsync(c->allocKey, 0, (size_t)numTiles * 8192 * sizeof(unsigned), c->stream));
  // every scan kernel family has its own ticket (tile = ticket % its own tile count): [0] allocation scan, [1] visible scan,
  // [2] the 8192-slot scans of FindVisibleBlocks / swapping / meshing
  const int allocTiles = (c->sp.nEntries + alloc_scan_tile() - 1) / alloc_scan_tile();
  CU(cudaMalloc(&c->scanTickets, 3 * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(c->scanTickets, 0, 3 * sizeof(unsigned long long), c->stream));
  CU(cudaMalloc(&c->allocTileState, allocTiles * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(c->allocTileState, 0, allocTiles * sizeof(unsigned long long), c->stream));
  CU(cudaMalloc(&c->visTileState, allocTiles * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(c->visTileState, 0, allocTiles * sizeof(unsigned long long), c->stream));
  {
    AllocLists &L = c->allocLists;
    L.numBins = (c->sp.nEntries + ITM_ALLOC_BIN - 1) / ITM_ALLOC_BIN;
    const size_t nWords = (size_t)L.numBins * ITM_ALLOC_BIN / 32;
    CU(cudaMalloc(&L.claimBits, nWords * sizeof(unsigned)));
    CU(cudaMemsetAsync(L.claimBits, 0, nWords * sizeof(unsigned), c->stream));
    CU(cudaMalloc(&L.reqList, (size_t)L.numBins * ITM_ALLOC_BIN * sizeof(int)));
    CU(cudaMalloc(&L.newVisList, (size_t)L.numBins * ITM_ALLOC_BIN * sizeof(int)));
    CU(cudaMalloc(&L.prevCopy, (size_t)(c->sp.nLocal + 32) * sizeof(int)));
    CU(cudaMalloc(&L.prevKeep, (size_t)(c->sp.nLocal + 32)));
    CU(cudaMalloc(&L.binCounts, (size_t)5 * L.numBins * sizeof(int)));
    CU(cudaMemsetAsync(L.binCounts, 0, (size_t)5 * L.numBins * sizeof(int), c->stream));
    CU(cudaMalloc(&L.done, sizeof(int)));
    CU(cudaMemsetAsync(L.done, 0, sizeof(int), c->stream));
  }
  CU(cudaMalloc(&c->otherTileState, numTiles * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(c->otherTileState, 0, numTiles * sizeof(unsigned long long), c->stream));
  CU(cudaMalloc(&c->icpPartials, (size_t)icp_max_ctas() * 32 * sizeof(double)));
  CU(cudaMalloc(&c->icpCounter, sizeof(unsigned)));
  CU(cudaMemsetAsync(c->icpCounter, 0, sizeof(unsigned), c->stream));
  CU(cudaMalloc(&c->icpRows, icp_rows_bytes()));
  CU(cudaMemsetAsync(c->icpRows, 0, icp_rows_bytes(), c->stream));
  CU(cudaMalloc(&c->icpEpochDev, sizeof(unsigned)));
  CU(cudaMemsetAsync(c->icpEpochDev, 0, sizeof(unsigned), c->stream));
  CU(cudaMalloc(&c->icpBcast, icp_bcast_bytes()));
  CU(cudaMemsetAsync(c->icpBcast, 0, icp_bcast_bytes(), c->stream));
  CU(cudaMalloc(&c->fwdKey, (size_t)c->p.width * c->p.height * sizeof(int)));
  CU(cudaMemsetAsync(c->fwdKey, 0, (size_t)c->p.width * c->p.height * sizeof(int), c->stream));
  CU(cudaMalloc(&c->icpOut, 44 * sizeof(float)));
  CU(cudaMalloc(&c->icpPoseIn, 16 * sizeof(float)));
  CU(cudaMalloc(&c->st, sizeof(FrameState)));
  CU(cudaMemsetAsync(c->st, 0, sizeof(FrameState), c->stream));
  CU(cudaMallocHost(&c->hst, sizeof(FrameState)));
  memset(c->hst, 0, sizeof(FrameState));
  for (int l = 1; l < c->nLevels; ++l) {
    const size_t n = (size_t)c->levels[l].w * c->levels[l].h;
    CU(cudaMalloc(&c->pyramid[l], (n ? n : 1) * sizeof(float)));
  }
  // occupancy / attribute queries of the persistent kernels happen here, never inside a stream capture
  (void)icp_track_grid();
  (void)integrate_grid();
  CU(cudaStreamSynchronize(c->stream));
  return ITM_B200_OK;
}

void ctx_free(itm_b200_ctx *c) {
  if (!c) return;
  RELEASE(cudaFree(c->allocKey));
  RELEASE(cudaFree(c->allocLists.claimBits));
  RELEASE(cudaFree(c->allocLists.reqList));
  RELEASE(cudaFree(c->allocLists.newVisList));
  RELEASE(cudaFree(c->allocLists.prevCopy));
  RELEASE(cudaFree(c->allocLists.prevKeep));
  RELEASE(cudaFree(c->allocLists.binCounts));
  RELEASE(cudaFree(c->allocLists.done));
  RELEASE(cudaFree(c->scanTickets));
  RELEASE(cudaFree(c->allocTileState));
  RELEASE(cudaFree(c->visTileState));
  RELEASE(cudaFree(c->otherTileState));
  RELEASE(cudaFree(c->pcTileState));
  RELEASE(cudaFree(c->icpPartials));
  RELEASE(cudaFree(c->icpCounter));
  RELEASE(cudaFree(c->icpRows));
  RELEASE(cudaFree(c->icpBcast));
  RELEASE(cudaFree(c->icpEpochDev));
  RELEASE(cudaFree(c->icpOut));
  RELEASE(cudaFree(c->fwdKey));
  RELEASE(cudaFree(c->meshBlockList)); RELEASE(cudaFree(c->meshCounts)); RELEASE(cudaFree(c->meshOffsets)); RELEASE(cudaFree(c->meshSt));
  if (c->meshHst) RELEASE(cudaFreeHost(c->meshHst));
  RELEASE(cudaFree(c->icpPoseIn));
  RELEASE(cudaFree(c->st));
  if (c->hst) RELEASE(cudaFreeHost(c->hst));
  for (int l = 1; l < ITM_MAX_LEVELS; ++l) RELEASE(cudaFree(c->pyramid[l]));
  for (int l = 1; l < ITM_MAX_LEVELS; ++l) RELEASE(cudaFree(c->weightPyramid[l]));
  if (c->ownStream && c->stream) RELEASE(cudaStreamDestroy(c->stream));
}

// identity pose etc.
void host_state_init(FrameState *h, const SceneParams &sp) {
  memset(h, 0, sizeof(FrameState));
  const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memcpy(h->M_d, I, sizeof(I));
  memcpy(h->invM_d, I, sizeof(I));
  memcpy(h->scenePose, I, sizeof(I));
  h->noVisibleEntries = 0;
  h->lastFreeBlockId = sp.nLocal - 1;
  h->lastFreeExcessId = sp.nExcess - 1;
  h->agePointCloud = -1;
  h->requiresFullRendering = 1;
}

void set_pose_host(FrameState *h, const float *M) {
  memcpy(h->M_d, M, 64);
  mat4_inv(M, h->invM_d);
  pose_M_to_params(M, h->poseParams);
}

int push_state(itm_b200_ctx *c) {
  CU(cudaMemcpyAsync(c->st, c->hst, sizeof(FrameState), cudaMemcpyHostToDevice, c->stream));
  return ITM_B200_OK;
}
int pull_state(itm_b200_ctx *c) {
  CU(cudaMemcpyAsync(c->hst, c->st, sizeof(FrameState), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

AllocArgs make_alloc_args(itm_b200_ctx *c, const float *depth, void *hash, const int *vbaList, const int *exList, int *visIds,
                          unsigned char *visType, int onlyVisible) {
  AllocArgs a;
  a.depth = depth;
  a.hashTable = hash;
  a.vbaAllocList = vbaList;
  a.excessAllocList = exList;
  a.visibleIds = visIds;
  a.visType = visType;
  a.visibleCapacity = c->sp.nLocal;
  a.allocKey = c->allocKey;
  a.lists = c->allocLists;
  a.scanTickets = c->scanTickets;
  a.allocTileState = c->allocTileState;
  a.visTileState = c->visTileState;
  a.st = c->st;
  a.vp = c->vp;
  a.sp = c->sp;
  a.onlyUpdateVisibleList = onlyVisible;
  a.swapStates = nullptr;
  a.prologueDone = 0;
  memset(&a.shard, 0, sizeof(a.shard));
  a.shard.world = 1;
  a.residentVisibleIds = nullptr;
  return a;
}

void fill_integrate_calib(const itm_b200_ctx *c, IntegrateArgs &a, const unsigned char *rgb) {
  a.rgb = rgb;
  a.rgbIntr[0] = c->p.rgb_fx; a.rgbIntr[1] = c->p.rgb_fy; a.rgbIntr[2] = c->p.rgb_cx; a.rgbIntr[3] = c->p.rgb_cy;
  memcpy(a.calibInv, c->p.trafo_rgb_to_depth_inv, sizeof(a.calibInv));
}

// image size of a render state (free-view render states may differ from the sensor) + the intrinsics of this call
ViewParams rs_view(const itm_b200_ctx *c, const itm_b200_render_state *rs, const float intrinsics[4]) {
  ViewParams vp = c->vp;
  if (rs->img_width > 0 && rs->img_height > 0) {
    vp.W = rs->img_width;
    vp.H = rs->img_height;
  }
  vp.fx = intrinsics[0]; vp.fy = intrinsics[1]; vp.cx = intrinsics[2]; vp.cy = intrinsics[3];
  return vp;
}

IcpLevelArgs make_level_args(const itm_b200_ctx *c, int l, const float *depth) {
  const LevelCfg &L = c->levels[l];
  IcpLevelArgs lv;
  lv.weight = nullptr;
  lv.depth = depth;
  lv.w = L.w; lv.h = L.h;
  lv.fx = L.fx; lv.fy = L.fy; lv.cx = L.cx; lv.cy = L.cy;
  lv.distThresh = L.distThresh;
  lv.iterationType = L.iterationType;
  return lv;
}

// ITMDepthTracker::TrackCamera: pyramid (unless already built) + LM loop, all enqueued
int ensure_weight_pyramid(itm_b200_ctx *c) {
  for (int l = 1; l < c->nLevels; ++l) {
    if (c->weightPyramid[l]) continue;
    const size_t n = (size_t)c->levels[l].w * c->levels[l].h;
    CU(cudaMalloc(&c->weightPyramid[l], (n ? n : 1) * sizeof(float)));
  }
  return ITM_B200_OK;
}

// weight0 != NULL: ITMWeightedICPTracker (weight hierarchy from view->depthUncertainty, Gauss-Newton)
int enqueue_track(itm_b200_ctx *c, const float *depth0, const float *points, const float *normals, bool buildPyramid,
                  bool epochBumped = false, const float *weight0 = nullptr) {
  cudaStream_t s = c->stream;
  if (buildPyramid && c->nLevels > 1) {
    float *lv[ITM_MAX_LEVELS];
    lv[0] = const_cast<float *>(depth0);
    for (int l = 1; l < c->nLevels; ++l) lv[l] = c->pyramid[l];
    launch_view_pyramid(nullptr, 0.f, 0.f, lv, c->vp.W, c->vp.H, c->nLevels, s);
    g_launches += 1 + (c->nLevels > 5 ? c->nLevels - 5 : 0);
  }
  if (weight0 && c->nLevels > 1) {
    // PrepareForEvaluation's second hierarchy (ITMWeightedICPTracker.cpp:84-85): the same hole-aware subsampling of sigma_z
    int rc = ensure_weight_pyramid(c);
    if (rc) return rc;
    float *lv[ITM_MAX_LEVELS];
    lv[0] = const_cast<float *>(weight0);
    for (int l = 1; l < c->nLevels; ++l) lv[l] = c->weightPyramid[l];
    launch_view_pyramid(nullptr, 0.f, 0.f, lv, c->vp.W, c->vp.H, c->nLevels, s);
    g_launches += 1 + (c->nLevels > 5 ? c->nLevels - 5 : 0);
  }
  IcpArgs a;
  a.pointsMap = points;
  a.normalsMap = normals;
  a.sceneVp = c->vp;
  a.st = c->st;
  a.partials = c->icpPartials;
  a.ctaCounter = c->icpCounter;
  a.terminationThreshold = c->p.depth_tracker_termination_threshold;
  IcpLevelArgs lv[ITM_MAX_LEVELS];
  int iters[ITM_MAX_LEVELS];
  for (int l = 0; l < c->nLevels; ++l) {
    lv[l] = make_level_args(c, l, l == 0 ? depth0 : c->pyramid[l]);
    if (weight0) lv[l].weight = l == 0 ? weight0 : c->weightPyramid[l];
    iters[l] = c->levels[l].noIterations;
  }
  const cudaError_t e = launch_icp_track(a, lv, iters, c->nLevels, c->p.no_icp_run_till_level, c->icpRows, c->icpBcast, c->icpEpochDev,
                                         !epochBumped, c->p.icp_max_ctas, s, weight0 != nullptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ITM_B200_ECUDA, std::string("cooperative launch of the ICP tracker failed: ") + cudaGetErrorString(e));
  }
  g_launches += epochBumped ? 1 : 2;
  return ITM_B200_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

void itm_b200_default_params(itm_b200_params *p, int width, int height) {
  memset(p, 0, sizeof(*p));
  p->width = width;
  p->height = height;
  const float s = (float)width / 640.0f;
  p->fx = 580.0f * s;
  p->fy = 580.0f * s;
  p->cx = (float)width / 2.0f;
  p->cy = (float)height / 2.0f;
  p->voxel_size = 0.005f;
  p->mu = 0.02f;
  p->max_w = 100;
  p->view_frustum_min = 0.35f;
  p->view_frustum_max = 3.0f;
  p->stop_integrating_at_max_w = 0;
  p->depth_calib_a = 1.0f / 1000.0f;
  p->depth_calib_b = 0.0f;
  p->sdf_local_block_num = 0x10000;
  p->sdf_bucket_num = 0x100000;
  p->sdf_excess_list_size = 0x20000;
  p->no_hierarchy_levels = 5;
  p->tracking_regime[0] = ITM_B200_ITER_BOTH;
  p->tracking_regime[1] = ITM_B200_ITER_BOTH;
  p->tracking_regime[2] = ITM_B200_ITER_ROTATION;
  p->tracking_regime[3] = ITM_B200_ITER_ROTATION;
  p->tracking_regime[4] = ITM_B200_ITER_ROTATION;
  p->no_icp_run_till_level = 0;
  p->depth_tracker_icp_threshold = 0.1f * 0.1f;
  p->depth_tracker_termination_threshold = 1e-3f;
  p->device = 0;
  p->voxel_type = ITM_B200_VOXEL_S;
  p->rgb_fx = p->fx; p->rgb_fy = p->fy; p->rgb_cx = p->cx; p->rgb_cy = p->cy;
  for (int i = 0; i < 16; ++i) p->trafo_rgb_to_depth_inv[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

const char *itm_b200_last_error(void) { return g_lastError.c_str(); }

int itm_b200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(ITM_B200_ENODEVICE, cudaGetErrorString(e));
  return n;
}

unsigned long long itm_b200_launch_count(void) { return g_launches.load(); }

int itm_b200_set_alloc_mode(int mode) {
  if (mode < 0 || mode > 2) return fail(ITM_B200_EINVAL, "alloc mode must be 0 (automatic), 1 (scans) or 2 (lists)");
  const int prev = alloc_mode();
  alloc_mode() = mode;
  return prev;
}

// the CUDA runtime's pending (non-sticky) error of the calling thread, cleared by the call; 0 = none.  Every entry point
// of this library is meant to leave none behind - tests check that.
int itm_b200_take_cuda_error(void) { return (int)cudaGetLastError(); }

int itm_b200_ctx_create(const itm_b200_params *params, void *stream, itm_b200_ctx **out) {
  if (!out) return fail(ITM_B200_EINVAL, "out is NULL");
  *out = nullptr;
  int rc = validate_params(params);
  if (rc) return rc;
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      return fail(ITM_B200_ENODEVICE, "no CUDA device available: this library has no CPU fallback");
    }
    if (params->device < 0 || params->device >= ndev) return fail(ITM_B200_EINVAL, "params.device is not a visible CUDA device");
  }
  DeviceScope deviceScope(params->device);
  itm_b200_ctx *c = new itm_b200_ctx();
  c->p = *params;
  derive(c);
  rc = ctx_alloc(c, stream);
  if (rc) {
    ctx_free(c);
    delete c;
    return rc;
  }
  host_state_init(c->hst, c->sp);
  *out = c;
  return ITM_B200_OK;
}

void itm_b200_ctx_destroy(itm_b200_ctx *ctx) {
  ON_DEVICE_OF_CTX(ctx);
  if (!ctx) return;
  ctx_free(ctx);
  delete ctx;
}

int itm_b200_reset_scene(itm_b200_ctx *c, itm_b200_scene *scene) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene) return fail(ITM_B200_EINVAL, "NULL argument");
  launch_reset_scene(scene->voxel_blocks_dev, scene->vba_allocation_list_dev, scene->hash_entries_dev, scene->excess_allocation_list_dev,
                     c->sp, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  scene->last_free_block_id = c->sp.nLocal - 1;
  scene->last_free_excess_list_id = c->sp.nExcess - 1;
  return ITM_B200_OK;
}

int itm_b200_allocate_scene_from_depth(itm_b200_ctx *c, itm_b200_scene *scene, itm_b200_render_state *rs, const float *depth_dev,
                                       const float pose_M[16], int only_update_visible_list) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !depth_dev || !pose_M) return fail(ITM_B200_EINVAL, "NULL argument");
  set_pose_host(c->hst, pose_M);
  c->hst->noVisibleEntries = rs->no_visible_entries;
  c->hst->lastFreeBlockId = scene->last_free_block_id;
  c->hst->lastFreeExcessId = scene->last_free_excess_list_id;
  c->hst->errorFlags = 0;
  int rc = push_state(c);
  if (rc) return rc;
  AllocArgs aa = make_alloc_args(c, depth_dev, scene->hash_entries_dev, scene->vba_allocation_list_dev, scene->excess_allocation_list_dev,
                                 rs->visible_entry_ids_dev, rs->entries_visible_type_dev, only_update_visible_list);
  aa.swapStates = scene->swap_states_dev;
  launch_allocate(aa, c->stream);
  g_launches += 4;
  rc = pull_state(c);
  if (rc) return rc;
  rs->no_visible_entries = c->hst->noVisibleEntries;
  scene->last_free_block_id = c->hst->lastFreeBlockId;
  scene->last_free_excess_list_id = c->hst->lastFreeExcessId;
  if (c->hst->errorFlags & 1) return fail(ITM_B200_EUNSUPPORTED, "allocation ray segment longer than the supported step bound");
  return ITM_B200_OK;
}

static int integrate_layer_a(itm_b200_ctx *c, itm_b200_scene *scene, const itm_b200_render_state *rs, const float *depth_dev,
                             const unsigned char *rgb_dev, const float pose_M[16]) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !depth_dev || !pose_M) return fail(ITM_B200_EINVAL, "NULL argument");
  if (c->sp.voxelWords == 2 && !rgb_dev)
    return fail(ITM_B200_EINVAL, "ITMVoxel_s_rgb context: use itm_b200_integrate_into_scene_rgb (needs view->rgb)");
  set_pose_host(c->hst, pose_M);
  c->hst->noVisibleEntries = rs->no_visible_entries;
  int rc = push_state(c);
  if (rc) return rc;
  IntegrateArgs a;
  a.residentList = 0;
  fill_integrate_calib(c, a, rgb_dev);
  a.depth = depth_dev;
  a.voxels = scene->voxel_blocks_dev;
  a.hashTable = scene->hash_entries_dev;
  a.visibleIds = rs->visible_entry_ids_dev;
  a.st = c->st;
  a.vp = c->vp;
  a.sp = c->sp;
  launch_integrate(a, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_integrate_into_scene(itm_b200_ctx *c, itm_b200_scene *scene, const itm_b200_render_state *rs, const float *depth_dev,
                                  const float pose_M[16]) {
  ON_DEVICE_OF_CTX(c);
  return integrate_layer_a(c, scene, rs, depth_dev, nullptr, pose_M);
}

int itm_b200_integrate_into_scene_rgb(itm_b200_ctx *c, itm_b200_scene *scene, const itm_b200_render_state *rs, const float *depth_dev,
                                      const unsigned char *rgb_dev, const float pose_M[16]) {
  ON_DEVICE_OF_CTX(c);
  if (!rgb_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  return integrate_layer_a(c, scene, rs, depth_dev, rgb_dev, pose_M);
}

int itm_b200_create_expected_depths(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                                    const float intrinsics[4]) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !pose_M || !intrinsics) return fail(ITM_B200_EINVAL, "NULL argument");
  set_pose_host(c->hst, pose_M);
  c->hst->noVisibleEntries = rs->no_visible_entries;
  int rc = push_state(c);
  if (rc) return rc;
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.shard.world = 1;
  a.hashTable = scene->hash_entries_dev;
  a.visibleIds = rs->visible_entry_ids_dev;
  a.minmax = rs->rendering_range_image_dev;
  a.st = c->st;
  a.vp = rs_view(c, rs, intrinsics);
  a.sp = c->sp;
  launch_expected_depths(a, c->stream);
  g_launches += 2;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_create_icp_maps(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, itm_b200_tracking_state *ts) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !ts) return fail(ITM_B200_EINVAL, "NULL argument");
  set_pose_host(c->hst, ts->pose_d);
  int rc = push_state(c);
  if (rc) return rc;
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.shard.world = 1;
  a.voxels = scene->voxel_blocks_dev;
  a.hashTable = scene->hash_entries_dev;
  a.minmax = rs->rendering_range_image_dev;
  a.raycastResult = rs->raycast_result_dev;
  a.raycastImage = rs->raycast_image_dev;
  a.pointsMap = ts->points_map_dev;
  a.normalsMap = ts->normals_map_dev;
  a.st = c->st;
  a.vp = c->vp;
  a.sp = c->sp;
  launch_raycast(a, c->stream);
  launch_icp_maps(a, c->stream);
  g_launches += 2;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  // trackingState->pose_pointCloud->SetFrom(trackingState->pose_d)  (ITMVisualisationEngine_CPU.cpp:273)
  memcpy(ts->pose_point_cloud, ts->pose_d, 64);
  return ITM_B200_OK;
}

// scan state of CreatePointCloud for a W x H image (kept until an image of another size comes along)
static int point_cloud_scan_state(itm_b200_ctx *c, int W, int H) {
  const int tiles = point_cloud_tiles(W, H);
  if (tiles == c->pcTiles) return ITM_B200_OK;
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(c->pcTileState);
  c->pcTileState = nullptr;
  c->pcTiles = 0;
  CU(cudaMalloc(&c->pcTileState, (size_t)(tiles + 1) * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(c->pcTileState, 0, (size_t)(tiles + 1) * sizeof(unsigned long long), c->stream));
  c->pcTiles = tiles;
  return ITM_B200_OK;
}

int itm_b200_create_point_cloud(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, itm_b200_tracking_state *ts,
                                const float inv_M[16], const float intrinsics[4], int skip_points, int *no_total_points) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !ts || !inv_M || !intrinsics || !no_total_points) return fail(ITM_B200_EINVAL, "NULL argument");
  if (!ts->points_map_dev || !ts->normals_map_dev || !rs->raycast_image_dev || !rs->raycast_result_dev || !rs->rendering_range_image_dev)
    return fail(ITM_B200_EINVAL, "CreatePointCloud needs the point cloud's locations / colours and the render state's images");
  // the caller's product pose_d->GetInvM() * trafo_rgb_to_depth.calib is used as it is (castRay and the light direction
  // read nothing else); M_d is only kept consistent with it
  memcpy(c->hst->invM_d, inv_M, 64);
  mat4_inv(inv_M, c->hst->M_d);
  int rc = push_state(c);
  if (rc) return rc;
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.shard.world = 1;
  a.voxels = scene->voxel_blocks_dev;
  a.hashTable = scene->hash_entries_dev;
  a.minmax = rs->rendering_range_image_dev;
  a.raycastResult = rs->raycast_result_dev;
  a.raycastImage = rs->raycast_image_dev;
  a.st = c->st;
  a.vp = rs_view(c, rs, intrinsics);
  a.sp = c->sp;
  rc = point_cloud_scan_state(c, a.vp.W, a.vp.H);
  if (rc) return rc;
  launch_render_image(a, rs->raycast_image_dev, ITM_B200_RENDER_SHADED_GREYSCALE, c->stream);
  launch_point_cloud(a, skip_points, ts->points_map_dev, ts->normals_map_dev, c->pcTileState, c->pcTiles, c->stream);
  g_launches += 2;
  rc = pull_state(c);
  if (rc) return rc;
  *no_total_points = c->hst->noTotalPoints;
  // trackingState->pose_pointCloud->SetFrom(trackingState->pose_d)  (ITMVisualisationEngine_CPU.cpp:250)
  memcpy(ts->pose_point_cloud, ts->pose_d, 64);
  return ITM_B200_OK;
}

int itm_b200_forward_render(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, const float *depth_dev,
                            const itm_b200_tracking_state *ts) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !depth_dev || !ts) return fail(ITM_B200_EINVAL, "NULL argument");
  if (!rs->forward_projection_dev || !rs->fwd_proj_missing_points_dev) return fail(ITM_B200_EINVAL, "render state without forward-projection buffers");
  if ((rs->img_width > 0 && rs->img_width != c->vp.W) || (rs->img_height > 0 && rs->img_height != c->vp.H))
    return fail(ITM_B200_EINVAL, "ForwardRender needs a render state of the sensor's size");
  set_pose_host(c->hst, ts->pose_d);
  int rc = push_state(c);
  if (rc) return rc;
  ForwardArgs f;
  memset(&f, 0, sizeof(f));
  f.render.shard.world = 1;
  f.render.voxels = scene->voxel_blocks_dev;
  f.render.hashTable = scene->hash_entries_dev;
  f.render.minmax = rs->rendering_range_image_dev;
  f.render.raycastResult = rs->raycast_result_dev;
  f.render.raycastImage = rs->raycast_image_dev;
  f.render.st = c->st;
  f.render.vp = c->vp;
  f.render.sp = c->sp;
  f.forwardProjection = rs->forward_projection_dev;
  f.missingPoints = rs->fwd_proj_missing_points_dev;
  f.key = c->fwdKey;
  f.depth = depth_dev;
  f.gated = 0;
  launch_forward_render(f, c->stream);
  g_launches += 4;
  rc = pull_state(c);
  if (rc) return rc;
  rs->no_fwd_proj_missing_points = c->hst->noFwdProjMissingPoints;
  return ITM_B200_OK;
}

int itm_b200_find_visible_blocks(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                                 const float intrinsics[4]) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !pose_M || !intrinsics) return fail(ITM_B200_EINVAL, "NULL argument");
  set_pose_host(c->hst, pose_M);
  c->hst->errorFlags = 0;
  int rc = push_state(c);
  if (rc) return rc;
  launch_find_visible_blocks(scene->hash_entries_dev, rs->visible_entry_ids_dev, c->st, rs_view(c, rs, intrinsics), c->sp, c->sp.nLocal,
                             c->scanTickets + 2, c->otherTileState, c->stream);
  g_launches += 1;
  rc = pull_state(c);
  if (rc) return rc;
  rs->no_visible_entries = c->hst->noVisibleEntries;
  return ITM_B200_OK;
}

static int raycast_layer_a(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                           const float intrinsics[4], unsigned char *out_image_dev, int type) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !pose_M || !intrinsics) return fail(ITM_B200_EINVAL, "NULL argument");
  set_pose_host(c->hst, pose_M);
  int rc = push_state(c);
  if (rc) return rc;
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.shard.world = 1;
  a.voxels = scene->voxel_blocks_dev;
  a.hashTable = scene->hash_entries_dev;
  a.minmax = rs->rendering_range_image_dev;
  a.raycastResult = rs->raycast_result_dev;
  a.st = c->st;
  a.vp = rs_view(c, rs, intrinsics);
  a.sp = c->sp;
  if (out_image_dev) launch_render_image(a, out_image_dev, type, c->stream);
  else launch_raycast(a, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_find_surface(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                          const float intrinsics[4]) {
  ON_DEVICE_OF_CTX(c);
  return raycast_layer_a(c, scene, rs, pose_M, intrinsics, nullptr, 0);
}

int itm_b200_render_image(itm_b200_ctx *c, const itm_b200_scene *scene, itm_b200_render_state *rs, const float pose_M[16],
                          const float intrinsics[4], unsigned char *out_image_dev, int type) {
  ON_DEVICE_OF_CTX(c);
  if (!out_image_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  if (type < 0 || type > 2) return fail(ITM_B200_EINVAL, "unknown RenderImageType");
  return raycast_layer_a(c, scene, rs, pose_M, intrinsics, out_image_dev, type);
}

// ---------------------------------------------------------------------------------------------
// swapping, Layer A
static SwapArgs swap_args_layer_a(itm_b200_ctx *c, const itm_b200_scene *scene, const itm_b200_render_state *rs, const itm_b200_swap_buffers *sw) {
  SwapArgs a;
  a.voxels = scene->voxel_blocks_dev;
  a.hashTable = scene->hash_entries_dev;
  a.vbaAllocList = scene->vba_allocation_list_dev;
  a.visType = rs ? rs->entries_visible_type_dev : nullptr;
  a.swapStates = scene->swap_states_dev;
  a.neededIds = sw->needed_entry_ids_dev;
  a.transfer = sw->synced_voxel_blocks_dev;
  a.hasSynced = sw->has_synced_data_dev;
  a.ticket = c->scanTickets + 2;
  a.tileState = c->otherTileState;
  a.st = c->st;
  a.sp = c->sp;
  a.cachePool = nullptr; a.cacheSlot = nullptr; a.cacheCount = nullptr; a.cachePoolBlocks = 0; a.movedCounts = nullptr;
  return a;
}

int itm_b200_swap_in_select(itm_b200_ctx *c, const itm_b200_scene *scene, const itm_b200_swap_buffers *sw, int *no_needed) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !sw || !no_needed) return fail(ITM_B200_EINVAL, "NULL argument");
  if (!scene->swap_states_dev) return fail(ITM_B200_EINVAL, "scene without swap states (scene->useSwapping is off)");
  c->hst->lastFreeBlockId = scene->last_free_block_id;
  int rc = push_state(c);
  if (rc) return rc;
  launch_swap_select(swap_args_layer_a(c, scene, nullptr, sw), 0, c->stream);
  g_launches += 1;
  rc = pull_state(c);
  if (rc) return rc;
  *no_needed = c->hst->swapCount;
  return ITM_B200_OK;
}

int itm_b200_swap_in_apply(itm_b200_ctx *c, itm_b200_scene *scene, const itm_b200_swap_buffers *sw, int no_needed) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !sw) return fail(ITM_B200_EINVAL, "NULL argument");
  if (no_needed <= 0) return ITM_B200_OK;
  c->hst->swapCount = no_needed;
  c->hst->lastFreeBlockId = scene->last_free_block_id;
  int rc = push_state(c);
  if (rc) return rc;
  launch_swap_in_apply(swap_args_layer_a(c, scene, nullptr, sw), c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_swap_out(itm_b200_ctx *c, itm_b200_scene *scene, const itm_b200_render_state *rs, const itm_b200_swap_buffers *sw, int *no_needed) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !rs || !sw || !no_needed) return fail(ITM_B200_EINVAL, "NULL argument");
  if (!scene->swap_states_dev) return fail(ITM_B200_EINVAL, "scene without swap states (scene->useSwapping is off)");
  c->hst->lastFreeBlockId = scene->last_free_block_id;
  int rc = push_state(c);
  if (rc) return rc;
  const SwapArgs a = swap_args_layer_a(c, scene, rs, sw);
  launch_swap_select(a, 1, c->stream);
  launch_swap_out_apply(a, c->stream);
  g_launches += 2;
  rc = pull_state(c);
  if (rc) return rc;
  *no_needed = c->hst->swapCount;
  scene->last_free_block_id = c->hst->lastFreeBlockId;
  return ITM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// meshing
static int mesh_scene_common(itm_b200_ctx *c, const void *voxels, const void *hash, float *triangles_dev, unsigned no_max_triangles,
                             unsigned *no_total_triangles) {
  ON_DEVICE_OF_CTX(c);
  if (no_max_triangles < 2) return fail(ITM_B200_EINVAL, "mesh capacity too small");
  if (!c->meshBlockList) {
    CU(cudaMalloc(&c->meshBlockList, (size_t)c->sp.nLocal * sizeof(int)));
    CU(cudaMalloc(&c->meshCounts, (size_t)c->sp.nLocal * sizeof(unsigned)));
    CU(cudaMalloc(&c->meshOffsets, ((size_t)c->sp.nLocal + 1) * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->meshSt, sizeof(FrameState)));
    CU(cudaMallocHost(&c->meshHst, sizeof(FrameState)));
  }
  memset(c->meshHst, 0, sizeof(FrameState));
  CU(cudaMemcpyAsync(c->meshSt, c->meshHst, sizeof(FrameState), cudaMemcpyHostToDevice, c->stream));
  MeshArgs a;
  a.voxels = voxels;
  a.hashTable = hash;
  a.blockList = c->meshBlockList;
  a.counts = c->meshCounts;
  a.offsets = c->meshOffsets;
  a.triangles = triangles_dev;
  a.noMaxTriangles = no_max_triangles;
  a.st = c->meshSt;
  a.sp = c->sp;
  a.ticket = c->scanTickets + 2;
  a.tileState = c->otherTileState;
  launch_mesh_scene(a, c->stream);
  g_launches += 4;
  CU(cudaMemcpyAsync(c->meshHst, c->meshSt, sizeof(FrameState), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (no_total_triangles) *no_total_triangles = (unsigned)c->meshHst->noMeshTriangles;
  return ITM_B200_OK;
}

int itm_b200_mesh_scene(itm_b200_ctx *c, const itm_b200_scene *scene, float *triangles_dev, unsigned no_max_triangles,
                        unsigned *no_total_triangles) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !scene || !triangles_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  return mesh_scene_common(c, scene->voxel_blocks_dev, scene->hash_entries_dev, triangles_dev, no_max_triangles, no_total_triangles);
}

// ITMMesh::WriteSTL (Objects/ITMMesh.h:66-118): 80 spaces, the triangle count, then per triangle a zero normal, the three
// vertices in REVERSE order (p2, p1, p0) and a zero attribute word
int itm_b200_write_stl(const char *file_name, const float *t, unsigned n) {
  if (!file_name || (!t && n)) return fail(ITM_B200_EINVAL, "NULL argument");
  FILE *f = fopen(file_name, "wb+");
  if (!f) return fail(ITM_B200_EINVAL, std::string("cannot open ") + file_name);
  char header[80];
  memset(header, ' ', sizeof(header));
  fwrite(header, 1, sizeof(header), f);
  fwrite(&n, sizeof(int), 1, f);
  std::vector<unsigned char> rec(50, 0);
  for (unsigned i = 0; i < n; ++i) {
    const float *p = t + (size_t)i * 9;
    memcpy(&rec[12], p + 6, 12);
    memcpy(&rec[24], p + 3, 12);
    memcpy(&rec[36], p + 0, 12);
    fwrite(rec.data(), 1, 50, f);
  }
  fclose(f);
  return ITM_B200_OK;
}

// ITMMesh::WriteOBJ (Objects/ITMMesh.h:34-64)
int itm_b200_write_obj(const char *file_name, const float *t, unsigned n) {
  if (!file_name || (!t && n)) return fail(ITM_B200_EINVAL, "NULL argument");
  FILE *f = fopen(file_name, "w+");
  if (!f) return fail(ITM_B200_EINVAL, std::string("cannot open ") + file_name);
  for (unsigned i = 0; i < n; ++i) {
    const float *p = t + (size_t)i * 9;
    for (int v = 0; v < 3; ++v) fprintf(f, "v %f %f %f\n", p[v * 3 + 0], p[v * 3 + 1], p[v * 3 + 2]);
  }
  for (unsigned i = 0; i < n; ++i) fprintf(f, "f %d %d %d\n", i * 3 + 2 + 1, i * 3 + 1 + 1, i * 3 + 0 + 1);
  fclose(f);
  return ITM_B200_OK;
}

int itm_b200_convert_depth_affine_to_float(itm_b200_ctx *c, float *out_dev, const short *in_dev, int w, int h, float a, float b) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !out_dev || !in_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  launch_convert_depth(in_dev, out_dev, w * h, a, b, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_convert_disparity_to_depth(itm_b200_ctx *c, float *out_dev, const short *in_dev, int w, int h, float c1, float c2, float fx_depth) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !out_dev || !in_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  if (fx_depth == 0.0f) return fail(ITM_B200_EINVAL, "fx_depth must not be 0");
  launch_convert_depth(in_dev, out_dev, w * h, c1, c2, c->stream, fx_depth);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_filter_subsample_with_holes(itm_b200_ctx *c, float *out_dev, const float *in_dev, int w_in, int h_in) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !out_dev || !in_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  launch_subsample_holes(out_dev, in_dev, w_in, h_in, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_copy_image(itm_b200_ctx *c, void *out_dev, const void *in_dev, size_t bytes) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !out_dev || !in_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  CU(cudaMemcpyAsync(out_dev, in_dev, bytes, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return ITM_B200_OK;
}

int itm_b200_filter_subsample_rgba(itm_b200_ctx *c, unsigned char *out_dev, const unsigned char *in_dev, int w_in, int h_in) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !out_dev || !in_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  if (w_in < 2 || h_in < 2) return ITM_B200_OK;  // newDims has a zero: the reference's loops do nothing
  launch_subsample_rgba(out_dev, in_dev, w_in, h_in, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_filter_subsample_with_holes_float4(itm_b200_ctx *c, float *out_dev, const float *in_dev, int w_in, int h_in) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !out_dev || !in_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  if (w_in < 2 || h_in < 2) return ITM_B200_OK;
  launch_subsample_holes4(out_dev, in_dev, w_in, h_in, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

static int gradient_common(itm_b200_ctx *c, short *grad_dev, const unsigned char *image_dev, int w, int h, int alongX) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !grad_dev || !image_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  if (w <= 0 || h <= 0) return fail(ITM_B200_EINVAL, "image size must be positive");
  CU(launch_gradient(grad_dev, image_dev, w, h, alongX, c->stream));
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}
int itm_b200_gradient_x(itm_b200_ctx *c, short *grad_dev, const unsigned char *image_dev, int w, int h) { return gradient_common(c, grad_dev, image_dev, w, h, 1); }
int itm_b200_gradient_y(itm_b200_ctx *c, short *grad_dev, const unsigned char *image_dev, int w, int h) { return gradient_common(c, grad_dev, image_dev, w, h, 0); }

static int compute_g_and_h_common(itm_b200_ctx *c, const float *level_weight_dev, const float *level_depth_dev, int w, int h, const float view_intrinsics[4],
                             const float *points_map_dev, const float *normals_map_dev, int scene_w, int scene_h,
                             const float scene_intrinsics[4], const float approx_inv_pose[16], const float scene_pose[16], float dist_thresh,
                             int iteration_type, float *f, float nabla[6], float hessian[36], int *no_valid_points) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !level_depth_dev || !points_map_dev || !normals_map_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  if (iteration_type == ITM_ITER_NONE) {
    if (no_valid_points) *no_valid_points = 0;
    return ITM_B200_OK;
  }
  memcpy(c->hst->scenePose, scene_pose, 64);
  int rc = push_state(c);
  if (rc) return rc;
  CU(cudaMemcpyAsync(c->icpPoseIn, approx_inv_pose, 64, cudaMemcpyHostToDevice, c->stream));
  IcpArgs a;
  a.pointsMap = points_map_dev;
  a.normalsMap = normals_map_dev;
  a.sceneVp.W = scene_w; a.sceneVp.H = scene_h;
  a.sceneVp.fx = scene_intrinsics[0]; a.sceneVp.fy = scene_intrinsics[1]; a.sceneVp.cx = scene_intrinsics[2]; a.sceneVp.cy = scene_intrinsics[3];
  a.st = c->st;
  a.partials = c->icpPartials;
  a.ctaCounter = c->icpCounter;
  a.terminationThreshold = c->p.depth_tracker_termination_threshold;
  IcpLevelArgs lv;
  lv.weight = level_weight_dev;
  lv.depth = level_depth_dev;
  lv.w = w; lv.h = h;
  lv.fx = view_intrinsics[0]; lv.fy = view_intrinsics[1]; lv.cx = view_intrinsics[2]; lv.cy = view_intrinsics[3];
  lv.distThresh = dist_thresh;
  lv.iterationType = iteration_type;
  launch_icp_eval_single(a, lv, c->icpOut, c->icpPoseIn, c->stream);
  g_launches += 1;
  float out[44];
  CU(cudaMemcpyAsync(out, c->icpOut, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (no_valid_points) *no_valid_points = (int)out[0];
  if (f) *f = out[1];
  if (nabla) memcpy(nabla, out + 2, 6 * sizeof(float));
  if (hessian) memcpy(hessian, out + 8, 36 * sizeof(float));
  return ITM_B200_OK;
}

int itm_b200_compute_g_and_h(itm_b200_ctx *c, const float *level_depth_dev, int w, int h, const float view_intrinsics[4],
                             const float *points_map_dev, const float *normals_map_dev, int scene_w, int scene_h,
                             const float scene_intrinsics[4], const float approx_inv_pose[16], const float scene_pose[16], float dist_thresh,
                             int iteration_type, float *f, float nabla[6], float hessian[36], int *no_valid_points) {
  ON_DEVICE_OF_CTX(c);
  return compute_g_and_h_common(c, nullptr, level_depth_dev, w, h, view_intrinsics, points_map_dev, normals_map_dev, scene_w, scene_h,
                                scene_intrinsics, approx_inv_pose, scene_pose, dist_thresh, iteration_type, f, nabla, hessian, no_valid_points);
}

int itm_b200_compute_g_and_h_weighted(itm_b200_ctx *c, const float *level_depth_dev, const float *level_weight_dev, int w, int h,
                                      const float view_intrinsics[4], const float *points_map_dev, const float *normals_map_dev, int scene_w,
                                      int scene_h, const float scene_intrinsics[4], const float approx_inv_pose[16],
                                      const float scene_pose[16], float dist_thresh, int iteration_type, float *f, float nabla[6],
                                      float hessian[36], int *no_valid_points) {
  ON_DEVICE_OF_CTX(c);
  if (!level_weight_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  return compute_g_and_h_common(c, level_weight_dev, level_depth_dev, w, h, view_intrinsics, points_map_dev, normals_map_dev, scene_w, scene_h,
                                scene_intrinsics, approx_inv_pose, scene_pose, dist_thresh, iteration_type, f, nabla, hessian, no_valid_points);
}

int itm_b200_depth_filtering(itm_b200_ctx *c, float *out_dev, const float *in_dev, int w, int h) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !out_dev || !in_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  launch_filter_depth(out_dev, in_dev, w, h, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_compute_normal_and_weights(itm_b200_ctx *c, float *normal_out_dev, float *sigma_z_out_dev, const float *depth_dev, int w, int h,
                                        const float intrinsics[4]) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !normal_out_dev || !sigma_z_out_dev || !depth_dev || !intrinsics) return fail(ITM_B200_EINVAL, "NULL argument");
  launch_normal_weight(normal_out_dev, sigma_z_out_dev, depth_dev, w, h, intrinsics, c->stream);
  g_launches += 1;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_track_camera(itm_b200_ctx *c, const float *depth_dev, itm_b200_tracking_state *ts) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !depth_dev || !ts) return fail(ITM_B200_EINVAL, "NULL argument");
  set_pose_host(c->hst, ts->pose_d);
  memcpy(c->hst->scenePose, ts->pose_point_cloud, 64);
  int rc = push_state(c);
  if (rc) return rc;
  rc = enqueue_track(c, depth_dev, ts->points_map_dev, ts->normals_map_dev, true);
  if (rc) return rc;
  rc = pull_state(c);
  if (rc) return rc;
  memcpy(ts->pose_d, c->hst->M_d, 64);
  return ITM_B200_OK;
}

int itm_b200_track_camera_weighted(itm_b200_ctx *c, const float *depth_dev, const float *depth_uncertainty_dev, itm_b200_tracking_state *ts) {
  ON_DEVICE_OF_CTX(c);
  if (!c || !depth_dev || !depth_uncertainty_dev || !ts) return fail(ITM_B200_EINVAL, "NULL argument");
  set_pose_host(c->hst, ts->pose_d);
  memcpy(c->hst->scenePose, ts->pose_point_cloud, 64);
  int rc = push_state(c);
  if (rc) return rc;
  rc = enqueue_track(c, depth_dev, ts->points_map_dev, ts->normals_map_dev, true, false, depth_uncertainty_dev);
  if (rc) return rc;
  rc = pull_state(c);
  if (rc) return rc;
  memcpy(ts->pose_d, c->hst->M_d, 64);
  return ITM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// host pose helpers
int itm_b200_mat4_inv(const float m[16], float out[16]) { return mat4_inv(m, out) ? ITM_B200_OK : ITM_B200_EINVAL; }

int itm_b200_pose_from_inv_m_coerced(const float inv_m[16], float m_out[16], float inv_out[16], float params_out[6]) {
  pose_set_invM_coerce(inv_m, m_out, params_out);
  mat4_inv_pose(m_out, inv_out);  // the routines the tracker's device loop runs (tests/test_capi_host.py pins them to the reference)
  return ITM_B200_OK;
}

int itm_b200_compute_delta(const float nabla[6], const float hessian[36], int short_iteration, float step_out[6]) {
  icp_compute_delta(step_out, nabla, hessian, short_iteration != 0);
  return ITM_B200_OK;
}

}  // extern "C"

// =============================================================================================
// Layer B

struct itm_b200_engine {
  itm_b200_ctx *c = nullptr;
  // scene
  void *voxels = nullptr;
  void *hash = nullptr;
  int *vbaAllocList = nullptr;
  int *excessAllocList = nullptr;
  // render state
  int *residentVisibleIds = nullptr;  // sharded scenes: visible entries whose block is resident on this rank
  int *visibleIds = nullptr;
  unsigned char *visType = nullptr;
  float *minmax = nullptr;
  float *raycastResult = nullptr;
  unsigned char *raycastImage = nullptr;
  float *forwardProjection = nullptr;
  int *fwdMissing = nullptr;
  // renderState_freeview (ITMMainEngine.cpp:176: created on the first free-view GetImage, with that image's size)
  int freeW = 0, freeH = 0;
  int *freeVisibleIds = nullptr;
  float *freeMinmax = nullptr;
  float *freeRaycastResult = nullptr;
  unsigned char *freeImage = nullptr;
  FrameState *stFree = nullptr;    // device: pose + visible count of the free-view camera
  FrameState *hstFree = nullptr;   // pinned
  float *meshTriangles = nullptr;      // device ITMMesh::triangles (noMaxTriangles), allocated by the first UpdateMesh
  float *cloudLocations = nullptr;     // CreatePointCloud query: Vector4f[W*H] each, sensor-sized render buffers of its own
  float *cloudColours = nullptr;
  float *cloudMinmax = nullptr;
  float *cloudRaycastResult = nullptr;
  unsigned char *cloudImage = nullptr;
  unsigned char *imageHost = nullptr;  // pinned staging for GetImage
  size_t imageHostBytes = 0;
  // tracking state
  float *points = nullptr;
  float *normals = nullptr;
  // view
  short *rawDepth = nullptr;
  unsigned char *rgb = nullptr;
  cudaStream_t copyStream = nullptr;  // view->rgb upload: not consumed by the depth-only path, kept off its critical path
  cudaEvent_t rgbDone = nullptr;
  // CreateExpectedDepths only needs the visible list and the pose, not the voxels: with ITM_B200_OVERLAP=1 it runs beside
  // IntegrateIntoScene on a second stream (fork / join by events; inside the frame graph: two parallel branches).  Measured:
  // frame 251.6 -> 249.4 us at 640x480, 436 -> 429 us at 1280x720 / 2 mm, while the integration itself gets 7 % slower from
  // sharing the SMs - a wash, so it is off by default and the stage times stay attributable.
  cudaStream_t sideStream = nullptr;
  cudaEvent_t forkEv = nullptr, joinEv = nullptr;
  bool overlapExpectedDepths = false;
  float *depth = nullptr;
  // TRACKER_WICP / useBilateralFilter: ITMViewBuilder's floatImage, view->depthNormal, view->depthUncertainty
  float *floatImage = nullptr, *depthNormal = nullptr, *depthUncertainty = nullptr;
  int agePointCloud = -1;  // host copy; its evolution does not depend on device results
  bool haveView = false;      // a frame has been given to the engine (ITMMainEngine::view != NULL)
  bool prologueDone = false;  // this frame's view kernel already did the FramePrologue chores
  ShardInfo shard;            // world == 1 unless created with itm_b200_engine_create_sharded
  // swapping (settings.useSwapping): ITMGlobalCache
  unsigned char *swapStates = nullptr;      // device, ITMHashSwapState[nEntries]
  int *neededIds = nullptr;                 // device
  void *transfer = nullptr;                 // device, syncedVoxelBlocks
  unsigned char *hasSynced = nullptr;       // device
  // the global cache: a pool of voxel blocks in host-mapped pinned memory (the swapping kernels move blocks themselves)
  char *cachePool = nullptr;                // host address
  void *cachePoolDev = nullptr;             // the same memory as the device sees it
  int cachePoolBlocks = 0;
  int *cacheSlot = nullptr;                 // device, [nEntries]: pool slot of an entry's stored block, -1 = none (hasStoredData)
  int *cacheCount = nullptr;                // device: slots handed out
  int *movedCounts = nullptr;               // device: [0] swapped in, [1] swapped out by the last frame
  // ITMGlobalCache's dense view, materialised by itm_b200_engine_global_cache only
  unsigned char *hasStoredData = nullptr;   // host, [nEntries]
  char *storedVoxelBlocks = nullptr;        // host, [nEntries] blocks (allocated lazily by the OS)
  int *cacheSlotHost = nullptr;
  unsigned long long *swapTileState = nullptr;
  unsigned long long *swapTicket = nullptr;
  int lastSwappedIn = 0, lastSwappedOut = 0;
  bool externalBuffers = false;  // voxels / raycastResult belong to the caller (sharded engines)
  unsigned barrierSeq = 0;
  // whole-frame CUDA graphs, one per (tracking on/off, profiling on/off); see enqueue_frame
  cudaGraphExec_t frameGraph[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int frameGraphLaunches[6] = {0, 0, 0, 0, 0, 0};
  bool graphsOff = false;   // swapping / sharded engines, ITM_B200_NO_GRAPH=1, or a failed capture
  bool capturing = false;
  // external pose (TRACKER_EXTERNAL / process_frame_with_pose): pinned staging ring for {M_d, invM_d, poseParams}
  float *poseStage = nullptr;          // [ITM_RESULT_RING][38]
  unsigned poseStageNext = 0;
  bool skipTrackThisFrame = false;
  // streaming API (submit_frame / wait_frame)
  FrameResult *resultRing = nullptr;     // pinned + mapped host memory, ITM_RESULT_RING slots
  FrameResult *resultRingDev = nullptr;  // its device address
  unsigned long long deviceFrameNo = 0;  // frames whose ICP-map kernel has been enqueued (= FrameState::frameNo once they ran)
  unsigned long long waitedFrameNo = 0;  // highest ticket handed back by wait_frame / drained by sync
  FrameResult saved[ITM_RESULT_RING];    // results collected early because their ring slot was about to be reused
  short *rawDepthStage[ITM_B200_MAX_IN_FLIGHT] = {nullptr};
  unsigned char *rgbStage[ITM_B200_MAX_IN_FLIGHT] = {nullptr};  // colour voxels only (integration reads view->rgb)
  cudaEvent_t h2dDone[ITM_B200_MAX_IN_FLIGHT] = {nullptr};
  cudaEvent_t stageFree[ITM_B200_MAX_IN_FLIGHT] = {nullptr};
  unsigned long long submitCount = 0;
  int profiling = 0;  // 0 off, 1 a time stamp at every stage boundary, 2 frame start and end only
  cudaEvent_t ev[9] = {nullptr};
  cudaEvent_t shardEv[4] = {nullptr};  // sharded + profiling: before the march, after it, after the barrier, after the composition
  float stageMs[8] = {0};
  size_t bytes[ITM_B200_BUF_COUNT] = {0};
};

namespace {

int engine_alloc(itm_b200_engine *e) {
  itm_b200_ctx *c = e->c;
  const size_t P = (size_t)c->vp.W * c->vp.H;
  if (!e->externalBuffers) CU(cudaMalloc(&e->voxels, (size_t)c->sp.nLocal * ITM_BLOCK_SIZE3 * 4 * c->sp.voxelWords));
  CU(cudaMalloc(&e->hash, (size_t)c->sp.nEntries * 16));
  CU(cudaMalloc(&e->vbaAllocList, (size_t)c->sp.nLocal * 4));
  CU(cudaMalloc(&e->excessAllocList, (size_t)c->sp.nExcess * 4));
  // the visible list of a sharded scene lists the whole scene's visible entries, not just this rank's pool
  const size_t visCap = (size_t)c->sp.nLocal * (e->shard.world > 1 ? e->shard.world : 1);
  CU(cudaMalloc(&e->visibleIds, visCap * 4));
  if (e->shard.world > 1) CU(cudaMalloc(&e->residentVisibleIds, (size_t)c->sp.nLocal * 4));
  CU(cudaMalloc(&e->visType, (size_t)((c->sp.nEntries + 8191) / 8192) * 8192));
  CU(cudaMalloc(&e->minmax, P * 8));
  if (!e->externalBuffers) CU(cudaMalloc(&e->raycastResult, P * 16));
  CU(cudaMalloc(&e->raycastImage, P * 4));
  CU(cudaMalloc(&e->forwardProjection, P * 16));
  CU(cudaMalloc(&e->fwdMissing, P * 4));
  CU(cudaMalloc(&e->stFree, sizeof(FrameState)));
  CU(cudaMallocHost(&e->hstFree, sizeof(FrameState)));
  CU(cudaMalloc(&e->points, P * 16));
  CU(cudaMalloc(&e->normals, P * 16));
  CU(cudaMalloc(&e->rawDepth, P * 2));
  CU(cudaMalloc(&e->rgb, P * 4));
  CU(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&e->rgbDone, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&e->sideStream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&e->forkEv, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&e->joinEv, cudaEventDisableTiming));
  e->overlapExpectedDepths = !c->p.use_swapping && e->shard.world == 1 && getenv("ITM_B200_OVERLAP") != nullptr;
  CU(cudaMalloc(&e->depth, P * 4));
  if (c->p.use_bilateral_filter) CU(cudaMalloc(&e->floatImage, P * 4));
  if (c->p.tracker_type == ITM_B200_TRACKER_WICP) {
    CU(cudaMalloc(&e->depthNormal, P * 16));
    CU(cudaMalloc(&e->depthUncertainty, P * 4));
    int rcw = ensure_weight_pyramid(c);
    if (rcw) return rcw;
  }
  for (int i = 0; i < 9; ++i) CU(cudaEventCreate(&e->ev[i]));
  if (e->shard.world > 1)
    for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&e->shardEv[i]));
  CU(cudaMallocHost(&e->poseStage, ITM_RESULT_RING * 38 * sizeof(float)));
  CU(cudaHostAlloc(&e->resultRing, ITM_RESULT_RING * sizeof(FrameResult), cudaHostAllocMapped));
  memset(e->resultRing, 0, ITM_RESULT_RING * sizeof(FrameResult));
  CU(cudaHostGetDevicePointer(&e->resultRingDev, e->resultRing, 0));
  if (c->p.use_swapping) {
    const size_t blockBytes = (size_t)ITM_BLOCK_SIZE3 * 4 * c->sp.voxelWords;
    const int numTiles = (c->sp.nEntries + 8191) / 8192;
    CU(cudaMalloc(&e->swapStates, (size_t)numTiles * 8192));
    CU(cudaMalloc(&e->neededIds, ITM_TRANSFER_BLOCK_NUM * sizeof(int)));
    CU(cudaMalloc(&e->transfer, ITM_TRANSFER_BLOCK_NUM * blockBytes));
    CU(cudaMalloc(&e->hasSynced, ITM_TRANSFER_BLOCK_NUM));
    long long poolBlocks = c->p.swap_cache_blocks > 0 ? c->p.swap_cache_blocks : 4ll * c->sp.nLocal;
    if (poolBlocks > c->sp.nEntries) poolBlocks = c->sp.nEntries;
    e->cachePoolBlocks = (int)poolBlocks;
    CU(cudaHostAlloc(&e->cachePool, (size_t)poolBlocks * blockBytes, cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer(&e->cachePoolDev, e->cachePool, 0));
    CU(cudaMalloc(&e->cacheSlot, (size_t)c->sp.nEntries * sizeof(int)));
    CU(cudaMemset(e->cacheSlot, 0xFF, (size_t)c->sp.nEntries * sizeof(int)));
    CU(cudaMalloc(&e->cacheCount, sizeof(int)));
    CU(cudaMemset(e->cacheCount, 0, sizeof(int)));
    CU(cudaMalloc(&e->movedCounts, 2 * sizeof(int)));
    CU(cudaMemset(e->movedCounts, 0, 2 * sizeof(int)));
    CU(cudaMalloc(&e->swapTileState, numTiles * sizeof(unsigned long long)));
    CU(cudaMemset(e->swapTileState, 0, numTiles * sizeof(unsigned long long)));
    CU(cudaMalloc(&e->swapTicket, sizeof(unsigned long long)));
    CU(cudaMemset(e->swapTicket, 0, sizeof(unsigned long long)));
    e->bytes[ITM_B200_BUF_SWAP_STATES] = (size_t)c->sp.nEntries;
  }
  e->bytes[ITM_B200_BUF_VOXELS] = (size_t)c->sp.nLocal * ITM_BLOCK_SIZE3 * 4 * c->sp.voxelWords;
  e->bytes[ITM_B200_BUF_RGB] = P * 4;
  e->bytes[ITM_B200_BUF_HASH] = (size_t)c->sp.nEntries * 16;
  e->bytes[ITM_B200_BUF_VBA_ALLOC_LIST] = (size_t)c->sp.nLocal * 4;
  e->bytes[ITM_B200_BUF_EXCESS_ALLOC_LIST] = (size_t)c->sp.nExcess * 4;
  e->bytes[ITM_B200_BUF_VISIBLE_IDS] = visCap * 4;
  e->bytes[ITM_B200_BUF_VISIBLE_TYPES] = (size_t)c->sp.nEntries;
  e->bytes[ITM_B200_BUF_DEPTH] = P * 4;
  e->bytes[ITM_B200_BUF_MINMAX] = P * 8;
  e->bytes[ITM_B200_BUF_RAYCAST_RESULT] = P * 16;
  e->bytes[ITM_B200_BUF_RAYCAST_IMAGE] = P * 4;
  e->bytes[ITM_B200_BUF_POINTS] = P * 16;
  e->bytes[ITM_B200_BUF_NORMALS] = P * 16;
  e->bytes[ITM_B200_BUF_RAW_DEPTH] = P * 2;
  e->bytes[ITM_B200_BUF_FORWARD_PROJECTION] = P * 16;
  e->bytes[ITM_B200_BUF_FWD_MISSING_POINTS] = P * 4;
  for (int l = 1; l <= 4; ++l)
    e->bytes[ITM_B200_BUF_PYRAMID_1 + l - 1] = l < c->nLevels ? (size_t)c->levels[l].w * c->levels[l].h * 4 : 0;
  return ITM_B200_OK;
}

void engine_free(itm_b200_engine *e) {
  if (!e->externalBuffers) { RELEASE(cudaFree(e->voxels)); RELEASE(cudaFree(e->raycastResult)); }
  RELEASE(cudaFree(e->hash)); RELEASE(cudaFree(e->vbaAllocList)); RELEASE(cudaFree(e->excessAllocList));
  RELEASE(cudaFree(e->visibleIds)); RELEASE(cudaFree(e->residentVisibleIds)); RELEASE(cudaFree(e->visType)); RELEASE(cudaFree(e->minmax));
  RELEASE(cudaFree(e->shard.unresolvedList));
  RELEASE(cudaFree(e->shard.remotePtr));
  RELEASE(cudaFree(e->raycastImage)); RELEASE(cudaFree(e->points)); RELEASE(cudaFree(e->normals)); RELEASE(cudaFree(e->rawDepth));
  RELEASE(cudaFree(e->rgb)); RELEASE(cudaFree(e->depth));
  RELEASE(cudaFree(e->floatImage)); RELEASE(cudaFree(e->depthNormal)); RELEASE(cudaFree(e->depthUncertainty));
  RELEASE(cudaFree(e->forwardProjection)); RELEASE(cudaFree(e->fwdMissing)); RELEASE(cudaFree(e->stFree));
  RELEASE(cudaFree(e->freeVisibleIds)); RELEASE(cudaFree(e->freeMinmax)); RELEASE(cudaFree(e->freeRaycastResult)); RELEASE(cudaFree(e->freeImage));
  RELEASE(cudaFree(e->cloudLocations)); RELEASE(cudaFree(e->cloudColours)); RELEASE(cudaFree(e->cloudMinmax));
  RELEASE(cudaFree(e->cloudRaycastResult)); RELEASE(cudaFree(e->cloudImage));
  if (e->hstFree) RELEASE(cudaFreeHost(e->hstFree));
  if (e->imageHost) RELEASE(cudaFreeHost(e->imageHost));
  if (e->poseStage) RELEASE(cudaFreeHost(e->poseStage));
  if (e->resultRing) RELEASE(cudaFreeHost(e->resultRing));
  for (int i = 0; i < ITM_B200_MAX_IN_FLIGHT; ++i) {
    RELEASE(cudaFree(e->rawDepthStage[i]));
    RELEASE(cudaFree(e->rgbStage[i]));
    if (e->h2dDone[i]) RELEASE(cudaEventDestroy(e->h2dDone[i]));
    if (e->stageFree[i]) RELEASE(cudaEventDestroy(e->stageFree[i]));
  }
  RELEASE(cudaFree(e->meshTriangles));
  RELEASE(cudaFree(e->swapStates)); RELEASE(cudaFree(e->neededIds)); RELEASE(cudaFree(e->transfer)); RELEASE(cudaFree(e->hasSynced));
  RELEASE(cudaFree(e->swapTileState)); RELEASE(cudaFree(e->swapTicket));
  if (e->cachePool) RELEASE(cudaFreeHost(e->cachePool));
  RELEASE(cudaFree(e->cacheSlot)); RELEASE(cudaFree(e->cacheCount)); RELEASE(cudaFree(e->movedCounts));
  free(e->hasStoredData); free(e->storedVoxelBlocks); free(e->cacheSlotHost);
  if (e->copyStream) RELEASE(cudaStreamDestroy(e->copyStream));
  if (e->rgbDone) RELEASE(cudaEventDestroy(e->rgbDone));
  if (e->sideStream) RELEASE(cudaStreamDestroy(e->sideStream));
  if (e->forkEv) RELEASE(cudaEventDestroy(e->forkEv));
  if (e->joinEv) RELEASE(cudaEventDestroy(e->joinEv));
  for (int i = 0; i < 9; ++i)
    if (e->ev[i]) RELEASE(cudaEventDestroy(e->ev[i]));
  for (int i = 0; i < 4; ++i)
    if (e->shardEv[i]) RELEASE(cudaEventDestroy(e->shardEv[i]));
  for (int i = 0; i < 6; ++i)
    if (e->frameGraph[i]) RELEASE(cudaGraphExecDestroy(e->frameGraph[i]));
  if (e->c) {
    ctx_free(e->c);
    delete e->c;
  }
}

int engine_reset(itm_b200_engine *e) {
  itm_b200_ctx *c = e->c;
  const size_t P = (size_t)c->vp.W * c->vp.H;
  launch_reset_scene(e->voxels, e->vbaAllocList, e->hash, e->excessAllocList, c->sp, c->stream);
  g_launches += 1;
  if (e->shard.remotePtr) CU(cudaMemsetAsync(e->shard.remotePtr, 0xFF, (size_t)c->sp.nEntries * sizeof(int), c->stream));
  // MemoryBlock constructors clear their memory (ORUtils/MemoryBlock.h:88-110)
  CU(cudaMemsetAsync(e->visType, 0, (size_t)((c->sp.nEntries + 8191) / 8192) * 8192, c->stream));
  CU(cudaMemsetAsync(e->visibleIds, 0, (size_t)c->sp.nLocal * 4 * (e->shard.world > 1 ? e->shard.world : 1), c->stream));
  CU(cudaMemsetAsync(e->raycastResult, 0, P * 16, c->stream));
  CU(cudaMemsetAsync(e->raycastImage, 0, P * 4, c->stream));
  CU(cudaMemsetAsync(e->forwardProjection, 0, P * 16, c->stream));
  CU(cudaMemsetAsync(e->fwdMissing, 0, P * 4, c->stream));
  CU(cudaMemsetAsync(c->fwdKey, 0, P * 4, c->stream));
  CU(cudaMemsetAsync(e->points, 0, P * 16, c->stream));
  CU(cudaMemsetAsync(e->normals, 0, P * 16, c->stream));
  CU(cudaMemsetAsync(e->minmax, 0, P * 8, c->stream));
  CU(cudaMemsetAsync(e->depth, 0, P * 4, c->stream));
  CU(cudaMemsetAsync(e->rgb, 0, P * 4, c->stream));
  if (e->floatImage) CU(cudaMemsetAsync(e->floatImage, 0, P * 4, c->stream));
  if (e->depthNormal) CU(cudaMemsetAsync(e->depthNormal, 0, P * 16, c->stream));
  if (e->depthUncertainty) CU(cudaMemsetAsync(e->depthUncertainty, 0, P * 4, c->stream));
  CU(cudaMemsetAsync(c->allocKey, 0, (size_t)((c->sp.nEntries + 8191) / 8192) * 8192 * sizeof(unsigned), c->stream));
  if (e->swapStates) {
    CU(cudaMemsetAsync(e->swapStates, 0, (size_t)((c->sp.nEntries + 8191) / 8192) * 8192, c->stream));
    CU(cudaMemsetAsync(e->cacheSlot, 0xFF, (size_t)c->sp.nEntries * sizeof(int), c->stream));
    CU(cudaMemsetAsync(e->cacheCount, 0, sizeof(int), c->stream));
    CU(cudaMemsetAsync(e->movedCounts, 0, 2 * sizeof(int), c->stream));
  }
  host_state_init(c->hst, c->sp);
  int rc = push_state(c);
  if (rc) return rc;
  e->agePointCloud = -1;
  e->deviceFrameNo = 0;
  e->waitedFrameNo = 0;
  if (e->resultRing) memset(e->resultRing, 0, ITM_RESULT_RING * sizeof(FrameResult));
  memset(e->saved, 0, sizeof(e->saved));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

// withPrologue: the whole frame follows in this stream (ProcessFrame path), so the view kernel also does the
// pose-independent first steps of the allocate and expected-depth stages
void stage_view(itm_b200_engine *e, bool withPrologue) {
  itm_b200_ctx *c = e->c;
  float *lv[ITM_MAX_LEVELS];
  lv[0] = e->depth;
  for (int l = 1; l < c->nLevels; ++l) lv[l] = c->pyramid[l];
  FramePrologue pro{c->st, e->visibleIds, e->visType, reinterpret_cast<float2 *>(e->minmax), c->vp.W * c->vp.H, c->icpEpochDev,
                    alloc_uses_lists(c->allocLists, false, e->swapStates != nullptr, e->shard.world, c->vp.W * c->vp.H, c->sp.nEntries)
                        ? c->allocLists.claimBits : nullptr};
  const float fxDisparity = c->p.depth_source == ITM_B200_DEPTH_KINECT_DISPARITY ? c->p.fx : 0.0f;
  const bool wicp = c->p.tracker_type == ITM_B200_TRACKER_WICP;
  if (!c->p.use_bilateral_filter && !wicp) {
    launch_view_pyramid(e->rawDepth, c->p.depth_calib_a, c->p.depth_calib_b, lv, c->vp.W, c->vp.H, c->nLevels, c->stream,
                        withPrologue ? &pro : nullptr, fxDisparity);
    g_launches += 1 + (c->nLevels > 5 ? c->nLevels - 5 : 0);
  } else {
    // ITMViewBuilder::UpdateView with useBilateralFilter / modelSensorNoise (ITMViewBuilder_CPU.cpp:38-63): conversion, five
    // filter passes, normals + sigma_z; then the tracker's two hierarchies (depth here, sigma_z in enqueue_track)
    const int W = c->vp.W, H = c->vp.H;
    launch_convert_depth(e->rawDepth, e->depth, W * H, c->p.depth_calib_a, c->p.depth_calib_b, c->stream, fxDisparity);
    g_launches += 1;
    if (c->p.use_bilateral_filter) {
      launch_filter_depth(e->floatImage, e->depth, W, H, c->stream);
      launch_filter_depth(e->depth, e->floatImage, W, H, c->stream);
      launch_filter_depth(e->floatImage, e->depth, W, H, c->stream);
      launch_filter_depth(e->depth, e->floatImage, W, H, c->stream);
      launch_filter_depth(e->floatImage, e->depth, W, H, c->stream);
      cudaMemcpyAsync(e->depth, e->floatImage, (size_t)W * H * 4, cudaMemcpyDeviceToDevice, c->stream);
      g_launches += 5;
    }
    if (wicp) {
      const float intr[4] = {c->vp.fx, c->vp.fy, c->vp.cx, c->vp.cy};
      launch_normal_weight(e->depthNormal, e->depthUncertainty, e->depth, W, H, intr, c->stream);
      g_launches += 1;
    }
    launch_view_pyramid(nullptr, 0.f, 0.f, lv, W, H, c->nLevels, c->stream, withPrologue ? &pro : nullptr);
    g_launches += 1 + (c->nLevels > 5 ? c->nLevels - 5 : 0);
  }
  e->prologueDone = withPrologue;
}

// trackingState->requiresFullRendering (ITMTrackingController.cpp:15).  Without useApproximateRaycast it is always true and
// nothing is launched: the render kernels are not gated then.
void stage_track_decide(itm_b200_engine *e) {
  if (!e->c->p.use_approximate_raycast) return;
  launch_track_decide(e->c->st, 1, e->c->stream);
  g_launches += 1;
}

// does this frame run the ICP tracker?  Not before the first point cloud exists (ITMTrackingController.cpp:13), not with
// ITMExternalTracker (its TrackCamera is empty) and not when the caller supplied the frame's pose
bool frame_tracks(const itm_b200_engine *e) {
  return e->agePointCloud != -1 && e->c->p.tracker_type != ITM_B200_TRACKER_EXTERNAL && !e->skipTrackThisFrame;
}

int stage_track(itm_b200_engine *e) {
  // ITMTrackingController::Track (ITMTrackingController.cpp:11-16)
  // in a whole frame the view kernel has already advanced the tracker's launch number (FramePrologue)
  int rc = ITM_B200_OK;
  if (frame_tracks(e))
    rc = enqueue_track(e->c, e->depth, e->points, e->normals, false, e->prologueDone,
                       e->c->p.tracker_type == ITM_B200_TRACKER_WICP ? e->depthUncertainty : nullptr);
  stage_track_decide(e);
  return rc;
}

void stage_allocate(itm_b200_engine *e) {
  itm_b200_ctx *c = e->c;
  AllocArgs a = make_alloc_args(c, e->depth, e->hash, e->vbaAllocList, e->excessAllocList, e->visibleIds, e->visType, 0);
  a.prologueDone = e->prologueDone ? 1 : 0;
  a.swapStates = e->swapStates;
  a.shard = e->shard;
  a.residentVisibleIds = e->residentVisibleIds;
  if (e->shard.world > 1) a.visibleCapacity = c->sp.nLocal * e->shard.world;
  launch_allocate(a, c->stream);
  g_launches += e->prologueDone ? 3 : 4;
}

void stage_integrate(itm_b200_engine *e) {
  itm_b200_ctx *c = e->c;
  IntegrateArgs a;
  fill_integrate_calib(c, a, e->rgb);
  a.depth = e->depth;
  a.voxels = e->voxels;
  a.hashTable = e->hash;
  a.visibleIds = e->residentVisibleIds ? e->residentVisibleIds : e->visibleIds;
  a.residentList = e->residentVisibleIds ? 1 : 0;
  a.st = c->st;
  a.vp = c->vp;
  a.sp = c->sp;
  // sharded scene with attached peers: nobody may still be reading this rank's voxels for the previous frame's fallback rays
  if (e->shard.world > 1 && e->shard.unresolvedList && e->barrierSeq > 0) {
    launch_shard_wait_readers(e->shard, e->barrierSeq, c->stream);
    g_launches += 1;
  }
  launch_integrate(a, c->stream);
  g_launches += 1;
}

RenderArgs engine_render_args(itm_b200_engine *e) {
  itm_b200_ctx *c = e->c;
  RenderArgs a;
  a.shard = e->shard;
  a.voxels = e->voxels;
  a.hashTable = e->hash;
  a.visibleIds = e->visibleIds;
  a.minmax = e->minmax;
  a.raycastResult = e->raycastResult;
  a.pointsMap = e->points;
  a.normalsMap = e->normals;
  a.raycastImage = e->raycastImage;
  a.minmaxReady = 0;
  a.residentList = 0;
  a.gated = c->p.use_approximate_raycast ? 1 : 0;
  a.resultRing = e->resultRingDev;
  a.st = c->st;
  a.vp = c->vp;
  a.sp = c->sp;
  return a;
}

void stage_expected_depths(itm_b200_engine *e, cudaStream_t stream = nullptr) {
  RenderArgs a = engine_render_args(e);
  // (a sharded scene renders the expected depths from ALL visible blocks - the index is replicated - so that every rank
  // marches the very ranges a single GPU would: launch_expected_depths accepts ptr = -1 entries there)
  a.minmaxReady = e->prologueDone ? 1 : 0;
  launch_expected_depths(a, stream ? stream : e->c->stream);
  g_launches += e->prologueDone ? 1 : 2;
  e->prologueDone = false;  // both consumers have run
}
void stage_shard_barrier(itm_b200_engine *e);
void stage_raycast(itm_b200_engine *e) {
  const RenderArgs a = engine_render_args(e);
  const bool stamps = e->shard.world > 1 && e->profiling == 1;
  if (stamps) cudaEventRecord(e->shardEv[0], e->c->stream);
  launch_raycast(a, e->c->stream);
  g_launches += 1;
  if (e->shard.world > 1) {
    // every rank's partial image and tile flags are complete and visible, then: nearest hit per pixel
    if (stamps) cudaEventRecord(e->shardEv[1], e->c->stream);
    stage_shard_barrier(e);
    if (stamps) cudaEventRecord(e->shardEv[2], e->c->stream);
    launch_raycast_compose(a, e->c->stream);
    g_launches += 1;
    if (e->shard.unresolvedList) {  // peers attached: the rays no rank could complete, with peer reads
      launch_raycast_fallback(a, e->barrierSeq, e->c->stream);
      g_launches += 1;
    }
    if (stamps) cudaEventRecord(e->shardEv[3], e->c->stream);
  }
}
void stage_icp_maps(itm_b200_engine *e) {
  // CreateICPMaps + the bookkeeping of ITMTrackingController::Prepare (:33-39); the copy
  // pose_pointCloud <- pose_d happens inside the kernel (FrameState::scenePose)
  launch_icp_maps(engine_render_args(e), e->c->stream);
  g_launches += 1;
  if (!e->capturing) e->deviceFrameNo++;  // the kernel advances FrameState::frameNo
  // the device copy (FrameState::agePointCloud, updated by the kernels) is the authoritative one; the host only needs to
  // know whether a point cloud exists at all (Track's "age != -1" test)
  if (e->agePointCloud == -1) e->agePointCloud = -2;
  else e->agePointCloud = 0;
}

// ITMTrackingController::Prepare's else-branch (:40-44).  gated: only when the device decided !requiresFullRendering
void stage_forward_render(itm_b200_engine *e, bool gated) {
  ForwardArgs f;
  f.render = engine_render_args(e);
  f.forwardProjection = e->forwardProjection;
  f.missingPoints = e->fwdMissing;
  f.key = e->c->fwdKey;
  f.depth = e->depth;
  f.gated = gated ? 1 : 0;
  launch_forward_render(f, e->c->stream);
  g_launches += 4;
}

// ITMSwappingEngine::IntegrateGlobalIntoLocal + SaveToGlobalMemory (ITMDenseMapper.cpp:59-64).  Unlike the rest of the
// frame this needs the host in the loop (the global cache is host memory): two list read-backs and the block transfers.
// ITMSwappingEngine::IntegrateGlobalIntoLocal + SaveToGlobalMemory (ITMDenseMapper.cpp:59-64) without the host: ordered
// selection, then the kernels combine from / copy to the host-mapped cache pool themselves (k_swap.cu)
int stage_swap(itm_b200_engine *e) {
  itm_b200_ctx *c = e->c;
  cudaStream_t s = c->stream;
  SwapArgs a;
  a.voxels = e->voxels; a.hashTable = e->hash; a.vbaAllocList = e->vbaAllocList; a.visType = e->visType; a.swapStates = e->swapStates;
  a.neededIds = e->neededIds; a.transfer = e->transfer; a.hasSynced = e->hasSynced; a.ticket = e->swapTicket; a.tileState = e->swapTileState;
  a.st = c->st; a.sp = c->sp;
  a.cachePool = e->cachePoolDev; a.cacheSlot = e->cacheSlot; a.cacheCount = e->cacheCount; a.cachePoolBlocks = e->cachePoolBlocks;
  a.movedCounts = e->movedCounts;
  launch_swap_select(a, 0, s);   // host -> active memory (ITMSwappingEngine_CPU.cpp:69-104)
  launch_swap_in_direct(a, s);
  launch_swap_select(a, 1, s);   // active memory -> host (:107-176)
  launch_swap_out_direct(a, s);
  g_launches += 4;
  return ITM_B200_OK;
}

void stage_shard_barrier(itm_b200_engine *e) {
  if (e->shard.world > 1) {
    launch_shard_barrier(e->shard, ++e->barrierSeq, e->c->stream);
    g_launches += 1;
  }
}

// stage-boundary time stamps; inside a stream capture they must become event-record NODES of the graph
void stamp(itm_b200_engine *e, int i) {
  if (!e->profiling) return;
  if (e->profiling == 2 && i != 0 && i != 8) return;  // every stamp is a node between two kernels: only the outer ones
  if (e->capturing) cudaEventRecordWithFlags(e->ev[i], e->c->stream, cudaEventRecordExternal);
  else cudaEventRecord(e->ev[i], e->c->stream);
}

int enqueue_frame_direct(itm_b200_engine *e) {
  int rc = ITM_B200_OK;
  stamp(e, 1);
  stage_view(e, true);
  stamp(e, 2);
  rc = stage_track(e);
  if (rc) return rc;
  stamp(e, 3);
  stage_allocate(e);
  stamp(e, 4);
  if (e->overlapExpectedDepths) {
    cudaStream_t s = e->c->stream;
    cudaEventRecord(e->forkEv, s);
    cudaStreamWaitEvent(e->sideStream, e->forkEv, 0);
    stage_expected_depths(e, e->sideStream);
    cudaEventRecord(e->joinEv, e->sideStream);
    stage_integrate(e);
    stamp(e, 5);
    cudaStreamWaitEvent(s, e->joinEv, 0);
    stamp(e, 6);  // "expected depths" = what is left of it after the integration has finished
  } else {
    stage_integrate(e);
    if (e->swapStates) {
      rc = stage_swap(e);  // the one stage with the host in the loop: a failed copy must not go unnoticed
      if (rc) return rc;
    }
    stamp(e, 5);
    stage_expected_depths(e);
    stamp(e, 6);
  }
  stage_raycast(e);
  if (e->c->p.use_approximate_raycast) stage_forward_render(e, true);
  stamp(e, 7);
  stage_icp_maps(e);
  stamp(e, 8);
  return ITM_B200_OK;
}

// One ProcessFrame = ~12 launches of 3..100 us each, all with launch parameters that never change (everything that varies
// from frame to frame lives in device memory: FrameState, the tracker's launch number).  So the frame is captured ONCE
// into a CUDA graph and replayed with a single cudaGraphLaunch: no per-kernel launch latency on the host and back-to-back
// scheduling on the device.  Two things do differ between frames and select the graph: whether the tracker runs (not on
// the very first frame) and whether stage time stamps are wanted.  Engines whose frame needs the host in the middle
// (swapping) or peers (sharding) keep the direct path.
int enqueue_frame(itm_b200_engine *e) {
  cudaStream_t s = e->c->stream;
  if (e->graphsOff) return enqueue_frame_direct(e);
  const int key = (frame_tracks(e) ? 1 : 0) + 2 * e->profiling;
  if (!e->frameGraph[key]) {
    const int ageBefore = e->agePointCloud;
    const unsigned long long launchesBefore = g_launches.load();
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
      e->capturing = true;
      const int rc = enqueue_frame_direct(e);
      e->capturing = false;
      ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && graph != nullptr && rc == ITM_B200_OK;
    }
    // the capture recorded the frame but ran nothing: undo its host-side bookkeeping
    e->frameGraphLaunches[key] = (int)(g_launches.load() - launchesBefore);
    g_launches -= (unsigned long long)e->frameGraphLaunches[key];
    e->agePointCloud = ageBefore;
    e->prologueDone = false;
    if (ok) ok = cudaGraphInstantiate(&e->frameGraph[key], graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      e->frameGraph[key] = nullptr;
      e->graphsOff = true;
      return enqueue_frame_direct(e);
    }
  }
  if (cudaGraphLaunch(e->frameGraph[key], s) != cudaSuccess) {
    cudaGetLastError();
    e->graphsOff = true;
    return enqueue_frame_direct(e);
  }
  g_launches += (unsigned long long)e->frameGraphLaunches[key];
  e->deviceFrameNo++;
  // what stage_icp_maps does on the host (ITMTrackingController::Prepare :35-37)
  if (e->agePointCloud == -1) e->agePointCloud = -2;
  else e->agePointCloud = 0;
  return ITM_B200_OK;
}

// trackingState->pose_d->SetM(M) for the coming frame, stream-ordered: {M_d, invM_d, poseParams} are the first 38 floats
// of FrameState.  The staging slot is pinned and only reused ITM_RESULT_RING frames later.
int enqueue_pose(itm_b200_engine *e, const float *M) {
  static_assert(offsetof(FrameState, M_d) == 0 && offsetof(FrameState, invM_d) == 64 && offsetof(FrameState, poseParams) == 128,
                "pose block of FrameState");
  FrameState tmp;
  set_pose_host(&tmp, M);
  float *slot = e->poseStage + (size_t)(e->poseStageNext++ % ITM_RESULT_RING) * 38;
  memcpy(slot, tmp.M_d, 64);
  memcpy(slot + 16, tmp.invM_d, 64);
  memcpy(slot + 32, tmp.poseParams, 24);
  CU(cudaMemcpyAsync(e->c->st, slot, 38 * sizeof(float), cudaMemcpyHostToDevice, e->c->stream));
  return ITM_B200_OK;
}

// result of frame `ticket` out of the host-mapped ring (or the early-collected copy): polls, no CUDA synchronisation
int collect_result(itm_b200_engine *e, unsigned long long ticket, FrameResult *out) {
  if (ticket == 0 || ticket > e->deviceFrameNo) return fail(ITM_B200_EINVAL, "wait_frame: no such frame has been submitted");
  if (ticket + ITM_RESULT_RING <= e->deviceFrameNo && e->saved[ticket % ITM_RESULT_RING].seq != ticket)
    return fail(ITM_B200_EINVAL, "wait_frame: the result of that frame has been overwritten");
  if (e->saved[ticket % ITM_RESULT_RING].seq == ticket) {
    *out = e->saved[ticket % ITM_RESULT_RING];
    return ITM_B200_OK;
  }
  volatile FrameResult *slot = e->resultRing + (ticket % ITM_RESULT_RING);
  unsigned long long spins = 0;
  while (true) {
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long *>(&slot->seq);
    if (seq == ticket) break;
    if ((++spins & 0xFFFF) == 0) {
      // a faulted or finished stream never publishes: do not spin forever
      const cudaError_t q = cudaStreamQuery(e->c->stream);
      if (q != cudaErrorNotReady) {
        if (q != cudaSuccess) return fail(ITM_B200_ECUDA, std::string("wait_frame: ") + cudaGetErrorString(q));
        if (*reinterpret_cast<volatile unsigned long long *>(&slot->seq) == ticket) break;
        return fail(ITM_B200_ECUDA, "wait_frame: the stream drained without publishing the frame");
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  memcpy(out, const_cast<FrameResult *>(slot), sizeof(FrameResult));
  return ITM_B200_OK;
}

}  // namespace

extern "C" {

int itm_b200_engine_create(const itm_b200_params *params, itm_b200_engine **out) {
  if (!out) return fail(ITM_B200_EINVAL, "out is NULL");
  *out = nullptr;
  itm_b200_ctx *c = nullptr;
  int rc = itm_b200_ctx_create(params, nullptr, &c);
  if (rc) return rc;
  DeviceScope deviceScope(params->device);
  itm_b200_engine *e = new itm_b200_engine();
  e->c = c;
  memset(&e->shard, 0, sizeof(e->shard));
  e->shard.world = 1;
  e->graphsOff = getenv("ITM_B200_NO_GRAPH") != nullptr;
  rc = engine_alloc(e);
  if (!rc) rc = engine_reset(e);
  if (rc) {
    engine_free(e);
    delete e;
    return rc;
  }
  *out = e;
  return ITM_B200_OK;
}

int itm_b200_engine_create_sharded(const itm_b200_params *params, const itm_b200_shard *shard, itm_b200_engine **out) {
  if (!out || !shard) return fail(ITM_B200_EINVAL, "NULL argument");
  *out = nullptr;
  if (shard->world < 1 || shard->world > ITM_MAX_SHARDS || shard->rank < 0 || shard->rank >= shard->world)
    return fail(ITM_B200_EINVAL, "rank / world out of range (at most 8 ranks)");
  if (params && (params->voxel_type != ITM_B200_VOXEL_S || params->use_swapping || params->use_approximate_raycast))
    return fail(ITM_B200_EUNSUPPORTED, "sharded engines support ITMVoxel_s without swapping and approximate raycast only");
  if (shard->axis < 0 || shard->axis > 2 || shard->thickness_blocks < 2)
    return fail(ITM_B200_EINVAL, "shard axis must be 0..2 and the slab thickness at least 2 blocks");
  for (int r = 0; r < shard->world; ++r)
    for (int q = 0; q < 2; ++q)
      if (!shard->partial_raycast_dev[q][r] || !shard->tile_hit_dev[q][r] || !shard->barrier_flags_dev[r])
        return fail(ITM_B200_EINVAL, "every rank's partial-image, tile-flag and barrier buffers must be given");
  itm_b200_ctx *c = nullptr;
  int rc = itm_b200_ctx_create(params, shard->stream, &c);
  if (rc) return rc;
  DeviceScope deviceScope(params->device);
  itm_b200_engine *e = new itm_b200_engine();
  e->c = c;
  e->graphsOff = true;
  memset(&e->shard, 0, sizeof(e->shard));
  e->shard.rank = shard->rank;
  e->shard.world = shard->world;
  e->shard.axis = shard->axis;
  e->shard.origin = shard->origin_block;
  e->shard.thickness = shard->thickness_blocks;
  e->shard.halo = shard->halo_blocks > 0 ? shard->halo_blocks : 1;
  for (int r = 0; r < shard->world; ++r) {
    for (int q = 0; q < 2; ++q) {
      e->shard.partial[q][r] = (float4 *)shard->partial_raycast_dev[q][r];
      e->shard.tileHit[q][r] = (unsigned char *)shard->tile_hit_dev[q][r];
    }
    e->shard.flags[r] = (unsigned *)shard->barrier_flags_dev[r];
  }
  rc = engine_alloc(e);
  if (!rc) rc = engine_reset(e);  // every rank resets its own copy of the replicated state
  if (rc) {
    engine_free(e);
    delete e;
    return rc;
  }
  *out = e;
  return ITM_B200_OK;
}

int itm_b200_engine_shard_export(itm_b200_engine *e, unsigned char voxels_handle[ITM_B200_IPC_HANDLE_BYTES],
                                 unsigned char hash_handle[ITM_B200_IPC_HANDLE_BYTES]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !voxels_handle || !hash_handle) return fail(ITM_B200_EINVAL, "NULL argument");
  if (e->shard.world <= 1) return fail(ITM_B200_EINVAL, "needs a sharded engine");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, e->voxels));
  memset(voxels_handle, 0, ITM_B200_IPC_HANDLE_BYTES);
  memcpy(voxels_handle, &h, sizeof(h));
  CU(cudaIpcGetMemHandle(&h, e->hash));
  memset(hash_handle, 0, ITM_B200_IPC_HANDLE_BYTES);
  memcpy(hash_handle, &h, sizeof(h));
  return ITM_B200_OK;
}

int itm_b200_engine_shard_attach(itm_b200_engine *e, void *const peer_voxels_dev[ITM_B200_MAX_SHARDS], void *const peer_hash_dev[ITM_B200_MAX_SHARDS]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !peer_voxels_dev || !peer_hash_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  if (e->shard.world <= 1) return fail(ITM_B200_EINVAL, "needs a sharded engine");
  for (int r = 0; r < e->shard.world; ++r) {
    if (r != e->shard.rank && (!peer_voxels_dev[r] || !peer_hash_dev[r])) return fail(ITM_B200_EINVAL, "every peer's voxel pool and hash table must be given");
    e->shard.peerVoxels[r] = r == e->shard.rank ? (const uint32_t *)e->voxels : (const uint32_t *)peer_voxels_dev[r];
    e->shard.peerTable[r] = r == e->shard.rank ? (const HashEntry *)e->hash : (const HashEntry *)peer_hash_dev[r];
  }
  if (!e->shard.unresolvedList) {
    CU(cudaMalloc(&e->shard.unresolvedList, (size_t)e->c->vp.W * e->c->vp.H * sizeof(int)));
    CU(cudaMalloc(&e->shard.remotePtr, (size_t)e->c->sp.nEntries * sizeof(int)));
    CU(cudaMemset(e->shard.remotePtr, 0xFF, (size_t)e->c->sp.nEntries * sizeof(int)));
  }
  return ITM_B200_OK;
}

int itm_b200_ipc_alloc(size_t bytes, void **dev_ptr, unsigned char handle[ITM_B200_IPC_HANDLE_BYTES]) {
  if (!dev_ptr || !handle || !bytes) return fail(ITM_B200_EINVAL, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= ITM_B200_IPC_HANDLE_BYTES, "handle size");
  CU(cudaMalloc(dev_ptr, bytes));
  CU(cudaMemset(*dev_ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, *dev_ptr));
  memset(handle, 0, ITM_B200_IPC_HANDLE_BYTES);
  memcpy(handle, &h, sizeof(h));
  return ITM_B200_OK;
}

int itm_b200_ipc_open(const unsigned char handle[ITM_B200_IPC_HANDLE_BYTES], void **dev_ptr) {
  if (!dev_ptr || !handle) return fail(ITM_B200_EINVAL, "NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ITM_B200_OK;
}

int itm_b200_ipc_close(void *dev_ptr) {
  CU(cudaIpcCloseMemHandle(dev_ptr));
  return ITM_B200_OK;
}

int itm_b200_ipc_free(void *dev_ptr) {
  CU(cudaFree(dev_ptr));
  return ITM_B200_OK;
}

int itm_b200_shard_owner_of_block(int x, int y, int z, int world, int axis, int origin_block, int thickness_blocks) {
  if (world < 1 || thickness_blocks < 1 || axis < 0 || axis > 2) return fail(ITM_B200_EINVAL, "world and thickness must be >= 1, axis 0..2");
  return shard_owner_of_block(x, y, z, world, axis, origin_block, thickness_blocks);
}

int itm_b200_shard_block_resident(int x, int y, int z, int rank, int world, int axis, int origin_block, int thickness_blocks) {
  return itm_b200_shard_block_resident_halo(x, y, z, rank, world, axis, origin_block, thickness_blocks, 1);
}

int itm_b200_shard_block_resident_halo(int x, int y, int z, int rank, int world, int axis, int origin_block, int thickness_blocks, int halo_blocks) {
  if (world < 1 || thickness_blocks < 1 || axis < 0 || axis > 2 || rank < 0 || rank >= world || halo_blocks < 1)
    return fail(ITM_B200_EINVAL, "world and thickness must be >= 1, axis 0..2, rank in [0, world)");
  ShardInfo sh;
  memset(&sh, 0, sizeof(sh));
  sh.rank = rank; sh.world = world; sh.axis = axis; sh.origin = origin_block; sh.thickness = thickness_blocks; sh.halo = halo_blocks;
  return shard_block_resident(x, y, z, sh) ? 1 : 0;
}

void itm_b200_engine_destroy(itm_b200_engine *e) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return;
  engine_free(e);
  delete e;
}

int itm_b200_engine_reset(itm_b200_engine *e) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  return engine_reset(e);
}

int itm_b200_engine_upload_depth(itm_b200_engine *e, const short *raw_depth_host) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !raw_depth_host) return fail(ITM_B200_EINVAL, "NULL argument");
  const size_t P = (size_t)e->c->vp.W * e->c->vp.H;
  CU(cudaMemcpyAsync(e->rawDepth, raw_depth_host, P * 2, cudaMemcpyHostToDevice, e->c->stream));
  e->haveView = true;
  return ITM_B200_OK;
}

int itm_b200_engine_sync(itm_b200_engine *e, float pose_out[16], int counters[6]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  int rc = pull_state(e->c);
  if (rc) return rc;
  e->waitedFrameNo = e->deviceFrameNo;
  const FrameState *h = e->c->hst;
  if (pose_out) memcpy(pose_out, h->M_d, 64);
  if (counters) {
    counters[0] = h->noVisibleEntries;
    counters[1] = h->lastFreeBlockId;
    counters[2] = h->lastFreeExcessId;
    counters[3] = h->allocFailures;
    counters[4] = h->errorFlags;
    counters[5] = h->icp.evalCount;
  }
  if (h->errorFlags & 1) return fail(ITM_B200_EUNSUPPORTED, "allocation ray segment longer than the supported step bound");
  return ITM_B200_OK;
}

static int process_frame_impl(itm_b200_engine *e, const unsigned char *rgb_host, const short *raw_depth_host, const float *pose_M_in,
                              bool external, float pose_out[16]) {
  if (!e || !raw_depth_host) return fail(ITM_B200_EINVAL, "NULL argument");
  cudaStream_t s = e->c->stream;
  const size_t P = (size_t)e->c->vp.W * e->c->vp.H;
  stamp(e, 0);
  // ITMViewBuilder::UpdateView: rgb + raw depth to the device (ITMViewBuilder_CUDA.cu:52-53).  The colour image is
  // not read by the ITMVoxel_s path, so it travels on a second stream while the frame is being fused; the call still
  // returns only after view->rgb is complete.
  CU(cudaMemcpyAsync(e->rawDepth, raw_depth_host, P * 2, cudaMemcpyHostToDevice, s));
  if (rgb_host) {
    // the depth image is on the frame's critical path, the (twice as large) colour image is not: it starts only when the
    // depth copy has finished, so the two do not share the host link
    if (e->c->sp.voxelWords != 2) {
      CU(cudaEventRecord(e->forkEv, s));
      CU(cudaStreamWaitEvent(e->copyStream, e->forkEv, 0));
    }
    CU(cudaMemcpyAsync(e->rgb, rgb_host, P * 4, cudaMemcpyHostToDevice, e->copyStream));
    CU(cudaEventRecord(e->rgbDone, e->copyStream));
  }
  // colour voxels read view->rgb during integration: then the frame waits for it up front
  if (rgb_host && e->c->sp.voxelWords == 2) CU(cudaStreamWaitEvent(s, e->rgbDone, 0));
  e->haveView = true;
  int rc = ITM_B200_OK;
  if (pose_M_in) rc = enqueue_pose(e, pose_M_in);
  if (rc) return rc;
  e->skipTrackThisFrame = external;
  rc = enqueue_frame(e);
  e->skipTrackThisFrame = false;
  if (rc) return rc;
  if (rgb_host && e->c->sp.voxelWords != 2) CU(cudaStreamWaitEvent(s, e->rgbDone, 0));
  return itm_b200_engine_sync(e, pose_out, nullptr);
}

int itm_b200_engine_process_frame(itm_b200_engine *e, const unsigned char *rgb_host, const short *raw_depth_host, float pose_out[16]) {
  ON_DEVICE_OF_ENGINE(e);
  return process_frame_impl(e, rgb_host, raw_depth_host, nullptr, false, pose_out);
}

int itm_b200_engine_process_frame_with_pose(itm_b200_engine *e, const unsigned char *rgb_host, const short *raw_depth_host,
                                            const float pose_M_in[16], float pose_out[16]) {
  ON_DEVICE_OF_ENGINE(e);
  return process_frame_impl(e, rgb_host, raw_depth_host, pose_M_in, true, pose_out);
}

int itm_b200_engine_submit_frame(itm_b200_engine *e, const unsigned char *rgb_host, const short *raw_depth_host,
                                 const float pose_M_in[16], unsigned long long *ticket) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !raw_depth_host || !ticket) return fail(ITM_B200_EINVAL, "NULL argument");
  if (e->swapStates || e->shard.world > 1) return fail(ITM_B200_EUNSUPPORTED, "submit_frame: swapping / sharded engines use process_frame");
  itm_b200_ctx *c = e->c;
  cudaStream_t s = c->stream;
  const size_t P = (size_t)c->vp.W * c->vp.H;
  const bool colour = c->sp.voxelWords == 2;
  // back-pressure: at most ITM_B200_MAX_IN_FLIGHT frames between the oldest one not yet waited for and this one
  const unsigned long long f = e->deviceFrameNo + 1;
  if (f > ITM_B200_MAX_IN_FLIGHT && f - ITM_B200_MAX_IN_FLIGHT > e->waitedFrameNo) {
    const unsigned long long old = f - ITM_B200_MAX_IN_FLIGHT;
    FrameResult r;
    int rc = collect_result(e, old, &r);
    if (rc) return rc;
    e->saved[old % ITM_RESULT_RING] = r;
  }
  const int slot = (int)(e->submitCount++ % ITM_B200_MAX_IN_FLIGHT);
  if (!e->rawDepthStage[0]) {
    // all staging slots at the first submit: cudaMalloc synchronises the device, it must not recur while frames are in flight
    for (int i = 0; i < ITM_B200_MAX_IN_FLIGHT; ++i) {
      CU(cudaMalloc(&e->rawDepthStage[i], P * 2));
      if (colour) CU(cudaMalloc(&e->rgbStage[i], P * 4));
      CU(cudaEventCreateWithFlags(&e->h2dDone[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&e->stageFree[i], cudaEventDisableTiming));
      CU(cudaEventRecord(e->stageFree[i], s));
    }
  }
  // upload on the copy stream into this frame's staging slot (free once the frame that used it last has copied it out) ...
  CU(cudaStreamWaitEvent(e->copyStream, e->stageFree[slot], 0));
  CU(cudaMemcpyAsync(e->rawDepthStage[slot], raw_depth_host, P * 2, cudaMemcpyHostToDevice, e->copyStream));
  if (rgb_host && colour) CU(cudaMemcpyAsync(e->rgbStage[slot], rgb_host, P * 4, cudaMemcpyHostToDevice, e->copyStream));
  CU(cudaEventRecord(e->h2dDone[slot], e->copyStream));
  // view->rgb of a depth-only scene is read by nothing on the device: it goes straight to its place, behind the depth image
  if (rgb_host && !colour) CU(cudaMemcpyAsync(e->rgb, rgb_host, P * 4, cudaMemcpyHostToDevice, e->copyStream));
  // ... and on the frame stream: staging slot -> view, then the frame
  stamp(e, 0);
  CU(cudaStreamWaitEvent(s, e->h2dDone[slot], 0));
  CU(cudaMemcpyAsync(e->rawDepth, e->rawDepthStage[slot], P * 2, cudaMemcpyDeviceToDevice, s));
  if (rgb_host && colour) CU(cudaMemcpyAsync(e->rgb, e->rgbStage[slot], P * 4, cudaMemcpyDeviceToDevice, s));
  CU(cudaEventRecord(e->stageFree[slot], s));
  e->haveView = true;
  int rc = ITM_B200_OK;
  if (pose_M_in) rc = enqueue_pose(e, pose_M_in);
  if (rc) return rc;
  e->skipTrackThisFrame = pose_M_in != nullptr;
  rc = enqueue_frame(e);
  e->skipTrackThisFrame = false;
  if (rc) return rc;
  *ticket = e->deviceFrameNo;
  return ITM_B200_OK;
}

int itm_b200_engine_wait_frame(itm_b200_engine *e, unsigned long long ticket, float pose_out[16], int counters[6]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  FrameResult r;
  int rc = collect_result(e, ticket, &r);
  if (rc) return rc;
  if (ticket > e->waitedFrameNo) e->waitedFrameNo = ticket;
  if (pose_out) memcpy(pose_out, r.M_d, 64);
  if (counters) memcpy(counters, r.counters, sizeof(r.counters));
  for (int l = 0; l < ITM_MAX_LEVELS; ++l) e->c->hst->icp.levelEvals[l] = r.levelEvals[l];
  if (r.counters[4] & 1) return fail(ITM_B200_EUNSUPPORTED, "allocation ray segment longer than the supported step bound");
  return ITM_B200_OK;
}

int itm_b200_engine_copy_to_buffer_dev(itm_b200_engine *e, int which, const void *src_dev, size_t bytes) {
  ON_DEVICE_OF_ENGINE(e);
  void *p = nullptr;
  size_t total = 0;
  int rc = itm_b200_engine_get_buffer(e, which, &p, &total);
  if (rc) return rc;
  if (!src_dev || bytes > total) return fail(ITM_B200_EINVAL, "copy_to_buffer_dev: range outside the buffer");
  CU(cudaMemcpyAsync(p, src_dev, bytes, cudaMemcpyDeviceToDevice, e->c->stream));
  return ITM_B200_OK;
}

int itm_b200_engine_get_stream(itm_b200_engine *e, void **stream) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !stream) return fail(ITM_B200_EINVAL, "NULL argument");
  *stream = (void *)e->c->stream;
  return ITM_B200_OK;
}

int itm_b200_engine_enqueue_frame_dev(itm_b200_engine *e, const short *raw_depth_dev) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !raw_depth_dev) return fail(ITM_B200_EINVAL, "NULL argument");
  cudaStream_t s = e->c->stream;
  const size_t P = (size_t)e->c->vp.W * e->c->vp.H;
  stamp(e, 0);
  if (raw_depth_dev != e->rawDepth) CU(cudaMemcpyAsync(e->rawDepth, raw_depth_dev, P * 2, cudaMemcpyDeviceToDevice, s));
  e->haveView = true;
  const int rc = enqueue_frame(e);
  if (rc) return rc;
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_engine_run_stage(itm_b200_engine *e, int stage) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  switch (stage) {
    case 0: stage_view(e, false); break;
    case 1: { int rc = stage_track(e); if (rc) return rc; } break;
    case 2: stage_allocate(e); break;
    case 3: stage_integrate(e); break;
    case 6: if (!e->swapStates) return fail(ITM_B200_EINVAL, "engine was created without use_swapping"); { int rc = stage_swap(e); if (rc) return rc; } break;
    case 4: stage_expected_depths(e); break;
    case 5: {
      // teacher forced: the full rendering is wanted, whatever the last decision was
      const int approx = e->c->p.use_approximate_raycast;
      e->c->p.use_approximate_raycast = 0;
      stage_raycast(e); stage_icp_maps(e);
      e->c->p.use_approximate_raycast = approx;
      break;
    }
    case 7: stage_forward_render(e, false); break;
    case 8: launch_track_decide(e->c->st, e->c->p.use_approximate_raycast, e->c->stream); g_launches += 1; break;
    default: return fail(ITM_B200_EINVAL, "unknown stage");
  }
  return itm_b200_engine_sync(e, nullptr, nullptr);
}

int itm_b200_engine_get_buffer(itm_b200_engine *e, int which, void **dev_ptr, size_t *bytes) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !dev_ptr) return fail(ITM_B200_EINVAL, "NULL argument");
  void *p = nullptr;
  switch (which) {
    case ITM_B200_BUF_VOXELS: p = e->voxels; break;
    case ITM_B200_BUF_HASH: p = e->hash; break;
    case ITM_B200_BUF_VBA_ALLOC_LIST: p = e->vbaAllocList; break;
    case ITM_B200_BUF_EXCESS_ALLOC_LIST: p = e->excessAllocList; break;
    case ITM_B200_BUF_VISIBLE_IDS: p = e->visibleIds; break;
    case ITM_B200_BUF_VISIBLE_TYPES: p = e->visType; break;
    case ITM_B200_BUF_DEPTH: p = e->depth; break;
    case ITM_B200_BUF_MINMAX: p = e->minmax; break;
    case ITM_B200_BUF_RAYCAST_RESULT: p = e->raycastResult; break;
    case ITM_B200_BUF_RAYCAST_IMAGE: p = e->raycastImage; break;
    case ITM_B200_BUF_POINTS: p = e->points; break;
    case ITM_B200_BUF_NORMALS: p = e->normals; break;
    case ITM_B200_BUF_RAW_DEPTH: p = e->rawDepth; break;
    case ITM_B200_BUF_RGB: p = e->rgb; break;
    case ITM_B200_BUF_SWAP_STATES: p = e->swapStates; break;
    case ITM_B200_BUF_FORWARD_PROJECTION: p = e->forwardProjection; break;
    case ITM_B200_BUF_FWD_MISSING_POINTS: p = e->fwdMissing; break;
    case ITM_B200_BUF_FREEVIEW_VISIBLE_IDS: p = e->freeVisibleIds; break;
    case ITM_B200_BUF_FREEVIEW_MINMAX: p = e->freeMinmax; break;
    case ITM_B200_BUF_FREEVIEW_RAYCAST_RESULT: p = e->freeRaycastResult; break;
    case ITM_B200_BUF_FREEVIEW_IMAGE: p = e->freeImage; break;
    case ITM_B200_BUF_PYRAMID_1: case ITM_B200_BUF_PYRAMID_2: case ITM_B200_BUF_PYRAMID_3: case ITM_B200_BUF_PYRAMID_4:
      p = e->c->pyramid[which - ITM_B200_BUF_PYRAMID_1 + 1]; break;
    default: return fail(ITM_B200_EINVAL, "unknown buffer id");
  }
  *dev_ptr = p;
  if (bytes) *bytes = e->bytes[which];
  return ITM_B200_OK;
}

int itm_b200_engine_read_buffer(itm_b200_engine *e, int which, void *host_dst, size_t bytes, size_t offset) {
  ON_DEVICE_OF_ENGINE(e);
  void *p = nullptr;
  size_t total = 0;
  int rc = itm_b200_engine_get_buffer(e, which, &p, &total);
  if (rc) return rc;
  if (!host_dst || offset + bytes > total) return fail(ITM_B200_EINVAL, "read_buffer: range outside the buffer");
  CU(cudaMemcpyAsync(host_dst, (const char *)p + offset, bytes, cudaMemcpyDeviceToHost, e->c->stream));
  CU(cudaStreamSynchronize(e->c->stream));
  return ITM_B200_OK;
}

int itm_b200_engine_write_buffer(itm_b200_engine *e, int which, const void *host_src, size_t bytes, size_t offset) {
  ON_DEVICE_OF_ENGINE(e);
  void *p = nullptr;
  size_t total = 0;
  int rc = itm_b200_engine_get_buffer(e, which, &p, &total);
  if (rc) return rc;
  if (!host_src || offset + bytes > total) return fail(ITM_B200_EINVAL, "write_buffer: range outside the buffer");
  CU(cudaMemcpyAsync((char *)p + offset, host_src, bytes, cudaMemcpyHostToDevice, e->c->stream));
  CU(cudaStreamSynchronize(e->c->stream));
  return ITM_B200_OK;
}

int itm_b200_engine_global_cache(itm_b200_engine *e, const unsigned char **has_stored_data, const void **stored_voxel_blocks,
                                 int *swapped_in, int *swapped_out) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  if (!e->swapStates) return fail(ITM_B200_EINVAL, "engine was created without use_swapping");
  itm_b200_ctx *c = e->c;
  const size_t blockBytes = (size_t)ITM_BLOCK_SIZE3 * 4 * c->sp.voxelWords;
  int moved[2] = {0, 0};
  if (!has_stored_data && !stored_voxel_blocks) {  // the counts only
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(moved, e->movedCounts, sizeof(moved), cudaMemcpyDeviceToHost));
    if (swapped_in) *swapped_in = moved[0];
    if (swapped_out) *swapped_out = moved[1];
    return ITM_B200_OK;
  }
  if (!e->hasStoredData) {
    e->hasStoredData = (unsigned char *)calloc(c->sp.nEntries, 1);
    e->storedVoxelBlocks = (char *)malloc((size_t)c->sp.nEntries * blockBytes);  // touched only where a block is stored
    e->cacheSlotHost = (int *)malloc((size_t)c->sp.nEntries * sizeof(int));
    if (!e->hasStoredData || !e->storedVoxelBlocks || !e->cacheSlotHost) return fail(ITM_B200_ECUDA, "host memory for the global cache view");
  }
  // the dense ITMGlobalCache view of the pool: every pending frame has to be through with it first
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(e->cacheSlotHost, e->cacheSlot, (size_t)c->sp.nEntries * sizeof(int), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(moved, e->movedCounts, sizeof(moved), cudaMemcpyDeviceToHost));
  for (int id = 0; id < c->sp.nEntries; ++id) {
    const int slot = e->cacheSlotHost[id];
    e->hasStoredData[id] = slot >= 0 ? 1 : 0;
    if (slot >= 0) memcpy(e->storedVoxelBlocks + (size_t)id * blockBytes, e->cachePool + (size_t)slot * blockBytes, blockBytes);
  }
  e->lastSwappedIn = moved[0];
  e->lastSwappedOut = moved[1];
  if (has_stored_data) *has_stored_data = e->hasStoredData;
  if (stored_voxel_blocks) *stored_voxel_blocks = e->storedVoxelBlocks;
  if (swapped_in) *swapped_in = e->lastSwappedIn;
  if (swapped_out) *swapped_out = e->lastSwappedOut;
  return ITM_B200_OK;
}

int itm_b200_engine_get_state(itm_b200_engine *e, float pose_d[16], float pose_point_cloud[16], int state6[6]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  int rc = pull_state(e->c);
  if (rc) return rc;
  const FrameState *h = e->c->hst;
  if (pose_d) memcpy(pose_d, h->M_d, 64);
  if (pose_point_cloud) memcpy(pose_point_cloud, h->scenePose, 64);
  if (state6) {
    state6[0] = h->noVisibleEntries;
    state6[1] = h->lastFreeBlockId;
    state6[2] = h->lastFreeExcessId;
    state6[3] = h->agePointCloud;
    state6[4] = h->requiresFullRendering;
    state6[5] = h->noFwdProjMissingPoints;
  }
  return ITM_B200_OK;
}

int itm_b200_engine_set_state(itm_b200_engine *e, const float pose_d[16], const float pose_point_cloud[16], const int state6[6]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  int rc = pull_state(e->c);
  if (rc) return rc;
  FrameState *h = e->c->hst;
  if (pose_d) set_pose_host(h, pose_d);
  if (pose_point_cloud) memcpy(h->scenePose, pose_point_cloud, 64);
  if (state6) {
    h->noVisibleEntries = state6[0];
    h->lastFreeBlockId = state6[1];
    h->lastFreeExcessId = state6[2];
    e->agePointCloud = state6[3];
    h->agePointCloud = state6[3];
  }
  rc = push_state(e->c);
  if (rc) return rc;
  CU(cudaStreamSynchronize(e->c->stream));
  return ITM_B200_OK;
}

// IITMVisualisationEngine::DepthToUchar4 (ITMLib/Engine/ITMVisualisationEngine.cpp:7-57) is host code in the reference for
// every device type (it runs after UpdateHostFromDevice); same here.  Colour ramp over the valid depth range.
static float ramp_segment(float v, float y0, float x0, float y1, float x1) { return (v - x0) * (y1 - y0) / (x1 - x0) + y0; }
static float ramp_base(float v) {
  if (v <= -0.75f) return 0.0f;
  if (v <= -0.25f) return ramp_segment(v, 0.0f, -0.75f, 1.0f, -0.25f);
  if (v <= 0.25f) return 1.0f;
  if (v <= 0.75f) return ramp_segment(v, 1.0f, 0.25f, 0.0f, 0.75f);
  return 0.0f;
}
static void depth_to_uchar4(unsigned char *dst, const float *src, size_t n) {
  memset(dst, 0, n * 4);
  float lo = 100000.0f, hi = -100000.0f;
  for (size_t i = 0; i < n; ++i) {
    const float v = src[i];
    if (v > 0.0f) {
      if (v < lo) lo = v;
      if (v > hi) hi = v;
    }
  }
  const float scale = ((hi - lo) != 0) ? 1.0f / (hi - lo) : 1.0f / hi;
  if (lo == hi) return;
  for (size_t i = 0; i < n; ++i) {
    float v = src[i];
    if (v > 0.0f) {
      v = (v - lo) * scale;
      dst[i * 4 + 0] = (unsigned char)(ramp_base(v - 0.5f) * 255.0f);
      dst[i * 4 + 1] = (unsigned char)(ramp_base(v) * 255.0f);
      dst[i * 4 + 2] = (unsigned char)(ramp_base(v + 0.5f) * 255.0f);
      dst[i * 4 + 3] = 255;
    }
  }
}

int itm_b200_engine_get_image(itm_b200_engine *e, int image_type, const float pose_M[16], const float intrinsics[4],
                              unsigned char *out_host, int out_w, int out_h) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !out_host || out_w <= 0 || out_h <= 0) return fail(ITM_B200_EINVAL, "NULL argument");
  if (!e->haveView) return ITM_B200_OK;  // "if (view == NULL) return;"
  itm_b200_ctx *c = e->c;
  cudaStream_t s = c->stream;
  const size_t P = (size_t)c->vp.W * c->vp.H, N = (size_t)out_w * out_h;
  const size_t need = (P > N ? P : N) * 4;
  if (e->imageHostBytes < need) {
    if (e->imageHost) cudaFreeHost(e->imageHost);
    e->imageHost = nullptr;
    e->imageHostBytes = 0;
    CU(cudaMallocHost(&e->imageHost, need));
    e->imageHostBytes = need;
  }
  const bool sensorSized = out_w == c->vp.W && out_h == c->vp.H;
  switch (image_type) {
    case ITM_B200_IMAGE_ORIGINAL_RGB:
    case ITM_B200_IMAGE_SCENERAYCAST:
      if (!sensorSized) return fail(ITM_B200_EINVAL, "GetImage: this image type has the sensor's size");
      CU(cudaMemcpyAsync(out_host, image_type == ITM_B200_IMAGE_ORIGINAL_RGB ? e->rgb : e->raycastImage, P * 4, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      return ITM_B200_OK;
    case ITM_B200_IMAGE_ORIGINAL_DEPTH:
      if (!sensorSized) return fail(ITM_B200_EINVAL, "GetImage: this image type has the sensor's size");
      CU(cudaMemcpyAsync(e->imageHost, e->depth, P * 4, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      depth_to_uchar4(out_host, reinterpret_cast<const float *>(e->imageHost), P);
      return ITM_B200_OK;
    case ITM_B200_IMAGE_FREECAMERA_SHADED:
    case ITM_B200_IMAGE_FREECAMERA_COLOUR_FROM_VOLUME:
    case ITM_B200_IMAGE_FREECAMERA_COLOUR_FROM_NORMAL:
      break;
    default:
      memset(out_host, 0, N * 4);  // out->Clear() is all InfiniTAM_IMAGE_UNKNOWN does
      return ITM_B200_OK;
  }
  if (!pose_M || !intrinsics) return fail(ITM_B200_EINVAL, "free-view images need a pose and intrinsics");
  if (e->freeW != out_w || e->freeH != out_h) {  // visualisationEngine->CreateRenderState(out->noDims)
    cudaFree(e->freeVisibleIds); cudaFree(e->freeMinmax); cudaFree(e->freeRaycastResult); cudaFree(e->freeImage);
    e->freeVisibleIds = nullptr; e->freeMinmax = nullptr; e->freeRaycastResult = nullptr; e->freeImage = nullptr;
    e->freeW = e->freeH = 0;
    CU(cudaMalloc(&e->freeVisibleIds, (size_t)c->sp.nLocal * 4));
    CU(cudaMalloc(&e->freeMinmax, N * 8));
    CU(cudaMalloc(&e->freeRaycastResult, N * 16));
    CU(cudaMalloc(&e->freeImage, N * 4));
    CU(cudaMemsetAsync(e->freeVisibleIds, 0, (size_t)c->sp.nLocal * 4, s));
    CU(cudaMemsetAsync(e->freeRaycastResult, 0, N * 16, s));
    CU(cudaMemsetAsync(e->freeImage, 0, N * 4, s));
    e->freeW = out_w;
    e->freeH = out_h;
    e->bytes[ITM_B200_BUF_FREEVIEW_VISIBLE_IDS] = (size_t)c->sp.nLocal * 4;
    e->bytes[ITM_B200_BUF_FREEVIEW_MINMAX] = N * 8;
    e->bytes[ITM_B200_BUF_FREEVIEW_RAYCAST_RESULT] = N * 16;
    e->bytes[ITM_B200_BUF_FREEVIEW_IMAGE] = N * 4;
  }
  memset(e->hstFree, 0, sizeof(FrameState));
  set_pose_host(e->hstFree, pose_M);
  CU(cudaMemcpyAsync(e->stFree, e->hstFree, sizeof(FrameState), cudaMemcpyHostToDevice, s));
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.shard.world = 1;
  a.voxels = e->voxels;
  a.hashTable = e->hash;
  a.visibleIds = e->freeVisibleIds;
  a.minmax = e->freeMinmax;
  a.raycastResult = e->freeRaycastResult;
  a.st = e->stFree;
  a.vp.W = out_w; a.vp.H = out_h;
  a.vp.fx = intrinsics[0]; a.vp.fy = intrinsics[1]; a.vp.cx = intrinsics[2]; a.vp.cy = intrinsics[3];
  a.sp = c->sp;
  launch_find_visible_blocks(e->hash, e->freeVisibleIds, e->stFree, a.vp, c->sp, c->sp.nLocal, c->scanTickets + 2, c->otherTileState, s);
  launch_expected_depths(a, s);
  const int type = image_type == ITM_B200_IMAGE_FREECAMERA_COLOUR_FROM_VOLUME ? ITM_B200_RENDER_COLOUR_FROM_VOLUME
                 : image_type == ITM_B200_IMAGE_FREECAMERA_COLOUR_FROM_NORMAL ? ITM_B200_RENDER_COLOUR_FROM_NORMAL
                                                                              : ITM_B200_RENDER_SHADED_GREYSCALE;
  launch_render_image(a, e->freeImage, type, s);
  g_launches += 4;
  CU(cudaMemcpyAsync(out_host, e->freeImage, N * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaGetLastError());
  return ITM_B200_OK;
}

int itm_b200_engine_create_point_cloud(itm_b200_engine *e, const float trafo_rgb_to_depth[16], const float intrinsics_rgb[4], int skip_points,
                                       float *locations_host, float *colours_host, int capacity_points, unsigned char *image_host,
                                       int *no_total_points) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !no_total_points) return fail(ITM_B200_EINVAL, "NULL argument");
  if ((locations_host || colours_host) && capacity_points < 0) return fail(ITM_B200_EINVAL, "capacity_points must not be negative");
  itm_b200_ctx *c = e->c;
  if (e->shard.world > 1) return fail(ITM_B200_EINVAL, "CreatePointCloud is not provided for sharded scenes");
  cudaStream_t s = c->stream;
  const size_t P = (size_t)c->vp.W * c->vp.H;
  if (!e->cloudLocations) {
    CU(cudaMalloc(&e->cloudLocations, P * 16));
    CU(cudaMalloc(&e->cloudColours, P * 16));
    CU(cudaMalloc(&e->cloudMinmax, P * 8));  // full-size allocation like ITMRenderState::renderingRangeImage
    CU(cudaMalloc(&e->cloudRaycastResult, P * 16));
    CU(cudaMalloc(&e->cloudImage, P * 4));
  }
  int rc = pull_state(c);  // the tracked pose and the visible list's length
  if (rc) return rc;
  e->waitedFrameNo = e->deviceFrameNo;
  // ITMExtrinsics::SetFrom (Objects/ITMExtrinsics.h:32-42): calib_inv = [R^T | -R^T t], the translation accumulated by subtraction
  float calib[16], calibInv[16];
  if (trafo_rgb_to_depth) memcpy(calib, trafo_rgb_to_depth, 64);
  else for (int i = 0; i < 16; ++i) calib[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  for (int i = 0; i < 16; ++i) calibInv[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 3; ++col) calibInv[r + 4 * col] = calib[col + 4 * r];
  for (int r = 0; r < 3; ++r) {
    float d = 0.0f;
    for (int col = 0; col < 3; ++col) d -= calib[col + 4 * r] * calib[col + 4 * 3];
    calibInv[r + 4 * 3] = d;
  }
  // Prepare (ITMTrackingController.cpp:24-26): pose_rgb = calib_inv * pose_d->GetM() for the ray ranges;
  // CreatePointCloud_common (ITMVisualisationEngine_CPU.cpp:247): invM = pose_d->GetInvM() * calib for the rays
  memset(e->hstFree, 0, sizeof(FrameState));
  mat4_mul(calibInv, c->hst->M_d, e->hstFree->M_d);
  mat4_mul(c->hst->invM_d, calib, e->hstFree->invM_d);
  e->hstFree->noVisibleEntries = c->hst->noVisibleEntries;
  CU(cudaMemcpyAsync(e->stFree, e->hstFree, sizeof(FrameState), cudaMemcpyHostToDevice, s));
  RenderArgs a;
  memset(&a, 0, sizeof(a));
  a.shard.world = 1;
  a.voxels = e->voxels;
  a.hashTable = e->hash;
  a.visibleIds = e->visibleIds;
  a.minmax = e->cloudMinmax;
  a.raycastResult = e->cloudRaycastResult;
  a.raycastImage = e->cloudImage;
  a.st = e->stFree;
  a.vp = c->vp;
  if (intrinsics_rgb) { a.vp.fx = intrinsics_rgb[0]; a.vp.fy = intrinsics_rgb[1]; a.vp.cx = intrinsics_rgb[2]; a.vp.cy = intrinsics_rgb[3]; }
  a.sp = c->sp;
  rc = point_cloud_scan_state(c, a.vp.W, a.vp.H);
  if (rc) return rc;
  launch_expected_depths(a, s);
  launch_render_image(a, e->cloudImage, ITM_B200_RENDER_SHADED_GREYSCALE, s);
  launch_point_cloud(a, skip_points, e->cloudLocations, e->cloudColours, c->pcTileState, c->pcTiles, s);
  g_launches += 4;
  CU(cudaMemcpyAsync(e->hstFree, e->stFree, sizeof(FrameState), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaGetLastError());
  const int n = e->hstFree->noTotalPoints;
  *no_total_points = n;
  const size_t k = (size_t)(n < capacity_points ? n : capacity_points);
  if (locations_host && k) CU(cudaMemcpy(locations_host, e->cloudLocations, k * 16, cudaMemcpyDeviceToHost));
  if (colours_host && k) CU(cudaMemcpy(colours_host, e->cloudColours, k * 16, cudaMemcpyDeviceToHost));
  if (image_host) CU(cudaMemcpy(image_host, e->cloudImage, P * 4, cudaMemcpyDeviceToHost));
  return ITM_B200_OK;
}

int itm_b200_engine_mesh_scene(itm_b200_engine *e, float *triangles_host, unsigned capacity_triangles, unsigned *no_total_triangles) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  itm_b200_ctx *c = e->c;
  const unsigned noMax = (unsigned)c->sp.nLocal * 32u;  // ITMMesh::noMaxTriangles (Objects/ITMMesh.h:22)
  if (!e->meshTriangles) CU(cudaMalloc(&e->meshTriangles, (size_t)noMax * 36));
  unsigned n = 0;
  int rc = mesh_scene_common(c, e->voxels, e->hash, e->meshTriangles, noMax, &n);
  if (rc) return rc;
  if (no_total_triangles) *no_total_triangles = n;
  if (triangles_host) {
    const unsigned k = n < capacity_triangles ? n : capacity_triangles;
    if (k) CU(cudaMemcpy(triangles_host, e->meshTriangles, (size_t)k * 36, cudaMemcpyDeviceToHost));
  }
  return ITM_B200_OK;
}

int itm_b200_engine_save_scene_to_mesh(itm_b200_engine *e, const char *file_name) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !file_name) return fail(ITM_B200_EINVAL, "NULL argument");
  unsigned n = 0;
  int rc = itm_b200_engine_mesh_scene(e, nullptr, 0, &n);
  if (rc) return rc;
  std::vector<float> host((size_t)n * 9);
  if (n) CU(cudaMemcpy(host.data(), e->meshTriangles, (size_t)n * 36, cudaMemcpyDeviceToHost));
  return itm_b200_write_stl(file_name, host.data(), n);
}

int itm_b200_engine_icp_stats(itm_b200_engine *e, int evals_per_level[ITM_B200_MAX_LEVELS]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !evals_per_level) return fail(ITM_B200_EINVAL, "NULL argument");
  for (int l = 0; l < ITM_B200_MAX_LEVELS; ++l) evals_per_level[l] = e->c->hst->icp.levelEvals[l];
  return ITM_B200_OK;
}

int itm_b200_engine_set_profiling(itm_b200_engine *e, int on) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e) return fail(ITM_B200_EINVAL, "NULL engine");
  e->profiling = on < 0 ? 0 : (on > 2 ? 2 : on);
  return ITM_B200_OK;
}

int itm_b200_engine_shard_times(itm_b200_engine *e, float ms3[3]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !ms3) return fail(ITM_B200_EINVAL, "NULL argument");
  if (e->shard.world <= 1 || e->profiling != 1) return fail(ITM_B200_EINVAL, "needs a sharded engine with set_profiling(1)");
  for (int i = 0; i < 3; ++i) CU(cudaEventElapsedTime(&ms3[i], e->shardEv[i], e->shardEv[i + 1]));
  return ITM_B200_OK;
}

int itm_b200_engine_shard_unresolved(itm_b200_engine *e, int *pixels) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !pixels) return fail(ITM_B200_EINVAL, "NULL argument");
  if (e->shard.world <= 1) return fail(ITM_B200_EINVAL, "needs a sharded engine");
  int rc = pull_state(e->c);
  if (rc) return rc;
  // the frame that incremented frameNo to its current value composed with parity (frameNo - 1) & 1
  *pixels = e->c->hst->shardUnresolved[(e->c->hst->frameNo - 1) & 1];
  return ITM_B200_OK;
}

int itm_b200_engine_stage_times(itm_b200_engine *e, float ms8[8]) {
  ON_DEVICE_OF_ENGINE(e);
  if (!e || !ms8) return fail(ITM_B200_EINVAL, "NULL argument");
  if (!e->profiling) return fail(ITM_B200_EINVAL, "profiling is off");
  // ev[0] frame start, ev[1] after H2D, ev[2] after view, ev[3] track, ev[4] allocate, ev[5] integrate,
  // ev[6] expected depths, ev[7] raycast, ev[8] icp maps
  float t;
  for (int i = 0; i < 8; ++i) ms8[i] = 0.0f;
  if (e->profiling == 1) {
    cudaEventElapsedTime(&t, e->ev[0], e->ev[2]); ms8[0] = t;
    for (int i = 1; i < 7; ++i) { cudaEventElapsedTime(&t, e->ev[i + 1], e->ev[i + 2]); ms8[i] = t; }
  }
  cudaEventElapsedTime(&t, e->ev[0], e->ev[8]); ms8[7] = t;
  return ITM_B200_OK;
}

}  // extern "C"
