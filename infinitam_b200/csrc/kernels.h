// Host-side launch interface of the CUDA kernels (internal to libitm_b200.so).
#pragma once
#include <cuda_runtime.h>
#include "itm_common.cuh"

#include <cstdlib>
#include <utility>

namespace itm {

// Launch of a kernel that belongs to a frame's chain: with programmatic stream serialisation the kernel's CTAs may be placed
// while its predecessor in the stream drains (itm_common.cuh, pdl_wait / pdl_trigger; a stream capture turns this into a
// programmatic edge of the frame graph).  ITM_B200_PDL=0 launches plainly (A/B measurements).
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("ITM_B200_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#define ITM_MAX_SHARDS 8

// Spatial sharding of ONE scene across the GPUs of an NVLink domain (DESIGN.md "Multi-GPU").
//  * The INDEX (hash table positions / offsets, excess list, visible list, pose, images) is replicated: every rank runs the
//    same deterministic allocation on the same broadcast depth frame, so positions and chain links are identical everywhere
//    and identical to a single-GPU run.
//  * The voxel PAYLOAD is not: a block's voxels exist only on the ranks where the block is RESIDENT - its owner (slab of
//    block coordinates along one axis) and, for the one-block halo the trilinear taps / normals of a ray cast need, the
//    neighbouring slab's rank.  Everywhere else the entry carries ptr = -1, the reference's own "allocated, but the payload
//    is not in active memory" state (ITMHashEntry::ptr, ITMLibDefines.h:78-81), which every kernel already honours.
//  * Each rank integrates its resident blocks (halo blocks redundantly: integration is a pure function of block, depth
//    and pose), renders the expected depths from ALL visible blocks and marches ALL rays over those ranges with its own voxels
//    (k_raycast_sharded).  A ray that never sampled a block held elsewhere is complete: its result is the single GPU's, bit for
//    bit.  k_raycast_compose takes any complete result per pixel (pulling peers' tiles over NVLink), k_raycast_fallback marches
//    the rays no rank could complete with peer reads of the voxels held elsewhere.  ICP maps and the tracker then run
//    replicated on the composed image - no pose broadcast, no G/H all-reduce is needed for the ranks to stay in step.
struct ShardInfo {
  int rank, world;                    // world == 1: single GPU, everything below unused
  int axis;                           // 0 / 1 / 2: block coordinate the slabs are cut along
  int origin, thickness;              // owner = clamp((coord - origin) / thickness, 0, world - 1)   (floor division)
  int halo;                           // blocks beyond its slab a rank keeps resident (>= 1)
  float4 *partial[2][ITM_MAX_SHARDS];        // per frame parity: every rank's partial raycast image ([rank] = the local one)
  unsigned char *tileHit[2][ITM_MAX_SHARDS]; // ... and its per-tile "contains a hit" flags (16x8-pixel tiles, raster order)
  unsigned *flags[ITM_MAX_SHARDS];    // every rank's barrier words: flags[r][src] = last barrier number src has reached
                                      // (flags[r][16 + src]: the frame up to which src is through with reading peers' voxels)
  // set by itm_b200_engine_shard_attach (else NULL: rays no rank can complete stay misses): every rank's voxel pool and hash
  // table ([rank] = the local ones), and this rank's list of unresolved pixels
  const uint32_t *peerVoxels[ITM_MAX_SHARDS];
  const HashEntry *peerTable[ITM_MAX_SHARDS];
  int *unresolvedList;
  int *remotePtr;                     // [nEntries] the owner's block number of an entry held elsewhere, once it has been looked up (-1: not yet)
};

__host__ __device__ __forceinline__ int shard_floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
// rank that owns the voxel block at block coordinate (x, y, z)
__host__ __device__ __forceinline__ int shard_owner_of_block(int x, int y, int z, int world, int axis, int origin, int thickness) {
  const int c = axis == 0 ? x : (axis == 1 ? y : z);
  int o = shard_floor_div(c - origin, thickness);
  return o < 0 ? 0 : (o >= world ? world - 1 : o);
}
// is the block's payload kept on `rank`?  (owned, or within one block of the rank's slab)
__host__ __device__ __forceinline__ bool shard_block_resident(int x, int y, int z, const ShardInfo &sh) {
  if (sh.world <= 1) return true;
  const int c = sh.axis == 0 ? x : (sh.axis == 1 ? y : z);
  const int lo = sh.origin + sh.rank * sh.thickness, hi = lo + sh.thickness;  // owned: [lo, hi), open-ended for the outer ranks
  const bool aboveLo = sh.rank == 0 || c >= lo - sh.halo;
  const bool belowHi = sh.rank == sh.world - 1 || c < hi + sh.halo;
  return aboveLo && belowHi;
}

// Scratch of the list-based allocation (k_alloc.cu, "compact lists"): empty / zero between frames.  The hash slots are cut
// into bins of ITM_ALLOC_BIN consecutive slots; a bin's lists hold at most one item per slot, so ITM_ALLOC_BIN items each.
#define ITM_ALLOC_BIN 8192
struct AllocLists {
  unsigned *claimBits;      // one bit per hash slot: a thread of this frame's per-pixel pass has taken care of its visible type
  int *reqList;             // [bins][ITM_ALLOC_BIN] slots that received an allocation request this frame (bit 31: excess-list request)
  int *newVisList;          // [bins][ITM_ALLOC_BIN] slots that became visible this frame
  int *prevCopy;            // [visibleCapacity] last frame's visible list (ascending)
  unsigned char *prevKeep;  // [visibleCapacity] 1 = that entry stays in the list
  int *binCounts;           // [5][bins]: requests, excess-list requests, newly visible, previous entries, kept previous entries
  int *done;                // CTAs of the merge kernel that have read the counts (the last one clears them)
  int numBins;
};

// Does an allocation pass with these properties take the list-based kernels (else the whole-table scans)?  The lists cost
// per ray-segment step that touches the table (a claim-bit test each), the scans per hash slot: the lists win while
// pixels x ~3 steps stay below the number of slots (measured on B200: 640x480 / 1.18 M slots 33.3 -> 29.4 us, 1280x720
// 51.9 -> 56.0 us).  ITM_B200_ALLOC=scan / lists forces one of them (A/B measurements).
int &alloc_mode();  // 0 automatic, 1 scans, 2 lists wherever they apply (itm_b200_set_alloc_mode; initial value from ITM_B200_ALLOC)
inline bool alloc_uses_lists(const AllocLists &l, bool onlyUpdateVisibleList, bool swapping, int world, int pixels, int nEntries) {
  const int force = alloc_mode();
  if (!l.claimBits || force == 1 || onlyUpdateVisibleList || swapping || world > 1) return false;
  return force == 2 || 3ll * pixels < (long long)nEntries;
}

struct AllocArgs {
  AllocLists lists;              // claimBits == NULL: always the full-table scans
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
  unsigned char *swapStates;     // ITMHashSwapState[nEntries] when the scene swaps (scene->useSwapping), else NULL
  int prologueDone;              // the marking pass already ran (FramePrologue)
  ShardInfo shard;               // world > 1: only resident blocks get a voxel block (local free list), the others ptr = -1
  int *residentVisibleIds;       // sharded scenes: the visible entries with ptr >= 0, ascending (st->noResidentVisible); else NULL
};

struct IntegrateArgs {
  int residentList;           // visibleIds is the resident-visible list: its length is st->noResidentVisible
  const unsigned char *rgb;   // view->rgb, Vector4u[W*H] (ITMVoxel_s_rgb only)
  float rgbIntr[4];           // intrinsics_rgb (fx, fy, cx, cy)
  float calibInv[16];         // trafo_rgb_to_depth.calib_inv
  const float *depth;
  void *voxels;
  const void *hashTable;
  const int *visibleIds;
  FrameState *st;
  ViewParams vp;
  SceneParams sp;
};

struct RenderArgs {
  ShardInfo shard;
  const void *voxels;
  const void *hashTable;
  const int *visibleIds;
  float *minmax;        // Vector2f[W*H] renderingRangeImage
  float *raycastResult; // Vector4f[W*H]
  float *pointsMap;     // Vector4f[W*H]
  float *normalsMap;    // Vector4f[W*H]
  unsigned char *raycastImage;  // Vector4u[W*H]
  int minmaxReady;      // the min/max image is already initialised (FramePrologue)
  int residentList;     // visibleIds is the resident-visible list of a sharded scene (length st->noResidentVisible)
  int gated;            // useApproximateRaycast engines: raycast / ICP maps run only when st->requiresFullRendering
  FrameResult *resultRing;  // host-mapped ring of ITM_RESULT_RING slots the ICP-map kernel publishes pose + counters into (or NULL)
  FrameState *st;
  ViewParams vp;
  SceneParams sp;
};

struct IcpLevelArgs {
  const float *weight;    // weighted ICP only: the level's depth-uncertainty image (weightHierarchy), else NULL
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
int alloc_scan_tile();  // slots per CTA of the two allocation scans (their tile-state arrays need one word per tile)
void launch_reset_scene(void *voxels, int *vbaAllocList, void *table, int *excessAllocList, const SceneParams &sp, cudaStream_t s);
void launch_allocate(const AllocArgs &a, cudaStream_t s);
void launch_integrate(const IntegrateArgs &a, cudaStream_t s);
void launch_expected_depths(const RenderArgs &a, cudaStream_t s);
void launch_raycast(const RenderArgs &a, cudaStream_t s);
// sharded engines: per-pixel nearest hit over every rank's partial image -> a.raycastResult (after the cross-GPU barrier)
void launch_raycast_compose(const RenderArgs &a, cudaStream_t s);
void launch_icp_maps(const RenderArgs &a, cudaStream_t s);

// ForwardRender (useApproximateRaycast): render.raycastResult is projected into forwardProjection at render.st's pose,
// holes are ray cast, render.raycastImage is shaded from the result.  gated: skip unless !st->requiresFullRendering.
struct ForwardArgs {
  RenderArgs render;
  float *forwardProjection;   // Vector4f[W*H]
  int *missingPoints;         // int[W*H] fwdProjMissingPoints
  int *key;                   // scratch int[W*H], all zero between calls
  const float *depth;         // view->depth
  int gated;
};
void launch_forward_render(const ForwardArgs &a, cudaStream_t s);
// st->requiresFullRendering = TrackerFarFromPointCloud() || !useApproximateRaycast
void launch_track_decide(FrameState *st, int useApproximateRaycast, cudaStream_t s);
// FindVisibleBlocks at st's pose with the intrinsics in vp: visibleIds in ascending slot order, st->noVisibleEntries
void launch_find_visible_blocks(const void *hashTable, int *visibleIds, FrameState *st, const ViewParams &vp, const SceneParams &sp,
                                int visibleCapacity, unsigned long long *ticket, unsigned long long *tileState, cudaStream_t s,
                                int allAllocated = 0);  // allAllocated: list every entry that owns a voxel block (meshing)
// ITMMeshingEngine::MeshScene: marching cubes over every allocated block, triangles in the reference's serial order
struct MeshArgs {
  const void *voxels;
  const void *hashTable;
  int *blockList;                 // scratch int[nLocal]: allocated entries, ascending
  unsigned *counts;               // scratch unsigned[nLocal]
  unsigned long long *offsets;    // scratch u64[nLocal + 1]
  void *triangles;                // ITMMesh::Triangle[noMaxTriangles] (36 B each)
  unsigned noMaxTriangles;
  FrameState *st;                 // a FrameState of its own: noVisibleEntries = list length, noMeshTriangles = result
  SceneParams sp;
  unsigned long long *ticket;     // scan scratch (see launch_find_visible_blocks)
  unsigned long long *tileState;
};
void launch_mesh_scene(const MeshArgs &a, cudaStream_t s);
// RenderImage: raycast at st's pose into a.raycastResult and shade into outImage (Vector4u[W*H]); type 0 grey, 1 colour
// from volume, 2 colour from normal
void launch_render_image(const RenderArgs &a, unsigned char *outImage, int type, cudaStream_t s);
// RenderPointCloud's compaction over a.raycastResult / a.raycastImage (grey, from launch_render_image): points in raster
// order into locations / colours, their number into a.st->noTotalPoints.  tileState: point_cloud_tiles(W, H) + 1 zeroed
// words that stay with this tile count (the last one is the ticket counter).
int point_cloud_tiles(int W, int H);
void launch_point_cloud(const RenderArgs &a, int skipPoints, float *locations, float *colours, unsigned long long *tileState,
                        int numTiles, cudaStream_t s);

// fxDisparity != 0: Kinect disparity conversion 8 * b * fxDisparity / (a - raw) instead of the affine raw * a + b
void launch_convert_depth(const short *raw, float *out, int n, float a, float b, cudaStream_t s, float fxDisparity = 0.0f);
void launch_subsample_holes(float *out, const float *in, int wIn, int hIn, cudaStream_t s);
// ITMLowLevelEngine's remaining image helpers (colour / Ren trackers): FilterSubsample(uchar4), FilterSubsampleWithHoles(Vector4f),
// GradientX / GradientY (Vector4s out; clears the first W*H*6 bytes like the reference driver, writes interior pixels)
void launch_subsample_rgba(unsigned char *out, const unsigned char *in, int wIn, int hIn, cudaStream_t s);
void launch_subsample_holes4(float *out, const float *in, int wIn, int hIn, cudaStream_t s);
cudaError_t launch_gradient(short *grad, const unsigned char *image, int W, int H, int alongX, cudaStream_t s);
// ITMViewBuilder::DepthFiltering / ComputeNormalAndWeights (useBilateralFilter / modelSensorNoise)
void launch_filter_depth(float *out, const float *in, int W, int H, cudaStream_t s);
void launch_normal_weight(float *normalOut, float *sigmaOut, const float *depth, int W, int H, const float intr[4], cudaStream_t s);
// Pose-independent first steps of AllocateSceneFromDepth (mark last frame's visible entries, snapshot the free-list
// heads) and of CreateExpectedDepths (min/max image initialisation); when given to launch_view_pyramid they are done
// by the view kernel and the allocate / expected-depth launches skip them (AllocArgs::prologueDone, RenderArgs::minmaxReady).
struct FramePrologue {
  FrameState *st;
  const int *visibleIds;
  unsigned char *visType;   // NULL: no marking
  float2 *minmax;           // NULL: no min/max initialisation
  int minmaxPixels;
  unsigned *icpEpoch;       // NULL: not bumped here (launch_icp_track then bumps it with a launch of its own)
  unsigned *claimBits;      // list-based allocation (AllocLists): the marked entries' claim bits are set; NULL: scans
};
void launch_view_pyramid(const short *raw, float a, float b, float *const *levels, int W, int H, int nLevels, cudaStream_t s,
                         const FramePrologue *prologue = nullptr, float fxDisparity = 0.0f);

// The whole ITMDepthTracker::TrackCamera LM loop as ONE persistent cooperative kernel (levels[l], iters[l] for
// l < nLevels).  rows / bcast: zero-initialised scratch of icp_rows_bytes() / icp_bcast_bytes().  epochDev: the launch
// number, a device word that is never reset (it tags every word CTAs exchange); bumpEpoch: advance it with a 1-thread
// launch first (false when the frame's view kernel already did, FramePrologue::icpEpoch).  Keeping it on the device
// makes the launch parameters identical from frame to frame, so a whole frame can be replayed as a CUDA graph.
cudaError_t launch_icp_track(const IcpArgs &a, const IcpLevelArgs *levels, const int *iters, int nLevels, int noIcpLevel,
                             unsigned long long *rows, unsigned long long *bcast, unsigned *epochDev, bool bumpEpoch, int gridCap,
                             cudaStream_t s, bool weighted = false);  // gridCap > 0: at most that many CTAs (several scenes sharing one GPU)
// weighted: ITMWeightedICPTracker - levels[l].weight = the level of the depth-uncertainty pyramid, Gauss-Newton instead of LM
__device__ __forceinline__ void icp_bump_epoch(unsigned *epochDev) {  // 1 .. 2^32 - 1, never 0 (the scratch starts zeroed)
  const unsigned n = *epochDev + 1u;
  *epochDev = n ? n : 1u;
}
size_t icp_rows_bytes();
size_t icp_bcast_bytes();
// One stand-alone evaluation at poseIn (16 floats, device); [n, f, nabla6, hessian36] left in out44 (device).
// lv.weight != NULL: ITMWeightedICPTracker's evaluation (per-pixel weight from the depth uncertainty)
void launch_icp_eval_single(const IcpArgs &a, const IcpLevelArgs &lv, float *out44, const float *poseIn, cudaStream_t s);
// ITMSwappingEngine (ITMLib/Engine/DeviceSpecific/CPU/ITMSwappingEngine_CPU.cpp); SDF_TRANSFER_BLOCK_NUM = 0x1000
#define ITM_TRANSFER_BLOCK_NUM 0x1000
struct SwapArgs {
  void *voxels;
  void *hashTable;
  int *vbaAllocList;
  const unsigned char *visType;
  unsigned char *swapStates;
  int *neededIds;                   // device: selected entry ids (ITM_TRANSFER_BLOCK_NUM)
  void *transfer;                   // device: syncedVoxelBlocks (ITM_TRANSFER_BLOCK_NUM blocks)
  const unsigned char *hasSynced;   // device: hasSyncedData for the swap-in direction
  unsigned long long *ticket;       // scan scratch
  unsigned long long *tileState;
  FrameState *st;
  SceneParams sp;
  // Layer B's global cache: a pool of blocks in host-mapped pinned memory + the slot of every hash entry (-1: nothing stored)
  void *cachePool;
  int *cacheSlot;
  int *cacheCount;                  // slots handed out so far
  int cachePoolBlocks;
  int *movedCounts;                 // [0] entries combined from the cache, [1] entries moved to the cache by the last frame
};
// ordered selection of the first ITM_TRANSFER_BLOCK_NUM entries (ascending slot order) that need swapping in (mode 0:
// state 1) or can be swapped out (mode 1: state 2, resident, not visible); st->swapCount = how many
void launch_swap_select(const SwapArgs &a, int mode, cudaStream_t s);
void launch_swap_in_apply(const SwapArgs &a, cudaStream_t s);   // IntegrateGlobalIntoLocal's combine loop
void launch_swap_out_apply(const SwapArgs &a, cudaStream_t s);  // SaveToGlobalMemory's device part
// the same two steps against the host-mapped cache pool: no transfer buffer, no host in the loop (Layer B)
void launch_swap_in_direct(const SwapArgs &a, cudaStream_t s);
void launch_swap_out_direct(const SwapArgs &a, cudaStream_t s);

// all ranks meet: returns (on the stream) once every rank has enqueued barrier number seq after its own prior work
void launch_shard_barrier(const ShardInfo &sh, unsigned seq, cudaStream_t s);
// rays no rank could complete, marched with peer reads of the blocks held elsewhere (needs the peers of shard_attach), then the
// "through with the peers' voxels" signal; launch_shard_wait_readers: before a rank's next integration overwrites voxels
void launch_raycast_fallback(const RenderArgs &a, unsigned seq, cudaStream_t s);
void launch_shard_wait_readers(const ShardInfo &sh, unsigned seq, cudaStream_t s);
int icp_max_ctas();
int icp_track_grid();
int integrate_grid();

void launch_set_pose(FrameState *st, cudaStream_t s);  // recompute invM_d from M_d on device

}  // namespace itm
