// Voxel-block hash allocation and visible-list construction.
//
// Replaces ITMSceneReconstructionEngine::AllocateSceneFromDepth
//   CPU driver      ITMLib/Engine/DeviceSpecific/CPU/ITMSceneReconstructionEngine_CPU.cpp:117-291
//   per pixel       buildHashAllocAndVisibleTypePP  ITMLib/Engine/DeviceAgnostic/ITMSceneReconstructionEngine.h:141-241
//   visibility      checkBlockVisibility / checkPointVisibility  same file :244-342
//
// Determinism.  The reference's per-pixel pass is "last writer wins" per hash slot in raster
// (pixel, step) order, and its allocation loop hands out free-list entries in ascending slot
// order.  Both are reproduced exactly, without serial loops:
//   1. k_alloc_pixels: every ray-segment step that misses the table does
//      atomicMax(allocKey[slot], pixel * stepBound + step + 1) - the largest key IS the last
//      writer of the serial loop.  The block coordinate is not stored; the winner's
//      coordinate is recomputed from its key in step 2 by the same device function.
//   2. k_alloc_scan: single-pass ordered scan over all slots (8192-slot tiles, one CTA per SM) (scan_util.cuh) gives every
//      request its rank, hence exactly the VBA / excess-list entries the serial loop would
//      pop.  The hash table (pos, offset, ptr) comes out bit-identical to the reference.
//   3. k_visible_scan: same scan machinery over entriesVisibleType; visibleEntryIDs come out
//      in ascending slot order like the reference's.  (The frustum re-check of last frame's entries
//      runs one thread per entry at the start of k_alloc_scan.)
//   2'/3'. Where the frame touches fewer table slots than the table has (kernels.h, alloc_uses_lists), steps 2 and 3 work on
//      compact per-bin lists of the requested and the newly visible slots instead of walking all slots (k_alloc_assign,
//      k_visible_merge, "Compact lists" below) - same results.
// No counter is read back by the host; the free-list heads and the visible count live in
// FrameState.
#include <cstdlib>
#include <cstring>

#include "itm_common.cuh"
#include "kernels.h"
#include "scan_util.cuh"
#include "visibility.cuh"

namespace {

using namespace itm;

struct RaySegment {
  float px, py, pz;  // start point in block units
  float dx, dy, dz;  // per-step increment
  int noSteps;
};

// ITMSceneReconstructionEngine.h:153-185; returns false when the pixel is rejected
__device__ __forceinline__ bool make_ray_segment(RaySegment &r, float depth_measure, int x, int y, const float *invM_d,
                                                 float invFx, float invFy, float cx, float cy, float mu, float oneOverVoxelSize,
                                                 float vfMin, float vfMax) {
  if (depth_measure <= 0 || (depth_measure - mu) < 0 || (depth_measure - mu) < vfMin || (depth_measure + mu) > vfMax) return false;
  const float cz = depth_measure;
  const float cxx = cz * (((float)x - cx) * invFx);
  const float cyy = cz * (((float)y - cy) * invFy);
  float norm = sqrtf(cxx * cxx + cyy * cyy + cz * cz);
  const float s0 = 1.0f - mu / norm;
  const float s1 = 1.0f + mu / norm;
  float ax, ay, az, bx, by, bz;
  mat4_mul_vec4(invM_d, cxx * s0, cyy * s0, cz * s0, 1.0f, ax, ay, az);
  ax *= oneOverVoxelSize; ay *= oneOverVoxelSize; az *= oneOverVoxelSize;
  mat4_mul_vec4(invM_d, cxx * s1, cyy * s1, cz * s1, 1.0f, bx, by, bz);
  bx *= oneOverVoxelSize; by *= oneOverVoxelSize; bz *= oneOverVoxelSize;
  float dx = bx - ax, dy = by - ay, dz = bz - az;
  norm = sqrtf(dx * dx + dy * dy + dz * dz);
  const int noSteps = (int)ceilf(2.0f * norm);
  const float div = (float)(noSteps - 1);
  r.px = ax; r.py = ay; r.pz = az;
  r.dx = dx / div; r.dy = dy / div; r.dz = dz / div;
  r.noSteps = noSteps;
  return true;
}

__device__ __forceinline__ void block_of(float px, float py, float pz, int &bx, int &by, int &bz) {
  bx = (short)(int)floorf(px);
  by = (short)(int)floorf(py);
  bz = (short)(int)floorf(pz);
}

// marks last frame's visible entries as "3" (:159-160) and snapshots the free-list heads the
// allocation scan will count down from
__global__ void k_mark_prev_visible(const int *__restrict__ visibleIds, unsigned char *__restrict__ visType, FrameState *st,
                                    unsigned *__restrict__ claimBits) {
  const int n = st->noVisibleEntries;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->allocBaseBlockId = st->lastFreeBlockId;
    st->allocBaseExcessId = st->lastFreeExcessId;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int id = visibleIds[i];
    visType[id] = 3;
    if (claimBits) atomicOr(claimBits + (id >> 5), 1u << (id & 31));  // list-based allocation: "was visible when the frame began"
  }
}

// LISTS (compact lists, see k_alloc_assign): besides the keys, the pass leaves per bin of ITM_ALLOC_BIN slots the list of slots
// that were requested and the list of slots that became visible.  "Became visible" = not in last frame's list: the marking
// pass that sets those entries' type to 3 also sets their claim bit, so a slot whose bit is still clear when this pass
// touches it is new - the threads that see it clear race for the bit and the one winner lists the slot.  The bit is
// tested through L1 (the array is 147 KB, every SM keeps its part): a stale line only costs a lost race.  (Deciding by
// reading the slot's visible type before overwriting it was measured at 100 us: every store evicts the line the next
// thread's load needs.)
__device__ __forceinline__ bool claim_slot(unsigned *claimBits, int slot) {
  unsigned *w = claimBits + (slot >> 5);
  const unsigned bit = 1u << (slot & 31);
  if (*w & bit) return false;        // through L1: set since the marking pass for nearly every slot a ray touches
  if (__ldcg(w) & bit) return false; // a new slot: most of the few hundred threads that touch it find it taken here
  return !(atomicOr(w, bit) & bit);
}
__device__ __forceinline__ void push_bin(int *list, int *counts, int slot, int value) {
  const int bin = slot / ITM_ALLOC_BIN;
  const int i = atomicAdd(counts + bin, 1);
  if (i < ITM_ALLOC_BIN) list[(size_t)bin * ITM_ALLOC_BIN + i] = value;
}
template <bool LISTS>
__device__ __forceinline__ void mark_visible(unsigned char *visType, int slot, unsigned char type, const itm::AllocLists &L) {
  visType[slot] = type;
  if (LISTS && claim_slot(L.claimBits, slot)) push_bin(L.newVisList, L.binCounts + 2 * L.numBins, slot, slot);
}

template <bool LISTS>
__global__ void __launch_bounds__(256) k_alloc_pixels(const float *__restrict__ depth, const HashEntry *__restrict__ table,
                                                      unsigned char *__restrict__ visType, unsigned *__restrict__ allocKey,
                                                      FrameState *st, ViewParams vp, SceneParams sp, float oneOverVoxelSize,
                                                      int stepBound, const itm::AllocLists L) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sInvM[16];
  if (threadIdx.x < 16) sInvM[threadIdx.x] = st->invM_d[threadIdx.x];
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= vp.W || y >= vp.H) return;
  const int locId = x + y * vp.W;
  RaySegment r;
  // invProjParams: x,y are 1/fx, 1/fy computed by the host as float divisions (:131-133)
  if (!make_ray_segment(r, __ldg(depth + locId), x, y, sInvM, 1.0f / vp.fx, 1.0f / vp.fy, vp.cx, vp.cy, sp.mu, oneOverVoxelSize, sp.vfMin,
                        sp.vfMax))
    return;
  if (r.noSteps > stepBound) {
    atomicOr(&st->errorFlags, 1);
    r.noSteps = stepBound;
  }
  float px = r.px, py = r.py, pz = r.pz;
  for (int i = 0; i < r.noSteps; ++i) {
    int bx, by, bz;
    block_of(px, py, pz, bx, by, bz);
    int hashIdx = (int)hash_index(bx, by, bz, sp.hashMask);
    HashEntry e = load_entry(table, hashIdx);
    bool isFound = false;
    if (e.px == bx && e.py == by && e.pz == bz && e.ptr >= -1) {
      mark_visible<LISTS>(visType, hashIdx, (e.ptr == -1) ? 2 : 1, L);
      isFound = true;
    }
    if (!isFound) {
      bool isExcess = false;
      if (e.ptr >= -1) {
        while (e.offset >= 1) {
          hashIdx = sp.nBuckets + e.offset - 1;
          e = load_entry(table, hashIdx);
          if (e.px == bx && e.py == by && e.pz == bz && e.ptr >= -1) {
            mark_visible<LISTS>(visType, hashIdx, (e.ptr == -1) ? 2 : 1, L);
            isFound = true;
            break;
          }
        }
        isExcess = true;
      }
      if (!isFound) {
        const unsigned old = atomicMax(allocKey + hashIdx, (unsigned)locId * (unsigned)stepBound + (unsigned)i + 1u);
        if (LISTS && old == 0) {  // the slot's first request of this frame
          push_bin(L.reqList, L.binCounts, hashIdx, isExcess ? (hashIdx | (int)0x80000000) : hashIdx);
          if (isExcess) atomicAdd(L.binCounts + L.numBins + hashIdx / ITM_ALLOC_BIN, 1);
        }
        if (!isExcess) mark_visible<LISTS>(visType, hashIdx, 1, L);
      }
    }
    px += r.dx; py += r.dy; pz += r.dz;
  }
}

#ifndef SCAN_PER_THREAD
#define SCAN_PER_THREAD 32    // consecutive slots per thread (256 threads); 8, 16 or 32
#endif
#define SCAN_TILE (256 * SCAN_PER_THREAD)  // slots per CTA

// Every thread owns SCAN_PER_THREAD consecutive slots; at 32 the 1.18 M slots are 144 tiles, i.e. one CTA per SM (smaller
// tiles - more CTAs, less serial work per thread - were measured slower).  A tile's offset is the sum of its predecessors'
// published aggregates (scan_util.cuh).
__global__ void __launch_bounds__(256) k_alloc_scan(unsigned *__restrict__ allocKey, HashEntry *__restrict__ table,
                                                    unsigned char *__restrict__ visType, const int *__restrict__ vbaAllocList,
                                                    const int *__restrict__ excessAllocList, const float *__restrict__ depth,
                                                    FrameState *st, ViewParams vp, SceneParams sp, float oneOverVoxelSize,
                                                    int stepBound, int doAllocate, unsigned long long *ticket,
                                                    unsigned long long *tileState, int numTiles,
                                                    const int *__restrict__ prevVisibleIds, int useSwapping, const itm::ShardInfo sh) {
  __shared__ unsigned sWarp[8];
  __shared__ unsigned sTotal;
  __shared__ unsigned sExA, sExB;
  __shared__ int sTile;
  __shared__ unsigned sEpoch;
  __shared__ float sInvM[16];
  __shared__ float sM[16];
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(ticket, 1ull);
    sTile = (int)(t % (unsigned long long)numTiles);
    sEpoch = (unsigned)((t / (unsigned long long)numTiles + 1ull) & 0xFFFFFull);
  }
  if (threadIdx.x < 16) sInvM[threadIdx.x] = st->invM_d[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 48) sM[threadIdx.x - 32] = st->M_d[threadIdx.x - 32];
  __syncthreads();
  // Entries that were visible last frame and that no ray hit this frame still carry type 3: they stay in the list only
  // while one of their corners projects into the image (:236-247).  One thread per previous entry, spread over the grid;
  // independent of the allocation scan below (new entries are never in the previous list).
  {
    const int nPrev = st->noVisibleEntries;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nPrev; i += gridDim.x * blockDim.x) {
      const int id = __ldg(prevVisibleIds + i);
      if (visType[id] == 3) {
        const HashEntry e = load_entry(table, id);
        const bool keep = useSwapping ? block_visible_enlarged(sM, e.px, e.py, e.pz, sp.voxelSize, vp)
                                      : block_visible(sM, e.px, e.py, e.pz, sp.voxelSize, vp);
        if (!keep) visType[id] = 0;
      }
    }
  }
  const int tile = sTile;
  const int slot0 = tile * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
  // the key array is padded to whole tiles, so the vector loads never run past it
  const uint4 *k4 = reinterpret_cast<const uint4 *>(allocKey + slot0);
  uint4 kv[SCAN_PER_THREAD / 4];
  unsigned any = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD / 4; ++i) {
    kv[i] = k4[i];
    any |= kv[i].x | kv[i].y | kv[i].z | kv[i].w;
  }
  unsigned typeMask1 = 0, typeMask2 = 0;  // bit j: slot0 + j requests a bucket entry / an excess-list entry
  unsigned packed = 0;                    // low 16: all requests, high 16: excess-list requests
  if (any && doAllocate) {
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD / 4; ++i) {
      const unsigned keys[4] = {kv[i].x, kv[i].y, kv[i].z, kv[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (keys[q] == 0) continue;
        const int j = i * 4 + q, slot = slot0 + j;
        if (slot < sp.nBuckets && table[slot].ptr < -1) {
          typeMask1 |= 1u << j;
          packed += 1u;
        } else {
          typeMask2 |= 1u << j;
          packed += 0x10001u;
        }
      }
    }
  }
  const unsigned excl = block_exclusive_scan_256(packed, sWarp, &sTotal);
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned exA, exB;
    scan_lookback(tileState, tile, sEpoch, sTotal & 0xFFFFu, sTotal >> 16, exA, exB);
    if (threadIdx.x == 0) {
      sExA = exA;
      sExB = exB;
      if (tile == numTiles - 1) {
        // counters always count down by the number of requests, successful or not (:186, :204-205)
        // (sharded engines pop their local free list once per RESIDENT block, in the loop below; the number of all requests is
        // not what their pool loses)
        if (sh.world <= 1) st->lastFreeBlockId = st->allocBaseBlockId - (int)(exA + (sTotal & 0xFFFFu));
        st->lastFreeExcessId = st->allocBaseExcessId - (int)(exB + (sTotal >> 16));
        st->reallocBaseBlockId = st->lastFreeBlockId;
      }
    }
  }
  __syncthreads();
  if (!any) return;
  int rankA = (int)(sExA + (excl & 0xFFFFu));
  int rankB = (int)(sExB + (excl >> 16));
  const int baseVba = st->allocBaseBlockId, baseExl = st->allocBaseExcessId;
#pragma unroll 1
  for (int j = 0; j < SCAN_PER_THREAD; ++j) {
    const unsigned keyRaw = allocKey[slot0 + j];
    if (keyRaw == 0) continue;
    const int slot = slot0 + j;
    allocKey[slot] = 0;  // leave the array clean for the next frame
    const bool t1 = (typeMask1 >> j) & 1u, t2 = (typeMask2 >> j) & 1u;
    if (!t1 && !t2) continue;
    // recompute the winning request's block coordinate from its (pixel, step) key
    const unsigned key = keyRaw - 1u;
    const int locId = (int)(key / (unsigned)stepBound), step = (int)(key % (unsigned)stepBound);
    const int y = locId / vp.W, x = locId - y * vp.W;
    RaySegment r;
    make_ray_segment(r, __ldg(depth + locId), x, y, sInvM, 1.0f / vp.fx, 1.0f / vp.fy, vp.cx, vp.cy, sp.mu, oneOverVoxelSize, sp.vfMin,
                     sp.vfMax);
    float px = r.px, py = r.py, pz = r.pz;
    for (int i = 0; i < step; ++i) { px += r.dx; py += r.dy; pz += r.dz; }
    int bx, by, bz;
    block_of(px, py, pz, bx, by, bz);
    int vbaIdx = baseVba - rankA;
    rankA++;
    int newPtr;
    if (sh.world > 1) {
      // Sharded scene: the entry is created on every rank (the index is replicated), the voxel block only where the block
      // is resident; elsewhere ptr = -1 (allocated, payload not here).  Local block numbers come from a local counter - their
      // order is of no consequence, no result depends on where in the local pool a block lives.
      if (shard_block_resident(bx, by, bz, sh)) {
        vbaIdx = atomicSub(&st->lastFreeBlockId, 1);
        if (vbaIdx < 0) atomicAdd(&st->lastFreeBlockId, 1);
        newPtr = vbaIdx >= 0 ? vbaAllocList[vbaIdx] : -1;
        if (vbaIdx < 0) atomicAdd(&st->allocFailures, 1);
        vbaIdx = 0;  // the entry itself is always created
      } else {
        newPtr = -1;
        vbaIdx = 0;
      }
    } else {
      newPtr = vbaIdx >= 0 ? vbaAllocList[vbaIdx] : -1;
    }
    if (t1) {
      if (vbaIdx >= 0) store_entry(table, slot, bx, by, bz, 0, newPtr);
      else atomicAdd(&st->allocFailures, 1);
    } else {
      const int exlIdx = baseExl - rankB;
      rankB++;
      if (vbaIdx >= 0 && exlIdx >= 0) {
        const int exlOffset = excessAllocList[exlIdx];
        table[slot].offset = exlOffset + 1;
        store_entry(table, sp.nBuckets + exlOffset, bx, by, bz, 0, newPtr);
        visType[sp.nBuckets + exlOffset] = 1;
      } else {
        atomicAdd(&st->allocFailures, 1);
      }
    }
  }
}

// swapStates != NULL (scene->useSwapping, not onlyUpdateVisibleList): visible entries whose data is not the most recent copy
// are flagged for swap-in (..._CPU.cpp:250-253), and visible entries that were swapped out (ptr == -1) get a voxel block
// again, in ascending slot order from the free list (:272-285) - the second count of the same scan.
__global__ void __launch_bounds__(256) k_visible_scan(const unsigned char *__restrict__ visType,
                                                      int *__restrict__ visibleIds, FrameState *st, SceneParams sp,
                                                      int visibleCapacity, unsigned long long *ticket, unsigned long long *tileState,
                                                      int numTiles, unsigned char *__restrict__ swapStates, HashEntry *__restrict__ table,
                                                      const int *__restrict__ vbaAllocList, int *__restrict__ residentIds) {
  __shared__ unsigned sWarp[8];
  __shared__ unsigned sTotal;
  __shared__ unsigned sExA, sExB;
  __shared__ int sTile;
  __shared__ unsigned sEpoch;
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(ticket, 1ull);
    sTile = (int)(t % (unsigned long long)numTiles);
    sEpoch = (unsigned)((t / (unsigned long long)numTiles + 1ull) & 0xFFFFFull);
  }
  __syncthreads();
  const int tile = sTile;
  const int slot0 = tile * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
  // 32 type bytes per thread.  entriesVisibleType is caller-owned and exactly nEntries long: the last tile may be ragged.
  unsigned w[SCAN_PER_THREAD / 4];
  if (slot0 + SCAN_PER_THREAD <= sp.nEntries) {
#if SCAN_PER_THREAD == 32
    const uint4 a = *reinterpret_cast<const uint4 *>(visType + slot0);
    const uint4 b = *reinterpret_cast<const uint4 *>(visType + slot0 + 16);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
#elif SCAN_PER_THREAD == 16
    const uint4 a = *reinterpret_cast<const uint4 *>(visType + slot0);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
#else
    const uint2 a = *reinterpret_cast<const uint2 *>(visType + slot0);
    w[0] = a.x; w[1] = a.y;
#endif
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD / 4; ++i) {
      w[i] = 0;
      for (int q = 0; q < 4; ++q)
        if (slot0 + i * 4 + q < sp.nEntries) w[i] |= (unsigned)visType[slot0 + i * 4 + q] << (8 * q);
    }
  }
  unsigned cnt = 0;
  unsigned liveMask = 0;  // bit j: slot0 + j is in the visible list (type > 0; type 3 was re-checked by k_alloc_scan)
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD / 4; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if ((w[i] >> (8 * q)) & 0xFFu) {
        cnt++;
        liveMask |= 1u << (i * 4 + q);
      }
    }
  }
  unsigned reallocMask = 0;  // bit j: slot0 + j is visible but swapped out
  if (swapStates) {
    unsigned m = liveMask;
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      if (swapStates[slot0 + j] != 2) swapStates[slot0 + j] = 1;
      if (table[slot0 + j].ptr == -1) reallocMask |= 1u << j;
    }
  }
  // sharded scenes (never swapping): the second count of the scan ranks the visible entries whose block is resident here
  unsigned residentMask = 0;
  if (residentIds) {
    unsigned m = liveMask;
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      if (table[slot0 + j].ptr >= 0) residentMask |= 1u << j;
    }
  }
  const unsigned excl = block_exclusive_scan_256(cnt | ((unsigned)__popc(reallocMask | residentMask) << 16), sWarp, &sTotal);
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned exA, exB;
    scan_lookback(tileState, tile, sEpoch, sTotal & 0xFFFFu, sTotal >> 16, exA, exB);
    if (threadIdx.x == 0) {
      sExA = exA;
      sExB = exB;
      if (tile == numTiles - 1) {
        int total = (int)(exA + (sTotal & 0xFFFFu));
        if (total > visibleCapacity) {
          atomicOr(&st->errorFlags, 2);
          total = visibleCapacity;
        }
        st->noVisibleEntries = total;
        if (swapStates) st->lastFreeBlockId = st->reallocBaseBlockId - (int)(exB + (sTotal >> 16));
        if (residentIds) st->noResidentVisible = min((int)(exB + (sTotal >> 16)), sp.nLocal);
      }
    }
  }
  __syncthreads();
  if (cnt == 0) return;
  int pos = (int)(sExA + (excl & 0xFFFFu));
  while (liveMask) {
    const int j = __ffs(liveMask) - 1;
    liveMask &= liveMask - 1;
    if (pos < visibleCapacity) visibleIds[pos] = slot0 + j;
    pos++;
  }
  if (residentMask) {
    int posB = (int)(sExB + (excl >> 16));
    while (residentMask) {
      const int j = __ffs(residentMask) - 1;
      residentMask &= residentMask - 1;
      if (posB < sp.nLocal) residentIds[posB] = slot0 + j;
      posB++;
    }
  }
  if (reallocMask) {
    int vbaIdx = st->reallocBaseBlockId - (int)(sExB + (excl >> 16));
    while (reallocMask) {
      const int j = __ffs(reallocMask) - 1;
      reallocMask &= reallocMask - 1;
      if (vbaIdx >= 0) table[slot0 + j].ptr = vbaAllocList[vbaIdx];
      vbaIdx--;
    }
  }
}

// =====================================================================================================================
// Compact lists.  The two ordered scans above walk all ~1.18 M slots every frame to rank a few hundred requests and a few
// thousand visible entries; their cost is the chain of dependent, mostly cold memory round trips of two whole-table
// kernels.  The reference's order - ascending slot index - can be had from what the per-pixel pass already knows:
//   * the slots are cut into bins of ITM_ALLOC_BIN consecutive slots (144 bins for the default table, one CTA each);
//     k_alloc_pixels<true> drops every requested slot and every newly visible slot into its bin's list;
//   * k_alloc_assign: rank of a request = requests in lower bins (a 144-term sum) + requests of its own bin with a lower
//     slot (a handful); the winner's block coordinate is recomputed from its key and the entry filled exactly as above.
//     The same kernel re-checks last frame's visible entries that no ray touched (grid-wide, one thread per entry), copies
//     the list aside and counts, per bin, its entries and the ones that stay;
//   * k_visible_merge: last frame's list is ascending already, so the new list is the merge of its kept entries with the
//     sorted newly-visible slots: position = kept + new entries in lower bins, + kept entries of the own bin below, + new
//     entries of the own bin below.
// Results (hash table, free-list heads, entriesVisibleType, visibleEntryIDs) are bit-identical to the scans'.  Engines that
// swap, shard, or only update the visible list keep the scans (their second ranking rides on the same pass there).
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum of counts[0..bin) and of all numBins counts, by warp 0 of the CTA; results in out[0], out[1] (shared)
__device__ __forceinline__ void bin_prefix(const int *__restrict__ counts, int numBins, int bin, int *out) {
  if (threadIdx.x < 32) {
    int lo = 0, all = 0;
    for (int b = threadIdx.x; b < numBins; b += 32) {
      const int c = min(__ldcg(counts + b), ITM_ALLOC_BIN);
      all += c;
      if (b < bin) lo += c;
    }
    lo = warp_sum(lo);
    all = warp_sum(all);
    if (threadIdx.x == 0) { out[0] = lo; out[1] = all; }
  }
}

__global__ void __launch_bounds__(256) k_alloc_assign(unsigned *__restrict__ allocKey, HashEntry *__restrict__ table,
                                                      unsigned char *__restrict__ visType, const int *__restrict__ vbaAllocList,
                                                      const int *__restrict__ excessAllocList, const float *__restrict__ depth,
                                                      FrameState *st, ViewParams vp, SceneParams sp, float oneOverVoxelSize,
                                                      int stepBound, const int *__restrict__ prevVisibleIds, const itm::AllocLists L) {
  __shared__ float sInvM[16];
  __shared__ float sM[16];
  __shared__ int sReq[2], sEx[2];
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x < 16) sInvM[threadIdx.x] = st->invM_d[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 48) sM[threadIdx.x - 32] = st->M_d[threadIdx.x - 32];
  const int bin = blockIdx.x, numBins = L.numBins;
  int *cntReq = L.binCounts, *cntEx = L.binCounts + numBins, *cntNew = L.binCounts + 2 * numBins;
  int *cntPrev = L.binCounts + 3 * numBins, *cntKept = L.binCounts + 4 * numBins;
  bin_prefix(cntReq, numBins, bin, sReq);
  __syncthreads();
  if (threadIdx.x < 32) {  // (excess-list requests: a second pass of the same warp)
    int lo = 0, all = 0;
    for (int b = threadIdx.x; b < numBins; b += 32) {
      const int c = __ldcg(cntEx + b);
      all += c;
      if (b < bin) lo += c;
    }
    lo = warp_sum(lo);
    all = warp_sum(all);
    if (threadIdx.x == 0) { sEx[0] = lo; sEx[1] = all; }
  }
  __syncthreads();

  // ---- last frame's visible entries: frustum re-check of the untouched ones (..._CPU.cpp:236-247), copy, per-bin counts
  {
    const int nPrev = st->noVisibleEntries;
    const int nThreads = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < nPrev; i0 += nThreads) {
      const int i = i0 + (threadIdx.x & 31);
      bool keep = false;
      int b = -1;
      if (i < nPrev) {
        const int id = __ldg(prevVisibleIds + i);
        unsigned char t = visType[id];
        if (t == 3) {
          const HashEntry e = load_entry(table, id);
          if (!block_visible(sM, e.px, e.py, e.pz, sp.voxelSize, vp)) {
            visType[id] = 0;
            t = 0;
          }
        }
        keep = t != 0;
        b = id / ITM_ALLOC_BIN;
        L.prevCopy[i] = id;
        L.prevKeep[i] = keep ? 1 : 0;
      }
      // one atomic per (warp, bin): the list is ascending, a warp's 32 entries fall into one or two bins
      const unsigned same = __match_any_sync(0xffffffffu, b);
      const unsigned kept = __ballot_sync(0xffffffffu, keep);
      if (b >= 0 && (threadIdx.x & 31) == __ffs(same) - 1) {
        atomicAdd(cntPrev + b, __popc(same));
        const int k = __popc(same & kept);
        if (k) atomicAdd(cntKept + b, k);
      }
    }
  }

  // ---- this bin's requests
  const int nReq = min(__ldcg(cntReq + bin), ITM_ALLOC_BIN);
  const int *__restrict__ myList = L.reqList + (size_t)bin * ITM_ALLOC_BIN;
  const int baseVba = st->allocBaseBlockId, baseExl = st->allocBaseExcessId;
  if (bin == numBins - 1 && threadIdx.x == 0) {
    // counters always count down by the number of requests, successful or not (:186, :204-205)
    st->lastFreeBlockId = baseVba - sReq[1];
    st->lastFreeExcessId = baseExl - sEx[1];
    st->reallocBaseBlockId = st->lastFreeBlockId;
  }
  for (int r = threadIdx.x; r < nReq; r += blockDim.x) {
    const int mine = __ldcg(myList + r);
    const int slot = mine & 0x7fffffff;
    const bool t2 = mine < 0;
    int rankA = sReq[0], rankB = sEx[0];
    for (int j = 0; j < nReq; ++j) {
      const int v = __ldcg(myList + j);
      const bool less = (v & 0x7fffffff) < slot;
      rankA += less ? 1 : 0;
      rankB += (less && v < 0) ? 1 : 0;
    }
    const unsigned keyRaw = allocKey[slot];
    allocKey[slot] = 0;  // leave the array clean for the next frame
    // recompute the winning request's block coordinate from its (pixel, step) key
    const unsigned key = keyRaw - 1u;
    const int locId = (int)(key / (unsigned)stepBound), step = (int)(key % (unsigned)stepBound);
    const int y = locId / vp.W, x = locId - y * vp.W;
    RaySegment rs;
    make_ray_segment(rs, __ldg(depth + locId), x, y, sInvM, 1.0f / vp.fx, 1.0f / vp.fy, vp.cx, vp.cy, sp.mu, oneOverVoxelSize, sp.vfMin,
                     sp.vfMax);
    float px = rs.px, py = rs.py, pz = rs.pz;
    for (int i = 0; i < step; ++i) { px += rs.dx; py += rs.dy; pz += rs.dz; }
    int bx, by, bz;
    block_of(px, py, pz, bx, by, bz);
    const int vbaIdx = baseVba - rankA;
    const int newPtr = vbaIdx >= 0 ? vbaAllocList[vbaIdx] : -1;
    if (!t2) {
      if (vbaIdx >= 0) store_entry(table, slot, bx, by, bz, 0, newPtr);
      else atomicAdd(&st->allocFailures, 1);
    } else {
      const int exlIdx = baseExl - rankB;
      if (vbaIdx >= 0 && exlIdx >= 0) {
        const int exlOffset = excessAllocList[exlIdx];
        table[slot].offset = exlOffset + 1;
        const int newSlot = sp.nBuckets + exlOffset;
        store_entry(table, newSlot, bx, by, bz, 0, newPtr);
        visType[newSlot] = 1;
        push_bin(L.newVisList, cntNew, newSlot, newSlot);
      } else {
        atomicAdd(&st->allocFailures, 1);
      }
    }
  }
}

// exclusive scan of one value per thread over the CTA (256 threads); *total = the CTA's sum.  sTmp: 8 ints of shared memory
__device__ __forceinline__ int cta_exclusive_scan(int v, int *sTmp, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();  // sTmp of the previous call has been consumed
  if (lane == 31) sTmp[warp] = inc;
  __syncthreads();
  int base = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int t = sTmp[w];
    if (w < warp) base += t;
    total += t;
  }
  return base + inc - v;
}

// index of the first element of the ascending run a[0..n) that is >= key
__device__ __forceinline__ int lower_bound_ldg(const int *__restrict__ a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_visible_merge(int *__restrict__ visibleIds, FrameState *st, int visibleCapacity,
                                                       const itm::AllocLists L) {
  __shared__ int sNew[2], sKept[2], sPrev[2];
  __shared__ int sTmp[8];
  __shared__ bool sLast;
  // per bin: newAt[b] (16 bit, two per word) = new slots whose place is right before previous entry b;
  // keptBefore[b] = kept previous entries of this bin with an index below b
  __shared__ unsigned sNewAt[ITM_ALLOC_BIN / 2 + 2];
  __shared__ unsigned short sKeptBefore[ITM_ALLOC_BIN + 2];
  pdl_wait();
  pdl_trigger();
  const int bin = blockIdx.x, numBins = L.numBins;
  int *cntNew = L.binCounts + 2 * numBins, *cntPrev = L.binCounts + 3 * numBins, *cntKept = L.binCounts + 4 * numBins;
  bin_prefix(cntNew, numBins, bin, sNew);
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    int lo = 0, all = 0, plo = 0;
    for (int b = threadIdx.x - 32; b < numBins; b += 32) {
      const int c = __ldcg(cntKept + b), p = __ldcg(cntPrev + b);
      all += c;
      if (b < bin) { lo += c; plo += p; }
    }
    lo = warp_sum(lo);
    all = warp_sum(all);
    plo = warp_sum(plo);
    if (threadIdx.x == 32) { sKept[0] = lo; sKept[1] = all; sPrev[0] = plo; }
  }
  const int nNew = min(__ldcg(cntNew + bin), ITM_ALLOC_BIN);
  const int nPrevBin = min(__ldcg(cntPrev + bin), ITM_ALLOC_BIN);  // (a bin holds at most ITM_ALLOC_BIN distinct slots)
  for (int i = threadIdx.x; i < ITM_ALLOC_BIN / 2 + 2; i += 256) sNewAt[i] = 0;
  __syncthreads();
  // every count this CTA needs has been read: the last CTA to get here clears them for the next frame
  if (threadIdx.x == 0) {
    __threadfence();
    sLast = atomicAdd(L.done, 1) == (int)gridDim.x - 1;
  }
  const int *__restrict__ newList = L.newVisList + (size_t)bin * ITM_ALLOC_BIN;
  const int *__restrict__ prevIds = L.prevCopy + sPrev[0];
  const unsigned char *__restrict__ prevKeep = L.prevKeep + sPrev[0];
  const int outBase = sKept[0] + sNew[0];
  if (bin == 0 && threadIdx.x == 0) {
    int total = sKept[1] + sNew[1];
    if (total > visibleCapacity) {
      atomicOr(&st->errorFlags, 2);
      total = visibleCapacity;
    }
    st->noVisibleEntries = total;
  }
  // (1) where in the previous run does each new slot belong?
  for (int k = threadIdx.x; k < nNew; k += 256) {
    const int b = lower_bound_ldg(prevIds, nPrevBin, __ldg(newList + k));
    atomicAdd(&sNewAt[b >> 1], 1u << (16 * (b & 1)));
  }
  __syncthreads();
  // (2) the kept previous entries: position = kept entries below + new slots below, both running sums over the run
  int keptRun = 0, newRun = 0;  // totals of the earlier rounds (the same in every thread)
  for (int j0 = 0; j0 < nPrevBin; j0 += 1024) {   // 4 consecutive entries per thread and round
    int id[4], keep[4], at[4];
    int keptMine = 0, newMine = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + threadIdx.x * 4 + q;
      const bool valid = j < nPrevBin;
      id[q] = valid ? __ldg(prevIds + j) : 0;
      keep[q] = (valid && __ldg(prevKeep + j) != 0) ? 1 : 0;
      at[q] = valid ? (int)((sNewAt[j >> 1] >> (16 * (j & 1))) & 0xFFFFu) : 0;
      keptMine += keep[q];
      newMine += at[q];
    }
    int keptTotal, newTotal;
    int keptEx = keptRun + cta_exclusive_scan(keptMine, sTmp, keptTotal);
    int newEx = newRun + cta_exclusive_scan(newMine, sTmp, newTotal);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + threadIdx.x * 4 + q;
      if (j < nPrevBin) sKeptBefore[j] = (unsigned short)keptEx;
      newEx += at[q];  // new slots placed before entry j are below it
      if (keep[q]) {
        const int pos = outBase + keptEx + newEx;
        if (pos < visibleCapacity) visibleIds[pos] = id[q];
      }
      keptEx += keep[q];
    }
    keptRun += keptTotal;
    newRun += newTotal;
  }
  if (threadIdx.x == 0) sKeptBefore[nPrevBin] = (unsigned short)keptRun;
  __syncthreads();
  // (3) the new slots: kept previous entries below + new slots below
  for (int k = threadIdx.x; k < nNew; k += 256) {
    const int slot = __ldg(newList + k);
    const int b = lower_bound_ldg(prevIds, nPrevBin, slot);
    int below = (int)sKeptBefore[b];
    for (int q = 0; q < nNew; ++q) below += (__ldg(newList + q) < slot) ? 1 : 0;
    const int pos = outBase + below;
    if (pos < visibleCapacity) visibleIds[pos] = slot;
  }
  __syncthreads();
  if (sLast) {
    for (int i = threadIdx.x; i < 5 * numBins; i += 256) L.binCounts[i] = 0;
    if (threadIdx.x == 0) *L.done = 0;
  }
  // the claim bits are this frame's only (the per-pixel pass is long over): every CTA clears its bin's
  unsigned *bits = L.claimBits + (size_t)bin * (ITM_ALLOC_BIN / 32);
  for (int i = threadIdx.x; i < ITM_ALLOC_BIN / 32; i += 256) bits[i] = 0;
}

// ResetScene, ITMSceneReconstructionEngine_CPU.cpp:25-45
__global__ void k_reset_scene(uint32_t *__restrict__ voxels, size_t nVectors, int *__restrict__ vbaAllocList, int nLocal,
                              HashEntry *__restrict__ table, int nEntries, int *__restrict__ excessAllocList, int nExcess,
                              uint4 ev) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 *v4 = reinterpret_cast<uint4 *>(voxels);
  for (size_t i = tid; i < nVectors; i += stride) v4[i] = ev;
  for (size_t i = tid; i < (size_t)nLocal; i += stride) vbaAllocList[i] = (int)i;
  for (size_t i = tid; i < (size_t)nEntries; i += stride) store_entry(table, (int)i, 0, 0, 0, 0, -2);
  for (size_t i = tid; i < (size_t)nExcess; i += stride) excessAllocList[i] = (int)i;
}

}  // namespace

namespace itm {

int alloc_scan_tile() { return SCAN_TILE; }

int &alloc_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("ITM_B200_ALLOC");
    mode = !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'l' ? 2 : 0));
  }
  return mode;
}

int alloc_step_bound(const SceneParams &sp) {
  // noSteps = ceil(2 * |segment| / blockSize); |segment| = 2*mu up to rounding
  const float len = 2.0f * sp.mu / (sp.voxelSize * (float)ITM_BLOCK_SIZE);
  return (int)ceilf(2.0f * len * 1.01f) + 2;
}

void launch_reset_scene(void *voxels, int *vbaAllocList, void *table, int *excessAllocList, const SceneParams &sp, cudaStream_t s) {
  // ITMVoxel_s(): sdf = 32767, w_depth = 0 (ITMLibDefines.h:175-178); ITMVoxel_s_rgb(): + clr = 0, w_color = 0 (:149-154)
  const uint32_t emptyVoxel = 0x00007FFFu;
  const uint4 ev = sp.voxelWords == 2 ? make_uint4(emptyVoxel, 0u, emptyVoxel, 0u) : make_uint4(emptyVoxel, emptyVoxel, emptyVoxel, emptyVoxel);
  k_reset_scene<<<148 * 8, 256, 0, s>>>(reinterpret_cast<uint32_t *>(voxels), (size_t)sp.nLocal * ITM_BLOCK_SIZE3 * sp.voxelWords / 4, vbaAllocList,
                                       sp.nLocal, reinterpret_cast<HashEntry *>(table), sp.nEntries, excessAllocList, sp.nExcess, ev);
}

void launch_allocate(const AllocArgs &a, cudaStream_t s) {
  const float oneOverVoxelSize = 1.0f / (a.sp.voxelSize * ITM_BLOCK_SIZE);
  const int stepBound = alloc_step_bound(a.sp);
  HashEntry *table = reinterpret_cast<HashEntry *>(a.hashTable);
  const bool lists = alloc_uses_lists(a.lists, a.onlyUpdateVisibleList != 0, a.swapStates != nullptr, a.shard.world, a.vp.W * a.vp.H, a.sp.nEntries);
  if (!a.prologueDone) k_mark_prev_visible<<<64, 256, 0, s>>>(a.visibleIds, a.visType, a.st, lists ? a.lists.claimBits : nullptr);
  dim3 g((a.vp.W + 31) / 32, (a.vp.H + 7) / 8);
  if (lists) {
    launch_pdl(k_alloc_pixels<true>, g, dim3(256), s, a.depth, (const HashEntry *)table, a.visType, a.allocKey, a.st, a.vp, a.sp, oneOverVoxelSize,
               stepBound, a.lists);
    launch_pdl(k_alloc_assign, dim3(a.lists.numBins), dim3(256), s, a.allocKey, table, a.visType, (const int *)a.vbaAllocList,
               (const int *)a.excessAllocList, (const float *)a.depth, a.st, a.vp, a.sp, oneOverVoxelSize, stepBound, (const int *)a.visibleIds,
               a.lists);
    launch_pdl(k_visible_merge, dim3(a.lists.numBins), dim3(256), s, a.visibleIds, a.st, a.visibleCapacity, a.lists);
    return;
  }
  launch_pdl(k_alloc_pixels<false>, g, dim3(256), s, a.depth, (const HashEntry *)table, a.visType, a.allocKey, a.st, a.vp, a.sp, oneOverVoxelSize,
             stepBound, a.lists);
  const int numTiles = (a.sp.nEntries + SCAN_TILE - 1) / SCAN_TILE;
  launch_pdl(k_alloc_scan, dim3(numTiles), dim3(256), s, a.allocKey, table, a.visType, (const int *)a.vbaAllocList, (const int *)a.excessAllocList,
             (const float *)a.depth, a.st, a.vp, a.sp, oneOverVoxelSize, stepBound, a.onlyUpdateVisibleList ? 0 : 1, a.scanTickets, a.allocTileState,
             numTiles, (const int *)a.visibleIds, (a.swapStates && !a.onlyUpdateVisibleList) ? 1 : 0, a.shard);
  launch_pdl(k_visible_scan, dim3(numTiles), dim3(256), s, (const unsigned char *)a.visType, a.visibleIds, a.st, a.sp, a.visibleCapacity,
             a.scanTickets + 1, a.visTileState, numTiles, a.onlyUpdateVisibleList ? (unsigned char *)nullptr : a.swapStates, table,
             (const int *)a.vbaAllocList, a.shard.world > 1 ? a.residentVisibleIds : (int *)nullptr);
}

}  // namespace itm
