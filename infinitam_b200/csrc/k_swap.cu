// Host swapping of voxel blocks (out-of-core scenes).
//
// Replaces ITMSwappingEngine
//   IntegrateGlobalIntoLocal / LoadFromGlobalMemory  ITMLib/Engine/DeviceSpecific/CPU/ITMSwappingEngine_CPU.cpp:20-104
//   SaveToGlobalMemory                               same file :107-176
//   combineVoxelDepth/ColorInformation               ITMLib/Engine/DeviceAgnostic/ITMSwappingEngine.h:8-43
//
// The reference walks all hash slots serially and takes the first SDF_TRANSFER_BLOCK_NUM that qualify; the order decides
// which free-list positions the swapped-out blocks return to, so it is reproduced exactly: one single-pass ordered scan
// (scan_util.cuh) ranks the qualifying slots, ranks below the transfer limit are selected.  The per-block work (copy to /
// merge from the transfer buffer, block reset, free-list push) then runs one CTA per selected block.  The host side
// (global cache in host memory, transfers) is in engine.cu.
#include "itm_common.cuh"
#include "kernels.h"
#include "scan_util.cuh"

namespace {

using namespace itm;

#define SWAP_TILE 8192
#define SWAP_PER_THREAD 32

__global__ void __launch_bounds__(256) k_swap_select(const HashEntry *__restrict__ table, const unsigned char *__restrict__ visType,
                                                     const unsigned char *__restrict__ swapStates, int *__restrict__ neededIds,
                                                     FrameState *st, int nEntries, int mode, unsigned long long *ticket,
                                                     unsigned long long *tileState, int numTiles) {
  __shared__ unsigned sWarp[8];
  __shared__ unsigned sTotal;
  __shared__ unsigned sExA;
  __shared__ int sTile;
  __shared__ unsigned sEpoch;
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(ticket, 1ull);
    sTile = (int)(t % (unsigned long long)numTiles);
    sEpoch = (unsigned)((t / (unsigned long long)numTiles + 1ull) & 0xFFFFFull);
  }
  __syncthreads();
  const int tile = sTile;
  const int slot0 = tile * SWAP_TILE + threadIdx.x * SWAP_PER_THREAD;
  unsigned mask = 0;
  for (int j = 0; j < SWAP_PER_THREAD; ++j) {
    const int slot = slot0 + j;
    if (slot >= nEntries) break;
    const unsigned state = swapStates[slot];
    bool take;
    if (mode == 0) take = state == 1;  // data both on host and in active memory, not yet combined
    else take = state == 2 && visType[slot] == 0 && table[slot].ptr >= 0;
    if (take) mask |= 1u << j;
  }
  const unsigned cnt = __popc(mask);
  const unsigned excl = block_exclusive_scan_256(cnt, sWarp, &sTotal);
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned exA, exB;
    scan_lookback(tileState, tile, sEpoch, sTotal, 0, exA, exB);
    if (threadIdx.x == 0) {
      sExA = exA;
      if (tile == numTiles - 1) {
        const int total = (int)(exA + sTotal);
        st->swapCount = total < ITM_TRANSFER_BLOCK_NUM ? total : ITM_TRANSFER_BLOCK_NUM;
        st->swapBaseBlockId = st->lastFreeBlockId;
      }
    }
  }
  __syncthreads();
  int pos = (int)(sExA + excl);
  while (mask) {
    const int j = __ffs(mask) - 1;
    mask &= mask - 1;
    if (pos < ITM_TRANSFER_BLOCK_NUM) neededIds[pos] = slot0 + j;
    pos++;
  }
}

// combineVoxelDepthInformation: src = the copy from the host, dst = the block in active memory
__device__ __forceinline__ uint32_t combine_depth(uint32_t src, uint32_t dst, int maxW) {
  int newW = (int)((dst >> 16) & 0xFFu);
  const int oldW = (int)((src >> 16) & 0xFFu);
  float newF = (float)(short)(dst & 0xFFFFu) / 32767.0f;
  const float oldF = (float)(short)(src & 0xFFFFu) / 32767.0f;
  if (oldW == 0) return dst;
  newF = (float)oldW * oldF + (float)newW * newF;
  newW = oldW + newW;
  newF /= (float)newW;
  newW = (newW < maxW) ? newW : maxW;
  const int sdf = (short)(int)(newF * 32767.0f);
  return (dst & 0xFF000000u) | ((uint32_t)sdf & 0xFFFFu) | (((uint32_t)newW & 0xFFu) << 16);
}

__device__ __forceinline__ unsigned to_uchar_round(float v) {  // Vector3::toUChar
  const int i = (int)((v < 0) ? (v - 0.5f) : (v + 0.5f));
  return (unsigned)(i < 0 ? 0 : (i > 255 ? 255 : i));
}

// combineVoxelColorInformation on the two words of an ITMVoxel_s_rgb
__device__ __forceinline__ void combine_colour(uint32_t srcLo, uint32_t srcHi, uint32_t &dstLo, uint32_t &dstHi, int maxW) {
  int newW = (int)((dstHi >> 16) & 0xFFu);
  const int oldW = (int)((srcHi >> 16) & 0xFFu);
  float nr = (float)(dstLo >> 24) / 255.0f, ng = (float)(dstHi & 0xFFu) / 255.0f, nb = (float)((dstHi >> 8) & 0xFFu) / 255.0f;
  const float orr = (float)(srcLo >> 24) / 255.0f, og = (float)(srcHi & 0xFFu) / 255.0f, ob = (float)((srcHi >> 8) & 0xFFu) / 255.0f;
  if (oldW == 0) return;
  nr = orr * (float)oldW + nr * (float)newW;
  ng = og * (float)oldW + ng * (float)newW;
  nb = ob * (float)oldW + nb * (float)newW;
  newW = oldW + newW;
  nr /= (float)newW; ng /= (float)newW; nb /= (float)newW;
  newW = (newW < maxW) ? newW : maxW;
  dstLo = (dstLo & 0x00FFFFFFu) | (to_uchar_round(nr * 255.0f) << 24);
  dstHi = (dstHi & 0xFF000000u) | to_uchar_round(ng * 255.0f) | (to_uchar_round(nb * 255.0f) << 8) | (((unsigned)(unsigned char)newW) << 16);
}

// IntegrateGlobalIntoLocal's loop (:82-101): one CTA per needed entry
__global__ void __launch_bounds__(256) k_swap_in_apply(uint32_t *__restrict__ voxels, const HashEntry *__restrict__ table,
                                                       unsigned char *__restrict__ swapStates, const int *__restrict__ neededIds,
                                                       const uint32_t *__restrict__ transfer, const unsigned char *__restrict__ hasSynced,
                                                       const FrameState *__restrict__ st, int voxelWords, int maxW) {
  const int n = st->swapCount;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int id = neededIds[i];
    if (hasSynced[i]) {
      const int ptr = table[id].ptr;
      if (ptr >= 0) {
        uint32_t *dst = voxels + (size_t)ptr * ITM_BLOCK_SIZE3 * voxelWords;
        const uint32_t *src = transfer + (size_t)i * ITM_BLOCK_SIZE3 * voxelWords;
        for (int v = threadIdx.x; v < ITM_BLOCK_SIZE3; v += 256) {
          if (voxelWords == 1) {
            dst[v] = combine_depth(src[v], dst[v], maxW);
          } else {
            uint32_t lo = dst[2 * v], hi = dst[2 * v + 1];
            const uint32_t slo = src[2 * v], shi = src[2 * v + 1];
            lo = combine_depth(slo, lo, maxW);
            combine_colour(slo, shi, lo, hi, maxW);
            dst[2 * v] = lo;
            dst[2 * v + 1] = hi;
          }
        }
      }
    }
    if (threadIdx.x == 0) swapStates[id] = 2;
  }
}

// SaveToGlobalMemory's loop (:136-163): one CTA per entry that leaves active memory
__global__ void __launch_bounds__(256) k_swap_out_apply(uint32_t *__restrict__ voxels, HashEntry *__restrict__ table,
                                                        unsigned char *__restrict__ swapStates, const int *__restrict__ neededIds,
                                                        uint32_t *__restrict__ transfer, int *__restrict__ vbaAllocList, FrameState *st,
                                                        int voxelWords, int nBuckets) {
  const int n = st->swapCount;
  const int base = st->swapBaseBlockId;  // noAllocatedVoxelEntries at the start of the loop
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int id = neededIds[i];
    const int ptr = table[id].ptr;
    uint32_t *blk = voxels + (size_t)ptr * ITM_BLOCK_SIZE3 * voxelWords;
    uint32_t *dst = transfer + (size_t)i * ITM_BLOCK_SIZE3 * voxelWords;
    const int vbaIdx = base + i;
    const bool release = vbaIdx < nBuckets - 1;
    for (int w = threadIdx.x; w < ITM_BLOCK_SIZE3 * voxelWords; w += 256) {
      dst[w] = blk[w];
      // TVoxel(): sdf = 32767, everything else 0
      if (release) blk[w] = (voxelWords == 1 || (w & 1) == 0) ? 0x00007FFFu : 0u;
    }
    __syncthreads();  // everybody has read table[id].ptr before it changes
    if (threadIdx.x == 0) {
      swapStates[id] = 0;
      if (release) {
        vbaAllocList[vbaIdx + 1] = ptr;
        table[id].ptr = -1;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int released = n;
    if (base + released > nBuckets - 1) released = (nBuckets - 1 - base) > 0 ? (nBuckets - 1 - base) : 0;
    st->lastFreeBlockId = base + released;
  }
}

// ---- Layer B: the global cache is a pool in host-mapped pinned memory; the kernels move blocks to / from it themselves
// (16-byte accesses over PCIe), so a swapping frame has no host round trip and sits in the frame graph like any other.

// IntegrateGlobalIntoLocal (:69-104): LoadFromGlobalMemory's copy and the combine loop in one pass
__global__ void __launch_bounds__(256) k_swap_in_direct(uint32_t *__restrict__ voxels, const HashEntry *__restrict__ table,
                                                        unsigned char *__restrict__ swapStates, const int *__restrict__ neededIds,
                                                        const uint32_t *__restrict__ pool, const int *__restrict__ cacheSlot,
                                                        const FrameState *__restrict__ st, int voxelWords, int maxW, int *movedCounts) {
  const int n = st->swapCount;
  if (blockIdx.x == 0 && threadIdx.x == 0) movedCounts[0] = n;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int id = neededIds[i];
    const int slot = cacheSlot[id];   // hasStoredData[id]
    const int ptr = table[id].ptr;
    if (slot >= 0 && ptr >= 0) {
      uint32_t *dst = voxels + (size_t)ptr * ITM_BLOCK_SIZE3 * voxelWords;
      const uint32_t *src = pool + (size_t)slot * ITM_BLOCK_SIZE3 * voxelWords;
      if (voxelWords == 1) {
        // 512 words: two per thread, one 8-byte read from the host each
        const uint2 sv = *reinterpret_cast<const uint2 *>(src + 2 * threadIdx.x);
        uint2 dv = *reinterpret_cast<uint2 *>(dst + 2 * threadIdx.x);
        dv.x = combine_depth(sv.x, dv.x, maxW);
        dv.y = combine_depth(sv.y, dv.y, maxW);
        *reinterpret_cast<uint2 *>(dst + 2 * threadIdx.x) = dv;
      } else {
        // 512 two-word voxels: two voxels (16 bytes) per thread
        const uint4 sv = *reinterpret_cast<const uint4 *>(src + 4 * threadIdx.x);
        uint4 dv = *reinterpret_cast<uint4 *>(dst + 4 * threadIdx.x);
        dv.x = combine_depth(sv.x, dv.x, maxW);
        combine_colour(sv.x, sv.y, dv.x, dv.y, maxW);
        dv.z = combine_depth(sv.z, dv.z, maxW);
        combine_colour(sv.z, sv.w, dv.z, dv.w, maxW);
        *reinterpret_cast<uint4 *>(dst + 4 * threadIdx.x) = dv;
      }
    }
    if (threadIdx.x == 0) swapStates[id] = 2;
  }
}

// SaveToGlobalMemory (:107-176): the copy goes straight to the entry's pool slot (SetStoredData), handed out on first use
__global__ void __launch_bounds__(256) k_swap_out_direct(uint32_t *__restrict__ voxels, HashEntry *__restrict__ table,
                                                         unsigned char *__restrict__ swapStates, const int *__restrict__ neededIds,
                                                         uint32_t *__restrict__ pool, int *__restrict__ cacheSlot, int *cacheCount,
                                                         int poolBlocks, int *__restrict__ vbaAllocList, FrameState *st, int voxelWords,
                                                         int nBuckets, int *movedCounts) {
  __shared__ int sSlot;
  const int n = st->swapCount;
  const int base = st->swapBaseBlockId;  // noAllocatedVoxelEntries at the start of the loop
  if (blockIdx.x == 0 && threadIdx.x == 0) movedCounts[1] = n;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int id = neededIds[i];
    const int ptr = table[id].ptr;
    if (threadIdx.x == 0) {
      int slot = cacheSlot[id];
      if (slot < 0) {
        slot = atomicAdd(cacheCount, 1);
        if (slot >= poolBlocks) {
          atomicOr(&st->errorFlags, 4);
          slot = -1;
        } else {
          cacheSlot[id] = slot;
        }
      }
      sSlot = slot;
    }
    __syncthreads();
    const int slot = sSlot;
    uint32_t *blk = voxels + (size_t)ptr * ITM_BLOCK_SIZE3 * voxelWords;
    const int vbaIdx = base + i;
    const bool release = vbaIdx < nBuckets - 1;
    const int words = ITM_BLOCK_SIZE3 * voxelWords;
    for (int w4 = threadIdx.x * 4; w4 < words; w4 += 256 * 4) {
      const uint4 v = *reinterpret_cast<const uint4 *>(blk + w4);
      if (slot >= 0) *reinterpret_cast<uint4 *>(pool + (size_t)slot * words + w4) = v;
      // TVoxel(): sdf = 32767, everything else 0
      if (release)
        *reinterpret_cast<uint4 *>(blk + w4) = voxelWords == 1 ? make_uint4(0x7FFFu, 0x7FFFu, 0x7FFFu, 0x7FFFu) : make_uint4(0x7FFFu, 0u, 0x7FFFu, 0u);
    }
    __syncthreads();  // everybody has read table[id].ptr and sSlot before they change
    if (threadIdx.x == 0) {
      swapStates[id] = 0;
      if (release) {
        vbaAllocList[vbaIdx + 1] = ptr;
        table[id].ptr = -1;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int released = n;
    if (base + released > nBuckets - 1) released = (nBuckets - 1 - base) > 0 ? (nBuckets - 1 - base) : 0;
    st->lastFreeBlockId = base + released;
  }
}

}  // namespace

namespace itm {

void launch_swap_select(const SwapArgs &a, int mode, cudaStream_t s) {
  const int numTiles = (a.sp.nEntries + SWAP_TILE - 1) / SWAP_TILE;
  k_swap_select<<<numTiles, 256, 0, s>>>(reinterpret_cast<const HashEntry *>(a.hashTable), a.visType, a.swapStates, a.neededIds, a.st,
                                         a.sp.nEntries, mode, a.ticket, a.tileState, numTiles);
}

void launch_swap_in_apply(const SwapArgs &a, cudaStream_t s) {
  k_swap_in_apply<<<148 * 4, 256, 0, s>>>(reinterpret_cast<uint32_t *>(a.voxels), reinterpret_cast<const HashEntry *>(a.hashTable), a.swapStates,
                                          a.neededIds, reinterpret_cast<const uint32_t *>(a.transfer), a.hasSynced, a.st, a.sp.voxelWords,
                                          a.sp.maxW);
}

void launch_swap_out_apply(const SwapArgs &a, cudaStream_t s) {
  k_swap_out_apply<<<148 * 4, 256, 0, s>>>(reinterpret_cast<uint32_t *>(a.voxels), reinterpret_cast<HashEntry *>(a.hashTable), a.swapStates,
                                           a.neededIds, reinterpret_cast<uint32_t *>(a.transfer), a.vbaAllocList, a.st, a.sp.voxelWords,
                                           a.sp.nBuckets);
}

void launch_swap_in_direct(const SwapArgs &a, cudaStream_t s) {
  k_swap_in_direct<<<148 * 4, 256, 0, s>>>(reinterpret_cast<uint32_t *>(a.voxels), reinterpret_cast<const HashEntry *>(a.hashTable), a.swapStates,
                                           a.neededIds, reinterpret_cast<const uint32_t *>(a.cachePool), a.cacheSlot, a.st, a.sp.voxelWords,
                                           a.sp.maxW, a.movedCounts);
}

void launch_swap_out_direct(const SwapArgs &a, cudaStream_t s) {
  k_swap_out_direct<<<148 * 4, 256, 0, s>>>(reinterpret_cast<uint32_t *>(a.voxels), reinterpret_cast<HashEntry *>(a.hashTable), a.swapStates,
                                            a.neededIds, reinterpret_cast<uint32_t *>(a.cachePool), a.cacheSlot, a.cacheCount, a.cachePoolBlocks,
                                            a.vbaAllocList, a.st, a.sp.voxelWords, a.sp.nBuckets, a.movedCounts);
}

}  // namespace itm
