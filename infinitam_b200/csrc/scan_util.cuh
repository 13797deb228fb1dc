// Single-pass ordered compaction helpers (tile aggregates published in tagged status words).
//
// The reference CPU engine walks all hash slots serially in ascending order and hands out
// free-list entries / visible-list positions in that order
// (ITMSceneReconstructionEngine_CPU.cpp:179-226, 230-269).  To reproduce the very same
// assignment on the GPU every slot needs its rank among the flagged slots: an exclusive
// prefix sum over ~1.18 M flags.  One kernel does it: each CTA scans its 8192-slot tile in
// shared memory, publishes the tile aggregate in a 64-bit status word and resolves its
// global offset from its predecessors' aggregates.
//
// Status word: [63:62] status  [61:42] epoch (20 bit)  [41:21] count A  [20:0] count B.
// Tiles are handed out by a 64-bit ticket (never reset): tile = ticket % numTiles,
// epoch = ticket / numTiles + 1, so no per-frame clearing of the status array is needed
// and a CTA's predecessors are guaranteed to have started.
#pragma once
#include <cuda_runtime.h>

namespace itm {

#define SCAN_STATUS_INVALID 0ull
#define SCAN_STATUS_AGGREGATE 1ull

__device__ __forceinline__ unsigned long long scan_pack(unsigned long long status, unsigned epoch, unsigned a, unsigned b) {
  return (status << 62) | ((unsigned long long)(epoch & 0xFFFFFu) << 42) | ((unsigned long long)(a & 0x1FFFFFu) << 21) |
         (unsigned long long)(b & 0x1FFFFFu);
}
__device__ __forceinline__ unsigned scan_status(unsigned long long v) { return (unsigned)(v >> 62); }
__device__ __forceinline__ unsigned scan_epoch(unsigned long long v) { return (unsigned)(v >> 42) & 0xFFFFFu; }
__device__ __forceinline__ unsigned scan_a(unsigned long long v) { return (unsigned)(v >> 21) & 0x1FFFFFu; }
__device__ __forceinline__ unsigned scan_b(unsigned long long v) { return (unsigned)v & 0x1FFFFFu; }

__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void scan_st(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by warp 0 of the CTA (all 32 lanes).  Publishes this tile's aggregate (a, b) and returns the exclusive prefix
// over all previous tiles in (exA, exB) (valid in every lane).
// The table is 144 tiles at the shipped size - about one CTA per SM, all resident - so instead of the classic chained
// look-back (a tile waits for its predecessor's inclusive prefix: up to 5 dependent global round trips here) every tile
// simply adds up the AGGREGATES of all its predecessors: lane l reads tiles l, l + 32, ... with all loads in flight at
// once and re-polls only the ones not yet published.  No tile ever waits for another tile's look-back, so the critical
// path is one publish + one read, whatever the number of tiles.
__device__ __forceinline__ void scan_lookback(unsigned long long *state, int tile, unsigned epoch, unsigned aggA, unsigned aggB,
                                              unsigned &exA, unsigned &exB) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) scan_st(state + tile, scan_pack(SCAN_STATUS_AGGREGATE, epoch, aggA, aggB));
  unsigned accA = 0, accB = 0;
  constexpr int CH = 8;  // loads in flight per lane
  for (int j0 = lane; j0 < tile; j0 += 32 * CH) {
    unsigned long long v[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int j = j0 + 32 * k;
      v[k] = j < tile ? scan_ld(state + j) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int j = j0 + 32 * k;
      if (j < tile) {
        while (scan_epoch(v[k]) != epoch || scan_status(v[k]) == SCAN_STATUS_INVALID) v[k] = scan_ld(state + j);
        accA += scan_a(v[k]);
        accB += scan_b(v[k]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    accA += __shfl_xor_sync(0xffffffffu, accA, o);
    accB += __shfl_xor_sync(0xffffffffu, accB, o);
  }
  exA = accA;
  exB = accB;
}

// Exclusive scan of a packed per-thread value over a 256-thread CTA.  Returns the exclusive
// prefix for this thread; total (all threads) is left in *total (shared).
__device__ __forceinline__ unsigned block_exclusive_scan_256(unsigned v, unsigned *sWarp /*[8]*/, unsigned *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) sWarp[warp] = inc;
  __syncthreads();
  unsigned warpBase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const unsigned s = sWarp[w];
    if (w < warp) warpBase += s;
    tot += s;
  }
  if (threadIdx.x == 0) *total = tot;
  return warpBase + inc - v;
}

}  // namespace itm
