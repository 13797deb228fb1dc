// Point-to-plane ICP: per-level G/H normal-equation reduction and the Levenberg-Marquardt loop.
//
// Replaces:
//   ITMDepthTracker::TrackCamera          ITMLib/Engine/ITMDepthTracker.cpp:145-199  (LM loop)
//   ITMDepthTracker_CPU::ComputeGandH     ITMLib/Engine/DeviceSpecific/CPU/ITMDepthTracker_CPU.cpp:14-79
//   computePerPointGH_Depth(_Ab)          ITMLib/Engine/DeviceAgnostic/ITMDepthTracker.h:9-105
//   interpolateBilinear_withHoles         ITMLib/Engine/DeviceAgnostic/ITMPixelUtils.h:41-71
//
// B200 design.  The reference CUDA tracker does memset + kernel + blocking 29-word D2H copy + host
// Cholesky for each of up to 30 evaluations per frame; most of a frame at kHz rates is those round
// trips.  Here the WHOLE TrackCamera is one persistent, cooperatively launched kernel:
//   * the grid (<= 2 CTAs per SM, all co-resident) walks the pyramid levels and LM iterations itself;
//   * per evaluation each thread accumulates its pixels' (count, b^2, b*A, A*A^T) in fp32 registers,
//     warps reduce with shuffles, the CTA combines its warps in fp64 and writes one 32-value partial;
//   * CTAs then meet at a grid barrier built from one atomic counter + a generation word.  The LAST
//     CTA to arrive sums the partials in CTA order (fp64: bit-reproducible run to run), runs the
//     accept/reject + Cholesky + SE(3) update of the LM loop (pose_math.cuh, everything in registers)
//     on the pose kept in FrameState, and releases the barrier; the others spin on the generation.
//   * HasConverged() simply ends the level's loop - nothing is launched for skipped iterations.
// k_icp_eval_single is the stand-alone evaluation behind the stage-level ComputeGandH entry point.
#include <cstdlib>

#include "itm_common.cuh"
#include "kernels.h"
#include "pose_math.cuh"

namespace {

using namespace itm;

#ifndef ICP_THREADS
#define ICP_THREADS 512
#endif
#ifndef ICP_CTAS_PER_SM
#define ICP_CTAS_PER_SM 1
#endif
#define ICP_MAX_CTAS (148 * ICP_CTAS_PER_SM)
#ifndef ICP_CHUNKS_PER_CTA
#define ICP_CHUNKS_PER_CTA 2
#endif
#define ICP_NVALS 32  // 1 count + 1 f + 6 nabla + 21 hessian, padded
#ifndef ICP_EARLY_NORMALS
#define ICP_EARLY_NORMALS 0  // issue the normal-map taps together with the point-map taps: 0 never, 1 always (the 29-value
#endif                       // evaluation then spills), 2 in the short (rotation- / translation-only) evaluations

__device__ __forceinline__ bool bilinear_holes(const float4 *__restrict__ src, float px, float py, int W, float &rx, float &ry,
                                               float &rz, float &rw) {
  const int ix = (short)(int)floorf(px), iy = (short)(int)floorf(py);
  const float dx = px - (float)ix, dy = py - (float)iy;
  const float4 a = __ldg(src + ix + iy * W);
  const float4 b = __ldg(src + (ix + 1) + iy * W);
  const float4 c = __ldg(src + ix + (iy + 1) * W);
  const float4 d = __ldg(src + (ix + 1) + (iy + 1) * W);
  if (a.w < 0 || b.w < 0 || c.w < 0 || d.w < 0) {
    rx = 0; ry = 0; rz = 0; rw = -1.0f;
    return false;
  }
  rx = (a.x * (1.0f - dx) * (1.0f - dy) + b.x * dx * (1.0f - dy) + c.x * (1.0f - dx) * dy + d.x * dx * dy);
  ry = (a.y * (1.0f - dx) * (1.0f - dy) + b.y * dx * (1.0f - dy) + c.y * (1.0f - dx) * dy + d.y * dx * dy);
  rz = (a.z * (1.0f - dx) * (1.0f - dy) + b.z * dx * (1.0f - dy) + c.z * (1.0f - dx) * dy + d.z * dx * dy);
  rw = (a.w * (1.0f - dx) * (1.0f - dy) + b.w * dx * (1.0f - dy) + c.w * (1.0f - dx) * dy + d.w * dx * dy);
  return true;
}

struct IcpConsts {
  float approxInvPose[16];
  float scenePose[16];
};

// weighted: computePerPointGH_wICP (ITMLib/Engine/DeviceAgnostic/ITMWeightedICPTracker.h:9-105) - the interpolated normal is
// scaled by localWeight AFTER b has been taken with the unscaled one
template <bool shortIteration, bool rotationOnly, bool weighted = false>
__device__ __forceinline__ bool per_point_Ab(float *A, float &b, int x, int y, float depth, const IcpLevelArgs &lv, const ViewParams &sv,
                                             const IcpConsts &c, const float4 *__restrict__ pointsMap,
                                             const float4 *__restrict__ normalsMap, float localWeight = 1.0f) {
  if (depth <= 1e-8f) return false;
  float tx = depth * (((float)x - lv.cx) / lv.fx);
  float ty = depth * (((float)y - lv.cy) / lv.fy);
  float tz = depth;
  // transform to previous frame coordinates
  float wx, wy, wz;
  mat4_mul_vec4(c.approxInvPose, tx, ty, tz, 1.0f, wx, wy, wz);
  // project into previous rendered image
  float rx, ry, rz;
  mat4_mul_vec4(c.scenePose, wx, wy, wz, 1.0f, rx, ry, rz);
  if (rz <= 0.0f) return false;
  const float u = sv.fx * rx / rz + sv.cx;
  const float v = sv.fy * ry / rz + sv.cy;
  if (!((u >= 0.0f) && (u <= (float)(sv.W - 2)) && (v >= 0.0f) && (v <= (float)(sv.H - 2)))) return false;
  float cxp, cyp, czp, cwp;
  bilinear_holes(pointsMap, u, v, sv.W, cxp, cyp, czp, cwp);
  if (cwp < 0.0f) return false;
  const float dx = cxp - wx, dy = cyp - wy, dz = czp - wz;
  const float dist = dx * dx + dy * dy + dz * dz;
  if (dist > lv.distThresh) return false;
  float nx, ny, nz, nw;
  bilinear_holes(normalsMap, u, v, sv.W, nx, ny, nz, nw);
  b = nx * dx + ny * dy + nz * dz;
  if (weighted) { nx *= localWeight; ny *= localWeight; nz *= localWeight; }
  if (shortIteration) {
    if (rotationOnly) {
      A[0] = +wz * ny - wy * nz;
      A[1] = -wz * nx + wx * nz;
      A[2] = +wy * nx - wx * ny;
    } else {
      A[0] = nx; A[1] = ny; A[2] = nz;
    }
  } else {
    A[0] = +wz * ny - wy * nz;
    A[1] = -wz * nx + wx * nz;
    A[2] = +wy * nx - wx * ny;
    A[3] = nx; A[4] = ny; A[5] = nz;
  }
  return true;
}

// Sums this CTA's share of one evaluation and writes it to partialOut[0..NV).  All threads take part.
// Layout of a partial: [n, sum b^2, nabla(noPara), hessian lower triangle(noParaSQ)].
template <bool shortIteration, bool rotationOnly, bool weighted = false>
__device__ __forceinline__ void eval_to_partial(const IcpLevelArgs &lv, const ViewParams &sv, const IcpConsts &c,
                                                const float4 *__restrict__ pointsMap, const float4 *__restrict__ normalsMap,
                                                double (*sPart)[ICP_NVALS], double *__restrict__ partialOut, int nCtas) {
  constexpr int noPara = shortIteration ? 3 : 6;
  constexpr int noParaSQ = shortIteration ? 6 : 21;
  constexpr int NV = 2 + noPara + noParaSQ;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
  const int n = lv.w * lv.h;
  for (int i = blockIdx.x * ICP_THREADS + threadIdx.x; i < n; i += nCtas * ICP_THREADS) {
    const int y = i / lv.w, x = i - y * lv.w;
    float A[noPara], b;
    float localWeight = 1.0f;
    if (weighted) {
      // ITMWeightedICPTracker_CPU.cpp:46: minSigmaZ / sigma_z * 0.5 + 0.5, minSigmaZ = 0.0012
      const float sz = __ldg(lv.weight + i);
      localWeight = sz > 0 ? 0.0012f / sz * 0.5f + 0.5f : 0.0f;
    }
    if (per_point_Ab<shortIteration, rotationOnly, weighted>(A, b, x, y, __ldg(lv.depth + i), lv, sv, c, pointsMap, normalsMap, localWeight)) {
      acc[0] += 1.0f;
      acc[1] += weighted ? b * b * localWeight * localWeight : b * b;
#pragma unroll
      for (int r = 0, counter = 0; r < noPara; r++) {
        acc[2 + r] += b * A[r];
#pragma unroll
        for (int cc = 0; cc <= r; cc++, counter++) acc[2 + noPara + counter] += A[r] * A[cc];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[i] = v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sPart[warp][i] = (double)acc[i];
  }
  __syncthreads();
  if (threadIdx.x < ICP_NVALS) {
    double s = 0.0;
    if (threadIdx.x < NV) {
#pragma unroll
      for (int w = 0; w < ICP_THREADS / 32; ++w) s += sPart[w][threadIdx.x];
    }
    partialOut[threadIdx.x] = s;
  }
  __syncthreads();  // sPart is reused by the caller
}

// Sum of the first nRows CTA partials (fp64) by the warps of one CTA; result (float) in sSums[0..32).
// Lane = value index, warp w owns a contiguous chunk of rows; loads are issued 8 deep so the chain of L2
// round trips stays short.  The summation order is fixed (rows ascending inside a chunk, chunks ascending),
// hence bit-reproducible run to run.
__device__ __forceinline__ void reduce_partials(const double *__restrict__ partials, int nRows, double (*sPart)[ICP_NVALS], float *sSums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = (nRows + ICP_THREADS / 32 - 1) / (ICP_THREADS / 32);
  const int r0 = warp * chunk, r1 = min(nRows, r0 + chunk);
  double s = 0.0;
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(partials + (size_t)(r + k) * ICP_NVALS + lane);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; r < r1; ++r) s += __ldcg(partials + (size_t)r * ICP_NVALS + lane);
  sPart[warp][lane] = s;
  __syncthreads();
  if (threadIdx.x < ICP_NVALS) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < ICP_THREADS / 32; ++w) t += sPart[w][threadIdx.x];
    sSums[threadIdx.x] = (float)t;
  }
}

#ifdef ITM_ICP_TRACE
__device__ unsigned long long g_icpTrace[64 * 32];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot, idx) if ((slot) < 64) g_icpTrace[(slot) * 32 + (idx)] = gtimer()
#define TRACE_VAL(slot, idx, v) if ((slot) < 64) g_icpTrace[(slot) * 32 + (idx)] = (unsigned long long)(v)
#ifndef ICP_TRACE_THREAD
#define ICP_TRACE_THREAD 44
#endif
#define TRACE0(slot, idx) if (blockIdx.x == 0 && threadIdx.x == ICP_TRACE_THREAD) { TRACE(slot, idx); }
__device__ __forceinline__ unsigned long long gtimer_after(float dep) {  // timestamp taken once dep is available
  unsigned long long t;
  asm volatile("{ .reg .f32 d; mov.f32 d, %1; mov.u64 %0, %globaltimer; }" : "=l"(t) : "f"(dep));
  return t;
}
#define TRACE_DEP(slot, idx, dep) if (blockIdx.x == 0 && threadIdx.x == ICP_TRACE_THREAD && (slot) < 64) g_icpTrace[(slot) * 32 + (idx)] = gtimer_after(dep)
__device__ unsigned long long g_icpCtaTrace[64 * 160 * 2];  // [evaluation][CTA]{pose received, row stored}
#define TRACE_CTA(slot, idx) if ((slot) < 64 && threadIdx.x == 0) g_icpCtaTrace[((slot) * 160 + blockIdx.x) * 2 + (idx)] = gtimer()
#else
#define TRACE_CTA(slot, idx)
#define TRACE_DEP(slot, idx, dep)
#define TRACE(slot, idx)
#define TRACE_VAL(slot, idx, v)
#define TRACE0(slot, idx)
#endif

// Levenberg-Marquardt state of one TrackCamera call; lives in the master CTA's shared memory.
struct LmShared {
  float M_d[16], params[6];           // trackingState->pose_d
  float approxInvPose[16];
  float lastGoodM[16], lastGoodParams[6];
  float Hgood[36], ngood[6];
  float step[6];
  float fOld, lambda;
  int evalCount;
  int levelEvals[ITM_MAX_LEVELS];
};

// Second half of one tracker iteration, the same for every parameter count and for both trackers (kept out of line so that
// its instructions are fetched once, not once per caller: the first call of each caller in a launch runs from a cold
// instruction cache): ApplyDelta, pose_d->SetInvM + Coerce, approxInvPose = pose_d->GetInvM(), HasConverged
// (ITMDepthTracker.cpp:190-196).  In: L.approxInvPose (the pose the step applies to), L.step.  Out: L.M_d, L.params, L.approxInvPose.
#ifndef ICP_LM_FINISH_ATTR
#define ICP_LM_FINISH_ATTR __noinline__
#endif
__device__ ICP_LM_FINISH_ATTR bool lm_finish(LmShared &L, int iterationType, float terminationThreshold, int traceSlot) {
  float approxInvPose[16], M_d[16], params[6], step[6];
#pragma unroll
  for (int i = 0; i < 16; ++i) approxInvPose[i] = L.approxInvPose[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) step[i] = L.step[i];
  TRACE(traceSlot, 13);
  icp_apply_delta(approxInvPose, step, iterationType, approxInvPose);
  TRACE(traceSlot, 14);
  pose_set_invM_coerce(approxInvPose, M_d, params);
  TRACE(traceSlot, 15);
  mat4_inv_pose(M_d, approxInvPose);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    L.M_d[i] = M_d[i];
    L.approxInvPose[i] = approxInvPose[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) L.params[i] = params[i];
  return icp_has_converged(step, terminationThreshold);
}

// The LM bookkeeping of one iteration, ITMDepthTracker.cpp:167-197.  Run by one thread; noPara is a template
// parameter so that every array index is static and the 6x6 system lives in registers.  Returns HasConverged().
template <int noPara>
__device__ __noinline__ bool lm_update(LmShared &L, const float *sSums, const float *sMeans, int iterationType, int level, bool firstIterOfLevel,
                                       float terminationThreshold, int traceSlot) {
  // (the state lives in shared memory and is touched only where an iteration needs it: this runs on one thread, every
  // instruction is on the frame's critical path)
  float fOld = L.fOld, lambda = L.lambda;
  if (firstIterOfLevel) {
    // approxInvPose = pose_d->GetInvM(); lastKnownGoodPose(*pose_d); f_old = 1e20f; lambda = 1.0  (:161-165)
#pragma unroll
    for (int i = 0; i < 16; ++i) L.lastGoodM[i] = L.M_d[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) L.lastGoodParams[i] = L.params[i];
    fOld = 1e20f;
    lambda = 1.0f;
  }
  const int noValid = (int)sSums[0];
  const float fNew = (noValid > 100) ? sqrtf(sSums[1]) / (float)noValid : 1e5f;
  L.evalCount++;
  L.levelEvals[level]++;

  // hessian_good / nabla_good: only the noPara x noPara block is ever read (entries outside it are garbage in the reference)
  float A[noPara * noPara], ngood[noPara];
  if ((noValid <= 0) || (fNew > fOld)) {
    // revert to the last known good pose (:173-177)
    float M_d[16], inv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { M_d[i] = L.lastGoodM[i]; L.M_d[i] = M_d[i]; }
#pragma unroll
    for (int i = 0; i < 6; ++i) L.params[i] = L.lastGoodParams[i];
    mat4_inv_pose(M_d, inv);
#pragma unroll
    for (int i = 0; i < 16; ++i) L.approxInvPose[i] = inv[i];
    lambda *= 10.0f;
#pragma unroll
    for (int r = 0; r < noPara; ++r) {
      ngood[r] = L.ngood[r];
#pragma unroll
      for (int c = 0; c < noPara; ++c) A[r + c * noPara] = L.Hgood[r + c * 6];
    }
  } else {
    // (L.approxInvPose == pose_d->GetInvM() already: it was set from the very same M_d)
#pragma unroll
    for (int i = 0; i < 16; ++i) L.lastGoodM[i] = L.M_d[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) L.lastGoodParams[i] = L.params[i];
    fOld = fNew;
    // hessian_good / nabla_good = new / noValidPoints (the quotients were taken by gather_rows, one per lane)
#pragma unroll
    for (int r = 0, counter = 0; r < noPara; r++) {
#pragma unroll
      for (int c = 0; c <= r; c++, counter++) {
        const float h = sMeans[2 + noPara + counter];
        A[r + c * noPara] = h;
        A[c + r * noPara] = h;
      }
    }
#pragma unroll
    for (int r = 0; r < noPara; ++r) {
      ngood[r] = sMeans[2 + r];
      L.ngood[r] = ngood[r];
#pragma unroll
      for (int c = 0; c < noPara; ++c) L.Hgood[r + c * 6] = A[r + c * noPara];
    }
    lambda /= 10.0f;
  }
#pragma unroll
  for (int i = 0; i < noPara; ++i) A[i + i * noPara] *= 1.0f + lambda;
  float step[noPara];
  TRACE(traceSlot, 12);
  cholesky_solve<noPara>(A, ngood, step);  // ComputeDelta (:85-102)
#pragma unroll
  for (int i = 0; i < 6; ++i) L.step[i] = i < noPara ? step[i < noPara ? i : 0] : 0.0f;
  L.fOld = fOld;
  L.lambda = lambda;
#ifdef ITM_ICP_TRACE_WARM  // experiment: how long does the second half take when its code was fetched a moment ago?
  {
    float save[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) save[i] = L.approxInvPose[i];
    lm_finish(L, iterationType, terminationThreshold, 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) L.approxInvPose[i] = save[i];
    TRACE(traceSlot, 21);
  }
#endif
  return lm_finish(L, iterationType, terminationThreshold, traceSlot);
}

// One iteration of ITMWeightedICPTracker::TrackCamera (ITMWeightedICPTracker.cpp:164-192): plain Gauss-Newton on the raw
// sums (no division by the point count, no damping, no step rejection - f_old stays at its initial 1e10).  Returns true when
// the level's loop ends (no valid points, f_new > f_old, or HasConverged).
template <int noPara>
__device__ __noinline__ bool gn_update(LmShared &L, const float *sSums, int iterationType, int level, float terminationThreshold) {
  const int noValid = (int)sSums[0];
  const float fNew = (noValid > 100) ? sqrtf(sSums[1]) / (float)noValid : 1e5f;
  L.evalCount++;
  L.levelEvals[level]++;
  if (noValid <= 0) return true;
  if (fNew > 1e10f) return true;
  float H[36], nabla[6];
#pragma unroll
  for (int i = 0; i < 36; ++i) H[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < 6; ++i) nabla[i] = 0.0f;
#pragma unroll
  for (int r = 0, counter = 0; r < noPara; r++) {
#pragma unroll
    for (int c = 0; c <= r; c++, counter++) {
      const float h = sSums[2 + noPara + counter];
      H[r + c * 6] = h;
      H[c + r * 6] = h;
    }
  }
#pragma unroll
  for (int r = 0; r < noPara; ++r) nabla[r] = sSums[2 + r];
  float step[6];
  icp_compute_delta(step, nabla, H, noPara == 3);
#pragma unroll
  for (int i = 0; i < 6; ++i) L.step[i] = step[i];
  return lm_finish(L, iterationType, terminationThreshold, 64);
}


// ---- self-validating 64-bit words: [63:32] tag, [31:0] payload ------------------------------------------------
// Everything that crosses CTAs inside the tracker travels as 8-byte words carrying their own tag (launch epoch and
// evaluation number), written and polled with relaxed gpu-scope accesses: a reader that sees the tag has the payload,
// so no fence, no arrival counter and no second round trip for the data are needed.
#define ICP_RING 128  // broadcast slots per launch (>= the largest possible number of evaluations, 72 for 8 levels)
#define ICP_BCAST_WORDS 20  // 16 pose words + 1 flag word, padded
// Row words are rewritten at every evaluation, so their tag carries the evaluation number next to the (truncated) launch
// epoch: a CTA that is waited for wrote its row in this very launch, and the previous content is at most one launch old,
// so 25 epoch bits cannot alias.  Broadcast slot k is written once per launch (by evaluation k) but slots beyond a launch's
// last evaluation keep older content indefinitely: their tag is the full 32-bit epoch (never 0), which only repeats after
// 2^32 - 1 launches of slots that are rewritten by every launch reaching them.
__device__ __forceinline__ unsigned icp_tag(unsigned epoch, int evalNo) { return (epoch << 7) | (unsigned)(evalNo + 1); }
__device__ __forceinline__ void word_st(unsigned long long *p, unsigned payload, unsigned tag) {
  const unsigned long long v = ((unsigned long long)tag << 32) | payload;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long word_ld(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct TrackArgs {
  IcpArgs a;
  IcpLevelArgs lv[ITM_MAX_LEVELS];
  int iters[ITM_MAX_LEVELS];
  int nLevels, noIcpLevel;
  unsigned long long *rows;   // [maxCtas][32] CTA partial sums (float payload)
  unsigned long long *bcast;  // [ICP_RING][ICP_BCAST_WORDS] pose for the next evaluation + flags
  const unsigned *epochDev;   // launch number (never 0), advanced on the device before this launch
  int chunksPerCta;           // coarse levels: 32-pixel chunks (warps) per active CTA
};

// What a thread needs to finish one pixel once the gathers have landed
struct IcpPixel {
  float wx, wy, wz;   // point in world coordinates
  float u, v;         // its projection into the raycast maps
  bool inside;
};

// what a level's pixels share: refined reciprocals of the level's focal lengths and the multiplier that turns
// i / w into one IMAD.HI: magic = ceil(2^32 / w) is exact for every i with i * (magic * w - 2^32) < 2^32, which
// w * w * h < 2^32 guarantees (1280x720: 1.2e9); wMagic = 0 (larger images, w = 1) means "divide"
struct IcpLevelDerived {
  float rcpFx, rcpFy;
  unsigned wMagic;
  bool focalOk;  // fx, fy far from the exponent limits: the inline division sequence is exact
};
__device__ __forceinline__ int icp_row_of(int i, int w, unsigned wMagic) { return wMagic ? (int)__umulhi((unsigned)i, wMagic) : i / w; }
__device__ __forceinline__ IcpLevelDerived icp_level_derived(const IcpLevelArgs &lv) {
  IcpLevelDerived d;
  const float ax = fabsf(lv.fx), ay = fabsf(lv.fy);
  d.focalOk = ax > 1e-3f && ax < 1e6f && ay > 1e-3f && ay < 1e6f;
  d.rcpFx = refined_rcp(d.focalOk ? lv.fx : 1.0f);
  d.rcpFy = refined_rcp(d.focalOk ? lv.fy : 1.0f);
  const bool magicOk = lv.w > 1 && (unsigned long long)lv.w * (unsigned long long)lv.w * (unsigned long long)lv.h < 0x100000000ull;
  d.wMagic = magicOk ? (unsigned)((0x100000000ull + (unsigned)lv.w - 1) / (unsigned)lv.w) : 0u;
  return d;
}

// first half of computePerPointGH_Depth_Ab (ITMDepthTracker.h:17-38): back-project, move into the world, project into the
// raycast maps.  The four divisions run the compiler's own IEEE fast-path sequence inline (shared reciprocals); operands
// outside its comfortable range take the ordinary `/`.
__device__ __forceinline__ void icp_project(IcpPixel &q, int x, int y, float depth, const IcpLevelArgs &lv, const IcpLevelDerived &ld,
                                            const ViewParams &sv, const IcpConsts &c) {
  q.inside = false;
  q.u = 0.0f; q.v = 0.0f;
  if (depth <= 1e-8f) return;
  const float tx = depth * safe_div((float)x - lv.cx, lv.fx, ld.rcpFx, ld.focalOk);
  const float ty = depth * safe_div((float)y - lv.cy, lv.fy, ld.rcpFy, ld.focalOk);
  mat4_mul_vec4(c.approxInvPose, tx, ty, depth, 1.0f, q.wx, q.wy, q.wz);
  float rx, ry, rz;
  mat4_mul_vec4(c.scenePose, q.wx, q.wy, q.wz, 1.0f, rx, ry, rz);
  if (rz <= 0.0f) return;
  const bool zOk = rz > 1e-3f && rz < 1e4f;
  const float yz = refined_rcp(zOk ? rz : 1.0f);
  q.u = safe_div(sv.fx * rx, rz, yz, zOk) + sv.cx;
  q.v = safe_div(sv.fy * ry, rz, yz, zOk) + sv.cy;
  q.inside = (q.u >= 0.0f) && (q.u <= (float)(sv.W - 2)) && (q.v >= 0.0f) && (q.v <= (float)(sv.H - 2));
}

struct IcpTaps {
  float4 p[4], n[4];
};

// the 4 + 4 taps of interpolateBilinear_withHoles for points and normals, issued together
template <bool EARLY>
__device__ __forceinline__ void icp_gather(IcpTaps &t, const IcpPixel &q, const float4 *__restrict__ pointsMap,
                                           const float4 *__restrict__ normalsMap, int W) {
  if (!q.inside) return;
  const int ix = (short)(int)floorf(q.u), iy = (short)(int)floorf(q.v);
  const int o = ix + iy * W;
  t.p[0] = __ldg(pointsMap + o); t.p[1] = __ldg(pointsMap + o + 1); t.p[2] = __ldg(pointsMap + o + W); t.p[3] = __ldg(pointsMap + o + W + 1);
  if (EARLY) {
    t.n[0] = __ldg(normalsMap + o); t.n[1] = __ldg(normalsMap + o + 1); t.n[2] = __ldg(normalsMap + o + W); t.n[3] = __ldg(normalsMap + o + W + 1);
  }
}

__device__ __forceinline__ float bilerp(float a, float b, float c, float d, float dx, float dy) {
  return (a * (1.0f - dx) * (1.0f - dy) + b * dx * (1.0f - dy) + c * (1.0f - dx) * dy + d * dx * dy);
}

// rest of computePerPointGH_Depth_Ab (ITMDepthTracker.h:40-77) and the accumulation of ComputeGandH (..._CPU.cpp:60-66)
// weighted (computePerPointGH_wICP, ITMWeightedICPTracker.h:66-69): sum b^2 w^2, and the normal is scaled by w after b was taken
template <bool shortIteration, bool rotationOnly, int NV, bool weighted, bool EARLY>
__device__ __forceinline__ void icp_accumulate(float *acc, const IcpPixel &q, IcpTaps &t, float distThresh, const float4 *__restrict__ normalsMap, int W,
                                               float localWeight = 1.0f) {
  constexpr int noPara = shortIteration ? 3 : 6;
  if (!q.inside) return;
  if (t.p[0].w < 0 || t.p[1].w < 0 || t.p[2].w < 0 || t.p[3].w < 0) return;
  const int ix = (short)(int)floorf(q.u), iy = (short)(int)floorf(q.v);
  const float fx = q.u - (float)ix, fy = q.v - (float)iy;
  const float cx = bilerp(t.p[0].x, t.p[1].x, t.p[2].x, t.p[3].x, fx, fy);
  const float cy = bilerp(t.p[0].y, t.p[1].y, t.p[2].y, t.p[3].y, fx, fy);
  const float cz = bilerp(t.p[0].z, t.p[1].z, t.p[2].z, t.p[3].z, fx, fy);
  const float cw = bilerp(t.p[0].w, t.p[1].w, t.p[2].w, t.p[3].w, fx, fy);
  if (cw < 0.0f) return;
  const float dx = cx - q.wx, dy = cy - q.wy, dz = cz - q.wz;
  const float dist = dx * dx + dy * dy + dz * dz;
  if (dist > distThresh) return;
  if (!EARLY) {
    const int o = ix + iy * W;
    t.n[0] = __ldg(normalsMap + o); t.n[1] = __ldg(normalsMap + o + 1); t.n[2] = __ldg(normalsMap + o + W); t.n[3] = __ldg(normalsMap + o + W + 1);
  }
  float nx, ny, nz;
  if (t.n[0].w < 0 || t.n[1].w < 0 || t.n[2].w < 0 || t.n[3].w < 0) {
    nx = 0; ny = 0; nz = 0;  // interpolateBilinear_withHoles returns (0,0,0,-1); the reference does not test it here
  } else {
    nx = bilerp(t.n[0].x, t.n[1].x, t.n[2].x, t.n[3].x, fx, fy);
    ny = bilerp(t.n[0].y, t.n[1].y, t.n[2].y, t.n[3].y, fx, fy);
    nz = bilerp(t.n[0].z, t.n[1].z, t.n[2].z, t.n[3].z, fx, fy);
  }
  const float b = nx * dx + ny * dy + nz * dz;
  if (weighted) { nx *= localWeight; ny *= localWeight; nz *= localWeight; }
  float A[noPara];
  if (shortIteration) {
    if (rotationOnly) {
      A[0] = +q.wz * ny - q.wy * nz;
      A[1] = -q.wz * nx + q.wx * nz;
      A[2] = +q.wy * nx - q.wx * ny;
    } else {
      A[0] = nx; A[1] = ny; A[2] = nz;
    }
  } else {
    A[0] = +q.wz * ny - q.wy * nz;
    A[1] = -q.wz * nx + q.wx * nz;
    A[2] = +q.wy * nx - q.wx * ny;
    A[3] = nx; A[4] = ny; A[5] = nz;
  }
  acc[0] += 1.0f;
  acc[1] += weighted ? b * b * localWeight * localWeight : b * b;
#pragma unroll
  for (int r = 0, counter = 0; r < noPara; r++) {
    acc[2 + r] += b * A[r];
#pragma unroll
    for (int cc = 0; cc <= r; cc++, counter++) acc[2 + noPara + counter] += A[r] * A[cc];
  }
}

// Warp reduction of NV (<= 32) per-thread values: instead of 5 butterfly steps per value (5*NV shuffles) the values are
// transposed while they are summed - each step halves the number of values a lane holds - so that 31 (NV > 16) or
// NV + 15 (NV <= 16) shuffles suffice and lane l ends up with the warp total of value l (l and l+16 both, if NV <= 16).
template <int NV>
__device__ __forceinline__ float warp_transpose_reduce(const float *acc) {
  const int lane = threadIdx.x & 31;
  float v[16];
  if (NV > 16) {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float lo = acc[i], hi = (i + 16 < NV) ? acc[i + 16] : 0.0f;
      const float send = up ? lo : hi, keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (i < NV) ? acc[i] + __shfl_xor_sync(0xffffffffu, acc[i], 16) : 0.0f;
  }
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = lane & h;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h], keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

// This CTA's share of one evaluation.  Pixels are dealt in chunks of 32 consecutive ones (one warp-wide, coalesced depth
// read); chunk k goes to CTA k mod nActive, warp slot k / nActive, so that a coarse level is spread over many SMs with
// one or two warps each instead of filling a few SMs: its map taps are 2^level pixels apart (one 128-byte line per lane
// and tap), and thousands of such line requests from one SM queue up behind each other (measured: 1.2-1.7 us for the
// four point taps with 16 warps per SM against 0.5 us spread out).  The CTA sum (fp32 per thread and warp, fp64 across the
// warps) is published as tagged words in rowOut[0..32).
template <bool shortIteration, bool rotationOnly, bool weighted = false>
__device__ __forceinline__ void eval_to_row(const IcpLevelArgs &lv, const IcpLevelDerived &ld, const ViewParams &sv, const IcpConsts &c,
                                            const float4 *__restrict__ pointsMap, const float4 *__restrict__ normalsMap,
                                            double (*sPart)[ICP_NVALS], unsigned long long *rowOut, unsigned tag, int cta, int nActive,
                                            int traceSlot) {
  constexpr int noPara = shortIteration ? 3 : 6;
  constexpr int noParaSQ = shortIteration ? 6 : 21;
  constexpr int NV = 2 + noPara + noParaSQ;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
  const int n = lv.w * lv.h;
  const int nChunks = (n + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chStride = nActive * (ICP_THREADS / 32);
  int ch = cta + warp * nActive;
  // the depth of a warp's next chunk is fetched while the current one is worked on (fine levels: several chunks per warp)
  float dNext = (ch < nChunks && ch * 32 + lane < n) ? __ldg(lv.depth + ch * 32 + lane) : 0.0f;
  for (; ch < nChunks; ch += chStride) {
    const int i = ch * 32 + lane;
    if (i >= n) break;
    IcpPixel q1;
    const int y1 = icp_row_of(i, lv.w, ld.wMagic);
    TRACE0(traceSlot, 20);
    const float d1 = dNext;
    {
      const int iN = i + chStride * 32;
      if (ch + chStride < nChunks && iN < n) dNext = __ldg(lv.depth + iN);
    }
    TRACE_DEP(traceSlot, 16, d1);
    icp_project(q1, i - y1 * lv.w, y1, d1, lv, ld, sv, c);
    TRACE_DEP(traceSlot, 17, q1.u + q1.v);
    float localWeight = 1.0f;
    if (weighted) {
      // ITMWeightedICPTracker_CPU.cpp:46: minSigmaZ / sigma_z * 0.5 + 0.5, minSigmaZ = 0.0012
      const float sz = __ldg(lv.weight + i);
      localWeight = sz > 0 ? 0.0012f / sz * 0.5f + 0.5f : 0.0f;
    }
    IcpTaps t1;
    constexpr bool EARLY = ICP_EARLY_NORMALS == 1 || (ICP_EARLY_NORMALS == 2 && shortIteration);
    icp_gather<EARLY>(t1, q1, pointsMap, normalsMap, sv.W);
    if (q1.inside) { TRACE_DEP(traceSlot, 18, t1.p[0].x + t1.p[1].y + t1.p[2].z + t1.p[3].w); }
    icp_accumulate<shortIteration, rotationOnly, NV, weighted, EARLY>(acc, q1, t1, lv.distThresh, normalsMap, sv.W, localWeight);
    TRACE_DEP(traceSlot, 19, acc[0] + acc[1] + acc[2] + acc[NV - 1]);
  }
  TRACE0(traceSlot, 8);
  const float tot = warp_transpose_reduce<NV>(acc);  // lane l: warp total of value l
  if (lane < NV) sPart[warp][lane] = (double)tot;
  TRACE0(traceSlot, 9);
  __syncthreads();
  TRACE0(traceSlot, 10);
  if (threadIdx.x < ICP_NVALS) {
    double s = 0.0;
    if (threadIdx.x < NV) {
#pragma unroll
      for (int w = 0; w < ICP_THREADS / 32; ++w) s += sPart[w][threadIdx.x];
    }
    word_st(rowOut + threadIdx.x, __float_as_uint((float)s), tag);
  }
  TRACE0(traceSlot, 11);
  TRACE_CTA(traceSlot, 1);
  __syncthreads();  // sPart is reused by the caller
}

// Master only: waits for the first nRows tagged rows and sums them (fp64, fixed order: rows ascending inside each of
// the interleaved per-warp groups, groups ascending) -> sSums[0..32).  Each thread has all of its rows' loads in flight at once.
#define ICP_GROUPS (ICP_THREADS / 32)
#define ICP_ROWS_PER_THREAD ((ICP_MAX_CTAS + ICP_GROUPS - 1) / ICP_GROUPS)
__device__ __forceinline__ void gather_rows(const unsigned long long *rows, int nRows, unsigned tag, double (*sPart)[ICP_NVALS], float *sSums,
                                            float *sMeans) {
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  double s = 0.0;
  static_assert(ICP_ROWS_PER_THREAD <= 10, "all of a thread's rows are kept in flight together");
  constexpr int CH = ICP_ROWS_PER_THREAD;
  unsigned long long w[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) w[k] = 0ull;  // tag 0 is never used
  // poll every outstanding row in the same round: the loads of one round are independent, so a round costs one L2 round
  // trip whatever the number of rows still missing (polling row after row cost one trip per late row: up to 3 us)
  bool pending;
  do {
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int r = g + ICP_GROUPS * k;
      if (r < nRows && (unsigned)(w[k] >> 32) != tag) w[k] = word_ld(rows + (size_t)r * ICP_NVALS + lane);
    }
    pending = false;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int r = g + ICP_GROUPS * k;
      pending = pending || (r < nRows && (unsigned)(w[k] >> 32) != tag);
    }
  } while (pending);
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int r = g + ICP_GROUPS * k;
    if (r < nRows) s += (double)__uint_as_float((unsigned)w[k]);
  }
  sPart[g][lane] = s;
  __syncthreads();
  if (threadIdx.x < ICP_NVALS) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < ICP_THREADS / 32; ++q) t += sPart[q][threadIdx.x];
    const float v = (float)t;
    sSums[threadIdx.x] = v;
    // hessian / noValidPoints and nabla / noValidPoints of the LM loop (ITMDepthTracker.cpp:181-182), one quotient per lane
    // instead of 27 dependent-issue divisions on the thread that runs the update
    const float cnt = (float)(int)__shfl_sync(0xffffffffu, v, 0);
    sMeans[threadIdx.x] = v / cnt;
  }
  __syncthreads();
}

// One launch = one TrackCamera.  Must be launched cooperatively (all CTAs co-resident: they wait for each other).
// CTA 0 is the master: it keeps the LM state in shared memory, collects every active CTA's row, runs the LM update and
// broadcasts the next pose together with the "level finished" flag in slot evalNo of the ring; everybody (active or
// not) follows the ring, so all CTAs walk the same sequence of levels and iterations.
// (A fixed master keeps the long straight-line LM code warm in one SM's instruction cache.)
// WICP: ITMWeightedICPTracker (per-pixel weights from the depth uncertainty, Gauss-Newton) instead of ITMDepthTracker (LM).
template <bool WICP>
__global__ void __launch_bounds__(ICP_THREADS, ICP_CTAS_PER_SM) k_icp_track(TrackArgs t) {
  __shared__ IcpConsts c;
  __shared__ double sPart[ICP_THREADS / 32][ICP_NVALS];
  __shared__ float sSums[ICP_NVALS], sMeans[ICP_NVALS];
  __shared__ LmShared L;
  __shared__ IcpLevelDerived sLd;
  __shared__ unsigned sFlags[2];  // double buffered by evaluation parity: a warp may run one evaluation ahead of a reader
  __shared__ IcpLevelArgs sLv;
  __shared__ ViewParams sSv;
  FrameState *st = t.a.st;
  const float4 *pointsMap = reinterpret_cast<const float4 *>(t.a.pointsMap);
  const float4 *normalsMap = reinterpret_cast<const float4 *>(t.a.normalsMap);
  const int nCtas = gridDim.x;
  pdl_wait();  // (launched cooperatively, without programmatic serialisation: a no-op that keeps the chain's rule)
  // (a master that evaluates no pixels - polling from the start of an evaluation, its instruction caches holding the update
  // code only - was measured: 86.1 against 84.4 us per frame, the 148th evaluating CTA is worth more)
  const bool master = blockIdx.x == 0;
  const int evalCta = (int)blockIdx.x;

  if (master && threadIdx.x == 0) { TRACE(63, 0); }
  if (threadIdx.x < 16) {
    c.scenePose[threadIdx.x] = st->scenePose[threadIdx.x];
    c.approxInvPose[threadIdx.x] = st->invM_d[threadIdx.x];  // pose_d->GetInvM() on entering the first level
  }
  if (master) {
    if (threadIdx.x < 16) {
      L.M_d[threadIdx.x] = st->M_d[threadIdx.x];
      L.approxInvPose[threadIdx.x] = st->invM_d[threadIdx.x];
    }
    if (threadIdx.x < 6) {
      L.params[threadIdx.x] = st->poseParams[threadIdx.x];
      L.ngood[threadIdx.x] = 0.0f;
    }
    // hessian_good / nabla_good are uninitialised stack variables in the reference (:151-153); start from zero
    if (threadIdx.x >= 32 && threadIdx.x < 68) L.Hgood[threadIdx.x - 32] = 0.0f;
    if (threadIdx.x == 0) { L.fOld = 1e10f; L.lambda = 1.0f; L.evalCount = 0; }
    if (threadIdx.x >= 96 && threadIdx.x < 96 + ITM_MAX_LEVELS) L.levelEvals[threadIdx.x - 96] = 0;
  }
  __syncthreads();

  const unsigned epoch = *t.epochDev;  // stable for the whole launch: only earlier launches write it
  int evalNo = 0;
  for (int level = t.nLevels - 1; level >= t.noIcpLevel; --level) {
    const int type = t.lv[level].iterationType;
    if (type == ITM_ITER_NONE) continue;
    __syncthreads();
    if (threadIdx.x == 0) { sLv = t.lv[level]; sSv = t.a.sceneVp; sLd = icp_level_derived(t.lv[level]); }
    __syncthreads();
    const IcpLevelArgs &lv = sLv;
    // coarse levels have fewer 32-pixel chunks than the grid has warps: the first nActive CTAs evaluate, with about
    // chunksPerCta warps each (see eval_to_row)
    const int nChunks = (lv.w * lv.h + 31) >> 5;
    const int nActive = max(1, min(nCtas, (nChunks + t.chunksPerCta - 1) / t.chunksPerCta));
    const int NV = (type == ITM_ITER_BOTH) ? 29 : 11;
    const int nIters = t.iters[level];
    for (int it = 0; it < nIters; ++it, ++evalNo) {
      const unsigned tag = icp_tag(epoch, evalNo);
      if (master && threadIdx.x == 0) { TRACE(evalNo, 0); }
      if (evalCta >= 0 && evalCta < nActive) {
        TRACE_CTA(evalNo, 0);
        unsigned long long *myRow = t.rows + (size_t)evalCta * ICP_NVALS;
        if (type == ITM_ITER_ROTATION) eval_to_row<true, true, WICP>(lv, sLd, sSv, c, pointsMap, normalsMap, sPart, myRow, tag, evalCta, nActive, evalNo);
        else if (type == ITM_ITER_TRANSLATION) eval_to_row<true, false, WICP>(lv, sLd, sSv, c, pointsMap, normalsMap, sPart, myRow, tag, evalCta, nActive, evalNo);
        else eval_to_row<false, false, WICP>(lv, sLd, sSv, c, pointsMap, normalsMap, sPart, myRow, tag, evalCta, nActive, evalNo);
      }
      unsigned long long *slot = t.bcast + (size_t)(evalNo & (ICP_RING - 1)) * ICP_BCAST_WORDS;
      if (master) {
        if (threadIdx.x == 0) { TRACE(evalNo, 1); }
        gather_rows(t.rows, nActive, tag, sPart, sSums, sMeans);
        if (threadIdx.x == 0) {
          TRACE(evalNo, 3);
          bool conv;
          if (WICP) conv = (NV == 11) ? gn_update<3>(L, sSums, type, level, t.a.terminationThreshold) : gn_update<6>(L, sSums, type, level, t.a.terminationThreshold);
          else conv = (NV == 11) ? lm_update<3>(L, sSums, sMeans, type, level, it == 0, t.a.terminationThreshold, evalNo)
                                 : lm_update<6>(L, sSums, sMeans, type, level, it == 0, t.a.terminationThreshold, evalNo);
          TRACE(evalNo, 4);
          TRACE_VAL(evalNo, 6, level);
          sFlags[evalNo & 1] = (conv || it == nIters - 1) ? 1u : 0u;
        }
      }
      unsigned long long w = 0ull;
      if (!master && threadIdx.x < 17) {
        while ((unsigned)((w = word_ld(slot + threadIdx.x)) >> 32) != epoch) { /* spin */ }
      }
      __syncthreads();  // the update is done (master) / every warp has finished reading the old pose (dry run)
      if (master) {
        if (threadIdx.x < 16) {
          const float v = L.approxInvPose[threadIdx.x];
          c.approxInvPose[threadIdx.x] = v;
          word_st(slot + threadIdx.x, __float_as_uint(v), epoch);
        } else if (threadIdx.x == 16) {
          word_st(slot + 16, sFlags[evalNo & 1], epoch);
        }
      } else if (threadIdx.x < 16) {
        c.approxInvPose[threadIdx.x] = __uint_as_float((unsigned)w);
      } else if (threadIdx.x == 16) {
        sFlags[evalNo & 1] = (unsigned)w;
      }
      __syncthreads();
      if (master && threadIdx.x == 0) { TRACE(evalNo, 5); }
      if (sFlags[evalNo & 1] & 1u) { ++evalNo; break; }
    }
  }
  if (master && threadIdx.x == 0) {
    float inv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { st->M_d[i] = L.M_d[i]; inv[i] = L.approxInvPose[i]; }
#pragma unroll
    for (int i = 0; i < 16; ++i) st->invM_d[i] = inv[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) st->poseParams[i] = L.params[i];
    st->icp.evalCount = L.evalCount;
#pragma unroll
    for (int i = 0; i < ITM_MAX_LEVELS; ++i) st->icp.levelEvals[i] = L.levelEvals[i];
    TRACE(63, 1);
  }
}

// Stand-alone evaluation at poseIn (16 floats, device): leaves ComputeGandH's results in out44.
template <bool shortIteration, bool rotationOnly, bool weighted = false>
__global__ void __launch_bounds__(ICP_THREADS) k_icp_eval_single(IcpArgs a, IcpLevelArgs lv, float *__restrict__ out44,
                                                                 const float *__restrict__ poseIn) {
  constexpr int noPara = shortIteration ? 3 : 6;
  __shared__ IcpConsts c;
  __shared__ double sPart[ICP_THREADS / 32][ICP_NVALS];
  __shared__ float sSums[ICP_NVALS];
  __shared__ bool sIsLast;
  if (threadIdx.x < 16) c.approxInvPose[threadIdx.x] = poseIn[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 48) c.scenePose[threadIdx.x - 32] = a.st->scenePose[threadIdx.x - 32];
  __syncthreads();
  eval_to_partial<shortIteration, rotationOnly, weighted>(lv, a.sceneVp, c, reinterpret_cast<const float4 *>(a.pointsMap),
                                                reinterpret_cast<const float4 *>(a.normalsMap), sPart,
                                                a.partials + (size_t)blockIdx.x * ICP_NVALS, gridDim.x);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sIsLast = (atomicAdd(a.ctaCounter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!sIsLast) return;
  __threadfence();
  reduce_partials(a.partials, gridDim.x, sPart, sSums);
  __syncthreads();
  if (threadIdx.x == 0) {
    *a.ctaCounter = 0;
    // ComputeGandH's return values (ITMDepthTracker_CPU.cpp:72-78)
    const int noValid = (int)sSums[0];
    out44[0] = sSums[0];
    out44[1] = (noValid > 100) ? sqrtf(sSums[1]) / (float)noValid : 1e5f;
    for (int r = 0; r < 6; ++r) out44[2 + r] = r < noPara ? sSums[2 + r] : 0.0f;
    for (int i = 0; i < 36; ++i) out44[8 + i] = 0.0f;
    for (int r = 0, counter = 0; r < noPara; r++)
      for (int cc = 0; cc <= r; cc++, counter++) {
        out44[8 + r + cc * 6] = sSums[2 + noPara + counter];
        out44[8 + cc + r * 6] = sSums[2 + noPara + counter];
      }
  }
}

__global__ void k_set_pose(FrameState *st) {
  float inv[16];
  mat4_inv(st->M_d, inv);
  for (int i = 0; i < 16; ++i) st->invM_d[i] = inv[i];
}

}  // namespace

namespace itm {

int icp_max_ctas() { return ICP_MAX_CTAS; }

void launch_set_pose(FrameState *st, cudaStream_t s) { k_set_pose<<<1, 1, 0, s>>>(st); }

// grid for the persistent tracker: as many CTAs as can be co-resident, capped at 2 per SM
int icp_track_grid() {
  static int gridOf[64] = {0};  // per device
  int dev = 0, sms = 0, perSm = 0;
  cudaGetDevice(&dev);
  int &grid = gridOf[dev & 63];
  if (grid) return grid;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_icp_track<false>, ICP_THREADS, 0);
  if (perSm > ICP_CTAS_PER_SM) perSm = ICP_CTAS_PER_SM;
  if (perSm < 1) perSm = 1;
  grid = sms * perSm;
  if (grid > icp_max_ctas()) grid = icp_max_ctas();
  return grid;
}

__global__ void k_icp_bump(unsigned *epochDev) { icp_bump_epoch(epochDev); }

cudaError_t launch_icp_track(const IcpArgs &a, const IcpLevelArgs *levels, const int *iters, int nLevels, int noIcpLevel,
                             unsigned long long *rows, unsigned long long *bcast, unsigned *epochDev, bool bumpEpoch, int gridCap,
                             cudaStream_t s, bool weighted) {
  TrackArgs t;
  t.a = a;
  for (int l = 0; l < ITM_MAX_LEVELS; ++l) {
    if (l < nLevels) {
      t.lv[l] = levels[l];
      t.iters[l] = iters[l];
    } else {
      t.lv[l] = IcpLevelArgs();
      t.lv[l].iterationType = ITM_ITER_NONE;
      t.iters[l] = 0;
    }
  }
  t.nLevels = nLevels;
  t.noIcpLevel = noIcpLevel;
  t.rows = rows;
  t.bcast = bcast;
  t.epochDev = epochDev;
  static int chunks = 0;  // ITM_B200_ICP_CHUNKS: A/B measurements
  if (!chunks) {
    const char *e = getenv("ITM_B200_ICP_CHUNKS");
    chunks = e ? atoi(e) : ICP_CHUNKS_PER_CTA;
    if (chunks < 1 || chunks > ICP_THREADS / 32) chunks = ICP_CHUNKS_PER_CTA;
  }
  t.chunksPerCta = chunks;
  if (bumpEpoch) k_icp_bump<<<1, 1, 0, s>>>(epochDev);
  void *args[] = {&t};
  int grid = icp_track_grid();
  if (gridCap > 0 && gridCap < grid) grid = gridCap;
  return cudaLaunchCooperativeKernel(weighted ? (const void *)k_icp_track<true> : (const void *)k_icp_track<false>, dim3(grid), dim3(ICP_THREADS),
                                     args, 0, s);
}

size_t icp_rows_bytes() { return (size_t)icp_max_ctas() * ICP_NVALS * sizeof(unsigned long long); }
size_t icp_bcast_bytes() { return (size_t)ICP_RING * ICP_BCAST_WORDS * sizeof(unsigned long long); }

#ifdef ITM_ICP_TRACE
extern "C" int itm_b200_debug_icp_cta_trace(unsigned long long *out) {
  return (int)cudaMemcpyFromSymbol(out, g_icpCtaTrace, sizeof(unsigned long long) * 64 * 160 * 2);
}
extern "C" int itm_b200_debug_icp_trace(unsigned long long *out2048) {
  return (int)cudaMemcpyFromSymbol(out2048, g_icpTrace, sizeof(unsigned long long) * 2048);
}
#endif

void launch_icp_eval_single(const IcpArgs &a, const IcpLevelArgs &lv, float *out44, const float *poseIn, cudaStream_t s) {
  const int n = lv.w * lv.h;
  int ctas = (n + ICP_THREADS - 1) / ICP_THREADS;
  if (ctas > icp_max_ctas()) ctas = icp_max_ctas();
  if (ctas < 1) ctas = 1;
  if (lv.weight) {
    switch (lv.iterationType) {
      case ITM_ITER_ROTATION: k_icp_eval_single<true, true, true><<<ctas, ICP_THREADS, 0, s>>>(a, lv, out44, poseIn); break;
      case ITM_ITER_TRANSLATION: k_icp_eval_single<true, false, true><<<ctas, ICP_THREADS, 0, s>>>(a, lv, out44, poseIn); break;
      case ITM_ITER_BOTH: k_icp_eval_single<false, false, true><<<ctas, ICP_THREADS, 0, s>>>(a, lv, out44, poseIn); break;
      default: break;
    }
    return;
  }
  switch (lv.iterationType) {
    case ITM_ITER_ROTATION:
      k_icp_eval_single<true, true><<<ctas, ICP_THREADS, 0, s>>>(a, lv, out44, poseIn);
      break;
    case ITM_ITER_TRANSLATION:
      k_icp_eval_single<true, false><<<ctas, ICP_THREADS, 0, s>>>(a, lv, out44, poseIn);
      break;
    case ITM_ITER_BOTH:
      k_icp_eval_single<false, false><<<ctas, ICP_THREADS, 0, s>>>(a, lv, out44, poseIn);
      break;
    default:
      break;
  }
}

}  // namespace itm
