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
#include "itm_common.cuh"
#include "kernels.h"
#include "pose_math.cuh"

namespace {

using namespace itm;

#define ICP_THREADS 256
#define ICP_NVALS 32  // 1 count + 1 f + 6 nabla + 21 hessian, padded

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

template <bool shortIteration, bool rotationOnly>
__device__ __forceinline__ bool per_point_Ab(float *A, float &b, int x, int y, float depth, const IcpLevelArgs &lv, const ViewParams &sv,
                                             const IcpConsts &c, const float4 *__restrict__ pointsMap,
                                             const float4 *__restrict__ normalsMap) {
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
template <bool shortIteration, bool rotationOnly>
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
    if (per_point_Ab<shortIteration, rotationOnly>(A, b, x, y, __ldg(lv.depth + i), lv, sv, c, pointsMap, normalsMap)) {
      acc[0] += 1.0f;
      acc[1] += b * b;
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

// Sum of the first nRows CTA partials (fp64) by the 8 warps of one CTA; result (float) in sSums[0..32).
// Lane = value index, warp w owns a contiguous chunk of rows; loads are issued 8 deep so the chain of L2
// round trips stays short.  The summation order is fixed (rows ascending inside a chunk, chunks ascending),
// hence bit-reproducible run to run.
__device__ __forceinline__ void reduce_partials(const double *__restrict__ partials, int nRows, double (*sPart)[ICP_NVALS], float *sSums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = (nRows + 7) >> 3;
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

// Levenberg-Marquardt state of one TrackCamera call; lives in the master CTA's shared memory.
struct LmShared {
  float M_d[16], params[6];           // trackingState->pose_d
  float approxInvPose[16];
  float lastGoodM[16], lastGoodParams[6];
  float Hgood[36], ngood[6];
  float fOld, lambda;
  int evalCount;
  int levelEvals[ITM_MAX_LEVELS];
};

// The LM bookkeeping of one iteration, ITMDepthTracker.cpp:167-197.  Run by one thread; noPara is a template
// parameter so that every array index is static and the 6x6 system lives in registers.  Returns HasConverged().
template <int noPara>
__device__ bool lm_update(LmShared &L, const float *sSums, int iterationType, int level, bool firstIterOfLevel, float terminationThreshold) {
  float M_d[16], params[6], approxInvPose[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) M_d[i] = L.M_d[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) params[i] = L.params[i];
  float fOld = L.fOld, lambda = L.lambda;
  if (firstIterOfLevel) {
    // approxInvPose = pose_d->GetInvM(); lastKnownGoodPose(*pose_d); f_old = 1e20f; lambda = 1.0  (:161-165)
#pragma unroll
    for (int i = 0; i < 16; ++i) L.lastGoodM[i] = M_d[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) L.lastGoodParams[i] = params[i];
    fOld = 1e20f;
    lambda = 1.0f;
  }
  const int noValid = (int)sSums[0];
  const float fNew = (noValid > 100) ? sqrtf(sSums[1]) / (float)noValid : 1e5f;
  L.evalCount++;
  L.levelEvals[level]++;

  float Hgood[36], ngood[6];
  if ((noValid <= 0) || (fNew > fOld)) {
    // revert to the last known good pose (:173-177)
#pragma unroll
    for (int i = 0; i < 16; ++i) M_d[i] = L.lastGoodM[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) params[i] = L.lastGoodParams[i];
    mat4_inv(M_d, approxInvPose);
    lambda *= 10.0f;
#pragma unroll
    for (int i = 0; i < 36; ++i) Hgood[i] = L.Hgood[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) ngood[i] = L.ngood[i];
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) approxInvPose[i] = L.approxInvPose[i];  // == pose_d->GetInvM() on entering a level
#pragma unroll
    for (int i = 0; i < 16; ++i) L.lastGoodM[i] = M_d[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) L.lastGoodParams[i] = params[i];
    fOld = fNew;
    // hessian_good / nabla_good = new / noValidPoints.  Entries outside the noPara block are garbage in
    // the reference (never read); zero here.
#pragma unroll
    for (int i = 0; i < 36; ++i) Hgood[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) ngood[i] = 0.0f;
#pragma unroll
    for (int r = 0, counter = 0; r < noPara; r++) {
#pragma unroll
      for (int c = 0; c <= r; c++, counter++) {
        const float h = sSums[2 + noPara + counter] / (float)noValid;
        Hgood[r + c * 6] = h;
        Hgood[c + r * 6] = h;
      }
    }
#pragma unroll
    for (int r = 0; r < noPara; ++r) ngood[r] = sSums[2 + r] / (float)noValid;
#pragma unroll
    for (int i = 0; i < 36; ++i) L.Hgood[i] = Hgood[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) L.ngood[i] = ngood[i];
    lambda /= 10.0f;
  }
  float A[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) A[i] = Hgood[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) A[i + i * 6] *= 1.0f + lambda;
  float step[6];
  icp_compute_delta(step, ngood, A, noPara == 3);
  icp_apply_delta(approxInvPose, step, iterationType, approxInvPose);
  pose_set_invM_coerce(approxInvPose, M_d, params);
  mat4_inv(M_d, approxInvPose);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    L.M_d[i] = M_d[i];
    L.approxInvPose[i] = approxInvPose[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) L.params[i] = params[i];
  L.fOld = fOld;
  L.lambda = lambda;
  return icp_has_converged(step, terminationThreshold);
}

#ifdef ITM_ICP_TRACE
__device__ unsigned long long g_icpTrace[64 * 8];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot, idx) if ((slot) < 64) g_icpTrace[(slot) * 8 + (idx)] = gtimer()
#define TRACE_VAL(slot, idx, v) if ((slot) < 64) g_icpTrace[(slot) * 8 + (idx)] = (unsigned long long)(v)
#else
#define TRACE(slot, idx)
#define TRACE_VAL(slot, idx, v)
#endif

struct TrackArgs {
  IcpArgs a;
  IcpLevelArgs lv[ITM_MAX_LEVELS];
  int iters[ITM_MAX_LEVELS];
  int nLevels, noIcpLevel;
  unsigned *barrier;  // [0] arrival count, [1] release word: (sequence << 1) | levelDone
};

// One launch = one TrackCamera.  Must be launched cooperatively (all CTAs co-resident).
// CTA 0 is the master: it keeps the LM state in shared memory, waits for every CTA's partial sums, runs the LM
// update and publishes the next pose + the "level finished" bit through the release word the others spin on.
// (A fixed master keeps the long straight-line LM code warm in one SM's instruction cache.)
__global__ void __launch_bounds__(ICP_THREADS, 2) k_icp_track(TrackArgs t) {
  __shared__ IcpConsts c;
  __shared__ double sPart[ICP_THREADS / 32][ICP_NVALS];
  __shared__ float sSums[ICP_NVALS];
  __shared__ LmShared L;
  __shared__ unsigned sRelease;
  FrameState *st = t.a.st;
  const float4 *pointsMap = reinterpret_cast<const float4 *>(t.a.pointsMap);
  const float4 *normalsMap = reinterpret_cast<const float4 *>(t.a.normalsMap);
  unsigned *bCount = t.barrier;
  volatile unsigned *bRelease = t.barrier + 1;
  const int nCtas = gridDim.x;
  const bool master = blockIdx.x == 0;

  if (master && threadIdx.x == 0) { TRACE(63, 0); }
  if (threadIdx.x < 16) c.scenePose[threadIdx.x] = st->scenePose[threadIdx.x];
  if (threadIdx.x == 0) sRelease = *bRelease;  // nobody can release before this CTA has arrived
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
  unsigned seq = sRelease >> 1;

  int evalNo = 0;
  for (int level = t.nLevels - 1; level >= t.noIcpLevel; --level) {
    const IcpLevelArgs lv = t.lv[level];
    const int type = lv.iterationType;
    if (type == ITM_ITER_NONE) continue;
    // coarse levels have fewer pixels than the grid has threads: only the first nActive CTAs evaluate
    const int nActive = min(nCtas, (lv.w * lv.h + ICP_THREADS - 1) / ICP_THREADS);
    const int NV = (type == ITM_ITER_BOTH) ? 29 : 11;
    for (int it = 0; it < t.iters[level]; ++it, ++evalNo) {
      // pose to evaluate at: pose_d->GetInvM() on entering a level, the LM loop's approxInvPose afterwards - the
      // master published either one in st->icp.approxInvPose before the last release (st->invM_d before the first)
      if (blockIdx.x < nActive) {
        if (threadIdx.x < 16) {
          c.approxInvPose[threadIdx.x] = master ? L.approxInvPose[threadIdx.x]
                                                : __ldcg((evalNo == 0 ? st->invM_d : st->icp.approxInvPose) + threadIdx.x);
        }
        __syncthreads();
        if (master && threadIdx.x == 0) { TRACE(evalNo, 0); }
        double *myPartial = t.a.partials + (size_t)blockIdx.x * ICP_NVALS;
        if (type == ITM_ITER_ROTATION) eval_to_partial<true, true>(lv, t.a.sceneVp, c, pointsMap, normalsMap, sPart, myPartial, nActive);
        else if (type == ITM_ITER_TRANSLATION) eval_to_partial<true, false>(lv, t.a.sceneVp, c, pointsMap, normalsMap, sPart, myPartial, nActive);
        else eval_to_partial<false, false>(lv, t.a.sceneVp, c, pointsMap, normalsMap, sPart, myPartial, nActive);
        __threadfence();
      }
      // every CTA arrives (idle ones at once): the master can then never run a whole iteration ahead of a CTA
      // that has not started yet, which keeps the sequence number read at kernel start consistent grid-wide
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(bCount, 1u);
      ++seq;
      if (master) {
        if (threadIdx.x == 0) {
          TRACE(evalNo, 1);
          while (*((volatile unsigned *)bCount) != (unsigned)nCtas) { /* wait for every partial */ }
          __threadfence();
          TRACE(evalNo, 2);
        }
        __syncthreads();
        reduce_partials(t.a.partials, nActive, sPart, sSums);
        __syncthreads();
        if (threadIdx.x == 0) {
          TRACE(evalNo, 3);
          const bool conv = (NV == 11) ? lm_update<3>(L, sSums, type, level, it == 0, t.a.terminationThreshold)
                                       : lm_update<6>(L, sSums, type, level, it == 0, t.a.terminationThreshold);
          TRACE(evalNo, 4);
          TRACE_VAL(evalNo, 6, level);
          const bool lastOfLevel = conv || it == t.iters[level] - 1;
#pragma unroll
          for (int i = 0; i < 16; ++i) st->icp.approxInvPose[i] = L.approxInvPose[i];
          *bCount = 0;
          __threadfence();
          sRelease = (seq << 1) | (lastOfLevel ? 1u : 0u);
          *bRelease = sRelease;  // release
        }
      } else if (threadIdx.x == 0) {
        unsigned v;
        while (((v = *bRelease) >> 1) != seq) { /* spin */ }
        __threadfence();
        sRelease = v;
      }
      __syncthreads();
      if (master && threadIdx.x == 0) { TRACE(evalNo, 5); }
      const bool done = (sRelease & 1u) != 0;
      if (done) { ++evalNo; break; }
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
template <bool shortIteration, bool rotationOnly>
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
  eval_to_partial<shortIteration, rotationOnly>(lv, a.sceneVp, c, reinterpret_cast<const float4 *>(a.pointsMap),
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

int icp_max_ctas() { return 148 * 2; }

void launch_set_pose(FrameState *st, cudaStream_t s) { k_set_pose<<<1, 1, 0, s>>>(st); }

// grid for the persistent tracker: as many CTAs as can be co-resident, capped at 2 per SM
int icp_track_grid() {
  static int grid = 0;
  if (grid) return grid;
  int dev = 0, sms = 0, perSm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_icp_track, ICP_THREADS, 0);
  if (perSm > 2) perSm = 2;
  if (perSm < 1) perSm = 1;
  grid = sms * perSm;
  if (grid > icp_max_ctas()) grid = icp_max_ctas();
  return grid;
}

cudaError_t launch_icp_track(const IcpArgs &a, const IcpLevelArgs *levels, const int *iters, int nLevels, int noIcpLevel,
                             unsigned *barrier, cudaStream_t s) {
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
  t.barrier = barrier;
  void *args[] = {&t};
  return cudaLaunchCooperativeKernel((const void *)k_icp_track, dim3(icp_track_grid()), dim3(ICP_THREADS), args, 0, s);
}

#ifdef ITM_ICP_TRACE
extern "C" int itm_b200_debug_icp_trace(unsigned long long *out512) {
  return (int)cudaMemcpyFromSymbol(out512, g_icpTrace, sizeof(unsigned long long) * 512);
}
#endif

void launch_icp_eval_single(const IcpArgs &a, const IcpLevelArgs &lv, float *out44, const float *poseIn, cudaStream_t s) {
  const int n = lv.w * lv.h;
  int ctas = (n + ICP_THREADS - 1) / ICP_THREADS;
  if (ctas > icp_max_ctas()) ctas = icp_max_ctas();
  if (ctas < 1) ctas = 1;
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
