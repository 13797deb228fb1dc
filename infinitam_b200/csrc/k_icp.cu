// Point-to-plane ICP: per-level G/H normal-equation reduction and the Levenberg-Marquardt loop.
//
// Replaces:
//   ITMDepthTracker::TrackCamera          ITMLib/Engine/ITMDepthTracker.cpp:145-199  (LM loop)
//   ITMDepthTracker_CPU::ComputeGandH     ITMLib/Engine/DeviceSpecific/CPU/ITMDepthTracker_CPU.cpp:14-79
//   computePerPointGH_Depth(_Ab)          ITMLib/Engine/DeviceAgnostic/ITMDepthTracker.h:9-105
//   interpolateBilinear_withHoles         ITMLib/Engine/DeviceAgnostic/ITMPixelUtils.h:41-71
//
// B200 design.  The reference CUDA tracker does memset + kernel + blocking 29-word D2H copy +
// host Cholesky for each of up to 30 evaluations per frame.  Here one launch per evaluation
// does everything on the device:
//   * each thread accumulates its pixels' (count, b^2, b*A, A*A^T) in registers (fp32),
//   * warps reduce with shuffles, the CTA combines its warps in fp64 and writes one 30-value
//     partial, then takes a ticket;
//   * the last CTA to finish sums the partials in CTA order (fp64, bit-reproducible run to
//     run) and runs the accept/reject + Cholesky + SE(3) update of the LM loop (pose_math.cuh)
//     on the pose kept in FrameState, including the early-exit flag of HasConverged().
// So a whole TrackCamera is <= 30 back-to-back launches with no host involvement.
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

// The LM bookkeeping of one iteration, ITMDepthTracker.cpp:167-197.  Run by one thread.
__device__ void lm_update(FrameState *st, const float *sums /*[0]=n [1]=sumF [2..7]=nabla [8..28]=hessian lower tri*/, int noPara,
                          int iterationType, bool firstIterOfLevel, float terminationThreshold) {
  IcpState &s = st->icp;
  if (firstIterOfLevel) {
    // approxInvPose = pose_d->GetInvM(); lastKnownGoodPose(*pose_d); f_old = 1e20f; lambda = 1.0  (:161-165)
    mat4_inv(st->M_d, s.approxInvPose);
    for (int i = 0; i < 16; ++i) s.lastGoodM[i] = st->M_d[i];
    for (int i = 0; i < 6; ++i) s.lastGoodParams[i] = st->poseParams[i];
    s.fOld = 1e20f;
    s.lambda = 1.0f;
    s.levelDone = 0;
  }
  const int noValid = (int)sums[0];
  const float fNew = (noValid > 100) ? sqrtf(sums[1]) / (float)noValid : 1e5f;
  float hessianNew[36], nablaNew[6];
  for (int i = 0; i < 36; ++i) hessianNew[i] = 0.0f;
  for (int i = 0; i < 6; ++i) nablaNew[i] = 0.0f;
  for (int r = 0, counter = 0; r < noPara; r++)
    for (int c = 0; c <= r; c++, counter++) hessianNew[r + c * 6] = sums[8 + counter];
  for (int r = 0; r < noPara; ++r)
    for (int c = r + 1; c < noPara; c++) hessianNew[r + c * 6] = hessianNew[c + r * 6];
  for (int r = 0; r < noPara; ++r) nablaNew[r] = sums[2 + r];
  s.lastNoValid = noValid;
  s.lastF = fNew;
  s.evalCount++;

  float approxInvPose[16];
  if ((noValid <= 0) || (fNew > s.fOld)) {
    // revert
    for (int i = 0; i < 16; ++i) st->M_d[i] = s.lastGoodM[i];
    for (int i = 0; i < 6; ++i) st->poseParams[i] = s.lastGoodParams[i];
    mat4_inv(st->M_d, approxInvPose);
    s.lambda *= 10.0f;
  } else {
    for (int i = 0; i < 16; ++i) { s.lastGoodM[i] = st->M_d[i]; approxInvPose[i] = s.approxInvPose[i]; }
    for (int i = 0; i < 6; ++i) s.lastGoodParams[i] = st->poseParams[i];
    s.fOld = fNew;
    for (int i = 0; i < 36; ++i) s.hessianGood[i] = hessianNew[i] / (float)noValid;
    for (int i = 0; i < 6; ++i) s.nablaGood[i] = nablaNew[i] / (float)noValid;
    s.lambda /= 10.0f;
  }
  float A[36];
  for (int i = 0; i < 36; ++i) A[i] = s.hessianGood[i];
  for (int i = 0; i < 6; ++i) A[i + i * 6] *= 1.0f + s.lambda;
  float step[6];
  icp_compute_delta(step, s.nablaGood, A, iterationType != ITM_ITER_BOTH);
  icp_apply_delta(approxInvPose, step, iterationType, approxInvPose);
  pose_set_invM_coerce(approxInvPose, st->M_d, st->poseParams);
  mat4_inv(st->M_d, s.approxInvPose);
  for (int i = 0; i < 16; ++i) st->invM_d[i] = s.approxInvPose[i];
  if (icp_has_converged(step, terminationThreshold)) s.levelDone = 1;
}

// mode 0: tracking fast path (LM update on device).  mode 1: evaluate at the pose given in
// out44[0..15] (approxInvPose, device memory) and leave [n, f, nabla6, hessian36] in out44.
template <bool shortIteration, bool rotationOnly>
__global__ void __launch_bounds__(ICP_THREADS) k_icp_eval(IcpArgs a, IcpLevelArgs lv, int firstIterOfLevel, int mode,
                                                          float *__restrict__ out44, const float *__restrict__ poseIn) {
  constexpr int noPara = shortIteration ? 3 : 6;
  constexpr int noParaSQ = shortIteration ? 6 : 21;
  constexpr int NV = 2 + noPara + noParaSQ;
  __shared__ IcpConsts c;
  __shared__ double sPart[ICP_THREADS / 32][ICP_NVALS];
  __shared__ bool sIsLast;
  FrameState *st = a.st;
  if (mode == 0 && !firstIterOfLevel && st->icp.levelDone) return;  // HasConverged() broke out of this level
  if (threadIdx.x == 0) {
    if (mode == 1) {
      for (int i = 0; i < 16; ++i) c.approxInvPose[i] = poseIn[i];
    } else if (firstIterOfLevel) {
      mat4_inv(st->M_d, c.approxInvPose);  // approxInvPose = pose_d->GetInvM()  (:161)
    } else {
      for (int i = 0; i < 16; ++i) c.approxInvPose[i] = st->icp.approxInvPose[i];
    }
  }
  if (threadIdx.x >= 32 && threadIdx.x < 48) c.scenePose[threadIdx.x - 32] = st->scenePose[threadIdx.x - 32];
  __syncthreads();

  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
  const float4 *pointsMap = reinterpret_cast<const float4 *>(a.pointsMap);
  const float4 *normalsMap = reinterpret_cast<const float4 *>(a.normalsMap);
  const int n = lv.w * lv.h;
  for (int i = blockIdx.x * ICP_THREADS + threadIdx.x; i < n; i += gridDim.x * ICP_THREADS) {
    const int y = i / lv.w, x = i - y * lv.w;
    float A[noPara], b;
    if (per_point_Ab<shortIteration, rotationOnly>(A, b, x, y, __ldg(lv.depth + i), lv, a.sceneVp, c, pointsMap, normalsMap)) {
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
  // warp reduce (fp32), CTA reduce (fp64)
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
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < ICP_THREADS / 32; ++w) s += sPart[w][threadIdx.x];
    a.partials[(size_t)blockIdx.x * ICP_NVALS + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(a.ctaCounter, 1u);
    sIsLast = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!sIsLast) return;
  __threadfence();
  __shared__ float sSums[ICP_NVALS];
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (unsigned cta = 0; cta < gridDim.x; ++cta) s += __ldcg(a.partials + (size_t)cta * ICP_NVALS + threadIdx.x);
    sSums[threadIdx.x] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *a.ctaCounter = 0;
    // unpack to the common layout [n, sumF, nabla(6), hessian lower triangle(21)]
    float sums[2 + 6 + 21];
    for (int i = 0; i < 29; ++i) sums[i] = 0.0f;
    sums[0] = sSums[0];
    sums[1] = sSums[1];
    for (int r = 0; r < noPara; ++r) sums[2 + r] = sSums[2 + r];
    for (int i = 0; i < noParaSQ; ++i) sums[8 + i] = sSums[2 + noPara + i];
    if (mode == 0) {
      lm_update(st, sums, noPara, lv.iterationType, firstIterOfLevel != 0, a.terminationThreshold);
    } else {
      // ComputeGandH's return values (ITMDepthTracker_CPU.cpp:72-78)
      const int noValid = (int)sums[0];
      out44[0] = sums[0];
      out44[1] = (noValid > 100) ? sqrtf(sums[1]) / (float)noValid : 1e5f;
      for (int r = 0; r < 6; ++r) out44[2 + r] = r < noPara ? sums[2 + r] : 0.0f;
      for (int i = 0; i < 36; ++i) out44[8 + i] = 0.0f;
      for (int r = 0, counter = 0; r < noPara; r++)
        for (int cc = 0; cc <= r; cc++, counter++) out44[8 + r + cc * 6] = sums[8 + counter];
      for (int r = 0; r < noPara; ++r)
        for (int cc = r + 1; cc < noPara; cc++) out44[8 + r + cc * 6] = out44[8 + cc + r * 6];
    }
  }
}

__global__ void k_icp_begin_frame(FrameState *st) {
  IcpState &s = st->icp;
  for (int i = 0; i < 36; ++i) s.hessianGood[i] = 0.0f;
  for (int i = 0; i < 6; ++i) s.nablaGood[i] = 0.0f;
  s.levelDone = 0;
  s.evalCount = 0;
  s.fOld = 1e10f;
  s.lambda = 1.0f;
}

__global__ void k_set_pose(FrameState *st) {
  float inv[16];
  mat4_inv(st->M_d, inv);
  for (int i = 0; i < 16; ++i) st->invM_d[i] = inv[i];
}

}  // namespace

namespace itm {

int icp_max_ctas() { return 148 * 2; }

void launch_icp_begin_frame(FrameState *st, cudaStream_t s) { k_icp_begin_frame<<<1, 1, 0, s>>>(st); }

void launch_set_pose(FrameState *st, cudaStream_t s) { k_set_pose<<<1, 1, 0, s>>>(st); }

void launch_icp_eval(const IcpArgs &a, const IcpLevelArgs &lv, int firstIterOfLevel, int mode, float *out44, const float *poseIn,
                     cudaStream_t s) {
  const int n = lv.w * lv.h;
  int ctas = (n + ICP_THREADS - 1) / ICP_THREADS;
  if (ctas > icp_max_ctas()) ctas = icp_max_ctas();
  if (ctas < 1) ctas = 1;
  switch (lv.iterationType) {
    case ITM_ITER_ROTATION:
      k_icp_eval<true, true><<<ctas, ICP_THREADS, 0, s>>>(a, lv, firstIterOfLevel, mode, out44, poseIn);
      break;
    case ITM_ITER_TRANSLATION:
      k_icp_eval<true, false><<<ctas, ICP_THREADS, 0, s>>>(a, lv, firstIterOfLevel, mode, out44, poseIn);
      break;
    case ITM_ITER_BOTH:
      k_icp_eval<false, false><<<ctas, ICP_THREADS, 0, s>>>(a, lv, firstIterOfLevel, mode, out44, poseIn);
      break;
    default:
      break;
  }
}

}  // namespace itm
