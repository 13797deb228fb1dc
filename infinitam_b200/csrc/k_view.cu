// View building: raw short depth -> float metres, and the tracker's depth pyramid.
//
// Replaces:
//   convertDepthAffineToFloat   ITMLib/Engine/DeviceAgnostic/ITMViewBuilder.h:22-28
//   filterSubsampleWithHoles    ITMLib/Engine/DeviceAgnostic/ITMLowLevelEngine.h:26-47
//   ITMDepthTracker::PrepareForEvaluation   ITMLib/Engine/ITMDepthTracker.cpp:62-75
//
// B200 design: the reference launches one kernel for the conversion and one per
// pyramid level (5 launches, each re-reading the previous level from memory).
// Here one CTA owns a 32x32 tile of the full-resolution image, converts it, and
// reduces it through shared memory to 16x16, 8x8, 4x4 and 2x2 - the raw frame is
// read once (2 B/px) and every level written once, in a single launch.
#include "itm_common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ float convert_affine(short d, float a, float b) {
  return ((d <= 0) || (d > 32000)) ? -1.0f : (float)d * a + b;
}

// convertDisparityToDepth (ITMViewBuilder.h:7-20): Kinect raw disparity, depth = 8 * c2 * fx / (c1 - d)
__device__ __forceinline__ float convert_disparity(short d, float c1, float c2, float fxDepth) {
  const float disparity_tmp = c1 - (float)d;
  float depth;
  if (disparity_tmp == 0) depth = 0.0f;
  else depth = 8.0f * c2 * fxDepth / disparity_tmp;
  return (depth > 0) ? depth : -1.0f;
}
// fxDisparity != 0 selects the Kinect disparity conversion with (a, b) = disparityCalib.params
__device__ __forceinline__ float convert_raw(short d, float a, float b, float fxDisparity) {
  return fxDisparity != 0.0f ? convert_disparity(d, a, b, fxDisparity) : convert_affine(d, a, b);
}

// mean of the >0 entries of a 2x2 quad, accumulation order (0,0) (1,0) (0,1) (1,1)
__device__ __forceinline__ float subsample4(float p00, float p10, float p01, float p11) {
  float out = 0.0f, n = 0.0f;
  if (p00 > 0.0f) { out += p00; n++; }
  if (p10 > 0.0f) { out += p10; n++; }
  if (p01 > 0.0f) { out += p01; n++; }
  if (p11 > 0.0f) { out += p11; n++; }
  if (n > 0) out /= n;
  return out;
}

struct PyramidArgs {
  float *level[ITM_MAX_LEVELS];  // level[0] = full resolution
  int w[ITM_MAX_LEVELS], h[ITM_MAX_LEVELS];
  int nLevels;                   // <= 5 handled by the fused kernel
  itm::FramePrologue pro;        // pose-independent chores of later stages, folded into this launch (all NULL: none)
};

// One CTA (256 threads) per 32x32 full-res tile.  raw may be NULL (then level 0 is
// taken as already converted and only the pyramid is built).
__global__ void __launch_bounds__(256) k_convert_pyramid(const short *__restrict__ raw, float a, float b, float fxDisparity, PyramidArgs args) {
  __shared__ float s0[32][33];
  __shared__ float s1[16][17];
  __shared__ float s2[8][9];
  __shared__ float s3[4][5];
  pdl_wait();
  pdl_trigger();
  const int W = args.w[0], H = args.h[0];
  const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * 32;
  const int tid = threadIdx.x;
  float *__restrict__ out0 = args.level[0];

  // Frame prologue (ProcessFrame path only).  Two small launches of later stages do not depend on the camera pose that
  // the tracker is about to compute, so they ride along here instead of sitting on the frame's critical path:
  //  * AllocateSceneFromDepth's first loop: last frame's visible entries become type 3 (..._CPU.cpp:159-160), and the
  //    free-list heads the allocation scan counts down from are snapshot;
  //  * CreateExpectedDepths' initialisation of the whole min/max image (ITMVisualisationEngine_CPU.cpp:104-108).
  {
    const int nThreads = gridDim.x * gridDim.y * 256, gtid = (blockIdx.y * gridDim.x + blockIdx.x) * 256 + tid;
    if (args.pro.visType) {
      FrameState *st = args.pro.st;
      if (gtid == 0) {
        st->allocBaseBlockId = st->lastFreeBlockId;
        st->allocBaseExcessId = st->lastFreeExcessId;
      }
      const int n = st->noVisibleEntries;
      for (int i = gtid; i < n; i += nThreads) {
        const int id = __ldg(args.pro.visibleIds + i);
        args.pro.visType[id] = 3;
        if (args.pro.claimBits) atomicOr(args.pro.claimBits + (id >> 5), 1u << (id & 31));
      }
    }
    if (args.pro.icpEpoch && gtid == 0) itm::icp_bump_epoch(args.pro.icpEpoch);
    if (args.pro.minmax) {
      float4 *mm4 = reinterpret_cast<float4 *>(args.pro.minmax);
      const float4 init = make_float4(ITM_FAR_AWAY, ITM_VERY_CLOSE, ITM_FAR_AWAY, ITM_VERY_CLOSE);
      for (int i = gtid; i < args.pro.minmaxPixels / 2; i += nThreads) mm4[i] = init;
      if (gtid == 0 && (args.pro.minmaxPixels & 1)) args.pro.minmax[args.pro.minmaxPixels - 1] = make_float2(ITM_FAR_AWAY, ITM_VERY_CLOSE);
    }
  }

  // level 0: 4 pixels per thread, rows coalesced
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int lx = tid & 31, ly = (tid >> 5) + 8 * k;
    const int x = tx0 + lx, y = ty0 + ly;
    float v = 0.0f;
    if (x < W && y < H) {
      if (raw) {
        v = convert_raw(__ldg(raw + x + y * W), a, b, fxDisparity);
        out0[x + y * W] = v;
      } else {
        v = out0[x + y * W];
      }
    }
    s0[ly][lx] = v;
  }
  __syncthreads();
  if (args.nLevels > 1) {
    const int lx = tid & 15, ly = tid >> 4;
    const float v = subsample4(s0[2 * ly][2 * lx], s0[2 * ly][2 * lx + 1], s0[2 * ly + 1][2 * lx], s0[2 * ly + 1][2 * lx + 1]);
    s1[ly][lx] = v;
    const int x = (tx0 >> 1) + lx, y = (ty0 >> 1) + ly;
    if (x < args.w[1] && y < args.h[1]) args.level[1][x + y * args.w[1]] = v;
  }
  __syncthreads();
  if (args.nLevels > 2 && tid < 64) {
    const int lx = tid & 7, ly = tid >> 3;
    const float v = subsample4(s1[2 * ly][2 * lx], s1[2 * ly][2 * lx + 1], s1[2 * ly + 1][2 * lx], s1[2 * ly + 1][2 * lx + 1]);
    s2[ly][lx] = v;
    const int x = (tx0 >> 2) + lx, y = (ty0 >> 2) + ly;
    if (x < args.w[2] && y < args.h[2]) args.level[2][x + y * args.w[2]] = v;
  }
  __syncthreads();
  if (args.nLevels > 3 && tid < 16) {
    const int lx = tid & 3, ly = tid >> 2;
    const float v = subsample4(s2[2 * ly][2 * lx], s2[2 * ly][2 * lx + 1], s2[2 * ly + 1][2 * lx], s2[2 * ly + 1][2 * lx + 1]);
    s3[ly][lx] = v;
    const int x = (tx0 >> 3) + lx, y = (ty0 >> 3) + ly;
    if (x < args.w[3] && y < args.h[3]) args.level[3][x + y * args.w[3]] = v;
  }
  __syncthreads();
  if (args.nLevels > 4 && tid < 4) {
    const int lx = tid & 1, ly = tid >> 1;
    const float v = subsample4(s3[2 * ly][2 * lx], s3[2 * ly][2 * lx + 1], s3[2 * ly + 1][2 * lx], s3[2 * ly + 1][2 * lx + 1]);
    const int x = (tx0 >> 4) + lx, y = (ty0 >> 4) + ly;
    if (x < args.w[4] && y < args.h[4]) args.level[4][x + y * args.w[4]] = v;
  }
}

// stand-alone pieces for the stage-level C ABI
__global__ void k_convert_only(const short *__restrict__ raw, float *__restrict__ out, int n, float a, float b, float fxDisparity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = convert_raw(__ldg(raw + i), a, b, fxDisparity);
}

__global__ void k_subsample_holes(float *__restrict__ out, const float *__restrict__ in, int wOut, int hOut, int wIn) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wOut || y >= hOut) return;
  const float *p = in + 2 * x + 2 * y * wIn;
  out[x + y * wOut] = subsample4(__ldg(p), __ldg(p + 1), __ldg(p + wIn), __ldg(p + wIn + 1));
}

// ---- the image helpers of ITMLowLevelEngine that only the colour / Ren trackers call (DeviceAgnostic/ITMLowLevelEngine.h) ----

// filterSubsample (:7-24): 2x2 box on uchar4, integer mean
__global__ void k_subsample_rgba(uchar4 *__restrict__ out, const uchar4 *__restrict__ in, int wOut, int hOut, int wIn) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wOut || y >= hOut) return;
  const uchar4 *p = in + 2 * x + 2 * y * wIn;
  const uchar4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + wIn), d = __ldg(p + wIn + 1);
  out[x + y * wOut] = make_uchar4((unsigned char)(((int)a.x + b.x + c.x + d.x) / 4), (unsigned char)(((int)a.y + b.y + c.y + d.y) / 4),
                                  (unsigned char)(((int)a.z + b.z + c.z + d.z) / 4), (unsigned char)(((int)a.w + b.w + c.w + d.w) / 4));
}

// filterSubsampleWithHoles for Vector4f (:49-72): mean of the taps with w >= 0 (all four components, w included), in tap
// order; (0, 0, 0, -1) if there is none
__global__ void k_subsample_holes4(float4 *__restrict__ out, const float4 *__restrict__ in, int wOut, int hOut, int wIn) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wOut || y >= hOut) return;
  const float4 *p = in + 2 * x + 2 * y * wIn;
  const float4 t[4] = {__ldg(p), __ldg(p + 1), __ldg(p + wIn), __ldg(p + wIn + 1)};
  float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  float good = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t[k].w >= 0) {
      acc.x += t[k].x; acc.y += t[k].y; acc.z += t[k].z; acc.w += t[k].w;
      good += 1.0f;
    }
  if (good > 0) { acc.x /= good; acc.y /= good; acc.z /= good; acc.w /= good; }
  else acc.w = -1.0f;
  out[x + y * wOut] = acc;
}

// gradientX / gradientY (:74-124): 3x3 Sobel on the colour channels, integer arithmetic, / 8 truncating; w = 255.  Only
// interior pixels are written: the reference driver clears the first W*H*sizeof(Vector3s) bytes of the Vector4s image
// beforehand (ITMLowLevelEngine_CPU.cpp:87,101), i.e. border pixels in its last quarter keep what they held - the
// launcher does the same.
template <bool ALONG_X>
__global__ void k_gradient(short4 *__restrict__ grad, const uchar4 *__restrict__ image, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < 1 || y < 1 || x >= W - 1 || y >= H - 1) return;
  int dx[3], dy[3], dz[3];
#pragma unroll
  for (int k = -1; k <= 1; ++k) {
    const uchar4 hi = ALONG_X ? __ldg(image + (x + 1) + (y + k) * W) : __ldg(image + (x + k) + (y + 1) * W);
    const uchar4 lo = ALONG_X ? __ldg(image + (x - 1) + (y + k) * W) : __ldg(image + (x + k) + (y - 1) * W);
    // each difference is stored in a short before it is combined (no truncation: |difference| <= 255)
    dx[k + 1] = (int)hi.x - (int)lo.x; dy[k + 1] = (int)hi.y - (int)lo.y; dz[k + 1] = (int)hi.z - (int)lo.z;
  }
  grad[x + y * W] = make_short4((short)((dx[0] + 2 * dx[1] + dx[2]) / 8), (short)((dy[0] + 2 * dy[1] + dy[2]) / 8),
                                (short)((dz[0] + 2 * dz[1] + dz[2]) / 8), (short)((2 * 255 + 2 * (2 * 255) + 2 * 255) / 8));
}

// filterDepth (ITMLib/Engine/DeviceAgnostic/ITMViewBuilder.h:31-56): 5x5 bilateral filter with the Kinect noise model as range
// sigma; interior pixels only, the 2-pixel border stays 0 (DepthFiltering clears the output first, ITMViewBuilder_CPU.cpp:116-128)
#define ITM_MEAN_SIGMA_L 1.2232f
__global__ void __launch_bounds__(256) k_filter_depth(float *__restrict__ out, const float *__restrict__ in, int W, int H) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  if (x < 2 || y < 2 || x >= W - 2 || y >= H - 2) {
    out[x + y * W] = 0.0f;
    return;
  }
  const float z = __ldg(in + x + y * W);
  if (z < 0.0f) {
    out[x + y * W] = -1.0f;
    return;
  }
  const float sigma_z = 1.0f / (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * 0.25f);
  float final_depth = 0.0f, w_sum = 0.0f;
#pragma unroll
  for (int i = -2; i <= 2; i++) {
#pragma unroll
    for (int j = -2; j <= 2; j++) {
      const float tmpz = __ldg(in + (x + j) + (y + i) * W);
      if (tmpz < 0.0f) continue;
      float dz = (tmpz - z);
      dz *= dz;
      const float w = expf(-0.5f * ((float)(abs(i) + abs(j)) * ITM_MEAN_SIGMA_L * ITM_MEAN_SIGMA_L + dz * sigma_z * sigma_z));
      w_sum += w;
      final_depth += w * tmpz;
    }
  }
  out[x + y * W] = final_depth / w_sum;
}

// computeNormalAndWeight (ITMViewBuilder.h:59-114): normal from the 4-neighbourhood of the unprojected depth and the
// depth uncertainty sigma_z of the Kinect noise model; interior pixels only (ComputeNormalAndWeights, ..._CPU.cpp:130-143),
// everything else keeps its previous content
__global__ void __launch_bounds__(256) k_normal_weight(float4 *__restrict__ normalOut, float *__restrict__ sigmaOut,
                                                       const float *__restrict__ depth, int W, int H, float4 intr) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x < 2 || y < 2 || x >= W - 2 || y >= H - 2) return;
  const int idx = x + y * W;
  const float z = __ldg(depth + idx);
  const float zxp = __ldg(depth + idx + 1), zyp = __ldg(depth + idx + W), zxm = __ldg(depth + idx - 1), zym = __ldg(depth + idx - W);
  if (z < 0.0f || zxp <= 0 || zyp <= 0 || zxm <= 0 || zym <= 0) {
    normalOut[idx].w = -1.0f;
    sigmaOut[idx] = -1.0f;
    return;
  }
  // "unprojected" exactly as the reference writes it: z * (u - c) * f (it multiplies by the focal length)
  const float fx = (float)x, fy = (float)y;
  const float xp1x = zxp * ((fx + 1.0f) - intr.z) * intr.x, xp1y = zxp * (fy - intr.w) * intr.y;
  const float xm1x = zxm * ((fx - 1.0f) - intr.z) * intr.x, xm1y = zxm * (fy - intr.w) * intr.y;
  const float yp1x = zyp * (fx - intr.z) * intr.x, yp1y = zyp * ((fy + 1.0f) - intr.w) * intr.y;
  const float ym1x = zym * (fx - intr.z) * intr.x, ym1y = zym * ((fy - 1.0f) - intr.w) * intr.y;
  const float dxx = xp1x - xm1x, dxy = xp1y - xm1y, dxz = zxp - zxm;
  const float dyx = yp1x - ym1x, dyy = yp1y - ym1y, dyz = zyp - zym;
  float nx = (dxy * dyz - dxz * dyy);
  float ny = (dxz * dyx - dxx * dyz);
  float nz = (dxx * dyy - dxy * dyx);
  if (nx == 0.0f && ny == 0 && nz == 0) {
    normalOut[idx].w = -1.0f;
    sigmaOut[idx] = -1.0f;
    return;
  }
  const float norm = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
  nx *= norm; ny *= norm; nz *= norm;
  normalOut[idx] = make_float4(nx, ny, nz, 1.0f);
  const float theta = acosf(nz);
  const float theta_diff = theta / (3.1415926535897932384626433832795f * 0.5f - theta);
  sigmaOut[idx] = (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * theta_diff * theta_diff);
}

}  // namespace

namespace itm {

void launch_filter_depth(float *out, const float *in, int W, int H, cudaStream_t s) {
  dim3 g((W + 31) / 32, (H + 7) / 8);
  k_filter_depth<<<g, 256, 0, s>>>(out, in, W, H);
}

void launch_normal_weight(float *normalOut, float *sigmaOut, const float *depth, int W, int H, const float intr[4], cudaStream_t s) {
  dim3 g((W + 31) / 32, (H + 7) / 8);
  k_normal_weight<<<g, 256, 0, s>>>(reinterpret_cast<float4 *>(normalOut), sigmaOut, depth, W, H, make_float4(intr[0], intr[1], intr[2], intr[3]));
}

void launch_convert_depth(const short *raw, float *out, int n, float a, float b, cudaStream_t s, float fxDisparity) {
  k_convert_only<<<(n + 255) / 256, 256, 0, s>>>(raw, out, n, a, b, fxDisparity);
}

void launch_subsample_holes(float *out, const float *in, int wIn, int hIn, cudaStream_t s) {
  const int wOut = wIn / 2, hOut = hIn / 2;
  dim3 b(32, 8), g((wOut + 31) / 32, (hOut + 7) / 8);
  k_subsample_holes<<<g, b, 0, s>>>(out, in, wOut, hOut, wIn);
}

void launch_subsample_rgba(unsigned char *out, const unsigned char *in, int wIn, int hIn, cudaStream_t s) {
  const int wOut = wIn / 2, hOut = hIn / 2;
  dim3 b(32, 8), g((wOut + 31) / 32, (hOut + 7) / 8);
  k_subsample_rgba<<<g, b, 0, s>>>(reinterpret_cast<uchar4 *>(out), reinterpret_cast<const uchar4 *>(in), wOut, hOut, wIn);
}

void launch_subsample_holes4(float *out, const float *in, int wIn, int hIn, cudaStream_t s) {
  const int wOut = wIn / 2, hOut = hIn / 2;
  dim3 b(32, 8), g((wOut + 31) / 32, (hOut + 7) / 8);
  k_subsample_holes4<<<g, b, 0, s>>>(reinterpret_cast<float4 *>(out), reinterpret_cast<const float4 *>(in), wOut, hOut, wIn);
}

cudaError_t launch_gradient(short *grad, const unsigned char *image, int W, int H, int alongX, cudaStream_t s) {
  const cudaError_t e = cudaMemsetAsync(grad, 0, (size_t)W * H * 6, s);  // sizeof(Vector3s), like the reference driver
  if (e != cudaSuccess) return e;
  dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
  if (alongX) k_gradient<true><<<g, b, 0, s>>>(reinterpret_cast<short4 *>(grad), reinterpret_cast<const uchar4 *>(image), W, H);
  else k_gradient<false><<<g, b, 0, s>>>(reinterpret_cast<short4 *>(grad), reinterpret_cast<const uchar4 *>(image), W, H);
  return cudaSuccess;
}

// Fused conversion + pyramid.  levels[0] is the full-resolution float depth; levels 1.. are
// the tracker's view hierarchy.  Falls back to per-level launches above 5 levels or when a
// level's children would straddle tiles (never for even dims >= level count).
void launch_view_pyramid(const short *raw, float a, float b, float *const *levels, int W, int H, int nLevels, cudaStream_t s,
                         const FramePrologue *prologue, float fxDisparity) {
  PyramidArgs args;
  if (prologue) args.pro = *prologue;
  else args.pro = FramePrologue{nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  int w = W, h = H;
  for (int l = 0; l < ITM_MAX_LEVELS; ++l) {
    args.level[l] = l < nLevels ? levels[l] : nullptr;
    args.w[l] = w;
    args.h[l] = h;
    w /= 2;
    h /= 2;
  }
  const int fused = nLevels < 5 ? nLevels : 5;
  args.nLevels = fused;
  dim3 g((W + 31) / 32, (H + 31) / 32);
  launch_pdl(k_convert_pyramid, g, dim3(256), s, raw, a, b, fxDisparity, args);
  for (int l = fused; l < nLevels; ++l) launch_subsample_holes(levels[l], levels[l - 1], args.w[l - 1], args.h[l - 1], s);
}

}  // namespace itm
