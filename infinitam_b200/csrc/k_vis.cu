// Visualisation beyond the tracking raycast: forward projection of the last raycast
// (useApproximateRaycast), free-view rendering and the visible list of an arbitrary camera.
//
// Replaces
//   ForwardRender_common     ITMLib/Engine/DeviceSpecific/CPU/ITMVisualisationEngine_CPU.cpp:289-354
//     forwardProjectPixel    ITMLib/Engine/DeviceAgnostic/ITMVisualisationEngine.h:160-173
//     processPixelForwardRender<true>                                          :351-366
//   FindVisibleBlocks        ITMVisualisationEngine_CPU.cpp:40-77  (checkBlockVisibility<false>)
//   RenderImage_common       ITMVisualisationEngine_CPU.cpp:191-240
//     processPixelGrey / Colour / Normal   DeviceAgnostic/ITMVisualisationEngine.h:368-410
//     computeSingleNormalFromSDF, readFromSDF_color4u_interpolated  DeviceAgnostic/ITMRepresentationAccess.h:186-337
//   ITMTrackingState::TrackerFarFromPointCloud  ITMLib/Objects/ITMTrackingState.h:41-59
//
// B200 design notes
//  * Forward projection is a scatter with write conflicts that the reference resolves by raster order (the last
//    source pixel wins).  atomicMax of (source index + 1) into a per-pixel key gives exactly that winner; the gather
//    kernel then copies the winning point, resets the key for the next frame, applies the "missing point" test and
//    appends the pixels that need a ray to a list with one warp-aggregated atomic.  Only the listed pixels are marched
//    (dense warps instead of a mostly idle full-image launch).  The list order is scratch (the reference's is raster
//    order); the images do not depend on it.
//  * FindVisibleBlocks reuses the single-pass ordered scan of the allocation kernels, so visibleEntryIDs come out in
//    ascending slot order like the reference's serial loop.
//  * RenderImage fuses the raycast and the shading of a pixel: the ray's end point never leaves registers before the
//    32-tap normal is taken, and the taps hit the voxel blocks the march just pulled into L1.
//  * Whether a frame needs a full raycast (ITMTrackingController::Track) is decided on the device from the pose the
//    tracker just produced, so that Layer B still enqueues a whole frame without a host round trip: both variants are
//    enqueued and the kernels of the one not taken return at once.
#include "itm_common.cuh"
#include "kernels.h"
#include "raycast.cuh"
#include "scan_util.cuh"
#include "visibility.cuh"

namespace {

using namespace itm;

// ---------------------------------------------------------------- full / approximate decision

// trackingState->requiresFullRendering = TrackerFarFromPointCloud() || !useApproximateRaycast   (ITMTrackingController.cpp:15)
__global__ void k_track_decide(FrameState *st, int useApproximateRaycast) {
  if (threadIdx.x != 0) return;
  bool far = false;
  const int age = st->agePointCloud;
  if (age < 0) far = true;
  else if (age > 5) far = true;
  else {
    // cameraCenter = -1.0f * (R^T * T) for both poses; (R^T)(r, c) = M[r*4 + c] in column-major storage
    const float *A = st->scenePose, *B = st->M_d;
    float ca[3], cb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      ca[r] = -1.0f * (A[r * 4 + 0] * A[12] + A[r * 4 + 1] * A[13] + A[r * 4 + 2] * A[14]);
      cb[r] = -1.0f * (B[r * 4 + 0] * B[12] + B[r * 4 + 1] * B[13] + B[r * 4 + 2] * B[14]);
    }
    const float dx = ca[0] - cb[0], dy = ca[1] - cb[1], dz = ca[2] - cb[2];
    const float diff = dx * dx + dy * dy + dz * dz;
    if (diff > 0.0005f) far = true;
  }
  st->requiresFullRendering = (far || !useApproximateRaycast) ? 1 : 0;
}

// ---------------------------------------------------------------- ForwardRender

__global__ void __launch_bounds__(256) k_fwd_project(const float4 *__restrict__ pointsRay, int *__restrict__ key,
                                                     FrameState *__restrict__ st, ViewParams vp, float voxelSize, int gated) {
  if (gated && st->requiresFullRendering) return;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= vp.W || y >= vp.H) return;
  const int locId = x + y * vp.W;
  if (locId == 0) st->noFwdProjMissingPoints = 0;
  const float4 p = __ldg(pointsRay + locId);
  // forwardProjectPixel(pixel * voxelSize, M, projParams, imgSize)
  float cx3, cy3, cz3;
  mat4_mul_vec4(st->M_d, p.x * voxelSize, p.y * voxelSize, p.z * voxelSize, 1.0f, cx3, cy3, cz3);
  const float ix = vp.fx * cx3 / cz3 + vp.cx;
  const float iy = vp.fy * cy3 / cz3 + vp.cy;
  // a NaN passes the reference's range test, but its float -> int conversion then yields a negative index: no write either way
  if (!(ix >= 0) || !(ix <= (float)(vp.W - 1)) || !(iy >= 0) || !(iy <= (float)(vp.H - 1))) return;
  const int locNew = (int)(ix + 0.5f) + (int)(iy + 0.5f) * vp.W;
  atomicMax(key + locNew, locId + 1);
}

// one thread per destination pixel: fetch the winning source point, reset the key, list the pixel when it needs a ray
__global__ void __launch_bounds__(256) k_fwd_gather(const float4 *__restrict__ pointsRay, int *__restrict__ key,
                                                    float4 *__restrict__ fwd, const float *__restrict__ depth,
                                                    const float2 *__restrict__ minmax, int *__restrict__ missing, FrameState *st,
                                                    ViewParams vp, int gated) {
  if (gated && st->requiresFullRendering) return;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  bool isMissing = false;
  int locId = 0;
  if (x < vp.W && y < vp.H) {
    locId = x + y * vp.W;
    const int k = key[locId];
    float4 p = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // renderState->forwardProjection->Clear()
    if (k > 0) {
      key[locId] = 0;
      p = __ldg(pointsRay + (k - 1));
    }
    fwd[locId] = p;
    const int locId2 = (int)floorf((float)x / (float)ITM_MINMAX_SUBSAMPLE) + (int)floorf((float)y / (float)ITM_MINMAX_SUBSAMPLE) * vp.W;
    const float2 mm = __ldg(minmax + locId2);
    const float d = __ldg(depth + locId);
    isMissing = (p.w <= 0) && ((p.x == 0 && p.y == 0 && p.z == 0) || (d >= 0)) && (mm.x < mm.y);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, isMissing);
  if (ballot) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&st->noFwdProjMissingPoints, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (isMissing) missing[base + __popc(ballot & ((1u << lane) - 1u))] = locId;
  }
}

template <int VW>
__global__ void __launch_bounds__(128) k_fwd_cast(const void *__restrict__ voxels, const void *__restrict__ table,
                                                  const float2 *__restrict__ minmax, const int *__restrict__ missing,
                                                  float4 *__restrict__ fwd, const FrameState *__restrict__ st, ViewParams vp,
                                                  SceneParams sp, int gated) {
  if (gated && st->requiresFullRendering) return;
  __shared__ float sInvM[16];
  if (threadIdx.x < 16) sInvM[threadIdx.x] = st->invM_d[threadIdx.x];
  __syncthreads();
  const int n = st->noFwdProjMissingPoints;
  VoxelReader<VW> rd;
  rd.init(voxels, table, sp.nBuckets, sp.hashMask);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int locId = __ldg(missing + i);
    const int y = locId / vp.W, x = locId - y * vp.W;
    const int locId2 = (int)floorf((float)x / (float)ITM_MINMAX_SUBSAMPLE) + (int)floorf((float)y / (float)ITM_MINMAX_SUBSAMPLE) * vp.W;
    fwd[locId] = cast_ray(rd, x, y, __ldg(minmax + locId2), sInvM, vp, sp);
  }
}

// processPixelForwardRender<true> over the forward projection; also the bookkeeping of ITMTrackingController::Prepare's
// else-branch (age_pointCloud++)
__global__ void __launch_bounds__(256) k_fwd_shade(const float4 *__restrict__ fwd, uchar4 *__restrict__ outRendering, FrameState *st,
                                                   ViewParams vp, float voxelSize, int gated) {
  if (gated && st->requiresFullRendering) return;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (gated && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) st->agePointCloud = st->agePointCloud + 1;
  if (x >= vp.W || y >= vp.H) return;
  const int locId = x + y * vp.W;
  const float lx = -st->invM_d[8], ly = -st->invM_d[9], lz = -st->invM_d[10];
  const float4 point = __ldg(fwd + locId);
  bool foundPoint = point.w > 0.0f;
  float nx, ny, nz, angle = 0.0f;
  normal_angle_from_points(foundPoint, x, y, fwd, lx, ly, lz, voxelSize, vp.W, vp.H, nx, ny, nz, angle);
  if (foundPoint) {
    const float outRes = (0.8f * angle + 0.2f) * 255.0f;
    const unsigned char g = (unsigned char)outRes;
    outRendering[locId] = make_uchar4(g, g, g, g);
  } else {
    outRendering[locId] = make_uchar4(0, 0, 0, 0);
  }
}

// ---------------------------------------------------------------- FindVisibleBlocks

#define VIS_TILE 8192
#define VIS_PER_THREAD 32

__global__ void __launch_bounds__(256) k_find_visible(const HashEntry *__restrict__ table, int *__restrict__ visibleIds, FrameState *st,
                                                      ViewParams vp, SceneParams sp, int visibleCapacity, unsigned long long *ticket,
                                                      unsigned long long *tileState, int numTiles, int allAllocated) {
  __shared__ unsigned sWarp[8];
  __shared__ unsigned sTotal;
  __shared__ unsigned sExA;
  __shared__ int sTile;
  __shared__ unsigned sEpoch;
  __shared__ float sM[16];
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(ticket, 1ull);
    sTile = (int)(t % (unsigned long long)numTiles);
    sEpoch = (unsigned)((t / (unsigned long long)numTiles + 1ull) & 0xFFFFFull);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 48) sM[threadIdx.x - 32] = st->M_d[threadIdx.x - 32];
  __syncthreads();
  const int tile = sTile;
  // lanes read consecutive entries (coalesced 512 B per warp and step); lane l of warp w owns slots base + step*32 + l,
  // so its rank within the tile is NOT contiguous - ranks are resolved per step with ballots below
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warpSlot0 = tile * VIS_TILE + warp * (32 * VIS_PER_THREAD);
  __shared__ unsigned sBallot[8][VIS_PER_THREAD];  // visible lanes of warp w at step i
  unsigned cnt = 0;                                // visible entries of the whole warp (same in every lane)
#pragma unroll 1
  for (int i = 0; i < VIS_PER_THREAD; ++i) {
    const int slot = warpSlot0 + i * 32 + lane;
    bool vis = false;
    if (slot < sp.nEntries) {
      const HashEntry e = load_entry(table, slot);
      if (e.ptr >= 0) vis = allAllocated ? true : block_visible(sM, e.px, e.py, e.pz, sp.voxelSize, vp);
    }
    const unsigned b = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) sBallot[warp][i] = b;
    cnt += __popc(b);
  }
  // exclusive scan over the 8 warps of the CTA (every lane of a warp carries the warp's count: use lane 0's)
  if (lane == 0) sWarp[warp] = cnt;
  __syncthreads();
  unsigned warpBase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const unsigned s = sWarp[w];
    if (w < warp) warpBase += s;
    tot += s;
  }
  if (threadIdx.x == 0) sTotal = tot;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned exA, exB;
    scan_lookback(tileState, tile, sEpoch, sTotal, 0u, exA, exB);
    if (threadIdx.x == 0) {
      sExA = exA;
      if (tile == numTiles - 1) {
        int total = (int)(exA + sTotal);
        if (total > visibleCapacity) {
          atomicOr(&st->errorFlags, 2);
          total = visibleCapacity;
        }
        st->noVisibleEntries = total;
      }
    }
  }
  __syncthreads();
  if (cnt == 0) return;
  int pos = (int)(sExA + warpBase);
  for (int i = 0; i < VIS_PER_THREAD; ++i) {
    const unsigned b = sBallot[warp][i];
    if ((b >> lane) & 1u) {
      const int p = pos + __popc(b & ((1u << lane) - 1u));
      if (p < visibleCapacity) visibleIds[p] = warpSlot0 + i * 32 + lane;
    }
    pos += __popc(b);
  }
}

// ---------------------------------------------------------------- RenderImage

// computeSingleNormalFromSDF (ITMRepresentationAccess.h:225-337); the 32 taps are raw shorts (readVoxel(...).sdf)
template <int VW>
__device__ __forceinline__ void single_normal_from_sdf(VoxelReader<VW> &rd, float px, float py, float pz, float &rx, float &ry, float &rz) {
  const float flx = floorf(px), fly = floorf(py), flz = floorf(pz);
  const float cx = px - flx, cy = py - fly, cz = pz - flz;
  const int x = (int)flx, y = (int)fly, z = (int)flz;
  const float ncx = 1.0f - cx, ncy = 1.0f - cy, ncz = 1.0f - cz;
  bool f;
#define SDF_AT(dx, dy, dz) ((float)rd.read_sdf(x + (dx), y + (dy), z + (dz), f))
  const float f_x = SDF_AT(0, 0, 0), f_y = SDF_AT(1, 0, 0), f_z = SDF_AT(0, 1, 0), f_w = SDF_AT(1, 1, 0);
  const float b_x = SDF_AT(0, 0, 1), b_y = SDF_AT(1, 0, 1), b_z = SDF_AT(0, 1, 1), b_w = SDF_AT(1, 1, 1);
  float t_x, t_y, t_z, t_w, p1, p2, v1;
  // gradient x
  p1 = f_x * ncy * ncz + f_z * cy * ncz + b_x * ncy * cz + b_z * cy * cz;
  t_x = SDF_AT(-1, 0, 0); t_y = SDF_AT(-1, 1, 0); t_z = SDF_AT(-1, 0, 1); t_w = SDF_AT(-1, 1, 1);
  p2 = t_x * ncy * ncz + t_y * cy * ncz + t_z * ncy * cz + t_w * cy * cz;
  v1 = p1 * cx + p2 * ncx;
  p1 = f_y * ncy * ncz + f_w * cy * ncz + b_y * ncy * cz + b_w * cy * cz;
  t_x = SDF_AT(2, 0, 0); t_y = SDF_AT(2, 1, 0); t_z = SDF_AT(2, 0, 1); t_w = SDF_AT(2, 1, 1);
  p2 = t_x * ncy * ncz + t_y * cy * ncz + t_z * ncy * cz + t_w * cy * cz;
  rx = div32767(p1 * ncx + p2 * cx - v1, rd.y32767);
  // gradient y
  p1 = f_x * ncx * ncz + f_y * cx * ncz + b_x * ncx * cz + b_y * cx * cz;
  t_x = SDF_AT(0, -1, 0); t_y = SDF_AT(1, -1, 0); t_z = SDF_AT(0, -1, 1); t_w = SDF_AT(1, -1, 1);
  p2 = t_x * ncx * ncz + t_y * cx * ncz + t_z * ncx * cz + t_w * cx * cz;
  v1 = p1 * cy + p2 * ncy;
  p1 = f_z * ncx * ncz + f_w * cx * ncz + b_z * ncx * cz + b_w * cx * cz;
  t_x = SDF_AT(0, 2, 0); t_y = SDF_AT(1, 2, 0); t_z = SDF_AT(0, 2, 1); t_w = SDF_AT(1, 2, 1);
  p2 = t_x * ncx * ncz + t_y * cx * ncz + t_z * ncx * cz + t_w * cx * cz;
  ry = div32767(p1 * ncy + p2 * cy - v1, rd.y32767);
  // gradient z
  p1 = f_x * ncx * ncy + f_y * cx * ncy + f_z * ncx * cy + f_w * cx * cy;
  t_x = SDF_AT(0, 0, -1); t_y = SDF_AT(1, 0, -1); t_z = SDF_AT(0, 1, -1); t_w = SDF_AT(1, 1, -1);
  p2 = t_x * ncx * ncy + t_y * cx * ncy + t_z * ncx * cy + t_w * cx * cy;
  v1 = p1 * cz + p2 * ncz;
  p1 = b_x * ncx * ncy + b_y * cx * ncy + b_z * ncx * cy + b_w * cx * cy;
  t_x = SDF_AT(0, 0, 2); t_y = SDF_AT(1, 0, 2); t_z = SDF_AT(0, 1, 2); t_w = SDF_AT(1, 1, 2);
  p2 = t_x * ncx * ncy + t_y * cx * ncy + t_z * ncx * cy + t_w * cx * cy;
  rz = div32767(p1 * ncz + p2 * cz - v1, rd.y32767);
#undef SDF_AT
}

// readFromSDF_color4u_interpolated (ITMRepresentationAccess.h:186-222) for ITMVoxel_s_rgb: word0 = sdf | w_depth<<16 | r<<24,
// word1 = g | b<<8 | w_color<<16.  Missing voxels read as ITMVoxel_s_rgb() (clr = 0).
__device__ __forceinline__ void colour_interpolated(VoxelReader<2> &rd, float px, float py, float pz, float &r, float &g, float &b) {
  const float flx = floorf(px), fly = floorf(py), flz = floorf(pz);
  const float cx = px - flx, cy = py - fly, cz = pz - flz;
  const int x = (int)flx, y = (int)fly, z = (int)flz;
  r = g = b = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const float wgt = (dx ? cx : (1.0f - cx)) * (dy ? cy : (1.0f - cy)) * (dz ? cz : (1.0f - cz));
    const int vx = x + dx, vy = y + dy, vz = z + dz;
    float cr = 0.0f, cg = 0.0f, cb = 0.0f;
    if (rd.find_block(vx >> 3, vy >> 3, vz >> 3)) {
      const uint32_t *p = rd.voxels + (size_t)(rd.cptr + (vx & 7) + ((vy & 7) << 3) + ((vz & 7) << 6)) * 2;
      const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1);
      cr = (float)(w0 >> 24); cg = (float)(w1 & 0xFFu); cb = (float)((w1 >> 8) & 0xFFu);
    }
    r += wgt * cr; g += wgt * cg; b += wgt * cb;
  }
  r = r / 255.0f; g = g / 255.0f; b = b / 255.0f;
}

// GenericRaycast + processPixel{Grey,Colour,Normal} fused.  type: 0 grey, 1 colour from volume, 2 colour from normal
// (IITMVisualisationEngine::RenderImageType, Engine/ITMVisualisationEngine.h:27-31)
template <int VW>
__global__ void __launch_bounds__(128) k_render_image(const void *__restrict__ voxels, const void *__restrict__ table,
                                                      const float2 *__restrict__ minmax, float4 *__restrict__ raycastResult,
                                                      uchar4 *__restrict__ outImage, const FrameState *__restrict__ st, ViewParams vp,
                                                      SceneParams sp, int type) {
  __shared__ float sInvM[16];
  if (threadIdx.x < 16) sInvM[threadIdx.x] = st->invM_d[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
  if (x >= vp.W || y >= vp.H) return;
  const int locId = x + y * vp.W;
  const int locId2 = (int)floorf((float)x / (float)ITM_MINMAX_SUBSAMPLE) + (int)floorf((float)y / (float)ITM_MINMAX_SUBSAMPLE) * vp.W;
  VoxelReader<VW> rd;
  rd.init(voxels, table, sp.nBuckets, sp.hashMask);
  const float4 pt = cast_ray(rd, x, y, __ldg(minmax + locId2), sInvM, vp, sp);
  raycastResult[locId] = pt;
  bool foundPoint = pt.w > 0.0f;
  uchar4 out = make_uchar4(0, 0, 0, 0);
  if (foundPoint) {
    // computeNormalAndAngle<TVoxel, TIndex>
    float nx, ny, nz;
    single_normal_from_sdf(rd, pt.x, pt.y, pt.z, nx, ny, nz);
    const float normScale = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    nx *= normScale; ny *= normScale; nz *= normScale;
    const float lx = -sInvM[8], ly = -sInvM[9], lz = -sInvM[10];
    const float angle = nx * lx + ny * ly + nz * lz;
    if (angle > 0.0f) {
      if (type == 1 && VW == 2) {
        float r, g, b;
        if constexpr (VW == 2) colour_interpolated(rd, pt.x, pt.y, pt.z, r, g, b);
        else r = g = b = 0.0f;
        out = make_uchar4((unsigned char)(r * 255.0f), (unsigned char)(g * 255.0f), (unsigned char)(b * 255.0f), 255);
      } else if (type == 2) {
        // drawPixelNormal writes r, g, b only: the pixel's w keeps whatever the image held before
        out = make_uchar4((unsigned char)((0.3f + (-nx + 1.0f) * 0.35f) * 255.0f), (unsigned char)((0.3f + (-ny + 1.0f) * 0.35f) * 255.0f),
                          (unsigned char)((0.3f + (-nz + 1.0f) * 0.35f) * 255.0f), outImage[locId].w);
      } else {
        const unsigned char g = (unsigned char)((0.8f * angle + 0.2f) * 255.0f);
        out = make_uchar4(g, g, g, g);
      }
    }
  }
  outImage[locId] = out;
}


// ---------------------------------------------------------------- CreatePointCloud

// RenderPointCloud (ITMVisualisationEngine_CPU.cpp:424-462) once k_render_image (grey) has shaded `image` from the same
// rays: a pixel carries a point iff its grey value is non-zero (drawPixelGrey writes >= 51 wherever
// computeNormalAndAngle kept the point, everything else is cleared), and with skipPoints only odd columns of odd rows
// count.  The serial loop numbers the points in raster order; here 8192-pixel tiles do that with one look-back scan, then
// every kept pixel writes location = point * voxelSize (w = 1) and colour = the interpolated voxel colour (w = 1;
// all zero for voxels without colour) at its rank.
template <int VW>
__global__ void __launch_bounds__(256) k_point_cloud(const void *__restrict__ voxels, const void *__restrict__ table,
                                                     const float4 *__restrict__ pointsRay, const uchar4 *__restrict__ image,
                                                     float4 *__restrict__ locations, float4 *__restrict__ colours, FrameState *st,
                                                     ViewParams vp, SceneParams sp, int skipPoints, unsigned long long *ticket,
                                                     unsigned long long *tileState, int numTiles) {
  __shared__ unsigned sWarp[8];
  __shared__ unsigned sTotal;
  __shared__ unsigned sExA;
  __shared__ int sTile;
  __shared__ unsigned sEpoch;
  __shared__ unsigned sBallot[8][VIS_PER_THREAD];
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(ticket, 1ull);
    sTile = (int)(t % (unsigned long long)numTiles);
    sEpoch = (unsigned)((t / (unsigned long long)numTiles + 1ull) & 0xFFFFFull);
  }
  __syncthreads();
  const int tile = sTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nPixels = vp.W * vp.H;
  const int warpPix0 = tile * VIS_TILE + warp * (32 * VIS_PER_THREAD);
  unsigned cnt = 0;
#pragma unroll 1
  for (int i = 0; i < VIS_PER_THREAD; ++i) {
    const int pix = warpPix0 + i * 32 + lane;
    bool keep = false;
    if (pix < nPixels) {
      keep = __ldg(&image[pix]).w != 0;
      if (skipPoints) {
        const int y = pix / vp.W, x = pix - y * vp.W;
        if ((x % 2 == 0) || (y % 2 == 0)) keep = false;
      }
    }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sBallot[warp][i] = b;
    cnt += __popc(b);
  }
  if (lane == 0) sWarp[warp] = cnt;
  __syncthreads();
  unsigned warpBase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const unsigned s = sWarp[w];
    if (w < warp) warpBase += s;
    tot += s;
  }
  if (threadIdx.x == 0) sTotal = tot;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned exA, exB;
    scan_lookback(tileState, tile, sEpoch, sTotal, 0u, exA, exB);
    if (threadIdx.x == 0) {
      sExA = exA;
      if (tile == numTiles - 1) st->noTotalPoints = (int)(exA + sTotal);
    }
  }
  __syncthreads();
  if (cnt == 0) return;
  VoxelReader<VW> rd;
  rd.init(voxels, table, sp.nBuckets, sp.hashMask);
  int pos = (int)(sExA + warpBase);
  for (int i = 0; i < VIS_PER_THREAD; ++i) {
    const unsigned b = sBallot[warp][i];
    if ((b >> lane) & 1u) {
      const int p = pos + __popc(b & ((1u << lane) - 1u));
      const float4 pt = __ldg(&pointsRay[warpPix0 + i * 32 + lane]);
      float4 clr = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if constexpr (VW == 2) {
        // readFromSDF_color4u_interpolated returns (r, g, b, 255) / 255, and "tmp /= tmp.w" divides by 1
        colour_interpolated(rd, pt.x, pt.y, pt.z, clr.x, clr.y, clr.z);
        clr.w = 1.0f;
      }
      colours[p] = clr;
      locations[p] = make_float4(pt.x * sp.voxelSize, pt.y * sp.voxelSize, pt.z * sp.voxelSize, 1.0f);
    }
    pos += __popc(b);
  }
}

}  // namespace

namespace itm {

void launch_track_decide(FrameState *st, int useApproximateRaycast, cudaStream_t s) { k_track_decide<<<1, 32, 0, s>>>(st, useApproximateRaycast); }

void launch_forward_render(const ForwardArgs &a, cudaStream_t s) {
  const RenderArgs &r = a.render;
  dim3 g((r.vp.W + 31) / 32, (r.vp.H + 7) / 8);
  const float4 *rays = reinterpret_cast<const float4 *>(r.raycastResult);
  float4 *fwd = reinterpret_cast<float4 *>(a.forwardProjection);
  const float2 *minmax = reinterpret_cast<const float2 *>(r.minmax);
  k_fwd_project<<<g, 256, 0, s>>>(rays, a.key, r.st, r.vp, r.sp.voxelSize, a.gated);
  k_fwd_gather<<<g, 256, 0, s>>>(rays, a.key, fwd, a.depth, minmax, a.missingPoints, r.st, r.vp, a.gated);
  if (r.sp.voxelWords == 2) k_fwd_cast<2><<<148 * 4, 128, 0, s>>>(r.voxels, r.hashTable, minmax, a.missingPoints, fwd, r.st, r.vp, r.sp, a.gated);
  else k_fwd_cast<1><<<148 * 4, 128, 0, s>>>(r.voxels, r.hashTable, minmax, a.missingPoints, fwd, r.st, r.vp, r.sp, a.gated);
  k_fwd_shade<<<g, 256, 0, s>>>(fwd, reinterpret_cast<uchar4 *>(r.raycastImage), r.st, r.vp, r.sp.voxelSize, a.gated);
}

void launch_find_visible_blocks(const void *hashTable, int *visibleIds, FrameState *st, const ViewParams &vp, const SceneParams &sp,
                                int visibleCapacity, unsigned long long *ticket, unsigned long long *tileState, cudaStream_t s,
                                int allAllocated) {
  const int numTiles = (sp.nEntries + VIS_TILE - 1) / VIS_TILE;
  k_find_visible<<<numTiles, 256, 0, s>>>(reinterpret_cast<const HashEntry *>(hashTable), visibleIds, st, vp, sp, visibleCapacity, ticket,
                                          tileState, numTiles, allAllocated);
}

void launch_render_image(const RenderArgs &a, unsigned char *outImage, int type, cudaStream_t s) {
  dim3 g((a.vp.W + 15) / 16, (a.vp.H + 7) / 8);
  if (a.sp.voxelWords == 2)
    k_render_image<2><<<g, 128, 0, s>>>(a.voxels, a.hashTable, reinterpret_cast<const float2 *>(a.minmax), reinterpret_cast<float4 *>(a.raycastResult),
                                        reinterpret_cast<uchar4 *>(outImage), a.st, a.vp, a.sp, type);
  else
    k_render_image<1><<<g, 128, 0, s>>>(a.voxels, a.hashTable, reinterpret_cast<const float2 *>(a.minmax), reinterpret_cast<float4 *>(a.raycastResult),
                                        reinterpret_cast<uchar4 *>(outImage), a.st, a.vp, a.sp, type);
}

int point_cloud_tiles(int W, int H) { return (W * H + VIS_TILE - 1) / VIS_TILE; }

void launch_point_cloud(const RenderArgs &a, int skipPoints, float *locations, float *colours, unsigned long long *tileState,
                        int numTiles, cudaStream_t s) {
  // tileState[numTiles] is the ticket counter of this tile count
  const float4 *rays = reinterpret_cast<const float4 *>(a.raycastResult);
  const uchar4 *img = reinterpret_cast<const uchar4 *>(a.raycastImage);
  if (a.sp.voxelWords == 2)
    k_point_cloud<2><<<numTiles, 256, 0, s>>>(a.voxels, a.hashTable, rays, img, reinterpret_cast<float4 *>(locations), reinterpret_cast<float4 *>(colours),
                                              a.st, a.vp, a.sp, skipPoints, tileState + numTiles, tileState, numTiles);
  else
    k_point_cloud<1><<<numTiles, 256, 0, s>>>(a.voxels, a.hashTable, rays, img, reinterpret_cast<float4 *>(locations), reinterpret_cast<float4 *>(colours),
                                              a.st, a.vp, a.sp, skipPoints, tileState + numTiles, tileState, numTiles);
}

}  // namespace itm
