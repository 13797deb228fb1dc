// Frustum tests of voxel blocks, shared by the allocation kernels (k_alloc.cu) and FindVisibleBlocks (k_vis.cu).
#pragma once
#include "itm_common.cuh"

namespace itm {

// checkPointVisibility<false>, ITMSceneReconstructionEngine.h:244-274
__device__ __forceinline__ bool point_visible(const float *M, float x, float y, float z, const ViewParams &vp) {
  float bx, by, bz;
  mat4_mul_vec4(M, x, y, z, 1.0f, bx, by, bz);
  if (bz < 1e-10f) return false;
  bx = vp.fx * bx / bz + vp.cx;
  by = vp.fy * by / bz + vp.cy;
  return bx >= 0 && bx < (float)vp.W && by >= 0 && by < (float)vp.H;
}

// checkPointVisibility<true>'s second answer (:262-273): inside the image enlarged by 1/8 on every side
__device__ __forceinline__ bool point_visible_enlarged(const float *M, float x, float y, float z, const ViewParams &vp) {
  float bx, by, bz;
  mat4_mul_vec4(M, x, y, z, 1.0f, bx, by, bz);
  if (bz < 1e-10f) return false;
  bx = vp.fx * bx / bz + vp.cx;
  by = vp.fy * by / bz + vp.cy;
  const int lx = -vp.W / 8, ux = vp.W + vp.W / 8, ly = -vp.H / 8, uy = vp.H + vp.H / 8;
  return bx >= (float)lx && bx < (float)ux && by >= (float)ly && by < (float)uy;
}

// checkBlockVisibility<true>'s isVisibleEnlarged: some corner lies in the enlarged image (a corner inside the image
// proper ends the reference's walk early, but it is inside the enlarged one too)
static __device__ __noinline__ bool block_visible_enlarged(const float *M, int hx, int hy, int hz, float voxelSize, const ViewParams &vp) {
  const float factor = (float)ITM_BLOCK_SIZE * voxelSize;
  float x = (float)hx * factor, y = (float)hy * factor, z = (float)hz * factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 0 0 0
  z += factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 0 0 1
  y += factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 0 1 1
  x += factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 1 1 1
  z -= factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 1 1 0
  y -= factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 1 0 0
  x -= factor;
  y += factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 0 1 0
  x += factor;
  y -= factor;
  z += factor;
  if (point_visible_enlarged(M, x, y, z, vp)) return true;  // 1 0 1
  return false;
}

// checkBlockVisibility<false>, :277-342 - the corner coordinates are built by the same chain of
// += / -= as the reference so that they round identically
static __device__ __noinline__ bool block_visible(const float *M, int hx, int hy, int hz, float voxelSize, const ViewParams &vp) {
  const float factor = (float)ITM_BLOCK_SIZE * voxelSize;
  float x = (float)hx * factor, y = (float)hy * factor, z = (float)hz * factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 0 0 0
  z += factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 0 0 1
  y += factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 0 1 1
  x += factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 1 1 1
  z -= factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 1 1 0
  y -= factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 1 0 0
  x -= factor;
  y += factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 0 1 0
  x += factor;
  y -= factor;
  z += factor;
  if (point_visible(M, x, y, z, vp)) return true;  // 1 0 1
  return false;
}

}  // namespace itm
