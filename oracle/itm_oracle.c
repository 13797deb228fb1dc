/* oracle/itm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the reference's per-frame dense-fusion path
 * (view -> ICP tracking -> allocation -> integration -> expected depths -> raycast -> ICP maps) and of
 * the rows next to it (ForwardRender / useApproximateRaycast, free-view FindVisibleBlocks + RenderImage, MeshScene,
 * the weighted ICP tracker with its depth filter and sensor-noise model),
 * written from the reference's algorithm with run-time pool sizes so that configurations the
 * reference can only reach by editing #defines (BASELINE configs[2]: 2 mm voxels, larger pools)
 * have an oracle too.  Every function cites the reference lines it restates.
 *
 * Parity status: PINNED.  tests/test_oracle_port.py checks this file stage by stage and bit for
 * bit against the real reference CPU engines (oracle/_ref/libitm_ref.so, built from the unmodified
 * sources by oracle/build_ref.py) when that library is present, and against the golden vectors in
 * tests/golden/ (generated from the real reference by tests/golden/make_golden.py and make_golden_8f.py) everywhere.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/build_port.py).  No FMA contraction and no
 * -ffast-math: every float operation is a single IEEE fp32 operation in source order, which is what
 * the parity flavour of the reference build does too.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <time.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK 8
#define BLOCK3 512
#define FAR_AWAY 999999.9f
#define VERY_CLOSE 0.05f
#define MINMAX_SUB 8
#define MAX_RENDERING_BLOCKS (65536 * 4)
#define MAX_LEVELS 8
/* the reference calls exp() / acos() unqualified on float arguments from C++ (DeviceAgnostic/ITMViewBuilder.h:48, 110) */
#ifndef WICP_EXP
#define WICP_EXP(x) expf(x)
#define WICP_ACOS(x) acosf(x)
#endif

enum { ITER_ROTATION = 1, ITER_TRANSLATION = 2, ITER_BOTH = 3, ITER_NONE = 4 }; /* ITMLibDefines.h:278-283 */

typedef struct { short x, y, z, pad; int offset; int ptr; } HashEntry; /* ITMLibDefines.h:71-82 */
typedef struct { short sdf; unsigned char w_depth; unsigned char pad; } Voxel; /* ITMVoxel_s, ITMLibDefines.h:157-179 */
typedef struct { float x, y, z, w; } V4;
typedef struct { float x, y; } V2;

typedef struct {
  int width, height;
  float fx, fy, cx, cy;
  float voxel_size, mu;
  int max_w;
  float vf_min, vf_max;
  int stop_at_max_w;
  float calib_a, calib_b;
  int n_local, n_bucket, n_excess;
  int n_levels;
  int regime[MAX_LEVELS];
  int no_icp_run_till_level;
  float icp_dist_thresh, icp_termination;
} port_params;

static int g_next_wicp = 0, g_next_bilateral = 0;  /* port_set_tracker_wicp */

typedef struct port_engine {
  port_params p;
  int n_entries;
  /* scene */
  Voxel *voxels; HashEntry *hash; int *vba_list; int *excess_list;
  int last_free_block, last_free_excess;
  /* render state */
  int *visible_ids; unsigned char *visible_type; int n_visible;
  V2 *minmax; V4 *raycast; unsigned char *raycast_image;
  /* tracking state */
  V4 *points, *normals;
  float pose_M[16], pose_params[6];
  float pose_pc_M[16];
  int age;
  /* view */
  short *raw; float *depth;
  float *level_depth[MAX_LEVELS]; int level_w[MAX_LEVELS], level_h[MAX_LEVELS]; float level_intr[MAX_LEVELS][4];
  int iters[MAX_LEVELS]; float dist_thresh[MAX_LEVELS];
  /* engine scratch */
  unsigned char *alloc_type; short *block_coords; /* Vector4s per slot */
  /* ITMRenderState::forwardProjection / fwdProjMissingPoints, ITMMesh, renderState_freeview */
  V4 *fwd; int *fwd_missing; int n_fwd_missing; int requires_full_rendering; int use_approximate_raycast;
  float *mesh; int n_mesh;
  /* TRACKER_WICP: settings.modelSensorNoise / useBilateralFilter, view->depthUncertainty / depthNormal, weight hierarchy */
  int wicp, bilateral; float *float_tmp; float *sigma; V4 *dnormal; float *level_sigma[MAX_LEVELS];
  int free_w, free_h; int *free_visible_ids; int free_n_visible; V2 *free_minmax; V4 *free_raycast; unsigned char *free_image;
  float trafo_rgb_to_depth[16]; /* ITMRGBDCalib::trafo_rgb_to_depth.calib (identity unless port_create_point_cloud installs one) */
} port_engine;

/* ------------------------------------------------------------------------------------------------
 * small matrix / pose arithmetic
 * ---------------------------------------------------------------------------------------------- */

/* Matrix4::inv by cofactors, ORUtils/Matrix.h:162-218 */
static int m4_inverse(const float *m, float *dst) {
  float t[12], s[16], det;
  int i;
  for (i = 0; i < 4; i++) { s[i] = m[i * 4]; s[i + 4] = m[i * 4 + 1]; s[i + 8] = m[i * 4 + 2]; s[i + 12] = m[i * 4 + 3]; }
  t[0] = s[10] * s[15]; t[1] = s[11] * s[14]; t[2] = s[9] * s[15]; t[3] = s[11] * s[13];
  t[4] = s[9] * s[14]; t[5] = s[10] * s[13]; t[6] = s[8] * s[15]; t[7] = s[11] * s[12];
  t[8] = s[8] * s[14]; t[9] = s[10] * s[12]; t[10] = s[8] * s[13]; t[11] = s[9] * s[12];
  dst[0] = (t[0] * s[5] + t[3] * s[6] + t[4] * s[7]) - (t[1] * s[5] + t[2] * s[6] + t[5] * s[7]);
  dst[1] = (t[1] * s[4] + t[6] * s[6] + t[9] * s[7]) - (t[0] * s[4] + t[7] * s[6] + t[8] * s[7]);
  dst[2] = (t[2] * s[4] + t[7] * s[5] + t[10] * s[7]) - (t[3] * s[4] + t[6] * s[5] + t[11] * s[7]);
  dst[3] = (t[5] * s[4] + t[8] * s[5] + t[11] * s[6]) - (t[4] * s[4] + t[9] * s[5] + t[10] * s[6]);
  det = s[0] * dst[0] + s[1] * dst[1] + s[2] * dst[2] + s[3] * dst[3];
  if (det == 0.0f) return 0;
  dst[4] = (t[1] * s[1] + t[2] * s[2] + t[5] * s[3]) - (t[0] * s[1] + t[3] * s[2] + t[4] * s[3]);
  dst[5] = (t[0] * s[0] + t[7] * s[2] + t[8] * s[3]) - (t[1] * s[0] + t[6] * s[2] + t[9] * s[3]);
  dst[6] = (t[3] * s[0] + t[6] * s[1] + t[11] * s[3]) - (t[2] * s[0] + t[7] * s[1] + t[10] * s[3]);
  dst[7] = (t[4] * s[0] + t[9] * s[1] + t[10] * s[2]) - (t[5] * s[0] + t[8] * s[1] + t[11] * s[2]);
  t[0] = s[2] * s[7]; t[1] = s[3] * s[6]; t[2] = s[1] * s[7]; t[3] = s[3] * s[5];
  t[4] = s[1] * s[6]; t[5] = s[2] * s[5]; t[6] = s[0] * s[7]; t[7] = s[3] * s[4];
  t[8] = s[0] * s[6]; t[9] = s[2] * s[4]; t[10] = s[0] * s[5]; t[11] = s[1] * s[4];
  dst[8] = (t[0] * s[13] + t[3] * s[14] + t[4] * s[15]) - (t[1] * s[13] + t[2] * s[14] + t[5] * s[15]);
  dst[9] = (t[1] * s[12] + t[6] * s[14] + t[9] * s[15]) - (t[0] * s[12] + t[7] * s[14] + t[8] * s[15]);
  dst[10] = (t[2] * s[12] + t[7] * s[13] + t[10] * s[15]) - (t[3] * s[12] + t[6] * s[13] + t[11] * s[15]);
  dst[11] = (t[5] * s[12] + t[8] * s[13] + t[11] * s[14]) - (t[4] * s[12] + t[9] * s[13] + t[10] * s[14]);
  dst[12] = (t[2] * s[10] + t[5] * s[11] + t[1] * s[9]) - (t[4] * s[11] + t[0] * s[9] + t[3] * s[10]);
  dst[13] = (t[8] * s[11] + t[0] * s[8] + t[7] * s[10]) - (t[6] * s[10] + t[9] * s[11] + t[1] * s[8]);
  dst[14] = (t[6] * s[9] + t[11] * s[11] + t[3] * s[8]) - (t[10] * s[11] + t[2] * s[8] + t[7] * s[9]);
  dst[15] = (t[10] * s[10] + t[4] * s[8] + t[9] * s[9]) - (t[8] * s[9] + t[11] * s[10] + t[5] * s[8]);
  { const float k = 1 / det; for (i = 0; i < 16; ++i) dst[i] *= k; }
  return 1;
}

/* Matrix4 * Matrix4, ORUtils/Matrix.h:102-107 (column-major, accumulate from zero over k) */
static void m4_product(const float *a, const float *b, float *out) {
  float r[16];
  int col, row, k;
  for (col = 0; col < 4; col++) for (row = 0; row < 4; row++) {
    float acc = 0.0f;
    for (k = 0; k < 4; k++) acc += a[row + 4 * k] * b[k + 4 * col];
    r[row + 4 * col] = acc;
  }
  memcpy(out, r, sizeof(r));
}

/* Matrix4 * Vector4, ORUtils/Matrix.h:112-119; only xyz are needed by the callers */
static void m4_apply(const float *m, float x, float y, float z, float w, float *ox, float *oy, float *oz) {
  *ox = m[0] * x + m[4] * y + m[8] * z + m[12] * w;
  *oy = m[1] * x + m[5] * y + m[9] * z + m[13] * w;
  *oz = m[2] * x + m[6] * y + m[10] * z + m[14] * w;
}

static float dot3(const float *a, const float *b) { float r = 0; r += a[0] * b[0]; r += a[1] * b[1]; r += a[2] * b[2]; return r; }
static void cross3(const float *a, const float *b, float *r) {
  r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}

/* SE(3) exponential, ITMPose::SetModelViewFromParams, ITMLib/Objects/ITMPose.cpp:84-152.  p = tx ty tz rx ry rz */
static void se3_exp(const float *p, float *M) {
  const float sixth = 1.0f / 6.0f, twentieth = 1.0f / 20.0f;
  float w[3], t[3], wxt[3], T[3], A, B, R[9];
  float th2, th, a, b, wx2, wy2, wz2;
  int r, c;
  w[0] = p[3]; w[1] = p[4]; w[2] = p[5]; t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
  th2 = dot3(w, w); th = sqrtf(th2);
  cross3(w, t, wxt);
  if (th2 < 1e-8f) {
    A = 1.0f - sixth * th2; B = 0.5f;
    T[0] = t[0] + 0.5f * wxt[0]; T[1] = t[1] + 0.5f * wxt[1]; T[2] = t[2] + 0.5f * wxt[2];
  } else {
    float C, wwxt[3];
    if (th2 < 1e-6f) {
      C = sixth * (1.0f - twentieth * th2); A = 1.0f - th2 * C; B = 0.5f - 0.25f * sixth * th2;
    } else {
      const float inv = 1.0f / th;
      A = sinf(th) * inv; B = (1.0f - cosf(th)) * (inv * inv); C = (1.0f - A) * (inv * inv);
    }
    cross3(w, wxt, wwxt);
    T[0] = t[0] + B * wxt[0] + C * wwxt[0]; T[1] = t[1] + B * wxt[1] + C * wwxt[1]; T[2] = t[2] + B * wxt[2] + C * wwxt[2];
  }
  wx2 = w[0] * w[0]; wy2 = w[1] * w[1]; wz2 = w[2] * w[2];
  R[0] = 1.0f - B * (wy2 + wz2); R[4] = 1.0f - B * (wx2 + wz2); R[8] = 1.0f - B * (wx2 + wy2);
  a = A * w[2]; b = B * (w[0] * w[1]); R[0 + 3 * 1] = b - a; R[1 + 3 * 0] = b + a;
  a = A * w[1]; b = B * (w[0] * w[2]); R[0 + 3 * 2] = b + a; R[2 + 3 * 0] = b - a;
  a = A * w[0]; b = B * (w[1] * w[2]); R[1 + 3 * 2] = b - a; R[2 + 3 * 1] = b + a;
  for (c = 0; c < 3; ++c) for (r = 0; r < 3; ++r) M[r + 4 * c] = R[r + 3 * c];
  M[12] = T[0]; M[13] = T[1]; M[14] = T[2];
  M[3] = 0.0f; M[7] = 0.0f; M[11] = 0.0f; M[15] = 1.0f;
}

/* SE(3) logarithm, ITMPose::SetParamsFromModelView, ITMLib/Objects/ITMPose.cpp:154-234 */
static void se3_log(const float *M, float *p) {
  float R[9], T[3], rot[3], half[6], Mh[16], rt[3];
  float cosang, sinabs, shtot, theta;
  const double HALF_SQRT2 = 0.707106781186547524401;
  int r, c;
  for (c = 0; c < 3; ++c) for (r = 0; r < 3; ++r) R[r + 3 * c] = M[r + 4 * c];
  T[0] = M[12]; T[1] = M[13]; T[2] = M[14];
  cosang = (R[0] + R[4] + R[8] - 1.0f) * 0.5f;
  rot[0] = (R[2 + 3 * 1] - R[1 + 3 * 2]) * 0.5f;
  rot[1] = (R[0 + 3 * 2] - R[2 + 3 * 0]) * 0.5f;
  rot[2] = (R[1 + 3 * 0] - R[0 + 3 * 1]) * 0.5f;
  sinabs = sqrtf(dot3(rot, rot));
  if (cosang > HALF_SQRT2) {
    if (sinabs) { const float k = asinf(sinabs) / sinabs; rot[0] *= k; rot[1] *= k; rot[2] *= k; }
  } else if (cosang > -HALF_SQRT2) {
    const float k = acosf(cosang) / sinabs; rot[0] *= k; rot[1] *= k; rot[2] *= k;
  } else {
    const float angle = (float)3.14159265358979323846 - asinf(sinabs);
    const float d0 = R[0] - cosang, d1 = R[4] - cosang, d2 = R[8] - cosang;
    float r2[3], len;
    if (fabsf(d0) > fabsf(d1) && fabsf(d0) > fabsf(d2)) {
      r2[0] = d0; r2[1] = (R[1] + R[3]) * 0.5f; r2[2] = (R[6] + R[2]) * 0.5f;
    } else if (fabsf(d1) > fabsf(d2)) {
      r2[0] = (R[1] + R[3]) * 0.5f; r2[1] = d1; r2[2] = (R[5] + R[7]) * 0.5f;
    } else {
      r2[0] = (R[6] + R[2]) * 0.5f; r2[1] = (R[5] + R[7]) * 0.5f; r2[2] = d2;
    }
    if (dot3(r2, rot) < 0.0f) { r2[0] *= -1.0f; r2[1] *= -1.0f; r2[2] *= -1.0f; }
    len = sqrtf(dot3(r2, r2));
    if (len == 0) { r2[0] = r2[1] = r2[2] = 0; } else { r2[0] /= len; r2[1] /= len; r2[2] /= len; }
    rot[0] = angle * r2[0]; rot[1] = angle * r2[1]; rot[2] = angle * r2[2];
  }
  shtot = 0.5f;
  theta = sqrtf(dot3(rot, rot));
  if (theta > 0.00001f) shtot = sinf(theta * 0.5f) / theta;
  half[0] = half[1] = half[2] = 0.0f; half[3] = rot[0] * -0.5f; half[4] = rot[1] * -0.5f; half[5] = rot[2] * -0.5f;
  se3_exp(half, Mh);
  rt[0] = Mh[0] * T[0] + Mh[4] * T[1] + Mh[8] * T[2];
  rt[1] = Mh[1] * T[0] + Mh[5] * T[1] + Mh[9] * T[2];
  rt[2] = Mh[2] * T[0] + Mh[6] * T[1] + Mh[10] * T[2];
  if (theta > 0.001f) {
    const float denom = dot3(rot, rot);
    const float k = dot3(T, rot) * (1 - 2 * shtot) / denom;
    rt[0] -= rot[0] * k; rt[1] -= rot[1] * k; rt[2] -= rot[2] * k;
  } else {
    const float k = dot3(T, rot) / 24;
    rt[0] -= rot[0] * k; rt[1] -= rot[1] * k; rt[2] -= rot[2] * k;
  }
  rt[0] /= 2 * shtot; rt[1] /= 2 * shtot; rt[2] /= 2 * shtot;
  p[3] = rot[0]; p[4] = rot[1]; p[5] = rot[2]; p[0] = rt[0]; p[1] = rt[1]; p[2] = rt[2];
}

/* ORUtils::Cholesky + Backsub, ORUtils/Cholesky.h:9-71 */
static void chol_solve(const float *mat, int n, const float *v, float *x) {
  float L[36], y[6];
  int c, r, k, i, j;
  for (i = 0; i < n * n; i++) L[i] = mat[i];
  for (c = 0; c < n; c++) {
    float inv_diag = 1;
    for (r = c; r < n; r++) {
      float val = L[c + r * n];
      for (k = 0; k < c; k++) val -= L[c + k * n] * L[k + r * n];
      if (r == c) { L[c + r * n] = val; inv_diag = 1.0f / val; }
      else { L[r + c * n] = val; L[c + r * n] = val * inv_diag; }
    }
  }
  for (i = 0; i < n; i++) { float val = v[i]; for (j = 0; j < i; j++) val -= L[j + i * n] * y[j]; y[i] = val; }
  for (i = 0; i < n; i++) y[i] /= L[i + i * n];
  for (i = n - 1; i >= 0; i--) { float val = y[i]; for (j = i + 1; j < n; j++) val -= L[i + j * n] * x[j]; x[i] = val; }
}

/* ------------------------------------------------------------------------------------------------
 * view: depth conversion and pyramid
 * ---------------------------------------------------------------------------------------------- */

/* convertDepthAffineToFloat, DeviceAgnostic/ITMViewBuilder.h:22-28 */
static void view_convert(port_engine *e) {
  const int n = e->p.width * e->p.height;
  int i;
  for (i = 0; i < n; ++i) {
    const short d = e->raw[i];
    e->depth[i] = (d <= 0 || d > 32000) ? -1.0f : (float)d * e->p.calib_a + e->p.calib_b;
  }
}

/* filterSubsampleWithHoles, DeviceAgnostic/ITMLowLevelEngine.h:26-47; PrepareForEvaluation, ITMDepthTracker.cpp:62-75 */
static void pyramid_of(port_engine *e, float **levels) {
  int l, x, y;
  for (l = 1; l < e->p.n_levels; ++l) {
    const float *src = levels[l - 1];
    float *dst = levels[l];
    const int sw = e->level_w[l - 1], dw = e->level_w[l], dh = e->level_h[l];
    for (y = 0; y < dh; ++y) for (x = 0; x < dw; ++x) {
      float sum = 0.0f, cnt = 0.0f, v;
      v = src[(2 * x) + (2 * y) * sw]; if (v > 0.0f) { sum += v; cnt++; }
      v = src[(2 * x + 1) + (2 * y) * sw]; if (v > 0.0f) { sum += v; cnt++; }
      v = src[(2 * x) + (2 * y + 1) * sw]; if (v > 0.0f) { sum += v; cnt++; }
      v = src[(2 * x + 1) + (2 * y + 1) * sw]; if (v > 0.0f) { sum += v; cnt++; }
      if (cnt > 0) sum /= cnt;
      dst[x + y * dw] = sum;
    }
  }
}

static void view_pyramid(port_engine *e) { pyramid_of(e, e->level_depth); }

/* DepthFiltering / filterDepth: ITMViewBuilder_CPU.cpp:116-128, DeviceAgnostic/ITMViewBuilder.h:31-56 */
static void filter_depth(const port_engine *e, float *out, const float *in) {
  const int W = e->p.width, H = e->p.height;
  int x, y, i, j;
  memset(out, 0, (size_t)W * H * sizeof(float));
  for (y = 2; y < H - 2; y++) for (x = 2; x < W - 2; x++) {
    const float z = in[x + y * W];
    float sigma_z, final_depth = 0.0f, w_sum = 0.0f;
    if (z < 0.0f) { out[x + y * W] = -1.0f; continue; }
    sigma_z = 1.0f / (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * 0.25f);
    for (i = -2; i <= 2; i++) for (j = -2; j <= 2; j++) {
      const float tmpz = in[(x + j) + (y + i) * W];
      float dz, w;
      if (tmpz < 0.0f) continue;
      dz = (tmpz - z); dz *= dz;
      w = WICP_EXP(-0.5f * ((abs(i) + abs(j)) * 1.2232f * 1.2232f + dz * sigma_z * sigma_z));
      w_sum += w;
      final_depth += w * tmpz;
    }
    final_depth /= w_sum;
    out[x + y * W] = final_depth;
  }
}

/* ComputeNormalAndWeights / computeNormalAndWeight: ITMViewBuilder_CPU.cpp:130-143, DeviceAgnostic/ITMViewBuilder.h:59-114 */
static void normal_and_weights(port_engine *e) {
  const int W = e->p.width, H = e->p.height;
  const float kx = e->p.fx, ky = e->p.fy, kz = e->p.cx, kw = e->p.cy;
  int x, y;
  for (y = 2; y < H - 2; y++) for (x = 2; x < W - 2; x++) {
    const int idx = x + y * W;
    const float z = e->depth[idx];
    float zxp, zyp, zxm, zym, xp1x, xp1y, xm1x, xm1y, yp1x, yp1y, ym1x, ym1y, ax, ay, az, bx, by, bz, nx, ny, nz, norm, theta, td;
    if (z < 0.0f) { e->dnormal[idx].w = -1.0f; e->sigma[idx] = -1; continue; }
    zxp = e->depth[idx + 1]; zyp = e->depth[idx + W]; zxm = e->depth[idx - 1]; zym = e->depth[idx - W];
    if (zxp <= 0 || zyp <= 0 || zxm <= 0 || zym <= 0) { e->dnormal[idx].w = -1.0f; e->sigma[idx] = -1; continue; }
    xp1x = zxp * ((x + 1.0f) - kz) * kx; xp1y = zxp * (y - kw) * ky;
    xm1x = zxm * ((x - 1.0f) - kz) * kx; xm1y = zxm * (y - kw) * ky;
    yp1x = zyp * (x - kz) * kx; yp1y = zyp * ((y + 1.0f) - kw) * ky;
    ym1x = zym * (x - kz) * kx; ym1y = zym * ((y - 1.0f) - kw) * ky;
    ax = xp1x - xm1x; ay = xp1y - xm1y; az = zxp - zxm;
    bx = yp1x - ym1x; by = yp1y - ym1y; bz = zyp - zym;
    nx = (ay * bz - az * by); ny = (az * bx - ax * bz); nz = (ax * by - ay * bx);
    if (nx == 0.0f && ny == 0 && nz == 0) { e->dnormal[idx].w = -1.0f; e->sigma[idx] = -1; continue; }
    norm = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    nx *= norm; ny *= norm; nz *= norm;
    e->dnormal[idx].x = nx; e->dnormal[idx].y = ny; e->dnormal[idx].z = nz; e->dnormal[idx].w = 1.0f;
    theta = WICP_ACOS(nz);
    td = theta / ((float)3.1415926535897932384626433832795 * 0.5f - theta);
    e->sigma[idx] = (0.0012f + 0.0019f * (z - 0.4f) * (z - 0.4f) + 0.0001f / sqrtf(z) * td * td);
  }
}

/* ITMViewBuilder_CPU::UpdateView's filter part, ITMViewBuilder_CPU.cpp:50-63 */
static void view_filters(port_engine *e) {
  const size_t P = (size_t)e->p.width * e->p.height;
  if (e->bilateral) {
    if (!e->float_tmp) e->float_tmp = (float *)calloc(P, sizeof(float));
    filter_depth(e, e->float_tmp, e->depth);
    filter_depth(e, e->depth, e->float_tmp);
    filter_depth(e, e->float_tmp, e->depth);
    filter_depth(e, e->depth, e->float_tmp);
    filter_depth(e, e->float_tmp, e->depth);
    memcpy(e->depth, e->float_tmp, P * sizeof(float));
  }
  if (e->wicp) {
    int l;
    if (!e->sigma) {
      e->sigma = (float *)calloc(P, sizeof(float));
      e->dnormal = (V4 *)calloc(P, sizeof(V4));
      e->level_sigma[0] = e->sigma;
      for (l = 1; l < e->p.n_levels; ++l) e->level_sigma[l] = (float *)calloc((size_t)e->level_w[l] * e->level_h[l] + 1, sizeof(float));
    }
    normal_and_weights(e);
  }
}

/* ------------------------------------------------------------------------------------------------
 * ICP tracker
 * ---------------------------------------------------------------------------------------------- */

/* interpolateBilinear_withHoles, DeviceAgnostic/ITMPixelUtils.h:41-71 */
static V4 bilerp_holes(const V4 *img, float px, float py, int W) {
  V4 r;
  const short ix = (short)floorf(px), iy = (short)floorf(py);
  const float dx = px - (float)ix, dy = py - (float)iy;
  const V4 a = img[ix + iy * W], b = img[(ix + 1) + iy * W], c = img[ix + (iy + 1) * W], d = img[(ix + 1) + (iy + 1) * W];
  if (a.w < 0 || b.w < 0 || c.w < 0 || d.w < 0) { r.x = 0; r.y = 0; r.z = 0; r.w = -1.0f; return r; }
  r.x = (a.x * (1.0f - dx) * (1.0f - dy) + b.x * dx * (1.0f - dy) + c.x * (1.0f - dx) * dy + d.x * dx * dy);
  r.y = (a.y * (1.0f - dx) * (1.0f - dy) + b.y * dx * (1.0f - dy) + c.y * (1.0f - dx) * dy + d.y * dx * dy);
  r.z = (a.z * (1.0f - dx) * (1.0f - dy) + b.z * dx * (1.0f - dy) + c.z * (1.0f - dx) * dy + d.z * dx * dy);
  r.w = (a.w * (1.0f - dx) * (1.0f - dy) + b.w * dx * (1.0f - dy) + c.w * (1.0f - dx) * dy + d.w * dx * dy);
  return r;
}

/* One evaluation: ITMDepthTracker_CPU::ComputeGandH (DeviceSpecific/CPU/ITMDepthTracker_CPU.cpp:14-79) over
 * computePerPointGH_Depth(_Ab) (DeviceAgnostic/ITMDepthTracker.h:9-105).  Sums are serial fp32 in raster order. */
static int icp_evaluate(const port_engine *e, int level, const float *approxInvPose, float *f, float *nabla, float *hessian) {
  const int type = e->p.regime[level];
  const int shortIter = (type == ITER_ROTATION) || (type == ITER_TRANSLATION);
  const int np = shortIter ? 3 : 6, nh = shortIter ? 6 : 21;
  const float *depth = e->level_depth[level];
  const int w = e->level_w[level], h = e->level_h[level];
  const float *vi = e->level_intr[level];
  const int SW = e->p.width, SH = e->p.height;
  const float sfx = e->p.fx, sfy = e->p.fy, scx = e->p.cx, scy = e->p.cy; /* scene maps stay at level 0, ITMDepthTracker.cpp:81 */
  const float thresh = e->dist_thresh[level];
  float sumH[21], sumN[6], sumF = 0.0f;
  int nValid = 0, x, y, i, r, c, k;
  if (type == ITER_NONE) return 0;
  memset(sumH, 0, sizeof(sumH)); memset(sumN, 0, sizeof(sumN));
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) {
    const float d = depth[x + y * w];
    float A[6], b, px, py, pz, qx, qy, qz, rx, ry, rz, u, v, ex, ey, ez, dist;
    V4 P, N;
    if (d <= 1e-8f) continue;
    px = d * (((float)x - vi[2]) / vi[0]); py = d * (((float)y - vi[3]) / vi[1]); pz = d;
    m4_apply(approxInvPose, px, py, pz, 1.0f, &qx, &qy, &qz);
    m4_apply(e->pose_pc_M, qx, qy, qz, 1.0f, &rx, &ry, &rz);
    if (rz <= 0.0f) continue;
    u = sfx * rx / rz + scx; v = sfy * ry / rz + scy;
    if (!((u >= 0.0f) && (u <= SW - 2) && (v >= 0.0f) && (v <= SH - 2))) continue;
    P = bilerp_holes(e->points, u, v, SW);
    if (P.w < 0.0f) continue;
    ex = P.x - qx; ey = P.y - qy; ez = P.z - qz;
    dist = ex * ex + ey * ey + ez * ez;
    if (dist > thresh) continue;
    N = bilerp_holes(e->normals, u, v, SW);
    b = N.x * ex + N.y * ey + N.z * ez;
    if (shortIter && type == ITER_TRANSLATION) { A[0] = N.x; A[1] = N.y; A[2] = N.z; }
    else {
      A[0] = +qz * N.y - qy * N.z; A[1] = -qz * N.x + qx * N.z; A[2] = +qy * N.x - qx * N.y;
      if (!shortIter) { A[3] = N.x; A[4] = N.y; A[5] = N.z; }
    }
    nValid++; sumF += b * b;
    for (r = 0, k = 0; r < np; r++) { sumN[r] += b * A[r]; for (c = 0; c <= r; c++, k++) sumH[k] += A[r] * A[c]; }
  }
  for (r = 0, k = 0; r < np; r++) for (c = 0; c <= r; c++, k++) hessian[r + c * 6] = sumH[k];
  for (r = 0; r < np; ++r) for (c = r + 1; c < np; c++) hessian[r + c * 6] = hessian[c + r * 6];
  for (i = 0; i < np; ++i) nabla[i] = sumN[i];
  (void)nh;
  *f = (nValid > 100) ? sqrtf(sumF) / nValid : 1e5f;
  return nValid;
}

/* ITMWeightedICPTracker_CPU::ComputeGandH (ITMWeightedICPTracker_CPU.cpp:14-85) + computePerPointGH_wICP
 * (DeviceAgnostic/ITMWeightedICPTracker.h:9-105): per pixel local sums, then added to the totals (the order matters in fp32) */
static int wicp_evaluate(const port_engine *e, int level, const float *approxInvPose, float *f, float *nabla, float *hessian) {
  const int type = e->p.regime[level];
  const int shortIter = (type == ITER_ROTATION) || (type == ITER_TRANSLATION);
  const int np = shortIter ? 3 : 6, nh = shortIter ? 6 : 21;
  const float *depth = e->level_depth[level], *weight = e->level_sigma[level];
  const int w = e->level_w[level], h = e->level_h[level];
  const float *vi = e->level_intr[level];
  const int SW = e->p.width, SH = e->p.height;
  const float sfx = e->p.fx, sfy = e->p.fy, scx = e->p.cx, scy = e->p.cy;
  const float thresh = e->dist_thresh[level], minSigmaZ = 0.0012f;
  float sumH[21], sumN[6], sumF = 0.0f;
  int nValid = 0, x, y, i, r, c, k;
  if (type == ITER_NONE) return 0;
  memset(sumH, 0, sizeof(sumH)); memset(sumN, 0, sizeof(sumN));
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) {
    const float d = depth[x + y * w];
    const float lw = weight[x + y * w] > 0 ? minSigmaZ / weight[x + y * w] * 0.5f + 0.5f : 0.0f;
    float A[6], b, px, py, pz, qx, qy, qz, rx, ry, rz, u, v, ex, ey, ez, dist, lH[21], lN[6], lF = 0;
    V4 P, N;
    for (i = 0; i < np; i++) lN[i] = 0.0f;
    for (i = 0; i < nh; i++) lH[i] = 0.0f;
    if (d <= 1e-8f) continue;
    px = d * (((float)x - vi[2]) / vi[0]); py = d * (((float)y - vi[3]) / vi[1]); pz = d;
    m4_apply(approxInvPose, px, py, pz, 1.0f, &qx, &qy, &qz);
    m4_apply(e->pose_pc_M, qx, qy, qz, 1.0f, &rx, &ry, &rz);
    if (rz <= 0.0f) continue;
    u = sfx * rx / rz + scx; v = sfy * ry / rz + scy;
    if (!((u >= 0.0f) && (u <= SW - 2) && (v >= 0.0f) && (v <= SH - 2))) continue;
    P = bilerp_holes(e->points, u, v, SW);
    if (P.w < 0.0f) continue;
    ex = P.x - qx; ey = P.y - qy; ez = P.z - qz;
    dist = ex * ex + ey * ey + ez * ez;
    if (dist > thresh) continue;
    N = bilerp_holes(e->normals, u, v, SW);
    b = N.x * ex + N.y * ey + N.z * ez;
    lF += b * b * lw * lw;
    N.x *= lw; N.y *= lw; N.z *= lw; N.w *= lw;
    if (shortIter && type == ITER_TRANSLATION) { A[0] = N.x; A[1] = N.y; A[2] = N.z; }
    else {
      A[0] = +qz * N.y - qy * N.z; A[1] = -qz * N.x + qx * N.z; A[2] = +qy * N.x - qx * N.y;
      if (!shortIter) { A[3] = N.x; A[4] = N.y; A[5] = N.z; }
    }
    for (r = 0, k = 0; r < np; r++) { lN[r] += b * A[r]; for (c = 0; c <= r; c++, k++) lH[k] += A[r] * A[c]; }
    nValid++; sumF += lF;
    for (i = 0; i < np; i++) sumN[i] += lN[i];
    for (i = 0; i < nh; i++) sumH[i] += lH[i];
  }
  for (r = 0, k = 0; r < np; r++) for (c = 0; c <= r; c++, k++) hessian[r + c * 6] = sumH[k];
  for (r = 0; r < np; ++r) for (c = r + 1; c < np; c++) hessian[r + c * 6] = hessian[c + r * 6];
  for (i = 0; i < np; ++i) nabla[i] = sumN[i];
  *f = (nValid > 100) ? sqrtf(sumF) / nValid : 1e5f;
  return nValid;
}

/* ITMDepthTracker::ComputeDelta, ITMDepthTracker.cpp:85-102 */
static void icp_delta(float *step, const float *nabla, const float *hessian, int shortIter) {
  int i, r, c;
  for (i = 0; i < 6; i++) step[i] = 0;
  if (shortIter) {
    float small[9];
    for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) small[r + c * 3] = hessian[r + c * 6];
    chol_solve(small, 3, nabla, step);
  } else chol_solve(hessian, 6, nabla, step);
}

/* ITMDepthTracker::ApplyDelta, ITMDepthTracker.cpp:114-143 */
static void icp_apply(const float *old, const float *delta, int type, float *out) {
  float s[6], T[16];
  if (type == ITER_ROTATION) { s[0] = delta[0]; s[1] = delta[1]; s[2] = delta[2]; s[3] = s[4] = s[5] = 0.0f; }
  else if (type == ITER_TRANSLATION) { s[0] = s[1] = s[2] = 0.0f; s[3] = delta[0]; s[4] = delta[1]; s[5] = delta[2]; }
  else memcpy(s, delta, sizeof(s));
  T[0] = 1.0f; T[4] = s[2]; T[8] = -s[1]; T[12] = s[3];
  T[1] = -s[2]; T[5] = 1.0f; T[9] = s[0]; T[13] = s[4];
  T[2] = s[1]; T[6] = -s[0]; T[10] = 1.0f; T[14] = s[5];
  T[3] = 0.0f; T[7] = 0.0f; T[11] = 0.0f; T[15] = 1.0f;
  m4_product(T, old, out);
}

/* ITMDepthTracker::TrackCamera, ITMDepthTracker.cpp:145-199 */
static void icp_track(port_engine *e) {
  float f_old, f_new, Hgood[36], Hnew[36], A[36], ngood[6], nnew[6], step[6];
  int level, it, i, nValid;
  view_pyramid(e);
  memset(Hgood, 0, sizeof(Hgood)); memset(ngood, 0, sizeof(ngood));
  for (level = e->p.n_levels - 1; level >= e->p.no_icp_run_till_level; level--) {
    const int type = e->p.regime[level];
    float inv[16], goodM[16], goodP[6], lambda = 1.0f;
    if (type == ITER_NONE) continue;
    m4_inverse(e->pose_M, inv);
    memcpy(goodM, e->pose_M, sizeof(goodM)); memcpy(goodP, e->pose_params, sizeof(goodP));
    f_old = 1e20f;
    for (it = 0; it < e->iters[level]; it++) {
      float stepLen = 0.0f, tmpM[16];
      memset(Hnew, 0, sizeof(Hnew)); memset(nnew, 0, sizeof(nnew));
      nValid = icp_evaluate(e, level, inv, &f_new, nnew, Hnew);
      if (nValid <= 0 || f_new > f_old) {
        memcpy(e->pose_M, goodM, sizeof(goodM)); memcpy(e->pose_params, goodP, sizeof(goodP));
        m4_inverse(e->pose_M, inv);
        lambda *= 10.0f;
      } else {
        memcpy(goodM, e->pose_M, sizeof(goodM)); memcpy(goodP, e->pose_params, sizeof(goodP));
        f_old = f_new;
        for (i = 0; i < 36; ++i) Hgood[i] = Hnew[i] / nValid;
        for (i = 0; i < 6; ++i) ngood[i] = nnew[i] / nValid;
        lambda /= 10.0f;
      }
      for (i = 0; i < 36; ++i) A[i] = Hgood[i];
      for (i = 0; i < 6; ++i) A[i + i * 6] *= 1.0f + lambda;
      icp_delta(step, ngood, A, type != ITER_BOTH);
      icp_apply(inv, step, type, inv);
      /* pose_d->SetInvM(inv); Coerce(); inv = GetInvM()   (ITMPose.cpp:309-326) */
      m4_inverse(inv, tmpM);
      se3_log(tmpM, e->pose_params);
      se3_exp(e->pose_params, e->pose_M);
      m4_inverse(e->pose_M, inv);
      for (i = 0; i < 6; i++) stepLen += step[i] * step[i];
      if (sqrtf(stepLen) / 6 < e->p.icp_termination) break;
    }
  }
}

/* ITMWeightedICPTracker::TrackCamera, ITMWeightedICPTracker.cpp:164-192: plain Gauss-Newton, f_old is never updated */
static void wicp_track(port_engine *e) {
  float f_old = 1e10f, f_new, H[36], nabla[6], step[6], inv[16];
  int level, it, i, nValid;
  view_pyramid(e);
  pyramid_of(e, e->level_sigma);
  m4_inverse(e->pose_M, inv);
  memset(H, 0, sizeof(H)); memset(nabla, 0, sizeof(nabla));
  for (level = e->p.n_levels - 1; level >= e->p.no_icp_run_till_level; level--) {
    const int type = e->p.regime[level];
    if (type == ITER_NONE) continue;
    for (it = 0; it < e->iters[level]; it++) {
      float stepLen = 0.0f, tmpM[16];
      nValid = wicp_evaluate(e, level, inv, &f_new, nabla, H);
      if (nValid <= 0) break;
      if (f_new > f_old) break;
      icp_delta(step, nabla, H, type != ITER_BOTH);
      icp_apply(inv, step, type, inv);
      m4_inverse(inv, tmpM);
      se3_log(tmpM, e->pose_params);
      se3_exp(e->pose_params, e->pose_M);
      m4_inverse(e->pose_M, inv);
      for (i = 0; i < 6; i++) stepLen += step[i] * step[i];
      if (sqrtf(stepLen) / 6 < e->p.icp_termination) break;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * allocation + visible list
 * ---------------------------------------------------------------------------------------------- */

static unsigned hash_of(int x, int y, int z, unsigned mask) { /* ITMRepresentationAccess.h:8-10 */
  return (((unsigned)x * 73856093u) ^ ((unsigned)y * 19349669u) ^ ((unsigned)z * 83492791u)) & mask;
}

/* buildHashAllocAndVisibleTypePP, DeviceAgnostic/ITMSceneReconstructionEngine.h:141-241 */
static void alloc_pixel(port_engine *e, int x, int y, const float *invM, float invfx, float invfy, float oneOverBlock) {
  const port_params *p = &e->p;
  const float d = e->depth[x + y * p->width], mu = p->mu;
  const unsigned mask = (unsigned)p->n_bucket - 1u;
  float cx, cy, cz, norm, ax, ay, az, bx, by, bz, dx, dy, dz, k;
  int steps, i;
  if (d <= 0 || (d - mu) < 0 || (d - mu) < p->vf_min || (d + mu) > p->vf_max) return;
  cz = d; cx = cz * (((float)x - p->cx) * invfx); cy = cz * (((float)y - p->cy) * invfy);
  norm = sqrtf(cx * cx + cy * cy + cz * cz);
  k = 1.0f - mu / norm;
  m4_apply(invM, cx * k, cy * k, cz * k, 1.0f, &ax, &ay, &az);
  ax *= oneOverBlock; ay *= oneOverBlock; az *= oneOverBlock;
  k = 1.0f + mu / norm;
  m4_apply(invM, cx * k, cy * k, cz * k, 1.0f, &bx, &by, &bz);
  bx *= oneOverBlock; by *= oneOverBlock; bz *= oneOverBlock;
  dx = bx - ax; dy = by - ay; dz = bz - az;
  norm = sqrtf(dx * dx + dy * dy + dz * dz);
  steps = (int)ceilf(2.0f * norm);
  k = (float)(steps - 1);
  dx /= k; dy /= k; dz /= k;
  for (i = 0; i < steps; i++) {
    const short qx = (short)floorf(ax), qy = (short)floorf(ay), qz = (short)floorf(az);
    int idx = (int)hash_of(qx, qy, qz, mask), found = 0;
    HashEntry h = e->hash[idx];
    if (h.x == qx && h.y == qy && h.z == qz && h.ptr >= -1) { e->visible_type[idx] = (h.ptr == -1) ? 2 : 1; found = 1; }
    if (!found) {
      int excess = 0;
      if (h.ptr >= -1) {
        while (h.offset >= 1) {
          idx = p->n_bucket + h.offset - 1;
          h = e->hash[idx];
          if (h.x == qx && h.y == qy && h.z == qz && h.ptr >= -1) { e->visible_type[idx] = (h.ptr == -1) ? 2 : 1; found = 1; break; }
        }
        excess = 1;
      }
      if (!found) {
        e->alloc_type[idx] = excess ? 2 : 1;
        if (!excess) e->visible_type[idx] = 1;
        e->block_coords[4 * idx] = qx; e->block_coords[4 * idx + 1] = qy; e->block_coords[4 * idx + 2] = qz; e->block_coords[4 * idx + 3] = 1;
      }
    }
    ax += dx; ay += dy; az += dz;
  }
}

/* checkPointVisibility<false> / checkBlockVisibility<false>, DeviceAgnostic/ITMSceneReconstructionEngine.h:244-342 */
static int corner_visible(const port_engine *e, float x, float y, float z) {
  float bx, by, bz;
  m4_apply(e->pose_M, x, y, z, 1.0f, &bx, &by, &bz);
  if (bz < 1e-10f) return 0;
  bx = e->p.fx * bx / bz + e->p.cx; by = e->p.fy * by / bz + e->p.cy;
  return bx >= 0 && bx < e->p.width && by >= 0 && by < e->p.height;
}
static int block_visible(const port_engine *e, const HashEntry *h) {
  const float f = (float)BLOCK * e->p.voxel_size;
  float x = (float)h->x * f, y = (float)h->y * f, z = (float)h->z * f;
  if (corner_visible(e, x, y, z)) return 1;
  z += f; if (corner_visible(e, x, y, z)) return 1;
  y += f; if (corner_visible(e, x, y, z)) return 1;
  x += f; if (corner_visible(e, x, y, z)) return 1;
  z -= f; if (corner_visible(e, x, y, z)) return 1;
  y -= f; if (corner_visible(e, x, y, z)) return 1;
  x -= f; y += f; if (corner_visible(e, x, y, z)) return 1;
  x += f; y -= f; z += f; if (corner_visible(e, x, y, z)) return 1;
  return 0;
}

/* AllocateSceneFromDepth, DeviceSpecific/CPU/ITMSceneReconstructionEngine_CPU.cpp:117-291 (no swapping) */
static void scene_allocate(port_engine *e, int onlyVisible) {
  const port_params *p = &e->p;
  float invM[16];
  const float invfx = 1.0f / p->fx, invfy = 1.0f / p->fy;
  const float oneOverBlock = 1.0f / (p->voxel_size * BLOCK);
  int i, x, y, nvis = 0;
  m4_inverse(e->pose_M, invM);
  memset(e->alloc_type, 0, (size_t)e->n_entries);
  for (i = 0; i < e->n_visible; i++) e->visible_type[e->visible_ids[i]] = 3;
  for (y = 0; y < p->height; ++y) for (x = 0; x < p->width; ++x) alloc_pixel(e, x, y, invM, invfx, invfy, oneOverBlock);
  if (!onlyVisible) {
    for (i = 0; i < e->n_entries; i++) {
      const unsigned char t = e->alloc_type[i];
      if (t == 1) {
        const int vba = e->last_free_block--;
        if (vba >= 0) {
          HashEntry h; memset(&h, 0, sizeof(h));
          h.x = e->block_coords[4 * i]; h.y = e->block_coords[4 * i + 1]; h.z = e->block_coords[4 * i + 2];
          h.ptr = e->vba_list[vba]; h.offset = 0;
          e->hash[i] = h;
        }
      } else if (t == 2) {
        const int vba = e->last_free_block--, exl = e->last_free_excess--;
        if (vba >= 0 && exl >= 0) {
          HashEntry h; int off; memset(&h, 0, sizeof(h));
          h.x = e->block_coords[4 * i]; h.y = e->block_coords[4 * i + 1]; h.z = e->block_coords[4 * i + 2];
          h.ptr = e->vba_list[vba]; h.offset = 0;
          off = e->excess_list[exl];
          e->hash[i].offset = off + 1;
          e->hash[p->n_bucket + off] = h;
          e->visible_type[p->n_bucket + off] = 1;
        }
      }
    }
  }
  for (i = 0; i < e->n_entries; i++) {
    unsigned char t = e->visible_type[i];
    if (t == 3) { if (!block_visible(e, &e->hash[i])) t = 0; e->visible_type[i] = t; }
    if (t > 0) e->visible_ids[nvis++] = i;
  }
  e->n_visible = nvis;
}

/* ------------------------------------------------------------------------------------------------
 * integration
 * ---------------------------------------------------------------------------------------------- */

/* IntegrateIntoScene (ITMSceneReconstructionEngine_CPU.cpp:48-114) + computeUpdatedVoxelDepthInfo
 * (DeviceAgnostic/ITMSceneReconstructionEngine.h:10-56) */
static void scene_integrate(port_engine *e) {
  const port_params *p = &e->p;
  const float *M = e->pose_M;
  const float vs = p->voxel_size, mu = p->mu;
  int n, x, y, z;
  for (n = 0; n < e->n_visible; n++) {
    const HashEntry *h = &e->hash[e->visible_ids[n]];
    Voxel *blk;
    int gx, gy, gz;
    if (h->ptr < 0) continue;
    gx = h->x * BLOCK; gy = h->y * BLOCK; gz = h->z * BLOCK;
    blk = e->voxels + (size_t)h->ptr * BLOCK3;
    for (z = 0; z < BLOCK; z++) for (y = 0; y < BLOCK; y++) for (x = 0; x < BLOCK; x++) {
      Voxel *v = &blk[x + y * BLOCK + z * BLOCK * BLOCK];
      float mx, my, mz, cx, cy, cz, ix, iy, dm, eta, oldF, newF;
      int oldW, newW;
      if (p->stop_at_max_w && v->w_depth == p->max_w) continue;
      mx = (float)(gx + x) * vs; my = (float)(gy + y) * vs; mz = (float)(gz + z) * vs;
      m4_apply(M, mx, my, mz, 1.0f, &cx, &cy, &cz);
      if (cz <= 0) continue;
      ix = p->fx * cx / cz + p->cx; iy = p->fy * cy / cz + p->cy;
      if (ix < 1 || ix > p->width - 2 || iy < 1 || iy > p->height - 2) continue;
      dm = e->depth[(int)(ix + 0.5f) + (int)(iy + 0.5f) * p->width];
      if (dm <= 0.0) continue;
      eta = dm - cz;
      if (eta < -mu) continue;
      oldF = (float)v->sdf / 32767.0f; oldW = v->w_depth;
      newF = (1.0f < eta / mu) ? 1.0f : eta / mu; newW = 1;
      newF = oldW * oldF + newW * newF;
      newW = oldW + newW;
      newF /= newW;
      newW = (newW < p->max_w) ? newW : p->max_w;
      v->sdf = (short)(newF * 32767.0f);
      v->w_depth = (unsigned char)newW;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * rendering for tracking
 * ---------------------------------------------------------------------------------------------- */

/* CreateExpectedDepths (DeviceSpecific/CPU/ITMVisualisationEngine_CPU.cpp:94-152) with ProjectSingleBlock and
 * CreateRenderingBlocks (DeviceAgnostic/ITMVisualisationEngine.h:28-90).  The rendering-block list only splits the
 * bounding box into <=16x16 pieces; it is kept (as a counter) for its MAX_RENDERING_BLOCKS cut-off. */
/* for any camera (pose M, intrinsics k = fx fy cx cy), image size and visible list: the live view and the free view share it */
static void expected_depths_for(const port_engine *e, const float *M, const float *k, int W, int H, const int *visible_ids, int n_visible,
                                V2 *minmax) {
  const port_params *p = &e->p;
  int i, n, corner, x, y, numBlocks = 0;
  for (i = 0; i < W * H; ++i) { minmax[i].x = FAR_AWAY; minmax[i].y = VERY_CLOSE; }
  for (n = 0; n < n_visible; ++n) {
    const HashEntry *h = &e->hash[visible_ids[n]];
    int ulx = W / MINMAX_SUB, uly = H / MINMAX_SUB, lrx = -1, lry = -1, need;
    float zmin = FAR_AWAY, zmax = VERY_CLOSE;
    if (h->ptr < 0) continue;
    for (corner = 0; corner < 8; ++corner) {
      const short tx = (short)(h->x + ((corner & 1) ? 1 : 0)), ty = (short)(h->y + ((corner & 2) ? 1 : 0)), tz = (short)(h->z + ((corner & 4) ? 1 : 0));
      float cx, cy, cz, px, py;
      m4_apply(M, (float)tx * (float)BLOCK * p->voxel_size, (float)ty * (float)BLOCK * p->voxel_size,
               (float)tz * (float)BLOCK * p->voxel_size, 1.0f, &cx, &cy, &cz);
      if (cz < 1e-6) continue;
      px = (k[0] * cx / cz + k[2]) / MINMAX_SUB; py = (k[1] * cy / cz + k[3]) / MINMAX_SUB;
      if (ulx > floorf(px)) ulx = (int)floorf(px);
      if (lrx < ceilf(px)) lrx = (int)ceilf(px);
      if (uly > floorf(py)) uly = (int)floorf(py);
      if (lry < ceilf(py)) lry = (int)ceilf(py);
      if (zmin > cz) zmin = cz;
      if (zmax < cz) zmax = cz;
    }
    if (ulx < 0) ulx = 0;
    if (uly < 0) uly = 0;
    if (lrx >= W) lrx = W - 1;
    if (lry >= H) lry = H - 1;
    if (ulx > lrx || uly > lry) continue;
    if (zmin < VERY_CLOSE) zmin = VERY_CLOSE;
    if (zmax < VERY_CLOSE) continue;
    need = (int)ceilf((float)(lrx - ulx + 1) / 16.0f) * (int)ceilf((float)(lry - uly + 1) / 16.0f);
    if (numBlocks + need >= MAX_RENDERING_BLOCKS) continue;
    numBlocks += need;
    for (y = uly; y <= lry; ++y) for (x = ulx; x <= lrx; ++x) {
      V2 *px2 = &minmax[x + y * W];
      if (px2->x > zmin) px2->x = zmin;
      if (px2->y < zmax) px2->y = zmax;
    }
  }
}

static void render_expected_depths(port_engine *e) {
  const float k[4] = {e->p.fx, e->p.fy, e->p.cx, e->p.cy};
  expected_depths_for(e, e->pose_M, k, e->p.width, e->p.height, e->visible_ids, e->n_visible, e->minmax);
}

typedef struct { int bx, by, bz, base; } BlockCache; /* ITMVoxelBlockHash::IndexCache, Objects/ITMVoxelBlockHash.h:27-33 */

/* readVoxel(...).sdf, DeviceAgnostic/ITMRepresentationAccess.h:86-119 (pointToVoxelBlockPos :12-20) */
static int voxel_sdf(const port_engine *e, int x, int y, int z, int *found, BlockCache *cache) {
  const int bx = ((x < 0) ? x - BLOCK + 1 : x) / BLOCK, by = ((y < 0) ? y - BLOCK + 1 : y) / BLOCK, bz = ((z < 0) ? z - BLOCK + 1 : z) / BLOCK;
  const int lin = x + (y - bx) * BLOCK + (z - by) * BLOCK * BLOCK - bz * BLOCK3;
  int idx;
  if (bx == cache->bx && by == cache->by && bz == cache->bz) { *found = 1; return e->voxels[cache->base + lin].sdf; }
  idx = (int)hash_of(bx, by, bz, (unsigned)e->p.n_bucket - 1u);
  for (;;) {
    const HashEntry h = e->hash[idx];
    if (h.x == bx && h.y == by && h.z == bz && h.ptr >= 0) {
      *found = 1; cache->bx = bx; cache->by = by; cache->bz = bz; cache->base = h.ptr * BLOCK3;
      return e->voxels[cache->base + lin].sdf;
    }
    if (h.offset < 1) break;
    idx = e->p.n_bucket + h.offset - 1;
  }
  *found = 0;
  return 32767;
}

static float round_half(float v) { return (v < 0) ? (v - 0.5f) : (v + 0.5f); } /* ROUND, ORUtils/MathUtils.h:22 */

/* readFromSDF_float_uninterpolated, ITMRepresentationAccess.h:145-158 */
static float sdf_nearest(const port_engine *e, float x, float y, float z, int *found, BlockCache *c) {
  return (float)voxel_sdf(e, (int)round_half(x), (int)round_half(y), (int)round_half(z), found, c) / 32767.0f;
}

/* readFromSDF_float_interpolated, ITMRepresentationAccess.h:161-185 */
static float sdf_trilinear(const port_engine *e, float x, float y, float z, BlockCache *c) {
  const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
  const float cx = x - fx, cy = y - fy, cz = z - fz;
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  int f;
  float v1, v2, r1, r2;
  v1 = (float)voxel_sdf(e, ix, iy, iz, &f, c); v2 = (float)voxel_sdf(e, ix + 1, iy, iz, &f, c);
  r1 = (1.0f - cx) * v1 + cx * v2;
  v1 = (float)voxel_sdf(e, ix, iy + 1, iz, &f, c); v2 = (float)voxel_sdf(e, ix + 1, iy + 1, iz, &f, c);
  r1 = (1.0f - cy) * r1 + cy * ((1.0f - cx) * v1 + cx * v2);
  v1 = (float)voxel_sdf(e, ix, iy, iz + 1, &f, c); v2 = (float)voxel_sdf(e, ix + 1, iy, iz + 1, &f, c);
  r2 = (1.0f - cx) * v1 + cx * v2;
  v1 = (float)voxel_sdf(e, ix, iy + 1, iz + 1, &f, c); v2 = (float)voxel_sdf(e, ix + 1, iy + 1, iz + 1, &f, c);
  r2 = (1.0f - cy) * r2 + cy * ((1.0f - cx) * v1 + cx * v2);
  return ((1.0f - cz) * r1 + cz * r2) / 32767.0f;
}

/* castRay, DeviceAgnostic/ITMVisualisationEngine.h:93-158: one pixel of a camera with inverse pose invM and intrinsics k */
static void cast_ray(const port_engine *e, int x, int y, V2 mm, const float *invM, const float *k, V4 *out) {
  const port_params *p = &e->p;
  const float oneOverVoxel = 1.0f / p->voxel_size, invfx = 1.0f / k[0], invfy = 1.0f / k[1];
  const float stepScale = p->mu * oneOverVoxel;
  float cz, cx, cy, len, lenMax, sx, sy, sz, ex, ey, ez, dx, dy, dz, kn, px, py, pz, sdf = 1.0f, step;
  int found;
  BlockCache cache;
  cache.bx = cache.by = cache.bz = 0x7fffffff; cache.base = -1;
  cz = mm.x; cx = cz * (((float)x - k[2]) * invfx); cy = cz * (((float)y - k[3]) * invfy);
  len = sqrtf(cx * cx + cy * cy + cz * cz) * oneOverVoxel;
  m4_apply(invM, cx, cy, cz, 1.0f, &sx, &sy, &sz); sx *= oneOverVoxel; sy *= oneOverVoxel; sz *= oneOverVoxel;
  cz = mm.y; cx = cz * (((float)x - k[2]) * invfx); cy = cz * (((float)y - k[3]) * invfy);
  lenMax = sqrtf(cx * cx + cy * cy + cz * cz) * oneOverVoxel;
  m4_apply(invM, cx, cy, cz, 1.0f, &ex, &ey, &ez); ex *= oneOverVoxel; ey *= oneOverVoxel; ez *= oneOverVoxel;
  dx = ex - sx; dy = ey - sy; dz = ez - sz;
  kn = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
  dx *= kn; dy *= kn; dz *= kn;
  px = sx; py = sy; pz = sz;
  while (len < lenMax) {
    sdf = sdf_nearest(e, px, py, pz, &found, &cache);
    if (!found) step = BLOCK;
    else {
      if (sdf <= 0.1f && sdf >= -0.5f) sdf = sdf_trilinear(e, px, py, pz, &cache);
      if (sdf <= 0.0f) break;
      step = (sdf * stepScale < 1.0f) ? 1.0f : sdf * stepScale;
    }
    px += step * dx; py += step * dy; pz += step * dz; len += step;
  }
  if (sdf <= 0.0f) {
    step = sdf * stepScale; px += step * dx; py += step * dy; pz += step * dz;
    sdf = sdf_trilinear(e, px, py, pz, &cache);
    step = sdf * stepScale; px += step * dx; py += step * dy; pz += step * dz;
    out->w = 1.0f;
  } else out->w = 0.0f;
  out->x = px; out->y = py; out->z = pz;
}

/* GenericRaycast, ITMVisualisationEngine_CPU.cpp:155-188 */
static void raycast_with_inverse(const port_engine *e, const float *invM, const float *k, int W, int H, const V2 *minmax, V4 *out) {
  int x, y;
  for (y = 0; y < H; ++y) for (x = 0; x < W; ++x)
    cast_ray(e, x, y, minmax[(int)floorf((float)x / MINMAX_SUB) + (int)floorf((float)y / MINMAX_SUB) * W], invM, k, &out[x + y * W]);
}
static void raycast_for(const port_engine *e, const float *M, const float *k, int W, int H, const V2 *minmax, V4 *out) {
  float invM[16];
  m4_inverse(M, invM);
  raycast_with_inverse(e, invM, k, W, H, minmax, out);
}

static void render_raycast(port_engine *e) {
  const float k[4] = {e->p.fx, e->p.fy, e->p.cx, e->p.cy};
  raycast_for(e, e->pose_M, k, e->p.width, e->p.height, e->minmax, e->raycast);
}

/* CreateICPMaps_common (ITMVisualisationEngine_CPU.cpp:267-287) + processPixelICP<true> / computeNormalAndAngle<true>
 * (DeviceAgnostic/ITMVisualisationEngine.h:192-349) + the bookkeeping of ITMTrackingController::Prepare (:33-39) */
static void render_icp_maps(port_engine *e) {
  const port_params *p = &e->p;
  const int W = p->width, H = p->height;
  const float vs = p->voxel_size;
  float invM[16], lx, ly, lz;
  int x, y;
  render_raycast(e);
  memcpy(e->pose_pc_M, e->pose_M, sizeof(e->pose_pc_M));
  m4_inverse(e->pose_M, invM);
  lx = -invM[8]; ly = -invM[9]; lz = -invM[10];
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) {
    const int id = x + y * W;
    const V4 pt = e->raycast[id];
    int ok = pt.w > 0.0f;
    float nx = 0, ny = 0, nz = 0, angle = 0;
    if (ok) {
      if (y <= 2 || y >= H - 3 || x <= 2 || x >= W - 3) ok = 0;
      else {
        V4 xp = e->raycast[(x + 2) + y * W], yp = e->raycast[x + (y + 2) * W], xm = e->raycast[(x - 2) + y * W], ym = e->raycast[x + (y - 2) * W];
        float ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
        int plus1 = 0;
        if (xp.w <= 0 || yp.w <= 0 || xm.w <= 0 || ym.w <= 0) plus1 = 1;
        else {
          float la, lb, l;
          ax = xp.x - xm.x; ay = xp.y - xm.y; az = xp.z - xm.z; bx = yp.x - ym.x; by = yp.y - ym.y; bz = yp.z - ym.z;
          la = ax * ax + ay * ay + az * az; lb = bx * bx + by * by + bz * bz;
          l = (la < lb) ? lb : la;
          if (l * vs * vs > (0.15f * 0.15f)) plus1 = 1;
        }
        if (plus1) {
          xp = e->raycast[(x + 1) + y * W]; yp = e->raycast[x + (y + 1) * W]; xm = e->raycast[(x - 1) + y * W]; ym = e->raycast[x + (y - 1) * W];
          ax = xp.x - xm.x; ay = xp.y - xm.y; az = xp.z - xm.z; bx = yp.x - ym.x; by = yp.y - ym.y; bz = yp.z - ym.z;
          if (xp.w <= 0 || yp.w <= 0 || xm.w <= 0 || ym.w <= 0) ok = 0;
        }
        if (ok) {
          float s;
          nx = -(ay * bz - az * by); ny = -(az * bx - ax * bz); nz = -(ax * by - ay * bx);
          s = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
          nx *= s; ny *= s; nz *= s;
          angle = nx * lx + ny * ly + nz * lz;
          if (!(angle > 0.0)) ok = 0;
        }
      }
    }
    if (ok) {
      const float g = (0.8f * angle + 0.2f) * 255.0f;
      const unsigned char gc = (unsigned char)g;
      e->raycast_image[4 * id] = gc; e->raycast_image[4 * id + 1] = gc; e->raycast_image[4 * id + 2] = gc; e->raycast_image[4 * id + 3] = gc;
      e->points[id].x = pt.x * vs; e->points[id].y = pt.y * vs; e->points[id].z = pt.z * vs; e->points[id].w = 1.0f;
      e->normals[id].x = nx; e->normals[id].y = ny; e->normals[id].z = nz; e->normals[id].w = 0.0f;
    } else {
      const V4 bad = {0.0f, 0.0f, 0.0f, -1.0f};
      e->points[id] = bad; e->normals[id] = bad;
      memset(&e->raycast_image[4 * id], 0, 4);
    }
  }
  if (e->age == -1) e->age = -2; else e->age = 0;
}

/* ------------------------------------------------------------------------------------------------
 * engine object + exported C API (mirrors oracle/ref_harness.cpp so that one Python wrapper drives both)
 * ---------------------------------------------------------------------------------------------- */

/* ResetScene, ITMSceneReconstructionEngine_CPU.cpp:25-45 */
/* ------------------------------------------------------------------------------------------------
 * ForwardRender, free-view rendering, meshing
 * ---------------------------------------------------------------------------------------------- */

/* computeNormalAndAngle<useSmoothing = true>, DeviceAgnostic/ITMVisualisationEngine.h:192-253 */
static int normal_from_points(const V4 *img, int x, int y, int W, int H, float vs, const float *light, float *n, float *angle) {
  V4 xp, yp, xm, ym;
  float ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0, la, lb, scale;
  int plus1 = 0;
  if (y <= 2 || y >= H - 3 || x <= 2 || x >= W - 3) return 0;
  xp = img[(x + 2) + y * W]; yp = img[x + (y + 2) * W]; xm = img[(x - 2) + y * W]; ym = img[x + (y - 2) * W];
  if (xp.w <= 0 || yp.w <= 0 || xm.w <= 0 || ym.w <= 0) plus1 = 1;
  else {
    ax = xp.x - xm.x; ay = xp.y - xm.y; az = xp.z - xm.z;
    bx = yp.x - ym.x; by = yp.y - ym.y; bz = yp.z - ym.z;
    la = ax * ax + ay * ay + az * az; lb = bx * bx + by * by + bz * bz;
    if (((la < lb) ? lb : la) * vs * vs > (0.15f * 0.15f)) plus1 = 1;
  }
  if (plus1) {
    xp = img[(x + 1) + y * W]; yp = img[x + (y + 1) * W]; xm = img[(x - 1) + y * W]; ym = img[x + (y - 1) * W];
    ax = xp.x - xm.x; ay = xp.y - xm.y; az = xp.z - xm.z;
    bx = yp.x - ym.x; by = yp.y - ym.y; bz = yp.z - ym.z;
    if (xp.w <= 0 || yp.w <= 0 || xm.w <= 0 || ym.w <= 0) return 0;
  }
  n[0] = -(ay * bz - az * by); n[1] = -(az * bx - ax * bz); n[2] = -(ax * by - ay * bx);
  scale = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  n[0] *= scale; n[1] *= scale; n[2] *= scale;
  *angle = n[0] * light[0] + n[1] * light[1] + n[2] * light[2];
  return (*angle > 0.0) ? 1 : 0;
}

/* ITMTrackingState::TrackerFarFromPointCloud, Objects/ITMTrackingState.h:41-59 */
static int tracker_far_from_point_cloud(const port_engine *e) {
  float ca[3], cb[3], d0, d1, d2;
  const float *A = e->pose_pc_M, *B = e->pose_M;
  int r;
  if (e->age < 0) return 1;
  if (e->age > 5) return 1;
  for (r = 0; r < 3; ++r) {
    ca[r] = -1.0f * (A[r * 4 + 0] * A[12] + A[r * 4 + 1] * A[13] + A[r * 4 + 2] * A[14]);
    cb[r] = -1.0f * (B[r * 4 + 0] * B[12] + B[r * 4 + 1] * B[13] + B[r * 4 + 2] * B[14]);
  }
  d0 = ca[0] - cb[0]; d1 = ca[1] - cb[1]; d2 = ca[2] - cb[2];
  return (d0 * d0 + d1 * d1 + d2 * d2 > 0.0005f) ? 1 : 0;
}

/* ForwardRender_common, ITMVisualisationEngine_CPU.cpp:289-354 (forwardProjectPixel, DeviceAgnostic/...VisualisationEngine.h:160-173;
 * processPixelForwardRender<true> :351-366) */
static void render_forward(port_engine *e) {
  const port_params *p = &e->p;
  const int W = p->width, H = p->height;
  const float k[4] = {p->fx, p->fy, p->cx, p->cy};
  float invM[16], light[3];
  int x, y, n = 0, i;
  m4_inverse(e->pose_M, invM);
  light[0] = -invM[8]; light[1] = -invM[9]; light[2] = -invM[10];
  memset(e->fwd, 0, (size_t)W * H * sizeof(V4));
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) {
    const V4 pt = e->raycast[x + y * W];
    float cx, cy, cz, ix, iy;
    m4_apply(e->pose_M, pt.x * p->voxel_size, pt.y * p->voxel_size, pt.z * p->voxel_size, 1.0f, &cx, &cy, &cz);
    ix = k[0] * cx / cz + k[2]; iy = k[1] * cy / cz + k[3];
    if ((ix < 0) || (ix > W - 1) || (iy < 0) || (iy > H - 1)) continue;
    if (ix != ix || iy != iy) continue; /* NaN: the reference's float -> int conversion yields a negative index */
    e->fwd[(int)(ix + 0.5f) + (int)(iy + 0.5f) * W] = pt;
  }
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) {
    const int id = x + y * W;
    const V4 f = e->fwd[id];
    const V2 mm = e->minmax[(int)floorf((float)x / MINMAX_SUB) + (int)floorf((float)y / MINMAX_SUB) * W];
    if ((f.w <= 0) && ((f.x == 0 && f.y == 0 && f.z == 0) || (e->depth[id] >= 0)) && (mm.x < mm.y)) e->fwd_missing[n++] = id;
  }
  e->n_fwd_missing = n;
  for (i = 0; i < n; ++i) {
    const int id = e->fwd_missing[i];
    y = id / W; x = id - y * W;
    cast_ray(e, x, y, e->minmax[(int)floorf((float)x / MINMAX_SUB) + (int)floorf((float)y / MINMAX_SUB) * W], invM, k, &e->fwd[id]);
  }
  for (y = 0; y < H; y++) for (x = 0; x < W; x++) {
    const int id = x + y * W;
    float nrm[3], angle = 0;
    unsigned char g = 0;
    if (e->fwd[id].w > 0.0f && normal_from_points(e->fwd, x, y, W, H, p->voxel_size, light, nrm, &angle)) g = (unsigned char)((0.8f * angle + 0.2f) * 255.0f);
    e->raycast_image[id * 4 + 0] = g; e->raycast_image[id * 4 + 1] = g; e->raycast_image[id * 4 + 2] = g; e->raycast_image[id * 4 + 3] = g;
  }
}

/* FindVisibleBlocks, ITMVisualisationEngine_CPU.cpp:40-77 (checkBlockVisibility<false> with the given camera) */
static int find_visible_blocks(const port_engine *e, const float *M, const float *k, int W, int H, int *ids) {
  port_engine cam = *e;  /* block_visible reads pose and intrinsics from the engine: a by-value view with the free camera */
  int i, n = 0;
  memcpy(cam.pose_M, M, 64);
  cam.p.fx = k[0]; cam.p.fy = k[1]; cam.p.cx = k[2]; cam.p.cy = k[3]; cam.p.width = W; cam.p.height = H;
  for (i = 0; i < e->n_entries; ++i)
    if (e->hash[i].ptr >= 0 && block_visible(&cam, &e->hash[i])) ids[n++] = i;
  return n;
}

/* computeSingleNormalFromSDF, DeviceAgnostic/ITMRepresentationAccess.h:225-337 */
static void normal_from_sdf(const port_engine *e, float px, float py, float pz, float *r) {
  const float flx = floorf(px), fly = floorf(py), flz = floorf(pz);
  const float cx = px - flx, cy = py - fly, cz = pz - flz;
  const float ncx = 1.0f - cx, ncy = 1.0f - cy, ncz = 1.0f - cz;
  const int x = (int)flx, y = (int)fly, z = (int)flz;
  BlockCache c;
  int f;
  float fr[4], bk[4], t[4], p1, p2, v1;
  c.bx = c.by = c.bz = 0x7fffffff; c.base = -1;
#define S(dx, dy, dz) ((float)voxel_sdf(e, x + (dx), y + (dy), z + (dz), &f, &c))
  fr[0] = S(0, 0, 0); fr[1] = S(1, 0, 0); fr[2] = S(0, 1, 0); fr[3] = S(1, 1, 0);
  bk[0] = S(0, 0, 1); bk[1] = S(1, 0, 1); bk[2] = S(0, 1, 1); bk[3] = S(1, 1, 1);
  p1 = fr[0] * ncy * ncz + fr[2] * cy * ncz + bk[0] * ncy * cz + bk[2] * cy * cz;
  t[0] = S(-1, 0, 0); t[1] = S(-1, 1, 0); t[2] = S(-1, 0, 1); t[3] = S(-1, 1, 1);
  p2 = t[0] * ncy * ncz + t[1] * cy * ncz + t[2] * ncy * cz + t[3] * cy * cz;
  v1 = p1 * cx + p2 * ncx;
  p1 = fr[1] * ncy * ncz + fr[3] * cy * ncz + bk[1] * ncy * cz + bk[3] * cy * cz;
  t[0] = S(2, 0, 0); t[1] = S(2, 1, 0); t[2] = S(2, 0, 1); t[3] = S(2, 1, 1);
  p2 = t[0] * ncy * ncz + t[1] * cy * ncz + t[2] * ncy * cz + t[3] * cy * cz;
  r[0] = (p1 * ncx + p2 * cx - v1) / 32767.0f;
  p1 = fr[0] * ncx * ncz + fr[1] * cx * ncz + bk[0] * ncx * cz + bk[1] * cx * cz;
  t[0] = S(0, -1, 0); t[1] = S(1, -1, 0); t[2] = S(0, -1, 1); t[3] = S(1, -1, 1);
  p2 = t[0] * ncx * ncz + t[1] * cx * ncz + t[2] * ncx * cz + t[3] * cx * cz;
  v1 = p1 * cy + p2 * ncy;
  p1 = fr[2] * ncx * ncz + fr[3] * cx * ncz + bk[2] * ncx * cz + bk[3] * cx * cz;
  t[0] = S(0, 2, 0); t[1] = S(1, 2, 0); t[2] = S(0, 2, 1); t[3] = S(1, 2, 1);
  p2 = t[0] * ncx * ncz + t[1] * cx * ncz + t[2] * ncx * cz + t[3] * cx * cz;
  r[1] = (p1 * ncy + p2 * cy - v1) / 32767.0f;
  p1 = fr[0] * ncx * ncy + fr[1] * cx * ncy + fr[2] * ncx * cy + fr[3] * cx * cy;
  t[0] = S(0, 0, -1); t[1] = S(1, 0, -1); t[2] = S(0, 1, -1); t[3] = S(1, 1, -1);
  p2 = t[0] * ncx * ncy + t[1] * cx * ncy + t[2] * ncx * cy + t[3] * cx * cy;
  v1 = p1 * cz + p2 * ncz;
  p1 = bk[0] * ncx * ncy + bk[1] * cx * ncy + bk[2] * ncx * cy + bk[3] * cx * cy;
  t[0] = S(0, 0, 2); t[1] = S(1, 0, 2); t[2] = S(0, 1, 2); t[3] = S(1, 1, 2);
  p2 = t[0] * ncx * ncy + t[1] * cx * ncy + t[2] * ncx * cy + t[3] * cx * cy;
  r[2] = (p1 * ncz + p2 * cz - v1) / 32767.0f;
#undef S
}

/* the free-view branch of ITMMainEngine::GetImage (ITMMainEngine.cpp:167-186): FindVisibleBlocks + CreateExpectedDepths +
 * RenderImage_common (ITMVisualisationEngine_CPU.cpp:191-240; processPixelGrey / processPixelNormal,
 * DeviceAgnostic/ITMVisualisationEngine.h:368-410).  renderType 0 grey, 2 colour from normal (ITMVoxel_s has no colour: 1 = grey) */
static void render_free_view(port_engine *e, int renderType, const float *M, const float *k, int W, int H) {
  float invM[16], light[3];
  int i;
  if (e->free_w != W || e->free_h != H) {
    free(e->free_visible_ids); free(e->free_minmax); free(e->free_raycast); free(e->free_image);
    e->free_visible_ids = (int *)calloc((size_t)e->n_entries, sizeof(int));
    e->free_minmax = (V2 *)calloc((size_t)W * H, sizeof(V2));
    e->free_raycast = (V4 *)calloc((size_t)W * H, sizeof(V4));
    e->free_image = (unsigned char *)calloc((size_t)W * H, 4);
    e->free_w = W; e->free_h = H;
  }
  e->free_n_visible = find_visible_blocks(e, M, k, W, H, e->free_visible_ids);
  expected_depths_for(e, M, k, W, H, e->free_visible_ids, e->free_n_visible, e->free_minmax);
  raycast_for(e, M, k, W, H, e->free_minmax, e->free_raycast);
  m4_inverse(M, invM);
  light[0] = -invM[8]; light[1] = -invM[9]; light[2] = -invM[10];
  for (i = 0; i < W * H; ++i) {
    const V4 pt = e->free_raycast[i];
    unsigned char *o = e->free_image + (size_t)i * 4;
    int found = pt.w > 0;
    float n[3], angle = 0;
    if (found) {
      float scale;
      normal_from_sdf(e, pt.x, pt.y, pt.z, n);
      scale = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      n[0] *= scale; n[1] *= scale; n[2] *= scale;
      angle = n[0] * light[0] + n[1] * light[1] + n[2] * light[2];
      if (!(angle > 0.0)) found = 0;
    }
    if (!found) { o[0] = o[1] = o[2] = o[3] = 0; }
    else if (renderType == 2) {  /* drawPixelNormal leaves the alpha byte as it was */
      o[0] = (unsigned char)((0.3f + (-n[0] + 1.0f) * 0.35f) * 255.0f);
      o[1] = (unsigned char)((0.3f + (-n[1] + 1.0f) * 0.35f) * 255.0f);
      o[2] = (unsigned char)((0.3f + (-n[2] + 1.0f) * 0.35f) * 255.0f);
    } else {
      const unsigned char g = (unsigned char)((0.8f * angle + 0.2f) * 255.0f);
      o[0] = o[1] = o[2] = o[3] = g;
    }
  }
}

/* The TRACKER_COLOR branch of ITMTrackingController::Prepare (ITMTrackingController.cpp:22-28): CreateExpectedDepths at
 * pose_rgb = trafo_rgb_to_depth.calib_inv * pose_d, then CreatePointCloud_common + RenderPointCloud
 * (ITMVisualisationEngine_CPU.cpp:242-264, 424-462) with invM = pose_d^-1 * trafo_rgb_to_depth.calib; the colour camera has the
 * depth camera's intrinsics here.  calib_inv as ITMExtrinsics::SetFrom builds it (Objects/ITMExtrinsics.h:32-42).  ITMVoxel_s
 * has no colour: VoxelColorReader<false> returns zeros.  Points land in e->points, colours in e->normals (the point cloud's
 * two images); returns noTotalPoints. */
static int render_point_cloud(port_engine *e, int skipPoints) {
  const port_params *p = &e->p;
  const int W = p->width, H = p->height;
  const float k[4] = {p->fx, p->fy, p->cx, p->cy};
  const float *T = e->trafo_rgb_to_depth;
  float Tinv[16], pose_rgb[16], invPose[16], invM[16], light[3];
  int r, c, x, y, n = 0;
  for (r = 0; r < 16; ++r) Tinv[r] = (r % 5 == 0) ? 1.0f : 0.0f;
  for (r = 0; r < 3; ++r) for (c = 0; c < 3; ++c) Tinv[r + 4 * c] = T[c + 4 * r];
  for (r = 0; r < 3; ++r) {
    float d = 0.0f;
    for (c = 0; c < 3; ++c) d -= T[c + 4 * r] * T[c + 4 * 3];
    Tinv[r + 4 * 3] = d;
  }
  m4_product(Tinv, e->pose_M, pose_rgb);
  expected_depths_for(e, pose_rgb, k, W, H, e->visible_ids, e->n_visible, e->minmax);
  m4_inverse(e->pose_M, invPose);
  m4_product(invPose, T, invM);
  raycast_with_inverse(e, invM, k, W, H, e->minmax, e->raycast);
  memcpy(e->pose_pc_M, e->pose_M, sizeof(e->pose_pc_M));
  light[0] = -invM[8]; light[1] = -invM[9]; light[2] = -invM[10];
  for (y = 0; y < H; ++y) for (x = 0; x < W; ++x) {
    const int i = x + y * W;
    const V4 pt = e->raycast[i];
    unsigned char *o = e->raycast_image + (size_t)i * 4;
    int found = pt.w > 0;
    if (found) {
      float nrm[3], scale, angle;
      normal_from_sdf(e, pt.x, pt.y, pt.z, nrm);
      scale = 1.0f / sqrtf(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      angle = nrm[0] * scale * light[0] + nrm[1] * scale * light[1] + nrm[2] * scale * light[2];
      if (!(angle > 0.0)) found = 0;
      else o[0] = o[1] = o[2] = o[3] = (unsigned char)((0.8f * angle + 0.2f) * 255.0f);
    }
    if (!found) { o[0] = o[1] = o[2] = o[3] = 0; }
    if (skipPoints && ((x % 2 == 0) || (y % 2 == 0))) found = 0;
    if (found) {
      V4 zero = {0.0f, 0.0f, 0.0f, 0.0f}, loc;
      e->normals[n] = zero;
      loc.x = pt.x * p->voxel_size; loc.y = pt.y * p->voxel_size; loc.z = pt.z * p->voxel_size; loc.w = 1.0f;
      e->points[n] = loc;
      n++;
    }
  }
  return n;
}

/* ITMLowLevelEngine_CPU's colour-tracker helpers (ITMLowLevelEngine_CPU.cpp:12-108; filterSubsample, filterSubsampleWithHoles
 * for Vector4f, gradientX / gradientY: DeviceAgnostic/ITMLowLevelEngine.h:7-124).  op: 0 CopyImage, 1 FilterSubsample,
 * 2 FilterSubsampleWithHoles(Vector4f), 3 GradientX, 4 GradientY; `out` is filled with prefillByte first (the gradient drivers
 * clear only the first w*h*sizeof(Vector3s) bytes of their Vector4s image).  Returns the number of output bytes. */
static long long low_level(int op, const void *in, int w, int h, void *out, int prefillByte) {
  const int w2 = w / 2, h2 = h / 2;
  int x, y, ch, j;
  if (op == 0) { memcpy(out, in, (size_t)w * h * 4); return (long long)w * h * 4; }
  if (op == 1) {
    const unsigned char *s = (const unsigned char *)in;
    unsigned char *d = (unsigned char *)out;
    for (y = 0; y < h2; ++y) for (x = 0; x < w2; ++x) for (ch = 0; ch < 4; ++ch) {
      const int a = s[((2 * x) + (2 * y) * w) * 4 + ch], b = s[((2 * x + 1) + (2 * y) * w) * 4 + ch];
      const int c = s[((2 * x) + (2 * y + 1) * w) * 4 + ch], e4 = s[((2 * x + 1) + (2 * y + 1) * w) * 4 + ch];
      d[(x + y * w2) * 4 + ch] = (unsigned char)((a + b + c + e4) / 4);
    }
    return (long long)w2 * h2 * 4;
  }
  if (op == 2) {
    const V4 *s = (const V4 *)in;
    V4 *d = (V4 *)out;
    for (y = 0; y < h2; ++y) for (x = 0; x < w2; ++x) {
      const V4 tap[4] = {s[(2 * x) + (2 * y) * w], s[(2 * x + 1) + (2 * y) * w], s[(2 * x) + (2 * y + 1) * w], s[(2 * x + 1) + (2 * y + 1) * w]};
      V4 acc = {0.0f, 0.0f, 0.0f, 0.0f};
      float good = 0.0f;
      for (j = 0; j < 4; ++j) if (tap[j].w >= 0) { acc.x += tap[j].x; acc.y += tap[j].y; acc.z += tap[j].z; acc.w += tap[j].w; good++; }
      if (good > 0) { acc.x /= good; acc.y /= good; acc.z /= good; acc.w /= good; } else acc.w = -1.0f;
      d[x + y * w2] = acc;
    }
    return (long long)w2 * h2 * 16;
  }
  {
    const unsigned char *s = (const unsigned char *)in;
    short *g = (short *)out;
    memset(out, prefillByte, (size_t)w * h * 8);
    memset(out, 0, (size_t)w * h * 6);
    for (y = 1; y < h - 1; ++y) for (x = 1; x < w - 1; ++x) {
      for (ch = 0; ch < 3; ++ch) {
        short d3[3];
        for (j = -1; j <= 1; ++j) {
          const int hi = (op == 3) ? s[((x + 1) + (y + j) * w) * 4 + ch] : s[((x + j) + (y + 1) * w) * 4 + ch];
          const int lo = (op == 3) ? s[((x - 1) + (y + j) * w) * 4 + ch] : s[((x + j) + (y - 1) * w) * 4 + ch];
          d3[j + 1] = (short)(hi - lo);
        }
        g[(x + y * w) * 4 + ch] = (short)((d3[0] + 2 * d3[1] + d3[2]) / 8);
      }
      g[(x + y * w) * 4 + 3] = (short)((2 * 255 + 2 * (2 * 255) + 2 * 255) / 8);
    }
    return (long long)w * h * 8;
  }
}

/* The marching-cubes case table (the classic Lorensen-Cline / Bourke triangulation the reference uses,
 * DeviceAgnostic/ITMMeshingEngine.h:9-151), one 64-bit word per case: nibble i = i-th edge index, 0xF terminates. */
static const unsigned long long MC_CASE[256] = {
    0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFF380ULL, 0xFFFFFFFFFFFFF910ULL, 0xFFFFFFFFFF189381ULL,
    0xFFFFFFFFFFFFFA21ULL, 0xFFFFFFFFFFA21380ULL, 0xFFFFFFFFFF920A29ULL, 0xFFFFFFF89A8A2382ULL,
    0xFFFFFFFFFFFFF2B3ULL, 0xFFFFFFFFFF0B82B0ULL, 0xFFFFFFFFFFB32091ULL, 0xFFFFFFFB89B912B1ULL,
    0xFFFFFFFFFF3AB1A3ULL, 0xFFFFFFFAB8A801A0ULL, 0xFFFFFFF9AB9B3093ULL, 0xFFFFFFFFFFB8AA89ULL,
    0xFFFFFFFFFFFFF874ULL, 0xFFFFFFFFFF437034ULL, 0xFFFFFFFFFF748910ULL, 0xFFFFFFF137174914ULL,
    0xFFFFFFFFFF748A21ULL, 0xFFFFFFFA21403743ULL, 0xFFFFFFF748209A29ULL, 0xFFFF4973727929A2ULL,
    0xFFFFFFFFFF2B3748ULL, 0xFFFFFFF40242B74BULL, 0xFFFFFFFB32748109ULL, 0xFFFF1292B9B49B74ULL,
    0xFFFFFFF487AB31A3ULL, 0xFFFF4B7401B41AB1ULL, 0xFFFF30BAB9B09874ULL, 0xFFFFFFFAB99B4B74ULL,
    0xFFFFFFFFFFFFF459ULL, 0xFFFFFFFFFF380459ULL, 0xFFFFFFFFFF051450ULL, 0xFFFFFFF513538458ULL,
    0xFFFFFFFFFF459A21ULL, 0xFFFFFFF594A21803ULL, 0xFFFFFFF204245A25ULL, 0xFFFF8434535235A2ULL,
    0xFFFFFFFFFFB32459ULL, 0xFFFFFFF594B802B0ULL, 0xFFFFFFFB32510450ULL, 0xFFFF584B82852512ULL,
    0xFFFFFFF45931AB3AULL, 0xFFFFAB81A8180594ULL, 0xFFFF30BAB5B05045ULL, 0xFFFFFFFB8AA85845ULL,
    0xFFFFFFFFFF975879ULL, 0xFFFFFFF375359039ULL, 0xFFFFFFF751710870ULL, 0xFFFFFFFFFF753351ULL,
    0xFFFFFFF21A759879ULL, 0xFFFF37503505921AULL, 0xFFFF25A758528208ULL, 0xFFFFFFF7533525A2ULL,
    0xFFFFFFF2B3987597ULL, 0xFFFFB72029279759ULL, 0xFFFF751871810B32ULL, 0xFFFFFFF51771B12BULL,
    0xFFFFB3A31A758859ULL, 0xF0ABA010B7905075ULL, 0xF07570805A30B0ABULL, 0xFFFFFFFFFF5B75ABULL,
    0xFFFFFFFFFFFFF56AULL, 0xFFFFFFFFFF6A5380ULL, 0xFFFFFFFFFF6A5109ULL, 0xFFFFFFF6A5891381ULL,
    0xFFFFFFFFFF162561ULL, 0xFFFFFFF803621561ULL, 0xFFFFFFF620609569ULL, 0xFFFF823625285895ULL,
    0xFFFFFFFFFF56AB32ULL, 0xFFFFFFF56A02B80BULL, 0xFFFFFFF6A5B32910ULL, 0xFFFFB892B92916A5ULL,
    0xFFFFFFF315356B36ULL, 0xFFFF6B51505B0B80ULL, 0xFFFF9505606306B3ULL, 0xFFFFFFF89BB96956ULL,
    0xFFFFFFFFFF8746A5ULL, 0xFFFFFFFA56374034ULL, 0xFFFFFFF7486A5091ULL, 0xFFFF49737179156AULL,
    0xFFFFFFF874156216ULL, 0xFFFF743403625521ULL, 0xFFFF620560509748ULL, 0xF962695923497937ULL,
    0xFFFFFFF56A4872B3ULL, 0xFFFFB720242746A5ULL, 0xFFFF6A5B32874910ULL, 0xF6A54B7B492B9129ULL,
    0xFFFF6B51535B3748ULL, 0xFB404B7B016B5B15ULL, 0xF74836B630560950ULL, 0xFFFF9B7974B96956ULL,
    0xFFFFFFFFFFA4694AULL, 0xFFFFFFF380A946A4ULL, 0xFFFFFFF04606A10AULL, 0xFFFFA16468618138ULL,
    0xFFFFFFF462421941ULL, 0xFFFF462942921803ULL, 0xFFFFFFFFFF624420ULL, 0xFFFFFFF624428238ULL,
    0xFFFFFFF32B46A94AULL, 0xFFFF6A4A94B82280ULL, 0xFFFFA164606102B3ULL, 0xF1B8B12184A16146ULL,
    0xFFFF36B319639469ULL, 0xF14641916B0181B8ULL, 0xFFFFFFF4600636B3ULL, 0xFFFFFFFFFF86B846ULL,
    0xFFFFFFFA98A876A7ULL, 0xFFFFA76A907A0370ULL, 0xFFFF0818717A176AULL, 0xFFFFFFF37117A76AULL,
    0xFFFF768981861621ULL, 0xF937390976192962ULL, 0xFFFFFFF206607087ULL, 0xFFFFFFFFFF276237ULL,
    0xFFFF76898A86AB32ULL, 0xF7A9A76790B72702ULL, 0xFB32A767A1871081ULL, 0xFFFF17616A71B12BULL,
    0xF63136B619768698ULL, 0xFFFFFFFFFF76B190ULL, 0xFFFF06B0B3607087ULL, 0xFFFFFFFFFFFFF6B7ULL,
    0xFFFFFFFFFFFFFB67ULL, 0xFFFFFFFFFF67B803ULL, 0xFFFFFFFFFF67B910ULL, 0xFFFFFFF67B138918ULL,
    0xFFFFFFFFFF7B621AULL, 0xFFFFFFF7B6803A21ULL, 0xFFFFFFF7B69A2092ULL, 0xFFFF89A38A3A27B6ULL,
    0xFFFFFFFFFF726327ULL, 0xFFFFFFF026067807ULL, 0xFFFFFFF910732672ULL, 0xFFFF678891681261ULL,
    0xFFFFFFF73171A67AULL, 0xFFFF801781A7167AULL, 0xFFFF7A69A0A70730ULL, 0xFFFFFFF9A88A7A67ULL,
    0xFFFFFFFFFF68B486ULL, 0xFFFFFFF640603B63ULL, 0xFFFFFFF109648B68ULL, 0xFFFF63B139369649ULL,
    0xFFFFFFF1A28B6486ULL, 0xFFFF640B60B03A21ULL, 0xFFFF9A2920B648B4ULL, 0xF36463B34923A39AULL,
    0xFFFFFFF264248328ULL, 0xFFFFFFFFFF264240ULL, 0xFFFF834642432091ULL, 0xFFFFFFF642241491ULL,
    0xFFFF1A6648168318ULL, 0xFFFFFFF40660A01AULL, 0xF39A9303A6834364ULL, 0xFFFFFFFFFF4A649AULL,
    0xFFFFFFFFFFB67594ULL, 0xFFFFFFF67B594380ULL, 0xFFFFFFFB67045105ULL, 0xFFFF51345343867BULL,
    0xFFFFFFFB6721A459ULL, 0xFFFF594380A217B6ULL, 0xFFFF204A24A45B67ULL, 0xF67B25A523453843ULL,
    0xFFFFFFF945267327ULL, 0xFFFF786260680459ULL, 0xFFFF045051673263ULL, 0xF851584812786826ULL,
    0xFFFF73167161A459ULL, 0xF459078701671A61ULL, 0xFA737A6A305A4A04ULL, 0xFFFFA84A458A7A67ULL,
    0xFFFFFFF98B9B6596ULL, 0xFFFF590650360B63ULL, 0xFFFFB65510B508B0ULL, 0xFFFFFFF1355363B6ULL,
    0xFFFF65B8B9B59A21ULL, 0xFA21965690B603B0ULL, 0xF52025A50865B58BULL, 0xFFFF35A3A25363B6ULL,
    0xFFFF283265825985ULL, 0xFFFFFFF260069659ULL, 0xF826283865081851ULL, 0xFFFFFFFFFF612651ULL,
    0xF698965683A61631ULL, 0xFFFF06505960A01AULL, 0xFFFFFFFFFFA65830ULL, 0xFFFFFFFFFFFFF65AULL,
    0xFFFFFFFFFFB57A5BULL, 0xFFFFFFF03857BA5BULL, 0xFFFFFFF091BA57B5ULL, 0xFFFF1381897BA57AULL,
    0xFFFFFFF15717B21BULL, 0xFFFFB27571721380ULL, 0xFFFF7B2209729579ULL, 0xF289823295B27257ULL,
    0xFFFFFFF573532A52ULL, 0xFFFF52A578258028ULL, 0xFFFF2A37353A5109ULL, 0xF25752A278129289ULL,
    0xFFFFFFFFFF573531ULL, 0xFFFFFFF571170780ULL, 0xFFFFFFF735539309ULL, 0xFFFFFFFFFF795789ULL,
    0xFFFFFFF8BA8A5485ULL, 0xFFFF03BBA50B5405ULL, 0xFFFF54ABA8A48910ULL, 0xF41314943B54A4BAULL,
    0xFFFF8548B2582152ULL, 0xFB151B2B543B0B40ULL, 0xF58B8545B2950520ULL, 0xFFFFFFFFFF3B2549ULL,
    0xFFFF483543253A52ULL, 0xFFFFFFF0244252A5ULL, 0xF910854583A532A3ULL, 0xFFFF2492914252A5ULL,
    0xFFFFFFF153358548ULL, 0xFFFFFFFFFF501540ULL, 0xFFFF530509358548ULL, 0xFFFFFFFFFFFFF549ULL,
    0xFFFFFFFBA9B947B4ULL, 0xFFFFBA97B9794380ULL, 0xFFFFB470414B1BA1ULL, 0xF4BAB474A1843413ULL,
    0xFFFF219B294B97B4ULL, 0xF3801B2B197B9479ULL, 0xFFFFFFF04224B47BULL, 0xFFFF42343824B47BULL,
    0xFFFF947732972A92ULL, 0xF70207872A4797A9ULL, 0xFA040A1A472A3A73ULL, 0xFFFFFFFFFF4782A1ULL,
    0xFFFFFFF317714194ULL, 0xFFFF178180714194ULL, 0xFFFFFFFFFF347304ULL, 0xFFFFFFFFFFFFF784ULL,
    0xFFFFFFFFFF8BA8A9ULL, 0xFFFFFFFA9BB93903ULL, 0xFFFFFFFBA88A0A10ULL, 0xFFFFFFFFFFA3BA13ULL,
    0xFFFFFFF8B99B1B21ULL, 0xFFFF9B2921B93903ULL, 0xFFFFFFFFFFB08B20ULL, 0xFFFFFFFFFFFFFB23ULL,
    0xFFFFFFF98AA82832ULL, 0xFFFFFFFFFF2902A9ULL, 0xFFFF8A1810A82832ULL, 0xFFFFFFFFFFFFF2A1ULL,
    0xFFFFFFFFFF819831ULL, 0xFFFFFFFFFFFFF190ULL, 0xFFFFFFFFFFFFF830ULL, 0xFFFFFFFFFFFFFFFFULL,
};

/* ITMMeshingEngine_CPU::MeshScene, ITMMeshingEngine_CPU.cpp:19-58 (findPointNeighbors / sdfInterp / buildVertList,
 * DeviceAgnostic/ITMMeshingEngine.h:153-232) */
static void mesh_scene(port_engine *e) {
  static const int CX[8] = {0, 1, 1, 0, 0, 1, 1, 0}, CY[8] = {0, 0, 1, 1, 0, 0, 1, 1}, CZ[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  static const int EA[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, EB[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
  const int noMax = e->p.n_local * 32;
  const float factor = e->p.voxel_size;
  int entry, x, y, z, n = 0;
  if (!e->mesh) e->mesh = (float *)malloc((size_t)noMax * 9 * sizeof(float));
  memset(e->mesh, 0, (size_t)noMax * 9 * sizeof(float));
  for (entry = 0; entry < e->n_entries; ++entry) {
    const HashEntry *h = &e->hash[entry];
    if (h->ptr < 0) continue;
    for (z = 0; z < BLOCK; z++) for (y = 0; y < BLOCK; y++) for (x = 0; x < BLOCK; x++) {
      const int gx = h->x * BLOCK + x, gy = h->y * BLOCK + y, gz = h->z * BLOCK + z;
      float pts[8][3], sdf[8], vert[12][3];
      unsigned long long tris;
      int kk, cube = 0, ok = 1, edges = 0;
      BlockCache c;
      c.bx = c.by = c.bz = 0x7fffffff; c.base = -1;
      for (kk = 0; kk < 8; ++kk) {
        int found;
        pts[kk][0] = (float)(gx + CX[kk]); pts[kk][1] = (float)(gy + CY[kk]); pts[kk][2] = (float)(gz + CZ[kk]);
        sdf[kk] = (float)voxel_sdf(e, gx + CX[kk], gy + CY[kk], gz + CZ[kk], &found, &c) / 32767.0f;
        if (!found || sdf[kk] == 1.0f) { ok = 0; break; }
      }
      if (!ok) continue;
      for (kk = 0; kk < 8; ++kk) if (sdf[kk] < 0) cube |= 1 << kk;
      tris = MC_CASE[cube];
      { unsigned long long t = tris; while ((t & 0xF) != 0xF) { edges |= 1 << (int)(t & 0xF); t >>= 4; } }
      if (edges == 0) continue;
      for (kk = 0; kk < 12; ++kk) {
        const float *p1 = pts[EA[kk]], *p2 = pts[EB[kk]];
        const float v1 = sdf[EA[kk]], v2 = sdf[EB[kk]];
        if (!(edges & (1 << kk))) continue;
        if (fabs(0.0f - v1) < 0.00001f) { vert[kk][0] = p1[0]; vert[kk][1] = p1[1]; vert[kk][2] = p1[2]; }
        else if (fabs(0.0f - v2) < 0.00001f) { vert[kk][0] = p2[0]; vert[kk][1] = p2[1]; vert[kk][2] = p2[2]; }
        else if (fabs(v1 - v2) < 0.00001f) { vert[kk][0] = p1[0]; vert[kk][1] = p1[1]; vert[kk][2] = p1[2]; }
        else {
          const float t = (0.0f - v1) / (v2 - v1);
          vert[kk][0] = p1[0] + t * (p2[0] - p1[0]); vert[kk][1] = p1[1] + t * (p2[1] - p1[1]); vert[kk][2] = p1[2] + t * (p2[2] - p1[2]);
        }
      }
      while ((tris & 0xF) != 0xF) {
        float *o = e->mesh + (size_t)n * 9;
        for (kk = 0; kk < 3; ++kk) {
          const int ed = (int)((tris >> (4 * kk)) & 0xF);
          o[kk * 3 + 0] = vert[ed][0] * factor; o[kk * 3 + 1] = vert[ed][1] * factor; o[kk * 3 + 2] = vert[ed][2] * factor;
        }
        if (n < noMax - 1) n++;
        tris >>= 12;
      }
    }
  }
  e->n_mesh = n;
}

static void scene_reset(port_engine *e) {
  const size_t nv = (size_t)e->p.n_local * BLOCK3;
  size_t i;
  for (i = 0; i < nv; ++i) { e->voxels[i].sdf = 32767; e->voxels[i].w_depth = 0; e->voxels[i].pad = 0; }
  for (i = 0; i < (size_t)e->p.n_local; ++i) e->vba_list[i] = (int)i;
  e->last_free_block = e->p.n_local - 1;
  for (i = 0; i < (size_t)e->n_entries; ++i) { memset(&e->hash[i], 0, sizeof(HashEntry)); e->hash[i].ptr = -2; }
  for (i = 0; i < (size_t)e->p.n_excess; ++i) e->excess_list[i] = (int)i;
  e->last_free_excess = e->p.n_excess - 1;
}

void port_default_params(port_params *p, int w, int h) {
  const float s = (float)w / 640.0f;
  memset(p, 0, sizeof(*p));
  p->width = w; p->height = h;
  p->fx = 580.0f * s; p->fy = 580.0f * s; p->cx = (float)w / 2.0f; p->cy = (float)h / 2.0f;
  p->voxel_size = 0.005f; p->mu = 0.02f; p->max_w = 100; p->vf_min = 0.35f; p->vf_max = 3.0f; p->stop_at_max_w = 0;
  p->calib_a = 1.0f / 1000.0f; p->calib_b = 0.0f;
  p->n_local = 0x10000; p->n_bucket = 0x100000; p->n_excess = 0x20000;
  p->n_levels = 5;
  p->regime[0] = ITER_BOTH; p->regime[1] = ITER_BOTH; p->regime[2] = ITER_ROTATION; p->regime[3] = ITER_ROTATION; p->regime[4] = ITER_ROTATION;
  p->no_icp_run_till_level = 0;
  p->icp_dist_thresh = 0.1f * 0.1f; p->icp_termination = 1e-3f;
}

port_engine *port_create(const port_params *pp) {
  port_engine *e = (port_engine *)calloc(1, sizeof(port_engine));
  const size_t P = (size_t)pp->width * pp->height;
  const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  int l, w, h;
  float fx, fy, cx, cy, stepT;
  e->p = *pp;
  e->n_entries = pp->n_bucket + pp->n_excess;
  e->voxels = (Voxel *)malloc((size_t)pp->n_local * BLOCK3 * sizeof(Voxel));
  e->hash = (HashEntry *)malloc((size_t)e->n_entries * sizeof(HashEntry));
  e->vba_list = (int *)malloc((size_t)pp->n_local * sizeof(int));
  e->excess_list = (int *)malloc((size_t)pp->n_excess * sizeof(int));
  /* the reference sizes this list SDF_LOCAL_BLOCK_NUM and would overrun it when more blocks are visible than the pool
   * holds; the restatement over-allocates instead of reproducing the overrun */
  e->visible_ids = (int *)calloc((size_t)e->n_entries, sizeof(int));
  e->visible_type = (unsigned char *)calloc((size_t)e->n_entries, 1);
  e->minmax = (V2 *)calloc(P, sizeof(V2));
  e->raycast = (V4 *)calloc(P, sizeof(V4));
  e->raycast_image = (unsigned char *)calloc(P, 4);
  e->points = (V4 *)calloc(P, sizeof(V4));
  e->normals = (V4 *)calloc(P, sizeof(V4));
  e->raw = (short *)calloc(P, sizeof(short));
  e->depth = (float *)calloc(P, sizeof(float));
  e->fwd = (V4 *)calloc(P, sizeof(V4));
  e->fwd_missing = (int *)calloc(P, sizeof(int));
  e->requires_full_rendering = 1;
  e->wicp = g_next_wicp; e->bilateral = g_next_bilateral;
  e->alloc_type = (unsigned char *)calloc((size_t)e->n_entries, 1);
  e->block_coords = (short *)calloc((size_t)e->n_entries * 4, sizeof(short));
  /* ITMRenderState constructor fills the range image with the frustum limits (Objects/ITMRenderState.h:60-72) */
  { size_t i; for (i = 0; i < P; ++i) { e->minmax[i].x = pp->vf_min; e->minmax[i].y = pp->vf_max; } }
  memcpy(e->pose_M, ident, sizeof(ident)); memcpy(e->pose_pc_M, ident, sizeof(ident));
  memcpy(e->trafo_rgb_to_depth, ident, sizeof(ident));
  e->age = -1;
  /* hierarchy: ITMDepthTracker constructor (ITMDepthTracker.cpp:11-36), intrinsics halving (:62-75) */
  w = pp->width; h = pp->height; fx = pp->fx; fy = pp->fy; cx = pp->cx; cy = pp->cy;
  for (l = 0; l < pp->n_levels; ++l) {
    e->level_w[l] = w; e->level_h[l] = h;
    e->level_intr[l][0] = fx; e->level_intr[l][1] = fy; e->level_intr[l][2] = cx; e->level_intr[l][3] = cy;
    e->level_depth[l] = (l == 0) ? e->depth : (float *)calloc((size_t)w * h + 1, sizeof(float));
    e->iters[l] = 2 + 2 * l;
    w /= 2; h /= 2; fx = fx * 0.5f; fy = fy * 0.5f; cx = cx * 0.5f; cy = cy * 0.5f;
  }
  stepT = pp->icp_dist_thresh / pp->n_levels;
  e->dist_thresh[pp->n_levels - 1] = pp->icp_dist_thresh;
  for (l = pp->n_levels - 2; l >= 0; --l) e->dist_thresh[l] = e->dist_thresh[l + 1] - stepT;
  scene_reset(e);
  return e;
}

void port_destroy(port_engine *e) {
  int l;
  if (!e) return;
  for (l = 1; l < e->p.n_levels; ++l) free(e->level_depth[l]);
  free(e->voxels); free(e->hash); free(e->vba_list); free(e->excess_list); free(e->visible_ids); free(e->visible_type);
  free(e->minmax); free(e->raycast); free(e->raycast_image); free(e->points); free(e->normals); free(e->raw); free(e->depth);
  free(e->alloc_type); free(e->block_coords);
  free(e->float_tmp); free(e->sigma); free(e->dnormal);
  if (e->level_sigma[0]) for (l = 1; l < e->p.n_levels; ++l) free(e->level_sigma[l]);
  free(e->fwd); free(e->fwd_missing); free(e->mesh); free(e->free_visible_ids); free(e->free_minmax); free(e->free_raycast); free(e->free_image);
  free(e);
}

void port_update_view(port_engine *e, const short *depth) {
  memcpy(e->raw, depth, (size_t)e->p.width * e->p.height * sizeof(short));
  view_convert(e);
  view_filters(e);
}
/* engines created from now on: TRACKER_WICP + settings.modelSensorNoise, optionally settings.useBilateralFilter */
void port_set_tracker_wicp(int on, int bilateral) { g_next_wicp = on; g_next_bilateral = bilateral; }
void port_wicp_prepare(port_engine *e) { view_pyramid(e); pyramid_of(e, e->level_sigma); }
int port_wicp_gandh(port_engine *e, int level, const float *approxInvPose, float *out44) {
  float f = 0, nabla[6] = {0, 0, 0, 0, 0, 0}, H[36];
  int n, i;
  memset(H, 0, sizeof(H));
  n = wicp_evaluate(e, level, approxInvPose, &f, nabla, H);
  out44[0] = (float)n; out44[1] = f;
  for (i = 0; i < 6; ++i) out44[2 + i] = nabla[i];
  for (i = 0; i < 36; ++i) out44[8 + i] = H[i];
  return n;
}
float *port_depth_uncertainty(port_engine *e) { return e->sigma; }
float *port_depth_normal(port_engine *e) { return (float *)e->dnormal; }
/* ITMTrackingController::Track, ITMTrackingController.cpp:11-16 */
void port_track(port_engine *e) {
  if (e->age != -1) { if (e->wicp) wicp_track(e); else icp_track(e); }
  e->requires_full_rendering = tracker_far_from_point_cloud(e) || !e->use_approximate_raycast;
}
void port_allocate(port_engine *e, int onlyVisible) { scene_allocate(e, onlyVisible); }
void port_integrate(port_engine *e) { scene_integrate(e); }
void port_expected_depths(port_engine *e) { render_expected_depths(e); }
void port_icp_maps(port_engine *e) { render_icp_maps(e); }
/* ITMMainEngine::ProcessFrame, ITMMainEngine.cpp:111-127 */
void port_prepare(port_engine *e);

void port_process_frame(port_engine *e, const short *depth) {
  port_update_view(e, depth);
  port_track(e);
  scene_allocate(e, 0);
  scene_integrate(e);
  port_prepare(e);
}

/* ITMTrackingController::Prepare, ITMTrackingController.cpp:18-46 (depth trackers) */
void port_prepare(port_engine *e) {
  render_expected_depths(e);
  if (e->requires_full_rendering) render_icp_maps(e);
  else { render_forward(e); e->age++; }
}
void port_set_use_approximate_raycast(port_engine *e, int on) { e->use_approximate_raycast = on != 0; }
int port_requires_full_rendering(port_engine *e) { return e->requires_full_rendering; }
void port_forward_render(port_engine *e) { render_forward(e); e->age++; }
float *port_forward_projection(port_engine *e) { return (float *)e->fwd; }
int *port_fwd_missing_points(port_engine *e) { return e->fwd_missing; }
int port_no_fwd_missing_points(port_engine *e) { return e->n_fwd_missing; }
int port_mesh_scene(port_engine *e, float **triangles, int *noMaxTriangles) {
  mesh_scene(e);
  *triangles = e->mesh; *noMaxTriangles = e->p.n_local * 32;
  return e->n_mesh;
}
/* free-view image types of ITMMainEngine::GetImage: 3 shaded, 4 colour from volume (= shaded for ITMVoxel_s), 5 colour from normal */
int port_get_image(port_engine *e, int type, const float *M, const float *k, int w, int h, unsigned char *out) {
  if (type < 3 || type > 5) return -2;
  render_free_view(e, type == 5 ? 2 : 0, M, k, w, h);
  memcpy(out, e->free_image, (size_t)w * h * 4);
  return 0;
}
int port_create_point_cloud(port_engine *e, const float *trafo16, int skipPoints) {
  if (trafo16) memcpy(e->trafo_rgb_to_depth, trafo16, 64);
  return render_point_cloud(e, skipPoints);
}
long long port_low_level(port_engine *e, int op, const void *in, int w, int h, void *out, int prefillByte) {
  (void)e;
  if (op != 3 && op != 4) memset(out, prefillByte, op == 0 ? (size_t)w * h * 4 : op == 1 ? (size_t)(w / 2) * (h / 2) * 4 : (size_t)(w / 2) * (h / 2) * 16);
  return low_level(op, in, w, h, out, prefillByte);
}
int *port_free_visible_ids(port_engine *e) { return e->free_visible_ids; }
int port_free_no_visible(port_engine *e) { return e->free_n_visible; }
float *port_free_minmax(port_engine *e) { return (float *)e->free_minmax; }
float *port_free_raycast_result(port_engine *e) { return (float *)e->free_raycast; }

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}
/* per-stage wall clock: view, track, allocate, integrate, expected depths, raycast + ICP maps */
void port_process_frame_timed(port_engine *e, const short *depth, double *ms6) {
  double t0 = now_ms(), t1, t2, t3, t4, t5, t6;
  port_update_view(e, depth); t1 = now_ms();
  port_track(e); t2 = now_ms();
  scene_allocate(e, 0); t3 = now_ms();
  scene_integrate(e); t4 = now_ms();
  render_expected_depths(e); t5 = now_ms();
  render_icp_maps(e); t6 = now_ms();
  ms6[0] = t1 - t0; ms6[1] = t2 - t1; ms6[2] = t3 - t2; ms6[3] = t4 - t3; ms6[4] = t5 - t4; ms6[5] = t6 - t5;
}

void port_icp_prepare(port_engine *e) { view_pyramid(e); }
int port_icp_gandh(port_engine *e, int level, const float *approxInvPose, float *out44) {
  float f = 0, nabla[6] = {0, 0, 0, 0, 0, 0}, H[36];
  int i, n;
  memset(H, 0, sizeof(H));
  n = icp_evaluate(e, level, approxInvPose, &f, nabla, H);
  out44[0] = (float)n; out44[1] = f;
  for (i = 0; i < 6; ++i) out44[2 + i] = nabla[i];
  for (i = 0; i < 36; ++i) out44[8 + i] = H[i];
  return n;
}
int port_pyramid_level(port_engine *e, int level, float **data, int *w, int *h, float *intr4) {
  *data = e->level_depth[level]; *w = e->level_w[level]; *h = e->level_h[level];
  memcpy(intr4, e->level_intr[level], 4 * sizeof(float));
  return 0;
}
void port_icp_config(port_engine *e, int *n, int *iters, float *thr, int *types) {
  int i;
  *n = e->p.n_levels;
  for (i = 0; i < e->p.n_levels; ++i) { iters[i] = e->iters[i]; thr[i] = e->dist_thresh[i]; types[i] = e->p.regime[i]; }
}

void port_get_pose(port_engine *e, float *M) { memcpy(M, e->pose_M, 64); }
void port_set_pose(port_engine *e, const float *M) { memcpy(e->pose_M, M, 64); se3_log(M, e->pose_params); } /* ITMPose::SetM */
void port_get_pose_pointcloud(port_engine *e, float *M) { memcpy(M, e->pose_pc_M, 64); }
void port_set_pose_pointcloud(port_engine *e, const float *M) { memcpy(e->pose_pc_M, M, 64); }
void port_get_pose_params(port_engine *e, float *p6) { memcpy(p6, e->pose_params, 24); }
int port_get_age(port_engine *e) { return e->age; }
void port_set_age(port_engine *e, int a) { e->age = a; }
void port_mat_inv(const float *in, float *out) { m4_inverse(in, out); }
void port_pose_from_invm_coerced(const float *inv, float *M, float *invOut, float *p6) {
  float tmp[16];
  m4_inverse(inv, tmp); se3_log(tmp, p6); se3_exp(p6, M); m4_inverse(M, invOut);
}
void port_pose_from_params(const float *p6, float *M) { se3_exp(p6, M); }
void port_compute_delta(port_engine *e, const float *nabla, const float *H, int shortIter, float *step) { (void)e; icp_delta(step, nabla, H, shortIter); }

void *port_hash_entries(port_engine *e) { return e->hash; }
void *port_voxels(port_engine *e) { return e->voxels; }
int *port_vba_alloc_list(port_engine *e) { return e->vba_list; }
int *port_excess_alloc_list(port_engine *e) { return e->excess_list; }
int *port_visible_ids(port_engine *e) { return e->visible_ids; }
unsigned char *port_visible_types(port_engine *e) { return e->visible_type; }
void port_get_counters(port_engine *e, int *c3) { c3[0] = e->n_visible; c3[1] = e->last_free_block; c3[2] = e->last_free_excess; }
void port_set_counters(port_engine *e, const int *c3) { e->n_visible = c3[0]; e->last_free_block = c3[1]; e->last_free_excess = c3[2]; }
float *port_depth(port_engine *e) { return e->depth; }
float *port_minmax(port_engine *e) { return (float *)e->minmax; }
float *port_raycast_result(port_engine *e) { return (float *)e->raycast; }
unsigned char *port_raycast_image(port_engine *e) { return e->raycast_image; }
float *port_points(port_engine *e) { return (float *)e->points; }
float *port_normals(port_engine *e) { return (float *)e->normals; }
