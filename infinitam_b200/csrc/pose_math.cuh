// SE(3) pose arithmetic used by the ICP tracker, usable on host and device.
//
// Semantics follow the reference's host code (operation order included):
//   mat4_inv            ORUtils/Matrix.h:162-218            (cofactor inverse)
//   mat4_mul            ORUtils/Matrix.h:102-107
//   pose_params_to_M    ITMLib/Objects/ITMPose.cpp:84-152   (SetModelViewFromParams, SE3 exp)
//   pose_M_to_params    ITMLib/Objects/ITMPose.cpp:154-234  (SetParamsFromModelView, SE3 log)
//   cholesky_solve      ORUtils/Cholesky.h:9-71
//   icp_compute_delta   ITMLib/Engine/ITMDepthTracker.cpp:85-102
//   icp_apply_delta     ITMLib/Engine/ITMDepthTracker.cpp:114-143
// Trigonometric functions are the only place where host libm and CUDA libm may
// differ by an ulp; everything else is IEEE fp32 with no contraction.
#pragma once
#include <math.h>

#define PM_HD __host__ __device__ inline

PM_HD bool mat4_inv(const float *m, float *dst) {
  float tmp[12], src[16], det;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    src[i] = m[i * 4];
    src[i + 4] = m[i * 4 + 1];
    src[i + 8] = m[i * 4 + 2];
    src[i + 12] = m[i * 4 + 3];
  }
  tmp[0] = src[10] * src[15];
  tmp[1] = src[11] * src[14];
  tmp[2] = src[9] * src[15];
  tmp[3] = src[11] * src[13];
  tmp[4] = src[9] * src[14];
  tmp[5] = src[10] * src[13];
  tmp[6] = src[8] * src[15];
  tmp[7] = src[11] * src[12];
  tmp[8] = src[8] * src[14];
  tmp[9] = src[10] * src[12];
  tmp[10] = src[8] * src[13];
  tmp[11] = src[9] * src[12];

  dst[0] = (tmp[0] * src[5] + tmp[3] * src[6] + tmp[4] * src[7]) - (tmp[1] * src[5] + tmp[2] * src[6] + tmp[5] * src[7]);
  dst[1] = (tmp[1] * src[4] + tmp[6] * src[6] + tmp[9] * src[7]) - (tmp[0] * src[4] + tmp[7] * src[6] + tmp[8] * src[7]);
  dst[2] = (tmp[2] * src[4] + tmp[7] * src[5] + tmp[10] * src[7]) - (tmp[3] * src[4] + tmp[6] * src[5] + tmp[11] * src[7]);
  dst[3] = (tmp[5] * src[4] + tmp[8] * src[5] + tmp[11] * src[6]) - (tmp[4] * src[4] + tmp[9] * src[5] + tmp[10] * src[6]);

  det = src[0] * dst[0] + src[1] * dst[1] + src[2] * dst[2] + src[3] * dst[3];
  if (det == 0.0f) return false;

  dst[4] = (tmp[1] * src[1] + tmp[2] * src[2] + tmp[5] * src[3]) - (tmp[0] * src[1] + tmp[3] * src[2] + tmp[4] * src[3]);
  dst[5] = (tmp[0] * src[0] + tmp[7] * src[2] + tmp[8] * src[3]) - (tmp[1] * src[0] + tmp[6] * src[2] + tmp[9] * src[3]);
  dst[6] = (tmp[3] * src[0] + tmp[6] * src[1] + tmp[11] * src[3]) - (tmp[2] * src[0] + tmp[7] * src[1] + tmp[10] * src[3]);
  dst[7] = (tmp[4] * src[0] + tmp[9] * src[1] + tmp[10] * src[2]) - (tmp[5] * src[0] + tmp[8] * src[1] + tmp[11] * src[2]);

  tmp[0] = src[2] * src[7];
  tmp[1] = src[3] * src[6];
  tmp[2] = src[1] * src[7];
  tmp[3] = src[3] * src[5];
  tmp[4] = src[1] * src[6];
  tmp[5] = src[2] * src[5];
  tmp[6] = src[0] * src[7];
  tmp[7] = src[3] * src[4];
  tmp[8] = src[0] * src[6];
  tmp[9] = src[2] * src[4];
  tmp[10] = src[0] * src[5];
  tmp[11] = src[1] * src[4];

  dst[8] = (tmp[0] * src[13] + tmp[3] * src[14] + tmp[4] * src[15]) - (tmp[1] * src[13] + tmp[2] * src[14] + tmp[5] * src[15]);
  dst[9] = (tmp[1] * src[12] + tmp[6] * src[14] + tmp[9] * src[15]) - (tmp[0] * src[12] + tmp[7] * src[14] + tmp[8] * src[15]);
  dst[10] = (tmp[2] * src[12] + tmp[7] * src[13] + tmp[10] * src[15]) - (tmp[3] * src[12] + tmp[6] * src[13] + tmp[11] * src[15]);
  dst[11] = (tmp[5] * src[12] + tmp[8] * src[13] + tmp[11] * src[14]) - (tmp[4] * src[12] + tmp[9] * src[13] + tmp[10] * src[14]);
  dst[12] = (tmp[2] * src[10] + tmp[5] * src[11] + tmp[1] * src[9]) - (tmp[4] * src[11] + tmp[0] * src[9] + tmp[3] * src[10]);
  dst[13] = (tmp[8] * src[11] + tmp[0] * src[8] + tmp[7] * src[10]) - (tmp[6] * src[10] + tmp[9] * src[11] + tmp[1] * src[8]);
  dst[14] = (tmp[6] * src[9] + tmp[11] * src[11] + tmp[3] * src[8]) - (tmp[10] * src[11] + tmp[2] * src[8] + tmp[7] * src[9]);
  dst[15] = (tmp[10] * src[10] + tmp[4] * src[8] + tmp[9] * src[9]) - (tmp[8] * src[9] + tmp[11] * src[10] + tmp[5] * src[8]);

  const float s = 1 / det;
#pragma unroll
  for (int i = 0; i < 16; ++i) dst[i] *= s;
  return true;
}

// mat4_inv for a matrix whose bottom row is exactly (0, 0, 0, 1) - every pose the tracker handles.  In the cofactor
// formula above the products with that row are x * 0 (= +-0, dropped here: adding a zero changes nothing) and x * 1
// (= x), so what remains is the same arithmetic on the same operands in the same order: bit-identical results for
// finite inputs (only the sign of a zero may differ) at well under half the instructions - this runs on one thread in
// the middle of the tracker's serial chain.  Falls back to the general routine for any other bottom row.
PM_HD bool mat4_inv_pose(const float *m, float *dst) {
  if (!(m[3] == 0.0f && m[7] == 0.0f && m[11] == 0.0f && m[15] == 1.0f)) return mat4_inv(m, dst);
  // rows of m (the general routine's transposed copy): a = src[0..3], b = src[4..7], c = src[8..11]
  const float a0 = m[0], a1 = m[4], a2 = m[8], a3 = m[12];
  const float b0 = m[1], b1 = m[5], b2 = m[9], b3 = m[13];
  const float c0 = m[2], c1 = m[6], c2 = m[10], c3 = m[14];
  dst[0] = c2 * b1 - c1 * b2;
  dst[1] = c0 * b2 - c2 * b0;
  dst[2] = c1 * b0 - c0 * b1;
  dst[3] = 0.0f;
  const float det = a0 * dst[0] + a1 * dst[1] + a2 * dst[2];
  if (det == 0.0f) return false;
  dst[4] = c1 * a2 - c2 * a1;
  dst[5] = c2 * a0 - c0 * a2;
  dst[6] = c0 * a1 - c1 * a0;
  dst[7] = 0.0f;
  const float t0 = a2 * b3, t1 = a3 * b2, t2 = a1 * b3, t3 = a3 * b1, t4 = a1 * b2, t5 = a2 * b1;
  const float t6 = a0 * b3, t7 = a3 * b0, t8 = a0 * b2, t9 = a2 * b0, t10 = a0 * b1, t11 = a1 * b0;
  dst[8] = t4 - t5;
  dst[9] = t9 - t8;
  dst[10] = t10 - t11;
  dst[11] = 0.0f;
  dst[12] = (t2 * c2 + t5 * c3 + t1 * c1) - (t4 * c3 + t0 * c1 + t3 * c2);
  dst[13] = (t8 * c3 + t0 * c0 + t7 * c2) - (t6 * c2 + t9 * c3 + t1 * c0);
  dst[14] = (t6 * c1 + t11 * c3 + t3 * c0) - (t10 * c3 + t2 * c0 + t7 * c1);
  dst[15] = (t10 * c2 + t4 * c0 + t9 * c1) - (t8 * c1 + t11 * c2 + t5 * c0);
  const float s = 1 / det;
#pragma unroll
  for (int i = 0; i < 16; ++i) dst[i] *= s;
  return true;
}

// r = lhs * rhs, column-major; each element accumulated from 0 in k order like the reference
PM_HD void mat4_mul(const float *lhs, const float *rhs, float *r) {
#pragma unroll
  for (int x = 0; x < 4; x++)
#pragma unroll
    for (int y = 0; y < 4; y++) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; k++) acc += lhs[y + 4 * k] * rhs[k + 4 * x];
      r[y + 4 * x] = acc;
    }
}

PM_HD void pm_cross(const float *a, const float *b, float *r) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
PM_HD float pm_dot(const float *a, const float *b) {
  float r = 0;
  r += a[0] * b[0];
  r += a[1] * b[1];
  r += a[2] * b[2];
  return r;
}

// params = tx ty tz rx ry rz ; M column-major
PM_HD void pose_params_to_M(const float *p, float *M) {
  const float one_6th = 1.0f / 6.0f;
  const float one_20th = 1.0f / 20.0f;
  float w[3] = {p[3], p[4], p[5]};
  float t[3] = {p[0], p[1], p[2]};
  float theta_sq = pm_dot(w, w);
  float theta = sqrtf(theta_sq);
  float A, B;
  float R[9], T[3];
  float crossV[3];
  pm_cross(w, t, crossV);
  if (theta_sq < 1e-8f) {
    A = 1.0f - one_6th * theta_sq;
    B = 0.5f;
    T[0] = t[0] + 0.5f * crossV[0];
    T[1] = t[1] + 0.5f * crossV[1];
    T[2] = t[2] + 0.5f * crossV[2];
  } else {
    float C;
    if (theta_sq < 1e-6f) {
      C = one_6th * (1.0f - one_20th * theta_sq);
      A = 1.0f - theta_sq * C;
      B = 0.5f - 0.25f * one_6th * theta_sq;
    } else {
      float inv_theta = 1.0f / theta;
      float sn, cs;
#ifdef __CUDA_ARCH__
      sincosf(theta, &sn, &cs);  // one argument reduction for both (same values as sinf / cosf)
#else
      sn = sinf(theta); cs = cosf(theta);
#endif
      A = sn * inv_theta;
      B = (1.0f - cs) * (inv_theta * inv_theta);
      C = (1.0f - A) * (inv_theta * inv_theta);
    }
    float cross2[3];
    pm_cross(w, crossV, cross2);
    T[0] = t[0] + B * crossV[0] + C * cross2[0];
    T[1] = t[1] + B * crossV[1] + C * cross2[1];
    T[2] = t[2] + B * crossV[2] + C * cross2[2];
  }
  float wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
  R[0 + 3 * 0] = 1.0f - B * (wy2 + wz2);
  R[1 + 3 * 1] = 1.0f - B * (wx2 + wz2);
  R[2 + 3 * 2] = 1.0f - B * (wx2 + wy2);
  float a, b;
  a = A * w[2], b = B * (w[0] * w[1]);
  R[0 + 3 * 1] = b - a;
  R[1 + 3 * 0] = b + a;
  a = A * w[1], b = B * (w[0] * w[2]);
  R[0 + 3 * 2] = b + a;
  R[2 + 3 * 0] = b - a;
  a = A * w[0], b = B * (w[1] * w[2]);
  R[1 + 3 * 2] = b - a;
  R[2 + 3 * 1] = b + a;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) M[r + 4 * c] = R[r + 3 * c];
  M[0 + 4 * 3] = T[0];
  M[1 + 4 * 3] = T[1];
  M[2 + 4 * 3] = T[2];
  M[3 + 4 * 0] = 0.0f;
  M[3 + 4 * 1] = 0.0f;
  M[3 + 4 * 2] = 0.0f;
  M[3 + 4 * 3] = 1.0f;
}

PM_HD void pose_M_to_params(const float *M, float *p) {
  float R[9], T[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) R[r + 3 * c] = M[r + 4 * c];
  T[0] = M[12];
  T[1] = M[13];
  T[2] = M[14];
  float rot[3];
  // Matrix3 member names are column-major too: m00 = m[0], m11 = m[4], m22 = m[8]
  float cos_angle = (R[0] + R[4] + R[8] - 1.0f) * 0.5f;
  rot[0] = (R[2 + 3 * 1] - R[1 + 3 * 2]) * 0.5f;
  rot[1] = (R[0 + 3 * 2] - R[2 + 3 * 0]) * 0.5f;
  rot[2] = (R[1 + 3 * 0] - R[0 + 3 * 1]) * 0.5f;
  float sin_angle_abs = sqrtf(pm_dot(rot, rot));
  const double SQRT1_2 = 0.707106781186547524401;
  if ((double)cos_angle > SQRT1_2) {
    if (sin_angle_abs) {
      float s = asinf(sin_angle_abs) / sin_angle_abs;
      rot[0] *= s; rot[1] *= s; rot[2] *= s;
    }
  } else {
    if ((double)cos_angle > -SQRT1_2) {
      float s = acosf(cos_angle) / sin_angle_abs;
      rot[0] *= s; rot[1] *= s; rot[2] *= s;
    } else {
      float angle = (float)3.14159265358979323846 - asinf(sin_angle_abs);
      float d0 = R[0 + 3 * 0] - cos_angle;
      float d1 = R[1 + 3 * 1] - cos_angle;
      float d2 = R[2 + 3 * 2] - cos_angle;
      float r2[3];
      if (fabsf(d0) > fabsf(d1) && fabsf(d0) > fabsf(d2)) {
        r2[0] = d0;
        r2[1] = (R[1 + 3 * 0] + R[0 + 3 * 1]) * 0.5f;
        r2[2] = (R[0 + 3 * 2] + R[2 + 3 * 0]) * 0.5f;
      } else if (fabsf(d1) > fabsf(d2)) {
        r2[0] = (R[1 + 3 * 0] + R[0 + 3 * 1]) * 0.5f;
        r2[1] = d1;
        r2[2] = (R[2 + 3 * 1] + R[1 + 3 * 2]) * 0.5f;
      } else {
        r2[0] = (R[0 + 3 * 2] + R[2 + 3 * 0]) * 0.5f;
        r2[1] = (R[2 + 3 * 1] + R[1 + 3 * 2]) * 0.5f;
        r2[2] = d2;
      }
      if (pm_dot(r2, rot) < 0.0f) { r2[0] *= -1.0f; r2[1] *= -1.0f; r2[2] *= -1.0f; }
      float len = sqrtf(pm_dot(r2, r2));
      if (len == 0) { r2[0] = r2[1] = r2[2] = 0.0f; } else { r2[0] /= len; r2[1] /= len; r2[2] /= len; }
      rot[0] = angle * r2[0]; rot[1] = angle * r2[1]; rot[2] = angle * r2[2];
    }
  }
  float shtot = 0.5f;
  float theta = sqrtf(pm_dot(rot, rot));
  if (theta > 0.00001f) shtot = sinf(theta * 0.5f) / theta;

  float hp[6] = {0.0f, 0.0f, 0.0f, rot[0] * -0.5f, rot[1] * -0.5f, rot[2] * -0.5f};
  float Mh[16];
  pose_params_to_M(hp, Mh);
  // halfrotor.GetR() * T   (Matrix3 * Vector3, ORUtils/Matrix.h)
  float rt[3];
  rt[0] = Mh[0] * T[0] + Mh[4] * T[1] + Mh[8] * T[2];
  rt[1] = Mh[1] * T[0] + Mh[5] * T[1] + Mh[9] * T[2];
  rt[2] = Mh[2] * T[0] + Mh[6] * T[1] + Mh[10] * T[2];
  if (theta > 0.001f) {
    float denom = pm_dot(rot, rot);
    float param = pm_dot(T, rot) * (1 - 2 * shtot) / denom;
    rt[0] -= rot[0] * param; rt[1] -= rot[1] * param; rt[2] -= rot[2] * param;
  } else {
    float param = pm_dot(T, rot) / 24;
    rt[0] -= rot[0] * param; rt[1] -= rot[1] * param; rt[2] -= rot[2] * param;
  }
  rt[0] /= 2 * shtot; rt[1] /= 2 * shtot; rt[2] /= 2 * shtot;
  p[3] = rot[0]; p[4] = rot[1]; p[5] = rot[2];
  p[0] = rt[0]; p[1] = rt[1]; p[2] = rt[2];
}

// pose_d->SetInvM(invM); pose_d->Coerce();  (ITMPose.cpp:316-326)
PM_HD void pose_set_invM_coerce(const float *invM, float *M, float *params) {
  float Mtmp[16];
  mat4_inv_pose(invM, Mtmp);
  // SetInvM -> SetParamsFromModelView, then Coerce: SetParamsFromModelView once more on the same M (a pure
  // function of M, so one evaluation gives the identical params), then SetModelViewFromParams
  pose_M_to_params(Mtmp, params);
  pose_params_to_M(params, M);
}

// ORUtils::Cholesky constructor + Backsub for size N (3 or 6), row/col conventions as in the reference.
// Templated on the size so that every index is a compile-time constant after unrolling (the arrays then
// live in registers when this runs on the device).
template <int N>
PM_HD void cholesky_solve(const float *mat, const float *v, float *result) {
  float ch[N * N];
#pragma unroll
  for (int i = 0; i < N * N; i++) ch[i] = mat[i];
#pragma unroll
  for (int c = 0; c < N; c++) {
    float inv_diag = 1;
#pragma unroll
    for (int r = c; r < N; r++) {
      float val = ch[c + r * N];
#pragma unroll
      for (int c2 = 0; c2 < c; c2++) val -= ch[c + c2 * N] * ch[c2 + r * N];
      if (r == c) {
        ch[c + r * N] = val;
        inv_diag = 1.0f / val;
      } else {
        ch[r + c * N] = val;
        ch[c + r * N] = val * inv_diag;
      }
    }
  }
  float y[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    float val = v[i];
#pragma unroll
    for (int j = 0; j < i; j++) val -= ch[j + i * N] * y[j];
    y[i] = val;
  }
#pragma unroll
  for (int i = 0; i < N; i++) y[i] /= ch[i + i * N];
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    float val = y[i];
#pragma unroll
    for (int j = i + 1; j < N; j++) val -= ch[i + j * N] * result[j];
    result[i] = val;
  }
}

PM_HD void icp_compute_delta(float *step, const float *nabla, const float *hessian, bool shortIteration) {
#pragma unroll
  for (int i = 0; i < 6; i++) step[i] = 0;
  if (shortIteration) {
    float small[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) small[r + c * 3] = hessian[r + c * 6];
    cholesky_solve<3>(small, nabla, step);
  } else {
    cholesky_solve<6>(hessian, nabla, step);
  }
}

PM_HD void icp_apply_delta(const float *para_old, const float *delta, int iterationType, float *para_new) {
  float step[6];
  if (iterationType == 1) {  // rotation
    step[0] = delta[0]; step[1] = delta[1]; step[2] = delta[2];
    step[3] = 0.0f; step[4] = 0.0f; step[5] = 0.0f;
  } else if (iterationType == 2) {  // translation
    step[0] = 0.0f; step[1] = 0.0f; step[2] = 0.0f;
    step[3] = delta[0]; step[4] = delta[1]; step[5] = delta[2];
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) step[i] = delta[i];
  }
  float Tinc[16];
  // Matrix4 member mCR lives at m[R + 4*C]  (ORUtils/Matrix.h:24-29)
  Tinc[0] = 1.0f;      Tinc[4] = step[2];   Tinc[8] = -step[1];  Tinc[12] = step[3];
  Tinc[1] = -step[2];  Tinc[5] = 1.0f;      Tinc[9] = step[0];   Tinc[13] = step[4];
  Tinc[2] = step[1];   Tinc[6] = -step[0];  Tinc[10] = 1.0f;     Tinc[14] = step[5];
  Tinc[3] = 0.0f;      Tinc[7] = 0.0f;      Tinc[11] = 0.0f;     Tinc[15] = 1.0f;
  float r[16];
  mat4_mul(Tinc, para_old, r);
#pragma unroll
  for (int i = 0; i < 16; ++i) para_new[i] = r[i];
}

PM_HD bool icp_has_converged(const float *step, float terminationThreshold) {
  float stepLength = 0.0f;
#pragma unroll
  for (int i = 0; i < 6; i++) stepLength += step[i] * step[i];
  return sqrtf(stepLength) / 6 < terminationThreshold;
}
