/*
 * oracle/se3_math.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
 *
 * Small fixed-size double-precision helpers that restate, on plain C arrays, the Eigen / g2o
 * rigid-body arithmetic the reference's optimisers rely on.  Parity status: UNPINNED -- the reference
 * ships no tests or golden vectors, and it cannot be compiled here (no Eigen,
 * OpenCV or CSparse).  Each function cites the reference lines it follows.
 *
 * Conventions: an "iso" is an Isometry3 stored as 12 doubles, R row-major (r[0..8]) followed by t (r[9..11]).
 */
#ifndef VDO_ORACLE_SE3_MATH_H
#define VDO_ORACLE_SE3_MATH_H

#include <math.h>
#include <string.h>

/* ---- 3x3 / 3-vector basics (row-major) ---- */
static inline void m3_mul(const double *a, const double *b, double *c) {
  double o[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(c, o, sizeof o);
}
static inline void m3_tr(const double *a, double *c) {
  double o[9] = {a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]};
  memcpy(c, o, sizeof o);
}
static inline void m3_vec(const double *a, const double *v, double *o) {
  double x = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
  double y = a[3] * v[0] + a[4] * v[1] + a[5] * v[2];
  double z = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void m3_tvec(const double *a, const double *v, double *o) { /* a^T v */
  double x = a[0] * v[0] + a[3] * v[1] + a[6] * v[2];
  double y = a[1] * v[0] + a[4] * v[1] + a[7] * v[2];
  double z = a[2] * v[0] + a[5] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}

/* ---- Isometry3 (Eigen::Transform<double,3,Isometry>) ---- */
static inline void iso_identity(double *T) {
  memset(T, 0, 12 * sizeof(double));
  T[0] = T[4] = T[8] = 1.0;
}
/* Eigen Isometry inverse: R^T, -R^T t */
static inline void iso_inv(const double *T, double *o) {
  double Rt[9], t[3];
  m3_tr(T, Rt);
  m3_vec(Rt, T + 9, t);
  memcpy(o, Rt, sizeof Rt);
  o[9] = -t[0]; o[10] = -t[1]; o[11] = -t[2];
}
static inline void iso_mul(const double *A, const double *B, double *o) {
  double R[9], t[3];
  m3_mul(A, B, R);
  m3_vec(A, B + 9, t);
  t[0] += A[9]; t[1] += A[10]; t[2] += A[11];
  memcpy(o, R, sizeof R);
  o[9] = t[0]; o[10] = t[1]; o[11] = t[2];
}
static inline void iso_apply(const double *T, const double *p, double *o) {
  double q[3];
  m3_vec(T, p, q);
  o[0] = q[0] + T[9]; o[1] = q[1] + T[10]; o[2] = q[2] + T[11];
}

/* ---- quaternion <-> rotation (Eigen's published algorithms; q = {x,y,z,w}) ---- */
/* Eigen::Quaternion(Matrix3): trace branch, else largest-diagonal branch. */
static inline void quat_from_m3(const double *R, double *q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}
/* Eigen::Quaternion::toRotationMatrix */
static inline void quat_to_m3(const double *q, double *R) {
  double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
static inline void quat_normalize(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

/* g2o::internal::toCompactQuaternion (isometry3d_mappings.cpp:78-83 with normalize :38-44) */
static inline void compact_quat_from_m3(const double *R, double *v) {
  double q[4];
  quat_from_m3(R, q);
  quat_normalize(q);
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
}
/* g2o::internal::fromCompactQuaternion (isometry3d_mappings.cpp:85-92) */
static inline void m3_from_compact_quat(const double *v, double *R) {
  double w = 1.0 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (w < 0) { R[0] = R[4] = R[8] = 1; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0; return; }
  double q[4] = {v[0], v[1], v[2], sqrt(w)};
  quat_to_m3(q, R);
}
/* toVectorMQT / fromVectorMQT (isometry3d_mappings.cpp:95-100, 118-123): [t, qx,qy,qz] */
static inline void iso_to_mqt(const double *T, double *v) {
  compact_quat_from_m3(T, v + 3);
  v[0] = T[9]; v[1] = T[10]; v[2] = T[11];
}
static inline void iso_from_mqt(const double *v, double *T) {
  m3_from_compact_quat(v + 3, T);
  T[9] = v[0]; T[10] = v[1]; T[11] = v[2];
}
/* VertexSE3::oplusImpl (vertex_se3.h:105-114): T <- T * fromVectorMQT(update) */
static inline void iso_oplus(double *T, const double *upd) {
  double inc[12];
  iso_from_mqt(upd, inc);
  iso_mul(T, inc, T);
}
/* approximateNearestOrthogonalMatrix (isometry3d_mappings.h:83-89): R -= 0.5 R (R^T R - I) */
static inline void m3_approx_orthogonalize(double *R) {
  double Rt[9], E[9], RE[9];
  m3_tr(R, Rt);
  m3_mul(Rt, R, E);
  E[0] -= 1; E[4] -= 1; E[8] -= 1;
  m3_mul(R, E, RE);
  for (int i = 0; i < 9; ++i) R[i] -= 0.5 * RE[i];
}
/* SE3Quat(R,t) then cast to Isometry3d (se3quat.h:58-60, 286-301; Converter.cc:25-35):
 * quaternion from R, flip to w>=0, normalise, back to a rotation matrix. */
static inline void iso_from_Rt_via_quat(const double *R, const double *t, double *T) {
  double q[4];
  quat_from_m3(R, q);
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  quat_normalize(q);
  quat_to_m3(q, T);
  T[9] = t[0]; T[10] = t[1]; T[11] = t[2];
}

#endif
