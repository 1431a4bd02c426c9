// ba_math.cuh -- per-edge / per-vertex arithmetic of the batch factor-graph path (fp64).
//
// Every function is VDO_HD so that the CUDA kernels (ba_kernels.cu) and the serial kernel emulation used by the
// CPU-only host-logic tests (tests/emul/) execute the same arithmetic.  Reference semantics:
//   EdgeSE3PointXYZ            g2o/types/edge_se3_pointxyz.cpp:99-140
//   LandmarkMotionTernaryEdge  g2o/types/types_dyn_slam3d.cpp:53-85
//   EdgeSE3 / EdgeSE3Prior     g2o/types/edge_se3.cpp:77-104, edge_se3_prior.cpp:89-102,
//                              isometry3d_gradients.h:191-325, dquat2mat.cpp:35-84
//   VertexSE3::oplusImpl       g2o/types/vertex_se3.h:105-114, isometry3d_mappings.cpp:78-123
//   RobustKernelHuber          g2o/core/robust_kernel_impl.cpp:65-91 (dsqr kept in float, robust_kernel_impl.h:84)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VDO_HD __host__ __device__ __forceinline__
#else
#define VDO_HD inline
#endif

namespace vdo {

struct Iso {            // world <- local: x_w = R x_l + t ; R row-major
  double R[9];
  double t[3];
};

VDO_HD void iso_load(const double* __restrict__ p, Iso& T) {
#pragma unroll
  for (int i = 0; i < 9; ++i) T.R[i] = p[i];
  T.t[0] = p[9]; T.t[1] = p[10]; T.t[2] = p[11];
}
VDO_HD void iso_store(double* p, const Iso& T) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = T.R[i];
  p[9] = T.t[0]; p[10] = T.t[1]; p[11] = T.t[2];
}
// R^T (p - t)
VDO_HD void iso_inv_apply(const Iso& T, const double* p, double* o) {
  double dx = p[0] - T.t[0], dy = p[1] - T.t[1], dz = p[2] - T.t[2];
  o[0] = T.R[0] * dx + T.R[3] * dy + T.R[6] * dz;
  o[1] = T.R[1] * dx + T.R[4] * dy + T.R[7] * dz;
  o[2] = T.R[2] * dx + T.R[5] * dy + T.R[8] * dz;
}
VDO_HD void rot_apply(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
VDO_HD void rot_t_apply(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
VDO_HD void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
VDO_HD void m3_mul(const double* a, const double* b, double* c) {
  double o[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) c[i] = o[i];
}
VDO_HD void iso_inv(const Iso& T, Iso& o) {
  double Rt[9] = {T.R[0], T.R[3], T.R[6], T.R[1], T.R[4], T.R[7], T.R[2], T.R[5], T.R[8]};
  double t[3];
  rot_apply(Rt, T.t, t);
#pragma unroll
  for (int i = 0; i < 9; ++i) o.R[i] = Rt[i];
  o.t[0] = -t[0]; o.t[1] = -t[1]; o.t[2] = -t[2];
}
VDO_HD void iso_mul(const Iso& A, const Iso& B, Iso& o) {
  double R[9], t[3];
  m3_mul(A.R, B.R, R);
  rot_apply(A.R, B.t, t);
#pragma unroll
  for (int i = 0; i < 9; ++i) o.R[i] = R[i];
  o.t[0] = t[0] + A.t[0]; o.t[1] = t[1] + A.t[1]; o.t[2] = t[2] + A.t[2];
}

// Huber weight rho'(e2) and cost rho(e2); delta <= 0 means "no kernel".
VDO_HD void huber(double e2, double delta, double& rho, double& w) {
  if (delta > 0.0) {
    double dsqr = (double)(float)(delta * delta);
    if (e2 > dsqr) {
      double s = sqrt(e2);
      rho = 2.0 * s * delta - dsqr;
      w = delta / s;
      return;
    }
  }
  rho = e2; w = 1.0;
}

// ---- quaternion helpers (Eigen's published conversions; q = {x,y,z,w}) ----
VDO_HD void quat_from_rot(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    double qq[4];
    qq[i] = 0.5 * t;
    t = 0.5 / t;
    qq[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    qq[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    qq[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
  }
}
VDO_HD void rot_from_quat(const double* q, double* R) {
  double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// toVectorMQT: [t, normalised quaternion xyz with w >= 0]
VDO_HD void iso_to_mqt(const Iso& T, double* v) {
  double q[4];
  quat_from_rot(T.R, q);
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double s = (q[3] < 0 ? -1.0 : 1.0) / n;
  v[0] = T.t[0]; v[1] = T.t[1]; v[2] = T.t[2];
  v[3] = q[0] * s; v[4] = q[1] * s; v[5] = q[2] * s;
}
// VertexSE3::oplusImpl: T <- T * fromVectorMQT(u)
VDO_HD void iso_oplus(Iso& T, const double* u) {
  Iso inc;
  double w = 1.0 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
  if (w < 0) {
    inc.R[0] = inc.R[4] = inc.R[8] = 1; inc.R[1] = inc.R[2] = inc.R[3] = inc.R[5] = inc.R[6] = inc.R[7] = 0;
  } else {
    double q[4] = {u[3], u[4], u[5], sqrt(w)};
    rot_from_quat(q, inc.R);
  }
  inc.t[0] = u[0]; inc.t[1] = u[1]; inc.t[2] = u[2];
  iso_mul(T, inc, T);
}
// approximateNearestOrthogonalMatrix (isometry3d_mappings.h:83-89)
VDO_HD void rot_reorthogonalize(double* R) {
  double Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]}, E[9], RE[9];
  m3_mul(Rt, R, E);
  E[0] -= 1; E[4] -= 1; E[8] -= 1;
  m3_mul(R, E, RE);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] -= 0.5 * RE[i];
}

// ---- d(quaternion xyz)/dR, 3x9, columns = column-major flattening of R (dquat2mat.cpp:35-84) ----
VDO_HD void dq_dR(const double* R, double* D) {
  const double r00 = R[0], r01 = R[1], r02 = R[2], r10 = R[3], r11 = R[4], r12 = R[5], r20 = R[6], r21 = R[7], r22 = R[8];
  for (int i = 0; i < 27; ++i) D[i] = 0.0;
  double tr = r00 + r11 + r22, qw;
  if (tr > 0) {
    double a = 0.5 * sqrt(tr + 1.0);
    qw = a;
    double i1 = 1.0 / a, i3 = i1 * i1 * i1;
    double gx = -0.03125 * (r21 - r12) * i3, gy = 0.03125 * (r20 - r02) * i3, gz = -0.03125 * (r10 - r01) * i3;
    D[0] = gx; D[4] = gx; D[8] = gx; D[5] = 0.25 * i1; D[7] = -0.25 * i1;
    D[9] = gy; D[13] = gy; D[17] = gy; D[11] = -0.25 * i1; D[15] = 0.25 * i1;
    D[18] = gz; D[22] = gz; D[26] = gz; D[19] = 0.25 * i1; D[21] = -0.25 * i1;
  } else if ((r00 > r11) & (r00 > r22)) {
    double S = sqrt(1.0 + r00 - r11 - r22) * 2, a = 0.25 * S;
    qw = (r21 - r12) / S;
    double i1 = 1.0 / a, i3 = i1 * i1 * i1, s1 = r10 + r01, s2 = r20 + r02;
    D[0] = 0.125 * i1; D[4] = -0.125 * i1; D[8] = -0.125 * i1;
    D[9] = -0.03125 * i3 * s1; D[13] = 0.03125 * i3 * s1; D[17] = 0.03125 * i3 * s1; D[10] = 0.25 * i1; D[12] = 0.25 * i1;
    D[18] = -0.03125 * i3 * s2; D[22] = 0.03125 * i3 * s2; D[26] = 0.03125 * i3 * s2; D[20] = 0.25 * i1; D[24] = 0.25 * i1;
  } else if (r11 > r22) {
    double S = sqrt(1.0 + r11 - r00 - r22) * 2, a = 0.25 * S;
    qw = (r02 - r20) / S;
    double i1 = 1.0 / a, i3 = i1 * i1 * i1, s1 = r10 + r01, s2 = r21 + r12;
    D[0] = 0.03125 * i3 * s1; D[4] = -0.03125 * i3 * s1; D[8] = 0.03125 * i3 * s1; D[1] = 0.25 * i1; D[3] = 0.25 * i1;
    D[9] = -0.125 * i1; D[13] = 0.125 * i1; D[17] = -0.125 * i1;
    D[18] = 0.03125 * i3 * s2; D[22] = -0.03125 * i3 * s2; D[26] = 0.03125 * i3 * s2; D[23] = 0.25 * i1; D[25] = 0.25 * i1;
  } else {
    double S = sqrt(1.0 + r22 - r00 - r11) * 2, a = 0.25 * S;
    qw = (r10 - r01) / S;
    double i1 = 1.0 / a, i3 = i1 * i1 * i1, s1 = r20 + r02, s2 = r21 + r12;
    D[0] = 0.03125 * i3 * s1; D[4] = 0.03125 * i3 * s1; D[8] = -0.03125 * i3 * s1; D[2] = 0.25 * i1; D[6] = 0.25 * i1;
    D[9] = 0.03125 * i3 * s2; D[13] = 0.03125 * i3 * s2; D[17] = -0.03125 * i3 * s2; D[14] = 0.25 * i1; D[16] = 0.25 * i1;
    D[18] = -0.125 * i1; D[22] = -0.125 * i1; D[26] = 0.125 * i1;
  }
  if (qw <= 0)
    for (int i = 0; i < 27; ++i) D[i] = -D[i];
}
// the three "skew" generator matrices of isometry3d_gradients.h:57-84 applied to R (sgn=+1: skew, -1: skewT)
VDO_HD void skew3(const double* R, double sgn, double* Sx, double* Sy, double* Sz) {
  double r[9];
  for (int i = 0; i < 9; ++i) r[i] = 2 * sgn * R[i];
  Sx[0] = Sx[1] = Sx[2] = 0; Sx[3] = -r[6]; Sx[4] = -r[7]; Sx[5] = -r[8]; Sx[6] = r[3]; Sx[7] = r[4]; Sx[8] = r[5];
  Sy[0] = r[6]; Sy[1] = r[7]; Sy[2] = r[8]; Sy[3] = Sy[4] = Sy[5] = 0; Sy[6] = -r[0]; Sy[7] = -r[1]; Sy[8] = -r[2];
  Sz[0] = -r[3]; Sz[1] = -r[4]; Sz[2] = -r[5]; Sz[3] = r[0]; Sz[4] = r[1]; Sz[5] = r[2]; Sz[6] = Sz[7] = Sz[8] = 0;
}
// out(3x3) = dq_dR * [vec(A Sx) vec(A Sy) vec(A Sz)], vec column-major
VDO_HD void dq_chain(const double* D, const double* A, const double* Sx, const double* Sy, const double* Sz, double* out) {
  const double* S[3] = {Sx, Sy, Sz};
  for (int c = 0; c < 3; ++c) {
    double M[9];
    m3_mul(A, S[c], M);
    for (int r = 0; r < 3; ++r) {
      double acc = 0;
      for (int k = 0; k < 9; ++k) acc += D[r * 9 + k] * M[3 * (k % 3) + k / 3];
      out[3 * r + c] = acc;
    }
  }
}

// EdgeSE3: error (6) and Jacobians (row-major 6x6).  Z is the measurement, Xi/Xj the vertex estimates.
VDO_HD void edge_se3_eval(const Iso& Z, const Iso& Xi, const Iso& Xj, double* e, double* Ji, double* Jj, bool jac) {
  Iso A, Xii, B, E;
  iso_inv(Z, A); iso_inv(Xi, Xii);
  iso_mul(Xii, Xj, B);
  iso_mul(A, B, E);
  iso_to_mqt(E, e);
  if (!jac) return;
  double D[27], S[9], T[9], Sx[9], Sy[9], Sz[9], blk[9];
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  dq_dR(E.R, D);
  for (int i = 0; i < 36; ++i) Ji[i] = Jj[i] = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { Ji[6 * r + c] = -A.R[3 * r + c]; Jj[6 * r + c] = E.R[3 * r + c]; }
  {  // Ra * skewT(tb),  skewT(v) = 2 [v]x
    double x = 2 * B.t[0], y = 2 * B.t[1], z = 2 * B.t[2];
    S[0] = 0; S[1] = -z; S[2] = y; S[3] = z; S[4] = 0; S[5] = -x; S[6] = -y; S[7] = x; S[8] = 0;
    m3_mul(A.R, S, T);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ji[6 * r + 3 + c] = T[3 * r + c];
  }
  skew3(B.R, -1.0, Sx, Sy, Sz); dq_chain(D, A.R, Sx, Sy, Sz, blk);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ji[6 * (3 + r) + 3 + c] = blk[3 * r + c];
  skew3(I3, 1.0, Sx, Sy, Sz); dq_chain(D, E.R, Sx, Sy, Sz, blk);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Jj[6 * (3 + r) + 3 + c] = blk[3 * r + c];
}
// EdgeSE3Prior with identity offset: error and Jacobian
VDO_HD void edge_prior_eval(const Iso& Z, const Iso& X, double* e, double* J, bool jac) {
  Iso Zi, A;
  iso_inv(Z, Zi); iso_mul(Zi, X, A);
  iso_to_mqt(A, e);
  if (!jac) return;
  double D[27], Sx[9], Sy[9], Sz[9], blk[9];
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  dq_dR(A.R, D);
  for (int i = 0; i < 36; ++i) J[i] = 0;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J[6 * r + c] = A.R[3 * r + c];
  skew3(I3, 1.0, Sx, Sy, Sz); dq_chain(D, A.R, Sx, Sy, Sz, blk);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J[6 * (3 + r) + 3 + c] = blk[3 * r + c];
}

// ---- symmetric 6x6 accumulators: 21 upper-triangle entries in row-major order (0,0),(0,1)..(0,5),(1,1).. ----
VDO_HD int sym6_idx(int r, int c) { return r * 6 - r * (r - 1) / 2 + (c - r); }  // r <= c

// EdgeSE3PointXYZ seen from its se3 vertex: with Zc = R^T(p - t), J_c = [-I | 2[Zc]x].
// Adds w * J_c^T J_c into A21 and -w * J_c^T err into g6.
VDO_HD void obs_accumulate_pose(const double* Zc, const double* err, double w, double* A21, double* g6) {
  // top-left I
  A21[sym6_idx(0, 0)] += w; A21[sym6_idx(1, 1)] += w; A21[sym6_idx(2, 2)] += w;
  // top-right block = -S, S = 2[Zc]x = [[0,-2z,2y],[2z,0,-2x],[-2y,2x,0]]
  const double x2 = 2 * Zc[0], y2 = 2 * Zc[1], z2 = 2 * Zc[2];
  A21[sym6_idx(0, 4)] += w * z2;  A21[sym6_idx(0, 5)] += -w * y2;
  A21[sym6_idx(1, 3)] += -w * z2; A21[sym6_idx(1, 5)] += w * x2;
  A21[sym6_idx(2, 3)] += w * y2;  A21[sym6_idx(2, 4)] += -w * x2;
  // bottom-right S^T S = 4(|Z|^2 I - Z Z^T)
  const double xx = x2 * x2, yy = y2 * y2, zz = z2 * z2;
  A21[sym6_idx(3, 3)] += w * (yy + zz); A21[sym6_idx(3, 4)] += -w * x2 * y2; A21[sym6_idx(3, 5)] += -w * x2 * z2;
  A21[sym6_idx(4, 4)] += w * (xx + zz); A21[sym6_idx(4, 5)] += -w * y2 * z2;
  A21[sym6_idx(5, 5)] += w * (xx + yy);
  // -J^T err = [err ; -S^T err] = [err ; S err] = [err ; 2 Zc x err]
  double c[3];
  cross3(Zc, err, c);
  g6[0] += w * err[0]; g6[1] += w * err[1]; g6[2] += w * err[2];
  g6[3] += 2 * w * c[0]; g6[4] += 2 * w * c[1]; g6[5] += 2 * w * c[2];
}
// J_c v (3) for v = [vt, vr]
VDO_HD void obs_Jc_mul(const double* Zc, const double* v, double* o) {
  double c[3];
  cross3(Zc, v + 3, c);
  o[0] = -v[0] + 2 * c[0]; o[1] = -v[1] + 2 * c[1]; o[2] = -v[2] + 2 * c[2];
}
// J_c^T a (6)
VDO_HD void obs_JcT_mul(const double* Zc, const double* a, double* o) {
  double c[3];
  cross3(Zc, a, c);
  o[0] = -a[0]; o[1] = -a[1]; o[2] = -a[2];
  o[3] = -2 * c[0]; o[4] = -2 * c[1]; o[5] = -2 * c[2];
}

// LandmarkMotionTernaryEdge seen from its motion vertex: q = H^-1 p2, J_h = [I | -[q]x].
VDO_HD void ter_accumulate_pose(const double* q, const double* err, double w, double* A21, double* g6) {
  A21[sym6_idx(0, 0)] += w; A21[sym6_idx(1, 1)] += w; A21[sym6_idx(2, 2)] += w;
  // top-right T = -[q]x = [[0,qz,-qy],[-qz,0,qx],[qy,-qx,0]]
  A21[sym6_idx(0, 4)] += w * q[2];  A21[sym6_idx(0, 5)] += -w * q[1];
  A21[sym6_idx(1, 3)] += -w * q[2]; A21[sym6_idx(1, 5)] += w * q[0];
  A21[sym6_idx(2, 3)] += w * q[1];  A21[sym6_idx(2, 4)] += -w * q[0];
  // T^T T = |q|^2 I - q q^T
  const double xx = q[0] * q[0], yy = q[1] * q[1], zz = q[2] * q[2];
  A21[sym6_idx(3, 3)] += w * (yy + zz); A21[sym6_idx(3, 4)] += -w * q[0] * q[1]; A21[sym6_idx(3, 5)] += -w * q[0] * q[2];
  A21[sym6_idx(4, 4)] += w * (xx + zz); A21[sym6_idx(4, 5)] += -w * q[1] * q[2];
  A21[sym6_idx(5, 5)] += w * (xx + yy);
  // -J^T err = [-err ; -(q x err)]
  double c[3];
  cross3(q, err, c);
  g6[0] -= w * err[0]; g6[1] -= w * err[1]; g6[2] -= w * err[2];
  g6[3] -= w * c[0]; g6[4] -= w * c[1]; g6[5] -= w * c[2];
}
VDO_HD void ter_Jh_mul(const double* q, const double* v, double* o) {  // vt + vr x q
  double c[3];
  cross3(v + 3, q, c);
  o[0] = v[0] + c[0]; o[1] = v[1] + c[1]; o[2] = v[2] + c[2];
}
VDO_HD void ter_JhT_mul(const double* q, const double* a, double* o) {  // [a ; q x a]
  double c[3];
  cross3(q, a, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}

// in-place Cholesky inverse of a symmetric positive definite 6x6 (row-major full storage). Returns false if not SPD.
VDO_HD bool spd6_inverse(double* M) {
  double L[36], Li[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) { L[i] = 0; Li[i] = 0; }
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = M[7 * j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k];
    if (!(d > 0)) { ok = false; d = 1.0; }
    d = sqrt(d);
    L[7 * j] = d;
    const double id = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = M[6 * i + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
      L[6 * i + j] = s * id;
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    Li[7 * j] = 1.0 / L[7 * j];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = 0;
#pragma unroll
      for (int k = j; k < i; ++k) s -= L[6 * i + k] * Li[6 * k + j];
      Li[6 * i + j] = s / L[7 * i];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = 0;
#pragma unroll
      for (int k = i; k < 6; ++k) s += Li[6 * k + i] * Li[6 * k + j];
      M[6 * i + j] = M[6 * j + i] = s;
    }
  return true;
}

}  // namespace vdo
