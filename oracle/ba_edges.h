/*
 * oracle/ba_edges.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Error functions and Jacobians of the four g2o edge types used by the reference's batch optimisers
 * (Optimizer::FullBatchOptimization / PartialBatchOptimization, src/Optimizer.cc:1232-2175, 42-1230),
 * restated on plain arrays.  Parity status: UNPINNED (no reference tests exist; see se3_math.h).
 *
 * All Jacobians are row-major, error-dimension x vertex-dimension.
 */
#ifndef VDO_ORACLE_BA_EDGES_H
#define VDO_ORACLE_BA_EDGES_H

#include "se3_math.h"

/* d(qx,qy,qz)/d(R) for the quaternion extracted from R, 3x9, columns indexed by the COLUMN-major
 * flattening of R (r00,r10,r20,r01,...).  Follows dquat2mat.cpp:35-84 (branch selection _q2m, sign flip
 * when the branch's w <= 0) and the closed forms tabulated in dquat2mat_maxima_generated.cpp. */
static inline void dq_dR(const double *R /*row-major*/, double *D /*3x9 row-major*/) {
  const double r00 = R[0], r01 = R[1], r02 = R[2], r10 = R[3], r11 = R[4], r12 = R[5], r20 = R[6], r21 = R[7], r22 = R[8];
  for (int i = 0; i < 27; ++i) D[i] = 0.0;
  double tr = r00 + r11 + r22, S, qw;
  /* column index helper: c(i,j) = i + 3 j */
  if (tr > 0) {
    S = sqrt(tr + 1.0) * 2; qw = 0.25 * S;
    double a = 0.25 * S;              /* = qw */
    double i3 = 1.0 / (a * a * a), i1 = 1.0 / a;
    double gx = -0.03125 * (r21 - r12) * i3, gy = 0.03125 * (r20 - r02) * i3, gz = -0.03125 * (r10 - r01) * i3;
    D[0 * 9 + 0] = gx; D[0 * 9 + 4] = gx; D[0 * 9 + 8] = gx; D[0 * 9 + 5] = 0.25 * i1; D[0 * 9 + 7] = -0.25 * i1;
    D[1 * 9 + 0] = gy; D[1 * 9 + 4] = gy; D[1 * 9 + 8] = gy; D[1 * 9 + 2] = -0.25 * i1; D[1 * 9 + 6] = 0.25 * i1;
    D[2 * 9 + 0] = gz; D[2 * 9 + 4] = gz; D[2 * 9 + 8] = gz; D[2 * 9 + 1] = 0.25 * i1; D[2 * 9 + 3] = -0.25 * i1;
  } else if ((r00 > r11) & (r00 > r22)) {
    S = sqrt(1.0 + r00 - r11 - r22) * 2; qw = (r21 - r12) / S;
    double a = 0.25 * S;              /* = qx */
    double i3 = 1.0 / (a * a * a), i1 = 1.0 / a;
    double s1 = r10 + r01, s2 = r20 + r02;
    D[0 * 9 + 0] = 0.125 * i1; D[0 * 9 + 4] = -0.125 * i1; D[0 * 9 + 8] = -0.125 * i1;
    D[1 * 9 + 0] = -0.03125 * i3 * s1; D[1 * 9 + 4] = 0.03125 * i3 * s1; D[1 * 9 + 8] = 0.03125 * i3 * s1;
    D[1 * 9 + 1] = 0.25 * i1; D[1 * 9 + 3] = 0.25 * i1;
    D[2 * 9 + 0] = -0.03125 * i3 * s2; D[2 * 9 + 4] = 0.03125 * i3 * s2; D[2 * 9 + 8] = 0.03125 * i3 * s2;
    D[2 * 9 + 2] = 0.25 * i1; D[2 * 9 + 6] = 0.25 * i1;
  } else if (r11 > r22) {
    S = sqrt(1.0 + r11 - r00 - r22) * 2; qw = (r02 - r20) / S;
    double a = 0.25 * S;              /* = qy */
    double i3 = 1.0 / (a * a * a), i1 = 1.0 / a;
    double s1 = r10 + r01, s2 = r21 + r12;
    D[0 * 9 + 0] = 0.03125 * i3 * s1; D[0 * 9 + 4] = -0.03125 * i3 * s1; D[0 * 9 + 8] = 0.03125 * i3 * s1;
    D[0 * 9 + 1] = 0.25 * i1; D[0 * 9 + 3] = 0.25 * i1;
    D[1 * 9 + 0] = -0.125 * i1; D[1 * 9 + 4] = 0.125 * i1; D[1 * 9 + 8] = -0.125 * i1;
    D[2 * 9 + 0] = 0.03125 * i3 * s2; D[2 * 9 + 4] = -0.03125 * i3 * s2; D[2 * 9 + 8] = 0.03125 * i3 * s2;
    D[2 * 9 + 5] = 0.25 * i1; D[2 * 9 + 7] = 0.25 * i1;
  } else {
    S = sqrt(1.0 + r22 - r00 - r11) * 2; qw = (r10 - r01) / S;
    double a = 0.25 * S;              /* = qz */
    double i3 = 1.0 / (a * a * a), i1 = 1.0 / a;
    double s1 = r20 + r02, s2 = r21 + r12;
    D[0 * 9 + 0] = 0.03125 * i3 * s1; D[0 * 9 + 4] = 0.03125 * i3 * s1; D[0 * 9 + 8] = -0.03125 * i3 * s1;
    D[0 * 9 + 2] = 0.25 * i1; D[0 * 9 + 6] = 0.25 * i1;
    D[1 * 9 + 0] = 0.03125 * i3 * s2; D[1 * 9 + 4] = 0.03125 * i3 * s2; D[1 * 9 + 8] = -0.03125 * i3 * s2;
    D[1 * 9 + 5] = 0.25 * i1; D[1 * 9 + 7] = 0.25 * i1;
    D[2 * 9 + 0] = -0.125 * i1; D[2 * 9 + 4] = -0.125 * i1; D[2 * 9 + 8] = 0.125 * i1;
  }
  if (qw <= 0)
    for (int i = 0; i < 27; ++i) D[i] = -D[i];
}

/* "skew" helpers of isometry3d_gradients.h:40-84.  Outputs row-major. */
static inline void g2o_skew(const double *v, double *s) { /* :41-46 */
  double x = 2 * v[0], y = 2 * v[1], z = 2 * v[2];
  s[0] = 0; s[1] = z; s[2] = -y; s[3] = -z; s[4] = 0; s[5] = x; s[6] = y; s[7] = -x; s[8] = 0;
}
static inline void g2o_skewT(const double *v, double *s) { /* :49-54 */
  double x = 2 * v[0], y = 2 * v[1], z = 2 * v[2];
  s[0] = 0; s[1] = -z; s[2] = y; s[3] = z; s[4] = 0; s[5] = -x; s[6] = -y; s[7] = x; s[8] = 0;
}
static inline void g2o_skew3(const double *R, double sgn, double *Sx, double *Sy, double *Sz) { /* :57-84, sgn=+1 skew, -1 skewT */
  double r[9];
  for (int i = 0; i < 9; ++i) r[i] = 2 * sgn * R[i];
  Sx[0] = Sx[1] = Sx[2] = 0; Sx[3] = -r[6]; Sx[4] = -r[7]; Sx[5] = -r[8]; Sx[6] = r[3]; Sx[7] = r[4]; Sx[8] = r[5];
  Sy[0] = r[6]; Sy[1] = r[7]; Sy[2] = r[8]; Sy[3] = Sy[4] = Sy[5] = 0; Sy[6] = -r[0]; Sy[7] = -r[1]; Sy[8] = -r[2];
  Sz[0] = -r[3]; Sz[1] = -r[4]; Sz[2] = -r[5]; Sz[3] = r[0]; Sz[4] = r[1]; Sz[5] = r[2]; Sz[6] = Sz[7] = Sz[8] = 0;
}
/* out(3x3) = dq_dR(3x9) * [vec(A*Sx) vec(A*Sy) vec(A*Sz)] with column-major vec (isometry3d_gradients.h:234-241) */
static inline void dq_chain(const double *D, const double *A, const double *Sx, const double *Sy, const double *Sz, double *out) {
  const double *S[3] = {Sx, Sy, Sz};
  for (int c = 0; c < 3; ++c) {
    double M[9];
    m3_mul(A, S[c], M);
    for (int r = 0; r < 3; ++r) {
      double acc = 0;
      for (int k = 0; k < 9; ++k) { /* k = i + 3 j, column-major */
        int i = k % 3, j = k / 3;
        acc += D[r * 9 + k] * M[3 * i + j];
      }
      out[3 * r + c] = acc;
    }
  }
}

/* ---------------- EdgeSE3 (edge_se3.cpp:77-82, 91-104; isometry3d_gradients.h:191-261) ---------------- */
static inline void edge_se3_error(const double *Z, const double *Xi, const double *Xj, double *e) {
  double Zi[12], Xii[12], B[12], E[12];
  iso_inv(Z, Zi); iso_inv(Xi, Xii);
  iso_mul(Zi, Xii, B);          /* (_inverseMeasurement * from^-1) * to, Eigen evaluates left to right */
  iso_mul(B, Xj, E);
  iso_to_mqt(E, e);
}
static inline void edge_se3_jac(const double *Z, const double *Xi, const double *Xj, double *Ji, double *Jj) {
  double A[12], Xii[12], B[12], E[12], D[27], S[9], T[9], Sx[9], Sy[9], Sz[9], blk[9], I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  iso_inv(Z, A); iso_inv(Xi, Xii); iso_mul(Xii, Xj, B); iso_mul(A, B, E);
  dq_dR(E, D);
  for (int i = 0; i < 36; ++i) Ji[i] = Jj[i] = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { Ji[6 * r + c] = -A[3 * r + c]; Jj[6 * r + c] = E[3 * r + c]; }
  g2o_skewT(B + 9, S); m3_mul(A, S, T);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ji[6 * r + 3 + c] = T[3 * r + c];
  g2o_skew3(B, -1.0, Sx, Sy, Sz); dq_chain(D, A, Sx, Sy, Sz, blk);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ji[6 * (3 + r) + 3 + c] = blk[3 * r + c];
  g2o_skew3(I3, 1.0, Sx, Sy, Sz); dq_chain(D, E, Sx, Sy, Sz, blk);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Jj[6 * (3 + r) + 3 + c] = blk[3 * r + c];
}

/* ---------------- EdgeSE3Prior with identity offset (edge_se3_prior.cpp:89-102; isometry3d_gradients.h:264-325) -------- */
static inline void edge_prior_error(const double *Z, const double *X, double *e) {
  double Zi[12], E[12];
  iso_inv(Z, Zi); iso_mul(Zi, X, E);
  iso_to_mqt(E, e);
}
static inline void edge_prior_jac(const double *Z, const double *X, double *J) {
  double Zi[12], A[12], D[27], Sx[9], Sy[9], Sz[9], blk[9], I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  iso_inv(Z, Zi); iso_mul(Zi, X, A);      /* E = A * P with P = identity */
  dq_dR(A, D);
  for (int i = 0; i < 36; ++i) J[i] = 0;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J[6 * r + c] = A[3 * r + c];
  /* dte/dq = Ra * skew(tb), tb = 0 -> zero block */
  g2o_skew3(I3, 1.0, Sx, Sy, Sz); dq_chain(D, A, Sx, Sy, Sz, blk);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J[6 * (3 + r) + 3 + c] = blk[3 * r + c];
}

/* ---------------- EdgeSE3PointXYZ with identity offset (edge_se3_pointxyz.cpp:99-140; parameter_se3_offset.cpp:77-82) --- */
static inline void edge_obs_error(const double *Xc, const double *p, const double *z, double *e) {
  double w2n[12], q[3];
  iso_inv(Xc, w2n);              /* n2w = X * offset(=I); w2n = n2w^-1 */
  iso_apply(w2n, p, q);
  e[0] = q[0] - z[0]; e[1] = q[1] - z[1]; e[2] = q[2] - z[2];
}
static inline void edge_obs_jac(const double *Xc, const double *p, double *Jc /*3x6*/, double *Jp /*3x3*/) {
  double w2l[12], Zc[3];
  iso_inv(Xc, w2l); iso_apply(w2l, p, Zc);
  for (int i = 0; i < 18; ++i) Jc[i] = 0;
  Jc[0] = Jc[7] = Jc[14] = -1.0;
  Jc[0 * 6 + 4] = -2 * Zc[2]; Jc[0 * 6 + 5] = 2 * Zc[1];
  Jc[1 * 6 + 3] = 2 * Zc[2];  Jc[1 * 6 + 5] = -2 * Zc[0];
  Jc[2 * 6 + 3] = -2 * Zc[1]; Jc[2 * 6 + 4] = 2 * Zc[0];
  for (int i = 0; i < 9; ++i) Jp[i] = w2l[i];
}

/* ---------------- LandmarkMotionTernaryEdge (types_dyn_slam3d.cpp:53-85), measurement = 0 ---------------- */
static inline void edge_ter_error(const double *p1, const double *p2, const double *H, double *e) {
  double Hi[12], q[3];
  iso_inv(H, Hi); iso_apply(Hi, p2, q);
  e[0] = p1[0] - q[0]; e[1] = p1[1] - q[1]; e[2] = p1[2] - q[2];
}
static inline void edge_ter_jac(const double *p2, const double *H, double *J1 /*3x3*/, double *J2 /*3x3*/, double *Jh /*3x6*/) {
  double Hi[12], q[3];
  iso_inv(H, Hi); iso_apply(Hi, p2, q);
  for (int i = 0; i < 9; ++i) { J1[i] = 0; J2[i] = -Hi[i]; }
  J1[0] = J1[4] = J1[8] = 1.0;
  for (int i = 0; i < 18; ++i) Jh[i] = 0;
  Jh[0] = Jh[7] = Jh[14] = 1.0;
  Jh[0 * 6 + 4] = q[2];  Jh[0 * 6 + 5] = -q[1];
  Jh[1 * 6 + 3] = -q[2]; Jh[1 * 6 + 5] = q[0];
  Jh[2 * 6 + 3] = q[1];  Jh[2 * 6 + 4] = -q[0];
}

/* Huber kernel (robust_kernel_impl.cpp:65-91; dsqr is stored as float, robust_kernel_impl.h:84) */
static inline void huber(double e2, double delta, double *rho /*3*/) {
  float dsqr_f = (float)(delta * delta);
  double dsqr = dsqr_f;
  if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.0; rho[2] = 0.0; }
  else {
    double s = sqrt(e2);
    rho[0] = 2 * s * delta - dsqr;
    rho[1] = delta / s;
    rho[2] = -0.5 * rho[1] / e2;
  }
}

#endif
