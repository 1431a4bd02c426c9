/*
 * oracle/ba_block.h -- TEST INFRASTRUCTURE ONLY (CPU oracle + timed CPU baseline; never on the product path).
 *
 * Second linear solver of the oracle: the SAME direct Cholesky solve of the full (3P + 6C)-dimensional system
 * (H + lambda I) x = b that the reference hands to LinearSolverCSparse (g2o/solvers/linear_solver_csparse.h:108-144),
 * organised as a blocked factorisation so that BASELINE configs 4 and 5 (50 k / 1 M point tracks) finish in
 * minutes instead of days:
 *
 *   elimination order   points first -- grouped by connected component of the point-point (ternary edge) graph,
 *                       i.e. one static landmark or one chain of per-frame copies of a dynamic landmark -- then
 *                       the se3 vertices in a caller-supplied order (bench / golden generator: by frame, which makes
 *                       the reduced matrix banded).  The reference orders by block-AMD (linear_solver_csparse.h:
 *                       248-346); an elimination order changes rounding only, never the solution.
 *   point blocks        dense Cholesky of each component's (3m x 3m) block, its Schur update -W^T A^-1 W scattered
 *                       into the reduced matrix in component order (deterministic for any thread count).
 *   reduced matrix      variable-band tiled right-looking Cholesky (96 x 96 tiles), OpenMP over the tiles of one
 *                       panel step.
 *
 * ba_lm.c's scalar up-looking sparse Cholesky (solver 0) stays the statement closest to CSparse; this file (solver 1)
 * is checked against it on small graphs in tests/test_oracle_ba.py (same x to ~1e-12 relative).
 *
 * Parity status: UNPINNED, like the rest of the batch oracle (see ba_lm.c).
 */
#ifndef VDO_ORACLE_BA_BLOCK_H
#define VDO_ORACLE_BA_BLOCK_H

#ifdef _OPENMP
#include <omp.h>
#else
static inline int omp_get_thread_num(void) { return 0; }
static inline int omp_get_num_threads(void) { return 1; }
static inline int omp_get_max_threads(void) { return 1; }
#endif

#define BLK_NB 96            /* tile edge in scalars */
#define BLK_VT 16            /* se3 vertices per tile (BLK_NB / 6) */

typedef struct {
  int P, C, nthreads;
  int *pos, *inv;                       /* se3 vertex <-> position in the elimination order */
  int ncomp; int *comp_ptr, *comp_pts, *loc_of_pt, *comp_of_pt;
  int64_t *pobs_ptr; int *pobs;         /* point -> EdgeSE3PointXYZ ids */
  int *cter_ptr, *cter;                 /* component -> ternary edge ids */
  int *cse3_ptr, *cse3;                 /* component -> se3 vertices it touches, ascending position */
  int max_m, max_k;
  double *Hll, *Hpp, *Wobs, *Wter, *Hse3e;
  int T; int *hi; int64_t *toff; double *S, *r;
  double t_schur, t_chol, t_back;
} blk_t;

static int blk_find(int *uf, int a) { while (uf[a] != a) { uf[a] = uf[uf[a]]; a = uf[a]; } return a; }
static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

static void blk_structure(const ba_t *g, blk_t *B, const int *se3_pos, int nthreads) {
  const int P = g->n_pt, C = g->n_se3;
  memset(B, 0, sizeof *B);
  B->P = P; B->C = C;
  B->nthreads = nthreads > 0 ? nthreads : omp_get_max_threads();
  B->pos = (int *)malloc(sizeof(int) * (C + 1)); B->inv = (int *)malloc(sizeof(int) * (C + 1));
  for (int v = 0; v < C; ++v) B->pos[v] = se3_pos ? se3_pos[v] : v;
  for (int v = 0; v < C; ++v) B->inv[B->pos[v]] = v;
  /* components of the point graph */
  int *uf = (int *)malloc(sizeof(int) * (P + 1));
  for (int p = 0; p < P; ++p) uf[p] = p;
  for (int e = 0; e < g->n_ter; ++e) {
    int a = blk_find(uf, g->ter_pph[3 * e]), b = blk_find(uf, g->ter_pph[3 * e + 1]);
    if (a != b) { if (a < b) uf[b] = a; else uf[a] = b; }      /* root = smallest point id */
  }
  B->comp_of_pt = (int *)malloc(sizeof(int) * (P + 1));
  int nc = 0;
  for (int p = 0; p < P; ++p) if (blk_find(uf, p) == p) B->comp_of_pt[p] = nc++;
  for (int p = 0; p < P; ++p) B->comp_of_pt[p] = B->comp_of_pt[blk_find(uf, p)];
  free(uf);
  B->ncomp = nc;
  B->comp_ptr = (int *)calloc((size_t)nc + 2, sizeof(int));
  for (int p = 0; p < P; ++p) B->comp_ptr[B->comp_of_pt[p] + 1]++;
  for (int c = 0; c < nc; ++c) B->comp_ptr[c + 1] += B->comp_ptr[c];
  B->comp_pts = (int *)malloc(sizeof(int) * (P + 1)); B->loc_of_pt = (int *)malloc(sizeof(int) * (P + 1));
  int *fill = (int *)calloc((size_t)nc + 1, sizeof(int));
  for (int p = 0; p < P; ++p) { int c = B->comp_of_pt[p]; B->loc_of_pt[p] = fill[c]; B->comp_pts[B->comp_ptr[c] + fill[c]++] = p; }
  /* point -> obs edges */
  B->pobs_ptr = (int64_t *)calloc((size_t)P + 2, sizeof(int64_t));
  for (int e = 0; e < g->n_obs; ++e) B->pobs_ptr[g->obs_cp[2 * e + 1] + 1]++;
  for (int p = 0; p < P; ++p) B->pobs_ptr[p + 1] += B->pobs_ptr[p];
  B->pobs = (int *)malloc(sizeof(int) * ((size_t)g->n_obs + 1));
  int64_t *pf = (int64_t *)calloc((size_t)P + 1, sizeof(int64_t));
  for (int e = 0; e < g->n_obs; ++e) { int p = g->obs_cp[2 * e + 1]; B->pobs[B->pobs_ptr[p] + pf[p]++] = e; }
  free(pf);
  /* component -> ternary edges */
  B->cter_ptr = (int *)calloc((size_t)nc + 2, sizeof(int));
  for (int e = 0; e < g->n_ter; ++e) B->cter_ptr[B->comp_of_pt[g->ter_pph[3 * e]] + 1]++;
  for (int c = 0; c < nc; ++c) B->cter_ptr[c + 1] += B->cter_ptr[c];
  B->cter = (int *)malloc(sizeof(int) * ((size_t)g->n_ter + 1));
  memset(fill, 0, sizeof(int) * ((size_t)nc + 1));
  for (int e = 0; e < g->n_ter; ++e) { int c = B->comp_of_pt[g->ter_pph[3 * e]]; B->cter[B->cter_ptr[c] + fill[c]++] = e; }
  free(fill);
  /* component -> se3 vertices (unique, by position) */
  B->cse3_ptr = (int *)calloc((size_t)nc + 2, sizeof(int));
  int64_t cap = (int64_t)g->n_obs + g->n_ter + 1;
  int *tmp = (int *)malloc(sizeof(int) * (size_t)cap);
  B->cse3 = (int *)malloc(sizeof(int) * (size_t)cap);
  int64_t w = 0;
  for (int c = 0; c < nc; ++c) {
    int n = 0;
    for (int q = B->comp_ptr[c]; q < B->comp_ptr[c + 1]; ++q) {
      int p = B->comp_pts[q];
      for (int64_t o = B->pobs_ptr[p]; o < B->pobs_ptr[p + 1]; ++o) tmp[n++] = B->pos[g->obs_cp[2 * B->pobs[o]]];
    }
    for (int q = B->cter_ptr[c]; q < B->cter_ptr[c + 1]; ++q) tmp[n++] = B->pos[g->ter_pph[3 * B->cter[q] + 2]];
    qsort(tmp, (size_t)n, sizeof(int), cmp_int);
    int k = 0;
    for (int i = 0; i < n; ++i) if (i == 0 || tmp[i] != tmp[i - 1]) B->cse3[w + k++] = B->inv[tmp[i]];
    w += k;
    B->cse3_ptr[c + 1] = (int)w;
    int m = B->comp_ptr[c + 1] - B->comp_ptr[c];
    if (m > B->max_m) B->max_m = m;
    if (k > B->max_k) B->max_k = k;
  }
  free(tmp);
  /* envelope of the reduced matrix in tiles */
  B->T = (C + BLK_VT - 1) / BLK_VT;
  B->hi = (int *)malloc(sizeof(int) * (B->T + 1));
  for (int J = 0; J < B->T; ++J) B->hi[J] = J;
  for (int c = 0; c < nc; ++c) {
    int a = B->cse3_ptr[c], b = B->cse3_ptr[c + 1];
    if (b > a) { int J = B->pos[B->cse3[a]] / BLK_VT, I = B->pos[B->cse3[b - 1]] / BLK_VT; if (I > B->hi[J]) B->hi[J] = I; }
  }
  for (int e = 0; e < g->n_se3e; ++e) {
    int a = B->pos[g->se3e_ij[2 * e]], b = B->pos[g->se3e_ij[2 * e + 1]];
    int J = (a < b ? a : b) / BLK_VT, I = (a < b ? b : a) / BLK_VT;
    if (I > B->hi[J]) B->hi[J] = I;
  }
  for (int J = 1; J < B->T; ++J) if (B->hi[J - 1] > B->hi[J]) B->hi[J] = B->hi[J - 1];
  B->toff = (int64_t *)malloc(sizeof(int64_t) * (B->T + 1));
  int64_t nt = 0;
  for (int J = 0; J < B->T; ++J) { B->toff[J] = nt; nt += B->hi[J] - J + 1; }
  B->toff[B->T] = nt;
  B->S = (double *)malloc(sizeof(double) * (size_t)nt * BLK_NB * BLK_NB);
  B->r = (double *)malloc(sizeof(double) * (size_t)(B->T > 0 ? B->T : 1) * BLK_NB);
  B->Hll = (double *)malloc(sizeof(double) * 9 * ((size_t)P + 1));
  B->Hpp = (double *)malloc(sizeof(double) * 36 * ((size_t)C + 1));
  B->Wobs = (double *)malloc(sizeof(double) * 18 * ((size_t)g->n_obs + 1));
  B->Wter = (double *)malloc(sizeof(double) * 45 * ((size_t)g->n_ter + 1));
  B->Hse3e = (double *)malloc(sizeof(double) * 36 * ((size_t)g->n_se3e + 1));
}

static void blk_free(blk_t *B) {
  free(B->pos); free(B->inv); free(B->comp_ptr); free(B->comp_pts); free(B->loc_of_pt); free(B->comp_of_pt);
  free(B->pobs_ptr); free(B->pobs); free(B->cter_ptr); free(B->cter); free(B->cse3_ptr); free(B->cse3);
  free(B->hi); free(B->toff); free(B->S); free(B->r); free(B->Hll); free(B->Hpp); free(B->Wobs); free(B->Wter); free(B->Hse3e);
}

/* out (da x db) (+)= w * Ja^T Jb;  Ja is dim x da, Jb is dim x db, row-major */
static inline void jtj(const double *Ja, int da, const double *Jb, int db, int dim, double w, double *out, int accumulate) {
  for (int r = 0; r < da; ++r)
    for (int c = 0; c < db; ++c) {
      double s = 0;
      for (int k = 0; k < dim; ++k) s += Ja[k * da + r] * Jb[k * db + c];
      if (accumulate) out[r * db + c] += w * s; else out[r * db + c] = w * s;
    }
}
static inline void jte(const double *J, int d, int dim, double w, const double *e, double *b) {   /* b -= w J^T e */
  for (int c = 0; c < d; ++c) {
    double s = 0;
    for (int k = 0; k < dim; ++k) s += J[k * d + c] * e[k];
    b[c] -= w * s;
  }
}

/* buildSystem (block_solver.hpp:501-560) into blocks: same edges, same order, same Jacobians as build_system() of ba_lm.c.
 * Returns max |H_jj|. */
static double blk_build_system(ba_t *g, blk_t *B) {
  const int P = g->n_pt, C = g->n_se3;
  memset(B->Hll, 0, sizeof(double) * 9 * (size_t)P);
  memset(B->Hpp, 0, sizeof(double) * 36 * (size_t)C);
  memset(g->b, 0, sizeof(double) * (size_t)g->n);
  double e[6], rho[3], Ji[36], Jj[36];
  double *bp = g->b + 3 * (size_t)P;
  for (int k = 0; k < g->n_prior; ++k) {
    int v = g->prior_v[k];
    edge_prior_error(g->prior_Z + 12 * k, g->se3 + 12 * v, e);
    edge_prior_jac(g->prior_Z + 12 * k, g->se3 + 12 * v, Ji);
    jtj(Ji, 6, Ji, 6, 6, g->prior_w[k], B->Hpp + 36 * (size_t)v, 1);
    jte(Ji, 6, 6, g->prior_w[k], e, bp + 6 * v);
  }
  for (int k = 0; k < g->n_se3e; ++k) {
    int i = g->se3e_ij[2 * k], j = g->se3e_ij[2 * k + 1];
    const double *Z = g->se3e_Z + 12 * k;
    edge_se3_error(Z, g->se3 + 12 * i, g->se3 + 12 * j, e);
    edge_se3_jac(Z, g->se3 + 12 * i, g->se3 + 12 * j, Ji, Jj);
    double w = g->se3e_w[k];
    if (g->se3e_delta[k] > 0) {
      double c = 0; for (int q = 0; q < 6; ++q) c += e[q] * e[q];
      huber(w * c, g->se3e_delta[k], rho); w *= rho[1];
    }
    jtj(Ji, 6, Ji, 6, 6, w, B->Hpp + 36 * (size_t)i, 1); jte(Ji, 6, 6, w, e, bp + 6 * i);
    jtj(Jj, 6, Jj, 6, 6, w, B->Hpp + 36 * (size_t)j, 1); jte(Jj, 6, 6, w, e, bp + 6 * j);
    jtj(Ji, 6, Jj, 6, 6, w, B->Hse3e + 36 * (size_t)k, 0);
  }
  for (int k = 0; k < g->n_obs; ++k) {
    int c = g->obs_cp[2 * k], p = g->obs_cp[2 * k + 1];
    edge_obs_error(g->se3 + 12 * c, g->pt + 3 * p, g->obs_z + 3 * k, e);
    edge_obs_jac(g->se3 + 12 * c, g->pt + 3 * p, Ji, Jj);
    double w = g->obs_w[k];
    if (g->obs_delta[k] > 0) { huber(w * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]), g->obs_delta[k], rho); w *= rho[1]; }
    jtj(Ji, 6, Ji, 6, 3, w, B->Hpp + 36 * (size_t)c, 1); jte(Ji, 6, 3, w, e, bp + 6 * c);
    jtj(Jj, 3, Jj, 3, 3, w, B->Hll + 9 * (size_t)p, 1); jte(Jj, 3, 3, w, e, g->b + 3 * (size_t)p);
    jtj(Jj, 3, Ji, 6, 3, w, B->Wobs + 18 * (size_t)k, 0);      /* point x camera, 3 x 6 */
  }
  for (int k = 0; k < g->n_ter; ++k) {
    int p1 = g->ter_pph[3 * k], p2 = g->ter_pph[3 * k + 1], h = g->ter_pph[3 * k + 2];
    double J1[9], J2[9], Jh[18];
    edge_ter_error(g->pt + 3 * p1, g->pt + 3 * p2, g->se3 + 12 * h, e);
    edge_ter_jac(g->pt + 3 * p2, g->se3 + 12 * h, J1, J2, Jh);
    double w = g->ter_w[k];
    if (g->ter_delta[k] > 0) { huber(w * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]), g->ter_delta[k], rho); w *= rho[1]; }
    jtj(J1, 3, J1, 3, 3, w, B->Hll + 9 * (size_t)p1, 1); jte(J1, 3, 3, w, e, g->b + 3 * (size_t)p1);
    jtj(J2, 3, J2, 3, 3, w, B->Hll + 9 * (size_t)p2, 1); jte(J2, 3, 3, w, e, g->b + 3 * (size_t)p2);
    jtj(Jh, 6, Jh, 6, 3, w, B->Hpp + 36 * (size_t)h, 1); jte(Jh, 6, 3, w, e, bp + 6 * h);
    double *W = B->Wter + 45 * (size_t)k;
    jtj(J1, 3, J2, 3, 3, w, W, 0);            /* p1 x p2 */
    jtj(J1, 3, Jh, 6, 3, w, W + 9, 0);        /* p1 x h */
    jtj(J2, 3, Jh, 6, 3, w, W + 27, 0);       /* p2 x h */
  }
  double md = 0;
  for (int p = 0; p < P; ++p) for (int d = 0; d < 3; ++d) { double a = fabs(B->Hll[9 * (size_t)p + 4 * d]); if (a > md) md = a; }
  for (int v = 0; v < C; ++v) for (int d = 0; d < 6; ++d) { double a = fabs(B->Hpp[36 * (size_t)v + 7 * d]); if (a > md) md = a; }
  return md;
}

/* ---- dense pieces ---- */
static int dense_chol(double *A, int n, int lda) {      /* lower, in place; 0 if not positive definite */
  for (int j = 0; j < n; ++j) {
    double d = A[j * lda + j];
    for (int k = 0; k < j; ++k) d -= A[j * lda + k] * A[j * lda + k];
    if (!(d > 0)) return 0;
    d = sqrt(d); A[j * lda + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * lda + j];
      for (int k = 0; k < j; ++k) s -= A[i * lda + k] * A[j * lda + k];
      A[i * lda + j] = s / d;
    }
  }
  return 1;
}

/* component c: A = H_ll(c) + lambda I factored in place (lower), Y = L^-1 W (3m x 6k), yb = L^-1 b_l.  lmap: scratch se3->local */
static int blk_comp_factor(const ba_t *g, const blk_t *B, int c, double lambda, double *A, double *Y, double *yb, int *lmap) {
  const int p0 = B->comp_ptr[c], m = B->comp_ptr[c + 1] - p0, n3 = 3 * m;
  const int s0 = B->cse3_ptr[c], k = B->cse3_ptr[c + 1] - s0, n6 = 6 * k;
  for (int a = 0; a < k; ++a) lmap[B->cse3[s0 + a]] = a;
  memset(A, 0, sizeof(double) * (size_t)n3 * n3);
  memset(Y, 0, sizeof(double) * (size_t)n3 * (n6 > 0 ? n6 : 1));
  for (int q = 0; q < m; ++q) {
    int p = B->comp_pts[p0 + q];
    const double *H = B->Hll + 9 * (size_t)p;
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) A[(3 * q + r) * n3 + 3 * q + cc] = H[3 * r + cc];
    for (int r = 0; r < 3; ++r) { A[(3 * q + r) * n3 + 3 * q + r] += lambda; yb[3 * q + r] = g->b[3 * (size_t)p + r]; }
    for (int64_t o = B->pobs_ptr[p]; o < B->pobs_ptr[p + 1]; ++o) {
      int e = B->pobs[o], a = lmap[g->obs_cp[2 * e]];
      const double *W = B->Wobs + 18 * (size_t)e;
      for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 6; ++cc) Y[(3 * q + r) * n6 + 6 * a + cc] += W[6 * r + cc];
    }
  }
  for (int t = B->cter_ptr[c]; t < B->cter_ptr[c + 1]; ++t) {
    int e = B->cter[t];
    int q1 = B->loc_of_pt[g->ter_pph[3 * e]], q2 = B->loc_of_pt[g->ter_pph[3 * e + 1]], a = lmap[g->ter_pph[3 * e + 2]];
    const double *W = B->Wter + 45 * (size_t)e;
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) {
      A[(3 * q1 + r) * n3 + 3 * q2 + cc] += W[3 * r + cc];
      A[(3 * q2 + cc) * n3 + 3 * q1 + r] += W[3 * r + cc];
    }
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 6; ++cc) {
      Y[(3 * q1 + r) * n6 + 6 * a + cc] += W[9 + 6 * r + cc];
      Y[(3 * q2 + r) * n6 + 6 * a + cc] += W[27 + 6 * r + cc];
    }
  }
  if (!dense_chol(A, n3, n3)) return 0;
  for (int i = 0; i < n3; ++i) {
    double *yi = Y + (size_t)i * n6, s = yb[i];
    for (int j = 0; j < i; ++j) {
      double l = A[i * n3 + j];
      if (l == 0) continue;
      const double *yj = Y + (size_t)j * n6;
      for (int cc = 0; cc < n6; ++cc) yi[cc] -= l * yj[cc];
      s -= l * yb[j];
    }
    double inv = 1.0 / A[i * n3 + i];
    for (int cc = 0; cc < n6; ++cc) yi[cc] *= inv;
    yb[i] = s * inv;
  }
  return 1;
}

/* ---- tiles ---- */
static inline double *blk_tile(const blk_t *B, int I, int J) { return B->S + (size_t)(B->toff[J] + (I - J)) * BLK_NB * BLK_NB; }

static int tile_potrf(double *A) { return dense_chol(A, BLK_NB, BLK_NB); }

/* X <- X L^-T  (row x of X solves  y L^T = x) */
static void tile_trsm(double *X, const double *L) {
  for (int r = 0; r < BLK_NB; ++r) {
    double *x = X + r * BLK_NB;
    for (int j = 0; j < BLK_NB; ++j) {
      const double *l = L + j * BLK_NB;
      double s = 0;
#pragma omp simd reduction(+ : s)
      for (int k = 0; k < j; ++k) s += x[k] * l[k];
      x[j] = (x[j] - s) / l[j];
    }
  }
}

typedef double v4d __attribute__((vector_size(32)));
/* C -= A * B^T, all BLK_NB x BLK_NB row-major.  Bt = scratch for the transpose of B. */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void tile_gemm_nt(double *__restrict__ Cm, const double *__restrict__ A, const double *__restrict__ Bm, double *__restrict__ Bt) {
  for (int i = 0; i < BLK_NB; ++i) for (int j = 0; j < BLK_NB; ++j) Bt[j * BLK_NB + i] = Bm[i * BLK_NB + j];
  for (int i = 0; i < BLK_NB; i += 4)
    for (int j = 0; j < BLK_NB; j += 16) {
      double acc[4][16];
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 16; ++c) acc[r][c] = 0;
      for (int k = 0; k < BLK_NB; ++k) {
        const double *b = Bt + k * BLK_NB + j;
        for (int r = 0; r < 4; ++r) {
          double a = A[(i + r) * BLK_NB + k];
#pragma omp simd
          for (int c = 0; c < 16; ++c) acc[r][c] += a * b[c];
        }
      }
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 16; ++c) Cm[(i + r) * BLK_NB + j + c] -= acc[r][c];
    }
}

/* Solve (H + lambda I) x = b.  g->x receives x in the natural order [points | se3].  Returns 0 if not positive definite. */
static int blk_solve(ba_t *g, blk_t *B, double lambda) {
  const int P = B->P, C = B->C, T = B->T, nth = B->nthreads;
  const size_t TS = (size_t)BLK_NB * BLK_NB;
  int fail = 0;
  double t0 = now_s();
  memset(B->S, 0, sizeof(double) * (size_t)B->toff[T] * TS);
  memset(B->r, 0, sizeof(double) * (size_t)T * BLK_NB);
  /* H_pp diagonal blocks, se3-se3 blocks, rhs; padding rows of the last tile get a unit diagonal */
  for (int v = 0; v < C; ++v) {
    int a = B->pos[v], I = a / BLK_VT, o = (a % BLK_VT) * 6;
    double *Tl = blk_tile(B, I, I);
    for (int r = 0; r < 6; ++r) {
      for (int c = 0; c < 6; ++c) Tl[(o + r) * BLK_NB + o + c] = B->Hpp[36 * (size_t)v + 6 * r + c];
      Tl[(o + r) * BLK_NB + o + r] += lambda;
      B->r[(size_t)a * 6 + r] = g->b[3 * (size_t)P + 6 * (size_t)v + r];
    }
  }
  for (int a = C; a < T * BLK_VT; ++a) {
    int I = a / BLK_VT, o = (a % BLK_VT) * 6;
    double *Tl = blk_tile(B, I, I);
    for (int r = 0; r < 6; ++r) Tl[(o + r) * BLK_NB + o + r] = 1.0;
  }
  for (int e = 0; e < g->n_se3e; ++e) {
    int i = g->se3e_ij[2 * e], j = g->se3e_ij[2 * e + 1], a = B->pos[i], b = B->pos[j];
    const double *H = B->Hse3e + 36 * (size_t)e;            /* rows i, cols j */
    if (a > b) {
      double *Tl = blk_tile(B, a / BLK_VT, b / BLK_VT); int ro = (a % BLK_VT) * 6, co = (b % BLK_VT) * 6;
      for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Tl[(ro + r) * BLK_NB + co + c] += H[6 * r + c];
    } else {
      double *Tl = blk_tile(B, b / BLK_VT, a / BLK_VT); int ro = (b % BLK_VT) * 6, co = (a % BLK_VT) * 6;
      for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Tl[(ro + r) * BLK_NB + co + c] += H[6 * c + r];
    }
  }
  /* Schur updates of all point components, in batches: compute in parallel, scatter in component order by tile-row owner */
  const int mm = 3 * B->max_m, kk = 6 * B->max_k;
  const int BATCH = 16384;
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (BATCH + 1));
  size_t cap = 0; double *cbuf = NULL;
  double **sA = (double **)malloc(sizeof(double *) * nth), **sY = (double **)malloc(sizeof(double *) * nth), **syb = (double **)malloc(sizeof(double *) * nth);
  int **sl = (int **)malloc(sizeof(int *) * nth);
  for (int t = 0; t < nth; ++t) {
    sA[t] = (double *)malloc(sizeof(double) * ((size_t)mm * mm + 1)); sY[t] = (double *)malloc(sizeof(double) * ((size_t)mm * kk + 1));
    syb[t] = (double *)malloc(sizeof(double) * (mm + 1)); sl[t] = (int *)malloc(sizeof(int) * (C + 1));
  }
  for (int c0 = 0; c0 < B->ncomp && !fail; c0 += BATCH) {
    int c1 = c0 + BATCH < B->ncomp ? c0 + BATCH : B->ncomp;
    int64_t tot = 0;
    for (int c = c0; c < c1; ++c) { int64_t n6 = 6 * (int64_t)(B->cse3_ptr[c + 1] - B->cse3_ptr[c]); off[c - c0] = tot; tot += n6 * n6 + n6; }
    if ((size_t)tot > cap) { free(cbuf); cap = (size_t)tot; cbuf = (double *)malloc(sizeof(double) * (cap + 1)); }
#pragma omp parallel num_threads(nth)
    {
      int tid = omp_get_thread_num();
#pragma omp for schedule(dynamic, 64)
      for (int c = c0; c < c1; ++c) {
        int k = B->cse3_ptr[c + 1] - B->cse3_ptr[c], n6 = 6 * k, n3 = 3 * (B->comp_ptr[c + 1] - B->comp_ptr[c]);
        if (k == 0) continue;
        if (!blk_comp_factor(g, B, c, lambda, sA[tid], sY[tid], syb[tid], sl[tid])) {
#pragma omp atomic write
          fail = 1;
          continue;
        }
        double *rc = cbuf + off[c - c0], *Sc = rc + n6;
        const double *Y = sY[tid], *yb = syb[tid];
        for (int a = 0; a < n6; ++a) {
          double s = 0;
          for (int i = 0; i < n3; ++i) s += Y[(size_t)i * n6 + a] * yb[i];
          rc[a] = s;
        }
        memset(Sc, 0, sizeof(double) * (size_t)n6 * n6);
        for (int i = 0; i < n3; ++i) {
          const double *y = Y + (size_t)i * n6;
          for (int a = 0; a < n6; ++a) {
            double ya = y[a];
            if (ya == 0) continue;
            double *row = Sc + (size_t)a * n6;
            for (int b = 0; b <= (a / 6) * 6 + 5; ++b) row[b] += ya * y[b];
          }
        }
      }
      /* implicit barrier; scatter: thread tid owns tile rows I with I % nthreads == tid */
      int nt = omp_get_num_threads();
      for (int c = c0; c < c1 && !fail; ++c) {
        int s0 = B->cse3_ptr[c], k = B->cse3_ptr[c + 1] - s0, n6 = 6 * k;
        const double *rc = cbuf + off[c - c0], *Sc = rc + n6;
        for (int a = 0; a < k; ++a) {
          int pa = B->pos[B->cse3[s0 + a]], I = pa / BLK_VT;
          if (I % nt != tid) continue;
          int ro = (pa % BLK_VT) * 6;
          for (int r = 0; r < 6; ++r) B->r[(size_t)pa * 6 + r] -= rc[6 * a + r];
          for (int b = 0; b <= a; ++b) {
            int pb = B->pos[B->cse3[s0 + b]], co = (pb % BLK_VT) * 6;
            double *Tl = blk_tile(B, I, pb / BLK_VT);
            for (int r = 0; r < 6; ++r) for (int cc = 0; cc < 6; ++cc) Tl[(ro + r) * BLK_NB + co + cc] -= Sc[(size_t)(6 * a + r) * n6 + 6 * b + cc];
          }
        }
      }
    }
  }
  free(off); free(cbuf);
  double t1 = now_s(); B->t_schur += t1 - t0;
  /* tiled right-looking Cholesky of the banded reduced matrix */
  if (!fail) {
    double **sB = (double **)malloc(sizeof(double *) * nth);
    for (int t = 0; t < nth; ++t) sB[t] = (double *)malloc(sizeof(double) * TS);
#pragma omp parallel num_threads(nth)
    {
      int tid = omp_get_thread_num();
      for (int J = 0; J < T; ++J) {
        int hiJ = B->hi[J], nI = hiJ - J;
#pragma omp single
        { if (!fail && !tile_potrf(blk_tile(B, J, J))) fail = 1; }
        if (fail) break;                              /* every thread sees the same value after the single's barrier */
#pragma omp for schedule(dynamic, 1)
        for (int I = J + 1; I <= hiJ; ++I) tile_trsm(blk_tile(B, I, J), blk_tile(B, J, J));
        int npair = nI * (nI + 1) / 2;
#pragma omp for schedule(dynamic, 1)
        for (int q = 0; q < npair; ++q) {
          int a = 0, rem = q;                          /* q -> (K, I) with J < K <= I <= hiJ, column-major over K */
          while (rem >= nI - a) { rem -= nI - a; ++a; }
          int K = J + 1 + a, I = K + rem;
          tile_gemm_nt(blk_tile(B, I, K), blk_tile(B, I, J), blk_tile(B, K, J), sB[tid]);
        }
      }
    }
    for (int t = 0; t < nth; ++t) free(sB[t]);
    free(sB);
  }
  double t2 = now_s(); B->t_chol += t2 - t1;
  if (!fail) {
    /* L y = r, L^T x = y (tile by tile) */
    double *r = B->r;
    for (int J = 0; J < T; ++J) {
      const double *L = blk_tile(B, J, J);
      double *y = r + (size_t)J * BLK_NB;
      for (int i = 0; i < BLK_NB; ++i) { double s = y[i]; for (int k = 0; k < i; ++k) s -= L[i * BLK_NB + k] * y[k]; y[i] = s / L[i * BLK_NB + i]; }
      for (int I = J + 1; I <= B->hi[J]; ++I) {
        const double *M = blk_tile(B, I, J); double *z = r + (size_t)I * BLK_NB;
        for (int i = 0; i < BLK_NB; ++i) { double s = 0; for (int k = 0; k < BLK_NB; ++k) s += M[i * BLK_NB + k] * y[k]; z[i] -= s; }
      }
    }
    for (int J = T - 1; J >= 0; --J) {
      double *y = r + (size_t)J * BLK_NB;
      for (int I = J + 1; I <= B->hi[J]; ++I) {
        const double *M = blk_tile(B, I, J); const double *z = r + (size_t)I * BLK_NB;
        for (int i = 0; i < BLK_NB; ++i) { double zi = z[i]; for (int k = 0; k < BLK_NB; ++k) y[k] -= M[i * BLK_NB + k] * zi; }
      }
      const double *L = blk_tile(B, J, J);
      for (int i = BLK_NB - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < BLK_NB; ++k) s -= L[k * BLK_NB + i] * y[k]; y[i] = s / L[i * BLK_NB + i]; }
    }
    for (int v = 0; v < C; ++v) for (int d = 0; d < 6; ++d) g->x[3 * (size_t)P + 6 * (size_t)v + d] = r[(size_t)B->pos[v] * 6 + d];
    /* points: x_l = L^-T (yb - Y x_p) */
#pragma omp parallel num_threads(nth)
    {
      int tid = omp_get_thread_num();
      double *xp = (double *)malloc(sizeof(double) * (kk + 1));
#pragma omp for schedule(dynamic, 64)
      for (int c = 0; c < B->ncomp; ++c) {
        int p0 = B->comp_ptr[c], m = B->comp_ptr[c + 1] - p0, n3 = 3 * m, s0 = B->cse3_ptr[c], k = B->cse3_ptr[c + 1] - s0, n6 = 6 * k;
        double *A = sA[tid], *Y = sY[tid], *yb = syb[tid];
        if (!blk_comp_factor(g, B, c, lambda, A, Y, yb, sl[tid])) continue;     /* cannot fail: it succeeded above */
        for (int a = 0; a < k; ++a) for (int d = 0; d < 6; ++d) xp[6 * a + d] = g->x[3 * (size_t)P + 6 * (size_t)B->cse3[s0 + a] + d];
        for (int i = 0; i < n3; ++i) { double s = yb[i]; const double *y = Y + (size_t)i * n6; for (int a = 0; a < n6; ++a) s -= y[a] * xp[a]; yb[i] = s; }
        for (int i = n3 - 1; i >= 0; --i) { double s = yb[i]; for (int j = i + 1; j < n3; ++j) s -= A[j * n3 + i] * yb[j]; yb[i] = s / A[i * n3 + i]; }
        for (int q = 0; q < m; ++q) for (int d = 0; d < 3; ++d) g->x[3 * (size_t)B->comp_pts[p0 + q] + d] = yb[3 * q + d];
      }
      free(xp);
    }
  }
  for (int t = 0; t < nth; ++t) { free(sA[t]); free(sY[t]); free(syb[t]); free(sl[t]); }
  free(sA); free(sY); free(syb); free(sl);
  B->t_back += now_s() - t2;
  return !fail;
}

#endif
