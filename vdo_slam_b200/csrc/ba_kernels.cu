// ba_kernels.cu -- sm_100a kernels of the batch factor-graph path and the CUDA implementation of BaBackend.
//
// All arithmetic is fp64 (g2o runs in double).  The path is HBM/L2-bound stream-gather-reduce work over
// edge streams, so the kernels are organised around coalesced / bulk-copied edge streams and shuffle reductions, not tensor
// cores.  Two layouts of the landmark / edge side share this backend (chosen per graph by BaGraph::finalize, d.tiled):
//   * tiled (default, ba_tile_kernels.cuh): one CTA per tile of whole tracklets, TMA-staged landmark blocks, landmark-side
//     sums in shared memory, se3-vertex-side sums in the world frame over vertex-sorted warp segments; edges stored once
//   * chunked (this file; graphs whose tracklets do not fit a tile): one thread per tracklet on the landmark-major stream
//     (k_lin_tracklets, k_schur_*), one CTA per <=512-edge chunk of a vertex's own copy of the edge stream (k_vertex_sym,
//     k_schur_vertex: 21+6 or 6 per-thread accumulators, warp-shuffle tree + one smem hop, <=42 atomics per chunk)
// Common to both: the reduced system is applied matrix-free inside a PCG whose preconditioner is block-tridiagonal along the
// se3-se3 edge chains and solved by parallel cyclic reduction (one thread-block cluster per chain); 8 PCG iterations are one
// CUDA-graph launch.  Kernel bodies live in ba_bodies.cuh / ba_tiles.cuh (shared with the serial emulation under tests/emul).
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <cstdio>
#include <cstring>

#include <dlfcn.h>
#include <nccl.h>

#include <map>
#include <set>
#include <utility>

#include "ba_bodies.cuh"

namespace cg = cooperative_groups;

namespace vdo {

constexpr int PCR_CL = 8;   // CTAs per thread-block cluster working on one LONG chain of the preconditioner (short chains: one CTA)

#define CK(x)                                                                                       \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "[vdo_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); } \
  } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; result valid in thread 0.  smem must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}

// deterministic sum of a[0..n) by one CTA (fixed per-thread strides, fixed tree): same inputs => same bits on every rank.
// All threads receive the result.  smem must hold >= 33 doubles.
__device__ __forceinline__ double det_sum(const double* __restrict__ a, int n, double* smem) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += a[i];
  v = block_sum(v, smem);
  if (threadIdx.x == 0) smem[32] = v;
  __syncthreads();
  return smem[32];
}

// ---------------------------------------------------------------------------------------------------------------
template <bool WRITE>
__global__ void __launch_bounds__(128) k_lin_tracklets(BaDev d) {
  __shared__ double red[32];
  const int t = d.Tstat + blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0;
  if (t < d.T) chi = body_lin_tracklet(d, t, WRITE);
  chi = block_sum(chi, red);
  if (threadIdx.x == 0 && chi != 0.0) atomicAdd(d.scal + SC_CHI2, chi);
}
template <bool WRITE>
__global__ void __launch_bounds__(256) k_lin_static(BaDev d) {
  __shared__ double red[32];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0;
  if (k < d.Tstat) chi = body_lin_static(d, k, WRITE);
  chi = block_sum(chi, red);
  if (threadIdx.x == 0 && chi != 0.0) atomicAdd(d.scal + SC_CHI2, chi);
}

// reduce NV per-thread values over the CTA into smem out[NV]
template <int NV>
__device__ __forceinline__ void block_reduce_vec(double* acc, double* smem /*[nwarps*NV]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double v = warp_sum(acc[i]);
    if (lane == 0) smem[w * NV + i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int k = 0; k < nw; ++k) s += smem[k * NV + threadIdx.x];
    smem[threadIdx.x] = s;   // safe: thread i only reads column i of every row, and row 0 column i is its own
  }
  __syncthreads();
}

// MODE 0: linearise (Hpp += A, bp += g) ; MODE 1: preconditioner (Minv -= A) ;  OBS: pointxyz vs ternary stream
template <int MODE, bool OBS>
__global__ void __launch_bounds__(128) k_vertex_sym(BaDev d) {
  __shared__ double sm[4 * 27];
  const Chunk ch = OBS ? d.obs_chunks[blockIdx.x] : d.ter_chunks[blockIdx.x];
  Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
  double acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.0;
  for (int e = ch.begin + threadIdx.x; e < ch.end; e += blockDim.x) {
    if (MODE == 0) { if (OBS) body_lin_vertex_obs(d, T, e, acc, acc + 21); else body_lin_vertex_ter(d, T, e, acc, acc + 21); }
    else { if (OBS) body_precond_vertex_obs(d, T, e, acc); else body_precond_vertex_ter(d, T, e, acc); }
  }
  block_reduce_vec<27>(acc, sm);
  if (threadIdx.x < 36) {
    const int r = threadIdx.x / 6, c = threadIdx.x % 6;
    const double v = sm[r <= c ? sym6_idx(r, c) : sym6_idx(c, r)];
    if (MODE == 0) atomicAdd(d.Hpp + 36 * (size_t)ch.v + threadIdx.x, v);
    else atomicAdd(d.Minv + 36 * (size_t)ch.v + threadIdx.x, -v);
  } else if (MODE == 0 && threadIdx.x < 42) {
    atomicAdd(d.bp + 6 * (size_t)ch.v + (threadIdx.x - 36), sm[21 + threadIdx.x - 36]);
  }
}

template <bool OBS>
__global__ void __launch_bounds__(128) k_schur_vertex(BaDev d, double sign, double* __restrict__ out, int check_done) {
  __shared__ double sm[4 * 6];
  if (check_done && d.scal[SC_DONE] != 0.0) return;
  const Chunk ch = OBS ? d.obs_chunks[blockIdx.x] : d.ter_chunks[blockIdx.x];
  Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int e = ch.begin + threadIdx.x; e < ch.end; e += blockDim.x) {
    if (OBS) body_schur_vertex_obs(d, T, e, acc); else body_schur_vertex_ter(d, T, e, acc);
  }
  block_reduce_vec<6>(acc, sm);
  if (threadIdx.x < 6) atomicAdd(out + 6 * (size_t)ch.v + threadIdx.x, sign * sm[threadIdx.x]);
}

template <bool WRITE>
__global__ void __launch_bounds__(64) k_lin_se3_edges(BaDev d) {
  __shared__ double red[32];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0;
  if (e < d.Ese) {
    double Hi[36], Hj[36], Ho[36], gi[6], gj[6];
    const bool binary = body_se3_edge(d, e, WRITE, chi, Hi, Hj, Ho, gi, gj);
    if (!d.own) chi = 0.0;      // sharded graphs: the se3-se3 edges are accumulated by rank 0 only (every rank keeps J_i^T W J_j)
    if (WRITE) {
      const int i = d.se_i[e];
      if (d.own) {
        for (int k = 0; k < 36; ++k) atomicAdd(d.Hpp + 36 * (size_t)i + k, Hi[k]);
        for (int k = 0; k < 6; ++k) atomicAdd(d.bp + 6 * (size_t)i + k, gi[k]);
      }
      if (binary) {
        const int j = d.se_j[e];
        for (int k = 0; k < 36; ++k) { if (d.own) atomicAdd(d.Hpp + 36 * (size_t)j + k, Hj[k]); d.se_Hoff[36 * (size_t)e + k] = Ho[k]; }
        if (d.own) for (int k = 0; k < 6; ++k) atomicAdd(d.bp + 6 * (size_t)j + k, gj[k]);
      }
    }
  }
  chi = block_sum(chi, red);
  if (threadIdx.x == 0 && chi != 0.0) atomicAdd(d.scal + SC_CHI2, chi);
}

__global__ void __launch_bounds__(256) k_max_diagonal(BaDev d) {
  __shared__ double red[32];
  const int n1 = d.C * 6, n = n1 + d.P;
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    m = fmax(m, i < n1 ? fabs(d.Hpp[36 * (size_t)(i / 6) + 7 * (i % 6)]) : fabs(d.hll[i - n1]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmax(m, red[k]);
    atomicMax(reinterpret_cast<unsigned long long*>(d.scal + SC_MAXDIAG), (unsigned long long)__double_as_longlong(m));  // m >= 0
  }
}

__global__ void __launch_bounds__(128) k_factor_landmarks(BaDev d, double lambda) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < d.Tstat) body_factor_static(d, t, lambda);
  else if (t < d.T) body_factor_tracklet(d, t, lambda);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_schur_landmarks(BaDev d, const double* __restrict__ v, double* __restrict__ out) {
  if (MODE == 1 && d.scal[SC_DONE] != 0.0) return;
  const int t = d.Tstat + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < d.T) body_schur_tracklet(d, t, MODE, v, out);
}
// Chains (dynamic tracklets): 8 lanes cooperate on one tracklet.  Each lane owns one landmark of the current 8-landmark
// segment and does that landmark's gathers (pointxyz edges through the per-vertex world-frame vectors, its ternary edge through the
// motion pose) in parallel with its neighbours; only the 3-vector recursions y_k = u_k + f_{k-1} R_{k-1} y_{k-1} (forward) and
// z_k = (y_k + omega_k R_k^T z_{k+1}) / s_k (backward) walk the lanes, one shuffle of 3 doubles per step.
// mode 0: out = Hll^-1 bl ; mode 1: out = Hll^-1 (Hlp v) ; mode 2: out = Hll^-1 (bl - Hlp v)
template <int MODE>
__global__ void __launch_bounds__(128) k_schur_chains8(BaDev d, const double* __restrict__ v, double* __restrict__ out) {
  if (MODE == 1 && d.scal[SC_DONE] != 0.0) return;
  const int lane8 = threadIdx.x & 7;
  const int t = d.Tstat + ((blockIdx.x * blockDim.x + threadIdx.x) >> 3);
  const bool live = t < d.T;
  const int kb = live ? d.tk_begin[t] : 0, ke = live ? d.tk_begin[t + 1] : 0;
  const unsigned FULL = 0xffffffffu;
  double seg_w[3] = {0, 0, 0};       // f_{k-1} R_{k-1} y_{k-1} entering the segment
  double seg_c[3] = {0, 0, 0};       // ternary contribution of edge (k-1, k) to u_k entering the segment
  // ---------------- forward ----------------
  for (int base = kb; __any_sync(FULL, base < ke); base += 8) {
    const int k = base + lane8;
    const bool valid = k < ke;
    double u[3] = {0, 0, 0}, cn[3] = {0, 0, 0}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double f = 0.0;
    if (valid) {
      const int h = d.tk_h[k];
      const double om = d.tk_omega[k];
      f = om / d.pt_s[k];
      if (MODE != 0) {
        const double p[3] = {d.pt[3 * (size_t)k], d.pt[3 * (size_t)k + 1], d.pt[3 * (size_t)k + 2]};
        for (int e = d.lm_obs_begin[k]; e < d.lm_obs_begin[k + 1]; ++e) {
          const double* w = d.vw + 6 * (size_t)d.lm_cam[e];
          double pxb[3]; cross3(p, w + 3, pxb);
          const double oe = d.lm_omega[e];
          u[0] += oe * (w[0] + 2 * pxb[0]); u[1] += oe * (w[1] + 2 * pxb[1]); u[2] += oe * (w[2] + 2 * pxb[2]);
        }
      }
      if (h >= 0) {
        Iso H; iso_load(d.se3 + 12 * (size_t)h, H);
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = H.R[i];
        if (MODE != 0) {
          const double pn[3] = {d.pt[3 * (size_t)(k + 1)], d.pt[3 * (size_t)(k + 1) + 1], d.pt[3 * (size_t)(k + 1) + 2]};
          double q[3]; iso_inv_apply(H, pn, q);
          double a[3]; ter_Jh_mul(q, v + 6 * (size_t)h, a);
          u[0] += om * a[0]; u[1] += om * a[1]; u[2] += om * a[2];
          double Ra[3]; rot_apply(H.R, a, Ra);
          cn[0] = -om * Ra[0]; cn[1] = -om * Ra[1]; cn[2] = -om * Ra[2];
        }
      }
    }
    // ternary contribution from the previous landmark's edge
    double cp[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { cp[i] = __shfl_up_sync(FULL, cn[i], 1, 8); if (lane8 == 0) cp[i] = seg_c[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { u[i] += cp[i]; seg_c[i] = __shfl_sync(FULL, cn[i], 7, 8); }
    double y[3];
    if (valid) {
      if (MODE == 0) { y[0] = d.bl[3 * (size_t)k]; y[1] = d.bl[3 * (size_t)k + 1]; y[2] = d.bl[3 * (size_t)k + 2]; }
      else if (MODE == 1) { y[0] = u[0]; y[1] = u[1]; y[2] = u[2]; }
      else { y[0] = d.bl[3 * (size_t)k] - u[0]; y[1] = d.bl[3 * (size_t)k + 1] - u[1]; y[2] = d.bl[3 * (size_t)k + 2] - u[2]; }
    } else { y[0] = y[1] = y[2] = 0; }
    double wout[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double win[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) win[i] = __shfl_up_sync(FULL, wout[i], 1, 8);
      if (lane8 == j) {
        if (j == 0) { win[0] = seg_w[0]; win[1] = seg_w[1]; win[2] = seg_w[2]; }
        y[0] += win[0]; y[1] += win[1]; y[2] += win[2];
        double Ry[3]; rot_apply(R, y, Ry);
        wout[0] = f * Ry[0]; wout[1] = f * Ry[1]; wout[2] = f * Ry[2];
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) seg_w[i] = __shfl_sync(FULL, wout[i], 7, 8);
    if (valid) { out[3 * (size_t)k] = y[0]; out[3 * (size_t)k + 1] = y[1]; out[3 * (size_t)k + 2] = y[2]; }
  }
  // ---------------- backward ----------------
  double seg_z[3] = {0, 0, 0};       // z of the first landmark of the following segment
  const int nseg = (ke - kb + 7) >> 3;
  int max_seg = nseg;
  { const int a = __shfl_xor_sync(FULL, max_seg, 8); if (a > max_seg) max_seg = a; }      // uniform trip count per warp (4 groups of 8 lanes)
  { const int a = __shfl_xor_sync(FULL, max_seg, 16); if (a > max_seg) max_seg = a; }
  for (int sgi = max_seg - 1; sgi >= 0; --sgi) {
    const int k = kb + 8 * sgi + lane8;
    const bool valid = sgi < nseg && k < ke;
    double y[3] = {0, 0, 0}, Rt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, om = 0.0, is = 0.0;
    bool last = true;
    if (valid) {
      y[0] = out[3 * (size_t)k]; y[1] = out[3 * (size_t)k + 1]; y[2] = out[3 * (size_t)k + 2];
      is = 1.0 / d.pt_s[k];
      const int h = d.tk_h[k];
      if (h >= 0 && k + 1 < ke) {
        last = false; om = d.tk_omega[k];
        const double* Rp = d.se3 + 12 * (size_t)h;
#pragma unroll
        for (int i = 0; i < 9; ++i) Rt[i] = Rp[i];
      }
    }
    double z[3] = {0, 0, 0};
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      double zn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) zn[i] = __shfl_down_sync(FULL, z[i], 1, 8);
      if (lane8 == j && valid) {
        if (j == 7) { zn[0] = seg_z[0]; zn[1] = seg_z[1]; zn[2] = seg_z[2]; }
        if (!last) {
          double t3[3]; rot_t_apply(Rt, zn, t3);
          y[0] += om * t3[0]; y[1] += om * t3[1]; y[2] += om * t3[2];
        }
        z[0] = y[0] * is; z[1] = y[1] * is; z[2] = y[2] * is;
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) seg_z[i] = __shfl_sync(FULL, z[i], 0, 8);
    if (valid) { out[3 * (size_t)k] = z[0]; out[3 * (size_t)k + 1] = z[1]; out[3 * (size_t)k + 2] = z[2]; }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_schur_static(BaDev d, double* __restrict__ out) {
  if (MODE == 1 && d.scal[SC_DONE] != 0.0) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < d.Tstat) body_schur_static(d, k, MODE, out);
}

__global__ void __launch_bounds__(128) k_precond_begin(BaDev d, double lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d.C * 36) { const int k = i % 36; d.Minv[i] = d.own ? d.Hpp[i] + ((k % 7) == 0 ? lambda : 0.0) : 0.0; }
}

__global__ void k_set_scalars(BaDev d, double lambda, double tol2) { d.scal[SC_LAMBDA] = lambda; d.scal[SC_TOL2] = tol2; }
__global__ void __launch_bounds__(128) k_vertex_transform(BaDev d, const double* __restrict__ x) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < d.C) body_vertex_transform(d, v, x, d.vw);
}
__global__ void __launch_bounds__(128) k_hpp_mul(BaDev d, const double* __restrict__ x, double* __restrict__ out) {
  if (d.scal[SC_DONE] != 0.0) return;
  const double lambda = d.scal[SC_LAMBDA];
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < d.C) body_hpp_mul(d, v, lambda, x, out);
}

// ---- parallel cyclic reduction: one thread-block CLUSTER (PCR_CL CTAs) per chain, cluster.sync() between levels ----
// CL = CTAs per cluster: PCR_CL for long paths, 1 (plain CTA, the cluster barrier degenerates to a CTA barrier) for paths of at most
// PCR_SHORT vertices -- most paths are short (objects seen for a few frames, single motion vertices) and a cluster of 8 CTAs each would only
// multiply the number of waves the launch needs.  path0: position of the launch's first path in own_paths (long paths first).

// In-place Gauss-Jordan inverse of an SPD 6x6 whose row i lives in lane i of an aligned 8-lane group (lanes 6, 7 of the group and groups
// without a vertex carry the identity; every lane of the warp executes the shuffles).  The pivots are those of the LDL^T factorisation:
// a non-positive (or non-finite) one reports "not SPD" exactly where the Cholesky of body_pcr_invert does.
__device__ __forceinline__ bool gj6_rows(double (&a)[6], int i) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double rk[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) rk[c] = __shfl_sync(0xffffffffu, a[c], k, 8);
    const double piv = rk[k];
    if (!(piv > 0.0) || !(piv < 1e300)) ok = false;
    const double pinv = 1.0 / piv;
    if (i == k) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a[c] = (c == k) ? pinv : rk[c] * pinv;
    } else {
      const double f = a[k] * pinv;
#pragma unroll
      for (int c = 0; c < 6; ++c) a[c] = (c == k) ? -f : a[c] - f * rk[c];
    }
  }
  return ok;
}
// Factorisation of the block-tridiagonal preconditioner of one se3 path by parallel cyclic reduction: ONE pass and one cluster barrier per
// level (the first version ran three passes of (vertex, row, column) items per level, each behind a barrier and a round trip through L2:
// 0.55 ms for the 1000-vertex camera path, 12 % of a solve).  An 8-lane group owns a vertex, lane i its row i of every block:
//   A_v = -L_v Dinv_{v-s},  G_v = -L_{v+s}^T Dinv_{v+s},  D'_v = D_v + A_v L_v^T + G_v L_{v+s},  L'_v = A_v L_{v-s},  Dinv'_v = (D'_v)^-1
// all from level-l data of v and v +- s (Dinv' is formed by the group at the end of the level, in registers, with shuffles), so the only
// exchange between groups is the barrier that closes the level.  D is updated in place (a group reads only its own rows); L and Dinv are
// double buffered (Dinv: pcr_Dinv / second half of pcr_D).  The arrays are read with plain loads: other CTAs of the cluster wrote them.
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 2) k_pcr_factor(BaDev d, double lambda, int path0) {
  cg::cluster_group cl = cg::this_cluster();
  const int path = d.own_paths[path0 + blockIdx.x / CL];
  const int pb = d.path_begin[path], pe = d.path_begin[path + 1];
  const int nl = pcr_num_levels(pe - pb);
  const int tid = cl.block_rank() * blockDim.x + threadIdx.x, nth = CL * blockDim.x;
  const int grp = tid >> 3, i = tid & 7, ngrp = nth >> 3;
  const size_t N36 = 36 * (size_t)d.C;
  __shared__ double sblk[32 * 180];                          // per 8-lane group: the five neighbour blocks of the current round
  double* const Dm = d.pcr_D;                                // level matrix D, in place
  double* const DI[2] = {d.pcr_Dinv, d.pcr_D + N36};
  int bad = 0;
  const int n_round = (pe - pb + ngrp - 1) / ngrp;
  // identity rows for the lanes without work (keeps gj6_rows finite)
  auto idle_row = [&](double (&a)[6]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = (c == (i % 6)) ? 1.0 : 0.0;
  };
  auto finish = [&](double (&a)[6], bool act, int v, double* out) {      // a = row i of D'_v on entry; writes row i of its inverse
    double dii = 1.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) if (act && c == i) dii = a[c];
    const bool ok = gj6_rows(a, i);
    if (!act) return;
    if (!ok) {
#pragma unroll
      for (int c = 0; c < 6; ++c) a[c] = (c == i) ? 1.0 / (fabs(dii) + lambda) : 0.0;
      if (i == 0) bad = 1;
    }
    double* o = out + 36 * (size_t)v + 6 * i;
#pragma unroll
    for (int c = 0; c < 6; ++c) o[c] = a[c];
  };
  // level 0: D = assembled diagonal block (in Minv), L = M(v, v-1) from the se3-se3 edge block; Dinv_0
  for (int r = 0; r < n_round; ++r) {
    const int v = pb + r * ngrp + grp;
    const bool act = v < pe && i < 6;
    double a[6];
    idle_row(a);
    if (act) {
      const double* S = d.Minv + 36 * (size_t)v + 6 * i;
      double* Dv = Dm + 36 * (size_t)v + 6 * i;
#pragma unroll
      for (int c = 0; c < 6; ++c) { a[c] = S[c]; Dv[c] = a[c]; }
      const int e = d.pcr_edge[v];
      double* Lv = d.pcr_L + 36 * (size_t)v + 6 * i;
      if (e < 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) Lv[c] = 0.0;
      } else {
        const double* B = d.se_Hoff + 36 * (size_t)e;
        const bool tr = d.pcr_tr[v] != 0;
#pragma unroll
        for (int c = 0; c < 6; ++c) Lv[c] = tr ? B[6 * c + i] : B[6 * i + c];
      }
    }
    finish(a, act, v, nl > 0 ? DI[0] : d.Minv);
  }
  if (nl > 0) cl.sync();
  int cur = 0;
  for (int l = 0; l < nl; ++l) {
    const double* L = d.pcr_L + cur * N36; double* Ln = d.pcr_L + (1 - cur) * N36;
    const double* Di = DI[cur]; double* Dout = (l + 1 < nl) ? DI[1 - cur] : d.Minv;
    double *A = d.pcr_A + l * N36, *G = d.pcr_G + l * N36;
    const int s = 1 << l;
    for (int r = 0; r < n_round; ++r) {
      const int v = pb + r * ngrp + grp;
      const bool has_v = v < pe, act = has_v && i < 6;
      // stage the five blocks the group needs (L_v, Dinv_{v-s}, L_{v-s}, L_{v+s}, Dinv_{v+s}; zeros outside the path) with all loads in flight
      // at once: read in place they cost one L2 round trip per block, serialised behind the branches (4 us per round, measured)
      double* blk = sblk + 180 * (threadIdx.x >> 3);
      __syncwarp();
      if (has_v) {
        const bool hm = v - s >= pb, hp = v + s < pe, hmm = v - 2 * s >= pb;
        const double *s0 = L + 36 * (size_t)v, *s1 = Di + 36 * (size_t)(hm ? v - s : v), *s2 = L + 36 * (size_t)(hm ? v - s : v), *s3 = L + 36 * (size_t)(hp ? v + s : v),
                     *s4 = Di + 36 * (size_t)(hp ? v + s : v);
        double tmp[23];
#pragma unroll
        for (int u = 0; u < 23; ++u) {
          const int e = (threadIdx.x & 7) + 8 * u, bq = e / 36, o = e - 36 * bq;
          const double* sp = bq == 0 ? s0 : bq == 1 ? s1 : bq == 2 ? s2 : bq == 3 ? s3 : s4;
          const bool on = bq == 0 ? true : bq == 1 ? hm : bq == 2 ? hmm : hp;
          tmp[u] = (e < 180 && on) ? sp[o] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 23; ++u) { const int e = (threadIdx.x & 7) + 8 * u; if (e < 180) blk[e] = tmp[u]; }
      }
      __syncwarp();
      double dn[6];
      idle_row(dn);
      if (act) {
        const double *Lv = blk, *Dim = blk + 36, *Lm = blk + 72, *Lp = blk + 108, *Dip = blk + 144;
        double av[6] = {0, 0, 0, 0, 0, 0}, gv[6] = {0, 0, 0, 0, 0, 0}, ln[6] = {0, 0, 0, 0, 0, 0};
        double* Dv = Dm + 36 * (size_t)v + 6 * i;
#pragma unroll
        for (int c = 0; c < 6; ++c) dn[c] = Dv[c];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const double lk = Lv[6 * i + k], lp = Lp[6 * k + i];
#pragma unroll
          for (int c = 0; c < 6; ++c) { av[c] -= lk * Dim[6 * k + c]; gv[c] -= lp * Dip[6 * k + c]; }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
          for (int c = 0; c < 6; ++c) { dn[c] += av[k] * Lv[6 * c + k] + gv[k] * Lp[6 * k + c]; ln[c] += av[k] * Lm[6 * k + c]; }
        }
        double* Ao = A + 36 * (size_t)v + 6 * i; double* Go = G + 36 * (size_t)v + 6 * i; double* Lo = Ln + 36 * (size_t)v + 6 * i;
#pragma unroll
        for (int c = 0; c < 6; ++c) { Ao[c] = av[c]; Go[c] = gv[c]; Lo[c] = ln[c]; Dv[c] = dn[c]; }
      }
      finish(dn, act, v, Dout);
    }
    if (l + 1 < nl) cl.sync();
    cur = 1 - cur;
  }
  if (bad) atomicAdd(d.scal + SC_BAD, 1.0);
}

// z = M^-1 r for the cluster's chain (r final for the whole chain on entry); returns this thread's share of r.z.
// Work item = (vertex, row): 6 items per vertex so that A / G rows are read coalesced.
template <int CL>
__device__ __forceinline__ double pcr_solve_path(const BaDev& d, cg::cluster_group& cl, int pb, int pe, const double* __restrict__ r, double* __restrict__ z) {
  const int nl = pcr_num_levels(pe - pb);
  const int tid = cl.block_rank() * blockDim.x + threadIdx.x, nth = CL * blockDim.x;
  const int n_items = 6 * (pe - pb);
  const size_t N6 = 6 * (size_t)d.C, N36 = 36 * (size_t)d.C;
  const double* src = r;
  int cur = 0;
  // The operator rows of level l + 1 do not depend on the vector, so they are fetched into registers BEFORE the cluster
  // barrier that closes level l: after the barrier only the (L2-resident) vector entries b[v - s], b[v + s] are on the critical
  // path.  KP items per thread are prefetched (paths up to KP * 2048 / 6 vertices); longer paths read the rest directly.
  constexpr int KP = 3;
  double pa[KP][6], pg[KP][6];
  auto fetch = [&](int l) {
    const int s = 1 << l;
    const double* A = d.pcr_A + l * N36; const double* G = d.pcr_G + l * N36;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int w = tid + k * nth;
      const int v = pb + w / 6, row = w % 6;
      const bool ia = w < n_items && v - s >= pb, ig = w < n_items && v + s < pe;
      const double* a = A + 36 * (size_t)v + 6 * row; const double* g = G + 36 * (size_t)v + 6 * row;
#pragma unroll
      for (int i = 0; i < 6; ++i) { pa[k][i] = ia ? a[i] : 0.0; pg[k][i] = ig ? g[i] : 0.0; }
    }
  };
  if (nl > 0) fetch(0);
  for (int l = 0; l < nl; ++l) {
    double* dst = d.pcr_b + cur * N6;
    const int s = 1 << l;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int w = tid + k * nth;
      if (w < n_items) {
        const int v = pb + w / 6, row = w % 6;
        double o = src[6 * (size_t)v + row];
        if (v - s >= pb) { const double* x = src + 6 * (size_t)(v - s); o += pa[k][0] * x[0] + pa[k][1] * x[1] + pa[k][2] * x[2] + pa[k][3] * x[3] + pa[k][4] * x[4] + pa[k][5] * x[5]; }
        if (v + s < pe) { const double* x = src + 6 * (size_t)(v + s); o += pg[k][0] * x[0] + pg[k][1] * x[1] + pg[k][2] * x[2] + pg[k][3] * x[3] + pg[k][4] * x[4] + pg[k][5] * x[5]; }
        dst[6 * (size_t)v + row] = o;
      }
    }
    for (int w = tid + KP * nth; w < n_items; w += nth) {
      const int v = pb + w / 6, row = w % 6;
      dst[6 * (size_t)v + row] = pcr_apply_row(v, row, pb, pe, s, d.pcr_A + l * N36, d.pcr_G + l * N36, src);
    }
    if (l + 1 < nl) fetch(l + 1);
    cl.sync();
    src = dst; cur = 1 - cur;
  }
  double rz = 0.0;
  for (int w = tid; w < n_items; w += nth) {
    const int v = pb + w / 6, row = w % 6;
    const double* m = d.Minv + 36 * (size_t)v + 6 * row; const double* x = src + 6 * (size_t)v;
    const double zz = m[0] * x[0] + m[1] * x[1] + m[2] * x[2] + m[3] * x[3] + m[4] * x[4] + m[5] * x[5];
    z[6 * (size_t)v + row] = zz;
    rz += zz * r[6 * (size_t)v + row];
  }
  return rz;
}

// Path-sharded preconditioner: the CTAs of one path publish their part of z (already in this rank's d.z) and their partial of r.z to
// every other rank; the last CTA of the launch to finish fences and raises this rank's flag of the second exchange on every rank.
template <int CL>
__device__ __forceinline__ void xchg_publish_z(const BaDev& d, cg::cluster_group& cl, int path, int pb, int pe, double rz_part, int* is_last_sm, unsigned int total_ctas) {
  const int tid = cl.block_rank() * blockDim.x + threadIdx.x, nth = CL * blockDim.x;
  const int pidx = path * PCR_CL + (int)cl.block_rank();
  if (threadIdx.x == 0) d.part_rz[pidx] = rz_part;
  if (!d.xg_paths) return;
  cl.sync();                                           // every CTA of the cluster wrote its share of z
  for (int r = 0; r < d.xg_world; ++r) {
    if (r == d.xg_rank) continue;
    double* zr = d.xg_slots[r] + d.xg_off_z;
    for (int w = tid; w < 6 * (pe - pb); w += nth) { const size_t q = 6 * (size_t)pb + w; zr[q] = d.z[q]; }
    if (threadIdx.x == 0) (d.xg_slots[r] + d.xg_off_prz)[pidx] = rz_part;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(d.ticket + 2, 1u);
    *is_last_sm = (t == total_ctas - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!*is_last_sm || threadIdx.x != 0) return;
  __threadfence_system();
  d.ticket[2] = 0u;
  const unsigned long long epoch = d.xg_epoch[192] + 1ull;
  d.xg_epoch[192] = epoch;
  for (int r = 0; r < d.xg_world; ++r) { volatile unsigned long long* f = d.xg_flags[r] + 320 + d.xg_rank; *f = epoch; }
}
// thread 0 of a CTA: wait until every rank's z part of the current epoch has arrived.  Returns false after ~2 s (a peer died).
__device__ __forceinline__ bool xchg_wait_z(const BaDev& d) {
  const unsigned long long epoch = d.xg_epoch[192];
  volatile unsigned long long* f = d.xg_flags[d.xg_rank] + 320;
  const long long t0 = clock64();
  for (int r = 0; r < d.xg_world; ++r)
    while (f[r] < epoch) if (clock64() - t0 > 4000000000ll) return false;
  __threadfence_system();
  return true;
}

template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256) k_pcg_init(BaDev d, int path0, unsigned int total_ctas) {
  __shared__ double red[32];
  __shared__ int is_last;
  cg::cluster_group cl = cg::this_cluster();
  const int path = d.own_paths[path0 + blockIdx.x / CL];
  const int pb = d.path_begin[path], pe = d.path_begin[path + 1];
  const int tid = cl.block_rank() * blockDim.x + threadIdx.x, nth = CL * blockDim.x;
  for (int w = tid; w < 6 * (pe - pb); w += nth) { const size_t q = 6 * (size_t)pb + w; d.r[q] = d.rhs[q]; d.xp[q] = 0.0; }
  if (pe - pb > 1) cl.sync();
  double rz = pcr_solve_path<CL>(d, cl, pb, pe, d.r, d.z);
  // p = z: each thread copies exactly the items it produced in the last loop of pcr_solve_path
  for (int w = tid; w < 6 * (pe - pb); w += nth) { const size_t q = 6 * (size_t)pb + w; d.p[q] = d.z[q]; }
  rz = block_sum(rz, red);
  if (threadIdx.x == 0) red[0] = rz;
  __syncthreads();
  xchg_publish_z<CL>(d, cl, path, pb, pe, red[0], &is_last, total_ctas);
}
__global__ void __launch_bounds__(256) k_pcg_init_fin(BaDev d) {
  __shared__ double red[33];
  __shared__ int okw;
  if (d.xg_paths) {
    if (threadIdx.x == 0) okw = xchg_wait_z(d) ? 1 : 0;
    __syncthreads();
    if (!okw) { if (threadIdx.x == 0) d.scal[SC_DONE] = 3.0; return; }
  }
  const double rz = det_sum(d.part_rz, d.n_paths * PCR_CL, red);
  if (threadIdx.x != 0) return;
  d.scal[SC_RZ] = rz; d.scal[SC_RZ0] = rz; d.scal[SC_RZ_NEW] = 0.0; d.scal[SC_PAP] = 0.0; d.scal[SC_ITERS] = 0.0; d.scal[SC_BETA] = 0.0;
  d.scal[SC_DONE] = (rz > 0.0) ? 0.0 : 1.0;
}
__global__ void __launch_bounds__(256) k_pcg_dot(BaDev d) {
  __shared__ double red[32];
  if (d.scal[SC_DONE] != 0.0) return;
  const int n = d.C * 6;
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += d.p[i] * d.Ap[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) d.part_pap[blockIdx.x] = s;
}
// x += alpha p ; r -= alpha Ap ; z = M^-1 r ; rz_new += r.z.  FUSED: p is passed explicitly (double-buffered) and the last CTA to
// finish does the work of k_pcg_step_b's beta and of k_pcg_scalars.
template <bool FUSED, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256) k_pcg_step_a(BaDev d, const double* __restrict__ p, int path0, unsigned int total_ctas) {
  __shared__ double red[33];
  __shared__ int is_last;
  if (d.scal[SC_DONE] != 0.0) return;
  cg::cluster_group cl = cg::this_cluster();
  const double pap = det_sum(d.part_pap, d.n_part_pap, red), rz = d.scal[SC_RZ];
  const double alpha = (pap > 0.0) ? rz / pap : 0.0;
  const int path = d.own_paths[path0 + blockIdx.x / CL];
  const int pb = d.path_begin[path], pe = d.path_begin[path + 1];
  const int tid = cl.block_rank() * blockDim.x + threadIdx.x, nth = CL * blockDim.x;
  for (int w = tid; w < 6 * (pe - pb); w += nth) { const size_t q = 6 * (size_t)pb + w; d.xp[q] += alpha * p[q]; d.r[q] -= alpha * d.Ap[q]; }
  if (pe - pb > 1) cl.sync();
  double rzn = pcr_solve_path<CL>(d, cl, pb, pe, d.r, d.z);
  rzn = block_sum(rzn, red);
  if (threadIdx.x == 0) red[0] = rzn;
  __syncthreads();
  xchg_publish_z<CL>(d, cl, path, pb, pe, red[0], &is_last, total_ctas);     // part_rz (and, path-sharded, z / part_rz on the other ranks)
  if (!FUSED || d.xg_paths) return;                           // path-sharded: k_pcg_scalars_x does the scalars once every part has arrived
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(d.ticket, 1u);
    is_last = (t == total_ctas - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double rz_new = det_sum(d.part_rz, d.n_paths * PCR_CL, red);
  if (threadIdx.x != 0) return;
  *d.ticket = 0u;
  if (!(pap > 0.0) || !isfinite(pap) || !isfinite(rz_new)) { d.scal[SC_DONE] = 2.0; return; }
  d.scal[SC_BETA] = rz_new / rz; d.scal[SC_RZ] = rz_new; d.scal[SC_ITERS] += 1.0;
  if (rz_new <= d.scal[SC_TOL2] * d.scal[SC_RZ0]) d.scal[SC_DONE] = 1.0;
}
// ---- fused PCG iteration (single GPU): 4 dependent launches per iteration instead of 8 ----
//   k_pcg_p_hpp            p_{k+1} = z + beta p_k (out of place, recomputed for the path neighbours), Ap = (Hpp + lambda I) p, vw / vh
//   k_tile_schur2 x 2      Hpl Hll^-1 Hlp p (static and chain tiles, forked)
//   k_tile_finalize_schur2 Ap -= B^T sums, partials of p.Ap
//   k_pcg_step_a<true>     alpha, x, r, z = M^-1 r (PCR), partials of r.z; the LAST CTA to finish sums them (fixed order) and sets beta, rz,
//                          the iteration count and the convergence flag
__global__ void __launch_bounds__(128) k_pcg_p_hpp(BaDev d, const double* __restrict__ p_in, double* __restrict__ p_out, double* __restrict__ out) {
  if (d.scal[SC_DONE] != 0.0) return;
  const double lambda = d.scal[SC_LAMBDA], beta = d.scal[SC_BETA];
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= d.C) return;
  double xv[6], o[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) { xv[r] = d.z[6 * (size_t)v + r] + beta * p_in[6 * (size_t)v + r]; p_out[6 * (size_t)v + r] = xv[r]; }
  const double* H = d.Hpp + 36 * (size_t)v;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s = lambda * xv[r];
#pragma unroll
    for (int c = 0; c < 6; ++c) s += H[6 * r + c] * xv[c];
    o[r] = s;
  }
  for (int n = d.nbr_begin[v]; n < d.nbr_begin[v + 1]; ++n) {
    const double* B = d.se_Hoff + 36 * (size_t)d.nbr_edge[n];
    const size_t u = 6 * (size_t)d.nbr_other[n];
    double xo[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) xo[c] = d.z[u + c] + beta * p_in[u + c];
    if (d.nbr_tr[n]) {
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) o[r] += B[6 * c + r] * xo[c];
    } else {
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) o[r] += B[6 * r + c] * xo[c];
    }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) out[6 * (size_t)v + r] = d.own ? o[r] : 0.0;
  body_vertex_transform(d, v, p_out, d.vw);
}

// ---- sharded PCG iteration: all-reduce of the 6C-vector S*p through peer memory (NVLink), inside the captured graph ----
// Every rank holds, for each sender r, a slot of 6C doubles (double-buffered by the parity of an epoch counter).
//   k_xchg_scatter  per vertex: this rank's partial (Hpp p on rank 0) - B^T (tile sums), stored straight into slot[rank] of EVERY
//                   rank (remote stores); the last CTA to finish fences (system scope), bumps the epoch and stores it into
//                   flag[rank] of every rank
//   k_xchg_reduce   waits until all world flags carry the epoch, then sums the world slots IN RANK ORDER -- every rank adds the
//                   same numbers in the same order, so Ap (and with it every PCG scalar and the convergence flag) is bit-identical
//                   on all ranks -- and forms the CTA's share of p.Ap
// Nothing here is enqueued by the host per iteration: the kernels are ordinary graph nodes.
__global__ void __launch_bounds__(128) k_xchg_scatter(BaDev d, double sign, const double* __restrict__ own_part) {
  __shared__ int is_last;
  if (d.scal[SC_DONE] != 0.0) return;
  const unsigned long long epoch = *d.xg_epoch + 1ull;            // the epoch this vector belongs to (bumped below by the last CTA)
  const size_t n6 = 6 * (size_t)d.C;
  const size_t slot = ((epoch & 1ull) * (size_t)d.xg_world + (size_t)d.xg_rank) * n6;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < d.C) {
    const double* T = d.se3 + 12 * (size_t)v;
    double* a = d.acc6 + 12 * (size_t)v;
    const double F[3] = {a[0] + a[6], a[1] + a[7], a[2] + a[8]};
    const double G[3] = {2 * a[0] + a[6], 2 * a[1] + a[7], 2 * a[2] + a[8]};
    double txg[3]; cross3(T + 9, G, txg);
    const double M[3] = {a[3] + a[9] - txg[0], a[4] + a[10] - txg[1], a[5] + a[11] - txg[2]};
    double o0[3], o1[3];
    rot_t_apply(T, F, o0); rot_t_apply(T, M, o1);
    const double* op = own_part + 6 * (size_t)v;
    const double o[6] = {op[0] + sign * o0[0], op[1] + sign * o0[1], op[2] + sign * o0[2], op[3] + sign * o1[0], op[4] + sign * o1[1], op[5] + sign * o1[2]};
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 0.0;
    for (int r = 0; r < d.xg_world; ++r) {
      double* dst = d.xg_slots[r] + slot + 6 * (size_t)v;
#pragma unroll
      for (int i = 0; i < 6; ++i) dst[i] = o[i];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(d.ticket + 1, 1u);
    is_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence_system();
  d.ticket[1] = 0u;
  *d.xg_epoch = epoch;
  for (int r = 0; r < d.xg_world; ++r) {
    volatile unsigned long long* f = d.xg_flags[r] + d.xg_rank;
    *f = epoch;
  }
}
__global__ void __launch_bounds__(128) k_xchg_reduce(BaDev d, double* __restrict__ out, const double* __restrict__ pdot) {
  __shared__ double red[32];
  __shared__ int failed;
  if (d.scal[SC_DONE] != 0.0) return;
  const unsigned long long epoch = *d.xg_epoch;
  if (threadIdx.x == 0) {
    failed = 0;
    volatile unsigned long long* f = d.xg_flags[d.xg_rank];
    const long long t0 = clock64();
    for (int r = 0; r < d.xg_world; ++r)
      while (f[r] < epoch) {
        if (clock64() - t0 > 4000000000ll) { failed = 1; break; }       // ~2 s: a peer died; do not hang the GPU
      }
    __threadfence_system();
  }
  __syncthreads();
  if (failed) { if (threadIdx.x == 0 && blockIdx.x == 0) d.scal[SC_DONE] = 3.0; return; }
  const size_t n6 = 6 * (size_t)d.C;
  const double* base = d.xg_slots[d.xg_rank] + (epoch & 1ull) * (size_t)d.xg_world * n6;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (v < d.C) {
    double o[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 0; r < d.xg_world; ++r) {
      const volatile double* src = base + (size_t)r * n6 + 6 * (size_t)v;
#pragma unroll
      for (int i = 0; i < 6; ++i) o[i] += src[i];
    }
    const double* pv = pdot + 6 * (size_t)v;
#pragma unroll
    for (int i = 0; i < 6; ++i) { out[6 * (size_t)v + i] = o[i]; s += pv[i] * o[i]; }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) d.part_pap[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_pcg_step_b(BaDev d) {
  __shared__ double red[33];
  if (d.scal[SC_DONE] != 0.0) return;
  const double beta = det_sum(d.part_rz, d.n_paths * PCR_CL, red) / d.scal[SC_RZ];
  const int n = d.C * 6;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d.p[i] = d.z[i] + beta * d.p[i];
}
__global__ void __launch_bounds__(256) k_pcg_scalars(BaDev d) {
  __shared__ double red[33];
  if (d.scal[SC_DONE] != 0.0) return;
  const double tol2 = d.scal[SC_TOL2];
  const double pap = det_sum(d.part_pap, d.n_part_pap, red);
  const double rzn = det_sum(d.part_rz, d.n_paths * PCR_CL, red);
  if (threadIdx.x != 0) return;
  if (!(pap > 0.0) || !isfinite(pap) || !isfinite(rzn)) { d.scal[SC_DONE] = 2.0; return; }
  d.scal[SC_RZ] = rzn; d.scal[SC_RZ_NEW] = 0.0; d.scal[SC_PAP] = 0.0; d.scal[SC_ITERS] += 1.0;
  if (rzn <= tol2 * d.scal[SC_RZ0]) d.scal[SC_DONE] = 1.0;
}

// path-sharded preconditioner: waits for every rank's part of z / r.z, then the scalars of k_pcg_step_a<true>'s last CTA (identical on all ranks)
__global__ void __launch_bounds__(256) k_pcg_scalars_x(BaDev d) {
  __shared__ double red[33];
  __shared__ int okw;
  if (d.scal[SC_DONE] != 0.0) return;
  if (threadIdx.x == 0) okw = xchg_wait_z(d) ? 1 : 0;
  __syncthreads();
  if (!okw) { if (threadIdx.x == 0) d.scal[SC_DONE] = 3.0; return; }
  const double pap = det_sum(d.part_pap, d.n_part_pap, red);
  const double rz_new = det_sum(d.part_rz, d.n_paths * PCR_CL, red);
  if (threadIdx.x != 0) return;
  const double rz = d.scal[SC_RZ];
  if (!(pap > 0.0) || !isfinite(pap) || !isfinite(rz_new)) { d.scal[SC_DONE] = 2.0; return; }
  d.scal[SC_BETA] = rz_new / rz; d.scal[SC_RZ] = rz_new; d.scal[SC_ITERS] += 1.0;
  if (rz_new <= d.scal[SC_TOL2] * d.scal[SC_RZ0]) d.scal[SC_DONE] = 1.0;
}

// ---- dense reduced system for small static-only graphs (NS1) ----
// S (n x n, n = 6C, row-major, lower triangle used) = Hpp + lambda I (+ se3-se3 off-diagonal blocks) - sum_j (1 / s_j) H_pl,j H_pl,j^T, rhs = bp - sum_j H_pl,j b_l,j / s_j
// with H_pl for an EdgeSE3PointXYZ = om J_c^T J_p, J_c = [-I | 2 [Zc]x], J_p = R_c^T (edge_se3_pointxyz.cpp:99-140).
__global__ void __launch_bounds__(128) k_dense_init(BaDev d, double lambda, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double* S = d.Sdense; double* rhs = S + (size_t)n * n;
  if (i < n * n) {
    const int r = i / n, c = i % n, vr = r / 6, vc = c / 6;
    double v = 0.0;
    if (vr == vc) { v = d.Hpp[36 * (size_t)vr + 6 * (r % 6) + (c % 6)]; if (r == c) v += lambda; }
    S[i] = v;
  }
  if (i < n) rhs[i] = d.bp[i];
  if (i == 0) rhs[n] = 0.0;                       // status word: != 0 after the factorisation means "not positive definite"
}
__global__ void __launch_bounds__(64) k_dense_se3_edges(BaDev d, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.Ese || d.se_j[e] < 0) return;
  const int i = d.se_i[e], j = d.se_j[e];
  const double* H = d.se_Hoff + 36 * (size_t)e;    // J_i^T W J_j: rows i, columns j
  double* S = d.Sdense;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) {
      if (i > j) atomicAdd(S + (size_t)(6 * i + r) * n + 6 * j + c, H[6 * r + c]);
      else atomicAdd(S + (size_t)(6 * j + c) * n + 6 * i + r, H[6 * r + c]);
    }
}
__global__ void __launch_bounds__(128) k_dense_schur(BaDev d, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d.P) return;
  const int eb = d.lm_obs_begin[k], ee = d.lm_obs_begin[k + 1];
  if (ee <= eb) return;
  const double is = 1.0 / d.pt_s[k];
  const double p[3] = {d.pt[3 * (size_t)k], d.pt[3 * (size_t)k + 1], d.pt[3 * (size_t)k + 2]};
  const double bls[3] = {d.bl[3 * (size_t)k] * is, d.bl[3 * (size_t)k + 1] * is, d.bl[3 * (size_t)k + 2] * is};
  double* S = d.Sdense; double* rhs = S + (size_t)n * n;
  // M_c = om J_c^T R_c^T (6 x 3): rows 0-2 = -om R^T, rows 3-5 = om (2 [Zc]x)^T R^T
  auto make_M = [&](int e, double* M) -> int {
    const int c = d.lm_cam[e];
    const double* T = d.se3 + 12 * (size_t)c;
    const double w[3] = {p[0] - T[9], p[1] - T[10], p[2] - T[11]};
    double Zc[3]; rot_t_apply(T, w, Zc);
    const double om = d.lm_omega[e];
    // 2 [Zc]x = [[0, -2z, 2y], [2z, 0, -2x], [-2y, 2x, 0]];  J_c^T rows 3-5 = (2 [Zc]x)^T
    const double A[9] = {0, 2 * Zc[2], -2 * Zc[1], -2 * Zc[2], 0, 2 * Zc[0], 2 * Zc[1], -2 * Zc[0], 0};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        M[3 * r + q] = -om * T[3 * q + r];                                                  // -om R^T
        M[9 + 3 * r + q] = om * (A[3 * r] * T[3 * q] + A[3 * r + 1] * T[3 * q + 1] + A[3 * r + 2] * T[3 * q + 2]);   // om A R^T
      }
    return c;
  };
  for (int e1 = eb; e1 < ee; ++e1) {
    double M1[18]; const int c1 = make_M(e1, M1);
#pragma unroll
    for (int r = 0; r < 6; ++r) atomicAdd(rhs + 6 * c1 + r, -(M1[3 * r] * bls[0] + M1[3 * r + 1] * bls[1] + M1[3 * r + 2] * bls[2]));
    for (int e2 = eb; e2 < ee; ++e2) {
      double M2[18]; const int c2 = make_M(e2, M2);
      if (c2 > c1) continue;                                                                  // lower triangle (blocks with c1 >= c2)
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          atomicAdd(S + (size_t)(6 * c1 + r) * n + 6 * c2 + q, -is * (M1[3 * r] * M2[3 * q] + M1[3 * r + 1] * M2[3 * q + 1] + M1[3 * r + 2] * M2[3 * q + 2]));
        }
    }
  }
}
// S += the static Schur term from the band (see k_band_form / k_band_mul): block (row vertex b = a + k, column vertex a) = Y_b Kc X_a with
//   X_a = [[-R, -2 [t]x R], [0, R]]  (local increment -> the world-frame vector vw of body_vertex_transform),
//   Kc  = [[M0 I, 2 [M1]x], [2 [M1]x, 4 (M2 - tr(M2) I)]]  (the band product, sign folded in),
//   Y_b = [[R^T, 0], [-2 R^T [t]x, R^T]]  (k_tile_finalize_schur2: torque moved to the vertex origin, rotated into the vertex frame).
// One thread per (a, k); every block is written by exactly one thread (no atomics -- the thread-per-landmark k_dense_schur issued
// ~600 fp64 atomics per landmark onto the 20 x 20 blocks: 0.4 ms per trial on the 20-camera window, mostly contention).
__global__ void __launch_bounds__(128) k_dense_from_band(BaDev d, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = d.band_W;
  if (idx >= d.band_n * W) return;
  const int a = idx / W, k = idx - a * W, b = a + k;
  if (b >= d.band_n) return;
  const double* m = d.band + (size_t)idx * 10;
  if (m[0] == 0.0) return;
  const int va = d.band_v0 + a, vb = d.band_v0 + b;
  const double* Ta = d.se3 + 12 * (size_t)va; const double* Tb = d.se3 + 12 * (size_t)vb;
  auto skew = [](const double* t, double* K) { K[0] = 0; K[1] = -t[2]; K[2] = t[1]; K[3] = t[2]; K[4] = 0; K[5] = -t[0]; K[6] = -t[1]; K[7] = t[0]; K[8] = 0; };
  double X[36], Kc[36], Z[36];
  {   // X_a
    double tx[9]; skew(Ta + 9, tx);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        const double txr = tx[3 * r] * Ta[c] + tx[3 * r + 1] * Ta[3 + c] + tx[3 * r + 2] * Ta[6 + c];     // ([t]x R)[r][c]
        X[6 * r + c] = -Ta[3 * r + c]; X[6 * r + 3 + c] = -2.0 * txr;
        X[6 * (r + 3) + c] = 0.0; X[6 * (r + 3) + 3 + c] = Ta[3 * r + c];
      }
  }
  {   // Kc
    const double m1[3] = {m[1], m[2], m[3]};
    double m1x[9]; skew(m1, m1x);
    const double M2[9] = {m[4], m[5], m[6], m[5], m[7], m[8], m[6], m[8], m[9]};
    const double tr = m[4] + m[7] + m[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        Kc[6 * r + c] = (r == c) ? m[0] : 0.0; Kc[6 * r + 3 + c] = 2.0 * m1x[3 * r + c];
        Kc[6 * (r + 3) + c] = 2.0 * m1x[3 * r + c]; Kc[6 * (r + 3) + 3 + c] = 4.0 * (M2[3 * r + c] - (r == c ? tr : 0.0));
      }
  }
  mat6_mul(Kc, X, Z);
  double Y[36];
  {   // Y_b
    double tx[9]; skew(Tb + 9, tx);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        const double rtx = Tb[r] * tx[c] + Tb[3 + r] * tx[3 + c] + Tb[6 + r] * tx[6 + c];                   // (R^T [t]x)[r][c]
        Y[6 * r + c] = Tb[3 * c + r]; Y[6 * r + 3 + c] = 0.0;
        Y[6 * (r + 3) + c] = -2.0 * rtx; Y[6 * (r + 3) + 3 + c] = Tb[3 * c + r];
      }
  }
  double B[36];
  mat6_mul(Y, Z, B);
  double* S = d.Sdense;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) S[(size_t)(6 * vb + r) * n + 6 * va + c] += B[6 * r + c];
}
// One CTA: blocked right-looking Cholesky of the lower triangle of S in shared memory (8-column panels; the trailing update
// A[i][j] -= L[i][k] L[j][k]^T over 8x8 tiles is two mma.sync.m8n8k4.f64 per tile), then the two triangular solves.  n <= DENSE_MAX.
constexpr int DENSE_MAX = 168;
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) k_dense_chol(BaDev d, int n) {
  extern __shared__ double sA[];                    // npad x ld, row-major; ld = npad + 1 (bank spread)
  const int npad = (n + 7) & ~7, ld = npad + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double* S = d.Sdense; double* rhs = S + (size_t)n * n;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int i = tid; i < npad * npad; i += blockDim.x) {
    const int r = i / npad, c = i % npad;
    double v = 0.0;
    if (r < n && c < n) v = (c <= r) ? S[(size_t)r * n + c] : S[(size_t)c * n + r];   // mirror the lower triangle (block (c1 == c2) parts were written in full)
    else if (r == c) v = 1.0;                                                        // padding: identity
    sA[r * ld + c] = v;
  }
  __syncthreads();
  const int nb = npad / 8;
  for (int kb = 0; kb < nb; ++kb) {
    const int k0 = 8 * kb;
    if (warp == 0) {                                // 8 x 8 diagonal block, unblocked (lanes = rows)
      for (int j = 0; j < 8; ++j) {
        double djj = sA[(k0 + j) * ld + k0 + j];
        if (!(djj > 0.0)) { if (lane == 0) bad = 1; djj = 1.0; }
        const double l = sqrt(djj);
        __syncwarp();
        if (lane == j) sA[(k0 + j) * ld + k0 + j] = l;
        if (lane > j && lane < 8) sA[(k0 + lane) * ld + k0 + j] /= l;
        __syncwarp();
        if (lane > j && lane < 8) {
          const double lij = sA[(k0 + lane) * ld + k0 + j];
          for (int c = j + 1; c <= lane; ++c) sA[(k0 + lane) * ld + k0 + c] -= lij * sA[(k0 + c) * ld + k0 + j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // panel: rows below the diagonal block solve x L_kk^T = a (one thread per row)
    for (int r = k0 + 8 + tid; r < npad; r += blockDim.x) {
      double x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double s = sA[r * ld + k0 + j];
#pragma unroll
        for (int c = 0; c < j; ++c) s -= x[c] * sA[(k0 + j) * ld + k0 + c];
        x[j] = s / sA[(k0 + j) * ld + k0 + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sA[r * ld + k0 + j] = x[j];
    }
    __syncthreads();
    // trailing update on the fp64 tensor cores: tile (ib, jb), ib >= jb > kb:  A_ij -= L_ik L_jk^T
    const int nt = nb - kb - 1, ntiles = nt * (nt + 1) / 2;
    for (int t = warp; t < ntiles; t += nw) {
      int a = 0, rem = t;
      while (rem >= nt - a) { rem -= nt - a; ++a; }
      const int jb = kb + 1 + a, ib = jb + rem;
      const int g = lane >> 2, q = lane & 3;        // fragment coordinates: A row g / col q (+4), B row q (+4) / col g, C row g / cols 2q, 2q + 1
      double c0 = 0.0, c1 = 0.0;
      dmma_m8n8k4(c0, c1, sA[(8 * ib + g) * ld + k0 + q], sA[(8 * jb + g) * ld + k0 + q]);
      dmma_m8n8k4(c0, c1, sA[(8 * ib + g) * ld + k0 + 4 + q], sA[(8 * jb + g) * ld + k0 + 4 + q]);
      sA[(8 * ib + g) * ld + 8 * jb + 2 * q] -= c0;
      sA[(8 * ib + g) * ld + 8 * jb + 2 * q + 1] -= c1;
    }
    __syncthreads();
  }
  // L y = rhs, L^T x = y (warp 0; n is small)
  __shared__ double y[DENSE_MAX];
  for (int i = tid; i < npad; i += blockDim.x) y[i] = i < n ? rhs[i] : 0.0;
  __syncthreads();
  if (warp == 0) {
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int c = lane; c < i; c += 32) s += sA[i * ld + c] * y[c];
      s = warp_sum(s);
      if (lane == 0) y[i] = (y[i] - s) / sA[i * ld + i];
      __syncwarp();
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = 0.0;
      for (int c = i + 1 + lane; c < n; c += 32) s += sA[c * ld + i] * y[c];
      s = warp_sum(s);
      if (lane == 0) y[i] = (y[i] - s) / sA[i * ld + i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) d.xp[i] = y[i];
  if (tid == 0) rhs[n] = bad ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(128) k_apply_update(BaDev d, double lambda, int reortho) {
  __shared__ double red[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (i < d.C) s = body_update_se3(d, i, lambda, reortho != 0);
  else if (i < d.C + d.P) s = body_update_pt(d, i - d.C, lambda);
  s = block_sum(s, red);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(d.scal + SC_SCALE, s);
}

}  // namespace vdo
#include "ba_tile_kernels.cuh"
namespace vdo {

// ---------------------------------------------------------------------------------------------------------------
// NCCL is resolved at run time (dlopen) so that the library links without it and picks up the copy torch already loaded
struct NcclApi {
  void* h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce;
  }
};
static NcclApi g_nccl;

struct CudaBackend : BaBackend {
  int dev = 0;
  ncclComm_t comm = nullptr;
  void allreduce_sum(double* b, size_t n) override {
    if (world <= 1 || n == 0) return;
    ncclResult_t r = g_nccl.AllReduce(b, b, n, ncclDouble, ncclSum, comm, st);
    if (r != ncclSuccess) std::fprintf(stderr, "[vdo_b200] ncclAllReduce failed: %s\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    ++n_coll;
  }
  void allreduce_max(double* b, size_t n) override {
    if (world <= 1 || n == 0) return;
    ncclResult_t r = g_nccl.AllReduce(b, b, n, ncclDouble, ncclMax, comm, st);
    if (r != ncclSuccess) std::fprintf(stderr, "[vdo_b200] ncclAllReduce failed: %s\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    ++n_coll;
  }
  int n_coll = 0;
  // ---- peer-memory exchange of the sharded PCG iteration: one buffer per rank (flags + 2 x world slots of 6C doubles), mapped into every
  //      other rank through CUDA IPC (handles all-gathered over NCCL once per buffer size) ----
  struct Xchg {
    char* local = nullptr; size_t bytes = 0, off_z = 0, off_prz = 0; int C = 0, n_paths = 0; bool ok = false, tried = false;
    std::vector<char*> peer;                 // peer[r]: this process' mapping of rank r's buffer (peer[rank] == local)
    double** d_slots = nullptr; unsigned long long** d_flags = nullptr; unsigned long long* d_epoch = nullptr;
  } xg;
  static constexpr size_t XG_HDR = 4096;     // flags (world x u64) + epoch live in the first page of the buffer
  void xchg_release() {
    drop_graphs();                           // captured PCG graphs carry the buffer's addresses
    for (size_t r = 0; r < xg.peer.size(); ++r) if ((int)r != rank && xg.peer[r]) cudaIpcCloseMemHandle(xg.peer[r]);
    xg.peer.clear();
    if (xg.local) cudaFree(xg.local);
    if (xg.d_slots) cudaFree(xg.d_slots);
    if (xg.d_flags) cudaFree(xg.d_flags);
    xg = Xchg();
  }
  // collective: every rank calls it with the same C.  Returns false (and the caller falls back to NCCL all-reduces) when peer mapping is unavailable.
  bool xchg_setup(int C, int n_paths) {
    if (xg.ok && xg.C >= C && xg.n_paths >= n_paths) return true;
    if (xg.tried && !xg.ok) return false;
    if (std::getenv("VDO_NO_PEER_EXCHANGE")) { xg.tried = true; return false; }
    CK(cudaStreamSynchronize(st));
    xchg_release();
    xg.tried = true; xg.C = C; xg.n_paths = n_paths;
    // header | 2 x world slots of 6C | z (6C) | part_rz (n_paths x PCR_CL)
    xg.off_z = 2 * (size_t)world * 6 * (size_t)C; xg.off_prz = xg.off_z + 6 * (size_t)C;
    xg.bytes = XG_HDR + sizeof(double) * (xg.off_prz + (size_t)n_paths * PCR_CL + 2);
    bool good = g_nccl.AllGather != nullptr;
    if (good && cudaMalloc(&xg.local, xg.bytes) != cudaSuccess) { cudaGetLastError(); xg.local = nullptr; good = false; }
    if (good) CK(cudaMemsetAsync(xg.local, 0, xg.bytes, st));
    cudaIpcMemHandle_t mine; std::memset(&mine, 0, sizeof mine);
    if (good && cudaIpcGetMemHandle(&mine, xg.local) != cudaSuccess) { cudaGetLastError(); good = false; }
    // all-gather {ok flag, handle} -- also the agreement on whether every rank got this far
    struct Rec { int ok; int pad; cudaIpcMemHandle_t h; };
    Rec rec; rec.ok = good ? 1 : 0; rec.pad = 0; rec.h = mine;
    char* d_all = nullptr; CK(cudaMalloc(&d_all, sizeof(Rec) * (size_t)world));
    CK(cudaMemcpyAsync(d_all + sizeof(Rec) * (size_t)rank, &rec, sizeof(Rec), cudaMemcpyHostToDevice, st));
    if (g_nccl.AllGather) {
      ncclResult_t r = g_nccl.AllGather(d_all + sizeof(Rec) * (size_t)rank, d_all, sizeof(Rec), ncclChar, comm, st);
      if (r != ncclSuccess) good = false;
    }
    std::vector<Rec> all(world);
    CK(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * (size_t)world, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_all);
    for (int r = 0; r < world; ++r) if (!all[r].ok) good = false;
    if (good) {
      xg.peer.assign(world, nullptr);
      for (int r = 0; r < world && good; ++r) {
        if (r == rank) { xg.peer[r] = xg.local; continue; }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; }
        xg.peer[r] = (char*)ptr;
      }
    }
    // agree on the outcome (a rank that failed to map a peer must take every rank down to the NCCL path)
    double flag = good ? 0.0 : 1.0, *d_flag = nullptr;
    CK(cudaMalloc(&d_flag, sizeof(double)));
    CK(cudaMemcpyAsync(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
    allreduce_sum(d_flag, 1);
    CK(cudaMemcpyAsync(&flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_flag);
    if (flag != 0.0) { xchg_release(); xg.tried = true; std::fprintf(stderr, "[vdo_b200] rank %d: peer-memory exchange unavailable, using NCCL all-reduces per PCG iteration\n", rank); return false; }
    std::vector<double*> hs(world); std::vector<unsigned long long*> hf(world);
    for (int r = 0; r < world; ++r) { hf[r] = (unsigned long long*)xg.peer[r]; hs[r] = (double*)(xg.peer[r] + XG_HDR); }
    CK(cudaMalloc(&xg.d_slots, sizeof(double*) * (size_t)world)); CK(cudaMalloc(&xg.d_flags, sizeof(unsigned long long*) * (size_t)world));
    CK(cudaMemcpyAsync(xg.d_slots, hs.data(), sizeof(double*) * (size_t)world, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(xg.d_flags, hf.data(), sizeof(unsigned long long*) * (size_t)world, cudaMemcpyHostToDevice, st));
    xg.d_epoch = (unsigned long long*)(xg.local + 8 * 256);       // behind the flags (world <= 256)
    CK(cudaStreamSynchronize(st));
    xg.ok = true;
    return true;
  }
  cudaStream_t st = nullptr, st2 = nullptr;   // st2: second branch inside the captured PCG graph
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int n_launch = 0;
  cudaEvent_t ev0[4], ev1[4];
  ~CudaBackend() override {
    for (int i = 0; i < 4; ++i) { cudaEventDestroy(ev0[i]); cudaEventDestroy(ev1[i]); }
    for (auto& kv : pool) cudaFree(kv.second);
    arena.destroy();
    xchg_release();
    if (comm) g_nccl.CommDestroy(comm);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (st2) cudaStreamDestroy(st2);
    if (st) cudaStreamDestroy(st);
  }
  // pinned staging arena (graph ingestion builds its upload streams in it) and a caching device allocator: graphs of the
  // same shape are created again and again by the callers (one per window / per solve), so freed blocks are kept and
  // handed back by exact size instead of going through cudaFree / cudaMalloc (both synchronise the device)
  struct PinnedArena : HostArena {
    char* raw_alloc(size_t b) override { void* p = nullptr; if (cudaHostAlloc(&p, b, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = std::malloc(b); pageable.insert(p); } return (char*)p; }
    void raw_free(char* p) override { if (pageable.count(p)) { pageable.erase(p); std::free(p); } else cudaFreeHost(p); }
    std::set<void*> pageable;
  } arena;
  HostArena& staging() override { return arena; }
  std::multimap<size_t, void*> pool;      // free device blocks by size
  std::map<void*, size_t> live;           // size of every block handed out
  size_t pool_bytes = 0;
  static constexpr size_t POOL_MAX = (size_t)24 << 30;
  // size classes: 1/16 steps of the enclosing power of two (<= 12.5 % slack), powers of two up to 4 KB -- callers such as the per-window
  // optimiser build graphs whose array sizes differ by a few elements from run to run; with exact sizes every buffer missed the cache
  // (~100 cudaMalloc per graph, occasional 30-80 ms stalls when the driver grew its heap)
  static size_t size_class(size_t b) {
    size_t p2 = 256;
    while (p2 < b) p2 <<= 1;
    if (p2 <= 4096) return p2;
    const size_t step = p2 >> 4;
    return (b + step - 1) / step * step;
  }
  void* alloc(size_t b) override {
    b = b ? b : 8;
    const size_t cls = size_class(b);
    void* p = nullptr;
    auto it = pool.find(cls);
    if (it != pool.end()) { p = it->second; pool_bytes -= cls; pool.erase(it); }
    else CK(cudaMalloc(&p, cls));
    live[p] = cls;
    CK(cudaMemsetAsync(p, 0, b, st));
    return p;
  }
  void free_(void* p) override {
    auto it = live.find(p);
    if (it == live.end()) { cudaFree(p); return; }
    const size_t b = it->second;
    live.erase(it);
    if (pool_bytes + b <= POOL_MAX) { pool.emplace(b, p); pool_bytes += b; }   // all users are ordered on st: no synchronisation needed
    else cudaFree(p);
  }
  void h2d_async(void* d, const void* s, size_t b) override { CK(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st)); }
  void h2d(void* d, const void* s, size_t b) override { CK(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); }
  void d2h(void* d, const void* s, size_t b) override { CK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
  void d2d(void* d, const void* s, size_t b) override { CK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, st)); }
  void zero(void* d, size_t b) override { CK(cudaMemsetAsync(d, 0, b, st)); }
  void sync() override { CK(cudaStreamSynchronize(st)); }
  int launches() const override { return n_launch; }
  void* stream() const override { return (void*)st; }
  void timer_start(int s) override { CK(cudaEventRecord(ev0[s], st)); }
  float timer_stop_ms(int s) override { float ms = 0; CK(cudaEventRecord(ev1[s], st)); CK(cudaEventSynchronize(ev1[s])); CK(cudaEventElapsedTime(&ms, ev0[s], ev1[s])); return ms; }

  static int nblk(int n, int b) { return n > 0 ? (n + b - 1) / b : 0; }
#define LAUNCH(kern, grid, block, ...)                         \
  do {                                                         \
    if ((grid) > 0) { kern<<<(grid), (block), 0, st>>>(__VA_ARGS__); ++n_launch; } \
  } while (0)

  template <typename K> void launch_tiles(K kern, size_t smem, const BaDev& d, int tile0, int n) {
    if (n > 0) { kern<<<n, VDO_TILE_L, smem, st>>>(d, tile0); ++n_launch; }
  }
  void tile_lin(BaDev& d, bool write, int part) {   // part: 0 static tiles, 1 chain tiles, -1 both
    const int ns = d.n_tiles_stat, nc = d.n_tiles - d.n_tiles_stat;
    if (part != 1) { if (write) launch_tiles(k_tile_lin<false, true>, SMEM_LIN_ST, d, 0, ns); else launch_tiles(k_tile_lin<false, false>, SMEM_LIN_ST, d, 0, ns); }
    if (part != 0) { if (write) launch_tiles(k_tile_lin<true, true>, SMEM_LIN_CH, d, ns, nc); else launch_tiles(k_tile_lin<true, false>, SMEM_LIN_CH, d, ns, nc); }
  }
  // modes 0 / 1 (rhs, S*p): k_tile_schur2 (one thread per (run, component) on the vertex side); mode 2 (back-substitution): k_tile_schur
  void tile_schur(BaDev& d, int mode, int part, cudaStream_t chain_stream) {
    const int ns = d.n_tiles_stat, nc = d.n_tiles - d.n_tiles_stat;
    const size_t bs = SMEM_SCH_ST, bc = SMEM_SCH_CH;
    if (part != 1 && ns > 0 && mode == 1 && d.band) {
      k_band_mul<<<nblk(d.band_n, 8), 256, 0, st>>>(d); ++n_launch;
    } else if (part != 1 && ns > 0) {
      const size_t sm2 = smem_sch2(false, d.capE_st, d.capV_st, 1);
      if (mode == 0) k_tile_schur2<false, 0><<<ns, VDO_TILE_L, sm2, st>>>(d, 0, d.capE_st, d.capV_st, 1);
      else if (mode == 1) k_tile_schur2<false, 1><<<ns, VDO_TILE_L, sm2, st>>>(d, 0, d.capE_st, d.capV_st, 1);
      else k_tile_schur<false, 2><<<ns, VDO_TILE_L, bs, st>>>(d, 0);
      ++n_launch;
    }
    if (part != 0 && nc > 0) {
      const size_t sm2 = smem_sch2(true, d.capE_ch, d.capV_ch, d.capH_ch);
      if (mode == 0) k_tile_schur2<true, 0><<<nc, VDO_TILE_L, sm2, chain_stream>>>(d, ns, d.capE_ch, d.capV_ch, d.capH_ch);
      else if (mode == 1) k_tile_schur2<true, 1><<<nc, VDO_TILE_L, sm2, chain_stream>>>(d, ns, d.capE_ch, d.capV_ch, d.capH_ch);
      else k_tile_schur<true, 2><<<nc, VDO_TILE_L, bc, chain_stream>>>(d, ns);
      ++n_launch;
    }
  }
  void lin_tracklets(BaDev& d, bool write) override {
    if (d.tiled) { tile_lin(d, write, -1); return; }
    if (write) { LAUNCH(k_lin_static<true>, nblk(d.Tstat, 256), 256, d); LAUNCH(k_lin_tracklets<true>, nblk(d.T - d.Tstat, 128), 128, d); }
    else { LAUNCH(k_lin_static<false>, nblk(d.Tstat, 256), 256, d); LAUNCH(k_lin_tracklets<false>, nblk(d.T - d.Tstat, 128), 128, d); }
  }
  // tiled layout: the tile kernels of lin_tracklets(write) have already formed the vertex-side sums in the world frame;
  // lin_vertex_obs turns them into H_pp / b_p, lin_vertex_ter has nothing left to do (same split for precond_* and schur_vertex_*)
  void lin_vertex_obs(BaDev& d) override {
    if (d.tiled) { LAUNCH(k_tile_finalize_lin, nblk(d.C, 128), 128, d); return; }
    auto k = k_vertex_sym<0, true>; LAUNCH(k, d.n_obs_chunks, 128, d);
  }
  void lin_vertex_ter(BaDev& d) override { if (d.tiled) return; auto k = k_vertex_sym<0, false>; LAUNCH(k, d.n_ter_chunks, 128, d); }
  void lin_se3_edges(BaDev& d, bool write) override {
    if (write) LAUNCH(k_lin_se3_edges<true>, nblk(d.Ese, 64), 64, d); else LAUNCH(k_lin_se3_edges<false>, nblk(d.Ese, 64), 64, d);
  }
  void max_diagonal(BaDev& d) override {
    zero(d.scal + SC_MAXDIAG, sizeof(double));
    int n = d.C * 6 + d.P;
    LAUNCH(k_max_diagonal, min(nblk(n, 256), 148 * 8), 256, d);
  }
  int band_max_width() const override { return 32; }
  void band_form(BaDev& d) override {
    if (!d.band || d.n_tiles_stat <= 0) return;
    zero(d.band, sizeof(double) * 10 * (size_t)d.band_n * d.band_W);
    const int per = max(1, (d.n_tiles_stat + 148 * 3 - 1) / (148 * 3));          // 3 CTAs per SM, each a run of consecutive tiles
    k_band_form<<<nblk(d.n_tiles_stat, per), VDO_TILE_L, smem_band(d.capE_st), st>>>(d, per, d.capE_st); ++n_launch;
  }
  void factor_landmarks(BaDev& d, double lambda) override { LAUNCH(k_factor_landmarks, nblk(d.T, 128), 128, d, lambda); }
  void precond_begin(BaDev& d, double lambda) override { LAUNCH(k_precond_begin, nblk(d.C * 36, 128), 128, d, lambda); }
  void precond_vertex_obs(BaDev& d) override {
    if (d.tiled) {
      launch_tiles(k_tile_precond<false>, SMEM_PRE_ST, d, 0, d.n_tiles_stat);
      launch_tiles(k_tile_precond<true>, SMEM_PRE_CH, d, d.n_tiles_stat, d.n_tiles - d.n_tiles_stat);
      LAUNCH(k_tile_finalize_precond, nblk(d.C, 128), 128, d);
      return;
    }
    auto k = k_vertex_sym<1, true>; LAUNCH(k, d.n_obs_chunks, 128, d);
  }
  void precond_vertex_ter(BaDev& d) override { if (d.tiled) return; auto k = k_vertex_sym<1, false>; LAUNCH(k, d.n_ter_chunks, 128, d); }
  // long paths (clusters of PCR_CL CTAs) first in own_paths, then the short ones (one CTA each)
  void precond_factor(BaDev& d, double lambda) override {
    LAUNCH(k_pcr_factor<PCR_CL>, d.n_own_long * PCR_CL, 256, d, lambda, 0);
    LAUNCH(k_pcr_factor<1>, d.n_own_paths - d.n_own_long, 256, d, lambda, d.n_own_long);
  }
  template <bool FUSED> void launch_step_a(BaDev& d, const double* p) {
    const unsigned int total = (unsigned int)(d.n_own_long * PCR_CL + (d.n_own_paths - d.n_own_long));
    LAUNCH((k_pcg_step_a<FUSED, PCR_CL>), d.n_own_long * PCR_CL, 256, d, p, 0, total);
    LAUNCH((k_pcg_step_a<FUSED, 1>), d.n_own_paths - d.n_own_long, 256, d, p, d.n_own_long, total);
  }
  void schur_landmarks(BaDev& d, int mode, const double* v) override {
    if (d.tiled) { tile_schur(d, mode, -1, st); return; }
    const int g = nblk((d.T - d.Tstat) * 8, 128), gs = nblk(d.Tstat, 256);
    if (mode == 0) { LAUNCH(k_schur_static<0>, gs, 256, d, d.zl); LAUNCH(k_schur_chains8<0>, g, 128, d, v, d.zl); }
    else if (mode == 1) { LAUNCH(k_schur_static<1>, gs, 256, d, d.zl); LAUNCH(k_schur_chains8<1>, g, 128, d, v, d.zl); }
    else { LAUNCH(k_schur_static<2>, gs, 256, d, d.xl); LAUNCH(k_schur_chains8<2>, g, 128, d, v, d.xl); }
  }
  void schur_landmarks_part(BaDev& d, int mode, const double* v, int part) override {
    if (d.tiled) { tile_schur(d, 1, part, st); return; }
    const int g = nblk((d.T - d.Tstat) * 8, 128), gs = nblk(d.Tstat, 256);
    if (part == 0) LAUNCH(k_schur_static<1>, gs, 256, d, d.zl); else LAUNCH(k_schur_chains8<1>, g, 128, d, v, d.zl);
    (void)mode;
  }
  void lin_tracklets_part(BaDev& d, bool write, int part) override {
    if (d.tiled) { tile_lin(d, true, part); return; }
    (void)write;
    if (part == 0) LAUNCH(k_lin_static<true>, nblk(d.Tstat, 256), 256, d); else LAUNCH(k_lin_tracklets<true>, nblk(d.T - d.Tstat, 128), 128, d);
  }
  void schur_vertex_obs(BaDev& d, double sign, double* out) override {
    if (d.tiled) { LAUNCH(k_tile_finalize_schur2, nblk(d.C, 128), 128, d, sign, out, out == d.Ap ? 1 : 0, (const double*)nullptr); return; }
    LAUNCH(k_schur_vertex<true>, d.n_obs_chunks, 128, d, sign, out, out == d.Ap ? 1 : 0);
  }
  void schur_vertex_ter(BaDev& d, double sign, double* out) override {
    if (d.tiled) return;
    LAUNCH(k_schur_vertex<false>, d.n_ter_chunks, 128, d, sign, out, out == d.Ap ? 1 : 0);
  }
  void set_scalars(BaDev& d, double lambda, double tol2) {
    if (lambda != cur_lambda || tol2 != cur_tol2 || d.scal != cur_scal) { LAUNCH(k_set_scalars, 1, 1, d, lambda, tol2); cur_lambda = lambda; cur_tol2 = tol2; cur_scal = d.scal; }
  }
  double cur_lambda = -1, cur_tol2 = -1; double* cur_scal = nullptr;
  void vertex_transform(BaDev& d, const double* v) override { LAUNCH(k_vertex_transform, nblk(d.C, 128), 128, d, v); }
  void hpp_mul(BaDev& d, double lambda, const double* x, double* out) override { set_scalars(d, lambda, cur_tol2 < 0 ? 0.0 : cur_tol2); LAUNCH(k_hpp_mul, nblk(d.C, 128), 128, d, x, out); }
  void pcg_init(BaDev& d) override {
    zero(d.scal + SC_PAP, 6 * sizeof(double));   // PAP, RZ, RZ_NEW, RZ0, DONE, ITERS
    if (d.xg_paths) zero(d.xp, 48 * (size_t)d.C);        // path-sharded: a rank touches x on its own paths only; the rest must read 0 in the final sum
    {
      const unsigned int total = (unsigned int)(d.n_own_long * PCR_CL + (d.n_own_paths - d.n_own_long));
      LAUNCH(k_pcg_init<PCR_CL>, d.n_own_long * PCR_CL, 256, d, 0, total);
      LAUNCH(k_pcg_init<1>, d.n_own_paths - d.n_own_long, 256, d, d.n_own_long, total);
    }
    LAUNCH(k_pcg_init_fin, 1, 256, d);
  }
  void pcg_dot_pAp(BaDev& d) override { LAUNCH(k_pcg_dot, d.n_part_pap, 256, d); }   // one CTA per slot of part_pap
  void pcg_step(BaDev& d, double tol2) override {
    set_scalars(d, cur_lambda, tol2);
    launch_step_a<false>(d, (const double*)d.p);
    LAUNCH(k_pcg_step_b, min(nblk(d.C * 6, 256), 148), 256, d);
    LAUNCH(k_pcg_scalars, 1, 256, d);
  }
  // n PCG iterations as ONE CUDA-graph launch (captured once per factor graph and batch size; lambda / tolerance travel
  // through device scalars so the captured kernel arguments never change)
  std::map<std::pair<const void*, int>, cudaGraphExec_t> graphs;
  void pcg_iterate(BaDev& d, double lambda, double tol2, int n) override {
    // sharded: the same captured graph with the all-reduce of S*p done by two kernels over peer memory (k_xchg_scatter / k_xchg_reduce);
    // without peer mapping (or with the chunked layout) plain launches + one NCCL all-reduce per iteration
    bool peer = false;
    if (world > 1) {
      peer = d.tiled && d.xg_paths && xg.ok;            // decided (collectively) at finalize: shard_paths
      if (!peer) { BaBackend::pcg_iterate(d, lambda, tol2, n); return; }
    }
    set_scalars(d, lambda, tol2);
    auto key = std::make_pair((const void*)d.scal, n);
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
      const int before = n_launch;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int gch = nblk((d.T - d.Tstat) * 8, 128), gst = nblk(d.Tstat, 256);
      for (int b = 0; b < n; ++b) {
        if (d.tiled) {
          // fused iteration (n is even: the search direction ends in d.p again)
          const double* p_in = (b & 1) ? d.p2 : d.p; double* p_out = (b & 1) ? d.p : d.p2;
          LAUNCH(k_pcg_p_hpp, nblk(d.C, 128), 128, d, p_in, p_out, d.Ap);
          // fork: static tiles on st, chain tiles on st2 (independent landmark sets); both scatter into acc6 with atomics
          CK(cudaEventRecord(ev_fork, st)); CK(cudaStreamWaitEvent(st2, ev_fork, 0));
          tile_schur(d, 1, -1, st2);
          CK(cudaEventRecord(ev_join, st2)); CK(cudaStreamWaitEvent(st, ev_join, 0));
          if (peer) {
            LAUNCH(k_xchg_scatter, nblk(d.C, 128), 128, d, -1.0, (const double*)d.Ap);      // partial S*p into slot[rank] of every rank
            LAUNCH(k_xchg_reduce, nblk(d.C, 128), 128, d, d.Ap, (const double*)p_out);      // sum of the slots in rank order, partials of p.Ap
          } else
            LAUNCH(k_tile_finalize_schur2, nblk(d.C, 128), 128, d, -1.0, d.Ap, 1, p_out);  // Ap -= B^T sums, and the partials of p.Ap
          launch_step_a<true>(d, (const double*)p_out);
          if (d.xg_paths) LAUNCH(k_pcg_scalars_x, 1, 256, d);
          continue;
        }
        LAUNCH(k_hpp_mul, nblk(d.C, 128), 128, d, (const double*)d.p, d.Ap);
        // fork: static landmarks on st, chains on st2 (independent landmark sets; the chain kernel is latency-bound)
        CK(cudaEventRecord(ev_fork, st)); CK(cudaStreamWaitEvent(st2, ev_fork, 0));
        LAUNCH(k_schur_static<1>, gst, 256, d, d.zl);
        if (gch > 0) { k_schur_chains8<1><<<gch, 128, 0, st2>>>(d, (const double*)d.p, d.zl); ++n_launch; }
        CK(cudaEventRecord(ev_join, st2)); CK(cudaStreamWaitEvent(st, ev_join, 0));
        // fork: the two vertex-major passes add into Ap with atomics and are independent of each other
        CK(cudaEventRecord(ev_fork, st)); CK(cudaStreamWaitEvent(st2, ev_fork, 0));
        schur_vertex_obs(d, -1.0, d.Ap);
        if (d.n_ter_chunks > 0) { k_schur_vertex<false><<<d.n_ter_chunks, 128, 0, st2>>>(d, -1.0, d.Ap, 1); ++n_launch; }
        CK(cudaEventRecord(ev_join, st2)); CK(cudaStreamWaitEvent(st, ev_join, 0));
        LAUNCH(k_pcg_dot, d.n_part_pap, 256, d);
        launch_step_a<false>(d, (const double*)d.p);
        LAUNCH(k_pcg_step_b, min(nblk(d.C * 6, 256), 148), 256, d);
        LAUNCH(k_pcg_scalars, 1, 256, d);
      }
      CK(cudaStreamEndCapture(st, &g));
      CK(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      per_batch[key] = n_launch - before;
      n_launch = before;
      it = graphs.emplace(key, ge).first;
    }
    CK(cudaGraphLaunch(it->second, st));
    n_launch += per_batch[key];
  }
  std::map<std::pair<const void*, int>, int> per_batch;
  // sharded graphs with the tiled layout: map the exchange buffer, move z / part_rz into it and shard the preconditioner by path
  bool shard_paths(BaDev& d) override {
    if (world <= 1 || !d.tiled || d.n_paths < world) return false;      // (every rank must own at least one path: it raises a flag per solve)
    if (!xchg_setup(d.C, d.n_paths)) return false;
    d.xg_rank = rank; d.xg_world = world; d.xg_slots = xg.d_slots; d.xg_flags = xg.d_flags; d.xg_epoch = xg.d_epoch;
    d.xg_paths = 1; d.xg_off_z = xg.off_z; d.xg_off_prz = xg.off_prz;
    d.z = (double*)(xg.local + XG_HDR) + xg.off_z;
    d.part_rz = (double*)(xg.local + XG_HDR) + xg.off_prz;
    return true;
  }
  void drop_graphs() {
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    graphs.clear(); per_batch.clear();
  }
  void release(BaDev& d) override {
    for (auto it = graphs.begin(); it != graphs.end();) {
      if (it->first.first == (const void*)d.scal) { cudaGraphExecDestroy(it->second); per_batch.erase(it->first); it = graphs.erase(it); } else ++it;
    }
    if (cur_scal == d.scal) cur_scal = nullptr;
  }
  // dense reduced system + tensor-core Cholesky (small static-only graphs)
  int dense_capacity() const override { return DENSE_MAX; }
  bool dense_solve(BaDev& d, double lambda) override {
    const int n = 6 * d.C, npad = (n + 7) & ~7;
    LAUNCH(k_dense_init, nblk(n * n, 128), 128, d, lambda, n);
    LAUNCH(k_dense_se3_edges, nblk(d.Ese, 64), 64, d, n);
    if (d.band) {
      // static Schur term from the band moments (one writer per block), right-hand side through the mode-0 tile kernel + finalize
      band_form(d);
      LAUNCH(k_dense_from_band, nblk(d.band_n * d.band_W, 128), 128, d, n);
      tile_schur(d, 0, -1, st);
      LAUNCH(k_tile_finalize_schur2, nblk(d.C, 128), 128, d, -1.0, d.Sdense + (size_t)n * n, 0, (const double*)nullptr);
    } else {
      LAUNCH(k_dense_schur, nblk(d.P, 128), 128, d, n);
    }
    const size_t smem = sizeof(double) * (size_t)npad * (npad + 1);
    k_dense_chol<<<1, 256, smem, st>>>(d, n); ++n_launch;
    double status = 0.0;
    d2h(&status, d.Sdense + (size_t)n * n + n, sizeof(double));
    return status == 0.0;
  }
  void apply_update(BaDev& d, double lambda, bool reortho) override { LAUNCH(k_apply_update, nblk(d.C + d.P, 128), 128, d, lambda, reortho ? 1 : 0); }
};

BaBackend* make_backend(int device, char* err, size_t errlen) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) { std::snprintf(err, errlen, "no CUDA device (%s); libvdo_b200 has no CPU path", cudaGetErrorString(e)); return nullptr; }
  if (device < 0 || device >= n) { std::snprintf(err, errlen, "device %d out of range (%d visible)", device, n); return nullptr; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { std::snprintf(err, errlen, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major != 10) { std::snprintf(err, errlen, "device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major, prop.minor); return nullptr; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { std::snprintf(err, errlen, "cudaSetDevice: %s", cudaGetErrorString(e)); return nullptr; }
  {
    auto optin = [&](const void* f, size_t bytes) { if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); };
    optin((const void*)k_tile_lin<false, true>, SMEM_LIN_ST); optin((const void*)k_tile_lin<false, false>, SMEM_LIN_ST);
    optin((const void*)k_tile_lin<true, true>, SMEM_LIN_CH); optin((const void*)k_tile_lin<true, false>, SMEM_LIN_CH);
    optin((const void*)k_tile_schur<false, 0>, SMEM_SCH_ST); optin((const void*)k_tile_schur<false, 1>, SMEM_SCH_ST); optin((const void*)k_tile_schur<false, 2>, SMEM_SCH_ST);
    optin((const void*)k_tile_precond<false>, SMEM_PRE_ST); optin((const void*)k_tile_precond<true>, SMEM_PRE_CH);
    optin((const void*)k_tile_schur2<false, 0>, smem_sch2(false, VDO_TILE_E, 255, 1)); optin((const void*)k_tile_schur2<false, 1>, smem_sch2(false, VDO_TILE_E, 255, 1));
    optin((const void*)k_band_form, smem_band(VDO_TILE_E));
    optin((const void*)k_tile_schur2<true, 0>, smem_sch2(true, VDO_TILE_E, 255, 255)); optin((const void*)k_tile_schur2<true, 1>, smem_sch2(true, VDO_TILE_E, 255, 255));
    optin((const void*)k_dense_chol, sizeof(double) * (size_t)DENSE_MAX * (DENSE_MAX + 1));
    optin((const void*)k_tile_schur<true, 0>, SMEM_SCH_CH); optin((const void*)k_tile_schur<true, 1>, SMEM_SCH_CH); optin((const void*)k_tile_schur<true, 2>, SMEM_SCH_CH);
  }
  CudaBackend* b = new CudaBackend;
  b->dev = device;
  if ((e = cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking)) != cudaSuccess) { std::snprintf(err, errlen, "cudaStreamCreate: %s", cudaGetErrorString(e)); delete b; return nullptr; }
  for (int i = 0; i < 4; ++i) { cudaEventCreate(&b->ev0[i]); cudaEventCreate(&b->ev1[i]); }
  cudaStreamCreateWithFlags(&b->st2, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming);
  return b;
}

}  // namespace vdo

// ---- multi-GPU bootstrap (C ABI, declared in include/vdo_b200.h) ----
struct vdo_ctx;
namespace vdo { BaBackend* ctx_backend(vdo_ctx* c); }
extern "C" int vdo_nccl_unique_id(char* out128) {
  if (!out128) return -2;
  if (!vdo::g_nccl.load()) return -5;
  ncclUniqueId id;
  if (vdo::g_nccl.GetUniqueId(&id) != ncclSuccess) return -5;
  std::memcpy(out128, &id, sizeof id);
  return 0;
}
extern "C" int vdo_ctx_init_comm(vdo_ctx* ctx, int rank, int world, const char* id128) {
  vdo::CudaBackend* be = static_cast<vdo::CudaBackend*>(vdo::ctx_backend(ctx));
  if (!be || !id128 || world < 1 || rank < 0 || rank >= world) return -2;
  if (world == 1) { be->rank = 0; be->world = 1; return 0; }
  if (!vdo::g_nccl.load()) return -5;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  cudaSetDevice(be->dev);
  ncclResult_t r = vdo::g_nccl.CommInitRank(&be->comm, world, id, rank);
  if (r != ncclSuccess) { std::fprintf(stderr, "[vdo_b200] ncclCommInitRank failed: %s\n", vdo::g_nccl.GetErrorString ? vdo::g_nccl.GetErrorString(r) : "?"); return -5; }
  be->rank = rank; be->world = world;
  return 0;
}
