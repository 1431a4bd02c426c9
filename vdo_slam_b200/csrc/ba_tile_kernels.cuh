// ba_tile_kernels.cuh -- sm_100a kernels of the tiled batch-LM layout (included by ba_kernels.cu; bodies in ba_tiles.cuh).
//
// One CTA (VDO_TILE_L = 256 threads) per tile.  Phases:
//   k_tile_lin     stage landmark blocks in smem -> per edge: residual, Huber weight (written once to HBM), e_w stash ->
//                  per landmark: H_ll / b_l -> per vertex-sorted segment: 16 world-frame sums, warp-transpose reduction, atomics
//   k_tile_precond stage -> per segment: 10 sums of the diagonal blocks of Hpl Hll^-1 Hlp
//   k_tile_schur   stage -> per edge / landmark: Hlp v -> tracklet solve in smem (chains: scalar tridiagonal in the Q-rotated
//                  frame) -> per segment: Hpl z, 6 sums.  z never leaves the SM for modes 0 / 1.
// Bytes per launch (algorithmic, every array touched once): see bench.py kernel_bytes and DESIGN.md section 5.
#pragma once
#include "ba_tiles.cuh"

namespace vdo {

enum { SM_P = 1, SM_OM = 2, SM_EW = 4, SM_Z = 8, SM_IS = 16, SM_F = 32, SM_OMT = 64, SM_Y = 128, SM_QS = 256, SM_TC = 512, SM_E2 = 1024,
       SM_HH = 2048, SM_LML = 4096 };

template <int FL>
__host__ __device__ constexpr size_t tile_smem_bytes() {
  size_t n = 0;
  n += (FL & SM_P) ? 3 * VDO_TILE_L * 8 : 0;
  n += (FL & SM_OM) ? VDO_TILE_E * 8 : 0;
  n += (FL & SM_EW) ? 3 * VDO_TILE_E * 8 : 0;
  n += (FL & SM_Z) ? 3 * VDO_TILE_L * 8 : 0;
  n += (FL & SM_IS) ? VDO_TILE_L * 8 : 0;
  n += (FL & SM_F) ? VDO_TILE_L * 8 : 0;
  n += (FL & SM_OMT) ? VDO_TILE_L * 8 : 0;
  n += (FL & SM_Y) ? 3 * VDO_TILE_L * 8 : 0;
  n += (FL & SM_QS) ? 9 * VDO_TILE_L * 8 : 0;
  n += (FL & SM_TC) ? 4 * VDO_TILE_L * 8 : 0;
  n += (FL & SM_E2) ? 3 * VDO_TILE_L * 8 : 0;
  n += (FL & SM_HH) ? VDO_TILE_L * 4 : 0;
  n += (FL & SM_LML) ? VDO_TILE_E : 0;
  return n;
}
template <int FL>
__device__ __forceinline__ void tile_carve(TileSm& sm, double* b) {
  if (FL & SM_P) { sm.P = b; b += 3 * VDO_TILE_L; }
  if (FL & SM_OM) { sm.OM = b; b += VDO_TILE_E; }
  if (FL & SM_EW) { sm.EW = b; b += 3 * VDO_TILE_E; }
  if (FL & SM_Z) { sm.Z = b; b += 3 * VDO_TILE_L; }
  if (FL & SM_IS) { sm.IS = b; b += VDO_TILE_L; }
  if (FL & SM_F) { sm.F = b; b += VDO_TILE_L; }
  if (FL & SM_OMT) { sm.OMT = b; b += VDO_TILE_L; }
  if (FL & SM_Y) { sm.Y = b; b += 3 * VDO_TILE_L; }
  if (FL & SM_QS) { sm.QS = b; b += 9 * VDO_TILE_L; }
  if (FL & SM_TC) { sm.TC = b; b += 4 * VDO_TILE_L; }
  if (FL & SM_E2) { sm.E2 = b; b += 3 * VDO_TILE_L; }
  int* ib = reinterpret_cast<int*>(b);
  if (FL & SM_HH) { sm.HH = ib; ib += VDO_TILE_L; }
  if (FL & SM_LML) { sm.LML = reinterpret_cast<uint8_t*>(ib); }
}

// Reduce N (power of two, <= 32) per-lane values over the warp with N/2 + N/4 + ... + 1 (+ log2(32/N)) shuffles instead of
// 5 N: at each level a lane keeps one half of its values and hands the other half to its partner.  On return v[0] is the
// warp total of value `idx`; lanes with (lane & (32/N - 1)) == 0 hold the N distinct totals.
template <int N>
__device__ __forceinline__ double warp_transpose_reduce(double (&v)[N], int lane, int& idx) {
  int off = 16;
  idx = 0;
#pragma unroll
  for (int n = N; n > 1; n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const double send = up ? v[i] : v[i + n / 2];
      const double keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    idx = (idx << 1) | (up ? 1 : 0);
    off >>= 1;
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}
template <int N, int NUSED>
__device__ __forceinline__ void seg_flush(double (&acc)[N], int lane, double* dst) {
  int idx;
  const double tot = warp_transpose_reduce<N>(acc, lane, idx);
  if ((lane & (32 / N - 1)) == 0 && idx < NUSED && tot != 0.0) atomicAdd(dst + idx, tot);
}

// -------------------------------------------------------------------------------------------------------------------------
constexpr int LIN_ST = SM_P | SM_OM | SM_EW | SM_LML;
constexpr int LIN_CH = LIN_ST | SM_TC | SM_E2 | SM_OMT | SM_HH;

template <bool CHAINS, bool WRITE>
__global__ void __launch_bounds__(VDO_TILE_L) k_tile_lin(BaDev d, int tile0) {
  extern __shared__ double tile_sh[];
  __shared__ double red[32];
  TileSm sm;
  tile_carve<CHAINS ? LIN_CH : LIN_ST>(sm, tile_sh);
  const Tile tl = d.tiles[tile0 + blockIdx.x];
  const int nl = tl.k1 - tl.k0, ne = tl.e1 - tl.e0, tid = threadIdx.x;
  if (tid < nl) tile_stage_p(d, tl, tid, sm);
  __syncthreads();
  double chi = 0.0;
  if (!CHAINS) {
    for (int i = tid; i < ne; i += VDO_TILE_L) chi += tile_lin_edge<WRITE>(d, tl, i, d.lm_lml[(size_t)tl.e0 + i], sm);
    if (WRITE) {
      __syncthreads();
      if (tid < nl) {
        double dsum = 0.0, b[3] = {0, 0, 0};
        tile_lin_landmark_obs(d, tl, tid, sm, dsum, b);
        const size_t k = (size_t)tl.k0 + tid;
        d.tk_omega[k] = 0.0; d.hll[k] = dsum; d.bl[3 * k] = b[0]; d.bl[3 * k + 1] = b[1]; d.bl[3 * k + 2] = b[2];
      }
    }
  } else {
    double dsum = 0.0, b[3] = {0, 0, 0};
    if (tid < nl) {
      const int k = tl.k0 + tid;
      const int ib = d.lm_obs_begin[k] - tl.e0, ie = d.lm_obs_begin[k + 1] - tl.e0;
      for (int i = ib; i < ie; ++i) chi += tile_lin_edge<WRITE>(d, tl, i, tid, sm);
      if (WRITE) tile_lin_landmark_obs(d, tl, tid, sm, dsum, b);
      chi += tile_lin_ternary<WRITE>(d, tl, tid, sm, dsum, b);
    }
    if (WRITE) {
      __syncthreads();
      if (tid < nl) {
        if (tid > 0) { dsum += sm.TC[4 * tid - 4]; b[0] += sm.TC[4 * tid - 3]; b[1] += sm.TC[4 * tid - 2]; b[2] += sm.TC[4 * tid - 1]; }
        const size_t k = (size_t)tl.k0 + tid;
        d.hll[k] = dsum; d.bl[3 * k] = b[0]; d.bl[3 * k + 1] = b[1]; d.bl[3 * k + 2] = b[2];
      }
      if (tid < tl.t1 - tl.t0) tile_chain_Q(d, tl, tid);
    }
  }
  if (WRITE) {
    const int lane = tid & 31, warp = tid >> 5;
    for (int s = tl.os0 + warp; s < tl.os1; s += VDO_TILE_L / 32) {
      const Seg sg = d.osegs[s];
      const double* T = d.se3 + 12 * (size_t)sg.v;
      const double t[3] = {T[9], T[10], T[11]};
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
      if (lane < sg.n) tile_lin_oseg_item(d, tl, sg, lane, sm, t, acc);
      seg_flush<16, 16>(acc, lane, d.accO + 16 * (size_t)sg.v);
    }
    if (CHAINS) {
      for (int s = tl.ts0 + warp; s < tl.ts1; s += VDO_TILE_L / 32) {
        const Seg sg = d.tsegs[s];
        const double* T = d.se3 + 12 * (size_t)sg.v;
        const double t[3] = {T[9], T[10], T[11]};
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0;
        if (lane < sg.n) tile_lin_tseg_item(d, tl, sg, lane, sm, t, acc);
        seg_flush<16, 16>(acc, lane, d.accT + 16 * (size_t)sg.v);
      }
    }
  }
  chi = block_sum(chi, red);
  if (tid == 0 && chi != 0.0) atomicAdd(d.scal + SC_CHI2, chi);
}

constexpr int PRE_ST = SM_P | SM_IS;
constexpr int PRE_CH = SM_P | SM_IS | SM_F | SM_OMT;
template <bool CHAINS>
__global__ void __launch_bounds__(VDO_TILE_L) k_tile_precond(BaDev d, int tile0) {
  extern __shared__ double tile_sh[];
  TileSm sm;
  tile_carve<CHAINS ? PRE_CH : PRE_ST>(sm, tile_sh);
  const Tile tl = d.tiles[tile0 + blockIdx.x];
  const int nl = tl.k1 - tl.k0, tid = threadIdx.x;
  if (tid < nl) {
    tile_stage_p(d, tl, tid, sm);
    sm.IS[tid] = d.pt_g[tl.k0 + tid];
    if (CHAINS) { sm.F[tid] = d.tk_gamma[tl.k0 + tid]; sm.OMT[tid] = d.tk_omega[tl.k0 + tid]; }
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int s = tl.os0 + warp; s < tl.os1; s += VDO_TILE_L / 32) {
    const Seg sg = d.osegs[s];
    const double* T = d.se3 + 12 * (size_t)sg.v;
    const double t[3] = {T[9], T[10], T[11]};
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    if (lane < sg.n) tile_pre_oseg_item(d, tl, sg, lane, sm, t, acc);
    seg_flush<16, 10>(acc, lane, d.accO + 16 * (size_t)sg.v);
  }
  if (CHAINS) {
    for (int s = tl.ts0 + warp; s < tl.ts1; s += VDO_TILE_L / 32) {
      const Seg sg = d.tsegs[s];
      const double* T = d.se3 + 12 * (size_t)sg.v;
      const double t[3] = {T[9], T[10], T[11]};
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
      if (lane < sg.n) tile_pre_tseg_item(d, tl, sg, lane, sm, t, acc);
      seg_flush<16, 10>(acc, lane, d.accT + 16 * (size_t)sg.v);
    }
  }
}

constexpr int SCH_ST = SM_P | SM_OM | SM_EW | SM_Z | SM_LML;
constexpr int SCH_CH = SM_P | SM_OM | SM_Z | SM_IS | SM_F | SM_OMT | SM_Y | SM_QS | SM_HH | SM_LML;
template <bool CHAINS, int MODE>
__global__ void __launch_bounds__(VDO_TILE_L) k_tile_schur(BaDev d, int tile0) {
  extern __shared__ double tile_sh[];
  if (MODE == 1 && d.scal[SC_DONE] != 0.0) return;
  TileSm sm;
  tile_carve<CHAINS ? SCH_CH : SCH_ST>(sm, tile_sh);
  const Tile tl = d.tiles[tile0 + blockIdx.x];
  const int nl = tl.k1 - tl.k0, ne = tl.e1 - tl.e0, tid = threadIdx.x;
  if (!CHAINS) {
    if (tid < nl) tile_stage_p(d, tl, tid, sm);
    __syncthreads();
    for (int i = tid; i < ne; i += VDO_TILE_L) tile_schur_edge<MODE>(d, tl, i, d.lm_lml[(size_t)tl.e0 + i], sm);
    __syncthreads();
    if (tid < nl) tile_schur_static_landmark<MODE>(d, tl, tid, sm);
  } else {
    if (tid < nl) tile_schur_chain_stage(d, tl, tid, sm);
    __syncthreads();
    if (tid < nl) tile_schur_chain_u<MODE>(d, tl, tid, sm);
    __syncthreads();
    if (tid < nl) tile_schur_chain_y<MODE>(d, tl, tid, sm);
    __syncthreads();
    if (tid < tl.t1 - tl.t0) tile_schur_chain_walk(d, tl, tid, sm);
    __syncthreads();
    if (tid < nl) tile_schur_chain_z<MODE>(d, tl, tid, sm);
  }
  if (MODE == 2) return;
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int s = tl.os0 + warp; s < tl.os1; s += VDO_TILE_L / 32) {
    const Seg sg = d.osegs[s];
    const double* T = d.se3 + 12 * (size_t)sg.v;
    const double t[3] = {T[9], T[10], T[11]};
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0;
    if (lane < sg.n) tile_schur_oseg_item(d, tl, sg, lane, sm, t, acc);
    seg_flush<8, 6>(acc, lane, d.acc6 + 6 * (size_t)sg.v);
  }
  if (CHAINS) {
    for (int s = tl.ts0 + warp; s < tl.ts1; s += VDO_TILE_L / 32) {
      const Seg sg = d.tsegs[s];
      const double* T = d.se3 + 12 * (size_t)sg.v;
      const double t[3] = {T[9], T[10], T[11]};
      double acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0;
      if (lane < sg.n) tile_schur_tseg_item(d, tl, sg, lane, sm, t, acc);
      seg_flush<8, 6>(acc, lane, d.acc6 + 6 * (size_t)sg.v);
    }
  }
}

__global__ void __launch_bounds__(128) k_tile_finalize_lin(BaDev d) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < d.C) tile_finalize_lin(d, v);
}
__global__ void __launch_bounds__(128) k_tile_finalize_precond(BaDev d) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < d.C) tile_finalize_precond(d, v);
}
__global__ void __launch_bounds__(128) k_tile_finalize_schur(BaDev d, double sign, double* __restrict__ out, int check_done) {
  if (check_done && d.scal[SC_DONE] != 0.0) return;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < d.C) tile_finalize_schur(d, v, sign, out);
}

}  // namespace vdo
