// tests/emul/ba_backend_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A serial stand-in for the CUDA backend (vdo_slam_b200/csrc/ba_kernels.cu): "device memory" is malloc, every
// "kernel" is a for-loop over the same VDO_HD per-thread bodies (ba_bodies.cuh) the CUDA kernels call, and the
// reductions the kernels do with warp shuffles + atomics are plain sums.  It exists so the -m "not gpu" test-suite
// can exercise the host logic of the product (graph ingestion, tracklet ordering, LM / Schur / PCG driver in
// ba_driver.cpp, the C ABI glue in vdo_capi.cpp) in a container without a GPU.  It is compiled into
// tests/emul/libvdo_emul.so, never into libvdo_b200.so; the product has no CPU path.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <vector>

#include "../../vdo_slam_b200/csrc/ba_bodies.cuh"
#include "../../vdo_slam_b200/csrc/ba_tiles.cuh"

namespace vdo {

typedef void (*vdo_collective_fn)(double* buf, size_t n, int op, void* user);
struct EmulBackend : BaBackend {
  int n_launch = 0;
  vdo_collective_fn coll = nullptr; void* coll_user = nullptr;
  void allreduce_sum(double* b, size_t n) override { if (world > 1 && coll) coll(b, n, 0, coll_user); }
  void allreduce_max(double* b, size_t n) override { if (world > 1 && coll) coll(b, n, 1, coll_user); }
  std::chrono::steady_clock::time_point t0[4];
  struct MallocArena : HostArena {
    char* raw_alloc(size_t b) override { return (char*)std::malloc(b); }
    void raw_free(char* p) override { std::free(p); }
    ~MallocArena() override { destroy(); }
  } arena;
  HostArena& staging() override { return arena; }
  void* alloc(size_t b) override { return std::calloc(1, b ? b : 1); }
  void free_(void* p) override { std::free(p); }
  void h2d(void* d, const void* s, size_t b) override { std::memcpy(d, s, b); }
  void d2h(void* d, const void* s, size_t b) override { std::memcpy(d, s, b); }
  void d2d(void* d, const void* s, size_t b) override { std::memcpy(d, s, b); }
  void zero(void* d, size_t b) override { std::memset(d, 0, b); }
  void sync() override {}
  int launches() const override { return n_launch; }
  void* stream() const override { return nullptr; }
  void timer_start(int s) override { t0[s] = std::chrono::steady_clock::now(); }
  float timer_stop_ms(int s) override { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0[s]).count(); }


  // ---- tiled layout: the phases of ba_tile_kernels.cuh as serial loops (segments summed lane by lane) ----
  struct SmBuf {   // stashes only; the views alias the global arrays (tile_views_global)
    std::vector<double> OMs, EW, Z, Y, IS, TC, E2, OMTs;
    TileSm sm;
    SmBuf() : OMs(VDO_TILE_E), EW(3 * VDO_TILE_E), Z(3 * VDO_TILE_L), Y(3 * VDO_TILE_L), IS(VDO_TILE_L), TC(4 * VDO_TILE_L), E2(3 * VDO_TILE_L), OMTs(VDO_TILE_L) {}
    TileSm& bind(const BaDev& d, const Tile& tl, bool precond, bool lin) {
      tile_views_global(d, tl, precond, sm);
      sm.EW = EW.data(); sm.Z = Z.data(); sm.Y = Y.data(); sm.IS = IS.data(); sm.TC = TC.data(); sm.E2 = E2.data();
      if (lin) { sm.OM = OMs.data(); sm.OMT = OMTs.data(); }
      return sm;
    }
  } sb;
  template <int N, typename F> static void seg_loop(const BaDev& d, const Seg* segs, int s0, int s1, double* dst_base, int stride, int nused, F item) {
    for (int s = s0; s < s1; ++s) {
      const Seg sg = segs[s];
      const double* T = d.se3 + 12 * (size_t)sg.v;
      const double t[3] = {T[9], T[10], T[11]};
      double acc[N] = {0};
      for (int l = 0; l < sg.n; ++l) item(sg, l, t, acc);
      for (int i = 0; i < nused; ++i) dst_base[(size_t)stride * sg.v + i] += acc[i];
    }
  }
  void tile_lin(BaDev& d, bool write) {
    double chi = 0;
    for (int ti = 0; ti < d.n_tiles; ++ti) {
      const Tile tl = d.tiles[ti];
      const bool chains = ti >= d.n_tiles_stat;
      const int nl = tl.k1 - tl.k0, ne = tl.e1 - tl.e0;
      TileSm& sm = sb.bind(d, tl, false, true);
      if (!chains) {
        for (int i = 0; i < ne; ++i) chi += write ? tile_lin_edge<true>(d, tl, i, sm.LML[i], sm) : tile_lin_edge<false>(d, tl, i, sm.LML[i], sm);
        if (write) for (int j = 0; j < nl; ++j) {
          double dsum = 0, b[3] = {0, 0, 0};
          tile_lin_landmark_obs(d, tl, j, sm, dsum, b);
          const size_t k = (size_t)tl.k0 + j;
          d.tk_omega[k] = 0; d.hll[k] = dsum; d.bl[3 * k] = b[0]; d.bl[3 * k + 1] = b[1]; d.bl[3 * k + 2] = b[2];
        }
      } else {
        std::vector<double> ds(nl, 0.0), bb(3 * (size_t)nl, 0.0);
        for (int j = 0; j < nl; ++j) {
          for (int i = sm.LB[j] - tl.e0; i < sm.LB[j + 1] - tl.e0; ++i) chi += write ? tile_lin_edge<true>(d, tl, i, j, sm) : tile_lin_edge<false>(d, tl, i, j, sm);
          if (write) tile_lin_landmark_obs(d, tl, j, sm, ds[j], &bb[3 * j]);
          chi += write ? tile_lin_ternary<true>(d, tl, j, sm, ds[j], &bb[3 * j]) : tile_lin_ternary<false>(d, tl, j, sm, ds[j], &bb[3 * j]);
        }
        if (write) {
          for (int j = 0; j < nl; ++j) {
            if (j > 0) { ds[j] += sm.TC[4 * j - 4]; bb[3 * j] += sm.TC[4 * j - 3]; bb[3 * j + 1] += sm.TC[4 * j - 2]; bb[3 * j + 2] += sm.TC[4 * j - 1]; }
            const size_t k = (size_t)tl.k0 + j;
            d.hll[k] = ds[j]; d.bl[3 * k] = bb[3 * j]; d.bl[3 * k + 1] = bb[3 * j + 1]; d.bl[3 * k + 2] = bb[3 * j + 2];
          }
          for (int jt = 0; jt < tl.t1 - tl.t0; ++jt) tile_chain_Q(d, tl, jt);
        }
      }
      if (write) {
        seg_loop<16>(d, d.osegs, tl.os0, tl.os1, d.accO, 16, 16, [&](const Seg& sg, int l, const double* t, double* acc) { tile_lin_oseg_item(d, tl, sg, l, sm, t, acc); });
        if (chains) seg_loop<16>(d, d.tsegs, tl.ts0, tl.ts1, d.accT, 16, 16, [&](const Seg& sg, int l, const double* t, double* acc) { tile_lin_tseg_item(d, tl, sg, l, sm, t, acc); });
      }
    }
    d.scal[SC_CHI2] += chi;
  }
  void tile_precond(BaDev& d) {
    for (int ti = 0; ti < d.n_tiles; ++ti) {
      const Tile tl = d.tiles[ti];
      const bool chains = ti >= d.n_tiles_stat;
      TileSm& sm = sb.bind(d, tl, true, false);
      seg_loop<16>(d, d.osegs, tl.os0, tl.os1, d.accO, 16, 10, [&](const Seg& sg, int l, const double* t, double* acc) { tile_pre_oseg_item(d, tl, sg, l, sm, t, acc); });
      if (chains) seg_loop<16>(d, d.tsegs, tl.ts0, tl.ts1, d.accT, 16, 10, [&](const Seg& sg, int l, const double* t, double* acc) { tile_pre_tseg_item(d, tl, sg, l, sm, t, acc); });
    }
    for (int v = 0; v < d.C; ++v) tile_finalize_precond(d, v);
  }
  // ---- banded static block of the reduced matrix, restated serially: the moments per (vertex, offset) over the static landmarks,
  //      and the product in this backend's accumulator convention (acc6 = 6 per vertex: force, torque about the VERTEX origin) ----
  int band_max_width() const override { return 32; }
  void band_form(BaDev& d) override {
    if (!d.band) return;
    ++n_launch;
    const int W = d.band_W;
    for (size_t i = 0; i < (size_t)d.band_n * W * 10; ++i) d.band[i] = 0.0;
    for (int k = 0; k < d.Tstat; ++k) {
      const int e0 = d.lm_obs_begin[k], e1 = d.lm_obs_begin[k + 1];
      const double is = 1.0 / d.pt_s[k];
      const double* p = d.pt + 3 * (size_t)k;
      const double m[10] = {1.0, p[0], p[1], p[2], p[0] * p[0], p[0] * p[1], p[0] * p[2], p[1] * p[1], p[1] * p[2], p[2] * p[2]};
      for (int i = e0; i < e1; ++i)
        for (int j = i; j < e1; ++j) {
          const double g = d.lm_omega[i] * d.lm_omega[j] * is;
          double* dst = d.band + ((size_t)(d.lm_cam[i] - d.band_v0) * W + (d.lm_cam[j] - d.lm_cam[i])) * 10;
          for (int q = 0; q < 10; ++q) dst[q] += g * m[q];
        }
    }
  }
  void band_mul(BaDev& d) {
    const int W = d.band_W;
    for (int a = 0; a < d.band_n; ++a) {
      double F[3] = {0, 0, 0}, M[3] = {0, 0, 0};
      for (int kk = -(W - 1); kk <= W - 1; ++kk) {
        const int b = a + kk;
        if (b < 0 || b >= d.band_n) continue;
        const double* m = d.band + (kk >= 0 ? ((size_t)a * W + kk) : ((size_t)b * W - kk)) * 10;
        if (m[0] == 0.0) continue;
        const double* w = d.vw + 6 * (size_t)(d.band_v0 + b);
        const double m1[3] = {m[1], m[2], m[3]};
        double c1[3], c2[3];
        cross3(m1, w + 3, c1); cross3(m1, w, c2);
        const double tr = m[4] + m[7] + m[9];
        const double q0 = m[4] * w[3] + m[5] * w[4] + m[6] * w[5] - tr * w[3], q1 = m[5] * w[3] + m[7] * w[4] + m[8] * w[5] - tr * w[4],
                     q2 = m[6] * w[3] + m[8] * w[4] + m[9] * w[5] - tr * w[5];
        F[0] -= m[0] * w[0] + 2 * c1[0]; F[1] -= m[0] * w[1] + 2 * c1[1]; F[2] -= m[0] * w[2] + 2 * c1[2];
        M[0] -= 2 * (c2[0] + 2 * q0); M[1] -= 2 * (c2[1] + 2 * q1); M[2] -= 2 * (c2[2] + 2 * q2);
      }
      const int v = d.band_v0 + a;
      const double* t = d.se3 + 12 * (size_t)v + 9;
      double txf[3]; cross3(t, F, txf);
      double* acc = d.acc6 + 6 * (size_t)v;
      acc[0] += F[0]; acc[1] += F[1]; acc[2] += F[2];
      acc[3] += M[0] - 2 * txf[0]; acc[4] += M[1] - 2 * txf[1]; acc[5] += M[2] - 2 * txf[2];       // torque moved from the world origin to the vertex origin
    }
  }
  template <int MODE> void tile_schur(BaDev& d) {
    if (MODE == 1 && d.band) band_mul(d);
    for (int ti = 0; ti < d.n_tiles; ++ti) {
      const Tile tl = d.tiles[ti];
      const bool chains = ti >= d.n_tiles_stat;
      if (MODE == 1 && d.band && !chains) continue;           // the static tiles' product comes from the band
      const int nl = tl.k1 - tl.k0, ne = tl.e1 - tl.e0;
      TileSm& sm = sb.bind(d, tl, false, false);
      if (!chains) {
        for (int i = 0; i < ne; ++i) tile_schur_edge<MODE>(d, tl, i, sm);
        for (int j = 0; j < nl; ++j) tile_schur_static_landmark<MODE>(d, tl, j, sm);
      } else {
        for (int j = 0; j < nl; ++j) tile_schur_chain_u<MODE>(d, tl, j, sm);
        for (int j = 0; j < nl; ++j) tile_schur_chain_y<MODE>(d, tl, j, sm);
        for (int jt = 0; jt < tl.t1 - tl.t0; ++jt) tile_schur_chain_walk(d, tl, jt, sm);
        for (int j = 0; j < nl; ++j) tile_schur_chain_z<MODE>(d, tl, j, sm);
      }
      if (MODE == 2) continue;
      seg_loop<8>(d, d.osegs, tl.os0, tl.os1, d.acc6, 6, 6, [&](const Seg& sg, int l, const double* t, double* acc) { tile_schur_oseg_item(d, tl, sg, l, sm, t, acc); });
      if (chains) seg_loop<8>(d, d.tsegs, tl.ts0, tl.ts1, d.acc6, 6, 6, [&](const Seg& sg, int l, const double* t, double* acc) { tile_schur_tseg_item(d, tl, sg, l, sm, t, acc); });
    }
  }
  void lin_tracklets(BaDev& d, bool write) override {
    ++n_launch;
    if (d.tiled) { tile_lin(d, write); return; }
    double chi = 0;
    for (int t = 0; t < d.Tstat; ++t) chi += body_lin_static(d, t, write);
    for (int t = d.Tstat; t < d.T; ++t) chi += body_lin_tracklet(d, t, write);
    d.scal[SC_CHI2] += chi;
  }
  static void add_sym(double* H36, const double* A21) {
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c) {
        double v = A21[sym6_idx(r, c)];
        H36[6 * r + c] += v;
        if (c != r) H36[6 * c + r] += v;
      }
  }
  void lin_vertex_obs(BaDev& d) override {
    ++n_launch;
    if (d.tiled) { for (int v = 0; v < d.C; ++v) tile_finalize_lin(d, v); return; }
    for (int ci = 0; ci < d.n_obs_chunks; ++ci) {
      Chunk ch = d.obs_chunks[ci];
      Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
      double A[21] = {0}, g[6] = {0};
      for (int e = ch.begin; e < ch.end; ++e) body_lin_vertex_obs(d, T, e, A, g);
      add_sym(d.Hpp + 36 * (size_t)ch.v, A);
      for (int i = 0; i < 6; ++i) d.bp[6 * (size_t)ch.v + i] += g[i];
    }
  }
  void lin_vertex_ter(BaDev& d) override {
    if (d.tiled) return;
    ++n_launch;
    for (int ci = 0; ci < d.n_ter_chunks; ++ci) {
      Chunk ch = d.ter_chunks[ci];
      Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
      double A[21] = {0}, g[6] = {0};
      for (int e = ch.begin; e < ch.end; ++e) body_lin_vertex_ter(d, T, e, A, g);
      add_sym(d.Hpp + 36 * (size_t)ch.v, A);
      for (int i = 0; i < 6; ++i) d.bp[6 * (size_t)ch.v + i] += g[i];
    }
  }
  void lin_se3_edges(BaDev& d, bool write) override {
    ++n_launch;
    double chi_tot = 0;
    for (int e = 0; e < d.Ese; ++e) {
      double chi, Hi[36], Hj[36], Ho[36], gi[6], gj[6];
      bool binary = body_se3_edge(d, e, write, chi, Hi, Hj, Ho, gi, gj);
      if (d.own) chi_tot += chi;
      if (!write) continue;
      int i = d.se_i[e];
      if (d.own) {
        for (int k = 0; k < 36; ++k) d.Hpp[36 * (size_t)i + k] += Hi[k];
        for (int k = 0; k < 6; ++k) d.bp[6 * (size_t)i + k] += gi[k];
      }
      if (binary) {
        int j = d.se_j[e];
        for (int k = 0; k < 36; ++k) { if (d.own) d.Hpp[36 * (size_t)j + k] += Hj[k]; d.se_Hoff[36 * (size_t)e + k] = Ho[k]; }
        if (d.own) for (int k = 0; k < 6; ++k) d.bp[6 * (size_t)j + k] += gj[k];
      }
    }
    d.scal[SC_CHI2] += chi_tot;
  }
  void max_diagonal(BaDev& d) override {
    ++n_launch;
    double m = 0;
    for (int v = 0; v < d.C; ++v) for (int i = 0; i < 6; ++i) m = std::fmax(m, std::fabs(d.Hpp[36 * (size_t)v + 7 * i]));
    for (int k = 0; k < d.P; ++k) m = std::fmax(m, std::fabs(d.hll[k]));
    d.scal[SC_MAXDIAG] = m;
  }
  void factor_landmarks(BaDev& d, double lambda) override { ++n_launch; for (int t = 0; t < d.Tstat; ++t) body_factor_static(d, t, lambda); for (int t = d.Tstat; t < d.T; ++t) body_factor_tracklet(d, t, lambda); }
  void precond_begin(BaDev& d, double lambda) override {
    ++n_launch;
    for (int v = 0; v < d.C; ++v) {
      for (int i = 0; i < 36; ++i) d.Minv[36 * (size_t)v + i] = d.own ? d.Hpp[36 * (size_t)v + i] : 0.0;
      if (d.own) for (int i = 0; i < 6; ++i) d.Minv[36 * (size_t)v + 7 * i] += lambda;
    }
  }
  void precond_vertex_obs(BaDev& d) override {
    ++n_launch;
    if (d.tiled) { tile_precond(d); return; }
    for (int ci = 0; ci < d.n_obs_chunks; ++ci) {
      Chunk ch = d.obs_chunks[ci];
      Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
      double A[21] = {0};
      for (int e = ch.begin; e < ch.end; ++e) body_precond_vertex_obs(d, T, e, A);
      for (int i = 0; i < 21; ++i) A[i] = -A[i];
      add_sym(d.Minv + 36 * (size_t)ch.v, A);
    }
  }
  void precond_vertex_ter(BaDev& d) override {
    if (d.tiled) return;
    ++n_launch;
    for (int ci = 0; ci < d.n_ter_chunks; ++ci) {
      Chunk ch = d.ter_chunks[ci];
      Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
      double A[21] = {0};
      for (int e = ch.begin; e < ch.end; ++e) body_precond_vertex_ter(d, T, e, A);
      for (int i = 0; i < 21; ++i) A[i] = -A[i];
      add_sym(d.Minv + 36 * (size_t)ch.v, A);
    }
  }
  void precond_factor(BaDev& d, double lambda) override {
    ++n_launch;
    const size_t N36 = 36 * (size_t)d.C;
    for (int pth = 0; pth < d.n_paths; ++pth) {
      const int pb = d.path_begin[pth], pe = d.path_begin[pth + 1], nl = pcr_num_levels(pe - pb);
      int cur = 0;
      for (int v = pb; v < pe; ++v) body_pcr_setup(d, v, d.pcr_D, d.pcr_L);
      for (int l = 0; l < nl; ++l) {
        const double *D = d.pcr_D + cur * N36, *L = d.pcr_L + cur * N36;
        double *Dn = d.pcr_D + (1 - cur) * N36, *Ln = d.pcr_L + (1 - cur) * N36;
        for (int v = pb; v < pe; ++v) { int bad = 0; body_pcr_invert(d, v, D, d.pcr_Dinv, lambda, &bad); d.scal[SC_BAD] += bad; }
        for (int v = pb; v < pe; ++v) for (int rc = 0; rc < 36; ++rc) body_pcr_AG(v, rc / 6, rc % 6, pb, pe, 1 << l, L, d.pcr_Dinv, d.pcr_A + l * N36, d.pcr_G + l * N36);
        for (int v = pb; v < pe; ++v) for (int rc = 0; rc < 36; ++rc) body_pcr_DL(v, rc / 6, rc % 6, pb, pe, 1 << l, D, L, d.pcr_A + l * N36, d.pcr_G + l * N36, Dn, Ln);
        cur = 1 - cur;
      }
      for (int v = pb; v < pe; ++v) { int bad = 0; body_pcr_invert(d, v, d.pcr_D + cur * N36, d.Minv, lambda, &bad); d.scal[SC_BAD] += bad; }
    }
  }
  // z = M^-1 r
  void precond_apply(BaDev& d, const double* r, double* z) {
    const size_t N6 = 6 * (size_t)d.C, N36 = 36 * (size_t)d.C;
    for (int pth = 0; pth < d.n_paths; ++pth) {
      const int pb = d.path_begin[pth], pe = d.path_begin[pth + 1], nl = pcr_num_levels(pe - pb);
      int cur = 0;
      for (int v = pb; v < pe; ++v) for (int i = 0; i < 6; ++i) d.pcr_b[6 * (size_t)v + i] = r[6 * (size_t)v + i];
      for (int l = 0; l < nl; ++l) {
        for (int v = pb; v < pe; ++v) body_pcr_apply(v, pb, pe, 1 << l, d.pcr_A + l * N36, d.pcr_G + l * N36, d.pcr_b + cur * N6, d.pcr_b + (1 - cur) * N6);
        cur = 1 - cur;
      }
      for (int v = pb; v < pe; ++v) mul6(d.Minv + 36 * (size_t)v, d.pcr_b + cur * N6 + 6 * (size_t)v, z + 6 * (size_t)v);
    }
  }
  void schur_landmarks(BaDev& d, int mode, const double* v) override {
    ++n_launch;
    if (mode == 1 && d.scal[SC_DONE] != 0.0) return;
    if (d.tiled) { if (mode == 0) tile_schur<0>(d); else if (mode == 1) tile_schur<1>(d); else tile_schur<2>(d); return; }
    double* out = mode == 2 ? d.xl : d.zl;
    for (int t = 0; t < d.Tstat; ++t) body_schur_static(d, t, mode, out);
    for (int t = d.Tstat; t < d.T; ++t) body_schur_tracklet(d, t, mode, v, out);
  }
  void schur_vertex_obs(BaDev& d, double sign, double* out) override {
    ++n_launch;
    if (d.tiled) { if (out == d.Ap && d.scal[SC_DONE] != 0.0) return; for (int v = 0; v < d.C; ++v) tile_finalize_schur(d, v, sign, out); return; }
    for (int ci = 0; ci < d.n_obs_chunks; ++ci) {
      Chunk ch = d.obs_chunks[ci];
      Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
      double a[6] = {0};
      for (int e = ch.begin; e < ch.end; ++e) body_schur_vertex_obs(d, T, e, a);
      for (int i = 0; i < 6; ++i) out[6 * (size_t)ch.v + i] += sign * a[i];
    }
  }
  void schur_vertex_ter(BaDev& d, double sign, double* out) override {
    if (d.tiled) return;
    ++n_launch;
    for (int ci = 0; ci < d.n_ter_chunks; ++ci) {
      Chunk ch = d.ter_chunks[ci];
      Iso T; iso_load(d.se3 + 12 * (size_t)ch.v, T);
      double a[6] = {0};
      for (int e = ch.begin; e < ch.end; ++e) body_schur_vertex_ter(d, T, e, a);
      for (int i = 0; i < 6; ++i) out[6 * (size_t)ch.v + i] += sign * a[i];
    }
  }
  void vertex_transform(BaDev& d, const double* v) override { ++n_launch; for (int c = 0; c < d.C; ++c) body_vertex_transform(d, c, v, d.vw); }
  void hpp_mul(BaDev& d, double lambda, const double* x, double* out) override {
    ++n_launch;
    for (int v = 0; v < d.C; ++v) body_hpp_mul(d, v, lambda, x, out);
  }
  void pcg_init(BaDev& d) override {
    ++n_launch;
    double rz = 0;
    for (size_t i = 0; i < 6 * (size_t)d.C; ++i) { d.r[i] = d.rhs[i]; d.xp[i] = 0; }
    precond_apply(d, d.r, d.z);
    for (size_t i = 0; i < 6 * (size_t)d.C; ++i) { d.p[i] = d.z[i]; rz += d.z[i] * d.r[i]; }
    d.scal[SC_RZ] = rz; d.scal[SC_RZ0] = rz; d.scal[SC_RZ_NEW] = 0; d.scal[SC_PAP] = 0; d.scal[SC_ITERS] = 0;
    d.scal[SC_DONE] = (rz > 0) ? 0.0 : 1.0;
  }
  void pcg_dot_pAp(BaDev& d) override {
    ++n_launch;
    if (d.scal[SC_DONE] != 0.0) return;
    double s = 0;
    for (size_t i = 0; i < 6 * (size_t)d.C; ++i) s += d.p[i] * d.Ap[i];
    d.scal[SC_PAP] = s;
  }
  void pcg_step(BaDev& d, double tol2) override {
    n_launch += 2;
    if (d.scal[SC_DONE] != 0.0) return;
    double pAp = d.scal[SC_PAP], rz = d.scal[SC_RZ];
    if (!(pAp > 0) || !std::isfinite(pAp)) { d.scal[SC_DONE] = 2.0; return; }
    double alpha = rz / pAp, rzn = 0;
    for (size_t q = 0; q < 6 * (size_t)d.C; ++q) { d.xp[q] += alpha * d.p[q]; d.r[q] -= alpha * d.Ap[q]; }
    precond_apply(d, d.r, d.z);
    for (size_t q = 0; q < 6 * (size_t)d.C; ++q) rzn += d.z[q] * d.r[q];
    double beta = rzn / rz;
    for (size_t i = 0; i < 6 * (size_t)d.C; ++i) d.p[i] = d.z[i] + beta * d.p[i];
    d.scal[SC_RZ] = rzn; d.scal[SC_ITERS] += 1;
    if (rzn <= tol2 * d.scal[SC_RZ0]) d.scal[SC_DONE] = 1.0;
    if (!std::isfinite(rzn)) d.scal[SC_DONE] = 2.0;
  }
  void apply_update(BaDev& d, double lambda, bool reortho) override {
    ++n_launch;
    double s = 0;
    for (int v = 0; v < d.C; ++v) s += body_update_se3(d, v, lambda, reortho);
    for (int k = 0; k < d.P; ++k) s += body_update_pt(d, k, lambda);
    d.scal[SC_SCALE] += s;
  }
};

BaBackend* make_backend(int, char*, size_t) { return new EmulBackend; }
// test-only: attach a host collective (e.g. torch.distributed over gloo) to the emulated backend
void emul_set_collective(BaBackend* be, int rank, int world, vdo_collective_fn fn, void* user) {
  EmulBackend* e = static_cast<EmulBackend*>(be);
  e->rank = rank; e->world = world; e->coll = fn; e->coll_user = user;
}

}  // namespace vdo

// ---- test-only C entry point: attach a host collective to an emulated context ----
struct vdo_ctx;
namespace vdo { BaBackend* ctx_backend(vdo_ctx* c); }
extern "C" int vdo_emul_set_collective(vdo_ctx* ctx, int rank, int world, vdo::vdo_collective_fn fn, void* user) {
  vdo::BaBackend* be = vdo::ctx_backend(ctx);
  if (!be || world < 1 || rank < 0 || rank >= world) return -2;
  vdo::emul_set_collective(be, rank, world, fn, user);
  return 0;
}
