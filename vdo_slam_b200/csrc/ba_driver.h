// ba_driver.h -- host side of the batch factor-graph optimiser: graph ingestion (the analogue of
// g2o::SparseOptimizer::initializeOptimization + BlockSolver::buildStructure) and the Levenberg-Marquardt loop
// (g2o/core/optimization_algorithm_levenberg.cpp:61-164, sparse_optimizer.cpp:354-427) driving backend kernels.
#pragma once
#include <string>
#include <vector>
#include "ba_types.h"
#include "../../include/vdo_b200.h"

namespace vdo {

// array in the backend's staging arena: plain memory, NOT zero-initialised, valid until the graph releases the arena
template <typename T> struct HostBuf {
  T* p = nullptr; size_t n = 0;
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T* begin() { return p; }
  T* end() { return p + n; }
};

class BaGraph {
 public:
  explicit BaGraph(BaBackend* be) : be_(be) {}
  ~BaGraph();
  int set_vertices(int n_se3, const double* se3, int n_pt, const double* pt);
  int add_prior(int n, const int* v, const double* Z, const double* w);
  int add_se3(int n, const int* ij, const double* Z, const double* w, const double* delta);
  int add_obs(int n, const int* cp, const double* z, const double* w, const double* delta);
  int add_ter(int n, const int* pph, const double* w, const double* delta);
  int finalize();
  int optimize(const vdo_lm_options& opt, vdo_lm_stats* stats, double* chi2_history);
  int get_vertices(double* se3, double* pt);
  int reset_vertices();
  int info(int64_t out[8]) const;
  int solver_info(int64_t out[8]) const;
  int debug_linearize(double* Hpp, double* bp, double* Hll, double* bl, double* chi2);
  int time_kernel(const char* name, int reps, float* ms_avg);
  const std::string& error() const { return err_; }

 private:
  template <typename T> T* dalloc(size_t n) { bytes_ += n * sizeof(T); T* p = (T*)be_->alloc(n * sizeof(T) + 16); owned_.push_back(p); return p; }   // +16: the tile kernels' bulk copies round ranges up to 16 B
  template <typename T> HostBuf<T> stage(size_t n) { hold_stage(); HostBuf<T> b; b.n = n; b.p = (T*)be_->staging().take(n * sizeof(T) + 16); return b; }
  template <typename T> HostBuf<T> stage_fill(size_t n, int byte) { HostBuf<T> b = stage<T>(n); fill_bytes(b.p, byte, n * sizeof(T)); return b; }
  template <typename T> void append(HostBuf<T>& b, const T* src, size_t n) {
    HostBuf<T> nb = stage<T>(b.n + n);
    if (b.n) copy_bytes(nb.p, b.p, b.n * sizeof(T));
    copy_bytes(nb.p + b.n, src, n * sizeof(T));
    b = nb;
  }
  template <typename T> T* upload(const HostBuf<T>& v) { T* p = dalloc<T>(v.size()); if (!v.empty()) be_->h2d_async(p, v.data(), v.size() * sizeof(T)); return p; }
  void hold_stage() { if (!holds_stage_) { be_->staging().acquire(); holds_stage_ = true; } }
  void drop_stage() { if (holds_stage_) { be_->staging().release(); holds_stage_ = false; } }
  static void copy_bytes(void* dst, const void* src, size_t bytes);   // threaded for large blocks
  static void fill_bytes(void* dst, int byte, size_t bytes);
  bool holds_stage_ = false;
  template <typename T> T* upload(const std::vector<T>& v) { T* p = dalloc<T>(v.size()); if (!v.empty()) be_->h2d(p, v.data(), v.size() * sizeof(T)); return p; }
  void linearize();                 // buildSystem
  double robust_chi2();             // computeActiveErrors + activeRobustChi2
  bool solve(double lambda, const vdo_lm_options& opt, int* pcg_iters);   // Schur + PCG + back-substitution -> xp, xl
  int fail(int code, const std::string& m) { err_ = m; return code; }

  BaBackend* be_;
  BaDev d_;
  std::vector<void*> owned_;
  size_t bytes_ = 0;
  bool finalized_ = false;
  std::string err_;
  long oplus_calls_ = 0;
  bool prof_on_ = false;
  float prof_ms_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double last_lambda_ = 1.0;
  double cur_pcg_tol_ = 0.0;        // tolerance of the current LM iteration's solves (forcing schedule), 0 = opt.pcg_rel_tol
  // host staging (until finalize)
  int n_se3_ = 0, n_pt_ = 0, P_all_ = 0;
  HostBuf<double> h_se3_, h_pt_;
  std::vector<int> pr_v_; std::vector<double> pr_Z_, pr_w_;
  std::vector<int> se_ij_; std::vector<double> se_Z_, se_w_, se_d_;
  HostBuf<int> ob_cp_; HostBuf<double> ob_z_, ob_w_, ob_d_;
  HostBuf<int> te_pph_; HostBuf<double> te_w_, te_d_;
  int n_prior_ = 0;
  std::vector<int> new_of_old_;     // landmark renumbering
  std::vector<int> new_se3_of_old_; // se3 renumbering (path order)
};

}  // namespace vdo
