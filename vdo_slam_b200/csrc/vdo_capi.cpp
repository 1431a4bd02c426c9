// vdo_capi.cpp -- extern "C" boundary of libvdo_b200.so (declarations and reference citations: include/vdo_b200.h).
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/vdo_b200.h"
#include "ba_driver.h"

struct vdo_ctx {
  vdo::BaBackend* be = nullptr;
  std::string err;
};
struct vdo_graph {
  vdo_ctx* ctx;
  vdo::BaGraph* g;
};

namespace vdo { BaBackend* ctx_backend(vdo_ctx* c) { return c ? c->be : nullptr; } }

extern "C" {

int vdo_ctx_create(int device, vdo_ctx** out) {
  if (!out) return VDO_ERR_ARG;
  *out = nullptr;
  vdo_ctx* c = new (std::nothrow) vdo_ctx;
  if (!c) return VDO_ERR_ARG;
  char msg[512] = {0};
  c->be = vdo::make_backend(device, msg, sizeof msg);
  if (!c->be) {
    // no CPU fallback: the caller gets an error, and the message through a static buffer
    static thread_local std::string last;
    last = msg;
    delete c;
    std::fprintf(stderr, "[vdo_b200] vdo_ctx_create failed: %s\n", last.c_str());
    return VDO_ERR_CUDA;
  }
  *out = c;
  return VDO_OK;
}
void vdo_ctx_destroy(vdo_ctx* ctx) {
  if (!ctx) return;
  delete ctx->be;
  delete ctx;
}
const char* vdo_last_error(const vdo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t vdo_ctx_stream(const vdo_ctx* ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->be->stream() : 0; }

int vdo_graph_create(vdo_ctx* ctx, vdo_graph** out) {
  if (!ctx || !out) return VDO_ERR_ARG;
  vdo_graph* g = new vdo_graph{ctx, new vdo::BaGraph(ctx->be)};
  *out = g;
  return VDO_OK;
}
void vdo_graph_destroy(vdo_graph* g) {
  if (!g) return;
  delete g->g;
  delete g;
}
#define VDO_FWD(call)                    \
  if (!g) return VDO_ERR_ARG;            \
  int rc_ = g->g->call;                  \
  if (rc_ != VDO_OK) g->ctx->err = g->g->error(); \
  return rc_;

int vdo_graph_set_vertices(vdo_graph* g, int n_se3, const double* se3, int n_pt, const double* pt) { VDO_FWD(set_vertices(n_se3, se3, n_pt, pt)) }
int vdo_graph_add_edges_se3_prior(vdo_graph* g, int n, const int* v, const double* Z, const double* w) { VDO_FWD(add_prior(n, v, Z, w)) }
int vdo_graph_add_edges_se3(vdo_graph* g, int n, const int* ij, const double* Z, const double* w, const double* delta) { VDO_FWD(add_se3(n, ij, Z, w, delta)) }
int vdo_graph_add_edges_se3_pointxyz(vdo_graph* g, int n, const int* cp, const double* z, const double* w, const double* delta) { VDO_FWD(add_obs(n, cp, z, w, delta)) }
int vdo_graph_add_edges_landmark_motion(vdo_graph* g, int n, const int* pph, const double* w, const double* delta) { VDO_FWD(add_ter(n, pph, w, delta)) }
int vdo_graph_finalize(vdo_graph* g) { VDO_FWD(finalize()) }

int vdo_abi_struct_size(const char* name) {
  if (!name) return -1;
  const std::string s(name);
  if (s == "vdo_lm_options") return (int)sizeof(vdo_lm_options);
  if (s == "vdo_lm_stats") return (int)sizeof(vdo_lm_stats);
  if (s == "vdo_tracker_params") return (int)sizeof(vdo_tracker_params);
  return -1;
}

void vdo_lm_options_default(vdo_lm_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);
  o->max_iterations = 300; o->gain_threshold = 1e-4; o->max_trials = 10;
  o->pcg_rel_tol = 1e-6; o->pcg_max_iterations = 2000; o->verbose = 0; o->force_all_iterations = 0; o->pcg_loose_tol = 0.0; o->pcg_switch_gain = 0.0;
}
int vdo_graph_optimize(vdo_graph* g, const vdo_lm_options* opt, vdo_lm_stats* stats, double* chi2_history) {
  vdo_lm_options o;
  if (opt) o = *opt; else vdo_lm_options_default(&o);
  VDO_FWD(optimize(o, stats, chi2_history))
}
int vdo_graph_get_vertices(const vdo_graph* g, double* se3, double* pt) { VDO_FWD(get_vertices(se3, pt)) }
int vdo_graph_reset_vertices(vdo_graph* g) { VDO_FWD(reset_vertices()) }
int vdo_graph_info(const vdo_graph* g, int64_t out[8]) { VDO_FWD(info(out)) }
int vdo_graph_solver_info(const vdo_graph* g, int64_t out[8]) { VDO_FWD(solver_info(out)) }
int vdo_graph_debug_linearize(vdo_graph* g, double* Hpp, double* bp, double* Hll, double* bl, double* chi2) { VDO_FWD(debug_linearize(Hpp, bp, Hll, bl, chi2)) }

int vdo_graph_time_kernel(vdo_graph* g, const char* name, int reps, float* ms_avg) { VDO_FWD(time_kernel(name, reps, ms_avg)) }

}  // extern "C"
