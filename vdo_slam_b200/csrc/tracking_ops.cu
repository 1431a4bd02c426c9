// tracking_ops.cu -- bookkeeping stages of Tracking that sit between the kernels: tracklet chaining (A16), mask propagation
// (A15) and dynamic-object classification / ID assignment (A13).  These are integer, order-dependent procedures over a few
// thousand to a few hundred thousand elements: the image-sized work (mask gather / forward warp) and the per-object
// reductions run on the device next to the resident frame, the decisions (which are sequential by definition in the
// reference: IDs are handed out in label order) run on the host side of the call.
//
// Reference semantics (restated in oracle/tracking_ops.py):
//   Tracking::GetStaticTrack / GetDynamicTrackNew   src/Tracking.cc:2201-2307, 2309-2421
//   Tracking::UpdateMask                            src/Tracking.cc:2997-3110
//   Tracking::DynObjTracking                        src/Tracking.cc:1366-1612  (ground-truth bookkeeping :1531-1544 excluded)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/vdo_b200.h"

namespace {
#define TRK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "[vdo_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return VDO_ERR_CUDA; } } while (0)

// label of the current mask at the (truncated) flow target of every last-frame object point; -1 = outside (u>0, v>0 strict)
__global__ void k_gather_mask(const int* __restrict__ mask, int w, int h, const float* __restrict__ cx, const float* __restrict__ cy, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u = (int)cx[i], v = (int)cy[i];
  out[i] = (u < w && u > 0 && v < h && v > 0) ? mask[(size_t)v * w + u] : -1;
}
// forward-warp one object of the last mask into the current mask by the (int-truncated) last flow
__global__ void k_mask_warp(const int* __restrict__ mask_last, const float* __restrict__ flow_last, int w, int h, int label, int* __restrict__ mask_cur) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (k >= w) return;
  const size_t p = (size_t)j * w + k;
  if (mask_last[p] != label) return;
  const int fx = (int)flow_last[2 * p], fy = (int)flow_last[2 * p + 1];
  const int x = k + fx, y = j + fy;
  if (x < w && x > 0 && y < h && y > 0) mask_cur[(size_t)y * w + x] = label;   // every colliding write stores the same value
}

struct ObjStat { float boundary, sf_count, depth_sum; int n; };
// one thread per object walks its points in index order: float sums round exactly like the reference's loops
__global__ void k_obj_stats(const int* __restrict__ obj_begin, const int* __restrict__ obj_idx, int n_obj, const float* __restrict__ kx,
                            const float* __restrict__ ky, const float* __restrict__ depth, const float* __restrict__ flow3d, int rows, int cols,
                            int shr_row, int shr_col, float sf_thres, ObjStat* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_obj) return;
  ObjStat s{0.f, 0.f, 0.f, obj_begin[o + 1] - obj_begin[o]};
  for (int q = obj_begin[o]; q < obj_begin[o + 1]; ++q) {
    const int i = obj_idx[q];
    const float u = kx[i], v = ky[i];
    if (v < (float)shr_row || v > (float)(rows - shr_row) || u < (float)shr_col || u > (float)(cols - shr_col)) s.boundary = __fadd_rn(s.boundary, 1.f);
    s.depth_sum = __fadd_rn(s.depth_sum, depth[i]);
    const float fx = flow3d[3 * i], fz = flow3d[3 * i + 2];
    const float nrm = sqrtf(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fz, fz)));
    if (nrm < sf_thres) s.sf_count = __fadd_rn(s.sf_count, 1.f);
  }
  out[o] = s;
}

// majority label with the reference's ordering: std::map (ascending label) then sort by count descending; ties keep the
// smaller label first (what std::sort's insertion-sort path does for the <= 16 distinct labels seen in practice)
int majority_label(const std::vector<int>& v) {
  std::map<int, int> dups;
  for (int k : v) ++dups[k];
  int best = 0, cnt = -1;
  for (auto& kv : dups) if (kv.second > cnt) { cnt = kv.second; best = kv.first; }
  return best;
}
}  // namespace

// frame internals needed here (defined in frame_kernels.cu)
extern "C" int vdo_frame_device_ptrs(vdo_frame* f, unsigned char** gray, float** depth, float** flow, int** mask, int* w, int* h, void** stream);

// ------------------------------------------------------------------------------------------------ A16
extern "C" int vdo_tracklets_build(int n_rows, const int* row_begin, const int* assoc, const int* labels, int max_tracklets, int max_entries,
                                   int* n_trk, int* trk_begin, int* trk_frame, int* trk_feat, int* obj_id) {
  if (n_rows < 0 || !row_begin || !n_trk || !trk_begin) return VDO_ERR_ARG;
  std::vector<std::vector<std::pair<int, int>>> T;
  std::vector<int> oid, pre, cur;
  for (int i = 0; i < n_rows; ++i) {
    const int n = row_begin[i + 1] - row_begin[i];
    cur.assign(n, -1);
    for (int j = 0; j < n; ++j) {
      const int a = assoc[row_begin[i] + j];
      if (a == -1) continue;
      if (i > 0 && a >= 0 && a < (int)pre.size() && pre[a] != -1) { T[pre[a]].push_back({i + 1, j}); cur[j] = pre[a]; }
      else {
        if (i > 0 && (a < 0 || a >= (int)pre.size())) return VDO_ERR_ARG;
        T.push_back({{i, a}, {i + 1, j}});
        if (labels) oid.push_back(labels[row_begin[i] + j]);
        cur[j] = (int)T.size() - 1;
      }
    }
    pre.swap(cur);
  }
  size_t tot = 0;
  for (auto& t : T) tot += t.size();
  *n_trk = (int)T.size();
  if ((int)T.size() > max_tracklets || (long)tot > (long)max_entries) return VDO_ERR_ARG;
  int off = 0;
  for (size_t t = 0; t < T.size(); ++t) {
    trk_begin[t] = off;
    for (auto& pr : T[t]) { trk_frame[off] = pr.first; trk_feat[off] = pr.second; ++off; }
    if (labels && obj_id) obj_id[t] = oid[t];
  }
  trk_begin[T.size()] = off;
  return VDO_OK;
}

// ------------------------------------------------------------------------------------------------ A15
// cur / last: resident frames (masks + last flow).  sem_label_last, corres_x/y: last frame's vSemObjLabel and mvObjCorres (n points).
// mask_out (may be NULL) receives the updated current mask so the caller's cv::Mat is mutated like in the reference.
extern "C" int vdo_update_mask(vdo_frame* cur, vdo_frame* last, int n, const int* sem_label_last, const float* corres_x, const float* corres_y,
                               int* mask_out, int* n_warped, int* warped_labels) {
  if (!cur || !last || n < 0) return VDO_ERR_ARG;
  int *mcur, *mlast, w, h, w2, h2; float *fl_last; void* stv;
  if (vdo_frame_device_ptrs(cur, nullptr, nullptr, nullptr, &mcur, &w, &h, &stv)) return VDO_ERR_ARG;
  if (vdo_frame_device_ptrs(last, nullptr, nullptr, &fl_last, &mlast, &w2, &h2, nullptr)) return VDO_ERR_ARG;
  if (w != w2 || h != h2) return VDO_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stv;
  if (n_warped) *n_warped = 0;
  std::vector<int> uni(sem_label_last, sem_label_last + n);
  std::sort(uni.begin(), uni.end());
  uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
  float *d_cx = nullptr, *d_cy = nullptr; int* d_lab = nullptr;
  if (n > 0) {
    TRK(cudaMalloc(&d_cx, sizeof(float) * n)); TRK(cudaMalloc(&d_cy, sizeof(float) * n)); TRK(cudaMalloc(&d_lab, sizeof(int) * n));
    TRK(cudaMemcpyAsync(d_cx, corres_x, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    TRK(cudaMemcpyAsync(d_cy, corres_y, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  }
  std::vector<int> lab(n);
  bool stale = true;
  for (size_t oi = 0; oi < uni.size(); ++oi) {
    if (stale && n > 0) {          // (re)gather: an earlier object's warp may have changed the labels this object votes on
      k_gather_mask<<<(n + 255) / 256, 256, 0, st>>>(mcur, w, h, d_cx, d_cy, n, d_lab);
      TRK(cudaMemcpyAsync(lab.data(), d_lab, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
      TRK(cudaStreamSynchronize(st));
      stale = false;
    }
    std::vector<int> tmp;
    for (int i = 0; i < n; ++i) if (sem_label_last[i] == uni[oi] && lab[i] != -1) tmp.push_back(lab[i]);
    if (tmp.size() < 100) continue;
    if (majority_label(tmp) != 0) continue;
    dim3 b(256), g((w + 255) / 256, h);
    k_mask_warp<<<g, b, 0, st>>>(mlast, fl_last, w, h, uni[oi], mcur);
    if (warped_labels && n_warped) warped_labels[*n_warped] = uni[oi];
    if (n_warped) ++*n_warped;
    stale = true;
  }
  if (mask_out) TRK(cudaMemcpyAsync(mask_out, mcur, sizeof(int) * (size_t)w * h, cudaMemcpyDeviceToHost, st));
  TRK(cudaStreamSynchronize(st));
  cudaFree(d_cx); cudaFree(d_cy); cudaFree(d_lab);
  return VDO_OK;
}

// ------------------------------------------------------------------------------------------------ A13
// Inputs are the current frame's object points (n): semantic label, tracking label (in/out), pixel, depth, scene flow (n x 3),
// the same points' semantic labels in the last frame, and the last frame's object table (nSemPosition, bObjStat, nModLabel).
// Outputs: obj_label (in place), the kept objects as CSR (obj_begin / obj_idx, indices into the point arrays) with their
// tracking ids (mod_label) and semantic labels (sem_position), and max_id (in/out).
extern "C" int vdo_dyn_obj_tracking(vdo_ctx* ctx, int n, const int* sem_label, int* obj_label, const float* kx, const float* ky, const float* depth,
                                    const float* flow3d, const int* sem_label_last, int n_last_obj, const int* last_sem_position,
                                    const unsigned char* last_obj_stat, const int* last_mod_label, int rows, int cols, int shrink_row, int shrink_col,
                                    float sf_mg_thres, float sf_ds_thres, float th_depth_obj, int f_id, int* max_id, int max_objects,
                                    int* n_objects, int* obj_begin, int* obj_idx, int* mod_label, int* sem_position) {
  if (!ctx || n < 0 || !max_id || !n_objects) return VDO_ERR_ARG;
  cudaStream_t st = (cudaStream_t)(uintptr_t)vdo_ctx_stream(ctx);
  *n_objects = 0;
  std::vector<int> uni(sem_label, sem_label + n);
  std::sort(uni.begin(), uni.end());
  uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
  std::vector<std::vector<int>> posi(uni.size());
  for (int i = 0; i < n; ++i) {
    if (obj_label[i] == -1) continue;
    const int j = (int)(std::lower_bound(uni.begin(), uni.end(), sem_label[i]) - uni.begin());
    posi[j].push_back(i);
  }
  // ---- per-object statistics on the device (one sequential thread per object, reference rounding) ----
  std::vector<int> ob(1, 0), oidx;
  for (auto& p : posi) { oidx.insert(oidx.end(), p.begin(), p.end()); ob.push_back((int)oidx.size()); }
  const int no = (int)posi.size();
  std::vector<ObjStat> stats(no);
  if (no > 0 && !oidx.empty()) {
    int *d_ob, *d_oi; float *d_kx, *d_ky, *d_dp, *d_f3; ObjStat* d_st;
    TRK(cudaMalloc(&d_ob, sizeof(int) * (no + 1))); TRK(cudaMalloc(&d_oi, sizeof(int) * oidx.size()));
    TRK(cudaMalloc(&d_kx, sizeof(float) * n)); TRK(cudaMalloc(&d_ky, sizeof(float) * n)); TRK(cudaMalloc(&d_dp, sizeof(float) * n)); TRK(cudaMalloc(&d_f3, sizeof(float) * 3 * n));
    TRK(cudaMalloc(&d_st, sizeof(ObjStat) * no));
    TRK(cudaMemcpyAsync(d_ob, ob.data(), sizeof(int) * (no + 1), cudaMemcpyHostToDevice, st));
    TRK(cudaMemcpyAsync(d_oi, oidx.data(), sizeof(int) * oidx.size(), cudaMemcpyHostToDevice, st));
    TRK(cudaMemcpyAsync(d_kx, kx, sizeof(float) * n, cudaMemcpyHostToDevice, st)); TRK(cudaMemcpyAsync(d_ky, ky, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    TRK(cudaMemcpyAsync(d_dp, depth, sizeof(float) * n, cudaMemcpyHostToDevice, st)); TRK(cudaMemcpyAsync(d_f3, flow3d, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st));
    k_obj_stats<<<(no + 31) / 32, 32, 0, st>>>(d_ob, d_oi, no, d_kx, d_ky, d_dp, d_f3, rows, cols, shrink_row, shrink_col, sf_mg_thres, d_st);
    TRK(cudaMemcpyAsync(stats.data(), d_st, sizeof(ObjStat) * no, cudaMemcpyDeviceToHost, st));
    TRK(cudaStreamSynchronize(st));
    cudaFree(d_ob); cudaFree(d_oi); cudaFree(d_kx); cudaFree(d_ky); cudaFree(d_dp); cudaFree(d_f3); cudaFree(d_st);
  }
  // ---- decisions, in label order like the reference ----
  std::vector<std::vector<int>> obj_new; std::vector<int> sem_new;
  for (int i = 0; i < no; ++i) {
    const float sz = (float)posi[i].size();
    if (posi[i].empty()) continue;                            // (the reference would divide 0/0 here: NaN > 0.5 is false -> kept, then dropped for size < 150)
    if (stats[i].boundary / sz > 0.5f) { for (int k : posi[i]) obj_label[k] = -1; continue; }
    if (stats[i].sf_count / sz > sf_ds_thres) { for (int k : posi[i]) obj_label[k] = 0; continue; }
    if (stats[i].depth_sum / sz > th_depth_obj || posi[i].size() < 150) { for (int k : posi[i]) obj_label[k] = -1; continue; }
    obj_new.push_back(posi[i]); sem_new.push_back(uni[i]);
  }
  if (f_id == 1) *max_id = 1;
  if ((int)obj_new.size() > max_objects) return VDO_ERR_ARG;
  int off = 0;
  for (size_t i = 0; i < obj_new.size(); ++i) {
    std::vector<int> lb;
    for (int k : obj_new[i]) lb.push_back(sem_label_last[k]);
    const int new_lab = majority_label(lb);
    int id = -1;
    if (*max_id == 1) { id = *max_id; *max_id += 1; }
    else {
      for (int k = 0; k < n_last_obj; ++k)
        if (last_sem_position[k] == new_lab && last_obj_stat[k]) { id = last_mod_label[k]; break; }
      if (id == -1) { id = *max_id; *max_id += 1; }
    }
    for (int k : obj_new[i]) obj_label[k] = id;
    mod_label[i] = id; sem_position[i] = sem_new[i];
    obj_begin[i] = off;
    for (int k : obj_new[i]) obj_idx[off++] = k;
  }
  obj_begin[obj_new.size()] = off;
  *n_objects = (int)obj_new.size();
  return VDO_OK;
}
