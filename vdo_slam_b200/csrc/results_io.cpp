// results_io.cpp -- the step AFTER the path: result files and error metrics of the reference.
//
//   System::SaveResults (src/System.cc:66-244) writes, with `fixed << setprecision(9)`:
//     <prefix>initial_stereo_new.txt / refined_stereo_new.txt / cam_pose_gt_stereo.txt   "frame  r00 r01 r02 tx ... r22 tz  0 0 0 1"
//     <prefix>obj_mot_stereo_new.txt / obj_mot_stereo_rf_new.txt                         "frame+1 label  <body-frame motion, 12 values>  0 0 0 1"
//         body-frame motion = toInvMatrix(ObjPosePre) * RigidMotion * ObjPosePre          (:92-93)
//     <prefix>obj_mot_gt.txt                                                              "frame+1 label  <ground-truth motion as stored>  0 0 0 1"
//     <prefix>obj_centre.txt                                                              "frame+1 label  x y z"
//     (entry j = 0 of every frame is the camera and is skipped, :88)
//   Tracking::GetMetricError (src/Tracking.cc:3243-3386): mean relative camera-pose error and mean object-motion error
//     (translation norm, rotation angle in degrees from the clipped trace), overall and per object id.
// All arithmetic is the reference's float arithmetic, including the two branches of OpenCV's gemm that cv::Mat expressions reach
// here (4x4 * 4x4: float accumulation; transposed 3x3 * 3x1: double accumulation) -- pinned against cv2.gemm in
// tests/test_results_io.py.  Host-only.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/vdo_b200.h"

namespace {
struct M4 { float m[16]; };
M4 mul(const M4& A, const M4& B) {            // cv::Mat A*B of two 4x4 CV_32F: OpenCV's small-matrix branch of gemm (flags == 0, inner
  M4 C;                                       // dimension <= 4) evaluates a0*b0 + a1*b1 + a2*b2 + a3*b3 in FLOAT, left to right
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = A.m[4 * r] * B.m[c];
      s = s + A.m[4 * r + 1] * B.m[4 + c];
      s = s + A.m[4 * r + 2] * B.m[8 + c];
      s = s + A.m[4 * r + 3] * B.m[12 + c];
      C.m[4 * r + c] = s;
    }
  return C;
}
M4 inv(const M4& T) {                         // Converter::toInvMatrix (src/Converter.cc:151-166): [R^T | -R^T t]; `-R.t()*t` is a gemm
                                              // with GEMM_1_T and alpha = -1: the generic branch, double accumulation, one rounding
  M4 I;
  std::memset(I.m, 0, sizeof I.m);
  I.m[15] = 1.0f;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) I.m[4 * r + c] = T.m[4 * c + r];
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)T.m[4 * k + r] * (double)T.m[4 * k + 3];
    I.m[4 * r + 3] = (float)(-s);
  }
  return I;
}
M4 load(const float* p) { M4 A; std::memcpy(A.m, p, sizeof A.m); return A; }
void put12(FILE* f, const M4& T) {
  for (int i = 0; i < 12; ++i) std::fprintf(f, "%.9f ", (double)T.m[i]);
  std::fprintf(f, "%.9f %.9f %.9f %.9f\n", 0.0, 0.0, 0.0, 1.0);
}
// translation norm and rotation angle (degrees) of an error transform, as GetMetricError computes them
void err_tr(const M4& E, float& t, float& r) {
  t = std::sqrt(E.m[3] * E.m[3] + E.m[7] * E.m[7] + E.m[11] * E.m[11]);
  float trace = 0;
  for (int j = 0; j < 3; ++j) {
    const float d = E.m[5 * j];
    if (d > 1.0) trace = (float)(trace + 1.0 - (d - 1.0));
    else trace = trace + d;
  }
  r = (float)(std::acos((trace - 1.0) / 2.0) * 180.0 / 3.1415926);
}
}  // namespace

extern "C" {

int vdo_results_write_poses(const char* path, int start_frame, int n, const float* T16) {
  if (!path || n < 0 || (n && !T16)) return VDO_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return VDO_ERR_ARG;
  for (int i = 0; i < n; ++i) { std::fprintf(f, "%d ", start_frame + i); put12(f, load(T16 + 16 * (size_t)i)); }
  const bool ok = !std::ferror(f);
  std::fclose(f);
  return ok ? VDO_OK : VDO_ERR_ARG;
}

// n_per_frame[i] entries for frame i (entry 0 = camera, skipped); labels / H16 / pose_pre16 hold all entries of all frames
// back to back.  pose_pre16 == NULL: the matrices are written as stored (obj_mot_gt.txt).
int vdo_results_write_object_motions(const char* path, int start_frame, int n_frames, const int* n_per_frame, const int* labels, const float* H16,
                                     const float* pose_pre16) {
  if (!path || n_frames < 0 || (n_frames && (!n_per_frame || !labels || !H16))) return VDO_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return VDO_ERR_ARG;
  size_t q = 0;
  for (int i = 0; i < n_frames; ++i) {
    for (int j = 0; j < n_per_frame[i]; ++j, ++q) {
      if (j == 0 || n_per_frame[i] <= 1) continue;
      M4 H = load(H16 + 16 * q);
      if (pose_pre16) { const M4 L = load(pose_pre16 + 16 * q); H = mul(mul(inv(L), H), L); }
      std::fprintf(f, "%d %d ", start_frame + i + 1, labels[q]);
      put12(f, H);
    }
  }
  const bool ok = !std::ferror(f);
  std::fclose(f);
  return ok ? VDO_OK : VDO_ERR_ARG;
}

int vdo_results_write_object_centres(const char* path, int start_frame, int n_frames, const int* n_per_frame, const int* labels, const float* centre3) {
  if (!path || n_frames < 0 || (n_frames && (!n_per_frame || !labels || !centre3))) return VDO_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return VDO_ERR_ARG;
  size_t q = 0;
  for (int i = 0; i < n_frames; ++i)
    for (int j = 0; j < n_per_frame[i]; ++j, ++q) {
      if (j == 0 || n_per_frame[i] <= 1) continue;
      std::fprintf(f, "%d %d %.9f %.9f %.9f\n", start_frame + i + 1, labels[q], (double)centre3[3 * q], (double)centre3[3 * q + 1], (double)centre3[3 * q + 2]);
    }
  const bool ok = !std::ferror(f);
  std::fclose(f);
  return ok ? VDO_OK : VDO_ERR_ARG;
}

// out4: mean camera translation / rotation error, mean object translation / rotation error; each_obj_*: max_id - 1 entries
// (object id k at index k - 1; entries with count 0 are NaN like the reference's 0/0).  obj_stat: 1 = estimated, 0 = failure.
int vdo_metric_error(int n_cam, const float* cam16, const float* cam_gt16, int n_frames, const int* n_per_frame, const int* labels,
                     const unsigned char* obj_stat, const float* H16, const float* pose_pre16, const float* H_gt16, int max_id, float out4[4],
                     float* each_obj_t, float* each_obj_r, int* each_obj_count) {
  if (!out4 || n_cam < 0 || n_frames < 0 || (n_cam && (!cam16 || !cam_gt16))) return VDO_ERR_ARG;
  float t_sum = 0, r_sum = 0;
  for (int i = 1; i < n_cam; ++i) {
    const M4 T_lc_inv = mul(load(cam16 + 16 * (size_t)i), inv(load(cam16 + 16 * (size_t)(i - 1))));
    const M4 T_lc_gt = mul(load(cam_gt16 + 16 * (size_t)(i - 1)), inv(load(cam_gt16 + 16 * (size_t)i)));
    float t, r;
    err_tr(mul(T_lc_inv, T_lc_gt), t, r);
    t_sum = t_sum + t; r_sum = r_sum + r;
  }
  out4[0] = t_sum / (n_cam - 1); out4[1] = r_sum / (n_cam - 1);
  const int nobj = max_id > 1 ? max_id - 1 : 0;
  std::vector<float> et(nobj, 0.f), er(nobj, 0.f);
  std::vector<int> ec(nobj, 0);
  float t_rpe = 0, r_rpe = 0, count = 0;
  size_t q = 0;
  for (int i = 0; i < n_frames; ++i)
    for (int j = 0; j < n_per_frame[i]; ++j, ++q) {
      if (j == 0 || n_per_frame[i] <= 1) continue;
      if (obj_stat && !obj_stat[q]) continue;
      const int id = labels[q];
      if (id < 1 || id > nobj) return VDO_ERR_ARG;
      const M4 L = load(pose_pre16 + 16 * q);
      const M4 body = mul(mul(inv(L), load(H16 + 16 * q)), L);
      float t, r;
      err_tr(mul(inv(body), load(H_gt16 + 16 * q)), t, r);
      et[id - 1] += t; er[id - 1] += r; ec[id - 1] += 1;
      t_rpe = t_rpe + t; r_rpe = r_rpe + r; count = count + 1;
    }
  out4[2] = t_rpe / count; out4[3] = r_rpe / count;
  for (int k = 0; k < nobj; ++k) {
    if (each_obj_t) each_obj_t[k] = et[k] / ec[k];
    if (each_obj_r) each_obj_r[k] = er[k] / ec[k];
    if (each_obj_count) each_obj_count[k] = ec[k];
  }
  return VDO_OK;
}

}  // extern "C"
