// converter.cpp -- the float <-> double boundary of the reference (src/Converter.cc) as host functions of the C ABI:
//   toSE3Quat   (:25-35)   4x4 CV_32F -> g2o::SE3Quat(R, t): Eigen::Quaterniond(R) (trace / largest-diagonal branches), flipped to w >= 0,
//                          normalised (se3quat.h:58-60, 286-301)
//   toCvMat     (:37-41, SE3Quat::to_homogeneous_matrix)  quaternion -> rotation matrix, rounded to float
//   toInvMatrix (:151-166) [R^T | -R^T t]: `-R.t()*t` is a cv::gemm with GEMM_1_T, alpha = -1 (generic branch: double accumulation, one rounding)
//   cv::Mat A*B of two 4x4 CV_32F: gemm's small-matrix branch (float accumulation, left to right)
// The device code of the flow LM (flow_lm.cu: rot_to_quat / quat_normalize_pos / quat_to_rot) and the tracker (tracker.cpp: to_iso / from_iso /
// mul4 / inv4) restate the same formulas; these entry points exist so the rounding rules can be tested in isolation (tests/test_converter.py)
// and so that reference-side code keeps a Converter to call.
#include <cmath>
#include <cstring>

#include "../../include/vdo_b200.h"
#include "ba_math.cuh"

extern "C" int vdo_convert_to_se3quat(const float* T16, double* q4, double* t3) {
  if (!T16 || !q4 || !t3) return VDO_ERR_ARG;
  double R[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = (double)T16[4 * i + j];
  vdo::quat_from_rot(R, q4);
  if (q4[3] < 0) { q4[0] = -q4[0]; q4[1] = -q4[1]; q4[2] = -q4[2]; q4[3] = -q4[3]; }
  const double n = std::sqrt(q4[0] * q4[0] + q4[1] * q4[1] + q4[2] * q4[2] + q4[3] * q4[3]);
  for (int k = 0; k < 4; ++k) q4[k] /= n;
  for (int k = 0; k < 3; ++k) t3[k] = (double)T16[4 * k + 3];
  return VDO_OK;
}
extern "C" int vdo_convert_to_cvmat(const double* q4, const double* t3, float* T16) {
  if (!q4 || !t3 || !T16) return VDO_ERR_ARG;
  double R[9];
  vdo::rot_from_quat(q4, R);
  std::memset(T16, 0, 64);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T16[4 * i + j] = (float)R[3 * i + j]; T16[4 * i + 3] = (float)t3[i]; }
  T16[15] = 1.f;
  return VDO_OK;
}
extern "C" int vdo_convert_inv_matrix(const float* T16, float* out16) {
  if (!T16 || !out16) return VDO_ERR_ARG;
  float o[16];
  std::memset(o, 0, sizeof o);
  o[15] = 1.f;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) o[4 * i + j] = T16[4 * j + i];
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)T16[4 * k + i] * (double)T16[4 * k + 3];
    o[4 * i + 3] = (float)(-s);
  }
  std::memcpy(out16, o, sizeof o);
  return VDO_OK;
}
extern "C" int vdo_convert_mul4(const float* A16, const float* B16, float* out16) {
  if (!A16 || !B16 || !out16) return VDO_ERR_ARG;
  float o[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = A16[4 * i] * B16[j];
      s = s + A16[4 * i + 1] * B16[4 + j];
      s = s + A16[4 * i + 2] * B16[8 + j];
      s = s + A16[4 * i + 3] * B16[12 + j];
      o[4 * i + j] = s;
    }
  std::memcpy(out16, o, sizeof o);
  return VDO_OK;
}
