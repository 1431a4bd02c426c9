// Optimizer.cc -- see Optimizer.h.
#include "Optimizer.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

#include "vdo_b200.h"

namespace VDO_SLAM {

float Frame::fx = 0, Frame::fy = 0, Frame::cx = 0, Frame::cy = 0, Frame::invfx = 0, Frame::invfy = 0;
bool Optimizer::msQuirk = true;

namespace {
vdo_ctx *g_ctx = nullptr;
void mat16(const cv::Mat &m, float *o) {
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o[4 * i + j] = (i < m.rows && j < m.cols) ? m.at<float>(i, j) : (i == j ? 1.f : 0.f);
}
cv::Mat from16(const float *v) {
  cv::Mat m = cv::Mat::eye(4, 4, CV_32F);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m.at<float>(i, j) = v[4 * i + j];
  return m;
}
}  // namespace

void Optimizer::SetContext(vdo_ctx *ctx) { g_ctx = ctx; }
vdo_ctx *Optimizer::Context() {
  if (!g_ctx && vdo_ctx_create(0, &g_ctx) != VDO_OK) {
    cerr << "vdo_b200: no usable CUDA device (there is no CPU fallback)" << endl;
    exit(-1);
  }
  return g_ctx;
}

int Optimizer::PoseOptimizationFlow2Cam(Frame *pCurFrame, Frame *pLastFrame, vector<int> &TemperalMatch) {
  const int N = (int)TemperalMatch.size();
  if (N < 3) return 0;                                                     // src/Optimizer.cc:2449-2450
  vector<float> pts(2 * (size_t)N), dep(N), flo(2 * (size_t)N);
  for (int i = 0; i < N; ++i) {
    const int k = TemperalMatch[i];
    pts[2 * i] = pLastFrame->mvStatKeys[k].pt.x; pts[2 * i + 1] = pLastFrame->mvStatKeys[k].pt.y;
    dep[i] = pLastFrame->mvStatDepth[k];                                   // Frame::ObtainFlowDepthCamera (src/Frame.cc:644-670)
    flo[2 * i] = pLastFrame->mvFlowNext[k].x; flo[2 * i + 1] = pLastFrame->mvFlowNext[k].y;
  }
  const float K[4] = {pCurFrame->fx, pCurFrame->fy, pCurFrame->cx, pCurFrame->cy};
  float Tl[16], Ti[16], To[16];
  mat16(pLastFrame->mTcw, Tl); mat16(pCurFrame->mTcw, Ti);
  vector<double> flow_out(2 * (size_t)N);
  vector<unsigned char> inl(N);
  double stats[8];
  if (vdo_pose_opt_flow2(Context(), 0, msQuirk ? 1 : 0, N, pts.data(), dep.data(), flo.data(), K, Tl, Ti, To, flow_out.data(), inl.data(), stats) != VDO_OK) {
    cerr << "vdo_b200: PoseOptimizationFlow2Cam failed: " << vdo_last_error(Context()) << endl;
    exit(-1);
  }
  pCurFrame->SetPose(from16(To));
  int nBad = 0;
  for (int i = 0; i < N; ++i) {                                            // :2512-2535
    const int k = TemperalMatch[i];
    if (inl[i]) {
      pCurFrame->mvStatKeys[k].pt.x = (float)(pLastFrame->mvStatKeys[k].pt.x + flow_out[2 * i]);
      pCurFrame->mvStatKeys[k].pt.y = (float)(pLastFrame->mvStatKeys[k].pt.y + flow_out[2 * i + 1]);
    } else { TemperalMatch[i] = -1; ++nBad; }
  }
  return N - nBad;
}

cv::Mat Optimizer::PoseOptimizationFlow2(Frame *pCurFrame, Frame *pLastFrame, const vector<int> &ObjId, std::vector<int> &InlierID) {
  const int N = (int)ObjId.size();
  InlierID.clear();
  if (N < 3) return cv::Mat::eye(4, 4, CV_32F);                            // :2872-2873
  vector<float> pts(2 * (size_t)N), dep(N), flo(2 * (size_t)N);
  for (int i = 0; i < N; ++i) {
    const int k = ObjId[i];
    pts[2 * i] = pLastFrame->mvObjKeys[k].pt.x; pts[2 * i + 1] = pLastFrame->mvObjKeys[k].pt.y;
    dep[i] = pLastFrame->mvObjDepth[k];                                    // Frame::ObtainFlowDepthObject (src/Frame.cc:617-642)
    flo[2 * i] = pLastFrame->mvObjFlowNext[k].x; flo[2 * i + 1] = pLastFrame->mvObjFlowNext[k].y;
  }
  const float K[4] = {pCurFrame->fx, pCurFrame->fy, pCurFrame->cx, pCurFrame->cy};
  float Tl[16], Ti[16], To[16];
  mat16(pLastFrame->mTcw, Tl); mat16(pCurFrame->mInitModel, Ti);
  vector<double> flow_out(2 * (size_t)N);
  vector<unsigned char> inl(N);
  double stats[8];
  if (vdo_pose_opt_flow2(Context(), 1, msQuirk ? 1 : 0, N, pts.data(), dep.data(), flo.data(), K, Tl, Ti, To, flow_out.data(), inl.data(), stats) != VDO_OK) {
    cerr << "vdo_b200: PoseOptimizationFlow2 failed: " << vdo_last_error(Context()) << endl;
    exit(-1);
  }
  for (int i = 0; i < N; ++i) {                                            // :2940-2965
    const int k = ObjId[i];
    if (inl[i]) {
      pCurFrame->mvObjKeys[k].pt.x = (float)(pLastFrame->mvObjKeys[k].pt.x + flow_out[2 * i]);
      pCurFrame->mvObjKeys[k].pt.y = (float)(pLastFrame->mvObjKeys[k].pt.y + flow_out[2 * i + 1]);
      InlierID.push_back(k);
    } else pCurFrame->vObjLabel[k] = -1;
  }
  return from16(To);
}

cv::Mat Optimizer::Get3DinCamera(const cv::KeyPoint &Feats2d, const float &Dpts, const cv::Mat &Calib_K) {   // :2995-3013
  const float invfx = 1.0f / Calib_K.at<float>(0, 0), invfy = 1.0f / Calib_K.at<float>(1, 1), cx = Calib_K.at<float>(0, 2), cy = Calib_K.at<float>(1, 2);
  cv::Mat x3D(3, 1, CV_32F);
  const float z = Dpts;
  x3D.at<float>(0, 0) = (Feats2d.pt.x - cx) * z * invfx;
  x3D.at<float>(1, 0) = (Feats2d.pt.y - cy) * z * invfy;
  x3D.at<float>(2, 0) = z;
  return x3D;
}

cv::Mat Optimizer::Get3DinWorld(const cv::KeyPoint &Feats2d, const float &Dpts, const cv::Mat &Calib_K, const cv::Mat &CameraPose) {   // :2974-2993
  // x3D = Rwc * x3Dc + twc: cv::Mat gemm of a 3x3 by a 3x1 (generic branch: double accumulation, one rounding), then a float add
  const cv::Mat c = Get3DinCamera(Feats2d, Dpts, Calib_K);
  cv::Mat x3D(3, 1, CV_32F);
  for (int r = 0; r < 3; ++r) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += (double)CameraPose.at<float>(r, k) * (double)c.at<float>(k, 0);
    x3D.at<float>(r, 0) = (float)s + CameraPose.at<float>(r, 3);
  }
  return x3D;
}

namespace {
// Map -> map-only tracker handle -> vdo_tracker_batch_optimize -> Map
void batch(Map *pMap, const cv::Mat &K, int mode, int window) {
  vdo_tracker_params p;
  vdo_tracker_params_default(&p);
  p.width = p.height = 0;                                                  // map-only handle
  p.fx = K.at<float>(0, 0); p.fy = K.at<float>(1, 1); p.cx = K.at<float>(0, 2); p.cy = K.at<float>(1, 2);
  if (window > 0) p.window_size = window;
  vdo_tracker *t = nullptr;
  if (vdo_tracker_create(Optimizer::Context(), &p, &t) != VDO_OK) { cerr << "vdo_b200: map handle creation failed" << endl; exit(-1); }
  const int N = (int)pMap->vpFeatSta.size();
  for (int i = 0; i < N; ++i) {
    const int ns = (int)pMap->vpFeatSta[i].size(), nd = (int)pMap->vpFeatDyn[i].size();
    vector<float> fs(2 * (size_t)ns + 2), ps(3 * (size_t)ns + 3), fd(2 * (size_t)nd + 2), pd(3 * (size_t)nd + 3);
    for (int j = 0; j < ns; ++j) {
      fs[2 * j] = pMap->vpFeatSta[i][j].pt.x; fs[2 * j + 1] = pMap->vpFeatSta[i][j].pt.y;
      for (int k = 0; k < 3; ++k) ps[3 * j + k] = pMap->vp3DPointSta[i][j].at<float>(k, 0);
    }
    for (int j = 0; j < nd; ++j) {
      fd[2 * j] = pMap->vpFeatDyn[i][j].pt.x; fd[2 * j + 1] = pMap->vpFeatDyn[i][j].pt.y;
      for (int k = 0; k < 3; ++k) pd[3 * j + k] = pMap->vp3DPointDyn[i][j].at<float>(k, 0);
    }
    float cam[16]; mat16(pMap->vmCameraPose[i], cam);
    vector<float> mot; vector<int> lab;
    if (i > 0) {
      lab = pMap->vnRMLabel[i - 1];
      mot.resize(16 * lab.size());
      for (size_t j = 0; j < lab.size(); ++j) mat16(pMap->vmRigidMotion[i - 1][j], &mot[16 * j]);
    }
    const int rc = vdo_tracker_map_push(t, ns, fs.data(), pMap->vfDepSta[i].data(), ps.data(), i > 0 ? pMap->vnAssoSta[i - 1].data() : nullptr, nd, fd.data(),
                                        pMap->vfDepDyn[i].data(), pd.data(), i > 0 ? pMap->vnAssoDyn[i - 1].data() : nullptr, i > 0 ? pMap->vnFeatLabel[i - 1].data() : nullptr,
                                        cam, (int)lab.size(), mot.data(), lab.data());
    if (rc != VDO_OK) { cerr << "vdo_b200: vdo_tracker_map_push failed (" << rc << "): " << vdo_tracker_last_error(t) << endl; exit(-1); }
  }
  vdo_lm_stats st;
  if (vdo_tracker_batch_optimize(t, mode, nullptr, &st, nullptr) != VDO_OK) {
    cerr << "vdo_b200: batch optimisation failed: " << vdo_tracker_last_error(t) << endl;
    vdo_tracker_destroy(t);
    return;                                                                // like the reference, the Map is left as it was
  }
  auto get = [&](const char *name) { int n = 0; vdo_tracker_map_get(t, name, nullptr, 0, &n); vector<float> v(n > 0 ? n : 1); vdo_tracker_map_get(t, name, v.data(), n, &n); v.resize(n); return v; };
  const vector<float> cam = get(mode == 1 ? "vmCameraPose_RF" : "vmCameraPose"), mot = get(mode == 1 ? "vmRigidMotion_RF" : "vmRigidMotion");
  const vector<float> p3s = get("vp3DPointSta"), p3d = get("vp3DPointDyn");
  vector<cv::Mat> &camOut = mode == 1 ? pMap->vmCameraPose_RF : pMap->vmCameraPose;
  vector<vector<cv::Mat> > &motOut = mode == 1 ? pMap->vmRigidMotion_RF : pMap->vmRigidMotion;
  size_t qm = 0, qs = 0, qd = 0;
  for (int i = 0; i < N; ++i) {
    if (i < (int)camOut.size()) camOut[i] = from16(&cam[16 * (size_t)i]);
    for (size_t j = 0; j < pMap->vp3DPointSta[i].size(); ++j, qs += 3) for (int k = 0; k < 3; ++k) pMap->vp3DPointSta[i][j].at<float>(k, 0) = p3s[qs + k];
    for (size_t j = 0; j < pMap->vp3DPointDyn[i].size(); ++j, qd += 3) for (int k = 0; k < 3; ++k) pMap->vp3DPointDyn[i][j].at<float>(k, 0) = p3d[qd + k];
    if (i > 0 && i - 1 < (int)motOut.size())
      for (size_t j = 0; j < motOut[i - 1].size(); ++j, qm += 16) motOut[i - 1][j] = from16(&mot[qm]);
  }
  vdo_tracker_destroy(t);
}
}  // namespace

void Optimizer::FullBatchOptimization(Map *pMap, const cv::Mat Calib_K) { batch(pMap, Calib_K, 1, 0); }
void Optimizer::PartialBatchOptimization(Map *pMap, const cv::Mat Calib_K, const int WINDOW_SIZE) { batch(pMap, Calib_K, 0, WINDOW_SIZE); }

}  // namespace VDO_SLAM
