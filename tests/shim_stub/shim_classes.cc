// Test driver for the class-level shims (VDO_SLAM::ORBextractor, VDO_SLAM::Optimizer over the C ABI), fed by tests/test_host_shim.py.
//   shim_classes orb  <gray.bin> <w> <h> <out.bin>            ORBextractor(2500, 1.2, 8, 20, 7)(image, Mat(), keypoints, descriptors)
//   shim_classes flow <problem.bin> <out.bin>                 Optimizer::PoseOptimizationFlow2 / Flow2Cam on a dumped problem
//   shim_classes ba   <map.bin> <mode> <window> <out.bin>     Optimizer::FullBatchOptimization / PartialBatchOptimization on a dumped Map
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "ORBextractor.h"
#include "Optimizer.h"

using namespace VDO_SLAM;

struct Reader {
  std::ifstream f;
  explicit Reader(const char* p) : f(p, std::ios::binary) {}
  int i32() { int v = 0; f.read((char*)&v, 4); return v; }
  float f32() { float v = 0; f.read((char*)&v, 4); return v; }
  void floats(float* p, size_t n) { f.read((char*)p, (std::streamsize)(4 * n)); }
  void ints(int* p, size_t n) { f.read((char*)p, (std::streamsize)(4 * n)); }
  cv::Mat mat4() { cv::Mat m(4, 4, CV_32F); floats((float*)m.data, 16); return m; }
};
static void put(std::ofstream& o, const void* p, size_t n) { o.write((const char*)p, (std::streamsize)n); }

static int run_orb(int argc, char** argv) {
  if (argc < 6) return 2;
  const int w = std::atoi(argv[3]), h = std::atoi(argv[4]);
  cv::Mat im(h, w, CV_8UC1);
  std::ifstream(argv[2], std::ios::binary).read((char*)im.data, (std::streamsize)((size_t)w * h));
  ORBextractor ext(2500, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> kps; cv::Mat desc;
  ext(im, cv::Mat(), kps, desc);
  std::ofstream o(argv[5], std::ios::binary);
  const int n = (int)kps.size(); put(o, &n, 4);
  for (const cv::KeyPoint& k : kps) { const float v[5] = {k.pt.x, k.pt.y, k.size, k.angle, k.response}; put(o, v, 20); put(o, &k.octave, 4); }
  if (n) put(o, desc.data, (size_t)n * 32);
  const int nl = ext.GetLevels(); put(o, &nl, 4);
  for (int l = 0; l < nl; ++l) { const int d[2] = {ext.mvImagePyramid[l].cols, ext.mvImagePyramid[l].rows}; put(o, d, 8); }
  return 0;
}

static int run_flow(int argc, char** argv) {
  if (argc < 4) return 2;
  Reader r(argv[2]);
  const int mode = r.i32(), n = r.i32();
  Frame cur, last;
  Frame::fx = r.f32(); Frame::fy = r.f32(); Frame::cx = r.f32(); Frame::cy = r.f32();
  last.mTcw = r.mat4();
  const cv::Mat Tinit = r.mat4();
  std::vector<float> pts(2 * (size_t)n), dep(n), flo(2 * (size_t)n);
  r.floats(pts.data(), pts.size()); r.floats(dep.data(), n); r.floats(flo.data(), flo.size());
  std::vector<int> ids(n);
  cv::Mat T; int n_in = 0;
  std::vector<int> inl;
  if (mode == 0) {
    last.mvStatKeys.resize(n); last.mvStatDepth = dep; last.mvFlowNext.resize(n); cur.mvStatKeys.resize(n);
    for (int i = 0; i < n; ++i) { last.mvStatKeys[i].pt = cv::Point2f(pts[2 * i], pts[2 * i + 1]); last.mvFlowNext[i] = cv::Point2f(flo[2 * i], flo[2 * i + 1]); ids[i] = i; }
    cur.mTcw = Tinit;
    n_in = Optimizer::PoseOptimizationFlow2Cam(&cur, &last, ids);
    T = cur.mTcw;
    for (int i = 0; i < n; ++i) if (ids[i] != -1) inl.push_back(i);
  } else {
    last.mvObjKeys.resize(n); last.mvObjDepth = dep; last.mvObjFlowNext.resize(n); cur.mvObjKeys.resize(n); cur.vObjLabel.assign(n, 1);
    for (int i = 0; i < n; ++i) { last.mvObjKeys[i].pt = cv::Point2f(pts[2 * i], pts[2 * i + 1]); last.mvObjFlowNext[i] = cv::Point2f(flo[2 * i], flo[2 * i + 1]); ids[i] = i; }
    cur.mInitModel = Tinit;
    T = Optimizer::PoseOptimizationFlow2(&cur, &last, ids, inl);
    n_in = (int)inl.size();
  }
  std::ofstream o(argv[3], std::ios::binary);
  put(o, T.data, 64); put(o, &n_in, 4);
  if (n_in) put(o, inl.data(), 4 * (size_t)n_in);
  for (int i = 0; i < n; ++i) { const cv::Point2f p = mode == 0 ? cur.mvStatKeys[i].pt : cur.mvObjKeys[i].pt; put(o, &p.x, 4); put(o, &p.y, 4); }
  return 0;
}

static int run_ba(int argc, char** argv) {
  if (argc < 6) return 2;
  Reader r(argv[2]);
  const int mode = std::atoi(argv[3]), window = std::atoi(argv[4]);
  Map map;
  cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
  K.at<float>(0, 0) = r.f32(); K.at<float>(1, 1) = r.f32(); K.at<float>(0, 2) = r.f32(); K.at<float>(1, 2) = r.f32();
  const int N = r.i32();
  auto pts3 = [&](int n) { std::vector<cv::Mat> v(n); for (int j = 0; j < n; ++j) { v[j].create(3, 1, CV_32F); r.floats((float*)v[j].data, 3); } return v; };
  for (int i = 0; i < N; ++i) {
    const int ns = r.i32();
    std::vector<cv::KeyPoint> ks(ns); for (int j = 0; j < ns; ++j) { ks[j].pt.x = r.f32(); ks[j].pt.y = r.f32(); }
    std::vector<float> ds(ns); r.floats(ds.data(), ns);
    map.vpFeatSta.push_back(ks); map.vfDepSta.push_back(ds); map.vp3DPointSta.push_back(pts3(ns));
    const int nd = r.i32();
    std::vector<cv::KeyPoint> kd(nd); for (int j = 0; j < nd; ++j) { kd[j].pt.x = r.f32(); kd[j].pt.y = r.f32(); }
    std::vector<float> dd(nd); r.floats(dd.data(), nd);
    map.vpFeatDyn.push_back(kd); map.vfDepDyn.push_back(dd); map.vp3DPointDyn.push_back(pts3(nd));
    map.vmCameraPose.push_back(r.mat4()); map.vmCameraPose_RF.push_back(map.vmCameraPose.back());
    if (i == 0) continue;
    std::vector<int> as(ns), ad(nd), fl(nd);
    r.ints(as.data(), ns); r.ints(ad.data(), nd); r.ints(fl.data(), nd);
    map.vnAssoSta.push_back(as); map.vnAssoDyn.push_back(ad); map.vnFeatLabel.push_back(fl);
    const int nm = r.i32();
    std::vector<cv::Mat> mot(nm); for (int j = 0; j < nm; ++j) mot[j] = r.mat4();
    std::vector<int> lab(nm); r.ints(lab.data(), nm);
    map.vmRigidMotion.push_back(mot); map.vmRigidMotion_RF.push_back(mot); map.vnRMLabel.push_back(lab);
  }
  if (mode == 1) Optimizer::FullBatchOptimization(&map, K); else Optimizer::PartialBatchOptimization(&map, K, window);
  std::ofstream o(argv[5], std::ios::binary);
  for (int i = 0; i < N; ++i) { put(o, map.vmCameraPose[i].data, 64); put(o, map.vmCameraPose_RF[i].data, 64); }
  for (size_t i = 0; i < map.vmRigidMotion.size(); ++i) for (size_t j = 0; j < map.vmRigidMotion[i].size(); ++j) { put(o, map.vmRigidMotion[i][j].data, 64); put(o, map.vmRigidMotion_RF[i][j].data, 64); }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string cmd = argv[1];
  if (cmd == "orb") return run_orb(argc, argv);
  if (cmd == "flow") return run_flow(argc, argv);
  if (cmd == "ba") return run_ba(argc, argv);
  return 2;
}
