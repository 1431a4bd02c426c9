// Test driver for the host shim: reads frames dumped by tests/test_host_shim.py (raw little-endian arrays), feeds them through
// VDO_SLAM::System::TrackRGBD exactly as example/vdo_slam.cc does, prints the returned poses and writes the mutated depth / mask back.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "System.h"

static bool slurp(const std::string& p, void* dst, size_t bytes) {
  std::ifstream f(p, std::ios::binary);
  f.read((char*)dst, (std::streamsize)bytes);
  return (size_t)f.gcount() == bytes;
}

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: shim_main settings.yaml dir n_frames width height\n"); return 2; }
  const std::string dir = argv[2];
  const int n = std::atoi(argv[3]), w = std::atoi(argv[4]), h = std::atoi(argv[5]);
  VDO_SLAM::System SLAM(argv[1], VDO_SLAM::System::RGBD);
  cv::Mat imTraj(4, 4, CV_8UC3);
  for (int t = 0; t < n; ++t) {
    const std::string b = dir + "/f" + std::to_string(t);
    cv::Mat im(h, w, CV_8UC3), depth(h, w, CV_32F), flow(h, w, CV_32FC2), mask(h, w, CV_32SC1);
    int ngt = 0;
    if (!slurp(b + ".rgb", im.data, (size_t)w * h * 3) || !slurp(b + ".depth", depth.data, (size_t)w * h * 4) || !slurp(b + ".flow", flow.data, (size_t)w * h * 8) ||
        !slurp(b + ".mask", mask.data, (size_t)w * h * 4) || !slurp(b + ".ngt", &ngt, 4)) { std::fprintf(stderr, "missing frame %d\n", t); return 3; }
    std::vector<int> ids(ngt);
    if (ngt) slurp(b + ".gt", ids.data(), (size_t)ngt * 4);
    std::vector<std::vector<float> > gt;
    // ground-truth rows in the layout of example/vdo_slam.cc (frame, id, box 4, t 3, yaw): a made-up pose per object and frame so that
    // the shim's ground-truth bookkeeping has something to chain
    for (int id : ids) gt.push_back(std::vector<float>{(float)t, (float)id, 0, 0, 0, 0, 1.0f * id, 0.5f, 10.0f + 0.8f * t, 0.01f * t});
    cv::Mat Tcw_gt = cv::Mat::eye(4, 4, CV_32F);
    Tcw_gt.at<float>(2, 3) = 0.8f * t;                       // camera-to-world ground truth as the driver passes it
    cv::Mat Tcw = SLAM.TrackRGBD(im, depth, flow, mask, Tcw_gt, gt, (double)t, imTraj, n);
    std::printf("POSE %d", t);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) std::printf(" %.9g", Tcw.at<float>(i, j));
    std::printf("\n");
    std::ofstream(b + ".depth_out", std::ios::binary).write((const char*)depth.data, (std::streamsize)((size_t)w * h * 4));
    std::ofstream(b + ".mask_out", std::ios::binary).write((const char*)mask.data, (std::streamsize)((size_t)w * h * 4));
  }
  SLAM.SaveResults(dir + "/out_");
  return 0;
}
