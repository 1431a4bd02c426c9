// ORBextractor.cc -- see ORBextractor.h.
#include "ORBextractor.h"

#include <cstdlib>
#include <iostream>

#include "Optimizer.h"
#include "vdo_b200.h"

namespace VDO_SLAM {

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST), mpFrame(nullptr), mW(0), mH(0) {
  mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);              // src/ORBextractor.cc:406-421
  mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
    mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
  }
  mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
  mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { if (mpFrame) vdo_frame_destroy(mpFrame); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray, std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors) {
  const cv::Mat &image = _image;
  _keypoints.clear();
  if (image.empty()) return;                                                  // :1037-1038
  if (image.type() != CV_8UC1 || !image.isContinuous()) { std::cerr << "ORBextractor: expects a continuous CV_8UC1 image" << std::endl; exit(-1); }
  if (!mpFrame || image.cols != mW || image.rows != mH) {
    if (mpFrame) vdo_frame_destroy(mpFrame);
    mpFrame = nullptr; mW = image.cols; mH = image.rows;
    if (vdo_frame_create(Optimizer::Context(), mW, mH, &mpFrame) != VDO_OK) { std::cerr << "vdo_b200: vdo_frame_create failed: " << vdo_last_error(Optimizer::Context()) << std::endl; exit(-1); }
  }
  const int cap = nfeatures * 2 + 4096;
  std::vector<float> x(cap), y(cap), resp(cap), ang(cap);
  std::vector<int> oct(cap), sz(cap);
  int n = 0;
  if (vdo_frame_upload(mpFrame, image.data, nullptr, nullptr, nullptr) != VDO_OK ||
      vdo_orb_extract(mpFrame, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, cap, x.data(), y.data(), oct.data(), resp.data(), ang.data(), sz.data(), &n, nullptr) != VDO_OK) {
    std::cerr << "vdo_b200: ORB extraction failed: " << vdo_last_error(Optimizer::Context()) << std::endl;
    exit(-1);
  }
  _keypoints.resize(n);
  for (int i = 0; i < n; ++i) {
    cv::KeyPoint &k = _keypoints[i];
    k.pt.x = x[i]; k.pt.y = y[i]; k.size = (float)sz[i]; k.angle = ang[i]; k.response = resp[i]; k.octave = oct[i]; k.class_id = -1;
  }
  cv::Mat &desc = _descriptors;
  if (n == 0) desc = cv::Mat();
  else {
    desc.create(n, 32, CV_8U);
    if (vdo_orb_describe(mpFrame, n, desc.data) != VDO_OK) {
      std::cerr << "vdo_b200: ORB descriptors failed: " << vdo_last_error(Optimizer::Context()) << std::endl;
      exit(-1);
    }
  }
  for (int l = 0; l < nlevels; ++l) {                                         // the public pyramid (:1112-1137), without the 19-px border
    int w = 0, h = 0;
    if (vdo_frame_debug_level(mpFrame, l, nullptr, nullptr, &w, &h) != VDO_OK) break;
    mvImagePyramid[l].create(h, w, CV_8UC1);
    vdo_frame_debug_level(mpFrame, l, mvImagePyramid[l].data, nullptr, &w, &h);
  }
}

}  // namespace VDO_SLAM
