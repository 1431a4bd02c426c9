// System.cc -- see System.h.  Mirrors src/System.cc:25-66 (construction, TrackRGBD) over the C ABI.
#include "System.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "vdo_b200.h"

namespace VDO_SLAM {
namespace {
// "Key: value  # comment" lines of an OpenCV FileStorage YAML (the only form the reference's settings files use)
map<string, double> read_settings(const string &path, bool &ok) {
  map<string, double> kv;
  ifstream f(path.c_str());
  ok = f.is_open();
  string line;
  while (ok && getline(f, line)) {
    const size_t h = line.find('#');
    if (h != string::npos) line.erase(h);
    const size_t c = line.find(':');
    if (c == string::npos || line.compare(0, 5, "%YAML") == 0) continue;
    string key = line.substr(0, c), val = line.substr(c + 1);
    key.erase(0, key.find_first_not_of(" \t")); key.erase(key.find_last_not_of(" \t") + 1);
    char *end = nullptr;
    const double v = strtod(val.c_str(), &end);
    if (end != val.c_str()) kv[key] = v;
  }
  return kv;
}
double get(const map<string, double> &kv, const char *k, double dflt = 0.0) {
  map<string, double>::const_iterator it = kv.find(k);
  return it == kv.end() ? dflt : it->second;      // cv::FileNode of a missing key converts to 0 as well
}
}  // namespace

System::System(const string &strSettingsFile, const eSensor sensor) : mSensor(sensor), mpCtx(nullptr), mpTracker(nullptr), mbRGB(true) {
  bool ok = false;
  const map<string, double> kv = read_settings(strSettingsFile, ok);
  if (!ok) {
    cerr << "Failed to open settings file at: " << strSettingsFile << endl;
    exit(-1);
  }
  vdo_tracker_params p;
  vdo_tracker_params_default(&p);
  p.fx = (float)get(kv, "Camera.fx"); p.fy = (float)get(kv, "Camera.fy"); p.cx = (float)get(kv, "Camera.cx"); p.cy = (float)get(kv, "Camera.cy");
  p.width = (int)get(kv, "Camera.width"); p.height = (int)get(kv, "Camera.height");
  p.bf = (float)get(kv, "Camera.bf"); p.depth_factor = (float)get(kv, "DepthMapFactor");
  p.th_depth_bg = (float)get(kv, "ThDepthBG"); p.th_depth_obj = (float)get(kv, "ThDepthOBJ");
  p.max_track_bg = (int)get(kv, "MaxTrackPointBG"); p.max_track_obj = (int)get(kv, "MaxTrackPointOBJ");
  p.sf_mg_thres = (float)get(kv, "SFMgThres"); p.sf_ds_thres = (float)get(kv, "SFDsThres");
  p.n_features = (int)get(kv, "ORBextractor.nFeatures"); p.scale_factor = (float)get(kv, "ORBextractor.scaleFactor");
  p.n_levels = (int)get(kv, "ORBextractor.nLevels"); p.ini_th_fast = (int)get(kv, "ORBextractor.iniThFAST"); p.min_th_fast = (int)get(kv, "ORBextractor.minThFAST");
  p.is_kitti = ((int)get(kv, "ChooseData") == 2) ? 1 : 0;
  p.window_size = (int)get(kv, "WINDOW_SIZE"); p.overlap_size = (int)get(kv, "OVERLAP_SIZE");
  mbRGB = (int)get(kv, "Camera.RGB") != 0;
  mbKitti = p.is_kitti != 0;
  if ((int)get(kv, "UseSampleFeature") == 1) {
    cerr << "UseSampleFeature: 1 draws its samples from cv::RNG(time(NULL)) in the reference and is not reproducible; not supported." << endl;
    exit(-1);
  }
  if (vdo_ctx_create(0, &mpCtx) != VDO_OK) {
    cerr << "vdo_b200: no usable CUDA device (there is no CPU fallback)" << endl;
    exit(-1);
  }
  if (vdo_tracker_create(mpCtx, &p, &mpTracker) != VDO_OK) {
    cerr << "vdo_b200: tracker creation failed: " << vdo_last_error(mpCtx) << endl;
    exit(-1);
  }
}

System::~System() {
  vdo_tracker_destroy(mpTracker);
  vdo_ctx_destroy(mpCtx);
}

cv::Mat System::TrackRGBD(const cv::Mat &im, cv::Mat &depthmap, const cv::Mat &flowmap, const cv::Mat &masksem, const cv::Mat &, const vector<vector<float> > &vObjPose_gt,
                          const double &, cv::Mat &, const int &nImage) {
  if (mSensor != RGBD) {
    cerr << "ERROR: you called TrackRGBD but input sensor was not set to RGBD." << endl;
    exit(-1);
  }
  const int rows = im.rows, cols = im.cols;
  if (!depthmap.isContinuous() || !flowmap.isContinuous() || !masksem.isContinuous() || depthmap.type() != CV_32F || flowmap.type() != CV_32FC2 ||
      masksem.type() != CV_32SC1 || depthmap.rows != rows || flowmap.rows != rows || masksem.rows != rows) {
    cerr << "ERROR: TrackRGBD expects continuous CV_32F depth, CV_32FC2 flow and CV_32SC1 mask of the image size." << endl;
    exit(-1);
  }
  // cvtColor(RGB/BGR(A) -> GRAY) of src/Tracking.cc:209-222 in OpenCV's 8-bit fixed point: (R*4899 + G*9617 + B*1868 + 2^13) >> 14
  mGray.resize((size_t)rows * cols);
  const int ch = im.channels();
  for (int r = 0; r < rows; ++r) {
    const unsigned char *src = im.data + (size_t)r * im.step;
    unsigned char *dst = &mGray[(size_t)r * cols];
    if (ch == 1) memcpy(dst, src, cols);
    else
      for (int c = 0; c < cols; ++c) {
        const unsigned char *px = src + (size_t)c * ch;
        const int R = mbRGB ? px[0] : px[2], G = px[1], B = mbRGB ? px[2] : px[0];
        dst[c] = (unsigned char)((R * 4899 + G * 9617 + B * 1868 + (1 << 13)) >> 14);
      }
  }
  vector<int> gt(vObjPose_gt.size());
  for (size_t i = 0; i < vObjPose_gt.size(); ++i) gt[i] = (int)vObjPose_gt[i][1];
  cv::Mat Tcw = cv::Mat::eye(4, 4, CV_32F);
  float T[16];
  // the mask is declared const in the reference's signature yet mutated through the shared cv::Mat buffer (mSegMap = maskSEM); same here
  const int rc = vdo_tracker_track(mpTracker, mGray.data(), (float *)depthmap.data, (const float *)flowmap.data, (int *)masksem.data, (int)gt.size(),
                                   gt.empty() ? nullptr : gt.data(), 1, T);
  if (rc != VDO_OK) {
    cerr << "vdo_b200: TrackRGBD failed (" << rc << "): " << vdo_tracker_last_error(mpTracker) << endl;
    exit(-1);
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) Tcw.at<float>(i, j) = T[4 * i + j];
  mTrajectory.push_back(vector<float>(T, T + 16));
  // StopFrame = nImage - 1: whole-sequence optimisation after the last frame (src/Tracking.cc:168, 1162-1176; KITTI only)
  if ((int)mTrajectory.size() == nImage && mbKitti && nImage > 2) {
    vdo_lm_stats st;
    if (vdo_tracker_batch_optimize(mpTracker, 1, nullptr, &st, nullptr) != VDO_OK)
      cerr << "vdo_b200: FullBatchOptimization failed: " << vdo_tracker_last_error(mpTracker) << endl;
  }
  return Tcw;
}

void System::SaveResults(const string &filename) {
  ofstream f(filename.c_str());
  f.precision(9);
  for (size_t k = 0; k < mTrajectory.size(); ++k) {
    for (int i = 0; i < 16; ++i) f << mTrajectory[k][i] << (i == 15 ? "\n" : " ");
  }
  // camera poses Twc of the map after the windowed / full batch optimisations (Map::vmCameraPose)
  int n = 0;
  if (vdo_tracker_map_get(mpTracker, "vmCameraPose", nullptr, 0, &n) == VDO_OK && n > 0) {
    vector<float> P(n);
    vdo_tracker_map_get(mpTracker, "vmCameraPose", P.data(), n, &n);
    f << "# refined Twc" << "\n";
    for (int k = 0; k < n / 16; ++k)
      for (int i = 0; i < 16; ++i) f << P[16 * k + i] << (i == 15 ? "\n" : " ");
  }
}

}  // namespace VDO_SLAM
