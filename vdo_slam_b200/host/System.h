// System.h -- host-side mirror of the reference's public entry point for the per-frame path.
//
// Same namespace, class name, enum and member signatures as the reference's include/System.h:29-67, so that example/vdo_slam.cc
// compiles against this header unchanged; the members marshal cv::Mat buffers into the extern "C" entry points of libvdo_b200.so
// (include/vdo_b200.h) and nothing else.  Build it in a tree that has OpenCV's core headers:
//     g++ -std=c++17 -I<repo>/include -I<repo>/vdo_slam_b200/host example/vdo_slam.cc <repo>/vdo_slam_b200/host/System.cc
//         -L<repo>/vdo_slam_b200 -lvdo_b200 `pkg-config --cflags --libs opencv`
// This image has no OpenCV C++ headers; tests/shim_stub/ holds a minimal stand-in for <opencv2/core/core.hpp> that the tests
// use to compile and run the shim here.
#ifndef VDO_B200_SYSTEM_H
#define VDO_B200_SYSTEM_H

#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

struct vdo_ctx;
struct vdo_tracker;

namespace VDO_SLAM {
using namespace std;

class System {
 public:
  enum eSensor { MONOCULAR = 0, STEREO = 1, RGBD = 2 };

  // reads the settings file (the YAML keys Tracking::Tracking reads, src/Tracking.cc:46-162); exits with -1 like the reference when it cannot be opened
  System(const string &strSettingsFile, const eSensor sensor);
  ~System();

  // Returns Tcw (4x4 CV_32F).  depthmap and masksem are updated in place like the reference does (src/Tracking.cc:180-204, :3062).
  // mTcw_gt / timestamp / imTraj only feed evaluation and drawing in the reference and are accepted and ignored here;
  // vObjPose_gt[i][1] (the semantic id of each ground-truth object) gates which objects get a motion estimate (:767-810).
  cv::Mat TrackRGBD(const cv::Mat &im, cv::Mat &depthmap, const cv::Mat &flowmap, const cv::Mat &masksem, const cv::Mat &mTcw_gt,
                    const vector<vector<float> > &vObjPose_gt, const double &timestamp, cv::Mat &imTraj, const int &nImage);

  // camera trajectory and per-frame object motions as text (the reference's SaveResults writes its evaluation files; N4 in SURVEY.md 8f)
  void SaveResults(const string &filename);

 private:
  eSensor mSensor;
  vdo_ctx *mpCtx;
  vdo_tracker *mpTracker;
  bool mbRGB, mbKitti;
  vector<unsigned char> mGray;
  vector<vector<float> > mTrajectory;
};

}  // namespace VDO_SLAM
#endif
