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
struct vdo_tracker_params;

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

  // `filename` is a path PREFIX (the reference passes a directory ending in '/'): writes obj_mot_stereo_new.txt, obj_mot_stereo_rf_new.txt,
  // obj_mot_gt.txt, obj_centre.txt, initial_stereo_new.txt, refined_stereo_new.txt, cam_pose_gt_stereo.txt in the reference's text format
  // (src/System.cc:66-193; the writers are vdo_results_* of the C ABI) and prints the mean stage timings (:196-237).  Failures are reported on cerr.
  void SaveResults(const string &filename);

  struct Mat16 { float v[16]; };     // a 4x4 CV_32F matrix of the ground-truth bookkeeping

 private:
  void UpdateGroundTruthMap(const cv::Mat &mTcw_gt, const vector<vector<float> > &vObjPose_gt);

  eSensor mSensor;
  vdo_ctx *mpCtx;
  vdo_tracker *mpTracker;            // created from the first frame's size (the reference never reads Camera.width / Camera.height)
  vdo_tracker_params *mpParams;
  bool mbRGB, mbKitti;
  int mnDataset;                     // ChooseData: 1 OMD, 2 KITTI, 3 VirtualKITTI
  vector<unsigned char> mGray;
  // ground-truth side of the Map (vmCameraPose_GT, vmObjPosePre, vmRigidMotion_GT; src/Tracking.cc:1101-1131), aligned with the tracker's map
  Mat16 mOriginInv, mLastTcwGT;
  vector<int> mLastSemGT;
  vector<Mat16> mLastPoseGT, mvCamPoseGT;
  vector<vector<Mat16> > mvObjPosePre, mvRigidMotionGT;
};

}  // namespace VDO_SLAM
#endif
