// Frame.h -- the public data members of the reference's Frame (include/Frame.h:88-196) that the Optimizer statics read and write.
// The reference's Frame also runs the ORB extractor and the sampling in its constructor (src/Frame.cc:61-260); on this path those stages
// live behind the C ABI (vdo_frame_*, vdo_tracker_*), so the shim's Frame is the plain data carrier the optimiser entry points need.
#ifndef VDO_B200_FRAME_H
#define VDO_B200_FRAME_H

#include <vector>

#include <opencv2/core/core.hpp>

namespace VDO_SLAM {

class Frame {
 public:
  Frame() : N_s(0) {}
  void SetPose(cv::Mat Tcw) { mTcw = Tcw; }                 // src/Frame.cc:262-266

  // calibration (static in the reference, include/Frame.h:88-96)
  static float fx, fy, cx, cy, invfx, invfy;

  cv::Mat mTcw;                                             // camera pose
  int N_s;
  std::vector<cv::KeyPoint> mvStatKeys;                     // static features ...
  std::vector<float> mvStatDepth;                           // ... their depths ...
  std::vector<cv::Point2f> mvFlowNext;                      // ... and optical flow to the next frame
  std::vector<cv::KeyPoint> mvObjKeys;                      // semi-dense object samples
  std::vector<float> mvObjDepth;
  std::vector<cv::Point2f> mvObjFlowNext;
  std::vector<int> vObjLabel;                               // -1 outlier / 0 static / 1..n object
  cv::Mat mInitModel;                                       // initial model of the object being refined (GetInitModelObj)
};

}  // namespace VDO_SLAM
#endif
