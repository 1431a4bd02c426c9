// Optimizer.h -- the reference's Optimizer statics (include/Optimizer.h:25-32) with the same names, argument meaning and effects on
// Frame / Map, marshalled into the extern "C" entry points of libvdo_b200.so (include/vdo_b200.h).  Not carried over:
// PoseOptimizationNew / PoseOptimizationObjMot (the bJoint == false branch the reference's shipped settings never take).
#ifndef VDO_B200_OPTIMIZER_H
#define VDO_B200_OPTIMIZER_H

#include "Frame.h"
#include "Map.h"

struct vdo_ctx;

namespace VDO_SLAM {
using namespace std;

class Optimizer {
 public:
  // src/Optimizer.cc:2333-2542: joint refinement of the camera pose and the static features' flow.  Updates pCurFrame->mTcw and
  // pCurFrame->mvStatKeys, marks outliers -1 in TemperalMatch, returns the inlier count.
  int static PoseOptimizationFlow2Cam(Frame *pCurFrame, Frame *pLastFrame, vector<int> &TemperalMatch);
  // :2755-2972: joint refinement of one object's motion (initialised from pCurFrame->mInitModel) and its points' flow.  Updates
  // pCurFrame->mvObjKeys / vObjLabel, fills InlierID, returns the refined transform.
  cv::Mat static PoseOptimizationFlow2(Frame *pCurFrame, Frame *pLastFrame, const vector<int> &ObjId, std::vector<int> &InlierID);
  // :1232-2175 / :42-1230: whole-sequence / sliding-window bundle adjustment of the Map (results written back like the reference:
  // full -> vmCameraPose_RF, vmRigidMotion_RF; partial -> vmCameraPose, vmRigidMotion; both -> vp3DPointSta / vp3DPointDyn)
  void static FullBatchOptimization(Map *pMap, const cv::Mat Calib_K);
  void static PartialBatchOptimization(Map *pMap, const cv::Mat Calib_K, const int WINDOW_SIZE);
  // :2974-3013
  cv::Mat static Get3DinWorld(const cv::KeyPoint &Feats2d, const float &Dpts, const cv::Mat &Calib_K, const cv::Mat &CameraPose);
  cv::Mat static Get3DinCamera(const cv::KeyPoint &Feats2d, const float &Dpts, const cv::Mat &Calib_K);

  // the device context the statics run on (created on first use on device 0 when none was set)
  static void SetContext(vdo_ctx *ctx);
  static vdo_ctx *Context();
  static bool msQuirk;     // true (default): the reference's arithmetic (see vdo_pose_opt_flow2); false: the intended 2x2 flow blocks
};

}  // namespace VDO_SLAM
#endif
