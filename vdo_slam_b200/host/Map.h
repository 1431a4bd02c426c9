// Map.h -- the reference's Map as the batch optimisers see it (include/Map.h:34-84): public vectors, filled by Tracking::Track
// (src/Tracking.cc:1016-1105) and read / refined by Optimizer::PartialBatchOptimization / FullBatchOptimization.
#ifndef VDO_B200_MAP_H
#define VDO_B200_MAP_H

#include <utility>
#include <vector>

#include <opencv2/core/core.hpp>

namespace VDO_SLAM {

class Map {
 public:
  // static features, depths and 3-D points per frame (k x n); temporal association (k-1) x n
  std::vector<std::vector<cv::KeyPoint> > vpFeatSta;
  std::vector<std::vector<float> > vfDepSta;
  std::vector<std::vector<cv::Mat> > vp3DPointSta;
  std::vector<std::vector<int> > vnAssoSta;
  std::vector<std::vector<std::pair<int, int> > > TrackletSta;
  // dynamic features
  std::vector<std::vector<cv::KeyPoint> > vpFeatDyn;
  std::vector<std::vector<float> > vfDepDyn;
  std::vector<std::vector<cv::Mat> > vp3DPointDyn;
  std::vector<std::vector<int> > vnAssoDyn;
  std::vector<std::vector<int> > vnFeatLabel;
  std::vector<std::vector<std::pair<int, int> > > TrackletDyn;
  std::vector<int> nObjID;
  // camera poses (k) and rigid motions ((k-1) x m, entry 0 = camera), tracking labels
  std::vector<cv::Mat> vmCameraPose, vmCameraPose_RF, vmCameraPose_GT;
  std::vector<std::vector<cv::Mat> > vmRigidCentre, vmRigidMotion, vmObjPosePre, vmRigidMotion_RF, vmRigidMotion_GT;
  std::vector<std::vector<int> > vnRMLabel, vnSMLabel;
  std::vector<std::vector<bool> > vbObjStat;
};

}  // namespace VDO_SLAM
#endif
