// ORBextractor.h -- the reference's ORBextractor (include/ORBextractor.h:36-110) over the C ABI: same constructor, operator(), getters and
// the public mvImagePyramid; the pyramid, FAST, octree distribution and orientation run on the device (vdo_orb_extract).
#ifndef VDO_B200_ORBEXTRACTOR_H
#define VDO_B200_ORBEXTRACTOR_H

#include <vector>

#include <opencv2/core/core.hpp>

struct vdo_ctx;
struct vdo_frame;

namespace VDO_SLAM {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();

  // Keypoints in the reference's order (level-major, octree order inside a level) with pt scaled to level 0, size, angle, response and
  // octave filled (src/ORBextractor.cc:1035-1110).  mask is ignored like in the reference.  descriptors: nkeypoints x 32 CV_8U; the
  // reference allocates it and never fills it (computeDescriptors is commented out, :1091) -- here it holds the rotated-BRIEF descriptors
  // the commented-out call would have produced (vdo_orb_describe: 7x7 sigma-2 blur + 256 pair tests).
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<cv::Mat> mvImagePyramid;

 protected:
  int nfeatures;
  double scaleFactor;
  int nlevels, iniThFAST, minThFAST;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  vdo_frame *mpFrame;
  int mW, mH;
};

}  // namespace VDO_SLAM
#endif
