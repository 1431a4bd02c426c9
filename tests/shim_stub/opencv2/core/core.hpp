// Minimal stand-in for <opencv2/core/core.hpp> (TEST INFRASTRUCTURE): just enough of cv::Mat for the host shim
// (vdo_slam_b200/host/System.cc) to compile and run in an image without OpenCV's C++ headers.  Same member names and
// semantics as the real class for the subset used: rows, cols, data, step, type(), channels(), isContinuous(), at<T>(), eye().
#ifndef VDO_TEST_OPENCV_CORE_STUB
#define VDO_TEST_OPENCV_CORE_STUB
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)

namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point3f { float x = 0, y = 0, z = 0; };
struct Point { int x = 0, y = 0; };
struct KeyPoint {
  Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
  KeyPoint() {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data((unsigned char*)ext), type_(type) { step = (size_t)c * elemSize(); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step, (unsigned char)0);
    data = buf_->data();
  }
  int type() const { return type_; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { const int d = type_ & 7; return (size_t)channels() * (d == CV_8U ? 1 : 4); }
  bool isContinuous() const { return true; }
  bool empty() const { return data == nullptr; }
  template <class T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < (r < c ? r : c); ++i) m.at<float>(i, i) = 1.f; return m; }
 private:
  int type_ = 0;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};
typedef const Mat& InputArray;      // enough for the shim: the reference passes cv::Mat for both
typedef Mat& OutputArray;
}  // namespace cv
#endif
