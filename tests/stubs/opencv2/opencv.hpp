// stub of <opencv2/opencv.hpp> (tests/stubs/README.md): the value types that cross vloam::VisualOdometry's interface — an 8-bit
// cv::Mat that owns or borrows its rows, cv::KeyPoint, cv::DMatch, cv::Point2f — and a CLAHE handle.  No image processing.
#pragma once
#include <cstring>
#include <memory>
#include <vector>
typedef unsigned char uchar;
#define CV_8UC1 0
namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Size { int width = 0, height = 0; Size() = default; Size(int w, int h) : width(w), height(h) {} };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0;
  DMatch() = default;
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), distance(d) {}
};
class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  Mat() = default;
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), own_(new std::vector<uchar>((size_t)r * c)) { data = own_->data(); }
  Mat(int r, int c, int /*type*/, void* external) : rows(r), cols(c), data(static_cast<uchar*>(external)) {}      // borrows, like OpenCV
  int type() const { return CV_8UC1; }
  bool isContinuous() const { return true; }
  bool empty() const { return rows == 0 || cols == 0; }
  Mat clone() const { Mat m(rows, cols, CV_8UC1); if (data) std::memcpy(m.data, data, (size_t)rows * cols); return m; }
 private:
  std::shared_ptr<std::vector<uchar>> own_;
};
template <typename T> using Ptr = std::shared_ptr<T>;
struct CLAHE { virtual ~CLAHE() = default; virtual void apply(const Mat& src, Mat& dst) { dst = src.clone(); } };
inline Ptr<CLAHE> createCLAHE(double = 40.0, Size = Size(8, 8)) { return std::make_shared<CLAHE>(); }
}  // namespace cv
