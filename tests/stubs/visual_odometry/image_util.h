// stub of the reference's <visual_odometry/image_util.h>: the four OpenCV-backed calls VisualOdometry::processImage may fall back
// to (image_util.h:60-72).  The stub has no OpenCV behind it: a call aborts, so a test that reaches one fails loudly.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <opencv2/opencv.hpp>
#include <tuple>
#include <vector>
namespace vloam {
class ImageUtil {
 public:
  std::vector<cv::KeyPoint> detKeypoints(cv::Mat&) { die("detKeypoints"); return {}; }
  cv::Mat descKeypoints(std::vector<cv::KeyPoint>&, cv::Mat&) { die("descKeypoints"); return {}; }
  std::vector<cv::DMatch> matchDescriptors(const cv::Mat&, const cv::Mat&) { die("matchDescriptors"); return {}; }
  std::tuple<std::vector<cv::Point2f>, std::vector<cv::Point2f>, std::vector<uchar>> calculateOpticalFlow(const cv::Mat&, const cv::Mat&,
                                                                                                        const std::vector<cv::KeyPoint>&) {
    die("calculateOpticalFlow"); return {};
  }
 private:
  static void die(const char* what) { std::fprintf(stderr, "stub ImageUtil::%s called: the OpenCV path is not available here\n", what); std::abort(); }
};
}  // namespace vloam
