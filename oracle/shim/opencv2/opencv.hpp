// TEST INFRASTRUCTURE ONLY (oracle build shim): the handful of OpenCV types box_fitting.cpp
// uses.  cv::Mat is a no-op (the reference allocates a 900x900 image per cluster and never reads
// it, box_fitting.cpp:217).  cv::minAreaRect / RotatedRect::points are DECLARED here and DEFINED in
// oracle/mar_contract.cpp -- OpenCV itself is not in /root/reference ("Open CV 3.2", README.md:88,
// un-vendored), so its arithmetic is restated as the documented MAR contract (DESIGN.md,
// "parity unpinned" for this one call).
#pragma once
#include <vector>
#include <cmath>
namespace cv {
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T x_, T y_) : x(x_), y(y_) {} };
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Size_ { T width, height; Size_() : width(0), height(0) {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<float> Size2f;
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0]=a; v[1]=b; v[2]=c; v[3]=d; } };
enum { CV_8UC1 = 0 };
struct Mat { Mat() {} Mat(int, int, int, const Scalar&) {} Mat(float, float, int, const Scalar&) {} };
struct RotatedRect {
  Point2f center; Size2f size; float angle;
  // contract extension: exact corners in pixel coordinates (see mar_contract.cpp)
  Point2f corner[4];
  RotatedRect() : angle(0) {}
  void points(Point2f pts[]) const;
};
RotatedRect minAreaRect(const std::vector<Point>& pts);
}  // namespace cv
