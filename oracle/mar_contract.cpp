// TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never linked into the product library.
//
// cv::minAreaRect / cv::RotatedRect::points for the oracle build.
//
// The reference calls OpenCV here (/root/reference/object_tracking/src/cluster/box_fitting.cpp:359-360,
// "Open CV 3.2" per /root/reference/README.md:88).  OpenCV is a system dependency that is NOT vendored
// in /root/reference, and no C++ OpenCV exists in this image, so the arithmetic of this one call is
// PARITY-UNPINNED.  This file restates the published algorithm (convex hull of the integer pixel set,
// then the minimum-area enclosing rectangle having one side collinear with a hull edge -- Freeman &
// Shapira 1975, which is what OpenCV's rotatingCalipers implements) as an EXACT-INTEGER contract:
//
//   1. H = strict convex hull of the integer points (Andrew monotone chain, collinear points dropped),
//      counter-clockwise in the (x right, y down) pixel frame, starting at the lexicographic minimum.
//   2. |H|==1: all four corners = the point.  |H|==2: corners = [h0,h0,h1,h1].
//   3. |H|>=3: for every hull edge d=(dx,dy), l=dx^2+dy^2, W=extent of H along d, T=extent along
//      (-dy,dx) (all exact integers); area = W*T/l compared exactly as a rational; smallest wins,
//      ties -> lowest edge index.
//   4. The winning direction is rotated by a multiple of 90 degrees into the half-open quadrant
//      {ux >= 0, uy < 0} -- the OpenCV<=4.5.0 angle convention theta in [-90,0) -- n=(-uy,ux).
//      Corner(s,t) = ((s*ux - t*uy)/l, (s*uy + t*ux)/l): exact int64 numerators, ONE double division,
//      then narrowed to float.  RotatedRect::points() order = OpenCV's: [P0+v2, P0, P0+v1, P0+v1+v2]
//      with P0=(smin,tmin), v1 along u (size.width), v2 along n (size.height).
//
// No trigonometry is involved, so a GPU implementation of the same contract is bit-identical.
// tests/test_mar_contract.py cross-checks the corner SET against cv2.minAreaRect+cv2.boxPoints (4.13).
#include <opencv2/opencv.hpp>
#include <algorithm>
#include <cstdint>
#include <cmath>

namespace {

struct IPt { int x, y; };

inline long long cross(const IPt& o, const IPt& a, const IPt& b) {
  return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

// strict hull, CCW (in a y-up reading; orientation is irrelevant to the contract), start = lexicographic min
std::vector<IPt> strict_hull(std::vector<IPt> p) {
  std::sort(p.begin(), p.end(), [](const IPt& a, const IPt& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  p.erase(std::unique(p.begin(), p.end(), [](const IPt& a, const IPt& b) { return a.x == b.x && a.y == b.y; }), p.end());
  int n = (int)p.size();
  if (n <= 2) return p;
  std::vector<IPt> h(2 * n);
  int k = 0;
  for (int i = 0; i < n; ++i) {  // lower chain
    while (k >= 2 && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k;
    h[k++] = p[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; --i) {  // upper chain
    while (k >= t && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k;
    h[k++] = p[i];
  }
  h.resize(k - 1);
  return h;
}

}  // namespace

extern "C" int lmot_oracle_mar(const int* xy, int n, float* corners8) {
  // corners8 = 4 x (x,y) in RotatedRect::points() order; returns hull size
  std::vector<IPt> pts(n);
  for (int i = 0; i < n; ++i) { pts[i].x = xy[2 * i]; pts[i].y = xy[2 * i + 1]; }
  std::vector<IPt> h = strict_hull(pts);
  int m = (int)h.size();
  if (m == 0) { for (int i = 0; i < 8; ++i) corners8[i] = 0.f; return 0; }
  if (m == 1) { for (int i = 0; i < 4; ++i) { corners8[2*i] = (float)h[0].x; corners8[2*i+1] = (float)h[0].y; } return 1; }
  if (m == 2) {
    corners8[0] = corners8[2] = (float)h[0].x; corners8[1] = corners8[3] = (float)h[0].y;
    corners8[4] = corners8[6] = (float)h[1].x; corners8[5] = corners8[7] = (float)h[1].y;
    return 2;
  }
  // exact min-area edge
  int best = -1; unsigned __int128 bestNum = 0; long long bestL = 1;
  for (int i = 0; i < m; ++i) {
    long long dx = h[(i + 1) % m].x - h[i].x, dy = h[(i + 1) % m].y - h[i].y;
    long long l = dx * dx + dy * dy;
    long long smin = 0, smax = 0, tmin = 0, tmax = 0;
    for (int j = 0; j < m; ++j) {
      long long s = h[j].x * dx + h[j].y * dy, t = -h[j].x * dy + h[j].y * dx;
      if (j == 0) { smin = smax = s; tmin = tmax = t; }
      else { smin = std::min(smin, s); smax = std::max(smax, s); tmin = std::min(tmin, t); tmax = std::max(tmax, t); }
    }
    unsigned __int128 num = (unsigned __int128)(unsigned long long)(smax - smin) * (unsigned long long)(tmax - tmin);
    // num/l < bestNum/bestL  <=>  num*bestL < bestNum*l
    if (best < 0 || num * (unsigned __int128)(unsigned long long)bestL < bestNum * (unsigned __int128)(unsigned long long)l) {
      best = i; bestNum = num; bestL = l;
    }
  }
  long long ux = h[(best + 1) % m].x - h[best].x, uy = h[(best + 1) % m].y - h[best].y;
  for (int r = 0; r < 4 && !(ux >= 0 && uy < 0); ++r) { long long t = ux; ux = -uy; uy = t; }  // rotate +90 deg
  long long l = ux * ux + uy * uy;
  long long smin = 0, smax = 0, tmin = 0, tmax = 0;
  for (int j = 0; j < m; ++j) {
    long long s = h[j].x * ux + h[j].y * uy, t = -h[j].x * uy + h[j].y * ux;
    if (j == 0) { smin = smax = s; tmin = tmax = t; }
    else { smin = std::min(smin, s); smax = std::max(smax, s); tmin = std::min(tmin, t); tmax = std::max(tmax, t); }
  }
  const long long S[4] = {smin, smin, smax, smax};
  const long long T[4] = {tmax, tmin, tmin, tmax};
  for (int c = 0; c < 4; ++c) {
    corners8[2 * c]     = (float)((double)(S[c] * ux - T[c] * uy) / (double)l);
    corners8[2 * c + 1] = (float)((double)(S[c] * uy + T[c] * ux) / (double)l);
  }
  return m;
}

namespace cv {

RotatedRect minAreaRect(const std::vector<Point>& pts) {
  RotatedRect r;
  float c[8];
  static_assert(sizeof(Point) == 2 * sizeof(int), "cv::Point stub layout");
  lmot_oracle_mar(pts.empty() ? nullptr : &pts[0].x, (int)pts.size(), c);
  for (int i = 0; i < 4; ++i) r.corner[i] = Point2f(c[2 * i], c[2 * i + 1]);
  // informational fields in OpenCV's (center, size, angle-in-degrees) form; the reference only calls points()
  r.center = Point2f((c[0] + c[4]) * 0.5f, (c[1] + c[5]) * 0.5f);
  double v1x = (double)c[4] - c[2], v1y = (double)c[5] - c[3], v2x = (double)c[0] - c[2], v2y = (double)c[1] - c[3];
  r.size = Size2f((float)std::sqrt(v1x * v1x + v1y * v1y), (float)std::sqrt(v2x * v2x + v2y * v2y));
  r.angle = (float)(std::atan2(v1y, v1x) * 180.0 / 3.141592653589793);
  return r;
}

void RotatedRect::points(Point2f pts[]) const {
  for (int i = 0; i < 4; ++i) pts[i] = corner[i];
}

}  // namespace cv
