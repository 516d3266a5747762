// TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product path.
// Minimal stand-in for the PCL containers the reference hot path touches
// (pcl::PointXYZ 16-byte layout, pcl::PointCloud<T> with std::vector storage).
// Written from the PCL public API; lets /root/reference/object_tracking sources
// compile unmodified into oracle/_ref/ (see oracle/Makefile).
#pragma once
#include <cmath>
#include <cstdint>
#include <array>
#include <vector>
#include <string>
#include <memory>
#include <algorithm>
#include <iostream>
#include <cassert>
#include "Eigen/Dense"

namespace pcl {

struct alignas(16) PointXYZ {
  float x, y, z, w;
  PointXYZ() : x(0.f), y(0.f), z(0.f), w(1.f) {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(1.f) {}
};

struct PointXY { float x, y; };

struct PCLHeader {
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
  std::string frame_id;
};

template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;

  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const PointT& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};

}  // namespace pcl
