// TEST INFRASTRUCTURE ONLY (oracle build shim): pcl::getMinMax3D for Vector4f,
// per-axis min/max over all points (PCL public API semantics).
#pragma once
#include <pcl/point_types.h>
#include <cfloat>
namespace pcl {
template <typename PointT>
inline void getMinMax3D(const PointCloud<PointT>& cloud, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (std::size_t i = 0; i < cloud.size(); ++i) {
    const PointT& p = cloud[i];
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  min_pt << mn[0], mn[1], mn[2], 1.f;
  max_pt << mx[0], mx[1], mx[2], 1.f;
}
}  // namespace pcl
