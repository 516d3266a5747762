// TEST INFRASTRUCTURE ONLY (oracle build shim): pcl::compute3DCentroid for Vector4f:
// float accumulation in cloud order, then divide by the count (PCL dense-cloud path).
#pragma once
#include <pcl/point_types.h>
namespace pcl {
template <typename PointT>
inline unsigned int compute3DCentroid(const PointCloud<PointT>& cloud, Eigen::Vector4f& centroid) {
  centroid.setZero();
  if (cloud.empty()) return 0;
  for (std::size_t i = 0; i < cloud.size(); ++i) {
    centroid[0] += cloud[i].x; centroid[1] += cloud[i].y; centroid[2] += cloud[i].z;
  }
  centroid /= static_cast<float>(cloud.size());
  centroid[3] = 1.f;
  return (unsigned int)cloud.size();
}
}  // namespace pcl
