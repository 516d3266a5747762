// ROS-free replay of the reference's three node callbacks through include/lmot_drop_in.hpp (host side in C++, like
// the reference).  The call sequence per frame is the reference's:
//   ground  node: groundRemove(cloud, elevated, ground)                         src/groundremove/main.cpp:120
//   cluster node: componentClustering(elevated, grid{}, numCluster = 0);        src/cluster/main.cpp:72-74
//                 boxFitting(elevated, grid, numCluster, ma)                    src/cluster/main.cpp:119
//   track   node: getOriginPoints(ts, ego, v, yaw); immUkfJpdaf(boxes, ts, ..)  tracking/main.cpp:74,166
// usage: drop_in_replay <frames.bin> <n_frames> <n_points> <out.bin>
//   frames.bin: n_frames x n_points x 4 float32; out.bin: per frame int32 {n_elev, n_ground, numCluster, n_boxes, n_tracks}
//   followed by n_boxes*24 float32 and n_tracks int32 trackManage.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include "lmot_drop_in.hpp"

namespace pcl {   // minimal stand-in with PCL's layout; with PCL installed include <pcl/point_cloud.h> instead
struct alignas(16) PointXYZ { float x = 0, y = 0, z = 0, w = 1.f; };
template <class P> struct PointCloud {
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  std::vector<P> points;
  void push_back(const P& p) { points.push_back(p); }
  size_t size() const { return points.size(); }
};
}  // namespace pcl
namespace visualization_msgs {
struct Marker { struct { struct { double x, y, z; } position; } pose; struct { double x, y, z; } scale; };
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs

using namespace lmot_drop_in;
typedef pcl::PointCloud<pcl::PointXYZ> Cloud;

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s frames.bin n_frames n_points out.bin\n", argv[0]); return 2; }
  const int nf = atoi(argv[2]), np = atoi(argv[3]);
  FILE* fi = fopen(argv[1], "rb"); FILE* fo = fopen(argv[4], "wb");
  if (!fi || !fo) return 3;
  std::vector<float> buf((size_t)np * 4);
  for (int f = 0; f < nf; ++f) {
    if (fread(buf.data(), sizeof(float), buf.size(), fi) != buf.size()) return 4;
    Cloud::Ptr cloud(new Cloud), elevated(new Cloud), ground(new Cloud);
    cloud->points.resize(np);
    for (int i = 0; i < np; ++i) { cloud->points[i].x = buf[4 * i]; cloud->points[i].y = buf[4 * i + 1]; cloud->points[i].z = buf[4 * i + 2]; }
    groundRemove(cloud, elevated, ground);
    CartesianGrid grid{};
    int numCluster = 0;
    componentClustering(elevated, grid, numCluster);
    visualization_msgs::MarkerArray ma;
    std::vector<Cloud> boxes = boxFitting(elevated, grid, numCluster, ma);
    const double ts = (f + 1) * 100000.0;
    std::vector<std::vector<double>> ego;
    getOriginPoints(ts, ego, 0.0, 0.0);
    Cloud targets; std::vector<std::vector<double>> vandyaw; std::vector<int> manage; std::vector<bool> isStatic, isVis; std::vector<Cloud> visBB;
    immUkfJpdaf(boxes, ts, targets, vandyaw, manage, isStatic, isVis, visBB);
    const int hdr[5] = {(int)elevated->size(), (int)ground->size(), numCluster, (int)boxes.size(), (int)manage.size()};
    fwrite(hdr, sizeof(int), 5, fo);
    for (const Cloud& b : boxes) for (int k = 0; k < 8; ++k) { const float v[3] = {b.points[k].x, b.points[k].y, b.points[k].z}; fwrite(v, sizeof(float), 3, fo); }
    if (!manage.empty()) fwrite(manage.data(), sizeof(int), manage.size(), fo);
  }
  fclose(fi); fclose(fo);
  shutdown();
  return 0;
}
