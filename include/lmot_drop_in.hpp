// lmot_drop_in.hpp -- the reference's own four C++ entry points, re-implemented as thin wrappers over the C ABI.
//
// A ROS node shell of the reference (object_tracking/src/groundremove/main.cpp, src/cluster/main.cpp,
// tracking/main.cpp) keeps its callbacks unchanged and only swaps
//     #include "ground_removal.h" / "component_clustering.h" / "box_fitting.h" / "imm_ukf_jpda.h"
// for
//     #include "lmot_drop_in.hpp"
//     using namespace lmot_drop_in;
// and links liblmot.so.  Signatures, argument meaning, output conventions (outputs APPENDED to the caller's clouds,
// one tracker entry per track ever created, ...) and the non-re-entrancy (one process-global context, like the
// reference's file-scope globals) are the reference's:
//
//   groundRemove         object_tracking/include/ground_removal.h:62-64
//   componentClustering  object_tracking/include/component_clustering.h:20-22
//   makeClusteredCloud / setObsMsg / createCostMap   object_tracking/include/component_clustering.h:27-37
//   boxFitting           object_tracking/include/box_fitting.h:34-36
//   getOriginPoints      object_tracking/include/imm_ukf_jpda.h:15
//   immUkfJpdaf          object_tracking/include/imm_ukf_jpda.h:19-22
//
// The wrappers are templates over the cloud type so that they compile against PCL
// (pcl::PointCloud<pcl::PointXYZ>, whose point is the 16-byte {x,y,z,pad} record the ABI uses) as well as against
// any container with the same shape (`points` vector of 16-byte points with float x,y,z, `push_back`, `size`).
// Errors: the reference has no error channel (asserts abort); a failing ABI call throws std::runtime_error here.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "lmot.h"

namespace lmot_drop_in {

inline lmot_ctx*& context_slot() { static lmot_ctx* ctx = nullptr; return ctx; }

// process-global context on CUDA device `device` (created on first use; pass parameters before the first call)
inline lmot_ctx* context(const lmot_params* params = nullptr, int device = 0) {
  lmot_ctx*& ctx = context_slot();
  if (!ctx) {
    const int rc = lmot_create(&ctx, params, device);
    if (rc != LMOT_OK) throw std::runtime_error(std::string("lmot_create: ") + lmot_strerror(rc));
  }
  return ctx;
}
inline void shutdown() { lmot_ctx*& ctx = context_slot(); if (ctx) { lmot_destroy(ctx); ctx = nullptr; } }

inline void check(int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + lmot_strerror(rc) + " | " + lmot_last_error(context_slot()));   // rc > 0: warning, outputs valid
}

namespace detail {
template <class Cloud> const float* data(const Cloud& c) {
  static_assert(sizeof(c.points[0]) == 16, "point type must be the 16-byte pcl::PointXYZ layout");
  return c.points.empty() ? nullptr : reinterpret_cast<const float*>(&c.points[0]);
}
template <class Cloud> void append(Cloud& c, const float* xyzw, int n) {
  typename std::remove_reference<decltype(c.points[0])>::type p{};
  for (int i = 0; i < n; ++i) { p.x = xyzw[4 * i]; p.y = xyzw[4 * i + 1]; p.z = xyzw[4 * i + 2]; c.push_back(p); }
}
struct EgoStash { double v = 0, yaw = 0; };
inline EgoStash& ego() { static EgoStash e; return e; }
}  // namespace detail

constexpr int numGrid = LMOT_NUM_GRID;
using CartesianGrid = std::array<std::array<int, numGrid>, numGrid>;

// void groundRemove(PointCloud<PointXYZ>::Ptr cloud, Ptr elevatedCloud, Ptr groundCloud)
template <class CloudPtr>
void groundRemove(CloudPtr cloud, CloudPtr elevatedCloud, CloudPtr groundCloud) {
  const int n = (int)cloud->points.size();
  std::vector<float> e((size_t)(n > 0 ? n : 1) * 4), g((size_t)(n > 0 ? n : 1) * 4);
  int ne = 0, ng = 0;
  check(lmot_ground_remove(context(), detail::data(*cloud), n, 4, nullptr, e.data(), &ne, g.data(), &ng), "lmot_ground_remove");
  detail::append(*elevatedCloud, e.data(), ne);
  detail::append(*groundCloud, g.data(), ng);
}

// void componentClustering(Ptr elevatedCloud, array<array<int,250>,250>& cartesianData, int& numCluster)
template <class CloudPtr>
void componentClustering(CloudPtr elevatedCloud, CartesianGrid& cartesianData, int& numCluster) {
  static_assert(sizeof(CartesianGrid) == sizeof(int) * numGrid * numGrid, "grid layout");
  check(lmot_component_cluster(context(), detail::data(*elevatedCloud), (int)elevatedCloud->points.size(), 4,
                               reinterpret_cast<int32_t*>(&cartesianData[0][0]), &numCluster), "lmot_component_cluster");
}

// The cluster node's other topics (src/cluster/main.cpp:62-99).  All three work on the elevated cloud and label grid the
// context still holds from the componentClustering call just above them in the callback (the arguments are the reference's
// and must be that cloud / that grid; they are not uploaded again).
// void makeClusteredCloud(Ptr& elevatedCloud, array<...> cartesianData, Ptr& clusterCloud)   component_clustering.h:27-29
template <class CloudPtr>
void makeClusteredCloud(CloudPtr& elevatedCloud, const CartesianGrid&, CloudPtr& clusterCloud) {
  const int cap = (int)elevatedCloud->points.size();
  std::vector<float> cl((size_t)(cap > 0 ? cap : 1) * 4);
  int n = 0;
  check(lmot_cluster_outputs(context(), cl.data(), cap, &n, nullptr, 0, nullptr, nullptr), "lmot_cluster_outputs");
  detail::append(*clusterCloud, cl.data(), n);
}
// void setObsMsg(Ptr& elevatedCloud, array<...> cartesianData, object_tracking::ObstacleList& clu_obs)   component_clustering.h:35-37
// ObstacleList: anything with cellLength, cellWidth and `obstacles` (push_back) of elements with x, y, z, cluster
template <class CloudPtr, class ObstacleList>
void setObsMsg(CloudPtr& elevatedCloud, const CartesianGrid&, ObstacleList& clu_obs) {
  std::vector<float> ob((size_t)numGrid * numGrid * 4);
  int n = 0;
  check(lmot_cluster_outputs(context(), nullptr, 0, nullptr, ob.data(), numGrid * numGrid, &n, nullptr), "lmot_cluster_outputs");
  (void)elevatedCloud;
  for (int i = 0; i < n; ++i) {
    typename std::remove_reference<decltype(clu_obs.obstacles[0])>::type o{};
    o.x = ob[4 * i]; o.y = ob[4 * i + 1]; o.z = ob[4 * i + 2]; o.cluster = (int)ob[4 * i + 3];
    clu_obs.cellLength = 0.2f; clu_obs.cellWidth = 0.2f;       // grid_size, component_clustering.h:15
    clu_obs.obstacles.push_back(o);
  }
}
// std::vector<int> createCostMap(const pcl::PointCloud<pcl::PointXYZ>& scan)   component_clustering.h:33
template <class Cloud>
std::vector<int> createCostMap(const Cloud&) {
  std::vector<int32_t> cm(50 * 50);
  check(lmot_cluster_outputs(context(), nullptr, 0, nullptr, nullptr, 0, nullptr, cm.data()), "lmot_cluster_outputs");
  return std::vector<int>(cm.begin(), cm.end());
}

// vector<PointCloud<PointXYZ>> boxFitting(Ptr elevatedCloud, array<...> cartesianData, int numCluster, MarkerArray& ma)
// MarkerArray: anything with `markers` (push_back) whose element has pose.position.{x,y,z} and scale.{x,y,z}
// (visualization_msgs::MarkerArray); the cube marker fields the reference fills besides those (frame "/velodyne",
// ns "cube", green, 1 s lifetime -- box_fitting.cpp:170-207) are constants the node shell can set.
template <class CloudPtr, class MarkerArray>
auto boxFitting(CloudPtr elevatedCloud, const CartesianGrid& cartesianData, int numCluster, MarkerArray& ma)
    -> std::vector<typename std::remove_reference<decltype(*elevatedCloud)>::type> {
  using Cloud = typename std::remove_reference<decltype(*elevatedCloud)>::type;
  const int cap = 1024;
  std::vector<float> boxes((size_t)cap * 24), markers((size_t)cap * 6);
  int nb = 0;
  check(lmot_box_fit(context(), detail::data(*elevatedCloud), (int)elevatedCloud->points.size(), 4,
                     reinterpret_cast<const int32_t*>(&cartesianData[0][0]), numCluster, boxes.data(), cap, &nb, markers.data()),
        "lmot_box_fit");
  std::vector<Cloud> out((size_t)nb);
  for (int b = 0; b < nb; ++b) {
    typename std::remove_reference<decltype(out[b].points[0])>::type p{};
    for (int k = 0; k < 8; ++k) { p.x = boxes[(b * 8 + k) * 3]; p.y = boxes[(b * 8 + k) * 3 + 1]; p.z = boxes[(b * 8 + k) * 3 + 2]; out[b].push_back(p); }
    typename std::remove_reference<decltype(ma.markers[0])>::type m{};
    m.pose.position.x = markers[b * 6]; m.pose.position.y = markers[b * 6 + 1]; m.pose.position.z = markers[b * 6 + 2];
    m.scale.x = markers[b * 6 + 3]; m.scale.y = markers[b * 6 + 4]; m.scale.z = markers[b * 6 + 5];
    ma.markers.push_back(m);
  }
  return out;
}

// void getOriginPoints(double timestamp, vector<vector<double>>& originPoints, double v_gps, double yaw_gps)
// must precede immUkfJpdaf every frame, like in the reference (tracking/main.cpp:74,166)
inline void getOriginPoints(double timestamp, std::vector<std::vector<double>>& originPoints, double v_gps, double yaw_gps) {
  double o[6];
  check(lmot_origin_points(context(), timestamp, v_gps, yaw_gps, o), "lmot_origin_points");
  originPoints.assign(2, std::vector<double>(3));
  for (int i = 0; i < 2; ++i) for (int k = 0; k < 3; ++k) originPoints[i][k] = o[3 * i + k];
  detail::ego().v = v_gps; detail::ego().yaw = yaw_gps;
}

// void immUkfJpdaf(vector<PointCloud<PointXYZ>> bBoxes, double timestamp, PointCloud<PointXYZ>& targets,
//                  vector<vector<double>>& targetVandYaw, vector<int>& trackManage, vector<bool>& isStaticVec,
//                  vector<bool>& isVisVec, vector<PointCloud<PointXYZ>>& visBB)
template <class Cloud>
void immUkfJpdaf(const std::vector<Cloud>& bBoxes, double timestamp, Cloud& targets, std::vector<std::vector<double>>& targetVandYaw,
                 std::vector<int>& trackManage, std::vector<bool>& isStaticVec, std::vector<bool>& isVisVec, std::vector<Cloud>& visBB) {
  const int m = (int)bBoxes.size();
  std::vector<float> boxes((size_t)(m > 0 ? m : 1) * 24);
  for (int b = 0; b < m; ++b)
    for (int k = 0; k < 8; ++k) { boxes[(b * 8 + k) * 3] = bBoxes[b].points[k].x; boxes[(b * 8 + k) * 3 + 1] = bBoxes[b].points[k].y; boxes[(b * 8 + k) * 3 + 2] = bBoxes[b].points[k].z; }
  const int cap = 8192;
  std::vector<float> tg((size_t)cap * 3), vbb((size_t)cap * 24);
  std::vector<double> vy((size_t)cap * 2);
  std::vector<int32_t> tm(cap);
  std::vector<uint8_t> st(cap), vis(cap);
  lmot_track_out o{};
  o.cap = cap; o.targets = tg.data(); o.vandyaw = vy.data(); o.track_manage = tm.data(); o.is_static = st.data(); o.is_vis = vis.data(); o.vis_bb = vbb.data();
  check(lmot_track_step(context(), boxes.data(), m, timestamp, detail::ego().v, detail::ego().yaw, &o), "lmot_track_step");
  typename std::remove_reference<decltype(targets.points[0])>::type p{};
  for (int i = 0; i < o.n_tracks; ++i) {
    p.x = tg[3 * i]; p.y = tg[3 * i + 1]; p.z = tg[3 * i + 2]; targets.push_back(p);
    targetVandYaw.push_back({vy[2 * i], vy[2 * i + 1]});
    trackManage.push_back(tm[i]);
    isStaticVec.push_back(st[i] != 0);
    isVisVec.push_back(vis[i] != 0);
  }
  for (int b = 0; b < o.n_vis; ++b) {
    Cloud c;
    for (int k = 0; k < 8; ++k) { p.x = vbb[(b * 8 + k) * 3]; p.y = vbb[(b * 8 + k) * 3 + 1]; p.z = vbb[(b * 8 + k) * 3 + 2]; c.push_back(p); }
    visBB.push_back(c);
  }
}

}  // namespace lmot_drop_in
