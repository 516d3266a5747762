// CPU test of ros/include/lmot_ros_codec.hpp against plain structs with the field names ROS generates (no ROS here).
#include <cassert>
#include <cstdio>
#include <string>
#include <vector>
#include "../../ros/include/lmot_ros_codec.hpp"

struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 0; };
struct PointCloud2 {
  uint32_t height = 0, width = 0; std::vector<PointField> fields; bool is_bigendian = false; uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data; bool is_dense = false;
};
struct TrackBox { uint8_t box_num = 0; std::vector<float> x1, x2, x3, x4, y1, y2, y3, y4; };
struct Pt { double x, y, z; };

int main() {
  using namespace lmot_ros;
  // library cloud -> PointCloud2 -> layout: zero copy, stride 4
  std::vector<float> cloud;
  for (int i = 0; i < 5; ++i) { cloud.push_back(1.f * i); cloud.push_back(10.f + i); cloud.push_back(-1.f * i); cloud.push_back(1.f); }
  PointCloud2 m;
  fill_pointcloud2_xyz(m, cloud.data(), 5);
  assert(m.width == 5 && m.height == 1 && m.point_step == 16 && m.data.size() == 80 && m.fields.size() == 3);
  XyzLayout L = xyz_layout(m);
  assert(L.n == 5 && L.zero_copy && L.stride_floats() == 4);
  assert(reinterpret_cast<const float*>(L.base)[4 * 3 + 1] == 13.f);
  // a KITTI-style XYZI cloud with an odd layout (intensity first, 20-byte records): gathered
  PointCloud2 k;
  k.height = 1; k.width = 3; k.point_step = 20; k.row_step = 60; k.data.resize(60);
  k.fields = {{"intensity", 0, 7, 1}, {"x", 4, 7, 1}, {"y", 8, 7, 1}, {"z", 12, 7, 1}};
  for (int i = 0; i < 3; ++i) { float rec[5] = {0.5f, 1.f + i, 2.f + i, 3.f + i, 0.f}; memcpy(k.data.data() + 20 * i, rec, 20); }
  XyzLayout K = xyz_layout(k);
  assert(!K.zero_copy && K.off_x == 4);
  std::vector<float> g(3 * 4);
  gather_xyz(K, g.data(), 4);
  assert(g[4] == 2.f && g[5] == 3.f && g[6] == 4.f && g[7] == 1.f);
  // error paths
  PointCloud2 bad = k; bad.fields[1].datatype = 8;
  bool threw = false; try { xyz_layout(bad); } catch (const std::runtime_error&) { threw = true; } assert(threw);
  bad = k; bad.fields.erase(bad.fields.begin() + 3);
  threw = false; try { xyz_layout(bad); } catch (const std::runtime_error&) { threw = true; } assert(threw);
  // trackbox round trip, 300 boxes -> 255 on the wire
  std::vector<float> boxes(300 * 24);
  for (size_t i = 0; i < boxes.size(); ++i) boxes[i] = 0.25f * (float)i;
  TrackBox tb;
  assert(pack_trackbox(tb, boxes.data(), 300) == 255 && tb.box_num == 255 && tb.x1.size() == 255 * 3 && tb.y4.size() == 255 * 3);
  assert(tb.x2[3 * 7 + 1] == boxes[(7 * 8 + 1) * 3 + 1] && tb.y1[3 * 2] == boxes[(2 * 8 + 4) * 3]);
  std::vector<float> back;
  assert(unpack_trackbox(tb, back) == 255);
  for (size_t i = 0; i < back.size(); ++i) assert(back[i] == boxes[i]);
  tb.y3.pop_back();
  threw = false; try { unpack_trackbox(tb, back); } catch (const std::runtime_error&) { threw = true; } assert(threw);
  // box edges
  std::vector<Pt> e;
  box_edges(boxes.data(), e);
  assert(e.size() == 24 && e[1].x == boxes[3] && e[5].z == boxes[4 * 3 + 2]);
  std::puts("ros codec ok");
  return 0;
}
