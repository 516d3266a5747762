// lmot_ros_codec.hpp -- message <-> flat-array codecs for the three node shells.
//
// ROS-free: every function is a template over the message type and only touches the public fields ROS generates
// (sensor_msgs/PointCloud2: height, width, fields[{name, offset, datatype, count}], is_bigendian, point_step, row_step, data,
// is_dense; object_tracking/trackbox: box_num, x1..x4, y1..y4; visualization_msgs/Marker: points, ...), so the same code
// compiles against the real headers inside catkin and against the plain structs of tests/cpp/ros_codec_test.cpp here.
//
// What the reference does at these places (for orientation; nothing below is taken from it):
//   pcl::fromROSMsg / pcl::toROSMsg around groundRemove and componentClustering   src/groundremove/main.cpp:102,125-126,
//                                                                                 src/cluster/main.cpp:65,82
//   trackbox packing: 8 corners x 3 floats per box spread over x1..x4 (bottom) and y1..y4 (top)   src/cluster/main.cpp:121-155
//   trackbox unpacking   tracking/main.cpp:101-140
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace lmot_ros {

constexpr uint8_t kFloat32 = 7;          // sensor_msgs/PointField FLOAT32

struct XyzLayout {
  const uint8_t* base = nullptr;         // first point
  int n = 0;                             // width * height
  int point_step = 0;                    // bytes between points
  int off_x = 0, off_y = 4, off_z = 8;
  // true: x, y, z are consecutive float32 at the start of a 4-byte-aligned record -> the buffer can be handed to
  // lmot_ground_remove / lmot_frame_submit as it is, with stride_floats = point_step / 4 (no repacking, no copy)
  bool zero_copy = false;
  int stride_floats() const { return point_step / 4; }
};

// Locate x/y/z in a PointCloud2.  Throws on clouds the hot path cannot read (non-float32 coordinates, big endian).
template <class PointCloud2>
XyzLayout xyz_layout(const PointCloud2& msg) {
  XyzLayout L;
  int found = 0;
  for (const auto& f : msg.fields) {
    int* dst = f.name == "x" ? &L.off_x : f.name == "y" ? &L.off_y : f.name == "z" ? &L.off_z : nullptr;
    if (!dst) continue;
    if (f.datatype != kFloat32 || f.count != 1) throw std::runtime_error("PointCloud2 field " + std::string(f.name) + " is not one float32");
    *dst = (int)f.offset;
    ++found;
  }
  if (found != 3) throw std::runtime_error("PointCloud2 without x/y/z fields");
  if (msg.is_bigendian) throw std::runtime_error("big-endian PointCloud2 is not supported");
  L.point_step = (int)msg.point_step;
  L.n = (int)(msg.width * msg.height);
  if ((size_t)L.n * (size_t)L.point_step > msg.data.size()) throw std::runtime_error("PointCloud2 data shorter than width*height*point_step");
  L.base = msg.data.empty() ? nullptr : msg.data.data();
  L.zero_copy = L.off_x == 0 && L.off_y == 4 && L.off_z == 8 && L.point_step % 4 == 0 && L.point_step >= 12 &&
                (reinterpret_cast<uintptr_t>(L.base) % 4 == 0);
  return L;
}

// General path: gather x, y, z into packed records of `stride` floats (3 or 4; the 4th float is set to 1).
inline void gather_xyz(const XyzLayout& L, float* out, int stride) {
  for (int i = 0; i < L.n; ++i) {
    const uint8_t* p = L.base + (size_t)i * L.point_step;
    std::memcpy(out + (size_t)i * stride, p + L.off_x, 4);
    std::memcpy(out + (size_t)i * stride + 1, p + L.off_y, 4);
    std::memcpy(out + (size_t)i * stride + 2, p + L.off_z, 4);
    if (stride > 3) out[(size_t)i * stride + 3] = 1.f;
  }
}

// A cloud returned by the library (n records of 4 floats: x, y, z, 1 == pcl::PointXYZ) as a PointCloud2 with the layout
// pcl::toROSMsg gives a PointCloud<PointXYZ>: unorganised, fields x/y/z float32 at 0/4/8, point_step 16.
template <class PointCloud2>
void fill_pointcloud2_xyz(PointCloud2& msg, const float* xyzw, int n) {
  msg.height = 1;
  msg.width = (uint32_t)n;
  msg.fields.resize(3);
  const char* names[3] = {"x", "y", "z"};
  for (int k = 0; k < 3; ++k) { msg.fields[k].name = names[k]; msg.fields[k].offset = 4u * k; msg.fields[k].datatype = kFloat32; msg.fields[k].count = 1; }
  msg.is_bigendian = false;
  msg.point_step = 16;
  msg.row_step = 16u * (uint32_t)n;
  msg.is_dense = true;
  msg.data.resize((size_t)n * 16);
  if (n > 0) std::memcpy(msg.data.data(), xyzw, (size_t)n * 16);
}

// boxes = float[m][8][3] (4 bottom corners, then 4 top corners) -> trackbox.  The wire format counts boxes in a uint8:
// at most 255 travel (the C ABI itself is limited only by max_boxes); returns how many were packed.
template <class TrackBox>
int pack_trackbox(TrackBox& msg, const float* boxes, int m) {
  const int n = m > 255 ? 255 : m;
  msg.box_num = (uint8_t)n;
  std::vector<decltype(&msg.x1)> arr{&msg.x1, &msg.x2, &msg.x3, &msg.x4, &msg.y1, &msg.y2, &msg.y3, &msg.y4};
  for (int k = 0; k < 8; ++k) {
    arr[k]->resize((size_t)n * 3);
    for (int b = 0; b < n; ++b)
      for (int c = 0; c < 3; ++c) (*arr[k])[(size_t)b * 3 + c] = boxes[((size_t)b * 8 + k) * 3 + c];
  }
  return n;
}

// trackbox -> boxes float[box_num][8][3]; returns box_num.  Throws if an array is shorter than box_num * 3.
template <class TrackBox>
int unpack_trackbox(const TrackBox& msg, std::vector<float>& boxes) {
  const int n = (int)msg.box_num;
  const std::vector<const decltype(msg.x1)*> arr{&msg.x1, &msg.x2, &msg.x3, &msg.x4, &msg.y1, &msg.y2, &msg.y3, &msg.y4};
  boxes.assign((size_t)n * 24, 0.f);
  for (int k = 0; k < 8; ++k) {
    if (arr[k]->size() < (size_t)n * 3) throw std::runtime_error("trackbox array shorter than box_num");
    for (int b = 0; b < n; ++b)
      for (int c = 0; c < 3; ++c) boxes[((size_t)b * 8 + k) * 3 + c] = (*arr[k])[(size_t)b * 3 + c];
  }
  return n;
}

// The 12 edges of a box as a LINE_LIST point sequence (24 points): bottom ring, top ring, 4 uprights.
// P: any point type with x, y, z.
template <class P>
void box_edges(const float* box /*[8][3]*/, std::vector<P>& out) {
  auto pt = [&](int k) { P p; p.x = box[k * 3]; p.y = box[k * 3 + 1]; p.z = box[k * 3 + 2]; return p; };
  for (int k = 0; k < 4; ++k) {
    out.push_back(pt(k)); out.push_back(pt((k + 1) % 4));
    out.push_back(pt(k + 4)); out.push_back(pt((k + 1) % 4 + 4));
    out.push_back(pt(k)); out.push_back(pt(k + 4));
  }
}

}  // namespace lmot_ros
