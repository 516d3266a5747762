// TEST INFRASTRUCTURE ONLY: API-shaped stand-ins so that ros/src/*_node.cpp can be syntax/type-checked without ROS
// (tests/test_host_cpu.py::test_ros_node_shells_compile).  Nothing here runs.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <memory>
namespace ros {
struct Time { double t = 0; double toSec() const { return t; } static Time now() { return Time(); } };
struct Duration { double d = 0; Duration() {} explicit Duration(double x) : d(x) {} };
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
struct Publisher { template <class M> void publish(const M&) const {} };
struct Subscriber {};
struct NodeHandle {
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <class T> bool param(const std::string&, T& v, const T& d) const { v = d; return false; }
  template <class M> Publisher advertise(const std::string&, int) { return Publisher(); }
  template <class M> Subscriber subscribe(const std::string&, int, void (*)(const std::shared_ptr<const M>&)) { return Subscriber(); }
  template <class M> Subscriber subscribe(const std::string&, int, void (*)(const M&)) { return Subscriber(); }
};
}  // namespace ros
namespace std_msgs { struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Twist { Vector3 linear, angular; };
}
#define ROS_ERROR_THROTTLE(p, ...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_WARN_THROTTLE(p, ...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_FATAL(...) std::fprintf(stderr, __VA_ARGS__)
