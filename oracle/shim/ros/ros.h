// TEST INFRASTRUCTURE ONLY (oracle build shim): the two ros:: symbols box_fitting.cpp touches.
#pragma once
#include <string>
#include <cstdint>
namespace ros {
struct Time { double sec = 0; static Time now() { return Time(); } double toSec() const { return sec; } };
struct Duration { double sec = 0; Duration() {} explicit Duration(double s) : sec(s) {} };
}  // namespace ros
namespace std_msgs { struct Header { std::uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
}  // namespace geometry_msgs
namespace std_msgs { struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
