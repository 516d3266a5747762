#pragma once
// API-shaped stand-in of tf/transform_broadcaster.h (type-checking the node shells without ROS; tests/test_host_cpu.py)
#include <ros/ros.h>
#include <string>
namespace tf {
struct Vector3 { double x, y, z; Vector3(double a = 0, double b = 0, double c = 0) : x(a), y(b), z(c) {} };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; void setRPY(double, double, double) {} };
struct Transform { void setOrigin(const Vector3&) {} void setRotation(const Quaternion&) {} };
struct StampedTransform : Transform {
  StampedTransform(const Transform& t, const ros::Time&, const std::string&, const std::string&) : Transform(t) {}
};
struct TransformBroadcaster { void sendTransform(const StampedTransform&) {} };
}  // namespace tf
