#pragma once
#include <ros/ros.h>
#include <cmath>
namespace tf {
inline double getYaw(const geometry_msgs::Quaternion& q) { return std::atan2(2 * (q.w * q.z + q.x * q.y), 1 - 2 * (q.y * q.y + q.z * q.z)); }
inline geometry_msgs::Quaternion createQuaternionMsgFromYaw(double yaw) { geometry_msgs::Quaternion q; q.z = std::sin(yaw / 2); q.w = std::cos(yaw / 2); return q; }
}
