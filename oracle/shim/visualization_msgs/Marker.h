// TEST INFRASTRUCTURE ONLY (oracle build shim): plain-struct visualization_msgs/Marker.
#pragma once
#include <ros/ros.h>
#include <vector>
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, POINTS = 8 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2 };
  std_msgs::Header header;
  std::string ns;
  std::int32_t id = 0;
  std::int32_t type = 0;
  std::int32_t action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  ros::Duration lifetime;
  bool frame_locked = false;
  std::vector<geometry_msgs::Point> points;
};
}  // namespace visualization_msgs
