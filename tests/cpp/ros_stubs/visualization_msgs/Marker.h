#pragma once
#include <ros/ros.h>
namespace std_msgs { struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, LINE_LIST = 5, POINTS = 8, ADD = 0 };
  std_msgs::Header header; std::string ns; int32_t id = 0, type = 0, action = 0; geometry_msgs::Pose pose; geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color; ros::Duration lifetime; std::vector<geometry_msgs::Point> points;
};
}
