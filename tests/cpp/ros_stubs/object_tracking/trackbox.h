#pragma once
#include <ros/ros.h>
namespace object_tracking { struct trackbox { std_msgs::Header header; uint8_t box_num = 0; std::vector<float> x1, x2, x3, x4, y1, y2, y3, y4; }; }
