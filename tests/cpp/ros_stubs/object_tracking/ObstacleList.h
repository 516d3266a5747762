#pragma once
#include <ros/ros.h>
namespace object_tracking {
struct Obstacle { double x = 0, y = 0, z = 0, yaw = 0, pitch = 0, roll = 0; int32_t cluster = 0; double speed = 0; };
struct ObstacleList { std_msgs::Header header; double cellLength = 0, cellWidth = 0; std::vector<Obstacle> obstacles; };
}
