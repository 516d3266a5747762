#pragma once
#include <object_tracking/Obstacle.h>
#include <vector>
namespace object_tracking {
struct ObstacleList { std_msgs::Header header; double cellLength = 0, cellWidth = 0; std::vector<Obstacle> obstacles; };
}
