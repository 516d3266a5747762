// TEST INFRASTRUCTURE ONLY (oracle build shim): plain-struct nav_msgs/OccupancyGrid.
#pragma once
#include <ros/ros.h>
#include <vector>
namespace nav_msgs {
struct MapMetaData { ros::Time map_load_time; float resolution = 0; std::uint32_t width = 0, height = 0; geometry_msgs::Pose origin; };
struct OccupancyGrid { std_msgs::Header header; MapMetaData info; std::vector<std::int8_t> data; };
}  // namespace nav_msgs
