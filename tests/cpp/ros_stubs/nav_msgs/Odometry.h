#pragma once
#include <ros/ros.h>
namespace nav_msgs {
struct PoseWithCovariance { geometry_msgs::Pose pose; };
struct TwistWithCovariance { geometry_msgs::Twist twist; };
struct Odometry { std_msgs::Header header; PoseWithCovariance pose; TwistWithCovariance twist; };
}
