// TEST INFRASTRUCTURE ONLY (oracle build shim): plain-struct object_tracking/Obstacle
// (fields of /root/reference/object_tracking/msg/Obstacle.msg).
#pragma once
#include <ros/ros.h>
namespace object_tracking {
struct Obstacle { double x = 0, y = 0, z = 0, yaw = 0, pitch = 0, roll = 0; std::int32_t cluster = 0; double speed = 0; };
}
