// TEST INFRASTRUCTURE ONLY (oracle/): tmsgs/TurtleSpline.msg (reference src/tmsgs/msg/TurtleSpline.msg).
#ifndef NSFS_ORACLE_STUB_TURTLESPLINE_H
#define NSFS_ORACLE_STUB_TURTLESPLINE_H
#include <cstdint>
#include "nav_msgs/Path.h"
namespace tmsgs { struct TurtleSpline { nav_msgs::Path spline; uint64_t spline_id = 0; double average_speed = 0, target_dt = 0; }; }
#endif
