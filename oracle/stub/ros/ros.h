// TEST INFRASTRUCTURE ONLY (oracle/): logging-only stand-in for <ros/ros.h> so the reference's
// planner sources compile unmodified outside a ROS workspace.  Nothing here is shipped.
//
// The real ros.h drags in the libstdc++ <math.h>/<stdlib.h> wrappers, which put the floating-point
// overloads of abs() into the global namespace.  The reference calls unqualified abs() on doubles
// (grid_planner_core.cpp:527, bot_utils.cpp:46-47), so the stub must do the same or sparsifyPath
// silently degenerates to int abs().
#ifndef NSFS_ORACLE_STUB_ROS_H
#define NSFS_ORACLE_STUB_ROS_H
#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>
#include <iostream>
#include <sstream>

#define NSFS_STUB_NOP(...) do { } while (0)
#define NSFS_STUB_COND(c, ...) do { (void)sizeof(c); } while (0)

#define ROS_DEBUG(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_INFO(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_WARN(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_ERROR(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_FATAL(...) NSFS_STUB_NOP(__VA_ARGS__)
#define ROS_DEBUG_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_INFO_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_WARN_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_ERROR_STREAM(x) NSFS_STUB_NOP(x)
#define ROS_DEBUG_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_INFO_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_WARN_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_ERROR_COND(c, ...) NSFS_STUB_COND(c, __VA_ARGS__)
#define ROS_INFO_STREAM_COND(c, x) NSFS_STUB_COND(c, x)
#define ROS_WARN_STREAM_COND(c, x) NSFS_STUB_COND(c, x)

#endif
