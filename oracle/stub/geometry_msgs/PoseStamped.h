// TEST INFRASTRUCTURE ONLY (oracle/): the two fields of geometry_msgs::PoseStamped that
// bot_utils.h:3,117 / bot_utils.cpp:104-110 touch, so the reference compiles without ROS.
#ifndef NSFS_ORACLE_STUB_POSESTAMPED_H
#define NSFS_ORACLE_STUB_POSESTAMPED_H
#include <memory>
namespace geometry_msgs {
struct PoseStamped {
    struct { struct { double x = 0, y = 0, z = 0; } position;
             struct { double x = 0, y = 0, z = 0, w = 1; } orientation; } pose;
    typedef std::shared_ptr<const PoseStamped> ConstPtr;
};
typedef std::shared_ptr<const PoseStamped> PoseStampedConstPtr;
}
#endif
