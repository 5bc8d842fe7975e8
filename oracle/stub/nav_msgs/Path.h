// TEST INFRASTRUCTURE ONLY (oracle/): nav_msgs::Path as commander.cpp:52,184-187,406-416 fills it.
#ifndef NSFS_ORACLE_STUB_PATH_H
#define NSFS_ORACLE_STUB_PATH_H
#include <string>
#include <vector>
#include "geometry_msgs/PoseStamped.h"
namespace nav_msgs {
struct Path {
    struct { std::string frame_id; } header;
    std::vector<geometry_msgs::PoseStamped> poses;
};
}
#endif
