// TEST INFRASTRUCTURE ONLY (oracle/): geometry_msgs::PointStamped as commander.cpp:53,418-423 fills it.
#ifndef NSFS_ORACLE_STUB_POINTSTAMPED_H
#define NSFS_ORACLE_STUB_POINTSTAMPED_H
#include <string>
namespace geometry_msgs {
struct PointStamped {
    struct { std::string frame_id; } header;
    struct { double x = 0, y = 0, z = 0; } point;
};
}
#endif
