// TEST INFRASTRUCTURE ONLY (oracle/): geometry_msgs::Twist as the commander publishes it.
#ifndef NSFS_ORACLE_STUB_TWIST_H
#define NSFS_ORACLE_STUB_TWIST_H
namespace geometry_msgs {
struct Twist {
    struct { double x = 0, y = 0, z = 0; } linear, angular;
};
}
#endif
