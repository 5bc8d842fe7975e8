// TEST INFRASTRUCTURE ONLY (oracle/): the one field of sensor_msgs::LaserScan occupancy_grid.cpp:68-71 reads.
#ifndef NSFS_ORACLE_STUB_LASERSCAN_H
#define NSFS_ORACLE_STUB_LASERSCAN_H
#include <memory>
#include <vector>
namespace sensor_msgs {
struct LaserScan {
    std::vector<float> ranges;
    typedef std::shared_ptr<const LaserScan> ConstPtr;
};
}
#endif
