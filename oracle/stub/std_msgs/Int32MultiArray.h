// TEST INFRASTRUCTURE ONLY (oracle/): std_msgs::Int32MultiArray as occupancy_grid.cpp:145-162 fills it.
#ifndef NSFS_ORACLE_STUB_INT32MULTIARRAY_H
#define NSFS_ORACLE_STUB_INT32MULTIARRAY_H
#include <memory>
#include <vector>
namespace std_msgs {
struct Int32MultiArray { std::vector<int> data; };
typedef std::shared_ptr<const Int32MultiArray> Int32MultiArrayConstPtr;
}
#endif
