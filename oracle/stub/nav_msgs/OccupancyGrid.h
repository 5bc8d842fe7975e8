// TEST INFRASTRUCTURE ONLY (oracle/): the fields of nav_msgs::OccupancyGrid that occupancy_grid.cpp:25-35,145-158 touches.
#ifndef NSFS_ORACLE_STUB_OCCUPANCYGRID_H
#define NSFS_ORACLE_STUB_OCCUPANCYGRID_H
#include <string>
#include <vector>
namespace nav_msgs {
struct OccupancyGrid {
    struct { std::string frame_id; } header;
    struct {
        double resolution = 0;
        unsigned width = 0, height = 0;
        struct { struct { double x = 0, y = 0, z = 0; } position; struct { double x = 0, y = 0, z = 0, w = 1; } orientation; } origin;
    } info;
    std::vector<signed char> data;
};
}
#endif
