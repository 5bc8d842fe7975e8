// TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference path consumer (src/commander/src/commander/commander.cpp, compiled
// where it lies under /root/reference together with its local planner and PID controller) outside ROS, for the one step of it
// that follows the planner path: Commander::checkTrajectorySafety (commander.cpp:316-357) with its own pos2idx / oob / checkCell
// (361-402).  Parameters go through the stub's table; map planes, trajectory and t_id are set directly (-fno-access-control).
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "commander.h"

extern "C" {

// xy: n_pts trajectory points in world metres.  Returns 0, or 1 when the reference threw std::out_of_range
// (trajectory_.at(t_id) past the end, or vector::at on a key outside the planes).
int nsref_cmd_check_trajectory(int H, int W, const int32_t* infl, const int32_t* lo, int lo_thresh, double cell, double ox, double oy,
                               int n_pts, const double* xy, int t_id, int* safe, int* bad_idx) {
    auto& p = nsfs_stub::params();
    p.clear();
    p["min_x"] = ox; p["min_y"] = oy; p["max_x"] = ox + H * cell; p["max_y"] = oy + W * cell; p["cell_size"] = cell;
    p["log_odds_thresh"] = lo_thresh;
    ros::NodeHandle nh;
    Commander c(nh);
    if (c.mapdata.map_size_.i != H || c.mapdata.map_size_.j != W) return 2;
    c.mapdata.grid_inflation_.assign(infl, infl + (size_t)H * W);
    c.mapdata.grid_logodds_.assign(lo, lo + (size_t)H * W);
    c.trajectory_.clear();
    for (int t = 0; t < n_pts; ++t) c.trajectory_.emplace_back(xy[2 * t], xy[2 * t + 1]);
    c.t_id = t_id;
    try {
        auto r = c.checkTrajectorySafety();
        *safe = r.first ? 1 : 0;
        *bad_idx = r.second;
    } catch (const std::out_of_range&) {
        return 1;
    }
    return 0;
}

// Commander::pos2idx alone (commander.cpp:386-391), so the host mirror's index conversion is pinned too.
void nsref_cmd_pos2idx(double cell, double ox, double oy, int n, const double* xy, int32_t* ij) {
    auto& p = nsfs_stub::params();
    p.clear();
    p["min_x"] = ox; p["min_y"] = oy; p["max_x"] = ox + 4 * cell; p["max_y"] = oy + 4 * cell; p["cell_size"] = cell;
    ros::NodeHandle nh;
    Commander c(nh);
    for (int t = 0; t < n; ++t) {
        bot_utils::Pos2D q(xy[2 * t], xy[2 * t + 1]);
        bot_utils::Index idx = c.pos2idx(q);
        ij[2 * t] = idx.i; ij[2 * t + 1] = idx.j;
    }
}

}  // extern "C"
