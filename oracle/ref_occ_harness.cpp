// TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference map producer (src/loco_mapping/src/occupancy_grid.cpp, compiled
// where it lies under /root/reference) outside ROS: parameters through the stub's table, scans through the spinOnce hook that
// stands in for the pose / scan callbacks, then OccupancyGrid::run() integrates them exactly as the node would
// (occupancy_grid.cpp:91-164).  Outputs: the log-odds and inflation planes the node publishes on grid/mapinfo/*.
#include <cstdint>
#include <cstring>
#include <stdexcept>

#include "occupancy_grid.h"

namespace {
struct Ctx {
    ros::NodeHandle nh;
    OccupancyGrid* grid = nullptr;
};
}

extern "C" {

void* nsref_occ_create(double min_x, double min_y, double max_x, double max_y, double cell, double inflation_radius, int thresh,
                       int cap, double max_scan_range) {
    auto& p = nsfs_stub::params();
    p.clear();
    p["min_x"] = min_x; p["min_y"] = min_y; p["max_x"] = max_x; p["max_y"] = max_y; p["cell_size"] = cell;
    p["inflation_radius"] = inflation_radius; p["log_odds_thresh"] = thresh; p["log_odds_cap"] = cap;
    p["max_scan_range"] = max_scan_range; p["occ_rate"] = 25.0; p["occ_verbose"] = 0;
    auto* c = new Ctx();
    c->grid = new OccupancyGrid(c->nh);
    return c;
}

void nsref_occ_destroy(void* v) { auto* c = static_cast<Ctx*>(v); delete c->grid; delete c; }

void nsref_occ_size(void* v, int* H, int* W, int* n_mask) {
    auto* g = static_cast<Ctx*>(v)->grid;
    *H = g->map_size_.i; *W = g->map_size_.j; *n_mask = (int)g->inflation_mask_.size();
}

// poses: (x, y, heading) per scan; ranges: 360 floats per scan.  Returns 0, or 1 when the reference threw std::out_of_range.
int nsref_occ_integrate(void* v, int n_scans, const double* poses, const float* ranges) {
    auto* g = static_cast<Ctx*>(v)->grid;
    int next = 0;
    auto load = [&](int s) {
        g->robot_position_.setCoords(poses[3 * s], poses[3 * s + 1]);
        g->robot_heading_ = poses[3 * s + 2];
        g->ranges.assign(ranges + 360 * (size_t)s, ranges + 360 * (size_t)(s + 1));
    };
    if (n_scans <= 0) return 0;
    load(0);                                              // the topics have arrived: run() leaves its waiting loop at once
    nsfs_stub::keep_running() = [&]() { return next < n_scans; };
    nsfs_stub::spin_hook() = [&]() { if (next < n_scans) load(next); ++next; };
    int rc = 0;
    try { g->run(); } catch (const std::out_of_range&) { rc = 1; }
    nsfs_stub::keep_running() = nullptr;
    nsfs_stub::spin_hook() = nullptr;
    return rc;
}

// A raw sequence of updateLogOdds(occupy, idx) events (occupancy_grid.cpp:166-196), for targeted edge cases.
int nsref_occ_update(void* v, int n, const int32_t* ij, const uint8_t* occupy) {
    auto* g = static_cast<Ctx*>(v)->grid;
    try {
        for (int t = 0; t < n; ++t) { bot_utils::Index idx(ij[2 * t], ij[2 * t + 1]); g->updateLogOdds(occupy[t] != 0, idx); }
    } catch (const std::out_of_range&) { return 1; }
    return 0;
}

void nsref_occ_get(void* v, int32_t* lo, int32_t* infl) {
    auto* g = static_cast<Ctx*>(v)->grid;
    std::memcpy(lo, g->grid_log_odds_.data(), g->grid_log_odds_.size() * sizeof(int));
    std::memcpy(infl, g->grid_inflation_.data(), g->grid_inflation_.size() * sizeof(int));
}

void nsref_occ_mask(void* v, int32_t* ab) {
    auto* g = static_cast<Ctx*>(v)->grid;
    for (size_t k = 0; k < g->inflation_mask_.size(); ++k) { ab[2 * k] = g->inflation_mask_[k].i; ab[2 * k + 1] = g->inflation_mask_[k].j; }
}

}  // extern "C"
