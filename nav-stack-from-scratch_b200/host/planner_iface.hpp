// planner_iface.hpp — the interface the GPU planners plug into.
//
// Two build modes:
//   -DNSFS_USE_REFERENCE_HEADERS   the children derive from the reference's OWN GridPlannerCore
//                                  (src/global_planner/include/algo/grid_planner_core.h:20-229) and use its
//                                  bot_utils types (src/bot_utils/include/bot_utils/{bot_utils.h,map_data.h}); this is
//                                  the drop-in build a maintainer makes inside the catkin workspace.
//   default                        a minimal interface-compatible mirror declared below, so this repository builds
//                                  and tests without the reference tree.  Same names, argument meaning and error
//                                  behaviour for the part of the interface the hot path touches; none of the
//                                  reference's CPU search machinery (Node store, OpenList) exists here.
#pragma once

#ifdef NSFS_USE_REFERENCE_HEADERS
#include "grid_planner_core.h"
#else

#include <cmath>
#include <string>
#include <vector>

namespace bot_utils {

struct Index {                       // bot_utils.h:93-100
    int i = 0, j = 0;
    Index() = default;
    Index(int i_, int j_) : i(i_), j(j_) {}
    void setIdx(int i_, int j_) { i = i_; j = j_; }
};

struct Pos2D {                       // bot_utils.h:14-75 (only what crosses the planner interface)
    double x = 0, y = 0;
    double EPS_ = 1e-6;
    Pos2D() = default;
    Pos2D(double x_, double y_) : x(x_), y(y_) {}
    void setCoords(double x_, double y_) { x = x_; y = y_; }
};

struct MapData {                     // map_data.h:8-20
    std::vector<int> grid_inflation_;
    std::vector<int> grid_logodds_;
    int lo_thresh_ = 0;
    int lo_cap_ = 0;
    Pos2D origin_, pos_min_, pos_max_;
    Index map_size_;
    double cell_size_ = 0;
    int total_cells_ = 0;
};

}  // namespace bot_utils

// Public + protected surface of GridPlannerCore that GlobalPlanner and the children use
// (grid_planner_core.h:146-150 pure virtuals, 221-227 public entry points).
class GridPlannerCore {
protected:
    GridPlannerCore() = default;
    virtual std::vector<bot_utils::Pos2D> plan(bot_utils::Index idx_start, bot_utils::Index idx_goal, bot_utils::MapData& map_data) = 0;
    virtual std::vector<bot_utils::Pos2D> plan(bot_utils::Pos2D pos_start, bot_utils::Pos2D pos_goal, bot_utils::MapData& map_data) = 0;
    virtual std::vector<bot_utils::Pos2D> post_process_path(std::vector<bot_utils::Pos2D>& raw_path, bot_utils::MapData& map_data) = 0;

    int map_height_ = 0;
    int map_width_ = 0;
    double cell_size_ = 0;
    bot_utils::Pos2D map_origin_;
    std::string cost_mode_ = "f";

public:
    virtual ~GridPlannerCore() = default;
    // generatePath = plan, then post_process_path (grid_planner_core.cpp:603-615)
    std::vector<bot_utils::Pos2D> generatePath(bot_utils::Index& idx_start, bot_utils::Index& idx_goal, bot_utils::MapData& map_data) {
        auto raw = plan(idx_start, idx_goal, map_data);
        return post_process_path(raw, map_data);
    }
    // grid_planner_core.cpp:610-615: positions are converted with pos2idx, then the Index overload runs
    std::vector<bot_utils::Pos2D> generatePath(bot_utils::Pos2D& pos_start, bot_utils::Pos2D& pos_goal, bot_utils::MapData& map_data) {
        bot_utils::Index a = pos2idx(pos_start), b = pos2idx(pos_goal);
        return generatePath(a, b, map_data);
    }

protected:
    // grid_planner_core.cpp:409-414
    bot_utils::Index pos2idx(bot_utils::Pos2D& pos) {
        const int i = (int)std::round((pos.x - map_origin_.x) / cell_size_);
        const int j = (int)std::round((pos.y - map_origin_.y) / cell_size_);
        return bot_utils::Index(i, j);
    }
};

#endif  // NSFS_USE_REFERENCE_HEADERS
