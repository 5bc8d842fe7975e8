// gpu_planners.hpp — B200 children of GridPlannerCore: GpuAstar, GpuDjikstra, GpuThetaStar.
//
// Each child overrides exactly what the reference's children override (astar.h:7-24, djikstra.h:9-32):
//   plan(Index, Index, MapData&), plan(Pos2D, Pos2D, MapData&), post_process_path(vector<Pos2D>&, MapData&)
// and GpuDjikstra adds find_better_point(Index|Pos2D, MapData&) (djikstra.h:29-31).  All search work is done by
// libnsfs_gpu.so through the C-ABI of include/nsfs.h; what stays on the host is the double-precision index<->world
// conversion (grid_planner_core.cpp:409-421) and the collinearity filter sparsifyPath (grid_planner_core.cpp:506-535),
// both a few dozen flops per path.  There is no CPU search fallback: if the CUDA library cannot run, construction throws.
//
// Error behaviour mirrors the reference: an expansion that would make vector::at throw surfaces as
// std::out_of_range; "no path" is the single start position; a failed fallback returns the input position.
#pragma once

#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "nsfs.h"
#include "planner_iface.hpp"

namespace nsfs_host {

// When does the planner re-upload MapData?  The reference reads the caller's vectors afresh on every call
// (global_planner.cpp:65-73 replaces them whenever a ROS message arrives), so every call uploads them by default
// (EveryCallPinned; EveryCall is the same without touching the caller's memory: a pageable copy, ~1.7x slower at 2048^2).
// EveryCallPinned: same contract as EveryCall, but the caller's two vectors are page-locked once (cudaHostRegister through
// nsfs_pin_host) and stay registered while their data pointers and sizes do not change (the ROS callbacks copy-assign into
// the same vectors, global_planner.cpp:65-73, so they do not), which makes the per-call upload a DMA at PCIe speed.
enum class MapSync { EveryCall, OnNotify, EveryCallPinned };

// Results of a batch of plan() calls on one map (no reference counterpart: GlobalPlanner plans one request at a time;
// the batch is what BASELINE.json's config "batched 65,536 Astar start/goal queries" runs).  Same per-query values as
// plan(): status (NSFS_OK / NSFS_E_RANGE where plan() would throw std::out_of_range), nodes[goal].g, raw path length in
// cells (1 = no path, astar.cpp:129-135), settled cells.
struct BatchPlan {
    std::vector<int32_t> status, path_len;
    std::vector<double> g_goal;
    std::vector<long long> settled;
    nsfs_stats stats{};
};

class GpuPlannerBase : public GridPlannerCore {
public:
    ~GpuPlannerBase() override;
    // reference: GridPlannerCore::prepPlanner(cost_mode, map) (grid_planner_core.cpp:570-600)
    void prepPlanner(std::string cost_mode, bot_utils::MapData& map_data);
    void setMapSync(MapSync m) { sync_ = m; }
    void notifyMapChanged() { dirty_ = true; }
    void setDevice(int device) { device_ = device; }
    // Opt-in: search the textbook 8-connected grid (rows and columns bounds-checked, diagonal steps
    // sqrt(2)) instead of the reference's quirk graph.  Default false = reference behaviour.
    void setTextbookGraph(bool on);
    // nq independent plan(Index, Index, map) calls of this child on one upload of the map (sg4 = {si, sj, gi, gj} per query)
    BatchPlan planBatch(const std::vector<int32_t>& sg4, bot_utils::MapData& map_data);
    // The same batch split over a DeviceGroup: `root` uploads its MapData and broadcasts the packed map, every rank plans its
    // slice on its own GPU and every rank receives all results.  map_data is read on the root only.
    BatchPlan planBatchSharded(class DeviceGroup& group, const std::vector<int32_t>& sg4, bot_utils::MapData& map_data, int root = 0);
    const nsfs_stats& lastStats() const { return stats_; }
    double lastGoalCost() const { return last_g_goal_; }   // nodes[goal].g of the last plan (1e5 = unreachable)
    // The path consumer's check of a trajectory against the current map: Commander::checkTrajectorySafety
    // (commander.cpp:316-357) with its pos2idx / checkCell / oob (361-397).  {safe, first unsafe index | -1 | -2}; throws
    // std::out_of_range where the reference's vector::at does.
    std::pair<bool, int> checkTrajectorySafety(const std::vector<bot_utils::Pos2D>& trajectory, int t_id, bot_utils::MapData& map_data);

protected:
    GpuPlannerBase();
    GpuPlannerBase(std::string cost_mode, bot_utils::MapData& map_data);

    std::vector<bot_utils::Pos2D> post_process_path(std::vector<bot_utils::Pos2D>& raw_path, bot_utils::MapData& map_data) override;

    // host-side helpers restating grid_planner_core.cpp:409-421 and 506-535 in double
    bot_utils::Index toIndex(const bot_utils::Pos2D& pos) const;
    bot_utils::Pos2D toPos(int i, int j) const;
    static std::vector<bot_utils::Pos2D> sparsify(const std::vector<bot_utils::Pos2D>& path_world);

    void ensureHandle(bot_utils::MapData& map_data);
    void uploadMap(bot_utils::MapData& map_data);
    std::vector<bot_utils::Pos2D> runPlan(int algo, bot_utils::Index s, bot_utils::Index g, bot_utils::MapData& map_data);
    virtual int algoCode() const = 0;
    static int modeCode(const std::string& cost_mode);
    [[noreturn]] void fail(int code, const char* where) const;

    nsfs_t* handle_ = nullptr;
    int device_ = 0;
    MapSync sync_ = MapSync::EveryCallPinned;     // the reference's contract (re-read the caller's vectors on every call) at PCIe speed
    bool dirty_ = true;
    bool textbook_graph_ = false;
    nsfs_stats stats_{};
    double last_g_goal_ = 0.0;
    std::vector<int32_t> path_buf_;
    const void* pinned_[2] = {nullptr, nullptr};     // EveryCallPinned: the registered vectors' data pointers
    size_t pinned_bytes_[2] = {0, 0};
#ifdef NSFS_USE_REFERENCE_HEADERS
    // the reference base class already owns map_height_, map_width_, cell_size_, map_origin_, cost_mode_
#endif
};

class GpuAstar : public GpuPlannerBase {
public:
    GpuAstar() = default;
    GpuAstar(std::string cost_mode, bot_utils::MapData& map_data) : GpuPlannerBase(std::move(cost_mode), map_data) {}
protected:
    int algoCode() const override { return NSFS_ASTAR; }
    std::vector<bot_utils::Pos2D> plan(bot_utils::Index idx_start, bot_utils::Index idx_goal, bot_utils::MapData& map_data) override;
    std::vector<bot_utils::Pos2D> plan(bot_utils::Pos2D pos_start, bot_utils::Pos2D pos_goal, bot_utils::MapData& map_data) override;
};

class GpuDjikstra : public GpuPlannerBase {
public:
    // Backend::Exact replays the reference's open-list order (identical path); Backend::Field relaxes the whole
    // cost-to-go field in parallel (identical reachability and cost, possibly another equal-cost path; much faster
    // on large maps).  Field falls back to Exact only where the reference's behaviour is order dependent (a cell whose
    // expansion throws is reachable), never to the CPU.
    enum class Backend { Exact, Field };
    GpuDjikstra() = default;
    GpuDjikstra(std::string cost_mode, bot_utils::MapData& map_data, Backend b = Backend::Exact)
        : GpuPlannerBase(std::move(cost_mode), map_data), backend_(b) {}
    void setBackend(Backend b) { backend_ = b; }
    bot_utils::Pos2D find_better_point(bot_utils::Index idx_bad, bot_utils::MapData& map_data);
    bot_utils::Pos2D find_better_point(bot_utils::Pos2D pos_bad, bot_utils::MapData& map_data);
protected:
    int algoCode() const override { return NSFS_DJIKSTRA; }
    std::vector<bot_utils::Pos2D> plan(bot_utils::Index idx_start, bot_utils::Index idx_goal, bot_utils::MapData& map_data) override;
    std::vector<bot_utils::Pos2D> plan(bot_utils::Pos2D pos_start, bot_utils::Pos2D pos_goal, bot_utils::MapData& map_data) override;
private:
    Backend backend_ = Backend::Exact;
};

// The reference ships no ThetaStar (thetastar.cpp is empty, thetastar.h:4-7 an empty class, the selection branch is
// commented out at global_planner.cpp:116-119).  This child is Basic Theta* assembled from the reference's own
// primitives (OpenList order, testPos, isLos/bresenham_los, dist_euc); DESIGN.md states the rule.
class GpuThetaStar : public GpuPlannerBase {
public:
    GpuThetaStar() = default;
    GpuThetaStar(std::string cost_mode, bot_utils::MapData& map_data) : GpuPlannerBase(std::move(cost_mode), map_data) {}
protected:
    int algoCode() const override { return NSFS_THETASTAR; }
    std::vector<bot_utils::Pos2D> plan(bot_utils::Index idx_start, bot_utils::Index idx_goal, bot_utils::MapData& map_data) override;
    std::vector<bot_utils::Pos2D> plan(bot_utils::Pos2D pos_start, bot_utils::Pos2D pos_goal, bot_utils::MapData& map_data) override;
};

// ---- several GPUs, one planner process per GPU ------------------------------------------------------
// RAII over nsfs_group_*: the NCCL communicator, the map broadcast and the gather of batch results all happen below the
// C-ABI, so a GlobalPlanner-side caller needs nothing but the 128-byte group id from rank 0.
class DeviceGroup {
public:
    static std::vector<unsigned char> makeId();                                   // rank 0
    DeviceGroup(const std::vector<unsigned char>& id, int rank, int world, int device);
    ~DeviceGroup();
    DeviceGroup(const DeviceGroup&) = delete;
    DeviceGroup& operator=(const DeviceGroup&) = delete;
    int rank() const { return nsfs_group_rank(g_); }
    int world() const { return nsfs_group_world(g_); }
    void barrier();
    nsfs_group_t* raw() const { return g_; }
private:
    nsfs_group_t* g_ = nullptr;
};

// ---- runtime selection (global_planner.cpp:102-130, 482-490) --------------------------------------------------
struct PlannerConfig {
    std::string planner_name = "astar";   // "astar" | "djikstra" | "thetastar"; anything else => astar (global_planner.cpp:121-125)
    std::string cost_mode = "fg";         // "f" | "g" | "fg"; anything else => fg (grid_planner_core.cpp:274-281)
    std::string djikstra_backend = "exact";  // "exact" | "field"
    std::string graph = "reference";      // "reference" (quirk graph, default) | "textbook" (plain 8-connected grid; optimal with djikstra)
    double gp_rate = 25.0;
    bool verbose_planner = false;
    int device = 0;
};
// Reads the planner keys of a rosparam-style YAML (flat "key: value" lines; core_turtle_config.yaml:33-36).
PlannerConfig load_planner_config(const std::string& yaml_path);
// GlobalPlanner::setupPlanner's choice (global_planner.cpp:102-125) with the GPU children.
std::shared_ptr<GridPlannerCore> make_main_planner(const PlannerConfig& cfg, bot_utils::MapData& map_data);
// The fallback planner: a Djikstra prepared with cost mode "g" (global_planner.cpp:127-128).
std::shared_ptr<GpuDjikstra> make_fallback_planner(const PlannerConfig& cfg, bot_utils::MapData& map_data);

// ---- one plan request, as GlobalPlanner::run serves it (global_planner.cpp:164-273) ---------------------------------
// Test the robot and goal cells (GlobalPlanner::testPos, global_planner.cpp:384-408, on the caller's MapData); repair
// whichever is not free with the fallback planner (cases 2-4: 188-273); then generatePath.  No ROS in here: the caller
// publishes `path` (writeToPathMsg) and, when goal_updated, tells the mission planner (192-205).
struct PlanRequestResult {
    std::vector<bot_utils::Pos2D> path;     // goal -> start, post-processed
    int which_case = 0;                     // 1: both ok, 2: goal repaired, 3: robot repaired (backup_mode_), 4: both repaired
    bool robot_ok = false, goal_ok = false;
    bool goal_updated = false;              // current_goal_ was replaced by `goal`
    bool backup_mode = false;               // the path starts at `backup_robot`, not at the robot
    bot_utils::Pos2D goal, backup_robot;
};
bool test_pos(bot_utils::Pos2D pos, bot_utils::MapData& map_data);
PlanRequestResult serve_plan_request(GridPlannerCore& main_planner, GpuDjikstra& fallback_planner, bot_utils::Pos2D robot_position,
                                     bot_utils::Pos2D current_goal, bot_utils::MapData& map_data);

}  // namespace nsfs_host
