// gpu_planners.cpp — B200 children of GridPlannerCore (see gpu_planners.hpp) and a small C shim for the tests.
//
// Host-side work kept here, all in double like the reference:
//   pos2idx / idx2pos         grid_planner_core.cpp:409-421
//   sparsifyPath              grid_planner_core.cpp:506-535 (collinearity filter, |cross| > 1e-5)
//   post_process_path         astar.cpp:150-165 == djikstra.cpp:147-163  (<= 2 points: unchanged)
// Everything that searches the grid (plan, find_better_point, the isLos loop of makeAnyAnglePath) is a call into
// libnsfs_gpu.so.  No CPU search exists in this file.
#include "gpu_planners.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace nsfs_host {

using bot_utils::Index;
using bot_utils::MapData;
using bot_utils::Pos2D;

// ---- GpuPlannerBase ----------------------------------------------------------------------------------------
GpuPlannerBase::GpuPlannerBase() : GridPlannerCore() {
    // reference default ctor (grid_planner_core.cpp:341-351): cost mode "f", geometry unset until prepPlanner
    cost_mode_ = "f";
    map_height_ = 0; map_width_ = 0; cell_size_ = 0.0;
}

// Deliberately NOT GridPlannerCore(cost_mode, map): that ctor allocates the 40-byte-per-cell CPU Node store
// (grid_planner_core.cpp:353-383), which the GPU planners never touch.
GpuPlannerBase::GpuPlannerBase(std::string cost_mode, MapData& map_data) : GridPlannerCore() {
    prepPlanner(std::move(cost_mode), map_data);
}

GpuPlannerBase::~GpuPlannerBase() {
    for (int k = 0; k < 2; ++k)
        if (pinned_[k] && handle_) nsfs_unpin_host(handle_, pinned_[k]);
    if (handle_) nsfs_destroy(handle_);
    handle_ = nullptr;
}

void GpuPlannerBase::prepPlanner(std::string cost_mode, MapData& map_data) {
    cost_mode_ = std::move(cost_mode);
    map_height_ = map_data.map_size_.i;
    map_width_ = map_data.map_size_.j;
    cell_size_ = map_data.cell_size_;
    map_origin_.setCoords(map_data.origin_.x, map_data.origin_.y);
    if (handle_) {
        for (int k = 0; k < 2; ++k) { if (pinned_[k]) nsfs_unpin_host(handle_, pinned_[k]); pinned_[k] = nullptr; pinned_bytes_[k] = 0; }
        nsfs_destroy(handle_);
        handle_ = nullptr;
    }
    dirty_ = true;
    ensureHandle(map_data);
}

[[noreturn]] void GpuPlannerBase::fail(int code, const char* where) const {
    if (code == NSFS_E_RANGE)     // the reference's vector::at inside testPos (grid_planner_core.cpp:434-455)
        throw std::out_of_range(std::string("vector::_M_range_check (testPos) in ") + where);
    std::string msg = std::string(where) + ": " + nsfs_strerror(code);
    if (code == NSFS_E_CUDA && handle_) msg += std::string(" | ") + nsfs_last_cuda_error(handle_);
    throw std::runtime_error(msg);
}

void GpuPlannerBase::ensureHandle(MapData& map_data) {
    if (handle_) return;
    if (map_height_ <= 0 || map_width_ <= 0) {
        map_height_ = map_data.map_size_.i;
        map_width_ = map_data.map_size_.j;
        cell_size_ = map_data.cell_size_;
        map_origin_.setCoords(map_data.origin_.x, map_data.origin_.y);
    }
    const int rc = nsfs_create(&handle_, map_height_, map_width_, device_);
    if (rc != NSFS_OK) { handle_ = nullptr; fail(rc, "nsfs_create"); }
    if (textbook_graph_) nsfs_set_option(handle_, "graph", 1);
    dirty_ = true;
}

void GpuPlannerBase::setTextbookGraph(bool on) {
    textbook_graph_ = on;
    if (handle_) {
        const int rc = nsfs_set_option(handle_, "graph", on ? 1 : 0);
        if (rc != NSFS_OK) fail(rc, "nsfs_set_option(graph)");
    }
}

void GpuPlannerBase::uploadMap(MapData& map_data) {
    ensureHandle(map_data);
    if (sync_ == MapSync::OnNotify && !dirty_) return;
    const size_t n = (size_t)map_height_ * (size_t)map_width_;
    if (map_data.grid_inflation_.size() < n || map_data.grid_logodds_.size() < n)
        throw std::out_of_range("MapData grids are smaller than map_size_");       // what .at() would report
    static_assert(sizeof(int) == sizeof(int32_t), "MapData stores int planes");
    if (sync_ == MapSync::EveryCallPinned) {
        // page-lock the caller's vectors once; re-register only when a vector was reallocated (pointer or size changed)
        const void* ptr[2] = {map_data.grid_inflation_.data(), map_data.grid_logodds_.data()};
        for (int k = 0; k < 2; ++k) {
            if (pinned_[k] == ptr[k] && pinned_bytes_[k] == n * sizeof(int)) continue;
            if (pinned_[k]) nsfs_unpin_host(handle_, pinned_[k]);
            pinned_[k] = nullptr; pinned_bytes_[k] = 0;
            if (nsfs_pin_host(handle_, ptr[k], n * sizeof(int)) == NSFS_OK) { pinned_[k] = ptr[k]; pinned_bytes_[k] = n * sizeof(int); }
            // a failed registration (e.g. the range is already registered by the caller) is not an error: the copy just stays pageable
        }
    }
    const int rc = nsfs_set_map(handle_, reinterpret_cast<const int32_t*>(map_data.grid_inflation_.data()),
                                reinterpret_cast<const int32_t*>(map_data.grid_logodds_.data()), map_data.lo_thresh_);
    if (rc != NSFS_OK) fail(rc, "nsfs_set_map");
    dirty_ = false;
}

Index GpuPlannerBase::toIndex(const Pos2D& pos) const {
    const int i = (int)std::round((pos.x - map_origin_.x) / cell_size_);
    const int j = (int)std::round((pos.y - map_origin_.y) / cell_size_);
    return Index(i, j);
}

Pos2D GpuPlannerBase::toPos(int i, int j) const {
    return Pos2D(i * cell_size_ + map_origin_.x, j * cell_size_ + map_origin_.y);
}

int GpuPlannerBase::modeCode(const std::string& cost_mode) {
    if (cost_mode == "f") return NSFS_MODE_F;
    if (cost_mode == "g") return NSFS_MODE_G;
    return NSFS_MODE_FG;                              // unknown => fg (grid_planner_core.cpp:274-281)
}

std::vector<Pos2D> GpuPlannerBase::sparsify(const std::vector<Pos2D>& path_world) {
    if (path_world.size() <= 2) return path_world;
    std::vector<Pos2D> turning{path_world.front()};
    for (size_t n = 2; n < path_world.size(); ++n) {
        const Pos2D& nx = path_world[n];
        const Pos2D& cu = path_world[n - 1];
        const Pos2D& pv = path_world[n - 2];
        const double dxn = nx.x - cu.x, dyn = nx.y - cu.y, dxp = cu.x - pv.x, dyp = cu.y - pv.y;
        if (std::fabs(dxn * dyp - dyn * dxp) > 1e-5) turning.push_back(cu);
    }
    turning.push_back(path_world.back());
    return turning;
}

std::vector<Pos2D> GpuPlannerBase::post_process_path(std::vector<Pos2D>& raw_path, MapData& map_data) {
    if (raw_path.size() <= 2) return raw_path;
    std::vector<Pos2D> tp = sparsify(raw_path);
    // makeAnyAnglePath (grid_planner_core.cpp:537-558): the greedy isLos chain runs on the device
    const int n = (int)tp.size();
    if (n <= 2) return tp;
    ensureHandle(map_data);
    std::vector<int32_t> ij(2 * (size_t)n);
    for (int t = 0; t < n; ++t) { const Index id = toIndex(tp[t]); ij[2 * t] = id.i; ij[2 * t + 1] = id.j; }
    std::vector<uint8_t> keep((size_t)n, 0);
    const int rc = nsfs_any_angle(handle_, n, ij.data(), keep.data());
    if (rc != NSFS_OK) fail(rc, "nsfs_any_angle");
    std::vector<Pos2D> out;
    for (int t = 0; t < n; ++t) if (keep[t]) out.push_back(tp[t]);
    return out;
}

std::pair<bool, int> GpuPlannerBase::checkTrajectorySafety(const std::vector<Pos2D>& trajectory, int t_id, MapData& map_data) {
    uploadMap(map_data);
    std::vector<int32_t> ij(2 * trajectory.size());
    for (size_t t = 0; t < trajectory.size(); ++t) {              // Commander::pos2idx (commander.cpp:386-391): the same rounding as core's
        const Index id = toIndex(trajectory[t]);
        ij[2 * t] = id.i; ij[2 * t + 1] = id.j;
    }
    int safe = 0, bad = -1;
    const int rc = nsfs_check_trajectory(handle_, (int)trajectory.size(), ij.data(), t_id, &safe, &bad);
    if (rc != NSFS_OK) fail(rc, "nsfs_check_trajectory");
    return {safe != 0, bad};
}

std::vector<Pos2D> GpuPlannerBase::runPlan(int algo, Index s, Index g, MapData& map_data) {
    uploadMap(map_data);
    if (s.i < 0 || s.i >= map_height_ || s.j < 0 || s.j >= map_width_)
        throw std::out_of_range("start index outside the node store");     // nodes[key] is undefined behaviour upstream
    if (path_buf_.empty()) path_buf_.resize(2 * (size_t)(4 * (map_height_ + map_width_) + 64));
    int n = 0;
    double g_goal = 0.0;
    int rc = nsfs_plan(handle_, algo, modeCode(cost_mode_), s.i, s.j, g.i, g.j, path_buf_.data(), (int)(path_buf_.size() / 2), &n,
                       &g_goal, &stats_);
    if (rc == NSFS_E_TRUNC) {              // path longer than the buffer: its exact length is known now; fetch it without searching again
        path_buf_.resize(2 * (size_t)n);
        int n2 = 0;
        rc = nsfs_last_path(handle_, path_buf_.data(), n, &n2);
        if (rc == NSFS_OK && n2 == n) {
            // the parent store (or predecessor field) of the search that just ran was walked again: same path, whole this time
        } else {
            rc = nsfs_plan(handle_, algo, modeCode(cost_mode_), s.i, s.j, g.i, g.j, path_buf_.data(), n, &n, &g_goal, &stats_);
        }
    }
    if (rc != NSFS_OK) fail(rc, "nsfs_plan");
    last_g_goal_ = g_goal;
    std::vector<Pos2D> path_world;
    path_world.reserve((size_t)n);
    for (int t = 0; t < n; ++t) path_world.push_back(toPos(path_buf_[2 * t], path_buf_[2 * t + 1]));
    return path_world;
}

BatchPlan GpuPlannerBase::planBatch(const std::vector<int32_t>& sg4, MapData& map_data) {
    uploadMap(map_data);
    const int nq = (int)(sg4.size() / 4);
    BatchPlan out;
    out.status.assign((size_t)nq, 0); out.path_len.assign((size_t)nq, 0); out.g_goal.assign((size_t)nq, 0.0); out.settled.assign((size_t)nq, 0);
    if (!nq) return out;
    const int rc = nsfs_plan_batch(handle_, algoCode(), modeCode(cost_mode_), nq, sg4.data(), out.status.data(), out.g_goal.data(),
                                   out.path_len.data(), out.settled.data(), nullptr, 0, &out.stats);
    if (rc != NSFS_OK) fail(rc, "nsfs_plan_batch");
    stats_ = out.stats;
    return out;
}

BatchPlan GpuPlannerBase::planBatchSharded(DeviceGroup& group, const std::vector<int32_t>& sg4, MapData& map_data, int root) {
    ensureHandle(map_data);
    if (group.rank() == root) uploadMap(map_data);
    int rc = nsfs_group_broadcast_map(group.raw(), handle_, root);
    if (rc != NSFS_OK) fail(rc, "nsfs_group_broadcast_map");
    const int nq = (int)(sg4.size() / 4);
    BatchPlan out;
    out.status.assign((size_t)nq, 0); out.path_len.assign((size_t)nq, 0); out.g_goal.assign((size_t)nq, 0.0); out.settled.assign((size_t)nq, 0);
    if (!nq) return out;
    rc = nsfs_group_plan_batch(group.raw(), handle_, algoCode(), modeCode(cost_mode_), nq, sg4.data(), out.status.data(), out.g_goal.data(),
                               out.path_len.data(), out.settled.data(), &out.stats);
    if (rc != NSFS_OK) fail(rc, "nsfs_group_plan_batch");
    stats_ = out.stats;
    return out;
}

// ---- DeviceGroup ------------------------------------------------------------------------------------------
std::vector<unsigned char> DeviceGroup::makeId() {
    std::vector<unsigned char> id(128);
    if (nsfs_group_unique_id(id.data()) != NSFS_OK) throw std::runtime_error("nsfs_group_unique_id: NCCL (libnccl.so.2) is not available");
    return id;
}
DeviceGroup::DeviceGroup(const std::vector<unsigned char>& id, int rank, int world, int device) {
    if (id.size() != 128) throw std::invalid_argument("a group id is 128 bytes");
    const int rc = nsfs_group_create(&g_, id.data(), rank, world, device);
    if (rc != NSFS_OK) { g_ = nullptr; throw std::runtime_error(std::string("nsfs_group_create: ") + nsfs_strerror(rc)); }
}
DeviceGroup::~DeviceGroup() { if (g_) nsfs_group_destroy(g_); }
void DeviceGroup::barrier() {
    if (nsfs_group_barrier(g_) != NSFS_OK) throw std::runtime_error(std::string("nsfs_group_barrier: ") + nsfs_group_error(g_));
}

// ---- children ---------------------------------------------------------------------------------------------
std::vector<Pos2D> GpuAstar::plan(Index idx_start, Index idx_goal, MapData& map_data) { return runPlan(NSFS_ASTAR, idx_start, idx_goal, map_data); }
std::vector<Pos2D> GpuAstar::plan(Pos2D pos_start, Pos2D pos_goal, MapData& map_data) {
    return plan(toIndex(pos_start), toIndex(pos_goal), map_data);       // astar.cpp:19-25
}

std::vector<Pos2D> GpuDjikstra::plan(Index idx_start, Index idx_goal, MapData& map_data) {
    if (backend_ == Backend::Field) {
        try {
            return runPlan(NSFS_DJIKSTRA_FIELD, idx_start, idx_goal, map_data);
        } catch (const std::out_of_range&) {
            // a reachable cell's expansion throws in the reference; whether that happens before the goal is popped
            // depends on the pop order, so let the exact-order search decide (still on the GPU)
        }
    }
    return runPlan(NSFS_DJIKSTRA, idx_start, idx_goal, map_data);
}
std::vector<Pos2D> GpuDjikstra::plan(Pos2D pos_start, Pos2D pos_goal, MapData& map_data) {
    return plan(toIndex(pos_start), toIndex(pos_goal), map_data);       // djikstra.cpp:19-25
}

Pos2D GpuDjikstra::find_better_point(Index idx_bad, MapData& map_data) {
    uploadMap(map_data);
    if (idx_bad.i < 0 || idx_bad.i >= map_height_ || idx_bad.j < 0 || idx_bad.j >= map_width_)
        throw std::out_of_range("bad index outside the node store");
    int fi = idx_bad.i, fj = idx_bad.j;
    const int rc = nsfs_find_better_point(handle_, idx_bad.i, idx_bad.j, &fi, &fj, &stats_);
    if (rc != NSFS_OK) fail(rc, "nsfs_find_better_point");
    return toPos(fi, fj);
}
Pos2D GpuDjikstra::find_better_point(Pos2D pos_bad, MapData& map_data) {
    return find_better_point(toIndex(pos_bad), map_data);               // djikstra.cpp:275-280
}

std::vector<Pos2D> GpuThetaStar::plan(Index idx_start, Index idx_goal, MapData& map_data) { return runPlan(NSFS_THETASTAR, idx_start, idx_goal, map_data); }
std::vector<Pos2D> GpuThetaStar::plan(Pos2D pos_start, Pos2D pos_goal, MapData& map_data) {
    return plan(toIndex(pos_start), toIndex(pos_goal), map_data);
}

// ---- runtime selection --------------------------------------------------------------------------------------
static std::string trim(std::string s) {
    const char* ws = " \t\r\n";
    const size_t a = s.find_first_not_of(ws);
    if (a == std::string::npos) return "";
    const size_t b = s.find_last_not_of(ws);
    s = s.substr(a, b - a + 1);
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) s = s.substr(1, s.size() - 2);
    return s;
}

PlannerConfig load_planner_config(const std::string& yaml_path) {
    PlannerConfig cfg;
    std::ifstream in(yaml_path);
    if (!in) throw std::runtime_error("cannot open " + yaml_path);
    std::string line;
    while (std::getline(in, line)) {
        // strip a trailing comment that is not inside quotes
        bool q1 = false, q2 = false;
        for (size_t c = 0; c < line.size(); ++c) {
            if (line[c] == '"' && !q1) q2 = !q2;
            else if (line[c] == '\'' && !q2) q1 = !q1;
            else if (line[c] == '#' && !q1 && !q2) { line.resize(c); break; }
        }
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (val.empty()) continue;
        if (key == "planner_name") cfg.planner_name = val;
        else if (key == "cost_mode") cfg.cost_mode = val;
        else if (key == "djikstra_backend") cfg.djikstra_backend = val;
        else if (key == "graph") cfg.graph = val;
        else if (key == "gp_rate") cfg.gp_rate = std::atof(val.c_str());
        else if (key == "verbose_planner") cfg.verbose_planner = (val == "true" || val == "True" || val == "1");
        else if (key == "gpu_device") cfg.device = std::atoi(val.c_str());
    }
    return cfg;
}

std::shared_ptr<GridPlannerCore> make_main_planner(const PlannerConfig& cfg, MapData& map_data) {
    // GlobalPlanner::setupPlanner (global_planner.cpp:102-125): astar | djikstra | (thetastar) | anything else => astar
    if (cfg.planner_name == "djikstra") {
        if (cfg.cost_mode != "g") fprintf(stderr, "[GlobalPlanner]: Djikstra requires g cost. f/fg is set. Planner is guaranteed to fail\n");
        auto p = std::make_shared<GpuDjikstra>();
        p->setDevice(cfg.device);
        p->setTextbookGraph(cfg.graph == "textbook");
        p->setBackend(cfg.djikstra_backend == "field" ? GpuDjikstra::Backend::Field : GpuDjikstra::Backend::Exact);
        p->prepPlanner(cfg.cost_mode, map_data);
        return p;
    }
    if (cfg.planner_name == "thetastar") {     // the branch the reference left commented out (global_planner.cpp:116-119)
        auto p = std::make_shared<GpuThetaStar>();
        p->setDevice(cfg.device);
        p->setTextbookGraph(cfg.graph == "textbook");
        p->prepPlanner(cfg.cost_mode, map_data);
        return p;
    }
    if (cfg.cost_mode == "g") fprintf(stderr, "[GlobalPlanner]: Astar requires f/fg cost. g cost is set. Planner is guaranteed to fail\n");
    auto p = std::make_shared<GpuAstar>();
    p->setDevice(cfg.device);
    p->setTextbookGraph(cfg.graph == "textbook");
    p->prepPlanner(cfg.cost_mode, map_data);
    return p;
}

std::shared_ptr<GpuDjikstra> make_fallback_planner(const PlannerConfig& cfg, MapData& map_data) {
    auto p = std::make_shared<GpuDjikstra>();
    p->setDevice(cfg.device);
    p->prepPlanner("g", map_data);             // global_planner.cpp:127-128
    return p;
}

// ---- one plan request ---------------------------------------------------------------------------------------------
bool test_pos(Pos2D pos, MapData& map_data) {
    // GlobalPlanner::testPos (global_planner.cpp:384-408) with its own pos2idx / flatten / oob (360-376): the key is formed
    // before the row-only bounds test, and vector::at throws for a key outside the grids
    const int i = (int)std::round((pos.x - map_data.origin_.x) / map_data.cell_size_);
    const int j = (int)std::round((pos.y - map_data.origin_.y) / map_data.cell_size_);
    const int key = i * map_data.map_size_.j + j;
    if (i < 0 || i >= map_data.map_size_.i) return false;
    if (map_data.grid_inflation_.at((size_t)key) > 0) return false;
    if (map_data.grid_logodds_.at((size_t)key) > map_data.lo_thresh_) return false;
    return true;
}

PlanRequestResult serve_plan_request(GridPlannerCore& main_planner, GpuDjikstra& fallback_planner, Pos2D robot_position, Pos2D current_goal,
                                     MapData& map_data) {
    PlanRequestResult r;
    r.robot_ok = test_pos(robot_position, map_data);
    r.goal_ok = test_pos(current_goal, map_data);
    r.goal = current_goal;
    r.backup_robot = robot_position;
    Pos2D start = robot_position;
    if (r.robot_ok && r.goal_ok) {
        r.which_case = 1;                                                       // global_planner.cpp:179-185
    } else if (r.robot_ok && !r.goal_ok) {
        r.which_case = 2;                                                       // 188-214
        r.goal = fallback_planner.find_better_point(current_goal, map_data);
        r.goal_updated = true;
    } else if (!r.robot_ok && r.goal_ok) {
        r.which_case = 3;                                                       // 217-236
        r.backup_robot = fallback_planner.find_better_point(robot_position, map_data);
        r.backup_mode = true;
        start = r.backup_robot;
    } else {
        r.which_case = 4;                                                       // 238-273: robot first, then the goal
        r.backup_robot = fallback_planner.find_better_point(robot_position, map_data);
        r.goal = fallback_planner.find_better_point(current_goal, map_data);
        r.goal_updated = true;
        start = r.backup_robot;                                                 // (upstream leaves backup_mode_ as it was here)
    }
    r.path = main_planner.generatePath(start, r.goal, map_data);
    return r;
}

}  // namespace nsfs_host

// ---- C shim: lets the Python tests drive the C++ children exactly as GlobalPlanner::run would -----------------
// (global_planner.cpp:183/212/234/271 generatePath, 191/225/244/247 find_better_point).  Test plumbing for the
// adapter layer; the product boundary is include/nsfs.h.
namespace {
struct Session {
    bot_utils::MapData map;
    nsfs_host::PlannerConfig cfg;
    std::shared_ptr<GridPlannerCore> main_planner;
    std::shared_ptr<nsfs_host::GpuDjikstra> fallback;
    std::string error;
};
// public generatePath is reachable through the base; plan() is protected, so raw plans go through this accessor
template <class P> struct Raw : P { using P::plan; };
}  // namespace

extern "C" {

// returns 0 ok, 1 std::out_of_range, 2 any other exception (text via nsfsh_error)
void* nsfsh_create(const char* planner_name, const char* cost_mode, const char* djikstra_backend, int H, int W, double cell,
                   double ox, double oy, int lo_thresh, const int32_t* infl, const int32_t* lo, int device, int* rc_out) {
    auto* s = new Session();
    int rc = 0;
    try {
        s->cfg.planner_name = planner_name ? planner_name : "astar";
        s->cfg.cost_mode = cost_mode ? cost_mode : "fg";
        s->cfg.djikstra_backend = djikstra_backend ? djikstra_backend : "exact";
        s->cfg.device = device;
        s->map.map_size_.setIdx(H, W);
        s->map.cell_size_ = cell;
        s->map.origin_.setCoords(ox, oy);
        s->map.lo_thresh_ = lo_thresh;
        s->map.total_cells_ = H * W;
        s->map.grid_inflation_.assign(infl, infl + (size_t)H * W);
        s->map.grid_logodds_.assign(lo, lo + (size_t)H * W);
        s->main_planner = nsfs_host::make_main_planner(s->cfg, s->map);
        s->fallback = nsfs_host::make_fallback_planner(s->cfg, s->map);
    } catch (const std::out_of_range& e) { s->error = e.what(); rc = 1; }
    catch (const std::exception& e) { s->error = e.what(); rc = 2; }
    if (rc_out) *rc_out = rc;
    return s;
}

void nsfsh_destroy(void* p) { delete static_cast<Session*>(p); }
const char* nsfsh_error(void* p) { return p ? static_cast<Session*>(p)->error.c_str() : ""; }

int nsfsh_set_map(void* p, const int32_t* infl, const int32_t* lo) {
    auto* s = static_cast<Session*>(p);
    const size_t n = (size_t)s->map.map_size_.i * s->map.map_size_.j;
    s->map.grid_inflation_.assign(infl, infl + n);      // what the ROS callbacks do (global_planner.cpp:65-73)
    s->map.grid_logodds_.assign(lo, lo + n);
    return 0;
}

// 0 = EveryCall (pageable copy of the caller's vectors on every plan: the reference's contract), 1 = OnNotify, 2 = EveryCallPinned
int nsfsh_set_map_sync(void* p, int mode) {
    auto* s = static_cast<Session*>(p);
    const nsfs_host::MapSync m = mode == 1 ? nsfs_host::MapSync::OnNotify : (mode == 2 ? nsfs_host::MapSync::EveryCallPinned : nsfs_host::MapSync::EveryCall);
    if (auto* g = dynamic_cast<nsfs_host::GpuPlannerBase*>(s->main_planner.get())) g->setMapSync(m);
    s->fallback->setMapSync(m);
    return 0;
}

int nsfsh_notify_map_changed(void* p) {
    auto* s = static_cast<Session*>(p);
    if (auto* g = dynamic_cast<nsfs_host::GpuPlannerBase*>(s->main_planner.get())) g->notifyMapChanged();
    s->fallback->notifyMapChanged();
    return 0;
}

// nq plan(Index, Index) calls of the session's main planner on one upload of its MapData (BASELINE config C4)
int nsfsh_plan_batch_idx(void* p, int nq, const int32_t* sg4, int32_t* status, double* g_goal, int32_t* path_len, long long* settled,
                         double* gpu_ms) {
    auto* s = static_cast<Session*>(p);
    try {
        auto* g = dynamic_cast<nsfs_host::GpuPlannerBase*>(s->main_planner.get());
        if (!g) return 2;
        const std::vector<int32_t> q(sg4, sg4 + 4 * (size_t)nq);
        const nsfs_host::BatchPlan r = g->planBatch(q, s->map);
        memcpy(status, r.status.data(), (size_t)nq * 4); memcpy(path_len, r.path_len.data(), (size_t)nq * 4);
        memcpy(g_goal, r.g_goal.data(), (size_t)nq * 8); memcpy(settled, r.settled.data(), (size_t)nq * 8);
        if (gpu_ms) *gpu_ms = r.stats.gpu_ms;
        return 0;
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
}

// A DeviceGroup for the session (rank 0 made the id with nsfsh_group_id and shipped it to the other ranks)
int nsfsh_group_id(unsigned char* id128) {
    try { const auto id = nsfs_host::DeviceGroup::makeId(); memcpy(id128, id.data(), 128); return 0; }
    catch (const std::exception&) { return 2; }
}
void* nsfsh_group_create(const unsigned char* id128, int rank, int world, int device) {
    try { return new nsfs_host::DeviceGroup(std::vector<unsigned char>(id128, id128 + 128), rank, world, device); }
    catch (const std::exception&) { return nullptr; }
}
void nsfsh_group_destroy(void* g) { delete static_cast<nsfs_host::DeviceGroup*>(g); }

// GpuPlannerBase::planBatchSharded: all nq queries in, all nq results out on every rank
int nsfsh_plan_batch_sharded(void* p, void* group, int nq, const int32_t* sg4, int32_t* status, double* g_goal, int32_t* path_len,
                             long long* settled, double* gpu_ms) {
    auto* s = static_cast<Session*>(p);
    try {
        auto* g = dynamic_cast<nsfs_host::GpuPlannerBase*>(s->main_planner.get());
        if (!g || !group) return 2;
        const std::vector<int32_t> q(sg4, sg4 + 4 * (size_t)nq);
        const nsfs_host::BatchPlan r = g->planBatchSharded(*static_cast<nsfs_host::DeviceGroup*>(group), q, s->map);
        memcpy(status, r.status.data(), (size_t)nq * 4); memcpy(path_len, r.path_len.data(), (size_t)nq * 4);
        memcpy(g_goal, r.g_goal.data(), (size_t)nq * 8); memcpy(settled, r.settled.data(), (size_t)nq * 8);
        if (gpu_ms) *gpu_ms = r.stats.gpu_ms;
        return 0;
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
}

static int emit(const std::vector<bot_utils::Pos2D>& path, double* out_xy, int cap, int* n) {
    if (n) *n = (int)path.size();
    for (size_t t = 0; t < path.size() && (int)t < cap; ++t) { out_xy[2 * t] = path[t].x; out_xy[2 * t + 1] = path[t].y; }
    return 0;
}

int nsfsh_generate_path_idx(void* p, int si, int sj, int gi, int gj, double* out_xy, int cap, int* n) {
    auto* s = static_cast<Session*>(p);
    try {
        bot_utils::Index a(si, sj), b(gi, gj);
        return emit(s->main_planner->generatePath(a, b, s->map), out_xy, cap, n);
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
}

int nsfsh_generate_path_pos(void* p, double sx, double sy, double gx, double gy, double* out_xy, int cap, int* n) {
    auto* s = static_cast<Session*>(p);
    try {
        bot_utils::Pos2D a(sx, sy), b(gx, gy);
        return emit(s->main_planner->generatePath(a, b, s->map), out_xy, cap, n);
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
}

int nsfsh_find_better_point_pos(void* p, double x, double y, double* fx, double* fy) {
    auto* s = static_cast<Session*>(p);
    try {
        const bot_utils::Pos2D r = s->fallback->find_better_point(bot_utils::Pos2D(x, y), s->map);
        *fx = r.x; *fy = r.y;
        return 0;
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
}

int nsfsh_find_better_point_idx(void* p, int bi, int bj, double* fx, double* fy) {
    auto* s = static_cast<Session*>(p);
    try {
        const bot_utils::Pos2D r = s->fallback->find_better_point(bot_utils::Index(bi, bj), s->map);
        *fx = r.x; *fy = r.y;
        return 0;
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
}

// serve_plan_request through the session's planners.  info[0..3] = case, goal_updated, backup_mode, robot_ok | goal_ok << 1
int nsfsh_plan_request(void* p, double rx, double ry, double gx, double gy, double* out_xy, int cap, int* n, int* info, double* goal_xy,
                       double* backup_xy) {
    auto* s = static_cast<Session*>(p);
    try {
        const nsfs_host::PlanRequestResult r = nsfs_host::serve_plan_request(*s->main_planner, *s->fallback, bot_utils::Pos2D(rx, ry),
                                                                             bot_utils::Pos2D(gx, gy), s->map);
        info[0] = r.which_case; info[1] = r.goal_updated; info[2] = r.backup_mode; info[3] = (int)r.robot_ok | ((int)r.goal_ok << 1);
        goal_xy[0] = r.goal.x; goal_xy[1] = r.goal.y; backup_xy[0] = r.backup_robot.x; backup_xy[1] = r.backup_robot.y;
        return emit(r.path, out_xy, cap, n);
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
}

// Parses a rosparam-style YAML; strings are copied into 32-byte buffers.
// The opt-in textbook graph for the session's main planner (the fallback keeps the reference's graph).
int nsfsh_set_graph(void* p, int textbook) {
    auto* s = static_cast<Session*>(p);
    try {
        auto* g = dynamic_cast<nsfs_host::GpuPlannerBase*>(s->main_planner.get());
        if (!g) return 2;
        g->setTextbookGraph(textbook != 0);
    } catch (const std::exception& e) { s->error = e.what(); return 2; }
    return 0;
}

// Commander::checkTrajectorySafety through the session's fallback planner handle (same map).
int nsfsh_check_trajectory(void* p, int n, const double* xy, int t_id, int* safe, int* bad_idx) {
    auto* s = static_cast<Session*>(p);
    try {
        std::vector<bot_utils::Pos2D> traj;
        for (int t = 0; t < n; ++t) traj.emplace_back(xy[2 * t], xy[2 * t + 1]);
        const auto r = s->fallback->checkTrajectorySafety(traj, t_id, s->map);
        *safe = r.first ? 1 : 0;
        *bad_idx = r.second;
    } catch (const std::out_of_range& e) { s->error = e.what(); return 1; }
    catch (const std::exception& e) { s->error = e.what(); return 2; }
    return 0;
}

int nsfsh_load_config(const char* yaml_path, char* planner_name, char* cost_mode, char* djikstra_backend, double* gp_rate,
                      int* verbose, int* device) {
    try {
        const nsfs_host::PlannerConfig c = nsfs_host::load_planner_config(yaml_path);
        snprintf(planner_name, 32, "%s", c.planner_name.c_str());
        snprintf(cost_mode, 32, "%s", c.cost_mode.c_str());
        snprintf(djikstra_backend, 32, "%s", c.djikstra_backend.c_str());
        *gp_rate = c.gp_rate; *verbose = c.verbose_planner ? 1 : 0; *device = c.device;
        return 0;
    } catch (const std::exception&) { return 2; }
}

// Which child did setupPlanner's rule pick?  0 astar, 1 djikstra, 2 thetastar
int nsfsh_main_planner_kind(void* p) {
    auto* s = static_cast<Session*>(p);
    if (dynamic_cast<nsfs_host::GpuDjikstra*>(s->main_planner.get())) return 1;
    if (dynamic_cast<nsfs_host::GpuThetaStar*>(s->main_planner.get())) return 2;
    return 0;
}

// sparsifyPath alone (host arithmetic; no device needed): lets the CPU suite check it against the golden finals' inputs
int nsfsh_sparsify(const double* xy_in, int n_in, double* xy_out, int cap) {
    struct Access : nsfs_host::GpuAstar { using nsfs_host::GpuPlannerBase::sparsify; };
    std::vector<bot_utils::Pos2D> in;
    for (int t = 0; t < n_in; ++t) in.emplace_back(xy_in[2 * t], xy_in[2 * t + 1]);
    const auto out = Access::sparsify(in);
    for (size_t t = 0; t < out.size() && (int)t < cap; ++t) { xy_out[2 * t] = out[t].x; xy_out[2 * t + 1] = out[t].y; }
    return (int)out.size();
}

}  // extern "C"
