// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// C-ABI harness around the UNMODIFIED reference planner sources.  oracle/Makefile compiles this file
// together with the reference's own translation units where they lie under /root/reference:
//     src/bot_utils/src/bot_utils.cpp
//     src/global_planner/src/algo/{grid_planner_core,astar,djikstra,thetastar}.cpp
// against oracle/stub/ (logging-only ros.h) with -fno-access-control so the harness can reach the
// protected plan()/post_process_path()/nodes/open_list_ members.  Output: oracle/_ref/libnsfs_ref.so.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// What each entry point drives (reference file:line):
//   nsref_plan            Astar::plan astar.cpp:27-148 / Djikstra::plan djikstra.cpp:27-145,
//                         optionally followed by post_process_path (astar.cpp:150-165, djikstra.cpp:147-163)
//                         exactly as GridPlannerCore::generatePath does (grid_planner_core.cpp:603-608)
//   nsref_field           Djikstra::plan with an unreachable goal (-7,-7) => full cost-to-go field
//   nsref_find_better_point  Djikstra::find_better_point djikstra.cpp:165-273
//   nsref_bresenham / nsref_is_los   bot_utils.cpp:137-192, grid_planner_core.cpp:481-496
//   nsref_sparsify / nsref_any_angle grid_planner_core.cpp:506-558
//   nsref_heap_trace      GridPlannerCore::OpenList grid_planner_core.cpp:55-122,286-306
//
// One deliberate harness intervention: OpenList::clearQ() (grid_planner_core.cpp:335-338) does not
// reset size_, so a planner object that found its goal with entries left in the queue carries a stale
// positive size_ into the next call; if that next search exhausts the queue the reference pops an
// empty std::priority_queue (undefined behaviour).  The harness zeroes open_list_.size_ before every
// call so each query behaves like the first query on a fresh planner.  DESIGN.md records this.

#include "astar.h"
#include "djikstra.h"
#include "thetastar.h"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <stdexcept>
#include <thread>
#include <vector>

using bot_utils::Index;
using bot_utils::MapData;
using bot_utils::Pos2D;

namespace {

const char* kModes[3] = {"f", "g", "fg"};

struct Ctx {
    MapData map;
    // planners[algo][mode], built on first use (each holds H*W 40-byte nodes)
    std::unique_ptr<Astar> astar[3];
    std::unique_ptr<Djikstra> djik[3];
};

GridPlannerCore* planner(Ctx* c, int algo, int mode) {
    if (mode < 0 || mode > 2) return nullptr;
    if (algo == 0) {
        if (!c->astar[mode]) c->astar[mode] = std::make_unique<Astar>(std::string(kModes[mode]), c->map);
        return c->astar[mode].get();
    }
    if (algo == 1) {
        if (!c->djik[mode]) c->djik[mode] = std::make_unique<Djikstra>(std::string(kModes[mode]), c->map);
        return c->djik[mode].get();
    }
    return nullptr;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

long long count_visited(GridPlannerCore* p) {
    long long n = 0;
    for (auto& nd : p->nodes) n += nd.visited ? 1 : 0;
    return n;
}

int emit_path(GridPlannerCore* p, const std::vector<Pos2D>& path, double* xy, int32_t* ij, int cap, int* n) {
    *n = (int)path.size();
    int m = (int)path.size() < cap ? (int)path.size() : cap;
    for (int t = 0; t < m; ++t) {
        if (xy) { xy[2 * t] = path[t].x; xy[2 * t + 1] = path[t].y; }
        if (ij) {
            Pos2D q = path[t];
            Index id = p->pos2idx(q);
            ij[2 * t] = id.i; ij[2 * t + 1] = id.j;
        }
    }
    return (int)path.size() <= cap ? 0 : -3;
}

}  // namespace

extern "C" {

typedef struct Ctx nsref_t;

nsref_t* nsref_create(int H, int W, double cell, double ox, double oy,
                      const int32_t* infl, const int32_t* lo, int lo_thresh) {
    Ctx* c = new Ctx();
    c->map.grid_inflation_.assign(infl, infl + (size_t)H * W);
    c->map.grid_logodds_.assign(lo, lo + (size_t)H * W);
    c->map.lo_thresh_ = lo_thresh;
    c->map.lo_cap_ = 100;
    c->map.origin_.setCoords(ox, oy);
    c->map.pos_min_.setCoords(ox, oy);
    c->map.pos_max_.setCoords(ox + H * cell, oy + W * cell);
    c->map.map_size_.setIdx(H, W);
    c->map.cell_size_ = cell;
    c->map.total_cells_ = H * W;
    return c;
}

void nsref_set_map(nsref_t* c, const int32_t* infl, const int32_t* lo) {
    size_t n = (size_t)c->map.map_size_.i * c->map.map_size_.j;
    c->map.grid_inflation_.assign(infl, infl + n);
    c->map.grid_logodds_.assign(lo, lo + n);
}

void nsref_destroy(nsref_t* c) { delete c; }

// status: 0 ok | -1 std::out_of_range | -2 other exception | -3 output truncated | -4 bad args
// post = 0 -> raw plan(); post = 1 -> plan() then post_process_path() (== generatePath)
int nsref_plan(nsref_t* c, int algo, int mode, int si, int sj, int gi, int gj, int post,
               double* xy, int32_t* ij, int cap, int* n, double* g_goal, long long* settled,
               double* seconds) {
    GridPlannerCore* p = planner(c, algo, mode);
    if (!p) return -4;
    p->open_list_.size_ = 0;  // see header comment
    p->open_list_.clearQ();
    Index s(si, sj), g(gi, gj);
    try {
        double t0 = now_s();
        std::vector<Pos2D> path = p->plan(s, g, c->map);
        if (post) path = p->post_process_path(path, c->map);
        double t1 = now_s();
        if (seconds) *seconds = t1 - t0;
        if (g_goal) {
            bool inside = gi >= 0 && gi < c->map.map_size_.i && gj >= 0 && gj < c->map.map_size_.j;
            *g_goal = inside ? p->nodes[(size_t)gi * c->map.map_size_.j + gj].g : 1e5;
        }
        if (settled) *settled = count_visited(p);
        return emit_path(p, path, xy, ij, cap, n);
    } catch (const std::out_of_range&) {
        return -1;
    } catch (...) {
        return -2;
    }
}

// Full Djikstra field: g (double), parent key (-1 if none) and visited for every cell.
int nsref_field(nsref_t* c, int mode, int si, int sj, double* g, int32_t* parent, uint8_t* visited,
                long long* settled, double* seconds) {
    GridPlannerCore* p = planner(c, 1, mode);
    if (!p) return -4;
    p->open_list_.size_ = 0;
    p->open_list_.clearQ();
    Index s(si, sj), goal(-7, -7);
    try {
        double t0 = now_s();
        (void)p->plan(s, goal, c->map);
        double t1 = now_s();
        if (seconds) *seconds = t1 - t0;
    } catch (const std::out_of_range&) {
        return -1;
    } catch (...) {
        return -2;
    }
    const int W = c->map.map_size_.j;
    size_t N = p->nodes.size();
    long long cnt = 0;
    for (size_t k = 0; k < N; ++k) {
        auto& nd = p->nodes[k];
        if (g) g[k] = nd.g;
        if (visited) visited[k] = nd.visited ? 1 : 0;
        if (parent) parent[k] = (nd.g < 1e5 && nd.parent.i >= 0 && (int)k != si * W + sj) ? nd.parent.i * W + nd.parent.j : -1;
        cnt += nd.visited ? 1 : 0;
    }
    if (settled) *settled = cnt;
    return 0;
}

int nsref_find_better_point(nsref_t* c, int bi, int bj, int* fi, int* fj, double* x, double* y,
                            long long* settled, double* seconds) {
    Djikstra* p = static_cast<Djikstra*>(planner(c, 1, 1));  // fallback planner runs in "g" mode (global_planner.cpp:127-128)
    p->open_list_.size_ = 0;
    p->open_list_.clearQ();
    Index bad(bi, bj);
    try {
        double t0 = now_s();
        Pos2D r = p->find_better_point(bad, c->map);
        double t1 = now_s();
        if (seconds) *seconds = t1 - t0;
        Index id = p->pos2idx(r);
        if (fi) *fi = id.i;
        if (fj) *fj = id.j;
        if (x) *x = r.x;
        if (y) *y = r.y;
        if (settled) *settled = count_visited(p);
        return 0;
    } catch (const std::out_of_range&) {
        return -1;
    } catch (...) {
        return -2;
    }
}

int nsref_bresenham(int si, int sj, int ti, int tj, int32_t* out_ij, int cap) {
    Index s(si, sj), t(ti, tj);
    std::vector<Index> ray = bot_utils::bresenham_los(s, t);
    int m = (int)ray.size() < cap ? (int)ray.size() : cap;
    for (int k = 0; k < m; ++k) { out_ij[2 * k] = ray[k].i; out_ij[2 * k + 1] = ray[k].j; }
    return (int)ray.size();
}

// 1 = line of sight, 0 = blocked, -1 = std::out_of_range
int nsref_is_los(nsref_t* c, int si, int sj, int ti, int tj) {
    GridPlannerCore* p = planner(c, 0, 0);
    Index s(si, sj), t(ti, tj);
    try { return p->isLos(s, t, c->map) ? 1 : 0; } catch (const std::out_of_range&) { return -1; }
}

// 1 = free, 0 = not free / out of the row range, -1 = std::out_of_range
int nsref_test_pos(nsref_t* c, int i, int j) {
    GridPlannerCore* p = planner(c, 0, 0);
    Index id(i, j);
    try { return p->testPos(id, c->map) ? 1 : 0; } catch (const std::out_of_range&) { return -1; }
}

double nsref_dist_oct(int si, int sj, int ti, int tj) { return bot_utils::dist_oct(Index(si, sj), Index(ti, tj)); }
double nsref_dist_euc(int si, int sj, int ti, int tj) { return bot_utils::dist_euc(Index(si, sj), Index(ti, tj)); }

void nsref_idx2pos(nsref_t* c, int i, int j, double* x, double* y) {
    GridPlannerCore* p = planner(c, 0, 0);
    Index id(i, j);
    Pos2D q = p->idx2pos(id);
    *x = q.x; *y = q.y;
}

void nsref_pos2idx(nsref_t* c, double x, double y, int* i, int* j) {
    GridPlannerCore* p = planner(c, 0, 0);
    Pos2D q(x, y);
    Index id = p->pos2idx(q);
    *i = id.i; *j = id.j;
}

static std::vector<Pos2D> to_path(const double* xy, int n) {
    std::vector<Pos2D> v;
    v.reserve(n);
    for (int t = 0; t < n; ++t) v.emplace_back(xy[2 * t], xy[2 * t + 1]);
    return v;
}

int nsref_sparsify(nsref_t* c, const double* xy_in, int n_in, double* xy_out, int cap) {
    GridPlannerCore* p = planner(c, 0, 0);
    auto in = to_path(xy_in, n_in);
    auto out = p->sparsifyPath(in);
    int n = 0;
    emit_path(p, out, xy_out, nullptr, cap, &n);
    return n;
}

int nsref_any_angle(nsref_t* c, const double* xy_in, int n_in, double* xy_out, int cap) {
    GridPlannerCore* p = planner(c, 0, 0);
    auto in = to_path(xy_in, n_in);
    try {
        auto out = p->makeAnyAnglePath(in, c->map);
        int n = 0;
        emit_path(p, out, xy_out, nullptr, cap, &n);
        return n;
    } catch (const std::out_of_range&) { return -1; }
}

int nsref_post_process(nsref_t* c, int algo, const double* xy_in, int n_in, double* xy_out, int cap) {
    GridPlannerCore* p = planner(c, algo, algo == 1 ? 1 : 0);
    auto in = to_path(xy_in, n_in);
    try {
        auto out = p->post_process_path(in, c->map);
        int n = 0;
        emit_path(p, out, xy_out, nullptr, cap, &n);
        return n;
    } catch (const std::out_of_range&) { return -1; }
}

// Drive the reference OpenList directly.  ops[4*t..] = {op, f, g, id}: op 1 = push OpenNode(f, g, id, 0),
// op 0 = pop (its id is appended to out).  Returns the number of pops.
int nsref_heap_trace(int mode, int n_ops, const int32_t* ops, int32_t* out) {
    std::string m(kModes[mode]);
    GridPlannerCore::OpenList ol;
    ol.setMode(m);
    int np = 0;
    for (int t = 0; t < n_ops; ++t) {
        const int32_t* o = ops + 4 * t;
        if (o[0] == 1) {
            ol.pushNode(GridPlannerCore::OpenNode(o[1], o[2], o[3], 0));
        } else if (ol.getSizeQ() > 0) {
            GridPlannerCore::OpenNode nd = ol.popNode();
            out[np++] = nd.idx.i;
        }
    }
    return np;
}

// Batch of independent plan() queries over `nthreads` host threads, one planner object per thread
// ("one planner instance per host core over disjoint query shards").
// sg4[4*q..] = {si, sj, gi, gj}.  Outputs per query: g at goal, path length (points), settled cells.
int nsref_plan_batch(nsref_t* c, int algo, int mode, int nq, const int32_t* sg4, int nthreads,
                     double* g_goal, int32_t* path_len, long long* settled, double* seconds) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = nq > 0 ? nq : 1;
    std::vector<int> status(nthreads, 0);
    const int W = c->map.map_size_.j;
    // construction is excluded from the timed region
    std::vector<std::unique_ptr<GridPlannerCore>> ps(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        if (algo == 0) ps[t] = std::make_unique<Astar>(std::string(kModes[mode]), c->map);
        else ps[t] = std::make_unique<Djikstra>(std::string(kModes[mode]), c->map);
    }
    double t0 = now_s();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            GridPlannerCore* p = ps[t].get();
            for (int q = t; q < nq; q += nthreads) {
                p->open_list_.size_ = 0;
                p->open_list_.clearQ();
                Index s(sg4[4 * q], sg4[4 * q + 1]), g(sg4[4 * q + 2], sg4[4 * q + 3]);
                try {
                    auto path = p->plan(s, g, c->map);
                    if (g_goal) g_goal[q] = p->nodes[(size_t)g.i * W + g.j].g;
                    if (path_len) path_len[q] = (int)path.size();
                    if (settled) settled[q] = count_visited(p);
                } catch (...) {
                    status[t] = -1;
                    if (g_goal) g_goal[q] = -1;
                    if (path_len) path_len[q] = -1;
                    if (settled) settled[q] = -1;
                }
            }
        });
    }
    for (auto& x : th) x.join();
    if (seconds) *seconds = now_s() - t0;
    for (int s : status) if (s) return s;
    return 0;
}

// The map generator's random stream as SURVEY.md §8d specifies it (std::mt19937 + uniform_real_distribution),
// so tests can pin nsfs/mapgen.py (numpy) to the C++ definition.
void nsref_uniform01(unsigned seed, int n, double* out) {
    std::mt19937 g(seed);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    for (int t = 0; t < n; ++t) out[t] = u(g);
}
void nsref_mt_raw(unsigned seed, int n, uint32_t* out) {
    std::mt19937 g(seed);
    for (int t = 0; t < n; ++t) out[t] = (uint32_t)g();
}

// n_src full Djikstra fields (plan() with an unreachable goal) spread over `nthreads` host threads, one planner
// object per thread; used by bench.py --impl reference to load every host core with the reference's own code.
// settled[s] = visited cells of field s.  Construction is outside the timed region.
int nsref_field_batch(nsref_t* c, int n_src, const int32_t* src_ij, int nthreads, long long* settled, double* seconds) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n_src) nthreads = n_src > 0 ? n_src : 1;
    std::vector<std::unique_ptr<Djikstra>> ps(nthreads);
    for (int t = 0; t < nthreads; ++t) ps[t] = std::make_unique<Djikstra>(std::string("g"), c->map);
    std::vector<int> status(nthreads, 0);
    double t0 = now_s();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            Djikstra* p = ps[t].get();
            for (int s = t; s < n_src; s += nthreads) {
                p->open_list_.size_ = 0;
                p->open_list_.clearQ();
                Index a(src_ij[2 * s], src_ij[2 * s + 1]), goal(-7, -7);
                try {
                    (void)p->plan(a, goal, c->map);
                    if (settled) settled[s] = count_visited(p);
                } catch (...) {
                    status[t] = -1;
                    if (settled) settled[s] = -1;
                }
            }
        });
    }
    for (auto& x : th) x.join();
    if (seconds) *seconds = now_s() - t0;
    for (int s : status) if (s) return s;
    return 0;
}

int nsref_sizeof_node() { return (int)sizeof(GridPlannerCore::Node); }

}  // extern "C"
