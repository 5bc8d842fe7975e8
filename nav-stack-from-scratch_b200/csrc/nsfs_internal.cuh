// Internal declarations shared by the sm_100a translation units of libnsfs_gpu.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "nsfs.h"

namespace nsfs {

// ---- reference constants ---------------------------------------------------------------------------
// NB_LUT order (reference grid_planner_core.h:197): d -> (DI, DJ); flat offset OFF[d] = DI*W + DJ.
// Packed 2-bit tables (value + 1) so a run-time d costs a shift and a mask and a constant d folds away.
__host__ __device__ __forceinline__ int nb_di(int d) { return (int)((0x901Au >> (2 * d)) & 3u) - 1; }   // {1,1,0,-1,-1,-1,0,1}
__host__ __device__ __forceinline__ int nb_dj(int d) { return (int)((0x01A9u >> (2 * d)) & 3u) - 1; }   // {0,1,1,1,0,-1,-1,-1}
__host__ __device__ __forceinline__ long long nb_off(int d, int W) { return (long long)nb_di(d) * W + nb_dj(d); }

constexpr float  kInfF = 1e5f;                 // "infinity" of the reference (astar.cpp:40)
constexpr double kInfD = 1e5;
constexpr double kSlack = 1e-5;                // relaxation slack (astar.cpp:116)
constexpr double kSqrt2D = 1.41421356237309504880;
constexpr float  kSqrt2F = 1.41421356237309504880f;

// ---- per-cell edge masks ---------------------------------------------------------------------------
// out_mask[u] (uint16): bits 0..7  pass[d]  : neighbour u+OFF[d] is reached by plan()'s loop (row guard ok,
//                                             key in range, free)                    (astar.cpp:89-101)
//                       bits 9..15 diag[d]  : that edge costs sqrt(2) (odd number of passable directions
//                                             before d; is_cardinal toggle, astar.cpp:88,107,123); d = 0 is
//                                             never diagonal, so
//                       bit  8     throws   : some direction passes the row guard with a key outside [0,N)
//                                             => vector::at throws when u is expanded
// in_mask[v]  (uint16): bits 0..7  in-edge d exists from the FREE cell v-OFF[d]; bits 8..15 it costs sqrt(2)
constexpr uint32_t kThrowBit = 1u << 8;

// ---- cost-to-go field (field.cu): Q17.15 fixed point ---------------------------------------------------------------------
// g is kept as uint32 in units of 2^-15: the edge weights 1 and sqrt(2) become 32768 and 46341 (= round(sqrt(2) * 2^15); relative
// error 1.08e-6), the reference's "infinity" 1e5 (astar.cpp:40) becomes kInfQ = 1e5 * 2^15 < 2^32.  Integer sums are exact and
// order independent, so every scheduler / sharding returns the same bits and the error against the reference's fp64 sums stays
// at 1.08e-6 relative for any path length (an fp32 plane drifts past 1e-5 beyond ~3000 steps).
constexpr uint32_t kQOne = 32768u;
constexpr uint32_t kQDiag = 46341u;
constexpr uint32_t kInfQ = 3276800000u;      // 1e5 in Q17.15: exactly representable as fp32, converts back to 1e5f
constexpr uint32_t kNoKey = 0xffffffffu;     // "no pending update" (every value and key is <= kInfQ + kQDiag)
constexpr int kFieldTile = 64;               // tile side of the field solver

// Parameters of the field solve kernel.
struct FieldParams {
    unsigned long long* prof;   // visit trace when built with NSFS_FIELD_TRACE (dev/trace_field.py), else nullptr
    uint32_t* g;                // [N] Q17.15
    const uint16_t* in_mask;
    uint32_t* tile_key;   // per tile: smallest value a neighbour wrote into this tile's halo since it was last relaxed
    int* counts;          // 1: pending tiles + visits in flight; 3: tile visits; 4: error flags; 5: reached; 6: smallest pending key
    int H, W;
    long long N;
    int tiles_x, tiles_y;
    unsigned idle_ns;     // longest back-off of a warp that has nothing to relax
    unsigned wait_ns;     // ... and the shortest, right after it had work
    int inner;            // warp-local relaxation passes of a due sub-block before it yields
    uint32_t delta;       // width (Q17.15) of the cost band [key, key + delta] a visit may advance into (delta-stepping over tiles)
    // peer-to-peer shards (NVLink): side 0 = the shard above (smaller rows), 1 = below.  A shard pushes every improvement
    // of its first / last two OWNED rows straight into the neighbour's ghost rows and wakes the neighbour's tiles.
    int* gcount;          // pending tiles + visits in flight: counts + 1, or ONE counter shared by all shards (on the owner's GPU)
    uint32_t* peer_g[2];  // the neighbour's g plane (its local rows), nullptr = no neighbour on that side
    uint32_t* peer_key[2];
    int peer_row_off[2];  // my local row r is the neighbour's local row r + peer_row_off[side]
    int peer_tiles_y[2];
    int peer_rows[2];     // rows of the neighbour's plane
    int own_lo, own_hi;   // my owned local rows [own_lo, own_hi); the others are ghost rows (or inactive map rows)
    unsigned sched_ns;    // back-off of a block that owns no pending tile
    int merge;            // 1: a visit that is about to end first absorbs posts that arrived for its own tile meanwhile
    int light;            // 1: key posts are release atomics (MEMBAR.ALL.GPU) instead of __threadfence (MEMBAR.SC.GPU + L1 invalidate) + atomicMin
    uint32_t* blk_key;    // [workers] global gate: the smallest key each worker block holds (pending or in flight)
    uint32_t* gfloor;     // ... and their minimum, kept by the grid's last block
    uint32_t gate;        // a worker leaves a tile alone while its key exceeds *gfloor + gate (Q17.15; 0 = no gate)
    uint32_t cap;         // tiles whose pending key exceeds it stay pending and nothing is relaxed beyond it (kNoKey - 1 = no cap).
                          // Used by the multi-GPU solve to advance all shards band by band.
};

struct FieldWork {
    uint32_t* g = nullptr;         // [N] Q17.15 during a solve; fp32 cost units (same buffer) once nsfs_field_finish has run
    uint8_t*  pred = nullptr;      // [N]
    uint32_t* tile_key = nullptr;  // [nTiles] smallest pending halo update per tile, kNoKey = none
    int*      counts = nullptr;    // [16]
    uint32_t* blk_key = nullptr;   // global gate: per-block keys + the floor word
    int       tiles_x = 0, tiles_y = 0;
    int       src_i = -1, src_j = -1;
    bool      valid = false;       // solved and finished: nsfs_field_path may walk pred
    // peer-to-peer shards: neighbours' buffers opened through CUDA IPC (other processes) and the shared pending counter
    float*    peer_g[2] = {nullptr, nullptr};
    uint32_t* peer_key[2] = {nullptr, nullptr};
    int       peer_row_off[2] = {0, 0};
    int       peer_tiles_y[2] = {0, 0};
    int       peer_rows[2] = {0, 0};
    int*      peer_counts = nullptr;          // the counter owner's counts array (nullptr: mine)
    void*     ipc_opened[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int       own_lo = 0, own_hi = 0;         // owned local rows (0, 0 = all)
    float     min_pending = 0.f;   // smallest tile key left pending by the last relax, in cost units (+inf = none)
    bool      begun = false;       // g holds a field in progress (nsfs_field_begin ran, nsfs_field_finish has not)
    bool      finished = false;    // g holds a finished field in fp32 cost units
};

struct SearchWork {
    // one slot per concurrently running query (one warp each)
    int       slots = 0;
    int       target = 0;          // the slot count this allocation was made for (slots < target: free HBM clamped it)
    long long heap_cap = 0;        // entries per slot
    uint4*    cells = nullptr;     // [slots][N]   {g.lo, g.hi, parent key, (epoch << 1) | visited}
    unsigned long long* heap = nullptr;  // [slots][heap_cap]
    uint32_t* epoch = nullptr;     // [slots]
    int*      work_counter = nullptr;
    int32_t*  order = nullptr;     // start order of a batch larger than the slots (see team_order_kernel)
    int       order_cap = 0;
};

struct TeamWork {
    // batch kernel (search_team.cu): one slot per team of 8 lanes, four teams per warp
    int       slots = 0;
    int       target = 0;          // see SearchWork
    long long heap_cap = 0;
    void*     cells = nullptr;     // [slots][N]   8-byte {n_card | n_diag << 16, epoch << 4 | visited << 3 | parent direction} or 4-byte records
    int       cell_bytes = 8;
    unsigned long long* heap = nullptr;
    uint32_t* epoch = nullptr;
    int*      work_counter = nullptr;
    int32_t*  order = nullptr;     // start order of a batch larger than the slots (longest expected query first)
    int       order_cap = 0;
};

}  // namespace nsfs

struct nsfs_planner {
    int H = 0, W = 0;
    long long N = 0;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // whole call
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;      // dominant kernel of the call
    cudaStream_t stream2 = nullptr;                // second stream of a split batch (wide kernel beside the team kernel); made on first use
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int carveout = 0;                              // shared-memory carveout (percent) both search kernels ask for while a split batch runs; 0 = each its own
    int lo_thresh = 0;
    bool have_map = false;
    bool have_raw = false;          // d_infl / d_lo hold the int32 planes of the current map

    int32_t*  d_infl = nullptr;     // staging for host maps [N]
    int32_t*  d_lo = nullptr;
    uint32_t* d_free = nullptr;     // bit k = free(k)            [(N+31)/32 + 4 words]
    uint32_t* d_inflated = nullptr; // bit k = infl[k] > 0
    uint16_t* d_out_mask = nullptr; // [N]
    uint16_t* d_in_mask = nullptr;  // [N]

    nsfs::FieldWork field;
    nsfs::SearchWork search;
    nsfs::TeamWork team;

    // small device scratch + pinned host mirror for per-call scalars/results
    void* d_scratch = nullptr; size_t scratch_bytes = 0;
    void* h_pinned = nullptr;  size_t pinned_bytes = 0;

    long long launches = 0;
    // what nsfs_last_path can walk again: 0 nothing, 1 the parent store of slot 0 of the wide kernel (last nsfs_plan), 2 the predecessor field
    int last_plan_kind = 0;
    long long last_ks = 0, last_kg = 0;
    long long opt_max_slots = 0;        // 0 = auto
    long long opt_heap_cap = 0;         // 0 = auto (N entries)
    long long opt_team_entry_bytes = 0; // batch kernel: 0 = 4-byte open-list entries where cost mode "f" and N <= 2^20 allow, 8 = always 8-byte
    long long opt_team_cell_bytes = 0;  // batch kernel: 0 = by map size (4 up to 2048 cells a side), 4 | 8
    long long opt_team_heap_cap = 0;    // batch kernel: 0 = auto (max(N / 8, 4096) entries; overflow falls back to the wide kernel)
    long long opt_split_wide = 0;       // batches on the team kernel: k > 0 = the k longest expected queries run one per warp beside it (capi.cu: launch_split); 0 = off
    long long opt_search_order = 0;     // batches larger than the slots: 0 / 1 = start the longest expected queries first, 2 = input order
    long long opt_search_kernel = 0;    // 0 = auto (batches of Astar / Djikstra on the team kernel), 1 = always one query per warp, 2 = team kernel even for one query
    long long opt_field_inner = 0;      // 0 = default (16)
    long long opt_field_delta = 0;      // 0 = default (128)
    long long opt_graph = 0;            // 0 = the reference's quirk graph; 1 = textbook 8-connected grid
    long long opt_field_mode = 0;       // warps per resident tile: 0 = by map size, 1 = 4 (8 tiles per SM), 2 = 8 (4 tiles), 3 = 16 (2 tiles), 4 = 32 (1 tile)
    long long opt_field_tiles_per_sm = 0;   // cap on resident tiles per SM (0 = as many as fit)
    long long opt_field_gate = 0;       // global gate on tile keys, cost units (0 = default for the residency, < 0 = off)
    long long opt_field_idle_ns = 0;    // 0 = default (1000)
    long long opt_field_fence = 0;      // 0 = __threadfence before key posts, 1 = release atomics (lighter fence, no L1 invalidate)
    long long opt_field_merge = 0;      // a visit absorbs posts that arrive for its own tile before it ends: 0 / 1 = on, 2 = off
    long long opt_field_sched_ns = 0;   // back-off of a block with no pending tile (0 = default 400)
    long long opt_field_wait_ns = 0;    // shortest back-off of a warp with nothing due, right after it had work (0 = default 200)
    // row-partitioned shards: rows outside [active_row_lo, active_row_hi) take no in-edges (they only exist
    // so the masks of the rows next to them see the right neighbourhood); counts cover [count_row_lo, count_row_hi)
    long long opt_active_row_lo = 0, opt_active_row_hi = 0;   // hi == 0: all rows
    long long opt_count_row_lo = 0, opt_count_row_hi = 0;
    char cuda_error[256] = {0};
};

namespace nsfs {

// Every launcher returns NSFS_OK or NSFS_E_CUDA and bumps h->launches by the kernels it launched.
int launch_pack_map(nsfs_planner* h, const int32_t* d_infl, const int32_t* d_lo);
int launch_build_masks(nsfs_planner* h);
int launch_update_rows(nsfs_planner* h, int row0, int nrows);
int launch_test_pos(nsfs_planner* h, int n, const int32_t* d_ij, int8_t* d_out);
int launch_trajectory_check(nsfs_planner* h, int n, const int32_t* d_ij, int t_id, int* d_first_bad);

int field_alloc(nsfs_planner* h);
void field_free(nsfs_planner* h);
int launch_field(nsfs_planner* h, int si, int sj, nsfs_stats* st);
int launch_field_begin(nsfs_planner* h, int si, int sj);
int launch_field_inject(nsfs_planner* h, int row0, int nrows, const float* d_vals, int* d_improved);
int launch_field_relax(nsfs_planner* h, float cap);
int launch_field_relax_p2p(nsfs_planner* h);
int launch_field_count_pending(nsfs_planner* h);
int launch_field_finish(nsfs_planner* h);
int launch_field_min_pending(nsfs_planner* h);
int launch_field_path(nsfs_planner* h, int gi, int gj, int32_t* d_path, int cap, int* d_n, double* d_cost, int* d_status);

int search_alloc(nsfs_planner* h, int want_slots);
void search_free(nsfs_planner* h);
// kind: 0 plan (algo/mode), 1 find_better_point.  All pointers are device pointers; paths optional.
// A split batch (capi.cu: launch_split) passes its own start order (nq = entries of ext_order, d_q = the whole batch), the stream
// to launch on, and part = true: no order of its own, no ev2 / ev3 records.
int launch_search(nsfs_planner* h, int kind, int algo, int mode, int nq, const int32_t* d_q,
                  int32_t* d_status, double* d_g_goal, int32_t* d_path_len, long long* d_settled,
                  int32_t* d_paths, int path_cap, long long* d_stats5,
                  const int32_t* ext_order = nullptr, cudaStream_t ext_stream = nullptr, bool part = false);

int team_alloc(nsfs_planner* h, int nq);        // slots for a batch of nq (search_team.cu: team_slots_wanted)
void team_free(nsfs_planner* h);
// order[k] = the k-th query to start, longest expected first (search_team.cu: team_order_kernel); both batch launchers use it
void launch_start_order(nsfs_planner* h, const int32_t* d_q, int nq, bool heuristic, int32_t* d_order);
int launch_search_team(nsfs_planner* h, int algo, int mode, int nq, const int32_t* d_q, int32_t* d_status, double* d_g_goal,
                       int32_t* d_path_len, long long* d_settled, int32_t* d_paths, int path_cap, long long* d_stats5,
                       const int32_t* ext_order = nullptr, cudaStream_t ext_stream = nullptr, bool part = false);

int launch_inflate(nsfs_planner* h, const int32_t* d_lo, int occ_thresh, double inflation_radius, double cell_size, int32_t* d_infl,
                   uint32_t* d_occ_scratch, int* d_flags);

int launch_fallback_small(nsfs_planner* h, int nq, const int32_t* d_q, int32_t* d_status, long long* d_settled, int32_t* d_cells,
                          long long* d_stats5);

int launch_path_walk(nsfs_planner* h, long long ks, long long kg, int32_t* d_path, int cap, int* d_n);
int launch_los_batch(nsfs_planner* h, int n, const int32_t* d_seg4, int8_t* d_out);
int launch_any_angle(nsfs_planner* h, int n, const int32_t* d_pts, uint8_t* d_keep);

int set_cuda_error(nsfs_planner* h, cudaError_t e, const char* where);

#define NSFS_CUDA_TRY(h, expr)                                                      \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return ::nsfs::set_cuda_error((h), _e, #expr);       \
    } while (0)

// ---- device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ bool bit_at(const uint32_t* __restrict__ plane, long long k) {
    return (__ldg(plane + (k >> 5)) >> (k & 31)) & 1u;
}

}  // namespace nsfs
