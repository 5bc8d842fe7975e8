/* nsfs.h — C-ABI of the B200 grid planner (libnsfs_gpu.so).
 *
 * Drop-in boundary for the grid-search hot path of mukundbala/Nav-Stack-From-Scratch's global_planner.
 * Every entry point names the reference interface it stands in for (paths relative to the reference
 * root; "core.cpp" = src/global_planner/src/algo/grid_planner_core.cpp).  Plain pointers and sizes only:
 * no STL, no exceptions, no torch types cross this boundary.  The C++17 children in
 * nav-stack-from-scratch_b200/host/ (GpuAstar / GpuDjikstra / GpuThetaStar : GridPlannerCore) are the only
 * intended callers on the reference side; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - Cells are (i, j) with flat key k = i*W + j, i <-> x, j <-> y (reference occupancy_grid.cpp:215-218).
 *   - A cell is free iff infl[k] <= 0 && lo[k] <= lo_thresh (testPos, core.cpp:434-455).
 *   - Paths are int32 (i, j) pairs ordered goal -> start, as plan() returns them (astar.cpp:68-85,138-147);
 *     "no path" is the single start cell (astar.cpp:129-135).  World coordinates are formed by the C++
 *     adapter in double (idx2pos, core.cpp:416-421).
 *   - Return value: NSFS_OK or a negative NSFS_E_* code.  NSFS_E_RANGE is the reference's
 *     std::out_of_range (vector::at in testPos); the adapter re-throws it.
 *   - One CUDA stream per handle; calls on one handle are not re-entrant (the reference planners are
 *     single-threaded objects too).  Host pointers are borrowed for the duration of the call.
 *   - There is no CPU fallback: every compute call fails with NSFS_E_CUDA when no sm_100 device is usable.
 */
#ifndef NSFS_H
#define NSFS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSFS_OK          0
#define NSFS_E_RANGE    -1   /* reference would throw std::out_of_range */
#define NSFS_E_CUDA     -2   /* CUDA runtime error; nsfs_last_cuda_error() has the text */
#define NSFS_E_TRUNC    -3   /* caller's output buffer too small (n_path still reports the full length) */
#define NSFS_E_ARG      -4   /* bad argument (start outside the grid, unknown algo, NULL handle, ...) */
#define NSFS_E_NOMEM    -5
#define NSFS_E_HEAP     -6   /* open list outgrew its device capacity even after the internal retry */
#define NSFS_E_NOMAP    -7   /* no map has been set on this handle */
#define NSFS_E_NOFIELD  -8   /* nsfs_field_path() before nsfs_field() */

/* planner_name (global_planner.cpp:104-125) */
#define NSFS_ASTAR           0   /* Astar::plan, astar.cpp:27-148: exact open-list order emulation */
#define NSFS_DJIKSTRA        1   /* Djikstra::plan, djikstra.cpp:27-145: exact open-list order emulation */
#define NSFS_THETASTAR       2   /* no reference implementation (thetastar.cpp is empty): designed here */
#define NSFS_DJIKSTRA_FIELD  3   /* Djikstra::plan by bulk relaxation of the cost-to-go field + predecessor walk:
                                    same reachability and cost, possibly another equal-cost path */

/* cost_mode (core.cpp:243-283): OpenList comparator */
#define NSFS_MODE_F   0
#define NSFS_MODE_G   1
#define NSFS_MODE_FG  2

typedef struct nsfs_planner nsfs_t;

typedef struct nsfs_stats {
    long long pops;          /* OpenList::popNode calls               (search calls) */
    long long expansions;    /* non-stale pops = settled cells        (search calls); reached cells (field) */
    long long pushes;        /* OpenList::pushNode calls */
    long long heap_peak;     /* largest open-list length */
    long long los_checks;    /* isLos evaluations (ThetaStar) */
    long long rounds;        /* global relaxation rounds (field) */
    long long tile_visits;   /* tile relaxations (field) */
    long long launches;      /* kernels launched by this call */
    float     gpu_ms;        /* device time of the call's kernels (CUDA events on the handle's stream) */
    float     main_kernel_ms;/* device time of the call's dominant kernel alone (field_solve / search), same clock */
} nsfs_stats;

/* ---- lifetime ------------------------------------------------------------------------------------
 * replaces: GridPlannerCore(cost_mode, map) / prepPlanner (core.cpp:353-383,570-600), which size the
 * per-cell Node store from map_size_.  device = CUDA ordinal. */
int  nsfs_create(nsfs_t** out, int H, int W, int device);
void nsfs_destroy(nsfs_t* h);

/* ---- map -----------------------------------------------------------------------------------------
 * replaces: the MapData& argument every reference call takes (map_data.h:8-20).  The reference re-reads
 * the caller's vectors on every call; here the caller hands them over whenever they changed
 * (global_planner.cpp:65-73 replaces them per ROS callback).  Packs to 1 bit/cell free + 1 bit/cell
 * inflated planes and the per-cell edge masks on the device. */
int nsfs_set_map(nsfs_t* h, const int32_t* infl, const int32_t* lo, int lo_thresh);          /* host pointers   */
int nsfs_set_map_device(nsfs_t* h, const int32_t* d_infl, const int32_t* d_lo, int lo_thresh); /* device pointers */
/* Only rows [row0, row0 + nrows) changed since the last nsfs_set_map (host pointers to nrows x W cells of each plane):
 * uploads those rows and rebuilds only the bit-plane words and edge masks that can see them (rows +-4).  Same result as
 * nsfs_set_map with the whole updated planes.  Needs a map set by nsfs_set_map (NSFS_E_NOMAP otherwise). */
int nsfs_update_map_rows(nsfs_t* h, int row0, int nrows, const int32_t* infl_rows, const int32_t* lo_rows);
/* The map producer's last step on the device (reference loco_mapping, occupancy_grid.cpp:73-89,166-213): the caller hands over
 * only the log-odds plane; the inflation plane (number of cells with log-odds >= occ_thresh whose disc mask covers the cell, with
 * the reference's flat-key / row-0 quirks) is computed on the GPU, then the map is packed as by nsfs_set_map.  Halves the upload.
 * inflation_radius / cell_size as in generateMask (0.2 / 0.05 upstream).  NSFS_E_RANGE: the reference's updateInflation would
 * have thrown (an occupied cell's disc runs past the last cell).  nsfs_get_inflation copies the plane back (tests). */
int nsfs_set_map_logodds(nsfs_t* h, const int32_t* lo, int lo_thresh, int occ_thresh, double inflation_radius, double cell_size);
int nsfs_get_inflation(nsfs_t* h, int32_t* infl_out);
/* Packed planes for shipping a map between GPUs (N/4 bytes instead of 8N): words = (H*W + 31) / 32 each. */
int nsfs_map_words(const nsfs_t* h);
int nsfs_get_packed_map_device(nsfs_t* h, const uint32_t** d_free, const uint32_t** d_inflated);
int nsfs_set_packed_map_device(nsfs_t* h, const uint32_t* d_free, const uint32_t* d_inflated);
/* testPos(Index) for a batch of cells (core.cpp:434-455): out = 1 free, 0 not free / row out of range, -1 would throw */
int nsfs_test_pos_batch(nsfs_t* h, int n, const int32_t* ij, int8_t* out);
/* The path consumer's check of what the planner handed it.
 * replaces: Commander::checkTrajectorySafety + checkCell + oob (commander.cpp:316-384,394-397) on the trajectory's cells
 * (ij = Commander::pos2idx of each point, 386-391; n points, host pointer).  Same results: n == 0 -> (safe 0, bad_idx -2);
 * n == 1 -> point 0 decides, bad_idx -1; else points t_id, t_id-1 .. 0 and the first unsafe one is bad_idx (safe 1, -1 when
 * none).  The commander's oob treats row 0 as outside and never tests the column.  NSFS_E_RANGE where the reference
 * throws std::out_of_range (t_id >= n, or a key outside the planes) before meeting an unsafe point. */
int nsfs_check_trajectory(nsfs_t* h, int n, const int32_t* ij, int t_id, int* safe, int* bad_idx);

/* ---- single query --------------------------------------------------------------------------------
 * replaces: GridPlannerCore::plan(Index, Index, MapData&) of the selected child.
 * path_ij: cap (i,j) pairs; n_path = number of path cells; g_goal = nodes[goal].g after the search
 * (1e5 when unreachable).  Any of path_ij / n_path / g_goal / st may be NULL. */
int nsfs_plan(nsfs_t* h, int algo, int cost_mode, int si, int sj, int gi, int gj,
              int32_t* path_ij, int cap, int* n_path, double* g_goal, nsfs_stats* st);

/* The path of the last nsfs_plan call on this handle once more, without searching again: for a caller whose buffer was too
 * small (NSFS_E_TRUNC told it the length).  replaces: nothing upstream (plan() returns a std::vector of whatever length,
 * astar.cpp:70-85); it spares the drop-in children a second search.  NSFS_E_NOFIELD when the search state is gone (another
 * search or a map change came in between): plan again. */
int nsfs_last_path(nsfs_t* h, int32_t* path_ij, int cap, int* n_path);

/* ---- batched queries -----------------------------------------------------------------------------
 * nq independent plan() calls on the current map; sg4[4q..] = {si, sj, gi, gj}.
 * Per query: status (NSFS_OK / NSFS_E_RANGE / NSFS_E_HEAP), g_goal, path_len (cells), settled (expansions).
 * paths_ij (optional) receives min(path_len, path_cap) cells per query at stride path_cap. */
int nsfs_plan_batch(nsfs_t* h, int algo, int cost_mode, int nq, const int32_t* sg4,
                    int32_t* status, double* g_goal, int32_t* path_len, long long* settled,
                    int32_t* paths_ij, int path_cap, nsfs_stats* st);
/* Same, with every array already resident on the handle's device (no copies inside the call). */
int nsfs_plan_batch_device(nsfs_t* h, int algo, int cost_mode, int nq, const int32_t* d_sg4,
                           int32_t* d_status, double* d_g_goal, int32_t* d_path_len, long long* d_settled,
                           int32_t* d_paths_ij, int path_cap, nsfs_stats* st);

/* ---- cost-to-go field ----------------------------------------------------------------------------
 * replaces: Djikstra::plan with an unreachable goal, i.e. nodes[k].g for every k (djikstra.cpp:27-145).
 * g: fp32 [H*W], 1e5 = unreached.  pred: direction code d in 0..7 of the incoming edge (parent key =
 * k - OFF[d], OFF = {+W, +W+1, +1, -W+1, -W, -W-1, -1, +W-1}, NB_LUT order grid_planner_core.h:197),
 * 255 = none (unreached or the source).  Either host pointer may be NULL. */
int nsfs_field(nsfs_t* h, int si, int sj, float* g, uint8_t* pred, nsfs_stats* st);
/* Same, result left on the device; the pointers stay valid until the next field call on this handle. */
int nsfs_field_device(nsfs_t* h, int si, int sj, const float** d_g, const uint8_t** d_pred, nsfs_stats* st);
/* Walk the predecessor field of the last nsfs_field* call from (gi, gj) back to its source.
 * cost = (#cardinal-cost steps) + (#diagonal-cost steps) * sqrt(2) in double. */
int nsfs_field_path(nsfs_t* h, int gi, int gj, int32_t* path_ij, int cap, int* n_path, double* cost);

/* ---- row-partitioned field (one shard of a map split by row blocks across GPUs) ---------------
 * The shard's handle covers its owned rows plus 4 rows of map on either side (2 ghost rows whose g is exchanged
 * + 2 more so the ghost rows' edge masks see their true neighbourhood).  Before nsfs_set_map*, restrict
 * propagation with nsfs_set_option("active_row_lo"/"active_row_hi") and the counters with "count_row_lo"/"count_row_hi"
 * (local row numbers, hi exclusive, 0 = unset).  Then per solve:
 *   nsfs_field_begin(h, si, sj)            g = 1e5 everywhere; (si, sj) local source, or si < 0 when another shard owns it
 *   repeat { nsfs_field_inject_rows(..)    g[row0 .. row0+nrows) = min(g, vals): ghost rows received from a neighbour shard
 *            nsfs_field_relax[_until](..) } relax to the local fixed point (or up to a cost cap); the caller exchanges
 *                                          boundary rows between shards
 *   nsfs_field_finish(h, st)               predecessor field + reached count of the counted rows; then nsfs_field_device*
 * All three leave results on the device (nsfs_field_device's pointers stay valid). */
int nsfs_field_begin(nsfs_t* h, int si, int sj);
int nsfs_field_inject_rows(nsfs_t* h, int row0, int nrows, const float* d_vals, int* n_improved);   /* device values: rows read from
                                            another shard's plane (nsfs_field_pointers) DURING its solve, i.e. in the plane's working
                                            representation (32-bit words, Q17.15 fixed point), not fp32 cost units */
int nsfs_field_relax(nsfs_t* h, nsfs_stats* st);
/* Same, but nothing is relaxed beyond the cost `cap`: tiles whose smallest pending value exceeds it stay pending.
 * min_pending = the smallest value still pending afterwards (+inf when the shard is at its fixed point).  Lets all
 * shards of a partitioned map advance band by band (cap = global min pending + band width) instead of one after another. */
int nsfs_field_relax_until(nsfs_t* h, float cap, float* min_pending, nsfs_stats* st);
int nsfs_field_min_pending(nsfs_t* h, float* min_pending);   /* e.g. after nsfs_field_inject_rows */
int nsfs_field_finish(nsfs_t* h, nsfs_stats* st);
int nsfs_field_pointers(nsfs_t* h, const float** d_g, const uint8_t** d_pred);   /* valid after nsfs_field_begin; the g plane holds
                                            Q17.15 words until nsfs_field_finish turns it into fp32 cost units in place */

/* ---- peer-to-peer shards over NVLink (one process per GPU) ----------------------------------------------------
 * Instead of exchange rounds, each shard's solve kernel pushes every improvement of its first / last two owned rows
 * straight into the neighbour shard's ghost rows (remote atomicMin on the neighbour's g), wakes the neighbour's tiles
 * (remote atomicMin on its tile keys) and all kernels share ONE counter of pending tiles, so a single launch per GPU runs
 * every shard to the global fixed point.  Buffers travel between processes as CUDA IPC handles.
 *   nsfs_field_peer_export(h, &info)                     after nsfs_set_map*; send info to the neighbours and rank 0's to all
 *   nsfs_field_peer_attach(h, side, &nb_info, row_off)   side 0 = shard above, 1 = below; my local row r = its local row r + row_off
 *   nsfs_field_peer_counter(h, &owner_info | NULL)       whose counter is shared (NULL: mine)
 *   nsfs_field_peer_rows(h, own_lo, own_hi)              my owned local rows
 *   per solve: nsfs_field_begin; nsfs_field_pending_count on every shard; sum; nsfs_field_set_counter(total) on the owner;
 *              barrier; nsfs_field_relax_p2p on every shard; barrier; nsfs_field_finish. */
typedef struct nsfs_peer_info {
    unsigned char g[64], key[64], counts[64];             /* cudaIpcMemHandle_t of the g, tile-key and counter arrays */
    int rows, W, tiles_x, tiles_y;
} nsfs_peer_info;
int nsfs_field_peer_export(nsfs_t* h, nsfs_peer_info* out);
int nsfs_field_peer_attach(nsfs_t* h, int side, const nsfs_peer_info* nb, int row_off);
int nsfs_field_peer_counter(nsfs_t* h, const nsfs_peer_info* owner);
int nsfs_field_peer_rows(nsfs_t* h, int own_lo, int own_hi);
int nsfs_field_peer_detach(nsfs_t* h);
int nsfs_field_pending_count(nsfs_t* h, int* n_pending);
int nsfs_field_set_counter(nsfs_t* h, int total);
int nsfs_field_relax_p2p(nsfs_t* h, nsfs_stats* st);

/* ---- groups: one process per GPU, for the two sharded configurations --------------------------------------
 * replaces: nothing upstream (the reference is single-threaded and single-device); this is the C++/C entry a GlobalPlanner
 * process per GPU would use instead of looping GridPlannerCore::generatePath over requests (global_planner.cpp:102-130, 183).
 * NCCL (libnccl.so.2) is bound at run time; without it nsfs_group_unique_id / nsfs_group_create return NSFS_E_CUDA.
 *   rank 0: nsfs_group_unique_id(id); ship the 128 bytes to the other ranks out of band (a ROS parameter, a file, a socket)
 *   every rank: nsfs_group_create(&g, id, rank, world, device)
 * Batched queries (BASELINE config "65,536 Astar queries sharded across 1/2/4/8 B200"):
 *   nsfs_set_map*(h) on the root only; nsfs_group_broadcast_map(g, h, root) ships the packed bit planes (N/4 bytes);
 *   nsfs_group_plan_batch(g, h, ...): every rank passes all nq queries and gets all nq results; rank r plans the contiguous
 *   slice [r*nq/world ..) on its own GPU, results are all-gathered.  No exchange during the search.
 * Row-block field (BASELINE config "16384x16384 tile-partitioned across 2/4/8 B200 with NVLink halo exchange"):
 *   each rank's handle covers its owned rows + 4 rows of map per side (see "row-partitioned field" above);
 *   nsfs_group_field_attach(g, h, held_lo, own_lo, own_hi) once (global row numbers; exchanges the CUDA IPC handles over NCCL);
 *   nsfs_group_field_solve(g, h, si, sj, &reached, &st) per field (global source): one solve kernel per GPU, boundary rows
 *   posted peer to peer over NVLink from inside it; then nsfs_field_pointers / nsfs_field_path on the owning shard. */
typedef struct nsfs_group nsfs_group_t;
int  nsfs_group_unique_id(unsigned char id[128]);
int  nsfs_group_create(nsfs_group_t** out, const unsigned char id[128], int rank, int world, int device);
void nsfs_group_destroy(nsfs_group_t* g);
int  nsfs_group_rank(const nsfs_group_t* g);
int  nsfs_group_world(const nsfs_group_t* g);
int  nsfs_group_barrier(nsfs_group_t* g);
const char* nsfs_group_error(const nsfs_group_t* g);
int  nsfs_group_broadcast_map(nsfs_group_t* g, nsfs_t* h, int root);
int  nsfs_group_plan_batch(nsfs_group_t* g, nsfs_t* h, int algo, int cost_mode, int nq, const int32_t* sg4, int32_t* status,
                           double* g_goal, int32_t* path_len, long long* settled, nsfs_stats* st);
int  nsfs_group_field_attach(nsfs_group_t* g, nsfs_t* h, int held_lo, int own_lo, int own_hi);
int  nsfs_group_field_solve(nsfs_group_t* g, nsfs_t* h, int si, int sj, long long* reached_total, nsfs_stats* st);

/* ---- fallback ------------------------------------------------------------------------------------
 * replaces: Djikstra::find_better_point(Index, MapData&) (djikstra.cpp:165-273) on a planner prepared
 * with cost mode "g" (global_planner.cpp:127-128).  (fi, fj) = first popped free cell, or the input cell
 * when none is found (djikstra.cpp:272). */
int nsfs_find_better_point(nsfs_t* h, int bi, int bj, int* fi, int* fj, nsfs_stats* st);
int nsfs_find_better_point_batch(nsfs_t* h, int n, const int32_t* bad_ij, int32_t* out_ij, int32_t* status,
                                 nsfs_stats* st);

/* ---- line of sight and post-processing -----------------------------------------------------------
 * replaces: GridPlannerCore::isLos(Index, Index, MapData&) (core.cpp:481-496) over bresenham_los
 * (bot_utils.cpp:137-192).  seg4[4s..] = {si, sj, ti, tj}; out = 1 LOS, 0 blocked, -1 would throw. */
int nsfs_los_batch(nsfs_t* h, int n, const int32_t* seg4, int8_t* out);
/* replaces: the isLos loop of makeAnyAnglePath (core.cpp:537-558) over turning points already converted
 * with pos2idx.  keep[t] = 1 if point t survives (first and last always do). */
int nsfs_any_angle(nsfs_t* h, int n, const int32_t* pts_ij, uint8_t* keep);

/* Page-lock / release a caller-owned host range (cudaHostRegister) so that nsfs_set_map on it is a DMA at PCIe speed instead of a
 * staged pageable copy.  replaces: nothing upstream; serves MapData's two std::vector<int> planes, which the ROS callbacks
 * overwrite in place (global_planner.cpp:65-73).  Optional: every host-pointer call works on pageable memory too. */
int nsfs_pin_host(nsfs_t* h, const void* p, size_t bytes);
int nsfs_unpin_host(nsfs_t* h, const void* p);

/* ---- service -------------------------------------------------------------------------------------*/
int         nsfs_sync(nsfs_t* h);
long long   nsfs_kernel_launches(const nsfs_t* h);   /* kernels launched through this handle so far */
void*       nsfs_stream(nsfs_t* h);                  /* cudaStream_t of the handle */
/* Options (all have working defaults; 0 = default unless stated):
 *   "graph"          0 the reference's quirk graph | 1 textbook 8-connected grid; rebuilds the edge masks
 *   "field_mode"     warps per resident tile of the field solver: 0 = chosen from the map size | 1 = 4 warps (8 tiles resident per
 *                    SM) | 2 = 8 (4 tiles) | 3 = 16 (2 tiles) | 4 = 32 (1 tile);  "field_gate"  how far (cost units) a block may run
 *                    ahead of the slowest pending tile (0 = default for the residency, -1 = no gate)
 *   "field_tiles_per_sm"  cap on resident tiles per SM (0 = as many as fit)
 *   "field_delta", "field_inner", "field_idle_ns", "field_wait_ns", "field_sched_ns", "field_merge", "field_fence"
 *                    scheduling knobs of the field solver (band width in cost units, passes per sub-block turn, back-offs,
 *                    merged visits 2 = off, release-atomic key posts 1 = on); the field is solved in exact integer arithmetic,
 *                    so the result is the same bits whatever they are set to
 *   "search_kernel"  0 auto | 1 one query per warp | 2 four queries per warp;   "search_order"  2 = start a batch in input order
 *   "max_slots", "heap_cap", "team_heap_cap"   query slots (in warps) and open-list capacity per slot
 *   "team_cell_bytes"  4 | 8: per-cell record of the batch kernel (0 = 4 up to 2048 cells a side);  "team_entry_bytes"  8 = always
 *                    8-byte open-list entries (0 = 4-byte ones where cost mode "f" and at most 2^20 cells allow)
 *   "split_wide"     k > 0: the k longest expected queries of a batch run one per warp on a second stream beside the batch
 *                    kernel (measured slower than the unsplit batch, so off by default)
 *   "active_row_lo/hi", "count_row_lo/hi"      row windows of a shard (see the row-partitioned field above)
 * Environment overrides read by nsfs_create (experiments / running the test-suite per solver): NSFS_FIELD_MODE, NSFS_FIELD_FENCE,
 * NSFS_SEARCH_ORDER, NSFS_MAX_SLOTS, NSFS_SPLIT_WIDE, NSFS_FIELD_GATE. */
int         nsfs_set_option(nsfs_t* h, const char* key, long long value);
const char* nsfs_strerror(int code);
const char* nsfs_last_cuda_error(const nsfs_t* h);
const char* nsfs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NSFS_H */
