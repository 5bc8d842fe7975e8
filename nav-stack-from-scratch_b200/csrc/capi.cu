// extern "C" surface of libnsfs_gpu.so (include/nsfs.h): argument checks, staging copies, launches, results.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "nsfs_internal.cuh"

namespace nsfs {

int set_cuda_error(nsfs_planner* h, cudaError_t e, const char* where) {
    if (h) snprintf(h->cuda_error, sizeof(h->cuda_error), "%s: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e), where);
    cudaGetLastError();   // clear the sticky-less error state
    return NSFS_E_CUDA;
}

static int ensure_scratch(nsfs_planner* h, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes <= h->scratch_bytes) return NSFS_OK;
    if (h->d_scratch) cudaFree(h->d_scratch);
    h->d_scratch = nullptr; h->scratch_bytes = 0;
    NSFS_CUDA_TRY(h, cudaMalloc(&h->d_scratch, bytes));
    h->scratch_bytes = bytes;
    return NSFS_OK;
}

struct Timer {
    nsfs_planner* h;
    explicit Timer(nsfs_planner* h_) : h(h_) { cudaEventRecord(h->ev0, h->stream); }
    float stop_sync() {
        cudaEventRecord(h->ev1, h->stream);
        cudaEventSynchronize(h->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        return ms;
    }
    float main_kernel_ms() {     // valid after stop_sync(): ev2/ev3 bracket the dominant kernel on the same stream
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev2, h->ev3) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
        return ms;
    }
};

static int enter(nsfs_planner* h, bool need_map) {
    if (!h) return NSFS_E_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return set_cuda_error(h, cudaGetLastError(), "cudaSetDevice");
    if (need_map && !h->have_map) return NSFS_E_NOMAP;
    return NSFS_OK;
}

static int finish_map(nsfs_planner* h) {
    int rc = launch_build_masks(h);
    if (rc != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_map = true;
    h->field.valid = false;
    h->field.begun = false;
    return NSFS_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace nsfs

using namespace nsfs;

extern "C" {

int nsfs_create(nsfs_t** out, int H, int W, int device) {
    if (!out || H <= 0 || W <= 0) return NSFS_E_ARG;
    const long long N = (long long)H * W;
    if (N >= (1ll << 29)) return NSFS_E_ARG;            // open-list entries carry a 29-bit key
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return NSFS_E_CUDA;                              // no CPU fallback exists
    }
    nsfs_planner* h = new (std::nothrow) nsfs_planner();
    if (!h) return NSFS_E_NOMEM;
    h->H = H; h->W = W; h->N = N; h->device = device;
    int rc = NSFS_OK;
    do {
        if (cudaSetDevice(device) != cudaSuccess) { rc = set_cuda_error(h, cudaGetLastError(), "cudaSetDevice"); break; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { rc = set_cuda_error(h, cudaGetLastError(), "cudaGetDeviceProperties"); break; }
        h->sm_count = prop.multiProcessorCount;
        const size_t words = (size_t)((N + 31) / 32) + 8;
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess ||
            (e = cudaEventCreate(&h->ev2)) != cudaSuccess || (e = cudaEventCreate(&h->ev3)) != cudaSuccess ||
            (e = cudaMalloc(&h->d_free, words * 4)) != cudaSuccess || (e = cudaMalloc(&h->d_inflated, words * 4)) != cudaSuccess ||
            (e = cudaMalloc(&h->d_out_mask, (size_t)N * 2)) != cudaSuccess || (e = cudaMalloc(&h->d_in_mask, (size_t)N * 2)) != cudaSuccess ||
            (e = cudaMemsetAsync(h->d_free, 0, words * 4, h->stream)) != cudaSuccess ||
            (e = cudaMemsetAsync(h->d_inflated, 0, words * 4, h->stream)) != cudaSuccess) {
            rc = set_cuda_error(h, e, "nsfs_create allocations");
            break;
        }
        rc = ensure_scratch(h, 1 << 16);
    } while (0);
    if (rc != NSFS_OK) { nsfs_destroy(h); return rc; }
    if (const char* e = getenv("NSFS_FIELD_FENCE")) h->opt_field_fence = atoi(e);
    if (const char* e = getenv("NSFS_SEARCH_ORDER")) h->opt_search_order = atoi(e);
    if (const char* e = getenv("NSFS_SPLIT_WIDE")) h->opt_split_wide = atoi(e);
    if (const char* e = getenv("NSFS_MAX_SLOTS")) h->opt_max_slots = atoll(e);    // warps of query slots (experiments)
    if (const char* e = getenv("NSFS_FIELD_GATE")) h->opt_field_gate = atoi(e);
    if (const char* e = getenv("NSFS_FIELD_MODE")) h->opt_field_mode = atoi(e);   // default field solver of new handles (tests run the suite per solver)
    *out = h;
    return NSFS_OK;
}

void nsfs_destroy(nsfs_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int k = 0; k < 5; ++k)
        if (h->field.ipc_opened[k]) cudaIpcCloseMemHandle(h->field.ipc_opened[k]);
    field_free(h);
    search_free(h);
    team_free(h);
    cudaFree(h->d_infl); cudaFree(h->d_lo); cudaFree(h->d_free); cudaFree(h->d_inflated);
    cudaFree(h->d_out_mask); cudaFree(h->d_in_mask); cudaFree(h->d_scratch);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->ev3) cudaEventDestroy(h->ev3);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

int nsfs_set_map(nsfs_t* h, const int32_t* infl, const int32_t* lo, int lo_thresh) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    if (!infl || !lo) return NSFS_E_ARG;
    const size_t bytes = (size_t)h->N * sizeof(int32_t);
    if (!h->d_infl) {
        NSFS_CUDA_TRY(h, cudaMalloc(&h->d_infl, bytes + 16));
        NSFS_CUDA_TRY(h, cudaMalloc(&h->d_lo, bytes + 16));
    }
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->d_infl, infl, bytes, cudaMemcpyHostToDevice, h->stream));
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->d_lo, lo, bytes, cudaMemcpyHostToDevice, h->stream));
    h->lo_thresh = lo_thresh;
    rc = launch_pack_map(h, h->d_infl, h->d_lo);
    if (rc != NSFS_OK) return rc;
    rc = finish_map(h);
    h->have_raw = rc == NSFS_OK;
    return rc;
}

int nsfs_set_map_logodds(nsfs_t* h, const int32_t* lo, int lo_thresh, int occ_thresh, double inflation_radius, double cell_size) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    if (!lo || !(cell_size > 0.0) || !(inflation_radius >= 0.0)) return NSFS_E_ARG;
    const size_t bytes = (size_t)h->N * sizeof(int32_t);
    if (!h->d_infl) {
        NSFS_CUDA_TRY(h, cudaMalloc(&h->d_infl, bytes + 16));
        NSFS_CUDA_TRY(h, cudaMalloc(&h->d_lo, bytes + 16));
    }
    const size_t occ_bytes = ((size_t)((h->N + 31) / 32) + 8) * 4;
    if ((rc = ensure_scratch(h, 256 + occ_bytes)) != NSFS_OK) return rc;
    h->have_map = false; h->have_raw = false;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->d_lo, lo, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = launch_inflate(h, h->d_lo, occ_thresh, inflation_radius, cell_size, h->d_infl, (uint32_t*)((char*)h->d_scratch + 256), (int*)h->d_scratch);
    if (rc != NSFS_OK) return rc;
    int flags = 0;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(&flags, h->d_scratch, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->lo_thresh = lo_thresh;
    if ((rc = launch_pack_map(h, h->d_infl, h->d_lo)) != NSFS_OK) return rc;
    if ((rc = finish_map(h)) != NSFS_OK) return rc;
    h->have_raw = true;
    return (flags & 1) ? NSFS_E_RANGE : NSFS_OK;
}

int nsfs_get_inflation(nsfs_t* h, int32_t* infl_out) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->have_raw || !infl_out) return NSFS_E_ARG;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(infl_out, h->d_infl, (size_t)h->N * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NSFS_OK;
}

// Rows [row0, row0 + nrows) of the caller's planes changed since the last nsfs_set_map: upload only those rows and
// rebuild only what can see them (the reference re-reads the whole MapData on every call; its maps arrive at 25 Hz and
// a lidar sweep touches a few rows' worth of cells, global_planner.cpp:65-73, occupancy_grid.cpp:166-196).
int nsfs_update_map_rows(nsfs_t* h, int row0, int nrows, const int32_t* infl_rows, const int32_t* lo_rows) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->have_raw) return NSFS_E_NOMAP;                 // needs the int32 planes of nsfs_set_map on the device
    if (row0 < 0 || nrows < 0 || row0 + nrows > h->H || (nrows && (!infl_rows || !lo_rows))) return NSFS_E_ARG;
    if (!nrows) return NSFS_OK;
    const size_t off = (size_t)row0 * h->W, bytes = (size_t)nrows * h->W * sizeof(int32_t);
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->d_infl + off, infl_rows, bytes, cudaMemcpyHostToDevice, h->stream));
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->d_lo + off, lo_rows, bytes, cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_update_rows(h, row0, nrows)) != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->field.valid = false;
    h->field.begun = false;
    return NSFS_OK;
}

int nsfs_set_map_device(nsfs_t* h, const int32_t* d_infl, const int32_t* d_lo, int lo_thresh) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    h->have_raw = false;
    if (!d_infl || !d_lo || ((uintptr_t)d_infl & 15) || ((uintptr_t)d_lo & 15)) return NSFS_E_ARG;   // 128-bit loads
    h->lo_thresh = lo_thresh;
    rc = launch_pack_map(h, d_infl, d_lo);
    if (rc != NSFS_OK) return rc;
    return finish_map(h);
}

int nsfs_map_words(const nsfs_t* h) { return h ? (int)((h->N + 31) / 32) : 0; }

int nsfs_get_packed_map_device(nsfs_t* h, const uint32_t** d_free, const uint32_t** d_inflated) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (d_free) *d_free = h->d_free;
    if (d_inflated) *d_inflated = h->d_inflated;
    return NSFS_OK;
}

int nsfs_set_packed_map_device(nsfs_t* h, const uint32_t* d_free, const uint32_t* d_inflated) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    h->have_raw = false;
    if (!d_free || !d_inflated) return NSFS_E_ARG;
    const size_t bytes = (size_t)((h->N + 31) / 32) * 4;
    if (d_free != h->d_free) NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->d_free, d_free, bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (d_inflated != h->d_inflated) NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->d_inflated, d_inflated, bytes, cudaMemcpyDeviceToDevice, h->stream));
    return finish_map(h);
}

int nsfs_test_pos_batch(nsfs_t* h, int n, const int32_t* ij, int8_t* out) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (n < 0 || (n && (!ij || !out))) return NSFS_E_ARG;
    if (!n) return NSFS_OK;
    const size_t in_b = align_up((size_t)n * 8, 256);
    if ((rc = ensure_scratch(h, in_b + n)) != NSFS_OK) return rc;
    char* d = (char*)h->d_scratch;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(d, ij, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_test_pos(h, n, (const int32_t*)d, (int8_t*)(d + in_b))) != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(out, d + in_b, n, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NSFS_OK;
}

int nsfs_check_trajectory(nsfs_t* h, int n, const int32_t* ij, int t_id, int* safe, int* bad_idx) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (n < 0 || (n && !ij) || !safe || !bad_idx) return NSFS_E_ARG;
    if (n == 0) { *safe = 0; *bad_idx = -2; return NSFS_OK; }                      // commander.cpp:320-324
    const bool single = n == 1;                                                     // 326-340: only point 0, reported as index -1
    const int last = single ? 0 : t_id;
    *safe = 1; *bad_idx = -1;
    if (last < 0) return NSFS_OK;                                                   // the loop body never runs
    if (last >= n) return NSFS_E_RANGE;                                             // trajectory_.at(t_id) is the first thing the loop does
    const size_t in_b = align_up((size_t)n * 8, 256);
    if ((rc = ensure_scratch(h, in_b + 256)) != NSFS_OK) return rc;
    char* d = (char*)h->d_scratch;
    int first_bad = -1;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(d, ij, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(d + in_b, &first_bad, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_trajectory_check(h, n, (const int32_t*)d, last, (int*)(d + in_b))) != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(&first_bad, d + in_b, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (first_bad < 0) return NSFS_OK;
    if (first_bad & 1) return NSFS_E_RANGE;                                         // the first point the reference would trip over throws
    *safe = 0;
    *bad_idx = single ? -1 : (first_bad >> 1);
    return NSFS_OK;
}

// ---- searches ------------------------------------------------------------------------------------------
static void fill_stats(nsfs_stats* st, const long long* s5, long long launches, float ms, float main_ms) {
    if (!st) return;
    memset(st, 0, sizeof(*st));
    st->pops = s5[0]; st->expansions = s5[1]; st->pushes = s5[2]; st->heap_peak = s5[3]; st->los_checks = s5[4];
    st->launches = launches; st->gpu_ms = ms; st->main_kernel_ms = main_ms;
}

// Batches of Astar / Djikstra go to the four-queries-per-warp kernel (search_team.cu); single queries, ThetaStar, the
// fallback search and degenerate widths stay on the one-query-per-warp kernel (search.cu).
static bool use_team_kernel(const nsfs_planner* h, int kind, int algo, int nq) {
    if (kind != 0 || (algo != NSFS_ASTAR && algo != NSFS_DJIKSTRA) || h->W < 3) return false;
    if (h->opt_search_kernel == 1) return false;
    if (h->opt_search_kernel == 2) return true;
    // A launch lasts as long as its slowest query, and one pop is a dependent chain that a 32-lane warp walks faster than an
    // 8-lane team (5 instead of 3 heap levels per round trip, fewer instructions in sequence): measured on C4, 8 192 queries
    // take 4.2 s one per warp against 5.5 s four per warp, while at 65 536 the team kernel's four-fold query residency wins
    // (profiles/r02_search.md).  So batches that fit two waves of warp slots stay on the wide kernel.
    return nq > 2 * h->sm_count * 32;                  // (9 472 on a B200)
}

// Split batches (option "split_wide" = k > 0; off by default).  A launch lasts as long as its slowest query, and the slowest C4
// queries (900 000 expansions) take several seconds on an 8-lane team under full load while a 32-lane warp walks the same chain
// about twice as fast; so the head of the start order (the k longest expected queries, team_order_kernel) can go to the wide kernel
// on a second stream while the team kernel takes the rest.  Statuses, paths and settled counts do not depend on which kernel ran a
// query (costs agree to the last few ulps: fp64 sums there, a + b * sqrt(2) here).  MEASURED (profiles/r02_search.md section 4): the
// first split launch of a process serialises (the second kernel's module is loaded lazily, which waits for the first); from the
// second launch on the kernels overlap, and the batch is STILL slower than unsplit - 16 384 C4 queries 5.8 - 6.0 s against 5.4 s,
// 65 536 queries 12.1 s against 10.4 s: next to the team kernel's load the wide kernel's long queries run no faster than they
// would on a team.  So nothing selects it; it stays as a tested option.
static int split_wide_count(const nsfs_planner* h, int nq) {
    if (h->opt_split_wide <= 0 || h->opt_search_order == 2) return 0;
    long long k = h->opt_split_wide;
    if (k > nq / 4) k = nq / 4;
    if (h->opt_max_slots > 0 && k > h->opt_max_slots) k = h->opt_max_slots;
    return k < 4 ? 0 : (int)k;
}

// Per-query state for the kernel(s) a batch of nq will run on.  The two kernels' state competes for the same HBM: a plain batch
// frees the other kernel's; a split batch keeps k wide slots and gives the team kernel what the rest of HBM holds.
static int alloc_for_search(nsfs_planner* h, bool team, int nq) {
    if (!team) {
        if (nq > 1 && h->team.slots > 0) team_free(h);
        return search_alloc(h, nq);
    }
    const int k = split_wide_count(h, nq);
    if (k == 0) {
        if (h->search.slots > 1) search_free(h);
        return team_alloc(h, nq);
    }
    if (h->search.slots > 2 * k) search_free(h);
    const int rc = search_alloc(h, k);
    return rc != NSFS_OK ? rc : team_alloc(h, nq - k);
}

static int launch_split(nsfs_planner* h, int k_wide, int algo, int mode, int nq, const int32_t* d_q, int32_t* d_status,
                        double* d_g_goal, int32_t* d_path_len, long long* d_settled, int32_t* d_paths, int path_cap, long long* d_s5) {
    int rc;
    if (!h->stream2) {
        NSFS_CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
        NSFS_CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        NSFS_CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    if ((rc = alloc_for_search(h, true, nq)) != NSFS_OK) return rc;
    nsfs::TeamWork& t = h->team;
    if (t.order_cap < nq) {
        cudaFree(t.order); t.order = nullptr; t.order_cap = 0;
        NSFS_CUDA_TRY(h, cudaMalloc(&t.order, (size_t)nq * sizeof(int32_t)));
        t.order_cap = nq;
    }
    nsfs::launch_start_order(h, d_q, nq, algo == NSFS_ASTAR, t.order);
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev_fork, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    // two wide blocks (33 KB each) and up to seven team blocks (17 KB) per SM.  Blocks of two kernels share an SM only under ONE
    // shared-memory configuration, and the driver raises a kernel's preference to what its own full occupancy needs (228 KB for
    // the wide kernel): both ask for the maximum
    h->carveout = 100;
    rc = nsfs::launch_search(h, 0, algo, mode, k_wide, d_q, d_status, d_g_goal, d_path_len, d_settled, d_paths, path_cap, d_s5,
                             t.order, h->stream2, true);
    if (rc == NSFS_OK)
        rc = nsfs::launch_search_team(h, algo, mode, nq - k_wide, d_q, d_status, d_g_goal, d_path_len, d_settled, d_paths, path_cap, d_s5,
                                      t.order + k_wide, h->stream, true);
    h->carveout = 0;
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev_join, h->stream2));
    NSFS_CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    return rc;
}

static int launch_any_search(nsfs_planner* h, bool team, int kind, int algo, int mode, int nq, const int32_t* d_q, int32_t* d_status,
                             double* d_g_goal, int32_t* d_path_len, long long* d_settled, int32_t* d_paths, int path_cap, long long* d_s5) {
    if (team) {
        const int k_wide = split_wide_count(h, nq);
        if (k_wide > 0) return launch_split(h, k_wide, algo, mode, nq, d_q, d_status, d_g_goal, d_path_len, d_settled, d_paths, path_cap, d_s5);
        const int rc = alloc_for_search(h, true, nq);
        if (rc != NSFS_OK) return rc;
        return launch_search_team(h, algo, mode, nq, d_q, d_status, d_g_goal, d_path_len, d_settled, d_paths, path_cap, d_s5);
    }
    const int rc = alloc_for_search(h, false, nq);
    if (rc != NSFS_OK) return rc;
    return launch_search(h, kind, algo, mode, nq, d_q, d_status, d_g_goal, d_path_len, d_settled, d_paths, path_cap, d_s5);
}

// Runs nq queries given as host arrays; kind 0 = plan, 1 = find_better_point.  Retries open-list overflows once
// with the provable upper bound on the open-list length (every directed edge is relaxed at most once => 8N+1).
static int run_search_host(nsfs_planner* h, int kind, int algo, int mode, int nq, const int32_t* q_host,
                           int32_t* status, double* g_goal, int32_t* path_len, long long* settled,
                           int32_t* paths, int path_cap, nsfs_stats* st) {
    const int qw = kind == 1 ? 2 : 4;
    const long long launches0 = h->launches;
    const size_t o_q = 0;
    const size_t o_status = align_up(o_q + (size_t)nq * qw * 4, 256);
    const size_t o_g = align_up(o_status + (size_t)nq * 4, 256);
    const size_t o_len = align_up(o_g + (size_t)nq * 8, 256);
    const size_t o_set = align_up(o_len + (size_t)nq * 4, 256);
    const size_t o_s5 = align_up(o_set + (size_t)nq * 8, 256);
    const size_t o_paths = align_up(o_s5 + 5 * 8, 256);
    const size_t path_elems = kind == 1 ? (size_t)nq * 2 : (paths ? (size_t)nq * (size_t)path_cap * 2 : 0);
    int rc = ensure_scratch(h, o_paths + path_elems * 4 + 256);
    if (rc != NSFS_OK) return rc;
    char* d = (char*)h->d_scratch;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(d + o_q, q_host, (size_t)nq * qw * 4, cudaMemcpyHostToDevice, h->stream));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(d + o_s5, 0, 5 * 8, h->stream));
    if (path_elems) NSFS_CUDA_TRY(h, cudaMemsetAsync(d + o_paths, 0, path_elems * 4, h->stream));   // cells beyond path_len read as 0
    const bool team = use_team_kernel(h, kind, algo, nq);
    // find_better_point runs with its state in shared memory unless the wide kernel is forced (the re-run path below)
    const bool small_fb = kind == 1 && h->opt_search_kernel != 1 && h->W >= 3;
    if (!small_fb && (rc = alloc_for_search(h, team, nq)) != NSFS_OK) return rc;      // outside the timed region
    Timer tm(h);
    if (small_fb)
        rc = launch_fallback_small(h, nq, (const int32_t*)(d + o_q), (int32_t*)(d + o_status), (long long*)(d + o_set),
                                   (int32_t*)(d + o_paths), (long long*)(d + o_s5));
    else
        rc = launch_any_search(h, team, kind, algo, mode, nq, (const int32_t*)(d + o_q), (int32_t*)(d + o_status), (double*)(d + o_g),
                               (int32_t*)(d + o_len), (long long*)(d + o_set), path_elems ? (int32_t*)(d + o_paths) : nullptr, path_cap,
                               (long long*)(d + o_s5));
    if (rc != NSFS_OK) return rc;
    const float ms = tm.stop_sync();
    NSFS_CUDA_TRY(h, cudaGetLastError());
    std::vector<int32_t> stat_h(nq);
    long long s5[5];
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(stat_h.data(), d + o_status, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(s5, d + o_s5, sizeof(s5), cudaMemcpyDeviceToHost, h->stream));
    if (g_goal) NSFS_CUDA_TRY(h, cudaMemcpyAsync(g_goal, d + o_g, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    if (path_len) NSFS_CUDA_TRY(h, cudaMemcpyAsync(path_len, d + o_len, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    if (settled) NSFS_CUDA_TRY(h, cudaMemcpyAsync(settled, d + o_set, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    if (paths && path_elems) NSFS_CUDA_TRY(h, cudaMemcpyAsync(paths, d + o_paths, path_elems * 4, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (status) memcpy(status, stat_h.data(), (size_t)nq * 4);
    fill_stats(st, s5, h->launches - launches0, ms, tm.main_kernel_ms());

    // ---- open-list overflow: redo those queries one at a time with the worst-case capacity ----------------
    long long worst = 8 * h->N + 16;
    if (worst > (1ll << 31) - 64) worst = (1ll << 31) - 64;      // search_alloc clamps to 32-bit heap positions
    for (int q = 0; q < nq; ++q) {
        if (stat_h[q] != NSFS_E_HEAP || (!team && !small_fb && h->search.heap_cap >= worst)) continue;
        const long long keep_cap = h->opt_heap_cap, keep_slots = h->opt_max_slots, keep_kernel = h->opt_search_kernel;
        h->opt_heap_cap = worst; h->opt_max_slots = 1; h->opt_search_kernel = 1;
        search_free(h);
        int32_t s1 = 0; double g1 = 0; int32_t l1 = 0; long long e1 = 0;
        std::vector<int32_t> p1(paths ? (size_t)path_cap * 2 : (kind == 1 ? 2 : 0));
        rc = run_search_host(h, kind, algo, mode, 1, q_host + (size_t)q * qw, &s1, &g1, &l1, &e1, p1.empty() ? nullptr : p1.data(),
                             path_cap, nullptr);
        h->opt_heap_cap = keep_cap; h->opt_max_slots = keep_slots; h->opt_search_kernel = keep_kernel;
        search_free(h);
        if (rc != NSFS_OK) return rc;
        if (status) status[q] = s1;
        if (g_goal) g_goal[q] = g1;
        if (path_len) path_len[q] = l1;
        if (settled) settled[q] = e1;
        if (paths && !p1.empty()) memcpy(paths + (size_t)q * (kind == 1 ? 2 : (size_t)path_cap * 2), p1.data(), p1.size() * 4);
    }
    return NSFS_OK;
}

static int field_read_counts(nsfs_planner* h, nsfs_stats* st, long long launches0, float ms, float main_ms);
int nsfs_field_device(nsfs_t* h, int si, int sj, const float** d_g, const uint8_t** d_pred, nsfs_stats* st);
int nsfs_field_path(nsfs_t* h, int gi, int gj, int32_t* path_ij, int cap, int* n_path, double* cost);

int nsfs_plan(nsfs_t* h, int algo, int cost_mode, int si, int sj, int gi, int gj, int32_t* path_ij, int cap, int* n_path,
              double* g_goal, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (si < 0 || si >= h->H || sj < 0 || sj >= h->W || cap < 0) return NSFS_E_ARG;
    h->last_plan_kind = 0;
    if (algo == NSFS_DJIKSTRA_FIELD) {
        nsfs_stats fs;
        rc = nsfs_field_device(h, si, sj, nullptr, nullptr, &fs);
        if (rc != NSFS_OK) return rc;
        int n = 0; double cost = 0;
        rc = nsfs_field_path(h, gi, gj, path_ij, path_ij ? cap : 0, &n, &cost);
        h->last_plan_kind = 2; h->last_ks = (long long)gi; h->last_kg = (long long)gj;     // kind 2 keeps the goal as (i, j)
        if (n_path) *n_path = n;
        if (g_goal) *g_goal = cost;
        if (st) { *st = fs; st->launches += 1; }
        if (rc == NSFS_E_TRUNC && !path_ij) rc = NSFS_OK;
        return rc;
    }
    if (algo != NSFS_ASTAR && algo != NSFS_DJIKSTRA && algo != NSFS_THETASTAR) return NSFS_E_ARG;
    const int32_t q[4] = {si, sj, gi, gj};
    int32_t status = 0, len = 0;
    long long cap_ll = cap;
    if (cap_ll > h->N + 1) cap_ll = h->N + 1;
    rc = run_search_host(h, 0, algo, cost_mode, 1, q, &status, g_goal, &len, nullptr, path_ij, path_ij ? (int)cap_ll : 0, st);
    if (rc != NSFS_OK) return rc;
    if (n_path) *n_path = len;
    if (status != NSFS_OK) return status;
    if (len > 1 && h->search.cells && h->search.slots >= 1 && gi >= 0 && gi < h->H && gj >= 0 && gj < h->W) {
        // the wide kernel ran this query in slot 0 and its parent store is still there: nsfs_last_path can walk it again
        h->last_plan_kind = 1; h->last_ks = (long long)si * h->W + sj; h->last_kg = (long long)gi * h->W + gj;
    }
    if (path_ij && len > cap) return NSFS_E_TRUNC;
    return NSFS_OK;
}

int nsfs_last_path(nsfs_t* h, int32_t* path_ij, int cap, int* n_path) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!path_ij || cap <= 0) return NSFS_E_ARG;
    if (h->last_plan_kind == 2) return nsfs_field_path(h, (int)h->last_ks, (int)h->last_kg, path_ij, cap, n_path, nullptr);
    if (h->last_plan_kind != 1) return NSFS_E_NOFIELD;
    long long dcap = cap;
    if (dcap > h->N + 1) dcap = h->N + 1;
    const size_t o_path = 256;
    if ((rc = ensure_scratch(h, o_path + (size_t)dcap * 8 + 256)) != NSFS_OK) return rc;
    char* d = (char*)h->d_scratch;
    if ((rc = launch_path_walk(h, h->last_ks, h->last_kg, (int32_t*)(d + o_path), (int)dcap, (int*)d)) != NSFS_OK) return rc;
    int n = 0;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(&n, d, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (n_path) *n_path = n;
    const long long ncopy = n < dcap ? n : dcap;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(path_ij, d + o_path, (size_t)ncopy * 8, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return n > cap ? NSFS_E_TRUNC : NSFS_OK;
}

int nsfs_pin_host(nsfs_t* h, const void* p, size_t bytes) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    if (!p || !bytes) return NSFS_E_ARG;
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) return set_cuda_error(h, cudaGetLastError(), "cudaHostRegister");
    return NSFS_OK;
}

int nsfs_unpin_host(nsfs_t* h, const void* p) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    if (!p) return NSFS_E_ARG;
    if (cudaHostUnregister(const_cast<void*>(p)) != cudaSuccess) return set_cuda_error(h, cudaGetLastError(), "cudaHostUnregister");
    return NSFS_OK;
}

int nsfs_plan_batch(nsfs_t* h, int algo, int cost_mode, int nq, const int32_t* sg4, int32_t* status, double* g_goal,
                    int32_t* path_len, long long* settled, int32_t* paths_ij, int path_cap, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (nq < 0 || (nq && !sg4) || path_cap < 0) return NSFS_E_ARG;
    if (algo != NSFS_ASTAR && algo != NSFS_DJIKSTRA && algo != NSFS_THETASTAR) return NSFS_E_ARG;
    if (!nq) return NSFS_OK;
    return run_search_host(h, 0, algo, cost_mode, nq, sg4, status, g_goal, path_len, settled, paths_ij, path_cap, st);
}

int nsfs_plan_batch_device(nsfs_t* h, int algo, int cost_mode, int nq, const int32_t* d_sg4, int32_t* d_status, double* d_g_goal,
                           int32_t* d_path_len, long long* d_settled, int32_t* d_paths_ij, int path_cap, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (nq < 0 || (nq && !d_sg4) || path_cap < 0) return NSFS_E_ARG;
    if (algo != NSFS_ASTAR && algo != NSFS_DJIKSTRA && algo != NSFS_THETASTAR) return NSFS_E_ARG;
    if (!nq) return NSFS_OK;
    const long long launches0 = h->launches;
    const bool team = use_team_kernel(h, 0, algo, nq);
    if ((rc = alloc_for_search(h, team, nq)) != NSFS_OK) return rc;      // outside the timed region
    if ((rc = ensure_scratch(h, 256)) != NSFS_OK) return rc;
    long long* d_s5 = (long long*)h->d_scratch;
    NSFS_CUDA_TRY(h, cudaMemsetAsync(d_s5, 0, 5 * 8, h->stream));
    Timer tm(h);
    rc = launch_any_search(h, team, 0, algo, cost_mode, nq, d_sg4, d_status, d_g_goal, d_path_len, d_settled, d_paths_ij, path_cap, d_s5);
    if (rc != NSFS_OK) return rc;
    const float ms = tm.stop_sync();
    const float main_ms = tm.main_kernel_ms();
    long long s5[5];
    NSFS_CUDA_TRY(h, cudaMemcpy(s5, d_s5, sizeof(s5), cudaMemcpyDeviceToHost));
    fill_stats(st, s5, h->launches - launches0, ms, main_ms);
    long long worst = 8 * h->N + 16;
    if (worst > (1ll << 31) - 64) worst = (1ll << 31) - 64;
    if (d_status && (team || h->search.heap_cap < worst)) {
        // queries that could not be finished exactly (open list beyond its slot; on the batch kernel also 16-bit step counts / sort
        // keys): re-run each on the wide kernel with the worst-case capacity, results written into the same device arrays
        // (the host-pointer entry point does the same, so both return the same statuses)
        std::vector<int32_t> stat_h(nq);
        NSFS_CUDA_TRY(h, cudaMemcpy(stat_h.data(), d_status, (size_t)nq * 4, cudaMemcpyDeviceToHost));
        bool any = false;
        for (int q = 0; q < nq; ++q) any = any || stat_h[q] == NSFS_E_HEAP;
        if (any) {
            const long long keep_cap = h->opt_heap_cap, keep_slots = h->opt_max_slots;
            h->opt_heap_cap = worst; h->opt_max_slots = 1;
            search_free(h);
            for (int q = 0; q < nq && rc == NSFS_OK; ++q) {
                if (stat_h[q] != NSFS_E_HEAP) continue;
                rc = launch_search(h, 0, algo, cost_mode, 1, d_sg4 + 4 * (size_t)q, d_status + q, d_g_goal ? d_g_goal + q : nullptr,
                                   d_path_len ? d_path_len + q : nullptr, d_settled ? d_settled + q : nullptr,
                                   d_paths_ij ? d_paths_ij + (size_t)q * path_cap * 2 : nullptr, path_cap, nullptr);
            }
            if (rc == NSFS_OK) NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            h->opt_heap_cap = keep_cap; h->opt_max_slots = keep_slots;
            search_free(h);
        }
    }
    return rc;
}

int nsfs_find_better_point_batch(nsfs_t* h, int n, const int32_t* bad_ij, int32_t* out_ij, int32_t* status, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (n < 0 || (n && (!bad_ij || !out_ij))) return NSFS_E_ARG;
    if (!n) return NSFS_OK;
    for (int t = 0; t < n; ++t)
        if (bad_ij[2 * t] < 0 || bad_ij[2 * t] >= h->H || bad_ij[2 * t + 1] < 0 || bad_ij[2 * t + 1] >= h->W) return NSFS_E_ARG;
    return run_search_host(h, 1, NSFS_DJIKSTRA, NSFS_MODE_G, n, bad_ij, status, nullptr, nullptr, nullptr, out_ij, 1, st);
}

int nsfs_find_better_point(nsfs_t* h, int bi, int bj, int* fi, int* fj, nsfs_stats* st) {
    const int32_t q[2] = {bi, bj};
    int32_t out[2] = {bi, bj}, status = 0;
    int rc = nsfs_find_better_point_batch(h, 1, q, out, &status, st);
    if (rc != NSFS_OK) return rc;
    if (fi) *fi = out[0];
    if (fj) *fj = out[1];
    return status;
}

// ---- field ------------------------------------------------------------------------------------------------
int nsfs_field_device(nsfs_t* h, int si, int sj, const float** d_g, const uint8_t** d_pred, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (si < 0 || si >= h->H || sj < 0 || sj >= h->W) return NSFS_E_ARG;
    const long long launches0 = h->launches;
    Timer tm(h);
    rc = launch_field(h, si, sj, st);
    if (rc != NSFS_OK) return rc;
    const float ms = tm.stop_sync();
    NSFS_CUDA_TRY(h, cudaGetLastError());
    h->field.finished = true;
    if (d_g) *d_g = reinterpret_cast<const float*>(h->field.g);
    if (d_pred) *d_pred = h->field.pred;
    h->field.begun = false;                       // finished: the plane holds fp32 cost units now
    if ((rc = field_read_counts(h, st, launches0, ms, tm.main_kernel_ms())) != NSFS_OK) return rc;
    h->field.valid = true;
    return NSFS_OK;
}

// ---- row-partitioned field: the same solve in caller-driven phases ----------------------------------------
int nsfs_field_begin(nsfs_t* h, int si, int sj) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (si >= 0 && (si >= h->H || sj < 0 || sj >= h->W)) return NSFS_E_ARG;
    if ((rc = launch_field_begin(h, si, sj)) != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->field.begun = true;
    h->field.finished = false;
    return NSFS_OK;
}

int nsfs_field_inject_rows(nsfs_t* h, int row0, int nrows, const float* d_vals, int* n_improved) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun) return NSFS_E_NOFIELD;
    if (row0 < 0 || nrows < 0 || row0 + nrows > h->H || (nrows && !d_vals)) return NSFS_E_ARG;
    if ((rc = ensure_scratch(h, 256)) != NSFS_OK) return rc;
    int* d_cnt = (int*)h->d_scratch;
    NSFS_CUDA_TRY(h, cudaMemsetAsync(d_cnt, 0, sizeof(int), h->stream));
    h->field.valid = false;
    if ((rc = launch_field_inject(h, row0, nrows, d_vals, d_cnt)) != NSFS_OK) return rc;
    int cnt = 0;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (n_improved) *n_improved = cnt;
    return NSFS_OK;
}

static int field_read_counts(nsfs_planner* h, nsfs_stats* st, long long launches0, float ms, float main_ms) {
    int counts[16];
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(counts, h->field.counts, sizeof(counts), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (st) {
        memset(st, 0, sizeof(*st));
        st->rounds = counts[2]; st->tile_visits = counts[3]; st->expansions = counts[5];
        st->launches = h->launches - launches0; st->gpu_ms = ms; st->main_kernel_ms = main_ms;
    }
    if (counts[4] & 2) { snprintf(h->cuda_error, sizeof(h->cuda_error), "field_solve_async_kernel: scheduler bail-out"); return NSFS_E_CUDA; }
    const uint32_t kb = (uint32_t)counts[6];            // smallest key still pending (Q17.15; kNoKey = none)
    h->field.min_pending = kb == kNoKey ? INFINITY : (float)((double)kb / (double)kQOne);
    if (counts[4] & 1) return NSFS_E_RANGE; // a reached cell's expansion throws in the reference
    return NSFS_OK;
}

int nsfs_field_relax_until(nsfs_t* h, float cap, float* min_pending, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun) return NSFS_E_NOFIELD;
    if (!(cap >= 0.f)) return NSFS_E_ARG;
    const long long launches0 = h->launches;
    Timer tm(h);
    if ((rc = launch_field_relax(h, cap)) != NSFS_OK) return rc;
    const float ms = tm.stop_sync();
    NSFS_CUDA_TRY(h, cudaGetLastError());
    rc = field_read_counts(h, st, launches0, ms, tm.main_kernel_ms());
    if (min_pending) *min_pending = h->field.min_pending;
    return rc;
}

// ---- peer-to-peer shards -------------------------------------------------------------------------------------
int nsfs_field_peer_export(nsfs_t* h, nsfs_peer_info* out) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!out) return NSFS_E_ARG;
    if ((rc = field_alloc(h)) != NSFS_OK) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "nsfs_peer_info carries 64-byte IPC handles");
    cudaIpcMemHandle_t hg, hk, hc;
    NSFS_CUDA_TRY(h, cudaIpcGetMemHandle(&hg, h->field.g));
    NSFS_CUDA_TRY(h, cudaIpcGetMemHandle(&hk, h->field.tile_key));
    NSFS_CUDA_TRY(h, cudaIpcGetMemHandle(&hc, h->field.counts));
    memcpy(out->g, &hg, 64); memcpy(out->key, &hk, 64); memcpy(out->counts, &hc, 64);
    out->rows = h->H; out->W = h->W; out->tiles_x = h->field.tiles_x; out->tiles_y = h->field.tiles_y;
    return NSFS_OK;
}

static int ipc_open(nsfs_planner* h, const unsigned char* raw, void** out, int slot) {
    cudaIpcMemHandle_t hh;
    memcpy(&hh, raw, 64);
    if (h->field.ipc_opened[slot]) { cudaIpcCloseMemHandle(h->field.ipc_opened[slot]); h->field.ipc_opened[slot] = nullptr; }
    NSFS_CUDA_TRY(h, cudaIpcOpenMemHandle(out, hh, cudaIpcMemLazyEnablePeerAccess));
    h->field.ipc_opened[slot] = *out;
    return NSFS_OK;
}

int nsfs_field_peer_attach(nsfs_t* h, int side, const nsfs_peer_info* nb, int row_off) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (side < 0 || side > 1 || !nb || nb->W != h->W || nb->tiles_x != (h->W + 63) / 64) return NSFS_E_ARG;
    if ((rc = field_alloc(h)) != NSFS_OK) return rc;
    void *pg = nullptr, *pk = nullptr;
    if ((rc = ipc_open(h, nb->g, &pg, 2 * side)) != NSFS_OK) return rc;
    if ((rc = ipc_open(h, nb->key, &pk, 2 * side + 1)) != NSFS_OK) return rc;
    h->field.peer_g[side] = (float*)pg; h->field.peer_key[side] = (uint32_t*)pk;
    h->field.peer_row_off[side] = row_off; h->field.peer_tiles_y[side] = nb->tiles_y; h->field.peer_rows[side] = nb->rows;
    return NSFS_OK;
}

int nsfs_field_peer_counter(nsfs_t* h, const nsfs_peer_info* owner) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if ((rc = field_alloc(h)) != NSFS_OK) return rc;
    if (!owner) {
        if (h->field.ipc_opened[4]) { cudaIpcCloseMemHandle(h->field.ipc_opened[4]); h->field.ipc_opened[4] = nullptr; }
        h->field.peer_counts = nullptr;
        return NSFS_OK;
    }
    void* pc = nullptr;
    if ((rc = ipc_open(h, owner->counts, &pc, 4)) != NSFS_OK) return rc;
    h->field.peer_counts = (int*)pc;
    return NSFS_OK;
}

int nsfs_field_peer_rows(nsfs_t* h, int own_lo, int own_hi) {
    if (!h || own_lo < 0 || own_hi > h->H || own_hi - own_lo < 2) return NSFS_E_ARG;
    h->field.own_lo = own_lo; h->field.own_hi = own_hi;
    return NSFS_OK;
}

int nsfs_field_peer_detach(nsfs_t* h) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 5; ++k)
        if (h->field.ipc_opened[k]) { cudaIpcCloseMemHandle(h->field.ipc_opened[k]); h->field.ipc_opened[k] = nullptr; }
    for (int sd = 0; sd < 2; ++sd) { h->field.peer_g[sd] = nullptr; h->field.peer_key[sd] = nullptr; }
    h->field.peer_counts = nullptr;
    return NSFS_OK;
}

int nsfs_field_pending_count(nsfs_t* h, int* n_pending) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun || !n_pending) return NSFS_E_NOFIELD;
    if ((rc = launch_field_count_pending(h)) != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(n_pending, h->field.counts + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NSFS_OK;
}

int nsfs_field_set_counter(nsfs_t* h, int total) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun) return NSFS_E_NOFIELD;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(h->field.counts + 1, &total, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NSFS_OK;
}

int nsfs_field_relax_p2p(nsfs_t* h, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun) return NSFS_E_NOFIELD;
    const long long launches0 = h->launches;
    Timer tm(h);
    if ((rc = launch_field_relax_p2p(h)) != NSFS_OK) return rc;
    const float ms = tm.stop_sync();
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return field_read_counts(h, st, launches0, ms, tm.main_kernel_ms());
}

int nsfs_field_min_pending(nsfs_t* h, float* min_pending) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun) return NSFS_E_NOFIELD;
    if ((rc = launch_field_min_pending(h)) != NSFS_OK) return rc;
    uint32_t kb = 0;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(&kb, h->field.counts + 6, sizeof(kb), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->field.min_pending = kb == kNoKey ? INFINITY : (float)((double)kb / (double)kQOne);
    if (min_pending) *min_pending = h->field.min_pending;
    return NSFS_OK;
}

int nsfs_field_relax(nsfs_t* h, nsfs_stats* st) { return nsfs_field_relax_until(h, 3.4e38f, nullptr, st); }

int nsfs_field_finish(nsfs_t* h, nsfs_stats* st) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun) return NSFS_E_NOFIELD;
    const long long launches0 = h->launches;
    Timer tm(h);
    if ((rc = launch_field_finish(h)) != NSFS_OK) return rc;
    const float ms = tm.stop_sync();
    NSFS_CUDA_TRY(h, cudaGetLastError());
    h->field.begun = false;                       // the plane is fp32 cost units from here on: begin again before relaxing
    rc = field_read_counts(h, st, launches0, ms, 0.f);
    if (rc == NSFS_OK) h->field.valid = true;
    h->field.finished = true;
    return rc;
}

int nsfs_field_pointers(nsfs_t* h, const float** d_g, const uint8_t** d_pred) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.begun && !h->field.finished) return NSFS_E_NOFIELD;
    if (d_g) *d_g = reinterpret_cast<const float*>(h->field.g);
    if (d_pred) *d_pred = h->field.pred;
    return NSFS_OK;
}

int nsfs_field(nsfs_t* h, int si, int sj, float* g, uint8_t* pred, nsfs_stats* st) {
    int rc = nsfs_field_device(h, si, sj, nullptr, nullptr, st);
    if (rc != NSFS_OK) return rc;
    if (g) NSFS_CUDA_TRY(h, cudaMemcpyAsync(g, h->field.g, (size_t)h->N * 4, cudaMemcpyDeviceToHost, h->stream));
    if (pred) NSFS_CUDA_TRY(h, cudaMemcpyAsync(pred, h->field.pred, (size_t)h->N, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NSFS_OK;
}

int nsfs_field_path(nsfs_t* h, int gi, int gj, int32_t* path_ij, int cap, int* n_path, double* cost) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (!h->field.valid) return NSFS_E_NOFIELD;
    if (cap < 0) return NSFS_E_ARG;
    long long dcap = cap;
    if (dcap > h->N + 1) dcap = h->N + 1;
    if (!path_ij) dcap = 0;
    const size_t o_path = 256;
    if ((rc = ensure_scratch(h, o_path + (size_t)dcap * 8 + 256)) != NSFS_OK) return rc;
    char* d = (char*)h->d_scratch;
    // header: int n @0, int status @4, double cost @8
    if ((rc = launch_field_path(h, gi, gj, (int32_t*)(d + o_path), (int)dcap, (int*)d, (double*)(d + 8), (int*)(d + 4))) != NSFS_OK) return rc;
    struct { int n, status; double cost; } hd;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(&hd, d, sizeof(hd), cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (n_path) *n_path = hd.n;
    if (cost) *cost = hd.cost;
    const long long ncopy = hd.n < dcap ? hd.n : dcap;
    if (path_ij && ncopy > 0) {
        NSFS_CUDA_TRY(h, cudaMemcpyAsync(path_ij, d + o_path, (size_t)ncopy * 8, cudaMemcpyDeviceToHost, h->stream));
        NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    return hd.status;
}

// ---- line of sight -------------------------------------------------------------------------------------------
int nsfs_los_batch(nsfs_t* h, int n, const int32_t* seg4, int8_t* out) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (n < 0 || (n && (!seg4 || !out))) return NSFS_E_ARG;
    if (!n) return NSFS_OK;
    const size_t in_b = align_up((size_t)n * 16, 256);
    if ((rc = ensure_scratch(h, in_b + n)) != NSFS_OK) return rc;
    char* d = (char*)h->d_scratch;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(d, seg4, (size_t)n * 16, cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_los_batch(h, n, (const int32_t*)d, (int8_t*)(d + in_b))) != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(out, d + in_b, n, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NSFS_OK;
}

int nsfs_any_angle(nsfs_t* h, int n, const int32_t* pts_ij, uint8_t* keep) {
    int rc = enter(h, true);
    if (rc != NSFS_OK) return rc;
    if (n < 1 || !pts_ij || !keep) return NSFS_E_ARG;
    const size_t o_pts = 256, o_keep = align_up(o_pts + (size_t)n * 8, 256);
    if ((rc = ensure_scratch(h, o_keep + n + 256)) != NSFS_OK) return rc;
    char* d = (char*)h->d_scratch;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(d + o_pts, pts_ij, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_any_angle(h, n, (const int32_t*)(d + o_pts), (uint8_t*)(d + o_keep))) != NSFS_OK) return rc;
    int status = 0;
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(&status, d, 4, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(keep, d + o_keep, n, cudaMemcpyDeviceToHost, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return status;
}

// ---- service ---------------------------------------------------------------------------------------------------
int nsfs_sync(nsfs_t* h) {
    int rc = enter(h, false);
    if (rc != NSFS_OK) return rc;
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return NSFS_OK;
}

long long nsfs_kernel_launches(const nsfs_t* h) { return h ? h->launches : 0; }
void* nsfs_stream(nsfs_t* h) { return h ? (void*)h->stream : nullptr; }

int nsfs_set_option(nsfs_t* h, const char* key, long long value) {
    if (!h || !key) return NSFS_E_ARG;
    if (!strcmp(key, "max_slots")) { h->opt_max_slots = value; search_free(h); team_free(h); return NSFS_OK; }
    if (!strcmp(key, "heap_cap")) { h->opt_heap_cap = value; search_free(h); return NSFS_OK; }
    if (!strcmp(key, "team_heap_cap")) { h->opt_team_heap_cap = value; team_free(h); return NSFS_OK; }
    if (!strcmp(key, "team_entry_bytes")) { h->opt_team_entry_bytes = value; return NSFS_OK; }
    if (!strcmp(key, "team_cell_bytes")) { h->opt_team_cell_bytes = value; team_free(h); return NSFS_OK; }
    if (!strcmp(key, "search_kernel")) { h->opt_search_kernel = value; return NSFS_OK; }
    if (!strcmp(key, "search_order")) { h->opt_search_order = value; return NSFS_OK; }
    if (!strcmp(key, "split_wide")) { h->opt_split_wide = value; return NSFS_OK; }
    if (!strcmp(key, "graph")) {                            // 0 reference quirk graph, 1 textbook 8-connected grid: rebuild the edge masks
        if (value != 0 && value != 1) return NSFS_E_ARG;
        if (h->opt_graph == value) return NSFS_OK;
        h->opt_graph = value;
        h->field.valid = false; h->field.begun = false;
        if (!h->have_map) return NSFS_OK;
        cudaSetDevice(h->device);
        int rc = launch_build_masks(h);
        if (rc != NSFS_OK) return rc;
        NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        return NSFS_OK;
    }
    if (!strcmp(key, "field_inner")) { h->opt_field_inner = value; return NSFS_OK; }
    if (!strcmp(key, "field_delta")) { h->opt_field_delta = value; return NSFS_OK; }
    if (!strcmp(key, "field_mode")) { h->opt_field_mode = value; return NSFS_OK; }
    if (!strcmp(key, "field_tiles_per_sm")) { h->opt_field_tiles_per_sm = value; return NSFS_OK; }
    if (!strcmp(key, "field_gate")) { h->opt_field_gate = value; return NSFS_OK; }
    if (!strcmp(key, "field_idle_ns")) { h->opt_field_idle_ns = value; return NSFS_OK; }
    if (!strcmp(key, "field_fence")) { h->opt_field_fence = value; return NSFS_OK; }
    if (!strcmp(key, "field_merge")) { h->opt_field_merge = value; return NSFS_OK; }
    if (!strcmp(key, "field_sched_ns")) { h->opt_field_sched_ns = value; return NSFS_OK; }
    if (!strcmp(key, "field_wait_ns")) { h->opt_field_wait_ns = value; return NSFS_OK; }
    if (!strcmp(key, "active_row_lo")) { h->opt_active_row_lo = value; h->have_map = false; return NSFS_OK; }   // masks depend on it: set the map again
    if (!strcmp(key, "active_row_hi")) { h->opt_active_row_hi = value; h->have_map = false; return NSFS_OK; }
    if (!strcmp(key, "count_row_lo")) { h->opt_count_row_lo = value; return NSFS_OK; }
    if (!strcmp(key, "count_row_hi")) { h->opt_count_row_hi = value; return NSFS_OK; }
    return NSFS_E_ARG;
}

const char* nsfs_strerror(int code) {
    switch (code) {
        case NSFS_OK: return "ok";
        case NSFS_E_RANGE: return "the reference would throw std::out_of_range (vector::at in testPos)";
        case NSFS_E_CUDA: return "CUDA error (see nsfs_last_cuda_error); there is no CPU fallback";
        case NSFS_E_TRUNC: return "output buffer too small";
        case NSFS_E_ARG: return "bad argument";
        case NSFS_E_NOMEM: return "out of memory";
        case NSFS_E_HEAP: return "open list exceeded its device capacity";
        case NSFS_E_NOMAP: return "no map set";
        case NSFS_E_NOFIELD: return "no cost-to-go field computed on this handle";
        default: return "unknown nsfs error";
    }
}

const char* nsfs_last_cuda_error(const nsfs_t* h) { return h ? h->cuda_error : ""; }
const char* nsfs_version(void) { return "nsfs-b200 0.1 (sm_100a)"; }

}  // extern "C"
