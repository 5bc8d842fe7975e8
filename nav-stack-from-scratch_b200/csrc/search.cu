// Exact-order search: Astar / Djikstra / ThetaStar plan() and the Djikstra fallback, one query per warp.
//
// The reference's A* (and its fallback search) is order dependent: keys are truncated to int in the open list
// (grid_planner_core.h:43,45), the heuristic is inadmissible (bot_utils.cpp:47), closed cells keep being
// re-labelled (astar.cpp:116-121), and ties between equal int keys are broken only by the mechanics of
// libstdc++'s binary heap.  Matching its paths, costs and fallback cells therefore means replaying its pop
// order exactly (SURVEY.md "three facts" #3).  Each warp replays one query:
//   - the open list is a device binary heap that performs std::push_heap / std::pop_heap step for step
//     (bits/stl_heap.h __push_heap / __adjust_heap: the hole always sinks to a leaf preferring the right child
//     unless it sorts after the left one, then the displaced last element is sifted up);
//   - entries are one 64-bit word  f:18 | g:17 | key:29  so a comparator is a shift and a compare;
//   - the node store (grid_planner_core.h:26-33) is one 16-byte record per cell {g (fp64), parent key,
//     epoch<<1 | visited}: the fp64 sums are the reference's own sums in the reference's own order, so g is
//     bit-identical, and the epoch stamp replaces the O(N) reset of astar.cpp:37-42;
//   - lanes 0..7 evaluate the 8 neighbours of the expanded cell together; pushes then happen in NB_LUT order.
// Batches run thousands of such warps at once (nsfs_plan_batch); queries are handed out by an atomic counter.
#include "nsfs_internal.cuh"

namespace nsfs {

constexpr int kAlgoFallback = 4;         // internal: Djikstra::find_better_point
constexpr int SEARCH_WARPS_PER_BLOCK = 4;
constexpr int kKeyBits = 29, kGBits = 17;
constexpr unsigned long long kKeyMask = (1ull << kKeyBits) - 1ull;

struct SearchParams {
    const uint32_t* free_bits;
    const uint32_t* infl_bits;
    const uint16_t* out_mask;
    uint4* cells;                 // [slots][N]
    unsigned long long* heap;     // [slots][heap_cap]
    uint32_t* epoch;              // [slots]
    int* work_counter;
    int H, W;
    long long N;
    long long heap_cap;
    int slots;
    int mode;
    int nq;
    const int32_t* q;             // plan: {si, sj, gi, gj} per query; fallback: {bi, bj}
    int32_t* status;
    double* g_goal;
    int32_t* path_len;
    long long* settled;
    int32_t* paths;               // optional [nq][path_cap][2]; fallback: [nq][2] result cell
    int path_cap;
    long long* stats5;            // pops, expansions, pushes, heap_peak (max), los_checks
};

__device__ __forceinline__ unsigned long long entry_prio(unsigned long long e, int mode) {
    // f_comp / g_comp / f_g_comp, grid_planner_core.cpp:55-75
    if (mode == NSFS_MODE_F) return e >> (kKeyBits + kGBits);
    if (mode == NSFS_MODE_G) return (e >> kKeyBits) & ((1ull << kGBits) - 1ull);
    return e >> kKeyBits;
}

// std::priority_queue::push: append, then __push_heap (stl_heap.h:135-148)
__device__ __forceinline__ void heap_push(unsigned long long* a, long long& n, unsigned long long v, int mode) {
    long long hole = n++;
    const unsigned long long pv = entry_prio(v, mode);
    while (hole > 0) {
        const long long parent = (hole - 1) >> 1;
        const unsigned long long e = a[parent];
        if (!(entry_prio(e, mode) > pv)) break;
        a[hole] = e;
        hole = parent;
    }
    a[hole] = v;
}

// top() then pop(): __pop_heap -> __adjust_heap (stl_heap.h:224-262) -> __push_heap, then pop_back
__device__ __forceinline__ unsigned long long heap_pop(unsigned long long* a, long long& n, int mode) {
    const unsigned long long top = a[0];
    const long long len = n - 1;
    if (len > 0) {
        const unsigned long long v = a[len];
        long long hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            const unsigned long long r = a[child], l = a[child - 1];
            unsigned long long pick = r;
            if (entry_prio(r, mode) > entry_prio(l, mode)) { child--; pick = l; }
            a[hole] = pick;
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[hole] = a[child - 1];
            hole = child - 1;
        }
        const unsigned long long pv = entry_prio(v, mode);
        while (hole > 0) {
            const long long parent = (hole - 1) >> 1;
            const unsigned long long e = a[parent];
            if (!(entry_prio(e, mode) > pv)) break;
            a[hole] = e;
            hole = parent;
        }
        a[hole] = v;
    }
    n = len;
    return top;
}

__device__ __forceinline__ double cell_g(const uint4& c) { return __hiloint2double((int)c.y, (int)c.x); }
__device__ __forceinline__ uint4 make_cell(double g, uint32_t parent, uint32_t tag) {
    return make_uint4((uint32_t)__double2loint(g), (uint32_t)__double2hiint(g), parent, tag);
}

// dist_oct(Index, Index), bot_utils.cpp:44-65 with the line-47 slip: both deltas are taken against the ROW.
__device__ __forceinline__ double h_oct(int i, int gi, int gj) {
    const double ax = fabs((double)gi - (double)i), ay = fabs((double)gj - (double)i);
    const double ord = ax > ay ? ay : ax, card = ax > ay ? ax - ay : ay - ax;
    return __dadd_rn(__dmul_rn(ord, kSqrt2D), card);          // no FMA: the host computes mul then add
}
// dist_euc(Index, Index), bot_utils.cpp:72-87
__device__ __forceinline__ double d_euc(int ai, int aj, int bi, int bj) {
    const double dx = (double)bi - (double)ai, dy = (double)bj - (double)aj;
    return sqrt(dx * dx + dy * dy);                            // both products and their sum are exact integers
}

template <int ALGO>
__device__ __forceinline__ double heuristic(int i, int j, int gi, int gj) {
    if (ALGO == NSFS_ASTAR) return h_oct(i, gi, gj);
    if (ALGO == NSFS_THETASTAR) return d_euc(i, j, gi, gj);
    return 0.0;
}

// pushToOpenList, grid_planner_core.cpp:385-391: f = (mode == "g") ? 0 : g + h, both truncated to int
__device__ __forceinline__ unsigned long long make_entry(double g, double h, long long key, int mode) {
    const double f = (mode == NSFS_MODE_G) ? 0.0 : __dadd_rn(g, h);
    const unsigned long long fi = (unsigned long long)(unsigned)__double2int_rz(f);
    const unsigned long long gi = (unsigned long long)(unsigned)__double2int_rz(g);
    return (fi << (kKeyBits + kGBits)) | (gi << kKeyBits) | (unsigned long long)key;
}

// Warp-parallel isLos (grid_planner_core.cpp:481-496) over the closed form of bresenham_los
// (bot_utils.cpp:137-192): cell t of the ray has l = l0 + t*dl, s = s0 + ds*floor((2t|Ds|+|Dl|)/(2|Dl|)).
// Both end points are in-range cells here, so no ray cell can leave the grid.
__device__ __forceinline__ bool warp_los(const uint32_t* __restrict__ free_bits, int W, int si, int sj, int ti, int tj, int lane) {
    const int Di = ti - si, Dj = tj - sj;
    const bool long_i = abs(Di) > abs(Dj);
    const int Dl = long_i ? Di : Dj, Ds = long_i ? Dj : Di;
    const int l0 = long_i ? si : sj, s0 = long_i ? sj : si;
    const int aDl = abs(Dl), aDs = abs(Ds);
    const int dl = (Dl > 0) - (Dl < 0), ds = (Ds > 0) - (Ds < 0);
    for (int base = 0; base <= aDl; base += 32) {
        const int t = base + lane;
        bool ok = true;
        if (t <= aDl) {
            const int l = l0 + t * dl;
            const int s = s0 + (aDl ? ds * (int)((2ll * t * aDs + aDl) / (2ll * aDl)) : 0);
            const long long key = long_i ? (long long)l * W + s : (long long)s * W + l;
            ok = bit_at(free_bits, key);
        }
        if (!__all_sync(0xffffffffu, ok)) return false;
    }
    return true;
}

template <int ALGO>
__global__ void __launch_bounds__(SEARCH_WARPS_PER_BLOCK * 32) search_kernel(SearchParams p) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SEARCH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (slot >= p.slots) return;
    const int W = p.W, H = p.H, mode = p.mode;
    const long long N = p.N;
    uint4* cells = p.cells + (size_t)slot * (size_t)N;
    unsigned long long* heap = p.heap + (size_t)slot * (size_t)p.heap_cap;
    uint32_t epoch = p.epoch[slot];
    const bool serial = W < 3;      // for W < 3 two directions can alias one key: evaluate them one after another

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= p.nq) break;
        ++epoch;

        int si, sj, gi = -1, gj = -1;
        if (ALGO == kAlgoFallback) { si = p.q[2 * qi]; sj = p.q[2 * qi + 1]; }
        else { si = p.q[4 * qi]; sj = p.q[4 * qi + 1]; gi = p.q[4 * qi + 2]; gj = p.q[4 * qi + 3]; }

        long long n = 0;                       // open-list length (warp-uniform)
        long long pops = 0, expansions = 0, pushes = 0, peak = 0, los_checks = 0;
        int status = NSFS_OK;
        bool found = false;
        long long k_found = -1;
        const long long ks = (long long)si * W + sj;
        const long long kg = (gi >= 0 && gi < H && gj >= 0 && gj < W) ? (long long)gi * W + gj : -1;

        if (si < 0 || si >= H || sj < 0 || sj >= W) {
            status = NSFS_E_ARG;               // the reference indexes nodes[] out of bounds here (undefined behaviour)
        } else {
            if (lane == 0) {
                cells[ks] = make_cell(0.0, (uint32_t)ks, epoch << 1);
                heap_push(heap, n, make_entry(0.0, heuristic<ALGO>(si, sj, gi, gj), ks, mode), mode);
            }
            n = 1; pushes = 1; peak = 1;
            __syncwarp();
        }

        while (status == NSFS_OK && n > 0) {
            // ---- pop (astar.cpp:56) --------------------------------------------------------------------
            unsigned long long top = 0;
            if (lane == 0) top = heap_pop(heap, n, mode);
            top = __shfl_sync(0xffffffffu, top, 0);
            n = __shfl_sync(0xffffffffu, n, 0);
            pops++;
            const long long ku = (long long)(top & kKeyMask);
            const uint4 cu = cells[ku];        // same address in every lane: one broadcast load
            if (cu.w & 1u) continue;           // stale entry (astar.cpp:59-62)
            if (lane == 0) cells[ku].w = cu.w | 1u;
            expansions++;
            const int i = (int)(ku / W), j = (int)(ku - (long long)i * W);
            if (ALGO == kAlgoFallback) {
                if (bit_at(p.free_bits, ku)) { found = true; k_found = ku; break; }     // djikstra.cpp:205-211
            } else if (i == gi && j == gj) { found = true; k_found = ku; break; }       // astar.cpp:68

            const double gu = cell_g(cu);
            const uint32_t om = p.out_mask[ku];
            if (om & kThrowBit) { status = NSFS_E_RANGE; break; }                       // vector::at throws (R6)

            // ---- neighbours (astar.cpp:89-124 / djikstra.cpp:214-267) ------------------------------------
            uint32_t rowok = 0;                // fallback: directions whose row is in range (djikstra.cpp:222-226)
            if (ALGO == kAlgoFallback) {
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    const int ni = i + nb_di(d);
                    if (ni >= 0 && ni < H) rowok |= 1u << d;
                }
            }
            uint32_t pkey = cu.z;              // ThetaStar: parent of the expanded cell
            double gp = 0.0;
            int pi = 0, pj = 0;
            if (ALGO == NSFS_THETASTAR) {
                gp = cell_g(cells[pkey]);
                pi = (int)(pkey / (uint32_t)W); pj = (int)(pkey - (uint32_t)pi * (uint32_t)W);
            }
            uint32_t los_bits = 0;
            if (ALGO == NSFS_THETASTAR) {
                for (int d = 0; d < 8; ++d) {
                    if (!((om >> d) & 1u)) continue;
                    const long long kv = ku + nb_off(d, W);
                    const int vi = (int)(kv / W), vj = (int)(kv - (long long)vi * W);
                    los_checks++;
                    if (warp_los(p.free_bits, W, pi, pj, vi, vj, lane)) los_bits |= 1u << d;
                }
            }

            for (int round = 0; round < (serial ? 8 : 1); ++round) {
                const int d = lane;
                const bool mine = serial ? (lane == round) : (lane < 8);
                bool relax = false;
                long long kv = 0;
                double cand = 0.0;
                uint32_t par = (uint32_t)ku;
                if (mine) {
                    bool considered;
                    if (ALGO == kAlgoFallback) considered = (rowok >> d) & 1u;
                    else considered = (om >> d) & 1u;
                    if (considered) {
                        kv = ku + nb_off(d, W);
                        const uint4 cv = cells[kv];
                        const bool valid = (cv.w >> 1) == epoch;
                        const double gv = valid ? cell_g(cv) : kInfD;
                        if (ALGO == kAlgoFallback) {
                            const bool card = !(__popc(rowok & ((1u << d) - 1u)) & 1);
                            if (!bit_at(p.free_bits, kv)) {
                                // +100 inflated / +1000 occupied, no step cost (djikstra.cpp:230-246)
                                cand = bit_at(p.infl_bits, kv) ? __dadd_rn(gu, 100.0) : __dadd_rn(gu, 1000.0);
                            } else {
                                cand = __dadd_rn(gu, card ? 1.0 : kSqrt2D);
                            }
                        } else if (ALGO == NSFS_THETASTAR) {
                            const int vi = (int)(kv / W), vj = (int)(kv - (long long)vi * W);
                            if ((los_bits >> d) & 1u) { cand = __dadd_rn(gp, d_euc(pi, pj, vi, vj)); par = pkey; }
                            else cand = __dadd_rn(gu, d_euc(i, j, vi, vj));
                        } else {
                            cand = __dadd_rn(gu, ((om >> (8 + d)) & 1u) && d ? kSqrt2D : 1.0);     // astar.cpp:104-107
                        }
                        relax = gv > __dadd_rn(cand, kSlack);                                         // astar.cpp:116
                        if (relax) cells[kv] = make_cell(cand, par, (epoch << 1) | (valid ? (cv.w & 1u) : 0u));
                    }
                }
                uint32_t rb = __ballot_sync(0xffffffffu, relax);
                // ---- pushes, in NB_LUT order (astar.cpp:120) ----------------------------------------------
                while (rb) {
                    const int src = __ffs(rb) - 1;
                    rb &= rb - 1;
                    const long long pk = __shfl_sync(0xffffffffu, kv, src);
                    const double pg = __shfl_sync(0xffffffffu, cand, src);
                    if (n >= p.heap_cap) { status = NSFS_E_HEAP; break; }
                    if (lane == 0) {
                        const int vi = (int)(pk / W), vj = (int)(pk - (long long)vi * W);
                        heap_push(heap, n, make_entry(pg, heuristic<ALGO>(vi, vj, gi, gj), pk, mode), mode);
                    }
                    n = __shfl_sync(0xffffffffu, n, 0);
                    pushes++;
                    if (n > peak) peak = n;
                }
                __syncwarp();
                if (status != NSFS_OK) break;
            }
        }

        // ---- results ------------------------------------------------------------------------------------
        if (lane == 0) {
            int plen = 1;
            if (ALGO == kAlgoFallback) {
                const long long kr = (found && status == NSFS_OK) ? k_found : ks;     // failure => the input cell (djikstra.cpp:272)
                if (p.paths) { p.paths[2 * qi] = (int)(kr / W); p.paths[2 * qi + 1] = (int)(kr % W); }
            } else if (status == NSFS_OK) {
                int32_t* out = p.paths ? p.paths + (size_t)qi * p.path_cap * 2 : nullptr;
                if (found) {                                                        // back-track (astar.cpp:70-82)
                    long long k = kg;
                    plen = 0;
                    for (;;) {
                        if (out && plen < p.path_cap) { out[2 * plen] = (int)(k / W); out[2 * plen + 1] = (int)(k % W); }
                        plen++;
                        if (k == ks || plen > N) break;
                        k = cells[k].z;
                    }
                } else if (out && p.path_cap > 0) {                                 // astar.cpp:129-135
                    out[0] = si; out[1] = sj;
                }
                if (p.g_goal) {
                    double gg = kInfD;
                    if (kg >= 0) { const uint4 c = cells[kg]; if ((c.w >> 1) == epoch) gg = cell_g(c); }
                    p.g_goal[qi] = gg;
                }
            } else if (p.g_goal && ALGO != kAlgoFallback) {
                p.g_goal[qi] = kInfD;
            }
            if (p.status) p.status[qi] = status;
            if (p.path_len) p.path_len[qi] = plen;
            if (p.settled) p.settled[qi] = expansions;
            if (p.stats5) {
                atomicAdd((unsigned long long*)p.stats5 + 0, (unsigned long long)pops);
                atomicAdd((unsigned long long*)p.stats5 + 1, (unsigned long long)expansions);
                atomicAdd((unsigned long long*)p.stats5 + 2, (unsigned long long)pushes);
                atomicMax((unsigned long long*)p.stats5 + 3, (unsigned long long)peak);
                atomicAdd((unsigned long long*)p.stats5 + 4, (unsigned long long)los_checks);
            }
        }
        __syncwarp();
    }
    if (lane == 0) p.epoch[slot] = epoch;
}

// ---- host side --------------------------------------------------------------------------------------------
int search_alloc(nsfs_planner* h, int want_slots) {
    SearchWork& s = h->search;
    long long heap_cap = h->opt_heap_cap > 0 ? h->opt_heap_cap : h->N + 8;
    if (want_slots < 1) want_slots = 1;
    if (s.cells && s.slots >= want_slots && s.heap_cap >= heap_cap) return NSFS_OK;
    search_free(h);
    size_t free_b = 0, total_b = 0;
    NSFS_CUDA_TRY(h, cudaMemGetInfo(&free_b, &total_b));
    const size_t per_slot = (size_t)h->N * sizeof(uint4) + (size_t)heap_cap * sizeof(unsigned long long);
    long long fit = (long long)((double)free_b * 0.80 / (double)per_slot);
    long long cap = h->opt_max_slots > 0 ? h->opt_max_slots : (long long)h->sm_count * 32;
    long long slots = want_slots;
    if (slots > cap) slots = cap;
    if (slots > fit) slots = fit;
    if (slots < 1) return NSFS_E_NOMEM;
    NSFS_CUDA_TRY(h, cudaMalloc(&s.cells, (size_t)slots * (size_t)h->N * sizeof(uint4)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.heap, (size_t)slots * (size_t)heap_cap * sizeof(unsigned long long)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.epoch, (size_t)slots * sizeof(uint32_t)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.work_counter, sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.cells, 0, (size_t)slots * (size_t)h->N * sizeof(uint4), h->stream));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.epoch, 0, (size_t)slots * sizeof(uint32_t), h->stream));
    s.slots = (int)slots;
    s.heap_cap = heap_cap;
    return NSFS_OK;
}

void search_free(nsfs_planner* h) {
    SearchWork& s = h->search;
    cudaFree(s.cells); cudaFree(s.heap); cudaFree(s.epoch); cudaFree(s.work_counter);
    s = SearchWork();
}

int launch_search(nsfs_planner* h, int kind, int algo, int mode, int nq, const int32_t* d_q, int32_t* d_status,
                  double* d_g_goal, int32_t* d_path_len, long long* d_settled, int32_t* d_paths, int path_cap,
                  long long* d_stats5) {
    if (nq <= 0) return NSFS_OK;
    int rc = search_alloc(h, nq);
    if (rc != NSFS_OK) return rc;
    SearchWork& s = h->search;
    SearchParams p;
    p.free_bits = h->d_free; p.infl_bits = h->d_inflated; p.out_mask = h->d_out_mask;
    p.cells = s.cells; p.heap = s.heap; p.epoch = s.epoch; p.work_counter = s.work_counter;
    p.H = h->H; p.W = h->W; p.N = h->N; p.heap_cap = s.heap_cap;
    p.slots = nq < s.slots ? nq : s.slots;
    p.mode = (mode == NSFS_MODE_F || mode == NSFS_MODE_G) ? mode : NSFS_MODE_FG;     // unknown => fg (grid_planner_core.cpp:274-281)
    p.nq = nq; p.q = d_q; p.status = d_status; p.g_goal = d_g_goal; p.path_len = d_path_len; p.settled = d_settled;
    p.paths = d_paths; p.path_cap = path_cap; p.stats5 = d_stats5;
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.work_counter, 0, sizeof(int), h->stream));
    const int block = SEARCH_WARPS_PER_BLOCK * 32;
    const int grid = (p.slots + SEARCH_WARPS_PER_BLOCK - 1) / SEARCH_WARPS_PER_BLOCK;
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    if (kind == 1) { p.mode = NSFS_MODE_G; search_kernel<kAlgoFallback><<<grid, block, 0, h->stream>>>(p); }
    else if (algo == NSFS_ASTAR) search_kernel<NSFS_ASTAR><<<grid, block, 0, h->stream>>>(p);
    else if (algo == NSFS_DJIKSTRA) search_kernel<NSFS_DJIKSTRA><<<grid, block, 0, h->stream>>>(p);
    else if (algo == NSFS_THETASTAR) search_kernel<NSFS_THETASTAR><<<grid, block, 0, h->stream>>>(p);
    else return NSFS_E_ARG;
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
