// Exact-order search: Astar / Djikstra / ThetaStar plan() and the Djikstra fallback, one query per warp.
//
// The reference's A* (and its fallback search) is order dependent: keys are truncated to int in the open list
// (grid_planner_core.h:43,45), the heuristic is inadmissible (bot_utils.cpp:47), closed cells keep being
// re-labelled (astar.cpp:116-121), and ties between equal int keys are broken only by the mechanics of
// libstdc++'s binary heap.  Matching its paths, costs and fallback cells therefore means replaying its pop
// order exactly (SURVEY.md "three facts" #3).  Each warp replays one query:
//   - the open list is a device binary heap that ends every push and pop in exactly the array std::push_heap /
//     std::pop_heap would leave (bits/stl_heap.h __push_heap / __adjust_heap: the hole sinks to a leaf preferring the
//     right child unless it sorts after the left one, then the displaced last element is sifted back up), but the
//     WARP walks it, not one thread:
//       pop   the 31 internal nodes of a 5-level subtree under the hole are examined at once, lane t loading both
//             children of node t with one 128-bit load; two ballots (which child wins / where the displaced element
//             stops) give the path through the five levels with register bit operations, so a sift-down costs one
//             memory round trip per five levels instead of one per level.  Stopping at the first picked child that
//             sorts after the displaced element is the same final array as "sink to the leaf, then sift up" because
//             priorities never decrease along a root-to-leaf path;
//       push  lane L loads the L-th ancestor of the new leaf, one ballot finds how far the element rises, and the
//             lanes shift that run of ancestors down in parallel: one round trip per push whatever the distance;
//   - the loads an expansion needs (the popped cell, its edge mask, its 8 neighbours' records) depend only on the key
//     at the top of the heap, so they are issued BEFORE the sift-down and their latency hides behind it;
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
    long long heap_cap;           // entries a query's open list may hold
    long long heap_stride;        // array positions per slot (heap_cap + slack for position 0 and the pair over-read; even)
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
    const int32_t* order;         // order[k] = the k-th query to start (longest expected first), or nullptr = input order
};

// Heap storage: heap index x lives at array position x + 1 (position 0 unused), so the two children of x
// (heap indices 2x+1, 2x+2 = positions 2x+2, 2x+3) form one 16-byte aligned pair.
struct Prio {                     // comparator of the selected cost mode as a shift and a mask (grid_planner_core.cpp:55-75)
    int sh;
    unsigned long long msk;
    __device__ __forceinline__ unsigned long long operator()(unsigned long long e) const { return (e >> sh) & msk; }
};
__device__ __forceinline__ Prio make_prio(int mode) {
    Prio p;
    if (mode == NSFS_MODE_F) { p.sh = kKeyBits + kGBits; p.msk = ~0ull; }                 // f only
    else if (mode == NSFS_MODE_G) { p.sh = kKeyBits; p.msk = (1ull << kGBits) - 1ull; }   // g only
    else { p.sh = kKeyBits; p.msk = ~0ull; }                                              // f, then the smaller g
    return p;
}

// One open list: positions below kWideTop (the top ten levels, 8 KB per warp) in shared memory, the rest in global memory, both
// reached through one generic pointer chosen with a select.  A sift-down of a 10^4-entry list then makes its first two 5-level
// chunks at shared-memory latency and only the last one goes to L2 / HBM, and the root is read without a round trip: this is
// the kernel that serves single queries and small batches, where the time of a launch IS the dependent chain of its longest query.
constexpr uint32_t kWideTop = 1024;
struct WideHeap {
    unsigned long long* a;        // global part (whole array; positions < kWideTop unused when s != a)
    unsigned long long* s;        // this warp's shared-memory top (== a for a list that lives in one place)
    __device__ __forceinline__ unsigned long long* at(uint32_t pos) const { return (pos < kWideTop ? s : a) + pos; }
};

// std::priority_queue::push = push_back + __push_heap (stl_heap.h:135-148), by the whole warp (n is warp-uniform).
// Positions fit 32 bits (heap_cap <= 8N + 16 < 2^32 / 2 for N < 2^28; larger capacities are refused by search_alloc).
__device__ __forceinline__ void heap_push(const WideHeap hp, uint32_t& n, unsigned long long v, const Prio prio, int lane) {
    const uint32_t pos = n + 1u;                          // array position of the new leaf
    const unsigned long long pv = prio(v);
    const uint32_t q = pos >> lane;                       // lane L >= 1: position of the L-th ancestor (0 = above the root)
    const bool valid = lane >= 1 && q >= 1u;
    unsigned long long e = 0;
    if (valid) e = *hp.at(q);
    const unsigned rise = __ballot_sync(0xffffffffu, valid && prio(e) > pv);    // ancestor sorts after v: it moves down
    const int m = __ffs(~(rise >> 1)) - 1;                // length of the run of such ancestors starting at the parent
    if (lane >= 1 && lane <= m) *hp.at(pos >> (lane - 1)) = e;
    if (lane == 0) *hp.at(pos >> m) = v;
    n = n + 1u;
    __syncwarp();
}

struct LaneNode {                 // my internal node of a 5-level subtree: level r (0..4), offset o within the level
    int r;
    uint32_t o;
};
__device__ __forceinline__ LaneNode lane_node(int lane) {
    LaneNode ln;
    ln.r = 31 - __clz(lane + 1);
    ln.o = (uint32_t)(lane + 1) - (1u << ln.r);
    return ln;
}

// top() + pop() = __pop_heap -> __adjust_heap -> __push_heap (stl_heap.h:224-262), by the whole warp.
__device__ __forceinline__ void heap_pop(const WideHeap hp, uint32_t& n, const Prio prio, int lane, const LaneNode ln) {
    const uint32_t len = n - 1u;                          // elements left after the pop
    n = len;
    if (len == 0u) return;
    const unsigned long long v = *hp.at(len + 1u);        // the displaced last element
    const unsigned long long pv = prio(v);
    uint32_t c1 = 1u;                                     // position (= heap index + 1) of the hole
    for (;;) {
        const uint32_t pos = (c1 << ln.r) + ln.o;         // position of my node; its children sit at positions 2*pos, 2*pos + 1
        const uint32_t lpos = 2u * pos;                   // left child's position; it exists iff its heap index lpos-1 < len
        const bool has = lane < 31 && (c1 >> (30 - ln.r)) == 0u && lpos <= len;   // (the shift test keeps pos from wrapping)
        ulonglong2 ch = make_ulonglong2(0ull, 0ull);
        if (has) ch = *reinterpret_cast<const ulonglong2*>(hp.at(lpos));     // lpos even: both children on one side of kWideTop
        // __adjust_heap: take the right child unless it sorts after the left one; a lone left child is taken as is
        const bool left = !(lpos + 1u <= len) || prio(ch.y) > prio(ch.x);
        const unsigned long long pick = left ? ch.x : ch.y;
        const bool stop = !has || prio(pick) > pv;        // the displaced element comes to rest in my node
        const unsigned m_left = __ballot_sync(0xffffffffu, left);
        const unsigned m_stop = __ballot_sync(0xffffffffu, stop);
        uint32_t node = 0, steps = 0, path = 0;
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) {
            if ((m_stop >> node) & 1u) break;
            path |= 1u << node;
            node = 2u * node + 2u - ((m_left >> node) & 1u);
            ++steps;
        }
        if ((path >> lane) & 1u) *hp.at(pos) = pick;      // the picked child moves up into my node
        const int rq = 31 - __clz(node + 1u);
        const uint32_t npos = (c1 << rq) + (node + 1u - (1u << rq));
        if (steps < 5u) {
            if (lane == 0) *hp.at(npos) = v;
            break;
        }
        c1 = npos;                                        // five levels down and still sinking
    }
    __syncwarp();
}

__device__ __forceinline__ double cell_g(const uint4& c) { return __hiloint2double((int)c.y, (int)c.x); }
__device__ __forceinline__ uint4 make_cell(double g, uint32_t parent, uint32_t tag) {
    return make_uint4((uint32_t)__double2loint(g), (uint32_t)__double2hiint(g), parent, tag);
}

// dist_oct(Index, Index), bot_utils.cpp:44-65 with the line-47 slip: both deltas are taken against the ROW.
__device__ __forceinline__ double h_oct(int i, int gi, int gj) {
    const int ax = abs(gi - i), ay = abs(gj - i);              // integers: |.|, min and difference are exact either way
    const int ord = min(ax, ay), card = abs(ax - ay);
    return __dadd_rn(__dmul_rn((double)ord, kSqrt2D), (double)card);   // no FMA: the host computes mul then add
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
__device__ __forceinline__ unsigned long long make_entry(double g, double h, uint32_t key, int mode) {
    const double f = (mode == NSFS_MODE_G) ? 0.0 : __dadd_rn(g, h);
    const unsigned long long fi = (unsigned long long)(unsigned)__double2int_rz(f);
    const unsigned long long gi = (unsigned long long)(unsigned)__double2int_rz(g);
    return (fi << (kKeyBits + kGBits)) | (gi << kKeyBits) | (unsigned long long)key;
}

// An open-list entry holds (int)f in 18 bits and (int)g in 17: g stays below the reference's "infinity" 1e5 < 2^17, but f = g + h
// grows with the heuristic, and the goal is the caller's (an off-grid or very distant goal, a very elongated map).  A key that
// does not fit would wrap and silently change the pop order, so such a query ends with NSFS_E_HEAP instead.
__device__ __forceinline__ bool entry_overflows(double g, double h, int mode) {
    const double f = (mode == NSFS_MODE_G) ? 0.0 : __dadd_rn(g, h);
    return !(f < (double)(1 << kGBits << 1)) || !(g < (double)(1 << kGBits));
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
__global__ void __launch_bounds__(SEARCH_WARPS_PER_BLOCK * 32, 7) search_kernel(SearchParams p) {
    __shared__ __align__(16) unsigned long long s_top[SEARCH_WARPS_PER_BLOCK][kWideTop];
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SEARCH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (slot >= p.slots) return;
    const int W = p.W, H = p.H, mode = p.mode;
    const long long N = p.N;
    uint4* cells = p.cells + (size_t)slot * (size_t)N;
    WideHeap heap;
    heap.a = p.heap + (size_t)slot * (size_t)p.heap_stride;
    heap.s = s_top[threadIdx.x >> 5];
    const Prio prio = make_prio(mode);
    const LaneNode ln = lane_node(lane);
    const uint32_t uW = (uint32_t)W, uN = (uint32_t)N;       // N < 2^29 (nsfs_create)
    uint32_t epoch = p.epoch[slot];
    const bool serial = W < 3;      // for W < 3 two directions can alias one key: evaluate them one after another

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= p.nq) break;
        if (p.order) qi = p.order[qi];
        ++epoch;

        int si, sj, gi = -1, gj = -1;
        if (ALGO == kAlgoFallback) { si = p.q[2 * qi]; sj = p.q[2 * qi + 1]; }
        else { si = p.q[4 * qi]; sj = p.q[4 * qi + 1]; gi = p.q[4 * qi + 2]; gj = p.q[4 * qi + 3]; }

        uint32_t n = 0;                        // open-list length (warp-uniform)
        long long pops = 0, expansions = 0, pushes = 0, peak = 0, los_checks = 0;
        int status = NSFS_OK;
        bool found = false;
        uint32_t k_found = 0;
        const uint32_t ks = (uint32_t)(si * W + sj);
        const uint32_t kg = (gi >= 0 && gi < H && gj >= 0 && gj < W) ? (uint32_t)(gi * W + gj) : 0xffffffffu;   // no cell has this key

        if (si < 0 || si >= H || sj < 0 || sj >= W) {
            status = NSFS_E_ARG;               // the reference indexes nodes[] out of bounds here (undefined behaviour)
        } else {
            if (lane == 0) __stcg(cells + ks, make_cell(0.0, ks, epoch << 1));
            __syncwarp();
            const double h0 = heuristic<ALGO>(si, sj, gi, gj);
            if (entry_overflows(0.0, h0, mode)) status = NSFS_E_HEAP;
            else {
                heap_push(heap, n, make_entry(0.0, h0, ks, mode), prio, lane);
                pushes = 1; peak = 1;
            }
        }

        while (status == NSFS_OK && n > 0) {
            // ---- pop (astar.cpp:56) --------------------------------------------------------------------
            const unsigned long long top = heap.s[1];          // the root lives in shared memory: one broadcast load, no round trip
            const uint32_t ku = (uint32_t)(top & kKeyMask);
            const int i = (int)(ku / uW), j = (int)(ku - (uint32_t)i * uW);
            // Everything the expansion reads depends only on ku: issue those loads now, use them after the sift-down.
            //   lanes 0..7  record of neighbour d (key range only; whether d is passable comes with the mask)
            //   lane 8      record of the popped cell          lane 9   its edge mask
            // my neighbour (lanes 0..7): flat key, and its (row, column) without another division - a column that leaves
            // [0, W) wraps into the next / previous row, which is exactly what the flat key does
            const int dd = lane & 7;
            const uint32_t kv_mine = ku + (uint32_t)(nb_di(dd) * W + nb_dj(dd));
            int vi_mine = i + nb_di(dd), vj_mine = j + nb_dj(dd);
            if (vj_mine >= W) { vj_mine -= W; vi_mine++; } else if (vj_mine < 0) { vj_mine += W; vi_mine--; }
            uint4 pre = make_uint4(0u, 0u, 0u, 0u);
            uint32_t pre_mask = 0;
            if (lane < 8) {
                if (kv_mine < uN) pre = __ldcg(cells + kv_mine);       // unsigned: a negative key wraps above N
            } else if (lane == 8) {
                pre = __ldcg(cells + ku);
            } else if (lane == 9) {
                pre_mask = __ldg(p.out_mask + ku);
            }
            heap_pop(heap, n, prio, lane, ln);
            pops++;
            uint4 cu;
            cu.x = __shfl_sync(0xffffffffu, pre.x, 8); cu.y = __shfl_sync(0xffffffffu, pre.y, 8);
            cu.z = __shfl_sync(0xffffffffu, pre.z, 8); cu.w = __shfl_sync(0xffffffffu, pre.w, 8);
            if (cu.w & 1u) continue;           // stale entry (astar.cpp:59-62)
            if (lane == 0) __stcg(&cells[ku].w, cu.w | 1u);
            expansions++;
            if (ALGO == kAlgoFallback) {
                if (bit_at(p.free_bits, ku)) { found = true; k_found = ku; break; }     // djikstra.cpp:205-211
            } else if (ku == kg) { found = true; k_found = ku; break; }                 // astar.cpp:68

            const double gu = cell_g(cu);
            const uint32_t om = __shfl_sync(0xffffffffu, pre_mask, 9);
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
                gp = cell_g(__ldcg(cells + pkey));
                pi = (int)(pkey / (uint32_t)W); pj = (int)(pkey - (uint32_t)pi * (uint32_t)W);
            }
            uint32_t los_bits = 0;
            if (ALGO == NSFS_THETASTAR) {
                for (int d = 0; d < 8; ++d) {
                    if (!((om >> d) & 1u)) continue;
                    const int vi = __shfl_sync(0xffffffffu, vi_mine, d), vj = __shfl_sync(0xffffffffu, vj_mine, d);
                    los_checks++;
                    if (warp_los(p.free_bits, W, pi, pj, vi, vj, lane)) los_bits |= 1u << d;
                }
            }

            for (int round = 0; round < (serial ? 8 : 1); ++round) {
                const int d = lane;
                const bool mine = serial ? (lane == round) : (lane < 8);
                bool relax = false;
                const uint32_t kv = kv_mine;
                double cand = 0.0;
                uint32_t par = (uint32_t)ku;
                if (mine) {
                    bool considered;
                    if (ALGO == kAlgoFallback) considered = (rowok >> d) & 1u;
                    else considered = (om >> d) & 1u;
                    if (considered) {
                        const uint4 cv = serial ? __ldcg(cells + kv) : pre;      // W < 3: an earlier round may have re-labelled it
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
                            if ((los_bits >> d) & 1u) { cand = __dadd_rn(gp, d_euc(pi, pj, vi_mine, vj_mine)); par = pkey; }
                            else cand = __dadd_rn(gu, d_euc(i, j, vi_mine, vj_mine));
                        } else {
                            cand = __dadd_rn(gu, ((om >> (8 + d)) & 1u) && d ? kSqrt2D : 1.0);     // astar.cpp:104-107
                        }
                        relax = gv > __dadd_rn(cand, kSlack);                                         // astar.cpp:116
                        if (relax) __stcg(cells + kv, make_cell(cand, par, (epoch << 1) | (valid ? (cv.w & 1u) : 0u)));
                    }
                }
                uint32_t rb = __ballot_sync(0xffffffffu, relax);
                // ---- pushes, in NB_LUT order (astar.cpp:120) ----------------------------------------------
                while (rb) {
                    const int src = __ffs(rb) - 1;
                    rb &= rb - 1;
                    const uint32_t pk = __shfl_sync(0xffffffffu, kv, src);
                    const double pg = __shfl_sync(0xffffffffu, cand, src);
                    if ((long long)n >= p.heap_cap) { status = NSFS_E_HEAP; break; }
                    const int vi = __shfl_sync(0xffffffffu, vi_mine, src), vj = __shfl_sync(0xffffffffu, vj_mine, src);
                    const double ph = heuristic<ALGO>(vi, vj, gi, gj);
                    if (entry_overflows(pg, ph, mode)) { status = NSFS_E_HEAP; break; }
                    heap_push(heap, n, make_entry(pg, ph, pk, mode), prio, lane);
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
                const uint32_t kr = (found && status == NSFS_OK) ? k_found : ks;      // failure => the input cell (djikstra.cpp:272)
                if (p.paths) { p.paths[2 * qi] = (int)(kr / uW); p.paths[2 * qi + 1] = (int)(kr % uW); }
            } else if (status == NSFS_OK) {
                int32_t* out = p.paths ? p.paths + (size_t)qi * p.path_cap * 2 : nullptr;
                if (found) {                                                        // back-track (astar.cpp:70-82)
                    uint32_t k = kg;
                    plen = 0;
                    for (;;) {
                        if (out && plen < p.path_cap) { out[2 * plen] = (int)(k / uW); out[2 * plen + 1] = (int)(k % uW); }
                        plen++;
                        if (k == ks || (long long)plen > N) break;
                        k = __ldcg(cells + k).z;
                    }
                } else if (out && p.path_cap > 0) {                                 // astar.cpp:129-135
                    out[0] = si; out[1] = sj;
                }
                if (p.g_goal) {
                    double gg = kInfD;
                    if (kg != 0xffffffffu) { const uint4 c = __ldcg(cells + kg); if ((c.w >> 1) == epoch) gg = cell_g(c); }
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

// ---- find_better_point with its whole state in shared memory -----------------------------------------------------
// The fallback search touches a handful of cells around the bad cell (10-16 on average), so a dense per-cell
// store (16 B x N per slot: 4 GB on the 16384^2 map) is the wrong shape for it.  Here each warp keeps the touched cells in
// a 256-slot open-addressed table and the open list in 256 heap positions, both in shared memory; the replay itself (same
// heap primitives, same fp64 sums, same NB_LUT push order) is unchanged.  A search that outgrows the table ends with
// NSFS_E_HEAP and the C-ABI re-runs it on the wide kernel, so results never depend on which kernel ran.
constexpr int FB_CAP = 256;
constexpr int FB_MAX_CELLS = 128;
constexpr uint32_t FB_EMPTY = 0xffffffffu;
struct FbState {
    unsigned long long heap[FB_CAP + 8];
    double g[FB_CAP];
    uint32_t key[FB_CAP];
    uint8_t vis[FB_CAP];
};

__device__ __forceinline__ int fb_find(const FbState& st, uint32_t key) {           // slot of key, or -1
    uint32_t s0 = (key * 2654435761u) >> 24;
    for (int probe = 0; probe < FB_CAP; ++probe) {
        const uint32_t k = st.key[s0];
        if (k == key) return (int)s0;
        if (k == FB_EMPTY) return -1;
        s0 = (s0 + 1) & (FB_CAP - 1);
    }
    return -1;
}

__device__ __forceinline__ int fb_insert(FbState& st, uint32_t key) {               // one thread; the key is known to be absent
    uint32_t s0 = (key * 2654435761u) >> 24;
    while (st.key[s0] != FB_EMPTY) s0 = (s0 + 1) & (FB_CAP - 1);
    st.key[s0] = key;
    st.vis[s0] = 0;
    return (int)s0;
}

__global__ void __launch_bounds__(SEARCH_WARPS_PER_BLOCK * 32) fallback_small_kernel(SearchParams p) {
    __shared__ __align__(16) FbState sm[SEARCH_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31;
    FbState& st = sm[threadIdx.x >> 5];
    WideHeap fbh;
    fbh.a = st.heap; fbh.s = st.heap;                           // the whole list in one (shared) array
    const int W = p.W, H = p.H;
    const uint32_t uW = (uint32_t)W;
    const Prio prio = make_prio(NSFS_MODE_G);                   // the fallback planner is prepared with cost mode "g" (global_planner.cpp:127)
    const LaneNode ln = lane_node(lane);
    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= p.nq) break;
        const int si = p.q[2 * qi], sj = p.q[2 * qi + 1];
        for (int t = lane; t < FB_CAP; t += 32) { st.key[t] = FB_EMPTY; st.vis[t] = 0; }
        __syncwarp();
        uint32_t n = 0, peak = 0;
        int ncells = 0, status = NSFS_OK;
        long long pops = 0, expansions = 0, pushes = 0;
        bool found = false;
        uint32_t k_found = 0;
        const uint32_t ks = (uint32_t)(si * W + sj);
        if (si < 0 || si >= H || sj < 0 || sj >= W) status = NSFS_E_ARG;
        else {
            if (lane == 0) { const int s0 = fb_insert(st, ks); st.g[s0] = 0.0; }
            ncells = 1;
            __syncwarp();
            heap_push(fbh, n, make_entry(0.0, 0.0, ks, NSFS_MODE_G), prio, lane);
            pushes = 1; peak = 1;
        }
        while (status == NSFS_OK && n > 0) {
            const uint32_t ku = (uint32_t)(st.heap[1] & kKeyMask);
            heap_pop(fbh, n, prio, lane, ln);
            pops++;
            const int su = fb_find(st, ku);                    // every key on the open list is in the table
            if (st.vis[su]) continue;                           // stale entry (djikstra.cpp:197-200)
            __syncwarp();
            if (lane == 0) st.vis[su] = 1;
            expansions++;
            if (bit_at(p.free_bits, ku)) { found = true; k_found = ku; break; }          // djikstra.cpp:205-211
            const uint32_t om = __ldg(p.out_mask + ku);
            if (om & kThrowBit) { status = NSFS_E_RANGE; break; }                        // vector::at throws (R6)
            const int i = (int)(ku / uW);
            const double gu = st.g[su];
            uint32_t rowok = 0;                                 // directions whose row is in range (djikstra.cpp:222-226)
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int ni = i + nb_di(d);
                if (ni >= 0 && ni < H) rowok |= 1u << d;
            }
            bool relax = false;
            uint32_t kv = 0;
            double cand = 0.0;
            if (lane < 8 && ((rowok >> lane) & 1u)) {
                const int d = lane;
                kv = ku + (uint32_t)(nb_di(d) * W + nb_dj(d));
                const int sv = fb_find(st, kv);
                const double gv = sv >= 0 ? st.g[sv] : kInfD;
                const bool card = !(__popc(rowok & ((1u << d) - 1u)) & 1);
                if (!bit_at(p.free_bits, kv)) cand = bit_at(p.infl_bits, kv) ? __dadd_rn(gu, 100.0) : __dadd_rn(gu, 1000.0);   // djikstra.cpp:230-246
                else cand = __dadd_rn(gu, card ? 1.0 : kSqrt2D);
                relax = gv > __dadd_rn(cand, kSlack);
            }
            uint32_t rb = __ballot_sync(0xffffffffu, relax);
            __syncwarp();
            while (rb) {                                        // pushes in NB_LUT order
                const int src = __ffs(rb) - 1;
                rb &= rb - 1;
                const uint32_t pk = __shfl_sync(0xffffffffu, kv, src);
                const double pg = __shfl_sync(0xffffffffu, cand, src);
                int sv = fb_find(st, pk);
                if (sv < 0) {
                    if (ncells >= FB_MAX_CELLS) { status = NSFS_E_HEAP; break; }
                    if (lane == 0) fb_insert(st, pk);
                    ncells++;
                    __syncwarp();
                    sv = fb_find(st, pk);
                }
                if (n >= (uint32_t)(FB_CAP - 2)) { status = NSFS_E_HEAP; break; }
                if (lane == 0) st.g[sv] = pg;                  // the label changes even when the cell is closed (djikstra.cpp:255-259)
                __syncwarp();
                heap_push(fbh, n, make_entry(pg, 0.0, pk, NSFS_MODE_G), prio, lane);
                pushes++;
                if (n > peak) peak = n;
            }
            __syncwarp();
        }
        if (lane == 0) {
            const uint32_t kr = (found && status == NSFS_OK) ? k_found : ks;          // failure => the input cell (djikstra.cpp:272)
            if (p.paths) { p.paths[2 * qi] = (int)(kr / uW); p.paths[2 * qi + 1] = (int)(kr % uW); }
            if (p.status) p.status[qi] = status;
            if (p.path_len) p.path_len[qi] = 1;
            if (p.settled) p.settled[qi] = expansions;
            if (p.stats5) {
                atomicAdd((unsigned long long*)p.stats5 + 0, (unsigned long long)pops);
                atomicAdd((unsigned long long*)p.stats5 + 1, (unsigned long long)expansions);
                atomicAdd((unsigned long long*)p.stats5 + 2, (unsigned long long)pushes);
                atomicMax((unsigned long long*)p.stats5 + 3, (unsigned long long)peak);
            }
        }
        __syncwarp();
    }
}

// ---- host side --------------------------------------------------------------------------------------------
int search_alloc(nsfs_planner* h, int want_slots) {
    SearchWork& s = h->search;
    long long heap_cap = h->opt_heap_cap > 0 ? h->opt_heap_cap : h->N + 8;
    if (heap_cap > (1ll << 31) - 64) heap_cap = (1ll << 31) - 64;            // heap positions are 32-bit
    if (want_slots < 1) want_slots = 1;
    const long long cap = h->opt_max_slots > 0 ? h->opt_max_slots : (long long)h->sm_count * 24;     // six resident blocks of four warps per SM
    long long slots = want_slots;
    if (slots > cap) slots = cap;
    // reuse: enough slots, or an allocation that was made for at least this many and clamped by free HBM (asking again would
    // only free and re-allocate the same ~100 GB)
    if (s.cells && s.heap_cap >= heap_cap && (s.slots >= slots || s.target >= slots)) return NSFS_OK;
    search_free(h);
    size_t free_b = 0, total_b = 0;
    NSFS_CUDA_TRY(h, cudaMemGetInfo(&free_b, &total_b));
    const long long heap_stride = (heap_cap + 5) & ~1ll;
    const size_t per_slot = (size_t)h->N * sizeof(uint4) + (size_t)heap_stride * sizeof(unsigned long long);
    long long fit = (long long)((double)free_b * 0.80 / (double)per_slot);
    const long long target = slots;
    if (slots > fit) slots = fit;
    if (slots < 1) return NSFS_E_NOMEM;
    NSFS_CUDA_TRY(h, cudaMalloc(&s.cells, (size_t)slots * (size_t)h->N * sizeof(uint4)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.heap, (size_t)slots * (size_t)heap_stride * sizeof(unsigned long long)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.epoch, (size_t)slots * sizeof(uint32_t)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.work_counter, sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.cells, 0, (size_t)slots * (size_t)h->N * sizeof(uint4), h->stream));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.epoch, 0, (size_t)slots * sizeof(uint32_t), h->stream));
    s.slots = (int)slots;
    s.target = (int)target;
    s.heap_cap = heap_cap;
    return NSFS_OK;
}

void search_free(nsfs_planner* h) {
    SearchWork& s = h->search;
    cudaFree(s.cells); cudaFree(s.heap); cudaFree(s.epoch); cudaFree(s.work_counter); cudaFree(s.order);
    s = SearchWork();
    if (h->last_plan_kind == 1) h->last_plan_kind = 0;      // the parent store nsfs_last_path would walk is gone
}

// The back-track of the last single query again (astar.cpp:70-82), for a caller whose path buffer was too small: the parent
// store of slot 0 still holds that search.  One thread chases the chain.
__global__ void path_walk_kernel(const uint4* __restrict__ cells, int W, long long N, uint32_t ks, uint32_t kg, int32_t* __restrict__ out,
                                 int cap, int* __restrict__ n_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t k = kg;
    int plen = 0;
    for (;;) {
        if (plen < cap) { out[2 * plen] = (int)(k / (uint32_t)W); out[2 * plen + 1] = (int)(k % (uint32_t)W); }
        plen++;
        if (k == ks || (long long)plen > N) break;
        k = __ldcg(cells + k).z;
    }
    *n_out = plen;
}

int launch_path_walk(nsfs_planner* h, long long ks, long long kg, int32_t* d_path, int cap, int* d_n) {
    if (!h->search.cells) return NSFS_E_NOFIELD;
    path_walk_kernel<<<1, 1, 0, h->stream>>>(h->search.cells, h->W, h->N, (uint32_t)ks, (uint32_t)kg, d_path, cap, d_n);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

// find_better_point on the shared-memory kernel: no per-query HBM state at all.
int launch_fallback_small(nsfs_planner* h, int nq, const int32_t* d_q, int32_t* d_status, long long* d_settled, int32_t* d_cells,
                          long long* d_stats5) {
    if (nq <= 0) return NSFS_OK;
    SearchWork& s = h->search;
    if (!s.work_counter) NSFS_CUDA_TRY(h, cudaMalloc(&s.work_counter, sizeof(int)));
    SearchParams p = {};
    p.free_bits = h->d_free; p.infl_bits = h->d_inflated; p.out_mask = h->d_out_mask; p.work_counter = s.work_counter;
    p.H = h->H; p.W = h->W; p.N = h->N; p.mode = NSFS_MODE_G;
    p.nq = nq; p.q = d_q; p.status = d_status; p.settled = d_settled; p.paths = d_cells; p.path_cap = 1; p.stats5 = d_stats5;
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.work_counter, 0, sizeof(int), h->stream));
    const int block = SEARCH_WARPS_PER_BLOCK * 32;
    long long warps = nq;
    const long long cap = (long long)h->sm_count * 32;
    if (warps > cap) warps = cap;
    const int grid = (int)((warps + SEARCH_WARPS_PER_BLOCK - 1) / SEARCH_WARPS_PER_BLOCK);
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    fallback_small_kernel<<<grid, block, 0, h->stream>>>(p);
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

int launch_search(nsfs_planner* h, int kind, int algo, int mode, int nq, const int32_t* d_q, int32_t* d_status,
                  double* d_g_goal, int32_t* d_path_len, long long* d_settled, int32_t* d_paths, int path_cap,
                  long long* d_stats5, const int32_t* ext_order, cudaStream_t ext_stream, bool part) {
    if (nq <= 0) return NSFS_OK;
    cudaStream_t stream = ext_stream ? ext_stream : h->stream;
    int rc = search_alloc(h, nq);
    if (rc != NSFS_OK) return rc;
    SearchWork& s = h->search;
    SearchParams p;
    p.free_bits = h->d_free; p.infl_bits = h->d_inflated; p.out_mask = h->d_out_mask;
    p.cells = s.cells; p.heap = s.heap; p.epoch = s.epoch; p.work_counter = s.work_counter;
    p.H = h->H; p.W = h->W; p.N = h->N; p.heap_cap = s.heap_cap; p.heap_stride = (s.heap_cap + 5) & ~1ll;
    p.slots = nq < s.slots ? nq : s.slots;
    p.mode = (mode == NSFS_MODE_F || mode == NSFS_MODE_G) ? mode : NSFS_MODE_FG;     // unknown => fg (grid_planner_core.cpp:274-281)
    p.nq = nq; p.q = d_q; p.status = d_status; p.g_goal = d_g_goal; p.path_len = d_path_len; p.settled = d_settled;
    p.paths = d_paths; p.path_cap = path_cap; p.stats5 = d_stats5;
    p.order = ext_order;
    if (!part && kind != 1 && nq > p.slots && h->opt_search_order != 2) {     // more queries than slots: start the long ones first
        if (s.order_cap < nq) {
            cudaFree(s.order); s.order = nullptr; s.order_cap = 0;
            NSFS_CUDA_TRY(h, cudaMalloc(&s.order, (size_t)nq * sizeof(int32_t)));
            s.order_cap = nq;
        }
        launch_start_order(h, d_q, nq, algo != NSFS_DJIKSTRA, s.order);
        p.order = s.order;
    }
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.work_counter, 0, sizeof(int), stream));
    const int block = SEARCH_WARPS_PER_BLOCK * 32;
    const int grid = (p.slots + SEARCH_WARPS_PER_BLOCK - 1) / SEARCH_WARPS_PER_BLOCK;
    if (kind != 1 && algo != NSFS_ASTAR && algo != NSFS_DJIKSTRA && algo != NSFS_THETASTAR) return NSFS_E_ARG;
    if (!part) NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, stream));
    const int carve = h->carveout > 0 ? h->carveout : (int)cudaSharedmemCarveoutMaxShared;
    // 32 KB of shared memory per block (the open lists' top levels): ask for the carveout that keeps six blocks resident
#define NSFS_WIDE_LAUNCH(A) do { cudaFuncSetAttribute(search_kernel<A>, cudaFuncAttributePreferredSharedMemoryCarveout, carve); \
                                 search_kernel<A><<<grid, block, 0, stream>>>(p); } while (0)
    if (kind == 1) { p.mode = NSFS_MODE_G; NSFS_WIDE_LAUNCH(kAlgoFallback); }
    else if (algo == NSFS_ASTAR) NSFS_WIDE_LAUNCH(NSFS_ASTAR);
    else if (algo == NSFS_DJIKSTRA) NSFS_WIDE_LAUNCH(NSFS_DJIKSTRA);
    else NSFS_WIDE_LAUNCH(NSFS_THETASTAR);
#undef NSFS_WIDE_LAUNCH
    if (!part) NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, stream));
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
