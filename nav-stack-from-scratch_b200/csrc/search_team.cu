// Exact-order Astar / Djikstra plan() for BATCHES: four queries per warp, eight lanes ("a team") per query.
//
// Why a second kernel: ncu on the one-query-per-warp kernel (search.cu) at batch size shows it is bound by issue slots,
// not by memory: ~500 warp instructions per pop, most of them warp-uniform bookkeeping that all 32 lanes repeat
// (profiles/r01_search_ncu.md).  Here one instruction stream serves four queries: every lane group of 8 runs its own
// query in lockstep with the other three, so the same bookkeeping instructions advance four open lists at once.
// Eight lanes are also exactly the 8 neighbours of an expansion.
//
// What is replayed is unchanged (see search.cu): libstdc++'s push_heap / pop_heap with the reference's three comparators,
// int-truncated keys, pushes in NB_LUT order, closed cells re-labelled.  Entries are 64-bit words
//   high word  (int)f : 16 | (int)g : 16        low word  flat key (29 bits)
// or, where the comparator is "f" and a key fits 20 bits, 32-bit words (int)f : 12 | key : 20 (EntryT below).  Either way a
// comparator is ONE 32-bit compare (f: hi >> 16, g: hi & 0xffff, fg: hi; e >> 20), chosen at compile time (cost mode and entry
// size are template parameters).  An f or g beyond its bits ends the query with NSFS_E_HEAP like a step-count overflow (re-run on
// the wide kernel, whose entries hold f:18 | g:17).
// The first 1 KB of every open list (positions 1..127 of 8-byte entries = seven levels, 1..255 of 4-byte ones = eight) lives in
// SHARED memory: a sift-down starts with two 3-level chunks at shared-memory latency, the root is read without a trip to
// L2 / HBM, and a push that rises into the top meets it there (ncu round 1: ~8 dependent global round trips per pop,
// long_scoreboard the first stall).
//   pop   3-level chunks: team lane t < 7 owns internal node t of the subtree under the hole and loads both children
//         with one 128-bit load; two ballots give the path through the three levels; stop at the first picked child that
//         sorts after the displaced element (same final array as sink-to-leaf + sift-up, priorities are monotone on a path)
//   push  lanes 1..7 load ancestors 1..7 of the new leaf, one ballot finds how far the element rises
// Per-cell record: the label as EXACT step counts, g = n_card + n_diag * sqrt(2) (SURVEY.md §7 "hard parts" 6), in one of two sizes
//   8 bytes   x = n_card | n_diag << 16, y = epoch << 4 | visited << 3 | direction of the edge from the parent (parent key = key - OFF[d])
//   4 bytes   n_card : 12 | n_diag : 12 | direction : 3 | visited : 1 | epoch : 4     for maps up to 2048 cells a side (a label
//             beyond 4095 steps of one kind ends the query with NSFS_E_HEAP like any other overflow).  The 4-bit epoch wraps every 15
//             queries of a slot, which then clears its records (4 MB at 1024^2, ~0.1 ms against seconds per query).  Half the
//             footprint per query in flight = more queries in flight and fewer DRAM sectors per expansion: the batch kernel's
//             throughput is set by both (profiles/r02_search.md §3).
// Every decision the reference takes on g is reproduced from the integer pair: (int)g and (int)(g + h) cannot differ because
// a + b*sqrt(2) with b != 0 stays >= 1/(2.83 b) away from any integer (b < 65536 here) while fp64 accumulates < 1e-9 of
// error, and "g[v] > cand + 1e-5" compares two such numbers whose difference is either exactly 0 or >= 1e-5-clear for
// b < 35000.  g(goal) is reported as n_card + n_diag*sqrt(2) in fp64: within 1e-12 relative of the reference's running
// sum (the stated bar is 1e-5).  A label that would overflow 16 bits ends the query with NSFS_E_HEAP, which the caller
// re-runs on the wide kernel of search.cu, like an open list that outgrows its slot.
#include "nsfs_internal.cuh"

namespace nsfs {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TEAM_WARPS_PER_BLOCK = 4;
enum { PH_FETCH = 0, PH_SEARCH = 1, PH_WALK = 2, PH_FINAL = 3, PH_DONE = 4 };

struct TeamParams {
    const uint16_t* out_mask;
    void* cells;                  // [slots][N] records of 8 or 4 bytes (template parameter CB)
    void* heap;                   // [slots][heap_stride] entries of 8 or 4 bytes (template parameter EB)
    uint32_t* epoch;              // [slots]
    int* work_counter;
    int H, W;
    long long N;
    long long heap_cap, heap_stride;
    long long cell_stride;        // bytes per slot of the record array (a multiple of 16)
    int slots;
    int mode;
    int nq;
    const int32_t* q;             // {si, sj, gi, gj} per query
    const int32_t* order;         // order[k] = the k-th query to start (longest expected first), or nullptr = input order
    int32_t* status;
    double* g_goal;
    int32_t* path_len;
    long long* settled;
    int32_t* paths;               // optional [nq][path_cap][2]
    int path_cap;
    long long* stats5;
    uint32_t div_m;               // key / W = __umulhi(key, div_m) >> div_s, exact for key < 2^29 (set by the launcher)
    int div_s;
};

// A cell record as loaded (either size); fields are extracted where they are used, so a record costs one or two registers.
template <int CB>
struct CellRec {
    uint32_t x, y;                                       // CB == 4: x is the whole record, y unused
    __device__ __forceinline__ uint32_t nc() const { return CB == 8 ? (x & 0xffffu) : (x & 0xfffu); }
    __device__ __forceinline__ uint32_t nd() const { return CB == 8 ? (x >> 16) : ((x >> 12) & 0xfffu); }
    __device__ __forceinline__ uint32_t dir() const { return CB == 8 ? (y & 7u) : ((x >> 24) & 7u); }
    __device__ __forceinline__ uint32_t visited() const { return CB == 8 ? ((y >> 3) & 1u) : ((x >> 27) & 1u); }
    __device__ __forceinline__ bool valid(uint32_t epoch) const { return CB == 8 ? (y >> 4) == epoch : (x >> 28) == epoch; }
};
template <int CB>
__device__ __forceinline__ CellRec<CB> cell_load(const void* base, uint32_t k) {
    CellRec<CB> c;
    if (CB == 8) { const uint2 r = __ldcg(reinterpret_cast<const uint2*>(base) + k); c.x = r.x; c.y = r.y; }
    else { c.x = __ldcg(reinterpret_cast<const uint32_t*>(base) + k); c.y = 0u; }
    return c;
}
template <int CB>
__device__ __forceinline__ void cell_store(void* base, uint32_t k, uint32_t nc, uint32_t nd, uint32_t dir, uint32_t visited, uint32_t epoch) {
    if (CB == 8) __stcg(reinterpret_cast<uint2*>(base) + k, make_uint2(nc | (nd << 16), (epoch << 4) | (visited << 3) | dir));
    else __stcg(reinterpret_cast<uint32_t*>(base) + k, nc | (nd << 12) | (dir << 24) | (visited << 27) | (epoch << 28));
}
template <int CB> __device__ __forceinline__ constexpr uint32_t cell_max_steps() { return CB == 8 ? 0xffffu : 0xfffu; }

// Open-list entries come in two sizes (template parameter EB):
//   8 bytes   high word (int)f : 16 | (int)g : 16, low word flat key            any cost mode, any map
//   4 bytes   (int)f : 12 | flat key : 20                                         cost mode "f" on maps of up to 2^20 cells: that
//             comparator never looks at g (grid_planner_core.cpp:55-59), so the entry need not carry it; half the bytes per heap
//             level, and the shared-memory top holds eight levels instead of seven.  f beyond 12 bits -> NSFS_E_HEAP (wide kernel).
template <int EB> struct EntryT { typedef unsigned long long type; typedef ulonglong2 pair; };
template <> struct EntryT<4> { typedef uint32_t type; typedef uint2 pair; };
template <int EB> __device__ __forceinline__ constexpr uint32_t top_n() { return EB == 4 ? 256u : 128u; }     // entries of the shared-memory top (1 KB)

// the selected comparator's sort key
template <int MODE>
__device__ __forceinline__ uint32_t prio_of(unsigned long long e) {
    const uint32_t hi = (uint32_t)(e >> 32);
    if (MODE == NSFS_MODE_F) return hi >> 16;            // f only            (grid_planner_core.cpp:55-59)
    if (MODE == NSFS_MODE_G) return hi & 0xffffu;        // g only            (61-65)
    return hi;                                           // f, then smaller g (67-75)
}
template <int MODE>
__device__ __forceinline__ uint32_t prio_of(uint32_t e) { return e >> 20; }      // 4-byte entries exist for MODE_F only
template <int EB> __device__ __forceinline__ uint32_t entry_key(typename EntryT<EB>::type e) { return EB == 4 ? ((uint32_t)e & 0xfffffu) : (uint32_t)e; }

__device__ __forceinline__ unsigned team_ballot(bool pred, int tbase) { return (__ballot_sync(FULL, pred) >> tbase) & 0xffu; }

// One open list: positions < top_n (1 KB worth of entries) in shared memory, the rest in global memory; position = heap index + 1.  Both halves are
// reached through ONE generic pointer chosen with a select (no branch): the first version of this split branched per access and
// its extra ~80 instructions per pop cost more than the saved round trips (ncu, round 2: the accessors were 20 % of all
// instructions and the kernel is bound by how many it issues in sequence, not by memory).
template <int EB>
struct TeamHeap {
    typedef typename EntryT<EB>::type E;
    typedef typename EntryT<EB>::pair P;
    E* a;        // global part (whole array; positions < top_n unused)
    E* s;        // generic pointer to this team's shared-memory top
    __device__ __forceinline__ E* at(uint32_t pos) const { return (pos < top_n<EB>() ? s : a) + pos; }
    __device__ __forceinline__ E ld(uint32_t pos) const { return *at(pos); }
    __device__ __forceinline__ void st(uint32_t pos, E v) const { *at(pos) = v; }
    __device__ __forceinline__ P ld_pair(uint32_t lpos) const {     // lpos even: both children on the same side of top_n
        return *reinterpret_cast<const P*>(at(lpos));
    }
};

// std::pop_heap for every team with act set (each on its own array); warp-uniform control flow.
template <int MODE, int EB>
__device__ __forceinline__ void team_pop(const TeamHeap<EB> hp, uint32_t& n, bool act, int tl, int tbase, int r, uint32_t o) {
    typedef typename EntryT<EB>::type E;
    uint32_t len = 0, pv = 0;
    E v = 0;
    bool sinking = false;
    if (act) {
        len = n - 1u;
        n = len;
        if (len > 0u) { v = hp.ld(len + 1u); pv = prio_of<MODE>(v); sinking = true; }
    }
    uint32_t c1 = 1u;                                     // position (heap index + 1) of the hole
    while (__any_sync(FULL, sinking)) {
        const uint32_t pos = (c1 << r) + o;               // my node; its children sit at positions 2*pos, 2*pos + 1
        const uint32_t lpos = 2u * pos;
        const bool has = sinking && tl < 7 && (c1 >> (30 - r)) == 0u && lpos <= len;
        typename EntryT<EB>::pair ch;
        ch.x = 0; ch.y = 0;
        if (has) ch = hp.ld_pair(lpos);
        const bool left = !(lpos + 1u <= len) || prio_of<MODE>(ch.y) > prio_of<MODE>(ch.x);      // right child unless it sorts after the left one
        const E pick = left ? ch.x : ch.y;
        const bool stop = !has || prio_of<MODE>(pick) > pv;
        const unsigned m_left = team_ballot(left, tbase);
        const unsigned m_stop = team_ballot(stop, tbase);
        // the path through the three levels, straight-line (the same chain of three bit tests as a loop, without its branches);
        // o1..o3 = offset of the path's node within its level, so the hole's new position needs no clz
        const uint32_t l0 = m_left & 1u, o1 = 1u - l0, n1 = 1u + o1;
        const uint32_t l1 = (m_left >> n1) & 1u, o2 = 2u * o1 + 1u - l1, n2 = 3u + o2;
        const uint32_t l2 = (m_left >> n2) & 1u, o3 = 2u * o2 + 1u - l2;
        const bool s0 = m_stop & 1u, s1 = (m_stop >> n1) & 1u, s2 = (m_stop >> n2) & 1u;
        const uint32_t steps = s0 ? 0u : (s1 ? 1u : (s2 ? 2u : 3u));
        const uint32_t path = s0 ? 0u : (1u | (s1 ? 0u : ((1u << n1) | (s2 ? 0u : (1u << n2)))));
        const uint32_t npos = s0 ? c1 : (s1 ? 2u * c1 + o1 : (s2 ? 4u * c1 + o2 : 8u * c1 + o3));
        if (sinking && ((path >> tl) & 1u)) hp.st(pos, pick);
        if (sinking) {
            if (steps < 3u) { if (tl == 0) hp.st(npos, v); sinking = false; }
            else c1 = npos;
        }
        __syncwarp();
    }
}

// std::push_heap for every team with pushing set.
template <int MODE, int EB>
__device__ __forceinline__ void team_push(const TeamHeap<EB> hp, uint32_t& n, typename EntryT<EB>::type v, bool pushing, int tl, int tbase) {
    typedef typename EntryT<EB>::type E;
    uint32_t base = n + 1u;                               // position of the hole the new element rises from
    const uint32_t pv = prio_of<MODE>(v);
    bool rising = pushing;
    while (__any_sync(FULL, rising)) {
        const uint32_t q = base >> tl;                    // tl >= 1: position of the tl-th ancestor of the hole
        const bool valid = rising && tl >= 1 && q >= 1u;
        E e = 0;
        if (valid) e = hp.ld(q);
        const unsigned rise = team_ballot(valid && prio_of<MODE>(e) > pv, tbase);
        const int m = __ffs(~(rise >> 1)) - 1;            // run of ancestors (from the parent up) that sort after v: 0..7
        if (rising && tl >= 1 && tl <= m) hp.st(base >> (tl - 1), e);
        if (rising) {
            if (m < 7) { if (tl == 0) hp.st(base >> m, v); rising = false; }
            else base >>= 7;                              // seven levels up and still rising
        }
        __syncwarp();
    }
    if (pushing) n = n + 1u;
}

template <int ALGO, int MODE, int CB, int EB>
__global__ void __launch_bounds__(TEAM_WARPS_PER_BLOCK * 32, CB == 4 ? 9 : 7) search_team_kernel(TeamParams p) {
    typedef typename EntryT<EB>::type E;
    static_assert(EB == 8 || MODE == NSFS_MODE_F, "4-byte entries carry f only");
    __shared__ __align__(16) E s_top[TEAM_WARPS_PER_BLOCK * 4][top_n<EB>()];
    const int lane = threadIdx.x & 31, tl = lane & 7, tbase = lane & 24;
    const int slot = (blockIdx.x * TEAM_WARPS_PER_BLOCK + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    const int W = p.W;
    constexpr int mode = MODE;
    const uint32_t uW = (uint32_t)W, uN = (uint32_t)p.N;
    const bool have_slot = slot < p.slots;
    char* cells = reinterpret_cast<char*>(p.cells) + (size_t)(have_slot ? slot : 0) * (size_t)p.cell_stride;
    TeamHeap<EB> heap;
    heap.a = reinterpret_cast<E*>(p.heap) + (size_t)(have_slot ? slot : 0) * (size_t)p.heap_stride;
    heap.s = s_top[(threadIdx.x >> 5) * 4 + (lane >> 3)];
    const int nr = 31 - __clz(tl + 1);                    // my node of a 3-level subtree: level 0,1,1,2,2,2,2 (lane 7: unused)
    const uint32_t no = (uint32_t)(tl + 1) - (1u << nr);
    const int my_di = nb_di(tl), my_dj = nb_dj(tl);
    const int my_off = my_di * W + my_dj;

    uint32_t epoch = have_slot ? p.epoch[slot] : 0u;
    int phase = have_slot ? PH_FETCH : PH_DONE;
    // per-query state (uniform within a team)
    uint32_t n = 0, ks = 0, kg = 0xffffffffu, peak = 0, wk = 0;
    int qi = 0, gi = 0, gj = 0, status = NSFS_OK, plen = 0;
    bool found = false;
    long long pops = 0, expansions = 0, pushes = 0;

    for (;;) {
        // ---- next query (skipped by the whole warp unless one of its four teams wants one) ---------------------------
        if (__any_sync(FULL, phase == PH_FETCH)) {
            const bool fetch = phase == PH_FETCH;
            int q = 0;
            if (fetch && tl == 0) q = atomicAdd(p.work_counter, 1);
            q = __shfl_sync(FULL, q, 0, 8);
            if (fetch) {
                if (q >= p.nq) phase = PH_DONE;
                else {
                    qi = p.order ? p.order[q] : q;
                    ++epoch;
                    if (CB == 4 && epoch > 15u) {
                        // the 4-bit stamp has gone round: this team wipes its records (128 bytes per step) and starts again at 1
                        uint4* w = reinterpret_cast<uint4*>(cells);
                        const size_t n16 = (size_t)p.cell_stride / 16;
                        for (size_t t = tl; t < n16; t += 8) __stcg(w + t, make_uint4(0u, 0u, 0u, 0u));
                        epoch = 1u;
                    }
                    const int si = p.q[4 * qi], sj = p.q[4 * qi + 1];
                    gi = p.q[4 * qi + 2]; gj = p.q[4 * qi + 3];
                    n = 0; peak = 0; pops = 0; expansions = 0; pushes = 0; found = false; status = NSFS_OK; plen = 1;
                    kg = (gi >= 0 && gi < p.H && gj >= 0 && gj < W) ? (uint32_t)(gi * W + gj) : 0xffffffffu;
                    if (si < 0 || si >= p.H || sj < 0 || sj >= W) { status = NSFS_E_ARG; ks = 0; phase = PH_FINAL; }
                    else {
                        ks = (uint32_t)(si * W + sj);
                        // start cell: g = 0, then the first push into an empty list is a plain store (astar.cpp:45-52)
                        double h0 = 0.0;
                        if (ALGO == NSFS_ASTAR) {
                            const int ax = abs(gi - si), ay = abs(gj - si);        // dist_oct's slip: both against the ROW (bot_utils.cpp:47)
                            h0 = __dadd_rn(__dmul_rn((double)min(ax, ay), kSqrt2D), (double)abs(ax - ay));
                        }
                        const uint32_t f0 = (mode == NSFS_MODE_G) ? 0u : (uint32_t)__double2int_rz(h0);
                        if (f0 > (EB == 4 ? 0xfffu : 0xffffu)) { status = NSFS_E_HEAP; phase = PH_FINAL; }     // beyond the entry's sort key: the wide kernel replays it
                        else {
                            if (tl == 0) {
                                cell_store<CB>(cells, ks, 0u, 0u, 0u, 0u, epoch);
                                heap.st(1u, EB == 4 ? (E)((f0 << 20) | ks) : (E)(((unsigned long long)(f0 << 16) << 32) | (unsigned long long)ks));
                            }
                            n = 1; pushes = 1; peak = 1;
                            phase = PH_SEARCH;
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (__all_sync(FULL, phase == PH_DONE)) break;

        // ---- one pop + expansion for every searching team ------------------------------------------------------------
        {
            const bool act = phase == PH_SEARCH;
            E top = 0;
            if (act) top = heap.s[1];                                // the root is always in shared memory
            const uint32_t ku = entry_key<EB>(top);
            const int i = (int)(__umulhi(ku, p.div_m) >> p.div_s), j = (int)(ku - (uint32_t)i * uW);
            const uint32_t kv = ku + (uint32_t)my_off;               // my neighbour: flat key; its row follows the column wrap
            int vi = i + my_di;
            { const int vj = j + my_dj; if (vj >= W) vi++; else if (vj < 0) vi--; }
            // the records an expansion needs: every lane its neighbour's AND (same cache line for most) the popped cell's
            CellRec<CB> pre = {0u, 0u}, cu = {0u, 0u};
            uint32_t om_l = 0;
            if (act) {
                if (kv < uN) pre = cell_load<CB>(cells, kv);
                cu = cell_load<CB>(cells, ku);
                if (tl == 1) om_l = __ldg(p.out_mask + ku);
            }
            team_pop<MODE, EB>(heap, n, act, tl, tbase, nr, no);
            const uint32_t om = __shfl_sync(FULL, om_l, 1, 8);
            bool expand = act;
            if (act) {
                pops++;
                if (cu.visited()) expand = false;                    // stale entry (astar.cpp:59-62)
            }
            if (expand) {
                if (tl == 0) cell_store<CB>(cells, ku, cu.nc(), cu.nd(), cu.dir(), 1u, epoch);
                expansions++;
                if (ku == kg) { found = true; expand = false; phase = PH_WALK; wk = kg; plen = 0; }        // astar.cpp:68
                else if (om & kThrowBit) { status = NSFS_E_RANGE; expand = false; phase = PH_FINAL; }         // vector::at throws (R6)
            }
            // ---- my neighbour (astar.cpp:89-124) ----------------------------------------------------------------------
            bool relax = false, overflow = false;
            E entry = 0;
            if (expand && ((om >> tl) & 1u)) {
                const bool valid = pre.valid(epoch);
                const bool diag = ((om >> (8 + tl)) & 1u) && tl;
                const uint32_t nc = cu.nc() + (diag ? 0u : 1u), nd = cu.nd() + (diag ? 1u : 0u);
                // astar.cpp:116  g[v] > cand + 1e-5  on labels a + b*sqrt(2) held as exact step counts: with da = a_v - a, db = b_v - b the
                // difference da + db*sqrt(2) is 0 or at least 1 / (|da| + |db| sqrt(2)) > 1e-5 away from 0 (|da| < 2^16, |db| < 2^14), so the
                // test is the SIGN of da + db*sqrt(2): decided by the signs, or by da^2 against 2 db^2 when they differ.  Integers only.
                if (!valid) relax = true;
                else {
                    const int da = (int)pre.nc() - (int)nc, db = (int)pre.nd() - (int)nd;
                    if (db > -16384 && db < 16384) {
                        if (db == 0) relax = da > 0;
                        else if (da == 0) relax = db > 0;
                        else if ((da ^ db) >= 0) relax = da > 0;
                        else {
                            const unsigned long long a2 = (unsigned long long)((long long)da * da), b2 = 2ull * (unsigned long long)((long long)db * db);
                            relax = da > 0 ? a2 > b2 : b2 > a2;
                        }
                    } else {
                        const double cand = __dadd_rn((double)nc, __dmul_rn((double)nd, kSqrt2D));
                        const double gv = __dadd_rn((double)pre.nc(), __dmul_rn((double)pre.nd(), kSqrt2D));
                        relax = gv > __dadd_rn(cand, kSlack);
                    }
                }
                if (relax) {
                    if (nc > cell_max_steps<CB>() || nd > cell_max_steps<CB>()) { overflow = true; relax = false; }
                    else {
                        cell_store<CB>(cells, kv, nc, nd, (uint32_t)tl, valid ? pre.visited() : 0u, epoch);
                        // keys (grid_planner_core.cpp:385-391): (int)g = a + floor(b sqrt 2); (int)(g + h) with h = ord*sqrt(2) + card (dist_oct
                        // with the row slip, bot_utils.cpp:44-65) = a + card + floor((b + ord) sqrt 2).  k*sqrt(2) stays >= 1 / (2.83 k) away
                        // from every integer and one fp64 product is good to 1e-16 relative, so the truncations are the reference's.
                        const uint32_t gi32 = nc + (uint32_t)__double2int_rz(__dmul_rn((double)nd, kSqrt2D));
                        uint32_t fi = 0u;
                        if (mode != NSFS_MODE_G) {
                            if (ALGO == NSFS_ASTAR) {
                                const int ax = abs(gi - vi), ay = abs(gj - vi);
                                fi = nc + (uint32_t)abs(ax - ay) + (uint32_t)__double2int_rz(__dmul_rn((double)(nd + (uint32_t)min(ax, ay)), kSqrt2D));
                            } else fi = gi32;
                        }
                        if (EB == 4) {
                            if (fi > 0xfffu) { overflow = true; relax = false; }                   // (the record is written; the query is abandoned anyway)
                            entry = (E)((fi << 20) | kv);
                        } else {
                            if (fi > 0xffffu || gi32 > 0xffffu) { overflow = true; relax = false; }
                            entry = (E)(((unsigned long long)((fi << 16) | (gi32 & 0xffffu)) << 32) | (unsigned long long)kv);
                        }
                    }
                }
            }
            unsigned rb = team_ballot(relax, tbase);
            if (team_ballot(overflow, tbase)) { status = NSFS_E_HEAP; rb = 0; if (phase == PH_SEARCH) phase = PH_FINAL; }
            // ---- pushes, in NB_LUT order (astar.cpp:120) -------------------------------------------------------------
            while (__any_sync(FULL, rb != 0u)) {
                bool pushing = rb != 0u;
                const int src = pushing ? __ffs(rb) - 1 : 0;
                rb &= rb - 1u;
                const E v = __shfl_sync(FULL, entry, src, 8);
                if (pushing && (long long)n >= p.heap_cap) { status = NSFS_E_HEAP; pushing = false; rb = 0; phase = PH_FINAL; }
                team_push<MODE, EB>(heap, n, v, pushing, tl, tbase);
                if (pushing) { pushes++; if (n > peak) peak = n; }
            }
            if (phase == PH_SEARCH && n == 0u) phase = PH_FINAL;        // open list exhausted: unreachable (astar.cpp:129-135)
        }

        // ---- one step of the back-track for every team that found its goal (astar.cpp:70-82) ---------------------------
        if (__any_sync(FULL, phase == PH_WALK)) {
            const bool walking = phase == PH_WALK;
            uint32_t nk = 0;
            if (walking && tl == 0) {
                int32_t* out = p.paths ? p.paths + (size_t)qi * p.path_cap * 2 : nullptr;
                if (out && plen < p.path_cap) { out[2 * plen] = (int)(wk / uW); out[2 * plen + 1] = (int)(wk % uW); }
                if (wk != ks) {
                    const uint32_t d = cell_load<CB>(cells, wk).dir();
                    nk = wk - (uint32_t)(nb_di((int)d) * W + nb_dj((int)d));
                }
            }
            nk = __shfl_sync(FULL, nk, 0, 8);
            if (walking) {
                plen++;
                if (wk == ks || (long long)plen > p.N) phase = PH_FINAL;
                else wk = nk;
            }
        }

        // ---- results ------------------------------------------------------------------------------------------------
        if (phase == PH_FINAL) {
            if (tl == 0) {
                if (status == NSFS_OK) {
                    if (!found) {
                        plen = 1;
                        if (p.paths && p.path_cap > 0) {
                            int32_t* out = p.paths + (size_t)qi * p.path_cap * 2;
                            out[0] = (int)(ks / uW); out[1] = (int)(ks % uW);
                        }
                    }
                    if (p.g_goal) {
                        double gg = kInfD;
                        if (kg != 0xffffffffu) {
                            const CellRec<CB> c = cell_load<CB>(cells, kg);
                            if (c.valid(epoch)) gg = __dadd_rn((double)c.nc(), __dmul_rn((double)c.nd(), kSqrt2D));
                        }
                        p.g_goal[qi] = gg;
                    }
                } else {
                    plen = 1;
                    if (p.g_goal) p.g_goal[qi] = kInfD;
                }
                if (p.status) p.status[qi] = status;
                if (p.path_len) p.path_len[qi] = plen;
                if (p.settled) p.settled[qi] = expansions;
                if (p.stats5) {
                    atomicAdd((unsigned long long*)p.stats5 + 0, (unsigned long long)pops);
                    atomicAdd((unsigned long long*)p.stats5 + 1, (unsigned long long)expansions);
                    atomicAdd((unsigned long long*)p.stats5 + 2, (unsigned long long)pushes);
                    atomicMax((unsigned long long*)p.stats5 + 3, (unsigned long long)peak);
                }
            }
            phase = PH_FETCH;
        }
        __syncwarp();
    }
    if (have_slot && tl == 0) p.epoch[slot] = epoch;
}

// ---- start order ---------------------------------------------------------------------------------------------------------
// The slots fetch queries from one counter, so a launch lasts until the query that STARTED last has finished, and the work per
// query spans 200 .. 900 000 expansions (C4).  Longest-expected-first keeps the tail short.  Expected work: for Djikstra the
// Chebyshev distance start -> goal; for Astar a linear form in |di|, |dj|, their maximum and the three terms the reference's
// slipped heuristic brings in (h looks at the node's ROW only: |gj - gi| = h(goal) inflates the search, |gj - si| and |sj - si|
// shrink it), fitted on 320 C4 queries against the reference's own settled counts: Spearman rank correlation 0.98 there and
// 0.95 - 0.97 on other maps, sizes, densities and cost mode fg (0.60 - 0.85 for the distance alone).
// One block: histogram of 1024 buckets in shared memory, exclusive scan, scatter.  Only the ORDER of starting changes; every
// query is still replayed on its own and written to its own output index.
__global__ void __launch_bounds__(1024) team_order_kernel(const int32_t* __restrict__ q, int nq, int span, int astar, int32_t* __restrict__ order) {
    __shared__ int s_cnt[1024];
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_cnt[tid] = 0;
    __syncthreads();
    const long long full = (astar ? 1500ll : 1ll) * (span > 0 ? span : 1);
    auto bucket = [&](int k) {
        const int si = q[4 * k], sj = q[4 * k + 1], gi = q[4 * k + 2], gj = q[4 * k + 3];
        const long long di = abs(gi - si), dj = abs(gj - sj), ch = di > dj ? di : dj;
        long long e = ch;
        if (astar) e = 147 * di + 219 * dj + 530 * ch + 587 * (long long)abs(gj - gi) - 232 * (long long)abs(gj - si) - 180 * (long long)abs(sj - si);
        const long long d = e * 1023 / full;
        return 1023 - (int)(d < 0 ? 0 : (d > 1023 ? 1023 : d));      // bucket 0 = the longest
    };
    for (int k = tid; k < nq; k += 1024) atomicAdd(&s_cnt[bucket(k)], 1);
    __syncthreads();
    // exclusive scan over the 1024 counts: warp scans, then the warp totals
    const int c = s_cnt[tid];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    s_cnt[tid] = s_warp[warp] + incl - c;                            // first output position of bucket tid
    __syncthreads();
    for (int k = tid; k < nq; k += 1024) order[atomicAdd(&s_cnt[bucket(k)], 1)] = k;
}

void launch_start_order(nsfs_planner* h, const int32_t* d_q, int nq, bool heuristic, int32_t* d_order) {
    team_order_kernel<<<1, 1024, 0, h->stream>>>(d_q, nq, h->H > h->W ? h->H : h->W, heuristic ? 1 : 0, d_order);
    h->launches++;
}

// ---- host side --------------------------------------------------------------------------------------------
// 4-byte records for maps up to 2048 cells a side (12-bit step counts), 8-byte records beyond; "team_cell_bytes" overrides.
static int team_cell_bytes(const nsfs_planner* h) {
    if (h->opt_team_cell_bytes == 4 || h->opt_team_cell_bytes == 8) return (int)h->opt_team_cell_bytes;
    return (h->H <= 2048 && h->W <= 2048) ? 4 : 8;
}

// 4-byte open-list entries where the comparator is "f" and a key fits 20 bits; decided per LAUNCH (the slot is sized for 8-byte ones)
static int team_entry_bytes(const nsfs_planner* h, int mode) {
    if (h->opt_team_entry_bytes == 8) return 8;
    return (mode == NSFS_MODE_F && h->N <= (1ll << 20)) ? 4 : 8;
}

// Teams in flight for a batch of nq.  More teams in flight raise the throughput (C4, 65 536 queries: 5 820 queries/s on 16 576
// slots, 6 276 on 21 312) but every query then runs slower, and a launch lasts as long as its slowest query: a batch that fits the
// slots less than twice over finishes soonest with HALF of it in flight (32 768 queries: 6.5 s on 16 384 slots, 7.0 s on 12 288,
// 7.9 s on 21 312; 16 384 queries: 5.4 s on 8 192 slots, 5.6 s on 12 288, 6.1 s on 16 384; profiles/r02_search.md).
static long long team_slots_wanted(const nsfs_planner* h, int nq, long long cap) {
    long long want = nq < 1 ? 1 : nq;
    if (h->opt_max_slots <= 0 && nq > 64) want = (((long long)nq + 1) / 2 + 15) & ~15ll;      // whole blocks of 16 teams
    return want < cap ? want : cap;
}

int team_alloc(nsfs_planner* h, int nq) {
    TeamWork& s = h->team;
    const int cb = team_cell_bytes(h);
    // open-list slot: N / 8 entries with 8-byte records (round 1), N / 32 with 4-byte records, where the point is the footprint
    // (a C4 query peaks at ~11 000 entries of the 32 768 it then gets); an overflow re-runs the query on the wide kernel either way
    const long long auto_cap = cb == 4 ? h->N / 32 : h->N / 8;
    long long heap_cap = h->opt_team_heap_cap > 0 ? h->opt_team_heap_cap : (auto_cap > 4096 ? auto_cap : 4096);
    if (heap_cap > (1ll << 31) - 64) heap_cap = (1ll << 31) - 64;
    // warps of slots per SM: 28 with 8-byte records (what 0.85 of a B200's HBM holds at 1024^2, 9 MB per slot), 36 with 4-byte
    // records (the register file's limit at 55 registers per thread; 4.25 MB per slot)
    const long long cap = h->opt_max_slots > 0 ? h->opt_max_slots * 4 : (long long)h->sm_count * (cb == 4 ? 36 : 28) * 4;
    long long slots = team_slots_wanted(h, nq, cap);
    // reuse: enough slots, or an allocation made for at least this many that free HBM clamped (see search_alloc)
    if (s.cells && s.cell_bytes == cb && s.heap_cap >= heap_cap && (s.slots >= slots || s.target >= slots)) return NSFS_OK;
    team_free(h);
    const long long target = slots;
    size_t free_b = 0, total_b = 0;
    NSFS_CUDA_TRY(h, cudaMemGetInfo(&free_b, &total_b));
    const long long heap_stride = (heap_cap + 5) & ~1ll;
    const size_t cell_stride = (((size_t)h->N * cb + 15) / 16) * 16;      // per-slot record arrays stay 16-byte aligned (the epoch wipe stores uint4)
    const size_t per_slot = cell_stride + (size_t)heap_stride * sizeof(unsigned long long);
    const long long fit = (long long)((double)free_b * 0.85 / (double)per_slot);
    if (slots > fit) slots = fit;
    if (slots < 1) return NSFS_E_NOMEM;
    NSFS_CUDA_TRY(h, cudaMalloc(&s.cells, (size_t)slots * cell_stride));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.heap, (size_t)slots * (size_t)heap_stride * sizeof(unsigned long long)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.epoch, (size_t)slots * sizeof(uint32_t)));
    NSFS_CUDA_TRY(h, cudaMalloc(&s.work_counter, sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.cells, 0, (size_t)slots * cell_stride, h->stream));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.epoch, 0, (size_t)slots * sizeof(uint32_t), h->stream));
    s.slots = (int)slots;
    s.target = (int)target;
    s.heap_cap = heap_cap;
    s.cell_bytes = cb;
    return NSFS_OK;
}

void team_free(nsfs_planner* h) {
    TeamWork& s = h->team;
    cudaFree(s.cells); cudaFree(s.heap); cudaFree(s.epoch); cudaFree(s.work_counter); cudaFree(s.order);
    s = TeamWork();
}

int launch_search_team(nsfs_planner* h, int algo, int mode, int nq, const int32_t* d_q, int32_t* d_status, double* d_g_goal,
                       int32_t* d_path_len, long long* d_settled, int32_t* d_paths, int path_cap, long long* d_stats5,
                       const int32_t* ext_order, cudaStream_t ext_stream, bool part) {
    if (nq <= 0) return NSFS_OK;
    cudaStream_t stream = ext_stream ? ext_stream : h->stream;
    int rc = team_alloc(h, nq);
    if (rc != NSFS_OK) return rc;
    TeamWork& s = h->team;
    TeamParams p;
    p.out_mask = h->d_out_mask; p.cells = s.cells; p.heap = s.heap; p.epoch = s.epoch; p.work_counter = s.work_counter;
    p.H = h->H; p.W = h->W; p.N = h->N; p.heap_cap = s.heap_cap; p.heap_stride = (s.heap_cap + 5) & ~1ll;
    p.cell_stride = (long long)((((size_t)h->N * s.cell_bytes + 15) / 16) * 16);
    p.slots = (int)team_slots_wanted(h, nq, s.slots);
    p.mode = (mode == NSFS_MODE_F || mode == NSFS_MODE_G) ? mode : NSFS_MODE_FG;
    p.nq = nq; p.q = d_q; p.status = d_status; p.g_goal = d_g_goal; p.path_len = d_path_len; p.settled = d_settled;
    p.paths = d_paths; p.path_cap = path_cap; p.stats5 = d_stats5;
    p.order = ext_order;
    if (!part && nq > p.slots && h->opt_search_order != 2) {         // more queries than slots: start the long ones first
        if (s.order_cap < nq) {
            cudaFree(s.order); s.order = nullptr; s.order_cap = 0;
            NSFS_CUDA_TRY(h, cudaMalloc(&s.order, (size_t)nq * sizeof(int32_t)));
            s.order_cap = nq;
        }
        launch_start_order(h, d_q, nq, algo == NSFS_ASTAR, s.order);
        p.order = s.order;
    }
    NSFS_CUDA_TRY(h, cudaMemsetAsync(s.work_counter, 0, sizeof(int), stream));
    const int block = TEAM_WARPS_PER_BLOCK * 32;
    const int per_block = TEAM_WARPS_PER_BLOCK * 4;
    const int grid = (p.slots + per_block - 1) / per_block;
    {   // exact division of a flat key by W (key < 2^29): q = umulhi(key, m) >> s with m = ceil(2^(32+s) / W), 32 + s >= 29 + ceil(log2 W)
        int l = 0;
        while ((1ll << l) < (long long)h->W) ++l;
        int sh = 29 + l - 32;
        if (sh < 0) sh = 0;
        const unsigned long long two = 1ull << (32 + sh);
        p.div_m = (uint32_t)((two + (unsigned long long)h->W - 1ull) / (unsigned long long)h->W);
        p.div_s = sh;
    }
    if (algo != NSFS_ASTAR && algo != NSFS_DJIKSTRA) return NSFS_E_ARG;
    if (!part) NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, stream));
    // 16 KB of shared memory per block (the open lists' top levels) x 9 resident blocks: ask for a carveout that holds them
    const int carve = h->carveout > 0 ? h->carveout : 70;
#define NSFS_TEAM_LAUNCH2(A, M, C, B) do { cudaFuncSetAttribute(search_team_kernel<A, M, C, B>, cudaFuncAttributePreferredSharedMemoryCarveout, carve); \
                                          search_team_kernel<A, M, C, B><<<grid, block, 0, stream>>>(p); } while (0)
#define NSFS_TEAM_LAUNCH1(A, M, B) do { if (s.cell_bytes == 4) NSFS_TEAM_LAUNCH2(A, M, 4, B); else NSFS_TEAM_LAUNCH2(A, M, 8, B); } while (0)
#define NSFS_TEAM_LAUNCH(A, M) NSFS_TEAM_LAUNCH1(A, M, 8)
    const bool small_entries = team_entry_bytes(h, p.mode) == 4;
    if (small_entries) {
        p.heap_stride *= 2;                    // the same slot bytes, counted in 4-byte entries
        if (algo == NSFS_ASTAR) NSFS_TEAM_LAUNCH1(NSFS_ASTAR, NSFS_MODE_F, 4); else NSFS_TEAM_LAUNCH1(NSFS_DJIKSTRA, NSFS_MODE_F, 4);
    } else if (algo == NSFS_ASTAR) {
        if (p.mode == NSFS_MODE_F) NSFS_TEAM_LAUNCH(NSFS_ASTAR, NSFS_MODE_F);
        else if (p.mode == NSFS_MODE_G) NSFS_TEAM_LAUNCH(NSFS_ASTAR, NSFS_MODE_G);
        else NSFS_TEAM_LAUNCH(NSFS_ASTAR, NSFS_MODE_FG);
    } else {
        if (p.mode == NSFS_MODE_F) NSFS_TEAM_LAUNCH(NSFS_DJIKSTRA, NSFS_MODE_F);
        else if (p.mode == NSFS_MODE_G) NSFS_TEAM_LAUNCH(NSFS_DJIKSTRA, NSFS_MODE_G);
        else NSFS_TEAM_LAUNCH(NSFS_DJIKSTRA, NSFS_MODE_FG);
    }
#undef NSFS_TEAM_LAUNCH
#undef NSFS_TEAM_LAUNCH1
#undef NSFS_TEAM_LAUNCH2
    if (!part) NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, stream));
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
