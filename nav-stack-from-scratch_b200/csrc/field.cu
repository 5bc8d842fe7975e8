// Djikstra cost-to-go field by tile-local relaxation, predecessor field, predecessor walk.
//
// What is computed (reference djikstra.cpp:27-145 run with an unreachable goal): the unique
// fixed point of   g[v] = min over in-edges (u -> v) of g[u] + w(u, v),   g[source] = 0,
// on the reference's quirk graph: flat-key stencil, row-only bounds test, direction cost 1 / sqrt(2) decided
// by the parity of passable directions before it (all folded into in_mask by map_pack.cu).  Because every
// edge costs >= 1 and the reference's open list is keyed on (int)g, its label-setting order cannot change
// the result, so a label-correcting solve in any order reproduces it.
//
// Arithmetic: Q17.15 fixed point in uint32 (kQOne = 2^15 per cost unit, sqrt(2) -> 46341 / 32768, 1e5 -> kInfQ).
// Integer sums are exact, so the fixed point does not depend on the schedule at all and every path's cost differs from
// the reference's fp64 sum by at most |46341/32768 - sqrt(2)| / sqrt(2) = 1.08e-6 relative whatever its length
// (fp32 sums drift past the 1e-5 bar beyond ~3000 steps: measured 1.08e-5 at 4096^2).  nsfs_field_finish converts
// the plane to fp32 cost units in place.
//
// Structure: the grid is cut into 64x64 tiles.  A persistent kernel keeps SEVERAL tiles resident per SM (a thread block
// of WPT warps per tile; WPT = 4 => 8 tiles per SM): a wavefront is a line through a tile, so one visit only has work
// for a couple of warps, and what the machine needs is many visits in flight (round-1 trace: one 1024-thread visit per SM
// kept ~2 of 32 warps busy; every block was busy 70-90 % of the time on C2 / C5, i.e. tiles QUEUED for their owner).
//   between tiles  delta-stepping on tile keys: tile_key[t] = the smallest value a neighbour moved in t's halo since t was
//                  last relaxed; tile t belongs to block t % grid; a block takes its pending tile with the smallest key and
//                  relaxes it within the band [key, key + delta]; candidates beyond the band become the tile's next key
//   inside a tile  the tile and a 1-cell halo (addressed by FLAT key: the reference's column wrap comes for free) and its
//                  edge masks are staged in shared memory; a warp owns NSB / WPT sub-blocks of 8x4 cells (one lane per
//                  cell), each with a due KEY; the warp finds its smallest due key with __reduce_min_sync / __ballot_sync
//                  over its lanes, relaxes that sub-block for up to `inner` warp-local passes (8 shared loads, a min tree,
//                  one band test; __ballot_sync tells when it is quiet), writes the cells that changed back to L2 and
//                  lowers the due keys of the neighbouring sub-blocks / the keys of the neighbouring tiles (cut-through:
//                  the next tile starts on the wavefront while this one is still relaxing)
//   termination    one counter of pending tiles + visits in flight; a visit that is about to end first absorbs posts
//                  that arrived for its own tile meanwhile (re-reads the halo ring, stays resident)
#include <cstdio>
#include <cstring>
#include <vector>

#include "nsfs_internal.cuh"

namespace nsfs {

#ifndef NSFS_FIELD_TRACE
#define NSFS_FIELD_TRACE 0        // 1 = instrumented dev build (dev/trace_field.py): per-visit timestamps + first key post per tile
#endif
#if NSFS_FIELD_TRACE
// trace layout (unsigned long long words, p.prof): [0] number of visit records, [16 + t] first key post to tile t
// (globaltimer ns), then 8-word visit records from word 16 + n_tiles rounded up to 8
constexpr int kTraceMaxVisits = 1 << 19;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TRACE_POST(p, tile) do { if ((p).prof) atomicMin((p).prof + 16 + (tile), gtime()); } while (0)
#else
#define TRACE_POST(p, tile) do { } while (0)
#endif

constexpr int TS = kFieldTile;         // tile side in cells
constexpr int SBW = 8, SBH = 4;        // sub-block: 8 columns x 4 rows = one lane per cell
constexpr int SB_X = TS / SBW;         // 8
constexpr int SB_Y = TS / SBH;         // 16
constexpr int NSB = SB_X * SB_Y;       // 128 sub-blocks per tile
constexpr int SROW = TS + 8;           // smem row stride in words; SROW % 32 == 8 => the 4 rows of a sub-block use disjoint banks
constexpr int SCOL0 = 4;               // smem column of tile column 0 (halo column -1 at 3)
constexpr int SROWS = TS + 2;
constexpr int kDummySlot = 0;          // s[0] (row 0, column 0: padding left of the halo column) always holds kInfQ: "no in-edge"

// Which sub-blocks a warp owns.  Only the sub-blocks on the wavefront have work, so what counts is how many WARPS the
// front's sub-blocks belong to.  warp = (sy + sx) mod WPT on the 16 x 8 grid of sub-blocks: the sides of the wave's octagon are
// rows, columns and cell diagonals, i.e. sub-block lines (sy), (sx) and (sx, +-2 sx) (a sub-block is 8 wide, 4 tall); their
// increments 1, 1, 1 +- 2 are all odd, so each such line is spread evenly over the WPT warps.
// (32 warps: (sy + sx) only takes 23 values, so the lattice (2 sy + 5 sx) mod 32 of round 1 is used: rows, columns and both cell
// diagonals of sub-blocks fall on distinct warps, two sub-blocks of one warp are >= 25 cells apart.)
template <int WPT>
__device__ __forceinline__ int sb_of(int warp, int slot) {           // (warp, slot) -> sub-block index sy * SB_X + sx
    if (WPT == 32) { const int sx = 2 * slot + (warp & 1); return (((warp - 5 * sx) & 31) >> 1) * SB_X + sx; }
    if (WPT == 4) { const int sy = slot >> 1; return sy * SB_X + ((warp - sy) & 3) + 4 * (slot & 1); }
    if (WPT == 8) { return slot * SB_X + ((warp - slot) & 7); }
    return ((warp - slot) & 15) * SB_X + slot;                        // WPT == 16: slot = sx
}
template <int WPT>
__device__ __forceinline__ int slot_of(int t) {                      // sub-block index -> index in s_due (warp-major)
    const int sy = t / SB_X, sx = t % SB_X;
    constexpr int SPW = NSB / WPT;
    if (WPT == 32) return ((2 * sy + 5 * sx) & 31) * SPW + (sx >> 1);
    const int warp = (sy + sx) & (WPT - 1);
    if (WPT == 4) return warp * SPW + sy * 2 + (sx >> 2);
    if (WPT == 8) return warp * SPW + sy;
    return warp * SPW + sx;
}

__global__ void field_fill_kernel(uint4* __restrict__ g4, long long n4, uint32_t* __restrict__ g, long long N,
                                  uint32_t* __restrict__ tile_key, int n_tiles) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4 v = make_uint4(kInfQ, kInfQ, kInfQ, kInfQ);
    for (long long k = t; k < n4; k += stride) g4[k] = v;
    for (long long k = n4 * 4 + t; k < N; k += stride) g[k] = kInfQ;
    for (long long k = t; k < n_tiles; k += stride) tile_key[k] = kNoKey;
}

// One thread: g[source] = 0 and the first pending tiles; a non-free source is expanded here because in_mask only
// carries edges out of free cells (plan() pushes the start cell unconditionally, astar.cpp:45-52).
__global__ void field_seed_kernel(FieldParams p, const uint32_t* __restrict__ free_bits,
                                  const uint16_t* __restrict__ out_mask, int si, int sj) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (si < 0) {                               // no source on this shard: everything arrives through field_inject_kernel
        for (int c = 0; c < 16; ++c) p.counts[c] = 0;
        return;
    }
    const long long ks = (long long)si * p.W + sj;
    p.g[ks] = 0u;
    auto tile_of = [&](long long k) { return (int)(k / p.W) / TS * p.tiles_x + (int)(k % p.W) / TS; };
    int pend = 0;
    auto seed_tile = [&](int t) { if (p.tile_key[t] == kNoKey) ++pend; p.tile_key[t] = 0u; };
    seed_tile(tile_of(ks));
    for (int d = 0; d < 8; ++d) {              // and the tiles that read the source as halo (it never "changes" in a visit, so nobody posts it)
        const long long kv = ks + nb_off(d, p.W);
        if (kv >= 0 && kv < p.N) seed_tile(tile_of(kv));
    }
    if (!bit_at(free_bits, ks)) {
        const uint32_t om = out_mask[ks];
        for (int d = 0; d < 8; ++d) {
            if (!((om >> d) & 1u)) continue;
            const long long kv = ks + nb_off(d, p.W);
            const uint32_t w = ((om >> (8 + d)) & 1u) && d ? kQDiag : kQOne;
            if (p.g[kv] > w) p.g[kv] = w;
            seed_tile(tile_of(kv));
        }
    }
    for (int c = 0; c < 16; ++c) p.counts[c] = 0;
    p.counts[1] = pend;                        // pending tiles + visits in flight
}

// Row injection for the row-partitioned multi-GPU field: g[row0*W + t] = min(g, vals[t]) over whole rows
// (vals in the plane's working representation, i.e. rows read from another shard's plane during a solve).
// An improved cell makes its tile pending with that value as key, exactly like a halo post from a neighbouring tile.
__global__ void __launch_bounds__(256) field_inject_kernel(FieldParams p, long long k0, long long n, const uint32_t* __restrict__ vals,
                                                           int* __restrict__ improved) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool better = false;
    if (t < n) {
        const long long k = k0 + t;
        const uint32_t v = vals[t], old = p.g[k];
        if (v < old) {
            p.g[k] = v;
            better = true;
            // pending: the tile that holds the cell AND every tile that reads it as halo, i.e. the tiles of its out-neighbours
            // (flat-key arithmetic, so the column wrap of the reference's stencil is covered too)
            int last = -1;
#pragma unroll
            for (int d = -1; d < 8; ++d) {
                const long long kv = d < 0 ? k : k + nb_off(d, p.W);
                if (kv < 0 || kv >= p.N) continue;
                const int tile = (int)(kv / p.W) / TS * p.tiles_x + (int)(kv % p.W) / TS;
                if (tile == last) continue;
                last = tile;
                const uint32_t was = atomicMin(p.tile_key + tile, v);
                if (was == kNoKey) atomicAdd(p.counts + 1, 1);
            }
        }
    }
    const uint32_t b = __ballot_sync(0xffffffffu, better);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(improved, __popc(b));
}

// Peer-to-peer shards: before the solve kernel, whatever is already finite in my first / last two OWNED rows goes to the
// neighbour's ghost rows.  The solve kernel only pushes cells that CHANGE in a visit, and the source (g = 0 from the seed, and
// the cells a non-free source seeds) never does: a source in a shard's boundary rows would stay unknown to the neighbour, whose
// cells next to it would then settle on a detour.  An improved ghost cell makes the neighbour's tiles that hold or read it
// pending, exactly as field_inject_kernel does locally, and is added to the shared counter.
__global__ void __launch_bounds__(256) field_push_boundary_kernel(FieldParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;            // 4 rows x W: rows own_lo, own_lo + 1 (side 0), own_hi - 2, own_hi - 1 (side 1)
    if (t >= 4 * p.W) return;
    const int which = t / p.W, c = t - which * p.W, side = which >> 1;
    const int r = side == 0 ? p.own_lo + (which & 1) : p.own_hi - 2 + (which & 1);
    if (!p.peer_g[side] || r < p.own_lo || r >= p.own_hi) return;
    const uint32_t v = __ldcg(p.g + (long long)r * p.W + c);
    if (v >= kInfQ) return;
    const int rp = r + p.peer_row_off[side];                          // the neighbour's local row
    const long long Np = (long long)p.peer_rows[side] * p.W, kp = (long long)rp * p.W + c;
    if (rp < 0 || kp >= Np) return;
    const uint32_t old = atomicMin(p.peer_g[side] + kp, v);
    if (old <= v) return;
    __threadfence_system();
    int last = -1;
#pragma unroll
    for (int d = -1; d < 8; ++d) {
        const long long kv = d < 0 ? kp : kp + nb_off(d, p.W);
        if (kv < 0 || kv >= Np) continue;
        const int tile = (int)(kv / p.W) / TS * p.tiles_x + (int)(kv % p.W) / TS;
        if (tile == last) continue;
        last = tile;
        const uint32_t was = atomicMin(p.peer_key[side] + tile, v);
        if (was == kNoKey) atomicAdd(p.gcount, 1);
    }
}

// Key post with release semantics: orders this thread's earlier stores, and those of other threads it synchronised with
// (__syncwarp / __syncthreads), before the key: MEMBAR.ALL.GPU + ATOMG, against MEMBAR.SC.GPU + CCTL.IVALL (an L1 invalidate)
// + ATOMG for __threadfence() followed by a relaxed atomicMin.  Message passing needs no more than release / acquire.
__device__ __forceinline__ uint32_t atomic_min_release(uint32_t* addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.release.gpu.global.min.u32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}

// Wake the neighbour shard's tiles that hold, or read as halo, the two boundary rows this shard pushes: its rows R-3 .. R+4 lie in
// at most two tile rows; columns tx-1 .. tx+1 plus the far edge column when this tile touches a map edge (flat-key wrap).
// slot 0..7 = (tile row, column choice); returns 1 if the post made a tile pending.
template <int WPT>
__device__ __forceinline__ int peer_post(const FieldParams& p, int side, int slot, int tx, uint32_t m) {
    if (!p.peer_key[side] || m == kNoKey) return 0;
    const int R = (side == 0 ? p.own_lo : p.own_hi - 2) + p.peer_row_off[side];      // first pushed row, neighbour-local
    const int tr0 = max(R - 3, 0) / TS, tr1 = min(R + 4, p.peer_tiles_y[side] * TS - 1) / TS;
    const int trow = (slot & 1) ? tr1 : tr0;
    const int cs = slot >> 1;                                      // 0..2: tx-1..tx+1, 3: the wrap column
    const int tcol = cs < 3 ? tx + cs - 1 : (tx == 0 ? p.tiles_x - 1 : (tx == p.tiles_x - 1 ? 0 : -1));
    bool dup = ((slot & 1) && tr1 == tr0) || tcol < 0 || tcol >= p.tiles_x;
    if (cs == 3 && !dup) dup = abs(tcol - tx) <= 1;                // already covered by the three neighbour columns
    if (dup) return 0;
    const uint32_t old = atomicMin(p.peer_key[side] + trow * p.tiles_x + tcol, m);
    return (old > p.cap && m <= p.cap) ? 1 : 0;
}

template <int WPT>
struct __align__(16) TileSmem {
    uint32_t s[SROWS * SROW];      // the tile and its halo: position (rr, cc) holds g[(r0-1+rr)*W + (c0-1+cc)] (flat key!)
    uint16_t mask[TS * TS];        // in_mask of the tile's cells (0 for cells beyond the map)
    uint32_t due[NSB];             // per sub-block: the smallest value that changed next to it (kNoKey = not due), warp-major
    uint32_t emin[16];             // 9 / 10: map-edge wrap posts, 11: this tile's own next key, 12 / 13: peer shards
    int ctl[4];                    // 0 idle warps, 1 done, 2 net change of the pending counter, 3 merges
    unsigned long long best;
    uint32_t key;
#if NSFS_FIELD_TRACE
    unsigned long long tr[8];
    unsigned tr_it[2];
    unsigned long long cyc[8];     // warp-cycles summed over the visit's warps: 0 due scan + idle, 1 claim + set-up, 2 passes, 3 in-tile posts, 4 write-back, 5 fence + tile posts, 6 sub-block turns
#endif
};

// Relax one tile to its local fixed point under the band limit thr.  The warps run ASYNCHRONOUSLY: no block barrier per
// sweep.  Relaxation is monotone (values only decrease towards the fixed point), so a stale read only costs an extra pass
// and a due key lowered after the owner cleared it simply re-queues the sub-block.  Termination: ctl[0] counts warps that
// found nothing due; the warp that makes it WPT verifies that every due key is clear and that the count still stands.
template <int WPT>
__device__ __forceinline__ void relax_tile(const FieldParams& p, int tile, uint32_t thr, TileSmem<WPT>& sm, uint32_t key0) {
    constexpr int THREADS = WPT * 32;
    constexpr int SPW = NSB / WPT;                                    // sub-blocks per warp (<= 32: one lane each in the due scan)
    volatile uint32_t* s = sm.s;
    volatile uint32_t* due = sm.due;
    volatile int* ctl = sm.ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tile / p.tiles_x, tx = tile % p.tiles_x;
    const int r0 = ty * TS, c0 = tx * TS;
    const int W = p.W, H = p.H;
    const long long N = p.N;
    const bool peers = p.peer_g[0] || p.peer_g[1];
#if NSFS_FIELD_TRACE
    unsigned tr_iters = 0;
    if (tid == 0) { sm.tr[1] = gtime(); sm.tr[6] = ~0ull; sm.tr[7] = 0; sm.tr_it[0] = 0; sm.tr_it[1] = 0; }
    if (tid < 8) sm.cyc[tid] = 0;
    long long c_scan = 0, c_setup = 0, c_pass = 0, c_post = 0, c_wb = 0, c_fence = 0, c_turns = 0, c_t = 0;
#define CYC_MARK(acc) do { const long long _n = clock64(); acc += _n - c_t; c_t = _n; } while (0)
#else
#define CYC_MARK(acc) do { } while (0)
#endif
    // ---- stage the tile, its halo and its edge masks -----------------------------------------------------------
    for (int idx = tid; idx < SROWS * (TS + 2); idx += THREADS) {
        const int rr = idx / (TS + 2), cc = idx - rr * (TS + 2);
        const long long key = (long long)(r0 - 1 + rr) * W + (c0 - 1 + cc);
        sm.s[rr * SROW + (SCOL0 - 1) + cc] = (key >= 0 && key < N) ? __ldcg(p.g + key) : kInfQ;   // L2: other SMs write g
    }
    for (int idx = tid; idx < TS * TS; idx += THREADS) {
        const int ly = idx / TS, lx = idx - ly * TS;
        const int r = r0 + ly, c = c0 + lx;
        sm.mask[idx] = (r < H && c < W) ? __ldg(p.in_mask + (long long)r * W + c) : (uint16_t)0;
    }
    if (tid == 0) sm.s[kDummySlot] = kInfQ;
    for (int t = tid; t < NSB; t += THREADS) sm.due[t] = key0;       // first pass: everything is due at the tile's own key
    if (tid < 16) sm.emin[tid] = kNoKey;
    if (tid < 4) sm.ctl[tid] = 0;
    __syncthreads();
#if NSFS_FIELD_TRACE
    if (tid == 0) sm.tr[2] = gtime();
    c_t = clock64();
#endif

    const int lr = lane >> 3, lc = lane & 7;
    uint32_t rej = kNoKey;                                            // smallest candidate the band rejected (this lane, any sub-block)
    bool counted_idle = false;
    unsigned nap = p.wait_ns;                                         // idle back-off: short right after work, doubling up to idle_ns
    for (int spin = 0; spin < (1 << 26); ++spin) {
        // ---- my due sub-block with the smallest key: one lane per owned sub-block, reduce + ballot ----------------
        uint32_t kq = kNoKey;
        if (lane < SPW) kq = due[warp * SPW + lane];
        const uint32_t mymin = __reduce_min_sync(0xffffffffu, kq);
        if (mymin != kNoKey) {
            CYC_MARK(c_scan);
            const int q = __ffs(__ballot_sync(0xffffffffu, kq == mymin)) - 1;
            const int sb = sb_of<WPT>(warp, q);
            if (counted_idle) {                                       // leave the idle count BEFORE consuming the key
                if (lane == 0) atomicSub(sm.ctl, 1);
                counted_idle = false;
            }
            __syncwarp();
            if (lane == 0) atomicExch(sm.due + warp * SPW + q, kNoKey);   // clear first, then read the neighbours
            __syncwarp();
            const int sby = sb / SB_X, sbx = sb % SB_X;
            const int ly = sby * SBH + lr, lx = sbx * SBW + lc;
            const int so = (ly + 1) * SROW + SCOL0 + lx;
            const uint32_t m = sm.mask[ly * TS + lx];
            uint32_t vq = s[so];
            // Per direction, decided ONCE per turn and kept in registers: the weight (1 or sqrt(2) in Q17.15) and WHERE the candidate's
            // source is read from: the neighbour's slot, or, for a direction without an in-edge, a dummy slot that holds kInfQ
            // (kInfQ + weight never undercuts a cell).  A pass is then 8 shared loads, 8 adds, a min tree and one band test: the
            // single warp that walks a sub-block issues dependent instructions ~5 cycles apart, so the pass's LENGTH in instructions
            // is its latency (trace, round 2: 75 instructions and ~650 cycles per pass when the compiler re-derived the weights
            // from the mask inside the loop).  The empty asm keeps the compiler from doing that again.
            uint32_t wgt[8];
            int src[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                wgt[d] = ((m >> (8 + d)) & 1u) ? kQDiag : kQOne;
                src[d] = ((m >> d) & 1u) ? so - nb_di(d) * SROW - nb_dj(d) : kDummySlot;
                asm volatile("" : "+r"(wgt[d]), "+r"(src[d]));
            }
            uint32_t bits = 0, last = 0, kchg = kNoKey;
            CYC_MARK(c_setup);
            for (int it = 0; it < p.inner; ++it) {
                uint32_t cd[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) cd[d] = s[src[d]] + wgt[d];                                   // <= kInfQ + kQDiag: no wrap
                // min as a tree (3 levels instead of a chain of 8); the band test on the minimum equals the test on every candidate
                const uint32_t best = min(min(min(cd[0], cd[1]), min(cd[2], cd[3])), min(min(cd[4], cd[5]), min(cd[6], cd[7])));
                const bool lower = best < vq;
                const bool changed = lower && best <= thr;
                if (changed) { vq = best; s[so] = best; kchg = min(kchg, best); }
                else if (lower) rej = min(rej, best);                 // beyond the band: the tile's own next key (at most)
#if NSFS_FIELD_TRACE
                tr_iters++;
#endif
                last = __ballot_sync(0xffffffffu, changed);
                if (!last) break;
                bits |= last;
                __syncwarp();
            }
            CYC_MARK(c_pass);
#if NSFS_FIELD_TRACE
            c_turns++;
#endif
            if (bits) {
#if NSFS_FIELD_TRACE
                if (lane == 0) { const unsigned long long tn = gtime(); atomicMin(&sm.tr[6], tn); atomicMax(&sm.tr[7], tn); }
#endif
                const uint32_t kmin = __reduce_min_sync(0xffffffffu, kchg);
                __syncwarp();
                // ---- wake the neighbouring sub-blocks whose border moved (self: only if the pass budget ran out mid-change)
                if (lane < 9) {
                    const int dy = lane / 3 - 1, dx = lane % 3 - 1;
                    uint32_t need = (dy == 0 && dx == 0) ? last : bits;
                    if (dy < 0) need &= 0x000000ffu; else if (dy > 0) need &= 0xff000000u;
                    if (dx < 0) need &= 0x01010101u; else if (dx > 0) need &= 0x80808080u;
                    const int sy = sby + dy, sx = sbx + dx;
                    if (need && sy >= 0 && sy < SB_Y && sx >= 0 && sx < SB_X) atomicMin(sm.due + slot_of<WPT>(sy * SB_X + sx), kmin);
                }
                CYC_MARK(c_post);
                // ---- write the cells that changed back to L2 --------------------------------------------------------
                const int r = r0 + ly, c = c0 + lx;
                const bool mine = (bits >> lane) & 1u;
                unsigned pushed = 0;                                  // bit side: my cell went to that neighbour shard
                if (mine && r < H && c < W) {
                    const long long gk = (long long)r * W + c;
                    if (peers && (r < p.own_lo || r >= p.own_hi)) {
                        // a ghost row: the neighbour shard lowers it too (remote atomicMin), so my store must not raise it
                        atomicMin(p.g + gk, vq);
                    } else {
                        __stcg(p.g + gk, vq);
                        if (peers) {
                            // my first / last two owned rows ARE the neighbour's ghost rows: push the value over NVLink
                            if (p.peer_g[0] && r < p.own_lo + 2) { atomicMin(p.peer_g[0] + (long long)(r + p.peer_row_off[0]) * W + c, vq); pushed = 1; }
                            if (p.peer_g[1] && r >= p.own_hi - 2) { atomicMin(p.peer_g[1] + (long long)(r + p.peer_row_off[1]) * W + c, vq); pushed |= 2; }
                        }
                    }
                    // flat-key wrap at the map's column edges: (r, W-1) feeds (r..r+2, 0), (r, 0) feeds (r-2..r, W-1); posted at the visit's end
                    if (c == W - 1) atomicMin(sm.emin + 9, vq);
                    if (c == 0) atomicMin(sm.emin + 10, vq);
                }
                // ---- cut-through between tiles: a changed cell that faces another tile is announced NOW (value, fence, key),
                // so the neighbouring tile starts on the wavefront while this one is still relaxing
                if (peers) {
                    // cut-through ACROSS GPUs: rows pushed in this turn are announced to the neighbour shard now (system-scope fence,
                    // then remote key posts), like the local posts below; only tiles on a shard boundary ever get here.  Waiting for the
                    // visit's end made every exchange across the boundary cost a whole visit, and a source on the boundary (16384 rows
                    // over an even number of ranks) sends the main wave ALONG it.
                    const unsigned pm = __reduce_or_sync(0xffffffffu, pushed);
                    if (pm) {
                        __threadfence_system();
                        __syncwarp();
                        if (lane < 16) {
                            const int side = lane >> 3;
                            if ((pm >> side) & 1u) {
                                const int add = peer_post<WPT>(p, side, lane & 7, tx, kmin);
                                if (add) atomicAdd(p.gcount, add);
                            }
                        }
                    }
                }
                CYC_MARK(c_wb);
                uint32_t outer = 0;
                if (sby == 0) outer |= bits & 0x000000ffu;
                if (sby == SB_Y - 1) outer |= bits & 0xff000000u;
                if (sbx == 0) outer |= bits & 0x01010101u;
                if (sbx == SB_X - 1) outer |= bits & 0x80808080u;
                if (outer) {
                    if (!p.light) __threadfence();
                    __syncwarp();
                    if (lane < 9 && lane != 4) {
                        const int dy = lane / 3 - 1, dx = lane % 3 - 1;
                        uint32_t need = bits;
                        if (dy < 0) need &= 0x000000ffu; else if (dy > 0) need &= 0xff000000u;
                        if (dx < 0) need &= 0x01010101u; else if (dx > 0) need &= 0x80808080u;
                        const int sy = sby + dy, sx = sbx + dx;
                        const int tdy = sy < 0 ? -1 : (sy >= SB_Y ? 1 : 0), tdx = sx < 0 ? -1 : (sx >= SB_X ? 1 : 0);
                        const int ny = ty + tdy, nx = tx + tdx;
                        if (need && (tdy || tdx) && ny >= 0 && ny < p.tiles_y && nx >= 0 && nx < p.tiles_x) {
                            uint32_t* kp = p.tile_key + ny * p.tiles_x + nx;
                            const uint32_t old = p.light ? atomic_min_release(kp, kmin) : atomicMin(kp, kmin);
                            TRACE_POST(p, ny * p.tiles_x + nx);
                            if (old > p.cap && kmin <= p.cap) atomicAdd(p.gcount, 1);
                        }
                    }
                }
                CYC_MARK(c_fence);
            }
            nap = p.wait_ns;
            continue;
        }
        // ---- nothing due for this warp ------------------------------------------------------------------------------
        if (!counted_idle) {
            int old = 0;
            if (lane == 0) old = atomicAdd(sm.ctl, 1);
            old = __shfl_sync(0xffffffffu, old, 0);
            counted_idle = true;
            if (old == WPT - 1) {                                     // every warp is idle-counted: verify and finish
                uint32_t g4 = kNoKey;
#pragma unroll
                for (int t = 0; t < NSB / 32; ++t) g4 = min(g4, due[lane + 32 * t]);
                const bool any = __any_sync(0xffffffffu, g4 != kNoKey);
                if (!any && ctl[0] == WPT) {
                    // Before ending the visit: did a neighbour post to THIS tile meanwhile (within the band)?  Then stay on it:
                    // take the key, re-read the halo ring, wake the sub-blocks next to halo cells that dropped.  Saves the
                    // publish / claim / stage round trip of a separate visit, which sits on the wavefront's critical path.
                    // (Not for a tile that sticks out over the map's right edge: the staged positions beyond column W-1 mirror cells of
                    // the NEXT row's first columns - the flat-key wrap - and sit inside the square, not on the ring that is re-read
                    // here; round 2 found such a tile keeping a stale mirror on an unframed 300x130 map.  Those tiles end and are staged afresh.)
                    uint32_t pk = kNoKey;
                    if (p.merge && c0 + TS <= W) {
                        if (lane == 0) {
                            pk = __ldcg(p.tile_key + tile);
                            pk = (pk <= thr && pk <= p.cap) ? atomicExch(p.tile_key + tile, kNoKey) : kNoKey;
                            if (pk != kNoKey) {
                                __threadfence();
                                sm.ctl[3] += 1;                       // two counted units (pending + in flight) become one
                                atomicSub(sm.ctl, 1);                 // I am busy again: nobody else can end the visit while I refresh the halo
                            }
                        }
                        pk = __shfl_sync(0xffffffffu, pk, 0);
                    }
                    if (pk != kNoKey) {
                        counted_idle = false;
                        for (int hx = lane; hx < 4 * TS + 4; hx += 32) {
                            int rr, cc2;                                      // position on the ring of the (TS+2)^2 staged square
                            if (hx < TS + 2) { rr = 0; cc2 = hx; }
                            else if (hx < 2 * (TS + 2)) { rr = TS + 1; cc2 = hx - (TS + 2); }
                            else if (hx < 2 * (TS + 2) + TS) { rr = hx - 2 * (TS + 2) + 1; cc2 = 0; }
                            else { rr = hx - 2 * (TS + 2) - TS + 1; cc2 = TS + 1; }
                            const long long gk = (long long)(r0 - 1 + rr) * W + (c0 - 1 + cc2);
                            if (gk < 0 || gk >= N) continue;
                            const uint32_t nv = __ldcg(p.g + gk);
                            const int so2 = rr * SROW + (SCOL0 - 1) + cc2;
                            if (nv < s[so2]) {
                                s[so2] = nv;
                                // interior cells within one step of the halo cell: rows rr-2 .. rr, columns cc-2 .. cc (tile coordinates)
                                const int y0 = max(rr - 2, 0), y1 = min(rr, TS - 1), x0 = max(cc2 - 2, 0), x1 = min(cc2, TS - 1);
                                for (int yy = 0; yy < 2; ++yy)
                                    for (int xx = 0; xx < 2; ++xx) {
                                        const int t2 = ((yy ? y1 : y0) / SBH) * SB_X + (xx ? x1 : x0) / SBW;
                                        atomicMin(sm.due + slot_of<WPT>(t2), nv);
                                    }
                            }
                        }
                        if (peers && (r0 < p.own_lo || r0 + TS > p.own_hi)) {
                            // a shard's tile that holds ghost rows: the neighbour lowers those cells remotely INSIDE the square, so they are
                            // re-read as well (rows of the tile outside my owned range; at most a few)
                            for (int ly2 = 0; ly2 < TS; ++ly2) {
                                const int r2 = r0 + ly2;
                                if ((r2 >= p.own_lo && r2 < p.own_hi) || r2 >= H) continue;
                                for (int lx2 = lane; lx2 < TS; lx2 += 32) {
                                    if (c0 + lx2 >= W) continue;
                                    const uint32_t nv = __ldcg(p.g + (long long)r2 * W + c0 + lx2);
                                    const int so2 = (ly2 + 1) * SROW + SCOL0 + lx2;
                                    if (nv < s[so2]) {
                                        s[so2] = nv;
                                        // the cell's own sub-block and the ones within one step of it
                                        const int y0 = max(ly2 - 1, 0), y1 = min(ly2 + 1, TS - 1), x0 = max(lx2 - 1, 0), x1 = min(lx2 + 1, TS - 1);
                                        for (int yy = 0; yy < 2; ++yy)
                                            for (int xx = 0; xx < 2; ++xx)
                                                atomicMin(sm.due + slot_of<WPT>(((yy ? y1 : y0) / SBH) * SB_X + (xx ? x1 : x0) / SBW), nv);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        continue;                                     // poll again: my own sub-blocks may be due; idle + finisher check come after
                    }
                    if (lane == 0) ctl[1] = 1;
                }
            }
        }
        if (ctl[1]) break;
        __nanosleep(nap);
        nap = min(nap * 2u, p.idle_ns);
    }
    __syncthreads();
#if NSFS_FIELD_TRACE
    if (tid == 0) sm.tr[3] = gtime();
    CYC_MARK(c_scan);
    if (lane == 0) {
        atomicAdd(&sm.tr_it[0], tr_iters); atomicMax(&sm.tr_it[1], tr_iters);
        atomicAdd(&sm.cyc[0], (unsigned long long)c_scan); atomicAdd(&sm.cyc[1], (unsigned long long)c_setup); atomicAdd(&sm.cyc[2], (unsigned long long)c_pass);
        atomicAdd(&sm.cyc[3], (unsigned long long)c_post); atomicAdd(&sm.cyc[4], (unsigned long long)c_wb); atomicAdd(&sm.cyc[5], (unsigned long long)c_fence);
        atomicAdd(&sm.cyc[6], (unsigned long long)c_turns);
    }
#endif
    // ---- candidates the band rejected: the smallest one is when this tile wants to be relaxed again --------------------
    rej = __reduce_min_sync(0xffffffffu, rej);
    if (lane == 0 && rej != kNoKey) atomicMin(sm.emin + 11, rej);
    __syncthreads();
    // Every changed cell was stored during the visit; the keys posted from here on announce values whose stores a block
    // barrier has ordered before this point, and a fence is cumulative over what happened before it.
    if (tid < 32 && !p.light) __threadfence();       // (rows pushed to a neighbour shard were fenced and announced in the turn that pushed them)
    int newly = 0;        // tiles this thread made pending: summed over warp 0 and folded into ONE update of the pending counter
    {
        // key posts: 9..11 right map edge -> left-edge tiles of tile rows ty-1..ty+1, 12..14 left map edge -> right-edge tiles
        // of rows ty-1..ty+1, 15 this tile's own rejected candidates.  (The 8 direct neighbours were posted during the visit.)
        int target = -1;
        uint32_t m = kNoKey;
        if (tid >= 9 && tid < 12) {
            const int ny = ty + (tid - 10);
            m = sm.emin[9];
            if (ny >= 0 && ny < p.tiles_y) target = ny * p.tiles_x;
        } else if (tid >= 12 && tid < 15) {
            const int ny = ty + (tid - 13);
            m = sm.emin[10];
            if (ny >= 0 && ny < p.tiles_y) target = ny * p.tiles_x + p.tiles_x - 1;
        } else if (tid == 15) {
            m = sm.emin[11];
            target = tile;
        }
        if (target >= 0 && m != kNoKey) {
            const uint32_t old = p.light ? atomic_min_release(p.tile_key + target, m) : atomicMin(p.tile_key + target, m);
            TRACE_POST(p, target);
            if (old > p.cap && m <= p.cap) newly = 1;                 // a tile became pending (within the cap)
        }
    }
    if (tid < 32) {
        // the visit's net effect on "pending tiles + visits in flight": +newly pending, -1 for this visit.  One atomic, so the
        // counter never dips below what is really outstanding and no second fence is needed before it.
        const int net = __reduce_add_sync(0xffffffffu, newly) - 1 - sm.ctl[3];
        if (tid == 0) sm.ctl[2] = net;
    }
    if (tid == 0) atomicAdd(p.counts + 3, 1);
#if NSFS_FIELD_TRACE
    __syncthreads();
    if (tid == 0 && p.prof) {
        const unsigned long long rec = atomicAdd(p.prof, 1ull);
        if (rec < (unsigned long long)kTraceMaxVisits) {
            unsigned long long* o = p.prof + ((16 + p.tiles_x * p.tiles_y + 7) & ~7) + rec * 8;
            o[0] = (unsigned long long)(unsigned)tile | ((unsigned long long)blockIdx.x << 32);
            o[1] = sm.tr[1]; o[2] = sm.tr[2]; o[3] = sm.tr[3]; o[4] = gtime();
            o[5] = (unsigned long long)key0 | ((unsigned long long)thr << 32);
            o[6] = (unsigned long long)sm.tr_it[0] | ((unsigned long long)(sm.tr_it[1] & 0xffffu) << 32) | ((unsigned long long)(unsigned)sm.ctl[3] << 48);
            o[7] = sm.tr[6] == ~0ull ? 0ull : ((sm.tr[6] - sm.tr[1]) | ((sm.tr[7] - sm.tr[1]) << 32));   // first / last sub-block change, ns after the claim
        }
        for (int c = 0; c < 7; ++c) atomicAdd(p.prof + 1 + c, sm.cyc[c]);          // totals over all visits
    }
#endif
    __syncthreads();   // smem is reused by the next tile of this block
}

// No rounds and no grid-wide barrier.  Tile t is owned by block t % gridDim.x (round-robin, so the tiles along the
// wavefront belong to different SMs).  Each block repeatedly takes its pending tile with the smallest key, relaxes it
// within the band [key, key + delta], and posts keys to the tiles whose halo it moved.  *gcount = tiles with a pending
// key + visits in flight; the kernel ends when it reaches zero.  Every spin loop has a bail-out (counts[4] bit 1) so the
// kernel cannot hang.  Cooperative launch: blocks spin on each other's posts, so all of them must be resident.
template <int WPT>
__global__ void __launch_bounds__(WPT * 32, WPT == 32 ? 1 : 32 / WPT) field_solve_async_kernel(FieldParams p) {
    constexpr int THREADS = WPT * 32;
    __shared__ TileSmem<WPT> sm;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_tiles = p.tiles_x * p.tiles_y;
    // Global gate (p.gate != 0): with many tiles resident, a block must not run far AHEAD of the wavefront (it would relax its
    // tile against values that are still going to drop, and every such visit is repeated later).  Each worker publishes the
    // smallest key it holds (pending or in flight), the last block of the grid does nothing but reduce those to one "floor",
    // and a worker leaves a tile alone while its key exceeds floor + gate.  The floor is a hint: stale values cost time, never bits.
    const bool gated = p.gate != 0u && gridDim.x > 1;
    const int workers = gated ? (int)gridDim.x - 1 : (int)gridDim.x;
    if (gated && (int)blockIdx.x == workers) {
        for (long long iter = 0; iter < (1ll << 40); ++iter) {
            if (tid == 0) sm.key = kNoKey;
            __syncthreads();
            uint32_t m = kNoKey;
            for (int b = tid; b < workers; b += THREADS) m = min(m, __ldcg(p.blk_key + b));
            m = __reduce_min_sync(0xffffffffu, m);
            if (lane == 0) atomicMin(&sm.key, m);
            __syncthreads();
            if (tid == 0) {
                __stcg(p.gfloor, sm.key);
                const int pend = atomicAdd(p.gcount, 0);
                sm.emin[0] = (uint32_t)pend;
                // the workers read the pending count HERE (0 = not published yet), not from the counter itself: with shards the
                // counter lives on another GPU, and hundreds of idle blocks polling it over NVLink would crowd out the real posts
                __stcg(p.gfloor + 1, (uint32_t)(pend > 0 ? pend : 0) + 1u);
            }
            __syncthreads();
            const int pending = (int)sm.emin[0];
            __syncthreads();
            if (pending <= 0 || (atomicAdd(p.counts + 4, 0) & 2)) break;
            __nanosleep(200);
        }
        return;
    }
    // Picking the next tile is WARP 0's job; the other warps of the block wait at the barrier, where they issue nothing.  (Round 2,
    // ncu on C2: with every warp of every idle or gated block running this scan and its shuffles every 400 ns, the scan was 25 % of
    // all instructions executed; with several blocks resident per SM it competed with the visits next door for issue slots.)
    const int warp_id = tid >> 5;
    int idle_spins = 0;
    for (long long iter = 0; iter < (1ll << 40); ++iter) {
        if (warp_id == 0) {
            unsigned long long best = ~0ull;
            for (int t = blockIdx.x + lane * workers; t < n_tiles; t += workers * 32) {
                const uint32_t k = __ldcg(p.tile_key + t);
                if (k <= p.cap) best = min(best, ((unsigned long long)k << 32) | (unsigned)t);
            }
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            int verdict = 0;                       // 0 take the tile, 1 nothing pending anywhere: leave, 2 wait and look again, 3 bail out
            if (best == ~0ull) {
                int pending = 1;
                if (lane == 0) {
                    const uint32_t pub = gated ? __ldcg(p.gfloor + 1) : 0u;
                    pending = pub ? (int)pub - 1 : atomicAdd(p.gcount, 0);
                }
                pending = __shfl_sync(0xffffffffu, pending, 0);
                verdict = pending <= 0 ? 1 : 2;
            } else if (gated) {
                uint32_t fl = 0;
                if (lane == 0) fl = __ldcg(p.gfloor);
                fl = __shfl_sync(0xffffffffu, fl, 0);
                const uint32_t k = (uint32_t)(best >> 32);
                if (fl != kNoKey && k > fl && k - fl > p.gate) verdict = 2;      // ahead of the wavefront: wait
            }
            if (gated && lane == 0) __stcg(p.blk_key + blockIdx.x, best == ~0ull ? kNoKey : (uint32_t)(best >> 32));
            if (verdict == 2) {
                if (++idle_spins > (1 << 22)) verdict = 3;
                else __nanosleep(p.sched_ns);
            } else idle_spins = 0;
            if (lane == 0) { sm.best = best; sm.key = (uint32_t)verdict; }
        }
        __syncthreads();
        const unsigned long long pick = sm.best;
        const int verdict = (int)sm.key;
        __syncthreads();
        if (verdict == 1) break;
        if (verdict == 3) { if (tid == 0) atomicOr(p.counts + 4, 2); break; }
        if (verdict == 2) continue;
        const int tile = (int)(pick & 0xffffffffu);
        if (tid == 0) {
            sm.key = atomicExch(p.tile_key + tile, kNoKey);
            __threadfence();
        }
        __syncthreads();
        const uint32_t my_key = sm.key;
        __syncthreads();
        if (my_key == kNoKey) continue;
        const unsigned long long lim = (unsigned long long)my_key + p.delta;
        const uint32_t thr = lim < p.cap ? (uint32_t)lim : p.cap;
        relax_tile<WPT>(p, tile, thr, sm, my_key);
        if (tid == 0 && sm.ctl[2] != 0) atomicAdd(p.gcount, sm.ctl[2]);
    }
}

// Before a (capped) solve: counts[1] = tiles whose pending key lies within the cap.  After it: counts[6] = the smallest
// key still pending (Q17.15, kNoKey = none), which the multi-GPU driver all-reduces to place the next band.
__global__ void __launch_bounds__(256) field_pending_kernel(const uint32_t* __restrict__ tile_key, int n_tiles, uint32_t cap,
                                                            int* __restrict__ counts, int what) {
    __shared__ uint32_t s_min;
    __shared__ int s_cnt;
    if (threadIdx.x == 0) { s_min = kNoKey; s_cnt = 0; }
    __syncthreads();
    uint32_t mn = kNoKey;
    int cnt = 0;
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        const uint32_t k = tile_key[t];
        mn = min(mn, k);
        cnt += k <= cap;
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_min, mn); atomicAdd(&s_cnt, cnt); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (what == 0) counts[1] = s_cnt;
        else counts[6] = (int)s_min;
    }
}

// ---- predecessor field + reached count + "would throw" check ------------------------------------------------
// pred[v] = lowest direction d whose in-edge attains g[v] exactly (integer sums: equality is exact).  The reference keeps
// whichever tied parent its heap order met first; any of them is a valid optimal predecessor (north_star: path validity
// exact, cost within 1e-5).
__global__ void __launch_bounds__(256) field_pred_kernel(const uint32_t* __restrict__ g, const uint16_t* __restrict__ in_mask,
                                                         const uint16_t* __restrict__ out_mask, int W, int N, int ks,
                                                         int count_lo, int count_hi,
                                                         uint8_t* __restrict__ pred, int* __restrict__ counts) {
    // grid-stride over the cells (N < 2^29: 32-bit keys); the reached count and the throw flag leave the block through ONE
    // atomic each (one atomicAdd per warp on a single address made this kernel 6 ms on the 16384^2 map: the L2 serialises them)
    __shared__ int s_reached, s_thrown;
    if (threadIdx.x == 0) { s_reached = 0; s_thrown = 0; }
    __syncthreads();
    int reached = 0, thrown = 0;
    int off[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) off[d] = nb_di(d) * W + nb_dj(d);
    for (int v = blockIdx.x * 256 + threadIdx.x; v < N; v += gridDim.x * 256) {
        const uint32_t gv = g[v];
        uint32_t pd = 255u;
        if (gv < kInfQ) {
            const bool counted = v >= count_lo && v < count_hi;     // a row-partitioned shard counts its own rows only
            reached += counted;
            if (counted && (__ldg(out_mask + v) & kThrowBit)) thrown = 1;   // a reached cell is expanded; expanding this one throws
            if (v != ks) {
                const uint32_t m = __ldg(in_mask + v);
#pragma unroll
                for (int d = 7; d >= 0; --d) {
                    if (!((m >> d) & 1u)) continue;
                    const uint32_t cand = g[v - off[d]] + (((m >> (8 + d)) & 1u) ? kQDiag : kQOne);
                    if (cand == gv) pd = (uint32_t)d;
                }
            }
        }
        pred[v] = (uint8_t)pd;
    }
    reached = __reduce_add_sync(0xffffffffu, reached);
    thrown = __reduce_or_sync(0xffffffffu, (unsigned)thrown);
    if ((threadIdx.x & 31) == 0) {
        if (reached) atomicAdd(&s_reached, reached);
        if (thrown) atomicOr(&s_thrown, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_reached) atomicAdd(counts + 5, s_reached);
        if (s_thrown) atomicOr(counts + 4, 1);
    }
}

// The plane leaves the solver as fp32 cost units (what nsfs_field* hands out): in place, 128 bits per thread.
__global__ void __launch_bounds__(256) field_to_float_kernel(uint4* __restrict__ g4, long long n4, uint32_t* __restrict__ g, long long N) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr float kScale = 1.0f / (float)kQOne;
    for (long long k = t; k < n4; k += stride) {
        const uint4 q = g4[k];
        g4[k] = make_uint4(__float_as_uint((float)q.x * kScale), __float_as_uint((float)q.y * kScale),
                           __float_as_uint((float)q.z * kScale), __float_as_uint((float)q.w * kScale));
    }
    for (long long k = n4 * 4 + t; k < N; k += stride) g[k] = __float_as_uint((float)g[k] * kScale);
}

// ---- predecessor walk (goal -> source) -----------------------------------------------------------------------
// A dependent chain of 1-byte loads: one thread chases it (stores are fire-and-forget).  Runs on the finished (fp32) plane.
__global__ void field_path_kernel(const float* __restrict__ g, const uint8_t* __restrict__ pred,
                                  const uint16_t* __restrict__ in_mask, int W, long long N, long long ks, long long kg,
                                  int32_t* __restrict__ path, int cap, int* __restrict__ n_out, double* __restrict__ cost_out,
                                  int* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const bool reachable = kg >= 0 && kg < N && g[kg] < kInfF;
    if (!reachable) {                                     // "no path" is the single start cell (astar.cpp:129-135)
        if (cap > 0) { path[0] = (int)(ks / W); path[1] = (int)(ks % W); }
        *n_out = 1; *cost_out = kInfD; *status = cap > 0 ? NSFS_OK : NSFS_E_TRUNC;
        return;
    }
    long long k = kg;
    int n = 0, n_card = 0, n_diag = 0, st = NSFS_OK;
    for (;;) {
        if (n < cap) { path[2 * n] = (int)(k / W); path[2 * n + 1] = (int)(k % W); }
        n++;
        if (k == ks) break;
        const uint32_t d = pred[k];
        if (d > 7u || n > N) { st = NSFS_E_ARG; break; }  // broken chain: cannot happen on a converged field
        if ((__ldg(in_mask + k) >> (8 + d)) & 1u) n_diag++; else n_card++;
        k -= nb_off((int)d, W);
    }
    *n_out = n;
    *cost_out = (double)n_card + (double)n_diag * kSqrt2D;
    *status = st != NSFS_OK ? st : (n > cap ? NSFS_E_TRUNC : NSFS_OK);
}

// ---- host side --------------------------------------------------------------------------------------------
int field_alloc(nsfs_planner* h) {
    FieldWork& f = h->field;
    if (f.g) return NSFS_OK;
    f.tiles_x = (h->W + TS - 1) / TS;
    f.tiles_y = (h->H + TS - 1) / TS;
    const int n_tiles = f.tiles_x * f.tiles_y;
    NSFS_CUDA_TRY(h, cudaMalloc(&f.g, (size_t)(h->N + 4) * sizeof(uint32_t)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.pred, (size_t)h->N));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.tile_key, (size_t)(n_tiles + 32) * sizeof(uint32_t)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.counts, 16 * sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.blk_key, (size_t)(h->sm_count * 32 + 64) * sizeof(uint32_t)));     // [grid] + the floor word
    return NSFS_OK;
}

void field_free(nsfs_planner* h) {
    FieldWork& f = h->field;
    cudaFree(f.g); cudaFree(f.pred); cudaFree(f.tile_key); cudaFree(f.counts); cudaFree(f.blk_key);
    f = FieldWork();
}

static uint32_t to_q(double cost) {                // cost units -> Q17.15, saturating below kNoKey
    if (!(cost > 0.0)) return 0u;
    const double q = cost * (double)kQOne;
    return q >= (double)(kNoKey - 1u) ? kNoKey - 1u : (uint32_t)q;
}

static FieldParams field_params(nsfs_planner* h) {
    FieldWork& f = h->field;
    FieldParams p;
    p.g = f.g; p.in_mask = h->d_in_mask; p.tile_key = f.tile_key;
    p.light = h->opt_field_fence == 1 ? 1 : 0;                // key posts as release atomics instead of __threadfence + atomicMin
    p.merge = h->opt_field_merge == 2 ? 0 : 1;
    p.sched_ns = h->opt_field_sched_ns > 0 ? (unsigned)h->opt_field_sched_ns : 400u;
    p.gcount = f.counts + 1;
    for (int sd = 0; sd < 2; ++sd) { p.peer_g[sd] = nullptr; p.peer_key[sd] = nullptr; p.peer_row_off[sd] = 0; p.peer_tiles_y[sd] = 0; p.peer_rows[sd] = 0; }
    p.own_lo = 0; p.own_hi = h->H;
    p.counts = f.counts; p.H = h->H; p.W = h->W; p.N = h->N; p.tiles_x = f.tiles_x; p.tiles_y = f.tiles_y;
    p.prof = nullptr;
    p.inner = h->opt_field_inner > 0 ? (int)h->opt_field_inner : 16;
    p.idle_ns = h->opt_field_idle_ns > 0 ? (unsigned)h->opt_field_idle_ns : 1000u;
    p.delta = to_q(h->opt_field_delta > 0 ? (double)h->opt_field_delta : 128.0);
    p.wait_ns = h->opt_field_wait_ns > 0 ? (unsigned)h->opt_field_wait_ns : 200u;
    p.cap = kNoKey - 1u;
    p.blk_key = f.blk_key; p.gfloor = f.blk_key + h->sm_count * 32 + 32;
    p.gate = h->opt_field_gate > 0 ? to_q((double)h->opt_field_gate) : 0u;     // 0 = default for the residency chosen, < 0 = off
    return p;
}

// Phase 1: g = 1e5 everywhere, no pending tile; (si, sj) is the source or si < 0 for "no source on this shard".
int launch_field_begin(nsfs_planner* h, int si, int sj) {
    int rc = field_alloc(h);
    if (rc != NSFS_OK) return rc;
    FieldWork& f = h->field;
    f.valid = false;
    FieldParams p = field_params(h);
    const long long n4 = h->N / 4;
    long long fill_blocks = (n4 + 255) / 256;
    if (fill_blocks < 1) fill_blocks = 1;
    if (fill_blocks > (long long)h->sm_count * 8) fill_blocks = (long long)h->sm_count * 8;
    field_fill_kernel<<<(int)fill_blocks, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(f.g), n4, f.g, h->N, f.tile_key, f.tiles_x * f.tiles_y);
    field_seed_kernel<<<1, 32, 0, h->stream>>>(p, h->d_free, h->d_out_mask, si, sj);
    h->launches += 2;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    f.src_i = si; f.src_j = sj;
    return NSFS_OK;
}

// Phase 2 (optional, repeatable): lower whole rows of g from device values; d_improved accumulates the cells lowered.
int launch_field_inject(nsfs_planner* h, int row0, int nrows, const float* d_vals, int* d_improved) {
    FieldParams p = field_params(h);
    const long long n = (long long)nrows * h->W;
    if (n <= 0) return NSFS_OK;
    field_inject_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(p, (long long)row0 * h->W, n, reinterpret_cast<const uint32_t*>(d_vals), d_improved);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

// The solve kernel for a tile-residency setting: field_mode 1 = 4 warps per tile (8 tiles resident per SM), 2 = 8 warps (4 tiles),
// 3 = 16 warps (2 tiles), 4 = 32 warps (1 tile); 0 = chosen from the map size (launch_solve_any).  Same bits whichever runs.
template <int WPT>
static int launch_solve(nsfs_planner* h, FieldParams& p) {
    FieldWork& f = h->field;
    auto kern = field_solve_async_kernel<WPT>;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    NSFS_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WPT * 32, 0));
    if (per_sm < 1) return set_cuda_error(h, cudaErrorLaunchOutOfResources, "field_solve_async_kernel occupancy");
    if (h->opt_field_tiles_per_sm > 0 && per_sm > h->opt_field_tiles_per_sm) per_sm = (int)h->opt_field_tiles_per_sm;
    int grid = per_sm * h->sm_count;
    const int n_tiles = f.tiles_x * f.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    void* args[] = {&p};
    if (p.gate) NSFS_CUDA_TRY(h, cudaMemsetAsync(f.blk_key, 0, (size_t)(h->sm_count * 32 + 64) * sizeof(uint32_t), h->stream));   // floor 0 until the workers report
#if NSFS_FIELD_TRACE
    static unsigned long long* d_trace = nullptr;
    const size_t trace_words = (size_t)((16 + n_tiles + 7) & ~7) + (size_t)kTraceMaxVisits * 8;
    if (getenv("NSFS_FIELD_TRACE_OUT")) {
        if (!d_trace) cudaMalloc(&d_trace, trace_words * 8);
        cudaMemsetAsync(d_trace, 0xff, trace_words * 8, h->stream);
        cudaMemsetAsync(d_trace, 0, 16 * 8, h->stream);
        p.prof = d_trace;
    }
#endif
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    NSFS_CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(WPT * 32), args, 0, h->stream));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    h->launches++;
#if NSFS_FIELD_TRACE
    if (p.prof) {
        std::vector<unsigned long long> hp(trace_words);
        cudaMemcpyAsync(hp.data(), p.prof, trace_words * 8, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        if (FILE* fo = fopen(getenv("NSFS_FIELD_TRACE_OUT"), "wb")) {
            const unsigned long long hdr[4] = {(unsigned long long)f.tiles_x, (unsigned long long)f.tiles_y, hp[0], (unsigned long long)grid};
            fwrite(hdr, 8, 4, fo);
            const size_t nrec = hp[0] < (unsigned long long)kTraceMaxVisits ? (size_t)hp[0] : (size_t)kTraceMaxVisits;
            fwrite(hp.data(), 8, (size_t)((16 + n_tiles + 7) & ~7) + nrec * 8, fo);
            fclose(fo);
        }
    }
#endif
    return NSFS_OK;
}

// Which residency?  Measured on one B200 (profiles/r02_field_solve.md): a map whose wavefront is never wider than the machine
// (2048^2: <= ~130 tiles on the front for 148 SMs) is bound by how fast ONE visit crosses its tile, so it gets all 32 warps of
// an SM per tile; a large map (16384^2: thousands of tiles on the front) is bound by visits per second, so it gets 4 resident
// tiles of 8 warps per SM plus the global gate that keeps the extra concurrency from running ahead of the wave.
static int launch_solve_any(nsfs_planner* h, FieldParams& p, bool allow_gate = true) {
    long long mode = h->opt_field_mode;
    if (mode <= 0) {
        const long long n_tiles = (long long)p.tiles_x * p.tiles_y;
        mode = n_tiles > 16ll * h->sm_count ? 2 : 4;
        if (mode == 2 && h->opt_field_gate == 0) p.gate = to_q(192.0);
        if (mode == 4 && h->opt_field_gate == 0) p.gate = 0u;
    }
    if (h->opt_field_gate < 0 || !allow_gate) p.gate = 0u;
    if (mode == 1) return launch_solve<4>(h, p);
    if (mode == 2) return launch_solve<8>(h, p);
    if (mode == 3) return launch_solve<16>(h, p);
    return launch_solve<32>(h, p);
}

// Phase 3 (repeatable): relax every pending tile whose key is <= cap until none is left; nothing is relaxed beyond cap
// (cap = +inf: to the fixed point).  Leaves the smallest key still pending in counts[6].
int launch_field_relax(nsfs_planner* h, float cap) {
    FieldWork& f = h->field;
    FieldParams p = field_params(h);
    if (cap < 3.0e38f) p.cap = to_q((double)cap);
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, p.cap, f.counts, 0);
    h->launches++;
    int rc = launch_solve_any(h, p);
    if (rc != NSFS_OK) return rc;
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, p.cap, f.counts, 1);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

// Peer-to-peer solve: the shared pending counter was set by the host (sum of every shard's pending tiles), every shard's
// kernel runs until the GLOBAL count is zero.  No cap, no rounds.
int launch_field_relax_p2p(nsfs_planner* h) {
    FieldWork& f = h->field;
    FieldParams p = field_params(h);
    for (int sd = 0; sd < 2; ++sd) {
        p.peer_g[sd] = reinterpret_cast<uint32_t*>(f.peer_g[sd]); p.peer_key[sd] = f.peer_key[sd];
        p.peer_row_off[sd] = f.peer_row_off[sd]; p.peer_tiles_y[sd] = f.peer_tiles_y[sd]; p.peer_rows[sd] = f.peer_rows[sd];
    }
    if (f.own_hi > f.own_lo) { p.own_lo = f.own_lo; p.own_hi = f.own_hi; }
    p.gcount = (f.peer_counts ? f.peer_counts : f.counts) + 1;
    p.light = 0;            // peer posts need system scope: keep the fences
    if (p.peer_g[0] || p.peer_g[1]) {
        field_push_boundary_kernel<<<(4 * h->W + 255) / 256, 256, 0, h->stream>>>(p);
        h->launches++;
    }
    // The gate's floor is this GPU's own smallest key: the block that holds it is always eligible, posts from the neighbour
    // shards lower keys (and with them the floor) like local ones, so gating each shard on its own floor cannot stall the group.
    int rc = launch_solve_any(h, p, true);
    if (rc != NSFS_OK) return rc;
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, p.cap, f.counts, 1);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

int launch_field_count_pending(nsfs_planner* h) {
    FieldWork& f = h->field;
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, kNoKey - 1u, f.counts, 0);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

int launch_field_min_pending(nsfs_planner* h) {
    FieldWork& f = h->field;
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, 0u, f.counts, 1);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

// Phase 4: predecessor field, reached count and "would throw" flag over the rows [count_row_lo, count_row_hi); then the
// plane becomes fp32 cost units in place (the solve is over: begin again before relaxing further).
int launch_field_finish(nsfs_planner* h) {
    FieldWork& f = h->field;
    const long long ks = f.src_i >= 0 ? (long long)f.src_i * h->W + f.src_j : -1;
    const long long lo = (long long)(h->opt_count_row_lo > 0 ? h->opt_count_row_lo : 0) * h->W;
    const long long hi = h->opt_count_row_hi > 0 ? (long long)h->opt_count_row_hi * h->W : h->N;
    long long blocks = (h->N + 255) / 256;
    if (blocks > (long long)h->sm_count * 16) blocks = (long long)h->sm_count * 16;      // 8 resident blocks per SM, two waves
    field_pred_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(f.g, h->d_in_mask, h->d_out_mask, h->W, (int)h->N, (int)ks, (int)lo, (int)hi,
                                                               f.pred, f.counts);
    const long long n4 = h->N / 4;
    long long cblocks = (n4 + 255) / 256;
    if (cblocks < 1) cblocks = 1;
    if (cblocks > (long long)h->sm_count * 16) cblocks = (long long)h->sm_count * 16;
    field_to_float_kernel<<<(unsigned)cblocks, 256, 0, h->stream>>>(reinterpret_cast<uint4*>(f.g), n4, f.g, h->N);
    h->launches += 2;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

int launch_field(nsfs_planner* h, int si, int sj, nsfs_stats* st) {
    (void)st;
    int rc = launch_field_begin(h, si, sj);
    if (rc == NSFS_OK) rc = launch_field_relax(h, 3.4e38f);
    if (rc == NSFS_OK) rc = launch_field_finish(h);
    return rc;
}

int launch_field_path(nsfs_planner* h, int gi, int gj, int32_t* d_path, int cap, int* d_n, double* d_cost, int* d_status) {
    FieldWork& f = h->field;
    const long long ks = (long long)f.src_i * h->W + f.src_j;
    const long long kg = (gi >= 0 && gi < h->H && gj >= 0 && gj < h->W) ? (long long)gi * h->W + gj : -1;
    field_path_kernel<<<1, 1, 0, h->stream>>>(reinterpret_cast<const float*>(f.g), f.pred, h->d_in_mask, h->W, h->N, ks, kg, d_path, cap, d_n, d_cost, d_status);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
