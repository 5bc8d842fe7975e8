// Djikstra cost-to-go field by tile-local relaxation, predecessor field, predecessor walk.
//
// What is computed (reference djikstra.cpp:27-145 run with an unreachable goal): the unique
// fixed point of   g[v] = min over in-edges (u -> v) of g[u] + w(u, v),   g[source] = 0,
// on the reference's quirk graph: flat-key stencil, row-only bounds test, direction cost 1 / sqrt(2) decided
// by the parity of passable directions before it (all folded into in_mask by map_pack.cu).  Because every
// edge costs >= 1 and the reference's open list is keyed on (int)g, its label-setting order cannot change
// the result, so a bulk-synchronous label-correcting solve reproduces it (fp32 here; tests bound the
// difference to the fp64 reference at 1e-5 relative, reachability exact).
//
// Structure: the grid is cut into 64x64 tiles.  A persistent cooperative kernel runs rounds; in a round each
// active tile is staged in shared memory with a 1-cell halo (flat-key addressed, so the reference's column
// wrap-around comes for free), relaxed to its local fixed point, written back, and the tiles whose halo it
// changed are flagged for the next round.  The next round's tile list is compacted from the flags with
// __ballot_sync/__popc.  Inside a tile every warp owns four 8x4 sub-blocks, one cell per lane; a sub-block is
// only re-relaxed while it or a neighbouring sub-block changed in the previous sweep (activity bit masks in
// shared memory), so the work follows the wavefront instead of sweeping the whole tile.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstring>

#include "nsfs_internal.cuh"

namespace cg = cooperative_groups;

namespace nsfs {

#ifndef NSFS_FIELD_PROF
#define NSFS_FIELD_PROF 0
#endif
#ifndef NSFS_FIELD_ORDERED
#define NSFS_FIELD_ORDERED 1      // 0 = the first (flag-based, unordered) in-tile scheduler, kept for A/B measurements
#endif

#ifndef NSFS_TS
#define NSFS_TS 64
#endif
constexpr int TS = NSFS_TS;            // tile side in cells (64: one 1024-thread block per SM; 32: four 256-thread blocks per SM)
constexpr int SBW = 8, SBH = 4;        // sub-block: 8 columns x 4 rows = one lane per cell
constexpr int SB_X = TS / SBW;
constexpr int SB_Y = TS / SBH;
constexpr int NSB = SB_X * SB_Y;       // sub-blocks per tile (128 / 32)
#ifndef NSFS_SB_PER_WARP
#define NSFS_SB_PER_WARP 4
#endif
constexpr int SB_PER_WARP = NSFS_SB_PER_WARP;   // warp w owns sub-blocks w, w + Wn, w + 2 Wn, ... (Wn = FIELD_WARPS); 8 => two 512-thread blocks per SM
constexpr int FIELD_WARPS = NSB / SB_PER_WARP;
constexpr int FIELD_THREADS = FIELD_WARPS * 32;
constexpr int FIELD_BLOCKS_PER_SM = 1024 / FIELD_THREADS;
static_assert(FIELD_WARPS <= 32 && NSB % SB_PER_WARP == 0 && (SB_PER_WARP == 4 || NSFS_FIELD_ORDERED), "the flag-based scheduler polls four flags as one word");
constexpr int SROW = TS + 8;           // smem row stride in floats; SROW % 32 == 8 => the 4 rows of a sub-block use disjoint banks
constexpr int SCOL0 = 4;               // smem column of tile column 0 (halo column -1 at 3)
constexpr int SROWS = TS + 2;


// Which sub-blocks a warp owns.  The wavefront is a line through the tile, and only the sub-blocks on it have work: the more
// warps those sub-blocks belong to, the more of the front advances at once.
//   0: warp w owns sub-blocks w, w + 32, w + 64, w + 96 (a COLUMN of 16 sub-blocks falls on 4 warps; two sub-blocks of one warp
//      are 16 cells apart)
//   1: a 2 x 2 group of adjacent sub-blocks (fewest hand-overs between warps: measured 2.2x SLOWER - front parallelism is what counts)
//   2: warp = (2 sy + b sx) mod 32 on the 16 x 8 grid of sub-blocks (default, b = 5): every row, column and 45-degree line of
//      sub-blocks falls on distinct warps, two sub-blocks of one warp are >= 25 cells apart.  C2 solve 1.51 -> 1.39 ms, C5 48.4 -> 45.6 ms
//      (b = 7, the lattice with the largest minimum distance, 32 cells: the same within noise).
#ifndef NSFS_SB_BLOCKED
#define NSFS_SB_BLOCKED ((NSFS_TS == 64 && NSFS_SB_PER_WARP == 4) ? 2 : 0)
#endif
#ifndef NSFS_SB_B
#define NSFS_SB_B 5
#endif
__device__ __forceinline__ int sb_of(int warp, int q) {
#if NSFS_SB_BLOCKED == 1
    return (((warp >> 2) << 1) + (q >> 1)) * SB_X + ((warp & 3) << 1) + (q & 1);
#elif NSFS_SB_BLOCKED == 2
    // warp = (2 sy + 5 sx) mod 32 on the 16 x 8 grid of sub-blocks: any row (8), column (16) or 45-degree line of sub-blocks falls
    // on distinct warps (w, w + 32 k of the default puts the 16 sub-blocks of a column on 4 warps)
    const int sx = 2 * q + (warp & 1);
    return (((warp - NSFS_SB_B * sx) & 31) >> 1) * SB_X + sx;
#else
    return warp + q * FIELD_WARPS;
#endif
}
__device__ __forceinline__ int slot_of(int t) {              // index of sub-block t in s_due / s_flag (warp-major)
#if NSFS_SB_BLOCKED == 1
    const int by = t / SB_X, bx = t % SB_X;
    return (((by >> 1) << 2) + (bx >> 1)) * SB_PER_WARP + ((by & 1) << 1) + (bx & 1);
#elif NSFS_SB_BLOCKED == 2
    const int sy = t / SB_X, sx = t % SB_X;
    return ((2 * sy + NSFS_SB_B * sx) & 31) * SB_PER_WARP + (sx >> 1);
#else
    return (t % FIELD_WARPS) * SB_PER_WARP + t / FIELD_WARPS;
#endif
}

__global__ void field_fill_kernel(float4* __restrict__ g4, long long n4, float* __restrict__ g, float4* __restrict__ gx4, float* __restrict__ gx,
                                  long long N, uint32_t* __restrict__ tile_key, int* __restrict__ tile_busy, int n_tiles) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4 v = make_float4(kInfF, kInfF, kInfF, kInfF);
    if (gx4) {                                                                // gx: "never expanded" (field_bucket.cu only)
        for (long long k = t; k < n4; k += stride) { g4[k] = v; gx4[k] = v; }
        for (long long k = n4 * 4 + t; k < N; k += stride) { g[k] = kInfF; gx[k] = kInfF; }
    } else {
        for (long long k = t; k < n4; k += stride) g4[k] = v;
        for (long long k = n4 * 4 + t; k < N; k += stride) g[k] = kInfF;
    }
    for (long long k = t; k < n_tiles; k += stride) { tile_key[k] = kNoKey; tile_busy[k] = 0; }
}

// One thread: g[source] = 0 and the first pending tiles; a non-free source is expanded here because in_mask only
// carries edges out of free cells (plan() pushes the start cell unconditionally, astar.cpp:45-52).
__global__ void field_seed_kernel(FieldParams p, const uint32_t* __restrict__ free_bits,
                                  const uint16_t* __restrict__ out_mask, int si, int sj) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (si < 0) {                               // no source on this shard: everything arrives through field_inject_kernel
        for (int c = 0; c < 16; ++c) p.counts[c] = 0;
        p.counts[7] = __float_as_int(-1e30f);
        return;
    }
    const long long ks = (long long)si * p.W + sj;
    p.g[ks] = 0.0f;
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
            const float w = ((om >> (8 + d)) & 1u) && d ? kSqrt2F : 1.0f;
            if (p.g[kv] > w) p.g[kv] = w;
            seed_tile(tile_of(kv));
        }
    }
    for (int c = 0; c < 16; ++c) p.counts[c] = 0;
    p.counts[1] = pend;                        // asynchronous scheduler: pending tiles + visits in flight
    p.counts[7] = __float_as_int(-1e30f);      // previous round's threshold
}

// Row injection for the row-partitioned multi-GPU field: g[row0*W + t] = min(g, vals[t]) over whole rows.
// An improved cell makes its tile pending with that value as key, exactly like a halo post from a neighbouring tile.
__global__ void __launch_bounds__(256) field_inject_kernel(FieldParams p, long long k0, long long n, const float* __restrict__ vals,
                                                           int* __restrict__ improved) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool better = false;
    if (t < n) {
        const long long k = k0 + t;
        const float v = vals[t], old = p.g[k];
        if (v < old) {
            p.g[k] = v;
            better = true;
            // pending: the tile that holds the cell AND every tile that reads it as halo, i.e. the tiles of its out-neighbours
            // (flat-key arithmetic, so the column wrap of the reference's stencil is covered too)
            const uint32_t key = __float_as_uint(v);
            int last = -1;
#pragma unroll
            for (int d = -1; d < 8; ++d) {
                const long long kv = d < 0 ? k : k + nb_off(d, p.W);
                if (kv < 0 || kv >= p.N) continue;
                const int tile = (int)(kv / p.W) / TS * p.tiles_x + (int)(kv % p.W) / TS;
                if (tile == last) continue;
                last = tile;
                const uint32_t was = atomicMin(p.tile_key + tile, key);
                if (was == kNoKey) atomicAdd(p.counts + 1, 1);
            }
        }
    }
    const uint32_t b = __ballot_sync(0xffffffffu, better);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(improved, __popc(b));
}

// Key post with release semantics: orders this thread's earlier stores, and those of other threads it synchronised with
// (__syncwarp / __syncthreads), before the key: MEMBAR.ALL.GPU + ATOMG, against MEMBAR.SC.GPU + CCTL.IVALL (an L1 invalidate)
// + ATOMG for __threadfence() followed by a relaxed atomicMin.  Message passing needs no more than release / acquire.
__device__ __forceinline__ uint32_t atomic_min_release(uint32_t* addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.release.gpu.global.min.u32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}

// Relax one tile to its local fixed point (under the band limit thr).
//
// Inside the tile the warps run ASYNCHRONOUSLY: there is no block barrier per sweep.  Each warp owns four sub-blocks
// and polls one 32-bit word holding their four "due" flags (bytes, plain stores).  The owner clears a flag, relaxes
// that sub-block for up to p.inner warp-local passes, and flags the neighbouring sub-blocks whose border it moved.
// Relaxation is monotone (values only decrease towards the fixed point), so a stale read only costs an extra pass
// and a flag raised after the owner's clear simply re-queues the sub-block.  Termination: s_ctl[0] counts warps that
// found nothing due; the warp that makes it FIELD_WARPS verifies that every flag is zero and that the count still
// stands, then raises s_ctl[1].  (Measured on B200, C2 input: 4.7 ms against 5.3 ms for barrier-per-sweep.)
template <bool ASYNC_TILES>
__device__ __forceinline__ void relax_tile(const FieldParams& p, int tile, float thr, float* s, uint8_t* s_flag_, int* s_ctl_,
                                           uint32_t* s_emin, uint32_t* s_due, uint32_t key0) {
    volatile uint8_t* s_flag = s_flag_;
    volatile int* s_ctl = s_ctl_;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tile / p.tiles_x, tx = tile % p.tiles_x;
    const int r0 = ty * TS, c0 = tx * TS;
    const int W = p.W, H = p.H;
    const long long N = p.N;
#if NSFS_FIELD_PROF
    long long t_a = clock64(); unsigned my_iters = 0;
#endif
    // ---- stage the tile and its halo: position (rr, cc) holds g[(r0-1+rr)*W + (c0-1+cc)] (flat key!) -------
    for (int idx = tid; idx < SROWS * (TS + 2); idx += FIELD_THREADS) {
        const int rr = idx / (TS + 2), cc = idx - rr * (TS + 2);
        const long long key = (long long)(r0 - 1 + rr) * W + (c0 - 1 + cc);
        s[rr * SROW + (SCOL0 - 1) + cc] = (key >= 0 && key < N) ? __ldcg(p.g + key) : kInfF;   // L2: other SMs write g
    }
    if (tid < NSB) { s_flag[tid] = 1; s_due[tid] = key0; }            // first pass: everything is due (ordered: at the tile's own key)
    if (tid < 16) s_emin[tid] = kNoKey;                               // slot 11: this tile's own pending key; 12 / 13: peer shards
    if (tid < 4) s_ctl[tid] = 0;                                      // 0 idle warps, 1 done, 2 net change of the pending counter, 3 merges

    // ---- this thread's four cells ---------------------------------------------------------------------
    const int lr = lane >> 3, lc = lane & 7;
    uint32_t mask[SB_PER_WARP];
    int key[SB_PER_WARP];
    int soff[SB_PER_WARP];
#pragma unroll
    for (int q = 0; q < SB_PER_WARP; ++q) {
        const int sb = sb_of(warp, q);
        const int ly = (sb / SB_X) * SBH + lr, lx = (sb % SB_X) * SBW + lc;
        const int r = r0 + ly, c = c0 + lx;
        soff[q] = (ly + 1) * SROW + SCOL0 + lx;
        const bool real = r < H && c < W;
        key[q] = real ? r * W + c : -1;                               // N < 2^29 (nsfs_create)
        mask[q] = real ? (uint32_t)__ldg(p.in_mask + key[q]) : 0u;
    }
    __syncthreads();
    float v[SB_PER_WARP];
    uint32_t ever = 0;                                                // bit q: my cell of sub-block q improved in this visit
#pragma unroll
    for (int q = 0; q < SB_PER_WARP; ++q) v[q] = s[soff[q]];
#if NSFS_FIELD_PROF
    long long t_b = clock64();
#endif

#if NSFS_FIELD_ORDERED
    // ---- asynchronous relaxation in near-Dijkstra order -------------------------------------------------------
    // Each sub-block carries a due KEY (the smallest value that changed next to it) instead of a flag.  A warp takes its
    // due sub-block with the smallest key, but only while that key lies within delta_in of the smallest key due anywhere in
    // the tile: the wavefront is relaxed in order, so a sub-block is rarely relaxed against values that are about to drop.
    // (The warp holding the tile-wide minimum always qualifies, so there is always progress.)
    {
    volatile uint32_t* due = s_due;
    bool counted_idle = false;
    unsigned nap = p.wait_ns;                                         // idle back-off: short right after work, doubling up to idle_ns
    for (int spin = 0; spin < (1 << 26); ++spin) {
        uint32_t mymin = kNoKey;
        int q = 0;
#pragma unroll
        for (int qq = 0; qq < SB_PER_WARP; ++qq) {
            const uint32_t kq = due[warp * SB_PER_WARP + qq];
            if (kq < mymin) { mymin = kq; q = qq; }
        }
        bool did = false;
        if (mymin != kNoKey) {
            uint32_t gmin = mymin;
            if (p.delta_in < 1e8f) {                                  // optional tile-wide gate (off by default: measured no gain)
                uint32_t g4 = kNoKey;
#pragma unroll
                for (int t = 0; t < NSB / 32; ++t) g4 = min(g4, due[lane + 32 * t]);
                gmin = min(mymin, __reduce_min_sync(0xffffffffu, g4));
            }
            if (__uint_as_float(mymin) <= __uint_as_float(gmin) + p.delta_in) {
                const int sb = sb_of(warp, q);
                if (counted_idle) {                                   // leave the idle count BEFORE consuming the key
                    if (lane == 0) atomicSub(s_ctl_, 1);
                    counted_idle = false;
                }
                __syncwarp();
                if (lane == 0) atomicExch(s_due + warp * SB_PER_WARP + q, kNoKey);   // clear first, then read the neighbours
                __syncwarp();
                // registers cannot be indexed by a run-time q: select
                uint32_t m = 0;
                float vq = 0.f;
                int so = 0;
#pragma unroll
                for (int qq = 0; qq < SB_PER_WARP; ++qq) if (qq == q) { m = mask[qq]; vq = v[qq]; so = soff[qq]; }
                const volatile float* cc = s + so;
                uint32_t bits = 0, last = 0;
                float kchg = __uint_as_float(kNoKey);
                // edge weights of my cell for this visit of the sub-block: 1, sqrt(2), or +inf where there is no in-edge
                float wgt[8];
#pragma unroll
                for (int d = 0; d < 8; ++d)
                    wgt[d] = ((m >> d) & 1u) ? (((m >> (8 + d)) & 1u) ? kSqrt2F : 1.0f) : __uint_as_float(kNoKey);
                for (int it = 0; it < p.inner; ++it) {
                    // (measured and rejected: two passes per vote - no change; neighbours of the same sub-block through shuffles
                    //  with the outside ones cached per visit - 2.5 ms instead of 1.5: the live view of the other warps' cells matters)
                    float cd[8];
#pragma unroll
                    for (int d = 0; d < 8; ++d) cd[d] = cc[-nb_di(d) * SROW - nb_dj(d)] + wgt[d];
                    // min as a tree (3 levels instead of a chain of 8); the band test on the minimum equals the test on every candidate
                    const float best = fminf(fminf(fminf(cd[0], cd[1]), fminf(cd[2], cd[3])), fminf(fminf(cd[4], cd[5]), fminf(cd[6], cd[7])));
                    const bool changed = best < vq && best <= thr;
                    if (changed) { vq = best; s[so] = best; ever |= 1u << q; kchg = fminf(kchg, best); }
#if NSFS_FIELD_PROF
                    my_iters++;
#endif
                    last = __ballot_sync(0xffffffffu, changed);
                    if (!last) break;
                    bits |= last;
                    __syncwarp();
                }
#pragma unroll
                for (int qq = 0; qq < SB_PER_WARP; ++qq) if (qq == q) v[qq] = vq;
                if (bits) {
                    const uint32_t kmin = __reduce_min_sync(0xffffffffu, __float_as_uint(kchg));
                    __syncwarp();
                    if (lane < 9) {
                        const int dy = lane / 3 - 1, dx = lane % 3 - 1;
                        uint32_t need = (dy == 0 && dx == 0) ? last : bits;
                        if (dy < 0) need &= 0x000000ffu; else if (dy > 0) need &= 0xff000000u;
                        if (dx < 0) need &= 0x01010101u; else if (dx > 0) need &= 0x80808080u;
                        const int sy = sb / SB_X + dy, sx = sb % SB_X + dx;
                        const int t = sy * SB_X + sx;
                        if (need && sy >= 0 && sy < SB_Y && sx >= 0 && sx < SB_X)
                            atomicMin(s_due + slot_of(t), kmin);
                    }
                    if (ASYNC_TILES && p.early) {
                        // cut-through between tiles: a changed cell that faces another tile is published NOW (value, fence, key)
                        // instead of at the end of the visit, so the neighbouring tile starts on the wavefront while this one is
                        // still relaxing; the end-of-visit posts to the 8 direct neighbours are then redundant and skipped
                        const int sby = sb / SB_X, sbx = sb % SB_X;
                        uint32_t outer = 0;
                        if (sby == 0) outer |= bits & 0x000000ffu;
                        if (sby == SB_Y - 1) outer |= bits & 0xff000000u;
                        if (sbx == 0) outer |= bits & 0x01010101u;
                        if (sbx == SB_X - 1) outer |= bits & 0x80808080u;
                        if (outer) {
                            if ((outer >> lane) & 1u) {
                                int gk = -1;
#pragma unroll
                                for (int qq = 0; qq < SB_PER_WARP; ++qq) if (qq == q) gk = key[qq];
                                if (gk >= 0) __stcg(p.g + gk, vq);
                            }
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
                                    if (old > p.cap && kmin <= p.cap) atomicAdd(p.gcount, 1);
                                }
                            }
                        }
                    }
                }
                did = true;
                nap = p.wait_ns;
            } else {
                __nanosleep(p.wait_ns);                               // due, but ahead of the wavefront: not idle, just early
                continue;
            }
        }
        if (did) continue;
        if (!counted_idle) {
            int old = 0;
            if (lane == 0) old = atomicAdd(s_ctl_, 1);
            old = __shfl_sync(0xffffffffu, old, 0);
            counted_idle = true;
            if (old == FIELD_WARPS - 1) {                             // every warp is idle-counted: verify and finish
                uint32_t g4 = kNoKey;
#pragma unroll
                for (int t = 0; t < NSB / 32; ++t) g4 = min(g4, due[lane + 32 * t]);
                const bool any = __any_sync(0xffffffffu, g4 != kNoKey);
                if (!any && s_ctl[0] == FIELD_WARPS) {
                    // Before ending the visit: did a neighbour post to THIS tile meanwhile (within the band)?  Then stay on it:
                    // take the key, re-read the halo ring, wake the sub-blocks next to halo cells that dropped.  Saves the
                    // publish / claim / stage round trip of a separate visit, which sits on the wavefront's critical path.
                    uint32_t pk = kNoKey;
                    if (ASYNC_TILES && p.merge) {
                        if (lane == 0) {
                            pk = __ldcg(p.tile_key + tile);
                            pk = (pk <= __float_as_uint(thr) && pk <= p.cap) ? atomicExch(p.tile_key + tile, kNoKey) : kNoKey;
                            if (pk != kNoKey) {
                                __threadfence();
                                s_ctl_[3] += 1;                       // two counted units (pending + in flight) become one
                                atomicSub(s_ctl_, 1);                 // I am busy again: nobody else can end the visit while I refresh the halo
                            }
                        }
                        pk = __shfl_sync(0xffffffffu, pk, 0);
                    }
                    if (pk != kNoKey) {
                        counted_idle = false;
                        for (int hx = lane; hx < 4 * TS + 4; hx += 32) {
                            int rr, cc;                                       // position on the ring of the (TS+2)^2 staged square
                            if (hx < TS + 2) { rr = 0; cc = hx; }
                            else if (hx < 2 * (TS + 2)) { rr = TS + 1; cc = hx - (TS + 2); }
                            else if (hx < 2 * (TS + 2) + TS) { rr = hx - 2 * (TS + 2) + 1; cc = 0; }
                            else { rr = hx - 2 * (TS + 2) - TS + 1; cc = TS + 1; }
                            const long long gk = (long long)(r0 - 1 + rr) * W + (c0 - 1 + cc);
                            if (gk < 0 || gk >= N) continue;
                            const float nv = __ldcg(p.g + gk);
                            const int so2 = rr * SROW + (SCOL0 - 1) + cc;
                            if (nv < s[so2]) {
                                s[so2] = nv;
                                // interior cells within one step of the halo cell: rows rr-2 .. rr, columns cc-2 .. cc (tile coordinates)
                                const int y0 = max(rr - 2, 0), y1 = min(rr, TS - 1), x0 = max(cc - 2, 0), x1 = min(cc, TS - 1);
                                const uint32_t kb = __float_as_uint(nv);
                                for (int yy = 0; yy < 2; ++yy)
                                    for (int xx = 0; xx < 2; ++xx) {
                                        const int t2 = ((yy ? y1 : y0) / SBH) * SB_X + (xx ? x1 : x0) / SBW;
                                        atomicMin(s_due + slot_of(t2), kb);
                                    }
                            }
                        }
                        __syncwarp();
                        continue;                                     // poll again: my own sub-blocks may be due; idle + finisher check come after
                    }
                    if (lane == 0) s_ctl[1] = 1;
                }
            }
        }
        if (s_ctl[1]) break;
        __nanosleep(nap);
        nap = min(nap * 2u, p.idle_ns);
    }
    }
#else
    // ---- asynchronous relaxation -----------------------------------------------------------------------
    bool counted_idle = false;
    for (int spin = 0; spin < (1 << 26); ++spin) {
        bool did = false;
        const uint32_t due = reinterpret_cast<const volatile uint32_t*>(s_flag_)[warp];   // my four flags in one load
#pragma unroll
        for (int q = 0; q < SB_PER_WARP && due != 0u; ++q) {
            const int sb = sb_of(warp, q);
            if (!((due >> (8 * q)) & 0xffu)) continue;                // warp-uniform
            if (counted_idle) {                                       // leave the idle count BEFORE consuming the flag
                if (lane == 0) atomicSub(s_ctl_, 1);
                counted_idle = false;
            }
            __syncwarp();
            if (lane == 0) s_flag[warp * SB_PER_WARP + q] = 0;       // clear first, then read the neighbours (volatile: program order)
            __syncwarp();
            const volatile float* c = s + soff[q];
            const uint32_t m = mask[q];
            uint32_t bits = 0, last = 0;
            for (int it = 0; it < p.inner; ++it) {
                float best = v[q];
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    // in-edge d arrives from u = v - OFF[d], i.e. tile position (ly - DI, lx - DJ)
                    const float nb = c[-nb_di(d) * SROW - nb_dj(d)];
                    const float cand = nb + (((m >> (8 + d)) & 1u) ? kSqrt2F : 1.0f);
                    if (((m >> d) & 1u) && cand <= thr) best = fminf(best, cand);    // beyond the band: a later round
                }
                const bool changed = best < v[q];
                if (changed) { v[q] = best; s[soff[q]] = best; ever |= 1u << q; }
#if NSFS_FIELD_PROF
                my_iters++;
#endif
                last = __ballot_sync(0xffffffffu, changed);
                if (!last) break;
                bits |= last;
                __syncwarp();                                         // this warp's stores are visible to its next pass
            }
            if (bits) {
                __syncwarp();                                         // values first, then the flags that announce them
                if (lane < 9) {
                    // lane l decides for neighbour (dy, dx) = (l/3 - 1, l%3 - 1) whether a changed cell borders it
                    const int dy = lane / 3 - 1, dx = lane % 3 - 1;
                    uint32_t need = (dy == 0 && dx == 0) ? last : bits;   // self: only if the pass budget ran out mid-change
                    if (dy < 0) need &= 0x000000ffu; else if (dy > 0) need &= 0xff000000u;
                    if (dx < 0) need &= 0x01010101u; else if (dx > 0) need &= 0x80808080u;
                    const int sy = sb / SB_X + dy, sx = sb % SB_X + dx;
                    const int t = sy * SB_X + sx;                     // flag slot = owner warp * 4 + its q
                    if (need && sy >= 0 && sy < SB_Y && sx >= 0 && sx < SB_X) s_flag[slot_of(t)] = 1;
                }
            }
            did = true;
        }
        if (did) continue;
        if (!counted_idle) {
            int old = 0;
            if (lane == 0) old = atomicAdd(s_ctl_, 1);
            old = __shfl_sync(0xffffffffu, old, 0);
            counted_idle = true;
            if (old == FIELD_WARPS - 1) {                             // every warp is idle-counted: verify and finish
                const uint32_t f4 = lane < FIELD_WARPS ? reinterpret_cast<const volatile uint32_t*>(s_flag_)[lane] : 0u;
                const bool any = __any_sync(0xffffffffu, f4 != 0u);
                if (!any && s_ctl[0] == FIELD_WARPS && lane == 0) s_ctl[1] = 1;
            }
        }
        if (s_ctl[1]) break;
        __nanosleep(p.idle_ns);
    }
#endif
    __syncthreads();
#if NSFS_FIELD_PROF
    long long t_c = clock64();
#endif
    // ---- candidates the band rejected: the smallest one is when this tile wants to be relaxed again ------------
    {
        uint32_t rej = kNoKey;
#pragma unroll
        for (int q = 0; q < SB_PER_WARP; ++q) {
            const float* c = s + soff[q];
            const uint32_t m = mask[q];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const float cand = c[-nb_di(d) * SROW - nb_dj(d)] + (((m >> (8 + d)) & 1u) ? kSqrt2F : 1.0f);
                if (((m >> d) & 1u) && cand > thr && cand < v[q]) rej = min(rej, __float_as_uint(cand));
            }
        }
        rej = __reduce_min_sync(0xffffffffu, rej);
        if (lane == 0 && rej != kNoKey) atomicMin(s_emin + 11, rej);
    }

    // ---- write back what improved; per tile border, the smallest value that moved ---------------------------
#pragma unroll
    for (int q = 0; q < SB_PER_WARP; ++q) {
        const bool ch = (ever >> q) & 1u;
        const int sb = sb_of(warp, q);
        const int sy = sb / SB_X, sx = sb % SB_X;
        const int ly = sy * SBH + lr, lx = sx * SBW + lc;
        const uint32_t vb = ch ? __float_as_uint(v[q]) : kNoKey;
        if (ch) {
            const int r = r0 + ly;
            if (ASYNC_TILES && (p.peer_g[0] || p.peer_g[1])) {
                if (r < p.own_lo || r >= p.own_hi) {
                    // a ghost row: the neighbour shard lowers it too (remote atomicMin), so my store must not raise it
                    atomicMin(reinterpret_cast<uint32_t*>(p.g) + key[q], vb);
                } else {
                    __stcg(p.g + key[q], v[q]);
                    // my first / last two owned rows ARE the neighbour's ghost rows: push the value over NVLink
                    if (p.peer_g[0] && r < p.own_lo + 2) {
                        atomicMin(reinterpret_cast<uint32_t*>(p.peer_g[0]) + (long long)(r + p.peer_row_off[0]) * W + (c0 + lx), vb);
                        atomicMin(s_emin + 12, vb);
                    }
                    if (p.peer_g[1] && r >= p.own_hi - 2) {
                        atomicMin(reinterpret_cast<uint32_t*>(p.peer_g[1]) + (long long)(r + p.peer_row_off[1]) * W + (c0 + lx), vb);
                        atomicMin(s_emin + 13, vb);
                    }
                }
            } else {
                __stcg(p.g + key[q], v[q]);
            }
        }
        if (!__any_sync(0xffffffffu, ch)) continue;
        // slot (dy+1)*3 + (dx+1): neighbour tile (dy, dx) reads a cell of ours that changed; 9 / 10: map edge wrap
        auto post = [&](bool cond, int slot) {
            const uint32_t m = __reduce_min_sync(0xffffffffu, cond ? vb : kNoKey);
            if (lane == 0 && m != kNoKey) atomicMin(s_emin + slot, m);
        };
        if (sy == 0) { post(ly == 0, 1); if (sx == 0) post(ly == 0 && lx == 0, 0); if (sx == SB_X - 1) post(ly == 0 && lx == TS - 1, 2); }
        if (sy == SB_Y - 1) { post(ly == TS - 1, 7); if (sx == 0) post(ly == TS - 1 && lx == 0, 6); if (sx == SB_X - 1) post(ly == TS - 1 && lx == TS - 1, 8); }
        if (sx == 0) post(lx == 0, 3);
        if (sx == SB_X - 1) post(lx == TS - 1, 5);
        if (tx == p.tiles_x - 1) post(c0 + lx == W - 1, 9);          // flat-key wrap: (r, W-1) feeds (r..r+2, 0)
        if (tx == 0 && sx == 0) post(lx == 0, 10);                    // flat-key wrap: (r, 0) feeds (r-2..r, W-1)
    }
    // The improved cells must be visible before the keys that announce them.  Only warp 0 posts keys, so only warp 0 fences:
    // the block barrier orders every thread's stores before it, and a fence is cumulative over what happened before it
    // (1 fence per visit instead of 32: __threadfence is MEMBAR.SC.GPU + an L1 invalidate on sm_100).
    __syncthreads();
    if (tid < 32) {
        if (ASYNC_TILES && (p.peer_g[0] || p.peer_g[1])) __threadfence_system();   // ... also on the neighbour GPU
        else if (!p.light) __threadfence();
    }
    int newly = 0;        // tiles this thread made pending: summed over warp 0 and folded into ONE update of the pending counter
    if (ASYNC_TILES && tid >= 16 && tid < 32) {
        // wake the neighbour shard's tiles that hold, or read as halo, the two rows just pushed: its rows R-3 .. R+4 lie in at
        // most two tile rows; columns tx-1 .. tx+1 plus the far edge column when this tile touches a map edge (flat-key wrap)
        const int side = (tid - 16) >> 3, slot = (tid - 16) & 7;
        const uint32_t m = s_emin[12 + side];
        if (p.peer_key[side] && m != kNoKey) {
            const int R = (side == 0 ? p.own_lo : p.own_hi - 2) + p.peer_row_off[side];      // first pushed row, neighbour-local
            const int tr0 = max(R - 3, 0) / TS, tr1 = min(R + 4, p.peer_tiles_y[side] * TS - 1) / TS;
            const int trow = (slot & 1) ? tr1 : tr0;
            const int cs = slot >> 1;                                      // 0..2: tx-1..tx+1, 3: the wrap column
            int tcol = cs < 3 ? tx + cs - 1 : (tx == 0 ? p.tiles_x - 1 : (tx == p.tiles_x - 1 ? 0 : -1));
            bool dup = ((slot & 1) && tr1 == tr0) || tcol < 0 || tcol >= p.tiles_x;
            if (cs == 3 && !dup) dup = abs(tcol - tx) <= 1;                // already covered by the three neighbour columns
            if (!dup) {
                const uint32_t old = atomicMin(p.peer_key[side] + trow * p.tiles_x + tcol, m);
                if (old > p.cap && m <= p.cap) newly = 1;
            }
        }
    }
    {
        // key posts: <9 neighbour tiles, 9..11 right map edge -> left-edge tiles of tile rows ty-1..ty+1,
        // 12..14 left map edge -> right-edge tiles of rows ty-1..ty+1, 15 this tile's own rejected candidates
        int target = -1;
        uint32_t m = kNoKey;
        if (tid < 9 && tid != 4) {
            const int ny = ty + tid / 3 - 1, nx = tx + tid % 3 - 1;
            m = s_emin[tid];
            if (ny >= 0 && ny < p.tiles_y && nx >= 0 && nx < p.tiles_x && !(ASYNC_TILES && p.early)) target = ny * p.tiles_x + nx;
        } else if (tid >= 9 && tid < 12) {
            const int ny = ty + (tid - 10);
            m = s_emin[9];
            if (ny >= 0 && ny < p.tiles_y) target = ny * p.tiles_x;
        } else if (tid >= 12 && tid < 15) {
            const int ny = ty + (tid - 13);
            m = s_emin[10];
            if (ny >= 0 && ny < p.tiles_y) target = ny * p.tiles_x + p.tiles_x - 1;
        } else if (tid == 15) {
            m = s_emin[11];
            target = tile;
        }
        if (target >= 0 && m != kNoKey) {
            const uint32_t old = p.light ? atomic_min_release(p.tile_key + target, m) : atomicMin(p.tile_key + target, m);
            if (ASYNC_TILES && old > p.cap && m <= p.cap) newly = 1;     // a tile became pending (within the cap)
        }
    }
    if (ASYNC_TILES && tid < 32) {
        // the visit's net effect on "pending tiles + visits in flight": +newly pending, -1 for this visit.  One atomic, so the
        // counter never dips below what is really outstanding and no second fence is needed before it.
        const int net = __reduce_add_sync(0xffffffffu, newly) - 1 - s_ctl_[3];
        if (tid == 0) s_ctl_[2] = net;
    }
    if (tid == 0) atomicAdd(p.counts + 3, 1);
#if NSFS_FIELD_PROF
    { long long t_d = clock64();
      if (lane == 0) atomicAdd(p.prof + 3, (unsigned long long)my_iters);
      { __shared__ unsigned s_it; __shared__ unsigned s_itmax; if (tid == 0) { s_it = 0; s_itmax = 0; } __syncthreads(); if (lane == 0) { atomicAdd(&s_it, my_iters); atomicMax(&s_itmax, my_iters); } __syncthreads();
        if (tid == 0) atomicMax(p.prof + 8, ((unsigned long long)(t_c - t_b) << 32) | ((unsigned long long)(s_itmax & 0xffff) << 16) | (s_it & 0xffff)); }
      if (tid == 0) { atomicAdd(p.prof + 0, (unsigned long long)(t_b - t_a)); atomicAdd(p.prof + 1, (unsigned long long)(t_c - t_b)); atomicAdd(p.prof + 2, (unsigned long long)(t_d - t_c)); } }
#endif
    __syncthreads();   // smem is reused by the next tile of this block
}

// Block 0 picks the tiles of the next round: all tiles whose pending key lies within delta of the smallest pending
// key (a delta-stepping bucket over tiles), compacted with __ballot_sync/__popc.  Near-Dijkstra order keeps a tile
// from being relaxed against a halo the wavefront has not reached yet.
__device__ __forceinline__ void select_tiles(const FieldParams& p, int* s_int) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_tiles = p.tiles_x * p.tiles_y;
    if (tid == 0) { s_int[0] = (int)kNoKey; s_int[1] = 0; }
    __syncthreads();
    uint32_t mn = kNoKey;
    for (int i = tid; i < n_tiles; i += FIELD_THREADS) mn = min(mn, p.tile_key[i]);
    mn = __reduce_min_sync(0xffffffffu, mn);
    if (lane == 0 && mn != kNoKey) atomicMin((uint32_t*)s_int, mn);
    __syncthreads();
    const uint32_t gmin = (uint32_t)s_int[0];
    if (gmin <= p.cap) {
        const float prev = __int_as_float(p.counts[7]);
        const float thr_f = fminf(fmaxf(__uint_as_float(gmin) + p.delta, prev + p.alpha * p.delta), __uint_as_float(p.cap));
        const uint32_t thr = __float_as_uint(thr_f);
        if (tid == 0) p.counts[7] = __float_as_int(thr_f);
        for (int base = 0; base < n_tiles; base += FIELD_THREADS) {
            const int i = base + tid;
            const bool f = i < n_tiles && p.tile_key[i] <= thr;
            const uint32_t b = __ballot_sync(0xffffffffu, f);
            if (b) {
                int off = 0;
                if (lane == 0) off = atomicAdd(s_int + 1, __popc(b));
                off = __shfl_sync(0xffffffffu, off, 0);
                if (f) {
                    p.list[off + __popc(b & ((1u << lane) - 1u))] = i;
                    p.tile_key[i] = kNoKey;
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) p.counts[0] = s_int[1];
}

__global__ void __launch_bounds__(FIELD_THREADS, FIELD_BLOCKS_PER_SM) field_solve_kernel(FieldParams p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s[SROWS * SROW];
    __shared__ __align__(16) uint8_t s_flag[NSB];
    __shared__ __align__(16) uint32_t s_due[NSB];
    __shared__ int s_ctl[4];
    __shared__ uint32_t s_emin[16];
    __shared__ int s_int[2];

#if NSFS_FIELD_PROF
    long long t_k0 = clock64(), t_relax = 0;
#endif
    for (int round = 0; round < p.max_rounds; ++round) {
#if NSFS_FIELD_PROF
        long long t_s0 = clock64();
#endif
        if (blockIdx.x == 0) { select_tiles(p, s_int); __threadfence(); }
#if NSFS_FIELD_PROF
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.prof + 7, (unsigned long long)(clock64() - t_s0));
#endif
        grid.sync();
        const int n_cur = *((volatile int*)p.counts);
        if (n_cur == 0) break;
        const float thr = __int_as_float(*((volatile int*)(p.counts + 7)));
#if NSFS_FIELD_PROF
        long long t_r0 = clock64();
#endif
        for (int t = blockIdx.x; t < n_cur; t += gridDim.x) relax_tile<false>(p, p.list[t], thr, s, s_flag, s_ctl, s_emin, s_due, __float_as_uint(fmaxf(thr - p.delta, 0.f)));
#if NSFS_FIELD_PROF
        { long long dt = clock64() - t_r0; t_relax += dt;
          if (threadIdx.x == 0 && round < 4000) atomicMax(p.prof + 16 + round, (unsigned long long)dt); }
#endif
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) p.counts[2] = round + 1;
    }
#if NSFS_FIELD_PROF
    if (threadIdx.x == 0) { atomicAdd(p.prof + 4, (unsigned long long)(clock64() - t_k0)); atomicAdd(p.prof + 5, (unsigned long long)t_relax); }
#endif
}

// Asynchronous variant: no rounds and no grid-wide barrier.  Tile t is owned by block t % gridDim.x (round-robin, so
// the tiles along the wavefront belong to different SMs).  Each block repeatedly takes its pending tile with the
// smallest key, relaxes it within the band [key, key + delta], and posts keys to the tiles whose halo it moved.
// counts[1] = tiles with a pending key + visits in flight; the kernel ends when it reaches zero.  Every spin loop has a
// bail-out (counts[4] bit 1) so the kernel cannot hang.
__global__ void __launch_bounds__(FIELD_THREADS, FIELD_BLOCKS_PER_SM) field_solve_async_kernel(FieldParams p) {
    __shared__ float s[SROWS * SROW];
    __shared__ __align__(16) uint8_t s_flag[NSB];
    __shared__ __align__(16) uint32_t s_due[NSB];
    __shared__ int s_ctl[4];
    __shared__ uint32_t s_emin[16];
    __shared__ unsigned long long s_best;
    __shared__ uint32_t s_key;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_tiles = p.tiles_x * p.tiles_y;
    int idle_spins = 0;
    for (long long iter = 0; iter < (1ll << 40); ++iter) {
        if (tid == 0) s_best = ~0ull;
        __syncthreads();
        unsigned long long best = ~0ull;
        if (p.dynamic) {
            // block b competes for the tiles of its group (t % groups == b % groups): groups = 1 is one global queue,
            // groups = gridDim.x is static ownership
            const int G = p.groups;
            for (int t = (int)(blockIdx.x % G) + tid * G; t < n_tiles; t += G * FIELD_THREADS) {
                const uint32_t k = __ldcg(p.tile_key + t);
                if (k <= p.cap && !__ldcg(p.tile_busy + t)) best = min(best, ((unsigned long long)k << 32) | (unsigned)t);
            }
        } else {
            for (int t = blockIdx.x + tid * gridDim.x; t < n_tiles; t += gridDim.x * FIELD_THREADS) {
                const uint32_t k = __ldcg(p.tile_key + t);
                if (k <= p.cap) best = min(best, ((unsigned long long)k << 32) | (unsigned)t);
            }
        }
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0 && best != ~0ull) atomicMin(&s_best, best);
        __syncthreads();
        unsigned long long pick = s_best;
        if (p.gate > 0.f) {
            // global gate: publish my smallest pending key, read everybody's, and leave my tile alone if it is far ahead of the
            // slowest part of the wavefront (it would only be relaxed against values that are still going to drop)
            if (tid == 0) { __stcg(p.blk_min + blockIdx.x, (uint32_t)(pick >> 32)); s_key = kNoKey; }   // ~0ull >> 32 == kNoKey-or-larger: "nothing"
            __syncthreads();
            uint32_t gm = kNoKey;
            for (int b = tid; b < (int)gridDim.x; b += FIELD_THREADS) gm = min(gm, __ldcg(p.blk_min + b));
            gm = __reduce_min_sync(0xffffffffu, gm);
            if (lane == 0 && gm < kNoKey) atomicMin(&s_key, gm);
            __syncthreads();
            const uint32_t gmin = s_key;
            __syncthreads();
            if (pick != ~0ull && gmin < kNoKey && __uint_as_float((uint32_t)(pick >> 32)) > __uint_as_float(gmin) + p.gate) {
                if (++idle_spins > (1 << 22)) { if (tid == 0) atomicOr(p.counts + 4, 2); break; }
                __nanosleep(200);
                continue;
            }
        }
        if (pick == ~0ull) {
            int pending = 1;
            if (tid == 0) { pending = atomicAdd(p.gcount, 0); s_key = (uint32_t)pending; }
            __syncthreads();
            pending = (int)s_key;
            __syncthreads();
            if (pending <= 0) break;
            if (++idle_spins > (1 << 22)) { if (tid == 0) atomicOr(p.counts + 4, 2); break; }
            __nanosleep(p.sched_ns);
            continue;
        }
        idle_spins = 0;
        const int tile = (int)(pick & 0xffffffffu);
        if (tid == 0) {
            uint32_t k = kNoKey;
            if (!p.dynamic || atomicCAS(p.tile_busy + tile, 0, 1) == 0) {
                k = atomicExch(p.tile_key + tile, kNoKey);
                if (p.dynamic && k > p.cap) {                  // someone relaxed it between my scan and my claim
                    if (k != kNoKey) atomicMin(p.tile_key + tile, k);
                    k = kNoKey;
                    __threadfence();
                    atomicExch(p.tile_busy + tile, 0);
                }
            }
            s_key = k;
            __threadfence();
        }
        __syncthreads();
        const uint32_t my_key = s_key;
        __syncthreads();
        if (my_key == kNoKey) continue;                        // lost the race for this tile: look again
        const float thr = fminf(__uint_as_float(my_key) + p.delta, __uint_as_float(p.cap));
        relax_tile<true>(p, tile, thr, s, s_flag, s_ctl, s_emin, s_due, my_key);
        if (tid == 0) {
            if (p.dynamic) { __threadfence(); atomicExch(p.tile_busy + tile, 0); }
            if (s_ctl[2] != 0) atomicAdd(p.gcount, s_ctl[2]);
        }
    }
}

// Before a (capped) solve: counts[1] = tiles whose pending key lies within the cap.  After it: counts[6] = the smallest
// key still pending (float bits, kNoKey = none), which the multi-GPU driver all-reduces to place the next band.
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
// pred[v] = lowest direction d whose in-edge attains g[v] exactly (the relaxation computed g[v] as exactly that
// fp32 sum, so equality is exact).  The reference keeps whichever tied parent its heap order met first; any
// of them is a valid optimal predecessor (north_star: path validity exact, cost within 1e-5).
__global__ void __launch_bounds__(256) field_pred_kernel(const float* __restrict__ g, const uint16_t* __restrict__ in_mask,
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
        const float gv = g[v];
        uint32_t pd = 255u;
        if (gv < kInfF) {
            const bool counted = v >= count_lo && v < count_hi;     // a row-partitioned shard counts its own rows only
            reached += counted;
            if (counted && (__ldg(out_mask + v) & kThrowBit)) thrown = 1;   // a reached cell is expanded; expanding this one throws
            if (v != ks) {
                const uint32_t m = __ldg(in_mask + v);
#pragma unroll
                for (int d = 7; d >= 0; --d) {
                    if (!((m >> d) & 1u)) continue;
                    const float cand = g[v - off[d]] + (((m >> (8 + d)) & 1u) ? kSqrt2F : 1.0f);
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

// ---- predecessor walk (goal -> source) -----------------------------------------------------------------------
// A dependent chain of 1-byte loads: one thread chases it (stores are fire-and-forget).
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
    NSFS_CUDA_TRY(h, cudaMalloc(&f.g, (size_t)(h->N + 4) * sizeof(float)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.pred, (size_t)h->N));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.tile_key, (size_t)(n_tiles + 32) * sizeof(uint32_t)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.list[0], (size_t)(n_tiles + 32) * sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.tile_busy, (size_t)(n_tiles + 32) * sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.counts, 16 * sizeof(int)));
    return NSFS_OK;
}

void field_free(nsfs_planner* h) {
    FieldWork& f = h->field;
    cudaFree(f.g); cudaFree(f.gx); cudaFree(f.pred); cudaFree(f.tile_key); cudaFree(f.list[0]); cudaFree(f.tile_busy); cudaFree(f.counts);
    f = FieldWork();
}

static FieldParams field_params(nsfs_planner* h) {
    FieldWork& f = h->field;
    FieldParams p;
    p.g = f.g; p.gx = f.gx; p.out_mask = h->d_out_mask; p.win_lo = 0; p.win_hi = h->N; p.in_mask = h->d_in_mask; p.tile_key = f.tile_key; p.list = f.list[0];
    p.tile_busy = f.tile_busy;
    p.blk_min = reinterpret_cast<uint32_t*>(f.list[0]);       // the bulk-synchronous variant's tile list is free in the asynchronous ones
    p.gate = h->opt_field_gate > 0 ? (float)h->opt_field_gate : 0.f;
    p.early = h->opt_field_early == 2 ? 0 : 1;                // 0 / 1 = on (default), 2 = off
    p.light = h->opt_field_fence == 1 ? 1 : 0;                // key posts as release atomics instead of __threadfence + atomicMin
    p.merge = h->opt_field_merge == 2 ? 0 : 1;
    p.sched_ns = h->opt_field_sched_ns > 0 ? (unsigned)h->opt_field_sched_ns : 400u;
    p.gcount = f.counts + 1;
    for (int sd = 0; sd < 2; ++sd) { p.peer_g[sd] = nullptr; p.peer_key[sd] = nullptr; p.peer_row_off[sd] = 0; p.peer_tiles_y[sd] = 0; }
    p.own_lo = 0; p.own_hi = h->H;
    // dynamic tile claiming (field_mode 2) is an option, not the default: on C2 one global queue cuts tile visits by 30 % but
    // herds the blocks onto the same few tiles (4.6 ms against 2.5 ms); 37 groups measured 2.33 ms, within noise of static
    p.dynamic = h->opt_field_mode == 2 ? 1 : 0;
    p.groups = h->opt_field_groups > 0 ? (int)h->opt_field_groups : 37;
    p.counts = f.counts; p.H = h->H; p.W = h->W; p.N = h->N; p.tiles_x = f.tiles_x; p.tiles_y = f.tiles_y;
    p.max_rounds = 1 << 30;
    p.prof = nullptr;
    p.inner = h->opt_field_inner > 0 ? (int)h->opt_field_inner : 16;
    p.idle_ns = h->opt_field_idle_ns > 0 ? (unsigned)h->opt_field_idle_ns : 1000u;
    p.delta = h->opt_field_delta > 0 ? (float)h->opt_field_delta : 128.0f;
    p.alpha = h->opt_field_alpha >= 0 ? (float)h->opt_field_alpha / 100.0f : 0.5f;
    p.delta_in = h->opt_field_delta_in > 0 ? (float)h->opt_field_delta_in : 1e9f;
    p.wait_ns = h->opt_field_wait_ns > 0 ? (unsigned)h->opt_field_wait_ns : 200u;
    p.cap = kNoKey - 1u;
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
    const bool bucket = h->opt_field_mode == 3;       // only the bucket solver keeps the second plane
    if (bucket && !f.gx) NSFS_CUDA_TRY(h, cudaMalloc(&f.gx, (size_t)(h->N + 4) * sizeof(float)));
    f.gx_ready = bucket;
    field_fill_kernel<<<(int)fill_blocks, 256, 0, h->stream>>>(reinterpret_cast<float4*>(f.g), n4, f.g, bucket ? reinterpret_cast<float4*>(f.gx) : nullptr,
                                                              f.gx, h->N, f.tile_key, f.tile_busy, f.tiles_x * f.tiles_y);
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
    field_inject_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(p, (long long)row0 * h->W, n, d_vals, d_improved);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

// Phase 3 (repeatable): relax every pending tile whose key is <= cap until none is left; nothing is relaxed beyond cap
// (cap = +inf: to the fixed point).  Leaves the smallest key still pending in counts[6].
int launch_field_relax(nsfs_planner* h, float cap) {
    FieldWork& f = h->field;
    FieldParams p = field_params(h);
    if (cap < 3.0e38f) { const float c = cap > 0.f ? cap : 0.f; memcpy(&p.cap, &c, sizeof(c)); }   // non-negative floats order like their bits
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, p.cap, f.counts, 0);
    h->launches++;
#if NSFS_FIELD_PROF
    static unsigned long long* d_prof = nullptr;
    if (!d_prof) cudaMalloc(&d_prof, 8 * 4100);
    cudaMemsetAsync(d_prof, 0, 8 * 4100, h->stream);
    p.prof = d_prof;
#endif
    if (h->opt_field_mode == 3 && f.gx_ready) {      // bucket-ordered tile solver (field_bucket.cu); a field begun under another
                                                     // mode has no gx plane and stays on the pull solver
        int rc = launch_field_relax_bucket(h, p);
        if (rc != NSFS_OK) return rc;
        field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, p.cap, f.counts, 1);
        h->launches++;
        NSFS_CUDA_TRY(h, cudaGetLastError());
        return NSFS_OK;
    }
    int per_sm = 0;
    const bool async_tiles = h->opt_field_mode != 1;
    NSFS_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, async_tiles ? field_solve_async_kernel : field_solve_kernel, FIELD_THREADS, 0));
    if (per_sm < 1) return set_cuda_error(h, cudaErrorLaunchOutOfResources, "field_solve_kernel occupancy");
    int grid = per_sm * h->sm_count;
    const int n_tiles = f.tiles_x * f.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    void* args[] = {&p};
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    // both variants need every block co-resident (grid barrier / spin-waiting on other blocks' posts): cooperative launch
    NSFS_CUDA_TRY(h, cudaLaunchCooperativeKernel(async_tiles ? (void*)field_solve_async_kernel : (void*)field_solve_kernel, dim3(grid),
                                                 dim3(FIELD_THREADS), args, 0, h->stream));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, f.tiles_x * f.tiles_y, p.cap, f.counts, 1);
    h->launches += 2;
#if NSFS_FIELD_PROF
    { static unsigned long long hp[4100]; cudaMemcpyAsync(hp, p.prof, 8 * 4100, cudaMemcpyDeviceToHost, h->stream); cudaStreamSynchronize(h->stream);
      unsigned long long summax = 0; int nr = 0; for (int r = 0; r < 4000; ++r) { summax += hp[16 + r]; if (hp[16 + r]) nr = r + 1; }
      fprintf(stderr, "[field prof] slowest visit: %llu cycles in the relax phase, %llu sub-block iterations in total, %llu in its busiest warp\n", hp[8] >> 32, hp[8] & 0xffff, (hp[8] >> 16) & 0xffff);
      fprintf(stderr, "[field prof] rounds %d: sum over rounds of slowest block's relax %llu cycles; block0 select total %llu cycles\n", nr, summax, hp[7]);
      fprintf(stderr, "[field prof] grid %d: per-block avg cycles: kernel %.0f relax %.0f | tile sums: load %.0f sweeps %.0f tail %.0f (cycles, summed over visits) | sub-block iterations %llu\n",
              grid, (double)hp[4] / grid, (double)hp[5] / grid, (double)hp[0], (double)hp[1], (double)hp[2], hp[3]); }
#endif
    return NSFS_OK;
}

// Peer-to-peer solve: the shared pending counter was set by the host (sum of every shard's pending tiles), every shard's
// kernel runs until the GLOBAL count is zero.  No cap, no rounds.
int launch_field_relax_p2p(nsfs_planner* h) {
    FieldWork& f = h->field;
    FieldParams p = field_params(h);
    for (int sd = 0; sd < 2; ++sd) {
        p.peer_g[sd] = f.peer_g[sd]; p.peer_key[sd] = f.peer_key[sd];
        p.peer_row_off[sd] = f.peer_row_off[sd]; p.peer_tiles_y[sd] = f.peer_tiles_y[sd];
    }
    if (f.own_hi > f.own_lo) { p.own_lo = f.own_lo; p.own_hi = f.own_hi; }
    p.gcount = (f.peer_counts ? f.peer_counts : f.counts) + 1;
    p.dynamic = 0;
    p.merge = 0;
    p.light = 0;            // peer posts need system scope: keep the fences
    p.early = 0;            // ghost rows are lowered with atomics by both sides; keep their publication at the end of the visit
    int per_sm = 0;
    NSFS_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, field_solve_async_kernel, FIELD_THREADS, 0));
    if (per_sm < 1) return set_cuda_error(h, cudaErrorLaunchOutOfResources, "field_solve_async_kernel occupancy");
    int grid = per_sm * h->sm_count;
    const int n_tiles = f.tiles_x * f.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    void* args[] = {&p};
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    NSFS_CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)field_solve_async_kernel, dim3(grid), dim3(FIELD_THREADS), args, 0, h->stream));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    field_pending_kernel<<<1, 256, 0, h->stream>>>(f.tile_key, n_tiles, p.cap, f.counts, 1);
    h->launches += 2;
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

// Phase 4: predecessor field, reached count and "would throw" flag over the rows [count_row_lo, count_row_hi).
int launch_field_finish(nsfs_planner* h) {
    FieldWork& f = h->field;
    const long long ks = f.src_i >= 0 ? (long long)f.src_i * h->W + f.src_j : -1;
    const long long lo = (long long)(h->opt_count_row_lo > 0 ? h->opt_count_row_lo : 0) * h->W;
    const long long hi = h->opt_count_row_hi > 0 ? (long long)h->opt_count_row_hi * h->W : h->N;
    long long blocks = (h->N + 255) / 256;
    if (blocks > (long long)h->sm_count * 16) blocks = (long long)h->sm_count * 16;      // 8 resident blocks per SM, two waves
    field_pred_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(f.g, h->d_in_mask, h->d_out_mask, h->W, (int)h->N, (int)ks, (int)lo, (int)hi,
                                                               f.pred, f.counts);
    h->launches++;
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
    field_path_kernel<<<1, 1, 0, h->stream>>>(f.g, f.pred, h->d_in_mask, h->W, h->N, ks, kg, d_path, cap, d_n, d_cost, d_status);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
