// Djikstra cost-to-go field, bucketed push expansion (field_mode 3): the in-tile solver that replaces the pull
// relaxation of field.cu where it applies (one GPU, or row-partitioned shards driven in rounds).
//
// Same result as field.cu (the unique fp32 fixed point of g[v] = min_u g[u] + w(u, v) on the reference's quirk graph,
// djikstra.cpp:27-145 / SURVEY.md R13), different schedule.  The reference's open list is keyed on (int)g and every edge
// costs >= 1, so all cells of one integer bucket [b, b+1) can be expanded together once the buckets below are done
// (delta-stepping with delta = 1 = the smallest edge weight: no light edges, one pass per bucket).  A tile visit is
// therefore a label-SETTING sweep over buckets, each cell expanded once per visit, instead of a label-correcting
// relaxation that re-reads 8 neighbours per cell per pass:
//
//   * 64x64 tile + 1-cell halo in shared memory, addressed by FLAT key (the reference's column wrap comes for free);
//     per cell: g (fp32), the 16-bit out-edge mask of map_pack.cu, one state byte (open / touched).
//   * 128 bucket lists of 16-bit cell ids; a bit map of the non-empty ones.  One step = one barrier among the 256
//     worker threads: find the lowest non-empty bucket, one thread per entry, 8 shared-memory atomicMin pushes, append
//     the lowered neighbours to the lists of buckets b+1 / b+2 (__ffs over the bit map picks the next bucket).
//   * edges that leave the tile lower the halo copy and mark it dirty; a ninth "courier" warp owns ALL global-memory
//     traffic during the visit: it drains dirty halo cells (atomicMin into g, fence, key post to the target tile), and
//     polls the tile's own key: when a neighbour has lowered border cells of this tile meanwhile, the courier reads the
//     border and hands the lower values to the workers through an inbox.  The workers never wait on HBM / L2.
//   * between visits a cell is "open" iff g < gx, gx = the value it was last expanded at (a second fp32 plane), so a visit
//     knows what is left to expand from the planes alone: whoever lowers g[v] only has to post the tile key.
//
// Tile scheduling between visits is the asynchronous scheme of field.cu: tile t belongs to block t % gridDim.x, a block
// takes its pending tile with the smallest key; counts[1] = pending tiles + visits in flight ends the kernel.
#include <cstdio>

#include "nsfs_internal.cuh"

namespace nsfs {

namespace {

constexpr int BT = kFieldTile;
constexpr int B_SROW = BT + 8;                  // smem row stride (floats)
constexpr int B_SCOL0 = 4;                      // smem column of tile column 0 (halo column -1 at 3)
constexpr int B_SROWS = BT + 2;
constexpr int B_POS = B_SROWS * B_SROW;         // staged positions (4752)
constexpr int B_HDW = (B_POS + 31) / 32;        // words of the dirty-halo bit map
constexpr int B_NB = 128;                       // bucket slots of a visit: [B0, B0 + B_NB)
constexpr int B_CAP = 144;                      // entries per bucket list; beyond: the bucket is found by a tile scan
constexpr int B_WORKERS = 256;
constexpr int B_THREADS = B_WORKERS + 32;       // + the courier warp
constexpr int B_CREDIT = 16;                    // a visit in flight counts 1 + B_CREDIT: covers key posts not yet accounted
constexpr int B_OUTCAP = 320;
constexpr int kIn = 1, kStop = 2, kEnd = 4;
static_assert(B_CAP <= B_WORKERS && B_NB % 32 == 0 && B_NB / 32 == 4, "one worker per list entry; the bit map is one uint4");

struct __align__(16) BucketSmem {
    float g[B_POS];
    uint32_t nonempty[B_NB / 32];
    int cnt[B_NB];
    uint32_t hdirty[B_HDW + 3];
    uint32_t inbox_val[256];
    uint32_t post_min[16];
    int post_tile[16];
    int dec[4];
    int ctl;                 // kIn: the inbox holds lower border values; kStop: end the visit now; kEnd: nothing more to merge
    int want_end;
    int in_cnt;
    int out_cnt;
    int B0;
    int smax;
    uint32_t red_min;
    uint32_t key;
    unsigned long long best;
    uint16_t inbox_cell[256];
    uint16_t out_list[B_OUTCAP];
    uint16_t om[BT * BT];
    uint16_t list[B_NB * B_CAP];
    uint8_t st[BT * BT];     // bit 0: open (lowered, not expanded at its value); bit 1: touched in this visit (write back)
    uint8_t ovf[B_NB];
};

__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(B_WORKERS) : "memory"); }

__device__ __forceinline__ int cell_pos(int c) { return ((c >> 6) + 1) * B_SROW + B_SCOL0 + (c & 63); }

__device__ __forceinline__ void bk_push(BucketSmem& S, int sv, int c) {
    const int idx = atomicAdd(&S.cnt[sv], 1);
    if (idx < B_CAP) S.list[sv * B_CAP + idx] = (uint16_t)c;
    else *(volatile uint8_t*)&S.ovf[sv] = 1;
    if (idx == 0) atomicOr(&S.nonempty[sv >> 5], 1u << (sv & 31));
}

// g[cell] was lowered from old_bits to val: mark it open and queue it unless it already waits in the same bucket.
__device__ __forceinline__ void bk_lowered(BucketSmem& S, int cv, float val, int old_bits, int B0, int smax) {
    volatile uint8_t* st = S.st;
    const uint32_t stv = st[cv];
    st[cv] = 3;
    const int sv = (int)val - B0;
    if (sv <= smax && !((stv & 1u) && (int)__int_as_float(old_bits) - B0 == sv)) bk_push(S, sv, cv);
}

// Expand cell c if it is open and sits in bucket slot s (stale list entries fail one of the two tests).
__device__ __forceinline__ void bk_expand(BucketSmem& S, int c, int s, int B0, int smax, int rh, int rw) {
    volatile uint8_t* st = S.st;
    if (!(st[c] & 1u)) return;
    const int ly = c >> 6, lx = c & 63;
    const int pos = (ly + 1) * B_SROW + B_SCOL0 + lx;
    const float gu = *(volatile float*)&S.g[pos];
    if ((int)gu - B0 != s) return;
    st[c] = 2;
    const uint32_t m = S.om[c];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        if (!((m >> d) & 1u)) continue;
        const float w = (d != 0 && ((m >> (8 + d)) & 1u)) ? kSqrt2F : 1.0f;
        const float cand = gu + w;
        const int di = nb_di(d), dj = nb_dj(d);
        const int pv = pos + di * B_SROW + dj;
        const int cb = __float_as_int(cand);
        const int old = atomicMin(reinterpret_cast<int*>(S.g) + pv, cb);
        if (old > cb) {
            if ((unsigned)(ly + di) < (unsigned)rh && (unsigned)(lx + dj) < (unsigned)rw) bk_lowered(S, c + di * BT + dj, cand, old, B0, smax);
            else atomicOr(&S.hdirty[pv >> 5], 1u << (pv & 31));
        }
    }
}

// ---- the courier warp's two jobs ---------------------------------------------------------------------------------
// (1) dirty halo cells -> g (atomicMin with the result: only a value that really lowered the neighbour's cell earns a key
// post; a neighbour that was lower already refreshes the halo copy instead), fence, key posts, pending-counter update.
__device__ __forceinline__ bool courier_drain(const FieldParams& p, BucketSmem& S, int ty, int tx, int r0, int c0, int lane) {
    bool any = false;
    for (int i = lane; i < B_HDW; i += 32) {
        uint32_t wd = *(volatile uint32_t*)&S.hdirty[i];
        if (wd) {
            wd = atomicExch(&S.hdirty[i], 0u);
            const int n = __popc(wd);
            int at = atomicAdd(&S.out_cnt, n);
            while (wd) {
                const int b = __ffs(wd) - 1;
                wd &= wd - 1;
                if (at < B_OUTCAP) S.out_list[at] = (uint16_t)(i * 32 + b);
                else atomicOr(&S.hdirty[i], 1u << b);                     // cannot happen (fewer halo cells than B_OUTCAP); keep it dirty
                ++at;
            }
            any = true;
        }
    }
    any = __any_sync(0xffffffffu, any);
    if (!any) return false;
    const int n = min(*(volatile int*)&S.out_cnt, B_OUTCAP);
    for (int e = lane; e < n; e += 32) {
        const int pv = S.out_list[e];
        const int rr = pv / B_SROW, cc = pv - rr * B_SROW - (B_SCOL0 - 1);
        const uint32_t vb = *(volatile uint32_t*)&S.g[pv];
        int row = r0 - 1 + rr, col = c0 - 1 + cc;
        const long long key = (long long)row * p.W + col;                 // in [0, N): out_mask only passes in-range free targets
        if (col < 0) { col += p.W; row -= 1; } else if (col >= p.W) { col -= p.W; row += 1; }
        const uint32_t old = atomicMin(reinterpret_cast<uint32_t*>(p.g) + key, vb);
        if (old > vb) {
            const int tty = row / BT, ttx = col / BT;
            const int dxx = ttx - tx;
            const int cls = (dxx >= -1 && dxx <= 1) ? dxx + 1 : (ttx == 0 ? 3 : 4);
            const int slot = (tty - ty + 1) * 5 + cls;
            atomicMin(&S.post_min[slot], vb);
            *(volatile int*)&S.post_tile[slot] = tty * p.tiles_x + ttx;
        } else if (old < vb) {
            atomicMin(reinterpret_cast<uint32_t*>(S.g) + pv, old);
        }
    }
    __threadfence();                  // the values are visible before the keys that announce them
    __syncwarp();
    int newly = 0;
    if (lane < 15) {
        const uint32_t m = atomicExch(&S.post_min[lane], kNoKey);
        if (m != kNoKey) {
            const int t = *(volatile int*)&S.post_tile[lane];
            const uint32_t old = atomicMin(p.tile_key + t, m);
            if (old > p.cap && m <= p.cap) newly = 1;
        }
    }
    newly = __reduce_add_sync(0xffffffffu, newly);
    if (lane == 0) {
        if (newly) atomicAdd(p.gcount, newly);
        *(volatile int*)&S.out_cnt = 0;
    }
    __syncwarp();
    return true;
}

// (2) the border cells of this tile as they are in g now; the ones below the staged copy go to the inbox.
__device__ __forceinline__ int courier_read_border(const FieldParams& p, BucketSmem& S, int r0, int c0, int rh, int rw, int lane) {
    const int nb = 2 * (rw + rh);                                         // <= 256: eight per lane, all loads in flight together
    uint32_t vals[8];
    int cells[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int i = q * 32 + lane;
        vals[q] = kNoKey;
        cells[q] = 0;
        if (i < nb) {
            int ly, lx;
            if (i < rw) { ly = 0; lx = i; }
            else if (i < 2 * rw) { ly = rh - 1; lx = i - rw; }
            else if (i < 2 * rw + rh) { ly = i - 2 * rw; lx = 0; }
            else { ly = i - 2 * rw - rh; lx = rw - 1; }
            cells[q] = ly * BT + lx;
            vals[q] = __float_as_uint(__ldcg(p.g + (long long)(r0 + ly) * p.W + (c0 + lx)));
        }
    }
    int base = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const bool lower = vals[q] < *(volatile uint32_t*)&S.g[cell_pos(cells[q])];
        const uint32_t bal = __ballot_sync(0xffffffffu, lower);
        if (lower) {
            const int k = base + __popc(bal & ((1u << lane) - 1u));
            S.inbox_cell[k] = (uint16_t)cells[q];
            S.inbox_val[k] = vals[q];
        }
        base += __popc(bal);
    }
    return base;
}

__device__ __forceinline__ void courier_loop(const FieldParams& p, BucketSmem& S, int tile, int ty, int tx, int r0, int c0, int rh, int rw,
                                             int B0, int smax, int lane) {
    const float lo_f = (float)B0, hi_f = (float)(B0 + smax + 1);
    for (int it = 0; it < (1 << 24); ++it) {
        const int we = *(volatile int*)&S.want_end;
        const int cf = *(volatile int*)&S.ctl;
        if (cf & (kStop | kEnd)) return;
        uint32_t pk = kNoKey;
        if (lane == 0 && !(cf & kIn)) pk = __ldcg(p.tile_key + tile);     // issued now, looked at after the drain
        const bool any = courier_drain(p, S, ty, tx, r0, c0, lane);
        pk = __shfl_sync(0xffffffffu, pk, 0);
        bool took = false;
        if (pk != kNoKey && pk <= p.cap) {
            const float pkf = __uint_as_float(pk);
            if (pkf < lo_f) {                                             // below this visit's buckets: end it, the tile is re-claimed
                if (lane == 0) atomicOr(&S.ctl, kStop);
                return;
            }
            if (pkf < hi_f) {
                if (lane == 0) {
                    atomicExch(p.tile_key + tile, kNoKey);
                    __threadfence();
                    atomicSub(p.gcount, 1);                               // pending + in flight become one unit
                    atomicAdd(p.counts + 9, 1);
                }
                __syncwarp();
                const int n = courier_read_border(p, S, r0, c0, rh, rw, lane);
                if (n > 0) {
                    __syncwarp();
                    if (lane == 0) {
                        *(volatile int*)&S.in_cnt = n;
                        __threadfence_block();
                        atomicExch(&S.want_end, 0);
                        atomicOr(&S.ctl, kIn);
                    }
                    took = true;
                }
            }
        }
        if (we && !took && !any) {
            if (lane == 0) atomicOr(&S.ctl, kEnd);
            return;
        }
        __syncwarp();
    }
    if (lane == 0) { atomicOr(p.counts + 4, 2); atomicOr(&S.ctl, kStop); }
}

// ---- the workers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void worker_loop(const FieldParams& p, BucketSmem& S, int B0, int smax, int rh, int rw, int tid) {
    int s_prev = -1, it = 0, steps = 0;
    for (;;) {
        if (tid == 0) *(volatile int*)&S.dec[it & 1] = *(volatile int*)&S.ctl;
        bar_workers();
        int f = *(volatile int*)&S.dec[it & 1];
        ++it;
        if (tid == 0 && s_prev >= 0) {                                    // the bucket of the last step is spent
            *(volatile int*)&S.cnt[s_prev] = 0;
            *(volatile uint8_t*)&S.ovf[s_prev] = 0;
            atomicAnd(&S.nonempty[s_prev >> 5], ~(1u << (s_prev & 31)));
        }
        if (f & kStop) break;
        int s = -1;
        if (!(f & kIn)) {
            const uint4 ne = *reinterpret_cast<const uint4*>(S.nonempty);   // after the barrier (memory clobber): re-read every step
            const int a = s_prev + 1;                                     // slots <= s_prev are spent (their bits may still read 1)
            auto from = [](uint32_t w, int a0) { return a0 >= 32 ? 0u : (a0 > 0 ? w & (0xffffffffu << a0) : w); };
            const uint32_t w0 = from(ne.x, a), w1 = from(ne.y, a - 32), w2 = from(ne.z, a - 64), w3 = from(ne.w, a - 96);
            s = w0 ? __ffs(w0) - 1 : (w1 ? 31 + __ffs(w1) : (w2 ? 63 + __ffs(w2) : (w3 ? 95 + __ffs(w3) : -1)));
            if (s < 0) {
                // out of work: ask the courier whether anything arrived for this tile meanwhile
                if (tid == 0) {
                    atomicExch(&S.want_end, 1);
                    int ff = 0, spins = 0;
                    while (!(ff = *(volatile int*)&S.ctl)) {
                        __nanosleep(40);
                        if (++spins > (1 << 24)) { atomicOr(p.counts + 4, 2); atomicOr(&S.ctl, kStop); }
                    }
                    *(volatile int*)&S.dec[2] = ff;
                }
                bar_workers();
                f = *(volatile int*)&S.dec[2];
                if (!(f & kIn)) break;                                    // kEnd or kStop
            }
        }
        if (f & kIn) {
            // lower border values from the courier: merge them, queue the cells, start again from the lowest bucket
            bar_workers();                                                // thread 0's reset of s_prev is done
            const int n = *(volatile int*)&S.in_cnt;
            for (int e = tid; e < n; e += B_WORKERS) {
                const int c = S.inbox_cell[e];
                const uint32_t vb = S.inbox_val[e];
                const int old = atomicMin(reinterpret_cast<int*>(S.g) + cell_pos(c), (int)vb);
                if (old > (int)vb) {
                    const float val = __uint_as_float(vb);
                    if ((int)val - B0 < 0) { *(volatile uint8_t*)&S.st[c] = 3; atomicOr(&S.ctl, kStop); }
                    else bk_lowered(S, c, val, old, B0, smax);
                }
            }
            bar_workers();
            if (tid == 0) {
                *(volatile int*)&S.in_cnt = 0;
                *(volatile int*)&S.want_end = 0;
                __threadfence_block();
                atomicAnd(&S.ctl, ~kIn);
            }
            s_prev = -1;
            continue;
        }
        const int n = *(volatile int*)&S.cnt[s];
        if (!*(volatile uint8_t*)&S.ovf[s]) {
            if (tid < n) bk_expand(S, S.list[s * B_CAP + tid], s, B0, smax, rh, rw);
        } else {
            for (int c = tid; c < BT * BT; c += B_WORKERS) bk_expand(S, c, s, B0, smax, rh, rw);
        }
        s_prev = s;
        ++steps;
    }
    if (tid == 0) atomicAdd(p.counts + 8, steps);
}

// One visit of one tile: stage, expand bucket by bucket, write back.
__device__ void bucket_visit(const FieldParams& p, BucketSmem& S, int tile, uint32_t key0) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int r0 = ty * BT, c0 = tx * BT;
    const int W = p.W;
    const long long N = p.N;
    const int rh = min(BT, p.H - r0), rw = min(BT, W - c0);
    const bool win = p.win_lo > 0 || p.win_hi < N;

    if (tid < B_NB) { S.cnt[tid] = 0; S.ovf[tid] = 0; }
    for (int i = tid; i < B_HDW + 3; i += B_THREADS) S.hdirty[i] = 0;
    if (tid < 4) S.nonempty[tid] = 0;
    if (tid < 16) S.post_min[tid] = kNoKey;
    if (tid == 0) { S.ctl = 0; S.want_end = 0; S.in_cnt = 0; S.out_cnt = 0; S.red_min = kNoKey; }
    __syncthreads();

    // ---- stage: position (rr, cc) holds g[(r0-1+rr)*W + (c0-1+cc)] (flat key); my cells also get mask and state ----
    uint32_t mn = kNoKey;
    for (int idx = tid; idx < B_SROWS * (BT + 2); idx += B_THREADS) {
        const int rr = idx / (BT + 2), cc = idx - rr * (BT + 2);
        const int ly = rr - 1, lx = cc - 1;
        const long long key = (long long)(r0 + ly) * W + (c0 + lx);
        const bool in_sq = (unsigned)ly < (unsigned)BT && (unsigned)lx < (unsigned)BT;
        const bool mine = ly >= 0 && ly < rh && lx >= 0 && lx < rw;
        float gv = kInfF;
        if (key >= 0 && key < N) gv = __ldcg(p.g + key);                  // L2: other SMs write g
        S.g[rr * B_SROW + (B_SCOL0 - 1) + cc] = gv;
        if (mine) {
            const float gxv = __ldcg(p.gx + key);
            uint32_t om = __ldg(p.out_mask + key);
            if (win) {                                                    // shards: a target outside the window takes no in-edge
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    const long long kv = key + nb_off(d, W);
                    if (kv < p.win_lo || kv >= p.win_hi) om &= ~(1u << d);
                }
            }
            const bool open = gv < gxv;
            S.om[ly * BT + lx] = (uint16_t)om;
            S.st[ly * BT + lx] = open ? 3 : 0;
            if (open) mn = min(mn, __float_as_uint(gv));
        } else if (in_sq) {
            S.om[ly * BT + lx] = 0;
            S.st[ly * BT + lx] = 0;
        }
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    if (lane == 0 && mn != kNoKey) atomicMin(&S.red_min, mn);
    __syncthreads();
    mn = S.red_min;
    const bool work = mn != kNoKey;
    int B0 = 0, smax = -1;
    if (work) {
        B0 = (int)__uint_as_float(mn);
        smax = min(B_NB - 1, (int)p.delta);
        if (p.cap < kNoKey - 1u) smax = min(smax, (int)__uint_as_float(p.cap) - B0);
        for (int c = tid; c < BT * BT; c += B_THREADS)
            if (S.st[c] & 1u) {
                const int sv = (int)S.g[cell_pos(c)] - B0;
                if (sv <= smax) bk_push(S, sv, c);
            }
    }
    __syncthreads();
    if (tid == 0) S.red_min = kNoKey;
    if (work) {
        if (tid < B_WORKERS) worker_loop(p, S, B0, smax, rh, rw, tid);
        else courier_loop(p, S, tile, ty, tx, r0, c0, rh, rw, B0, smax, lane);
    }
    __syncthreads();

    // ---- end of the visit: the courier flushes what is still dirty, the workers write the tile back --------------------
    if (work) {
        if (tid >= B_WORKERS) {
            courier_drain(p, S, ty, tx, r0, c0, lane);
        } else {
            if (*(volatile int*)&S.ctl & kIn) {
                // the courier took a post for this tile that the workers never merged (the visit was stopped): the lower border
                // values stay open, so the key this visit leaves behind covers them
                const int n = *(volatile int*)&S.in_cnt;
                for (int e = tid; e < n; e += B_WORKERS) {
                    const int c = S.inbox_cell[e];
                    const int old = atomicMin(reinterpret_cast<int*>(S.g) + cell_pos(c), (int)S.inbox_val[e]);
                    if (old > (int)S.inbox_val[e]) *(volatile uint8_t*)&S.st[c] = 3;
                }
            }
            bar_workers();
            uint32_t rej = kNoKey;
            for (int c = tid; c < BT * BT; c += B_WORKERS) {
                const uint32_t st = S.st[c];
                if (!(st & 2u)) continue;
                const int ly = c >> 6, lx = c & 63;
                const long long key = (long long)(r0 + ly) * W + (c0 + lx);
                const float gv = S.g[cell_pos(c)];
                // border cells are lowered by the neighbours' couriers too: never raise them
                if (ly == 0 || ly == rh - 1 || lx == 0 || lx == rw - 1) atomicMin(reinterpret_cast<uint32_t*>(p.g) + key, __float_as_uint(gv));
                else __stcg(p.g + key, gv);
                if (st & 1u) rej = min(rej, __float_as_uint(gv));         // still open: this tile wants another visit at that key
                else __stcg(p.gx + key, gv);
            }
            rej = __reduce_min_sync(0xffffffffu, rej);
            if (lane == 0 && rej != kNoKey) atomicMin(&S.red_min, rej);
        }
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        int net = -1 - B_CREDIT;
        const uint32_t rej = S.red_min;
        if (rej != kNoKey) {
            const uint32_t old = atomicMin(p.tile_key + tile, rej);
            if (old > p.cap && rej <= p.cap) net += 1;
        }
        atomicAdd(p.gcount, net);
        atomicAdd(p.counts + 3, 1);
        if (!work) atomicAdd(p.counts + 10, 1);
    }
    __syncthreads();
    (void)key0;
}

__global__ void __launch_bounds__(B_THREADS, 3) field_solve_bucket_kernel(FieldParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BucketSmem& S = *reinterpret_cast<BucketSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_tiles = p.tiles_x * p.tiles_y;
    int idle_spins = 0;
    for (long long iter = 0; iter < (1ll << 40); ++iter) {
        if (tid == 0) S.best = ~0ull;
        __syncthreads();
        unsigned long long best = ~0ull;
        for (int t = blockIdx.x + tid * gridDim.x; t < n_tiles; t += gridDim.x * B_THREADS) {
            const uint32_t k = __ldcg(p.tile_key + t);
            if (k <= p.cap) best = min(best, ((unsigned long long)k << 32) | (unsigned)t);
        }
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0 && best != ~0ull) atomicMin(&S.best, best);
        __syncthreads();
        const unsigned long long pick = S.best;
        if (pick == ~0ull) {
            if (tid == 0) S.key = (uint32_t)atomicAdd(p.gcount, 0);
            __syncthreads();
            const int pending = (int)S.key;
            __syncthreads();
            if (pending <= 0) break;
            if (++idle_spins > (1 << 22)) { if (tid == 0) atomicOr(p.counts + 4, 2); break; }
            __nanosleep(p.sched_ns);
            continue;
        }
        idle_spins = 0;
        const int tile = (int)(pick & 0xffffffffu);
        if (tid == 0) {
            const uint32_t k = atomicExch(p.tile_key + tile, kNoKey);
            if (k != kNoKey) atomicAdd(p.gcount, B_CREDIT);
            S.key = k;
            __threadfence();
        }
        __syncthreads();
        const uint32_t my_key = S.key;
        __syncthreads();
        if (my_key == kNoKey) continue;
        bucket_visit(p, S, tile, my_key);
    }
}

}  // namespace

// Same contract as launch_field_relax (field.cu): relax every pending tile whose key is <= cap until none is left.
int launch_field_relax_bucket(nsfs_planner* h, FieldParams p) {
    FieldWork& f = h->field;
    p.gx = f.gx;
    p.out_mask = h->d_out_mask;
    p.win_lo = (long long)(h->opt_active_row_lo > 0 ? h->opt_active_row_lo : 0) * h->W;
    p.win_hi = h->opt_active_row_hi > 0 ? (long long)h->opt_active_row_hi * h->W : h->N;
    p.delta = h->opt_field_delta > 0 ? (float)(h->opt_field_delta < B_NB - 1 ? h->opt_field_delta : B_NB - 1) : 100.0f;
    const size_t smem = sizeof(BucketSmem);
    NSFS_CUDA_TRY(h, cudaFuncSetAttribute(field_solve_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
    int per_sm = 0;
    NSFS_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, field_solve_bucket_kernel, B_THREADS, smem));
    if (per_sm < 1) return set_cuda_error(h, cudaErrorLaunchOutOfResources, "field_solve_bucket_kernel occupancy");
    if (h->opt_field_groups > 0 && per_sm > h->opt_field_groups) per_sm = (int)h->opt_field_groups;   // blocks per SM (tuning)
    int grid = per_sm * h->sm_count;
    const int n_tiles = f.tiles_x * f.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    void* args[] = {&p};
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    NSFS_CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)field_solve_bucket_kernel, dim3(grid), dim3(B_THREADS), args, smem, h->stream));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    h->launches++;
    return NSFS_OK;
}

}  // namespace nsfs
