// Djikstra cost-to-go field, bucket-ordered tile solver (field_mode 3): the in-tile schedule that replaces the
// asynchronous pull relaxation of field.cu where it applies (one GPU, or row-partitioned shards driven in rounds).
//
// Same result as field.cu (the unique fp32 fixed point of g[v] = min_u g[u] + w(u, v) on the reference's quirk graph,
// djikstra.cpp:27-145 / SURVEY.md R13), different schedule.  The reference's open list is keyed on (int)g and every edge
// costs >= 1, so all cells of one integer bucket [b, b+1) can be settled together once the buckets below are done
// (delta-stepping with delta = 1 = the smallest edge weight: no light edges, one pass per bucket).  A tile visit is
// therefore a label-SETTING sweep over buckets: a cell is pulled when it is due, settled once, and tells its
// out-neighbours when THEY are due, instead of a label-correcting relaxation that re-reads 8 neighbours per cell per pass.
//
//   * 64x64 tile + 1-cell halo in shared memory, addressed by FLAT key (the reference's column wrap comes for free);
//     per position: g (fp32) and the 16-bit in-edge mask of map_pack.cu; per cell a DUE byte (0 = nothing to do,
//     1 + bucket the cell has to be looked at in) and a state byte (open / touched).
//   * no atomics on the tile's own cells ("owner computes"): thread t owns the 16 due bytes of cells 16t .. 16t+15 and
//     reads them with one 128-bit load per step; the cells due in the current bucket are compacted into the warp's
//     work list, pulled (8 shared loads, min tree), settled by a plain store, and their out-neighbours that would
//     improve get their due byte set by a plain byte store.  One named barrier per bucket among the 256 workers.
//   * edges that leave the tile lower the halo copy (shared atomicMin: several border cells may feed one halo cell) and
//     mark it dirty; a ninth "courier" warp owns ALL global-memory traffic during the visit: it drains dirty halo cells
//     (atomicMin into g, fence, key post to the target tile) and polls the tile's own key: when a neighbour has lowered
//     border cells of this tile meanwhile, the courier reads the border and hands the lower values to the workers
//     through an inbox.  The workers never wait on HBM / L2.
//   * between visits a cell is "open" iff g < gx, gx = the value it was last expanded at (a second fp32 plane), so a visit
//     knows what is left to expand from the planes alone: whoever lowers g[v] only has to post the tile key.
//
// Tile scheduling between visits is the asynchronous scheme of field.cu: tile t belongs to block t % gridDim.x, a block
// takes its pending tile with the smallest key; counts[1] = pending tiles + visits in flight ends the kernel.
#include <cstdio>
#include <cstdlib>

#include "nsfs_internal.cuh"

namespace nsfs {

namespace {

constexpr int BT = kFieldTile;
constexpr int B_SROW = BT + 8;                  // smem row stride (floats)
constexpr int B_SCOL0 = 4;                      // smem column of tile column 0 (halo column -1 at 3)
constexpr int B_SROWS = BT + 2;
constexpr int B_POS = B_SROWS * B_SROW;         // staged positions (4752)
constexpr int B_HDW = (B_POS + 31) / 32;        // words of the dirty-halo bit map
constexpr int B_MAXBAND = 125;                  // due bytes hold 1 + (bucket - B0) <= 126 (0x7f stands for "none" in the min scan)
#ifndef NSFS_BUCKET_FENCE
#define NSFS_BUCKET_FENCE 0                     // 1: a CTA fence between clearing due bytes and pulling (A/B check of the ordering argument)
#endif
constexpr int B_WORKERS = 256;
constexpr int B_WARPS = B_WORKERS / 32;
constexpr int B_THREADS = B_WORKERS + 32;       // + the courier warp
constexpr int B_CREDIT = 16;                    // a visit in flight counts 1 + B_CREDIT: covers key posts not yet accounted
constexpr int B_OUTCAP = 320;
constexpr int kIn = 1, kStop = 2, kEnd = 4;
static_assert(B_WORKERS * 16 == BT * BT, "one worker per 16 due bytes");

struct __align__(16) BucketSmem {
    float g[B_POS];
    uint32_t mask[BT * BT];  // in-edge mask | out-edge mask << 16 (map_pack.cu), own cells only
    uint8_t due[BT * BT];    // 0 = idle, else 1 + (bucket - B0) the cell is due in (<= 127: the scan compares 4 bytes per instruction)
    uint8_t st[BT * BT];     // bit 0: open (lowered / re-opened, not expanded at its value); bit 1: touched in this visit (write back)
    uint16_t wl[B_WARPS][512];
    uint16_t wcnt[2][B_WARPS];
    int wmin[B_WARPS];
    uint32_t hdirty[B_HDW + 3];
    uint32_t inbox_val[256];
    uint32_t post_min[16];
    int post_tile[16];
    int dec[4];
    int ctl;                 // kIn: the inbox holds lower border values; kStop: end the visit now; kEnd: nothing more to merge
    int want_end;
    int in_cnt;
    int out_cnt;
    uint32_t red_min;
    uint32_t key;
    unsigned long long best;
    uint16_t inbox_cell[256];
    uint16_t out_list[B_OUTCAP];
};

__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(B_WORKERS) : "memory"); }

__device__ __forceinline__ int cell_of(int ly, int lx) { return ly * BT + lx; }
__device__ __forceinline__ int cell_row(int c) { return c >> 6; }
__device__ __forceinline__ int cell_pos(int c) { return c + ((c >> 6) << 3) + (B_SROW + B_SCOL0); }   // (ly + 1) * 72 + 4 + lx

// Cell c is due: pull it; if its value belongs to bucket <= b settle it and tell the out-neighbours it improves when they
// are due; otherwise put it back with the bucket it belongs to.  The caller cleared due[c] BEFORE (clear, then read).
// All shared loads are issued up front (one round trip); the 8 neighbour positions serve both as sources of the pull and as
// targets of the notifications (target d is the source of in-edge d ^ 4).  INTERIOR: all 8 neighbours are cells of this tile.
template <bool INTERIOR>
__device__ __forceinline__ void bk_pull(BucketSmem& S, int c, int b, int B0, int smax, int rh, int rw, uint32_t& rej) {
    const int pos = cell_pos(c);
    const float* gp = S.g + pos;
    const uint8_t* dp = S.due + c;
    float ng[8];          // g of the neighbour at pos + OFF[d]
    uint32_t nd[8];       // its due byte (own cells only)
    uint32_t mine = 0xffu;
    if (!INTERIOR) {
        const int ly = c >> 6, lx = c & 63;
        mine = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) mine |= (uint32_t)((unsigned)(ly + nb_di(d)) < (unsigned)rh && (unsigned)(lx + nb_dj(d)) < (unsigned)rw) << d;
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        ng[d] = gp[nb_di(d) * B_SROW + nb_dj(d)];
        nd[d] = (INTERIOR || ((mine >> d) & 1u)) ? dp[nb_di(d) * BT + nb_dj(d)] : 0u;
    }
    const uint32_t mm = S.mask[c];
    const float gv = gp[0];
    const uint32_t stv = S.st[c];
    float cd[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        // in-edge d arrives from u = v - OFF[d] = v + OFF[d ^ 4]
        const float w = ((mm >> d) & 1u) ? (((mm >> (8 + d)) & 1u) ? kSqrt2F : 1.0f) : __uint_as_float(kNoKey);
        cd[d] = ng[d ^ 4] + w;
    }
    const float best = fminf(fminf(fminf(cd[0], cd[1]), fminf(cd[2], cd[3])), fminf(fminf(cd[4], cd[5]), fminf(cd[6], cd[7])));
    const float nv = fminf(gv, best);
    if (!(nv < gv) && !(stv & 1u)) return;                                // a stale due byte: settled and expanded already
    const int bk = (int)nv - B0;
    if (bk > b) {                                                         // not yet: its bucket comes later
        if (bk <= smax) S.due[c] = (uint8_t)(1 + bk);
        else if (nv < gv) { S.g[pos] = nv; S.st[c] = 3; }                 // beyond the band: keep the value, the cell stays open
        return;
    }
    S.g[pos] = nv;
    asm volatile("" ::: "memory");                                        // the value before the due bytes that announce it
    const uint32_t om = mm >> 16;
    bool keep_open = false;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        if (!((om >> d) & 1u)) continue;                                  // no edge v -> w (w blocked, row guard, shard window)
        const float cnd = nv + ((d != 0 && ((om >> (8 + d)) & 1u)) ? kSqrt2F : 1.0f);
        if (!(ng[d] > cnd)) continue;
        if (INTERIOR || ((mine >> d) & 1u)) {
            const uint32_t e = (uint32_t)(1 + max((int)cnd - B0, 0));
            if (e <= (uint32_t)(1 + smax)) {
                // nd[d] was read BEFORE my value was stored, so w's owner may have consumed it meanwhile without seeing my value:
                // always store (the smaller of the two: too early only costs a look, too late would settle w out of order)
                S.due[c + nb_di(d) * BT + nb_dj(d)] = (uint8_t)((nd[d] != 0u && nd[d] < e) ? nd[d] : e);
            } else {
                keep_open = true;                                         // beyond the band: this cell is expanded again next visit,
                rej = min(rej, __float_as_uint(cnd));                     // which the tile asks for at the value it could not place
            }
        } else {
            const int wp = pos + nb_di(d) * B_SROW + nb_dj(d);
            const int cb = __float_as_int(cnd);
            const int old = atomicMin(reinterpret_cast<int*>(S.g) + wp, cb);
            if (old > cb) atomicOr(&S.hdirty[wp >> 5], 1u << (wp & 31));
        }
    }
    S.st[c] = keep_open ? 3 : 2;
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
            cells[q] = cell_of(ly, lx);
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
    const float hi_f = (float)(B0 + smax + 1);
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
            if (pkf < hi_f) {                                             // within the band (values below B0 are simply overdue)
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
// SWAR over the 4 due bytes of a word (all <= 127): bit 7 of a byte is set iff it is non-zero and <= the limit byte.
__device__ __forceinline__ uint32_t due_le(uint32_t w, uint32_t lim4) {
    return ((lim4 | 0x80808080u) - w) & (w + 0x7f7f7f7fu) & 0x80808080u;
}

__device__ __forceinline__ void worker_loop(const FieldParams& p, BucketSmem& S, int B0, int smax, int rh, int rw, int tid, uint32_t& rej) {
    const int lane = tid & 31, warp = tid >> 5;
    int b = 0;                    // staged open cells carry due byte 1: the first round files every one of them under its bucket
    int it = 0, steps = 0;
    long long pa = 0, pb = 0, pc = 0, npull = 0;
    for (;;) {
        const long long q0 = p.prof ? clock64() : 0;
        // ---- my 16 due bytes: which cells are due now ----
        const uint4 f4 = reinterpret_cast<const uint4*>(S.due)[tid];      // re-read every round (the barrier is a memory clobber)
        uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        if (f4.x | f4.y | f4.z | f4.w) {
            const uint32_t lim4 = (uint32_t)min(b + 1, 127) * 0x01010101u;
            m0 = due_le(f4.x, lim4); m1 = due_le(f4.y, lim4); m2 = due_le(f4.z, lim4); m3 = due_le(f4.w, lim4);
        }
        const int ndue = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
        int total = 0;
        if (__any_sync(0xffffffffu, ndue)) {
            int incl = ndue;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            total = __shfl_sync(0xffffffffu, incl, 31);
            int at = incl - ndue;
            uint32_t mw[4] = {m0, m1, m2, m3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t m = mw[q];
                while (m) {
                    const int bit = __ffs(m) - 1;                         // 7, 15, 23, 31
                    m &= m - 1;
                    const int c = tid * 16 + q * 4 + (bit >> 3);
                    S.wl[warp][at++] = (uint16_t)c;
                    S.due[c] = 0;                                         // clear first, then read the neighbours (shared memory
                }                                                         // executes one warp's accesses in order)
            }
            if (NSFS_BUCKET_FENCE || p.early == 2) __threadfence_block();  // (field_early 2: A/B check of the ordering argument)
            __syncwarp();
        }
        const long long q1 = p.prof ? clock64() : 0;
        npull += total;
        for (int e = lane; e < total; e += 32) {
            const int c = S.wl[warp][e];
            const int ly = c >> 6, lx = c & 63;
            if ((unsigned)(ly - 1) < (unsigned)(rh - 2) && (unsigned)(lx - 1) < (unsigned)(rw - 2)) bk_pull<true>(S, c, b, B0, smax, rh, rw, rej);
            else bk_pull<false>(S, c, b, B0, smax, rh, rw, rej);
        }
        if (lane == 0) *(volatile uint16_t*)&S.wcnt[it & 1][warp] = (uint16_t)total;
        if (tid == 0) *(volatile int*)&S.dec[it & 1] = *(volatile int*)&S.ctl;
        const long long q2 = p.prof ? clock64() : 0;
        bar_workers();
        int f = *(volatile int*)&S.dec[it & 1];
        const uint4 cn = *reinterpret_cast<const uint4*>(&S.wcnt[it & 1][0]);
        const bool tot = (cn.x | cn.y | cn.z | cn.w) != 0u;
        ++it;
        if (p.prof) { const long long q3 = clock64(); pa += q1 - q0; pb += q2 - q1; pc += q3 - q2; }
        if (f & kStop) break;
        int b_next = b + 1;
        if (tot) ++steps;
        if (!tot || b_next > smax) {
            // nothing was due, or the band is used up: find the next bucket that holds something (gaps, the start, the end of
            // the visit).  Cells settled out of order (overdue corrections) may have left due bytes BELOW b: they count too.
            const uint4 e4 = reinterpret_cast<const uint4*>(S.due)[tid];
            uint32_t mn4 = 0x7f7f7f7fu;
            const uint32_t ew[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t z7 = ((ew[q] + 0x7f7f7f7fu) & 0x80808080u) ^ 0x80808080u;   // bit 7 of the zero bytes
                mn4 = __vminu4(mn4, ew[q] | (z7 - (z7 >> 7)));                               // zero bytes count as 0x7f
            }
            int mymin = (int)min(min(mn4 & 0xffu, (mn4 >> 8) & 0xffu), min((mn4 >> 16) & 0xffu, mn4 >> 24));
            mymin = __reduce_min_sync(0xffffffffu, mymin);
            if (lane == 0) *(volatile int*)&S.wmin[warp] = mymin;
            bar_workers();
            const int4 w0 = *reinterpret_cast<const int4*>(&S.wmin[0]), w1 = *reinterpret_cast<const int4*>(&S.wmin[4]);
            const int mn = min(min(min(w0.x, w0.y), min(w0.z, w0.w)), min(min(w1.x, w1.y), min(w1.z, w1.w)));
            b_next = mn < 0x7f ? mn - 1 : 255;
            bar_workers();                                                // wmin is free again
        }
        if (!(f & kIn)) {
            if (b_next <= smax) { b = b_next; continue; }
            // out of work: ask the courier whether anything arrived for this tile meanwhile
            if (tid == 0) {
                atomicExch(&S.want_end, 1);
                int ff = 0, spins = 0;
                const long long t_w = clock64();
                while (!(ff = *(volatile int*)&S.ctl)) {
                    __nanosleep(40);
                    if (++spins > (1 << 24)) { atomicOr(p.counts + 4, 2); atomicOr(&S.ctl, kStop); }
                }
                *(volatile int*)&S.dec[2] = ff;
                atomicAdd(p.counts + 15, (int)((clock64() - t_w) >> 4));
            }
            bar_workers();
            f = *(volatile int*)&S.dec[2];
            if (!(f & kIn)) break;                                        // kEnd or kStop
        }
        // lower border values from the courier: take them over, the cells are open and due in their own buckets
        const int n = *(volatile int*)&S.in_cnt;
        int amin = 255;
        for (int e = tid; e < n; e += B_WORKERS) {
            const int c = S.inbox_cell[e];
            const uint32_t vb = S.inbox_val[e];
            const int pos = cell_pos(c);
            if (vb < __float_as_uint(S.g[pos])) {
                S.g[pos] = __uint_as_float(vb);
                S.st[c] = 3;
                const int bk = max((int)__uint_as_float(vb) - B0, 0);
                if (bk <= smax) { S.due[c] = (uint8_t)(1 + bk); amin = min(amin, bk); }
            }
        }
        amin = __reduce_min_sync(0xffffffffu, amin);
        if (lane == 0) *(volatile int*)&S.wmin[warp] = amin;
        bar_workers();
        {
            const int4 w0 = *reinterpret_cast<const int4*>(&S.wmin[0]), w1 = *reinterpret_cast<const int4*>(&S.wmin[4]);
            amin = min(min(min(w0.x, w0.y), min(w0.z, w0.w)), min(min(w1.x, w1.y), min(w1.z, w1.w)));
        }
        if (tid == 0) {
            *(volatile int*)&S.in_cnt = 0;
            *(volatile int*)&S.want_end = 0;
            __threadfence_block();
            atomicAnd(&S.ctl, ~kIn);
        }
        b = min(b_next, amin);
        bar_workers();                                                    // everybody has read the minima before they are overwritten
    }
    if (tid == 0) { atomicAdd(p.counts + 8, steps); atomicAdd(p.counts + 14, it); }
    if (lane == 0 && p.prof) {
        atomicAdd(p.prof + 0 + warp * 4, (unsigned long long)pa); atomicAdd(p.prof + 1 + warp * 4, (unsigned long long)pb);
        atomicAdd(p.prof + 2 + warp * 4, (unsigned long long)pc); atomicAdd(p.prof + 3 + warp * 4, (unsigned long long)npull);
    }
}

// One visit of one tile: stage, settle bucket by bucket, write back.
__device__ void bucket_visit(const FieldParams& p, BucketSmem& S, int tile) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int r0 = ty * BT, c0 = tx * BT;
    const int W = p.W;
    const long long N = p.N;
    const int rh = min(BT, p.H - r0), rw = min(BT, W - c0);

    const long long t_a = clock64();
    for (int i = tid; i < B_HDW + 3; i += B_THREADS) S.hdirty[i] = 0;
    if (tid < 16) S.post_min[tid] = kNoKey;
    if (tid == 0) { S.ctl = 0; S.want_end = 0; S.in_cnt = 0; S.out_cnt = 0; S.red_min = kNoKey; }
    __syncthreads();

    // ---- stage: position (rr, cc) holds g of flat key (r0-1+rr)*W + (c0-1+cc); my cells also get their masks and state ----
    const bool win = p.win_lo > 0 || p.win_hi < N;
    uint32_t mn = kNoKey;
    constexpr int NPOS = B_SROWS * (BT + 2);
    for (int base = 0; base < NPOS; base += 8 * B_THREADS) {
        // eight positions per thread with all their loads in flight together (one L2 round trip per batch, not per position)
        float gvv[8], gxv[8];
        uint32_t mkv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int idx = base + k * B_THREADS + tid;
            const int rr = idx / (BT + 2), cc = idx - rr * (BT + 2);
            const int ly = rr - 1, lx = cc - 1;
            const long long key = (long long)(r0 + ly) * W + (c0 + lx);
            gvv[k] = kInfF; gxv[k] = 0.f; mkv[k] = 0;
            if (idx < NPOS && key >= 0 && key < N) {
                gvv[k] = __ldcg(p.g + key);                               // through L2: other SMs write g
                if (ly >= 0 && ly < rh && lx >= 0 && lx < rw) {
                    gxv[k] = __ldcg(p.gx + key);
                    uint32_t om = __ldg(p.out_mask + key);
                    if (win) {                                            // shards: a target outside the window takes no in-edge
#pragma unroll
                        for (int d = 0; d < 8; ++d) {
                            const long long kv = key + nb_off(d, W);
                            if (kv < p.win_lo || kv >= p.win_hi) om &= ~(1u << d);
                        }
                    }
                    mkv[k] = (uint32_t)__ldg(p.in_mask + key) | (om << 16);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int idx = base + k * B_THREADS + tid;
            if (idx >= NPOS) continue;
            const int rr = idx / (BT + 2), cc = idx - rr * (BT + 2);
            const int ly = rr - 1, lx = cc - 1;
            const bool in_sq = (unsigned)ly < (unsigned)BT && (unsigned)lx < (unsigned)BT;
            const bool mine = ly >= 0 && ly < rh && lx >= 0 && lx < rw;
            S.g[rr * B_SROW + (B_SCOL0 - 1) + cc] = gvv[k];
            if (in_sq) {
                const bool open = mine && gvv[k] < gxv[k];
                S.mask[cell_of(ly, lx)] = mkv[k];
                S.st[cell_of(ly, lx)] = open ? 3 : 0;
                S.due[cell_of(ly, lx)] = open ? 1 : 0;                    // looked at in the first round, which files it under its bucket
                if (open) mn = min(mn, __float_as_uint(gvv[k]));
            }
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
        smax = min(B_MAXBAND, (int)p.delta);
        if (p.cap < kNoKey - 1u) smax = min(smax, (int)__uint_as_float(p.cap) - B0);
    }
    __syncthreads();
    if (tid == 0) S.red_min = kNoKey;
    uint32_t rej = kNoKey;        // the smallest value this visit could not place (beyond the band / the cap): the tile's next key
    const long long t_b = clock64();
    if (work) {
        if (tid < B_WORKERS) worker_loop(p, S, B0, smax, rh, rw, tid, rej);
        else courier_loop(p, S, tile, ty, tx, r0, c0, rh, rw, B0, smax, lane);
    }
    __syncthreads();
    const long long t_c = clock64();

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
                    const int pos = cell_pos(c);
                    if (S.inbox_val[e] < __float_as_uint(S.g[pos])) { S.g[pos] = __uint_as_float(S.inbox_val[e]); S.st[c] = 3; }
                }
            }
            bar_workers();
            const bool stopped = *(volatile int*)&S.ctl & kStop;
            for (int c = tid; c < BT * BT; c += B_WORKERS) {
                uint32_t st = S.st[c];
                const int ly = cell_row(c), lx = c & 63;
                const long long key = (long long)(r0 + ly) * W + (c0 + lx);
                const float gv = S.g[cell_pos(c)];
                // a due byte left over (the visit was stopped, or its bucket lies beyond the band): one of its in-neighbours holds a
                // value this cell has not taken yet; take it now so the cell itself is open and the tile's key covers it
                const bool left = S.due[c] != 0;
                if (left) {
                    const uint32_t m = S.mask[c];
                    float best = gv;
#pragma unroll
                    for (int d = 0; d < 8; ++d)
                        if ((m >> d) & 1u) best = fminf(best, S.g[cell_pos(c) - nb_di(d) * B_SROW - nb_dj(d)] + (((m >> (8 + d)) & 1u) ? kSqrt2F : 1.0f));
                    if (best < gv) S.g[cell_pos(c)] = best;
                    st = 3;
                }
                if (!(st & 2u)) continue;
                const float gw = S.g[cell_pos(c)];
                // border cells are lowered by the neighbours' couriers too: never raise them
                if (ly == 0 || ly == rh - 1 || lx == 0 || lx == rw - 1) atomicMin(reinterpret_cast<uint32_t*>(p.g) + key, __float_as_uint(gw));
                else __stcg(p.g + key, gw);
                // still open: beyond the band the tile wants another visit at that value; within it (a cell some of whose out-edges
                // reach beyond the band) the rejected candidates were collected while expanding
                if (st & 1u) { if ((int)gw - B0 > smax || stopped || left) rej = min(rej, __float_as_uint(gw)); }
                else __stcg(p.gx + key, gw);
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
        const long long t_d = clock64();
        atomicAdd(p.counts + 11, (int)((t_b - t_a) >> 4));    // cycles / 16: stage, bucket loop, write-back
        atomicAdd(p.counts + 12, (int)((t_c - t_b) >> 4));
        atomicAdd(p.counts + 13, (int)((t_d - t_c) >> 4));
    }
    __syncthreads();
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
        bucket_visit(p, S, tile);
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
    p.delta = h->opt_field_delta > 0 ? (float)(h->opt_field_delta < B_MAXBAND ? h->opt_field_delta : B_MAXBAND) : 128.0f;
    const size_t smem = sizeof(BucketSmem);
    NSFS_CUDA_TRY(h, cudaFuncSetAttribute(field_solve_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
    int per_sm = 0;
    NSFS_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, field_solve_bucket_kernel, B_THREADS, smem));
    if (per_sm < 1) return set_cuda_error(h, cudaErrorLaunchOutOfResources, "field_solve_bucket_kernel occupancy");
    if (h->opt_field_groups > 0 && per_sm > h->opt_field_groups) per_sm = (int)h->opt_field_groups;   // blocks per SM (tuning)
    int grid = per_sm * h->sm_count;
    const int n_tiles = f.tiles_x * f.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    static unsigned long long* d_prof = nullptr;
    const bool prof = getenv("NSFS_FIELD_DEBUG") != nullptr;
    if (prof) {
        if (!d_prof) cudaMalloc(&d_prof, 64 * 8);
        cudaMemsetAsync(d_prof, 0, 64 * 8, h->stream);
        p.prof = d_prof;
    }
    void* args[] = {&p};
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    NSFS_CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)field_solve_bucket_kernel, dim3(grid), dim3(B_THREADS), args, smem, h->stream));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    h->launches++;
    if (prof) {
        unsigned long long hp[64];
        cudaMemcpyAsync(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        for (int w = 0; w < 8; ++w)
            fprintf(stderr, "[bucket prof] warp %d: scan %llu pull %llu barrier %llu cycles, %llu cells pulled\n", w, hp[4 * w], hp[4 * w + 1], hp[4 * w + 2], hp[4 * w + 3]);
    }
    return NSFS_OK;
}

}  // namespace nsfs
