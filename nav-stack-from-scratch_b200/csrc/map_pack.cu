// Map ingestion: int32 inflation / log-odds planes -> bit planes + per-cell edge masks.
//
// Reference semantics restated here (never its code):
//   free(k)  := infl[k] <= 0 && lo[k] <= lo_thresh                   testPos, grid_planner_core.cpp:434-455
//   oob      := only the ROW is tested                                oob,     grid_planner_core.cpp:423-426
//   neighbours are addressed by the flat key k + DI*W + DJ            flatten, grid_planner_core.cpp:404-407
//   cost of direction d = 1 if an even number of passable directions precede it, else sqrt(2)
//                                                                     astar.cpp:88,97-101,107,123
#include "nsfs_internal.cuh"

namespace nsfs {

// ---- kernel 1: pack -------------------------------------------------------------------------------
// Each thread reads 4 consecutive cells of both planes with one 128-bit load each (fully coalesced:
// a warp covers 128 cells = 512 B per plane), forms a 4-bit nibble per plane, and 8 neighbouring lanes
// OR their nibbles into one 32-bit word with three xor-shuffles.  HBM-bound: 8 B/cell in, 0.25 B/cell out.
__global__ void __launch_bounds__(256) pack_map_kernel(const int32_t* __restrict__ infl, const int32_t* __restrict__ lo,
                                                       int lo_thresh, long long N, uint32_t* __restrict__ free_bits,
                                                       uint32_t* __restrict__ infl_bits, long long n_words, long long word0) {
    const long long t = word0 * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 4 cells (word0: partial re-pack)
    const long long k0 = t * 4;
    uint32_t fn = 0, in_ = 0;
    if (k0 + 3 < N) {
        const int4 a = __ldcs(reinterpret_cast<const int4*>(infl) + t);     // streamed once: evict-first
        const int4 b = __ldcs(reinterpret_cast<const int4*>(lo) + t);
        in_ = (a.x > 0) | ((a.y > 0) << 1) | ((a.z > 0) << 2) | ((a.w > 0) << 3);
        const uint32_t occ = (b.x > lo_thresh) | ((b.y > lo_thresh) << 1) | ((b.z > lo_thresh) << 2) | ((b.w > lo_thresh) << 3);
        fn = ~(in_ | occ) & 0xFu;
    } else {
        for (int c = 0; c < 4; ++c) {
            const long long k = k0 + c;
            if (k < N) {
                const bool inf = infl[k] > 0;
                const bool occ = lo[k] > lo_thresh;
                in_ |= (uint32_t)inf << c;
                fn |= (uint32_t)(!inf && !occ) << c;
            }
        }
    }
    const int sh = 4 * (threadIdx.x & 7);
    fn <<= sh;
    in_ <<= sh;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
        fn |= __shfl_xor_sync(0xffffffffu, fn, m);
        in_ |= __shfl_xor_sync(0xffffffffu, in_, m);
    }
    const long long w = t >> 3;
    if ((threadIdx.x & 7) == 0 && w < n_words) {
        free_bits[w] = fn;
        infl_bits[w] = in_;
    }
}

// ---- kernel 2: out-edge masks ----------------------------------------------------------------------
// textbook = 1 (opt-in "graph" option): the grid a textbook planner would search instead of the reference's
// quirk graph: rows AND columns are bounds-checked (no flat-key wrap, nothing throws) and a direction costs sqrt(2) iff it is a
// diagonal (odd d in NB_LUT order), whatever the neighbourhood.  Every kernel downstream reads only these masks.
__global__ void __launch_bounds__(256) out_mask_kernel(const uint32_t* __restrict__ free_bits, int H, int W, long long N,
                                                       uint16_t* __restrict__ out_mask, int row0, int textbook) {
    // 2-D launch: blockIdx.y = row (no 64-bit division per cell), x = columns
    const int i = row0 + (int)blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    const long long u = (long long)i * W + j;
    uint32_t pass = 0, diag = 0, thr = 0, card = 1;
    if (textbook) {
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const int ni = i + nb_di(d), nj = j + nb_dj(d);
            if (ni < 0 || ni >= H || nj < 0 || nj >= W) continue;
            if (!bit_at(free_bits, (long long)ni * W + nj)) continue;
            pass |= 1u << d;
            if (d & 1) diag |= 1u << (8 + d);
        }
        out_mask[u] = (uint16_t)(pass | diag);
        return;
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        const int ni = i + nb_di(d);
        if (ni < 0 || ni >= H) continue;                       // row guard; toggle untouched
        const long long kv = u + nb_off(d, W);
        if (kv < 0 || kv >= N) { thr = kThrowBit; continue; }  // vector::at would throw here
        if (!bit_at(free_bits, kv)) continue;                  // blocked; toggle untouched
        pass |= 1u << d;
        if (!card) diag |= 1u << (8 + d);
        card ^= 1u;
    }
    out_mask[u] = (uint16_t)(pass | diag | thr);
}

// ---- kernel 3: in-edge masks (pull form used by the field relaxation) --------------------------------
__global__ void __launch_bounds__(256) in_mask_kernel(const uint32_t* __restrict__ free_bits,
                                                      const uint16_t* __restrict__ out_mask, int H, int W, long long N,
                                                      long long active_lo, long long active_hi, uint16_t* __restrict__ in_mask,
                                                      long long k0, long long k1) {
    const long long v = k0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= k1) return;
    uint32_t m = 0;
    if (bit_at(free_bits, v) && v >= active_lo && v < active_hi) {    // rows outside the window take no in-edges (shards)
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const long long u = v - nb_off(d, W);
            if (u < 0 || u >= N) continue;
            if (!bit_at(free_bits, u)) continue;               // only free cells are ever expanded (the source is seeded apart)
            const uint32_t om = __ldg(out_mask + u);
            if (!((om >> d) & 1u)) continue;                   // row guard of u failed for d
            m |= 1u << d;
            m |= ((om >> (8 + d)) & 1u & (uint32_t)(d != 0)) << (8 + d);
        }
    }
    in_mask[v] = (uint16_t)m;
}

// ---- testPos for a batch of cells --------------------------------------------------------------------
__global__ void test_pos_kernel(const uint32_t* __restrict__ free_bits, int H, int W, long long N, int n,
                                const int32_t* __restrict__ ij, int8_t* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = ij[2 * t], j = ij[2 * t + 1];
    const long long key = (long long)i * W + j;
    int8_t r;
    if (i < 0 || i >= H) r = 0;
    else if (key < 0 || key >= N) r = -1;
    else r = bit_at(free_bits, key) ? 1 : 0;
    out[t] = r;
}

// ---- the path consumer's safety check (reference commander.cpp:316-402) -------------------------
// Commander::checkCell per trajectory cell with the commander's OWN oob (row 0 counts as outside, the column is never
// tested, so the key i*W + j may alias or leave the planes, where vector::at throws), and the scan order of
// checkTrajectorySafety: points t_id, t_id-1, ..., 0; the first one that is not safe decides.  One lane per point,
// warps agree with __ballot_sync, blocks with one atomicMax on (index << 1 | throws): the largest index wins.
__global__ void __launch_bounds__(256) trajectory_check_kernel(const uint32_t* __restrict__ free_bits, int H, int W, long long N, int n,
                                                               const int32_t* __restrict__ ij, int t_id, int* __restrict__ first_bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int cls = 1;                                           // 1 safe, 0 not safe, -1 throws
    if (t <= t_id) {
        if (t >= n) cls = -1;                              // trajectory_.at(t) past the end
        else {
            const int i = ij[2 * t], j = ij[2 * t + 1];
            const long long key = (long long)i * W + j;
            if (i <= 0 || i >= H) cls = 0;
            else if (key < 0 || key >= N) cls = -1;
            else cls = bit_at(free_bits, key) ? 1 : 0;
        }
    }
    const uint32_t bad = __ballot_sync(0xffffffffu, cls != 1);
    if (bad) {
        const int top = 31 - __clz(bad);                   // the highest index in this warp comes first in the reference's scan
        if ((threadIdx.x & 31) == top) atomicMax(first_bad, (t << 1) | (cls < 0));
    }
}

int launch_trajectory_check(nsfs_planner* h, int n, const int32_t* d_ij, int t_id, int* d_first_bad) {
    const int pts = t_id + 1;
    if (pts <= 0) return NSFS_OK;
    trajectory_check_kernel<<<(pts + 255) / 256, 256, 0, h->stream>>>(h->d_free, h->H, h->W, h->N, n, d_ij, t_id, d_first_bad);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

// Packs the words that hold cells [k0, k1) (whole map: 0, N).
static int pack_range(nsfs_planner* h, const int32_t* d_infl, const int32_t* d_lo, long long k0, long long k1) {
    const long long n_words = (h->N + 31) / 32;
    const long long w0 = k0 / 32, w1 = (k1 + 31) / 32;
    const long long threads = (w1 - w0) * 8;
    if (threads <= 0) return NSFS_OK;
    const int block = 256;
    const long long grid = (threads + block - 1) / block;
    pack_map_kernel<<<(unsigned)grid, block, 0, h->stream>>>(d_infl, d_lo, h->lo_thresh, h->N, h->d_free, h->d_inflated, w1 < n_words ? w1 : n_words, w0);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

int launch_pack_map(nsfs_planner* h, const int32_t* d_infl, const int32_t* d_lo) { return pack_range(h, d_infl, d_lo, 0, h->N); }

// Edge masks of cells [k0, k1) (out) and [j0, j1) (in).
static int mask_range(nsfs_planner* h, long long k0, long long k1, long long j0, long long j1) {
    const int block = 256;
    const long long a_lo = (h->opt_active_row_lo > 0 ? h->opt_active_row_lo : 0) * (long long)h->W;
    const long long a_hi = h->opt_active_row_hi > 0 ? h->opt_active_row_hi * (long long)h->W : h->N;
    if (k1 > k0)       // whole rows (callers pass row-aligned ranges)
        out_mask_kernel<<<dim3((unsigned)((h->W + block - 1) / block), (unsigned)((k1 - k0) / h->W)), block, 0, h->stream>>>(
            h->d_free, h->H, h->W, h->N, h->d_out_mask, (int)(k0 / h->W), h->opt_graph == 1 ? 1 : 0);
    if (j1 > j0) in_mask_kernel<<<(unsigned)((j1 - j0 + block - 1) / block), block, 0, h->stream>>>(h->d_free, h->d_out_mask, h->H, h->W, h->N, a_lo, a_hi, h->d_in_mask, j0, j1);
    h->launches += 2;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

int launch_build_masks(nsfs_planner* h) { return mask_range(h, 0, h->N, 0, h->N); }

// Rows [row0, row0 + nrows) of the int32 planes changed (already in d_infl / d_lo): re-pack their words and rebuild the
// masks that can see them.  A cell's out-edges depend on free bits up to 2 rows away (flat-key wrap at the column edges),
// a cell's in-edges on out-masks up to 2 rows away: out-masks of rows +-2, in-masks of rows +-4.
int launch_update_rows(nsfs_planner* h, int row0, int nrows) {
    const long long W = h->W;
    int rc = pack_range(h, h->d_infl, h->d_lo, row0 * W, (row0 + nrows) * W);
    if (rc != NSFS_OK) return rc;
    auto clampk = [&](long long r) { return (r < 0 ? 0 : (r > h->H ? h->H : r)) * W; };
    return mask_range(h, clampk(row0 - 2), clampk(row0 + nrows + 2), clampk(row0 - 4), clampk(row0 + nrows + 4));
}

int launch_test_pos(nsfs_planner* h, int n, const int32_t* d_ij, int8_t* d_out) {
    if (n <= 0) return NSFS_OK;
    test_pos_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_free, h->H, h->W, h->N, n, d_ij, d_out);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
