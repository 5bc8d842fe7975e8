// Disc-count inflation of a log-odds plane: the map producer's step right before the planner path.
//
// The reference keeps grid_inflation_ incrementally (occupancy_grid.cpp:166-213): whenever a cell's log-odds crosses
// log_odds_thresh upwards / downwards, updateInflation adds +1 / -1 to every cell of a disc mask around it (generateMask,
// occupancy_grid.cpp:73-89).  At any moment, therefore, inflation[k] = number of cells currently at or above the threshold
// whose mask covers k, and that can be computed in bulk from the log-odds plane alone.  The reference's index quirks carry over:
//   - oob() only tests the row, and with "<= 0" (occupancy_grid.cpp:220-223): row 0 is never updated and never inflated;
//   - the target is addressed by the FLAT key (ci + a) * W + (cj + b) with an unchecked column, so a disc near a column
//     edge spills into the neighbouring row; the row guard applies to ci + a, not to the row the key lands in;
//   - a flat key >= N makes vector::at throw (last row, column overflow): reported as NSFS_E_RANGE.
// Kernels: (1) occupancy bit plane, 4 B/cell in, 1 bit out.  (2) For target rows 2 .. H-2 every row guard is true whatever
// the column wrap does, so the count is a pure sum of FLAT shifts of the bit plane: one thread takes 32 consecutive targets,
// adds the 49 shifted 32-bit words into six bit-sliced counter planes (AND / XOR ripple adds: ~18 logic ops per cell) and
// the warp transposes the planes into coalesced int32 stores.  (3) The few remaining rows (0, 1, H-1, the words that
// straddle them, and the virtual keys past the grids that mean "the reference throws") take a per-cell path with explicit
// guards.  HBM-bound by design: 4 B in + 4 B out per cell.
#include "nsfs_internal.cuh"

namespace nsfs {

struct DiscMask {
    int R;                // rows a in [-R, R]
    int bmax[31];         // row a covers b in [-bmax[a + R], bmax[a + R]]; -1 = empty row
};

// occ bit k = log-odds[k] >= thresh, for cells the reference can ever update (row >= 1)
__global__ void __launch_bounds__(256) occ_bits_kernel(const int32_t* __restrict__ lo, int thresh, int W, long long N,
                                                       uint32_t* __restrict__ occ, long long n_words) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 4 cells
    const long long k0 = t * 4;
    uint32_t nib = 0;
    if (k0 + 3 < N) {
        const int4 a = __ldcs(reinterpret_cast<const int4*>(lo) + t);
        nib = (a.x >= thresh) | ((a.y >= thresh) << 1) | ((a.z >= thresh) << 2) | ((a.w >= thresh) << 3);
    } else {
        for (int c = 0; c < 4; ++c) if (k0 + c < N) nib |= (uint32_t)(lo[k0 + c] >= thresh) << c;
    }
    // row 0 is "out of bounds" for updateLogOdds: those cells never count
#pragma unroll
    for (int c = 0; c < 4; ++c) if (k0 + c < W) nib &= ~(1u << c);
    nib <<= 4 * (threadIdx.x & 7);
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) nib |= __shfl_xor_sync(0xffffffffu, nib, m);
    const long long w = t >> 3;
    if ((threadIdx.x & 7) == 0 && w < n_words) occ[w] = nib;
}

__device__ __forceinline__ uint32_t bits_at(const uint32_t* __restrict__ plane, long long s, long long n_words) {
    // 32 bits of the plane starting at bit s (s >= 0); words beyond the plane read as 0
    const long long w = s >> 5;
    const int sh = (int)(s & 31);
    const uint32_t lo = w < n_words ? __ldg(plane + w) : 0u;
    const uint32_t hi = w + 1 < n_words ? __ldg(plane + w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// 64 bits of the plane starting at flat bit position pos (may be negative or run past the plane: those bits read as 0)
__device__ __forceinline__ unsigned long long window64(const uint32_t* __restrict__ plane, long long pos, long long n_words) {
    if (pos <= -64) return 0ull;
    const long long p = pos < 0 ? 0 : pos;
    const long long w = p >> 5;
    const int sh = (int)(p & 31);
    const uint32_t a0 = w < n_words ? __ldg(plane + w) : 0u;
    const uint32_t a1 = w + 1 < n_words ? __ldg(plane + w + 1) : 0u;
    const uint32_t a2 = w + 2 < n_words ? __ldg(plane + w + 2) : 0u;
    const unsigned long long v = ((unsigned long long)__funnelshift_r(a1, a2, sh) << 32) | __funnelshift_r(a0, a1, sh);
    return pos < 0 ? v << (int)(-pos) : v;
}

// Fast path: words [w_lo, w_hi) of the target plane, all of whose cells lie in rows 2 .. H-2.
// NP = number of bit-sliced counter planes: a count can reach the number of cells of the mask, so 6 planes serve masks of up to
// 63 cells (the reference's radius-4 disc has 49) and 10 planes every mask build_mask accepts (radius 15: 709 cells).
template <int NP>
__global__ void __launch_bounds__(256) inflate_words_kernel(const uint32_t* __restrict__ occ, long long n_words, int W, int R,
                                                            const int* __restrict__ bmax, long long w_lo, long long w_hi,
                                                            int32_t* __restrict__ infl) {
    const int lane = threadIdx.x & 31;
    const long long wbase = w_lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x - lane);   // first word of this warp
    if (wbase >= w_hi) return;
    const long long w = wbase + lane;
    uint32_t pl[NP];                                            // bit-sliced counters: cell c of my word holds sum_p ((pl[p] >> c) & 1) << p
#pragma unroll
    for (int p = 0; p < NP; ++p) pl[p] = 0u;
    if (w < w_hi) {
        const long long k0 = w * 32;
        for (int a = -R; a <= R; ++a) {
            const int bm = __ldg(bmax + a + R);
            if (bm < 0) continue;
            const unsigned long long win = window64(occ, k0 - (long long)a * W - bm, n_words);   // sources k - a*W - b, b = bm .. -bm
            for (int t = 0; t <= 2 * bm; ++t) {
                uint32_t carry = (uint32_t)(win >> t);
#pragma unroll
                for (int p = 0; p < NP; ++p) { const uint32_t nc = pl[p] & carry; pl[p] ^= carry; carry = nc; }
            }
        }
    }
    // transpose through shuffles: in step s the warp writes the 32 cells of word wbase + s, lane = cell
#pragma unroll 4
    for (int s2 = 0; s2 < 32; ++s2) {
        int v = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) v |= (int)((__shfl_sync(0xffffffffu, pl[p], s2) >> lane) & 1u) << p;
        if (wbase + s2 < w_hi) infl[(wbase + s2) * 32 + lane] = v;
    }
}

__global__ void __launch_bounds__(256) inflate_kernel(const uint32_t* __restrict__ occ, long long n_words, int H, int W, long long N,
                                                      int R, const int* __restrict__ bmax, long long fast_lo, long long fast_hi,
                                                      int row_a, int row_b, int32_t* __restrict__ infl, int* __restrict__ flags) {
    // (the mask rows come from global memory: indexing a by-value struct with a run-time index would spill it to local memory)
    // target cell (i, j); row i == H stands for the flat keys N .. N + 31 that lie past the grids (the ones that throw)
    // blockIdx.y walks rows 0 .. row_a and row_b .. H; cells in [fast_lo, fast_hi) belong to inflate_words_kernel
    const int i = (int)blockIdx.y <= row_a ? (int)blockIdx.y : row_b + ((int)blockIdx.y - row_a - 1);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W || (i == H && j >= 32)) return;
    const long long k = (long long)i * W + j;
    if (k >= fast_lo && k < fast_hi) return;
    int count = 0;
    for (int a = -R; a <= R; ++a) {
        const int bm = __ldg(bmax + a + R);
        if (bm < 0) continue;
        // the sources of mask row a are the flat range [k - a*W - bm, k - a*W + bm]; it starts in source row cs0 at column c0
        const bool wraps = j < bm;                                           // the range starts in the row before
        const int cs0 = i - a - (wraps ? 1 : 0);
        const int c0 = wraps ? j - bm + W : j - bm;
        long long s0 = (long long)cs0 * W + c0;
        int width = 2 * bm + 1, first = W - c0;                              // sources that lie in row cs0
        if (s0 < 0) { width += (int)s0; first += (int)s0; s0 = 0; }
        if (width <= 0 || s0 >= N) continue;
        uint32_t bits = bits_at(occ, s0, n_words) & ((1u << width) - 1u);    // width <= 31
        if (!bits) continue;
        // the row guard is on (source row + a), whatever row the flat key lands in
        const bool ok0 = cs0 + a >= 1 && cs0 + a <= H - 1;
        const bool ok1 = cs0 + 1 + a >= 1 && cs0 + 1 + a <= H - 1;
        const uint32_t m0 = first >= 32 ? 0xffffffffu : (first <= 0 ? 0u : ((1u << first) - 1u));
        count += (ok0 ? __popc(bits & m0) : 0) + (ok1 ? __popc(bits & ~m0) : 0);
    }
    if (i < H) infl[k] = count;
    else if (count) atomicOr(flags, 1);                                      // some update would index past the grids: vector::at throws
}

// generateMask (occupancy_grid.cpp:73-89), same double arithmetic; rows of a disc are symmetric contiguous intervals
static bool build_mask(double inflation_radius, double cell_size, DiscMask& mk) {
    const double cells_in_radius = round(inflation_radius / cell_size);
    const double furthest_away = inflation_radius * inflation_radius / cell_size / cell_size;
    if (!(cells_in_radius >= 0) || cells_in_radius > 15) return false;
    mk.R = (int)cells_in_radius;
    for (int t = 0; t < 31; ++t) mk.bmax[t] = -1;
    for (double i = -cells_in_radius; i <= cells_in_radius; ++i)
        for (double j = -cells_in_radius; j <= cells_in_radius; ++j)
            if (i * i + j * j <= furthest_away) {
                int& b = mk.bmax[(int)i + mk.R];
                const int aj = (int)fabs(j);
                if (aj > b) b = aj;
            }
    return true;
}

int launch_inflate(nsfs_planner* h, const int32_t* d_lo, int occ_thresh, double inflation_radius, double cell_size, int32_t* d_infl,
                   uint32_t* d_occ_scratch, int* d_flags) {
    DiscMask mk;
    if (!build_mask(inflation_radius, cell_size, mk)) return NSFS_E_ARG;
    if (h->W < 64) return NSFS_E_ARG;                       // a mask row must not span more than two grid rows
    const long long n_words = (h->N + 31) / 32;
    const long long t4 = n_words * 8;
    NSFS_CUDA_TRY(h, cudaMemsetAsync(d_flags, 0, sizeof(int), h->stream));
    int* d_bmax = d_flags + 1;                               // 31 ints right after the flag word (the caller reserves 256 bytes)
    NSFS_CUDA_TRY(h, cudaMemcpyAsync(d_bmax, mk.bmax, sizeof(mk.bmax), cudaMemcpyHostToDevice, h->stream));
    NSFS_CUDA_TRY(h, cudaStreamSynchronize(h->stream));      // mk lives on this stack frame
    occ_bits_kernel<<<(unsigned)((t4 + 255) / 256), 256, 0, h->stream>>>(d_lo, occ_thresh, h->W, h->N, d_occ_scratch, n_words);
    // fast words: every cell in rows 2 .. H-2
    long long w_lo = (2ll * h->W + 31) / 32, w_hi = ((long long)(h->H - 1) * h->W) / 32;
    if (w_hi < w_lo) w_hi = w_lo;
    const long long fast_lo = w_lo * 32, fast_hi = w_hi * 32;
    int mask_cells = 0;
    for (int t = 0; t < 31; ++t) if (mk.bmax[t] >= 0) mask_cells += 2 * mk.bmax[t] + 1;
    if (w_hi > w_lo) {
        const unsigned blocks = (unsigned)((w_hi - w_lo + 255) / 256);
        if (mask_cells <= 63) inflate_words_kernel<6><<<blocks, 256, 0, h->stream>>>(d_occ_scratch, n_words, h->W, mk.R, d_bmax, w_lo, w_hi, d_infl);
        else inflate_words_kernel<10><<<blocks, 256, 0, h->stream>>>(d_occ_scratch, n_words, h->W, mk.R, d_bmax, w_lo, w_hi, d_infl);
    }
    // everything else, with the guards spelled out: rows 0 .. row_a and row_b .. H (row H = the keys past the grids)
    int row_a = (int)((fast_lo + h->W - 1) / h->W), row_b = (int)(fast_hi / h->W);
    if (w_hi == w_lo || row_a >= row_b) { row_a = h->H; row_b = h->H + 1; }     // no fast range: all rows 0 .. H in one list
    if (row_a > h->H) row_a = h->H;
    const unsigned n_rows = (unsigned)(row_a + 1) + (row_b <= h->H ? (unsigned)(h->H - row_b + 1) : 0u);
    inflate_kernel<<<dim3((unsigned)((h->W + 255) / 256), n_rows), 256, 0, h->stream>>>(d_occ_scratch, n_words, h->H, h->W, h->N, mk.R, d_bmax,
                                                                                         fast_lo, fast_hi, row_a, row_b, d_infl, d_flags);
    h->launches += 3;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
