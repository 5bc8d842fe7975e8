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

#include "nsfs_internal.cuh"

namespace cg = cooperative_groups;

namespace nsfs {

constexpr int TS = 64;                 // tile side in cells
constexpr int SBW = 8, SBH = 4;        // sub-block: 8 columns x 4 rows = one lane per cell
constexpr int SB_X = TS / SBW;         // 8
constexpr int SB_Y = TS / SBH;         // 16
constexpr int NSB = SB_X * SB_Y;       // 128 sub-blocks per tile
constexpr int FIELD_THREADS = 1024;    // 32 warps, warp w owns sub-blocks w, w+32, w+64, w+96
constexpr int SB_PER_WARP = NSB / (FIELD_THREADS / 32);
constexpr int SROW = 72;               // smem row stride in floats; 72 % 32 == 8 => the 4 rows of a sub-block use disjoint banks
constexpr int SCOL0 = 4;               // smem column of tile column 0 (halo column -1 at 3)
constexpr int SROWS = TS + 2;
constexpr int kMaxSweeps = 4096;       // safety net only; a tile converges in O(TS) sweeps

struct FieldParams {
    float* g;
    const uint16_t* in_mask;
    uint8_t* tile_flag;
    int* list0;
    int* list1;
    int* counts;          // 0,1: list lengths; 2: rounds; 3: tile visits; 6: sweeps
    int H, W;
    long long N;
    int tiles_x, tiles_y;
    int max_rounds;
};

__global__ void field_fill_kernel(float4* __restrict__ g4, long long n4, float* __restrict__ g, long long N) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4 v = make_float4(kInfF, kInfF, kInfF, kInfF);
    for (long long k = t; k < n4; k += stride) g4[k] = v;
    for (long long k = n4 * 4 + t; k < N; k += stride) g[k] = kInfF;
}

// One thread: g[source] = 0, first tile list; a non-free source is expanded here because in_mask only
// carries edges out of free cells (plan() pushes the start cell unconditionally, astar.cpp:45-52).
__global__ void field_seed_kernel(FieldParams p, const uint32_t* __restrict__ free_bits,
                                  const uint16_t* __restrict__ out_mask, int si, int sj) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long ks = (long long)si * p.W + sj;
    p.g[ks] = 0.0f;
    int n = 0;
    auto add_tile = [&](long long k) {
        const int t = (int)(k / p.W) / TS * p.tiles_x + (int)(k % p.W) / TS;
        for (int q = 0; q < n; ++q) if (p.list0[q] == t) return;
        p.list0[n++] = t;
    };
    add_tile(ks);
    if (!bit_at(free_bits, ks)) {
        const uint32_t om = out_mask[ks];
        for (int d = 0; d < 8; ++d) {
            if (!((om >> d) & 1u)) continue;
            const long long kv = ks + nb_off(d, p.W);
            const float w = ((om >> (8 + d)) & 1u) && d ? kSqrt2F : 1.0f;
            if (p.g[kv] > w) p.g[kv] = w;
            add_tile(kv);
        }
    }
    p.counts[0] = n;
    p.counts[1] = 0;
    p.counts[2] = 0;
    p.counts[3] = 0;
    p.counts[4] = 0;
    p.counts[5] = 0;
    p.counts[6] = 0;
}

__device__ __forceinline__ void relax_tile(const FieldParams& p, int tile, float* s, uint32_t (*s_act)[4], uint32_t* s_edges,
                                           int* s_sweeps) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tile / p.tiles_x, tx = tile % p.tiles_x;
    const int r0 = ty * TS, c0 = tx * TS;
    const int W = p.W, H = p.H;
    const long long N = p.N;

    // ---- stage the tile and its halo: position (rr, cc) holds g[(r0-1+rr)*W + (c0-1+cc)] (flat key!) -------
    for (int idx = tid; idx < SROWS * (TS + 2); idx += FIELD_THREADS) {
        const int rr = idx / (TS + 2), cc = idx - rr * (TS + 2);
        const long long key = (long long)(r0 - 1 + rr) * W + (c0 - 1 + cc);
        s[rr * SROW + (SCOL0 - 1) + cc] = (key >= 0 && key < N) ? p.g[key] : kInfF;
    }
    if (tid < 12) (&s_act[0][0])[tid] = tid < 4 ? 0xffffffffu : 0u;   // sweep 0: everything active
    if (tid == 12) *s_edges = 0u;

    // ---- this thread's four cells ---------------------------------------------------------------------
    const int lr = lane >> 3, lc = lane & 7;
    uint32_t mask[SB_PER_WARP];
    long long key[SB_PER_WARP];
    int soff[SB_PER_WARP];
#pragma unroll
    for (int q = 0; q < SB_PER_WARP; ++q) {
        const int sb = warp + q * (FIELD_THREADS / 32);
        const int ly = (sb / SB_X) * SBH + lr, lx = (sb % SB_X) * SBW + lc;
        const int r = r0 + ly, c = c0 + lx;
        soff[q] = (ly + 1) * SROW + SCOL0 + lx;
        const bool real = r < H && c < W;
        key[q] = real ? (long long)r * W + c : -1;
        mask[q] = real ? (uint32_t)__ldg(p.in_mask + key[q]) : 0u;
    }
    __syncthreads();
    float v[SB_PER_WARP], v0[SB_PER_WARP];
#pragma unroll
    for (int q = 0; q < SB_PER_WARP; ++q) v[q] = v0[q] = s[soff[q]];

    // ---- sweeps: relax active sub-blocks until a whole sweep changes nothing ----------------------------
    int sweep = 0;
    for (; sweep < kMaxSweeps; ++sweep) {
        const int cur = sweep % 3, nxt = (sweep + 1) % 3, clr = (sweep + 2) % 3;
        if (tid < 4) s_act[clr][tid] = 0u;
#pragma unroll
        for (int q = 0; q < SB_PER_WARP; ++q) {
            const int sb = warp + q * (FIELD_THREADS / 32);
            if (!((s_act[cur][sb >> 5] >> (sb & 31)) & 1u)) continue;       // warp-uniform
            const float* c = s + soff[q];
            const uint32_t m = mask[q];
            float best = v[q];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                // in-edge d arrives from u = v - OFF[d], i.e. tile position (ly - DI, lx - DJ)
                const float nb = c[-nb_di(d) * SROW - nb_dj(d)];
                const float cand = nb + (((m >> (8 + d)) & 1u) ? kSqrt2F : 1.0f);
                if ((m >> d) & 1u) best = fminf(best, cand);
            }
            const bool changed = best < v[q];
            if (changed) { v[q] = best; s[soff[q]] = best; }
            const uint32_t bits = __ballot_sync(0xffffffffu, changed);
            if (bits && lane < 9) {
                // lane l decides for neighbour (dy, dx) = (l/3 - 1, l%3 - 1) whether a changed cell borders it
                const int dy = lane / 3 - 1, dx = lane % 3 - 1;
                uint32_t need = bits;
                if (dy < 0) need &= 0x000000ffu; else if (dy > 0) need &= 0xff000000u;
                if (dx < 0) need &= 0x01010101u; else if (dx > 0) need &= 0x80808080u;
                const int sy = sb / SB_X + dy, sx = sb % SB_X + dx;
                if (need && sy >= 0 && sy < SB_Y && sx >= 0 && sx < SB_X) {
                    const int t = sy * SB_X + sx;
                    atomicOr(&s_act[nxt][t >> 5], 1u << (t & 31));
                }
            }
        }
        __syncthreads();
        const uint32_t any = s_act[nxt][0] | s_act[nxt][1] | s_act[nxt][2] | s_act[nxt][3];
        if (!any) break;
    }

    // ---- write back what improved; note which tile borders moved -----------------------------------------
    uint32_t edges = 0;
#pragma unroll
    for (int q = 0; q < SB_PER_WARP; ++q) {
        if (key[q] >= 0 && v[q] < v0[q]) {
            p.g[key[q]] = v[q];
            const int sb = warp + q * (FIELD_THREADS / 32);
            const int ly = (sb / SB_X) * SBH + lr, lx = (sb % SB_X) * SBW + lc;
            const int c = c0 + lx;
            const bool top = ly == 0, bot = ly == TS - 1, lef = lx == 0, rig = lx == TS - 1;
            // bit (dy+1)*3 + (dx+1) = neighbour tile (dy, dx) reads a cell of ours that changed
            if (top) edges |= 1u << 1;
            if (bot) edges |= 1u << 7;
            if (lef) edges |= 1u << 3;
            if (rig) edges |= 1u << 5;
            if (top && lef) edges |= 1u << 0;
            if (top && rig) edges |= 1u << 2;
            if (bot && lef) edges |= 1u << 6;
            if (bot && rig) edges |= 1u << 8;
            if (c == W - 1) edges |= 1u << 9;     // flat-key wrap: (r, W-1) feeds (r..r+2, 0)
            if (c == 0) edges |= 1u << 10;        // flat-key wrap: (r, 0) feeds (r-2..r, W-1)
        }
    }
    edges = __reduce_or_sync(0xffffffffu, edges);
    if (lane == 0 && edges) atomicOr(s_edges, edges);
    __syncthreads();
    const uint32_t e = *s_edges;
    if (tid < 9 && tid != 4 && ((e >> tid) & 1u)) {
        const int ny = ty + tid / 3 - 1, nx = tx + tid % 3 - 1;
        if (ny >= 0 && ny < p.tiles_y && nx >= 0 && nx < p.tiles_x) p.tile_flag[ny * p.tiles_x + nx] = 1;
    }
    if (tid >= 9 && tid < 12 && ((e >> 9) & 1u)) {       // right map edge -> left-edge tiles of rows ty, ty+1 (and ty-1 for safety)
        const int ny = ty + (tid - 10);
        if (ny >= 0 && ny < p.tiles_y) p.tile_flag[ny * p.tiles_x + 0] = 1;
    }
    if (tid >= 12 && tid < 15 && ((e >> 10) & 1u)) {     // left map edge -> right-edge tiles of rows ty-1, ty (and ty+1 for safety)
        const int ny = ty + (tid - 13);
        if (ny >= 0 && ny < p.tiles_y) p.tile_flag[ny * p.tiles_x + p.tiles_x - 1] = 1;
    }
    if (tid == 0) { atomicAdd(p.counts + 3, 1); atomicAdd(p.counts + 6, sweep + 1); }
    (void)s_sweeps;
    __syncthreads();   // smem is reused by the next tile of this block
}

__global__ void __launch_bounds__(FIELD_THREADS, 1) field_solve_kernel(FieldParams p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s[SROWS * SROW];
    __shared__ uint32_t s_act[3][4];
    __shared__ uint32_t s_edges;
    __shared__ int s_sweeps;

    const int n_tiles = p.tiles_x * p.tiles_y;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gthreads = (long long)gridDim.x * blockDim.x;

    for (int round = 0; round < p.max_rounds; ++round) {
        const int* cur = (round & 1) ? p.list1 : p.list0;
        int* nxt = (round & 1) ? p.list0 : p.list1;
        const int n_cur = *((volatile int*)(p.counts + (round & 1)));
        if (n_cur == 0) break;
        for (int t = blockIdx.x; t < n_cur; t += gridDim.x) relax_tile(p, cur[t], s, s_act, &s_edges, &s_sweeps);
        __threadfence();
        grid.sync();
        // ---- compact the flagged tiles into the next list (ballot + popc, one atomic per warp) -----------
        if (gtid == 0) { p.counts[round & 1] = 0; p.counts[2] = round + 1; }
        for (long long base = gtid - (gtid & 31); base < n_tiles; base += gthreads) {
            const long long i = base + (gtid & 31);
            const bool f = i < n_tiles && p.tile_flag[i];
            const uint32_t b = __ballot_sync(0xffffffffu, f);
            if (b) {
                int off = 0;
                if ((gtid & 31) == 0) off = atomicAdd(p.counts + ((round + 1) & 1), __popc(b));
                off = __shfl_sync(0xffffffffu, off, 0);
                if (f) {
                    nxt[off + __popc(b & ((1u << (gtid & 31)) - 1u))] = (int)i;
                    p.tile_flag[i] = 0;
                }
            }
        }
        __threadfence();
        grid.sync();
    }
}

// ---- predecessor field + reached count + "would throw" check ------------------------------------------------
// pred[v] = lowest direction d whose in-edge attains g[v] exactly (the relaxation computed g[v] as exactly that
// fp32 sum, so equality is exact).  The reference keeps whichever tied parent its heap order met first; any
// of them is a valid optimal predecessor (north_star: path validity exact, cost within 1e-5).
__global__ void __launch_bounds__(256) field_pred_kernel(const float* __restrict__ g, const uint16_t* __restrict__ in_mask,
                                                         const uint16_t* __restrict__ out_mask, int W, long long N, long long ks,
                                                         uint8_t* __restrict__ pred, int* __restrict__ counts) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int reached = 0, thrown = 0;
    if (v < N) {
        const float gv = g[v];
        uint8_t pd = 255;
        if (gv < kInfF) {
            reached = 1;
            if (__ldg(out_mask + v) & kThrowBit) thrown = 1;   // a reached cell is expanded; expanding this one throws
            if (v != ks) {
                const uint32_t m = __ldg(in_mask + v);
#pragma unroll
                for (int d = 7; d >= 0; --d) {
                    if (!((m >> d) & 1u)) continue;
                    const float cand = g[v - nb_off(d, W)] + (((m >> (8 + d)) & 1u) ? kSqrt2F : 1.0f);
                    if (cand == gv) pd = (uint8_t)d;
                }
            }
        }
        pred[v] = pd;
    }
    const uint32_t rb = __ballot_sync(0xffffffffu, reached);
    const uint32_t tb = __ballot_sync(0xffffffffu, thrown);
    if ((threadIdx.x & 31) == 0) {
        if (rb) atomicAdd(counts + 5, __popc(rb));
        if (tb) atomicOr(counts + 4, 1);
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
    NSFS_CUDA_TRY(h, cudaMalloc(&f.tile_flag, (size_t)n_tiles + 32));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.list[0], (size_t)(n_tiles + 32) * sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.list[1], (size_t)(n_tiles + 32) * sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMalloc(&f.counts, 16 * sizeof(int)));
    NSFS_CUDA_TRY(h, cudaMemsetAsync(f.tile_flag, 0, (size_t)n_tiles + 32, h->stream));
    return NSFS_OK;
}

void field_free(nsfs_planner* h) {
    FieldWork& f = h->field;
    cudaFree(f.g); cudaFree(f.pred); cudaFree(f.tile_flag); cudaFree(f.list[0]); cudaFree(f.list[1]); cudaFree(f.counts);
    f = FieldWork();
}

int launch_field(nsfs_planner* h, int si, int sj, nsfs_stats* st) {
    int rc = field_alloc(h);
    if (rc != NSFS_OK) return rc;
    FieldWork& f = h->field;
    f.valid = false;
    FieldParams p;
    p.g = f.g; p.in_mask = h->d_in_mask; p.tile_flag = f.tile_flag; p.list0 = f.list[0]; p.list1 = f.list[1];
    p.counts = f.counts; p.H = h->H; p.W = h->W; p.N = h->N; p.tiles_x = f.tiles_x; p.tiles_y = f.tiles_y;
    p.max_rounds = 1 << 30;

    const long long n4 = h->N / 4;
    long long fill_blocks = (n4 + 255) / 256;
    if (fill_blocks < 1) fill_blocks = 1;
    if (fill_blocks > (long long)h->sm_count * 8) fill_blocks = (long long)h->sm_count * 8;
    const int fill_grid = (int)fill_blocks;
    field_fill_kernel<<<fill_grid, 256, 0, h->stream>>>(reinterpret_cast<float4*>(f.g), n4, f.g, h->N);
    field_seed_kernel<<<1, 32, 0, h->stream>>>(p, h->d_free, h->d_out_mask, si, sj);
    h->launches += 2;
    NSFS_CUDA_TRY(h, cudaGetLastError());

    int per_sm = 0;
    NSFS_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, field_solve_kernel, FIELD_THREADS, 0));
    if (per_sm < 1) return set_cuda_error(h, cudaErrorLaunchOutOfResources, "field_solve_kernel occupancy");
    int grid = per_sm * h->sm_count;
    const int n_tiles = f.tiles_x * f.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    void* args[] = {&p};
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev2, h->stream));
    NSFS_CUDA_TRY(h, cudaLaunchCooperativeKernel((void*)field_solve_kernel, dim3(grid), dim3(FIELD_THREADS), args, 0, h->stream));
    NSFS_CUDA_TRY(h, cudaEventRecord(h->ev3, h->stream));
    const long long ks = (long long)si * h->W + sj;
    field_pred_kernel<<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(f.g, h->d_in_mask, h->d_out_mask, h->W, h->N, ks,
                                                                             f.pred, f.counts);
    h->launches += 2;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    f.src_i = si; f.src_j = sj;
    (void)st;
    return NSFS_OK;
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
