// Warp-parallel Bresenham line of sight and the any-angle shortcut pass.
//
// bresenham_los (bot_utils.cpp:137-192) walks the long axis one cell at a time and bumps the short axis when
// 2|err| >= |Dl|.  That recurrence has the closed form (checked cell-for-cell in tests/):
//     cell t:   l = l0 + t*dl,   s = s0 + ds * floor((2 t |Ds| + |Dl|) / (2 |Dl|)),   t = 0..|Dl|
// with the long axis = i iff |Di| > |Dj| (ties go to j).  So a ray needs no carried state: lane t evaluates
// cell t, tests it exactly like testPos (grid_planner_core.cpp:434-455: row-only bounds test, flat key,
// vector::at throwing for a key outside [0,N)), and a ballot finds the first failing cell, which is what
// isLos (grid_planner_core.cpp:481-496) reports.
#include "nsfs_internal.cuh"

namespace nsfs {

// 1 = line of sight, 0 = blocked, -1 = the reference would throw std::out_of_range.  Warp-uniform result.
__device__ __forceinline__ int warp_is_los(const uint32_t* __restrict__ free_bits, int H, int W, long long N,
                                           int si, int sj, int ti, int tj, int lane) {
    const int Di = ti - si, Dj = tj - sj;
    const bool long_i = abs(Di) > abs(Dj);
    const int Dl = long_i ? Di : Dj, Ds = long_i ? Dj : Di;
    const int l0 = long_i ? si : sj, s0 = long_i ? sj : si;
    const long long aDl = abs(Dl), aDs = abs(Ds);
    const int dl = (Dl > 0) - (Dl < 0), ds = (Ds > 0) - (Ds < 0);
    for (long long base = 0; base <= aDl; base += 32) {
        const long long t = base + lane;
        int r = 1;
        if (t <= aDl) {
            const int l = l0 + (int)t * dl;
            const int s = s0 + (aDl ? ds * (int)((2 * t * aDs + aDl) / (2 * aDl)) : 0);
            const int i = long_i ? l : s, j = long_i ? s : l;
            const long long key = (long long)i * W + j;
            if (i < 0 || i >= H) r = 0;
            else if (key < 0 || key >= N) r = -1;
            else r = bit_at(free_bits, key) ? 1 : 0;
        }
        const uint32_t bad = __ballot_sync(0xffffffffu, r != 1);
        if (bad) return __shfl_sync(0xffffffffu, r, __ffs(bad) - 1);     // first failing cell in ray order decides
    }
    return 1;
}

__global__ void __launch_bounds__(256) los_batch_kernel(const uint32_t* __restrict__ free_bits, int H, int W, long long N, int n,
                                                        const int32_t* __restrict__ seg4, int8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(seg4) + s);
        const int r = warp_is_los(free_bits, H, W, N, q.x, q.y, q.z, q.w, lane);
        if (lane == 0) out[s] = (int8_t)r;
    }
}

// makeAnyAnglePath (grid_planner_core.cpp:537-558): greedy and sequential over the turning points, each test a
// warp-wide ray.  status: NSFS_OK or NSFS_E_RANGE.
__global__ void any_angle_kernel(const uint32_t* __restrict__ free_bits, int H, int W, long long N, int n,
                                 const int32_t* __restrict__ pts, uint8_t* __restrict__ keep, int* __restrict__ status) {
    const int lane = threadIdx.x;
    int ci = pts[0], cj = pts[1], st = NSFS_OK;
    if (lane == 0) { keep[0] = 1; if (n > 1) keep[n - 1] = 1; }
    for (int t = 1; t < n - 1; ++t) {
        const int ti = pts[2 * (t + 1)], tj = pts[2 * (t + 1) + 1];
        const int r = warp_is_los(free_bits, H, W, N, ci, cj, ti, tj, lane);
        if (r < 0) { st = NSFS_E_RANGE; break; }
        if (lane == 0) keep[t] = r ? 0 : 1;
        if (!r) { ci = pts[2 * t]; cj = pts[2 * t + 1]; }
    }
    if (lane == 0) *status = st;
}

int launch_los_batch(nsfs_planner* h, int n, const int32_t* d_seg4, int8_t* d_out) {
    if (n <= 0) return NSFS_OK;
    long long blocks = ((long long)n * 32 + 255) / 256;
    const long long cap = (long long)h->sm_count * 8;
    if (blocks > cap) blocks = cap;
    los_batch_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(h->d_free, h->H, h->W, h->N, n, d_seg4, d_out);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

int launch_any_angle(nsfs_planner* h, int n, const int32_t* d_pts, uint8_t* d_keep) {
    // status lands in the first int of the device scratch
    any_angle_kernel<<<1, 32, 0, h->stream>>>(h->d_free, h->H, h->W, h->N, n, d_pts, d_keep, (int*)h->d_scratch);
    h->launches++;
    NSFS_CUDA_TRY(h, cudaGetLastError());
    return NSFS_OK;
}

}  // namespace nsfs
