// nsfs_group_*: one rank of a one-process-per-GPU group, for the two sharded configurations:
//   batched queries   the packed map is broadcast once (N/4 bytes), every rank plans its contiguous slice, results are
//                     all-gathered: no exchange during the search
//   large-map field   row blocks; shards post boundary rows into each other's memory over NVLink from inside the solve
//                     kernel (nsfs_field_relax_p2p); this file does the plumbing around it: CUDA IPC handles exchanged
//                     through an NCCL all-gather, the shared pending counter primed with an all-reduce, two barriers per solve
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the hosting process already loaded, or the system one), so
// libnsfs_gpu.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "nsfs_internal.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
#define NSFS_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym))
    NSFS_SYM(GetUniqueId, "ncclGetUniqueId");
    NSFS_SYM(CommInitRank, "ncclCommInitRank");
    NSFS_SYM(CommDestroy, "ncclCommDestroy");
    NSFS_SYM(Broadcast, "ncclBroadcast");
    NSFS_SYM(AllGather, "ncclAllGather");
    NSFS_SYM(AllReduce, "ncclAllReduce");
    NSFS_SYM(GetErrorString, "ncclGetErrorString");
#undef NSFS_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather && api.AllReduce;
    return api;
}

}  // namespace

struct nsfs_group {
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    void* d_buf = nullptr;        // staging for collectives
    size_t buf_bytes = 0;
    // field shards
    bool field_ready = false;
    int held_lo = 0, own_lo = 0, own_hi = 0;
    std::vector<int> owners_lo, owners_hi;     // every rank's owned global rows
    char error[256] = {0};
};

namespace {

int fail(nsfs_group* g, const char* what, const char* detail) {
    if (g) snprintf(g->error, sizeof(g->error), "%s: %s", what, detail ? detail : "");
    return NSFS_E_CUDA;
}
#define G_CUDA(g, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail((g), #expr, cudaGetErrorString(_e)); } while (0)
#define G_NCCL(g, expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) return fail((g), #expr, nccl().GetErrorString ? nccl().GetErrorString(_r) : "nccl error"); } while (0)

int ensure_buf(nsfs_group* g, size_t bytes) {
    if (bytes <= g->buf_bytes) return NSFS_OK;
    if (g->d_buf) cudaFree(g->d_buf);
    g->d_buf = nullptr; g->buf_bytes = 0;
    G_CUDA(g, cudaMalloc(&g->d_buf, bytes));
    g->buf_bytes = bytes;
    return NSFS_OK;
}

// contiguous, balanced: the first nq % world ranks take one extra query (same rule as nsfs/sharded_queries.py)
void query_slice(int nq, int world, int rank, int* a, int* b) {
    const int base = nq / world, extra = nq % world;
    *a = rank * base + (rank < extra ? rank : extra);
    *b = *a + base + (rank < extra ? 1 : 0);
}

}  // namespace

extern "C" {

int nsfs_group_unique_id(unsigned char id[128]) {
    if (!id || !nccl().ok) return NSFS_E_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "the id travels as 128 bytes");
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != ncclSuccess) return NSFS_E_CUDA;
    memcpy(id, &u, 128);
    return NSFS_OK;
}

int nsfs_group_create(nsfs_group_t** out, const unsigned char id[128], int rank, int world, int device) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return NSFS_E_ARG;
    if (!nccl().ok) return NSFS_E_CUDA;
    nsfs_group* g = new (std::nothrow) nsfs_group();
    if (!g) return NSFS_E_NOMEM;
    g->rank = rank; g->world = world; g->device = device;
    int rc = NSFS_OK;
    do {
        if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = NSFS_E_CUDA; break; }
        ncclUniqueId u;
        memcpy(&u, id, 128);
        if (nccl().CommInitRank(&g->comm, world, u, rank) != ncclSuccess) { rc = NSFS_E_CUDA; break; }
    } while (0);
    if (rc != NSFS_OK) { nsfs_group_destroy(g); return rc; }
    *out = g;
    return NSFS_OK;
}

void nsfs_group_destroy(nsfs_group_t* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    if (g->comm) nccl().CommDestroy(g->comm);
    if (g->d_buf) cudaFree(g->d_buf);
    if (g->stream) cudaStreamDestroy(g->stream);
    cudaGetLastError();
    delete g;
}

const char* nsfs_group_error(const nsfs_group_t* g) { return g ? g->error : ""; }
int nsfs_group_rank(const nsfs_group_t* g) { return g ? g->rank : -1; }
int nsfs_group_world(const nsfs_group_t* g) { return g ? g->world : 0; }

int nsfs_group_barrier(nsfs_group_t* g) {
    if (!g) return NSFS_E_ARG;
    G_CUDA(g, cudaSetDevice(g->device));
    int rc = ensure_buf(g, 256);
    if (rc != NSFS_OK) return rc;
    G_NCCL(g, nccl().AllReduce(g->d_buf, g->d_buf, 1, ncclInt32, ncclSum, g->comm, g->stream));
    G_CUDA(g, cudaStreamSynchronize(g->stream));
    return NSFS_OK;
}

// The packed free / inflated bit planes of the root's handle become every rank's map (N/4 bytes on the wire).
int nsfs_group_broadcast_map(nsfs_group_t* g, nsfs_t* h, int root) {
    if (!g || !h || root < 0 || root >= g->world) return NSFS_E_ARG;
    G_CUDA(g, cudaSetDevice(g->device));
    const size_t words = (size_t)nsfs_map_words(h);
    int rc = ensure_buf(g, 2 * words * 4 + 256);
    if (rc != NSFS_OK) return rc;
    uint32_t* planes = (uint32_t*)g->d_buf;
    if (g->rank == root) {
        const uint32_t *df = nullptr, *di = nullptr;
        if ((rc = nsfs_get_packed_map_device(h, &df, &di)) != NSFS_OK) return rc;
        if ((rc = nsfs_sync(h)) != NSFS_OK) return rc;
        G_CUDA(g, cudaMemcpyAsync(planes, df, words * 4, cudaMemcpyDeviceToDevice, g->stream));
        G_CUDA(g, cudaMemcpyAsync(planes + words, di, words * 4, cudaMemcpyDeviceToDevice, g->stream));
    }
    G_NCCL(g, nccl().Broadcast(planes, planes, 2 * words * 4, ncclUint8, root, g->comm, g->stream));
    G_CUDA(g, cudaStreamSynchronize(g->stream));
    if (g->rank != root) return nsfs_set_packed_map_device(h, planes, planes + words);
    return NSFS_OK;
}

// nq plan() calls split over the ranks: every rank passes ALL queries (host pointer) and receives ALL results (host arrays).
int nsfs_group_plan_batch(nsfs_group_t* g, nsfs_t* h, int algo, int cost_mode, int nq, const int32_t* sg4, int32_t* status,
                          double* g_goal, int32_t* path_len, long long* settled, nsfs_stats* st) {
    if (!g || !h || nq < 0 || (nq && !sg4)) return NSFS_E_ARG;
    if (!nq) return NSFS_OK;
    G_CUDA(g, cudaSetDevice(g->device));
    int a = 0, b = 0;
    query_slice(nq, g->world, g->rank, &a, &b);
    const int mine = b - a, pad = (nq + g->world - 1) / g->world;
    // per query, packed for the gather: { double g_goal; long long settled; int32 status; int32 path_len } = 24 bytes
    const size_t rec = 24;
    const size_t o_q = 0, o_loc = ((size_t)pad * 16 + 255) & ~(size_t)255, o_all = o_loc + (((size_t)pad * rec + 255) & ~(size_t)255);
    int rc = ensure_buf(g, o_all + (size_t)pad * g->world * rec + 256);
    if (rc != NSFS_OK) return rc;
    char* d = (char*)g->d_buf;
    int32_t* d_q = (int32_t*)(d + o_q);
    double* d_g = (double*)(d + o_loc);
    long long* d_set = (long long*)(d + o_loc + (size_t)pad * 8);
    int32_t* d_status = (int32_t*)(d + o_loc + (size_t)pad * 16);
    int32_t* d_len = (int32_t*)(d + o_loc + (size_t)pad * 20);
    G_CUDA(g, cudaMemsetAsync(d + o_loc, 0, (size_t)pad * rec, g->stream));
    if (mine) G_CUDA(g, cudaMemcpyAsync(d_q, sg4 + 4 * (size_t)a, (size_t)mine * 16, cudaMemcpyHostToDevice, g->stream));
    G_CUDA(g, cudaStreamSynchronize(g->stream));
    if (mine && (rc = nsfs_plan_batch_device(h, algo, cost_mode, mine, d_q, d_status, d_g, d_len, d_set, nullptr, 0, st)) != NSFS_OK) return rc;
    if ((rc = nsfs_sync(h)) != NSFS_OK) return rc;
    G_NCCL(g, nccl().AllGather(d + o_loc, d + o_all, (size_t)pad * rec, ncclUint8, g->comm, g->stream));
    std::vector<char> host((size_t)pad * g->world * rec);
    G_CUDA(g, cudaMemcpyAsync(host.data(), d + o_all, host.size(), cudaMemcpyDeviceToHost, g->stream));
    G_CUDA(g, cudaStreamSynchronize(g->stream));
    for (int r = 0; r < g->world; ++r) {
        int ra = 0, rb = 0;
        query_slice(nq, g->world, r, &ra, &rb);
        const char* blk = host.data() + (size_t)r * pad * rec;
        const size_t n = (size_t)(rb - ra);
        if (g_goal) memcpy(g_goal + ra, blk, n * 8);
        if (settled) memcpy(settled + ra, blk + (size_t)pad * 8, n * 8);
        if (status) memcpy(status + ra, blk + (size_t)pad * 16, n * 4);
        if (path_len) memcpy(path_len + ra, blk + (size_t)pad * 20, n * 4);
    }
    return NSFS_OK;
}

// Row-block field shards: this rank's handle covers global rows [held_lo, held_lo + H_local) and OWNS [own_lo, own_hi).
// Exports the shard's buffers (CUDA IPC), all-gathers every rank's handles and row ranges, attaches the neighbours above and
// below and the pending counter of rank 0.  Once per map geometry; nsfs_set_map* must have run on h.
int nsfs_group_field_attach(nsfs_group_t* g, nsfs_t* h, int held_lo, int own_lo, int own_hi) {
    if (!g || !h || own_hi - own_lo < 2 || own_lo < held_lo) return NSFS_E_ARG;
    G_CUDA(g, cudaSetDevice(g->device));
    struct Info { nsfs_peer_info peer; int held_lo, own_lo, own_hi, pad; };
    Info mine;
    memset(&mine, 0, sizeof(mine));
    int rc = nsfs_field_peer_export(h, &mine.peer);
    if (rc != NSFS_OK) return rc;
    mine.held_lo = held_lo; mine.own_lo = own_lo; mine.own_hi = own_hi;
    const size_t sz = sizeof(Info);
    if ((rc = ensure_buf(g, sz * (g->world + 1) + 256)) != NSFS_OK) return rc;
    char* d = (char*)g->d_buf;
    G_CUDA(g, cudaMemcpyAsync(d, &mine, sz, cudaMemcpyHostToDevice, g->stream));
    G_NCCL(g, nccl().AllGather(d, d + sz, sz, ncclUint8, g->comm, g->stream));
    std::vector<Info> all((size_t)g->world);
    G_CUDA(g, cudaMemcpyAsync(all.data(), d + sz, sz * g->world, cudaMemcpyDeviceToHost, g->stream));
    G_CUDA(g, cudaStreamSynchronize(g->stream));
    g->owners_lo.clear(); g->owners_hi.clear();
    for (const Info& i : all) { g->owners_lo.push_back(i.own_lo); g->owners_hi.push_back(i.own_hi); }
    if (g->rank > 0 && (rc = nsfs_field_peer_attach(h, 0, &all[g->rank - 1].peer, held_lo - all[g->rank - 1].held_lo)) != NSFS_OK) return rc;
    if (g->rank < g->world - 1 && (rc = nsfs_field_peer_attach(h, 1, &all[g->rank + 1].peer, held_lo - all[g->rank + 1].held_lo)) != NSFS_OK) return rc;
    if ((rc = nsfs_field_peer_counter(h, g->rank == 0 ? nullptr : &all[0].peer)) != NSFS_OK) return rc;
    if ((rc = nsfs_field_peer_rows(h, own_lo - held_lo, own_hi - held_lo)) != NSFS_OK) return rc;
    // a shard sees 1 / world of the wavefront, so the band of keys that keeps its SMs supplied with tiles is that much wider
    // (measured on C5: 2 shards 30.4 -> 28.1 ms with twice the single-GPU gate)
    if (h->opt_field_gate == 0) h->opt_field_gate = 192ll * g->world;
    g->held_lo = held_lo; g->own_lo = own_lo; g->own_hi = own_hi;
    g->field_ready = true;
    return nsfs_group_barrier(g);
}

// One field over all shards from the GLOBAL source (si, sj): seed on the owner, prime the shared counter, one solve kernel per
// GPU that ends when the whole map is at its fixed point, predecessors + reached count.  reached_total = cells reached on all shards.
int nsfs_group_field_solve(nsfs_group_t* g, nsfs_t* h, int si, int sj, long long* reached_total, nsfs_stats* st) {
    if (!g || !h || !g->field_ready) return NSFS_E_ARG;
    G_CUDA(g, cudaSetDevice(g->device));
    const bool mine = si >= g->own_lo && si < g->own_hi;
    int rc = nsfs_field_begin(h, mine ? si - g->held_lo : -1, sj);
    if (rc != NSFS_OK) return rc;
    int pending = 0;
    if ((rc = nsfs_field_pending_count(h, &pending)) != NSFS_OK) return rc;
    if ((rc = ensure_buf(g, 256)) != NSFS_OK) return rc;
    long long v[2] = {pending, 0};
    G_CUDA(g, cudaMemcpyAsync(g->d_buf, v, sizeof(v), cudaMemcpyHostToDevice, g->stream));
    G_NCCL(g, nccl().AllReduce(g->d_buf, g->d_buf, 1, ncclInt64, ncclSum, g->comm, g->stream));
    G_CUDA(g, cudaMemcpyAsync(v, g->d_buf, sizeof(long long), cudaMemcpyDeviceToHost, g->stream));
    G_CUDA(g, cudaStreamSynchronize(g->stream));
    if (g->rank == 0 && (rc = nsfs_field_set_counter(h, (int)v[0])) != NSFS_OK) return rc;
    if ((rc = nsfs_group_barrier(g)) != NSFS_OK) return rc;             // the counter is primed before any kernel may decrement it
    nsfs_stats relax_st;
    int status = nsfs_field_relax_p2p(h, &relax_st);
    if (status != NSFS_OK && status != NSFS_E_RANGE) return status;
    if ((rc = nsfs_group_barrier(g)) != NSFS_OK) return rc;             // every shard is at the global fixed point
    nsfs_stats fin;
    rc = nsfs_field_finish(h, &fin);
    if (rc != NSFS_OK && rc != NSFS_E_RANGE) return rc;
    if (rc == NSFS_E_RANGE) status = NSFS_E_RANGE;
    v[0] = fin.expansions; v[1] = status == NSFS_E_RANGE ? 1 : 0;
    G_CUDA(g, cudaMemcpyAsync(g->d_buf, v, sizeof(v), cudaMemcpyHostToDevice, g->stream));
    G_NCCL(g, nccl().AllReduce(g->d_buf, g->d_buf, 2, ncclInt64, ncclSum, g->comm, g->stream));
    G_CUDA(g, cudaMemcpyAsync(v, g->d_buf, sizeof(v), cudaMemcpyDeviceToHost, g->stream));
    G_CUDA(g, cudaStreamSynchronize(g->stream));
    if (reached_total) *reached_total = v[0];
    if (st) { *st = relax_st; st->expansions = fin.expansions; st->launches += fin.launches; st->gpu_ms += fin.gpu_ms; }
    return v[1] ? NSFS_E_RANGE : NSFS_OK;                               // a throwing cell anywhere fails the whole solve, like the exception
}

}  // extern "C"
