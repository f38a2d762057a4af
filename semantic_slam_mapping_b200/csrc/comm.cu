// comm.cu -- multi-GPU voxel fusion: spatially owned hash, NCCL all-to-all point routing.
//
// The reference has no distributed code.  Frames are sharded across ranks by the host; every rank turns its
// frames into labelled world-frame points (mapper.cu), buckets them by the owner rank of their voxel brick
// (voxel_owner: 8^3-voxel bricks hashed over the ranks), exchanges the buckets with one grouped
// ncclSend/ncclRecv all-to-all over NVLink, and fuses what it receives into its own table.  Counts, votes
// and the fixed-point centroid sums are integers, so the fused map is identical for any rank count.
//
// NCCL is resolved with dlopen at ssm_comm_init time (the process usually already holds torch's
// libnccl.so.2); libssm.so itself has no link-time NCCL dependency and single-GPU use never touches it.
#include <dlfcn.h>

#include <vector>

#include "ssm_internal.cuh"

namespace ssm {

// ---- minimal NCCL surface (matches nccl.h 2.x) ---------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclUint32 = 3 };
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.lib) return SSM_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        return SSM_ERR_COMM;
    }
#define SSM_SYM(field, name)                                                           \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));           \
    if (!g_nccl.field) { set_error("libnccl lacks " name); return SSM_ERR_COMM; }
    SSM_SYM(GetUniqueId, "ncclGetUniqueId")
    SSM_SYM(CommInitRank, "ncclCommInitRank")
    SSM_SYM(CommDestroy, "ncclCommDestroy")
    SSM_SYM(AllGather, "ncclAllGather")
    SSM_SYM(AllReduce, "ncclAllReduce")
    SSM_SYM(Send, "ncclSend")
    SSM_SYM(Recv, "ncclRecv")
    SSM_SYM(GroupStart, "ncclGroupStart")
    SSM_SYM(GroupEnd, "ncclGroupEnd")
    SSM_SYM(GetErrorString, "ncclGetErrorString")
#undef SSM_SYM
    g_nccl.lib = h;
    return SSM_OK;
}

static int nccl_fail(ncclResult_t r, const char* what)
{
    set_error(std::string("NCCL error: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?") + " at " + what);
    return SSM_ERR_COMM;
}
#define SSM_NCCL(expr)                                         \
    do {                                                       \
        ncclResult_t _r = (expr);                              \
        if (_r != 0) return nccl_fail(_r, #expr);              \
    } while (0)

// ---- bucketing kernels ---------------------------------------------------------------------------
constexpr int kMaxRanks = 64;

__device__ __forceinline__ int owner_of(const Point& pt, float inv_leaf, int nranks)
{
    if (!isfinite(pt.x) || !isfinite(pt.y) || !isfinite(pt.z)) return 0;   // dropped by the fuse kernel anyway
    const int i = (int)floorf(__fmul_rn(pt.x, inv_leaf));
    const int j = (int)floorf(__fmul_rn(pt.y, inv_leaf));
    const int k = (int)floorf(__fmul_rn(pt.z, inv_leaf));
    return voxel_owner(i, j, k, nranks);
}

// counts[r] += number of points owned by r (shared-memory histogram per CTA, one global atomic per rank per CTA)
__global__ void __launch_bounds__(256) k_route_count(const Point* __restrict__ pts, const uint32_t* __restrict__ n_ptr,
                                                     uint32_t max_n, float inv_leaf, int nranks, uint32_t* __restrict__ counts)
{
    __shared__ uint32_t h[kMaxRanks];
    if (threadIdx.x < kMaxRanks) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = min(*n_ptr, max_n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&h[owner_of(pts[i], inv_leaf, nranks)], 1u);
    __syncthreads();
    if (threadIdx.x < nranks && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}
// counts[0..R) -> offsets[R..2R), cursors[2R..3R)
__global__ void k_route_offsets(uint32_t* __restrict__ c, int nranks)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t acc = 0;
    for (int r = 0; r < nranks; ++r) {
        c[nranks + r] = acc;
        c[2 * nranks + r] = acc;
        acc += c[r];
    }
}
__global__ void __launch_bounds__(256) k_route_scatter(const Point* __restrict__ pts, const uint32_t* __restrict__ n_ptr,
                                                       uint32_t max_n, float inv_leaf, int nranks, uint32_t* __restrict__ c,
                                                       Point* __restrict__ send)
{
    const uint32_t n = min(*n_ptr, max_n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Point pt = pts[i];
        const int r = owner_of(pt, inv_leaf, nranks);
        send[atomicAdd(&c[2 * nranks + r], 1u)] = pt;
    }
}

int launch_route_bucket(ssm_ctx* c, uint32_t max_points, cudaStream_t s)
{
    const int R = c->nranks;
    SSM_CUDA(cudaMemsetAsync(c->d_send_counts, 0, sizeof(uint32_t) * 3 * R, s));
    const unsigned grid = (unsigned)c->sm_count * 8;
    k_route_count<<<grid, 256, 0, s>>>(c->d_points, c->d_counters, max_points, c->dp.inv_leaf, R, c->d_send_counts);
    SSM_LAUNCH_CHECK(c);
    k_route_offsets<<<1, 32, 0, s>>>(c->d_send_counts, R);
    SSM_LAUNCH_CHECK(c);
    k_route_scatter<<<grid, 256, 0, s>>>(c->d_points, c->d_counters, max_points, c->dp.inv_leaf, R, c->d_send_counts, c->d_send);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

// d_points[0..counters[0]) -> owners' tables
int route_and_fuse(ssm_ctx* c, cudaStream_t s)
{
    if (!c->comm) {
        set_error("ssm_comm_init has not been called on this context");
        return SSM_ERR_COMM;
    }
    const int R = c->nranks, me = c->rank;
    const uint32_t cap = (uint32_t)((size_t)c->cap_w * c->cap_h * c->cap_b);
    int rc = launch_route_bucket(c, cap, s);
    if (rc) return rc;
    ncclComm_t comm = (ncclComm_t)c->comm;
    uint32_t* d_all = c->d_send_counts + 3 * R;   // [R][R] gathered counts
    SSM_NCCL(g_nccl.AllGather(c->d_send_counts, d_all, R, ncclUint32, comm, s));
    std::vector<uint32_t> all((size_t)R * R), mine(3 * R);
    SSM_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(uint32_t) * R * R, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaMemcpyAsync(mine.data(), c->d_send_counts, sizeof(uint32_t) * 3 * R, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    size_t total_recv = 0;
    for (int r = 0; r < R; ++r) total_recv += all[(size_t)r * R + me];
    if (total_recv > c->route_cap) {   // grow the receive buffer (rare: sized for 2x a full local batch at init)
        if (c->d_recv) SSM_CUDA(cudaFree(c->d_recv));
        c->d_recv = nullptr;
        c->route_cap = total_recv + total_recv / 4;
        SSM_CUDA(cudaMalloc(&c->d_recv, sizeof(Point) * c->route_cap));
    }
    SSM_NCCL(g_nccl.GroupStart());
    size_t roff = 0;
    for (int r = 0; r < R; ++r) {
        const uint32_t scnt = mine[r], soff = mine[R + r], rcnt = all[(size_t)r * R + me];
        if (scnt) SSM_NCCL(g_nccl.Send(c->d_send + soff, (size_t)scnt * sizeof(Point), ncclUint8, r, comm, s));
        if (rcnt) SSM_NCCL(g_nccl.Recv(c->d_recv + roff, (size_t)rcnt * sizeof(Point), ncclUint8, r, comm, s));
        roff += rcnt;
    }
    SSM_NCCL(g_nccl.GroupEnd());
    return launch_fuse_points(c, c->d_recv, nullptr, (uint32_t)total_recv, s);
}

// stream-ordered barrier over the ranks: a one-word all-reduce (no host synchronisation)
int comm_barrier(ssm_ctx* c, cudaStream_t s)
{
    if (!c->comm) {
        set_error("ssm_comm_init has not been called on this context");
        return SSM_ERR_COMM;
    }
    uint32_t* word = c->d_send_counts;   // scratch
    SSM_NCCL(g_nccl.AllReduce(word, word, 1, ncclUint32, 0 /* ncclSum */, (ncclComm_t)c->comm, s));
    return SSM_OK;
}

// blocking: *value = max over the ranks (used to agree on collective round counts)
int comm_allreduce_max(ssm_ctx* c, uint32_t* value, cudaStream_t s)
{
    if (!c->comm) {
        set_error("ssm_comm_init has not been called on this context");
        return SSM_ERR_COMM;
    }
    uint32_t* word = c->d_send_counts + 1;   // scratch (word 0 is the barrier's)
    SSM_CUDA(cudaMemcpyAsync(word, value, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    SSM_NCCL(g_nccl.AllReduce(word, word, 1, ncclUint32, 2 /* ncclMax */, (ncclComm_t)c->comm, s));
    SSM_CUDA(cudaMemcpyAsync(value, word, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

// SURVEY 8e "Export": every rank compacts the occupied records of its table and sends them to rank 0, which returns the
// dense union (the tables are disjoint by ownership).  *d_all is a cudaMalloc'ed array the caller frees (NULL off rank 0).
int comm_gather_records(ssm_ctx* c, Voxel** d_all, uint64_t* n_all, cudaStream_t s)
{
    *d_all = nullptr;
    *n_all = 0;
    if (!c->comm) {
        set_error("ssm_comm_init has not been called on this context");
        return SSM_ERR_COMM;
    }
    const int R = c->nranks, me = c->rank;
    ncclComm_t comm = (ncclComm_t)c->comm;
    uint32_t mine = 0;
    SSM_CUDA(cudaMemcpyAsync(&mine, c->d_counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    uint32_t* d_cnt = c->d_send_counts + 3 * R;   // [R] gathered voxel counts (the [R][R] scratch of the routing path)
    SSM_NCCL(g_nccl.AllGather(c->d_counters + 1, d_cnt, 1, ncclUint32, comm, s));
    std::vector<uint32_t> cnt(R);
    SSM_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(uint32_t) * R, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    uint64_t total = 0;
    for (int r = 0; r < R; ++r) total += cnt[r];
    Voxel* buf = nullptr;
    const uint64_t need = me == 0 ? total : mine;
    if (need) SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&buf), sizeof(Voxel) * need));
    int rc = SSM_OK;
    if (mine) rc = launch_export(c, buf, mine, s);   // rank 0's own records come first
    if (rc == SSM_OK) {
        ncclResult_t nr = g_nccl.GroupStart();
        uint64_t off = cnt[0];
        for (int r = 1; r < R && nr == 0; ++r) {
            if (me == 0 && cnt[r]) nr = g_nccl.Recv(buf + off, (size_t)cnt[r] * sizeof(Voxel), ncclUint8, r, comm, s);
            if (me == r && mine) nr = g_nccl.Send(buf, (size_t)mine * sizeof(Voxel), ncclUint8, 0, comm, s);
            off += cnt[r];
        }
        const ncclResult_t ne = g_nccl.GroupEnd();
        if (nr == 0) nr = ne;
        if (nr != 0) rc = nccl_fail(nr, "record gather");
    }
    if (rc == SSM_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "record gather");
    if (rc != SSM_OK || me != 0) {
        if (buf) cudaFree(buf);
        return rc;
    }
    *d_all = buf;
    *n_all = total;
    return SSM_OK;
}

int points_route_p2p(ssm_ctx* c, int B, const uint16_t* d_depth, const uint8_t* d_sem, const uint8_t* d_rgb, const double* d_pose,
                     cudaStream_t s)
{
    const int parity = (int)(c->p2p_step & 1u);
    c->p2p_step++;
    int rc = launch_points_p2p(c, B, d_depth, d_sem, d_rgb, d_pose, c->d_peer_base, parity, s);
    if (rc) return rc;
    // every peer's stores for this step have landed: arrival flags in the peers' inbox headers (SSM_NCCL_BARRIER=1: the
    // one-word ncclAllReduce of round 1)
    static const bool nccl_barrier = [] { const char* e = getenv("SSM_NCCL_BARRIER"); return e && e[0] == '1'; }();
    if ((rc = nccl_barrier ? comm_barrier(c, s) : launch_flag_barrier(c, (uint32_t)c->p2p_step, s))) return rc;
    return launch_fuse_inbox(c, parity, s);
}

}  // namespace ssm

using namespace ssm;

extern "C" {

int ssm_voxel_owner(int32_t i, int32_t j, int32_t k, int nranks) { return voxel_owner(i, j, k, nranks); }

int ssm_comm_get_unique_id(uint8_t id[SSM_UNIQUE_ID_BYTES])
{
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId u;
    SSM_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, SSM_UNIQUE_ID_BYTES);
    return SSM_OK;
}

int ssm_comm_init(ssm_ctx* c, const uint8_t id[SSM_UNIQUE_ID_BYTES], int rank, int nranks)
{
    if (!c || !id || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) {
        set_error("bad rank / nranks");
        return SSM_ERR_INVALID_ARGUMENT;
    }
    if (c->comm) {
        set_error("this context already has a communicator: call ssm_comm_destroy first");
        return SSM_ERR_COMM;
    }
    int rc = load_nccl();
    if (rc) return rc;
    SSM_CUDA(cudaSetDevice(c->device));
    ncclUniqueId u;
    memcpy(u.internal, id, SSM_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    SSM_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
    c->comm = comm;
    c->rank = rank;
    c->nranks = nranks;
    const size_t cap = (size_t)c->cap_w * c->cap_h * c->cap_b;
    SSM_CUDA(cudaMalloc(&c->d_send, sizeof(Point) * cap));
    c->route_cap = 2 * cap;
    SSM_CUDA(cudaMalloc(&c->d_recv, sizeof(Point) * c->route_cap));
    SSM_CUDA(cudaMalloc(&c->d_send_counts, sizeof(uint32_t) * (3 * nranks + (size_t)nranks * nranks)));
    return SSM_OK;
}

int ssm_comm_ipc_export(ssm_ctx* c, uint8_t handle[SSM_IPC_HANDLE_BYTES])
{
    if (!c || !handle) { set_error("null argument"); return SSM_ERR_INVALID_ARGUMENT; }
    if (!c->comm) { set_error("call ssm_comm_init first"); return SSM_ERR_COMM; }
    static_assert(sizeof(cudaIpcMemHandle_t) <= SSM_IPC_HANDLE_BYTES, "handle size");
    SSM_CUDA(cudaSetDevice(c->device));
    if (!c->ipc_base) {
        // worst case every peer's whole batch is owned by this rank; sized for twice a local batch per parity,
        // overflow is reported (SSM_ERR_CAPACITY) rather than silently dropped
        c->inbox_cap = 2 * (size_t)c->cap_w * c->cap_h * c->cap_b;
        const size_t bytes = kInboxHeader + 2 * c->inbox_cap * 16;   // 16-byte records (mapper.cu: k_points_p2p)
        SSM_CUDA(cudaMalloc(&c->ipc_base, bytes));
    }
    SSM_CUDA(cudaMemset(c->ipc_base, 0, kInboxHeader));   // counts, overflow flag and the arrival flags (the step counter restarts at connect)
    cudaIpcMemHandle_t h;
    SSM_CUDA(cudaIpcGetMemHandle(&h, c->ipc_base));
    memset(handle, 0, SSM_IPC_HANDLE_BYTES);
    memcpy(handle, &h, sizeof(h));
    return SSM_OK;
}

int ssm_comm_ipc_connect(ssm_ctx* c, const uint8_t* handles, int nranks)
{
    if (!c || !handles) { set_error("null argument"); return SSM_ERR_INVALID_ARGUMENT; }
    if (!c->comm || !c->ipc_base || nranks != c->nranks || nranks > ssm_ctx::kMaxPeers) {
        set_error("ssm_comm_ipc_connect needs ssm_comm_init + ssm_comm_ipc_export first, with the same rank count");
        return SSM_ERR_COMM;
    }
    SSM_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < nranks; ++r) {
        if (r == c->rank) { c->peer_base[r] = c->ipc_base; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * SSM_IPC_HANDLE_BYTES, sizeof(h));
        SSM_CUDA(cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    SSM_CUDA(cudaMalloc(&c->d_peer_base, sizeof(void*) * nranks));
    SSM_CUDA(cudaMemcpy(c->d_peer_base, c->peer_base, sizeof(void*) * nranks, cudaMemcpyHostToDevice));
    c->p2p = true;
    c->p2p_step = 0;
    return SSM_OK;
}

int ssm_comm_destroy(ssm_ctx* c)
{
    if (c && c->p2p) {
        for (int r = 0; r < c->nranks; ++r)
            if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
        c->p2p = false;
    }
    if (c && c->d_peer_base) { cudaFree(c->d_peer_base); c->d_peer_base = nullptr; }
    if (c && c->ipc_base) { cudaFree(c->ipc_base); c->ipc_base = nullptr; }
    if (!c || !c->comm) return SSM_OK;
    g_nccl.CommDestroy((ncclComm_t)c->comm);
    // the routing buffers belong to the communicator (a later ssm_comm_init allocates them for its own rank count)
    if (c->d_send) { cudaFree(c->d_send); c->d_send = nullptr; }
    if (c->d_recv) { cudaFree(c->d_recv); c->d_recv = nullptr; }
    if (c->d_send_counts) { cudaFree(c->d_send_counts); c->d_send_counts = nullptr; }
    c->route_cap = 0;
    c->comm = nullptr;
    c->rank = 0;
    c->nranks = 1;
    return SSM_OK;
}

}  // extern "C"
