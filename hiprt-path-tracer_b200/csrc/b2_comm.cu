/*
 * b2_comm.cu -- multi-GPU part of the device layer (include/b2pt.h, group 1b): the pixel-row partition of a frame and the
 * exchanges between the GPUs of one box, issued from C on the caller's stream with NCCL (NVLink / NVSwitch).
 *
 * The reference is single-GPU; what crosses pixels in its hot path are the ReSTIR DI reuse passes only
 * (/root/reference/src/Device/kernels/ReSTIR/DI/SpatialReuse.h:142, FusedSpatiotemporalReuse.h:90,380,397 read the reservoir and
 * the G-buffer of pixels up to reuse_radius away; UtilsTemporal.h:33-96 up to neighbor_search_radius away), so a partitioned
 * frame needs, per frame, the reservoir rows next to a rank's bands (HALO exchange: ncclSend / ncclRecv between the ranks whose
 * bands touch) and, per step, the framebuffer rows of every rank everywhere (ncclAllGather).
 *
 * Partition: interleaved bands of `band_rows` rows, band b -> rank b % world (b2pt_partition_rows). Rows are contiguous in every
 * buffer involved (framebuffer: width*12 B per row; reservoirs: 9 SoA planes of width*4 B per row), so an exchange is
 * "pack rows -> one NCCL call (group) -> unpack rows" with one small copy kernel on either side, all on the same stream:
 * no host synchronisation and no Python in the data path.
 *
 * NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process -- torch's -- or the system one): libb2pt.so has
 * no link-time dependency on it and single-GPU users never load it.
 */
#include "b2pt.h"

#include "b2_comm.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

void b2_set_error(const char* fmt, ...);

#define B2_CUDA(call)                                                                                  \
    do                                                                                                 \
    {                                                                                                  \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
        {                                                                                              \
            b2_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return B2PT_ERR_CUDA;                                                                      \
        }                                                                                              \
    } while (0)

namespace
{
struct NcclApi
{
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        /* the copy already mapped into the process first (torch bundles its own libnccl.so.2), then the system's */
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h)
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
        {
            api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        api.handle = h;
        bool ok = true;
        auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) { ok = false; api.error = std::string("libnccl is missing ") + name; } return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        if (!ok)
            api.handle = nullptr;
    });
    return api;
}

#define B2_NCCL(call)                                                                                                 \
    do                                                                                                                \
    {                                                                                                                 \
        ncclResult_t r_ = (call);                                                                                     \
        if (r_ != ncclSuccess)                                                                                        \
        {                                                                                                             \
            b2_set_error("%s failed: %s (%s:%d)", #call, nccl().GetErrorString(r_), __FILE__, __LINE__);               \
            return B2PT_ERR_CUDA;                                                                                     \
        }                                                                                                             \
    } while (0)

/* rows[i] of `planes` planes: dst row list / src row list null = rows packed back to back (row i at i * row_words).
 * One thread per 16-byte chunk when the row size allows it, else per 4-byte word. */
template <class T>
__global__ void copy_rows_kernel(T* __restrict__ dst, const int* __restrict__ dst_rows, size_t dst_plane_stride, size_t dst_packed_rows,
                                 const T* __restrict__ src, const int* __restrict__ src_rows, size_t src_plane_stride, size_t src_packed_rows, int n_rows,
                                 int row_elems, int planes)
{
    const size_t per_plane = (size_t)n_rows * row_elems;
    const size_t total = per_plane * planes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    {
        const int plane = (int)(i / per_plane);
        const size_t in_plane = i - (size_t)plane * per_plane;
        const int r = (int)(in_plane / row_elems);
        const int x = (int)(in_plane - (size_t)r * row_elems);
        const size_t s = src_rows ? (size_t)plane * src_plane_stride + (size_t)src_rows[r] * row_elems + x : ((size_t)plane * src_packed_rows + r) * row_elems + x;
        const size_t d = dst_rows ? (size_t)plane * dst_plane_stride + (size_t)dst_rows[r] * row_elems + x : ((size_t)plane * dst_packed_rows + r) * row_elems + x;
        dst[d] = src[s];
    }
}

/* strides and sizes in BYTES; picks the widest element the layout allows */
cudaError_t copy_rows(void* dst, const int* dst_rows, size_t dst_plane_stride, size_t dst_packed_rows, const void* src, const int* src_rows, size_t src_plane_stride,
                      size_t src_packed_rows, int n_rows, size_t row_bytes, int planes, cudaStream_t stream)
{
    if (n_rows <= 0 || planes <= 0)
        return cudaSuccess;
    const bool wide = row_bytes % 16 == 0 && dst_plane_stride % 16 == 0 && src_plane_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(dst) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(src) % 16 == 0);
    const size_t elem = wide ? 16 : 4;
    const size_t total = (size_t)n_rows * (row_bytes / elem) * planes;
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, 148 * 16);
    if (wide)
        copy_rows_kernel<uint4><<<blocks, threads, 0, stream>>>((uint4*)dst, dst_rows, dst_plane_stride / 16, dst_packed_rows, (const uint4*)src, src_rows, src_plane_stride / 16,
                                                                 src_packed_rows, n_rows, (int)(row_bytes / 16), planes);
    else
        copy_rows_kernel<uint32_t><<<blocks, threads, 0, stream>>>((uint32_t*)dst, dst_rows, dst_plane_stride / 4, dst_packed_rows, (const uint32_t*)src, src_rows,
                                                                    src_plane_stride / 4, src_packed_rows, n_rows, (int)(row_bytes / 4), planes);
    return cudaGetLastError();
}

struct PeerPlan
{
    int peer = 0;
    std::vector<int> send_rows, recv_rows; /* ascending frame rows */
    int* d_send_rows = nullptr;
    int* d_recv_rows = nullptr;
};

struct HaloPlan
{
    std::vector<PeerPlan> peers;
};

struct GatherPlan
{
    int max_rows = 0;
    std::vector<std::vector<int>> rows; /* per rank */
    int* d_rows = nullptr;              /* [world][max_rows], -1 padded (padding rows are never copied: counts are per rank) */
};
} // namespace

struct b2pt_comm
{
    b2pt_ctx* ctx = nullptr;
    int device = 0;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    std::map<std::tuple<int, int, int>, HaloPlan> halo_plans;     /* (height, band_rows, halo) */
    std::map<std::tuple<int, int>, GatherPlan> gather_plans;      /* (height, band_rows) */
    void* staging_send = nullptr;
    void* staging_recv = nullptr;
    size_t staging_bytes = 0;
    std::vector<void*> allocations;
    uint64_t exchanges = 0, gathers = 0;
    uint64_t bytes_sent = 0;
};

/* ---- partition (pure host logic; tests/test_partition.py checks it against b2pt/partition.py) ------------------------------------ */
extern "C" int b2pt_partition_rows(int height, int rank, int world, int band_rows, int halo, int* out_rows, int* out_count)
{
    if (height <= 0 || world < 1 || rank < 0 || rank >= world || band_rows < 1 || halo < 0 || !out_count)
    {
        b2_set_error("b2pt_partition_rows: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    int n = 0;
    for (int y = 0; y < height; y++)
    {
        /* y is wanted when a row within `halo` of it belongs to one of rank's bands */
        bool wanted = false;
        const int lo = std::max(0, y - halo), hi = std::min(height - 1, y + halo);
        for (int band = lo / band_rows; band <= hi / band_rows && !wanted; band++)
            wanted = band % world == rank;
        if (wanted)
        {
            if (out_rows)
                out_rows[n] = y;
            n++;
        }
    }
    *out_count = n;
    return B2PT_OK;
}

std::vector<int> b2_partition_rows(int height, int rank, int world, int band_rows, int halo)
{
    std::vector<int> rows((size_t)height);
    int n = 0;
    b2pt_partition_rows(height, rank, world, band_rows, halo, rows.data(), &n);
    rows.resize((size_t)n);
    return rows;
}

extern "C" int b2pt_comm_unique_id(unsigned char out_id[B2PT_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == B2PT_COMM_ID_BYTES, "ncclUniqueId size");
    if (!out_id) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (!nccl().handle) { b2_set_error("%s", nccl().error.c_str()); return B2PT_ERR_NO_DEVICE; }
    ncclUniqueId id;
    B2_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(out_id, &id, sizeof(id));
    return B2PT_OK;
}

int b2_ctx_device(b2pt_ctx* ctx); /* b2_capi.cu */

extern "C" int b2pt_comm_create(b2pt_ctx* ctx, const unsigned char id_bytes[B2PT_COMM_ID_BYTES], int rank, int world, b2pt_comm** out_comm)
{
    if (!ctx || !id_bytes || !out_comm || world < 1 || rank < 0 || rank >= world)
    {
        b2_set_error("b2pt_comm_create: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    if (!nccl().handle) { b2_set_error("%s", nccl().error.c_str()); return B2PT_ERR_NO_DEVICE; }
    B2_CUDA(cudaSetDevice(b2_ctx_device(ctx)));
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof(id));
    b2pt_comm* c = new b2pt_comm();
    c->ctx = ctx;
    c->device = b2_ctx_device(ctx);
    c->rank = rank;
    c->world = world;
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess)
    {
        b2_set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
        delete c;
        return B2PT_ERR_CUDA;
    }
    *out_comm = c;
    return B2PT_OK;
}

extern "C" int b2pt_comm_destroy(b2pt_comm* c)
{
    if (!c) return B2PT_OK;
    cudaSetDevice(c->device);
    for (void* p : c->allocations)
        cudaFree(p);
    cudaFree(c->staging_send);
    cudaFree(c->staging_recv);
    if (c->comm)
        nccl().CommDestroy(c->comm);
    delete c;
    return B2PT_OK;
}

extern "C" int b2pt_comm_stats(b2pt_comm* c, uint64_t* out_exchanges, uint64_t* out_gathers, uint64_t* out_bytes_sent)
{
    if (!c) { b2_set_error("null argument"); return B2PT_ERR_INVALID_ARGUMENT; }
    if (out_exchanges) *out_exchanges = c->exchanges;
    if (out_gathers) *out_gathers = c->gathers;
    if (out_bytes_sent) *out_bytes_sent = c->bytes_sent;
    return B2PT_OK;
}

static int ensure_staging(b2pt_comm* c, size_t bytes)
{
    if (bytes <= c->staging_bytes)
        return B2PT_OK;
    cudaFree(c->staging_send);
    cudaFree(c->staging_recv);
    c->staging_send = c->staging_recv = nullptr;
    c->staging_bytes = 0;
    B2_CUDA(cudaMalloc(&c->staging_send, bytes));
    B2_CUDA(cudaMalloc(&c->staging_recv, bytes));
    c->staging_bytes = bytes;
    return B2PT_OK;
}

static int upload_rows(b2pt_comm* c, const std::vector<int>& rows, int** out)
{
    *out = nullptr;
    if (rows.empty())
        return B2PT_OK;
    B2_CUDA(cudaMalloc((void**)out, rows.size() * sizeof(int)));
    c->allocations.push_back(*out);
    B2_CUDA(cudaMemcpy(*out, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice));
    return B2PT_OK;
}

/* Host logic of the halo exchange, shared with tests: which rows rank sends to / receives from which peer. */
void b2_halo_plan_rows(int height, int rank, int world, int band_rows, int halo, std::vector<int>& peers, std::vector<std::vector<int>>& send_rows,
                       std::vector<std::vector<int>>& recv_rows)
{
    peers.clear();
    send_rows.clear();
    recv_rows.clear();
    const std::vector<int> need_mine = b2_partition_rows(height, rank, world, band_rows, halo);
    for (int q = 0; q < world; q++)
    {
        if (q == rank)
            continue;
        std::vector<int> send, recv;
        for (int y : b2_partition_rows(height, q, world, band_rows, halo))
            if ((y / band_rows) % world == rank)
                send.push_back(y); /* q wants it, I own it */
        for (int y : need_mine)
            if ((y / band_rows) % world == q)
                recv.push_back(y); /* I want it, q owns it */
        if (send.empty() && recv.empty())
            continue;
        peers.push_back(q);
        send_rows.push_back(send);
        recv_rows.push_back(recv);
    }
}

extern "C" int b2pt_comm_exchange_halo(b2pt_comm* c, void* base, int planes, size_t plane_stride_bytes, size_t row_bytes, int height, int band_rows, int halo,
                                       b2pt_stream stream)
{
    if (!c || !base || planes < 1 || row_bytes == 0 || row_bytes % 4 != 0 || plane_stride_bytes % 4 != 0 || height < 1 || band_rows < 1 || halo < 0)
    {
        b2_set_error("b2pt_comm_exchange_halo: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    if (c->world == 1 || halo == 0)
        return B2PT_OK;
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const auto key = std::make_tuple(height, band_rows, halo);
    auto it = c->halo_plans.find(key);
    if (it == c->halo_plans.end())
    {
        HaloPlan plan;
        std::vector<int> peers;
        std::vector<std::vector<int>> send_rows, recv_rows;
        b2_halo_plan_rows(height, c->rank, c->world, band_rows, halo, peers, send_rows, recv_rows);
        for (size_t i = 0; i < peers.size(); i++)
        {
            PeerPlan p;
            p.peer = peers[i];
            p.send_rows = send_rows[i];
            p.recv_rows = recv_rows[i];
            int rc = upload_rows(c, p.send_rows, &p.d_send_rows);
            if (rc == B2PT_OK)
                rc = upload_rows(c, p.recv_rows, &p.d_recv_rows);
            if (rc != B2PT_OK)
                return rc;
            plan.peers.push_back(p);
        }
        it = c->halo_plans.emplace(key, std::move(plan)).first;
    }
    const HaloPlan& plan = it->second;
    size_t send_total = 0, recv_total = 0;
    for (const PeerPlan& p : plan.peers)
    {
        send_total += p.send_rows.size();
        recv_total += p.recv_rows.size();
    }
    const size_t packed_row = row_bytes * planes; /* one frame row across all planes */
    int rc = ensure_staging(c, std::max(send_total, recv_total) * packed_row);
    if (rc != B2PT_OK)
        return rc;

    /* pack: per peer a block of [planes][rows][row_bytes] */
    size_t off = 0;
    for (const PeerPlan& p : plan.peers)
    {
        B2_CUDA(copy_rows(static_cast<char*>(c->staging_send) + off, nullptr, 0, p.send_rows.size(), base, p.d_send_rows, plane_stride_bytes, 0, (int)p.send_rows.size(),
                          row_bytes, planes, s));
        off += p.send_rows.size() * packed_row;
    }
    B2_NCCL(nccl().GroupStart());
    size_t send_off = 0, recv_off = 0;
    for (const PeerPlan& p : plan.peers)
    {
        if (!p.send_rows.empty())
            B2_NCCL(nccl().Send(static_cast<char*>(c->staging_send) + send_off, p.send_rows.size() * packed_row, ncclUint8, p.peer, c->comm, s));
        if (!p.recv_rows.empty())
            B2_NCCL(nccl().Recv(static_cast<char*>(c->staging_recv) + recv_off, p.recv_rows.size() * packed_row, ncclUint8, p.peer, c->comm, s));
        send_off += p.send_rows.size() * packed_row;
        recv_off += p.recv_rows.size() * packed_row;
    }
    B2_NCCL(nccl().GroupEnd());
    off = 0;
    for (const PeerPlan& p : plan.peers)
    {
        B2_CUDA(copy_rows(base, p.d_recv_rows, plane_stride_bytes, 0, static_cast<char*>(c->staging_recv) + off, nullptr, 0, p.recv_rows.size(), (int)p.recv_rows.size(),
                          row_bytes, planes, s));
        off += p.recv_rows.size() * packed_row;
    }
    c->exchanges++;
    c->bytes_sent += send_total * packed_row;
    return B2PT_OK;
}

extern "C" int b2pt_comm_allgather_rows(b2pt_comm* c, void* base, int planes, size_t plane_stride_bytes, size_t row_bytes, int height, int band_rows, b2pt_stream stream)
{
    if (!c || !base || planes < 1 || row_bytes == 0 || row_bytes % 4 != 0 || plane_stride_bytes % 4 != 0 || height < 1 || band_rows < 1)
    {
        b2_set_error("b2pt_comm_allgather_rows: invalid argument");
        return B2PT_ERR_INVALID_ARGUMENT;
    }
    if (c->world == 1)
        return B2PT_OK;
    B2_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const auto key = std::make_tuple(height, band_rows);
    auto it = c->gather_plans.find(key);
    if (it == c->gather_plans.end())
    {
        GatherPlan plan;
        for (int q = 0; q < c->world; q++)
        {
            plan.rows.push_back(b2_partition_rows(height, q, c->world, band_rows, 0));
            plan.max_rows = std::max(plan.max_rows, (int)plan.rows.back().size());
        }
        std::vector<int> table((size_t)c->world * plan.max_rows, -1);
        for (int q = 0; q < c->world; q++)
            std::copy(plan.rows[q].begin(), plan.rows[q].end(), table.begin() + (size_t)q * plan.max_rows);
        int rc = upload_rows(c, table, &plan.d_rows);
        if (rc != B2PT_OK)
            return rc;
        it = c->gather_plans.emplace(key, std::move(plan)).first;
    }
    const GatherPlan& plan = it->second;
    if (plan.max_rows == 0)
        return B2PT_OK;
    const size_t block = (size_t)plan.max_rows * row_bytes * planes; /* per rank, padded to the largest owner */
    /* staging_send holds this rank's block, staging_recv all of them */
    if (c->staging_bytes < block * c->world)
    {
        int rc = ensure_staging(c, block * c->world);
        if (rc != B2PT_OK)
            return rc;
    }
    const std::vector<int>& mine = plan.rows[c->rank];
    B2_CUDA(copy_rows(c->staging_send, nullptr, 0, plan.max_rows, base, plan.d_rows + (size_t)c->rank * plan.max_rows, plane_stride_bytes, 0, (int)mine.size(), row_bytes, planes, s));
    B2_NCCL(nccl().AllGather(c->staging_send, c->staging_recv, block, ncclUint8, c->comm, s));
    for (int q = 0; q < c->world; q++)
    {
        if (q == c->rank || plan.rows[q].empty())
            continue;
        B2_CUDA(copy_rows(base, plan.d_rows + (size_t)q * plan.max_rows, plane_stride_bytes, 0, static_cast<char*>(c->staging_recv) + (size_t)q * block, nullptr, 0, plan.max_rows,
                          (int)plan.rows[q].size(), row_bytes, planes, s));
    }
    c->gathers++;
    c->bytes_sent += mine.size() * row_bytes * planes;
    return B2PT_OK;
}

int b2_comm_rank(const b2pt_comm* c) { return c->rank; }
int b2_comm_world(const b2pt_comm* c) { return c->world; }
