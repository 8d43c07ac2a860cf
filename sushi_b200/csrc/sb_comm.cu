// Multi-GPU plumbing of the matcher: one process per GPU, an NCCL communicator owned by
// the library, collectives enqueued on the library's own streams -- no Python framework on the timed
// path.  The path shards by events, so there are exactly two collectives: a broadcast of the normalised
// streams and an all-gather of the per-event (diff, idx) results; neither sits inside a kernel's data
// path, hence no fused compute + communication kernel.
//
// NCCL is opened with dlopen() when the first communicator call arrives: a single-GPU user never needs
// libnccl, and the library loads on hosts without it.  Broadcasts run on a second stream (c.comm_stream)
// so that the broadcast of one stream overlaps the running sums / spectra of the other; an event per
// slot orders the consumer on the library stream behind its own broadcast only.
#include "sb_internal.h"
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

using namespace sb;

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

constexpr int kSlots = 4;

struct Comm {
    NcclApi api;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    cudaStream_t stream = nullptr;            // broadcasts
    cudaEvent_t ready[kSlots] = {};           // slot's broadcast has finished (recorded on `stream`)
    cudaEvent_t fence = nullptr;              // library stream -> comm stream ordering
    float* d_scratch = nullptr;               // small device buffer for barriers / reductions
    float* h_scratch = nullptr;               // pinned
};

Comm& comm() { static Comm c; return c; }

#define SB_NCCL(expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) { \
    sb::set_error("%s failed: %s (%s:%d)", #expr, comm().api.GetErrorString ? comm().api.GetErrorString(r__) : "nccl error", __FILE__, __LINE__); \
    return SB_ECUDA; } } while (0)

int load_nccl() {
    NcclApi& a = comm().api;
    if (a.handle) return SB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (a.handle) break; }
    if (!a.handle) SB_FAIL(SB_ESTATE, "multi-GPU calls need NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
#define SB_SYM(field, name) do { *(void**)(&a.field) = dlsym(a.handle, name); \
    if (!a.field) { sb::set_error("libnccl lacks %s", name); dlclose(a.handle); a.handle = nullptr; return SB_ESTATE; } } while (0)
    SB_SYM(GetVersion, "ncclGetVersion");
    SB_SYM(GetUniqueId, "ncclGetUniqueId");
    SB_SYM(CommInitRank, "ncclCommInitRank");
    SB_SYM(CommDestroy, "ncclCommDestroy");
    SB_SYM(Broadcast, "ncclBroadcast");
    SB_SYM(AllGather, "ncclAllGather");
    SB_SYM(AllReduce, "ncclAllReduce");
    SB_SYM(GetErrorString, "ncclGetErrorString");
#undef SB_SYM
    return SB_OK;
}

int need_comm(const char* who) {
    if (!ctx().inited) SB_FAIL(SB_ESTATE, "%s: library not initialised (call sb_init)", who);
    if (!comm().comm) SB_FAIL(SB_ESTATE, "%s: no communicator (call sb_comm_init)", who);
    return SB_OK;
}

}  // namespace

extern "C" {

int sb_comm_unique_id(void* id_out) {
    if (!id_out) SB_FAIL(SB_EINVAL, "sb_comm_unique_id: NULL output");
    SB_TRY(load_nccl());
    ncclUniqueId id;
    SB_NCCL(comm().api.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == SB_COMM_ID_BYTES, "SB_COMM_ID_BYTES must equal NCCL_UNIQUE_ID_BYTES");
    memcpy(id_out, &id, sizeof(id));
    return SB_OK;
}

int sb_comm_init(const void* id, int world_size, int rank) {
    Ctx& c = ctx();
    Comm& m = comm();
    if (!c.inited) SB_FAIL(SB_ESTATE, "sb_comm_init: library not initialised (call sb_init)");
    if (m.comm) SB_FAIL(SB_ESTATE, "sb_comm_init: communicator already exists");
    if (!id || world_size < 1 || rank < 0 || rank >= world_size) SB_FAIL(SB_EINVAL, "sb_comm_init: bad arguments (world %d, rank %d)", world_size, rank);
    SB_TRY(load_nccl());
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    SB_CUDA(cudaSetDevice(c.device));
    SB_NCCL(m.api.CommInitRank(&m.comm, world_size, uid, rank));
    m.world = world_size; m.rank = rank;
    SB_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) SB_CUDA(cudaEventCreateWithFlags(&m.ready[i], cudaEventDisableTiming));
    SB_CUDA(cudaEventCreateWithFlags(&m.fence, cudaEventDisableTiming));
    SB_CUDA(cudaMalloc((void**)&m.d_scratch, 256));
    SB_CUDA(cudaMallocHost((void**)&m.h_scratch, 256));
    return SB_OK;
}

int sb_comm_destroy(void) {
    Comm& m = comm();
    if (!m.comm) return SB_OK;
    if (ctx().inited) cudaStreamSynchronize(ctx().stream);
    cudaStreamSynchronize(m.stream);
    m.api.CommDestroy(m.comm);
    m.comm = nullptr;
    for (int i = 0; i < kSlots; ++i) { cudaEventDestroy(m.ready[i]); m.ready[i] = nullptr; }
    cudaEventDestroy(m.fence); m.fence = nullptr;
    cudaStreamDestroy(m.stream); m.stream = nullptr;
    cudaFree(m.d_scratch); m.d_scratch = nullptr;
    cudaFreeHost(m.h_scratch); m.h_scratch = nullptr;
    m.world = 1; m.rank = 0;
    return SB_OK;
}

int sb_comm_world_size(void) { return comm().comm ? comm().world : 1; }
int sb_comm_rank(void) { return comm().comm ? comm().rank : 0; }
int sb_comm_nccl_version(void) {
    if (load_nccl() != SB_OK) return -1;
    int v = 0;
    return comm().api.GetVersion(&v) == ncclSuccess ? v : -1;
}

int sb_comm_broadcast(void* dev_buf, int64_t bytes, int root, int slot) {
    SB_TRY(need_comm("sb_comm_broadcast"));
    Comm& m = comm();
    if (!dev_buf || bytes < 0 || root < 0 || root >= m.world || slot < 0 || slot >= kSlots)
        SB_FAIL(SB_EINVAL, "sb_comm_broadcast: bad arguments");
    Ctx& c = ctx();
    // whatever the library stream did with this buffer before (last step's device-to-device copy) comes first
    SB_CUDA(cudaEventRecord(m.fence, c.stream));
    SB_CUDA(cudaStreamWaitEvent(m.stream, m.fence, 0));
    SB_NCCL(m.api.Broadcast(dev_buf, dev_buf, (size_t)bytes, ncclUint8, root, m.comm, m.stream));
    SB_CUDA(cudaEventRecord(m.ready[slot], m.stream));
    return SB_OK;
}

int sb_comm_wait(int slot) {
    SB_TRY(need_comm("sb_comm_wait"));
    if (slot < 0 || slot >= kSlots) SB_FAIL(SB_EINVAL, "sb_comm_wait: slot %d out of range", slot);
    SB_CUDA(cudaStreamWaitEvent(ctx().stream, comm().ready[slot], 0));
    return SB_OK;
}

int sb_comm_all_gather(const void* dev_send, void* dev_recv, int64_t bytes_per_rank) {
    SB_TRY(need_comm("sb_comm_all_gather"));
    if (!dev_send || !dev_recv || bytes_per_rank < 0) SB_FAIL(SB_EINVAL, "sb_comm_all_gather: bad arguments");
    Comm& m = comm();
    ProfScope ps("nccl_all_gather", 0);
    SB_NCCL(m.api.AllGather(dev_send, dev_recv, (size_t)bytes_per_rank, ncclUint8, m.comm, ctx().stream));
    return SB_OK;
}

int sb_comm_max_f32(float* host_inout, int count) {
    SB_TRY(need_comm("sb_comm_max_f32"));
    if (!host_inout || count < 1 || count > 32) SB_FAIL(SB_EINVAL, "sb_comm_max_f32: 1..32 values");
    Comm& m = comm();
    Ctx& c = ctx();
    memcpy(m.h_scratch, host_inout, sizeof(float) * count);
    SB_CUDA(cudaMemcpyAsync(m.d_scratch, m.h_scratch, sizeof(float) * count, cudaMemcpyHostToDevice, c.stream));
    SB_NCCL(m.api.AllReduce(m.d_scratch, m.d_scratch, (size_t)count, ncclFloat32, ncclMax, m.comm, c.stream));
    SB_CUDA(cudaMemcpyAsync(m.h_scratch, m.d_scratch, sizeof(float) * count, cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(host_inout, m.h_scratch, sizeof(float) * count);
    return SB_OK;
}

int sb_comm_barrier(void) {
    SB_TRY(need_comm("sb_comm_barrier"));
    Comm& m = comm();
    SB_CUDA(cudaStreamSynchronize(m.stream));
    float one = 1.f;
    return sb_comm_max_f32(&one, 1);           // every rank's library stream has drained when this returns
}

}  // extern "C"
