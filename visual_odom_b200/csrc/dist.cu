// dist.cu -- SURVEY.md 8(e) "C1": the only communication of the multi-GPU layer, the gather of the fixed-size result
// records, as C-ABI calls over NCCL (one process per GPU; units are sharded with no data-path collective).
//
//   rank 0: vo_dist_unique_id(id)  ->  the 128 bytes travel out of band (bench.py broadcasts them with torch.distributed)
//   all   : vo_dist_init(ctx, id, rank, world)
//   loop  : vo_batch_wait(slot) ; vo_dist_gather_post(ctx, slot, n)      non-blocking: ncclAllGather straight from the device
//           ...                 ; vo_dist_gather_wait(ctx, all, ...)     snapshot of the submission's records, + one D2H into pinned memory
// NCCL is resolved with dlopen at vo_dist_init, so libvo_b200.so itself does not link it (the library must load on hosts
// without NCCL, e.g. the CPU test box); the process-wide libnccl.so.2 that torch already loaded is the one that is found.
#include "ctx.h"
#include <dlfcn.h>
#include <string.h>

namespace {
struct NcclUid { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUid, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

const char* load_nccl()
{
    if (g_nccl.h) return nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return "libnccl.so.2 not found (dlopen)";
    g_nccl.GetUniqueId = (int (*)(NcclUid*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclUid, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) return "libnccl.so.2 lacks the expected symbols";
    g_nccl.h = h;
    return nullptr;
}
const char* nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"; }
}  // namespace

extern "C" int vo_dist_unique_id(uint8_t id_out[128])
{
    if (!id_out) return VO_E_INVALID;
    if (load_nccl()) return VO_E_UNSUPPORTED;
    NcclUid id;
    if (g_nccl.GetUniqueId(&id) != 0) return VO_E_CUDA;
    memcpy(id_out, id.internal, 128);
    return VO_OK;
}

extern "C" int vo_dist_init(vo_ctx* ctx, const uint8_t id[128], int rank, int world)
{
    if (!ctx) return VO_E_INVALID;
    if (!id || world < 1 || rank < 0 || rank >= world) { vo_set_error(ctx, "vo_dist_init: bad argument"); return VO_E_INVALID; }
    if (const char* e = load_nccl()) { vo_set_error(ctx, "vo_dist_init: %s", e); return VO_E_UNSUPPORTED; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (ctx->dist_comm) { g_nccl.CommDestroy((NcclComm)ctx->dist_comm); ctx->dist_comm = nullptr; }
    NcclUid uid;
    memcpy(uid.internal, id, 128);
    NcclComm comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (rc != 0) { vo_set_error(ctx, "ncclCommInitRank: %s", nccl_err(rc)); return VO_E_CUDA; }
    ctx->dist_comm = comm; ctx->dist_rank = rank; ctx->dist_world = world;
    if (!ctx->dist_stream) VO_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->dist_stream, cudaStreamNonBlocking));
    for (int k = 0; k < VO_DIST_DEPTH; k++) {
        if (!ctx->dist_ev_read[k]) VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->dist_ev_read[k], cudaEventDisableTiming));
        if (!ctx->dist_ev_done[k]) VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->dist_ev_done[k], cudaEventDisableTiming));
        ctx->dist_posted[k] = false;
    }
    ctx->dist_head = ctx->dist_tail = 0;
    return VO_OK;
}

static int dist_buffers(vo_ctx* ctx, int n_units)
{
    const size_t need = (size_t)ctx->dist_world * n_units * sizeof(vo_unit_result_dev);
    if (need <= ctx->dist_bytes) return VO_OK;
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->dist_stream));
    for (int k = 0; k < VO_DIST_DEPTH; k++) {
        if (ctx->d_dist[k]) cudaFree(ctx->d_dist[k]);
        if (ctx->h_dist[k]) cudaFreeHost(ctx->h_dist[k]);
        VO_CUDA_CHECK(cudaMalloc(&ctx->d_dist[k], need));
        VO_CUDA_CHECK(cudaMallocHost(&ctx->h_dist[k], need));
    }
    ctx->dist_bytes = need;
    return VO_OK;
}

// all-gather of the result records of resident slots [first_unit, first_unit + n_units) (every rank posts the same n_units),
// asynchronous on the communication stream; up to VO_DIST_DEPTH posts may be outstanding.  The records are copied to their
// place in the ring entry's table first (device to device, this rank only) and gathered IN PLACE from there: the slots are
// free again as soon as that local copy is done, so refilling them never waits for the slowest rank's contribution.
extern "C" int vo_dist_gather_post(vo_ctx* ctx, int first_unit, int n_units)
{
    if (!ctx) return VO_E_INVALID;
    if (!ctx->dist_comm) { vo_set_error(ctx, "vo_dist_gather_post: vo_dist_init first"); return VO_E_INVALID; }
    if (first_unit < 0 || n_units <= 0 || first_unit + n_units > ctx->units) { vo_set_error(ctx, "vo_dist_gather_post: slots outside the state"); return VO_E_INVALID; }
    if (ctx->dist_head - ctx->dist_tail >= VO_DIST_DEPTH) { vo_set_error(ctx, "vo_dist_gather_post: %d gathers are already outstanding (vo_dist_gather_wait)", VO_DIST_DEPTH); return VO_E_INVALID; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc = dist_buffers(ctx, n_units);
    if (rc) return rc;
    const int k = (int)(ctx->dist_head % VO_DIST_DEPTH);
    // the records were written by work that the caller's stream has already been made to wait for (vo_batch_wait / vo_batch_run)
    if (!ctx->dist_ev_fork) VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->dist_ev_fork, cudaEventDisableTiming));
    VO_CUDA_CHECK(cudaEventRecord(ctx->dist_ev_fork, ctx->stream));
    VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->dist_stream, ctx->dist_ev_fork, 0));
    const size_t bytes = (size_t)n_units * sizeof(vo_unit_result_dev);
    uint8_t* table = (uint8_t*)ctx->d_dist[k];
    uint8_t* mine = table + (size_t)ctx->dist_rank * bytes;
    VO_CUDA_CHECK(cudaMemcpyAsync(mine, ctx->d_results + first_unit, bytes, cudaMemcpyDeviceToDevice, ctx->dist_stream));
    VO_CUDA_CHECK(cudaEventRecord(ctx->dist_ev_read[k], ctx->dist_stream));          // the slot's records have been read
    const int nrc = g_nccl.AllGather(mine, table, bytes, /*ncclUint8*/ 1, (NcclComm)ctx->dist_comm, ctx->dist_stream);
    if (nrc != 0) { vo_set_error(ctx, "ncclAllGather: %s", nccl_err(nrc)); return VO_E_CUDA; }
    VO_CUDA_CHECK(cudaMemcpyAsync(ctx->h_dist[k], table, bytes * ctx->dist_world, cudaMemcpyDeviceToHost, ctx->dist_stream));
    VO_CUDA_CHECK(cudaEventRecord(ctx->dist_ev_done[k], ctx->dist_stream));
    ctx->dist_posted[k] = true; ctx->dist_n[k] = n_units;
    ctx->dist_head++;
    return VO_OK;
}

// the oldest outstanding gather: all[r * n_units + i] = record i of rank r's posted slots
extern "C" int vo_dist_gather_wait(vo_ctx* ctx, vo_unit_result* all, int cap_records, int* n_records)
{
    if (!ctx) return VO_E_INVALID;
    if (ctx->dist_head == ctx->dist_tail) { vo_set_error(ctx, "vo_dist_gather_wait: nothing outstanding"); return VO_E_INVALID; }
    const int k = (int)(ctx->dist_tail % VO_DIST_DEPTH);
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    VO_CUDA_CHECK(cudaEventSynchronize(ctx->dist_ev_done[k]));
    const int n = ctx->dist_world * ctx->dist_n[k];
    if (n_records) *n_records = n;
    ctx->dist_tail++;
    if (all) {
        if (cap_records < n) { vo_set_error(ctx, "vo_dist_gather_wait: %d records, room for %d", n, cap_records); return VO_E_CAPACITY; }
        memcpy(all, ctx->h_dist[k], (size_t)n * sizeof(vo_unit_result));
    }
    return VO_OK;
}

// a submission that refills resident slots must not overtake a gather that still reads their records
int vo_dist_order_after_gathers(vo_ctx* ctx, cudaStream_t st)
{
    if (!ctx->dist_comm) return VO_OK;
    for (int k = 0; k < VO_DIST_DEPTH; k++)
        if (ctx->dist_posted[k]) VO_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->dist_ev_read[k], 0));
    return VO_OK;
}

void vo_dist_shutdown(vo_ctx* ctx)
{
    if (ctx->dist_stream) cudaStreamSynchronize(ctx->dist_stream);
    if (ctx->dist_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((NcclComm)ctx->dist_comm);
    ctx->dist_comm = nullptr;
    for (int k = 0; k < VO_DIST_DEPTH; k++) {
        if (ctx->d_dist[k]) cudaFree(ctx->d_dist[k]);
        if (ctx->h_dist[k]) cudaFreeHost(ctx->h_dist[k]);
        if (ctx->dist_ev_read[k]) cudaEventDestroy(ctx->dist_ev_read[k]);
        if (ctx->dist_ev_done[k]) cudaEventDestroy(ctx->dist_ev_done[k]);
        ctx->d_dist[k] = nullptr; ctx->h_dist[k] = nullptr; ctx->dist_ev_read[k] = nullptr; ctx->dist_ev_done[k] = nullptr;
    }
    if (ctx->dist_ev_fork) cudaEventDestroy(ctx->dist_ev_fork);
    if (ctx->dist_stream) cudaStreamDestroy(ctx->dist_stream);
    ctx->dist_ev_fork = nullptr; ctx->dist_stream = nullptr;
}
