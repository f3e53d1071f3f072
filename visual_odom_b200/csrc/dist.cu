// dist.cu -- SURVEY.md 8(e) "C1": the only communication of the multi-GPU layer, the gather of the fixed-size result
// records, as C-ABI calls over NCCL (one process per GPU; units are sharded with no data-path collective).
//
//   rank 0: vo_dist_unique_id(id)  ->  the 128 bytes travel out of band (bench.py broadcasts them with torch.distributed)
//   all   : vo_dist_init(ctx, id, rank, world)
//   loop  : vo_batch_wait(slot) ; vo_dist_gather_post(ctx, slot, n)      non-blocking: device snapshot of the records into a bucket;
//           ...                 ; vo_dist_gather_wait(ctx, all, ...)     one in-place ncclAllGather + one D2H per 4 posts
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
    if (ctx->dist_comm) vo_dist_shutdown(ctx);
    NcclUid uid;
    memcpy(uid.internal, id, 128);
    NcclComm comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (rc != 0) { vo_set_error(ctx, "ncclCommInitRank: %s", nccl_err(rc)); return VO_E_CUDA; }
    ctx->dist_comm = comm; ctx->dist_rank = rank; ctx->dist_world = world;
    VO_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->dist_stream, cudaStreamNonBlocking));
    VO_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->dist_snap_stream, cudaStreamNonBlocking));
    VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->dist_ev_read, cudaEventDisableTiming));
    VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->dist_ev_fork, cudaEventDisableTiming));
    for (auto& bk : ctx->dist_bk) {
        VO_CUDA_CHECK(cudaEventCreateWithFlags(&bk.done, cudaEventDisableTiming));
        bk.fill = bk.unwaited = bk.n_units = 0; bk.flushed = false;
    }
    ctx->dist_cur = 0;
    ctx->dist_head = ctx->dist_tail = 0;
    ctx->dist_bytes = 0;
    return VO_OK;
}

// a bucket's table: [rank][VO_DIST_BUCKET][units of the state] records, device + pinned host; sized for the whole state, so
// it is (re)allocated only when the state grows, which needs every post handed out first
static int dist_buffers(vo_ctx* ctx)
{
    const size_t need = (size_t)ctx->dist_world * VO_DIST_BUCKET * ctx->units * sizeof(vo_unit_result_dev);
    if (need <= ctx->dist_bytes) return VO_OK;
    if (ctx->dist_head != ctx->dist_tail) { vo_set_error(ctx, "vo_dist_gather_post: the state grew while gathers are outstanding"); return VO_E_INVALID; }
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->dist_stream));
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->dist_snap_stream));
    for (auto& bk : ctx->dist_bk) {
        if (bk.d) cudaFree(bk.d);
        if (bk.h) cudaFreeHost(bk.h);
        bk.d = bk.h = nullptr;
        VO_CUDA_CHECK(cudaMalloc(&bk.d, need));
        VO_CUDA_CHECK(cudaMallocHost(&bk.h, need));
        bk.fill = bk.unwaited = 0; bk.flushed = false;
    }
    ctx->dist_bytes = need;
    return VO_OK;
}

// the exchange of one bucket: in-place all-gather of VO_DIST_BUCKET step snapshots per rank (a constant size, whatever
// the fill, so every rank issues the same collective) + one copy of the whole table to pinned memory
static int dist_flush(vo_ctx* ctx, int b)
{
    vo_ctx::DistBucket& bk = ctx->dist_bk[b];
    const size_t per_rank = (size_t)VO_DIST_BUCKET * bk.n_units * sizeof(vo_unit_result_dev);
    uint8_t* table = (uint8_t*)bk.d;
    VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->dist_stream, ctx->dist_ev_read, 0));       // every snapshot of the bucket is in place
    const int nrc = g_nccl.AllGather(table + (size_t)ctx->dist_rank * per_rank, table, per_rank, /*ncclUint8*/ 1, (NcclComm)ctx->dist_comm, ctx->dist_stream);
    if (nrc != 0) { vo_set_error(ctx, "ncclAllGather: %s", nccl_err(nrc)); return VO_E_CUDA; }
    VO_CUDA_CHECK(cudaMemcpyAsync(bk.h, table, per_rank * ctx->dist_world, cudaMemcpyDeviceToHost, ctx->dist_stream));
    VO_CUDA_CHECK(cudaEventRecord(bk.done, ctx->dist_stream));
    bk.flushed = true;
    if (b == ctx->dist_cur) ctx->dist_cur = (ctx->dist_cur + 1) % VO_DIST_NB;      // closed: later posts open the next bucket
    return VO_OK;
}

// Posts the result records of resident slots [first_unit, first_unit + n_units) (every rank posts the same n_units) for the
// gather; asynchronous, up to VO_DIST_DEPTH posts may be outstanding.  The post itself is a device-to-device snapshot into
// the open bucket, so the slots may be refilled at once and no rank ever waits for another one here; the collective runs once
// per VO_DIST_BUCKET posts (or when vo_dist_gather_wait needs a step of a bucket that has not been exchanged yet).
extern "C" int vo_dist_gather_post(vo_ctx* ctx, int first_unit, int n_units)
{
    if (!ctx) return VO_E_INVALID;
    if (!ctx->dist_comm) { vo_set_error(ctx, "vo_dist_gather_post: vo_dist_init first"); return VO_E_INVALID; }
    if (first_unit < 0 || n_units <= 0 || first_unit + n_units > ctx->units) { vo_set_error(ctx, "vo_dist_gather_post: slots outside the state"); return VO_E_INVALID; }
    if (ctx->dist_head - ctx->dist_tail >= VO_DIST_DEPTH) { vo_set_error(ctx, "vo_dist_gather_post: %d gathers are already outstanding (vo_dist_gather_wait)", VO_DIST_DEPTH); return VO_E_INVALID; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc = dist_buffers(ctx);
    if (rc) return rc;
    if (ctx->dist_bk[ctx->dist_cur].fill > 0 && ctx->dist_bk[ctx->dist_cur].n_units != n_units && (rc = dist_flush(ctx, ctx->dist_cur))) return rc;
    const int b = ctx->dist_cur;
    vo_ctx::DistBucket& bk = ctx->dist_bk[b];
    if (bk.fill == 0) {
        if (bk.unwaited > 0) { vo_set_error(ctx, "vo_dist_gather_post: no free bucket (vo_dist_gather_wait)"); return VO_E_INVALID; }
        bk.flushed = false; bk.n_units = n_units;
    }
    // the records were written by work that the caller's stream has already been made to wait for (vo_batch_wait / vo_batch_run)
    // The snapshot runs on its own stream: a collective that is waiting for a slower rank (or for an SM) on the
    // communication stream must not hold back the snapshots behind it -- later submissions wait for those.
    VO_CUDA_CHECK(cudaEventRecord(ctx->dist_ev_fork, ctx->stream));
    VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->dist_snap_stream, ctx->dist_ev_fork, 0));
    const size_t bytes = (size_t)n_units * sizeof(vo_unit_result_dev);
    uint8_t* dst = (uint8_t*)bk.d + ((size_t)ctx->dist_rank * VO_DIST_BUCKET + bk.fill) * bytes;
    VO_CUDA_CHECK(cudaMemcpyAsync(dst, ctx->d_results + first_unit, bytes, cudaMemcpyDeviceToDevice, ctx->dist_snap_stream));
    VO_CUDA_CHECK(cudaEventRecord(ctx->dist_ev_read, ctx->dist_snap_stream));        // the slots' records have been read
    ctx->dist_steps[ctx->dist_head % (2 * VO_DIST_DEPTH)] = vo_ctx::DistStep{b, bk.fill};
    bk.fill++; bk.unwaited++;
    ctx->dist_head++;
    if (bk.fill == VO_DIST_BUCKET) return dist_flush(ctx, b);
    return VO_OK;
}

// the oldest outstanding post: all[r * n_units + i] = record i of rank r's posted slots
extern "C" int vo_dist_gather_wait(vo_ctx* ctx, vo_unit_result* all, int cap_records, int* n_records)
{
    if (!ctx) return VO_E_INVALID;
    if (ctx->dist_head == ctx->dist_tail) { vo_set_error(ctx, "vo_dist_gather_wait: nothing outstanding"); return VO_E_INVALID; }
    const vo_ctx::DistStep st = ctx->dist_steps[ctx->dist_tail % (2 * VO_DIST_DEPTH)];
    vo_ctx::DistBucket& bk = ctx->dist_bk[st.bucket];
    const int n = ctx->dist_world * bk.n_units;
    if (all && cap_records < n) { vo_set_error(ctx, "vo_dist_gather_wait: %d records, room for %d", n, cap_records); return VO_E_CAPACITY; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc;
    if (!bk.flushed && (rc = dist_flush(ctx, st.bucket))) return rc;
    VO_CUDA_CHECK(cudaEventSynchronize(bk.done));
    if (n_records) *n_records = n;
    if (all) {
        const size_t bytes = (size_t)bk.n_units * sizeof(vo_unit_result);
        for (int r = 0; r < ctx->dist_world; r++)
            memcpy((uint8_t*)all + (size_t)r * bytes, (const uint8_t*)bk.h + ((size_t)r * VO_DIST_BUCKET + st.index) * bytes, bytes);
    }
    ctx->dist_tail++;
    if (--bk.unwaited == 0) bk.fill = 0;          // every step handed out: the bucket may be reused
    return VO_OK;
}

// a submission that refills resident slots must not overtake the snapshot of a post that still reads their records
// (the snapshots are ordered on the communication stream, so the latest one stands for all of them)
int vo_dist_order_after_gathers(vo_ctx* ctx, cudaStream_t st)
{
    if (!ctx->dist_comm || ctx->dist_head == 0) return VO_OK;
    VO_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->dist_ev_read, 0));
    return VO_OK;
}

void vo_dist_shutdown(vo_ctx* ctx)
{
    if (ctx->dist_snap_stream) cudaStreamSynchronize(ctx->dist_snap_stream);
    if (ctx->dist_stream) cudaStreamSynchronize(ctx->dist_stream);
    if (ctx->dist_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((NcclComm)ctx->dist_comm);
    ctx->dist_comm = nullptr;
    for (auto& bk : ctx->dist_bk) {
        if (bk.d) cudaFree(bk.d);
        if (bk.h) cudaFreeHost(bk.h);
        if (bk.done) cudaEventDestroy(bk.done);
        bk = vo_ctx::DistBucket();
    }
    if (ctx->dist_ev_read) cudaEventDestroy(ctx->dist_ev_read);
    if (ctx->dist_ev_fork) cudaEventDestroy(ctx->dist_ev_fork);
    if (ctx->dist_stream) cudaStreamDestroy(ctx->dist_stream);
    if (ctx->dist_snap_stream) cudaStreamDestroy(ctx->dist_snap_stream);
    ctx->dist_ev_read = ctx->dist_ev_fork = nullptr; ctx->dist_stream = ctx->dist_snap_stream = nullptr;
    ctx->dist_head = ctx->dist_tail = 0; ctx->dist_cur = 0; ctx->dist_bytes = 0;
}
