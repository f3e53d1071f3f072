// batch.cu -- the batched, device-resident whole-path API (vo_batch_*, vo_frame_batch).
//
// One call runs, for every resident work unit and with no host round trip in between:
//   [FAST on L0 + even-stride selection]  ->  pyramids  ->  LK ring (L0->R0->R1->L1->L0)
//   ->  status / negative-coordinate / circular filters  ->  DLT triangulation  ->  PnP/RANSAC + LM
// i.e. reference src/visualOdometry.cpp:81-129 (matchingFeatures, with the bucketing replaced by
// the benchmark's stride selection), src/main.cpp:170-171 and src/visualOdometry.cpp:132-193.
#include "ctx.h"
#include <string.h>

__global__ void k_pack_counts(vo_unit_result_dev* res, const int* n_pts, const int* n_det, const int* n3, const int* n5,
                              int n_units, int detect)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    res[u].n_features = n_pts[u];
    res[u].n_detected = detect ? n_det[u] : 0;
    res[u].n_tracked = n3[u];
    res[u].n_valid = n5[u];
}

// Packed outputs of one unit: float2 pts4[4][per] (L0, R0, L1, R1 of the valid tracks) | int kept_idx[per] |
// float3 X[per] | int inliers[per], `per` point slots each (only the first n_valid / n_inliers are written).
#define VO_OUT_BYTES_PER_SLOT (4 * sizeof(float2) + sizeof(int) + sizeof(float3) + sizeof(int))
__global__ void k_pack_outputs(uint8_t* __restrict__ out, size_t stride, int per, const float2* __restrict__ valid4, size_t cs,
                               const int* __restrict__ idx5, const float3* __restrict__ X, const int* __restrict__ inliers,
                               const int* __restrict__ n5, const vo_unit_result_dev* __restrict__ res, int cap)
{
    const int u = blockIdx.x;
    const int nv = min(n5[u], per), ni = min(res[u].n_inliers, per);
    float2* o4 = reinterpret_cast<float2*>(out + (size_t)u * stride);
    int* ok = reinterpret_cast<int*>(o4 + 4 * (size_t)per);
    float* oX = reinterpret_cast<float*>(ok + per);
    int* oi = reinterpret_cast<int*>(oX + 3 * (size_t)per);
    const size_t ub = (size_t)u * cap;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 4; k++) o4[(size_t)k * per + i] = valid4[k * cs + ub + i];
        ok[i] = idx5[ub + i];
        const float3 x = X[ub + i];
        oX[3 * i] = x.x; oX[3 * i + 1] = x.y; oX[3 * i + 2] = x.z;
    }
    for (int i = threadIdx.x; i < ni; i += blockDim.x) oi[i] = inliers[ub + i];
}

// device + pinned blocks of the packed outputs, sized for the configured batch and the resident feature counts
static int ensure_outputs(vo_ctx* ctx)
{
    int per = ctx->batch_max_pts > 0 ? ctx->batch_max_pts : ctx->cap;
    per = (per + 63) / 64 * 64;
    if (per > ctx->cap) per = ctx->cap;
    const size_t stride = ((size_t)per * VO_OUT_BYTES_PER_SLOT + 255) / 256 * 256;
    if (ctx->d_out && ctx->out_per == per && ctx->h_out && ctx->out_units >= ctx->units) return VO_OK;
    int rc = vo_drain_pending(ctx);
    if (rc) return rc;
    if (!ctx->d_out || ctx->out_per != per) {
        void* q = nullptr;                      // freed with the batch state (ctx->allocs)
        VO_CUDA_CHECK(cudaMalloc(&q, stride * ctx->units + 256));
        ctx->allocs.push_back(q);
        ctx->d_out = (uint8_t*)q;
        vo_drop_graphs(ctx);                    // graphs hold the old pointer / stride
    }
    if (ctx->h_out) { cudaFreeHost(ctx->h_out); ctx->h_out = nullptr; }
    VO_CUDA_CHECK(cudaMallocHost(&ctx->h_out, stride * ctx->units + 256));
    ctx->out_units = ctx->units; ctx->out_per = per; ctx->out_stride = stride;
    return VO_OK;
}

extern "C" int vo_batch_configure(vo_ctx* ctx, int w, int h, int n_units, const float P_l[12], const float P_r[12])
{
    if (!ctx) return VO_E_INVALID;
    if (!P_l || !P_r || n_units <= 0) { vo_set_error(ctx, "bad argument"); return VO_E_INVALID; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc = vo_claim_buffers(ctx, "vo_batch_configure");
    if (rc) return rc;
    if ((rc = vo_ensure_state(ctx, w, h, n_units, 4))) return rc;
    vo_set_calibration(ctx, P_l, P_r);
    ctx->batch_units = n_units;
    ctx->batch_uploaded = 0;
    return VO_OK;
}

// H2D of units [u0, u0+n) on stream `st`; scalars go through the pinned staging block (disjoint per unit)
static int upload_range(vo_ctx* ctx, const vo_unit* units, int u0, int n, size_t pitch, cudaStream_t st, bool detect, int src0 = -1)
{
    if (src0 < 0) src0 = u0;           // units[src0 + i] fills resident slot u0 + i
    const int w = ctx->w, h = ctx->h, cap = ctx->cap;
    const int total = ctx->batch_units;
    double* h_tprev = (double*)ctx->h_pinned;
    int* h_cnt = (int*)(h_tprev + 3 * (size_t)total);
    for (int u = u0; u < u0 + n; u++) {
        const vo_unit& U = units[u - u0 + src0];
        const uint8_t* imgs[4] = {U.l0, U.r0, U.l1, U.r1};
        for (int k = 0; k < 4; k++) {
            uint8_t* dst = ctx->d_raw + ((size_t)u * 4 + k) * w * h;
            if (pitch == (size_t)w) VO_CUDA_CHECK(cudaMemcpyAsync(dst, imgs[k], (size_t)w * h, cudaMemcpyHostToDevice, st));
            else VO_CUDA_CHECK(cudaMemcpy2DAsync(dst, w, imgs[k], pitch, w, h, cudaMemcpyHostToDevice, st));
        }
        if (!detect && U.n_pts > 0)
            VO_CUDA_CHECK(cudaMemcpyAsync(ctx->d_pts_in + (size_t)u * cap, U.pts, (size_t)U.n_pts * sizeof(float2),
                                          cudaMemcpyHostToDevice, st));
        h_cnt[u] = U.n_pts;
        for (int k = 0; k < 3; k++) h_tprev[3 * u + k] = U.t_prev[k];
    }
    VO_CUDA_CHECK(cudaMemcpyAsync(ctx->d_tprev + 3 * (size_t)u0, h_tprev + 3 * (size_t)u0, (size_t)n * 3 * sizeof(double),
                                  cudaMemcpyHostToDevice, st));
    VO_CUDA_CHECK(cudaMemcpyAsync((detect ? ctx->d_want : ctx->d_npts) + u0, h_cnt + u0, (size_t)n * sizeof(int),
                                  cudaMemcpyHostToDevice, st));
    return VO_OK;
}

static int validate_units(vo_ctx* ctx, const vo_unit* units, int n_units, size_t pitch, bool* detect_out, int* max_pts_out)
{
    if (!ctx || !units) return VO_E_INVALID;
    if (n_units <= 0 || n_units > ctx->batch_units) { vo_set_error(ctx, "n_units=%d outside the configured batch (%d)", n_units, ctx->batch_units); return VO_E_INVALID; }
    if (pitch < (size_t)ctx->w) { vo_set_error(ctx, "pitch %zu < width %d", pitch, ctx->w); return VO_E_INVALID; }
    const bool detect = (units[0].pts == nullptr);
    int max_pts = 0;
    for (int u = 0; u < n_units; u++) {
        const vo_unit& U = units[u];
        if (!U.l0 || !U.r0 || !U.l1 || !U.r1) { vo_set_error(ctx, "unit %d: null image", u); return VO_E_INVALID; }
        if ((U.pts == nullptr) != detect) { vo_set_error(ctx, "units must all carry features or all request detection"); return VO_E_INVALID; }
        if (U.n_pts < 0 || U.n_pts > ctx->cap) { vo_set_error(ctx, "unit %d: n_pts=%d outside [0,%d]", u, U.n_pts, ctx->cap); return VO_E_CAPACITY; }
        if (U.n_pts > max_pts) max_pts = U.n_pts;
    }
    *detect_out = detect; *max_pts_out = max_pts;
    // pinned staging: t_prev, counts, result records
    const size_t bytes = (size_t)ctx->batch_units * (3 * sizeof(double) + sizeof(int) + sizeof(vo_unit_result_dev)) + 256;
    return vo_ensure_pinned(ctx, bytes);
}

static vo_unit_result_dev* pinned_results(vo_ctx* ctx)
{
    char* p = (char*)ctx->h_pinned + (size_t)ctx->batch_units * (3 * sizeof(double) + sizeof(int));
    p = (char*)(((uintptr_t)p + 63) & ~(uintptr_t)63);
    return (vo_unit_result_dev*)p;
}

extern "C" int vo_batch_upload(vo_ctx* ctx, const vo_unit* units, int n_units, size_t pitch)
{
    bool detect; int max_pts;
    int rc = validate_units(ctx, units, n_units, pitch, &detect, &max_pts);
    if (rc) return rc;
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if ((rc = vo_claim_buffers(ctx, "vo_batch_upload", true))) return rc;
    if ((rc = vo_drain_pending(ctx))) return rc;
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));      // staging block may still be in flight
    if ((rc = upload_range(ctx, units, 0, n_units, pitch, ctx->stream, detect))) return rc;
    ctx->batch_uploaded = n_units;
    ctx->batch_detect = detect;
    ctx->batch_max_pts = max_pts;
    return VO_OK;
}

static int ensure_side_streams(vo_ctx* ctx)
{
    if (ctx->fork_ev) return VO_OK;
    VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    for (int c = 0; c < VO_LANES; c++) {
        VO_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->side_stream[c], cudaStreamNonBlocking));
        VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->join_ev[c], cudaEventDisableTiming));
    }
    return VO_OK;
}

// high-priority helper streams: the short and the latency-bound kernels of a range (FAST, pyramids, filters,
// triangulation, PnP) are issued there, the LK ring at normal priority.  With two ranges in flight the helpers' few
// CTAs are then never queued behind the thousands of pending CTAs of the OTHER range's LK launch, so consecutive LK
// launches follow each other directly and the ramp-down of one (a feature-ring lasts ~0.3 ms) fills with the next.
static int ensure_hi_streams(vo_ctx* ctx)
{
    if (ctx->hi_stream[0]) return VO_OK;
    int lo = 0, hi = 0;
    VO_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // hi is the numerically smallest value
    for (int c = 0; c < VO_LANES; c++) {
        VO_CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->hi_stream[c], cudaStreamNonBlocking, hi));
        for (int k = 0; k < 4; k++) VO_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->hi_ev[c][k], cudaEventDisableTiming));
    }
    return VO_OK;
}

// the whole path for the resident units of `v`, asynchronous on v.s.  When v.s is one of the side streams and
// priorities are enabled, everything but the LK ring is forked to that side stream's high-priority helper.
static int run_range_launch(vo_ctx* ctx, const View& v)
{
    ctx->imgs_per_unit = 4;
    int rc;
    int c = -1;
    for (int k = 0; k < VO_LANES; k++) if (ctx->side_stream[k] && v.s == ctx->side_stream[k]) c = k;
    // With the SM partition on, a side-stream range runs its LK ring on the LK partition's stream and everything else on the
    // helper partition's stream; otherwise (priorities) everything but the LK ring goes to the side stream's high-priority helper.
    const bool part = ctx->part_on && c >= 0;
    const bool prio = !part && ctx->use_priorities && c >= 0;
    View h = v, lk = v;                          // the views the helper kernels / the LK ring run on
    cudaEvent_t* ev = nullptr;
    if (part) {
        h.s = ctx->part_hp_stream[c]; lk.s = ctx->part_lk_stream[c]; ev = ctx->part_ev[c];
    } else if (prio) {
        if ((rc = ensure_hi_streams(ctx))) return rc;
        h.s = ctx->hi_stream[c]; ev = ctx->hi_ev[c];
    }
    View pre = h;                                // FAST + pyramids
    if (part && ctx->part_pre_with_lk) pre.s = lk.s;
    if (ev) {
        VO_CUDA_CHECK(cudaEventRecord(ev[0], v.s));
        VO_CUDA_CHECK(cudaStreamWaitEvent(pre.s, ev[0], 0));
        if (pre.s != h.s) VO_CUDA_CHECK(cudaStreamWaitEvent(h.s, ev[0], 0));
    }
    if (ctx->batch_detect) {
        if ((rc = vo_run_fast(ctx, pre, 0, false))) return rc;
        if ((rc = vo_run_select(ctx, pre))) return rc;
    }
    if ((rc = vo_run_pyramid(ctx, v.u0 * ctx->imgs_per_unit, v.n * ctx->imgs_per_unit, pre.s))) return rc;
    if (ev && pre.s != lk.s) {
        VO_CUDA_CHECK(cudaEventRecord(ev[1], pre.s));
        VO_CUDA_CHECK(cudaStreamWaitEvent(lk.s, ev[1], 0));
    }
    const int ip[4] = {0, 1, 3, 2}, in[4] = {1, 3, 2, 0};      // ring L0->R0->R1->L1->L0 (planes L0,R0,L1,R1)
    ctx->lk_per_unit = ctx->batch_max_pts;                      // no unit of the resident batch has more live features
    rc = vo_run_lk_ring(ctx, lk, 4, ip, in, false);
    ctx->lk_per_unit = 0;
    if (rc) return rc;
    if (ev) {
        VO_CUDA_CHECK(cudaEventRecord(ev[2], lk.s));
        VO_CUDA_CHECK(cudaStreamWaitEvent(h.s, ev[2], 0));
    }
    if ((rc = vo_run_filter(ctx, h, false))) return rc;
    const size_t cs = (size_t)ctx->units * ctx->cap;
    if ((rc = vo_run_triangulate(ctx, h, ctx->d_valid4, ctx->d_valid4 + cs, ctx->d_n5))) return rc;
    float K9[9] = {ctx->P_l[0], ctx->P_l[1], ctx->P_l[2], ctx->P_l[4], ctx->P_l[5], ctx->P_l[6], ctx->P_l[8], ctx->P_l[9], ctx->P_l[10]};
    if ((rc = vo_run_pnp(ctx, h, ctx->d_valid4 + 2 * cs, ctx->d_n5, K9))) return rc;
    k_pack_counts<<<(v.n + 63) / 64, 64, 0, h.s>>>(ctx->d_results + v.u0, ctx->d_npts + v.u0, ctx->d_ndet + v.u0, ctx->d_n3 + v.u0,
                                                  ctx->d_n5 + v.u0, v.n, ctx->batch_detect ? 1 : 0);
    ctx->launches += 1;
    VO_CUDA_CHECK(cudaGetLastError());
    if (ev) {                                    // join: later work on v.s (result copy, the next submission) sees everything
        VO_CUDA_CHECK(cudaEventRecord(ev[3], h.s));
        VO_CUDA_CHECK(cudaStreamWaitEvent(v.s, ev[3], 0));
    }
    return VO_OK;
}

// Replays (or first captures) the kernel sequence of one unit range as a CUDA graph on v.s.  All kernel
// arguments are device pointers / sizes fixed by (range, detect, staging), so the graph is reusable until
// the device state is re-allocated.  The LK event timing is not part of graphs.
static int run_range(vo_ctx* ctx, const View& v)
{
    // Kernel nodes of a captured graph do not keep the capture streams' priorities (measured: no effect), and the
    // priority split is worth more (+7 % on the pipelined step) than the graph's launch savings (+2 %): ranges that run
    // on a side stream with priorities enabled are launched plainly unless "batch_graphs" forces graphs.
    bool on_side = false;
    for (int k = 0; k < VO_LANES; k++) on_side = on_side || (ctx->side_stream[k] && v.s == ctx->side_stream[k]);
    if (!ctx->use_graphs || ((ctx->use_priorities || ctx->part_on) && on_side && !ctx->batch_graphs)) return run_range_launch(ctx, v);
    for (auto& g : ctx->graphs)
        if (g.u0 == v.u0 && g.n == v.n && g.detect == ctx->batch_detect && g.tma == ctx->lk_use_tma && g.s == v.s && g.max_pts == ctx->batch_max_pts) {
            VO_CUDA_CHECK(cudaGraphLaunch(g.exec, v.s));
            ctx->launches += g.launches;
            return VO_OK;
        }
    const bool timing = ctx->lk_timing;
    const long long before = ctx->launches;
    ctx->lk_timing = false;
    cudaGraph_t graph = nullptr;
    VO_CUDA_CHECK(cudaStreamBeginCapture(v.s, cudaStreamCaptureModeThreadLocal));
    int rc = run_range_launch(ctx, v);
    cudaError_t e = cudaStreamEndCapture(v.s, &graph);
    ctx->lk_timing = timing;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    VO_CUDA_CHECK(e);
    vo_ctx::RangeGraph g;
    g.u0 = v.u0; g.n = v.n; g.detect = ctx->batch_detect; g.tma = ctx->lk_use_tma; g.s = v.s; g.max_pts = ctx->batch_max_pts;      // the LK launch geometry depends on max_pts
    g.launches = ctx->launches - before;
    VO_CUDA_CHECK(cudaGraphInstantiate(&g.exec, graph, 0));
    cudaGraphDestroy(graph);
    ctx->graphs.push_back(g);
    VO_CUDA_CHECK(cudaGraphLaunch(g.exec, v.s));
    return VO_OK;
}

extern "C" int vo_batch_run(vo_ctx* ctx)
{
    if (!ctx) return VO_E_INVALID;
    const int units = ctx->batch_uploaded;
    if (units <= 0) { vo_set_error(ctx, "vo_batch_run: nothing uploaded"); return VO_E_INVALID; }
    if (!ctx->have_P) { vo_set_error(ctx, "vo_batch_run: projection matrices not set"); return VO_E_INVALID; }
    { int rcc = vo_claim_buffers(ctx, "vo_batch_run"); if (rcc) return rcc; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (units < 2 || ctx->batch_streams < 2) return run_range(ctx, View{0, units, ctx->stream});
    // two unit ranges on two side streams: the latency-bound PnP kernels of one range run under the
    // LK ring of the other (fork from / join into the context's stream, so callers see one stream)
    int rc = ensure_side_streams(ctx);
    if (rc) return rc;
    VO_CUDA_CHECK(cudaEventRecord(ctx->fork_ev, ctx->stream));
    const int half = (units + 1) / 2;
    for (int c = 0; c < 2; c++) {
        const int u0 = c ? half : 0, n = c ? units - half : half;
        cudaStream_t st = ctx->side_stream[c];
        VO_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->fork_ev, 0));
        if ((rc = run_range(ctx, View{u0, n, st}))) return rc;
        VO_CUDA_CHECK(cudaEventRecord(ctx->join_ev[c], st));
        VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[c], 0));
    }
    return VO_OK;
}

extern "C" int vo_batch_download(vo_ctx* ctx, vo_unit_result* results, int n_units)
{
    if (!ctx || !results) return VO_E_INVALID;
    if (n_units <= 0 || n_units > ctx->batch_uploaded) { vo_set_error(ctx, "n_units=%d outside the resident batch (%d)", n_units, ctx->batch_uploaded); return VO_E_INVALID; }
    static_assert(sizeof(vo_unit_result) == sizeof(vo_unit_result_dev), "result record layout");
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    VO_CUDA_CHECK(cudaMemcpyAsync(results, ctx->d_results, (size_t)n_units * sizeof(vo_unit_result), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}

// End-to-end entry point: H2D + whole path + D2H.  The batch is split into two unit ranges on two
// side streams forked from / joined into the context's stream, so the H2D of the second range runs
// under the kernels of the first (the images arrive over PCIe; compute is ~6x longer than the copy).
extern "C" int vo_frame_batch(vo_ctx* ctx, const vo_unit* units, int n_units, size_t pitch, vo_unit_result* results)
{
    bool detect; int max_pts;
    int rc = validate_units(ctx, units, n_units, pitch, &detect, &max_pts);
    if (rc) return rc;
    if (!results) return VO_E_INVALID;
    if (!ctx->have_P) { vo_set_error(ctx, "vo_frame_batch: projection matrices not set"); return VO_E_INVALID; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if ((rc = vo_claim_buffers(ctx, "vo_frame_batch", true))) return rc;
    if ((rc = vo_drain_pending(ctx))) return rc;
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->batch_uploaded = n_units; ctx->batch_detect = detect; ctx->batch_max_pts = max_pts;
    vo_unit_result_dev* h_res = pinned_results(ctx);
    const int nchunks = (n_units >= 2 && ctx->batch_streams >= 2) ? 2 : 1;
    if (nchunks == 1) {
        if ((rc = upload_range(ctx, units, 0, n_units, pitch, ctx->stream, detect))) return rc;
        if ((rc = run_range(ctx, View{0, n_units, ctx->stream}))) return rc;
        VO_CUDA_CHECK(cudaMemcpyAsync(h_res, ctx->d_results, (size_t)n_units * sizeof(vo_unit_result_dev), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        if ((rc = ensure_side_streams(ctx))) return rc;
        VO_CUDA_CHECK(cudaEventRecord(ctx->fork_ev, ctx->stream));
        const int half = (n_units + 1) / 2;
        for (int c = 0; c < 2; c++) {
            const int u0 = c ? half : 0, n = c ? n_units - half : half;
            cudaStream_t st = ctx->side_stream[c];
            VO_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->fork_ev, 0));
            if ((rc = upload_range(ctx, units, u0, n, pitch, st, detect))) return rc;
            if ((rc = run_range(ctx, View{u0, n, st}))) return rc;
            VO_CUDA_CHECK(cudaMemcpyAsync(h_res + u0, ctx->d_results + u0, (size_t)n * sizeof(vo_unit_result_dev), cudaMemcpyDeviceToHost, st));
            VO_CUDA_CHECK(cudaEventRecord(ctx->join_ev[c], st));
            VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[c], 0));
        }
    }
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    memcpy(results, h_res, (size_t)n_units * sizeof(vo_unit_result));
    return VO_OK;
}

// ---- pipelined submissions -------------------------------------------------------------------------------------
// A submission fills the resident unit slots [first_unit, first_unit + n_units), runs the whole path on them and
// copies their result records to pinned staging, all asynchronously on one of the two side streams (alternating).
// Submissions on disjoint slot ranges overlap on the GPU: the H2D copy, FAST / pyramids and above all the
// latency-bound PnP tail (a few warps for ~0.4 ms) of one run under the issue-bound LK ring of the other, which a
// synchronous vo_frame_batch per batch cannot do for its last unit range.
extern "C" int vo_batch_submit(vo_ctx* ctx, const vo_unit* units, int first_unit, int n_units, size_t pitch)
{
    if (!ctx) return VO_E_INVALID;
    if (first_unit < 0 || n_units <= 0 || first_unit + n_units > ctx->batch_units) {
        vo_set_error(ctx, "vo_batch_submit: slots [%d, %d) outside the configured batch (%d)", first_unit, first_unit + n_units, ctx->batch_units);
        return VO_E_INVALID;
    }
    if (!ctx->have_P) { vo_set_error(ctx, "vo_batch_submit: projection matrices not set"); return VO_E_INVALID; }
    { int rcc = vo_claim_buffers(ctx, "vo_batch_submit", true); if (rcc) return rcc; }
    for (auto& p : ctx->pending)
        if (p.active && first_unit < p.u0 + p.n && p.u0 < first_unit + n_units) {
            vo_set_error(ctx, "vo_batch_submit: slots [%d, %d) overlap a submission that has not been waited for", first_unit, first_unit + n_units);
            return VO_E_INVALID;
        }
    bool detect = ctx->batch_detect; int max_pts = ctx->batch_max_pts;
    int rc;
    if (units) {
        if ((rc = validate_units(ctx, units, n_units, pitch, &detect, &max_pts))) return rc;
    } else {
        if (ctx->batch_uploaded < first_unit + n_units) {
            vo_set_error(ctx, "vo_batch_submit: units == NULL but slots [%d, %d) were never uploaded", first_unit, first_unit + n_units);
            return VO_E_INVALID;
        }
        const size_t bytes = (size_t)ctx->batch_units * (3 * sizeof(double) + sizeof(int) + sizeof(vo_unit_result_dev)) + 256;
        if ((rc = vo_ensure_pinned(ctx, bytes))) return rc;
    }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if ((rc = ensure_side_streams(ctx))) return rc;
    if (ctx->part_auto) {
        // Pipelined submissions: the persistent LK ring of one range would keep the latency-bound kernels after the other
        // range's ring (filters, triangulation, PnP) off the SMs until it ends.  8 SMs are set aside for them (green
        // contexts); FAST / pyramids stay with the ring.  Measured: value 4239 -> 4603, e2e 3854 -> 4540 frames/s.
        ctx->part_auto = false;
        ctx->part_pre_with_lk = true;
        if (vo_partition_enable(ctx, 8) != VO_OK) ctx->err[0] = 0;      // no green contexts on this driver: run unpartitioned
    }
    if (max_pts > ctx->batch_max_pts || units) ctx->batch_max_pts = max_pts;
    if (ctx->batch_outputs && (rc = ensure_outputs(ctx))) return rc;
    vo_ctx::Pending* slot = nullptr;
    for (auto& p : ctx->pending) if (!p.active) { slot = &p; break; }
    if (!slot) {
        ctx->pending.emplace_back();
        slot = &ctx->pending.back();
        VO_CUDA_CHECK(cudaEventCreateWithFlags(&slot->done, cudaEventDisableTiming));
    }
    const int c = (int)((ctx->submit_count++) % VO_LANES);     // up to VO_LANES submissions in flight, each on its own lane
    cudaStream_t st = ctx->side_stream[c];
    VO_CUDA_CHECK(cudaEventRecord(ctx->fork_ev, ctx->stream));
    VO_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->fork_ev, 0));
    if ((rc = vo_dist_order_after_gathers(ctx, st))) return rc;
    ctx->batch_detect = detect;
    if (max_pts > ctx->batch_max_pts || units) ctx->batch_max_pts = max_pts;
    if (units) {
        if ((rc = upload_range(ctx, units, first_unit, n_units, pitch, st, detect, 0))) return rc;
        if (ctx->batch_uploaded < first_unit + n_units) ctx->batch_uploaded = first_unit + n_units;
    }
    if ((rc = run_range(ctx, View{first_unit, n_units, st}))) return rc;
    vo_unit_result_dev* h_res = pinned_results(ctx);
    VO_CUDA_CHECK(cudaMemcpyAsync(h_res + first_unit, ctx->d_results + first_unit, (size_t)n_units * sizeof(vo_unit_result_dev), cudaMemcpyDeviceToHost, st));
    if (ctx->batch_outputs) {                    // the point lists of the submission: one packed block, one copy
        const size_t cs = (size_t)ctx->units * ctx->cap, ub = (size_t)first_unit * ctx->cap;
        k_pack_outputs<<<n_units, 256, 0, st>>>(ctx->d_out + (size_t)first_unit * ctx->out_stride, ctx->out_stride, ctx->out_per,
                                               ctx->d_valid4 + ub, cs, ctx->d_idx5 + ub, ctx->d_X + ub, ctx->d_inliers + ub,
                                               ctx->d_n5 + first_unit, ctx->d_results + first_unit, ctx->cap);
        ctx->launches += 1;
        VO_CUDA_CHECK(cudaGetLastError());
        VO_CUDA_CHECK(cudaMemcpyAsync(ctx->h_out + (size_t)first_unit * ctx->out_stride, ctx->d_out + (size_t)first_unit * ctx->out_stride,
                                      (size_t)n_units * ctx->out_stride, cudaMemcpyDeviceToHost, st));
    }
    VO_CUDA_CHECK(cudaEventRecord(slot->done, st));
    slot->u0 = first_unit; slot->n = n_units; slot->active = true;
    return VO_OK;
}

extern "C" int vo_batch_wait(vo_ctx* ctx, int first_unit, int n_units, vo_unit_result* results)
{
    if (!ctx) return VO_E_INVALID;
    for (auto& p : ctx->pending)
        if (p.active && p.u0 == first_unit && p.n == n_units) {
            VO_CUDA_CHECK(cudaSetDevice(ctx->device));
            VO_CUDA_CHECK(cudaEventSynchronize(p.done));
            VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, p.done, 0));     // later work on the caller's stream sees the results
            p.active = false;
            if (results) memcpy(results, pinned_results(ctx) + first_unit, (size_t)n_units * sizeof(vo_unit_result));
            return VO_OK;
        }
    vo_set_error(ctx, "vo_batch_wait: no pending submission for slots [%d, %d)", first_unit, first_unit + n_units);
    return VO_E_INVALID;
}

// The point lists of a waited submission, from the pinned block its single D2H copy filled ("batch_outputs" = 1):
// pts4 = [4][n_valid] (L0, R0, L1, R1), kept_idx / X = [n_valid], inliers = [n_inliers]; counts are in the unit's record.
extern "C" int vo_batch_outputs(vo_ctx* ctx, int unit, vo_point2f* pts4, int32_t* kept_idx, vo_point3f* X, int32_t* inliers,
                                size_t* d2h_bytes_per_unit)
{
    if (!ctx) return VO_E_INVALID;
    if (!ctx->batch_outputs || !ctx->h_out) { vo_set_error(ctx, "vo_batch_outputs: option batch_outputs is off"); return VO_E_INVALID; }
    if (unit < 0 || unit >= ctx->batch_uploaded || unit >= ctx->out_units) { vo_set_error(ctx, "unit %d outside the resident batch", unit); return VO_E_INVALID; }
    for (auto& p : ctx->pending)
        if (p.active && unit >= p.u0 && unit < p.u0 + p.n) { vo_set_error(ctx, "vo_batch_outputs: unit %d has not been waited for", unit); return VO_E_INVALID; }
    const vo_unit_result_dev& r = pinned_results(ctx)[unit];
    const int per = ctx->out_per;
    const int nv = r.n_valid < per ? r.n_valid : per, ni = r.n_inliers < per ? r.n_inliers : per;
    const uint8_t* base = ctx->h_out + (size_t)unit * ctx->out_stride;
    const float2* o4 = (const float2*)base;
    const int* ok = (const int*)(o4 + 4 * (size_t)per);
    const float* oX = (const float*)(ok + per);
    const int* oi = (const int*)(oX + 3 * (size_t)per);
    if (pts4) for (int k = 0; k < 4; k++) memcpy(pts4 + (size_t)k * nv, o4 + (size_t)k * per, (size_t)nv * sizeof(float2));
    if (kept_idx) memcpy(kept_idx, ok, (size_t)nv * sizeof(int));
    if (X) memcpy(X, oX, (size_t)nv * 3 * sizeof(float));
    if (inliers) memcpy(inliers, oi, (size_t)ni * sizeof(int));
    if (d2h_bytes_per_unit) *d2h_bytes_per_unit = ctx->out_stride;
    return VO_OK;
}

extern "C" int vo_batch_fetch(vo_ctx* ctx, int unit, vo_point2f* pts_in, vo_point2f* pts4, int32_t* kept_idx,
                              vo_point3f* X, int32_t* inliers)
{
    if (!ctx) return VO_E_INVALID;
    if (unit < 0 || unit >= ctx->batch_uploaded) { vo_set_error(ctx, "unit %d outside the resident batch", unit); return VO_E_INVALID; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    vo_unit_result_dev r;
    VO_CUDA_CHECK(cudaMemcpyAsync(&r, ctx->d_results + unit, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const size_t cs = (size_t)ctx->units * ctx->cap, ub = (size_t)unit * ctx->cap;
    if (pts_in && r.n_features > 0)
        VO_CUDA_CHECK(cudaMemcpyAsync(pts_in, ctx->d_pts_in + ub, (size_t)r.n_features * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    if (pts4 && r.n_valid > 0)
        for (int k = 0; k < 4; k++)
            VO_CUDA_CHECK(cudaMemcpyAsync(pts4 + (size_t)k * r.n_valid, ctx->d_valid4 + k * cs + ub, (size_t)r.n_valid * sizeof(float2),
                                          cudaMemcpyDeviceToHost, ctx->stream));
    if (kept_idx && r.n_valid > 0)
        VO_CUDA_CHECK(cudaMemcpyAsync(kept_idx, ctx->d_idx5 + ub, (size_t)r.n_valid * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (X && r.n_valid > 0)
        VO_CUDA_CHECK(cudaMemcpyAsync(X, ctx->d_X + ub, (size_t)r.n_valid * sizeof(float3), cudaMemcpyDeviceToHost, ctx->stream));
    if (inliers && r.n_inliers > 0)
        VO_CUDA_CHECK(cudaMemcpyAsync(inliers, ctx->d_inliers + ub, (size_t)r.n_inliers * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}
