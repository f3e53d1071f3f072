// seq_api.cu -- C-ABI of the streaming sequence mode (include/vo_b200.h, vo_seq_*).
#include "ctx.h"
#include <string.h>

static int upload_pair(vo_ctx* ctx, int slot, const uint8_t* left, const uint8_t* right, size_t pitch, int channels)
{
    const int w = ctx->w, h = ctx->h;
    const uint8_t* imgs[2] = {left, right};
    if (channels == 3) {                 // colour input: BGR bytes to the staging area, converted inside the frame's graph
        const size_t img = (size_t)3 * w * h;
        int rc = vo_ensure_bgr(ctx, 2 * img);
        if (rc) return rc;
        for (int k = 0; k < 2; k++)
            VO_CUDA_CHECK(cudaMemcpy2DAsync(ctx->d_bgr + k * img, (size_t)3 * w, imgs[k], pitch, (size_t)3 * w, h, cudaMemcpyHostToDevice, ctx->stream));
        return VO_OK;
    }
    for (int k = 0; k < 2; k++) {
        uint8_t* dst = ctx->d_raw + (size_t)(2 * slot + k) * w * h;
        if (pitch == (size_t)w) VO_CUDA_CHECK(cudaMemcpyAsync(dst, imgs[k], (size_t)w * h, cudaMemcpyHostToDevice, ctx->stream));
        else VO_CUDA_CHECK(cudaMemcpy2DAsync(dst, w, imgs[k], pitch, w, h, cudaMemcpyHostToDevice, ctx->stream));
    }
    return VO_OK;
}

// BGR staging -> the two raw gray planes of `slot` (cv::cvtColor(BGR2GRAY) of utils.cpp:179,189)
static int convert_pair(vo_ctx* ctx, int slot)
{
    const int w = ctx->w, h = ctx->h;
    ctx->launches += vo_launch_bgr_to_gray(ctx->d_bgr, (size_t)3 * w, (size_t)3 * w * h, ctx->d_raw + (size_t)(2 * slot) * w * h,
                                           (size_t)w * h, w, h, 2, ctx->stream);
    VO_CUDA_CHECK(cudaGetLastError());
    return VO_OK;
}

// the kernel sequence of one pushed frame; s0 = slot of the previous pair (planes 2*s0, 2*s0+1)
static int seq_launch(vo_ctx* ctx, int s0, bool bgr)
{
    ctx->imgs_per_unit = 4;
    const int s1 = 1 - s0;
    const int L0 = 2 * s0, R0 = 2 * s0 + 1, L1 = 2 * s1, R1 = 2 * s1 + 1;
    const View v{0, 1, ctx->stream};
    int rc;
    if (bgr && (rc = convert_pair(ctx, s1))) return rc;
    // the new pair's two pyramids (the previous pair's are already resident)
    if ((rc = vo_run_pyramid(ctx, 2 * s1, 2, ctx->stream))) return rc;
    // matchingFeatures(): FAST refill on the t0 left image -> bucketing -> circular matching -> filters
    if ((rc = vo_run_fast(ctx, v, L0, false))) return rc;
    SeqArgs a;
    memset(&a, 0, sizeof(a));
    a.corners = ctx->d_corners; a.n_det = ctx->d_ndet; a.corner_cap = ctx->corner_cap;
    a.feat_pts = ctx->d_feat_pts; a.feat_ages = ctx->d_feat_ages; a.cnt = ctx->d_feat_cnt; a.feat_cap = ctx->feat_cap;
    a.refill_below = 2000;                                   // visualOdometry.cpp:95
    a.rows = ctx->h; a.cols = ctx->w; a.bucket_size = ctx->h / 10;      // visualOdometry.cpp:106 (features_per_bucket = 1)
    a.bucket = ctx->d_bucket; a.bucket_cap = ctx->bucket_cap;
    a.out_pts = ctx->d_pts_in; a.out_ages = ctx->d_ages_in; a.out_n = ctx->d_npts; a.out_cap = ctx->cap;
    const size_t cs = (size_t)ctx->units * ctx->cap;
    a.valid_l1 = ctx->d_valid4 + 2 * cs; a.n5 = ctx->d_n5; a.ages_out = ctx->d_ages_out; a.n3 = ctx->d_n3;
    a.res = ctx->d_results; a.tprev = ctx->d_tprev; a.err = ctx->d_seq_err;
    ctx->launches += vo_launch_seq_append(a, ctx->stream);
    ctx->launches += vo_launch_seq_bucket(a, ctx->stream);
    const int ip[4] = {L0, R0, R1, L1}, in[4] = {R0, R1, L1, L0};
    if ((rc = vo_run_lk_ring(ctx, v, 4, ip, in, false))) return rc;
    if ((rc = vo_run_filter(ctx, v, true))) return rc;
    // triangulation + pose
    if ((rc = vo_run_triangulate(ctx, v, ctx->d_valid4, ctx->d_valid4 + cs, ctx->d_n5))) return rc;
    float K9[9] = {ctx->P_l[0], ctx->P_l[1], ctx->P_l[2], ctx->P_l[4], ctx->P_l[5], ctx->P_l[6], ctx->P_l[8], ctx->P_l[9], ctx->P_l[10]};
    if ((rc = vo_run_pnp(ctx, v, ctx->d_valid4 + 2 * cs, ctx->d_n5, K9))) return rc;
    // state carry: features.points = pointsLeft_t1, ages keep their A3 length, translation = tvec; counts -> result record
    ctx->launches += vo_launch_seq_update(a, ctx->stream);
    VO_CUDA_CHECK(cudaGetLastError());
    return VO_OK;
}

// replay (or first capture) the frame's kernel sequence as a CUDA graph; one graph per slot parity
static int seq_run(vo_ctx* ctx, int s0, bool bgr)
{
    if (!ctx->use_graphs) return seq_launch(ctx, s0, bgr);
    const int key = -1 - s0 - (bgr ? 2 : 0);
    for (auto& g : ctx->graphs)
        if (g.u0 == key && g.tma == ctx->lk_use_tma) {
            VO_CUDA_CHECK(cudaGraphLaunch(g.exec, ctx->stream));
            ctx->launches += g.launches;
            return VO_OK;
        }
    const bool timing = ctx->lk_timing;
    const long long before = ctx->launches;
    ctx->lk_timing = false;
    cudaGraph_t graph = nullptr;
    VO_CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = seq_launch(ctx, s0, bgr);
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->lk_timing = timing;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    VO_CUDA_CHECK(e);
    vo_ctx::RangeGraph g;
    g.u0 = key; g.n = 1; g.detect = true; g.tma = ctx->lk_use_tma;
    g.launches = ctx->launches - before;
    VO_CUDA_CHECK(cudaGraphInstantiate(&g.exec, graph, 0));
    cudaGraphDestroy(graph);
    ctx->graphs.push_back(g);
    VO_CUDA_CHECK(cudaGraphLaunch(g.exec, ctx->stream));
    return VO_OK;
}

extern "C" int vo_seq_begin(vo_ctx* ctx, int w, int h, const float P_l[12], const float P_r[12], const uint8_t* left0,
                            const uint8_t* right0, size_t pitch)
{
    return vo_seq_begin_ex(ctx, w, h, P_l, P_r, left0, right0, pitch, 1);
}

extern "C" int vo_seq_begin_ex(vo_ctx* ctx, int w, int h, const float P_l[12], const float P_r[12], const uint8_t* left0,
                               const uint8_t* right0, size_t pitch, int channels)
{
    if (!ctx) return VO_E_INVALID;
    if (channels != 1 && channels != 3) { vo_set_error(ctx, "vo_seq_begin: channels must be 1 (gray) or 3 (BGR)"); return VO_E_INVALID; }
    if (!P_l || !P_r || !left0 || !right0 || w <= 0 || h <= 0 || pitch < (size_t)w * channels) { vo_set_error(ctx, "vo_seq_begin: bad argument"); return VO_E_INVALID; }
    if (h / 10 <= 0) { vo_set_error(ctx, "vo_seq_begin: image too small for the rows/10 bucket size"); return VO_E_UNSUPPORTED; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc = vo_ensure_state(ctx, w, h, 1, 4);
    if (rc) return rc;
    memcpy(ctx->P_l, P_l, 12 * sizeof(float));
    memcpy(ctx->P_r, P_r, 12 * sizeof(float));
    ctx->have_P = true;
    ctx->imgs_per_unit = 4;
    ctx->seq_slot = 0;
    ctx->seq_frames = 0;
    for (int i = 0; i < 16; i++) ctx->seq_pose[i] = (i % 5 == 0) ? 1.0 : 0.0;
    VO_CUDA_CHECK(cudaMemsetAsync(ctx->d_feat_cnt, 0, 2 * sizeof(int), ctx->stream));
    VO_CUDA_CHECK(cudaMemsetAsync(ctx->d_seq_err, 0, sizeof(int), ctx->stream));
    VO_CUDA_CHECK(cudaMemsetAsync(ctx->d_tprev, 0, 3 * sizeof(double), ctx->stream));      // translation = zeros (main.cpp:82)
    if ((rc = upload_pair(ctx, 0, left0, right0, pitch, channels))) return rc;
    if (channels == 3 && (rc = convert_pair(ctx, 0))) return rc;
    if ((rc = vo_run_pyramid(ctx, 0, 2, ctx->stream))) return rc;
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->seq_active = true;
    return VO_OK;
}

extern "C" int vo_seq_push(vo_ctx* ctx, const uint8_t* left1, const uint8_t* right1, size_t pitch, vo_unit_result* out,
                           vo_point2f* pts4, int pts_cap)
{
    return vo_seq_push_ex(ctx, left1, right1, pitch, 1, out, pts4, pts_cap);
}

extern "C" int vo_seq_push_ex(vo_ctx* ctx, const uint8_t* left1, const uint8_t* right1, size_t pitch, int channels,
                              vo_unit_result* out, vo_point2f* pts4, int pts_cap)
{
    if (!ctx) return VO_E_INVALID;
    if (channels != 1 && channels != 3) { vo_set_error(ctx, "vo_seq_push: channels must be 1 (gray) or 3 (BGR)"); return VO_E_INVALID; }
    if (!ctx->seq_active) { vo_set_error(ctx, "vo_seq_push: call vo_seq_begin first"); return VO_E_INVALID; }
    if (!left1 || !right1 || !out || pitch < (size_t)ctx->w * channels) { vo_set_error(ctx, "vo_seq_push: bad argument"); return VO_E_INVALID; }
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int s0 = ctx->seq_slot, s1 = 1 - s0;
    int rc;
    // new stereo pair -> device; then the frame's kernel sequence (a CUDA graph per slot parity)
    if ((rc = upload_pair(ctx, s1, left1, right1, pitch, channels))) return rc;
    if ((rc = seq_run(ctx, s0, channels == 3))) return rc;

    // one pinned read-back: result record (counts packed by k_seq_update) + sticky error bits
    if ((rc = vo_ensure_pinned(ctx, sizeof(vo_unit_result_dev) + 16))) return rc;
    vo_unit_result_dev* hr = (vo_unit_result_dev*)ctx->h_pinned;
    int* herr = (int*)((char*)ctx->h_pinned + sizeof(vo_unit_result_dev));
    VO_CUDA_CHECK(cudaMemcpyAsync(hr, ctx->d_results, sizeof(*hr), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA_CHECK(cudaMemcpyAsync(herr, ctx->d_seq_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    const size_t cs = (size_t)ctx->units * ctx->cap;
    if (pts4 && pts_cap > 0) {
        // n_valid is not known on the host yet: copy up to min(cap, pts_cap) slots of each list in the same batch
        const int n = ctx->cap < pts_cap ? ctx->cap : pts_cap;
        for (int k = 0; k < 4; k++)
            VO_CUDA_CHECK(cudaMemcpyAsync(pts4 + (size_t)k * pts_cap, ctx->d_valid4 + k * cs, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    }
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    const vo_unit_result_dev r = *hr;
    const int err = *herr;
    memcpy(out, &r, sizeof(r));
    ctx->seq_slot = s1;                 // imageLeft_t0 = imageLeft_t1 (main.cpp:157-158)
    ctx->seq_frames++;
    if (r.pnp_status == VO_OK) vo_pose_step(ctx->seq_pose, r.R, r.tvec);   // main.cpp:196-208
    if (err) { vo_set_error(ctx, "vo_seq_push: glue kernel error bits 0x%x (1/2: capacity, 4: bucket grid, 8: feature outside the image)", err); return VO_E_CAPACITY; }
    return VO_OK;
}

extern "C" int vo_seq_state(vo_ctx* ctx, vo_point2f* points, int32_t* ages, int cap, int* n_points, int* n_ages, double t_out[3])
{
    if (!ctx || !ctx->seq_active) return VO_E_INVALID;
    VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int cnt[2] = {0, 0};
    VO_CUDA_CHECK(cudaMemcpyAsync(cnt, ctx->d_feat_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (t_out) VO_CUDA_CHECK(cudaMemcpyAsync(t_out, ctx->d_tprev, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (n_points) *n_points = cnt[0];
    if (n_ages) *n_ages = cnt[1];
    if (points && cnt[0] > 0) VO_CUDA_CHECK(cudaMemcpyAsync(points, ctx->d_feat_pts, (size_t)(cnt[0] < cap ? cnt[0] : cap) * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    if (ages && cnt[1] > 0) VO_CUDA_CHECK(cudaMemcpyAsync(ages, ctx->d_feat_ages, (size_t)(cnt[1] < cap ? cnt[1] : cap) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}
