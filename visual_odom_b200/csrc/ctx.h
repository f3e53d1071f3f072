// ctx.h -- the library context (device-resident state shared by every entry point)
#pragma once
#include "common.cuh"
#include "lk_ring.h"
#include "filter.h"
#include "fast.h"
#include "pnp.h"
#include "seq.h"
#include "ess.h"
#include "../../include/vo_b200.h"
#include <vector>
#include <stdarg.h>
#define LK_QUEUES 32
#define VO_DIST_BUCKET 4       // posted steps per collective
#define VO_DIST_NB 4           // buckets (ring)
#define VO_LANES 3            // submissions in flight, each with its own side stream / partition streams / events

struct vo_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    vo_params p;
    char err[1024] = {0};
    long long launches = 0;
    int sm_count = 0;

    // ---- geometry of the currently allocated batch state -----------------------------------
    int w = 0, h = 0;             // image size
    int units = 0;                // allocated work-unit slots
    int imgs_per_unit = 4;
    int cap = 0;                  // feature capacity per unit
    PyrGeom pg;                   // device plane pointers per level
    LkMaps maps;                  // TMA descriptors per level
    bool have_P = false;
    float P_l[12], P_r[12];

    // ---- device buffers ---------------------------------------------------------------------
    uint8_t* d_raw = nullptr;           // [units*4][h*w] raw images
    const uint8_t** d_raw_tab = nullptr;// [units*4] pointers into d_raw (or caller device images)
    float2* d_pts_in = nullptr;         // [units][cap]
    int* d_npts = nullptr;              // [units]
    float2* d_pts_out = nullptr;        // [4][units][cap]
    uint8_t* d_status = nullptr;        // [4][units][cap]
    float* d_err = nullptr;             // [4][units][cap]
    int* d_ages_in = nullptr;           // [units][cap]
    int* d_ages_out = nullptr;          // [units][cap]
    float2* d_kept5 = nullptr;          // [5][units][cap]
    int* d_idx3 = nullptr;              // [units][cap]
    int* d_n3 = nullptr;                // [units]
    float2* d_valid4 = nullptr;         // [4][units][cap]
    int* d_idx5 = nullptr;              // [units][cap]
    int* d_n5 = nullptr;                // [units]
    // FAST
    uint8_t* d_score = nullptr;         // [units][h*w]
    uint16_t* d_rowbuf = nullptr;       // [units][h][w]
    int* d_rowcount = nullptr;          // [units][h]
    int* d_rowoff = nullptr;            // [units][h]
    int* d_ndet = nullptr;              // [units]
    float2* d_corners = nullptr;        // [units][corner_cap]
    float* d_resp = nullptr;            // [units][corner_cap]
    int* d_want = nullptr;              // [units] features to select (batched path)
    int corner_cap = 0;
    // triangulation + PnP
    float3* d_X = nullptr;              // [units][cap]
    double* d_tprev = nullptr;          // [units][3]
    PnpState* d_pnp_state = nullptr;    // [units]
    int* d_subsets = nullptr;           // [units][iters][5]
    double* d_models = nullptr;         // [units][iters][12]
    int* d_counts = nullptr;            // [units][iters]
    int* d_inliers = nullptr;           // [units][cap]
    vo_unit_result_dev* d_results = nullptr;   // [units]
    // sequence mode (seq.cu): the reference main loop's state, device resident
    float2* d_feat_pts = nullptr;       // [feat_cap] currentVOFeatures.points
    int* d_feat_ages = nullptr;         // [feat_cap] currentVOFeatures.ages (may be longer than points)
    int* d_feat_cnt = nullptr;          // [2] sizes of the two vectors
    int* d_bucket = nullptr;            // [bucket_cap] scratch of bucketingFeatures
    int* d_seq_err = nullptr;           // sticky error bits of the glue kernels
    int feat_cap = 0, bucket_cap = 4096;
    bool seq_active = false;
    int seq_slot = 0;                   // raw/pyramid planes (2*slot, 2*slot+1) hold the previous stereo pair
    int seq_inflight = 0;               // frames submitted and not yet waited for (<= 2)
    long long seq_submitted = 0;        // frames submitted since vo_seq_begin (frame k uses buffer unit k & 1)
    int seq_channels[2] = {1, 1};
    cudaEvent_t seq_front_ev[2] = {nullptr, nullptr}, seq_back_ev[2] = {nullptr, nullptr};
    long long seq_frames = 0;
    double seq_pose[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};   // frame_pose of main.cpp:90, integrated per push
    uint8_t* d_bgr = nullptr;           // staging of colour (BGR) inputs, converted by k_bgr_to_gray (ingest.cu)
    size_t bgr_bytes = 0;
    // SM partition (green contexts, ctx.cu vo_partition_enable): the LK ring kernel -- persistent, 100 % of the registers of
    // every SM it runs on -- gets its own SMs, every other kernel of the batched path (FAST, pyramids, filters, triangulation,
    // PnP) runs on the rest, so the helper kernels of one unit range execute WHILE the other range's LK ring does
    bool part_on = false;
    bool part_auto = true;              // the first vo_batch_submit turns the partition on (8 SMs for the kernels after the ring)
    bool part_pre_with_lk = false;      // FAST / pyramids (throughput kernels) stay on the LK partition, only the latency-bound
                                        // kernels after the ring (filters, triangulation, PnP) go to the small one
    int part_helper_sms = 0, part_lk_sms = 0;
    void* part_gctx[2] = {nullptr, nullptr};            // CUgreenCtx: [0] helpers, [1] LK
    cudaStream_t part_lk_stream[VO_LANES] = {};     // per side stream
    cudaStream_t part_hp_stream[VO_LANES] = {};
    cudaEvent_t part_ev[VO_LANES][4] = {};
    // multi-GPU record gather over NCCL (dist.cu); NCCL is dlopen'ed at vo_dist_init
    void* dist_comm = nullptr;
    int dist_rank = 0, dist_world = 1;
    cudaStream_t dist_stream = nullptr, dist_snap_stream = nullptr;     // collectives / per-post snapshots (never queued behind a collective)
    // up to VO_DIST_DEPTH posted steps outstanding.  A post only snapshots the records (device to device) into the open
    // bucket; a bucket is exchanged with ONE in-place all-gather + ONE copy to pinned memory when it holds VO_DIST_BUCKET
    // steps, or earlier when the host asks for one of its steps.
    struct DistBucket { void* d = nullptr; void* h = nullptr; cudaEvent_t done = nullptr; int fill = 0, unwaited = 0, n_units = 0; bool flushed = false; };
    struct DistStep { int bucket = 0, index = 0; };
    DistBucket dist_bk[VO_DIST_NB];
    DistStep dist_steps[2 * VO_DIST_DEPTH];
    int dist_cur = 0;
    cudaEvent_t dist_ev_read = nullptr, dist_ev_fork = nullptr;
    size_t dist_bytes = 0;
    long long dist_head = 0, dist_tail = 0;
    // mono_rotation branch (ess.cu): scratch of the essential-matrix RANSAC, allocated on first use
    void* d_ess = nullptr;
    int ess_cap = 0;
    std::vector<void*> allocs;          // everything cudaMalloc'ed for the batch state

    // ---- pinned host staging ------------------------------------------------------------------
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;

    // ---- LK kernel timing (CUDA events on the launching stream) -------------------------------
    std::vector<cudaEvent_t> ev_pool;   // pairs: start, stop
    size_t ev_used = 0;
    double lk_ms = 0.0;
    long long lk_n = 0;
    bool lk_timing = true;
    bool lk_use_tma = true;
    int lk_ctas_per_sm = 0;             // 0 = the default instantiation (LK_CTAS_PER_SM)
    int lk_quota = 0;                   // work items a warp takes before its CTA retires (0 = persistent): retiring CTAs let the
                                        // high-priority helper kernels of the other unit range onto the SMs between LK work
    int lk_span = 0;                    // phases per LK work item: 0 = automatic (one level-solve per item when a launch has
                                        // more features than resident warps, else one item per feature-ring)
    int* d_lk_progress = nullptr;       // [units][cap] hand-over counters of the LK work items (zero between launches)
    int lk_per_unit = 0;                // upper bound of live features per unit known to the host (0 = cap)
    // work queues of the persistent LK warps: one (next, dry) pair per stream that launches the kernel,
    // so launches of different streams never share a pair; a pair resets itself at the end of a launch
    int* d_lk_queue = nullptr;          // [LK_QUEUES][2]
    std::vector<cudaStream_t> lk_queue_streams;

    // ---- batched path bookkeeping -----------------------------------------------------------
    int batch_units = 0;            // units configured by vo_batch_configure
    int batch_uploaded = 0;         // units currently resident
    bool batch_detect = false;      // features come from the on-GPU FAST + stride selection
    int batch_streams = 2;          // unit ranges run concurrently by the batched path
    bool use_graphs = true;         // replay the per-range kernel sequence as a CUDA graph (no LK event timing then)
    struct RangeGraph { int u0, n; bool detect, tma; cudaStream_t s; int max_pts; cudaGraphExec_t exec; long long launches; };   // s: the stream it was captured on (its LK work queue is that stream's)
    std::vector<RangeGraph> graphs; // invalidated when the device state is re-allocated
    int batch_max_pts = 0;          // largest per-unit feature count of the resident batch
    cudaStream_t hi_stream[VO_LANES] = {};     // high-priority helpers of the side streams (see run_range_launch)
    cudaEvent_t hi_ev[VO_LANES][4] = {};
    bool use_priorities = true;
    bool batch_graphs = false;      // force CUDA graphs for side-stream ranges even though they lose the priority split
    cudaStream_t side_stream[VO_LANES] = {};   // pipelining of vo_frame_batch (H2D of chunk k+1 under compute of chunk k) and of vo_batch_submit
    cudaEvent_t fork_ev = nullptr, join_ev[VO_LANES] = {};
    struct Pending { int u0 = 0, n = 0; bool active = false; cudaEvent_t done = nullptr; };
    std::vector<Pending> pending;   // vo_batch_submit / vo_batch_wait
    // full outputs of a submission (what matchingFeatures / trackingFrame2Frame hand back): packed per unit on the
    // device and copied with ONE D2H per submission into pinned staging (vo_set_option "batch_outputs")
    bool batch_outputs = false;
    uint8_t* d_out = nullptr;       // [units][out_stride]  (part of the batch state)
    uint8_t* h_out = nullptr;       // pinned, [out_units][out_stride]
    size_t out_stride = 0;
    int out_per = 0, out_units = 0; // point slots per unit in a packed block; units the pinned block holds
    unsigned submit_count = 0;
};

void vo_set_error(vo_ctx* ctx, const char* fmt, ...);
int vo_ensure_state(vo_ctx* ctx, int w, int h, int units, int imgs_per_unit);
void vo_free_state(vo_ctx* ctx);
void vo_drop_graphs(vo_ctx* ctx);
// new projection matrices: cached graphs carry the old calibration in their kernel arguments, so they are dropped
void vo_set_calibration(vo_ctx* ctx, const float P_l[12], const float P_r[12]);
int vo_drain_pending(vo_ctx* ctx);
// Entry points that overwrite the shared image planes / unit-0 buffers call this first: refused (VO_E_INVALID) while
// sequence frames or batch submissions are in flight; an idle sequence is ended (its planes are about to be reused).
int vo_partition_enable(vo_ctx* ctx, int helper_sms);       // 0 = off
void vo_partition_destroy(vo_ctx* ctx);
int vo_dist_order_after_gathers(vo_ctx* ctx, cudaStream_t st);
void vo_dist_shutdown(vo_ctx* ctx);
int vo_claim_buffers(vo_ctx* ctx, const char* who, bool allow_pending_batches = false);
int vo_ensure_pinned(vo_ctx* ctx, size_t bytes);
int vo_ensure_bgr(vo_ctx* ctx, size_t bytes);
int vo_launch_bgr_to_gray(const uint8_t* d_bgr, size_t pitch, size_t img_stride_in, uint8_t* d_gray, size_t img_stride_out,
                          int w, int h, int n_img, cudaStream_t s);
// a contiguous range of resident work units processed on one stream
// plane0 >= 0 overrides the image-plane base (default u0 * imgs_per_unit): the sequence mode ping-pongs its per-frame
// buffers between units 0 and 1 while both use the same four image planes
struct View { int u0, n; cudaStream_t s; int plane0 = -1; };
// run pyramids + LK (ncalls chained) for the units of `v`; images must already be in d_raw/d_raw_tab
int vo_run_lk(vo_ctx* ctx, const View& v, int ncalls, const int* img_prev, const int* img_next, bool want_err);
int vo_run_pyramid(vo_ctx* ctx, int plane0, int nplanes, cudaStream_t s);
int vo_run_lk_ring(vo_ctx* ctx, const View& v, int ncalls, const int* img_prev, const int* img_next, bool want_err);
int vo_run_filter(vo_ctx* ctx, const View& v, bool with_ages);
// FAST on raw plane `plane_in_unit` of each unit -> d_corners / d_ndet ; stride selection -> d_pts_in / d_npts
int vo_run_fast(vo_ctx* ctx, const View& v, int plane_in_unit, bool want_resp);
int vo_run_select(vo_ctx* ctx, const View& v);
// triangulate pts_l/pts_r ([units][cap], counts n) -> d_X ; PnP on (d_X, pts2d) -> d_results / d_inliers
int vo_run_triangulate(vo_ctx* ctx, const View& v, const float2* pts_l, const float2* pts_r, const int* n, float4* X4 = nullptr);
int vo_run_pnp(vo_ctx* ctx, const View& v, const float2* pts2d, const int* n, const float* K9);
