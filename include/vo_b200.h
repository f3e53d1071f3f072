/*
 * vo_b200.h -- C-ABI of the B200-native visual-odometry front-end (libvo_b200.so).
 *
 * This is the drop-in boundary for the per-frame hot path of ZhenghaoFei/visual_odom:
 * every entry point replaces one OpenCV-backed function of the reference's libfeature /
 * libvisualOdometry shared libraries (reference src/CMakeLists.txt:16-19,31-34).  The C++
 * facade in include/compat/ keeps the reference's own signatures (feature.h, visualOdometry.h)
 * and forwards to these functions; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C, no torch / OpenCV types; caller owns every buffer; outputs are capacity-passed.
 *   - images are 8-bit single channel (CV_8UC1), row pitch in bytes (reference utils.cpp:179,189).
 *   - points are {float x, y} (cv::Point2f), status is unsigned char, ages are int32.
 *   - every call is synchronous at return unless it says "async" (then it is ordered on the
 *     context's stream; see vo_set_stream / vo_sync).
 *   - return value: VO_OK (0) or a negative VO_E_* code; vo_last_error() gives the text.
 *   - there is NO CPU fallback: if no sm_100-class GPU is usable, vo_create fails.
 *   - a context is not thread-safe; distinct contexts are independent (one per host thread/GPU).
 */
#ifndef VO_B200_H
#define VO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VO_API __attribute__((visibility("default")))
#else
#define VO_API
#endif

#define VO_OK 0
#define VO_E_INVALID (-1)        /* bad argument                                             */
#define VO_E_CUDA (-2)           /* CUDA runtime/driver failure (text in vo_last_error)      */
#define VO_E_TOO_FEW_POINTS (-3) /* PnP with < 4 points (the reference aborts in cv::solvePnPRansac) */
#define VO_E_UNSUPPORTED (-4)    /* parameter outside what the kernels are built for         */
/* pnp_status only: RANSAC found no model with more than 4 inliers; the pose is what the reference is left with when it
 * ignores cv::solvePnPRansac's return value: R = I (rvec stays 0), t = the caller's guess */
#define VO_PNP_NO_MODEL 1
#define VO_E_CAPACITY (-5)       /* caller buffer / context capacity too small               */

typedef struct vo_ctx vo_ctx;

typedef struct vo_point2f { float x, y; } vo_point2f;
typedef struct vo_point3f { float x, y, z; } vo_point3f;

/* All literals the reference hard-codes, as run-time parameters ("Config"). */
typedef struct vo_params {
    int fast_threshold;      /* 20      reference src/feature.cpp:43                          */
    int fast_nonmax;         /* 1       reference src/feature.cpp:44                          */
    int lk_win;              /* 21      reference src/feature.cpp:127 (only 21 is built)      */
    int lk_max_level;        /* 3       reference src/feature.cpp:136 (0-based: 4 images)     */
    int lk_max_iters;        /* 30      reference src/feature.cpp:128                         */
    double lk_epsilon;       /* 0.01    reference src/feature.cpp:128                         */
    double lk_min_eig;       /* 0.001   reference src/feature.cpp:136                         */
    int circ_threshold;      /* 0       reference src/visualOdometry.cpp:120                  */
    int pnp_iterations;      /* 500     reference src/visualOdometry.cpp:168                  */
    float pnp_reproj_error;  /* 0.5     reference src/visualOdometry.cpp:169                  */
    double pnp_confidence;   /* 0.999   reference src/visualOdometry.cpp:170                  */
    int max_features;        /* per-unit feature capacity of the context (default 8192)       */
    int max_units;           /* work units the batched path can hold at once (default 1)      */
} vo_params;

VO_API void vo_default_params(vo_params* p);

/* Create a context on CUDA device `device`.  Fails (VO_E_CUDA) when no GPU is present. */
VO_API int vo_create(int device, const vo_params* params, vo_ctx** out);
VO_API void vo_destroy(vo_ctx* ctx);
VO_API const char* vo_last_error(const vo_ctx* ctx);
/* Run all subsequent work on `cuda_stream` (a cudaStream_t; NULL = the context's own stream). */
VO_API int vo_set_stream(vo_ctx* ctx, void* cuda_stream);
VO_API int vo_sync(vo_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's "gpu_launches"). */
VO_API long long vo_kernel_launches(const vo_ctx* ctx);
/* Accumulated device time (ms) of the LK ring kernel launches since the last reset,
 * measured with CUDA events on the launching stream; n = launches counted. */
VO_API int vo_lk_kernel_time(vo_ctx* ctx, double* ms_total, long long* n, int reset);
/* Run-time knobs (measurement / debugging): "batch_streams" = 1|2 (unit ranges the batched path runs
 * concurrently, default 2), "lk_staging" = 0 (TMA, default) | 1 (plain loads), "graphs" = 1 (default: kernel
 * sequences are replayed as CUDA graphs where that does not defeat the priority split) | 0 (plain launches; needed
 * for vo_lk_kernel_time), "priorities" = 1 (default: with several unit ranges in flight the short / latency-bound
 * kernels run on high-priority helper streams and only the LK ring at normal priority; such ranges are launched
 * plainly because captured graph nodes lose the stream priority) | 0, "batch_graphs" = 1 forces graphs for them. */
VO_API int vo_set_option(vo_ctx* ctx, const char* key, double value);

/* ---- A1: cv::FAST(image, kps, threshold, nonmax) + KeyPoint::convert ------------------------
 * replaces featureDetectionFast(), reference src/feature.cpp:39-47 (decl feature.h:48).
 * Output in raster order (y, then x), integer coordinates as float.  *n_out is the number of
 * corners found; at most `cap` are written (VO_E_CAPACITY if n_out > cap, buffer still filled). */
VO_API int vo_fast_detect(vo_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch,
                          vo_point2f* out, float* response /* optional */, int cap, int* n_out);

/* ---- one cv::calcOpticalFlowPyrLK call ---------------------------------------------------------
 * replaces the call inside featureTracking(), reference src/feature.cpp:72 (decl feature.h:52),
 * and is the unit the ring below chains.  err may be NULL. */
VO_API int vo_lk_track(vo_ctx* ctx, const uint8_t* prev, const uint8_t* next, int w, int h, size_t pitch,
                       const vo_point2f* prev_pts, int n, vo_point2f* next_pts, uint8_t* status, float* err);

/* ---- A2 + A3: circularMatching() ---------------------------------------------------------------
 * replaces circularMatching(), reference src/feature.cpp:118-148 (decl feature.h:61-65), i.e. the
 * four chained LK calls L0->R0->R1->L1->L0 and deleteUnmatchFeaturesCircle() (src/feature.cpp:76-116).
 *   pts_l0[n]            input features (points_l_0)
 *   ages_io[n]           in: current_features.ages; out: ages+1, compacted to *n_kept (may be NULL)
 *   o_l0,o_r0,o_l1,o_r1,o_l0_ret  capacity n each: the five point vectors after the erase loop
 *   status4              optional, 4*n bytes: raw status of the four calls, original indexing
 *   raw4                 optional, 4*n points: raw outputs of the four calls (R0,R1,L1,L0_ret), original indexing
 *   kept_idx[n]          original indices of the survivors, ascending; *n_kept their count */
VO_API int vo_circular_match(vo_ctx* ctx, const uint8_t* l0, const uint8_t* r0, const uint8_t* l1,
                             const uint8_t* r1, int w, int h, size_t pitch, const vo_point2f* pts_l0, int n,
                             int32_t* ages_io, vo_point2f* o_l0, vo_point2f* o_r0, vo_point2f* o_l1,
                             vo_point2f* o_r1, vo_point2f* o_l0_ret, uint8_t* status4, vo_point2f* raw4,
                             int32_t* kept_idx, int* n_kept);

/* ---- A8: cv::triangulatePoints + cv::convertPointsFromHomogeneous ------------------------------
 * replaces the call site reference src/main.cpp:170-171 (and Frame::triangulateFeaturePoints,
 * src/Frame.cpp:25-28).  P_l, P_r: row-major 3x4 float (CV_32F, main.cpp:73-74). */
VO_API int vo_triangulate(vo_ctx* ctx, const float P_l[12], const float P_r[12], const vo_point2f* pts_l,
                          const vo_point2f* pts_r, int n, vo_point3f* X);
/* The same triangulation, returned the way cv::triangulatePoints itself returns it (Frame::triangulateFeaturePoints,
 * reference src/Frame.cpp:25-28): X4 = n x 4 floats, point i = the unit-norm homogeneous vector (x, y, z, w) that is
 * column i of OpenCV's 4 x N CV_32F output (same bits, same sign). */
VO_API int vo_triangulate_homogeneous(vo_ctx* ctx, const float P_l[12], const float P_r[12], const vo_point2f* pts_l,
                                      const vo_point2f* pts_r, int n, float* X4);

/* ---- A9: cv::solvePnPRansac(..., SOLVEPNP_ITERATIVE, useExtrinsicGuess) + cv::Rodrigues --------
 * replaces the pose solve of trackingFrame2Frame(), reference src/visualOdometry.cpp:161-189.
 *   K            row-major 3x3 float intrinsics (built from P_l, visualOdometry.cpp:163-165)
 *   rvec_io      in: initial rvec (the reference resets it to 0 every call, :162); out: solution
 *   tvec_io      in: extrinsic guess (previous translation, main.cpp:82,181); out: solution
 *   inliers[n]   ascending inlier indices (CV_32S column in the reference), *n_inliers their count
 *   R_out        row-major 3x3 double = Rodrigues(rvec)
 * n == 4: as OpenCV, no RANSAC: the P3P pose of the first three points that reprojects the fourth best, unrefined, all
 * four reported as inliers, *ransac_iters = 0 (agrees with cv2 to <= 1e-6 on [R|t] in generic scenes; csrc/p3p_math.cuh).
 * Returns VO_E_TOO_FEW_POINTS for n < 4 (the reference would abort with cv::Exception). */
VO_API int vo_pnp_ransac(vo_ctx* ctx, const vo_point3f* X, const vo_point2f* x, int n, const float K[9],
                         double rvec_io[3], double tvec_io[3], int32_t* inliers, int* n_inliers,
                         double R_out[9], int* ransac_iters /* optional */);

/* ---- N5: the mono_rotation = true branch of trackingFrame2Frame (the header default of the reference's flag) -------
 * replaces reference src/visualOdometry.cpp:146-157:
 *     E = cv::findEssentialMat(pointsLeft_t0, pointsLeft_t1, focal, pp, cv::RANSAC, 0.999, 1.0, mask);
 *     cv::recoverPose(E, pointsLeft_t0, pointsLeft_t1, rotation, translation_mono, focal, pp, mask);
 * Five-point RANSAC (1000 iterations at most, adaptive bound) + the four-way cheirality vote.  R_out = `rotation`
 * (row-major 3x3); mask_out (optional, n bytes) = inliers of the best E; *n_inliers their count.
 * n < 5 or no model: VO_E_TOO_FEW_POINTS (OpenCV throws in both cases). */
VO_API int vo_mono_rotation(vo_ctx* ctx, const vo_point2f* pts_t0, const vo_point2f* pts_t1, int n, double focal, double ppx,
                            double ppy, double R_out[9], uint8_t* mask_out, int* n_inliers, int* ransac_iters);

/* ---- batched whole-path API (configs 4/5 of BASELINE.json, bench.py, multi-GPU sharding) ------
 * One work unit = one stereo pair-of-pairs + its feature list (SURVEY.md section 8d "Work unit").
 * The batched path keeps everything device-resident between stages:
 *   FAST on l0 (when pts == NULL) -> even-stride selection of select_n corners
 *   -> pyramids -> LK ring -> status/negative/circular filters -> triangulation -> PnP/RANSAC. */
typedef struct vo_unit {
    const uint8_t *l0, *r0, *l1, *r1;   /* HOST images (w x h, pitch) unless VO_UNIT_DEVICE_IMAGES   */
    const vo_point2f* pts;              /* HOST features of l0, or NULL = detect on the GPU         */
    int n_pts;                          /* features in pts; with pts==NULL: number to select        */
    double t_prev[3];                   /* extrinsic guess for the pose solve                       */
} vo_unit;

typedef struct vo_unit_result {
    int n_features;      /* features fed to the ring                                   */
    int n_detected;      /* FAST corners found on l0 (0 when pts were given)           */
    int n_tracked;       /* survivors of deleteUnmatchFeaturesCircle (A3)              */
    int n_valid;         /* survivors of checkValidMatch/removeInvalidPoints (A5/A6)   */
    int n_inliers;       /* RANSAC inliers                                             */
    int ransac_iters;    /* iterations the adaptive loop ran                           */
    int pnp_status;      /* VO_OK, VO_PNP_NO_MODEL or VO_E_TOO_FEW_POINTS (n < 4); n == 4 runs OpenCV's P3P case */
    double rvec[3], tvec[3], R[9];
} vo_unit_result;

/* Allocate/resize the device-resident batch state for n_units units of w x h images. */
VO_API int vo_batch_configure(vo_ctx* ctx, int w, int h, int n_units, const float P_l[12], const float P_r[12]);
/* async: copy the units' images (and features) host->device on the context's stream. */
VO_API int vo_batch_upload(vo_ctx* ctx, const vo_unit* units, int n_units, size_t pitch);
/* async: run the whole path for the uploaded units. */
VO_API int vo_batch_run(vo_ctx* ctx);
/* async D2H of the per-unit result records into pinned staging + sync + copy to `results`. */
VO_API int vo_batch_download(vo_ctx* ctx, vo_unit_result* results, int n_units);
/* upload + run + download in one call: the end-to-end entry point. */
VO_API int vo_frame_batch(vo_ctx* ctx, const vo_unit* units, int n_units, size_t pitch, vo_unit_result* results);
/* Pipelined form of vo_frame_batch.  vo_batch_submit fills the resident unit slots [first_unit, first_unit + n_units)
 * from units[0 .. n_units) (units == NULL: re-run what is resident there), runs the whole path on them and stages their
 * result records, all asynchronously; vo_batch_wait blocks until that submission is done and copies the records out.
 * Submissions on disjoint slot ranges overlap on the GPU (H2D and the latency-bound PnP tail of one under the LK ring of
 * the other).  The library runs up to three submissions concurrently (three lanes of streams): with inputs resident on
 * the device two in flight saturate the GPU; with inputs coming from the host configure 3 x B units and keep three
 * submissions of B in flight, so that the upload of step s+2 is already queued while the host reads step s (otherwise the
 * host's enqueue time and the H2D copy sit between two steps of the GPU).  The host images of a submission
 * must stay valid (and, for a true async copy, be pinned) until it has been waited for; a slot range must be waited
 * for before it is submitted again.  vo_batch_fetch of a waited slot is valid until that slot is resubmitted. */
VO_API int vo_batch_submit(vo_ctx* ctx, const vo_unit* units, int first_unit, int n_units, size_t pitch);
VO_API int vo_batch_wait(vo_ctx* ctx, int first_unit, int n_units, vo_unit_result* results);
/* Fetch the per-unit arrays of the last run (any pointer may be NULL). Capacity = max_features.
 *   pts4: 4 x n_valid points (L0,R0,L1,R1 after A6);  kept_idx: n_valid original indices;
 *   X: n_valid 3-D points;  inliers: n_inliers indices into the n_valid list. */
/* With vo_set_option(ctx, "batch_outputs", 1) every submission also returns what the reference's matchingFeatures() /
 * trackingFrame2Frame() hand back (reference src/visualOdometry.h:27-42): the four point lists, the tracked-feature
 * indices, points3D and the inlier list of every unit, packed on the device and copied with ONE device-to-host copy per
 * submission into pinned staging.  vo_batch_outputs reads a waited unit from that staging (layout as vo_batch_fetch;
 * any pointer may be NULL); *d2h_bytes_per_unit = bytes that crossed PCIe for the unit's packed block. */
VO_API int vo_batch_outputs(vo_ctx* ctx, int unit, vo_point2f* pts4, int32_t* kept_idx, vo_point3f* X, int32_t* inliers,
                            size_t* d2h_bytes_per_unit);
VO_API int vo_batch_fetch(vo_ctx* ctx, int unit, vo_point2f* pts_in, vo_point2f* pts4, int32_t* kept_idx,
                          vo_point3f* X, int32_t* inliers);

/* ---- multi-GPU: gather of the result records over NCCL (one process per GPU, units sharded) -----------
 * The path has no data-path collective; the only exchange is the gather of the fixed-size records.  NCCL is resolved with
 * dlopen at vo_dist_init (VO_E_UNSUPPORTED when the host has none).  Rank 0 makes the id with vo_dist_unique_id and hands the
 * 128 bytes to the other ranks out of band; every rank calls vo_dist_init once.  vo_dist_gather_post posts, without
 * blocking, the records of resident slots [first_unit, first_unit + n_units) (same n_units on every rank): the post is a
 * device-side snapshot, so the slots may be refilled at once and a later submission never waits for another rank.  The
 * exchange itself is batched: one in-place ncclAllGather + one copy into pinned staging per 4 posts, or as soon as
 * vo_dist_gather_wait needs a posted step that has not been exchanged yet.  Up to VO_DIST_DEPTH posts may be outstanding
 * (ranks may drift apart by that many steps before anyone blocks); vo_dist_gather_wait returns the oldest one:
 * all[r * n_units + i] = record i of rank r.  Every rank must post and wait in the same order. */
#define VO_DIST_DEPTH 8
VO_API int vo_dist_unique_id(uint8_t id_out[128]);
VO_API int vo_dist_init(vo_ctx* ctx, const uint8_t id[128], int rank, int world);
VO_API int vo_dist_gather_post(vo_ctx* ctx, int first_unit, int n_units);
VO_API int vo_dist_gather_wait(vo_ctx* ctx, vo_unit_result* all, int cap_records, int* n_records);

/* ---- streaming sequence mode ------------------------------------------------
 * The state of the reference's main loop (src/main.cpp:87-92,123-181: currentVOFeatures, the previous
 * stereo pair, `translation`) lives on the device.  vo_seq_begin uploads the first pair; each vo_seq_push
 * uploads only the NEW pair, builds only its two pyramids and runs matchingFeatures() (FAST refill,
 * bucketing rows/10 x 1, circular matching, 1-px round-trip check) -> triangulation ->
 * trackingFrame2Frame(mono_rotation=false), carrying features / ages / translation to the next frame
 * exactly as the reference does (including the ages-vs-points length skew).
 *   out      counts + pose of this frame pair
 *   pts4     optional: 4 arrays of pts_cap points (L0, R0, L1, R1 after the circular check); the first
 *            out->n_valid entries of each are meaningful, the rest of the arrays is scratch
 * The per-frame kernel sequence is replayed as two CUDA graphs (front half / pose solve; option "graphs"). */
VO_API int vo_seq_begin(vo_ctx* ctx, int w, int h, const float P_l[12], const float P_r[12], const uint8_t* left0,
                        const uint8_t* right0, size_t pitch);
VO_API int vo_seq_push(vo_ctx* ctx, const uint8_t* left1, const uint8_t* right1, size_t pitch, vo_unit_result* out,
                       vo_point2f* pts4, int pts_cap);
/* same, for inputs with `channels` interleaved bytes per pixel: 1 = gray, 3 = BGR as cv::imread(IMREAD_COLOR) returns
 * it -- the BGR bytes are uploaded as they are and converted on the device with cv::cvtColor(BGR2GRAY)'s fixed-point
 * formula (reference src/utils.cpp:178-179,188-189) inside the frame's graph. */
VO_API int vo_seq_begin_ex(vo_ctx* ctx, int w, int h, const float P_l[12], const float P_r[12], const uint8_t* left0,
                           const uint8_t* right0, size_t pitch, int channels);
VO_API int vo_seq_push_ex(vo_ctx* ctx, const uint8_t* left1, const uint8_t* right1, size_t pitch, int channels,
                          vo_unit_result* out, vo_point2f* pts4, int pts_cap);
/* Pipelined form: vo_seq_submit enqueues a frame and returns; vo_seq_wait blocks for the OLDEST frame in flight and
 * returns its record (and integrates frame_pose).  At most two frames may be in flight: only the pose solve of frame
 * k+1 depends on the pose solve of frame k (the extrinsic guess), so the upload, pyramids, FAST, bucketing, LK ring,
 * filters and triangulation of frame k+1 run under the latency-bound PnP of frame k -- one frame of result lag buys
 * ~1.7x the frame rate.  Results are identical to vo_seq_push (= submit + wait).  The host images of a submitted frame
 * must stay valid until the call returns for pageable memory, until its vo_seq_wait for pinned memory. */
VO_API int vo_seq_submit(vo_ctx* ctx, const uint8_t* left1, const uint8_t* right1, size_t pitch, int channels);
VO_API int vo_seq_wait(vo_ctx* ctx, vo_unit_result* out, vo_point2f* pts4, int pts_cap);
/* currentVOFeatures (points / ages may differ in length) and the carried translation (waits for frames in flight) */
VO_API int vo_seq_state(vo_ctx* ctx, vo_point2f* points, int32_t* ages, int cap, int* n_points, int* n_ages, double t_out[3]);

/* ---- image ingest ---------------------------------------------------------------
 * What loadImageLeft / loadImageRight do (reference src/utils.cpp:172-190): read <dir>/image_0/%06d.png and
 * <dir>/image_1/%06d.png with cv::imread(IMREAD_COLOR) and cvtColor(BGR2GRAY).  The decoder is host code
 * (zlib inflate is a serial stream); the reader decodes AHEAD of the consumer on worker threads into a ring of
 * pinned buffers, so vo_seq_push's upload is one async DMA per image.
 *   vo_png_info / vo_png_decode   one PNG held in memory -> BGR (3 B/px) and/or gray (1 B/px); either may be NULL.
 *                                 Non-interlaced PNGs of every colour type / bit depth; 16-bit -> 8-bit by the high
 *                                 byte, alpha dropped (what IMREAD_COLOR does).  Errors: vo_png_last_error().
 *   vo_reader_open                sequence_dir/image_{0,1}/%06d.png, frames first_frame .. first_frame+n_frames-1,
 *                                 `threads` decoders, `depth` (>= 3) frames of pinned ring.  force_channels: 0 = gray
 *                                 files are delivered as gray, colour files as BGR (device conversion); 1 / 3 force.
 *   vo_reader_next                blocks until the next frame is decoded; the returned pointers stay valid until the
 *                                 SECOND following vo_reader_next (so a frame handed to vo_seq_submit may still be
 *                                 uploading while the next one is requested) or vo_reader_close.
 *   vo_bgr_to_gray                the device conversion on host buffers (stage-level entry point, for parity tests) */
typedef struct vo_reader vo_reader;
VO_API int vo_png_info(const uint8_t* file_bytes, size_t n, int* w, int* h, int* color_type, int* bit_depth);
VO_API int vo_png_decode(const uint8_t* file_bytes, size_t n, uint8_t* bgr, size_t bgr_pitch, uint8_t* gray, size_t gray_pitch);
VO_API const char* vo_png_last_error(void);
VO_API vo_reader* vo_reader_open(const char* sequence_dir, int first_frame, int n_frames, int threads, int depth, int force_channels);
VO_API int vo_reader_next(vo_reader* rd, const uint8_t** left, const uint8_t** right, int* w, int* h, size_t* pitch,
                          int* channels, int* frame_id);
VO_API const char* vo_reader_error(vo_reader* rd);
VO_API void vo_reader_close(vo_reader* rd);
VO_API int vo_bgr_to_gray(vo_ctx* ctx, const uint8_t* bgr, size_t pitch, int w, int h, uint8_t* gray, size_t gray_pitch);

/* ---- pose bookkeeping -- host-only, O(1) per frame ---------------------------
 * R is row-major 3x3, t is 3x1, frame_pose / rigid_inv are row-major 4x4 (the reference's CV_64F Mats).
 *   vo_pose_is_rotation  replaces isRotationMatrix            (src/utils.cpp:93-102):  |I - R^T R|_F < 1e-6
 *   vo_pose_euler        replaces rotationMatrixToEulerAngles (src/utils.cpp:107-131): float x,y,z
 *   vo_pose_integrate    replaces integrateOdometryStereo     (src/utils.cpp:57-91):   rigid_inv (optional) =
 *                        [R|t;0 0 0 1]^-1; frame_pose *= rigid_inv iff 0.05 < |t| < 10.  Returns 1 when the
 *                        pose was advanced, 0 when the frame was skipped, <0 on a singular transform.
 *   vo_pose_step         the main loop's gate + integration    (src/main.cpp:196-208):  all |euler| < 0.1
 *   vo_seq_pose          frame_pose accumulated by vo_seq_push since vo_seq_begin */
VO_API int  vo_pose_is_rotation(const double R[9]);
VO_API void vo_pose_euler(const double R[9], float euler_xyz[3]);
VO_API int  vo_pose_integrate(double frame_pose[16], const double R[9], const double t[3], double rigid_inv[16]);
VO_API int  vo_pose_step(double frame_pose[16], const double R[9], const double t[3]);
VO_API int  vo_seq_pose(vo_ctx* ctx, double frame_pose[16]);

/* ---- KITTI accuracy evaluation -- host-only, offline -----------------------------
 * Poses are KITTI rows: 12 doubles = top 3 rows of the 4x4 camera-to-world matrix (what vo_seq_pose accumulates).
 *   vo_poses_load / vo_poses_save   replace loadPoses (src/evaluate/evaluate_odometry.cpp:17-33; also read by
 *                                   src/main.cpp for the ground-truth overlay) and the result writer
 *   vo_eval_segments                replaces calcSequenceErrors (evaluate_odometry.cpp:71-116): start every `step`
 *                                   (10) frames, segment lengths `lengths` (NULL = 100..800 m) along the ground-truth
 *                                   path; per segment r_err [rad/m], t_err [fraction], speed [m/s at 10 Hz]
 *   vo_eval_summary                 replaces saveStats (evaluate_odometry.cpp:376-395): mean t_err, r_err
 * With out == NULL vo_eval_segments / vo_poses_load only count. */
typedef struct vo_segment_error { int32_t first_frame; float r_err, t_err, len, speed; } vo_segment_error;
VO_API int vo_poses_load(const char* path, double* poses12, int cap, int* n_out);
VO_API int vo_poses_save(const char* path, const double* poses12, int n);
VO_API int vo_eval_segments(const double* gt12, const double* est12, int n_poses, const float* lengths, int n_lengths,
                            int step, vo_segment_error* out, int cap, int* n_out);
VO_API int vo_eval_summary(const vo_segment_error* seg, int n, float* t_err_avg, float* r_err_avg);

#ifdef __cplusplus
}
#endif
#endif /* VO_B200_H */
