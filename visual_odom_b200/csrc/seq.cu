// seq.cu -- streaming sequence mode: the reference's main loop state
// (src/main.cpp:87-92,123-181: currentVOFeatures, the previous stereo pair, `translation`) kept resident
// on the device, so one vo_seq_push() uploads only the NEW stereo pair, builds only its two pyramids
// (the previous pair's pyramids are reused as t0) and runs matchingFeatures() -> triangulation ->
// trackingFrame2Frame() without a host round trip.
//
// The glue between the OpenCV-backed stages is the reference's own, restated bug-for-bug on the device:
//   appendNewFeatures   src/feature.cpp:255-262  (refill when fewer than 2000 features: always, in practice)
//   bucketingFeatures   src/feature.cpp:206-253 + Bucket::add_feature src/bucket.cpp:14-45
//        (nh+1)*(nw+1) cells addressed with row stride nw (aliasing + duplicated read-back), ages >= 10
//        refused, a full one-slot cell is overwritten: the LAST admitted feature in input order wins
//   the ages / points length skew after the circular check (src/visualOdometry.cpp:122-127): point i is
//        paired with ages[i] even though the ages vector is longer and shifted.
#include "common.cuh"
#include "seq.h"

__global__ void __launch_bounds__(1024) k_seq_append(const float2* __restrict__ corners, const int* __restrict__ n_det,
                                                     int corner_cap, float2* feat_pts, int* feat_ages, int* cnt, int feat_cap,
                                                     int refill_below, int* err)
{
    // the error bits are per frame: this is the first glue kernel of a frame's front stage, it clears the frame's word
    if (threadIdx.x == 0) *err = 0;
    __syncthreads();
    const int n_pts = cnt[0], n_ages = cnt[1];
    if (n_pts >= refill_below) return;                 // `if (currentVOFeatures.size() < 2000)`
    int m = *n_det;
    if (m > corner_cap) { m = corner_cap; if (threadIdx.x == 0) atomicOr(err, 1); }
    if (n_pts + m > feat_cap || n_ages + m > feat_cap) { m = min(feat_cap - n_pts, feat_cap - n_ages); if (m < 0) m = 0; if (threadIdx.x == 0) atomicOr(err, 2); }
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        feat_pts[n_pts + i] = corners[i];
        feat_ages[n_ages + i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) { cnt[0] = n_pts + m; cnt[1] = n_ages + m; }
}

__global__ void __launch_bounds__(1024) k_seq_bucket(const float2* __restrict__ feat_pts, const int* __restrict__ feat_ages,
                                                     const int* __restrict__ cnt, int rows, int cols, int bucket_size,
                                                     int* bucket /* [nb] scratch */, int nb_cap,
                                                     float2* out_pts, int* out_ages, int* out_n, int out_cap, int* err)
{
    const int nh = rows / bucket_size, nw = cols / bucket_size;
    const int nb = (nh + 1) * (nw + 1);
    if (nb > nb_cap) { if (threadIdx.x == 0) { atomicOr(err, 4); *out_n = 0; } return; }
    for (int b = threadIdx.x; b < nb; b += blockDim.x) bucket[b] = -1;
    __syncthreads();
    const int n = cnt[0];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (feat_ages[i] < 10) {                                         // Bucket::add_feature: age_threshold = 10
            const float2 p = feat_pts[i];
            const int bh = (int)(p.y / (float)bucket_size), bw = (int)(p.x / (float)bucket_size);
            const int idx = bh * nw + bw;                                // row stride nw, not nw + 1
            if (idx >= 0 && idx < nb) atomicMax(&bucket[idx], i);        // one slot per cell: the last admitted wins
            else atomicOr(err, 8);                                       // outside the image: undefined behaviour in the reference
        }
    }
    __syncthreads();
    // ordered read-back (same aliased addressing as the reference): cell sequence q = h * (nw + 1) + w, h <= nh, w <= nw.
    // Each thread takes a contiguous chunk of q, a block-wide exclusive scan of the chunk counts gives its output offset.
    __shared__ int s_cnt[1024];
    const int per = (nb + blockDim.x - 1) / blockDim.x;
    const int q0 = threadIdx.x * per, q1 = min(nb, q0 + per);
    int mine = 0;
    for (int q = q0; q < q1; q++) {
        const int h = q / (nw + 1), w = q - h * (nw + 1);
        mine += bucket[h * nw + w] >= 0 ? 1 : 0;
    }
    s_cnt[threadIdx.x] = mine;
    __syncthreads();
    for (int d = 1; d < (int)blockDim.x; d <<= 1) {           // Hillis-Steele inclusive scan
        const int v = threadIdx.x >= d ? s_cnt[threadIdx.x - d] : 0;
        __syncthreads();
        s_cnt[threadIdx.x] += v;
        __syncthreads();
    }
    int m = s_cnt[threadIdx.x] - mine;
    for (int q = q0; q < q1; q++) {
        const int h = q / (nw + 1), w = q - h * (nw + 1);
        const int b = bucket[h * nw + w];
        if (b >= 0) {
            if (m < out_cap) { out_pts[m] = feat_pts[b]; out_ages[m] = feat_ages[b]; }
            m++;
        }
    }
    if (threadIdx.x == blockDim.x - 1) *out_n = min(s_cnt[threadIdx.x], out_cap);
}

// after the circular check: currentVOFeatures.points = pointsLeft_t1 (A5 survivors), the ages keep their A3 length
// (src/visualOdometry.cpp:122-127).  Runs before the pose solve, so the next frame's front half can start under it.
__global__ void __launch_bounds__(256) k_seq_carry(const float2* __restrict__ valid_l1, const int* __restrict__ n5,
                                                   const int* __restrict__ ages_out, const int* __restrict__ n3,
                                                   float2* feat_pts, int* feat_ages, int* cnt)
{
    const int np = *n5, na = *n3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(np, na); i += gridDim.x * blockDim.x) {
        if (i < np) feat_pts[i] = valid_l1[i];
        if (i < na) feat_ages[i] = ages_out[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { cnt[0] = np; cnt[1] = na; }
}

// after the pose solve: `translation` carries the solved tvec to the next frame's solve; counts into the result record
__global__ void k_seq_finish(vo_unit_result_dev* res, double* tprev_next, const int* __restrict__ n_feat,
                             const int* __restrict__ n_det, const int* __restrict__ n3, const int* __restrict__ n5,
                             const int* __restrict__ err, int* err_out)
{
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; k++) tprev_next[k] = res->tvec[k];
        res->n_features = *n_feat; res->n_detected = *n_det; res->n_tracked = *n3; res->n_valid = *n5;
        if (err_out != err) *err_out = *err;
    }
}

int vo_launch_seq_append(const SeqArgs& a, cudaStream_t s)
{
    k_seq_append<<<1, 1024, 0, s>>>(a.corners, a.n_det, a.corner_cap, a.feat_pts, a.feat_ages, a.cnt, a.feat_cap, a.refill_below, a.err);
    return 1;
}
int vo_launch_seq_bucket(const SeqArgs& a, cudaStream_t s)
{
    k_seq_bucket<<<1, 1024, 0, s>>>(a.feat_pts, a.feat_ages, a.cnt, a.rows, a.cols, a.bucket_size, a.bucket, a.bucket_cap,
                                    a.out_pts, a.out_ages, a.out_n, a.out_cap, a.err);
    return 1;
}
int vo_launch_seq_carry(const SeqArgs& a, cudaStream_t s)
{
    k_seq_carry<<<8, 256, 0, s>>>(a.valid_l1, a.n5, a.ages_out, a.n3, a.feat_pts, a.feat_ages, a.cnt);
    return 1;
}
int vo_launch_seq_finish(const SeqArgs& a, cudaStream_t s)
{
    k_seq_finish<<<1, 32, 0, s>>>(a.res, a.tprev, a.out_n, a.n_det, a.n3, a.n5, a.err, a.err_out);
    return 1;
}
