// ess.cu -- the `mono_rotation = true` branch of the reference's trackingFrame2Frame
// (reference src/visualOdometry.cpp:146-157):
//     E = cv::findEssentialMat(pointsLeft_t0, pointsLeft_t1, focal, pp, cv::RANSAC, 0.999, 1.0, mask);
//     cv::recoverPose(E, pointsLeft_t0, pointsLeft_t1, rotation, translation_mono, focal, pp, mask);
// The reference's main() passes mono_rotation = false (src/main.cpp:181), but the flag's header default is true
// (src/visualOdometry.h:42), so a drop-in must honour it.
//
// Structure (same shape as the PnP RANSAC of pnp.cu: waves of iterations, a unit that reached its adaptive bound skips
// the rest; every kernel reads the bound from device memory, so there is no host round trip):
//   k_ess_init          normalise the points ((p - pp) / focal in fp64), reset the RANSAC state
//   k_ess_subsets       1 thread: cv::RNG(2^64-1) stream -> 5 distinct indices per iteration (ptsetreg.cpp getSubset)
//   k_ess_hypotheses    1 thread / iteration: Nister five-point solver (ess_math.cuh) -> up to 10 E per sample
//   k_ess_count         1 CTA / (iteration, candidate): Sampson error of all N points, err <= (float)thr^2, count
//   k_ess_replay        1 thread: candidates in order, `count > max(best, 4)` -> new best, RANSACUpdateNumIters
//   k_ess_mask          inlier mask of the best E
//   k_ess_decompose     decomposeEssentialMat
//   k_ess_cheirality    1 thread / point: the four [R|t] hypotheses of recoverPose (fp64 DLT, distance threshold 50)
//   k_ess_pick          recoverPose's vote -> rotation
// Restated in oracle/essential_ref.py; the math of ess_math.cuh is checked on the host against cv2 4.13.0
// (tests/test_oracle_essential.py) and on the GPU through vo_mono_rotation (tests/test_gpu_stages.py).
#include "common.cuh"
#include "ess.h"
#include "ess_math.cuh"

using namespace vomath;

__global__ void k_ess_init(const EssArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        EssState& s = *a.state;
        s.rng_state = 0xffffffffffffffffULL;
        s.niters = a.max_iters;
        s.max_good = 0;
        s.best_it = -1; s.best_cand = -1;
        s.iters_run = 0;
        s.done = a.n < 5 ? 1 : 0;
        for (int k = 0; k < 4; k++) s.good4[k] = 0;
    }
    if (i >= a.n) return;
    const float2 p0 = a.pts0[i], p1 = a.pts1[i];
    a.q0[i] = make_double2(((double)p0.x - a.ppx) / a.focal, ((double)p0.y - a.ppy) / a.focal);
    a.q1[i] = make_double2(((double)p1.x - a.ppx) / a.focal, ((double)p1.y - a.ppy) / a.focal);
}

__global__ void k_ess_subsets(const EssArgs a, int it0, int it1)
{
    if (blockIdx.x || threadIdx.x) return;
    EssState& s = *a.state;
    if (s.done) return;
    const int n = a.n;
    Rng rng(s.rng_state);
    const int last = it1 < s.niters ? it1 : s.niters;
    for (int it = it0; it < last; it++) {
        int idx[5];
        if (n > 5) {
            for (int i = 0; i < 5; i++) {
                int v;
                bool dup;
                do {
                    v = (int)(rng.next() % (unsigned)n);
                    dup = false;
                    for (int j = 0; j < i; j++) dup |= (idx[j] == v);
                } while (dup);
                idx[i] = v;
            }
        } else {
            for (int i = 0; i < 5; i++) idx[i] = i;
        }
        for (int i = 0; i < 5; i++) a.subsets[it * 5 + i] = idx[i];
    }
    s.rng_state = rng.state;
}

__global__ void __launch_bounds__(32) k_ess_hypotheses(const EssArgs a, int it0, int it1)
{
    const int it = it0 + blockIdx.x * blockDim.x + threadIdx.x;
    const EssState& s = *a.state;
    if (s.done || it >= it1 || it >= s.niters) return;
    double q0[10], q1[10];
    for (int i = 0; i < 5; i++) {
        const int j = a.subsets[it * 5 + i];
        const double2 u = a.q0[j], v = a.q1[j];
        q0[2 * i] = u.x; q0[2 * i + 1] = u.y; q1[2 * i] = v.x; q1[2 * i + 1] = v.y;
    }
    a.nmodels[it] = five_point(q0, q1, a.models + (size_t)it * 90);
}

__global__ void __launch_bounds__(128) k_ess_count(const EssArgs a, int it0, int it1)
{
    const int it = it0 + blockIdx.x, cand = blockIdx.y;
    const EssState& s = *a.state;
    if (s.done || it >= it1 || it >= s.niters) return;
    if (cand >= a.nmodels[it]) return;
    __shared__ double E[9];
    __shared__ int total;
    if (threadIdx.x < 9) E[threadIdx.x] = a.models[(size_t)it * 90 + cand * 9 + threadIdx.x];
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        const double2 u = a.q0[i], v = a.q1[i];
        c += sampson_err(E, u.x, u.y, v.x, v.y) <= a.thr2;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&total, c);
    __syncthreads();
    if (threadIdx.x == 0) a.counts[it * 10 + cand] = total;
}

__global__ void k_ess_replay(const EssArgs a, int it0, int it1)
{
    if (blockIdx.x || threadIdx.x) return;
    EssState& s = *a.state;
    if (s.done) return;
    const int n = a.n;
    int it = it0;
    for (; it < it1 && it < s.niters; it++) {
        const int nm = a.nmodels[it];
        for (int m = 0; m < nm; m++) {
            const int good = a.counts[it * 10 + m];
            if (good > max(s.max_good, 4)) {
                s.best_it = it; s.best_cand = m;
                s.max_good = good;
                s.niters = ransac_update_num_iters(a.prob, (double)(n - good) / n, 5, s.niters);
            }
        }
    }
    s.iters_run = it;
    if (it >= s.niters || it1 >= a.max_iters) s.done = 1;
}

__global__ void k_ess_mask(const EssArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const EssState& s = *a.state;
    if (i >= a.n) return;
    if (s.best_it < 0) { a.mask[i] = 0; return; }
    const double* E = a.models + (size_t)s.best_it * 90 + s.best_cand * 9;
    const double2 u = a.q0[i], v = a.q1[i];
    a.mask[i] = sampson_err(E, u.x, u.y, v.x, v.y) <= a.thr2 ? 1 : 0;
}

__global__ void k_ess_decompose(const EssArgs a)
{
    if (blockIdx.x || threadIdx.x) return;
    const EssState& s = *a.state;
    if (s.best_it < 0) return;
    const double* E = a.models + (size_t)s.best_it * 90 + s.best_cand * 9;
    for (int k = 0; k < 9; k++) a.pose[21 + k] = E[k];
    decompose_essential(E, a.pose, a.pose + 9, a.pose + 18);      // R1 | R2 | t | E
}

__global__ void __launch_bounds__(128) k_ess_cheirality(const EssArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    EssState& s = *a.state;
    if (s.best_it < 0) return;
    int ok[4] = {0, 0, 0, 0};
    if (i < a.n && a.mask[i]) {
        const double2 u = a.q0[i], v = a.q1[i];
        const double* t = a.pose + 18;
        const double tn[3] = {-t[0], -t[1], -t[2]};
        ok[0] = cheirality_ok(a.pose, t, u.x, u.y, v.x, v.y, 50.0);
        ok[1] = cheirality_ok(a.pose + 9, t, u.x, u.y, v.x, v.y, 50.0);
        ok[2] = cheirality_ok(a.pose, tn, u.x, u.y, v.x, v.y, 50.0);
        ok[3] = cheirality_ok(a.pose + 9, tn, u.x, u.y, v.x, v.y, 50.0);
    }
    for (int k = 0; k < 4; k++) {
        int c = ok[k];
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s.good4[k], c);
    }
}

__global__ void k_ess_pick(const EssArgs a)
{
    if (blockIdx.x || threadIdx.x) return;
    const EssState& s = *a.state;
    EssResult& r = *a.result;
    r.n_inliers = s.max_good; r.iters = s.iters_run; r.ok = s.best_it >= 0 ? 1 : 0;
    if (s.best_it < 0) {
        for (int k = 0; k < 9; k++) { r.R[k] = (k % 4 == 0) ? 1.0 : 0.0; r.E[k] = 0.0; }
        r.t[0] = r.t[1] = r.t[2] = 0.0; r.n_good = 0;
        return;
    }
    const int* g = s.good4;
    int k = 3;                                    // recoverPose's order of preference on ties
    if (g[0] >= g[1] && g[0] >= g[2] && g[0] >= g[3]) k = 0;
    else if (g[1] >= g[0] && g[1] >= g[2] && g[1] >= g[3]) k = 1;
    else if (g[2] >= g[0] && g[2] >= g[1] && g[2] >= g[3]) k = 2;
    const double* R = a.pose + ((k & 1) ? 9 : 0);
    const double sgn = k >= 2 ? -1.0 : 1.0;
    for (int j = 0; j < 9; j++) { r.R[j] = R[j]; r.E[j] = a.pose[21 + j]; }
    for (int j = 0; j < 3; j++) r.t[j] = sgn * a.pose[18 + j];
    r.n_good = g[k];
}

int vo_launch_essential(const EssArgs& a, cudaStream_t s)
{
    int launches = 0;
    const int nb = (a.n + 127) / 128 > 0 ? (a.n + 127) / 128 : 1;
    k_ess_init<<<nb, 128, 0, s>>>(a); launches++;
    const int waves[4] = {0, 32, 128, a.max_iters};
    for (int w = 0; w < 3; w++) {
        const int it0 = waves[w], it1 = waves[w + 1] < a.max_iters ? waves[w + 1] : a.max_iters;
        if (it1 <= it0) break;
        k_ess_subsets<<<1, 32, 0, s>>>(a, it0, it1);
        k_ess_hypotheses<<<(it1 - it0 + 31) / 32, 32, 0, s>>>(a, it0, it1);
        k_ess_count<<<dim3(it1 - it0, 10), 128, 0, s>>>(a, it0, it1);
        k_ess_replay<<<1, 32, 0, s>>>(a, it0, it1);
        launches += 4;
    }
    k_ess_mask<<<nb, 128, 0, s>>>(a);
    k_ess_decompose<<<1, 32, 0, s>>>(a);
    k_ess_cheirality<<<nb, 128, 0, s>>>(a);
    k_ess_pick<<<1, 32, 0, s>>>(a);
    return launches + 4;
}
