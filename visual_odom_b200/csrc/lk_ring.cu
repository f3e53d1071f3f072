// lk_ring.cu -- K2: the pyramidal Lucas-Kanade kernel (the hot kernel of this library).
//
// Replaces the four chained cv::calcOpticalFlowPyrLK calls of the reference's circularMatching()
// (reference src/feature.cpp:136-139; window 21x21 :127, 30 iterations / eps 0.01 :128,
// maxLevel 3, minEigThreshold 1e-3 :136).  Arithmetic restated in oracle/lk_ref.c (lk_track),
// which is pinned bit-for-bit against cv2 4.13.0; this kernel reproduces the same bits:
//   * fixed-point bilinear patches (weights cvRound(w*2^14), DESCALE by 9 / 14)
//   * float32 normal equations accumulated in OpenCV's 4-SIMD-lane + scalar-tail order
//     (the order matters once partial sums pass 2^24; see "summation chains" below)
//   * per-level next = next*2 propagation, level-0-only status writes, final bounds re-check.
//
// Parallelisation: ONE WARP PER FEATURE, and one launch runs the WHOLE ring (up to 4 chained
// calls x all pyramid levels) for every feature of every unit -- a feature's track never
// depends on another feature, so nothing forces a launch boundary between levels or calls.
//
// Staging: per level, lane 0 issues three TMA (cp.async.bulk.tensor.3d) box loads into the
// warp's private shared memory: the 32x32 u8 window of the previous image, the 24x22 s16x2
// window of its Scharr derivative, and a 32x32 u8 tile of the next image around the current
// estimate (re-issued only if the 22x22 search window drifts out of the tile).  The planes are
// physically padded (see common.cuh), so no box ever needs border handling.
//
// Summation chains: OpenCV accumulates A11/A12/A22 and b1/b2 in float32 with 4 SIMD lanes over
// columns 0..15 (lane = x & 3) and a scalar tail over columns 16..20, rows outermost.  The 441
// window pixels are therefore split into 5 ordered chains (4 x 84 + 105 pixels).  Lanes 0..23
// own the four SIMD chains (6 lanes x 14 consecutive chain elements), lanes 24..31 own the tail
// (8 lanes x 14).  All addends are integers, so when the sum of |addend| over a chain is
// <= 2^24 every partial sum is exact and the chain total is order independent: that fast path
// uses integer shuffles.  Otherwise the chain is replayed faithfully: lane s adds its 14
// elements in order and hands the running float to lane s+1.
#include "common.cuh"
#include "lk_ring.h"

#define FULL 0xffffffffu
#define W_BITS 14

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// per-warp shared memory
// TMA tile loads need a 16-byte aligned global start address, so every box starts at the
// 16-byte boundary at or below the wanted column and is 16 bytes wider than the data it must hold:
//   u8 windows  : 48 bytes wide  (<= 15 bytes of lead-in + 22 (I) / 32 (J) bytes of payload)
//   s16x2 window: 28 elements wide (<= 3 elements of lead-in + 22)
#define RAW_W 48                // row pitch of the u8 boxes as TMA writes them (dense)
#define IW 52                   // row pitch of the u8 windows the kernel reads: 13 words (odd), so the rows a
                                // warp touches in one LDS fall into different banks (48 = 12 words made rows
                                // r and r+8 collide: 2.5 wavefronts per byte load, the kernel was LSU bound)
#define DW 28                   // row pitch (uint32) of the derivative box
#define I_ROWS 22
#define J_ROWS 32
#define CHS 132                 // floats per (quantity, chain) slot: >= 112 and == 4 (mod 32) so that the
                                // 128-bit loads of the runner lanes fall into distinct bank groups
#define CHN 15                  // slots: 3 quantities x 5 chains
struct __align__(128) WarpSmem {
    uint32_t dwin[DW * I_ROWS + 24];    // derivative window     (box 28 x 22 u32)  2464 -> 2560
    uint8_t iwin[IW * I_ROWS + 8];      // previous-image window, pitch 52          1144 -> 1152
    uint8_t jtile[IW * J_ROWS];         // next-image tile, pitch 52                1664
    float chain[CHN * CHS];             // chain-ordered float addends (faithful summation); ALSO the landing zone
                                        // of the dense TMA boxes (I at +0, J at +1152 bytes) before re-pitching
    uint64_t bar;                       // mbarrier for TMA completion
    uint64_t pad_[1];
};
#define RAW_I_OFF 0
#define RAW_J_OFF 1152
static_assert(sizeof(WarpSmem) % 128 == 0, "WarpSmem must keep 128B alignment");

#define I_BYTES (RAW_W * I_ROWS)
#define J_BYTES (RAW_W * J_ROWS)
#define D_BYTES (DW * I_ROWS * 4)

// plain-load staging of one box (debug / A-B path): rows x row_bytes from a padded plane
__device__ __forceinline__ void ldg_box_u8(uint8_t* dst, const uint8_t* plane, int pitch, int x, int y, int rows, int lane)
{
    const uint8_t* src = plane + (size_t)y * pitch + x;
    for (int r = 0; r < rows; r++) {
        dst[r * IW + lane] = __ldg(src + (size_t)r * pitch + lane);
        if (lane < RAW_W - 32) dst[r * IW + 32 + lane] = __ldg(src + (size_t)r * pitch + 32 + lane);
    }
}
// dense TMA box (rows x 48 B) -> window with row pitch IW: each lane moves 16-byte chunks
__device__ __forceinline__ void repitch(uint8_t* dst, const uint8_t* raw, int rows, int lane)
{
    for (int c = lane; c < rows * 3; c += 32) {
        const int row = c / 3, part = c - row * 3;
        const uint4 v = *reinterpret_cast<const uint4*>(raw + c * 16);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + row * IW + part * 16);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}
__device__ __forceinline__ void ldg_box_u32(uint32_t* dst, const uint32_t* plane, int pitch, int x, int y, int lane)
{
    const uint32_t* src = plane + (size_t)y * pitch + x;
    if (lane < DW)
        for (int r = 0; r < I_ROWS; r++) dst[r * DW + lane] = __ldg(src + (size_t)r * pitch + lane);
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10, int& w11)
{
    w00 = __float2int_rn((1.f - a) * (1.f - b) * (float)(1 << W_BITS));
    w01 = __float2int_rn(a * (1.f - b) * (float)(1 << W_BITS));
    w10 = __float2int_rn((1.f - a) * b * (float)(1 << W_BITS));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// tree reduction over the lanes of one chain group (6 or 8 consecutive lanes), result at sub==0
__device__ __forceinline__ int group_sum(int v, int sub, int gsize)
{
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        int o = __shfl_down_sync(FULL, v, d);
        if (sub + d < gsize) v += o;
    }
    return v;
}

// Faithful float chains.  The addends of every chain lie contiguously (in chain order, zero padded
// to a multiple of 4) in shared memory; one RUNNER lane per (quantity, chain) adds them strictly
// in order with 128-bit loads.  5 chains x NQ quantities run concurrently on 5*NQ lanes.
//   base  : this lane's chain start (floats), nvec : float4 count (0 for non-runner lanes)
__device__ __forceinline__ float run_chain(const float* buf, int base, int nvec)
{
    float acc = 0.f;
    const float4* p = reinterpret_cast<const float4*>(buf + base);
#pragma unroll 4
    for (int v = 0; v < nvec; v++) {
        const float4 t = p[v];
        acc = __fadd_rn(acc, t.x); acc = __fadd_rn(acc, t.y); acc = __fadd_rn(acc, t.z); acc = __fadd_rn(acc, t.w);
    }
    return acc;
}
// total of quantity q from the runner lanes 5q..5q+4:  tail + ((c0 + c2) + (c1 + c3))
__device__ __forceinline__ float combine_chains(float acc, int q)
{
    const float c0 = __shfl_sync(FULL, acc, 5 * q), c1 = __shfl_sync(FULL, acc, 5 * q + 1),
                c2 = __shfl_sync(FULL, acc, 5 * q + 2), c3 = __shfl_sync(FULL, acc, 5 * q + 3),
                t = __shfl_sync(FULL, acc, 5 * q + 4);
    return __fadd_rn(t, __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3)));
}

} // namespace

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LK_WARPS_PER_CTA * 32, LK_MIN_CTAS_PER_SM)
k_lk_ring(const __grid_constant__ LkMaps maps, const LkArgs args)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp_in_cta = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    WarpSmem& sm = reinterpret_cast<WarpSmem*>(smem_raw)[warp_in_cta];

    const int gwarp = blockIdx.x * LK_WARPS_PER_CTA + warp_in_cta;
    const int unit = gwarp / args.cap;
    const int f = gwarp - unit * args.cap;
    if (unit >= args.n_units) return;
    const int npts = args.n_pts ? args.n_pts[unit] : args.cap;
    if (f >= npts) return;

    // static lane -> chain-element assignment: lanes 0..23 = 4 SIMD chains x 6 lanes, 24..31 = tail
    const bool tail = lane >= 24;
    const int chain = tail ? 4 : lane / 6;
    const int sub = tail ? lane - 24 : lane - chain * 6;
    const int gsize = tail ? 8 : 6;
    const int e0 = sub * 14;                       // first chain element of this lane
    const int nel = tail ? (105 - e0 < 14 ? 105 - e0 : 14) : 14;   // elements owned (tail last lane: 7)
    int woff[14];                                  // element k -> offset (row*IW + col) in a u8 box
#pragma unroll
    for (int k = 0; k < 14; k++) {
        int e = e0 + k, row, col;
        if (tail) { row = e / 5; col = 16 + e - row * 5; }
        else { row = e >> 2; col = chain + 4 * (e & 3); }
        if (k >= nel) { row = 0; col = 0; }
        woff[k] = row * IW + col;
    }
    // chain-buffer slots: where this lane writes its addends, and (runner lanes) what it sums
    // slot of (quantity q, chain c) = (5q + c) * CHS
    const int a_slot = chain * CHS + sub * 14;                   // setup: 14 floats per lane and quantity
    const int b_slot = chain * CHS + (tail ? sub * 14 : sub * 8);// iteration: 7 pairs (+1 zero) / 14 singles
    const int rc = lane % 5;                                     // runner lane L sums slot L (quantity L/5, chain L%5)
    const int a_base = lane * CHS, a_nvec = lane < 15 ? (rc < 4 ? 21 : 28) : 0;
    const int b_base = lane * CHS, b_nvec = lane < 10 ? (rc < 4 ? 12 : 28) : 0;

    if (lane == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    uint8_t* const raw = reinterpret_cast<uint8_t*>(sm.chain);      // landing zone of the dense TMA boxes

    const size_t pbase = (size_t)unit * args.cap + f;
    float2 pt = args.pts_in[pbase];
    const float half = (VO_WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int max_level = args.nlevels - 1;

    for (int call = 0; call < args.ncalls; call++) {
        const int img_prev = args.img_plane0 + unit * args.imgs_per_unit + args.img_prev[call];
        const int img_next = args.img_plane0 + unit * args.imgs_per_unit + args.img_next[call];
        float2 nxt = make_float2(0.f, 0.f);
        int status = 1;
        float errv = 0.f;

        for (int level = max_level; level >= 0; level--) {
            const int lw = args.lw[level], lh = args.lh[level];
            const float sc = 1.f / (float)(1 << level);
            float px = pt.x * sc, py = pt.y * sc;
            if (level == max_level) { nxt.x = px; nxt.y = py; }
            else { nxt.x = nxt.x * 2.f; nxt.y = nxt.y * 2.f; }
            px -= half; py -= half;
            const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
            if (ipx < -VO_WIN || ipx >= lw || ipy < -VO_WIN || ipy >= lh) {
                if (level == 0) { status = 0; errv = 0.f; }
                continue;
            }
            // ---- stage windows: I (u8), dI (s16x2), J tile (u8) --------------------------------
            float npx = nxt.x - half, npy = nxt.y - half;
            int inx = __float2int_rd(npx), iny = __float2int_rd(npy);
            const bool j_ok0 = !(inx < -VO_WIN || inx >= lw || iny < -VO_WIN || iny >= lh);
            // box origins in padded-plane coordinates, x snapped down to the 16-byte boundary
            const int ibx = (ipx + VO_PAD) & ~15, iby = ipy + VO_PAD;       // u8 window of I
            const int dbx = (ipx + VO_PAD) & ~3;                            // derivative window (4 elements = 16 B)
            int jbx = (inx - 5 + VO_PAD) & ~15, jby = iny - 5 + VO_PAD;     // tile of J
            __syncwarp();
            if (args.use_tma) {
                if (lane == 0) {
                    mbar_expect_tx(&sm.bar, I_BYTES + D_BYTES + (j_ok0 ? J_BYTES : 0));
                    tma_load_3d(raw + RAW_I_OFF, &maps.img_i[level], &sm.bar, ibx, iby, img_prev);
                    tma_load_3d(sm.dwin, &maps.der[level], &sm.bar, dbx, iby, img_prev);
                    if (j_ok0)
                        tma_load_3d(raw + RAW_J_OFF, &maps.img_j[level], &sm.bar, jbx, jby, img_next);
                }
            } else {
                ldg_box_u8(sm.iwin, args.img_base[level] + args.plane[level] * img_prev, args.pitch[level], ibx, iby, I_ROWS, lane);
                ldg_box_u32(sm.dwin, args.der_base[level] + args.plane[level] * img_prev, args.pitch[level], dbx, iby, lane);
                if (j_ok0)
                    ldg_box_u8(sm.jtile, args.img_base[level] + args.plane[level] * img_next, args.pitch[level], jbx, jby, J_ROWS, lane);
                __syncwarp();
            }
            const uint8_t* ib = sm.iwin + (ipx + VO_PAD - ibx);
            const uint32_t* db = sm.dwin + (ipx + VO_PAD - dbx);
            float a = px - (float)ipx, b = py - (float)ipy;
            int w00, w01, w10, w11;
            bilinear_weights(a, b, w00, w01, w10, w11);
            if (args.use_tma) {
                mbar_wait(&sm.bar, phase); phase ^= 1;
                repitch(sm.iwin, raw + RAW_I_OFF, I_ROWS, lane);
                if (j_ok0) repitch(sm.jtile, raw + RAW_J_OFF, J_ROWS, lane);
                __syncwarp();
            }

            // ---- patch extraction: I (x32), Ix, Iy for the 14 owned elements ------------------
            int Ipk[7];           // two int16 patch intensities per register
            int dxy[14];          // lo16 = Ix, hi16 = Iy
            float A11, A12, A22;
            {
#pragma unroll
                for (int k = 0; k < 14; k++) {
                    const uint8_t* s0 = ib + woff[k];
                    const int ival = (s0[0] * w00 + s0[1] * w01 + s0[IW] * w10 + s0[IW + 1] * w11 + (1 << (W_BITS - 6))) >> (W_BITS - 5);
                    const int row = woff[k] / IW, col = woff[k] - row * IW;
                    const uint32_t* d0 = db + row * DW + col;
                    const uint32_t d00 = d0[0], d01 = d0[1], d10 = d0[DW], d11 = d0[DW + 1];
                    int ix = ((int)(short)(d00 & 0xffff) * w00 + (int)(short)(d01 & 0xffff) * w01 +
                              (int)(short)(d10 & 0xffff) * w10 + (int)(short)(d11 & 0xffff) * w11 + (1 << (W_BITS - 1))) >> W_BITS;
                    int iy = (((int)d00 >> 16) * w00 + ((int)d01 >> 16) * w01 +
                              ((int)d10 >> 16) * w10 + ((int)d11 >> 16) * w11 + (1 << (W_BITS - 1))) >> W_BITS;
                    if (k >= nel) { ix = 0; iy = 0; }
                    if (k & 1) Ipk[k >> 1] |= ival << 16; else Ipk[k >> 1] = ival & 0xffff;
                    dxy[k] = (ix & 0xffff) | (iy << 16);
                    const float fx = (float)ix, fy = (float)iy;
                    // chain-ordered addends of A11 / A12 / A22 (zero for the unused tail slots)
                    sm.chain[0 * 5 * CHS + a_slot + k] = __fmul_rn(fx, fx);
                    sm.chain[1 * 5 * CHS + a_slot + k] = __fmul_rn(fx, fy);
                    sm.chain[2 * 5 * CHS + a_slot + k] = __fmul_rn(fy, fy);
                }
                __syncwarp();
                const float acc = run_chain(sm.chain, a_base, a_nvec);
                const float iA11 = combine_chains(acc, 0), iA12 = combine_chains(acc, 1), iA22 = combine_chains(acc, 2);
                A11 = __fmul_rn(iA11, FLT_SCALE); A12 = __fmul_rn(iA12, FLT_SCALE); A22 = __fmul_rn(iA22, FLT_SCALE);
                __syncwarp();
            }
            float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
            {
                float dd = __fsub_rn(A11, A22);
                float rad = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
                float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * VO_WIN * VO_WIN));
                if ((double)minEig < args.min_eig || D < 1.1920928955078125e-07f) {
                    if (level == 0) status = 0;
                    continue;
                }
            }
            D = __fdiv_rn(1.f, D);

            // ---- Newton iterations ------------------------------------------------------------
            float pdx = 0.f, pdy = 0.f;
            bool tile_valid = j_ok0;
            for (int j = 0; j < args.max_iters; j++) {
                inx = __float2int_rd(npx); iny = __float2int_rd(npy);
                if (inx < -VO_WIN || inx >= lw || iny < -VO_WIN || iny >= lh) {
                    if (level == 0) status = 0;
                    break;
                }
                int rx = inx + VO_PAD - jbx, ry = iny + VO_PAD - jby;   // window origin inside the tile
                if (!tile_valid || rx < 0 || ry < 0 || rx > RAW_W - 22 || ry > J_ROWS - 22) {
                    jbx = (inx - 5 + VO_PAD) & ~15; jby = iny - 5 + VO_PAD;
                    rx = inx + VO_PAD - jbx; ry = 5;
                    __syncwarp();
                    if (args.use_tma) {
                        if (lane == 0) {
                            mbar_expect_tx(&sm.bar, J_BYTES);
                            tma_load_3d(raw + RAW_J_OFF, &maps.img_j[level], &sm.bar, jbx, jby, img_next);
                        }
                        mbar_wait(&sm.bar, phase); phase ^= 1;
                        repitch(sm.jtile, raw + RAW_J_OFF, J_ROWS, lane);
                        __syncwarp();
                    } else {
                        ldg_box_u8(sm.jtile, args.img_base[level] + args.plane[level] * img_next, args.pitch[level], jbx, jby, J_ROWS, lane);
                        __syncwarp();
                    }
                    tile_valid = true;
                }
                a = npx - (float)inx; b = npy - (float)iny;
                bilinear_weights(a, b, w00, w01, w10, w11);
                const uint8_t* jb = sm.jtile + ry * IW + rx;
                int dpk[7];                       // two int16 residuals per register
                int sx = 0, sy = 0;
                unsigned ax = 0, ay = 0;
#pragma unroll
                for (int k = 0; k < 14; k++) {
                    const uint8_t* s0 = jb + woff[k];
                    const int Iv = (k & 1) ? (Ipk[k >> 1] >> 16) : (int)(short)(Ipk[k >> 1] & 0xffff);
                    const int diff = ((s0[0] * w00 + s0[1] * w01 + s0[IW] * w10 + s0[IW + 1] * w11 + (1 << (W_BITS - 6))) >> (W_BITS - 5)) - Iv;
                    if (k & 1) dpk[k >> 1] |= diff << 16; else dpk[k >> 1] = diff & 0xffff;
                    const int vx = diff * (int)(short)(dxy[k] & 0xffff);      // Ix, Iy are 0 for unused slots
                    const int vy = diff * (dxy[k] >> 16);
                    sx += vx; sy += vy;
                    ax += (unsigned)abs(vx); ay += (unsigned)abs(vy);
                }
                // per-chain totals (exact integers) and per-chain sum of |addend|
                const int csx = group_sum(sx, sub, gsize), csy = group_sum(sy, sub, gsize);
                const int cax = group_sum((int)ax, sub, gsize), cay = group_sum((int)ay, sub, gsize);
                // |pair sum| <= |v0|+|v1|, so the bound is conservative for the SIMD chains
                const bool exact = __all_sync(FULL, (sub != 0) || ((unsigned)cax <= (1u << 24) && (unsigned)cay <= (1u << 24)));
                float ib1, ib2;
                if (exact) {
                    // every partial sum of every chain is an exactly representable integer:
                    // chain totals live in lanes 0, 6, 12, 18 (SIMD) and 24 (tail)
                    const float fx = (float)csx, fy = (float)csy;
                    float c0 = __shfl_sync(FULL, fx, 0), c1 = __shfl_sync(FULL, fx, 6), c2 = __shfl_sync(FULL, fx, 12),
                          c3 = __shfl_sync(FULL, fx, 18), t = __shfl_sync(FULL, fx, 24);
                    ib1 = __fadd_rn(t, __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3)));
                    c0 = __shfl_sync(FULL, fy, 0); c1 = __shfl_sync(FULL, fy, 6); c2 = __shfl_sync(FULL, fy, 12);
                    c3 = __shfl_sync(FULL, fy, 18); t = __shfl_sync(FULL, fy, 24);
                    ib2 = __fadd_rn(t, __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3)));
                } else {
                    // faithful replay: write the float addends in chain order, runner lanes add them
                    float* cx = sm.chain + b_slot;
                    float* cy = sm.chain + 5 * CHS + b_slot;
                    if (!tail) {
#pragma unroll
                        for (int k = 0; k < 7; k++) {
                            const int d0 = (int)(short)(dpk[k] & 0xffff), d1 = dpk[k] >> 16;
                            cx[k] = (float)(d0 * (int)(short)(dxy[2 * k] & 0xffff) + d1 * (int)(short)(dxy[2 * k + 1] & 0xffff));
                            cy[k] = (float)(d0 * (dxy[2 * k] >> 16) + d1 * (dxy[2 * k + 1] >> 16));
                        }
                        cx[7] = 0.f; cy[7] = 0.f;
                    } else {
#pragma unroll
                        for (int k = 0; k < 14; k++) {
                            const int d = (k & 1) ? (dpk[k >> 1] >> 16) : (int)(short)(dpk[k >> 1] & 0xffff);
                            cx[k] = (float)(d * (int)(short)(dxy[k] & 0xffff));
                            cy[k] = (float)(d * (dxy[k] >> 16));
                        }
                    }
                    __syncwarp();
                    const float acc = run_chain(sm.chain, b_base, b_nvec);
                    ib1 = combine_chains(acc, 0);
                    ib2 = combine_chains(acc, 1);
                    __syncwarp();
                }
                const float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
                const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
                const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
                npx = __fadd_rn(npx, dx); npy = __fadd_rn(npy, dy);
                nxt.x = __fadd_rn(npx, half); nxt.y = __fadd_rn(npy, half);
                if ((double)dx * (double)dx + (double)dy * (double)dy <= args.eps2) break;
                if (j > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
                    nxt.x = __fsub_rn(nxt.x, __fmul_rn(dx, 0.5f));
                    nxt.y = __fsub_rn(nxt.y, __fmul_rn(dy, 0.5f));
                    break;
                }
                pdx = dx; pdy = dy;
            }

            // ---- level 0 epilogue: final bounds re-check (+ err when requested) ---------------
            if (level == 0 && status) {
                float fxp = __fsub_rn(nxt.x, half), fyp = __fsub_rn(nxt.y, half);
                inx = __float2int_rd(fxp); iny = __float2int_rd(fyp);
                if (inx < -VO_WIN || inx >= lw || iny < -VO_WIN || iny >= lh) {
                    status = 0;
                } else if (args.err_out) {
                    int rx = inx + VO_PAD - jbx, ry = iny + VO_PAD - jby;
                    if (!tile_valid || rx < 0 || ry < 0 || rx > RAW_W - 22 || ry > J_ROWS - 22) {
                        jbx = (inx - 5 + VO_PAD) & ~15; jby = iny - 5 + VO_PAD;
                        rx = inx + VO_PAD - jbx; ry = 5;
                        __syncwarp();
                        if (args.use_tma) {
                            if (lane == 0) {
                                mbar_expect_tx(&sm.bar, J_BYTES);
                                tma_load_3d(raw + RAW_J_OFF, &maps.img_j[0], &sm.bar, jbx, jby, img_next);
                            }
                            mbar_wait(&sm.bar, phase); phase ^= 1;
                            repitch(sm.jtile, raw + RAW_J_OFF, J_ROWS, lane);
                            __syncwarp();
                        } else {
                            ldg_box_u8(sm.jtile, args.img_base[0] + args.plane[0] * img_next, args.pitch[0], jbx, jby, J_ROWS, lane);
                            __syncwarp();
                        }
                    }
                    a = fxp - (float)inx; b = fyp - (float)iny;
                    bilinear_weights(a, b, w00, w01, w10, w11);
                    const uint8_t* jb = sm.jtile + ry * IW + rx;
                    // errval += |diff| is a plain row-major float sum of small integers
                    // (<= 441 * 8160 < 2^24): exact, so any order gives the same float.
                    int s = 0;
#pragma unroll
                    for (int k = 0; k < 14; k++) {
                        const uint8_t* s0 = jb + woff[k];
                        const int Iv = (k & 1) ? (Ipk[k >> 1] >> 16) : (int)(short)(Ipk[k >> 1] & 0xffff);
                        const int diff = ((s0[0] * w00 + s0[1] * w01 + s0[IW] * w10 + s0[IW + 1] * w11 + (1 << (W_BITS - 6))) >> (W_BITS - 5)) - Iv;
                        if (k < nel) s += abs(diff);
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(FULL, s, d);
                    errv = __fdiv_rn(__fmul_rn((float)s, 1.f), (float)(32 * VO_WIN * VO_WIN));
                }
            }
        } // level

        if (lane == 0) {
            const size_t o = (size_t)call * args.call_stride + pbase;
            args.pts_out[o] = nxt;
            args.status_out[o] = (uint8_t)status;
            if (args.err_out) args.err_out[o] = errv;
        }
        pt = nxt;
    } // call
}

// ---------------------------------------------------------------------------------------------
size_t vo_lk_smem_bytes() { return sizeof(WarpSmem) * LK_WARPS_PER_CTA; }

cudaError_t vo_lk_prepare()
{
    return cudaFuncSetAttribute(k_lk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vo_lk_smem_bytes());
}

cudaError_t vo_launch_lk_ring(const LkMaps& maps, const LkArgs& args, cudaStream_t stream)
{
    const long warps = (long)args.n_units * args.cap;
    if (warps <= 0) return cudaSuccess;
    const int ctas = (int)((warps + LK_WARPS_PER_CTA - 1) / LK_WARPS_PER_CTA);
    k_lk_ring<<<ctas, LK_WARPS_PER_CTA * 32, vo_lk_smem_bytes(), stream>>>(maps, args);
    return cudaGetLastError();
}
