// lk_ring_v3.cu -- the ROUND-1 version of K2, kept only for A/B measurement (vo_set_option("lk_kernel", 3)).
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
// window pixels are therefore split into 5 ordered chains (4 x 84 + 105 pixels); the addends of
// b are float(int pair sum) of columns (x, x+4).
//
// Work mapping (who computes which pixel): COLUMN STRIPS.  Lane L owns window column L>>1, rows
// 0..10 (L even) or 11..20 (L odd), plus up to 4 rows of one tail column (lanes 0..29).  A strip
// walks down its column, so the two byte taps of a row are fetched once (two aligned 32-bit loads
// + a funnel shift), serve as the bottom taps of one pixel and the top taps of the next, and feed
// the fixed-point bilinear interpolation as packed operands of dp2a.
//
// Summation (who adds): all addends are integers, so when the sum of |addend| over every chain is
// <= 2^24 each partial sum is exact and the chain totals are integer warp reductions (REDUX) --
// the fast path.  Otherwise the float addends are written to shared memory in chain order and one
// RUNNER lane per (quantity, chain) adds them strictly in order with 128-bit loads.  The A sums
// always take the faithful path (they pass 2^24 on any corner-like texture).
#include "common.cuh"
#include "lk_ring.h"
#define LK3_WARPS_PER_CTA 1
#define LK3_MIN_CTAS_PER_SM 17

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
#define CHS 116                 // floats per (quantity, chain) slot: >= 112 and CHS/4 odd, so that the 128-bit loads
                                // of the runner lanes fall into distinct bank groups (132 worked as well; 116 lets
                                // 17 CTAs fit in the SM's shared memory: 17 x (12416 + 1024 reserved) <= 228 KB)
#define CHN 15                  // slots: 3 quantities x 5 chains
struct __align__(128) WarpSmem {
    uint32_t dwin[DW * I_ROWS + 24];    // derivative window     (box 28 x 22 u32)  2464 -> 2560
    uint8_t iwin[IW * I_ROWS + 8];      // previous-image window, pitch 52          1144 -> 1152
    uint8_t jtile[IW * J_ROWS];         // next-image tile, pitch 52                1664
    float chain[CHN * CHS];             // chain-ordered float addends (faithful summation); ALSO the landing zone
                                        // of the dense TMA boxes (I at +0, J at +1152 bytes) before re-pitching
    uint64_t bar;                       // mbarrier for TMA completion
    uint64_t pad_[9];
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

// a.lo * b.byte0 + a.hi * b.byte1 + c with SIGNED 16-bit halves of a (w11 = 2^14 - w00 - w01 - w10 can be -1)
// and UNSIGNED bytes of b: the two horizontal taps of the fixed-point bilinear interpolation
__device__ __forceinline__ int dp2a_taps(unsigned w_pair, unsigned taps, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w_pair), "r"(taps), "r"(c));
    return d;
}

// sum over the 8 lanes of a SIMD chain: lanes that differ only in bit 0 (row half) and bits 3,4
// (column group) -- every lane ends up with its own chain's total
__device__ __forceinline__ unsigned chain_sum_u(unsigned v)
{
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 16);
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
__global__ void __launch_bounds__(LK3_WARPS_PER_CTA * 32, LK3_MIN_CTAS_PER_SM)
k_lk_ring_v3(const __grid_constant__ LkMaps maps, const LkArgs args)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp_in_cta = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    WarpSmem& sm = reinterpret_cast<WarpSmem*>(smem_raw)[warp_in_cta];

    const int gwarp = blockIdx.x * LK3_WARPS_PER_CTA + warp_in_cta;
    const int unit = gwarp / args.cap;
    const int f = gwarp - unit * args.cap;
    if (unit >= args.n_units) return;
    const int npts = args.n_pts ? args.n_pts[unit] : args.cap;
    if (f >= npts) return;

    // ---- static work mapping -------------------------------------------------------------------
    const int col = lane >> 1, half = lane & 1;
    const int r0 = half ? 11 : 0;                  // strip rows [r0, r0 + 11) (row 21 of the odd lanes is a dummy)
    const int chain = col & 3, cpos = col >> 2;    // SIMD chain and position inside the row's group of 4
    const bool has_tail = lane < 30;
    const int tcol = has_tail ? 16 + lane / 6 : 16;
    const int seg = lane % 6;
    const int tr0 = has_tail ? (seg < 3 ? seg * 4 : 12 + (seg - 3) * 3) : 0;     // tail rows [tr0, tr0 + tn)
    const int tn = has_tail ? (seg < 3 ? 4 : 3) : 0;
    // chain-buffer positions (floats).  Slot of (quantity q, chain c) = (5q + c) * CHS.
    //   A (setup)     : SIMD element (row, col) at row*4 + cpos ; tail element at row*5 + (tcol-16)
    //   b (iteration) : SIMD pair (col, col+4) of a row at row*2 + cpos/2 (written by the even-cpos lane) ; tail as A
    // Dummy elements (row 21 of odd lanes, unused tail rows, lanes 30/31) carry zero gradients: they
    // write 0.0f exactly onto the zero padding the runner lanes read (positions 84.., 42..43, 105..111).
    const int a_pos = chain * CHS + r0 * 4 + cpos;
    const int b_pos = chain * CHS + r0 * 2 + (cpos >> 1);
    const int t_pos = 4 * CHS + (has_tail ? tr0 * 5 + (tcol - 16) : 105 + (lane - 30) * 4);
    const int t_stride = has_tail ? 5 : 1;
    const int rc = lane % 5;                                     // runner lane L sums slot L (quantity L/5, chain L%5)
    const int a_nvec = lane < 15 ? (rc < 4 ? 21 : 28) : 0;
    const int b_nvec = lane < 10 ? (rc < 4 ? 11 : 28) : 0;
    const int run_base = lane * CHS;

    if (lane == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    uint8_t* const raw = reinterpret_cast<uint8_t*>(sm.chain);      // landing zone of the dense TMA boxes

    const size_t pbase = (size_t)unit * args.cap + f;
    float2 pt = args.pts_in[pbase];
    const float half_win = (VO_WIN - 1) * 0.5f;
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
            px -= half_win; py -= half_win;
            const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
            if (ipx < -VO_WIN || ipx >= lw || ipy < -VO_WIN || ipy >= lh) {
                if (level == 0) { status = 0; errv = 0.f; }
                continue;
            }
            // ---- stage windows: I (u8), dI (s16x2), J tile (u8) --------------------------------
            float npx = nxt.x - half_win, npy = nxt.y - half_win;
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
            float a = px - (float)ipx, b = py - (float)ipy;
            int w00, w01, w10, w11;
            bilinear_weights(a, b, w00, w01, w10, w11);
            if (args.use_tma) {
                mbar_wait(&sm.bar, phase); phase ^= 1;
                repitch(sm.iwin, raw + RAW_I_OFF, I_ROWS, lane);
                if (j_ok0) repitch(sm.jtile, raw + RAW_J_OFF, J_ROWS, lane);
                __syncwarp();
            }

            // ---- patch extraction: I (x32), Ix, Iy of the strip elements; A addends in chain order ----
            int Ipk[8];           // int16 patch intensities, two per register: strip 0..10, tail 11..14
            int dxy[15];          // lo16 = Ix, hi16 = Iy
            float A11, A12, A22;
            {
                const unsigned wt = (unsigned)w00 | ((unsigned)w01 << 16), wb = (unsigned)w10 | ((unsigned)(w11 & 0xffff) << 16);
                const int ox = ipx + VO_PAD - ibx, odx = ipx + VO_PAD - dbx;
#pragma unroll
                for (int part = 0; part < 2; part++) {
                    const int c0 = part ? tcol : col, row0 = part ? tr0 : r0, ne = part ? 4 : 11, nvalid = part ? tn : (half ? 10 : 11);
                    const int ioff = row0 * IW + ox + c0;                      // byte offset of the strip's first tap
                    const uint32_t* iw = reinterpret_cast<const uint32_t*>(sm.iwin) + (ioff >> 2);
                    const int ish = 8 * (ioff & 3);
                    const uint32_t* dw = sm.dwin + row0 * DW + odx + c0;
                    unsigned ptop = __funnelshift_r(iw[0], iw[1], ish);
                    unsigned d00 = dw[0], d01 = dw[1];
#pragma unroll
                    for (int k = 0; k < ne; k++) {
                        // the padding element of a short strip / segment would read one row past the windows: skip its loads
                        const bool live = k < nvalid || (part && !has_tail);
                        const unsigned pbot = live ? __funnelshift_r(iw[(k + 1) * (IW / 4)], iw[(k + 1) * (IW / 4) + 1], ish) : 0u;
                        const unsigned d10 = live ? dw[(k + 1) * DW] : 0u, d11 = live ? dw[(k + 1) * DW + 1] : 0u;
                        const int ival = (dp2a_taps(wb, pbot, dp2a_taps(wt, ptop, 1 << (W_BITS - 6))) >> (W_BITS - 5));
                        int ix = ((int)(short)(d00 & 0xffff) * w00 + (int)(short)(d01 & 0xffff) * w01 +
                                  (int)(short)(d10 & 0xffff) * w10 + (int)(short)(d11 & 0xffff) * w11 + (1 << (W_BITS - 1))) >> W_BITS;
                        int iy = (((int)d00 >> 16) * w00 + ((int)d01 >> 16) * w01 +
                                  ((int)d10 >> 16) * w10 + ((int)d11 >> 16) * w11 + (1 << (W_BITS - 1))) >> W_BITS;
                        if (k >= nvalid) { ix = 0; iy = 0; }
                        const int e = part ? 11 + k : k;
                        if (e & 1) Ipk[e >> 1] |= ival << 16; else Ipk[e >> 1] = ival & 0xffff;
                        dxy[e] = (ix & 0xffff) | (iy << 16);
                        const float fx = (float)ix, fy = (float)iy;
                        const int pos = part ? t_pos + k * t_stride : a_pos + k * 4;
                        // an unused tail row of a 3-row segment is the first row of the next lane's segment: no store
                        if (!part || k < tn || !has_tail) {
                            sm.chain[0 * 5 * CHS + pos] = __fmul_rn(fx, fx);
                            sm.chain[1 * 5 * CHS + pos] = __fmul_rn(fx, fy);
                            sm.chain[2 * 5 * CHS + pos] = __fmul_rn(fy, fy);
                        }
                        ptop = pbot; d00 = d10; d01 = d11;
                    }
                }
                __syncwarp();
                const float acc = run_chain(sm.chain, run_base, a_nvec);
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
                const unsigned wt = (unsigned)w00 | ((unsigned)w01 << 16), wb = (unsigned)w10 | ((unsigned)(w11 & 0xffff) << 16);
                int dpk[8];                       // int16 residuals, two per register (same element order as Ipk)
                int sxs = 0, sys = 0, sxt = 0, syt = 0;         // signed sums: strip (my SIMD chain) / tail
                unsigned axs = 0, ays = 0, axt = 0, ayt = 0;    // sums of |addend|
#pragma unroll
                for (int part = 0; part < 2; part++) {
                    const int c0 = part ? tcol : col, row0 = part ? tr0 : r0, ne = part ? 4 : 11, nvalid = part ? tn : (half ? 10 : 11);
                    const int joff = (ry + row0) * IW + rx + c0;
                    const uint32_t* jw = reinterpret_cast<const uint32_t*>(sm.jtile) + (joff >> 2);
                    const int jsh = 8 * (joff & 3);
                    unsigned ptop = __funnelshift_r(jw[0], jw[1], jsh);
#pragma unroll
                    for (int k = 0; k < ne; k++) {
                        const bool live = k < nvalid || (part && !has_tail);        // see the patch extraction
                        const unsigned pbot = live ? __funnelshift_r(jw[(k + 1) * (IW / 4)], jw[(k + 1) * (IW / 4) + 1], jsh) : 0u;
                        const int e = part ? 11 + k : k;
                        const int Iv = (e & 1) ? (Ipk[e >> 1] >> 16) : (int)(short)(Ipk[e >> 1] & 0xffff);
                        const int diff = (dp2a_taps(wb, pbot, dp2a_taps(wt, ptop, 1 << (W_BITS - 6))) >> (W_BITS - 5)) - Iv;
                        if (e & 1) dpk[e >> 1] |= diff << 16; else dpk[e >> 1] = diff & 0xffff;
                        const int vx = diff * (int)(short)(dxy[e] & 0xffff);      // gradients are 0 for dummy elements
                        const int vy = diff * (dxy[e] >> 16);
                        // |v| + acc in one VABSDIFF (|v| <= 2^28, 15 of them per lane: no overflow)
                        if (part) { sxt += vx; syt += vy; axt = __sad(vx, 0, axt); ayt = __sad(vy, 0, ayt); }
                        else { sxs += vx; sys += vy; axs = __sad(vx, 0, axs); ays = __sad(vy, 0, ays); }
                        ptop = pbot;
                    }
                }
                // per-chain totals (exact integers) and per-chain sums of |addend| (REDUX over the chain's lanes)
                const unsigned cax = chain_sum_u(axs), cay = chain_sum_u(ays);
                const unsigned tax = __reduce_add_sync(FULL, axt), tay = __reduce_add_sync(FULL, ayt);
                // |pair sum| <= |v0| + |v1|, so the bound is conservative for the SIMD chains
                const bool exact = __all_sync(FULL, cax <= (1u << 24) && cay <= (1u << 24) && tax <= (1u << 24) && tay <= (1u << 24));
                float ib1, ib2;
                if (exact) {
                    // every partial sum of every chain is an exactly representable integer
                    const float fx = (float)(int)chain_sum_u((unsigned)sxs), fy = (float)(int)chain_sum_u((unsigned)sys);
                    const float tx = (float)__reduce_add_sync(FULL, sxt), ty = (float)__reduce_add_sync(FULL, syt);
                    float c0 = __shfl_sync(FULL, fx, 0), c1 = __shfl_sync(FULL, fx, 2), c2 = __shfl_sync(FULL, fx, 4), c3 = __shfl_sync(FULL, fx, 6);
                    ib1 = __fadd_rn(tx, __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3)));
                    c0 = __shfl_sync(FULL, fy, 0); c1 = __shfl_sync(FULL, fy, 2); c2 = __shfl_sync(FULL, fy, 4); c3 = __shfl_sync(FULL, fy, 6);
                    ib2 = __fadd_rn(ty, __fadd_rn(__fadd_rn(c0, c2), __fadd_rn(c1, c3)));
                } else {
                    // faithful replay: float addends in chain order, runner lanes add them
#pragma unroll
                    for (int k = 0; k < 11; k++) {
                        const int d = (k & 1) ? (dpk[k >> 1] >> 16) : (int)(short)(dpk[k >> 1] & 0xffff);
                        const int vx = d * (int)(short)(dxy[k] & 0xffff), vy = d * (dxy[k] >> 16);
                        const int px2 = vx + __shfl_down_sync(FULL, vx, 8), py2 = vy + __shfl_down_sync(FULL, vy, 8);   // + column x+4
                        if (!(cpos & 1)) {
                            sm.chain[b_pos + k * 2] = (float)px2;
                            sm.chain[5 * CHS + b_pos + k * 2] = (float)py2;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int e = 11 + k;
                        const int d = (e & 1) ? (dpk[e >> 1] >> 16) : (int)(short)(dpk[e >> 1] & 0xffff);
                        if (k < tn || !has_tail) {
                            sm.chain[t_pos + k * t_stride] = (float)(d * (int)(short)(dxy[e] & 0xffff));
                            sm.chain[5 * CHS + t_pos + k * t_stride] = (float)(d * (dxy[e] >> 16));
                        }
                    }
                    __syncwarp();
                    const float acc = run_chain(sm.chain, run_base, b_nvec);
                    ib1 = combine_chains(acc, 0);
                    ib2 = combine_chains(acc, 1);
                    __syncwarp();
                }
                const float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
                const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
                const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
                npx = __fadd_rn(npx, dx); npy = __fadd_rn(npy, dy);
                nxt.x = __fadd_rn(npx, half_win); nxt.y = __fadd_rn(npy, half_win);
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
                float fxp = __fsub_rn(nxt.x, half_win), fyp = __fsub_rn(nxt.y, half_win);
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
                    const unsigned wt = (unsigned)w00 | ((unsigned)w01 << 16), wb = (unsigned)w10 | ((unsigned)(w11 & 0xffff) << 16);
                    // errval += |diff| is a plain row-major float sum of small integers
                    // (<= 441 * 8160 < 2^24): exact, so any order gives the same float.
                    int s = 0;
#pragma unroll
                    for (int part = 0; part < 2; part++) {
                        const int c0 = part ? tcol : col, row0 = part ? tr0 : r0, ne = part ? 4 : 11, nvalid = part ? tn : (half ? 10 : 11);
                        const int joff = (ry + row0) * IW + rx + c0;
                        const uint32_t* jw = reinterpret_cast<const uint32_t*>(sm.jtile) + (joff >> 2);
                        const int jsh = 8 * (joff & 3);
                        unsigned ptop = __funnelshift_r(jw[0], jw[1], jsh);
#pragma unroll
                        for (int k = 0; k < ne; k++) {
                            const bool live = k < nvalid || (part && !has_tail);
                            const unsigned pbot = live ? __funnelshift_r(jw[(k + 1) * (IW / 4)], jw[(k + 1) * (IW / 4) + 1], jsh) : 0u;
                            const int e = part ? 11 + k : k;
                            const int Iv = (e & 1) ? (Ipk[e >> 1] >> 16) : (int)(short)(Ipk[e >> 1] & 0xffff);
                            const int diff = (dp2a_taps(wb, pbot, dp2a_taps(wt, ptop, 1 << (W_BITS - 6))) >> (W_BITS - 5)) - Iv;
                            if (k < nvalid) s += abs(diff);
                            ptop = pbot;
                        }
                    }
                    s = __reduce_add_sync(FULL, s);
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
size_t vo_lk_smem_bytes_v3() { return sizeof(WarpSmem) * LK3_WARPS_PER_CTA; }

cudaError_t vo_lk_prepare_v3()
{
    return cudaFuncSetAttribute(k_lk_ring_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vo_lk_smem_bytes_v3());
}

cudaError_t vo_launch_lk_ring_v3(const LkMaps& maps, const LkArgs& args, cudaStream_t stream)
{
    const long warps = (long)args.n_units * args.cap;
    if (warps <= 0) return cudaSuccess;
    const int ctas = (int)((warps + LK3_WARPS_PER_CTA - 1) / LK3_WARPS_PER_CTA);
    k_lk_ring_v3<<<ctas, LK3_WARPS_PER_CTA * 32, vo_lk_smem_bytes_v3(), stream>>>(maps, args);
    return cudaGetLastError();
}
