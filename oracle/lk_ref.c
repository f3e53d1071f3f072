/*
 * oracle/lk_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement of the arithmetic behind the reference's
 *   cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err,
 *                            Size(21,21), 3, TermCriteria(COUNT+EPS,30,0.01), 0, 0.001)
 * call sites (reference src/feature.cpp:136-139, src/feature.cpp:72).  The
 * arithmetic itself lives in OpenCV (un-vendored third-party dependency; this
 * build is pinned to OpenCV 4.13.0 as installed, modules/video/src/lkpyramid.cpp,
 * modules/imgproc/src/pyramids.cpp -- sources are not on disk, the published
 * algorithm is restated here and pinned bit-for-bit against cv2 4.13.0 by
 * tests/test_oracle_lk.py).
 *
 * Restated pieces:
 *   pyr_down_u8        5x5 [1 4 6 4 1]^2 REFLECT_101, (s+128)>>8, dst=((w+1)/2,(h+1)/2)
 *   scharr_deriv       int16x2 interleaved Scharr derivative, REFLECT_101
 *   lk_track           per-point coarse-to-fine LK with fixed-point patches,
 *                      float32 normal equations, OpenCV's SIMD-lane float
 *                      summation order (so the result is bit-identical, not
 *                      merely close), status / err semantics.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* cv::pyrDown for CV_8UC1, BORDER_REFLECT_101 (default). */
void pyr_down_u8(const uint8_t *src, int w, int h, int sstep,
                 uint8_t *dst, int dstep)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    static const int k[5] = {1, 4, 6, 4, 1};
    for (int y = 0; y < dh; y++) {
        for (int x = 0; x < dw; x++) {
            int s = 0;
            for (int j = 0; j < 5; j++) {
                int sy = reflect101(2 * y + j - 2, h);
                int r = 0;
                for (int i = 0; i < 5; i++) {
                    int sx = reflect101(2 * x + i - 2, w);
                    r += k[i] * src[sy * sstep + sx];
                }
                s += k[j] * r;
            }
            dst[y * dstep + x] = (uint8_t)((s + 128) >> 8);
        }
    }
}

/* calcScharrDeriv: d[2*x] = dI/dx, d[2*x+1] = dI/dy (int16), REFLECT_101. */
void scharr_deriv(const uint8_t *src, int w, int h, int sstep,
                  int16_t *dst, int dstep /* in int16 elements */)
{
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = src + reflect101(y - 1, h) * sstep;
        const uint8_t *r1 = src + y * sstep;
        const uint8_t *r2 = src + reflect101(y + 1, h) * sstep;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10;
            int t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
            int t1m = r2[xm] - r0[xm];
            int t1c = r2[x] - r0[x];
            int t1p = r2[xp] - r0[xp];
            dst[y * dstep + 2 * x] = (int16_t)(t0p - t0m);
            dst[y * dstep + 2 * x + 1] = (int16_t)((t1m + t1p) * 3 + t1c * 10);
        }
    }
}

/* ------------------------------------------------------------------------- */

typedef struct {
    int w, h;           /* image size of this level              */
    int pad;            /* border of the padded copies            */
    int istep;          /* row stride of padded u8 image          */
    uint8_t *img;       /* padded REFLECT_101 image, origin at (pad,pad) */
    int dstep;          /* row stride (int16 elements) of padded derivative */
    int16_t *deriv;     /* zero-padded Scharr derivative (interleaved) or NULL */
} lk_level_t;

typedef struct {
    int nlevels;        /* number of pyramid images (= effective maxLevel+1) */
    lk_level_t lv[16];
} lk_pyr_t;

static void pad_reflect101(const uint8_t *src, int w, int h, int sstep,
                           uint8_t *dst, int pad, int dstep)
{
    for (int y = -pad; y < h + pad; y++) {
        int sy = reflect101(y, h);
        for (int x = -pad; x < w + pad; x++) {
            int sx = reflect101(x, w);
            dst[(y + pad) * dstep + (x + pad)] = src[sy * sstep + sx];
        }
    }
}

/* buildOpticalFlowPyramid(img, pyr, winSize, maxLevel, withDerivatives, REFLECT_101, CONSTANT) */
lk_pyr_t *lk_pyr_build(const uint8_t *img, int w, int h, int step,
                       int win, int max_level, int with_deriv)
{
    lk_pyr_t *p = (lk_pyr_t *)calloc(1, sizeof(lk_pyr_t));
    int pad = win;
    const uint8_t *cur = img;
    int cw = w, ch = h, cstep = step;
    uint8_t *tmp_prev = NULL;
    for (int l = 0; l <= max_level; l++) {
        uint8_t *tmp = NULL;
        if (l > 0) {
            int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
            /* OpenCV stops when the level is not larger than the window */
            if (nw <= win || nh <= win) break;
            tmp = (uint8_t *)malloc((size_t)nw * nh);
            pyr_down_u8(cur, cw, ch, cstep, tmp, nw);
            cur = tmp; cw = nw; ch = nh; cstep = nw;
        }
        lk_level_t *L = &p->lv[l];
        L->w = cw; L->h = ch; L->pad = pad;
        L->istep = cw + 2 * pad;
        L->img = (uint8_t *)malloc((size_t)L->istep * (ch + 2 * pad));
        pad_reflect101(cur, cw, ch, cstep, L->img, pad, L->istep);
        if (with_deriv) {
            L->dstep = 2 * (cw + 2 * pad);
            L->deriv = (int16_t *)calloc((size_t)L->dstep * (ch + 2 * pad), sizeof(int16_t));
            scharr_deriv(cur, cw, ch, cstep,
                         L->deriv + pad * L->dstep + 2 * pad, L->dstep);
        }
        p->nlevels = l + 1;
        if (tmp_prev) free(tmp_prev);
        tmp_prev = tmp;
    }
    if (tmp_prev) free(tmp_prev);
    return p;
}

void lk_pyr_free(lk_pyr_t *p)
{
    if (!p) return;
    for (int l = 0; l < p->nlevels; l++) { free(p->lv[l].img); free(p->lv[l].deriv); }
    free(p);
}

int lk_pyr_nlevels(const lk_pyr_t *p) { return p->nlevels; }
void lk_pyr_level_info(const lk_pyr_t *p, int l, int *w, int *h, int *pad, int *istep)
{ *w = p->lv[l].w; *h = p->lv[l].h; *pad = p->lv[l].pad; *istep = p->lv[l].istep; }
const uint8_t *lk_pyr_level_img(const lk_pyr_t *p, int l) { return p->lv[l].img; }
const int16_t *lk_pyr_level_deriv(const lk_pyr_t *p, int l) { return p->lv[l].deriv; }

/* ------------------------------------------------------------------------- */
/* accumulation-order selector (the pinned one is mode 0; the others exist so
 * tests can show that the order matters and that mode 0 is the cv2 one)      */
static int g_sum_mode = 0;
void lk_set_sum_mode(int m) { g_sum_mode = m; }

/* Optional bookkeeping for kernel design (not part of the checker): how often are all partial sums of the b chains
 * exactly representable?  [0] iterations, [1] sum|addend| <= 2^24 on every chain (the kernel's current test),
 * [2] strip bound: sum over the kernel's strips of max|running strip sum| <= 2^24 and every addend <= 2^24,
 * [3] truth: every chain prefix and every addend <= 2^24 in magnitude. */
static long long g_stats[8];
static int g_stats_on = 0;
void lk_stats_enable(int on) { g_stats_on = on; for (int i = 0; i < 8; i++) g_stats[i] = 0; }
void lk_stats_get(long long *out) { for (int i = 0; i < 8; i++) out[i] = g_stats[i]; }

static void lk_chain_stats(const int *diffs_all, const int16_t *dIbuf, int win)
{
    /* diffs_all: win x win residuals; dIbuf: win x win x 2 gradients */
    const long long LIM = 1LL << 24;
    int ok_old = 1, ok_strip = 1, ok_true = 1;
    for (int comp = 0; comp < 2; comp++) {
        /* SIMD chains q = 0..3: per row the pairs (q, q+4) then (q+8, q+12) */
        for (int q = 0; q < 4; q++) {
            long long sabs = 0, pre = 0, maxpre = 0, maxadd = 0;
            for (int y = 0; y < win; y++)
                for (int g = 0; g < 2; g++) {
                    int x0 = q + 8 * g, x1 = x0 + 4;
                    long long v0 = (long long)diffs_all[y * win + x0] * dIbuf[(y * win + x0) * 2 + comp];
                    long long v1 = (long long)diffs_all[y * win + x1] * dIbuf[(y * win + x1) * 2 + comp];
                    long long pair = v0 + v1;
                    sabs += llabs(v0) + llabs(v1);
                    pre += pair;
                    if (llabs(pre) > maxpre) maxpre = llabs(pre);
                    if (llabs(pair) > maxadd) maxadd = llabs(pair);
                }
            if (sabs > LIM) ok_old = 0;
            if (maxpre > LIM || maxadd > LIM) ok_true = 0;
            /* strip bound: strips (col, rows 0..10) and (col, rows 11..20) of the chain's 4 columns */
            long long bound = 0, maxa = 0;
            for (int c = q; c < 16; c += 4)
                for (int half = 0; half < 2; half++) {
                    long long run = 0, mx = 0;
                    for (int y = half ? 11 : 0; y < (half ? win : 11); y++) {
                        long long v = (long long)diffs_all[y * win + c] * dIbuf[(y * win + c) * 2 + comp];
                        run += v;
                        if (llabs(run) > mx) mx = llabs(run);
                        if (llabs(v) > maxa) maxa = llabs(v);
                    }
                    bound += mx;
                }
            if (bound > LIM || 2 * maxa > LIM) ok_strip = 0;
        }
        /* tail chain: columns 16..20 row-major */
        {
            long long sabs = 0, pre = 0, maxpre = 0, maxadd = 0;
            for (int y = 0; y < win; y++)
                for (int x = 16; x < win; x++) {
                    long long v = (long long)diffs_all[y * win + x] * dIbuf[(y * win + x) * 2 + comp];
                    sabs += llabs(v); pre += v;
                    if (llabs(pre) > maxpre) maxpre = llabs(pre);
                    if (llabs(v) > maxadd) maxadd = llabs(v);
                }
            if (sabs > LIM) ok_old = 0;
            if (maxpre > LIM || maxadd > LIM) ok_true = 0;
            static const int seg_r0[6] = {0, 4, 8, 12, 15, 18}, seg_n[6] = {4, 4, 4, 3, 3, 3};
            long long bound = 0;
            for (int x = 16; x < win; x++)
                for (int sg = 0; sg < 6; sg++) {
                    long long run = 0, mx = 0;
                    for (int y = seg_r0[sg]; y < seg_r0[sg] + seg_n[sg]; y++) {
                        run += (long long)diffs_all[y * win + x] * dIbuf[(y * win + x) * 2 + comp];
                        if (llabs(run) > mx) mx = llabs(run);
                    }
                    bound += mx;
                }
            if (bound > LIM || maxadd > LIM) ok_strip = 0;
        }
    }
    g_stats[0]++; g_stats[1] += ok_old; g_stats[2] += ok_strip; g_stats[3] += ok_true;
}

static inline float reduce4(const float q[4])
{
    switch (g_sum_mode) {
    default:
    case 0: return (q[0] + q[2]) + (q[1] + q[3]);   /* SSE v_reduce_sum: movehl + shuffle */
    case 1: return ((q[0] + q[1]) + q[2]) + q[3];
    case 2: return (q[0] + q[1]) + (q[2] + q[3]);
    }
}

#define W_BITS 14
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

static inline int cv_floor(float v) { return (int)floorf(v); }
static inline int cv_round(float v) { return (int)lrintf(v); } /* round-half-even */

/*
 * One calcOpticalFlowPyrLK call, flags = 0, err requested.
 * next_pts is output only (no OPTFLOW_USE_INITIAL_FLOW).
 * iters_out (optional): total J-iterations per point, for workload statistics.
 */
void lk_track(const lk_pyr_t *prev, const lk_pyr_t *next,
              const float *prev_pts, float *next_pts, uint8_t *status, float *err,
              int npts, int win, int max_count, double epsilon, double min_eig_thr,
              int *iters_out)
{
    int nlev = prev->nlevels < next->nlevels ? prev->nlevels : next->nlevels;
    int max_level = nlev - 1;
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    if (epsilon < 0.) epsilon = 0.; if (epsilon > 10.) epsilon = 10.;
    double eps2 = epsilon * epsilon;
    int16_t *Ibuf = (int16_t *)malloc(sizeof(int16_t) * win * win * 3);
    int16_t *dIbuf = Ibuf + win * win;
    const int nsimd = (win / 8) * 8;   /* columns handled by the 8-wide SIMD body */

    for (int i = 0; i < npts; i++) { status[i] = 1; if (err) err[i] = 0.f; if (iters_out) iters_out[i] = 0; }

    for (int level = max_level; level >= 0; level--) {
        const lk_level_t *I = &prev->lv[level], *J = &next->lv[level];
        const uint8_t *Ibase = I->img + I->pad * I->istep + I->pad;
        const uint8_t *Jbase = J->img + J->pad * J->istep + J->pad;
        const int16_t *Dbase = I->deriv + I->pad * I->dstep + 2 * I->pad;
        for (int pi = 0; pi < npts; pi++) {
            float px = prev_pts[2 * pi] * (float)(1. / (1 << level));
            float py = prev_pts[2 * pi + 1] * (float)(1. / (1 << level));
            float nx, ny;
            if (level == max_level) { nx = px; ny = py; }
            else { nx = next_pts[2 * pi] * 2.f; ny = next_pts[2 * pi + 1] * 2.f; }
            next_pts[2 * pi] = nx; next_pts[2 * pi + 1] = ny;

            px -= half; py -= half;
            int ipx = cv_floor(px), ipy = cv_floor(py);
            if (ipx < -win || ipx >= I->w || ipy < -win || ipy >= I->h) {
                if (level == 0) { status[pi] = 0; if (err) err[pi] = 0.f; }
                continue;
            }
            float a = px - ipx, b = py - ipy;
            int iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
            int iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
            int iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
            int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;

            float qA11[4] = {0, 0, 0, 0}, qA12[4] = {0, 0, 0, 0}, qA22[4] = {0, 0, 0, 0};
            float iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win; y++) {
                const uint8_t *src = Ibase + (y + ipy) * I->istep + ipx;
                const int16_t *dsrc = Dbase + (y + ipy) * I->dstep + 2 * ipx;
                int16_t *Ip = Ibuf + y * win, *dIp = dIbuf + y * win * 2;
                for (int x = 0; x < win; x++, dsrc += 2, dIp += 2) {
                    int ival = DESCALE(src[x] * iw00 + src[x + 1] * iw01 +
                                       src[x + I->istep] * iw10 + src[x + I->istep + 1] * iw11, W_BITS - 5);
                    int ixval = DESCALE(dsrc[0] * iw00 + dsrc[2] * iw01 +
                                        dsrc[I->dstep] * iw10 + dsrc[I->dstep + 2] * iw11, W_BITS);
                    int iyval = DESCALE(dsrc[1] * iw00 + dsrc[3] * iw01 +
                                        dsrc[I->dstep + 1] * iw10 + dsrc[I->dstep + 3] * iw11, W_BITS);
                    Ip[x] = (int16_t)ival; dIp[0] = (int16_t)ixval; dIp[1] = (int16_t)iyval;
                    float fx = (float)ixval, fy = (float)iyval;
                    if (x < nsimd && g_sum_mode != 9) {
                        int l = x & 3;
                        /* v_muladd without FMA: separately rounded mul and add */
                        qA22[l] = fy * fy + qA22[l];
                        qA12[l] = fx * fy + qA12[l];
                        qA11[l] = fx * fx + qA11[l];
                    } else {
                        iA11 += (float)(ixval * ixval);
                        iA12 += (float)(ixval * iyval);
                        iA22 += (float)(iyval * iyval);
                    }
                }
            }
            if (g_sum_mode != 9) {
                iA11 += reduce4(qA11); iA12 += reduce4(qA12); iA22 += reduce4(qA22);
            }
            float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < min_eig_thr || D < FLT_EPSILON) {
                if (level == 0) status[pi] = 0;
                continue;
            }
            D = 1.f / D;

            float npx = nx - half, npy = ny - half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_count; j++) {
                int inx = cv_floor(npx), iny = cv_floor(npy);
                if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) {
                    if (level == 0) status[pi] = 0;
                    break;
                }
                if (iters_out) iters_out[pi]++;
                a = npx - inx; b = npy - iny;
                iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
                iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
                iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                float qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0};
                float ib1 = 0, ib2 = 0;
                int diffs_all[32 * 32];
                for (int y = 0; y < win; y++) {
                    const uint8_t *Jp = Jbase + (y + iny) * J->istep + inx;
                    const int16_t *Ip = Ibuf + y * win, *dIp = dIbuf + y * win * 2;
                    int *diffs = diffs_all + y * win;
                    for (int x = 0; x < win; x++)
                        diffs[x] = DESCALE(Jp[x] * iw00 + Jp[x + 1] * iw01 +
                                           Jp[x + J->istep] * iw10 + Jp[x + J->istep + 1] * iw11, W_BITS - 5) - Ip[x];
                    int x = 0;
                    if (g_sum_mode != 9) {
                        for (; x <= win - 8; x += 8) {
                            /* lanes: qb0 = [bx(0,4) by(0,4) bx(1,5) by(1,5)], qb1 = [bx(2,6) by(2,6) bx(3,7) by(3,7)] */
                            for (int l = 0; l < 4; l++) {
                                int p0 = x + l, p1 = x + l + 4;
                                int sx = dIp[2 * p0] * diffs[p0] + dIp[2 * p1] * diffs[p1];
                                int sy = dIp[2 * p0 + 1] * diffs[p0] + dIp[2 * p1 + 1] * diffs[p1];
                                float *q = (l < 2) ? qb0 : qb1;
                                int li = (l & 1) * 2;
                                q[li] += (float)sx;
                                q[li + 1] += (float)sy;
                            }
                        }
                    }
                    for (; x < win; x++) {
                        ib1 += (float)(diffs[x] * dIp[2 * x]);
                        ib2 += (float)(diffs[x] * dIp[2 * x + 1]);
                    }
                }
                if (g_stats_on && win == 21) lk_chain_stats(diffs_all, dIbuf, win);
                if (g_sum_mode != 9) {
                    /* qf0 = interleave_pairs(qb0+qb1) low half = [s0, s2, 0, 0], qf1 = [s1, s3, 0, 0] */
                    float s[4];
                    for (int l = 0; l < 4; l++) s[l] = qb0[l] + qb1[l];
                    float qf0[4] = {s[0], s[2], 0.f, 0.f}, qf1[4] = {s[1], s[3], 0.f, 0.f};
                    ib1 += reduce4(qf0);
                    ib2 += reduce4(qf1);
                }
                float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
                float dx = (float)((A12 * b2 - A22 * b1) * D);
                float dy = (float)((A12 * b1 - A11 * b2) * D);
                npx += dx; npy += dy;
                next_pts[2 * pi] = npx + half; next_pts[2 * pi + 1] = npy + half;
                if ((double)dx * dx + (double)dy * dy <= eps2) break;
                if (j > 0 && fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01) {
                    next_pts[2 * pi] -= dx * 0.5f;
                    next_pts[2 * pi + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }

            if (status[pi] && err && level == 0) {
                float fx = next_pts[2 * pi] - half, fy = next_pts[2 * pi + 1] - half;
                int inx = cv_floor(fx), iny = cv_floor(fy);
                if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) {
                    status[pi] = 0;
                    continue;
                }
                float aa = fx - inx, bb = fy - iny;
                iw00 = cv_round((1.f - aa) * (1.f - bb) * (1 << W_BITS));
                iw01 = cv_round(aa * (1.f - bb) * (1 << W_BITS));
                iw10 = cv_round((1.f - aa) * bb * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                float errval = 0.f;
                for (int y = 0; y < win; y++) {
                    const uint8_t *Jp = Jbase + (y + iny) * J->istep + inx;
                    const int16_t *Ip = Ibuf + y * win;
                    for (int x = 0; x < win; x++) {
                        int diff = DESCALE(Jp[x] * iw00 + Jp[x + 1] * iw01 +
                                           Jp[x + J->istep] * iw10 + Jp[x + J->istep + 1] * iw11, W_BITS - 5) - Ip[x];
                        errval += fabsf((float)diff);
                    }
                }
                err[pi] = errval * 1.f / (32 * win * win);
            }
        }
    }
    free(Ibuf);
}
