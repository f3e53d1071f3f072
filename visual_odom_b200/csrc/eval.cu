// eval.cu -- KITTI odometry accuracy metric, host-only.
// The benchmark's published definition, as the reference's bundled devkit evaluates it
// (reference src/evaluate/evaluate_odometry.cpp:17-116,376-395): for every 10th frame as a start and every segment
// length in {100..800 m} measured along the GROUND-TRUTH path, compare the relative motion of the estimate with the
// ground truth's; rotation error [rad/m] and translation error [fraction] per segment, averaged over all segments.
// Arithmetic is single precision where the devkit's is (path lengths, error values), double for the 4x4 algebra.
// Plotting (gnuplot), mail and directory handling of the devkit are out of scope.
#include "ctx.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

struct M4 { double v[16]; };

M4 from12(const double* p)
{
    M4 m;
    memcpy(m.v, p, 12 * sizeof(double));
    m.v[12] = m.v[13] = m.v[14] = 0.0; m.v[15] = 1.0;
    return m;
}

M4 mul(const M4& a, const M4& b)
{
    M4 c;
    for (int r = 0; r < 4; r++)
        for (int k = 0; k < 4; k++) {
            double s = 0;
            for (int j = 0; j < 4; j++) s += a.v[4 * r + j] * b.v[4 * j + k];
            c.v[4 * r + k] = s;
        }
    return c;
}

bool inv(const M4& m, M4& out)
{
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 8; c++) a[r][c] = c < 4 ? m.v[4 * r + c] : (c - 4 == r ? 1.0 : 0.0);
    for (int col = 0; col < 4; col++) {
        int piv = col;
        for (int r = col + 1; r < 4; r++)
            if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.0) return false;
        if (piv != col)
            for (int c = 0; c < 8; c++) { const double x = a[piv][c]; a[piv][c] = a[col][c]; a[col][c] = x; }
        const double d = a[col][col];
        for (int c = 0; c < 8; c++) a[col][c] /= d;
        for (int r = 0; r < 4; r++) {
            if (r == col) continue;
            const double f = a[r][col];
            if (f != 0.0)
                for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out.v[4 * r + c] = a[r][4 + c];
    return true;
}

}  // namespace

// KITTI pose file: one pose per line, 12 doubles = the top 3 rows of the 4x4 camera-to-world matrix.
extern "C" int vo_poses_load(const char* path, double* poses12, int cap, int* n_out)
{
    if (!path || !n_out) return VO_E_INVALID;
    FILE* f = fopen(path, "r");
    if (!f) return VO_E_INVALID;
    int n = 0;
    double p[12];
    for (;;) {
        int got = 0;
        for (; got < 12; got++)
            if (fscanf(f, "%lf", &p[got]) != 1) break;
        if (got < 12) break;
        if (poses12 && n < cap) memcpy(poses12 + (size_t)12 * n, p, sizeof(p));
        n++;
    }
    fclose(f);
    *n_out = n;
    return (poses12 && n > cap) ? VO_E_CAPACITY : VO_OK;
}

extern "C" int vo_poses_save(const char* path, const double* poses12, int n)
{
    if (!path || !poses12 || n < 0) return VO_E_INVALID;
    FILE* f = fopen(path, "w");
    if (!f) return VO_E_INVALID;
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 12; k++) fprintf(f, k ? " %.9e" : "%.9e", poses12[(size_t)12 * i + k]);
        fputc('\n', f);
    }
    fclose(f);
    return VO_OK;
}

extern "C" int vo_eval_segments(const double* gt12, const double* est12, int n_poses, const float* lengths, int n_lengths,
                                int step, vo_segment_error* out, int cap, int* n_out)
{
    if (!gt12 || !est12 || n_poses <= 0 || !n_out) return VO_E_INVALID;
    static const float kitti_lengths[8] = {100, 200, 300, 400, 500, 600, 700, 800};
    if (!lengths || n_lengths <= 0) { lengths = kitti_lengths; n_lengths = 8; }
    if (step <= 0) step = 10;
    // path length along the ground truth, accumulated in float
    std::vector<float> dist(n_poses);
    dist[0] = 0.f;
    for (int i = 1; i < n_poses; i++) {
        const float dx = (float)(gt12[12 * (i - 1) + 3] - gt12[12 * i + 3]);
        const float dy = (float)(gt12[12 * (i - 1) + 7] - gt12[12 * i + 7]);
        const float dz = (float)(gt12[12 * (i - 1) + 11] - gt12[12 * i + 11]);
        dist[i] = dist[i - 1] + std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    int n = 0;
    for (int first = 0; first < n_poses; first += step) {
        for (int li = 0; li < n_lengths; li++) {
            const float len = lengths[li];
            int last = -1;
            for (int i = first; i < n_poses; i++)
                if (dist[i] > dist[first] + len) { last = i; break; }
            if (last < 0) continue;
            M4 gi, ei, di;
            if (!inv(from12(gt12 + 12 * first), gi) || !inv(from12(est12 + 12 * first), ei)) return VO_E_INVALID;
            const M4 d_gt = mul(gi, from12(gt12 + 12 * last));
            const M4 d_est = mul(ei, from12(est12 + 12 * last));
            if (!inv(d_est, di)) return VO_E_INVALID;
            const M4 e = mul(di, d_gt);
            const float a = (float)e.v[0], b = (float)e.v[5], c = (float)e.v[10];
            float d = (float)(0.5 * ((double)(a + b + c) - 1.0));
            d = d > 1.f ? 1.f : (d < -1.f ? -1.f : d);
            const float r_err = std::acos(d);
            const float tx = (float)e.v[3], ty = (float)e.v[7], tz = (float)e.v[11];
            const float t_err = std::sqrt(tx * tx + ty * ty + tz * tz);
            const float frames = (float)(last - first + 1);
            if (out && n < cap) {
                out[n].first_frame = first;
                out[n].r_err = r_err / len;
                out[n].t_err = t_err / len;
                out[n].len = len;
                out[n].speed = (float)(len / (0.1 * frames));
            }
            n++;
        }
    }
    *n_out = n;
    return (out && n > cap) ? VO_E_CAPACITY : VO_OK;
}

extern "C" int vo_eval_summary(const vo_segment_error* seg, int n, float* t_err_avg, float* r_err_avg)
{
    if (!seg || n <= 0) return VO_E_INVALID;
    float t = 0.f, r = 0.f;
    for (int i = 0; i < n; i++) { t += seg[i].t_err; r += seg[i].r_err; }
    if (t_err_avg) *t_err_avg = t / (float)n;
    if (r_err_avg) *r_err_avg = r / (float)n;
    return VO_OK;
}
