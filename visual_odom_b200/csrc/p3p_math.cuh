// p3p_math.cuh -- the n == 4 case of cv::solvePnPRansac (fp64, host/device).
//
// reference src/visualOdometry.cpp:176-178 calls cv::solvePnPRansac; with exactly four correspondences OpenCV skips RANSAC:
// model_points = 4, kernel = SOLVEPNP_P3P, one cv::solvePnP on all four points (pose of the first three, the fourth picks
// among the up-to-four solutions by its squared pixel reprojection error), no refinement, inliers = {0, 1, 2, 3}.
// The pose set of three correspondences is a mathematical function of the input, so this is Grunert's formulation and not
// an operation-for-operation copy of OpenCV's solver: distances s1, s2 = u s1, s3 = v s1 along the unit bearings,
//   u = N(v) / D(v),  N = (K - 1) v^2 - 2 K cos(beta) v + (K + 1),  D = 2 (cos(gamma) - v cos(alpha)),  K = (a^2 - c^2) / b^2
//   quartic in v:  D^2 + N^2 - 2 cos(gamma) N D - (c^2 / b^2) (1 + v^2 - 2 v cos(beta)) D^2 = 0
// built by polynomial products, all roots by poly_roots (polished, near-double roots re-derived, see below), then the rigid
// motion from the two triangles' orthonormal frames.
// What IS reproduced from OpenCV: the image points are normalised as cv::undistortPoints does (f64 arithmetic, stored f32),
// and K is the float matrix the reference builds.  Agreement with cv2 (tests/test_oracle_pnp.py, this header compiled for
// the host): a pose whenever cv2 has one, the same solution picked, [R|t] within 4e-7 on generic scenes, 4e-5 on
// ill-conditioned 0.3 m point clusters (1e-4 asked).  Restated in oracle/pnp_ref.py (p3p_four_points).
// Where cv2's P3P returns NaN poses (~0.3 % of random sets) this returns "no model" instead.
#pragma once
#include "ess_math.cuh"

namespace vomath {

VO_HD void p3p_cross(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
// columns e1, e2, e3 of the orthonormal frame of triangle A (rows = points), row-major 3 x 3 (F[r][c] = e_c[r])
VO_HD bool p3p_frame(const double A[3][3], double* F)
{
    double e1[3], w[3], e3[3], e2[3];
    for (int k = 0; k < 3; k++) { e1[k] = A[1][k] - A[0][k]; w[k] = A[2][k] - A[0][k]; }
    const double n1 = sqrt(dot3(e1, e1));
    if (!(n1 > 0)) return false;
    for (int k = 0; k < 3; k++) e1[k] /= n1;
    p3p_cross(e1, w, e3);
    const double n3 = sqrt(dot3(e3, e3));
    if (!(n3 > 0)) return false;
    for (int k = 0; k < 3; k++) e3[k] /= n3;
    p3p_cross(e3, e1, e2);
    for (int r = 0; r < 3; r++) { F[r * 3 + 0] = e1[r]; F[r * 3 + 1] = e2[r]; F[r * 3 + 2] = e3[r]; }
    return true;
}

// X: three world points; y: their normalised image coordinates.  Up to four (R row-major, t).  Returns the count.
VO_HDN inline int p3p_solutions(const double X[3][3], const double y[3][2], double* R_out, double* t_out)
{
    double f[3][3];
    for (int i = 0; i < 3; i++) {
        const double n = sqrt(y[i][0] * y[i][0] + y[i][1] * y[i][1] + 1.0);
        f[i][0] = y[i][0] / n; f[i][1] = y[i][1] / n; f[i][2] = 1.0 / n;
    }
    double d12[3], d02[3], d01[3];
    for (int k = 0; k < 3; k++) { d12[k] = X[1][k] - X[2][k]; d02[k] = X[0][k] - X[2][k]; d01[k] = X[0][k] - X[1][k]; }
    const double a2 = dot3(d12, d12), b2 = dot3(d02, d02), c2 = dot3(d01, d01);
    if (!(b2 > 0)) return 0;
    const double ca = dot3(f[1], f[2]), cb = dot3(f[0], f[2]), cg = dot3(f[0], f[1]);
    const double K = (a2 - c2) / b2;
    const double N[3] = {K + 1, -2 * K * cb, K - 1};
    const double D[2] = {2 * cg, -2 * ca};
    const double Q[3] = {1, -2 * cb, 1};
    double D2[3] = {0, 0, 0}, QD2[5] = {0, 0, 0, 0, 0}, poly[5] = {0, 0, 0, 0, 0};
    poly_mac(D, 1, D, 1, 1.0, D2);
    poly_mac(Q, 2, D2, 2, 1.0, QD2);
    for (int k = 0; k < 3; k++) poly[k] += D2[k];
    poly_mac(N, 2, N, 2, 1.0, poly);
    poly_mac(N, 2, D, 1, -2 * cg, poly);
    for (int k = 0; k < 5; k++) poly[k] -= (c2 / b2) * QD2[k];
    cplx roots[10];
    const int nr = poly_roots(poly, 4, roots);
    double Fw[9];
    if (!p3p_frame(X, Fw)) return 0;
    // ---- real roots of the quartic ------------------------------------------------------------------------------------
    // Two real roots that nearly coincide (poses next to Grunert's singularity, clustered points) come out of any root finder
    // with errors ~ sqrt(eps), often as a conjugate pair with a small imaginary part: such a pair is re-derived from the local
    // parabola around the extremum of the quartic, p(v) ~ A ((v - vm)^2 + s), real iff s <= 0.  Every root is polished by
    // Newton steps on the real quartic (this alone takes the agreement with cv2 from 6e-5 to 1e-6 on generic scenes).
    auto pval = [&](double v) { return (((poly[4] * v + poly[3]) * v + poly[2]) * v + poly[1]) * v + poly[0]; };
    auto pd1 = [&](double v) { return ((4 * poly[4] * v + 3 * poly[3]) * v + 2 * poly[2]) * v + poly[1]; };
    auto pd2 = [&](double v) { return (12 * poly[4] * v + 6 * poly[3]) * v + 2 * poly[2]; };
    auto pabs = [&](double v) { v = fabs(v); return (((fabs(poly[4]) * v + fabs(poly[3])) * v + fabs(poly[2])) * v + fabs(poly[1])) * v + fabs(poly[0]); };
    auto polish = [&](double v) {
        for (int it = 0; it < 4; it++) {
            const double dv = pd1(v);
            if (dv == 0.0) break;
            v -= pval(v) / dv;
        }
        return v;
    };
    double vs[8];
    int nv = 0;
    for (int i = 0; i < nr && nv < 7; i++) {
        const double re = roots[i].re, im = roots[i].im;
        const double mag = fabs(re) > 1.0 ? fabs(re) : 1.0;
        if (fabs(im) <= 1e-9 * mag) {
            vs[nv++] = polish(re);
        } else if (im > 0 && im <= 1e-3 * mag) {
            double vm = re;
            for (int it = 0; it < 4; it++) {                  // extremum: Newton on p'
                const double dd = pd2(vm);
                if (dd == 0.0) break;
                vm -= pd1(vm) / dd;
            }
            const double A = 0.5 * pd2(vm);
            if (A != 0.0) {
                const double sq = pval(vm) / A;
                if (sq <= 0) {
                    const double dl = sqrt(-sq);
                    if (dl > 1e-12 * mag) { vs[nv++] = polish(vm - dl); vs[nv++] = polish(vm + dl); }
                    else vs[nv++] = vm;
                }
            }
        }
    }
    // ---- one pose per (v, u) ------------------------------------------------------------------------------------------
    int count = 0;
    double uv_seen[8];
    for (int i = 0; i < nv && count < 4; i++) {
        const double v = vs[i];
        if (!(fabs(pval(v)) <= 1e-9 * pabs(v)) || !(v > 0)) continue;
        const double q = 1 + v * v - 2 * v * cb;
        if (!(q > 0)) continue;
        // u = N(v) / D(v); next to the singularity D(v) -> 0 (N(v) -> 0 too) the quotient only selects which root of the
        // quadratic u^2 - 2 cos(gamma) u + 1 - (c^2 / b^2) q = 0 (the third distance equation) belongs to this v
        double us[2];
        int nu = 0;
        const double Dv = D[0] + D[1] * v;
        const double rel = fabs(Dv) / (fabs(D[0]) + fabs(D[1] * v));
        const double u_lin = Dv != 0.0 ? (N[0] + N[1] * v + N[2] * v * v) / Dv : 0.0;
        if (rel > 1e-3) {
            us[nu++] = u_lin;
        } else {
            const double disc = cg * cg - 1 + (c2 / b2) * q;
            if (disc >= 0) {
                const double sq = sqrt(disc), u0 = cg + sq, u1 = cg - sq;
                if (rel > 1e-9) us[nu++] = fabs(u0 - u_lin) <= fabs(u1 - u_lin) ? u0 : u1;
                else { us[nu++] = u0; us[nu++] = u1; }
            }
        }
        for (int k = 0; k < nu && count < 4; k++) {
            const double u = us[k];
            if (!(u > 0)) continue;
            // the other distance equation must hold as well (it does by construction on the linear branch)
            if (fabs(u * u + v * v - 2 * u * v * ca - (a2 / b2) * q) > 1e-5 * (u * u + v * v + (a2 / b2) * q)) continue;
            bool dup = false;
            for (int j = 0; j < count; j++) dup = dup || (fabs(uv_seen[2 * j] - u) <= 1e-7 * u && fabs(uv_seen[2 * j + 1] - v) <= 1e-7 * v);
            if (dup) continue;
            const double s1 = sqrt(b2 / q);
            const double s[3] = {s1, u * s1, v * s1};
            double P[3][3], Fc[9];
            for (int j = 0; j < 3; j++)
                for (int kk = 0; kk < 3; kk++) P[j][kk] = f[j][kk] * s[j];
            if (!p3p_frame(P, Fc)) continue;
            double* R = R_out + 9 * count;
            double* t = t_out + 3 * count;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) R[r * 3 + c] = Fc[r * 3 + 0] * Fw[c * 3 + 0] + Fc[r * 3 + 1] * Fw[c * 3 + 1] + Fc[r * 3 + 2] * Fw[c * 3 + 2];
            for (int r = 0; r < 3; r++) t[r] = P[0][r] - (R[r * 3 + 0] * X[0][0] + R[r * 3 + 1] * X[0][1] + R[r * 3 + 2] * X[0][2]);
            bool finite = true;
            for (int kk = 0; kk < 9; kk++) finite &= (R[kk] == R[kk]) && fabs(R[kk]) <= 2.0;
            for (int kk = 0; kk < 3; kk++) finite &= (t[kk] == t[kk]) && fabs(t[kk]) < 1e300;
            if (finite) { uv_seen[2 * count] = u; uv_seen[2 * count + 1] = v; count++; }
        }
    }
    return count;
}

// The four-point solvePnP: Xw / uv are the four float correspondences.  Writes R (row-major), t; false = no solution.
VO_HDN inline bool p3p_four_points(const float* Xw_f, const float* uv_f, double fu, double fv, double uc, double vc,
                                   double* R, double* t)
{
    double X[3][3], y[3][2];
    const double ifu = 1.0 / fu, ifv = 1.0 / fv;
    for (int i = 0; i < 3; i++) {
        for (int k = 0; k < 3; k++) X[i][k] = (double)Xw_f[3 * i + k];
        y[i][0] = (double)(float)(((double)uv_f[2 * i] - uc) * ifu);           // cv::undistortPoints: f64 arithmetic, f32 store
        y[i][1] = (double)(float)(((double)uv_f[2 * i + 1] - vc) * ifv);
    }
    double Rs[36], ts[12];
    const int ns = p3p_solutions(X, y, Rs, ts);
    int best = -1;
    double best_err = 0;
    const double X3[3] = {(double)Xw_f[9], (double)Xw_f[10], (double)Xw_f[11]};
    for (int i = 0; i < ns; i++) {
        const double* Ri = Rs + 9 * i;
        const double xc = Ri[0] * X3[0] + Ri[1] * X3[1] + Ri[2] * X3[2] + ts[3 * i];
        const double yc = Ri[3] * X3[0] + Ri[4] * X3[1] + Ri[5] * X3[2] + ts[3 * i + 1];
        const double zc = Ri[6] * X3[0] + Ri[7] * X3[1] + Ri[8] * X3[2] + ts[3 * i + 2];
        const double du = uc + fu * xc / zc - (double)uv_f[6], dv = vc + fv * yc / zc - (double)uv_f[7];
        const double e = du * du + dv * dv;
        if (!(e == e)) continue;
        if (best < 0 || e < best_err) { best = i; best_err = e; }
    }
    if (best < 0) return false;
    for (int k = 0; k < 9; k++) R[k] = Rs[9 * best + k];
    for (int k = 0; k < 3; k++) t[k] = ts[3 * best + k];
    return true;
}

}   // namespace vomath
