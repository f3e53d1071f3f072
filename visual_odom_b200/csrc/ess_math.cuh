// ess_math.cuh -- five-point essential-matrix solver, Sampson error, essential decomposition (fp64, host/device).
//
// The arithmetic behind the reference's `mono_rotation = true` branch (reference src/visualOdometry.cpp:146-157):
//   cv::findEssentialMat(p_t0, p_t1, focal, pp, RANSAC, 0.999, 1.0, mask)  +  cv::recoverPose(E, ..., mask)
// i.e. Nister's five-point algorithm as OpenCV's EMEstimatorCallback runs it: null space of the 5 x 9 epipolar system,
// the ten cubic constraints det(E) = 0, 2 E E^T E - tr(E E^T) E = 0 in Nister's monomial order, Gauss-Jordan, the 3 x 3
// polynomial matrix B(z) = <e> - z <f>, its degree-10 determinant, one E per real root.  The candidate set of a sample is
// a mathematical function of the five correspondences, so this is NOT an operation-for-operation copy (OpenCV's generated
// coefficient code and its root finder are not reproduced); oracle/essential_ref.py restates the same algorithm in numpy
// and is pinned against cv2 4.13.0 (identical inlier masks, rotations to 1e-12 on the stress sets), and
// tests/test_oracle_essential.py checks THIS code, compiled for the host, against both.
#pragma once
#include "pnp_math.cuh"

namespace vomath {

// monomials: degree <= 1: [x, y, z, 1]; degree <= 2: [x2, y2, z2, xy, xz, yz, x, y, z, 1];
// degree <= 3 in Nister's order: [x3, y3, x2y, xy2, x2z, x2, y2z, y2, xyz, xy | xz2, xz, x, yz2, yz, y, z3, z2, z, 1]
VO_HD int ess_t12(int a, int b)
{
    const int8_t T[4][4] = {{0, 3, 4, 6}, {3, 1, 5, 7}, {4, 5, 2, 8}, {6, 7, 8, 9}};
    return T[a][b];
}
VO_HD int ess_t23(int c, int a)
{
    const int8_t T[10][4] = {{0, 2, 4, 5}, {3, 1, 6, 7}, {10, 13, 16, 17}, {2, 3, 8, 9}, {4, 8, 10, 11},
                             {8, 6, 13, 14}, {5, 9, 11, 12}, {9, 7, 14, 15}, {11, 14, 17, 18}, {12, 15, 18, 19}};
    return T[c][a];
}
// o (degree 2, 10 coefficients) += s * a * b   (a, b degree 1)
VO_HD void ess_mul11(const double* a, const double* b, double s, double* o)
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) o[ess_t12(i, j)] += s * a[i] * b[j];
}
// o (degree 3, 20 coefficients) += s * c * a   (c degree 2, a degree 1)
VO_HD void ess_mul21(const double* c, const double* a, double s, double* o)
{
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 4; j++) o[ess_t23(i, j)] += s * c[i] * a[j];
}

struct cplx { double re, im; };
VO_HD cplx c_mul(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
VO_HD cplx c_sub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
VO_HD cplx c_div(cplx a, cplx b)
{
    const double d = b.re * b.re + b.im * b.im;
    return cplx{(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// all roots of c[0] + c[1] z + ... + c[n] z^n (n <= 10, c[n] != 0): Durand-Kerner from a circle of the Cauchy radius,
// then each root is polished by two Newton steps of the full polynomial in complex arithmetic
VO_HDN inline int poly_roots(const double* c, int n, cplx* r)
{
    while (n > 0 && c[n] == 0.0) n--;
    if (n <= 0) return 0;
    double a[11];
    double radius = 0;
    for (int k = 0; k <= n; k++) a[k] = c[k] / c[n];
    for (int k = 0; k < n; k++) { const double t = fabs(a[k]); if (t > radius) radius = t; }
    radius = 1.0 + radius;
    // a tighter start: geometric mean of the root moduli is |a0|^(1/n)
    double r0 = pow(fabs(a[0]) > 0 ? fabs(a[0]) : 1.0, 1.0 / n);
    if (!(r0 > 1e-3)) r0 = 1e-3;
    if (r0 > radius) r0 = radius;
    for (int k = 0; k < n; k++) {
        const double ang = 6.283185307179586 * k / n + 0.4;
        r[k] = cplx{r0 * cos(ang), r0 * sin(ang)};
    }
    for (int iter = 0; iter < 600; iter++) {
        double maxd = 0, maxr = 0;
        for (int i = 0; i < n; i++) {
            cplx p = r[i], num = cplx{1.0, 0.0};                     // Horner, monic
            for (int k = n - 1; k >= 0; k--) { num = c_mul(num, p); num.re += a[k]; }
            cplx den = cplx{1.0, 0.0};
            for (int j = 0; j < n; j++)
                if (j != i) den = c_mul(den, c_sub(p, r[j]));
            if (den.re == 0.0 && den.im == 0.0) { den.re = 1e-300; }
            const cplx d = c_div(num, den);
            r[i] = c_sub(p, d);
            const double ad = fabs(d.re) + fabs(d.im), ar = fabs(r[i].re) + fabs(r[i].im);
            if (ad > maxd) maxd = ad;
            if (ar > maxr) maxr = ar;
        }
        if (maxd <= 1e-15 * (maxr > 1.0 ? maxr : 1.0)) break;
    }
    for (int i = 0; i < n; i++)
        for (int it = 0; it < 2; it++) {
            cplx p = r[i], f = cplx{1.0, 0.0}, df = cplx{0.0, 0.0};
            for (int k = n - 1; k >= 0; k--) {
                df = c_mul(df, p); df.re += f.re; df.im += f.im;
                f = c_mul(f, p); f.re += a[k];
            }
            if (df.re == 0.0 && df.im == 0.0) break;
            const cplx d = c_div(f, df);
            if (!(fabs(d.re) + fabs(d.im) < 1e-6 * (1.0 + fabs(p.re) + fabs(p.im)))) break;     // polish only
            r[i] = c_sub(p, d);
        }
    return n;
}

// polynomial product (coefficients in ascending powers): o[0 .. na + nb] += s * a * b
VO_HD void poly_mac(const double* a, int na, const double* b, int nb, double s, double* o)
{
    for (int i = 0; i <= na; i++)
        for (int j = 0; j <= nb; j++) o[i + j] += s * a[i] * b[j];
}

// q1, q2: five normalised correspondences (x, y interleaved), x2^T E x1 = 0.  E_out: up to 10 matrices (row-major 3 x 3).
VO_HDN inline int five_point(const double* q1, const double* q2, double* E_out)
{
    // ---- null space of the epipolar system (rows of Vt with zero singular value) ----
    double At[9 * 5], W[9], Vt[81];
    for (int i = 0; i < 5; i++) {
        const double x1 = q1[2 * i], y1 = q1[2 * i + 1], x2 = q2[2 * i], y2 = q2[2 * i + 1];
        const double row[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
        for (int k = 0; k < 9; k++) At[k * 5 + i] = row[k];
    }
    jacobi_svd_t<5, 9>(At, W, Vt, 0);
    const double* EE = Vt + 5 * 9;               // 4 x 9: E(x, y, z) = x EE0 + y EE1 + z EE2 + EE3
    // entry (r, c) of E as a degree-1 polynomial [x, y, z, 1]
    double e1[9][4];
    for (int k = 0; k < 9; k++)
        for (int b = 0; b < 4; b++) e1[k][b] = EE[b * 9 + k];
    // ---- the ten cubic constraints ----
    double A[10][20];
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 20; j++) A[i][j] = 0;
    {
        const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}, {1, 0, 2}, {0, 2, 1}};
        for (int p = 0; p < 6; p++) {
            double t2[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            ess_mul11(e1[0 * 3 + perm[p][0]], e1[1 * 3 + perm[p][1]], 1.0, t2);
            ess_mul21(t2, e1[2 * 3 + perm[p][2]], p < 3 ? 1.0 : -1.0, A[0]);
        }
    }
    double eet[9][10];                           // E E^T, degree 2
    double tr[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double* o = eet[r * 3 + c];
            for (int k = 0; k < 10; k++) o[k] = 0;
            for (int k = 0; k < 3; k++) ess_mul11(e1[r * 3 + k], e1[c * 3 + k], 1.0, o);
            if (r == c) for (int k = 0; k < 10; k++) tr[k] += o[k];
        }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double* o = A[1 + r * 3 + c];
            for (int k = 0; k < 3; k++) ess_mul21(eet[r * 3 + k], e1[k * 3 + c], 2.0, o);
            ess_mul21(tr, e1[r * 3 + c], -1.0, o);
        }
    // ---- Gauss-Jordan with partial pivoting: [A_left | A_right] -> [I | A_left^-1 A_right] ----
    for (int col = 0; col < 10; col++) {
        int piv = col;
        double best = fabs(A[col][col]);
        for (int r = col + 1; r < 10; r++)
            if (fabs(A[r][col]) > best) { best = fabs(A[r][col]); piv = r; }
        if (!(best > 0)) return 0;
        if (piv != col)
            for (int k = 0; k < 20; k++) { const double t = A[col][k]; A[col][k] = A[piv][k]; A[piv][k] = t; }
        const double inv = 1.0 / A[col][col];
        for (int k = 0; k < 20; k++) A[col][k] *= inv;
        for (int r = 0; r < 10; r++)
            if (r != col) {
                const double f = A[r][col];
                if (f != 0)
                    for (int k = 0; k < 20; k++) A[r][k] -= f * A[col][k];
            }
    }
    // ---- B(z) = <e> - z <f> for the row pairs (4,5), (6,7), (8,9); ascending powers of z ----
    double Bx[3][4], By[3][4], B1[3][5];
    for (int i = 0; i < 3; i++) {
        const double* r1 = &A[2 * i + 4][10];
        const double* r2 = &A[2 * i + 5][10];
        // r = [xz2, xz, x, yz2, yz, y, z3, z2, z, 1]
        Bx[i][0] = r1[2];          Bx[i][1] = r1[1] - r2[2]; Bx[i][2] = r1[0] - r2[1]; Bx[i][3] = -r2[0];
        By[i][0] = r1[5];          By[i][1] = r1[4] - r2[5]; By[i][2] = r1[3] - r2[4]; By[i][3] = -r2[3];
        B1[i][0] = r1[9];          B1[i][1] = r1[8] - r2[9]; B1[i][2] = r1[7] - r2[8]; B1[i][3] = r1[6] - r2[7]; B1[i][4] = -r2[6];
    }
    double c[11];
    for (int k = 0; k < 11; k++) c[k] = 0;
    {
        const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}, {1, 0, 2}, {0, 2, 1}};
        for (int p = 0; p < 6; p++) {
            // columns: 0 = x (degree 3), 1 = y (degree 3), 2 = 1 (degree 4); row i takes column perm[p][i]
            const double* f[3]; int d[3];
            for (int i = 0; i < 3; i++) {
                const int col = perm[p][i];
                f[i] = col == 0 ? Bx[i] : (col == 1 ? By[i] : B1[i]);
                d[i] = col == 2 ? 4 : 3;
            }
            double t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            poly_mac(f[0], d[0], f[1], d[1], 1.0, t);
            poly_mac(t, d[0] + d[1], f[2], d[2], p < 3 ? 1.0 : -1.0, c);
        }
    }
    cplx roots[10];
    const int nr = poly_roots(c, 10, roots);
    int count = 0;
    for (int i = 0; i < nr && count < 10; i++) {
        if (fabs(roots[i].im) > 1e-10) continue;
        const double z = roots[i].re, z2 = z * z, z3 = z2 * z, z4 = z3 * z;
        double Bt[9], w3[3], vt3[9];          // Bt = Bz^T (rows = columns of Bz) for jacobi_svd_t<3,3>
        for (int j = 0; j < 3; j++) {
            Bt[0 * 3 + j] = Bx[j][0] + Bx[j][1] * z + Bx[j][2] * z2 + Bx[j][3] * z3;
            Bt[1 * 3 + j] = By[j][0] + By[j][1] * z + By[j][2] * z2 + By[j][3] * z3;
            Bt[2 * 3 + j] = B1[j][0] + B1[j][1] * z + B1[j][2] * z2 + B1[j][3] * z3 + B1[j][4] * z4;
        }
        jacobi_svd_t<3, 3>(Bt, w3, vt3, 0);
        const double* v = vt3 + 6;            // right singular vector of the smallest singular value
        if (fabs(v[2]) < 1e-10) continue;
        const double x = v[0] / v[2], y = v[1] / v[2];
        double* E = E_out + 9 * count;
        for (int k = 0; k < 9; k++) E[k] = x * EE[k] + y * EE[9 + k] + z * EE[18 + k] + EE[27 + k];
        count++;
    }
    return count;
}

// EMEstimatorCallback::computeError: squared Sampson distance of (x1, y1) <-> (x2, y2), stored as float
VO_HD float sampson_err(const double* E, double x1, double y1, double x2, double y2)
{
    const double Ex0 = E[0] * x1 + E[1] * y1 + E[2], Ex1 = E[3] * x1 + E[4] * y1 + E[5], Ex2 = E[6] * x1 + E[7] * y1 + E[8];
    const double Et0 = E[0] * x2 + E[3] * y2 + E[6], Et1 = E[1] * x2 + E[4] * y2 + E[7];
    const double x2tEx1 = x2 * Ex0 + y2 * Ex1 + Ex2;
    return (float)(x2tEx1 * x2tEx1 / (Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1));
}

VO_HD double det3(const double* M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
VO_HD void mat3_mul(const double* A, const double* B, double* C)
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

// cv::decomposeEssentialMat: E = U diag(1,1,0) V^T -> R1 = U W V^T, R2 = U W^T V^T, t = U[:, 2]
VO_HDN inline void decompose_essential(const double* E, double* R1, double* R2, double* t)
{
    double At[9], W[3], Vt[9], U[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) At[c * 3 + r] = E[r * 3 + c];
    jacobi_svd_t<3, 3>(At, W, Vt, 3);          // rows of At = U^T
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) U[r * 3 + c] = At[c * 3 + r];
    if (det3(U) < 0) for (int k = 0; k < 9; k++) U[k] = -U[k];
    if (det3(Vt) < 0) for (int k = 0; k < 9; k++) Vt[k] = -Vt[k];
    const double Wm[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    double T[9];
    mat3_mul(U, Wm, T); mat3_mul(T, Vt, R1);
    mat3_mul(U, Wt, T); mat3_mul(T, Vt, R2);
    t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
}

// recoverPose's cheirality test of one correspondence for the pose [R|t] (P0 = [I|0]), fp64 DLT as cv::triangulatePoints
VO_HDN inline bool cheirality_ok(const double* R, const double* t, double x1, double y1, double x2, double y2, double dist)
{
    double At[16], W[4], Vt[16];
    const double P1[12] = {R[0], R[1], R[2], t[0], R[3], R[4], R[5], t[1], R[6], R[7], R[8], t[2]};
    for (int k = 0; k < 4; k++) {
        const double p0r0 = k == 0 ? 1.0 : 0.0, p0r1 = k == 1 ? 1.0 : 0.0, p0r2 = k == 2 ? 1.0 : 0.0;
        At[k * 4 + 0] = x1 * p0r2 - p0r0;
        At[k * 4 + 1] = y1 * p0r2 - p0r1;
        At[k * 4 + 2] = x2 * P1[8 + k] - P1[k];
        At[k * 4 + 3] = y2 * P1[8 + k] - P1[4 + k];
    }
    jacobi_svd_t<4, 4>(At, W, Vt, 0);
    const double* Q = Vt + 12;
    bool ok = Q[2] * Q[3] > 0;
    const double X = Q[0] / Q[3], Y = Q[1] / Q[3], Z = Q[2] / Q[3];
    ok = ok && Z < dist;
    const double zc = P1[8] * X + P1[9] * Y + P1[10] * Z + P1[11];
    return ok && zc > 0 && zc < dist;
}

} // namespace vomath
