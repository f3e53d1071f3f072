"""
oracle/pnp_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy (float64) restatement of the arithmetic behind the reference's pose solve
  cv::solvePnPRansac(points3D_t0, pointsLeft_t1, K, dist=0, rvec=0, t_prev, useExtrinsicGuess=true,
                     500, 0.5, 0.999, inliers, SOLVEPNP_ITERATIVE)        reference src/visualOdometry.cpp:176-178
  cv::Rodrigues(rvec, rotation)                                            reference src/visualOdometry.cpp:188
and of the triangulation call site
  cv::triangulatePoints + cv::convertPointsFromHomogeneous                  reference src/main.cpp:170-171

The arithmetic lives in OpenCV (un-vendored third-party dependency, pinned to 4.13.0 as
installed; modules/calib3d/src/{solvepnp,ptsetreg,epnp,calibration,triangulate}.cpp and
modules/core/src/rand.cpp -- not on disk).  The published algorithms are restated here and
pinned against cv2 4.13.0 by tests/test_oracle_pnp.py:
  * RNG (multiply-with-carry, state 2^64-1), 5-distinct-index subsets, adaptive niters rule
  * EPnP (Lepetit/Moreno-Noguer/Fua) exactly as OpenCV structures it (control points by PCA,
    12x12 M^T M null space, beta approximations 1/2/3, 5 Gauss-Newton steps, Horn alignment)
  * reprojection in f64 -> f32, squared error in f32, inlier iff err <= (float)(0.5^2)
  * final pose: Levenberg-Marquardt on (rvec, t) over the inliers from (rvec=0, t_prev)
"""
import numpy as np

# ----------------------------------------------------------------------------- RNG
RNG_COEFF = 4164903690
MASK32 = 0xFFFFFFFF
MASK64 = 0xFFFFFFFFFFFFFFFF


class CvRNG:
    """cv::RNG: s = (uint32)s * 4164903690 + (s >> 32); next() returns (uint32)s."""

    def __init__(self, state=MASK64):
        self.state = state if state else MASK64

    def next(self):
        self.state = ((self.state & MASK32) * RNG_COEFF + (self.state >> 32)) & MASK64
        return self.state & MASK32

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a)) + a


def rng_raw_stream(n, state=MASK64):
    r = CvRNG(state)
    return np.array([r.next() for _ in range(n)], dtype=np.uint32)


def ransac_subset(rng, count, model_points=5):
    """RANSACPointSetRegistrator::getSubset: model_points distinct indices, redraw on duplicate."""
    idx = []
    for _ in range(model_points):
        while True:
            v = rng.uniform(0, count)
            if v not in idx:
                break
        idx.append(v)
    return idx


def ransac_update_num_iters(p, ep, model_points, max_iters):
    """cv::RANSACUpdateNumIters"""
    p = max(p, 0.0); p = min(p, 1.0)
    ep = max(ep, 0.0); ep = min(ep, 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)   # DBL_MIN
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num = np.log(num)
    denom = np.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))      # cvRound: round-half-even


# ----------------------------------------------------------------------------- Rodrigues / projection
def rodrigues(rvec):
    """cv::Rodrigues vector -> matrix (f64)."""
    r = np.asarray(rvec, np.float64).reshape(3)
    theta = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c, s = np.cos(theta), np.sin(theta)
    c1 = 1.0 - c
    k = r / theta
    rrt = np.outer(k, k)
    rx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return c * np.eye(3) + c1 * rrt + s * rx


def rodrigues_inv(R):
    """cv::Rodrigues matrix -> vector (f64) for proper rotations."""
    R = np.asarray(R, np.float64)
    u, _, vt = np.linalg.svd(R)
    R = u @ vt
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((r * r).sum() * 0.25)
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) * 0.5
    c = min(max(c, -1.0), 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        rx = np.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        ry = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        rz = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (R[1, 2] > 0) != (ry * rz > 0):
            rz = -rz
        v = np.array([rx, ry, rz])
        theta /= np.linalg.norm(v)
        return v * theta
    vth = 1.0 / (2 * s) * theta
    return r * vth


def project_points(X, rvec, tvec, K):
    """cv::projectPoints with zero distortion, f64 arithmetic; returns f64 (N,2)."""
    R = rodrigues(rvec)
    X = np.asarray(X, np.float64)
    Xc = X @ R.T + np.asarray(tvec, np.float64).reshape(1, 3)
    z = np.where(Xc[:, 2] != 0, 1.0 / Xc[:, 2], 1.0)
    x = Xc[:, 0] * z
    y = Xc[:, 1] * z
    return np.stack([x * K[0, 0] + K[0, 2], y * K[1, 1] + K[1, 2]], axis=1)


def reproj_err_f32(X, x, rvec, tvec, K):
    """PnPRansacCallback::computeError: projections stored f32, squared distance in f32."""
    p = project_points(X, rvec, tvec, K).astype(np.float32)
    d = np.asarray(x, np.float32) - p
    dx2 = d[:, 0] * d[:, 0]
    dy2 = d[:, 1] * d[:, 1]
    return (dx2 + dy2).astype(np.float32)


# ----------------------------------------------------------------------------- EPnP
def _svd_ut(a):
    """cvSVD(A, W, U^T) for symmetric PSD A: rows of the returned matrix are singular vectors,
    singular values descending."""
    u, w, _ = np.linalg.svd(a)
    return w, u.T


def epnp(X, x, K, null_rot=None):
    """OpenCV's epnp class driven the way solvePnP(SOLVEPNP_EPNP) does it.
    X: (n,3), x: (n,2) pixel coordinates, K f64 3x3.  Returns (R 3x3, t 3)."""
    X = np.asarray(X, np.float64)
    n = len(X)
    fu, fv, uc, vc = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # solvePnP: undistortPoints (zero distortion) -> normalised coords stored as f32,
    # epnp::init_points maps them back with fu/uc in f64
    xn = ((np.asarray(x, np.float32).astype(np.float64) - [uc, vc]) * [1.0 / fu, 1.0 / fv]).astype(np.float32)
    us = xn.astype(np.float64) * [fu, fv] + [uc, vc]

    # choose_control_points
    cws = np.zeros((4, 3))
    cws[0] = X.sum(axis=0) / n
    pw0 = X - cws[0]
    dc, uct = _svd_ut(pw0.T @ pw0)
    for i in range(1, 4):
        cws[i] = cws[0] + np.sqrt(dc[i - 1] / n) * uct[i - 1]
    # compute_barycentric_coordinates
    cc = (cws[1:4] - cws[0]).T
    ci = np.linalg.pinv(cc)
    al = np.zeros((n, 4))
    al[:, 1:4] = (X - cws[0]) @ ci.T
    al[:, 0] = 1.0 - al[:, 1] - al[:, 2] - al[:, 3]
    # M
    M = np.zeros((2 * n, 12))
    for i in range(4):
        M[0::2, 3 * i] = al[:, i] * fu
        M[0::2, 3 * i + 2] = al[:, i] * (uc - us[:, 0])
        M[1::2, 3 * i + 1] = al[:, i] * fv
        M[1::2, 3 * i + 2] = al[:, i] * (vc - us[:, 1])
    _, ut = _svd_ut(M.T @ M)
    if null_rot is not None:    # experiment hook: rotate the (degenerate) null-space basis
        c, s = np.cos(null_rot), np.sin(null_rot)
        v11, v10 = ut[11].copy(), ut[10].copy()
        ut = ut.copy()
        ut[11] = c * v11 + s * v10
        ut[10] = -s * v11 + c * v10
    v = [ut[11], ut[10], ut[9], ut[8]]
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        for j, (a, b) in enumerate(pairs):
            dv[i, j] = v[i][3 * a:3 * a + 3] - v[i][3 * b:3 * b + 3]
    L = np.zeros((6, 10))
    for i in range(6):
        d = dv[:, i]
        L[i] = [d[0] @ d[0], 2 * d[0] @ d[1], d[1] @ d[1], 2 * d[0] @ d[2], 2 * d[1] @ d[2], d[2] @ d[2],
                2 * d[0] @ d[3], 2 * d[1] @ d[3], 2 * d[2] @ d[3], d[3] @ d[3]]
    rho = np.array([((cws[a] - cws[b]) ** 2).sum() for a, b in pairs])

    def lstsq(A, b):
        return np.linalg.lstsq(A, b, rcond=None)[0]

    def approx1():
        b4 = lstsq(L[:, [0, 1, 3, 6]], rho)
        if b4[0] < 0:
            b0 = np.sqrt(-b4[0]); return np.array([b0, -b4[1] / b0, -b4[2] / b0, -b4[3] / b0])
        b0 = np.sqrt(b4[0]); return np.array([b0, b4[1] / b0, b4[2] / b0, b4[3] / b0])

    def approx2():
        b3 = lstsq(L[:, [0, 1, 2]], rho)
        if b3[0] < 0:
            b0 = np.sqrt(-b3[0]); b1 = np.sqrt(-b3[2]) if b3[2] < 0 else 0.0
        else:
            b0 = np.sqrt(b3[0]); b1 = np.sqrt(b3[2]) if b3[2] > 0 else 0.0
        if b3[1] < 0: b0 = -b0
        return np.array([b0, b1, 0.0, 0.0])

    def approx3():
        b5 = lstsq(L[:, [0, 1, 2, 3, 4]], rho)
        if b5[0] < 0:
            b0 = np.sqrt(-b5[0]); b1 = np.sqrt(-b5[2]) if b5[2] < 0 else 0.0
        else:
            b0 = np.sqrt(b5[0]); b1 = np.sqrt(b5[2]) if b5[2] > 0 else 0.0
        if b5[1] < 0: b0 = -b0
        return np.array([b0, b1, b5[3] / b0, 0.0])

    def gauss_newton(be):
        be = be.copy()
        for _ in range(5):
            A = np.zeros((6, 4)); b = np.zeros(6)
            for i in range(6):
                r = L[i]
                A[i] = [2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3],
                        r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3],
                        r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3],
                        r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3]]
                b[i] = rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1] +
                                 r[3] * be[0] * be[2] + r[4] * be[1] * be[2] + r[5] * be[2] * be[2] +
                                 r[6] * be[0] * be[3] + r[7] * be[1] * be[3] + r[8] * be[2] * be[3] +
                                 r[9] * be[3] * be[3])
            be = be + lstsq(A, b)
        return be

    def compute_R_and_t(be):
        ccs = np.zeros((4, 3))
        for i in range(4):
            ccs += be[i] * v[i].reshape(4, 3)
        pcs = al @ ccs
        if pcs[0, 2] < 0:
            ccs = -ccs; pcs = -pcs
        pc0 = pcs.sum(axis=0) / n
        pw0_ = X.sum(axis=0) / n
        abt = (pcs - pc0).T @ (X - pw0_)
        u, _, vt = np.linalg.svd(abt)
        R = u @ vt
        if np.linalg.det(R) < 0:
            R[2] = -R[2]
        t = pc0 - R @ pw0_
        Xc = X @ R.T + t
        ue = uc + fu * Xc[:, 0] / Xc[:, 2]
        ve = vc + fv * Xc[:, 1] / Xc[:, 2]
        err = np.sqrt((us[:, 0] - ue) ** 2 + (us[:, 1] - ve) ** 2).sum() / n
        return err, R, t

    sols = [compute_R_and_t(gauss_newton(f())) for f in (approx1, approx2, approx3)]
    N = 0
    if sols[1][0] < sols[N][0]: N = 1
    if sols[2][0] < sols[N][0]: N = 2
    return sols[N][1], sols[N][2]
