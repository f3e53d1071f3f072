"""KITTI-shaped synthetic stereo work units (SURVEY.md section 8d "Synthetic inputs").

There is no network and no dataset in this environment, so every test / benchmark input is
generated here, deterministically from a seed.  cv2 is used only as an image library
(resize / blur / remap); nothing in here is on the measured path.

scene "v1": three fronto-parallel textured planes (stereo disparities 4 / 12 / 40 px at KITTI-00
intrinsics) seen by a calibrated stereo rig that moves by a known ego-motion between t0 and t1;
each of the four views (L0, R0, L1, R1) is rendered through the exact plane homography and gets
independent N(0,1) sensor noise.  scene "v0": pure-translation crops of one texture.
"""
import numpy as np
import cv2

KITTI00 = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=-386.1448)      # calibration/kitti00.yaml:8-14
ZED = dict(fx=684.367919921875 * 1.5, fy=684.367919921875 * 1.5, cx=960.0, cy=540.0,
           bf=-82.12415128946304 * 1.5)                                               # calibration/zed.yaml scaled to 1920x1080
EGO_RVEC = np.array([0.004, -0.02, 0.001])
EGO_T = np.array([0.03, -0.01, -0.9])


def proj_matrices(cal=KITTI00):
    """P_l, P_r exactly as reference src/main.cpp:73-74 builds them (3x4 float32)."""
    fx, fy, cx, cy, bf = (np.float32(cal[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    P_l = np.array([[fx, 0, cx, 0], [0, fy, cy, 0], [0, 0, 1, 0]], np.float32)
    P_r = np.array([[fx, 0, cx, bf], [0, fy, cy, 0], [0, 0, 1, 0]], np.float32)
    return P_l, P_r


def texture(h, w, seed, margin=96):
    rng = np.random.default_rng(seed)
    lo = rng.integers(0, 256, size=((h + 2 * margin) // 4 + 2, (w + 2 * margin) // 4 + 2)).astype(np.uint8)
    up = cv2.resize(lo, (w + 2 * margin, h + 2 * margin), interpolation=cv2.INTER_CUBIC)
    return cv2.GaussianBlur(up, (0, 0), 1.0)


def _rodrigues(r):
    R, _ = cv2.Rodrigues(np.asarray(r, np.float64).reshape(3, 1))
    return R


def stereo_unit(w=1241, h=376, seed=0, cal=KITTI00, scene="v1", noise=1.0, rvec=EGO_RVEC, tvec=EGO_T):
    """Returns dict(l0, r0, l1, r1 : uint8 HxW, P_l, P_r, K, rvec, tvec)."""
    P_l, P_r = proj_matrices(cal)
    K = P_l[:, :3].astype(np.float64)
    rng = np.random.default_rng(1000 + seed)
    margin = 96
    if scene == "v0":
        T = texture(h, w, seed, margin)
        def crop(dx, dy):
            return T[margin + dy:margin + dy + h, margin + dx:margin + dx + w].copy()
        views = [crop(0, 0), crop(6, 0), crop(-2, -1), crop(4, -1)]          # L0, R0, L1, R1
    else:
        disp = [4.0, 12.0, 40.0]
        Z = [-float(cal["bf"]) / d for d in disp]
        # L0 band membership: far plane on top, near plane at the bottom
        bands = [(0, int(h * 0.40)), (int(h * 0.40), int(h * 0.72)), (int(h * 0.72), h)]
        texs = [texture(h, w, seed * 3 + k, margin) for k in range(3)]
        base = float(cal["bf"]) / float(cal["fx"])            # X_r = X_l + (base, 0, 0)
        R1 = _rodrigues(rvec)
        poses = [(np.eye(3), np.zeros(3)),                    # L0
                 (np.eye(3), np.array([base, 0, 0])),         # R0
                 (R1, np.asarray(tvec, np.float64)),          # L1:  X_1 = R X_0 + t
                 (R1, np.asarray(tvec, np.float64) + np.array([base, 0, 0]))]   # R1
        uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        pix = np.stack([uu.ravel(), vv.ravel(), np.ones(w * h)], 0)
        Kinv = np.linalg.inv(K)
        views = []
        for (R, t) in poses:
            img = np.zeros((h, w), np.uint8)
            filled = np.zeros((h, w), bool)
            for k in (2, 1, 0):                                # nearest plane wins
                n = np.array([0.0, 0.0, 1.0])
                H = K @ (R + np.outer(t, n) / Z[k]) @ Kinv   # L0 pixel -> view pixel
                p0 = np.linalg.inv(H) @ pix
                x0 = (p0[0] / p0[2]).reshape(h, w)
                y0 = (p0[1] / p0[2]).reshape(h, w)
                inside = (y0 >= bands[k][0]) & (y0 < bands[k][1]) & ~filled
                m = cv2.remap(texs[k], (x0 + margin).astype(np.float32), (y0 + margin).astype(np.float32),
                              cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
                img[inside] = m[inside]
                filled |= inside
            if not filled.all():                               # slivers between bands: far plane
                H = K @ (R + np.outer(t, np.array([0.0, 0.0, 1.0])) / Z[0]) @ Kinv
                p0 = np.linalg.inv(H) @ pix
                m = cv2.remap(texs[0], ((p0[0] / p0[2]).reshape(h, w) + margin).astype(np.float32),
                              ((p0[1] / p0[2]).reshape(h, w) + margin).astype(np.float32),
                              cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
                img[~filled] = m[~filled]
            views.append(img)
    out = []
    for v in views:
        if noise > 0:
            v = np.clip(v.astype(np.int32) + np.rint(rng.normal(0, noise, v.shape)).astype(np.int32), 0, 255).astype(np.uint8)
        out.append(np.ascontiguousarray(v))
    return dict(l0=out[0], r0=out[1], l1=out[2], r1=out[3], P_l=P_l, P_r=P_r,
                K=P_l[:, :3].copy(), rvec=np.asarray(rvec, np.float64), tvec=np.asarray(tvec, np.float64))


def select_features(corners, n):
    """Even-stride selection over the raster-ordered FAST list (SURVEY.md 8d "Features")."""
    corners = np.asarray(corners, np.float32).reshape(-1, 2)
    m = len(corners)
    if m == 0 or n <= 0:
        return corners[:0]
    if n >= m:
        return corners.copy()
    idx = (np.arange(n, dtype=np.int64) * (m - 1)) // (n - 1) if n > 1 else np.array([0])
    return corners[idx]


def pnp_stress_set(n=1500, sigma=0.15, outlier_frac=0.3, seed=0, cal=KITTI00, rvec=EGO_RVEC, tvec=EGO_T):
    """3-D points + noisy projections + gross outliers (SURVEY.md 8d "PnP stress set")."""
    rng = np.random.default_rng(seed)
    P_l, _ = proj_matrices(cal)
    K = P_l[:, :3].astype(np.float64)
    X = np.stack([rng.uniform(-30, 30, n), rng.uniform(-3, 6, n), rng.uniform(6, 80, n)], 1).astype(np.float32)
    R = _rodrigues(rvec)
    Xc = X.astype(np.float64) @ R.T + np.asarray(tvec, np.float64)
    x = np.stack([Xc[:, 0] / Xc[:, 2] * K[0, 0] + K[0, 2], Xc[:, 1] / Xc[:, 2] * K[1, 1] + K[1, 2]], 1)
    x += rng.normal(0, sigma, x.shape)
    out = rng.random(n) < outlier_frac
    x[out] += rng.uniform(-15, 15, (int(out.sum()), 2))
    return X, x.astype(np.float32), P_l[:, :3].copy(), out


def essential_stress_set(n, sigma, outlier_frac, seed, cal=KITTI00, rvec=EGO_RVEC, tvec=EGO_T):
    """Image points of n world points before / after the ego-motion (pixel noise sigma on both, a fraction of gross
    outliers in the second view) + the (focal, principal point) the reference's mono branch reads from P_l as floats
    (reference src/visualOdometry.cpp:144-145).  Returns (p_t0, p_t1, focal, pp)."""
    rng = np.random.default_rng(seed)
    P_l, _ = proj_matrices(cal)
    K = P_l[:, :3].astype(np.float64)
    X = np.stack([rng.uniform(-30, 30, n), rng.uniform(-3, 6, n), rng.uniform(6, 80, n)], 1)
    x0 = np.stack([X[:, 0] / X[:, 2] * K[0, 0] + K[0, 2], X[:, 1] / X[:, 2] * K[1, 1] + K[1, 2]], 1)
    Xc = X @ _rodrigues(rvec).T + np.asarray(tvec, np.float64)
    x1 = np.stack([Xc[:, 0] / Xc[:, 2] * K[0, 0] + K[0, 2], Xc[:, 1] / Xc[:, 2] * K[1, 1] + K[1, 2]], 1)
    x0 += rng.normal(0, sigma, x0.shape); x1 += rng.normal(0, sigma, x1.shape)
    o = rng.random(n) < outlier_frac
    x1[o] += rng.uniform(-15, 15, (int(o.sum()), 2))
    return x0.astype(np.float32), x1.astype(np.float32), float(np.float32(K[0, 0])), (float(np.float32(K[0, 2])), float(np.float32(K[1, 2])))
