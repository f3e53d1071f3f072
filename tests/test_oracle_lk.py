"""Pins the C restatement (oracle/lk_ref.c, fast_ref.c) bit-for-bit against cv2 4.13.0 -- the
third-party implementation the reference's calls resolve to."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from visual_odom_b200 import synth


def _cv_lk(a, b, pts):
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    o, s, e = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3, criteria=crit,
                                       flags=0, minEigThreshold=0.001)
    return o.reshape(-1, 2), s.ravel(), e.ravel()


def test_pyrdown_and_scharr_bit_exact(built):
    from oracle import cref
    u = synth.stereo_unit(333, 129, 0, scene="v0")
    img = u["l0"]
    for _ in range(3):
        d = cref.pyr_down(img)
        assert np.array_equal(d, cv2.pyrDown(img))
        img = d
    _, pyr = cv2.buildOpticalFlowPyramid(u["l0"], (21, 21), 3, withDerivatives=True)
    # 129 rows: level 3 would be 17 rows <= window -> OpenCV truncates the pyramid to 3 images
    assert len(pyr) // 2 == 3 and cref.Pyramid(u["l0"]).nlevels() == 3
    lvl = u["l0"]
    for l in range(len(pyr) // 2):
        assert np.array_equal(cref.scharr(lvl), pyr[2 * l + 1])
        lvl = cref.pyr_down(lvl)


@pytest.mark.parametrize("scene,seed", [("v1", 0), ("v0", 1)])
def test_lk_bit_exact_vs_cv2(built, scene, seed):
    from oracle import cref
    w, h = 1241, 376
    u = synth.stereo_unit(w, h, seed, scene=scene)
    corners = np.array([k.pt for k in cv2.FastFeatureDetector_create(20, True).detect(u["l0"])], np.float32)
    pts = synth.select_features(corners, 1200)
    rng = np.random.default_rng(seed)
    pts = np.concatenate([pts + rng.uniform(-0.5, 0.5, pts.shape).astype(np.float32),
                          np.array([[0, 0], [-5, 3], [w - 1, h - 1], [w + 5, 10], [3, h + 30], [-30, -30]], np.float32)])
    for a, b in ((u["l0"], u["r0"]), (u["r0"], u["r1"])):
        co, cs, ce = _cv_lk(a, b, pts)
        o, s, e = cref.lk_track(a, b, pts)
        assert np.array_equal(s, cs)
        assert np.array_equal(o, co)
        assert np.array_equal(e[cs == 1], ce[cs == 1])


def test_sum_order_matters(built):
    """The SIMD-lane summation order is part of the contract: a plain row-major float sum
    (mode 9) is NOT bit-identical to cv2, mode 0 is."""
    from oracle import cref
    u = synth.stereo_unit(640, 360, 4, scene="v0")
    corners = np.array([k.pt for k in cv2.FastFeatureDetector_create(20, True).detect(u["l0"])], np.float32)
    pts = synth.select_features(corners, 1500)
    co, cs, _ = _cv_lk(u["l0"], u["r0"], pts)
    try:
        cref.lib().lk_set_sum_mode(9)
        o9, _, _ = cref.lk_track(u["l0"], u["r0"], pts)
    finally:
        cref.lib().lk_set_sum_mode(0)
    o0, s0, _ = cref.lk_track(u["l0"], u["r0"], pts)
    assert np.array_equal(o0, co) and np.array_equal(s0, cs)
    assert not np.array_equal(o9, co)
    assert np.abs(o9 - co).max() < 1e-2


def test_fast_list_exact_vs_cv2(built):
    from oracle import cref
    for (w, h, seed) in [(1241, 376, 0), (320, 200, 1), (9, 8, 2)]:
        img = synth.stereo_unit(w, h, seed, scene="v0")["l0"]
        kps = cv2.FastFeatureDetector_create(20, True).detect(img)
        xy, resp = cref.fast_detect(img)
        assert len(kps) == len(xy)
        if len(kps):
            assert np.array_equal(xy, np.array([k.pt for k in kps], np.float32))
            assert np.array_equal(resp, np.array([k.response for k in kps], np.float32))
