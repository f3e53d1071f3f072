"""GPU parity: K1 pyramid + K2 LK ring + K3 filters vs the oracle (bit-exact).

The oracle (oracle/lk_ref.c) is pinned bit-for-bit against cv2 4.13.0 in test_oracle_lk.py, so
bit-equality here is bit-equality with the reference's cv::calcOpticalFlowPyrLK.
"""
import numpy as np
import pytest

from visual_odom_b200 import synth

pytestmark = pytest.mark.gpu


def _border_points(w, h):
    return np.array([[0, 0], [-5, 3], [w - 1, h - 1], [w + 5, 10], [3, h + 30], [-30, -30], [w - 0.5, 5.5],
                     [10.25, -21.5], [w + 20.9, h / 2], [5, -10.99], [-11.01, 7]], np.float32)


@pytest.mark.parametrize("w,h,seed,scene", [(1241, 376, 0, "v1"), (1241, 376, 1, "v0"), (640, 480, 2, "v1"), (333, 129, 3, "v0")])
def test_single_call_bit_exact(ctx, w, h, seed, scene):
    from oracle import cref
    u = synth.stereo_unit(w, h, seed, scene=scene)
    corners, _ = cref.fast_detect(u["l0"])
    pts = synth.select_features(corners, 1500)
    rng = np.random.default_rng(seed)
    pts = np.concatenate([pts + rng.uniform(-0.5, 0.5, pts.shape).astype(np.float32), _border_points(w, h)])
    for a, b in ((u["l0"], u["r0"]), (u["l0"], u["l1"])):
        ro, rs, re = cref.lk_track(a, b, pts)
        go, gs, ge = ctx.lk_track(a, b, pts)
        assert np.array_equal(gs, rs), f"status differs at {np.nonzero(gs != rs)[0][:10]}"
        assert np.array_equal(go, ro), f"positions differ: max {np.abs(go - ro).max()} at {np.nonzero((go != ro).any(1))[0][:10]}"
        ok = rs == 1
        assert np.array_equal(ge[ok], re[ok])
        assert ok.sum() > 0.8 * len(pts)


def test_ring_and_filters_bit_exact(ctx):
    from oracle import cref, ref_path
    w, h = 1241, 376
    u = synth.stereo_unit(w, h, 5)
    corners, _ = cref.fast_detect(u["l0"])
    pts = np.concatenate([synth.select_features(corners, 2000), _border_points(w, h)])
    fs = ref_path.FeatureSet()
    fs.points = pts.copy(); fs.ages = np.arange(len(pts), dtype=np.int32) % 7
    ages0 = fs.ages.copy()
    ref = ref_path.circular_matching(u["l0"], u["r0"], u["l1"], u["r1"], pts, fs, backend="c")
    got = ctx.circular_match(u["l0"], u["r0"], u["l1"], u["r1"], pts, ages=ages0)
    assert np.array_equal(got["status4"], ref["raw"]["status"])
    for k, name in enumerate(("r0", "r1", "l1", "l0_ret")):
        assert np.array_equal(got["raw4"][k], ref["raw"][name]), name
    assert np.array_equal(got["kept_idx"], ref["kept_idx"])
    for name in ("l0", "r0", "l1", "r1", "l0_ret"):
        assert np.array_equal(got[name], ref[name]), name
    assert np.array_equal(got["ages"], fs.ages)
    assert len(got["kept_idx"]) > 1000


def test_empty_and_tiny_inputs(ctx):
    u = synth.stereo_unit(320, 240, 7, scene="v0")
    o, s, e = ctx.lk_track(u["l0"], u["r0"], np.zeros((0, 2), np.float32))
    assert len(o) == 0 and len(s) == 0
    o, s, e = ctx.lk_track(u["l0"], u["r0"], np.array([[100.5, 80.25]], np.float32))
    from oracle import cref
    ro, rs, re = cref.lk_track(u["l0"], u["r0"], np.array([[100.5, 80.25]], np.float32))
    assert np.array_equal(o, ro) and np.array_equal(s, rs)


def test_random_points_everywhere(ctx):
    """Property-style sweep: points scattered far inside / on / outside the image (every early-out of
    the level loop: prev window out of range, next window leaving the image mid-iteration, low
    texture -> min-eigenvalue rejection) stay bit-identical to the oracle, status included."""
    from oracle import cref
    rng = np.random.default_rng(1234)
    for (w, h, seed, scene) in [(300, 200, 1, "v0"), (96, 64, 2, "v0"), (1241, 376, 3, "v1")]:
        u = synth.stereo_unit(w, h, seed, scene=scene)
        a, b = u["l0"].copy(), u["l1"].copy()
        a[h // 3: h // 3 + 40, w // 4: w // 4 + 60] = 128          # a flat patch: minEig < 1e-3
        b[h // 3: h // 3 + 40, w // 4: w // 4 + 60] = 128
        n = 1200
        pts = np.stack([rng.uniform(-60, w + 60, n), rng.uniform(-60, h + 60, n)], 1).astype(np.float32)
        pts[:50] = np.round(pts[:50])                               # integer positions (zero fractional weights)
        ro, rs, re = cref.lk_track(a, b, pts)
        go, gs, ge = ctx.lk_track(a, b, pts)
        assert np.array_equal(gs, rs) and np.array_equal(go, ro)
        assert np.array_equal(ge[rs == 1], re[rs == 1])
        assert 0 < rs.sum() < n                                     # both outcomes are exercised


def test_pyramid_truncation_small_image(ctx):
    """Levels not larger than the 21x21 window end the pyramid (OpenCV rule): 100x44 has 2 images."""
    from oracle import cref
    u = synth.stereo_unit(100, 44, 9, scene="v0")
    assert cref.Pyramid(u["l0"]).nlevels() == 2
    pts = np.array([[20.5, 12.25], [50, 22], [80.75, 30.5], [5, 40], [99, 43]], np.float32)
    ro, rs, re = cref.lk_track(u["l0"], u["r0"], pts)
    go, gs, ge = ctx.lk_track(u["l0"], u["r0"], pts)
    assert np.array_equal(gs, rs) and np.array_equal(go, ro)
