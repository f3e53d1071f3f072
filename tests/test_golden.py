"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from cv2 4.13.0
through the reference glue).  CPU: the oracle restatements reproduce them.  GPU: the library does."""
import glob
import os

import numpy as np
import pytest

from visual_odom_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
sys_path_cfg = None


def _load(path):
    g = np.load(path)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(path), "make_golden.py"))
    return g


def _config(path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(path), "make_golden.py"))
    # parse CONFIGS without importing cv2: the dict literal is small and static
    src = open(spec.origin).read()
    ns = {}
    start = src.index("CONFIGS = {"); end = src.index("}\n", start) + 1
    exec(src[start:end], ns)
    return ns["CONFIGS"][os.path.basename(path)[:-4]]


def _unit(cfg):
    w, h, seed, scene, n_sel, cal = cfg
    return synth.stereo_unit(w, h, seed, cal=synth.KITTI00 if cal == "kitti" else synth.ZED, scene=scene)


def _crc(u, w, h):
    return np.array([int(np.bitwise_xor.reduce(u[k].astype(np.uint32).ravel() * np.arange(1, w * h + 1, dtype=np.uint32)))
                     for k in ("l0", "r0", "l1", "r1")], np.uint32)


def _check_raw(g, raws):
    """Per-call raw LK outputs: stored in full up to 2000 features, as a CRC of the float bits above (fixtures made
    before the CRC existed carry only the arrays)."""
    import zlib
    for i, k in enumerate(("r0", "r1", "l1", "l0_ret")):
        if "raw_" + k in g.files:
            assert np.array_equal(raws[i], g["raw_" + k]), k
        if "raw_crc" in g.files:
            assert zlib.crc32(np.ascontiguousarray(raws[i], np.float32).tobytes()) == int(g["raw_crc"][i]), k


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(built, path):
    from oracle import ref_path, cref, pnp_ref
    g = np.load(path)
    cfg = _config(path)
    w, h, seed, scene, n_sel, cal = cfg
    u = _unit(cfg)
    assert np.array_equal(_crc(u, w, h), g["image_crc"]), "synthetic generator is not reproducing the golden inputs"
    corners, _ = cref.fast_detect(u["l0"])
    assert len(corners) == int(g["n_corners"]) and np.array_equal(corners[:64], g["corners_head"])
    pts = synth.select_features(corners, n_sel)
    assert np.array_equal(pts, g["pts"])
    fs = ref_path.FeatureSet(); fs.points = pts.copy(); fs.ages = np.zeros(len(pts), np.int32)
    cm = ref_path.circular_matching(u["l0"], u["r0"], u["l1"], u["r1"], pts, fs, backend="c")
    assert np.array_equal(cm["raw"]["status"], g["status"])
    _check_raw(g, [cm["raw"][k] for k in ("r0", "r1", "l1", "l0_ret")])
    assert np.array_equal(cm["kept_idx"], g["kept3"])
    ok = ref_path.check_valid_match(cm["l0"], cm["l0_ret"], 0)
    assert np.array_equal(cm["kept_idx"][ok], g["kept"])
    if len(g["kept"]) <= 600:          # pure-Python restatement: small cases only
        X = pnp_ref.triangulate(u["P_l"], u["P_r"], g["l0"], g["r0"])
        assert np.array_equal(X, g["X"])
        res = pnp_ref.solve_pnp_ransac(g["X"], g["l1"], u["K"], np.zeros(3), g["t_prev"], confidence=ref_path.PNP_CONFIDENCE)
        assert np.array_equal(res["inliers"], g["inliers"])
        assert np.linalg.norm(pnp_ref.rodrigues(res["rvec"]) - g["R"]) <= 1e-6  # LM stops at a FLT_EPSILON relative step
        assert np.linalg.norm(res["tvec"] - g["t"]) <= 1e-6 * np.linalg.norm(g["t"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_gpu_reproduces_golden(ctx, path):
    g = np.load(path)
    cfg = _config(path)
    w, h, seed, scene, n_sel, cal = cfg
    u = _unit(cfg)
    assert np.array_equal(_crc(u, w, h), g["image_crc"])
    ctx.batch_configure(w, h, 1, u["P_l"], u["P_r"])
    arr, keep, pitch = ctx.make_units([dict(u, n_select=n_sel, t_prev=tuple(g["t_prev"]))])
    res = ctx.frame_batch(arr, pitch)[0]
    got = ctx.batch_fetch(0, res)
    assert res["n_detected"] == int(g["n_corners"])
    assert np.array_equal(got["pts_in"], g["pts"])
    assert res["n_tracked"] == len(g["kept3"])
    assert np.array_equal(got["kept_idx"], g["kept"]), "tracked-feature indices"
    for k in ("l0", "r0", "l1", "r1"):
        assert np.array_equal(got[k], g[k]), k
    assert np.array_equal(got["X"], g["X"])
    assert np.array_equal(got["inliers"], g["inliers"]), "RANSAC inlier list"
    assert np.linalg.norm(res["R"] - g["R"]) / np.linalg.norm(g["R"]) <= 1e-4
    assert np.linalg.norm(res["tvec"] - g["t"]) / np.linalg.norm(g["t"]) <= 1e-4
    # raw per-call LK outputs through the single-unit C-ABI entry point
    cm = ctx.circular_match(u["l0"], u["r0"], u["l1"], u["r1"], g["pts"])
    assert np.array_equal(cm["status4"], g["status"])
    _check_raw(g, [cm["raw4"][i] for i in range(4)])
