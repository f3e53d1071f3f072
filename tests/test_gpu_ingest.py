"""Device side of the ingest row: k_bgr_to_gray against cv2.cvtColor, colour inputs through the
sequence mode, and the PNG reader -> pinned buffers -> vo_seq_push pipeline against the reference path on the decoded
images."""
import os

import numpy as np
import pytest

from visual_odom_b200 import capi, synth

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("w,h", [(1241, 376), (640, 480), (37, 5), (1, 1)])
def test_bgr_to_gray_matches_cvtcolor(ctx, w, h):
    rng = np.random.default_rng(w + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(ctx.bgr_to_gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def _colourise(gray, seed):
    """A colour image whose BGR2GRAY is NOT simply one of its channels."""
    rng = np.random.default_rng(seed)
    tint = rng.integers(-20, 21, gray.shape + (3,))
    return np.clip(gray[..., None].astype(np.int32) + tint, 0, 255).astype(np.uint8)


def test_colour_sequence_equals_gray_sequence_of_converted_images(ctx):
    w, h, nf = 1241, 376, 5
    base = synth.stereo_unit(w, h, 31)
    frames = [(base["l0"], base["r0"])]
    for k in range(1, nf):
        u = synth.stereo_unit(w, h, 31, rvec=np.array([0.001, -0.004, 0.0005]) * k, tvec=np.array([0.01, -0.003, -0.2]) * k)
        frames.append((u["l1"], u["r1"]))
    colour = [(_colourise(l, 2 * i), _colourise(r, 2 * i + 1)) for i, (l, r) in enumerate(frames)]
    gray = [(cv2.cvtColor(l, cv2.COLOR_BGR2GRAY), cv2.cvtColor(r, cv2.COLOR_BGR2GRAY)) for l, r in colour]
    ctx.seq_begin(gray[0][0], gray[0][1], base["P_l"], base["P_r"])
    ref = [ctx.seq_push(l, r) for l, r in gray[1:]]
    pose_ref = ctx.seq_pose()
    ctx.seq_begin_bgr(colour[0][0], colour[0][1], base["P_l"], base["P_r"])
    for k, (l, r) in enumerate(colour[1:]):
        got = ctx.seq_push_bgr(l, r)
        for key in ("n_features", "n_tracked", "n_valid", "n_inliers"):
            assert got[key] == ref[k][key], (k, key)
        for key in ("l0", "r0", "l1", "r1", "R", "tvec"):
            assert np.array_equal(got[key], ref[k][key]), (k, key)
    assert np.array_equal(ctx.seq_pose(), pose_ref)
    assert ref[-1]["n_inliers"] > 20


def test_png_reader_feeds_sequence_mode(ctx, tmp_path):
    """KITTI layout on disk -> SequenceReader (pinned ring) -> vo_seq_push on the raw pointers, against the
    reference path (cv2.imread + cvtColor + cv2 LK / PnP through oracle/ref_path.py) on the same files."""
    from oracle import ref_path
    w, h, nf = 1241, 376, 6
    base = synth.stereo_unit(w, h, 12)
    frames = [(base["l0"], base["r0"])]
    for k in range(1, nf):
        u = synth.stereo_unit(w, h, 12, rvec=np.array([0.001, -0.004, 0.0005]) * k, tvec=np.array([0.01, -0.003, -0.2]) * k)
        frames.append((u["l1"], u["r1"]))
    os.makedirs(tmp_path / "image_0"); os.makedirs(tmp_path / "image_1")
    for i, (l, r) in enumerate(frames):
        cv2.imwrite(str(tmp_path / "image_0" / ("%06d.png" % i)), l)
        cv2.imwrite(str(tmp_path / "image_1" / ("%06d.png" % i)), r)
    rd = capi.SequenceReader(str(tmp_path), 0, nf, threads=4, depth=3)
    lp, rp, rw, rh, pitch, ch, fid = rd.next_ptr()
    assert (rw, rh, ch, fid) == (w, h, 1, 0)
    ctx.seq_begin_ptr(w, h, lp, rp, pitch, base["P_l"], base["P_r"], ch)
    fs = ref_path.FeatureSet(); translation = np.zeros(3); pose = np.eye(4)

    def load(i):
        out = []
        for cam in range(2):
            c = cv2.imread(str(tmp_path / ("image_%d" % cam) / ("%06d.png" % i)), cv2.IMREAD_COLOR)
            out.append(cv2.cvtColor(c, cv2.COLOR_BGR2GRAY))
        return out
    prev = load(0)
    for k in range(1, nf):
        lp, rp, rw, rh, pitch, ch, fid = rd.next_ptr()
        got = ctx.seq_push_ptr(lp, rp, pitch, ch)
        cur = load(k)
        pL0, pR0, pL1, pR1, info = ref_path.matching_features(prev[0], prev[1], cur[0], cur[1], fs, backend="cv2")
        X = ref_path.triangulate(base["P_l"], base["P_r"], pL0, pR0, "cv2")
        R, translation, inl, rvec = ref_path.tracking_frame2frame(base["P_l"], pL0, pL1, X, translation, "cv2")
        pose = ref_path.integrate_pose(pose, R, translation)
        prev = cur
        assert got["n_valid"] == len(pL0) and got["n_inliers"] == len(inl), k
        assert np.linalg.norm(got["R"] - R) / np.linalg.norm(R) <= 1e-4
        assert np.linalg.norm(got["tvec"] - translation) / np.linalg.norm(translation) <= 1e-4
    assert np.abs(ctx.seq_pose() - pose).max() <= 1e-6 * max(1.0, np.abs(pose).max())
    rd.close()
