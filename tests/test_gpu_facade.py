"""The drop-in C++ facade (include/compat/*.h, reference signatures) driven by a miniature of the
reference's main loop (tests/cpp/facade_main.cpp), compared frame by frame -- with the
FeatureSet / translation state carried across frames -- against the reference path
(cv2 through oracle/ref_path.py): bucketing, ages, circular check, triangulation, PnP."""
import os
import struct
import subprocess

import numpy as np
import pytest

from visual_odom_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sequence(w, h, seed, n_frames):
    frames = []
    base = synth.stereo_unit(w, h, seed)
    frames.append((base["l0"], base["r0"]))
    for k in range(1, n_frames):
        u = synth.stereo_unit(w, h, seed, rvec=synth.EGO_RVEC * k, tvec=synth.EGO_T * k)
        frames.append((u["l1"], u["r1"]))
    return base, frames


def test_facade_sequence_matches_reference(built, tmp_path):
    pytest.importorskip("cv2")
    from oracle import ref_path
    w, h, nf = 1241, 376, 4
    base, frames = _sequence(w, h, 21, nf)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<iii", w, h, nf))
        f.write(base["P_l"].astype(np.float32).tobytes()); f.write(base["P_r"].astype(np.float32).tobytes())
        for l, r in frames:
            f.write(l.tobytes()); f.write(r.tobytes())
    exe = os.path.join(ROOT, "tests", "cpp", "facade_main")
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(fout, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off)
        off += a.nbytes
        return a

    fs = ref_path.FeatureSet()
    translation = np.zeros(3)
    frame_pose = np.eye(4)
    for k in range(1, nf):
        l0, r0 = frames[k - 1]; l1, r1 = frames[k]
        pL0, pR0, pL1, pR1, info = ref_path.matching_features(l0, r0, l1, r1, fs, backend="cv2")
        X = ref_path.triangulate(base["P_l"], base["P_r"], pL0, pR0, "cv2")
        R, translation, inl, rvec = ref_path.tracking_frame2frame(base["P_l"], pL0, pL1, X, translation, "cv2")
        n = int(take(np.int32, 1)[0])
        assert n == len(pL0), f"frame {k}: {n} vs {len(pL0)} matched features"
        for ref in (pL0, pR0, pL1, pR1):
            assert np.array_equal(take(np.float32, 2 * n).reshape(-1, 2), ref)
        assert np.array_equal(take(np.float32, 3 * n).reshape(-1, 3), X)
        ni = int(take(np.int32, 1)[0])
        assert np.array_equal(take(np.int32, ni), inl), f"frame {k}: inlier list"
        Rg = take(np.float64, 9).reshape(3, 3); tg = take(np.float64, 3)
        assert np.linalg.norm(Rg - R) / np.linalg.norm(R) <= 1e-4
        assert np.linalg.norm(tg - translation) / np.linalg.norm(translation) <= 1e-4
        assert int(take(np.int32, 1)[0]) == fs.size()
        frame_pose = ref_path.integrate_pose(frame_pose, R, translation)
        pose_g = take(np.float64, 16).reshape(4, 4)
        assert np.abs(pose_g - frame_pose).max() <= 1e-9 * max(1.0, np.abs(frame_pose).max())
        # Frame::triangulateFeaturePoints == cv::triangulatePoints: 4 x N unit-norm homogeneous columns, bit for bit
        import cv2
        X4 = cv2.triangulatePoints(base["P_l"], base["P_r"], pL0.T.copy(), pR0.T.copy())
        assert np.array_equal(take(np.float32, 4 * n).reshape(4, n), X4)
        # the flag's default (mono_rotation = true): rotation from the five-point branch, translation from the same PnP
        focal = float(base["P_l"][0, 0]); pp = (float(base["P_l"][0, 2]), float(base["P_l"][1, 2]))
        E, emask = cv2.findEssentialMat(pL0, pL1, focal, pp, cv2.RANSAC, 0.999, 1.0)
        _, R_m, _t, _ = cv2.recoverPose(E, pL0, pL1, focal=focal, pp=pp, mask=emask.copy())
        Rm_g = take(np.float64, 9).reshape(3, 3); tm_g = take(np.float64, 3)
        assert np.linalg.norm(Rm_g - R_m) <= 1e-4 * np.linalg.norm(R_m), f"frame {k}: mono rotation"
        assert np.array_equal(tm_g, tg), f"frame {k}: the PnP translation does not depend on the flag"
        assert n > 100 and ni > 50
