"""KITTI accuracy metric: the library's evaluator against a numpy restatement of the benchmark's
definition (as evaluated by the reference's bundled devkit, src/evaluate/evaluate_odometry.cpp:36-116,376-395).
Host-only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pnp_ref  # noqa: E402
from visual_odom_b200 import capi  # noqa: E402

f32 = np.float32


def _restated(gt, est, lengths=(100, 200, 300, 400, 500, 600, 700, 800), step=10):
    n = len(gt)
    dist = [f32(0)]
    for i in range(1, n):
        d = (gt[i - 1][:3, 3] - gt[i][:3, 3]).astype(f32)
        dist.append(f32(dist[-1] + np.sqrt(f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))))
    out = []
    for first in range(0, n, step):
        for L in lengths:
            L = f32(L)
            last = next((i for i in range(first, n) if dist[i] > f32(dist[first] + L)), -1)
            if last < 0:
                continue
            d_gt = np.linalg.inv(gt[first]) @ gt[last]
            d_est = np.linalg.inv(est[first]) @ est[last]
            e = np.linalg.inv(d_est) @ d_gt
            a, b, c = f32(e[0, 0]), f32(e[1, 1]), f32(e[2, 2])
            d = f32(0.5 * (float(f32(f32(a + b) + c)) - 1.0))
            r = np.arccos(np.clip(d, f32(-1), f32(1)))
            t = e[:3, 3].astype(f32)
            te = np.sqrt(f32(f32(t[0] * t[0] + t[1] * t[1]) + t[2] * t[2]))
            out.append((first, f32(r) / L, f32(te) / L, L, f32(float(L) / (0.1 * float(last - first + 1)))))
    return out


def _trajectory(n, seed, drift=0.0):
    rng = np.random.default_rng(seed)
    poses = [np.eye(4)]
    for i in range(1, n):
        rvec = np.array([0.0, 0.004 * np.sin(i / 40.0), 0.0]) + rng.normal(0, 2e-4 + drift, 3)
        t = np.array([0.0, 0.0, 1.1]) + rng.normal(0, 5e-3 + 10 * drift, 3)
        T = np.eye(4); T[:3, :3] = np.asarray(pnp_ref.rodrigues(list(rvec))).reshape(3, 3); T[:3, 3] = t
        poses.append(poses[-1] @ T)
    return poses


def test_segment_errors_match_the_definition(tmp_path):
    gt = _trajectory(1200, 0)
    est = _trajectory(1200, 0, drift=3e-4)          # same seed: gt + extra noise terms drawn differently -> a drifting estimate
    seg, t_avg, r_avg = capi.eval_segments(gt, est)
    ref = _restated(gt, est)
    assert len(seg) == len(ref) > 200
    for s, r in zip(seg, ref):
        assert s["first_frame"] == r[0] and s["len"] == r[3]
        assert abs(s["t_err"] - r[2]) <= 2e-6 * max(1e-3, r[2])
        assert abs(s["r_err"] - r[1]) <= 1e-6 * max(1e-4, r[1]) + 1e-9
        assert abs(s["speed"] - r[4]) <= 1e-5 * r[4]
    assert abs(t_avg - np.mean([r[2] for r in ref], dtype=np.float64)) < 1e-5
    assert abs(r_avg - np.mean([r[1] for r in ref], dtype=np.float64)) < 1e-7
    assert 0 < t_avg < 0.2
    # identical trajectories: zero translation error, rotation error at float round-off
    seg0, t0, r0 = capi.eval_segments(gt, gt)
    assert t0 < 1e-6 and r0 < 1e-5
    # too short for any 100 m segment
    seg1, t1, r1 = capi.eval_segments(gt[:50], est[:50])
    assert len(seg1) == 0 and np.isnan(t1)
    # custom lengths / step
    seg2, _, _ = capi.eval_segments(gt, est, lengths=[5, 10, 50], step=7)
    ref2 = _restated(gt, est, (5, 10, 50), 7)
    assert len(seg2) == len(ref2) and all(s["first_frame"] == r[0] for s, r in zip(seg2, ref2))


def test_pose_file_round_trip(tmp_path):
    gt = _trajectory(40, 2)
    p = tmp_path / "00.txt"
    capi.poses_save(p, gt)
    back = capi.poses_load(p)
    assert back.shape == (40, 12)
    assert np.abs(back - np.array([g[:3].reshape(12) for g in gt])).max() < 1e-8
    # the KITTI text format: 12 numbers per line, parsed like the devkit does (also by numpy)
    assert np.allclose(np.loadtxt(p), back)
    (tmp_path / "trunc.txt").write_text("1 0 0 0 0 1 0 0 0 0 1 0\n1 0 0 0 0 1\n")
    assert capi.poses_load(tmp_path / "trunc.txt").shape == (1, 12)
    import pytest
    with pytest.raises(RuntimeError):
        capi.poses_load(tmp_path / "missing.txt")
