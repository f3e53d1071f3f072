"""compat/utils.h: Euler gate + integrateOdometryStereo of the C++ facade against the
restatement in oracle/ref_path.py (reference src/utils.cpp:57-131, src/main.cpp:196-208). Host-only, no GPU."""
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pnp_ref, ref_path  # noqa: E402


def test_pose_integration_matches_reference_semantics(tmp_path):
    from visual_odom_b200 import build
    build.build_native(); build.build_facade()
    exe = os.path.join(ROOT, "tests", "cpp", "utils_main")
    rng = np.random.default_rng(5)
    steps = []
    for i in range(200):
        rvec = rng.normal(0, 0.02, 3)
        t = rng.normal(0, 0.5, 3)
        if i % 17 == 3:
            rvec = rng.normal(0, 0.3, 3)       # large rotation -> Euler gate rejects
        if i % 23 == 5:
            t = t * 1e-3                       # |t| < 0.05 -> skipped with the reference's warning
        if i % 29 == 7:
            t = t * 40                         # |t| > 10 -> skipped
        if i == 50:
            rvec = np.array([0.0, np.pi / 2, 0.0])  # singular branch (sy ~ 0)
        R = np.asarray(pnp_ref.rodrigues(list(rvec)), dtype=np.float64).reshape(3, 3)
        steps.append((R, t))
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<i", len(steps)))
        for R, t in steps:
            f.write(R.astype("<f8").tobytes()); f.write(t.astype("<f8").tobytes())
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    rec = 12 + 4 + 128
    assert len(raw) == rec * len(steps)
    pose = np.eye(4)
    n_gate = n_skip = 0
    for i, (R, t) in enumerate(steps):
        e = np.frombuffer(raw, np.float32, 3, i * rec)
        isrot = struct.unpack_from("<i", raw, i * rec + 12)[0]
        pg = np.frombuffer(raw, np.float64, 16, i * rec + 16).reshape(4, 4)
        e_ref = ref_path.rotation_matrix_to_euler_angles(R)
        assert np.abs(e - e_ref).max() <= 2e-7 * max(1.0, np.abs(e_ref).max())
        assert isrot == 1
        new = ref_path.integrate_pose(pose, R, t)
        if new is pose:
            if np.abs(e_ref).max() >= 0.1:
                n_gate += 1
            else:
                n_skip += 1
        pose = new
        assert np.abs(pg - pose).max() <= 1e-12 * max(1.0, np.abs(pose).max()), i
    assert n_gate > 3 and n_skip > 3
    assert "[WARNING]" in r.stdout


def test_pose_c_abi_helpers():
    """The same semantics through the C-ABI entry points (vo_pose_*), loaded without a GPU."""
    from visual_odom_b200 import capi
    rng = np.random.default_rng(11)
    pose = np.eye(4); ref = np.eye(4)
    adv = 0
    for i in range(300):
        rvec = rng.normal(0, 0.05 if i % 7 else 0.2, 3)
        t = rng.normal(0, 0.6, 3) * (1e-2 if i % 13 == 1 else 1.0)
        R = np.asarray(pnp_ref.rodrigues(list(rvec)), np.float64).reshape(3, 3)
        assert capi.pose_is_rotation(R)
        assert not capi.pose_is_rotation(R * 1.001)
        e = capi.pose_euler(R)
        assert np.abs(e - ref_path.rotation_matrix_to_euler_angles(R)).max() <= 2e-7
        pose, a = capi.pose_step(pose, R, t)
        new = ref_path.integrate_pose(ref, R, t)
        assert a == (new is not ref)
        adv += a
        ref = new
        assert np.abs(pose - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        p2, inv, a2 = capi.pose_integrate(np.eye(4), R, t)
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
        assert np.abs(inv @ T - np.eye(4)).max() < 1e-13
    assert 50 < adv < 300
