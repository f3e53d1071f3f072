"""The `mono_rotation = true` branch (reference src/visualOdometry.cpp:146-157: cv::findEssentialMat(RANSAC, 0.999, 1.0)
+ cv::recoverPose): the numpy restatement (oracle/essential_ref.py) and the kernels' own math compiled for the host
(visual_odom_b200/csrc/ess_math.cuh through libvo_hostcheck.so) against cv2 4.13.0 -- inlier masks identical, rotation to 1e-6."""
import ctypes as C

import numpy as np
import pytest

from visual_odom_b200 import synth

cv2 = pytest.importorskip("cv2")

CASES = [(300, 0.1, 0.1, 0), (500, 0.2, 0.3, 1), (1000, 0.3, 0.5, 2), (200, 0.05, 0.0, 3), (60, 0.2, 0.2, 4), (800, 0.25, 0.6, 6)]


def tracks(n, sigma, outl, seed):
    return synth.essential_stress_set(n, sigma, outl, seed)


def cv2_mono(p0, p1, focal, pp):
    E, mask = cv2.findEssentialMat(p0, p1, focal, pp, cv2.RANSAC, 0.999, 1.0)
    _, R, t, _ = cv2.recoverPose(E, p0, p1, focal=focal, pp=pp, mask=mask.copy())
    return R, mask.ravel().astype(bool)


@pytest.mark.parametrize("n,sigma,outl,seed", CASES)
def test_numpy_restatement_matches_cv2(n, sigma, outl, seed):
    from oracle import essential_ref as er
    p0, p1, focal, pp = tracks(n, sigma, outl, seed)
    R, mask = cv2_mono(p0, p1, focal, pp)
    Ro, mo, iters = er.mono_rotation(p0, p1, focal, pp)
    assert np.array_equal(mo, mask)
    assert np.abs(Ro - R).max() <= 1e-6


@pytest.mark.parametrize("n,sigma,outl,seed", CASES + [(1500, 0.15, 0.3, 5), (2000, 0.1, 0.2, 7)])
def test_kernel_math_on_the_host_matches_cv2(built, n, sigma, outl, seed):
    from visual_odom_b200 import build
    L = C.CDLL(build.build_hostcheck())
    L.vo_hostcheck_mono_rotation.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 5 + [C.c_int] + [C.c_void_p] * 4
    p0, p1, focal, pp = tracks(n, sigma, outl, seed)
    R, mask = cv2_mono(p0, p1, focal, pp)
    E = np.zeros(9); mo = np.zeros(n, np.uint8); Ro = np.zeros(9); it = C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    good = L.vo_hostcheck_mono_rotation(p(p0), p(p1), n, focal, pp[0], pp[1], 0.999, 1.0, 1000, p(E), p(mo), p(Ro), C.byref(it))
    assert good == int(mask.sum()) and np.array_equal(mo.astype(bool), mask)
    assert np.abs(Ro.reshape(3, 3) - R).max() <= 1e-6


def test_five_point_candidates_satisfy_the_constraints(built):
    """every E of a sample is a valid essential matrix through the five correspondences"""
    from visual_odom_b200 import build
    L = C.CDLL(build.build_hostcheck())
    p0, p1, focal, pp = tracks(50, 0.0, 0.0, 11)
    q0 = ((p0.astype(np.float64) - pp) / focal)[:5].copy(); q1 = ((p1.astype(np.float64) - pp) / focal)[:5].copy()
    Es = np.zeros(90)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = L.vo_hostcheck_five_point(p(q0), p(q1), p(Es))
    assert 1 <= n <= 10
    for E in Es[:9 * n].reshape(n, 3, 3):
        E = E / np.linalg.norm(E)
        x0 = np.concatenate([q0, np.ones((5, 1))], 1); x1 = np.concatenate([q1, np.ones((5, 1))], 1)
        assert np.abs(np.sum(x1 * (x0 @ E.T), 1)).max() < 1e-9            # epipolar constraint
        assert abs(np.linalg.det(E)) < 1e-7
        assert np.abs(2 * E @ E.T @ E - np.trace(E @ E.T) * E).max() < 1e-6
