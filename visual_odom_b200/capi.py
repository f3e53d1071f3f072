"""ctypes binding of libvo_b200.so (include/vo_b200.h).

This is plumbing only: it loads the in-tree shared library and exposes each C-ABI entry point
with numpy buffers.  There is no Python/CPU implementation behind it -- if the library is
missing or no B200-class GPU is present, the calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvo_b200.so")

VO_OK = 0
VO_E_INVALID = -1
VO_E_CUDA = -2
VO_E_TOO_FEW_POINTS = -3
VO_E_UNSUPPORTED = -4
VO_DIST_DEPTH = 8          # gathers that may be outstanding (include/vo_b200.h)
VO_E_CAPACITY = -5


class VoParams(C.Structure):
    _fields_ = [
        ("fast_threshold", C.c_int), ("fast_nonmax", C.c_int), ("lk_win", C.c_int),
        ("lk_max_level", C.c_int), ("lk_max_iters", C.c_int), ("lk_epsilon", C.c_double),
        ("lk_min_eig", C.c_double), ("circ_threshold", C.c_int), ("pnp_iterations", C.c_int),
        ("pnp_reproj_error", C.c_float), ("pnp_confidence", C.c_double),
        ("max_features", C.c_int), ("max_units", C.c_int),
    ]


class VoUnit(C.Structure):
    _fields_ = [
        ("l0", C.c_void_p), ("r0", C.c_void_p), ("l1", C.c_void_p), ("r1", C.c_void_p),
        ("pts", C.c_void_p), ("n_pts", C.c_int), ("t_prev", C.c_double * 3),
    ]


class VoUnitResult(C.Structure):
    _fields_ = [
        ("n_features", C.c_int), ("n_detected", C.c_int), ("n_tracked", C.c_int), ("n_valid", C.c_int),
        ("n_inliers", C.c_int), ("ransac_iters", C.c_int), ("pnp_status", C.c_int),
        ("rvec", C.c_double * 3), ("tvec", C.c_double * 3), ("R", C.c_double * 9),
    ]


# the same record as a numpy structured dtype (C layout, 152 bytes): arrays of records can be handed out without building
# one Python dict per record (a gathered table of a multi-GPU step has world x units of them)
RESULT_DTYPE = np.dtype([("n_features", "<i4"), ("n_detected", "<i4"), ("n_tracked", "<i4"), ("n_valid", "<i4"), ("n_inliers", "<i4"),
                         ("ransac_iters", "<i4"), ("pnp_status", "<i4"), ("rvec", "<f8", (3,)), ("tvec", "<f8", (3,)),
                         ("R", "<f8", (3, 3))], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(VoUnitResult)


# name -> (restype, argtypes); every symbol include/vo_b200.h declares must be listed here
SIGNATURES = {
    "vo_default_params": (None, [C.POINTER(VoParams)]),
    "vo_create": (C.c_int, [C.c_int, C.POINTER(VoParams), C.POINTER(C.c_void_p)]),
    "vo_destroy": (None, [C.c_void_p]),
    "vo_last_error": (C.c_char_p, [C.c_void_p]),
    "vo_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vo_sync": (C.c_int, [C.c_void_p]),
    "vo_kernel_launches": (C.c_longlong, [C.c_void_p]),
    "vo_lk_kernel_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    "vo_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "vo_fast_detect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                 C.c_int, C.POINTER(C.c_int)]),
    "vo_lk_track": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "vo_circular_match": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int,
                                    C.c_void_p] + [C.c_void_p] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int)]),
    "vo_triangulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vo_triangulate_homogeneous": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vo_pnp_ransac": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]),
    "vo_batch_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vo_batch_upload": (C.c_int, [C.c_void_p, C.POINTER(VoUnit), C.c_int, C.c_size_t]),
    "vo_batch_run": (C.c_int, [C.c_void_p]),
    "vo_batch_download": (C.c_int, [C.c_void_p, C.POINTER(VoUnitResult), C.c_int]),
    "vo_frame_batch": (C.c_int, [C.c_void_p, C.POINTER(VoUnit), C.c_int, C.c_size_t, C.POINTER(VoUnitResult)]),
    "vo_batch_submit": (C.c_int, [C.c_void_p, C.POINTER(VoUnit), C.c_int, C.c_int, C.c_size_t]),
    "vo_batch_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(VoUnitResult)]),
    "vo_seq_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "vo_seq_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(VoUnitResult), C.c_void_p, C.c_int]),
    "vo_seq_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "vo_seq_wait": (C.c_int, [C.c_void_p, C.POINTER(VoUnitResult), C.c_void_p, C.c_int]),
    "vo_seq_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "vo_seq_begin_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "vo_seq_push_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(VoUnitResult), C.c_void_p, C.c_int]),
    "vo_png_info": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vo_png_decode": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "vo_png_last_error": (C.c_char_p, []),
    "vo_reader_open": (C.c_void_p, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vo_reader_next": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vo_reader_error": (C.c_char_p, [C.c_void_p]),
    "vo_reader_close": (None, [C.c_void_p]),
    "vo_bgr_to_gray": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "vo_poses_load": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "vo_poses_save": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int]),
    "vo_eval_segments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "vo_eval_summary": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "vo_pose_is_rotation": (C.c_int, [C.c_void_p]),
    "vo_pose_euler": (None, [C.c_void_p, C.c_void_p]),
    "vo_pose_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vo_pose_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vo_seq_pose": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vo_batch_fetch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vo_mono_rotation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                   C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vo_dist_unique_id": (C.c_int, [C.c_void_p]),
    "vo_dist_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "vo_dist_gather_post": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vo_dist_gather_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "vo_batch_outputs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
}

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m visual_odom_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name, None)
            if fn is None:
                continue            # reported by tests/test_abi.py
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class VoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vo_b200 error {code}: {msg}")
        self.code = code


class Context:
    """Owns one vo_ctx (one per GPU / host thread)."""

    def __init__(self, device=0, **params):
        self.lib = load_library()
        p = VoParams()
        self.lib.vo_default_params(C.byref(p))
        for k, v in params.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown vo_params field {k}")
            setattr(p, k, v)
        self.params = p
        h = C.c_void_p()
        rc = self.lib.vo_create(device, C.byref(p), C.byref(h))
        self.h = h
        if rc != VO_OK:
            msg = self.lib.vo_last_error(h).decode() if h else "vo_create failed"
            if h:
                self.lib.vo_destroy(h)
            self.h = None
            raise VoError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, ok=(VO_OK,)):
        if rc not in ok:
            raise VoError(rc, self.lib.vo_last_error(self.h).decode())
        return rc

    # ---- plumbing --------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.vo_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._check(self.lib.vo_sync(self.h))

    def set_option(self, key, value):
        self._check(self.lib.vo_set_option(self.h, key.encode(), float(value)))

    def kernel_launches(self):
        return int(self.lib.vo_kernel_launches(self.h))

    def lk_kernel_time(self, reset=False):
        ms = C.c_double(); n = C.c_longlong()
        self._check(self.lib.vo_lk_kernel_time(self.h, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value

    # ---- single-call entry points (host buffers) -------------------------------------------------
    @staticmethod
    def _img(a):
        a = np.asarray(a)
        assert a.dtype == np.uint8 and a.ndim == 2, "images must be 2-D uint8 (CV_8UC1)"
        if a.strides[1] != 1:
            a = np.ascontiguousarray(a)
        return a

    def fast_detect(self, img, cap=None, with_response=False):
        img = self._img(img)
        h, w = img.shape
        cap = cap or self.params.max_features * 8
        out = np.zeros((cap, 2), np.float32)
        resp = np.zeros(cap, np.float32) if with_response else None
        n = C.c_int()
        self._check(self.lib.vo_fast_detect(self.h, _p(img), w, h, img.strides[0], _p(out), _p(resp), cap, C.byref(n)),
                    ok=(VO_OK, VO_E_CAPACITY))
        m = min(n.value, cap)
        return (out[:m], resp[:m], n.value) if with_response else (out[:m], n.value)

    def lk_track(self, prev, nxt, pts, want_err=True):
        prev = self._img(prev); nxt = self._img(nxt)
        assert prev.shape == nxt.shape and prev.strides == nxt.strides
        h, w = prev.shape
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        out = np.zeros((n, 2), np.float32); st = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32) if want_err else None
        self._check(self.lib.vo_lk_track(self.h, _p(prev), _p(nxt), w, h, prev.strides[0], _p(pts), n,
                                         _p(out), _p(st), _p(err)))
        return out, st, err

    def circular_match(self, l0, r0, l1, r1, pts, ages=None):
        imgs = [self._img(a) for a in (l0, r0, l1, r1)]
        assert all(a.shape == imgs[0].shape and a.strides == imgs[0].strides for a in imgs)
        h, w = imgs[0].shape
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        outs = [np.zeros((n, 2), np.float32) for _ in range(5)]
        status4 = np.zeros((4, n), np.uint8)
        raw4 = np.zeros((4, n, 2), np.float32)
        kept = np.zeros(n, np.int32)
        nk = C.c_int()
        ages_io = None if ages is None else np.ascontiguousarray(ages, np.int32).copy()
        self._check(self.lib.vo_circular_match(
            self.h, _p(imgs[0]), _p(imgs[1]), _p(imgs[2]), _p(imgs[3]), w, h, imgs[0].strides[0], _p(pts), n,
            _p(ages_io), _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]), _p(outs[4]), _p(status4), _p(raw4),
            _p(kept), C.byref(nk)))
        k = nk.value
        return {
            "l0": outs[0][:k], "r0": outs[1][:k], "l1": outs[2][:k], "r1": outs[3][:k], "l0_ret": outs[4][:k],
            "status4": status4, "raw4": raw4, "kept_idx": kept[:k],
            "ages": None if ages_io is None else ages_io[:k],
        }

    def triangulate(self, P_l, P_r, pts_l, pts_r):
        P_l = np.ascontiguousarray(P_l, np.float32).reshape(12); P_r = np.ascontiguousarray(P_r, np.float32).reshape(12)
        a = np.ascontiguousarray(pts_l, np.float32).reshape(-1, 2); b = np.ascontiguousarray(pts_r, np.float32).reshape(-1, 2)
        assert len(a) == len(b)
        X = np.zeros((len(a), 3), np.float32)
        self._check(self.lib.vo_triangulate(self.h, _p(P_l), _p(P_r), _p(a), _p(b), len(a), _p(X)))
        return X

    def triangulate_homogeneous(self, P_l, P_r, pts_l, pts_r):
        """(n, 4) float32: row i = column i of cv2.triangulatePoints' 4 x N output."""
        P_l = np.ascontiguousarray(P_l, np.float32).reshape(12); P_r = np.ascontiguousarray(P_r, np.float32).reshape(12)
        a = np.ascontiguousarray(pts_l, np.float32).reshape(-1, 2); b = np.ascontiguousarray(pts_r, np.float32).reshape(-1, 2)
        assert len(a) == len(b)
        X4 = np.zeros((len(a), 4), np.float32)
        self._check(self.lib.vo_triangulate_homogeneous(self.h, _p(P_l), _p(P_r), _p(a), _p(b), len(a), _p(X4)))
        return X4

    def mono_rotation(self, pts_t0, pts_t1, focal, pp):
        """findEssentialMat(RANSAC, 0.999, 1.0) + recoverPose: (R 3x3, inlier mask, RANSAC iterations)."""
        a = np.ascontiguousarray(pts_t0, np.float32).reshape(-1, 2); b = np.ascontiguousarray(pts_t1, np.float32).reshape(-1, 2)
        assert len(a) == len(b)
        R = np.zeros(9, np.float64); mask = np.zeros(max(len(a), 1), np.uint8)
        ni = C.c_int(0); it = C.c_int(0)
        self._check(self.lib.vo_mono_rotation(self.h, _p(a), _p(b), len(a), float(focal), float(pp[0]), float(pp[1]), _p(R), _p(mask),
                                              C.byref(ni), C.byref(it)))
        return R.reshape(3, 3), mask[:len(a)].astype(bool), it.value

    def pnp_ransac(self, X, x, K, rvec0=None, tvec0=None):
        X = np.ascontiguousarray(X, np.float32).reshape(-1, 3); x = np.ascontiguousarray(x, np.float32).reshape(-1, 2)
        assert len(X) == len(x)
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        rvec = np.zeros(3) if rvec0 is None else np.ascontiguousarray(rvec0, np.float64).reshape(3).copy()
        tvec = np.zeros(3) if tvec0 is None else np.ascontiguousarray(tvec0, np.float64).reshape(3).copy()
        inl = np.zeros(max(len(X), 1), np.int32); n_in = C.c_int(); R = np.zeros(9); iters = C.c_int()
        rc = self.lib.vo_pnp_ransac(self.h, _p(X), _p(x), len(X), _p(K), _p(rvec), _p(tvec), _p(inl), C.byref(n_in),
                                    _p(R), C.byref(iters))
        self._check(rc)
        return {"rvec": rvec, "tvec": tvec, "R": R.reshape(3, 3), "inliers": inl[:n_in.value], "iters": iters.value}

    # ---- batched whole-path API ------------------------------------------------------------------
    def batch_configure(self, w, h, n_units, P_l, P_r):
        P_l = np.ascontiguousarray(P_l, np.float32).reshape(12); P_r = np.ascontiguousarray(P_r, np.float32).reshape(12)
        self._check(self.lib.vo_batch_configure(self.h, w, h, n_units, _p(P_l), _p(P_r)))
        self._batch_geom = (w, h, n_units)

    def make_units(self, units):
        """units: list of dicts(l0,r0,l1,r1 uint8 HxW [, pts (n,2) f32 | n_select int] [, t_prev]).
        Returns (ctypes array, keep-alive list, pitch)."""
        arr = (VoUnit * len(units))()
        keep = []
        pitch = None
        for i, u in enumerate(units):
            imgs = [self._img(u[k]) for k in ("l0", "r0", "l1", "r1")]
            for a in imgs:
                assert a.shape == (self._batch_geom[1], self._batch_geom[0])
                pitch = pitch or a.strides[0]
                assert a.strides[0] == pitch
            keep.append(imgs)
            arr[i].l0, arr[i].r0, arr[i].l1, arr[i].r1 = (a.ctypes.data for a in imgs)
            if u.get("pts") is not None:
                pts = np.ascontiguousarray(u["pts"], np.float32).reshape(-1, 2)
                keep.append(pts)
                arr[i].pts = pts.ctypes.data
                arr[i].n_pts = len(pts)
            else:
                arr[i].pts = None
                arr[i].n_pts = int(u["n_select"])
            t = u.get("t_prev", (0.0, 0.0, 0.0))
            for k in range(3):
                arr[i].t_prev[k] = float(t[k])
        return arr, keep, pitch

    def batch_upload(self, arr, pitch):
        self._check(self.lib.vo_batch_upload(self.h, arr, len(arr), pitch))

    def batch_run(self):
        self._check(self.lib.vo_batch_run(self.h))

    def batch_download(self, n_units):
        res = (VoUnitResult * n_units)()
        self._check(self.lib.vo_batch_download(self.h, res, n_units))
        return [self._result_dict(r) for r in res]

    def frame_batch(self, arr, pitch):
        res = (VoUnitResult * len(arr))()
        self._check(self.lib.vo_frame_batch(self.h, arr, len(arr), pitch, res))
        return [self._result_dict(r) for r in res]

    def batch_submit(self, arr, first_unit, pitch, n_units=None):
        """Asynchronous upload + run + result staging of resident slots [first_unit, first_unit + len(arr));
        arr=None re-runs what is resident (n_units required)."""
        if arr is None:
            self._check(self.lib.vo_batch_submit(self.h, None, first_unit, n_units, pitch))
        else:
            self._check(self.lib.vo_batch_submit(self.h, arr, first_unit, len(arr), pitch))

    def batch_wait(self, first_unit, n_units, raw=False):
        """raw=True: the records as one numpy structured array (RESULT_DTYPE; fields are read as rec["n_valid"], rec["R"], ...)
        instead of a list of dicts."""
        res = (VoUnitResult * n_units)()
        self._check(self.lib.vo_batch_wait(self.h, first_unit, n_units, res))
        if raw:
            return np.frombuffer(res, dtype=RESULT_DTYPE)
        return [self._result_dict(r) for r in res]

    @staticmethod
    def _result_dict(r):
        return dict(n_features=r.n_features, n_detected=r.n_detected, n_tracked=r.n_tracked, n_valid=r.n_valid,
                    n_inliers=r.n_inliers, ransac_iters=r.ransac_iters, pnp_status=r.pnp_status,
                    rvec=np.array(r.rvec[:]), tvec=np.array(r.tvec[:]), R=np.array(r.R[:]).reshape(3, 3))

    @staticmethod
    def records_to_dicts(arr):
        """A RESULT_DTYPE array (batch_wait / dist_gather_wait with raw=True) as the list of dicts the other calls return."""
        ints = ("n_features", "n_detected", "n_tracked", "n_valid", "n_inliers", "ransac_iters", "pnp_status")
        return [dict({k: int(r[k]) for k in ints}, rvec=np.array(r["rvec"]), tvec=np.array(r["tvec"]), R=np.array(r["R"])) for r in arr]

    def batch_fetch(self, unit, res):
        nf, nv, ni = res["n_features"], res["n_valid"], res["n_inliers"]
        pts_in = np.zeros((max(nf, 1), 2), np.float32); pts4 = np.zeros((4, max(nv, 1), 2), np.float32)
        kept = np.zeros(max(nv, 1), np.int32); X = np.zeros((max(nv, 1), 3), np.float32); inl = np.zeros(max(ni, 1), np.int32)
        self._check(self.lib.vo_batch_fetch(self.h, unit, _p(pts_in), _p(pts4), _p(kept), _p(X), _p(inl)))
        return dict(pts_in=pts_in[:nf], l0=pts4[0, :nv], r0=pts4[1, :nv], l1=pts4[2, :nv], r1=pts4[3, :nv],
                    kept_idx=kept[:nv], X=X[:nv], inliers=inl[:ni])

    def batch_outputs(self, unit, res, into=None):
        """Point lists of a waited submission from the pinned block its single D2H filled (set_option("batch_outputs", 1)).
        `into` = preallocated dict(pts4, kept_idx, X, inliers) to avoid allocations in a timed loop."""
        nv, ni = res["n_valid"], res["n_inliers"]
        if into is None:
            into = dict(pts4=np.zeros((4, max(nv, 1), 2), np.float32), kept_idx=np.zeros(max(nv, 1), np.int32),
                        X=np.zeros((max(nv, 1), 3), np.float32), inliers=np.zeros(max(ni, 1), np.int32))
        nb = C.c_size_t(0)
        self._check(self.lib.vo_batch_outputs(self.h, unit, _p(into["pts4"]), _p(into["kept_idx"]), _p(into["X"]), _p(into["inliers"]),
                                              C.byref(nb)))
        flat = into["pts4"].reshape(-1, 2)            # the C side packs the four lists back to back (n_valid each)
        return dict(l0=flat[0:nv], r0=flat[nv:2 * nv], l1=flat[2 * nv:3 * nv], r1=flat[3 * nv:4 * nv], kept_idx=into["kept_idx"][:nv],
                    X=into["X"][:nv], inliers=into["inliers"][:ni], d2h_bytes=int(nb.value))

    # ---- multi-GPU record gather over NCCL (one process per GPU) -------------------------------------
    def dist_unique_id(self):
        buf = np.zeros(128, np.uint8)
        rc = self.lib.vo_dist_unique_id(_p(buf))
        if rc != VO_OK:
            raise VoError(rc, "vo_dist_unique_id: NCCL is not available on this host")
        return buf

    def dist_init(self, uid, rank, world):
        uid = np.ascontiguousarray(uid, np.uint8)
        assert uid.size == 128
        self._check(self.lib.vo_dist_init(self.h, _p(uid), int(rank), int(world)))
        self._dist_world = int(world)

    def dist_gather_post(self, first_unit, n_units):
        self._check(self.lib.vo_dist_gather_post(self.h, int(first_unit), int(n_units)))

    def dist_gather_wait(self, n_units, raw=False):
        """The oldest posted step's table (world x n_units records, rank-major).  raw=True: one numpy structured array
        (RESULT_DTYPE) instead of world x n_units dicts."""
        res = (VoUnitResult * (self._dist_world * n_units))()
        n = C.c_int(0)
        self._check(self.lib.vo_dist_gather_wait(self.h, res, len(res), C.byref(n)))
        if raw:
            return np.frombuffer(res, dtype=RESULT_DTYPE)[:n.value]
        return [self._result_dict(r) for r in res[:n.value]]

    # ---- streaming sequence mode -------------------------------------------------------------------
    def seq_begin(self, left0, right0, P_l, P_r):
        l = self._img(left0); r = self._img(right0)
        assert l.shape == r.shape and l.strides == r.strides
        P_l = np.ascontiguousarray(P_l, np.float32).reshape(12); P_r = np.ascontiguousarray(P_r, np.float32).reshape(12)
        self._check(self.lib.vo_seq_begin(self.h, l.shape[1], l.shape[0], _p(P_l), _p(P_r), _p(l), _p(r), l.strides[0]))

    def seq_push(self, left1, right1, pts_cap=4096, want_points=True):
        l = self._img(left1); r = self._img(right1)
        res = VoUnitResult()
        if not want_points:
            self._check(self.lib.vo_seq_push(self.h, _p(l), _p(r), l.strides[0], C.byref(res), None, 0))
            return self._result_dict(res)
        pts4 = np.zeros((4, pts_cap, 2), np.float32)
        self._check(self.lib.vo_seq_push(self.h, _p(l), _p(r), l.strides[0], C.byref(res), _p(pts4), pts_cap))
        d = self._result_dict(res)
        n = min(d["n_valid"], pts_cap)
        d.update(l0=pts4[0, :n].copy(), r0=pts4[1, :n].copy(), l1=pts4[2, :n].copy(), r1=pts4[3, :n].copy())
        return d

    def seq_begin_bgr(self, left0, right0, P_l, P_r):
        """Colour (H x W x 3, BGR) inputs: converted on the device like cv::cvtColor(BGR2GRAY)."""
        l = np.ascontiguousarray(left0, np.uint8); r = np.ascontiguousarray(right0, np.uint8)
        assert l.ndim == 3 and l.shape[2] == 3 and r.shape == l.shape
        Pl = np.ascontiguousarray(P_l, np.float32); Pr = np.ascontiguousarray(P_r, np.float32)
        self._check(self.lib.vo_seq_begin_ex(self.h, l.shape[1], l.shape[0], _p(Pl), _p(Pr), _p(l), _p(r), l.strides[0], 3))

    def seq_push_bgr(self, left1, right1, pts_cap=4096):
        l = np.ascontiguousarray(left1, np.uint8); r = np.ascontiguousarray(right1, np.uint8)
        res = VoUnitResult()
        pts4 = np.zeros((4, pts_cap, 2), np.float32)
        self._check(self.lib.vo_seq_push_ex(self.h, _p(l), _p(r), l.strides[0], 3, C.byref(res), _p(pts4), pts_cap))
        d = self._result_dict(res)
        n = min(d["n_valid"], pts_cap)
        d.update(l0=pts4[0, :n].copy(), r0=pts4[1, :n].copy(), l1=pts4[2, :n].copy(), r1=pts4[3, :n].copy())
        return d

    def seq_push_ptr(self, left_ptr, right_ptr, pitch, channels=1):
        """Raw host pointers (e.g. the pinned buffers a SequenceReader hands out); returns counts + pose only."""
        res = VoUnitResult()
        self._check(self.lib.vo_seq_push_ex(self.h, left_ptr, right_ptr, pitch, channels, C.byref(res), None, 0))
        return self._result_dict(res)

    def seq_begin_ptr(self, w, h, left_ptr, right_ptr, pitch, P_l, P_r, channels=1):
        Pl = np.ascontiguousarray(P_l, np.float32); Pr = np.ascontiguousarray(P_r, np.float32)
        self._check(self.lib.vo_seq_begin_ex(self.h, w, h, _p(Pl), _p(Pr), left_ptr, right_ptr, pitch, channels))

    def bgr_to_gray(self, bgr):
        """Stage-level entry point of the device gray conversion (H x W x 3 uint8 -> H x W uint8)."""
        a = np.ascontiguousarray(bgr, np.uint8)
        assert a.ndim == 3 and a.shape[2] == 3
        out = np.empty(a.shape[:2], np.uint8)
        self._check(self.lib.vo_bgr_to_gray(self.h, _p(a), a.strides[0], a.shape[1], a.shape[0], _p(out), out.strides[0]))
        return out

    def seq_submit(self, left1, right1):
        """Asynchronous push (gray H x W or BGR H x W x 3); at most two frames in flight. Keep the arrays alive."""
        l = np.asarray(left1); r = np.asarray(right1)
        ch = 3 if l.ndim == 3 else 1
        assert l.dtype == np.uint8 and l.shape == r.shape and l.strides[-1] == 1 and (ch == 1 or l.strides[1] == 3)
        self._check(self.lib.vo_seq_submit(self.h, _p(l), _p(r), l.strides[0], ch))

    def seq_submit_ptr(self, left_ptr, right_ptr, pitch, channels=1):
        self._check(self.lib.vo_seq_submit(self.h, left_ptr, right_ptr, pitch, channels))

    def seq_wait(self, pts_cap=4096, want_points=True):
        res = VoUnitResult()
        if not want_points:
            self._check(self.lib.vo_seq_wait(self.h, C.byref(res), None, 0))
            return self._result_dict(res)
        pts4 = np.zeros((4, pts_cap, 2), np.float32)
        self._check(self.lib.vo_seq_wait(self.h, C.byref(res), _p(pts4), pts_cap))
        d = self._result_dict(res)
        n = min(d["n_valid"], pts_cap)
        d.update(l0=pts4[0, :n].copy(), r0=pts4[1, :n].copy(), l1=pts4[2, :n].copy(), r1=pts4[3, :n].copy())
        return d

    def seq_state(self, cap=1 << 17):
        pts = np.zeros((cap, 2), np.float32); ages = np.zeros(cap, np.int32); t = np.zeros(3)
        npts = C.c_int(); nages = C.c_int()
        self._check(self.lib.vo_seq_state(self.h, _p(pts), _p(ages), cap, C.byref(npts), C.byref(nages), _p(t)))
        return pts[:npts.value].copy(), ages[:nages.value].copy(), t

    def seq_pose(self):
        """frame_pose (4x4) integrated by seq_push since seq_begin (reference main.cpp:196-208)."""
        pose = np.zeros((4, 4))
        self._check(self.lib.vo_seq_pose(self.h, _p(pose)))
        return pose


# ---- host-only pose bookkeeping (no GPU needed) ------------------------------------------------
def pose_euler(R):
    R = np.ascontiguousarray(R, np.float64); e = np.zeros(3, np.float32)
    load_library().vo_pose_euler(_p(R), _p(e))
    return e


def pose_is_rotation(R):
    R = np.ascontiguousarray(R, np.float64)
    return bool(load_library().vo_pose_is_rotation(_p(R)))


def pose_integrate(frame_pose, R, t):
    """integrateOdometryStereo: returns (new_pose, rigid_inv, advanced)."""
    pose = np.array(frame_pose, np.float64).reshape(4, 4).copy(); inv = np.zeros((4, 4))
    R = np.ascontiguousarray(R, np.float64); t = np.ascontiguousarray(t, np.float64).reshape(3)
    rc = load_library().vo_pose_integrate(_p(pose), _p(R), _p(t), _p(inv))
    if rc < 0:
        raise RuntimeError("vo_pose_integrate: singular transformation")
    return pose, inv, bool(rc)


def pose_step(frame_pose, R, t):
    """Euler gate + integration (main.cpp:196-208): returns (new_pose, advanced)."""
    pose = np.array(frame_pose, np.float64).reshape(4, 4).copy()
    R = np.ascontiguousarray(R, np.float64); t = np.ascontiguousarray(t, np.float64).reshape(3)
    rc = load_library().vo_pose_step(_p(pose), _p(R), _p(t))
    if rc < 0:
        raise RuntimeError("vo_pose_step: singular transformation")
    return pose, bool(rc)


# ---- image ingest: host-side PNG decode + prefetching KITTI-layout reader ----------------------
def png_info(data):
    buf = np.frombuffer(data, np.uint8)
    w = C.c_int(); h = C.c_int(); ct = C.c_int(); bd = C.c_int()
    lib = load_library()
    if lib.vo_png_info(_p(buf), buf.size, C.byref(w), C.byref(h), C.byref(ct), C.byref(bd)) != 0:
        raise RuntimeError(lib.vo_png_last_error().decode())
    return w.value, h.value, ct.value, bd.value


def png_decode(data, want_bgr=True, want_gray=True):
    """bytes of one PNG -> (bgr HxWx3 | None, gray HxW | None): imread(IMREAD_COLOR) and cvtColor(BGR2GRAY) of it."""
    buf = np.frombuffer(data, np.uint8)
    w, h, _, _ = png_info(data)
    bgr = np.empty((h, w, 3), np.uint8) if want_bgr else None
    gray = np.empty((h, w), np.uint8) if want_gray else None
    lib = load_library()
    rc = lib.vo_png_decode(_p(buf), buf.size, _p(bgr) if want_bgr else None, 3 * w, _p(gray) if want_gray else None, w)
    if rc != 0:
        raise RuntimeError(lib.vo_png_last_error().decode())
    return bgr, gray


class SequenceReader:
    """<dir>/image_0/%06d.png + <dir>/image_1/%06d.png decoded ahead on worker threads into pinned buffers."""

    def __init__(self, sequence_dir, first_frame, n_frames, threads=4, depth=4, force_channels=0):
        self.lib = load_library()
        self.h = self.lib.vo_reader_open(str(sequence_dir).encode(), first_frame, n_frames, threads, depth, force_channels)
        if not self.h:
            raise RuntimeError("vo_reader_open: " + self.lib.vo_png_last_error().decode())
        self.n_frames = n_frames

    def next_ptr(self):
        """(left_ptr, right_ptr, w, h, pitch, channels, frame_id); pointers valid until the second following call."""
        l = C.c_void_p(); r = C.c_void_p(); w = C.c_int(); h = C.c_int(); p = C.c_size_t(); ch = C.c_int(); fid = C.c_int()
        rc = self.lib.vo_reader_next(self.h, C.byref(l), C.byref(r), C.byref(w), C.byref(h), C.byref(p), C.byref(ch), C.byref(fid))
        if rc != 0:
            raise RuntimeError("vo_reader_next: " + self.lib.vo_reader_error(self.h).decode())
        return l.value, r.value, w.value, h.value, p.value, ch.value, fid.value

    def next(self):
        """Copies of the frame's two images as numpy arrays (H x W or H x W x 3) + frame id."""
        l, r, w, h, p, ch, fid = self.next_ptr()
        shape = (h, w) if ch == 1 else (h, w, 3)
        n = h * p
        la = np.ctypeslib.as_array((C.c_uint8 * n).from_address(l)).reshape(shape).copy()
        ra = np.ctypeslib.as_array((C.c_uint8 * n).from_address(r)).reshape(shape).copy()
        return la, ra, fid

    def close(self):
        if self.h:
            self.lib.vo_reader_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- KITTI accuracy evaluation (host-only) -------------------------------------------------------
SEGMENT_DTYPE = np.dtype([("first_frame", np.int32), ("r_err", np.float32), ("t_err", np.float32), ("len", np.float32),
                          ("speed", np.float32)])


def poses_load(path):
    lib = load_library()
    n = C.c_int()
    if lib.vo_poses_load(str(path).encode(), None, 0, C.byref(n)) != 0:
        raise RuntimeError(f"cannot read poses from {path}")
    out = np.zeros((n.value, 12))
    if n.value:
        lib.vo_poses_load(str(path).encode(), _p(out), n.value, C.byref(n))
    return out


def poses_save(path, poses):
    p = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(len(poses), -1)[:, :12])
    if load_library().vo_poses_save(str(path).encode(), _p(p), len(p)) != 0:
        raise RuntimeError(f"cannot write poses to {path}")


def eval_segments(gt, est, lengths=None, step=10):
    """KITTI segment errors of `est` against `gt` (n x 12 or n x 4 x 4 poses) -> (segments, t_err_avg, r_err_avg)."""
    def rows(a):
        a = np.asarray(a, np.float64)
        return np.ascontiguousarray(a.reshape(len(a), -1)[:, :12])
    g, e = rows(gt), rows(est)
    assert g.shape == e.shape
    lib = load_library()
    L = np.ascontiguousarray(lengths, np.float32) if lengths is not None else None
    n = C.c_int()
    args = (_p(g), _p(e), len(g), _p(L) if L is not None else None, len(L) if L is not None else 0, step)
    if lib.vo_eval_segments(*args, None, 0, C.byref(n)) != 0:
        raise RuntimeError("vo_eval_segments failed (singular pose?)")
    seg = np.zeros(n.value, SEGMENT_DTYPE)
    if n.value == 0:
        return seg, float("nan"), float("nan")
    lib.vo_eval_segments(*args, _p(seg), n.value, C.byref(n))
    t = C.c_float(); r = C.c_float()
    lib.vo_eval_summary(_p(seg), n.value, C.byref(t), C.byref(r))
    return seg, t.value, r.value
