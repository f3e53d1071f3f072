"""Whole-path parity (A1..A9) through the batched C-ABI on the GPU against
the reference path: cv2 4.13.0 driven through the verbatim glue restatement (oracle/ref_path.py).

Gates (BASELINE.json north_star): tracked-feature indices and RANSAC inlier lists bit-exact;
LK positions bit-exact (stronger than the 1e-4 asked); [R|t] within 1e-4 relative.
"""
import numpy as np
import pytest

from visual_odom_b200 import synth

pytestmark = pytest.mark.gpu


def reference_unit(u, n_select, t_prev, backend="cv2"):
    """The reference's per-frame sequence with the benchmark's stride selection in place of the
    bucketing: FAST -> select -> circularMatching -> checkValidMatch /
    removeInvalidPoints -> triangulate -> trackingFrame2Frame."""
    from oracle import ref_path
    fast, _ = ref_path._backend(backend)
    corners = fast(u["l0"])
    pts = synth.select_features(corners, n_select)
    fs = ref_path.FeatureSet(); fs.points = pts.copy(); fs.ages = np.zeros(len(pts), np.int32)
    cm = ref_path.circular_matching(u["l0"], u["r0"], u["l1"], u["r1"], pts, fs, backend)
    ok = ref_path.check_valid_match(cm["l0"], cm["l0_ret"], 0)
    pL0, pR0, pL1, pR1 = (ref_path.remove_invalid_points(cm[k], ok) for k in ("l0", "r0", "l1", "r1"))
    X = ref_path.triangulate(u["P_l"], u["P_r"], pL0, pR0, backend)
    R, t, inl, rvec = ref_path.tracking_frame2frame(u["P_l"], pL0, pL1, X, t_prev, backend)
    return dict(n_detected=len(corners), pts=pts, kept3=cm["kept_idx"], kept=cm["kept_idx"][ok], l0=pL0, r0=pR0, l1=pL1,
                r1=pR1, X=X, R=R, t=t, inliers=inl)


def check_unit(got_res, got, ref):
    assert got_res["n_detected"] == ref["n_detected"]
    assert np.array_equal(got["pts_in"], ref["pts"])
    assert got_res["n_tracked"] == len(ref["kept3"])
    assert np.array_equal(got["kept_idx"], ref["kept"]), "tracked-feature indices differ"
    for k in ("l0", "r0", "l1", "r1"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["X"], ref["X"])
    assert np.array_equal(got["inliers"], ref["inliers"]), "RANSAC inlier list differs"
    assert np.linalg.norm(got_res["R"] - ref["R"]) / np.linalg.norm(ref["R"]) <= 1e-4
    assert np.linalg.norm(got_res["tvec"] - ref["t"]) / np.linalg.norm(ref["t"]) <= 1e-4


@pytest.mark.parametrize("w,h,n_sel,cal", [(1241, 376, 2000, "kitti"), (1920, 1080, 4000, "zed")])
def test_whole_path_vs_cv2(ctx, w, h, n_sel, cal):
    pytest.importorskip("cv2")
    c = synth.KITTI00 if cal == "kitti" else synth.ZED
    seeds = [0, 1] if cal == "kitti" else [2]
    units = [synth.stereo_unit(w, h, s, cal=c) for s in seeds]
    t_prev = (0.0, 0.0, -0.8)
    ctx.batch_configure(w, h, len(units), units[0]["P_l"], units[0]["P_r"])
    arr, keep, pitch = ctx.make_units([dict(u, n_select=n_sel, t_prev=t_prev) for u in units])
    res = ctx.frame_batch(arr, pitch)
    for i, u in enumerate(units):
        ref = reference_unit(u, n_sel, np.array(t_prev))
        got = ctx.batch_fetch(i, res[i])
        check_unit(res[i], got, ref)
        assert res[i]["n_valid"] > 0.25 * n_sel and res[i]["n_inliers"] > 0.3 * res[i]["n_valid"]
        # the recovered motion is the synthetic ego-motion
        assert np.linalg.norm(res[i]["tvec"] - u["tvec"]) < 0.05


def test_given_features_and_batch_independence(ctx):
    """Results of a unit do not depend on what else is in the batch (needed for sharding)."""
    pytest.importorskip("cv2")
    from oracle import cref
    w, h = 640, 240
    units = [synth.stereo_unit(w, h, s) for s in (3, 4, 5)]
    feats = [synth.select_features(cref.fast_detect(u["l0"])[0], 500) for u in units]
    ctx.batch_configure(w, h, 3, units[0]["P_l"], units[0]["P_r"])
    arr, keep, pitch = ctx.make_units([dict(u, pts=f, t_prev=(0, 0, -0.8)) for u, f in zip(units, feats)])
    res3 = ctx.frame_batch(arr, pitch)
    got3 = [ctx.batch_fetch(i, res3[i]) for i in range(3)]
    for i in range(3):
        ctx.batch_configure(w, h, 1, units[0]["P_l"], units[0]["P_r"])
        a1, k1, p1 = ctx.make_units([dict(units[i], pts=feats[i], t_prev=(0, 0, -0.8))])
        r1 = ctx.frame_batch(a1, p1)[0]
        g1 = ctx.batch_fetch(0, r1)
        assert r1["n_valid"] == res3[i]["n_valid"] and r1["n_inliers"] == res3[i]["n_inliers"]
        for k in ("kept_idx", "l0", "l1", "X", "inliers"):
            assert np.array_equal(g1[k], got3[i][k]), k
        assert np.array_equal(r1["rvec"], res3[i]["rvec"]) and np.array_equal(r1["tvec"], res3[i]["tvec"])


def test_pipelined_submissions_equal_synchronous_batches(ctx):
    """vo_batch_submit / vo_batch_wait: two slot ranges in flight give exactly the records and arrays of
    vo_frame_batch on the same units, in any interleaving, including the re-run of resident slots."""
    w, h, B = 640, 240, 3
    sets = [[synth.stereo_unit(w, h, 10 * k + s) for s in range(B)] for k in range(4)]
    P_l, P_r = sets[0][0]["P_l"], sets[0][0]["P_r"]
    # synchronous references
    ctx.batch_configure(w, h, B, P_l, P_r)
    ref = []
    for us in sets:
        arr, keep, pitch = ctx.make_units([dict(u, n_select=400, t_prev=(0, 0, -0.8)) for u in us])
        res = ctx.frame_batch(arr, pitch)
        ref.append((res, [ctx.batch_fetch(i, res[i]) for i in range(B)]))
    # pipelined: slots [0,B) and [B,2B), submit k+1 before waiting for k
    ctx.batch_configure(w, h, 2 * B, P_l, P_r)
    arrs = [ctx.make_units([dict(u, n_select=400, t_prev=(0, 0, -0.8)) for u in us]) for us in sets]
    ctx.batch_submit(arrs[0][0], 0, arrs[0][2])
    for k in range(len(sets)):
        if k + 1 < len(sets):
            ctx.batch_submit(arrs[k + 1][0], ((k + 1) % 2) * B, arrs[k + 1][2])
        s0 = (k % 2) * B
        res = ctx.batch_wait(s0, B)
        for i in range(B):
            r, rr = res[i], ref[k][0][i]
            for key in ("n_features", "n_detected", "n_tracked", "n_valid", "n_inliers", "ransac_iters"):
                assert r[key] == rr[key], (k, i, key)
            assert np.array_equal(r["rvec"], rr["rvec"]) and np.array_equal(r["tvec"], rr["tvec"])
            g = ctx.batch_fetch(s0 + i, r)
            for key in ("kept_idx", "l0", "r0", "l1", "r1", "X", "inliers"):
                assert np.array_equal(g[key], ref[k][1][i][key]), (k, i, key)
    # re-run of what is resident (no upload): slot range [B, 2B) holds sets[3]
    ctx.batch_submit(None, B, 0, n_units=B)
    again = ctx.batch_wait(B, B)
    assert all(np.array_equal(a["tvec"], r["tvec"]) and a["n_inliers"] == r["n_inliers"] for a, r in zip(again, ref[3][0]))
    # misuse is reported, not undefined
    ctx.batch_submit(arrs[0][0], 0, arrs[0][2])
    with pytest.raises(RuntimeError, match="overlap"):
        ctx.batch_submit(arrs[1][0], 1, arrs[1][2])
    with pytest.raises(RuntimeError, match="no pending"):
        ctx.batch_wait(B, B)
    with pytest.raises(RuntimeError, match="outside"):
        ctx.batch_submit(arrs[1][0], 2 * B - 1, arrs[1][2])
    ctx.batch_wait(0, B)


def test_full_size_properties_static_camera_and_determinism(ctx):
    """BASELINE.json's largest feature count (8000 at 1241x376) through oracle-free properties: with a static camera
    (L1 = L0, R1 = R0) the R0 -> R1 leg of the ring sees identical images, so it returns its input to float round-off;
    the two stereo legs are mutually consistent to a fraction of a pixel; the solved motion is the identity to
    noise level; survivors keep the input order; and a second run of the same batch is bit-identical (no run-to-run
    nondeterminism from atomics or stream interleaving)."""
    w, h, n = 1241, 376, 8000
    us = [synth.stereo_unit(w, h, s) for s in (40, 41)]
    ctx.batch_configure(w, h, 2, us[0]["P_l"], us[0]["P_r"])
    arr, keep, pitch = ctx.make_units([dict(l0=u["l0"], r0=u["r0"], l1=u["l0"], r1=u["r0"], n_select=n, t_prev=(0, 0, 0)) for u in us])
    res = ctx.frame_batch(arr, pitch)
    got = [ctx.batch_fetch(i, res[i]) for i in range(2)]
    for r, g in zip(res, got):
        assert r["n_features"] == n and r["n_valid"] > 0.5 * n
        assert np.abs(g["r1"] - g["r0"]).max() <= 1e-4 * max(1.0, np.abs(g["r0"]).max())      # identical images: zero flow
        assert np.median(np.abs(g["l1"] - g["l0"])) < 0.05 and np.abs(g["l1"] - g["l0"]).max() < 1.5   # A5 keeps < 1 px round trips
        assert np.array_equal(g["l0"], g["pts_in"][g["kept_idx"]])   # survivors are the selected features, in order
        assert np.all(np.diff(g["kept_idx"]) > 0)
        assert r["n_inliers"] > 0.9 * r["n_valid"]
        assert np.linalg.norm(r["R"] - np.eye(3)) < 1e-3 and np.linalg.norm(r["tvec"]) < 0.05
    res2 = ctx.frame_batch(arr, pitch)
    got2 = [ctx.batch_fetch(i, res2[i]) for i in range(2)]
    for a, b, ga, gb in zip(res, res2, got, got2):
        assert a["n_valid"] == b["n_valid"] and a["n_inliers"] == b["n_inliers"]
        assert np.array_equal(a["rvec"], b["rvec"]) and np.array_equal(a["tvec"], b["tvec"])
        for key in ("kept_idx", "l0", "r0", "l1", "r1", "X", "inliers"):
            assert np.array_equal(ga[key], gb[key]), key


def test_record_gather_ring_world_of_one(built):
    """SURVEY.md 8(e), the C-ABI record gather (vo_dist_*), on one GPU as a world of one rank: a posted gather returns the
    records AS THEY WERE WHEN POSTED even though the slots are refilled right away (the device snapshot is what decouples a
    submission from the other ranks), VO_DIST_DEPTH posts may be outstanding, and one more is refused."""
    import torch                                         # makes libnccl.so.2 resident for the library's dlopen
    from visual_odom_b200.capi import Context, VO_DIST_DEPTH
    w, h, B = 640, 240, 2
    c = Context(0, max_features=2048, max_units=2 * B)
    try:
        try:
            uid = c.dist_unique_id()
        except RuntimeError:
            pytest.skip("no loadable NCCL on this host")
        c.dist_init(uid, 0, 1)
        sets = [[synth.stereo_unit(w, h, 7 * k + s) for s in range(B)] for k in range(VO_DIST_DEPTH)]
        c.batch_configure(w, h, 2 * B, sets[0][0]["P_l"], sets[0][0]["P_r"])
        arrs = [c.make_units([dict(u, n_select=300, t_prev=(0, 0, -0.8)) for u in us]) for us in sets]
        expect = []
        for k in range(VO_DIST_DEPTH):                    # every step reuses slot range (k % 2) * B: the gather must not see the refill
            s0 = (k % 2) * B
            c.batch_submit(arrs[k][0], s0, arrs[k][2])
            res = c.batch_wait(s0, B)
            c.dist_gather_post(s0, B)
            expect.append(res)
        with pytest.raises(RuntimeError, match="outstanding"):
            c.dist_gather_post(0, B)
        assert len({tuple(r["tvec"]) for res in expect for r in res}) == VO_DIST_DEPTH * B      # the steps really differ
        for k in range(VO_DIST_DEPTH):
            got = c.dist_gather_wait(B, raw=(k % 2 == 1))         # every other table as one structured array (RESULT_DTYPE)
            if k % 2 == 1:
                assert got.dtype.itemsize == 152 and got["n_valid"].shape == (B,)
                got = c.records_to_dicts(got)
            assert len(got) == B
            for a, b in zip(got, expect[k]):
                assert a["n_inliers"] == b["n_inliers"] and a["n_valid"] == b["n_valid"]
                assert np.array_equal(a["tvec"], b["tvec"]) and np.array_equal(a["R"], b["R"])
        with pytest.raises(RuntimeError, match="nothing outstanding"):
            c.dist_gather_wait(B)
    finally:
        c.close()
