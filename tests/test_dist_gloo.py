"""world_size-2 gloo test of the sharding plumbing on CPU: the unit table is broadcast from rank 0,
units are processed block-cyclically, result records are gathered, and the assembled table is
identical to a single-process run (results must not depend on the shard count)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _process_unit(seed):
    """Stand-in for the GPU path in this CPU test: the oracle ring on a tiny unit."""
    from oracle import cref, ref_path
    from visual_odom_b200 import synth
    u = synth.stereo_unit(200, 120, int(seed), scene="v0")
    pts = synth.select_features(cref.fast_detect(u["l0"])[0], 40)
    fs = ref_path.FeatureSet(); fs.points = pts; fs.ages = np.zeros(len(pts), np.int32)
    cm = ref_path.circular_matching(u["l0"], u["r0"], u["l1"], u["r1"], pts, fs, backend="c")
    ok = ref_path.check_valid_match(cm["l0"], cm["l0_ret"], 0)
    s = cm["l1"][ok].astype(np.float64).sum(axis=0) if ok.any() else np.zeros(2)
    return dict(n_features=len(pts), n_detected=0, n_tracked=len(cm["kept_idx"]), n_valid=int(ok.sum()), n_inliers=0,
                ransac_iters=0, rvec=np.array([s[0], s[1], 0.0]), tvec=np.zeros(3), R=np.eye(3))


def _worker(rank, world, port, n_units, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from visual_odom_b200 import dist as vd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = np.arange(100, 100 + n_units) if rank == 0 else np.zeros(n_units, np.int64)
    table = vd.broadcast_unit_table(table)
    mine = vd.unit_assignment(n_units, world)[rank]
    recs = [vd.result_to_record(_process_unit(table[u])) for u in mine]
    full = vd.gather_records(recs, mine, n_units)
    # the non-blocking gather of a pipelined loop: three posts (more than... fewer than the ring depth), then a ring wrap
    ag = vd.AsyncRecordGather(n_units, device="cpu", depth=2)
    for k in range(3):
        ag.post([r + k for r in recs], mine)
    tabs = ag.drain()
    assert len(tabs) == 3 and all(np.array_equal(t, full + k) for k, t in enumerate(tabs))
    if rank == 0:
        np.save(out_path, full)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(built, tmp_path):
    import torch.multiprocessing as mp
    from visual_odom_b200 import dist as vd
    n_units = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, n_units, out), nprocs=2, join=True)
    got = np.load(out)
    ref = np.stack([vd.result_to_record(_process_unit(100 + u)) for u in range(n_units)])
    assert np.array_equal(got, ref)
    assert vd.unit_assignment(5, 2) == [[0, 2, 4], [1, 3]]
    r = vd.record_to_result(got[0])
    assert r["n_features"] == int(ref[0, 0]) and np.array_equal(r["rvec"], ref[0, 6:9])
