"""Multi-GPU plumbing: independent work units sharded over ranks.

The path has no data-path collective -- a unit (stereo pair-of-pairs + feature list) never needs
another unit -- so the only communication is the trivial work-queue scatter (broadcast of the unit
table from rank 0) and the gather of fixed-size result records.  torch.distributed is the
plumbing: NCCL on the GPUs, gloo in the CPU tests (tests/test_dist_gloo.py).
"""
import numpy as np

RECORD_LEN = 6 + 3 + 3 + 9      # n_features n_detected n_tracked n_valid n_inliers ransac_iters | rvec | tvec | R


def unit_assignment(n_units, world):
    """Static block-cyclic partition: unit u -> rank u mod world."""
    return [list(range(r, n_units, world)) for r in range(world)]


def result_to_record(res):
    r = np.zeros(RECORD_LEN, np.float64)
    r[:6] = [res["n_features"], res["n_detected"], res["n_tracked"], res["n_valid"], res["n_inliers"], res["ransac_iters"]]
    r[6:9] = res["rvec"]; r[9:12] = res["tvec"]; r[12:21] = np.asarray(res["R"]).ravel()
    return r


def record_to_result(r):
    return dict(n_features=int(r[0]), n_detected=int(r[1]), n_tracked=int(r[2]), n_valid=int(r[3]), n_inliers=int(r[4]),
                ransac_iters=int(r[5]), rvec=r[6:9].copy(), tvec=r[9:12].copy(), R=r[12:21].reshape(3, 3).copy())


def broadcast_unit_table(table, device="cpu"):
    """Rank 0 owns the unit table (int64 array, e.g. seeds / frame ids); everyone gets a copy."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(table, np.int64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=0)
    return t.cpu().numpy()


def gather_records(local_records, local_unit_ids, n_units, device="cpu"):
    """all_gather of the per-unit records; returns the (n_units, RECORD_LEN) table in unit order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    per = (n_units + world - 1) // world
    host = np.zeros((per, RECORD_LEN + 1), np.float64)
    host[:, 0] = -1
    n_local = len(local_unit_ids)
    if n_local:
        host[:n_local, 0] = np.asarray(local_unit_ids, np.float64)
        host[:n_local, 1:] = np.asarray(local_records, np.float64).reshape(n_local, RECORD_LEN)
    buf = torch.from_numpy(host).to(device)
    if world > 1:
        out = torch.empty((world, per, RECORD_LEN + 1), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(out.view(world * per, RECORD_LEN + 1), buf)
        rows = out.view(world * per, RECORD_LEN + 1).cpu().numpy()
    else:
        rows = host
    rows = rows[rows[:, 0] >= 0]
    table = np.zeros((n_units, RECORD_LEN), np.float64)
    ids = rows[:, 0].astype(np.int64)
    table[ids] = rows[:, 1:]
    assert len(np.unique(ids)) == n_units, "some units were not processed by any rank"
    return table
