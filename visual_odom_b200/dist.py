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


class AsyncRecordGather:
    """Non-blocking form of gather_records for a pipelined loop: every post() stages the rank's records in pinned
    memory and enqueues H2D -> all_gather -> D2H on a side stream; nothing waits until the ring of `depth` slots wraps
    or drain() is called.  Tables come back in posting order."""

    def __init__(self, n_units, device="cuda", depth=4):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_units = n_units
        self.per = (n_units + self.world - 1) // self.world
        self.device = device
        cuda = str(device).startswith("cuda")
        self.side = torch.cuda.Stream() if cuda else None
        self.slots = []
        for _ in range(depth):
            h_in = torch.zeros((self.per, RECORD_LEN + 1), dtype=torch.float64, pin_memory=cuda)
            h_out = torch.zeros((self.world * self.per, RECORD_LEN + 1), dtype=torch.float64, pin_memory=cuda)
            self.slots.append(dict(h_in=h_in, h_out=h_out, d_in=torch.zeros_like(h_in, device=device),
                                   d_out=torch.zeros_like(h_out, device=device),
                                   ev=torch.cuda.Event() if cuda else None, busy=False))
        self.k = 0
        self.tables = []

    def _harvest(self, s):
        if s["ev"] is not None:
            s["ev"].synchronize()
        rows = s["h_out"].numpy()
        rows = rows[rows[:, 0] >= 0]
        table = np.zeros((self.n_units, RECORD_LEN), np.float64)
        ids = rows[:, 0].astype(np.int64)
        table[ids] = rows[:, 1:]
        assert len(np.unique(ids)) == self.n_units, "some units were not processed by any rank"
        self.tables.append(table)
        s["busy"] = False

    def post(self, local_records, local_unit_ids):
        torch, dist = self.torch, self.dist
        s = self.slots[self.k % len(self.slots)]
        self.k += 1
        if s["busy"]:
            self._harvest(s)
        h = s["h_in"].numpy()
        h[:, 0] = -1
        n = len(local_unit_ids)
        if n:
            h[:n, 0] = np.asarray(local_unit_ids, np.float64)
            h[:n, 1:] = np.asarray(local_records, np.float64).reshape(n, RECORD_LEN)
        if self.side is not None:
            with torch.cuda.stream(self.side):
                s["d_in"].copy_(s["h_in"], non_blocking=True)
                if self.world > 1:
                    dist.all_gather_into_tensor(s["d_out"], s["d_in"])
                else:
                    s["d_out"].copy_(s["d_in"])
                s["h_out"].copy_(s["d_out"], non_blocking=True)
                s["ev"].record(self.side)
        else:
            if self.world > 1:
                dist.all_gather_into_tensor(s["d_out"], s["d_in"].copy_(s["h_in"]))
            else:
                s["d_out"].copy_(s["h_in"])
            s["h_out"].copy_(s["d_out"])
        s["busy"] = True

    def drain(self):
        n = len(self.slots)
        for i in range(self.k - min(self.k, n), self.k):        # oldest first
            s = self.slots[i % n]
            if s["busy"]:
                self._harvest(s)
        out, self.tables = self.tables, []
        return out
