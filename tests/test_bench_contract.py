"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the driver's contract
(keys, types, the tier's extra objects).  Small workload so it stays in the CPU suite's budget."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytest.importorskip("cv2")


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "0", "--units", "2", "--features", "300", "--width", "640", "--height", "240", "--cpu-seconds", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "frames" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    # both ways of running the CPU path are always reported, each as median / min / max of 3 repetitions
    seq, pool = cb["sequential"], cb["pool"]
    assert seq["reps"] == 3 and seq["min"] <= seq["median"] <= seq["max"] and seq["seconds"] >= 1.9
    assert "error" in pool or (pool["reps"] == 3 and pool["min"] <= pool["median"] <= pool["max"] and pool["busy_workers"] >= 1)
    assert cb["value"] == max(seq["median"], pool.get("median", 0.0))


def test_non_rank0_reference_arm_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_gather_bookkeeping_never_exceeds_the_library_depth():
    """bench.NativeGather (host side of the C-ABI record gather): posts every step, harvests the oldest table only when
    VO_DIST_DEPTH posts are outstanding, drains in order -- checked against a stub context that enforces the library's rule."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from visual_odom_b200.capi import VO_DIST_DEPTH

    class Stub:
        def __init__(self):
            self.posted, self.waited, self.max_out = 0, 0, 0

        def dist_gather_post(self, slot, n):
            assert self.posted - self.waited < VO_DIST_DEPTH, "the library would refuse this post"
            self.posted += 1
            self.max_out = max(self.max_out, self.posted - self.waited)

        def dist_gather_wait(self, n, raw=False):
            assert self.waited < self.posted and raw
            self.waited += 1
            return self.waited - 1

    ctx = Stub()
    g = bench.NativeGather(ctx, 8)
    assert g.DEPTH == VO_DIST_DEPTH
    for s in range(20):
        g.post_step((s % 3) * 8, None, None)
    tables = g.drain()
    assert tables == list(range(20)) and ctx.posted == ctx.waited == 20 and ctx.max_out == VO_DIST_DEPTH
    assert g.drain() == []
