#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics + executed-instruction regions."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, val = r[0], r[2] if len(r) > 2 else r[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__pipe_alu_cycles_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_fmaheavy", 
        "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp",
        "smsp__average_warps_issue_stalled", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct"]
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want):
        print(f"{h} = {val[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2, data = rows[1], rows[2:]
iex, ist, isrc = h2.index("Instructions Executed"), h2.index("# Samples"), h2.index("Source")
stall_cols = [i for i, n in enumerate(h2) if n.startswith("stall_") or n.startswith("Warp Stall")]
tot = sum(int(x[iex]) for x in data); tots = sum(int(x[ist]) for x in data)
print("total warp instructions", tot, "samples", tots)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 150
for b in range(0, len(data), B):
    blk = data[b:b + B]
    ex = sum(int(x[iex]) for x in blk); st = sum(int(x[ist]) for x in blk)
    if ex / tot < 0.004 and st / max(tots, 1) < 0.004:
        continue
    ops = {}
    for x in blk:
        t = x[isrc].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op] = ops.get(op, 0) + int(x[iex])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
    print(f"[{b:4d}-{b+B:4d}] instr {ex/tot*100:5.1f}% samples {st/max(tots,1)*100:5.1f}% ", " ".join(f"{k}:{v/tot*100:.1f}" for k, v in top))
# stall reasons overall
names = [n for n in h2]
agg = {}
for i, n in enumerate(h2):
    if n.startswith("stall_"):
        agg[n] = sum(int(x[i] or 0) for x in data)
if agg:
    s = sum(agg.values())
    print("stall samples:", " ".join(f"{k[6:]}:{v/s*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
