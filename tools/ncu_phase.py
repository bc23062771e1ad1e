"""Summarise an ncu report of one kernel: headline metrics + stall samples / instruction counts per SASS range
delimited by BAR.SYNC instructions (the kernel's phases)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum"]
for k in want:
    if k in hdr:
        print(f"{k:90s} {vals[hdr.index(k)]} {rows[1][hdr.index(k)]}")
for i, k in enumerate(hdr):
    if "pipe" in k and "pct_of_peak_sustained_active" in k and k not in want:
        try:
            if float(vals[i]) > 5: print(f"{k:90s} {vals[i]}")
        except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = rows[2:]
ia, isamp, iex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
cols = [c for c in h2 if c.startswith("stall_") and "Not Issued" not in c]
tot_s = sum(int(r[isamp]) for r in data); tot_i = sum(int(r[iex]) for r in data)
print("total samples", tot_s, "warp instructions", tot_i)
bounds = [0] + [i + 1 for i, r in enumerate(data) if "BAR.SYNC" in r[ia]] + [len(data)]
for a, b in zip(bounds[:-1], bounds[1:]):
    s = sum(int(r[isamp]) for r in data[a:b]); e = sum(int(r[iex]) for r in data[a:b])
    d = {c: sum(int(r[h2.index(c)]) for r in data[a:b]) for c in cols}
    t = max(1, sum(d.values()))
    top = {k[6:]: round(100 * v / t) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:5] if v > 0.03 * t}
    print(f"[{a:5d},{b:5d}) samples {100*s/tot_s:5.1f}%  inst {100*e/tot_i:5.1f}%  {top}")
if len(sys.argv) > 2:
    thr = float(sys.argv[2])
    for i, r in enumerate(data):
        if int(r[isamp]) > thr * tot_s: print(i, r[isamp], r[iex], r[ia].strip()[:90])
