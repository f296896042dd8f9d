"""Summarise an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [kernel-instance]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for i, h in enumerate(hdr):
    if h in keys or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        vals = [r[i] for r in rows[2:]]
        if h.startswith("smsp__average") and all(float(v or 0) < 0.05 for v in vals):
            continue
        print(f"{h} [{units[i]}] = {vals}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    recs.append(r)
tot = sum(int(r[col["# Samples"]] or 0) for r in recs)
byop = collections.Counter(); cnt = collections.Counter()
for r in recs:
    t = r[col["Source"]].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    byop[op] += int(r[col["# Samples"]] or 0); cnt[op] += int(r[col["Instructions Executed"]] or 0)
print(f"-- stall samples by opcode (total {tot}, {len(recs)} SASS instructions)")
for op, s in byop.most_common(18):
    print(f"  {op:24s} {100 * s / tot:5.1f}%  executed {cnt[op]}")
# region split: main loop = between first and last DMMA
idx = [i for i, r in enumerate(recs) if "DMMA" in r[col["Source"]]]
if idx:
    lo, hi2 = idx[0], idx[-1]
    inl = sum(int(r[col["# Samples"]] or 0) for r in recs[lo - 12: hi2 + 2])
    print(f"-- samples inside the DMMA loop body: {100 * inl / tot:.1f}%  (rest: waits, producer, epilogue)")
    for name in ("stall_math", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_branch_resolving", "stall_not_selected", "stall_mio", "stall_dispatch", "stall_selected","stall_no_inst"):
        a = sum(int(r[col[name]] or 0) for r in recs[lo - 12: hi2 + 2]); b = sum(int(r[col[name]] or 0) for r in recs)
        print(f"   {name:24s} loop {100 * a / tot:5.1f}%   all {100 * b / tot:5.1f}%")
