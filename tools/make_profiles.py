"""Turn the scratch ncu output under gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py <round-tag> <gram .ncu-rep> [<pair .ncu-rep>|-] [<launch list csv>|-] [<label>]
    python tools/make_profiles.py <round-tag> --only <.ncu-rep> <name> [pairs per launch]   (just the summary of one capture;
                                                                                              with the pair count also <name>_traffic.json)

<label> (default "gram") names the Gram kernel the capture is of: "gram" = FP64 DMMA kernel, "gram_i8" = int8 tcgen05 kernel.
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
label = sys.argv[5] if len(sys.argv) > 5 else "gram"
out_dir = os.path.join(ROOT, "profiles")


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [{h: (r[i], units[i]) for i, h in enumerate(hdr)} for r in rows[2:]]


KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second"]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def summarise(rep, name):
    text = [f"# ncu --set full summary of {os.path.basename(rep)} ({name}); produced by tools/make_profiles.py", ""]
    ms = raw_metrics(rep)
    for k, m in enumerate(ms):
        text.append(f"## launch {k}")
        for key in KEEP:
            if key in m:
                text.append(f"{key:80s} {m[key][0]} {m[key][1]}")
        for key, (v, u) in sorted(m.items()):
            if key.startswith("smsp__average_warps_issue_stalled") and key.endswith("per_issue_active.ratio") and float(v or 0) >= 0.05:
                text.append(f"{key:80s} {v}")
        text.append("")
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    text.append("## stall samples (tools/ncu_summary.py)")
    text.append(summ)
    path = os.path.join(out_dir, f"{tag}_{name}_ncu_summary.txt")
    open(path, "w").write("\n".join(text))
    print("wrote", path)
    return ms


if sys.argv[2] == "--only":
    ms1 = summarise(sys.argv[3], sys.argv[4])
    if len(sys.argv) > 5:   # pairs per launch of the capture: also write <name>_traffic.json with DRAM bytes per pair
        pairs = int(sys.argv[5])
        m1 = ms1[-1]
        rd, wr = to_bytes(*m1["dram__bytes_read.sum"]), to_bytes(*m1["dram__bytes_write.sum"])
        json.dump({"kernel": m1["Kernel Name"][0], "workload": f"one launch = {pairs} pairs", "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "dram_bytes_per_launch": rd + wr, "pairs_per_launch": pairs, "per_pair": (rd + wr) / pairs,
                   "note": "per_pair scales the capture to other frame counts of the same kernel shape: the 8 B per pair written dominate, "
                           "the operand reads (proportional to the frame count, not the pair count) are the small remainder",
                   "duration_ms_under_ncu": float(m1["gpu__time_duration.sum"][0]), "source": os.path.basename(sys.argv[3])},
                  open(os.path.join(out_dir, f"{tag}_{sys.argv[4]}_traffic.json"), "w"), indent=1)
    sys.exit(0)
ms = summarise(sys.argv[2], label)
m = ms[-1]
traffic = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
json.dump({"kernel": m["Kernel Name"][0], "workload": "cfg2 20000 x 100 plain, one launch = whole matrix",
           "dram_bytes_read": to_bytes(*m["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(*m["dram__bytes_write.sum"]),
           "dram_bytes_per_launch": traffic, "duration_ms_under_ncu": float(m["gpu__time_duration.sum"][0]),
           "source": os.path.basename(sys.argv[2])}, open(os.path.join(out_dir, f"{tag}_{label}_traffic.json"), "w"), indent=1)
if len(sys.argv) > 3 and sys.argv[3] != "-":
    summarise(sys.argv[3], "pair")
if len(sys.argv) > 4 and sys.argv[4] != "-":
    rows = list(csv.reader(open(sys.argv[4])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    lines = ["# per-kernel device time of one `bench.py --steps 2 --warmup 3 --no-cpu-baseline` run under",
             "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)",
             f"{'kernel':70s} {'launches':>8s} {'total_ms':>10s} {'share':>7s}"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name[:70]:70s} {n:8d} {t / 1e6:10.3f} {100 * t / tot:6.1f}%")
    suffix = "" if label == "gram" else "_" + label
    open(os.path.join(out_dir, f"{tag}_launch_shares{suffix}.txt"), "w").write("\n".join(lines) + "\n")
    import shutil
    shutil.copy(sys.argv[4], os.path.join(out_dir, f"{tag}_launches_ncu{suffix}.csv"))
    print("\n".join(lines))
