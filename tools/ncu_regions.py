"""Per-region view of an .ncu-rep source page (development tool): python tools/ncu_regions.py rep [--list lo hi] [--top N]
Groups consecutive SASS instructions with similar execution counts into regions (loop bodies) and prints each region's
share of the executed warp instructions and of the stall samples; --list prints the instructions lo..hi with their samples;
--top the instructions with the most samples."""
import csv, io, subprocess, sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
recs = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0] != "Address"]
S = [int(r[col["# Samples"]] or 0) for r in recs]
X = [int(r[col["Instructions Executed"]] or 0) for r in recs]
tot, totx = sum(S), sum(X)
print(f"{len(recs)} SASS instructions, {totx} warp instructions executed, {tot} samples")
if "--list" in sys.argv:
    k = sys.argv.index("--list"); lo, hi2 = int(sys.argv[k + 1]), int(sys.argv[k + 2])
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for i in range(lo, min(hi2, len(recs))):
        r = recs[i]
        top = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        tops = " ".join(f"{n}:{v}" for v, n in top if v)
        print(f"{i:5d} {100 * S[i] / tot:5.2f} {X[i]:10d} {r[col['L1 Wavefronts Shared Excessive']]:>8s} {r[col['Source']][:88]:88s} {tops}")
    sys.exit()
if "--top" in sys.argv:
    n = int(sys.argv[sys.argv.index("--top") + 1])
    for i in sorted(range(len(recs)), key=lambda i: -S[i])[:n]:
        print(f"{i:5d} {100 * S[i] / tot:5.2f} {X[i]:10d} {recs[i][col['Source']][:90]}")
    sys.exit()
reg, cur = [], None
for i in range(len(recs)):
    x = X[i]
    if cur and 0.7 * cur["x"] <= x <= 1.4 * cur["x"]:
        cur["n"] += 1; cur["s"] += S[i]; cur["tx"] += x; cur["end"] = i
    else:
        cur = {"start": i, "end": i, "x": max(x, 1), "n": 1, "s": S[i], "tx": x}; reg.append(cur)
for c in reg:
    if c["tx"] / totx > 0.004 or c["s"] / tot > 0.005:
        print(f"{c['start']:5d}-{c['end']:5d} n={c['n']:4d} exec/instr~{c['x']:10d} exec {100 * c['tx'] / totx:5.1f}%  samples {100 * c['s'] / tot:5.1f}%")
