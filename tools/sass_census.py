"""Per-kernel census of the SASS mnemonics that prove which hardware paths the library uses:
    python tools/sass_census.py [lib.so] > profiles/rNN_sass_census.txt
UTCIMMA = tcgen05.mma kind::i8, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UBLKCP = cp.async.bulk (TMA unit),
DMMA = FP64 mma.sync, SYNCS = mbarrier, ELECT = elect.sync, REDUX = warp reductions, DFMA/DADD/DMUL = FP64 pipe."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "clusttraj_b200", "_libclusttraj_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCIMMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "ELECT", "DMMA", "HMMA", "IMMA", "CREDUX", "REDUX",
         "DFMA", "DADD", "DMUL", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "BAR"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or op.startswith(w):
                counts[kern][w] += 1
                break
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); instruction counts per kernel, produced by tools/sass_census.py")
print(f"{'kernel':58s} {'instrs':>7s}  watched mnemonics")
for k, c in counts.items():
    print(f"{k[:58]:58s} {total[k]:7d}  " + " ".join(f"{w}:{c[w]}" for w in WATCH if c[w]))
