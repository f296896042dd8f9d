"""Stall samples and executed instructions of an .ncu-rep by SOURCE LINE (development tool):
    python tools/ncu_lines.py rep kernel-substring [lib.so] [--top N]
The CLI's source page lists SASS only; the line table comes from nvdisasm -g of the same cubin (matched by instruction index),
so the library must be the build the capture was taken from.  Inlined code is attributed to the innermost line; the
`file:line <- caller line` chain is kept for the outer function."""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
rest = [a for a in sys.argv[3:] if a.endswith(".so")]
lib = rest[0] if rest else os.path.join(ROOT, "clusttraj_b200", "_libclusttraj_b200.so")
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
recs = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0] != "Address"]
S = [int(r[col["# Samples"]] or 0) for r in recs]
X = [int(r[col["Instructions Executed"]] or 0) for r in recs]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, out, name = None, {}, None
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            out[name] = []; cur = None
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), "inlined" in m.group(3))
            continue
        if name and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out[name].append(cur)
    for nm, tab in out.items():
        if kern in nm and len(tab) == len(recs):
            lines = tab
            break
    if lines:
        break
if not lines:
    sys.exit(f"no kernel matching {kern!r} with {len(recs)} instructions in {lib}")
tot, totx = sum(S), sum(X)
agg = collections.defaultdict(lambda: [0, 0, 0])
for (ln, s, x) in zip(lines, S, X):
    key = (ln[0], ln[1]) if ln else ("?", 0)
    a = agg[key]; a[0] += s; a[1] += x; a[2] += 1
text = {}
def source_line(f, n):
    if f not in text:
        p = os.path.join(ROOT, "clusttraj_b200", "csrc", f)
        text[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return text[f][n - 1].strip()[:100] if 0 < n <= len(text[f]) else ""
print(f"{len(recs)} SASS instructions, {totx} warp instructions, {tot} samples")
print(f"{'file:line':24s} {'samples':>8s} {'exec':>7s} {'sass':>5s}  source")
for (f, n), (s, x, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f + ':' + str(n):24s} {100 * s / tot:7.2f}% {100 * x / totx:6.2f}% {c:5d}  {source_line(f, n)}")
