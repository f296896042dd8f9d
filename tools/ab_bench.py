"""A/B timing of engine-library variants (development tool): python tools/ab_bench.py lib1.so lib2.so ... [--shapes 20000x100,16000x300]
Each library runs in its own process (CLUSTTRAJ_B200_LIB); prints the Gram kernel's best time per shape, the flagged-pair
count and the largest difference from the FIRST library's slab (bit-identical variants print 0)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_trajectory
shapes = sys.argv[1].split(",")
tag = sys.argv[2]
for sh in shapes:
    noh = sh.endswith("n")
    M, N = (int(x) for x in sh.rstrip("n").split("x")[:2])
    atoms, coords = make_trajectory(M, N, seed=20250602)
    eng = _engine.Engine(0)
    eng.load_frames(coords, atoms, _engine.make_opts(no_hydrogen=noh))
    ts = []
    for r in range(6):
        _, n, st = eng.rmsd_rows_device(0, M)
        ts.append(st["gram_ms"])
    t = min(ts[1:])
    rows = eng.copy_slab(0, min(n, 3_000_000))
    ref = "/tmp/ab_ref_%%s.npy" %% sh
    if tag == "0":
        np.save(ref, rows); diff = 0.0
    else:
        diff = float(np.abs(rows - np.load(ref)).max())
    print(f"  {sh}: gram {t:.3f} ms  {n / t * 1e-6:.3f} Gpairs/s  flagged {st['flagged']}  max|d vs first| {diff:.2e}", flush=True)
    eng.close()
''' % ROOT

libs = [a for a in sys.argv[1:] if not a.startswith("--")]
shapes = "20000x100,16000x300"
for k, a in enumerate(sys.argv):
    if a == "--shapes":
        shapes = sys.argv[k + 1]
        libs = [x for x in libs if x != shapes]
for k, spec in enumerate(libs):
    lib, *sets = spec.split("@")      # lib.so@CLUSTTRAJ_B200_PACE=4,0  runs the library with that environment variable set
    print(os.path.basename(lib), *sets, flush=True)
    env = dict(os.environ, CLUSTTRAJ_B200_LIB=os.path.abspath(lib))
    for kv in sets:
        key, val = kv.split("=", 1)
        env[key] = val
    subprocess.run([sys.executable, "-c", CHILD, shapes, str(k)], env=env)
