"""Kernel-only timing of the Gram path for tuning: python tools/quick_bench.py [M] [N] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_trajectory

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
atoms, coords = make_trajectory(M, N, seed=20250602)
for variant in os.environ.get("VARIANTS", "0,1").split(","):
    os.environ["CLUSTTRAJ_B200_GRAM_VARIANT"] = variant
    eng = _engine.Engine(0)
    eng.load_frames(coords, atoms, _engine.make_opts())
    ts = []
    for r in range(reps):
        _, n, st = eng.rmsd_rows_device(0, M)
        ts.append(st["gram_ms"])
    t = min(ts[1:])
    print(f"variant {variant}: M={M} N={N} gram {t:.3f} ms  {n / t * 1e-6:.3f} Gpairs/s  {18.0 * N * n / t * 1e-9:.2f} TFLOP/s  flagged {st['flagged']}", flush=True)
    eng.close()
