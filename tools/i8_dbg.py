"""Dev: time the int8 Gram kernel with parts switched off (CLUSTTRAJ_B200_OZ_DBG bits)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CLUSTTRAJ_B200_GRAM_VARIANT"] = "10"
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_trajectory
e = _engine.Engine(0)
for M, N in [(20000, 100), (12000, 300)]:
    atoms, coords = make_trajectory(M, N, seed=7)
    e.load_frames(coords, atoms, _engine.make_opts())
    for dbg in [int(x) for x in (sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "7"])]:
        os.environ["CLUSTTRAJ_B200_OZ_DBG"] = str(dbg)
        for rep in range(2):
            _, n, st = e.rmsd_rows_device(0, M)
        print(f"M={M} N={N} dbg={dbg}: {st['gram_ms']:.3f} ms  {n / st['gram_ms'] / 1e6:.2f} Gpairs/s", flush=True)
