"""Timing of the per-pair (stepwise / Hungarian) kernel on a cfg4-shaped sample: python tools/pair_bench.py [M] [solvent molecules]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_solute_solvent

M = int(sys.argv[1]) if len(sys.argv) > 1 else 600
NMOL = int(sys.argv[2]) if len(sys.argv) > 2 else 100
atoms, coords = make_solute_solvent(M, n_solute=60, n_solvent_mol=NMOL, seed=20250604)
eng = _engine.Engine(0)
for name, kw in [("ns60_e", dict(solute_natoms=60, reorder=True)), ("ns60_e_rs", dict(solute_natoms=60, reorder=True, reorder_solvent_only=True)),
                 ("e_only", dict(reorder=True)), ("ns60_noreorder_pairpath", None)]:
    if kw is None:
        eng.load_frames(coords, atoms, _engine.make_opts(solute_natoms=60))
        pairs = np.array([(i, j) for i in range(0, M, 7) for j in range(i + 1, M, 3)], dtype=np.int64)
        t0 = time.perf_counter(); vals, st = eng.pair_values(pairs); dt = time.perf_counter() - t0
        print(f"{name}: {len(pairs)} pairs kernel {st['pair_ms']:.2f} ms -> {len(pairs) / st['pair_ms'] * 1e3:.3e} pairs/s", flush=True)
        continue
    eng.load_frames(coords, atoms, _engine.make_opts(**kw))
    for rep in range(2):
        _, n, st = eng.rmsd_rows_device(0, M)
    print(f"{name}: M={M} pairs {n} kernel {st['pair_ms']:.2f} ms -> {n / st['pair_ms'] * 1e3:.3e} pairs/s, lsap steps/pair {st['lsap_steps'] / n:.1f}", flush=True)
