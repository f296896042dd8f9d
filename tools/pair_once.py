"""One launch of the per-pair kernel on a cfg4-shaped sample (profiling aid): python tools/pair_once.py [M] [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_solute_solvent

M = int(sys.argv[1]) if len(sys.argv) > 1 else 300
mode = sys.argv[2] if len(sys.argv) > 2 else "ns60_e"
kw = {"ns60_e": dict(solute_natoms=60, reorder=True), "ns60_e_rs": dict(solute_natoms=60, reorder=True, reorder_solvent_only=True),
      "e_only": dict(reorder=True)}[mode]
atoms, coords = make_solute_solvent(M, n_solute=60, n_solvent_mol=100, seed=20250604)
eng = _engine.Engine(0)
eng.load_frames(coords, atoms, _engine.make_opts(**kw))
for rep in range(2):
    _, n, st = eng.rmsd_rows_device(0, M)
print(mode, M, n, st["pair_ms"], n / st["pair_ms"] * 1e3, st["lsap_steps"] / n)
