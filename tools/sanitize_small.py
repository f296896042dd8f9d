"""Small end-to-end pass over every kernel family (linkage variants, consumers, per-pair modes, transforms): a quick
sanity run, and the workload to put under compute-sanitizer where that tool is available:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_solute_solvent, make_trajectory

eng = _engine.Engine(0)
atoms, coords = make_trajectory(300, 40, seed=5)
eng.load_frames(coords, atoms, _engine.make_opts())
_, n, st = eng.rmsd_rows_device(0, 300)
d = eng.copy_slab(0, n)
print("gram", st["gram_path"], float(d.sum()))
for variant in ("cluster", "graph"):
    os.environ["CLUSTTRAJ_B200_LINKAGE"] = variant
    for method in ("single", "complete", "average", "weighted", "ward"):
        Z, st = eng.linkage(300, method)
        print(variant, method, st["launches"], float(Z[:, 2].sum()))
total, med, _ = eng.matrix_consumers(300, clusters=np.arange(300) % 7 + 1)
print("consumers", total, med[:3])
atoms, coords = make_solute_solvent(10, n_solute=12, n_solvent_mol=40, seed=77, box=9.0)
pairs = [(i, j) for i in range(10) for j in range(i + 1, 10)]
for kw in (dict(reorder=True, solute_natoms=12), dict(reorder="distance", solute_natoms=12), dict(reorder=True),
           dict(reorder=True, solute_natoms=12, weight_solute=0.6, final_kabsch=True)):
    eng.load_frames(coords, atoms, _engine.make_opts(**kw))
    vals, perms, st = eng.pair_values(pairs, want_perms=True)
    v2, R, T = eng.pair_transforms(pairs)
    print("pairs", kw, float(vals.sum()), st["lsap_steps"])
eng.close()
print("done")
