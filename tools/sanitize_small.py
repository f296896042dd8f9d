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
# per-frame element lists: the relabelling pre-pass of the per-pair kernel
rng = np.random.default_rng(1)
atoms2d = np.tile(atoms, (10, 1))
shuf = coords.copy()
for m in range(1, 10):
    p = np.concatenate([rng.permutation(12), 12 + rng.permutation(len(atoms) - 12)])
    atoms2d[m] = atoms2d[m][p]
    shuf[m] = shuf[m][p]
eng.load_frames(shuf, atoms2d[0], _engine.make_opts(reorder=True, solute_natoms=12), atoms2d)
vals, perms, st = eng.pair_values(pairs, want_perms=True)
print("per-frame lists", float(vals.sum()), st["lsap_steps"])
# a larger solvent shell: the 8-column solver with the collision shortcut, and the K = 300 Gram shape
atoms, coords = make_solute_solvent(8, n_solute=20, n_solvent_mol=70, seed=78, box=12.0)
eng.load_frames(coords, atoms, _engine.make_opts(reorder=True, solute_natoms=20))
_, n, st = eng.rmsd_rows_device(0, 8)
print("shell", n, st["lsap_steps"])
atoms, coords = make_trajectory(200, 300, seed=6)
eng.load_frames(coords, atoms, _engine.make_opts())
_, n, st = eng.rmsd_rows_device(0, 200)
print("gram K=300", st["gram_path"], n)
eng.close()
print("done")
