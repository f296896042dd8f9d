import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from clusttraj_b200 import _engine
from clusttraj_b200.distmat import condensed_matrix
from clusttraj_b200.synth import make_trajectory
n = _engine.device_count()
atoms, coords = make_trajectory(3000, 60, seed=9)
one = condensed_matrix(atoms, coords, _engine.make_opts(), devices=[0]).copy()
allg, st = condensed_matrix(atoms, coords, _engine.make_opts(), devices=list(range(n)), return_stats=True)
print("devices", n, "identical to single-GPU:", bool(np.array_equal(one, allg)), [s["rows"] for s in st])
atoms, coords = make_trajectory(1500, 200, seed=10)
one = condensed_matrix(atoms, coords, _engine.make_opts(no_hydrogen=True), devices=[0]).copy()
allg = condensed_matrix(atoms, coords, _engine.make_opts(no_hydrogen=True), devices=list(range(n)))
print("noh K=100 identical:", bool(np.array_equal(one, allg)))
