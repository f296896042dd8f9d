"""The CLI end to end on a synthetic cfg2-sized trajectory file: python tools/cli_e2e.py [frames] [atoms]
Writes the file, then runs `python -m clusttraj_b200 <file> -rmsd <t> -m average -f` in-process and prints the
wall time of each stage (ingest, matrix, save, linkage + fcluster, consumers + summary)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clusttraj_b200 import _engine, classify, distmat, xyz
from clusttraj_b200.main import main
from clusttraj_b200.synth import make_trajectory

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
atoms, coords = make_trajectory(M, N, seed=20250602)
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "traj.xyz")
t0 = time.perf_counter()
with open(path, "wb") as fh:
    fh.write(_engine.format_xyz_native(atoms, coords, [f"frame {i}" for i in range(M)]))
print(f"wrote {os.path.getsize(path) / 1e6:.1f} MB in {time.perf_counter() - t0:.2f} s", flush=True)
_engine.Engine(0).close()          # CUDA context + library load outside the timed run
stages = {}
def timed(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); stages[name] = stages.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(mod, name, g)
timed(xyz, "read_trajectory"); timed(np, "save"); timed(classify, "linkage")
timed(classify, "sum_distmat"); timed(classify, "find_medoids_from_clusters")
import clusttraj_b200.main as mm
mm.sum_distmat = classify.sum_distmat; mm.find_medoids_from_clusters = classify.find_medoids_from_clusters
distmat.read_trajectory = xyz.read_trajectory
cwd = os.getcwd(); os.chdir(tmp)
t0 = time.perf_counter()
rc = main([path, "-rmsd", "3.0", "-m", "average", "-f"])
total = time.perf_counter() - t0
os.chdir(cwd)
print(f"rc={rc} total {total:.2f} s; stages: " + ", ".join(f"{k} {v:.2f} s" for k, v in stages.items()), flush=True)
print(open(os.path.join(tmp, "clusters.out")).read()[-300:])
