"""Dev: run the int8 Gram kernel a few times on cfg2 (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CLUSTTRAJ_B200_GRAM_VARIANT", "10")
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_trajectory
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
e = _engine.Engine(0)
atoms, coords = make_trajectory(M, N, seed=7)
e.load_frames(coords, atoms, _engine.make_opts())
for rep in range(2):
    _, n, st = e.rmsd_rows_device(0, M)
print(st["gram_ms"], n / st["gram_ms"] / 1e6)
