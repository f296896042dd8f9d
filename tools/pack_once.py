"""K1 (centre + quantise + pack) alone at a given size (profiling aid): python tools/pack_once.py [M] [N] [-n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_trajectory
M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
noh = "-n" in sys.argv
atoms, coords = make_trajectory(M, N, seed=20250602)
eng = _engine.Engine(0)
eng.load_frames(coords, atoms, _engine.make_opts(no_hydrogen=noh))
ts = [eng.pack_resident() for _ in range(5)]
print(f"M={M} N={N} noh={noh}: pack_resident {min(ts):.3f} ms (all {['%.3f' % t for t in ts]})")
