"""Where a kernel-only bench step goes (development tool): pack, Gram, the rest, on one slab of a large matrix.
python tools/step_probe.py [M] [N] [slabs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clusttraj_b200 import _engine
from clusttraj_b200.distmat import slab_bounds
from clusttraj_b200.synth import make_trajectory
M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 8
atoms, coords = make_trajectory(M, N, seed=20250602)
eng = _engine.Engine(0)
eng.load_frames(coords, atoms, _engine.make_opts())
cuts = slab_bounds(M, parts)
for k in (0, parts - 1):
    for rep in range(3):
        pk = eng.pack_resident()
        _, n, st = eng.rmsd_rows_device(cuts[k], cuts[k + 1])
    print(f"slab {k} rows {cuts[k]}..{cuts[k+1]}: pairs {n}  pack {pk:.3f} ms  kernel {st['kernel_ms']:.3f}  gram {st['gram_ms']:.3f}  "
          f"rest {st['kernel_ms'] - st['gram_ms']:.3f}  flagged {st['flagged']}  -> {n / (pk + st['kernel_ms']) * 1e-6:.2f} Gpairs/s", flush=True)
