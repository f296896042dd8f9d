"""GPU linkage vs SciPy on the GPU box's host: python tools/linkage_bench.py [M] [methods...] [--no-cpu]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_trajectory

args = [a for a in sys.argv[1:] if not a.startswith("--")]
M = int(args[0]) if args else 20000
methods = args[1:] or ["average", "ward", "complete", "weighted", "single"]
atoms, coords = make_trajectory(M, 30, seed=46, n_modes=12)
eng = _engine.Engine(0)
eng.load_frames(coords, atoms, _engine.make_opts())
_, n, st = eng.rmsd_rows_device(0, M)
d = None
for method in methods:
    Z, st = eng.linkage(M, method)
    Z, st = eng.linkage(M, method)
    line = f"{os.environ.get('CLUSTTRAJ_B200_LINKAGE', 'cluster'):7s} {method:9s} M={M} gpu {st['kernel_ms']:9.1f} ms  steps {st['launches']}  ({st['kernel_ms'] * 1e3 / st['launches']:.2f} us/step)"
    if "--no-cpu" not in sys.argv:
        import scipy.cluster.hierarchy as hcl
        if d is None:
            d = eng.copy_slab(0, n)
        t0 = time.perf_counter(); want = hcl.linkage(d, method); cpu = time.perf_counter() - t0
        line += f"  scipy {cpu * 1e3:9.1f} ms  x{cpu * 1e3 / st['kernel_ms']:.1f}  identical={Z.tobytes() == want.tobytes()}"
    print(line, flush=True)
