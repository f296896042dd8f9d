"""Where an end-to-end step goes (development tool): ct_load_frames vs ct_rmsd_rows on cfg2 with pinned host buffers."""
import sys, time, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clusttraj_b200 import _engine
from clusttraj_b200.synth import make_trajectory
M, N = 20000, 100
atoms, coords = make_trajectory(M, N, seed=20250602)
eng = _engine.Engine(0)
pin_in = _engine.pinned_empty(coords.size).reshape(coords.shape); pin_in[...] = coords
total = M * (M - 1) // 2
out = _engine.pinned_empty(total)
opts = _engine.make_opts()
for rep in range(5):
    t0 = time.perf_counter(); st1 = eng.load_frames(pin_in, atoms, opts); t1 = time.perf_counter()
    _, st = eng.rmsd_rows(0, M, out); t2 = time.perf_counter()
    print(f"load {1e3*(t1-t0):.2f} ms  rows {1e3*(t2-t1):.2f} ms  (stats total_ms {st['total_ms']:.2f} kernel_ms {st['kernel_ms']:.2f})  -> {8*total/(t2-t1)/1e9:.1f} GB/s in rows", flush=True)
