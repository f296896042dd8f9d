"""Dev check of the int8 tensor-core Gram path against the DMMA path (same engine API), on the GPU box.

    python tools/i8_check.py [quick|full|time]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clusttraj_b200 import _engine  # noqa: E402
from clusttraj_b200.synth import make_trajectory  # noqa: E402


def engine(variant):
    if variant is None:
        os.environ.pop("CLUSTTRAJ_B200_GRAM_VARIANT", None)
    else:
        os.environ["CLUSTTRAJ_B200_GRAM_VARIANT"] = str(variant)
    return _engine.Engine(0)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    e_ref = engine(-2)
    e_i8 = engine(10)
    shapes = [(96, 50, {}), (41, 7, {}), (333, 100, {}), (700, 130, {}), (260, 64, dict(solute_natoms=20)),
              (500, 300, dict(no_hydrogen=True)), (1000, 300, {})]
    if mode == "quick":
        shapes = shapes[:3]
    if mode in ("quick", "full"):
        for M, N, kw in shapes:
            atoms, coords = make_trajectory(M, N, seed=100 + M)
            e_ref.load_frames(coords, atoms, _engine.make_opts(**kw))
            want, _ = e_ref.rmsd_rows(0, M)
            want = want.copy()
            e_i8.load_frames(coords, atoms, _engine.make_opts(**kw))
            got, st = e_i8.rmsd_rows(0, M)
            d = np.abs(got - want)
            k = int(np.argmax(d))
            print(f"M={M} N={N} {kw}: max|d|={d.max():.3e} at {k} (got {got[k]:.9f} want {want[k]:.9f}) "
                  f"mean|d|={d.mean():.3e} flagged={st['flagged']} n={len(got)}", flush=True)
            # row-slab path (chunk cuts on 40-row boundaries)
            if M > 200:
                a, _ = e_i8.rmsd_rows(0, 120)
                b, _ = e_i8.rmsd_rows(120, M)
                both = np.concatenate([a, b])
                print("   slabs identical:", bool(np.array_equal(both, got)))
    if mode in ("time", "full"):
        for M, N in [(20000, 100), (16000, 300)]:
            atoms, coords = make_trajectory(M, N, seed=7)
            for name, e in (("dmma", e_ref), ("int8", e_i8)):
                e.load_frames(coords, atoms, _engine.make_opts())
                for rep in range(3):
                    _, n, st = e.rmsd_rows_device(0, M)
                print(f"{name} M={M} N={N}: gram {st['gram_ms']:.3f} ms  {n / st['gram_ms'] / 1e6:.2f} Gpairs/s  "
                      f"pack {e.pack_resident():.3f} ms flagged={st['flagged']}", flush=True)
            a = e_ref.copy_slab(0, 1000000)
            e_i8.rmsd_rows_device(0, M)
            b = e_i8.copy_slab(0, 1000000)
            print(f"   first 1e6 entries max|d| = {np.abs(a - b).max():.3e}")


if __name__ == "__main__":
    main()
