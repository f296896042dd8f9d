"""Minimal XYZ trajectory reader/writer for the oracle.  TEST INFRASTRUCTURE."""
import numpy as np

_SYMBOLS = (
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce "
    "Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn"
).split()
SYMBOL_TO_Z = {s: i + 1 for i, s in enumerate(_SYMBOLS)}
Z_TO_SYMBOL = {v: k for k, v in SYMBOL_TO_Z.items()}


def read_xyz(path):
    """-> (atoms int[M][N], coords float64[M,N,3])"""
    atoms, coords = [], []
    with open(path) as fh:
        while True:
            head = fh.readline()
            if not head.strip():
                break
            n = int(head.split()[0])
            fh.readline()
            za, xa = [], []
            for _ in range(n):
                t = fh.readline().split()
                za.append(int(t[0]) if t[0].isdigit() else SYMBOL_TO_Z[t[0].capitalize()])
                xa.append((float(t[1]), float(t[2]), float(t[3])))
            atoms.append(za)
            coords.append(xa)
    return np.asarray(atoms, dtype=np.int64), np.asarray(coords, dtype=np.float64)


def write_xyz(path, atoms, coords, decimals=5):
    atoms = np.asarray(atoms)
    with open(path, "w") as fh:
        for m, frame in enumerate(coords):
            za = atoms[m] if atoms.ndim == 2 else atoms
            fh.write(f"{len(frame)}\nframe {m}\n")
            for z, (x, y, zc) in zip(za, frame):
                fh.write(f"{Z_TO_SYMBOL[int(z)]:2s} {x:.{decimals}f} {y:.{decimals}f} {zc:.{decimals}f}\n")
