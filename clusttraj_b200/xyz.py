"""One-shot trajectory reader.  Replaces the reference's per-pair OpenBabel parse
(clusttraj/distmat.py:56-61,123-131 + utils.get_mol_info, utils.py:20-31) with a single pass that
returns the whole trajectory as ``(atomic_numbers int32[N], coords float64[M, N, 3])``.  Multi-frame XYZ
is read natively; other formats go through OpenBabel when it is installed."""
import os

import numpy as np

SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
           "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
           "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg "
           "Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og").split()
Z_OF = {s.lower(): z for z, s in enumerate(SYMBOLS)}


def _atomic_number(token):
    if token.isdigit():
        return int(token)
    z = Z_OF.get(token.lower())
    if z is None:
        z = Z_OF.get(token[:2].lower()) or Z_OF.get(token[:1].lower())
    if z is None:
        raise ValueError(f"unknown element symbol {token!r}")
    return z


def read_xyz(path):
    """All frames of a multi-frame XYZ file -> (int32[M, N], float64[M, N, 3])."""
    atoms, frames = [], []
    with open(path) as fh:
        lines = fh.read().split("\n")
    pos = 0
    while pos < len(lines):
        head = lines[pos].strip()
        if not head:
            pos += 1
            continue
        n = int(head.split()[0])
        body = lines[pos + 2: pos + 2 + n]
        if len(body) < n:
            raise ValueError(f"{path}: truncated frame at line {pos + 1}")
        z = np.empty(n, dtype=np.int32)
        x = np.empty((n, 3), dtype=np.float64)
        for k, line in enumerate(body):
            t = line.split()
            z[k] = _atomic_number(t[0])
            x[k] = (float(t[1]), float(t[2]), float(t[3]))
        atoms.append(z)
        frames.append(x)
        pos += 2 + n
    if not frames:
        raise ValueError(f"{path}: no frames")
    if any(len(f) != len(frames[0]) for f in frames):
        raise ValueError(f"{path}: frames have different atom counts")
    return np.stack(atoms), np.stack(frames)


def write_xyz(path, atomic_numbers, coords, decimals=5):
    z = np.asarray(atomic_numbers)
    with open(path, "w") as fh:
        for m, frame in enumerate(coords):
            zz = z[m] if z.ndim == 2 else z
            fh.write(f"{len(frame)}\nframe {m}\n")
            for a, (x, y, w) in zip(zz, frame):
                fh.write(f"{SYMBOLS[int(a)]:<2s} {x:.{decimals}f} {y:.{decimals}f} {w:.{decimals}f}\n")


def read_trajectory(path):
    """(atomic numbers int32[M, N], coords float64[M, N, 3]) for any supported trajectory file."""
    ext = os.path.splitext(path)[1][1:].lower()
    if ext == "xyz":
        return read_xyz(path)
    try:
        from openbabel import pybel  # optional dependency, exactly as in the reference
    except Exception as exc:
        raise ImportError(
            f"reading .{ext} trajectories needs OpenBabel (only .xyz is read natively): {exc}") from exc
    atoms, frames = [], []
    for mol in pybel.readfile(ext, path):
        atoms.append(np.array([a.atomicnum for a in mol], dtype=np.int32))
        frames.append(np.array([a.coords for a in mol], dtype=np.float64))
    return np.stack(atoms), np.stack(frames)
