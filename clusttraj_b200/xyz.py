"""One-shot trajectory reader.  Replaces the reference's per-pair OpenBabel parse
(clusttraj/distmat.py:56-61,123-131 + utils.get_mol_info, utils.py:20-31) with a single pass that
returns the whole trajectory as ``(atomic_numbers int32[M, N], coords float64[M, N, 3])``.  Multi-frame XYZ,
multi-MODEL PDB and multi-frame GRO are read natively (GRO coordinates are converted from nm to Angstrom, as
OpenBabel does); other formats go through OpenBabel when it is installed."""
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


def _read_xyz_slow(lines, path):
    """Frame-by-frame parse: the general case (frames may differ in their comment lines or be separated by blanks)."""
    atoms, frames = [], []
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


def read_xyz(path):
    """All frames of a multi-frame XYZ file -> (int32[M, N], float64[M, N, 3]).  Regular files (every frame N + 2
    lines) go through the engine library's parallel native parser (ct_read_xyz: all host cores, ~30x numpy's reader;
    host-only code, no GPU involved); when the library cannot be loaded they are parsed in bulk by numpy's C reader;
    anything irregular falls back to the general line loop, which also words the errors."""
    try:
        from . import _engine
        native = _engine.read_xyz_native(path)
    except Exception:   # library missing / cannot be built (EngineError), dlopen failure: the readers below need no library
        native = None
    if native is not None:
        return native
    with open(path) as fh:
        lines = fh.read().split("\n")
    nl = len(lines)
    while nl and not lines[nl - 1].strip():
        nl -= 1
    try:
        n = int(lines[0].split()[0])
        stride = n + 2
        if n < 1 or nl % stride:
            raise ValueError
        M = nl // stride
        if any(lines[f * stride].split()[0] != lines[0].split()[0] for f in range(M)):
            raise ValueError
        body = [ln for f in range(M) for ln in lines[f * stride + 2: (f + 1) * stride]]
        coords = np.loadtxt(body, usecols=(1, 2, 3), dtype=np.float64, ndmin=2).reshape(M, n, 3)
        syms = np.loadtxt(body, usecols=0, dtype="U8", ndmin=1).reshape(M, n)
    except (ValueError, IndexError):
        return _read_xyz_slow(lines[:nl], path)
    table = {t: _atomic_number(t) for t in np.unique(syms)}
    atoms = np.vectorize(table.get, otypes=[np.int32])(syms)
    return atoms, coords


def _native(path, fmt):
    try:
        from . import _engine
        return _engine.read_xyz_native(path, fmt=fmt)
    except (ImportError, OSError, AttributeError):
        return None
    except Exception:   # EngineError when the library cannot be built/loaded: the Python readers below need no library
        return None


def pdb_element_from_name(field):
    """Element symbol for an ATOM/HETATM record WITHOUT element columns, from the atom-name field (columns 13-16) by the
    alignment convention of the PDB format (wwPDB format guide, ATOM record: "Alignment of one-letter atom name such as C
    starts at column 14, while two-letter atom name such as FE starts at column 13"):
      * column 13 blank or a digit -> the first letter that follows is a one-letter element (" CA " alpha carbon -> C,
        "1HG2" -> H, " OW " -> O);
      * column 13 a letter -> the two-letter symbol in columns 13-14 when that is an element ("FE  ", "CL  ", "CA  " =
        calcium, "ZN  "), except the four-character hydrogen names that have to start in column 13 ("HG12", "HD21": H);
        otherwise the letter in column 13 ("C12 " from a writer that left-justifies -> C)."""
    name = (field + "    ")[:4]
    if name[0].isalpha():
        two = name[:2]
        if name[1].isalpha() and two.lower() in Z_OF and not (name[0] in "Hh" and " " not in name):
            return two
        return name[0]
    for ch in name[1:]:
        if ch.isalpha():
            return ch
    return name.strip()[:1]


def read_pdb(path, native=True):
    """MODEL/ENDMDL frames of a PDB file (or a single frame without MODEL records): ATOM/HETATM records, element
    from columns 77-78 when present, else from the atom name (columns 13-16) by the format's alignment convention
    (pdb_element_from_name).  Parsed by the
    engine library's parallel reader (ct_read_pdb) when it accepts the file, else by the line loop below."""
    got = _native(path, "pdb") if native else None
    if got is not None:
        return got
    atoms, frames = [], []
    z, x = [], []

    def flush():
        if z:
            atoms.append(np.array(z, dtype=np.int32))
            frames.append(np.array(x, dtype=np.float64))
            z.clear(); x.clear()

    with open(path) as fh:
        for line in fh:
            rec = line[:6]
            if rec in ("ATOM  ", "HETATM"):
                el = line[76:78].strip() or pdb_element_from_name(line[12:16])
                z.append(_atomic_number(el))
                x.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
            elif rec.startswith("ENDMDL") or rec.startswith("END"):
                flush()
    flush()
    if not frames:
        raise ValueError(f"{path}: no ATOM/HETATM records")
    if any(len(f) != len(frames[0]) for f in frames):
        raise ValueError(f"{path}: frames have different atom counts")
    return np.stack(atoms), np.stack(frames)


def read_gro(path, native=True):
    """Frames of a GROMACS .gro file: title, atom count, fixed-column atom lines (positions in nm), box line.
    The element is guessed from the atom name's leading letters.  Parsed by the engine library's parallel reader
    (ct_read_gro) when it accepts the file, else by the line loop below."""
    got = _native(path, "gro") if native else None
    if got is not None:
        return got
    atoms, frames = [], []
    with open(path) as fh:
        lines = fh.read().split("\n")
    pos = 0
    while pos + 1 < len(lines):
        if not lines[pos].strip() and not lines[pos + 1].strip():
            pos += 1
            continue
        n = int(lines[pos + 1].split()[0])
        body = lines[pos + 2: pos + 2 + n]
        if len(body) < n:
            raise ValueError(f"{path}: truncated frame at line {pos + 1}")
        z = np.empty(n, dtype=np.int32)
        x = np.empty((n, 3), dtype=np.float64)
        for k, line in enumerate(body):
            name = "".join(ch for ch in line[10:15] if ch.isalpha())
            el = name[:2] if name[:2].lower() in Z_OF and len(name) > 1 and name[1].islower() else name[:1]
            z[k] = _atomic_number(el)
            x[k] = (10.0 * float(line[20:28]), 10.0 * float(line[28:36]), 10.0 * float(line[36:44]))
        atoms.append(z)
        frames.append(x)
        pos += 3 + n
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
    if ext in ("pdb", "ent"):
        return read_pdb(path)
    if ext == "gro":
        return read_gro(path)
    try:
        from openbabel import pybel  # optional dependency, exactly as in the reference
    except Exception as exc:
        raise ImportError(
            f"reading .{ext} trajectories needs OpenBabel (.xyz, .pdb and .gro are read natively): {exc}") from exc
    atoms, frames = [], []
    for mol in pybel.readfile(ext, path):
        atoms.append(np.array([a.atomicnum for a in mol], dtype=np.int32))
        frames.append(np.array([a.coords for a in mol], dtype=np.float64))
    return np.stack(atoms), np.stack(frames)
