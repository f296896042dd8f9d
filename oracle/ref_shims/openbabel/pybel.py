"""XYZ-only stand-in for openbabel.pybel: ``readfile(fmt, path)`` yields
molecules whose atoms carry ``atomicnum`` and ``coords`` — all that
clusttraj/utils.py:20-31 reads.  TEST INFRASTRUCTURE."""
from oracle.xyz_min import SYMBOL_TO_Z

informats = {"xyz": "XYZ cartesian coordinates format"}
outformats = {"xyz": "XYZ cartesian coordinates format"}


class Atom:
    __slots__ = ("atomicnum", "coords")

    def __init__(self, z, xyz):
        self.atomicnum = z
        self.coords = xyz


class Molecule:
    def __init__(self, atoms, title=""):
        self.atoms = atoms
        self.title = title

    def __iter__(self):
        return iter(self.atoms)


def readfile(fmt, filename):
    if fmt != "xyz":
        raise ValueError(f"shim reads xyz only, got {fmt}")
    with open(filename) as fh:
        while True:
            head = fh.readline()
            if not head.strip():
                return
            n = int(head.split()[0])
            title = fh.readline().rstrip("\n")
            atoms = []
            for _ in range(n):
                tok = fh.readline().split()
                sym = tok[0]
                z = int(sym) if sym.isdigit() else SYMBOL_TO_Z[sym.capitalize()]
                atoms.append(Atom(z, (float(tok[1]), float(tok[2]), float(tok[3]))))
            yield Molecule(atoms, title)
