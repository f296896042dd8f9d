"""Placeholder for the SWIG module; the hot path never touches it.  GetSymbol is what clusttraj/io.py:750,861 call
when they print superposed structures.  TEST INFRASTRUCTURE."""
from oracle.xyz_min import SYMBOL_TO_Z

_Z_TO_SYMBOL = {z: s for s, z in SYMBOL_TO_Z.items()}


def GetSymbol(z):
    return _Z_TO_SYMBOL[int(z)]
