"""Molecule -> arrays helpers with the reference's names (clusttraj/utils.py:8-31).  They accept any
object that iterates over atoms exposing ``atomicnum`` and ``coords`` (a pybel.Molecule does)."""
import numpy as np


def get_mol_coords(mol):
    return np.asarray([atom.coords for atom in mol])


def get_mol_info(mol):
    return np.array([atom.atomicnum for atom in mol]), np.array([atom.coords for atom in mol])
