"""clusttraj_b200 — B200-native engine for clusttraj's all-pairs minimum-RMSD matrix.

Public surface (same names as the reference package for this path):
    clusttraj_b200.distmat.get_distmat / build_distance_matrix / compute_distmat_line
    clusttraj_b200.io.ClustOptions / configure_runtime / parse_args
    clusttraj_b200.main.main  (python -m clusttraj_b200 ...)
"""
__version__ = "0.1.0"


def main(args=None):
    from .main import main as _main

    return _main(args)
