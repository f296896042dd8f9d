"""Import shim: lets the UNMODIFIED reference module clusttraj/distmat.py be
imported in the authoring container, where the third-party ``rmsd`` package is
not installable.  Re-exports the oracle's restatement.  TEST INFRASTRUCTURE."""
from oracle.rmsd_restated import (  # noqa: F401
    centroid,
    hungarian,
    kabsch,
    kabsch_rmsd,
    kabsch_rotate,
    kabsch_weighted,
    kabsch_weighted_rmsd,
    reorder_hungarian,
    rmsd,
)


def _unsupported(*_a, **_k):
    raise NotImplementedError("out of the north-star scope (SURVEY.md §8a)")


reorder_distance = reorder_brute = reorder_similarity = _unsupported
