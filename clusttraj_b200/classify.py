"""Drop-in for the two consumers of the RMSD matrix that clusttraj runs right after the clustering
(clusttraj/classify.py:146-177), on the GPU and without the M x M ``squareform`` the reference builds on the
host (3.2 GB at 20k frames, 80 GB at 100k):

  find_medoids_from_clusters(distmat, clusters)  classify.py:146-165
  sum_distmat(distmat)                           classify.py:168-177

Both take the condensed float64 vector ``get_distmat`` returned (it is copied to the GPU; one pass over its 8 L
bytes).  Like the rest of the package there is no CPU fallback: without a GPU they raise EngineError.  Ties between
candidate medoids go to the lowest frame index, as ``np.argmin`` does in the reference; two candidates whose sums
agree to the last bit of a differently ordered floating-point sum can resolve differently (documented in DESIGN.md).
"""
import numpy as np

from .distmat import _engine_for, visible_devices


def _n_frames(distmat):
    L = int(np.asarray(distmat).shape[0])
    M = int(round((1.0 + np.sqrt(1.0 + 8.0 * L)) / 2.0))
    if M * (M - 1) // 2 != L:
        raise ValueError("distmat is not a condensed distance vector")   # what squareform raises in the reference
    return M


def find_medoids_from_clusters(distmat, clusters):
    """Indices of the medoids of the clusters (classify.py:146-165): for cluster c = 1..n the member with the
    smallest sum of distances to the members of c."""
    clusters = np.asarray(clusters)
    M = _n_frames(distmat)
    if clusters.shape != (M,):
        raise ValueError("clusters must hold one label per frame")
    eng = _engine_for(visible_devices()[0])
    _, medoids, _ = eng.matrix_consumers(M, clusters=clusters, distmat=distmat)
    return medoids.astype(int)


def sum_distmat(distmat):
    """Sum of the RMSD matrix (classify.py:168-177)."""
    M = _n_frames(distmat)
    eng = _engine_for(visible_devices()[0])
    total, _, _ = eng.matrix_consumers(M, clusters=None, distmat=distmat)
    return np.float64(total)
