"""CPU restatement of the two matrix consumers of clusttraj/classify.py.  TEST INFRASTRUCTURE ONLY: imported by
tests/ and never by the product.

  find_medoids_from_clusters  follows classify.py:146-165 (squareform, per-cluster submatrix, argmin of the column sums)
  sum_distmat                 follows classify.py:168-177 (np.sum)

Pinned by tests/golden/reference_consumers.json, produced by the UNMODIFIED reference functions imported over the
shims of oracle/ref_shims (oracle/make_golden.py); the reference's own test-suite has no test of either function.
"""
import numpy as np
from scipy.spatial.distance import squareform


def find_medoids_from_clusters(distmat, clusters):
    clusters = np.asarray(clusters)
    n_clusters = len(np.unique(clusters))                       # classify.py:157
    medoids = np.zeros(n_clusters, dtype=int)
    sq = squareform(distmat)                                    # :159
    for i in range(1, n_clusters + 1):                          # :161
        indices = np.where(clusters == i)[0]
        sub = sq[indices][:, indices]
        medoids[i - 1] = indices[np.argmin(np.sum(sub, axis=0))]  # :164
    return medoids


def sum_distmat(distmat):
    return np.sum(distmat)                                      # classify.py:177


def member_sums(distmat, clusters):
    """Per-frame sum of distances to the frames of the same cluster (what the argmin above ranks), for tests that
    need to tell a genuine mismatch from a floating-point tie."""
    clusters = np.asarray(clusters)
    sq = squareform(distmat)
    same = clusters[:, None] == clusters[None, :]
    return np.where(same, sq, 0.0).sum(axis=1)
