"""Drop-in for the clustering step and the two consumers of the RMSD matrix that clusttraj runs right after it
(clusttraj/classify.py), on the GPU and without the M x M ``squareform`` the reference builds on the host (3.2 GB at
20k frames, 80 GB at 100k):

  linkage(distmat, method, optimal_ordering)     the hcl.linkage call at classify.py:30,99,127
  classify_structures_silhouette(clust_opt, ..)  classify.py:11-81 (linkage on the GPU; the silhouette scan itself is the
                                                 reference's unchanged scikit-learn call on the host)
  classify_structures(clust_opt, distmat)        classify.py:84-110
  classify_structures_nclusters(clust_opt, ...)  classify.py:113-143
  find_medoids_from_clusters(distmat, clusters)  classify.py:146-165
  sum_distmat(distmat)                           classify.py:168-177

Both take the condensed float64 vector ``get_distmat`` returned (it is copied to the GPU; one pass over its 8 L
bytes).  Like the rest of the package there is no CPU fallback: without a GPU they raise EngineError.  Ties between
candidate medoids go to the lowest frame index, as ``np.argmin`` does in the reference; two candidates whose sums
agree to the last bit of a differently ordered floating-point sum can resolve differently (documented in DESIGN.md).
"""
import numpy as np

from .distmat import _engine_for, visible_devices


_trusted_ref = None


def trust_resident(distmat):
    """The caller (clusttraj_b200.main) vouches that it will not write to `distmat` between the calls of this module:
    the copy an engine call leaves on the GPU is then reused by the next one instead of being uploaded again (1.6 GB
    at 20k frames).  Without this every function uploads what it is given, like any library function would."""
    global _trusted_ref
    import weakref

    _trusted_ref = weakref.ref(distmat) if isinstance(distmat, np.ndarray) else None


def _trusted(distmat):
    return _trusted_ref is not None and _trusted_ref() is distmat


def _n_frames(distmat):
    L = int(np.asarray(distmat).shape[0])
    M = int(round((1.0 + np.sqrt(1.0 + 8.0 * L)) / 2.0))
    if M * (M - 1) // 2 != L:
        raise ValueError("distmat is not a condensed distance vector")   # what squareform raises in the reference
    return M


GPU_LINKAGE_METHODS = ("single", "complete", "average", "weighted", "ward")

def _value_errors(call, *args, **kw):
    """scipy.cluster.hierarchy.linkage raises ValueError on NaN / infinite distances (non-finite coordinates are a
    supported input of the matrix kernels, so such a matrix can reach us).  The engine checks the matrix where it lies
    (one pass over the resident copy, ct_linkage / ct_matrix_consumers; the medoid reduction also needs distances >= 0)
    and reports SciPy's wording: surfaced here as the ValueError the reference's callers expect (main.py:53-66)."""
    from ._engine import EngineError

    try:
        return call(*args, **kw)
    except EngineError as exc:
        msg = str(exc)
        if "condensed distance matrix must" in msg:
            raise ValueError(msg[msg.index("The condensed") if "The condensed" in msg else msg.index("the condensed"):]) from exc
        raise


def linkage(distmat, method="average", optimal_ordering=False):
    """scipy.cluster.hierarchy.linkage(distmat, method, optimal_ordering=...) with the clustering itself on the GPU:
    the linkage matrix is the one SciPy returns, bit for bit (same nearest-neighbour-chain / minimum-spanning-tree
    step order and tie rules).  `centroid` and `median` — SciPy's generic priority-queue algorithm — are
    passed to SciPy unchanged, as is the optional leaf reordering of the finished tree."""
    import scipy.cluster.hierarchy as hcl

    if method not in GPU_LINKAGE_METHODS:
        return hcl.linkage(distmat, method, optimal_ordering=bool(optimal_ordering))
    M = _n_frames(distmat)
    eng = _engine_for(visible_devices()[0])
    Z, _ = _value_errors(eng.linkage, M, method, distmat=distmat, trusted=_trusted(distmat))
    if optimal_ordering:
        Z = hcl.optimal_leaf_ordering(Z, distmat)
    return Z


def classify_structures_silhouette(clust_opt, distmat, dstep=0.1):
    """Linkage, then the RMSD threshold with the highest silhouette score among min(Z[:, 2]) + k * dstep
    (classify.py:11-81): ties keep every tying threshold in clust_opt.optimal_cut and adopt the smallest."""
    import scipy.cluster.hierarchy as hcl
    from scipy.spatial.distance import squareform
    from sklearn import metrics

    Z = linkage(distmat, clust_opt.method, optimal_ordering=clust_opt.opt_order)
    square = squareform(distmat)
    best_score, best_t, best_labels = np.float64(-1.0), [], []
    for t in np.arange(min(Z[:, 2]), max(Z[:, 2]), dstep):
        labels = hcl.fcluster(Z, t=t, criterion="distance")
        score = metrics.silhouette_score(square, labels, metric="precomputed")
        if score == best_score:
            best_t.append(t)
            best_labels.append(labels)
        elif score > best_score:
            best_score, best_t, best_labels = score, [t], [labels]
    clusters = best_labels[0]
    clust_opt.update({"optimal_cut": np.array(best_t)})
    np.savetxt(clust_opt.out_clust_name, clusters, fmt="%d")
    return Z, clusters


def compute_metrics(distmat, z_matrix, clusters):
    """(silhouette, Calinski-Harabasz, Davies-Bouldin, cophenetic correlation) of a clustering, as clusttraj/metrics.py:14-45
    computes them: unchanged scikit-learn / SciPy calls on the host."""
    from scipy.cluster.hierarchy import cophenet
    from scipy.spatial.distance import squareform
    from sklearn.metrics import calinski_harabasz_score, davies_bouldin_score, silhouette_score

    square = squareform(distmat)
    return (silhouette_score(square, clusters, metric="precomputed"), calinski_harabasz_score(square, clusters),
            davies_bouldin_score(square, clusters), cophenet(z_matrix, distmat)[0])


def classify_structures(clust_opt, distmat):
    """Linkage + flat clusters at the RMSD threshold, written to clust_opt.out_clust_name (classify.py:84-110)."""
    import scipy.cluster.hierarchy as hcl

    Z = linkage(distmat, clust_opt.method, optimal_ordering=clust_opt.opt_order)
    clusters = hcl.fcluster(Z, clust_opt.min_rmsd, criterion="distance")
    np.savetxt(clust_opt.out_clust_name, clusters, fmt="%d")
    return Z, clusters


def classify_structures_nclusters(clust_opt, distmat):
    """Linkage + the cut that gives clust_opt.n_clusters clusters (classify.py:113-143)."""
    import scipy.cluster.hierarchy as hcl

    Z = linkage(distmat, clust_opt.method, optimal_ordering=clust_opt.opt_order)
    n_snapshots = Z.shape[0] + 1
    if clust_opt.n_clusters > n_snapshots:
        raise ValueError(f"Cannot request {clust_opt.n_clusters} clusters from {n_snapshots} snapshots")
    clusters = hcl.cut_tree(Z, n_clusters=[clust_opt.n_clusters]).ravel() + 1
    np.savetxt(clust_opt.out_clust_name, clusters, fmt="%d")
    return Z, clusters


def find_medoids_from_clusters(distmat, clusters):
    """Indices of the medoids of the clusters (classify.py:146-165): for cluster c = 1..n the member with the
    smallest sum of distances to the members of c."""
    clusters = np.asarray(clusters)
    M = _n_frames(distmat)
    if clusters.shape != (M,):
        raise ValueError("clusters must hold one label per frame")
    eng = _engine_for(visible_devices()[0])
    _, medoids, _ = _value_errors(eng.matrix_consumers, M, clusters=clusters, distmat=distmat, trusted=_trusted(distmat))
    return medoids.astype(int)


def sum_distmat(distmat):
    """Sum of the RMSD matrix (classify.py:168-177)."""
    M = _n_frames(distmat)
    eng = _engine_for(visible_devices()[0])
    total, _, _ = eng.matrix_consumers(M, clusters=None, distmat=distmat, trusted=_trusted(distmat))
    return np.float64(total)
