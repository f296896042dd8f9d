"""Restatement of the subset of the third-party ``rmsd`` package (charnley/rmsd,
pinned by the reference only as ``rmsd>=1.6.3`` in requirements.txt:2) that the
reference hot path reaches.  TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

The package source is NOT under /root/reference and cannot be installed here
(no network), so these functions restate its published algorithm; each one
names the reference call sites (clusttraj/distmat.py) that reach it.  The
evidence that the first four are faithful is the bit-exact reproduction of the
reference's golden vector (tests/test_oracle_golden.py).
"""

import numpy as np
from scipy.optimize import linear_sum_assignment
from scipy.spatial.distance import cdist


def centroid(X):
    """Mean position of the rows of X.  Call sites: distmat.py:148-149,157-158,
    164-165,174."""
    return X.mean(axis=0)


def kabsch(P, Q):
    """Proper rotation U (3x3) minimising |P U - Q|, by SVD of the covariance.

    Call sites: distmat.py:178,220,227 (and inside kabsch_rmsd at :275).
    The covariance is P^T Q; a reflection is removed by flipping the last
    column of the left singular vectors when det(V) det(W) < 0.
    """
    cov = P.T @ Q
    left, sing, right_t = np.linalg.svd(cov)
    if np.linalg.det(left) * np.linalg.det(right_t) < 0.0:
        sing[-1] = -sing[-1]
        left[:, -1] = -left[:, -1]
    return left @ right_t


def rmsd(V, W):
    """sqrt(sum |V-W|^2 / n_rows).  Call site: distmat.py:270."""
    d = np.asarray(V) - np.asarray(W)
    return np.sqrt((d * d).sum() / len(V))


def kabsch_rotate(P, Q):
    return P @ kabsch(P, Q)


def kabsch_weighted(P, Q, W=None):
    """Weighted Kabsch (returns U, translation V, weighted rmsd).

    Call site: distmat.py:273 via kabsch_weighted_rmsd.  Restated from memory
    of rmsd 1.6.x; no reference test pins it ("parity unpinned").
    The weights are per atom; the covariance is sum_j w_j P_j Q_j^T, corrected
    by the weighted centres and scaled by iw = 3 / sum(W replicated on 3 axes).
    """
    n = len(P)
    w = np.full(n, 1.0 / n) if W is None else np.asarray(W, dtype=float)
    w3 = np.stack([w, w, w], axis=1)
    iw = 3.0 / w3.sum()
    cov = (P * w3).T @ Q
    cmp_ = (P * w3).sum(axis=0)
    cmq_ = (Q * w3).sum(axis=0)
    psq = (P * P * w3).sum() - (cmp_ * cmp_).sum() * iw
    qsq = (Q * Q * w3).sum() - (cmq_ * cmq_).sum() * iw
    cov = (cov - np.outer(cmp_, cmq_) * iw) * iw
    left, sing, right_t = np.linalg.svd(cov)
    if np.linalg.det(left) * np.linalg.det(right_t) < 0.0:
        sing[-1] = -sing[-1]
        left[:, -1] = -left[:, -1]
    U = left @ right_t
    msd = (psq + qsq) * iw - 2.0 * sing.sum()
    msd = max(msd, 0.0)
    trans = (cmp_ - U @ cmq_) * iw
    return U, trans, np.sqrt(msd)


def kabsch_weighted_rmsd(P, Q, W=None):
    """Call site: distmat.py:273."""
    return kabsch_weighted(P, Q, W)[2]


def kabsch_rmsd(P, Q, W=None, translate=False):
    """Rotate P onto Q with kabsch (NO re-centring unless translate) and return
    rmsd.  Call site: distmat.py:275."""
    if translate:
        Q = Q - centroid(Q)
        P = P - centroid(P)
    if W is not None:
        return kabsch_weighted_rmsd(P, Q, W)
    return rmsd(kabsch_rotate(P, Q), Q)


def hungarian(A, B):
    """Column assignment of the min-cost perfect matching on the EUCLIDEAN
    (not squared) distance matrix between rows of A and rows of B."""
    cost = cdist(A, B, "euclidean")
    return linear_sum_assignment(cost)[1]


def reorder_hungarian(p_atoms, q_atoms, p_coord, q_coord):
    """Per-element Hungarian reordering.  Bound at io.py:559-561, called at
    distmat.py:192,243 as reorder(ref_atoms, mov_atoms, ref_coords, mov_coords);
    returns ``view`` such that mov[view] lines up with ref."""
    view = np.full(np.asarray(q_atoms).shape, -1, dtype=int)
    for element in np.unique(p_atoms):
        (p_idx,) = np.where(p_atoms == element)
        (q_idx,) = np.where(q_atoms == element)
        view[p_idx] = q_idx[hungarian(p_coord[p_idx], q_coord[q_idx])]
    return view


def reorder_distance(p_atoms, q_atoms, p_coord, q_coord):
    """Per-element reordering by distance from the origin (the frames are centred when clusttraj calls it):
    the k-th closest atom of an element in the reference is paired with the k-th closest atom of that element in
    the moving frame.  Bound at io.py:562-563 (``--reorder-alg distance``), same call sites as reorder_hungarian.

    Restated from the published rmsd 1.6.x algorithm (np.linalg.norm per atom, np.argsort of both sets, the rank of
    each reference atom looked up in the moving frame's order).  PARITY UNPINNED: the package is not importable here
    and the reference's tests never run this mode; equal norms (measure zero) are ordered by index here."""
    view = np.zeros(np.asarray(q_atoms).shape, dtype=int)
    for element in np.unique(p_atoms):
        (p_idx,) = np.where(p_atoms == element)
        (q_idx,) = np.where(q_atoms == element)
        a_order = np.argsort(np.linalg.norm(p_coord[p_idx], axis=1), kind="stable")
        b_order = np.argsort(np.linalg.norm(q_coord[q_idx], axis=1), kind="stable")
        rank_of_a = np.argsort(a_order, kind="stable")
        view[p_idx] = q_idx[b_order[rank_of_a]]
    return view
