"""CPU restatement of the reference hot path, clusttraj/distmat.py:44-277, with
the trajectory held in memory.  TEST INFRASTRUCTURE — see oracle/__init__.py.

``pair_value`` follows compute_distmat_line's per-pair body (distmat.py:
133-275) step by step, including its quirks:
  * ``natoms`` (solute size after the H mask) is taken from the ROW frame only
    (distmat.py:116-118);
  * with ``-ns`` both frames are centred on the centroid of their first
    ``natoms`` kept atoms (:148-149,168-169) and the (already ~0) solute
    centroid of Q is subtracted from P once more (:174);
  * the solute reorder block runs whenever reorder is set and ``-rs`` is not
    (:184-223);
  * the final value is the un-rotated rmsd only for ``-ns`` + reorder without
    ``--final-kabsch`` (:265-270); everything else is kabsch_rmsd, which
    rotates again but never re-centres (:271-275);
  * truthiness decides every branch (None / 0 / 0.0 are "off").
It additionally tracks the composite atom permutation so tests can compare
permutations, which the reference does not return.
"""

import multiprocessing
import os

import numpy as np

from . import rmsd_restated as R


def _reinsert(reordered, excluded_positions, excluded_rows, total):
    """Restates the np.insert idiom at distmat.py:200-214 and :247-253:
    rebuild an array of ``total`` rows in which the sorted positions
    ``excluded_positions`` receive ``excluded_rows`` (in the order given) and
    the remaining positions receive ``reordered`` in order."""
    where = np.where(np.atleast_1d(np.isin(np.arange(total), excluded_positions)))[0]
    spots = [x - k for k, x in enumerate(where.tolist())]
    return np.insert(reordered, spots, excluded_rows, axis=0)


def pair_value(
    q_atoms,
    q_all,
    p_atoms,
    p_all,
    noh,
    reorder,
    reorder_solvent_only,
    nsatoms,
    weight_solute,
    reorderexcl,
    final_kabsch,
    return_perm=False,
):
    """One entry of the matrix: row frame (q_*) against column frame (p_*).

    ``p_all`` is modified in place exactly when the reference would modify the
    array it got from get_mol_info (it never re-uses it, so callers pass a
    copy); ``q_all`` is modified in place exactly as the reference does
    (distmat.py:169 — Q aliases the row frame when no H mask is applied).
    """
    reorderexcl = np.asarray(reorderexcl, dtype=np.int32)
    natoms = nsatoms
    if noh:
        natoms = len(np.where(q_atoms[:nsatoms] != 1)[0])

    # --- atom selection + centring (distmat.py:134-169) ---------------------
    if noh:
        keep_p = np.where(p_atoms != 1)
        keep_q = np.where(q_atoms != 1)
        P, Q = p_all[keep_p], q_all[keep_q]
        Pa, Qa = p_atoms[keep_p], q_atoms[keep_q]
    else:
        P, Q, Pa, Qa = p_all, q_all, p_atoms, q_atoms
    if nsatoms:
        pcenter, qcenter = R.centroid(P[:natoms]), R.centroid(Q[:natoms])
    else:
        pcenter, qcenter = R.centroid(P), R.centroid(Q)
    P -= pcenter
    Q -= qcenter

    perm = np.arange(len(P))  # Pr[k] == P_kept_centred_rotated[perm[k]]

    # --- rotation(s) before the solvent reorder (distmat.py:172-228) --------
    if nsatoms:
        P -= R.centroid(Q[:natoms])
        P = np.dot(P, R.kabsch(P[:natoms], Q[:natoms]))
        if reorder and not reorder_solvent_only:
            sol_excl = reorderexcl[np.where(reorderexcl < natoms)]
            sview = np.delete(np.arange(natoms), sol_excl)
            order = reorder(Qa[sview], Pa[sview], Q[sview], P[sview])
            Psolu = _reinsert(P[sview][order], sol_excl, P[sol_excl], natoms)
            Pasolu = _reinsert(Pa[sview][order], sol_excl, Pa[sol_excl], natoms)
            permsolu = _reinsert(perm[sview][order], sol_excl, perm[sol_excl], natoms)
            P = np.concatenate((Psolu, P[natoms:]))
            Pa = np.concatenate((Pasolu, Pa[natoms:]))
            perm = np.concatenate((permsolu, perm[natoms:]))
            P = np.dot(P, R.kabsch(P[:natoms], Q[:natoms]))
    else:
        P = np.dot(P, R.kabsch(P, Q))

    # --- reorder of everything not excluded (distmat.py:231-256) ------------
    if reorder:
        if nsatoms:
            exclusions = np.unique(np.concatenate((np.arange(natoms), reorderexcl)))
        else:
            exclusions = reorderexcl
        view = np.delete(np.arange(len(P)), exclusions)
        order = reorder(Qa[view], Pa[view], Q[view], P[view])
        Pr = _reinsert(P[view][order], exclusions, P[exclusions], len(P))
        perm = _reinsert(perm[view][order], exclusions, perm[exclusions], len(P))
    else:
        Pr = P

    # --- final value (distmat.py:259-275) -------------------------------------
    if weight_solute:
        W = np.zeros(Pr.shape[0])
        W[:natoms] = weight_solute / natoms
        W[natoms:] = (1.0 - weight_solute) / (Pr.shape[0] - natoms)
    if nsatoms and reorder and not final_kabsch:
        if weight_solute:
            d = Pr - Q
            val = np.sqrt(np.dot(W, np.sum(d * d, axis=1)))
        else:
            val = R.rmsd(Pr, Q)
    else:
        if weight_solute:
            val = R.kabsch_weighted_rmsd(Pr, Q, W)
        else:
            val = R.kabsch_rmsd(Pr, Q)
    if return_perm:
        return val, perm
    return val


def distmat_line(idx1, atoms, coords, noh, reorder, reorder_solvent_only, nsatoms,
                 weight_solute, reorderexcl, final_kabsch, columns=None, return_perm=False):
    """Row ``idx1`` of the matrix (all idx2 > idx1, or the given ``columns``),
    distmat.py:81-277.  ``atoms`` is int[M][N] or int[N]; ``coords`` float64[M,N,3].
    The row frame is copied once (the reference's q_info is the row's own
    arrays) and then aliased across the row like the reference does."""
    atoms = np.asarray(atoms)
    per_frame = atoms.ndim == 2
    q_atoms = np.array(atoms[idx1] if per_frame else atoms)
    q_all = np.array(coords[idx1], dtype=np.float64)
    out = []
    cols = range(idx1 + 1, len(coords)) if columns is None else columns
    for idx2 in cols:
        if idx1 >= idx2:
            continue
        p_atoms = np.array(atoms[idx2] if per_frame else atoms)
        p_all = np.array(coords[idx2], dtype=np.float64)
        out.append(
            pair_value(q_atoms, q_all, p_atoms, p_all, noh, reorder, reorder_solvent_only,
                       nsatoms, weight_solute, reorderexcl, final_kabsch,
                       return_perm=return_perm)
        )
    return out


def _line_star(args):
    return distmat_line(*args)


def build_distance_matrix(atoms, coords, noh=False, reorder=None, reorder_solvent_only=False,
                          nsatoms=None, weight_solute=None, reorderexcl=(), final_kabsch=False,
                          n_workers=1, rows=None):
    """Condensed vector (distmat.py:44-78): rows in order, each row idx2>idx1.
    ``rows`` restricts to a subset of rows (used by the bounded CPU baseline);
    ``n_workers`` > 1 uses a process pool over rows like distmat.py:73-76."""
    M = len(coords)
    reorderexcl = np.asarray(reorderexcl, dtype=np.int32)
    row_ids = range(M) if rows is None else rows
    tasks = [(i, atoms, coords, noh, reorder, reorder_solvent_only, nsatoms, weight_solute,
              reorderexcl, final_kabsch) for i in row_ids]
    if n_workers and n_workers > 1:
        ctx = multiprocessing.get_context("fork")
        with ctx.Pool(processes=n_workers) as pool:
            lines = pool.map(_line_star, tasks, chunksize=max(1, len(tasks) // (8 * n_workers)))
    else:
        lines = [distmat_line(*t) for t in tasks]
    return np.asarray([x for line in lines if len(line) > 0 for x in line], dtype=np.float64)


def closed_form_kabsch_rmsd(P, Q):
    """Specification of the GPU epilogue for the no-reorder modes:
    sqrt(max(G_P + G_Q - 2 (s1 + s2 + sign(det C) s3), 0) / K) with C = P^T Q,
    for P, Q already centred the way the mode requires."""
    C = P.T @ Q
    s = np.linalg.svd(C, compute_uv=False)
    if np.linalg.det(C) < 0.0:
        s[-1] = -s[-1]
    e = (P * P).sum() + (Q * Q).sum() - 2.0 * s.sum()
    return np.sqrt(max(e, 0.0) / len(P))


def cpu_count():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
