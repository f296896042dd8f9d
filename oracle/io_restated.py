"""CPU restatement of the structure-output helpers of clusttraj/io.py.  TEST INFRASTRUCTURE ONLY: imported by tests/
and never by the product.

  reference_frame(q_atoms, q_all, noh, nsatoms)   follows io.py:823-856 / :946-974 (the medoid: mask, centre)
  align_coords(...)                               follows io.py:582-746 (align_mol) and returns the rotated ALL-atom
                                                   coordinates that :748-756 prints

Pinned by tests/golden/reference_align.json, produced by the UNMODIFIED clusttraj.io.align_mol over the import shims
(oracle/make_golden.py); the -ws --final-kabsch branch goes through rmsd.kabsch_weighted, whose restatement
(oracle/rmsd_restated.py) is from memory of the rmsd package: parity unpinned for that branch.
"""
import numpy as np

from oracle import rmsd_restated as rmsd


def reference_frame(q_atoms, q_all, noh, nsatoms):
    """(Q centred kept atoms, Qa kept atomic numbers, ALL atoms minus the same centre) of the medoid."""
    q_atoms = np.asarray(q_atoms)
    q_all = np.asarray(q_all, dtype=float)
    natoms = nsatoms
    if noh:
        natoms = len(np.where(q_atoms[:nsatoms] != 1)[0])                       # io.py:833-835
    if nsatoms:
        if noh:
            keep = np.where(q_atoms != 1)
            Q, Qa = np.copy(q_all[keep]), np.copy(q_atoms[keep])
        else:
            Q, Qa = np.copy(q_all), np.copy(q_atoms)
        qcenter = rmsd.centroid(Q[:natoms])
    elif noh:
        keep = np.where(q_atoms != 1)
        Q = np.copy(q_all[keep]); qcenter = rmsd.centroid(Q); Qa = np.copy(q_atoms[keep])
    else:
        Q = np.copy(q_all); qcenter = rmsd.centroid(Q); Qa = np.copy(q_atoms)
    Q -= qcenter                                                                  # :856
    return Q, Qa, q_all - qcenter


def align_coords(p_atoms, p_all, Q, Qa, noh, reorder, reorder_solvent_only, nsatoms, weight_solute, reorderexcl,
                 final_kabsch):
    p_atoms = np.asarray(p_atoms)
    p_all = np.array(p_all, dtype=float)
    reorderexcl = np.asarray(reorderexcl, dtype=np.int32)
    natoms = nsatoms
    if noh:
        natoms = len(np.where(p_atoms[:nsatoms] != 1)[0])                       # io.py:620-622
    if nsatoms:
        if noh:
            keep = np.where(p_atoms != 1)
            P, Pa = np.copy(p_all[keep]), np.copy(p_atoms[keep])
        else:
            P, Pa = np.copy(p_all), np.copy(p_atoms)
        pcenter = rmsd.centroid(P[:natoms])
    elif noh:
        keep = np.where(p_atoms != 1)
        P = np.copy(p_all[keep]); pcenter = rmsd.centroid(P); Pa = np.copy(p_atoms[keep])
    else:
        P = np.copy(p_all); pcenter = rmsd.centroid(P); Pa = np.copy(p_atoms)
    P -= pcenter                                                                  # :646-647
    p_all -= pcenter
    if nsatoms:
        P -= rmsd.centroid(Q[:natoms])                                            # :652 (P only, not p_all)
        U = rmsd.kabsch(P[:natoms], Q[:natoms])                                   # :656
        P = np.dot(P, U)
        p_all = np.dot(p_all, U)
        if reorder and not reorder_solvent_only:                                  # :663
            soluexcl = np.where(reorderexcl < natoms)
            soluteview = np.delete(np.arange(natoms), reorderexcl[soluexcl])
            Pview, Paview = P[soluteview], Pa[soluteview]
            prr = reorder(Qa[soluteview], Paview, Q[soluteview], Pview)
            Pview, Paview = Pview[prr], Paview[prr]
            whereins = np.where(np.atleast_1d(np.isin(np.arange(natoms), reorderexcl)))
            at = [x - whereins[0].tolist().index(x) for x in whereins[0]]
            Psolu = np.insert(Pview, at, P[reorderexcl[soluexcl]], axis=0)
            Pasolu = np.insert(Paview, at, Pa[reorderexcl[soluexcl]], axis=0)
            P = np.concatenate((Psolu, P[np.arange(len(P) - natoms) + natoms]))
            Pa = np.concatenate((Pasolu, Pa[np.arange(len(Pa) - natoms) + natoms]))
            U = rmsd.kabsch(P[:natoms], Q[:natoms])                               # :692
            P = np.dot(P, U)
            p_all = np.dot(p_all, U)
    else:
        U = rmsd.kabsch(P, Q)                                                     # :699
        P = np.dot(P, U)
        p_all = np.dot(p_all, U)
    if reorder:                                                                   # :704
        if nsatoms:
            exclusions = np.unique(np.concatenate((np.arange(natoms), reorderexcl)))
        else:
            exclusions = reorderexcl
        view = np.delete(np.arange(len(P)), exclusions)
        Pview, Paview = P[view], Pa[view]
        prr = reorder(Qa[view], Paview, Q[view], Pview)
        Pview = Pview[prr]
        whereins = np.where(np.atleast_1d(np.isin(np.arange(len(P)), exclusions)))
        Pr = np.insert(Pview, [x - whereins[0].tolist().index(x) for x in whereins[0]], P[exclusions], axis=0)
    else:
        Pr = P
    if weight_solute and final_kabsch:                                            # :733
        W = np.zeros(Pr.shape[0])
        W[:natoms] = weight_solute / natoms
        W[natoms:] = (1.0 - weight_solute) / (Pr.shape[0] - natoms)
        R, T, _ = rmsd.kabsch_weighted(Q, Pr, W)
        p_all = np.dot(p_all, R.T) + T
    elif nsatoms and reorder and final_kabsch:                                    # :743
        U = rmsd.kabsch(Pr, Q)
        p_all = np.dot(p_all, U)
    return p_all
