"""Shared helpers of the parity tests: run the oracle (CPU restatement) and the engine on the same
in-memory frames."""
import numpy as np

from oracle import distmat_restated as O
from oracle import rmsd_restated as R


def _alg(reorder):
    if not reorder:
        return None
    return R.reorder_distance if reorder == "distance" else R.reorder_hungarian


def oracle_rows(atoms, coords, rows, noh=False, reorder=False, rs=False, ns=None, ws=None, excl=(), fk=False,
                n_workers=1):
    return O.build_distance_matrix(atoms, coords, noh=noh, reorder=_alg(reorder),
                                   reorder_solvent_only=rs, nsatoms=ns, weight_solute=ws,
                                   reorderexcl=np.asarray(excl, np.int32), final_kabsch=fk, rows=rows,
                                   n_workers=n_workers)


def oracle_pairs(atoms, coords, pairs, noh=False, reorder=False, rs=False, ns=None, ws=None, excl=(), fk=False):
    """values and composite permutations for explicit (i, j) pairs."""
    vals, perms = [], []
    for i, j in pairs:
        (v, p), = O.distmat_line(i, atoms, coords, noh, _alg(reorder), rs, ns, ws,
                                 np.asarray(excl, np.int32), fk, columns=[j], return_perm=True)
        vals.append(v)
        perms.append(p)
    return np.array(vals), np.array(perms)


def engine_opts(noh=False, reorder=False, rs=False, ns=None, ws=None, excl=(), fk=False):
    from clusttraj_b200 import _engine

    return _engine.make_opts(no_hydrogen=noh, solute_natoms=ns, reorder=reorder, reorder_solvent_only=rs,
                             final_kabsch=fk, weight_solute=ws, reorder_excl=excl)


def condensed_index(M, i, j):
    return M * i - i * (i + 1) // 2 + (j - i - 1)
