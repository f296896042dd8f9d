"""Generate tests/golden/*.json by running the UNMODIFIED reference module
/root/reference/clusttraj/distmat.py (compute_distmat_line / build_distance_matrix)
over the import shims in oracle/ref_shims.  TEST INFRASTRUCTURE.

Run in the authoring container only (the GPU box has no /root/reference):

    python -m oracle.make_golden

The committed outputs are what the CPU and GPU parity tests compare against.
Values are stored as float64 hex strings so they round-trip bit-exactly.
"""

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

# name -> (noh, reorder?, reorder_solvent_only, nsatoms, weight_solute, excl(0-based), final_kabsch)
MODES_TESTTRAJ = {
    "plain": (False, False, False, None, None, [], False),
    "noh": (True, False, False, None, None, [], False),
    "ns17": (False, False, False, 17, None, [], False),
    "noh_ns17": (True, False, False, 17, None, [], False),  # the reference's own golden config
    "e": (False, True, False, None, None, [], False),
    "e_noh": (True, True, False, None, None, [], False),
    "ns17_e": (False, True, False, 17, None, [], False),
    "ns17_e_rs": (False, True, True, 17, None, [], False),
    "ns17_e_noh": (True, True, False, 17, None, [], False),
    "ns17_e_fk": (False, True, False, 17, None, [], True),
    "ns17_e_rs_fk": (False, True, True, 17, None, [], True),
    "ns17_e_ws08": (False, True, False, 17, 0.8, [], False),
    "ns17_e_ws08_fk": (False, True, False, 17, 0.8, [], True),
    "ns17_ws08": (False, False, False, 17, 0.8, [], False),
    "e_excl012": (False, True, False, None, None, [0, 1, 2], False),
    "ns17_e_excl012": (False, True, False, 17, None, [0, 1, 2], False),
    "ns17_e_excl1920": (False, True, False, 17, None, [19, 20], False),
    "ns17_e_rs_excl19": (False, True, True, 17, None, [19], False),
}

MODES_SOLV = {
    "plain": (False, False, False, None, None, [], False),
    "noh": (True, False, False, None, None, [], False),
    "ns12": (False, False, False, 12, None, [], False),
    "e": (False, True, False, None, None, [], False),
    "ns12_e": (False, True, False, 12, None, [], False),
    "ns12_e_rs": (False, True, True, 12, None, [], False),
    "ns12_e_noh": (True, True, False, 12, None, [], False),
    "ns12_e_fk": (False, True, False, 12, None, [], True),
    "ns12_e_ws06": (False, True, False, 12, 0.6, [], False),
    "ns12_e_excl": (False, True, False, 12, None, [1, 14, 15], False),
}


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "ref_shims"))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REFERENCE)
    import clusttraj.distmat as ref_distmat  # the unmodified reference file
    import rmsd as shim_rmsd

    assert ref_distmat.__file__.startswith(REFERENCE), ref_distmat.__file__
    return ref_distmat, shim_rmsd


def _run_reference(ref_distmat, shim_rmsd, trajfile, mode):
    from clusttraj.utils import get_mol_info
    from openbabel import pybel

    noh, do_reorder, rs, ns, ws, excl, fk = mode
    reorder = shim_rmsd.reorder_hungarian if do_reorder else None
    lines = []
    for idx1, mol in enumerate(pybel.readfile("xyz", trajfile)):
        lines.append(
            ref_distmat.compute_distmat_line(
                idx1, get_mol_info(mol), trajfile, noh, reorder, rs, ns, ws,
                np.asarray(excl, np.int32), fk,
            )
        )
    return np.asarray([x for line in lines if len(line) > 0 for x in line], dtype=np.float64)


def main():
    from clusttraj_b200.synth import make_solute_solvent, make_trajectory
    from oracle.xyz_min import write_xyz

    ref_distmat, shim_rmsd = _import_reference()
    out = {}

    traj = os.path.join(GOLDEN, "ref_testtraj.xyz")
    out["testtraj"] = {"trajfile": "ref_testtraj.xyz", "modes": {}}
    for name, mode in MODES_TESTTRAJ.items():
        vec = _run_reference(ref_distmat, shim_rmsd, traj, mode)
        out["testtraj"]["modes"][name] = {"args": mode, "hex": [float(v).hex() for v in vec],
                                           "repr": [repr(float(v)) for v in vec]}

    # small synthetic solute+solvent trajectory (12 solute atoms + 8 waters, 6 frames)
    z, x = make_solute_solvent(6, n_solute=12, n_solvent_mol=8, seed=11, box=6.0)
    solv = os.path.join(GOLDEN, "synth_solv_6x36.xyz")
    write_xyz(solv, z, x)
    out["synth_solv"] = {"trajfile": "synth_solv_6x36.xyz", "modes": {}}
    for name, mode in MODES_SOLV.items():
        vec = _run_reference(ref_distmat, shim_rmsd, solv, mode)
        out["synth_solv"]["modes"][name] = {"args": mode, "hex": [float(v).hex() for v in vec],
                                             "repr": [repr(float(v)) for v in vec]}

    # small plain trajectory incl. duplicate + mirrored frames (10 frames x 24 atoms)
    z, x = make_trajectory(10, 24, seed=5, n_modes=3)
    plain = os.path.join(GOLDEN, "synth_plain_10x24.xyz")
    write_xyz(plain, z, x)
    out["synth_plain"] = {"trajfile": "synth_plain_10x24.xyz", "modes": {}}
    for name in ("plain", "noh", "e"):
        mode = MODES_SOLV[name]
        vec = _run_reference(ref_distmat, shim_rmsd, plain, mode)
        out["synth_plain"]["modes"][name] = {"args": mode, "hex": [float(v).hex() for v in vec],
                                              "repr": [repr(float(v)) for v in vec]}

    # the Pool path of the unmodified build_distance_matrix on the reference's own config
    from clusttraj.io import ClustOptions

    with tempfile.TemporaryDirectory() as tmp:
        opt = ClustOptions(trajfile=traj, no_hydrogen=True, solute_natoms=17, reorder_alg=None,
                           reorder_solvent_only=False, weight_solute=None,
                           reorder_excl=np.asarray([], np.int32), final_kabsch=False, n_workers=2,
                           distmat_name=os.path.join(tmp, "d.npy"))
        pool_vec = ref_distmat.build_distance_matrix(opt)
    out["testtraj"]["pool_noh_ns17_hex"] = [float(v).hex() for v in pool_vec]

    with open(os.path.join(GOLDEN, "reference_vectors.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(GOLDEN, "reference_vectors.json"))
    golden = np.load(os.path.join(GOLDEN, "ref_test_distmat.npy"))
    mine = np.array([float.fromhex(h) for h in out["testtraj"]["modes"]["noh_ns17"]["hex"]])
    print("max |reference-over-shims - reference golden| =", np.abs(mine - golden).max())



def make_consumers_golden():
    """tests/golden/reference_consumers.json: the UNMODIFIED clusttraj.classify.find_medoids_from_clusters /
    sum_distmat (classify.py:146-177) on small matrices (the reference's own tests never call them)."""
    import scipy.cluster.hierarchy as hcl

    _import_reference()
    import clusttraj.classify as ref_classify

    assert ref_classify.__file__.startswith(REFERENCE), ref_classify.__file__
    from clusttraj_b200.synth import make_trajectory
    from oracle import distmat_restated as O

    cases, matrices = {}, {}

    def add(name, matrix_name, distmat, clusters):
        distmat = np.asarray(distmat, dtype=np.float64)
        clusters = np.asarray(clusters, dtype=int)
        med = ref_classify.find_medoids_from_clusters(distmat, clusters)
        tot = ref_classify.sum_distmat(distmat)
        matrices[matrix_name] = [float(v).hex() for v in distmat]
        cases[name] = {"matrix": matrix_name, "clusters": [int(c) for c in clusters],
                       "medoids": [int(m) for m in med], "sum_hex": float(tot).hex(), "sum_repr": repr(float(tot))}

    golden = np.load(os.path.join(GOLDEN, "ref_test_distmat.npy"))
    add("ref_golden_3_clusters", "ref_golden", golden, [1, 2, 3])       # test/conftest.py:83-85
    add("ref_golden_2_clusters", "ref_golden", golden, [1, 1, 2])       # test/test_classify.py:44
    add("ref_golden_1_cluster", "ref_golden", golden, [1, 1, 1])
    z, x = make_trajectory(60, 12, seed=31, n_modes=4)
    d60 = O.build_distance_matrix(z, x)
    Z = hcl.linkage(d60, "average")
    for t in (0.5, 1.5, 3.0):
        add(f"synth60_average_t{t}", "synth60", d60, hcl.fcluster(Z, t, criterion="distance"))
    add("synth60_maxclust5", "synth60", d60, hcl.fcluster(hcl.linkage(d60, "ward"), 5, criterion="maxclust"))
    rng = np.random.default_rng(77)
    d45 = rng.uniform(0.1, 9.0, size=45 * 44 // 2)
    lab = rng.integers(1, 6, size=45)
    lab[:5] = (1, 2, 3, 4, 5)
    add("random45_5_labels", "random45", d45, lab)
    with open(os.path.join(GOLDEN, "reference_consumers.json"), "w") as fh:
        json.dump({"matrices": matrices, "cases": cases}, fh, indent=0)
    print("wrote", os.path.join(GOLDEN, "reference_consumers.json"), {k: v["medoids"] for k, v in cases.items()})




# name -> (noh, reorder?, reorder_solvent_only, nsatoms, weight_solute, excl(0-based), final_kabsch) for align_mol
ALIGN_MODES_TESTTRAJ = {k: MODES_TESTTRAJ[k] for k in ("plain", "noh", "ns17", "noh_ns17", "e", "ns17_e", "ns17_e_rs", "ns17_e_noh",
                                                       "ns17_e_fk", "ns17_e_ws08_fk", "ns17_e_ws08", "ns17_e_excl1920")}
ALIGN_MODES_SOLV = {k: MODES_SOLV[k] for k in ("plain", "ns12_e", "ns12_e_rs", "ns12_e_noh", "ns12_e_fk", "ns12_e_excl")}


def make_align_golden():
    """tests/golden/reference_align.json: the UNMODIFIED clusttraj.io.align_mol (io.py:582-756) superposing every
    later frame of the fixtures on frame 0 prepared as save_clusters_config prepares a medoid (io.py:823-856)."""
    _, shim_rmsd = _import_reference()
    import clusttraj.io as ref_io
    from openbabel import pybel
    from oracle import io_restated as IO

    assert ref_io.__file__.startswith(REFERENCE), ref_io.__file__
    out = {}
    for ds, fname, modes in (("testtraj", "ref_testtraj.xyz", ALIGN_MODES_TESTTRAJ), ("synth_solv", "synth_solv_6x36.xyz", ALIGN_MODES_SOLV)):
        traj = os.path.join(GOLDEN, fname)
        mols = list(pybel.readfile("xyz", traj))
        out[ds] = {"trajfile": fname, "modes": {}}
        for name, mode in modes.items():
            noh, do_reorder, rs, ns, ws, excl, fk = mode
            reorder = shim_rmsd.reorder_hungarian if do_reorder else None
            q_atoms = np.array([a.atomicnum for a in mols[0]])
            q_all = np.array([a.coords for a in mols[0]])
            Q, Qa, _ = IO.reference_frame(q_atoms, q_all, noh, ns)   # verbatim io.py:823-856 (no function of its own in the reference)
            frames = []
            for mol in mols[1:]:
                text = ref_io.align_mol(mol, Q, Qa, noh, reorder, rs, ns, ws, np.asarray(excl, np.int32), fk)
                rows = text.strip().split("\n")[2:]
                frames.append([[float(v).hex() for v in r.split("\t")[1:4]] for r in rows])
            out[ds]["modes"][name] = {"args": mode, "aligned_hex": frames}
    with open(os.path.join(GOLDEN, "reference_align.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", os.path.join(GOLDEN, "reference_align.json"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "consumers":
        make_consumers_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "align":
        make_align_golden()
    else:
        main()
        make_consumers_golden()
        make_align_golden()
