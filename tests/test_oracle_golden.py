"""CPU: pin the oracle against the reference's golden vector and against the
vectors the unmodified reference produced over shims (tests/golden)."""
import os

import numpy as np
import pytest

from oracle import distmat_restated as O
from oracle import rmsd_restated as R
from oracle.xyz_min import read_xyz


def _run(golden_dir, ds, args, **kw):
    atoms, coords = read_xyz(os.path.join(golden_dir, ds["trajfile"]))
    noh, do_reorder, rs, ns, ws, excl, fk = args
    return O.build_distance_matrix(
        atoms, coords, noh=noh, reorder=R.reorder_hungarian if do_reorder else None,
        reorder_solvent_only=rs, nsatoms=ns, weight_solute=ws,
        reorderexcl=np.asarray(excl, np.int32), final_kabsch=fk, **kw)


def test_reference_golden_vector_bit_exact(golden_dir, reference_vectors, ref_golden_distmat):
    """test/test_distmat.py:7-16 pins this config at abs 1e-8; the oracle is bit-exact."""
    ds = reference_vectors["testtraj"]
    got = _run(golden_dir, ds, ds["modes"]["noh_ns17"]["args"])
    assert got.shape == (3,)
    assert np.array_equal(got, ref_golden_distmat)


@pytest.mark.parametrize("dataset", ["testtraj", "synth_solv", "synth_plain"])
def test_oracle_matches_unmodified_reference(golden_dir, reference_vectors, dataset):
    ds = reference_vectors[dataset]
    for name, mode in ds["modes"].items():
        got = _run(golden_dir, ds, mode["args"])
        assert got.shape == mode["values"].shape, name
        assert np.abs(got - mode["values"]).max() <= 1e-12, name


def test_oracle_pool_path(golden_dir, reference_vectors, ref_golden_distmat):
    """distmat.py:73-76 process-pool path (test/test_distmat.py:70-75)."""
    ds = reference_vectors["testtraj"]
    got = _run(golden_dir, ds, ds["modes"]["noh_ns17"]["args"], n_workers=2)
    assert np.array_equal(got, ref_golden_distmat)
    pool = np.array([float.fromhex(h) for h in ds["pool_noh_ns17_hex"]])
    assert np.array_equal(pool, ref_golden_distmat)


def test_closed_form_is_the_no_reorder_value(golden_dir, reference_vectors):
    """SURVEY §8a closed forms: the GPU epilogue's specification."""
    atoms, coords = read_xyz(os.path.join(golden_dir, "ref_testtraj.xyz"))
    z = atoms[0]
    for name, keep, ncen in (("plain", z > 0, None), ("noh", z != 1, None),
                             ("ns17", z > 0, 17), ("noh_ns17", z != 1, 7)):
        want = reference_vectors["testtraj"]["modes"][name]["values"]
        got = []
        for i in range(3):
            for j in range(i + 1, 3):
                Q = coords[i][keep].copy()
                P = coords[j][keep].copy()
                Q -= Q[:ncen].mean(axis=0)
                P -= P[:ncen].mean(axis=0)
                got.append(O.closed_form_kabsch_rmsd(P, Q))
        assert np.abs(np.array(got) - want).max() < 1e-12, name


def test_linkage_labels_config1(reference_vectors):
    """BASELINE config 1: plain, -rmsd 1.0, average linkage -> labels [1 2 3]."""
    import scipy.cluster.hierarchy as hcl

    vec = reference_vectors["testtraj"]["modes"]["plain"]["values"]
    Z = hcl.linkage(vec, "average")
    assert list(hcl.fcluster(Z, 1.0, criterion="distance")) == [1, 2, 3]
    assert Z[0, 2] == pytest.approx(3.4161225575825, abs=1e-12)
    assert Z[1, 2] == pytest.approx(4.419679587281052, abs=1e-12)


# ---- consumers of the matrix (clusttraj/classify.py:146-177) -------------------------------------------------------

def _consumer_cases(golden_dir):
    import json

    with open(os.path.join(golden_dir, "reference_consumers.json")) as fh:
        data = json.load(fh)
    mats = {k: np.array([float.fromhex(h) for h in v]) for k, v in data["matrices"].items()}
    return [(name, mats[c["matrix"]], np.array(c["clusters"]), np.array(c["medoids"]), float.fromhex(c["sum_hex"]))
            for name, c in data["cases"].items()]


def test_consumers_oracle_matches_unmodified_reference(golden_dir):
    """oracle/classify_restated.py against what the unmodified classify.py produced (oracle/make_golden.py)."""
    from oracle import classify_restated as C

    cases = _consumer_cases(golden_dir)
    assert len(cases) >= 8
    for name, d, clusters, medoids, total in cases:
        assert (C.find_medoids_from_clusters(d, clusters) == medoids).all(), name
        assert C.sum_distmat(d) == total, name


# ---- structure output: align_mol (clusttraj/io.py:582-756) --------------------------------------------------------------

def _align_cases(golden_dir):
    import json

    with open(os.path.join(golden_dir, "reference_align.json")) as fh:
        data = json.load(fh)
    for ds in data.values():
        for mode in ds["modes"].values():
            mode["aligned"] = np.array([[[float.fromhex(h) for h in row] for row in fr] for fr in mode["aligned_hex"]])
    return data


def test_align_oracle_matches_unmodified_reference(golden_dir):
    """oracle/io_restated.align_coords against the coordinates the unmodified align_mol printed (oracle/make_golden.py)."""
    from oracle import io_restated as IO

    n = 0
    for ds in _align_cases(golden_dir).values():
        atoms, coords = read_xyz(os.path.join(golden_dir, ds["trajfile"]))
        for name, mode in ds["modes"].items():
            noh, do_reorder, rs, ns, ws, excl, fk = mode["args"]
            reorder = R.reorder_hungarian if do_reorder else None
            Q, Qa, _ = IO.reference_frame(atoms[0], coords[0], noh, ns)
            for k in range(1, len(coords)):
                got = IO.align_coords(atoms[k], coords[k], Q, Qa, noh, reorder, rs, ns, ws, np.asarray(excl, np.int32), fk)
                assert np.abs(got - mode["aligned"][k - 1]).max() <= 1e-12, (name, k)
                n += 1
    assert n >= 40


# ---- hierarchical clustering (the hcl.linkage call at clusttraj/classify.py:30,99,127) ------------------------------

def test_linkage_restatement_is_scipys_algorithm_bit_for_bit():
    """oracle/linkage_restated.py (the step order the GPU kernel follows) against the live SciPy module: nearest-neighbour
    chain for complete/average/weighted/ward, Prim for single, stable sort + union-find labels — identical Z, ties and
    zero distances included."""
    import scipy.cluster.hierarchy as hcl
    from scipy.spatial.distance import pdist

    from oracle.linkage_restated import linkage

    rng = np.random.default_rng(5)
    for trial in range(16):
        n = int(rng.integers(2, 48))
        kind = trial % 4
        if kind == 0:
            d = rng.random(n * (n - 1) // 2) * 5
        elif kind == 1:
            d = np.round(rng.random(n * (n - 1) // 2) * 3, 1)
        elif kind == 2:
            x = rng.normal(size=(n, 3))
            x[n // 2:] = x[: n - n // 2]
            d = pdist(x)
        else:
            d = pdist(rng.normal(size=(n, 5)))
        for method in ("single", "complete", "average", "weighted", "ward"):
            assert linkage(d, method).tobytes() == hcl.linkage(d, method).tobytes(), (trial, n, method)
