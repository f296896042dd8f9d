"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI, against
  * the reference's own golden vector (test/ref/test_distmat.npy, abs 1e-8 like test/test_distmat.py),
  * the vectors the UNMODIFIED reference produced over import shims (tests/golden/reference_vectors.json),
  * the CPU oracle on seeded synthetic frames,
and through size-independent properties at BASELINE.json's full sizes.
Tolerance: |dRMSD| <= 1e-6 A (north star); the observed error is asserted much tighter where the
arithmetic allows it: 1e-9 for the FP64 (DMMA) Gram kernel and the per-pair kernels, two fixed-point quanta
(~6e-8 A for a 30 A molecule) for the int8 tensor-core Gram kernel, whose only error is the one-off rounding of
the centred coordinates to 31-bit fixed point (csrc/ozaki.cu).  Every pure-Kabsch test runs on both kernels."""
import os

import numpy as np
import pytest

from tests.helpers import condensed_index, engine_opts, oracle_pairs, oracle_rows

pytestmark = pytest.mark.gpu
TOL = 1e-6


def _engine_with(variant):
    """An engine whose pure-Kabsch modes use the int8 tcgen05 kernel (None = the default, automatic choice) or
    the FP64 DMMA kernel ("-2")."""
    from clusttraj_b200 import _engine

    old = os.environ.pop("CLUSTTRAJ_B200_GRAM_VARIANT", None)
    if variant is not None:
        os.environ["CLUSTTRAJ_B200_GRAM_VARIANT"] = variant
    try:
        return _engine.Engine(0)
    finally:
        os.environ.pop("CLUSTTRAJ_B200_GRAM_VARIANT", None)
        if old is not None:
            os.environ["CLUSTTRAJ_B200_GRAM_VARIANT"] = old


@pytest.fixture(scope="module", params=["int8", "dmma"])
def eng(request):
    e = _engine_with(None if request.param == "int8" else "-2")
    e.expected_path = 2 if request.param == "int8" else 1
    return e


def ktol(eng):
    """Tolerance of the pure-Kabsch value of the frames loaded last: 1e-9 (FP64 kernels), or two quanta of the
    fixed-point coordinates (int8 kernel; the worst case of the rounding is 1.8 quanta, typical 0.2)."""
    return max(1e-9, 2.0 * eng.stats()["quant_step"])


def _xyz(golden_dir, name):
    from clusttraj_b200.xyz import read_xyz

    atoms, coords = read_xyz(os.path.join(golden_dir, name))
    return atoms[0], coords


def _full(eng, atoms, coords, **kw):
    eng.load_frames(coords, atoms, engine_opts(**kw))
    out, st = eng.rmsd_rows(0, len(coords))
    return out.copy(), st


def test_reference_golden_vector(eng, golden_dir, ref_golden_distmat):
    atoms, coords = _xyz(golden_dir, "ref_testtraj.xyz")
    got, st = _full(eng, atoms, coords, noh=True, ns=17)
    assert got.shape == (3,)
    assert np.abs(got - ref_golden_distmat).max() < 1e-8
    assert st["launches"] >= 1 and st["kept_atoms"] == 15


@pytest.mark.parametrize("dataset", ["testtraj", "synth_solv", "synth_plain"])
def test_all_modes_against_unmodified_reference(eng, golden_dir, reference_vectors, dataset):
    ds = reference_vectors[dataset]
    atoms, coords = _xyz(golden_dir, ds["trajfile"])
    worst = 0.0
    for name, mode in ds["modes"].items():
        noh, do_reorder, rs, ns, ws, excl, fk = mode["args"]
        got, _ = _full(eng, atoms, coords, noh=noh, reorder=do_reorder, rs=rs, ns=ns, ws=ws, excl=excl, fk=fk)
        assert got.shape == mode["values"].shape, name
        err = np.abs(got - mode["values"]).max()
        worst = max(worst, err)
        assert err <= TOL, (name, err)
        if not do_reorder:
            assert err <= ktol(eng), (name, err)
    print(f"{dataset}: worst |dRMSD| = {worst:.3e}")


def test_drop_in_entry_points(golden_dir, ref_golden_distmat, tmp_path):
    """The reference's own test_distmat.py, against this package's entry points."""
    from clusttraj_b200.distmat import build_distance_matrix, compute_distmat_line, get_distmat
    from clusttraj_b200.io import ClustOptions
    from clusttraj_b200.xyz import read_xyz

    traj = os.path.join(golden_dir, "ref_testtraj.xyz")
    opt = ClustOptions(trajfile=traj, no_hydrogen=True, solute_natoms=17, reorder_alg=None, reorder=False,
                       reorder_solvent_only=False, weight_solute=None, reorder_excl=np.asarray([], np.int32),
                       final_kabsch=False, n_workers=1, input_distmat=False,
                       distmat_name=str(tmp_path / "distmat.npy"), method="ward", min_rmsd=1.0)
    d = get_distmat(opt)
    assert d.dtype == np.float64 and d.flags.c_contiguous and d.flags.writeable
    assert np.abs(d - ref_golden_distmat).max() < 1e-8
    opt.update({"input_distmat": True})
    assert np.abs(get_distmat(opt) - ref_golden_distmat).max() < 1e-8
    assert np.abs(build_distance_matrix(opt) - ref_golden_distmat).max() < 1e-8
    atoms, coords = read_xyz(traj)
    line = compute_distmat_line(0, (atoms[0], coords[0]), traj, True, None, False, 17, None,
                                np.asarray([], np.int32), False)
    assert len(line) == 2
    assert np.abs(np.array(line) - ref_golden_distmat[:2]).max() < 1e-8
    line = compute_distmat_line(0, (atoms[0], coords[0]), traj, True, None, False, 17, 0.8,
                                np.asarray([], np.int32), True)
    assert len(line) == 2  # test_distmat.py:41-67 checks only the shape of the -ws --final-kabsch row


@pytest.mark.parametrize("M,N,kw", [
    (700, 100, dict()),
    (333, 57, dict(noh=True)),
    (260, 64, dict(ns=20)),
    (200, 301, dict(noh=True, ns=31)),
])
def test_kabsch_modes_vs_oracle(eng, M, N, kw):
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(M, N, seed=1234 + M)
    got, st = _full(eng, atoms, coords, **kw)
    rows = sorted({0, 1, 2, 31, 32, 33, 63, 64, M // 2, M - 3, M - 2})
    want = oracle_rows(atoms, coords, rows, **kw)
    pos = 0
    worst = 0.0
    for r in rows:
        n = M - 1 - r
        lo = condensed_index(M, r, r + 1)
        err = np.abs(got[lo: lo + n] - want[pos: pos + n]).max()
        worst = max(worst, err)
        pos += n
    print(f"M={M} N={N} {kw}: worst |dRMSD| = {worst:.3e}, flagged = {st['flagged']}, path = {st['gram_path']}")
    assert st["gram_path"] == eng.expected_path
    assert worst <= ktol(eng)
    assert st["flagged"] >= 1  # frames 1, 2 duplicate frame 0: the near-duplicate path was exercised
    assert np.isfinite(got).all() and (got >= 0).all()


def test_exact_duplicates_and_mirror(eng):
    """RMSD = 0 edge (cancellation in G_i + G_j - 2 lambda) and the det < 0 (mirror image) edge."""
    rng = np.random.default_rng(5)
    base = rng.normal(0, 12.0, size=(300, 3))
    th = 0.7
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    frames = np.stack([base, base @ Rz.T + 3.0, base * np.array([1, 1, -1.0]), base + 1e-7 * rng.normal(size=base.shape),
                       rng.normal(0, 12.0, size=(300, 3))])
    atoms = np.full(300, 6, np.int32)
    got, st = _full(eng, atoms, frames)
    want = oracle_rows(atoms, frames, range(5))
    assert np.abs(got - want).max() <= ktol(eng), np.abs(got - want)
    assert got[0] < 1e-7  # rotated+translated duplicate
    assert st["flagged"] >= 2


def test_fixup_list_overflow_path(golden_dir):
    """Force the flagged-pair list to overflow: the marker/scan path must give the same values."""
    from clusttraj_b200 import _engine

    rng = np.random.default_rng(9)
    base = rng.normal(0, 5.0, size=(40, 3))
    frames = np.stack([base + 1e-9 * rng.normal(size=base.shape) for _ in range(70)] +
                      [rng.normal(0, 5.0, size=(40, 3)) for _ in range(30)])
    atoms = np.full(40, 6, np.int32)
    os.environ["CLUSTTRAJ_B200_FIX_CAPACITY"] = "7"
    try:
        small = _engine.Engine(0)
    finally:
        del os.environ["CLUSTTRAJ_B200_FIX_CAPACITY"]
    small.load_frames(frames, atoms, engine_opts())
    got, st = small.rmsd_rows(0, 100)
    assert st["flagged"] >= 70 * 69 // 2
    want = oracle_rows(atoms, frames, [0, 1, 50, 69, 70, 98])
    pos = 0
    for r in [0, 1, 50, 69, 70, 98]:
        n = 99 - r
        lo = condensed_index(100, r, r + 1)
        assert np.abs(got[lo: lo + n] - want[pos: pos + n]).max() <= ktol(small)
        pos += n
    assert (got >= 0).all()
    small.close()


@pytest.mark.parametrize("kw", [
    dict(reorder=True),
    dict(reorder=True, ns=12),
    dict(reorder=True, ns=12, rs=True),
    dict(reorder=True, ns=12, fk=True),
    dict(reorder=True, ns=12, ws=0.6),
    dict(reorder=True, ns=12, noh=True),
    dict(reorder="distance"),
    dict(reorder="distance", ns=12),
    dict(reorder="distance", ns=12, rs=True, fk=True),
    dict(reorder="distance", ns=12, ws=0.6, excl=(1, 14, 15)),
])
def test_hungarian_permutations_identical(eng, kw):
    """Identical permutation on non-degenerate Hungarian reorders (and on --reorder-alg distance) + value parity."""
    from clusttraj_b200.synth import make_solute_solvent

    atoms, coords = make_solute_solvent(24, n_solute=12, n_solvent_mol=40, seed=77, box=9.0)
    pairs = [(i, j) for i in range(0, 24, 5) for j in range(i + 1, 24, 3)]
    eng.load_frames(coords, atoms, engine_opts(**kw))
    vals, perms, st = eng.pair_values(pairs, want_perms=True)
    want_v, want_p = oracle_pairs(atoms, coords, pairs, **kw)
    assert np.abs(vals - want_v).max() <= TOL
    assert np.abs(vals - want_v).max() <= 1e-9
    assert (perms == want_p).all()
    assert (st["lsap_steps"] > 0) == (kw["reorder"] is True)
    full, _ = eng.rmsd_rows(0, 24)
    for (i, j), v in zip(pairs, vals):
        assert full[condensed_index(24, i, j)] == v


def test_hungarian_config4_shape_sample(eng):
    """BASELINE config 4 shape (60 solute + 300 solvent atoms, -ns 60 -e) on a few frames."""
    from clusttraj_b200.synth import make_solute_solvent

    atoms, coords = make_solute_solvent(12, n_solute=60, n_solvent_mol=100, seed=20250604)
    kw = dict(reorder=True, ns=60)
    pairs = [(0, 1), (0, 7), (3, 4), (5, 11), (10, 11)]
    eng.load_frames(coords, atoms, engine_opts(**kw))
    vals, perms, _ = eng.pair_values(pairs, want_perms=True)
    want_v, want_p = oracle_pairs(atoms, coords, pairs, **kw)
    assert np.abs(vals - want_v).max() <= 1e-9
    assert (perms == want_p).all()


def test_hungarian_large_block_generic_solver(eng):
    """An element block above 256 atoms takes the shared-memory-state solver instead of the register one."""
    rng = np.random.default_rng(4)
    base = rng.uniform(-12, 12, size=(300, 3))
    frames = np.stack([(base + rng.normal(0, 0.25, size=base.shape))[rng.permutation(300)] for _ in range(4)])
    atoms = np.full(300, 6, np.int32)
    pairs = [(0, 1), (0, 3), (1, 2), (2, 3)]
    eng.load_frames(frames, atoms, engine_opts(reorder=True))
    vals, perms, _ = eng.pair_values(pairs, want_perms=True)
    want_v, want_p = oracle_pairs(atoms, frames, pairs, reorder=True)
    assert (perms == want_p).all()
    assert np.abs(vals - want_v).max() <= 1e-9


def test_row_slabs_concatenate(eng):
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(500, 40, seed=3)
    full, _ = _full(eng, atoms, coords)
    parts = [eng.rmsd_rows(a, b)[0].copy() for a, b in [(0, 17), (17, 64), (64, 65), (65, 333), (333, 500)]]
    assert np.array_equal(np.concatenate(parts), full)
    dev_ptr, n, _ = eng.rmsd_rows_device(100, 200)
    assert dev_ptr and n == eng.rows_pairs(100, 200)
    lo = condensed_index(500, 100, 101)
    assert np.array_equal(eng.copy_slab(0, n), full[lo: lo + n])


def test_streaming_chunks_match_single_chunk():
    from clusttraj_b200 import _engine
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(900, 30, seed=8)
    os.environ["CLUSTTRAJ_B200_CHUNK_ENTRIES"] = "20000"
    try:
        chunky = _engine.Engine(0)
    finally:
        del os.environ["CLUSTTRAJ_B200_CHUNK_ENTRIES"]
    one = _engine.Engine(0)
    for e in (chunky, one):
        e.load_frames(coords, atoms, engine_opts())
    a, sa = chunky.rmsd_rows(0, 900, _engine.pinned_empty(900 * 899 // 2))
    b, sb = one.rmsd_rows(0, 900)
    assert sa["launches"] > sb["launches"]
    assert np.array_equal(a, b)
    chunky.close(); one.close()


def test_identical_cluster_labels(eng):
    """SciPy linkage on the engine's matrix gives the labels the oracle's matrix gives."""
    import scipy.cluster.hierarchy as hcl
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(160, 30, seed=21, n_modes=5)
    got, _ = _full(eng, atoms, coords)
    want = oracle_rows(atoms, coords, range(160), n_workers=4)
    assert np.abs(got - want).max() <= ktol(eng)
    for method in ("average", "ward", "complete"):
        Zg, Zw = hcl.linkage(got, method), hcl.linkage(want, method)
        for t in (0.5, 1.0, 2.0):
            assert (hcl.fcluster(Zg, t, criterion="distance") == hcl.fcluster(Zw, t, criterion="distance")).all()


def test_config2_full_size_properties(eng):
    """BASELINE config 2 (20k frames x 100 atoms, plain) through size-independent properties."""
    from clusttraj_b200.synth import make_trajectory

    M, N = 20000, 100
    atoms, coords = make_trajectory(M, N, seed=20250602)
    eng.load_frames(coords, atoms, engine_opts())
    out, st = eng.rmsd_rows(0, M, __import__("clusttraj_b200")._engine.pinned_empty(M * (M - 1) // 2))
    assert st["pairs"] == M * (M - 1) // 2
    assert np.isfinite(out).all() and out.min() >= 0.0
    # duplicates of frame 0 (frames 1, 2) up to the 5-decimal rounding; mirrored last frame is NOT a duplicate
    assert out[condensed_index(M, 0, 1)] < 5e-5 and out[condensed_index(M, 1, 2)] < 5e-5
    assert out[condensed_index(M, 0, M - 1)] > 0.1
    # rows against a duplicated reference frame are identical up to the rounding of the input
    r0 = out[condensed_index(M, 0, 3): condensed_index(M, 0, 3) + 2000]
    r1 = out[condensed_index(M, 1, 3): condensed_index(M, 1, 3) + 2000]
    assert np.abs(r0 - r1).max() < 1e-4
    # sampled entries against the oracle
    rng = np.random.default_rng(0)
    pairs = [(int(min(a, b)), int(max(a, b))) for a, b in rng.integers(0, M, size=(300, 2)) if a != b]
    want, _ = oracle_pairs(atoms, coords, pairs)
    got = np.array([out[condensed_index(M, i, j)] for i, j in pairs])
    tol = ktol(eng)
    assert st["gram_path"] == eng.expected_path
    assert np.abs(got - want).max() <= tol
    # triangle inequality of the minimum RMSD on sampled triples
    tri = rng.integers(0, M, size=(2000, 3))
    for a, b, c in tri:
        a, b, c = sorted((int(a), int(b), int(c)))
        if a == b or b == c:
            continue
        ab, bc, ac = (out[condensed_index(M, a, b)], out[condensed_index(M, b, c)], out[condensed_index(M, a, c)])
        assert ac <= ab + bc + 3 * tol and ab <= ac + bc + 3 * tol and bc <= ab + ac + 3 * tol
    # invariance: the same frames rigidly moved give the same matrix rows
    th = 1.1
    Rx = np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]])
    sub = coords[:2048] @ Rx.T + np.array([5.0, -3.0, 2.0])
    eng.load_frames(sub, atoms, engine_opts())
    moved, _ = eng.rmsd_rows(0, 64)
    eng.load_frames(coords[:2048], atoms, engine_opts())
    still, _ = eng.rmsd_rows(0, 64)
    assert np.abs(moved - still).max() < 2 * tol


def _sampled_device_check(eng, atoms, coords, kw, n_samples, tol, seed=0):
    """Whole matrix left on the GPU; random entries copied back and compared with the oracle."""
    M = len(coords)
    eng.load_frames(coords, atoms, engine_opts(**kw))
    if eng.stats()["gram_path"]:
        tol = max(tol, ktol(eng))
    ptr, n, st = eng.rmsd_rows_device(0, M)
    assert n == M * (M - 1) // 2 and st["pairs"] == n
    rng = np.random.default_rng(seed)
    pairs = []
    while len(pairs) < n_samples:
        a, b = (int(v) for v in rng.integers(0, M, size=2))
        if a != b:
            pairs.append((min(a, b), max(a, b)))
    pairs += [(0, 1), (1, 2), (0, M - 1), (M - 2, M - 1)]
    got = np.array([eng.copy_slab(condensed_index(M, i, j), 1)[0] for i, j in pairs])
    want, _ = oracle_pairs(atoms, coords, pairs, **kw)
    assert np.abs(got - want).max() <= tol, np.abs(got - want).max()
    # a strided sweep over the slab: finite, non-negative, no fix-up marker left behind
    stride = max(1, n // 64)
    for off in range(0, n - 4096, stride):
        blk = eng.copy_slab(off, 4096)
        assert np.isfinite(blk).all() and blk.min() >= 0.0
    return st


def test_config3_full_size(eng):
    """BASELINE config 3: 50k frames x 300 atoms, -n (K = 150), 1.25e9 pairs, slab kept on the GPU."""
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(50000, 300, seed=20250603)
    st = _sampled_device_check(eng, atoms, coords, dict(noh=True), 120, 1e-9)
    assert st["kept_atoms"] == 150 and st["flagged"] >= 3


def test_config5_full_size_single_gpu(eng):
    """BASELINE config 5's matrix (100k frames x 300 atoms, plain, 5e9 pairs = 40 GB) on ONE GPU."""
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(100000, 300, seed=20250605)
    st = _sampled_device_check(eng, atoms, coords, dict(), 120, 1e-9)
    assert st["pairs"] == 4999950000


def test_config4_full_size_hungarian(eng):
    """BASELINE config 4: 10k frames, 60 solute + 300 solvent atoms, -ns 60 -e hungarian (5e7 pairs).
    Frames 1..3 are frame 0 with its solvent atoms relabelled (same-element permutation) plus 1e-4 A noise:
    the reorder must undo the relabelling (known permutation) and give an RMSD at the noise level."""
    from clusttraj_b200.synth import make_solute_solvent

    atoms, coords = make_solute_solvent(10000, n_solute=60, n_solvent_mol=100, seed=20250604)
    rng = np.random.default_rng(1)
    known = []
    for k in (1, 2, 3):
        perm = np.arange(360)
        for z in np.unique(atoms[60:]):
            idx = 60 + np.where(atoms[60:] == z)[0]
            perm[idx] = rng.permutation(idx)
        coords[k] = coords[0][perm] + rng.normal(0, 1e-4, size=(360, 3))
        known.append(perm)
    kw = dict(reorder=True, ns=60)
    st = _sampled_device_check(eng, atoms, coords, kw, 24, 1e-9, seed=3)
    assert st["lsap_steps"] > 300 * st["pairs"]
    vals, perms, _ = eng.pair_values([(0, 1), (0, 2), (0, 3)], want_perms=True)
    assert vals.max() < 1e-3
    for k in range(3):
        # perms[k][p] = index in frame k+1 of the atom that ends at position p; frame k+1 = frame 0[known]
        assert (known[k][perms[k]] == np.arange(360)).all()


# ---- the int8 tensor-core Gram kernel (csrc/ozaki.cu) against the FP64 DMMA kernel and the oracle ----------------

@pytest.mark.parametrize("M,N,kw", [
    (2, 5, dict()), (3, 1, dict()), (41, 7, dict()), (96, 50, dict()), (333, 100, dict()), (330, 128, dict()),
    (257, 129, dict()), (700, 130, dict()), (260, 64, dict(ns=20)), (500, 300, dict(noh=True)), (301, 300, dict()),
    (120, 400, dict()),
])
def test_int8_kernel_matches_dmma_kernel_and_oracle(M, N, kw):
    """Shapes around every boundary of the tcgen05 kernels: row tiles of 40 frames, column tiles of 32, 32-atom MMA
    chunks, the one-K-block (A in TMEM) / several-K-block (A in shared memory) switch at 128 atoms."""
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(M, N, seed=77 + M + N)
    ei, ed = _engine_with(None), _engine_with("-2")
    got, st = _full(ei, atoms, coords, **kw)
    ref, sd = _full(ed, atoms, coords, **kw)
    assert st["gram_path"] == 2 and sd["gram_path"] == 1
    q = st["quant_step"]
    assert 0 < q <= 2.0 ** -22
    assert np.isfinite(got).all() and (got >= 0).all()
    assert np.abs(got - ref).max() <= 2 * q, (np.abs(got - ref).max(), q)
    rows = sorted({0, 1, M // 2, max(M - 2, 0)})
    want = oracle_rows(atoms, coords, rows, **kw)
    pos = 0
    for r in rows:
        n = M - 1 - r
        if n == 0:
            continue
        lo = condensed_index(M, r, r + 1)
        assert np.abs(got[lo: lo + n] - want[pos: pos + n]).max() <= 2 * q
        pos += n
    ei.close(); ed.close()


def test_int8_kernel_falls_back_when_the_quantum_is_too_coarse():
    """Coordinates hundreds of Angstrom from the centroid would leave a quantum above 2^-22 A: the engine must take
    the FP64 kernel by itself (and does so for any input when the threshold is raised)."""
    from clusttraj_b200 import _engine
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(90, 40, seed=4)
    e = _engine_with(None)
    got_small, st = _full(e, atoms, coords)
    assert st["gram_path"] == 2
    big = coords * 40.0   # a "molecule" 700 A across
    got_big, st = _full(e, atoms, big)
    assert st["gram_path"] == 1 and st["quant_step"] == 0.0
    assert np.abs(got_big - 40.0 * got_small).max() <= 1e-6
    want = oracle_rows(atoms, big, [0, 45])
    assert np.abs(got_big[:89] - want[:89]).max() <= 1e-9
    e.close()
    os.environ["CLUSTTRAJ_B200_MIN_QUANT_BITS"] = "29"
    try:
        e2 = _engine.Engine(0)
    finally:
        del os.environ["CLUSTTRAJ_B200_MIN_QUANT_BITS"]
    _, st = _full(e2, atoms, coords)
    assert st["gram_path"] == 1
    e2.close()


def test_int8_kernel_error_is_a_rigid_input_perturbation():
    """The int8 path's only error is the rounding of the centred coordinates: the value must equal (to FP64 round-off)
    the oracle's value for the rounded coordinates themselves, also where G_i + G_j - 2 lambda cancels."""
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(64, 60, seed=15, noise=0.01, mode_spread=0.02)   # RMSDs of 0.01-0.05 A
    e = _engine_with(None)
    got, st = _full(e, atoms, coords)
    assert st["gram_path"] == 2
    q = st["quant_step"]
    cen = coords - coords.mean(axis=1, keepdims=True)
    rounded = np.rint(cen / q) * q
    want = oracle_rows(atoms, rounded, range(64))
    # the oracle re-centres the rounded frames (their centroid is off by < q / sqrt(N)): second-order effect.
    # Near-duplicates (frames 1, 2 copy frame 0) are re-done by the fix-up kernel from the UNrounded coordinates.
    exact = oracle_rows(atoms, coords, range(64))
    far = exact > 1e-3
    assert far.sum() >= 2000
    # (left over: the dropped weight-1 digit product and FP64 cancellation, both ~1e-10 at RMSD 0.01 A; q is ~1.5e-8)
    assert np.abs(got - want)[far].max() <= 1e-9 < q / 4, np.abs(got - want)[far].max()
    assert np.abs(got - exact)[~far].max() <= 1e-9
    assert np.abs(got - exact).max() <= 2 * q
    e.close()


@pytest.mark.parametrize("K", [100, 300])
def test_int8_kernel_accuracy_at_the_edge_of_its_acceptance_window(K):
    """The edge of the automatic int8 choice: coordinates just under 256 A from the centroid (the coarsest quantum the engine
    accepts, 2^-22 A = 2.4e-7 A) and near-duplicate frames whose cancelling E = G_i + G_j - 2 lambda straddles the flag
    threshold 1e-9 (G_i + G_j): pairs just above it are NOT re-done by the FP64 fix-up and carry the whole quantisation
    error amplified by 1 / (2 K RMSD).  Every pair (exact duplicates included) must still be within the north star's
    1e-6 A of the oracle's value for the UNROUNDED coordinates."""
    from clusttraj_b200.synth import _random_rotations

    rng = np.random.default_rng(300 + K)
    M = 56
    base = rng.normal(size=(K, 3))
    base *= (250.0 * rng.uniform(0.0, 1.0, size=(K, 1)) ** (1.0 / 3.0)) / np.linalg.norm(base, axis=1, keepdims=True)   # ball
    base -= base.mean(axis=0)
    base *= 254.0 / np.linalg.norm(base, axis=1).max()      # the farthest atom 254 A from the centroid, whatever the rotation
    g_half = 0.5 * (base ** 2).sum()
    # noise per frame such that pair RMSDs run from 0.4x to 2.5x the threshold RMSD sqrt(1e-9 * 2 (G_i + G_j) / 2 / K)
    thr = np.sqrt(1e-9 * 4.0 * g_half / K)
    sig = thr * np.linspace(0.25, 1.6, M) / np.sqrt(3.0)
    frames = base[None] + rng.normal(size=(M, K, 3)) * sig[:, None, None]
    frames[1] = frames[0]                      # exact duplicates (one left unrotated)
    frames[2] = frames[0]
    R = _random_rotations(rng, M)
    R[1] = np.eye(3)
    frames = np.einsum("mnk,mjk->mnj", frames, R)
    atoms = np.full(K, 6, dtype=np.int32)
    e = _engine_with(None)
    got, st = _full(e, atoms, frames)
    assert st["gram_path"] == 2 and st["quant_step"] == 2.0 ** -22, st      # int8 kernel at its coarsest accepted quantum
    want = oracle_rows(atoms, frames, range(M), n_workers=4)
    ratio = (want ** 2) * K / (4.0 * g_half) / 1e-9                          # E / (1e-9 (G_i + G_j)) per pair
    just_above = (ratio > 1.0) & (ratio < 1.5)
    assert just_above.sum() >= 50 and (ratio < 1.0).sum() >= 50, (just_above.sum(), (ratio < 1.0).sum())
    assert 0 < st["flagged"] < len(want)                                    # both sides of the threshold were exercised
    err = np.abs(got - want)
    assert err.max() <= 1e-6, (err.max(), err.max() / st["quant_step"])
    assert err[just_above].max() <= 1e-6 and got[0] <= 1e-6 and got[1] <= 1e-6   # (0,1), (0,2): duplicates
    e.close()


# ---- consumers of the matrix: sum_distmat / find_medoids_from_clusters (clusttraj/classify.py:146-177) -----------------

def _consumer_cases(golden_dir):
    import json

    with open(os.path.join(golden_dir, "reference_consumers.json")) as fh:
        data = json.load(fh)
    mats = {k: np.array([float.fromhex(h) for h in v]) for k, v in data["matrices"].items()}
    return [(name, mats[c["matrix"]], np.array(c["clusters"]), np.array(c["medoids"]), float.fromhex(c["sum_hex"]))
            for name, c in data["cases"].items()]


def test_consumers_against_unmodified_reference(golden_dir):
    """The package's classify entry points (same names and arguments as the reference's) on the vectors the unmodified
    clusttraj/classify.py produced."""
    from clusttraj_b200.classify import find_medoids_from_clusters, sum_distmat

    for name, d, clusters, medoids, total in _consumer_cases(golden_dir):
        got = find_medoids_from_clusters(d, clusters)
        assert got.dtype.kind == "i" and (got == medoids).all(), (name, got, medoids)
        assert abs(sum_distmat(d) - total) <= 1e-12 * abs(total), name


def test_consumers_on_the_device_resident_matrix():
    """Medoids and sum straight from the matrix ct_rmsd_rows_device left on the GPU (no host round trip), against the
    oracle; empty clusters and bad labels are errors like in the reference."""
    import scipy.cluster.hierarchy as hcl
    from clusttraj_b200 import _engine
    from clusttraj_b200.synth import make_trajectory
    from oracle import classify_restated as C

    M = 700
    atoms, coords = make_trajectory(M, 30, seed=44, n_modes=6)
    e = _engine_with("-2")
    e.load_frames(coords, atoms, engine_opts())
    _, n, _ = e.rmsd_rows_device(0, M)
    d = e.copy_slab(0, n)
    for method, k in (("average", 6), ("ward", 23), ("complete", 1)):
        clusters = hcl.fcluster(hcl.linkage(d, method), k, criterion="maxclust")
        total, med, st = e.matrix_consumers(M, clusters=clusters)        # distmat=None: the resident matrix
        want = C.find_medoids_from_clusters(d, clusters)
        if not (med == want).all():   # only a floating-point tie between two candidates may differ
            S = C.member_sums(d, clusters)
            assert np.abs(S[med] - S[want]).max() <= 1e-12 * S.max()
        assert abs(total - C.sum_distmat(d)) <= 1e-12 * total
        assert st["h2d_ms"] < 1.0   # nothing was uploaded
    with pytest.raises(_engine.EngineError, match="must lie in 1..n_clusters"):
        # 2 distinct labels but no cluster "2": the reference's np.argmin raises on the empty cluster (classify.py:164)
        e.matrix_consumers(M, clusters=np.where(np.arange(M) < 5, 1, 3))
    e.load_frames(coords[:50], atoms, engine_opts())
    e.rmsd_rows_device(0, 50)
    with pytest.raises(_engine.EngineError, match="no whole matrix"):
        e.matrix_consumers(M, clusters=np.ones(M, int))
    e.close()


def test_consumers_config2_size():
    """20k frames (2.0e8 entries, 1.6 GB): one pass over the vector on the GPU; the reference's squareform would need
    a 3.2 GB square matrix on the host.  Checked on sampled clusters against direct sums."""
    from clusttraj_b200 import _engine

    M = 20000
    rng = np.random.default_rng(3)
    d = _engine.pinned_empty(M * (M - 1) // 2)
    d[:] = rng.random(d.size)
    clusters = rng.integers(1, 41, size=M)
    clusters[:40] = np.arange(1, 41)
    e = _engine_with(None)
    total, med, st = e.matrix_consumers(M, clusters=clusters, distmat=d)
    assert abs(total - float(np.sum(d))) <= 1e-11 * total
    print(f"consumers on 2.0e8 entries: upload {st['h2d_ms']:.1f} ms, kernels {st['kernel_ms']:.2f} ms")

    def member_sum(i):
        js = np.where(clusters == clusters[i])[0]
        js = js[js != i]
        lo, hi = np.minimum(i, js), np.maximum(i, js)
        return d[M * lo - lo * (lo + 1) // 2 + (hi - lo - 1)].sum()

    for c in (1, 17, 40):
        members = np.where(clusters == c)[0]
        sums = np.array([member_sum(int(i)) for i in members])
        assert clusters[med[c - 1]] == c
        assert member_sum(int(med[c - 1])) <= sums.min() * (1 + 1e-12)
    e.close()


def test_cli_config1_end_to_end(golden_dir, tmp_path):
    """BASELINE config 1: the reference's test trajectory, -rmsd 1.0, no reorder, average linkage, through the CLI:
    matrix on the GPU -> SciPy linkage -> labels [1 2 3] -> summary with the matrix sum and medoids."""
    import shutil

    from clusttraj_b200.main import main

    traj = str(tmp_path / "testtraj.xyz")
    shutil.copy(os.path.join(golden_dir, "ref_testtraj.xyz"), traj)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        rc = main([traj, "-rmsd", "1.0", "-m", "average", "-f"])
    finally:
        os.chdir(cwd)
    assert rc == 0
    d = np.load(tmp_path / "distmat.npy")
    want = np.array([3.4161225575825, 4.370869363191556, 4.468489811370548])   # SURVEY.md 8c(2), plain mode
    assert np.abs(d - want).max() <= 1e-7
    assert (np.loadtxt(tmp_path / "clusters.dat", dtype=int) == [1, 2, 3]).all()
    text = open(tmp_path / "clusters.out").read()
    assert "3 cluster(s)" in text and "The total sum of the RMSD matrix is 12.2554" in text
    # the reference's summary text (main.py:107-135): every frame is the medoid of its own cluster
    assert "The index for the medoids are:\n0 1 2 \n\nThe cluster sizes are:\nCluster\tSize\n1\t1\n2\t1\n3\t1\n" in text


def test_cli_structure_output_native_formats_and_clean_errors(golden_dir, tmp_path):
    """ADVICE round 1 through the CLI: -cc xyz -mc xyz write the superposed structures (native writers); .pdb and .gro
    trajectories run without OpenBabel; -nc beyond the number of snapshots and a matrix with NaNs (-i) end in the
    reference's clean SystemExit (main.py:53-66; SciPy's ValueError for non-finite distances) instead of a traceback."""
    import shutil

    from clusttraj_b200.main import main

    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        traj = str(tmp_path / "testtraj.xyz")
        shutil.copy(os.path.join(golden_dir, "ref_testtraj.xyz"), traj)
        assert main([traj, "-rmsd", "4.4", "-m", "average", "-f", "-cc", "xyz", "-mc", "xyz", "-v"]) == 0
        clusters = np.loadtxt(tmp_path / "clusters.dat", dtype=int)
        assert list(clusters) == [1, 1, 2]        # Z = [[0,1,3.416..],[2,3,4.4197..]] cut at 4.4
        from clusttraj_b200.xyz import read_xyz

        a1, c1 = read_xyz(str(tmp_path / "clusters_confs_1.xyz"))
        a2, c2 = read_xyz(str(tmp_path / "clusters_confs_2.xyz"))
        am, cm = read_xyz(str(tmp_path / "clusters_medoids.xyz"))
        assert c1.shape == (2, 41, 3) and c2.shape == (1, 41, 3) and cm.shape == (2, 41, 3)
        # the members of cluster 1 are superposed on its medoid: their plain Kabsch RMSD is the matrix entry (0, 1)
        d = np.load(tmp_path / "distmat.npy")
        assert abs(np.sqrt(((c1[0] - c1[1]) ** 2).sum() / 41) - d[0]) < 2e-5    # files hold 5 decimals
        with pytest.raises(SystemExit, match="Cannot request 5 clusters from 3 snapshots"):
            main([traj, "-nc", "5", "-f"])
        bad = d.copy()
        bad[1] = np.nan
        np.save(tmp_path / "bad.npy", bad)
        with pytest.raises(SystemExit, match="must contain only finite values"):
            main([traj, "-rmsd", "1.0", "-f", "-i", str(tmp_path / "bad.npy")])
        for name, n_frames in (("hand_elements.pdb", 2), ("hand_noelements.pdb", 2), ("hand_water.gro", 2)):
            src = str(tmp_path / name)
            shutil.copy(os.path.join(golden_dir, name), src)
            assert main([src, "-rmsd", "0.01", "-f", "-od", str(tmp_path / (name + ".npy"))]) == 0
            dd = np.load(tmp_path / (name + ".npy"))
            assert dd.shape == (n_frames * (n_frames - 1) // 2,) and 0.0 < dd[0] < 1.0
            assert list(np.loadtxt(tmp_path / "clusters.dat", dtype=int)) == [1, 2]
    finally:
        os.chdir(cwd)


def test_non_finite_coordinates_take_the_fp64_kernel():
    """A NaN coordinate has no fixed-point image: the engine must not quantise (it would turn NaN into a number) but
    run the FP64 path, where the NaN reaches every value of that frame as it does in the reference."""
    from clusttraj_b200 import _engine
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(50, 20, seed=2)
    coords[7, 3, 1] = np.nan
    e = _engine_with(None)
    got, st = _full(e, atoms, coords)
    assert st["gram_path"] == 1
    bad = np.array([condensed_index(50, min(7, j), max(7, j)) for j in range(50) if j != 7])
    assert np.isnan(got[bad]).all()
    mask = np.ones(got.size, bool); mask[bad] = False
    assert np.isfinite(got[mask]).all()
    e.close()
    forced = _engine_with("10")
    with pytest.raises(_engine.EngineError, match="not finite"):
        forced.load_frames(coords, atoms, engine_opts())
    forced.close()


@pytest.mark.parametrize("kind", ["collinear", "planar", "two_atoms", "one_atom", "random_large_rmsd", "tiny_scale", "huge_scale"])
def test_degenerate_geometries_on_both_gram_kernels(kind):
    """Rank-deficient covariances (collinear / planar / 1-2 atoms: double roots of the quartic), structures with nothing in
    common (start of the Newton iteration far above the root) and extreme length scales, on both Gram kernels."""
    rng = np.random.default_rng(sum(kind.encode()))   # deterministic (str hashes are salted per process)
    M = 57
    if kind == "collinear":
        t = rng.normal(size=(M, 23, 1))
        frames = t * rng.normal(size=(M, 1, 3)) + rng.normal(size=(M, 1, 3))
    elif kind == "planar":
        uv = rng.normal(size=(M, 31, 2))
        basis = rng.normal(size=(M, 2, 3))
        frames = np.einsum("mnk,mkj->mnj", uv, basis) * 4.0
    elif kind == "two_atoms":
        frames = rng.normal(size=(M, 2, 3)) * 3.0
    elif kind == "one_atom":
        frames = rng.normal(size=(M, 1, 3)) * 3.0
    elif kind == "random_large_rmsd":
        frames = rng.normal(size=(M, 140, 3)) * 8.0
    elif kind == "tiny_scale":
        frames = rng.normal(size=(M, 40, 3)) * 1e-4
    else:
        frames = rng.normal(size=(M, 40, 3)) * 50.0    # stays inside the int8 path's 256 A range
    frames = np.ascontiguousarray(frames)
    atoms = np.full(frames.shape[1], 6, np.int32)
    want = oracle_rows(atoms, frames, range(M))
    scale = max(float(np.abs(want).max()), 1e-300)
    for variant, path in ((None, 2), ("-2", 1)):
        e = _engine_with(variant)
        got, st = _full(e, atoms, frames)
        assert st["gram_path"] == path
        assert np.isfinite(got).all() and (got >= 0).all()
        tol = max(1e-9 * max(scale, 1.0), 2.0 * st["quant_step"]) if kind != "tiny_scale" else 1e-12
        assert np.abs(got - want).max() <= tol, (kind, variant, np.abs(got - want).max(), tol, st["flagged"])
        e.close()


# ---- structure output (clusttraj/io.py:582-1015): transforms from the per-pair kernel --------------------------------

def _align_cases(golden_dir):
    import json

    with open(os.path.join(golden_dir, "reference_align.json")) as fh:
        data = json.load(fh)
    for ds in data.values():
        for mode in ds["modes"].values():
            mode["aligned"] = np.array([[[float.fromhex(h) for h in row] for row in fr] for fr in mode["aligned_hex"]])
    return data


def test_superposed_frames_against_unmodified_align_mol(golden_dir):
    """Every later frame of the fixtures superposed on frame 0, in every mode align_mol distinguishes, against what the
    unmodified clusttraj.io.align_mol printed (tests/golden/reference_align.json)."""
    from clusttraj_b200.io import superposed_frames
    from clusttraj_b200.reorder import reorder_hungarian

    worst = 0.0
    for ds in _align_cases(golden_dir).values():
        traj = os.path.join(golden_dir, ds["trajfile"])
        for name, mode in ds["modes"].items():
            noh, do_reorder, rs, ns, ws, excl, fk = mode["args"]
            n = len(mode["aligned"])
            atoms, ref, moved = superposed_frames(traj, 0, range(1, n + 1), noh, reorder_hungarian if do_reorder else None,
                                                  rs, ns, ws, np.asarray(excl, np.int32), fk)
            err = np.abs(moved - mode["aligned"]).max()
            worst = max(worst, err)
            assert err <= 1e-9, (name, err)
    print(f"superposed structures: worst |dx| = {worst:.2e} A")


def test_save_clusters_and_medoids_config(golden_dir, tmp_path):
    """The two writers on a clustered synthetic trajectory: medoid first, members in trajectory order, every member
    closer to its medoid after superposition than the matrix says (same RMSD for the plain mode), no overwrite."""
    from clusttraj_b200 import xyz
    from clusttraj_b200.classify import find_medoids_from_clusters
    from clusttraj_b200.io import save_clusters_config, save_medoids_config
    from clusttraj_b200.synth import make_trajectory
    import scipy.cluster.hierarchy as hcl

    atoms, coords = make_trajectory(40, 18, seed=12, n_modes=3, n_duplicates=0, mirrored=False)
    traj = str(tmp_path / "t.xyz")
    xyz.write_xyz(traj, atoms, coords)
    e = _engine_with(None)
    d, _ = _full(e, atoms, xyz.read_xyz(traj)[1])
    e.close()
    clusters = hcl.fcluster(hcl.linkage(d, "average"), 3, criterion="maxclust")
    base = str(tmp_path / "confs")
    save_clusters_config(traj, clusters, d, False, None, False, None, None, base, "xyz", np.asarray([], np.int32), False, False)
    medoids = find_medoids_from_clusters(d, clusters)
    for c in range(1, 4):
        a, x = xyz.read_xyz(f"{base}_{c}.xyz")
        members = [int(i) for i in np.where(clusters == c)[0] if i != medoids[c - 1]]
        assert len(x) == 1 + len(members) and (a == atoms).all()
        assert np.abs(x[0].mean(axis=0)).max() < 1e-5          # the medoid is written centred
        for k, m in enumerate(members):
            i, j = sorted((int(medoids[c - 1]), m))
            direct = np.sqrt(((x[1 + k] - x[0]) ** 2).sum() / len(atoms))
            assert abs(direct - d[condensed_index(40, i, j)]) < 2e-5   # 5 decimals in the file
    with pytest.raises(IOError):
        save_clusters_config(traj, clusters, d, False, None, False, None, None, base, "xyz", np.asarray([], np.int32), False, False)
    save_medoids_config(traj, medoids, False, None, False, None, None, str(tmp_path / "med"), "xyz", np.asarray([], np.int32), False, True)
    a, x = xyz.read_xyz(str(tmp_path / "med.xyz"))
    assert len(x) == 3


# ---- hierarchical clustering on the GPU (the hcl.linkage call at clusttraj/classify.py:30,99,127) -------------------

LINKAGE_METHODS = ["single", "complete", "average", "weighted", "ward"]


def _linkage_inputs():
    from scipy.spatial.distance import pdist

    rng = np.random.default_rng(77)
    cases = {}
    for M in (2, 3, 7, 64, 257, 1000):
        cases[f"random_{M}"] = rng.random(M * (M - 1) // 2) * 5.0
    cases["ties_120"] = np.round(rng.random(120 * 119 // 2) * 3.0, 1)              # many equal distances
    x = rng.normal(size=(90, 3))
    x[45:] = x[:45]                                                               # exact duplicates: zero distances
    cases["duplicates_90"] = pdist(x)
    cases["euclid_500"] = pdist(rng.normal(size=(500, 6)))
    cases["all_equal_33"] = np.full(33 * 32 // 2, 1.25)
    return cases


@pytest.mark.parametrize("variant", ["cluster", "graph"])
@pytest.mark.parametrize("method", LINKAGE_METHODS)
def test_linkage_is_scipys_linkage_bit_for_bit(method, variant, monkeypatch):
    """Z from the GPU == scipy.cluster.hierarchy.linkage(distmat, method): same merges in the same order, identical
    heights to the last bit, on random, tie-laden, duplicate-laden and constant matrices — for the persistent
    single-cluster kernel (default) and for the whole-GPU step-launch variant."""
    import scipy.cluster.hierarchy as hcl
    from clusttraj_b200.classify import linkage

    monkeypatch.setenv("CLUSTTRAJ_B200_LINKAGE", variant)
    for name, d in _linkage_inputs().items():
        want = hcl.linkage(d, method)
        got = linkage(d, method)
        assert got.shape == want.shape and got.dtype == np.float64
        assert got.tobytes() == want.tobytes(), (method, name, np.abs(got - want).max(), np.argwhere(got != want)[:3])
        assert hcl.is_valid_linkage(got)


def test_linkage_on_the_device_resident_matrix_and_downstream():
    """RMSD matrix -> linkage -> medoids without the matrix leaving the GPU; the matrix is left intact (the chain methods
    work on a copy); optimal ordering, centroid/median (SciPy) and the classify_* entry points behave as the reference's."""
    import types

    import scipy.cluster.hierarchy as hcl
    from clusttraj_b200 import _engine
    from clusttraj_b200.classify import classify_structures, classify_structures_nclusters, linkage
    from clusttraj_b200.synth import make_trajectory

    M = 900
    atoms, coords = make_trajectory(M, 40, seed=45, n_modes=7)
    e = _engine_with(None)
    e.load_frames(coords, atoms, engine_opts())
    _, n, _ = e.rmsd_rows_device(0, M)
    d = e.copy_slab(0, n)
    for method in LINKAGE_METHODS:
        Z, st = e.linkage(M, method)                       # distmat=None: the resident matrix
        assert Z.tobytes() == hcl.linkage(d, method).tobytes(), method
        assert st["h2d_ms"] < 1.0 and st["launches"] >= M - 1
    assert e.copy_slab(0, n).tobytes() == d.tobytes()      # untouched
    total, med, _ = e.matrix_consumers(M, clusters=hcl.fcluster(Z, 7, criterion="maxclust"))
    assert abs(total - d.sum()) <= 1e-12 * total and len(med) == 7
    with pytest.raises(ValueError, match="does not run on the GPU"):
        e.linkage(M, "centroid")
    e.load_frames(coords[:50], atoms, engine_opts())
    e.rmsd_rows_device(0, 50)
    with pytest.raises(_engine.EngineError, match="no whole matrix"):
        e.linkage(M, "average")
    e.close()
    # host entry points
    small = d[: 200 * 199 // 2]
    assert linkage(small, "average", optimal_ordering=True).tobytes() == hcl.linkage(small, "average", optimal_ordering=True).tobytes()
    assert linkage(small, "centroid").tobytes() == hcl.linkage(small, "centroid").tobytes()
    with pytest.raises(ValueError):
        linkage(d[:-1], "average")
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        opt = types.SimpleNamespace(method="ward", opt_order=False, min_rmsd=2.0, n_clusters=5,
                                    out_clust_name=os.path.join(tmp, "clusters.dat"))
        Zw = hcl.linkage(d, "ward")
        Z1, c1 = classify_structures(opt, d)
        assert Z1.tobytes() == Zw.tobytes() and (c1 == hcl.fcluster(Zw, 2.0, criterion="distance")).all()
        assert (np.loadtxt(opt.out_clust_name, dtype=int) == c1).all()
        Z2, c2 = classify_structures_nclusters(opt, d)
        assert (c2 == hcl.cut_tree(Zw, n_clusters=[5]).ravel() + 1).all() and c2.max() == 5
        opt.n_clusters = M + 1
        with pytest.raises(ValueError, match="Cannot request"):
            classify_structures_nclusters(opt, d)


def test_linkage_config2_size():
    """20000 observations (the cfg2 matrix size): Z identical to SciPy's for average linkage, and much faster."""
    import time

    import scipy.cluster.hierarchy as hcl
    from clusttraj_b200.synth import make_trajectory

    M = 20000
    atoms, coords = make_trajectory(M, 30, seed=46, n_modes=12)
    e = _engine_with(None)
    e.load_frames(coords, atoms, engine_opts())
    _, n, _ = e.rmsd_rows_device(0, M)
    Z, st = e.linkage(M, "average")
    d = e.copy_slab(0, n)
    e.close()
    t0 = time.perf_counter()
    want = hcl.linkage(d, "average")
    cpu_s = time.perf_counter() - t0
    assert Z.tobytes() == want.tobytes()
    assert st["kernel_ms"] * 1e-3 < cpu_s, (st["kernel_ms"], cpu_s)


def test_classify_on_golden_vector(tmp_path, ref_golden_distmat):
    """Reference test_classify.py:30-52: ward linkage of the golden matrix, labels at -rmsd 1.0 and for -nc 2."""
    from clusttraj_b200.io import ClustOptions
    from clusttraj_b200.main import classify

    opt = ClustOptions(method="ward", min_rmsd=1.0, opt_order=False, out_clust_name=str(tmp_path / "c.dat"))
    Z, clusters = classify(opt, ref_golden_distmat)
    assert np.allclose(Z, [[0.0, 1.0, 2.94126393, 2.0], [2.0, 3.0, 4.6348889, 3.0]])
    assert list(clusters) == [1, 2, 3]
    opt.update({"n_clusters": 2})
    assert list(classify(opt, ref_golden_distmat)[1]) == [1, 1, 2]


def test_linkage_abi_argument_errors():
    """ct_linkage refuses what SciPy would raise on (fewer than two observations) and what it does not implement
    (method codes of centroid / median), with the engine's error string instead of an exception across the ABI."""
    import ctypes as C

    from clusttraj_b200 import _engine

    e = _engine.Engine(0)
    lib = e._lib
    Z = np.zeros((4, 4))
    d = np.arange(10, dtype=np.float64)                 # 5 observations
    for M, method, zptr, needle in ((1, 2, Z.ctypes.data, "2 <= M"), (5, 3, Z.ctypes.data, "centroid"),
                                    (5, 4, Z.ctypes.data, "centroid"), (5, 2, None, "Z is NULL")):
        rc = lib.ct_linkage(e._h, d.ctypes.data, M, method, zptr, None)
        assert rc != 0 and needle in lib.ct_last_error(e._h).decode(), (M, method)
    assert lib.ct_linkage(e._h, d.ctypes.data, 5, 2, Z.ctypes.data, None) == 0
    import scipy.cluster.hierarchy as hcl
    assert Z.tobytes() == hcl.linkage(d, "average").tobytes()
    e.close()


def test_resident_matrix_is_reused_only_for_the_trusted_array():
    """One GPU, one shot: the matrix stays resident after condensed_matrix; clusttraj_b200.main vouches (trust_resident)
    that it does not touch the host copy, and linkage / sum / medoids then skip the 8 L-byte upload.  Any other array —
    a copy, a modified copy, the same values under another identity — is uploaded like before."""
    import scipy.cluster.hierarchy as hcl

    from clusttraj_b200 import classify
    from clusttraj_b200 import distmat as D
    from clusttraj_b200.synth import make_trajectory

    atoms, coords = make_trajectory(1500, 20, seed=9, n_modes=5)
    d = D.condensed_matrix(atoms, coords, engine_opts(), devices=[0])
    assert type(d) is np.ndarray and d.flags.writeable and d.flags.c_contiguous and d.dtype == np.float64
    want = oracle_rows(atoms, coords, [0, 700, 1498])
    assert np.abs(np.concatenate([d[:1499], d[condensed_index(1500, 700, 701):condensed_index(1500, 700, 701) + 799],
                                  d[-1:]]) - want).max() <= 1e-6
    eng = D._engine_for(0)
    Z0 = hcl.linkage(d, "average")
    labels = hcl.fcluster(Z0, 6, criterion="maxclust")
    classify.trust_resident(d)
    assert classify.linkage(d, "average").tobytes() == Z0.tobytes()
    assert eng.stats()["h2d_ms"] < 0.2                                    # nothing was uploaded (9 MB would take ~0.3 ms+)
    total = classify.sum_distmat(d)
    assert abs(total - d.sum()) <= 1e-12 * d.sum() and eng.stats()["h2d_ms"] < 0.2
    med = classify.find_medoids_from_clusters(d, labels)
    d2 = d.copy()
    d2[:1499] += 5.0                                                      # frame 0 moved away from everything
    Z2 = classify.linkage(d2, "average")                                  # not the trusted array: uploaded, honoured
    assert Z2.tobytes() == hcl.linkage(d2, "average").tobytes() and Z2.tobytes() != Z0.tobytes()
    assert abs(classify.sum_distmat(d2) - d2.sum()) <= 1e-12 * d2.sum()
    # the GPU now holds d2: the trusted array must be uploaded again, not confused with it
    assert classify.linkage(d, "average").tobytes() == Z0.tobytes()
    assert (classify.find_medoids_from_clusters(d, labels) == med).all()
    classify.trust_resident(None)
    assert classify.linkage(d, "average").tobytes() == Z0.tobytes()


def test_gather_into_a_reserved_matrix_on_one_gpu():
    """ct_matrix_reserve / ct_rmsd_rows_gather / ct_matrix_commit with two engines on the SAME GPU: each engine's row slab
    lands at its own place of the root's whole-matrix buffer (device-to-device on one GPU; over NVLink between GPUs, see
    the next test), the host copies are bit-identical to ct_rmsd_rows, and linkage / consumers then run on the gathered
    matrix without an upload."""
    import scipy.cluster.hierarchy as hcl

    from clusttraj_b200 import _engine
    from clusttraj_b200.distmat import slab_bounds
    from clusttraj_b200.synth import make_trajectory

    M = 1100
    atoms, coords = make_trajectory(M, 30, seed=77, n_modes=6)
    os.environ["CLUSTTRAJ_B200_CHUNK_ENTRIES"] = "50000"      # several chunks per slab
    try:
        root, other = _engine.Engine(0), _engine.Engine(0)
    finally:
        del os.environ["CLUSTTRAJ_B200_CHUNK_ENTRIES"]
    for e in (root, other):
        e.load_frames(coords, atoms, engine_opts())
    want, _ = root.rmsd_rows(0, M)
    want = want.copy()
    cuts = slab_bounds(M, 2)
    with pytest.raises(_engine.EngineError, match="ct_matrix_reserve"):
        other.rmsd_rows_gather(cuts[1], cuts[2], root)
    root.matrix_reserve(M)
    with pytest.raises(_engine.EngineError, match="no whole matrix"):
        root.linkage(M, "average")                             # reserved is not yet complete
    n0 = root.rows_pairs(cuts[0], cuts[1])
    a, st_a = root.rmsd_rows_gather(cuts[0], cuts[1], root)    # in place + host copy
    _, st_b = other.rmsd_rows_gather(cuts[1], cuts[2], root, to_host=False)   # device only
    assert st_a["launches"] > 2 and st_b["pairs"] == len(want) - n0
    assert np.array_equal(a, want[:n0])
    root.matrix_commit(M)
    assert np.array_equal(root.copy_slab(0, len(want)), want)
    Z, st = root.linkage(M, "average")
    assert Z.tobytes() == hcl.linkage(want, "average").tobytes() and st["h2d_ms"] < 0.5
    total, _, _ = root.matrix_consumers(M)
    assert abs(total - want.sum()) <= 1e-12 * want.sum()
    with pytest.raises(_engine.EngineError, match="reserved"):
        other.matrix_commit(M)
    root.close(); other.close()


def test_per_frame_element_lists(tmp_path):
    """Frames that list their atoms differently (reference distmat.py:131-141 reads each frame's own atoms): the drop-in entry
    points against the oracle run on the per-frame lists; what the engine cannot describe once per trajectory is refused."""
    from clusttraj_b200 import distmat as D
    from clusttraj_b200.io import ClustOptions
    from clusttraj_b200.reorder import reorder_hungarian
    from clusttraj_b200.synth import make_trajectory

    rng = np.random.default_rng(11)
    atoms0, coords = make_trajectory(12, 20, seed=5)
    M, N = coords.shape[:2]
    sym = {1: "H", 6: "C", 7: "N", 8: "O"}
    heavy = np.array([6, 7, 8], dtype=np.int32)
    none = np.asarray([], np.int32)

    def write(path, atoms2d):
        with open(path, "w") as fh:
            for m in range(M):
                fh.write(f"{N}\nframe {m}\n")
                for z, (x, y, z3) in zip(atoms2d[m], coords[m]):
                    fh.write(f"{sym[int(z)]} {x:.5f} {y:.5f} {z3:.5f}\n")

    # (1) the heavy atoms change element from frame to frame, the hydrogens stay where they are
    a1 = np.tile(atoms0, (M, 1))
    for m in range(1, M):
        hv = np.where(a1[m] != 1)[0]
        a1[m, hv] = rng.choice(heavy, size=len(hv))
    # (2) the hydrogens move around too (same count): -n keeps different positions in different frames
    a2 = a1.copy()
    for m in range(1, M):
        a2[m] = a2[m][rng.permutation(N)]
    for tag, atoms2d in (("a1", a1), ("a2", a2)):
        path = str(tmp_path / f"{tag}.xyz")
        write(path, atoms2d)
        for noh in (True, False):
            opt = ClustOptions(trajfile=path, no_hydrogen=noh, reorder=False, reorder_alg=None)
            got = D.build_distance_matrix(opt)
            want = oracle_rows(atoms2d, coords, range(M), noh=noh)
            assert got.shape == want.shape and np.abs(got - want).max() < 1e-8, (tag, noh)
        line = D.compute_distmat_line(2, (atoms2d[2], coords[2]), path, True, None, False, None, None, none, False)
        want = oracle_rows(atoms2d, coords, [2], noh=True)
        assert np.abs(np.asarray(line) - want).max() < 1e-8
        # -ns -n: the kept solute atoms are counted per frame; (1) keeps the count, so it runs
        if tag == "a1":
            line = D.compute_distmat_line(3, None, path, True, None, False, 6, None, none, False)
            want = oracle_rows(atoms2d, coords, [3], noh=True, ns=6)
            assert np.abs(np.asarray(line) - want).max() < 1e-8
    # a reorder over frames whose element COUNTS differ has no defined result (the reference's per-element assignment
    # fails on such a pair): refused by the engine
    from clusttraj_b200._engine import EngineError
    with pytest.raises(EngineError, match="element counts"):
        D.compute_distmat_line(0, None, str(tmp_path / "a1.xyz"), False, reorder_hungarian, False, None, None, none, False)
    # ... frames with different numbers of kept atoms have no RMSD (the reference's Kabsch fails on the shapes)
    a4 = a1.copy()
    a4[5, np.where(a4[5] != 1)[0][0]] = 1
    path = str(tmp_path / "a4.xyz")
    write(path, a4)
    with pytest.raises(ValueError):
        D.compute_distmat_line(0, None, path, True, None, False, None, None, none, False)


@pytest.mark.parametrize("mode", ["e", "ns_e", "ns_e_noh", "e_distance", "ns_rs_excl"])
def test_reorder_with_per_frame_element_lists(tmp_path, mode):
    """Frames that list the SAME atoms in different orders (reference distmat.py:131-141, 192, 243: the element blocks of a
    reorder come from each frame's own atoms): ct_set_frame_elements gives every frame its own block index lists.  Values
    and composite permutations against the oracle run on the per-frame lists."""
    from clusttraj_b200 import _engine
    from clusttraj_b200 import distmat as D
    from clusttraj_b200.reorder import reorder_distance, reorder_hungarian
    from clusttraj_b200.synth import make_solute_solvent

    rng = np.random.default_rng(23)
    ns = 9
    atoms0, coords = make_solute_solvent(7, n_solute=ns, n_solvent_mol=8, seed=31, box=6.0)
    M, N = coords.shape[:2]
    atoms2d = np.tile(atoms0, (M, 1))
    coords = coords.copy()
    kw = {"e": dict(reorder=True), "ns_e": dict(reorder=True, ns=ns), "ns_e_noh": dict(reorder=True, ns=ns, noh=True),
          "e_distance": dict(reorder="distance"), "ns_rs_excl": dict(reorder=True, ns=ns, rs=True, excl=(ns + 1, ns + 4))}[mode]
    fixed = set(kw.get("excl", ()))   # excluded positions keep their atoms (exclusions are positional)
    for m in range(1, M):   # shuffle the atom order inside the solute and inside the solvent, frame by frame
        solv = np.array([k for k in range(ns, N) if k not in fixed])
        p = np.arange(N)
        p[:ns] = rng.permutation(ns)
        p[solv] = rng.permutation(solv)
        atoms2d[m] = atoms2d[m][p]
        coords[m] = coords[m][p]
    want = oracle_rows(atoms2d, coords, range(M), **kw)
    handle = reorder_distance if kw["reorder"] == "distance" else reorder_hungarian
    a, c, noh, nsat, fz = D.frame_invariant_view(atoms2d, coords, kw.get("noh", False), handle, kw.get("ns"))
    assert fz is not None and fz.shape == (M, c.shape[1])
    opts = D._opts_from(noh, handle, kw.get("rs", False), nsat, None, np.asarray(kw.get("excl", ()), np.int32), False)
    got = D.condensed_matrix(a, c, opts, devices=[0], frame_elements=fz)
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-9, mode
    # the permutations too (pair_values on the same engine state)
    eng = _engine.Engine(0)
    eng.load_frames(c, a, opts, fz)
    pairs = [(0, 3), (2, 5), (1, 6), (4, 6)]
    vals, perms, _ = eng.pair_values(np.asarray(pairs, dtype=np.int64), want_perms=True)
    wv, wp = oracle_pairs(atoms2d, coords, pairs, **kw)
    assert np.abs(vals - wv).max() < 1e-9
    if mode != "ns_e_noh":   # (the oracle's permutation indexes the masked frame of each pair the same way)
        assert (perms == wp).all()
    else:
        assert (perms == wp).all()
    eng.close()


@pytest.mark.parametrize("M", [1, 2, 3, 41, 81])
def test_tiny_trajectories_and_empty_row_ranges(M):
    """Edge sizes (a single frame has no pairs; 41 and 81 frames cross the 40-frame row tiles of the Gram kernels by one) and
    empty / last-row ranges, on both Gram kernels and on the per-pair kernel, against the oracle."""
    from clusttraj_b200 import _engine
    from clusttraj_b200 import distmat as D
    from clusttraj_b200.synth import make_solute_solvent, make_trajectory

    atoms, coords = make_trajectory(max(M, 2), 17, seed=3)
    atoms, coords = atoms, coords[:M]
    total = M * (M - 1) // 2
    want = oracle_rows(atoms, coords, range(M)) if M > 1 else np.empty(0)
    for variant in (None, "-2"):
        eng = _engine_with(variant)
        eng.load_frames(coords, atoms, _engine.make_opts())
        got, st = eng.rmsd_rows(0, M)
        assert got.shape == (total,) and int(st["pairs"]) == total
        if total:
            assert np.abs(got - want).max() < 1e-6
        # an empty range, and the last row (which owns no entries)
        e0, _ = eng.rmsd_rows(0, 0)
        e1, _ = eng.rmsd_rows(M - 1, M)
        assert e0.size == 0 and e1.size == 0
        if M > 2:   # a ragged range in the middle
            mid, _ = eng.rmsd_rows(1, 2)
            lo = condensed_index(M, 1, 2)
            assert np.array_equal(mid, got[lo: lo + M - 2])
        eng.close()
    got = D.condensed_matrix(atoms, coords, _engine.make_opts(), devices=[0])
    assert got.shape == (total,) and (total == 0 or np.abs(got - want).max() < 1e-6)
    # the per-pair kernel (stepwise + Hungarian)
    if M <= 41:
        a2, c2 = make_solute_solvent(max(M, 2), n_solute=6, n_solvent_mol=5, seed=4, box=5.0)
        c2 = c2[:M]
        eng = _engine.Engine(0)
        eng.load_frames(c2, a2, _engine.make_opts(solute_natoms=6, reorder=True))
        got2, st2 = eng.rmsd_rows(0, M)
        assert got2.shape == (total,)
        if total:
            assert np.abs(got2 - oracle_rows(a2, c2, range(M), reorder=True, ns=6)).max() < 1e-9
        eng.close()


def test_frame_elements_abi_argument_errors():
    """ct_set_frame_elements through the binding: call order, shapes, the mask that must already be applied."""
    from clusttraj_b200 import _engine
    from clusttraj_b200.synth import make_solute_solvent

    atoms, coords = make_solute_solvent(4, n_solute=6, n_solvent_mol=4, seed=9, box=5.0)
    z2d = np.tile(atoms, (4, 1)).astype(np.int32)
    eng = _engine.Engine(0)
    rc = eng._lib.ct_set_frame_elements(eng._h, z2d.ctypes.data, 4, z2d.shape[1])
    assert rc == -3                                   # CT_ERR_STATE: nothing loaded yet
    eng.load_frames(coords, atoms, _engine.make_opts(reorder=True))
    assert eng._lib.ct_set_frame_elements(eng._h, z2d.ctypes.data, 5, z2d.shape[1]) == -2     # CT_ERR_ARG: M
    assert eng._lib.ct_set_frame_elements(eng._h, z2d.ctypes.data, 4, z2d.shape[1] - 1) == -2 # CT_ERR_ARG: K
    assert eng._lib.ct_set_frame_elements(eng._h, None, 4, z2d.shape[1]) == -2
    assert eng._lib.ct_set_frame_elements(eng._h, z2d.ctypes.data, 4, z2d.shape[1]) == 0
    with pytest.raises(_engine.EngineError, match="per-frame element lists"):
        eng.pair_transforms([(0, 1)])
    # without a reorder the elements are not used: accepted, nothing changes
    eng.load_frames(coords, atoms, _engine.make_opts())
    a, _ = eng.rmsd_rows(0, 4)
    assert eng._lib.ct_set_frame_elements(eng._h, z2d.ctypes.data, 4, z2d.shape[1]) == 0
    b, _ = eng.rmsd_rows(0, 4)
    assert np.array_equal(a, b)
    # the engine's own mask and per-frame lists do not combine: the host applies the masks (frame_invariant_view)
    eng.load_frames(coords, atoms, _engine.make_opts(reorder=True, no_hydrogen=True))
    K = eng.K
    zk = np.ascontiguousarray(np.tile(atoms[atoms != 1], (4, 1)), dtype=np.int32)
    assert zk.shape[1] == K
    with pytest.raises(_engine.EngineError, match="no_hydrogen"):
        eng._check(eng._lib.ct_set_frame_elements(eng._h, zk.ctypes.data, 4, K), "ct_set_frame_elements")
    eng.close()


def test_in_process_multi_gpu_slabs():
    """condensed_matrix(devices=[0, 1, ...]) — the in-process multi-GPU path the CLI uses (one engine + one host thread
    per GPU, equal-area row slabs): bit-identical to one GPU, and the whole matrix is left resident on the first GPU
    (gathered over NVLink) so that the clustering that follows uploads nothing.  With a single visible GPU the same code
    runs with both slabs on device 0 through two passes."""
    import scipy.cluster.hierarchy as hcl

    from clusttraj_b200 import _engine, classify
    from clusttraj_b200 import distmat as D
    from clusttraj_b200.synth import make_trajectory

    n = min(_engine.device_count(), 4)
    if n < 2:
        pytest.skip("needs two GPUs (the single-GPU gather test covers the same-device path)")
    M = 3000
    atoms, coords = make_trajectory(M, 40, seed=12, n_modes=6)
    one = D.condensed_matrix(atoms, coords, engine_opts(), devices=[0]).copy()
    many, stats = D.condensed_matrix(atoms, coords, engine_opts(), devices=list(range(n)), return_stats=True)
    assert len(stats) == n and all(s["gathered_on"] == 0 for s in stats)
    assert sum(s["pairs"] for s in stats) == M * (M - 1) // 2
    assert np.array_equal(many, one)
    root = D._engine_for(0)
    assert np.array_equal(root.copy_slab(0, len(one)), one)            # gathered copy on GPU 0
    classify.trust_resident(many)
    Z = classify.linkage(many, "average")
    assert root.stats()["h2d_ms"] < 0.5                                # no re-upload of the 36 MB matrix
    assert Z.tobytes() == hcl.linkage(one, "average").tobytes()
    classify.trust_resident(None)
    # Hungarian mode through the per-pair kernel, and the streaming-only fall-back (gather switched off)
    from clusttraj_b200.synth import make_solute_solvent

    atoms, coords = make_solute_solvent(40, n_solute=12, n_solvent_mol=16, seed=2, box=7.0)
    kw = dict(ns=12, reorder=True)
    assert np.array_equal(D.condensed_matrix(atoms, coords, engine_opts(**kw), devices=list(range(n))),
                          D.condensed_matrix(atoms, coords, engine_opts(**kw), devices=[0]))
    os.environ["CLUSTTRAJ_B200_GATHER"] = "0"
    try:
        plain = D.condensed_matrix(atoms, coords, engine_opts(**kw), devices=list(range(n)))
    finally:
        del os.environ["CLUSTTRAJ_B200_GATHER"]
    assert np.array_equal(plain, D.condensed_matrix(atoms, coords, engine_opts(**kw), devices=[0]))
