"""CPU tests: the C-ABI library loads and exports every symbol the header declares, the host-side logic
(slab partition, option parsing, trajectory reader, reorder handles) and the loud failure without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "clusttraj_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ct_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from clusttraj_b200 import _engine

    lib = _engine.load_library()
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert set(_engine.EXPORTS) == set(names)
    assert b"sm_100a" in lib.ct_version()


def test_rows_pairs_matches_condensed_layout():
    from clusttraj_b200 import _engine

    lib = _engine.load_library()
    for M in (1, 2, 3, 17, 1000):
        total = M * (M - 1) // 2
        assert lib.ct_rows_pairs(M, 0, M) == total
        acc = 0
        for r in range(M):
            n = lib.ct_rows_pairs(M, r, r + 1)
            assert n == M - 1 - r
            acc += n
        assert acc == total
    assert lib.ct_rows_pairs(100, 40, 40) == 0


def test_no_gpu_fails_loudly():
    from clusttraj_b200 import _engine

    if _engine.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_engine.EngineError, match="no CUDA device"):
        _engine.Engine(0)
    from clusttraj_b200.distmat import build_distance_matrix
    from clusttraj_b200.io import ClustOptions

    opt = ClustOptions(trajfile=os.path.join(ROOT, "tests", "golden", "ref_testtraj.xyz"), no_hydrogen=True,
                       solute_natoms=17, reorder_excl=np.asarray([], np.int32))
    with pytest.raises(_engine.EngineError):
        build_distance_matrix(opt)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "clusttraj_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


@pytest.mark.parametrize("M,parts", [(2, 1), (10, 3), (20000, 8), (100000, 8), (57, 7)])
def test_slab_bounds_equal_area(M, parts):
    from clusttraj_b200.distmat import slab_bounds

    cuts = slab_bounds(M, parts)
    assert cuts[0] == 0 and cuts[-1] == M and len(cuts) == parts + 1
    assert all(a <= b for a, b in zip(cuts, cuts[1:]))
    area = [b * M - b * (b + 1) // 2 - (a * M - a * (a + 1) // 2) for a, b in zip(cuts, cuts[1:])]
    total = M * (M - 1) // 2
    assert sum(area) == total
    if M >= 1000:
        assert max(area) - min(area) <= 2 * M  # within one row of perfectly equal


def test_xyz_round_trip(tmp_path):
    from clusttraj_b200.synth import make_trajectory
    from clusttraj_b200.xyz import read_trajectory, write_xyz

    z, x = make_trajectory(5, 9, seed=3)
    path = str(tmp_path / "t.xyz")
    write_xyz(path, z, x)
    atoms, coords = read_trajectory(path)
    assert atoms.shape == (5, 9) and (atoms == z).all()
    assert np.array_equal(coords, x)  # 5-decimal coordinates survive the text round trip


def test_reference_fixture_reads_like_the_oracle_reader(golden_dir):
    from clusttraj_b200.xyz import read_xyz
    from oracle.xyz_min import read_xyz as oracle_read

    a, c = read_xyz(os.path.join(golden_dir, "ref_testtraj.xyz"))
    oa, oc = oracle_read(os.path.join(golden_dir, "ref_testtraj.xyz"))
    assert a.shape == (3, 41) and np.array_equal(a, oa) and np.array_equal(c, oc)
    assert (a[0] == 1).sum() == 26 and (a[0] == 6).sum() == 10 and (a[0] == 8).sum() == 5


def test_cli_options_follow_the_reference(tmp_path, golden_dir):
    """Mirrors the reference's test_io.py expectations for the hot-path fields."""
    from clusttraj_b200 import reorder
    from clusttraj_b200.io import ClustOptions, configure_runtime

    traj = os.path.join(golden_dir, "ref_testtraj.xyz")
    common = ["--log", str(tmp_path / "c.log"), "-oc", str(tmp_path / "clusters.dat"), "-od", str(tmp_path / "d.npy")]
    opt = configure_runtime([traj, "-rmsd", "1.0"] + common)
    assert isinstance(opt, ClustOptions)
    assert opt.trajfile == traj and opt.min_rmsd == 1.0 and opt.method == "average" and opt.n_workers == 4
    assert opt.reorder_alg is None and opt.reorder is False and opt.no_hydrogen is False
    assert opt.reorder_excl.dtype == np.int32 and opt.reorder_excl.size == 0
    assert opt.summary_name == str(tmp_path / "clusters.out")
    opt = configure_runtime([traj, "-rmsd", "1.0", "-e", "-ns", "17", "-eex", "1", "2", "5", "-ws", "0.8", "-rs",
                             "--final-kabsch"] + common)
    assert opt.reorder_alg is reorder.reorder_hungarian and opt.reorder_alg_name == "hungarian"
    assert list(opt.reorder_excl) == [0, 1, 4] and opt.exclusions is True  # CLI is 1-based, options 0-based
    assert opt.solute_natoms == 17 and opt.weight_solute == 0.8 and opt.reorder_solvent_only and opt.final_kabsch
    opt = configure_runtime([traj, "-nc", "2", "-e", "--reorder-alg", "inertia-hungarian"] + common)
    assert opt.reorder_alg is None and opt.n_clusters == 2  # accepted but bound to nothing, as in the reference
    for bad in (["-rmsd", "1.0", "-m", "nonsense"], ["-rmsd", "1.0", "-rs"], ["-rmsd", "1.0", "-ws", "0.5"],
                ["-rmsd", "1.0", "-n", "-e", "-eex", "1"], ["-rmsd", "1.0", "-ns", "17", "-ws", "1.5"], []):
        with pytest.raises(SystemExit):
            configure_runtime([traj] + bad + common)
    open(tmp_path / "d.npy", "w").close()
    with pytest.raises(FileExistsError):
        configure_runtime([traj, "-rmsd", "1.0"] + common)
    configure_runtime([traj, "-rmsd", "1.0", "-f"] + common)


def test_reorder_handles():
    from clusttraj_b200 import reorder

    assert reorder.reorder_mode_of(None) == 0
    assert reorder.reorder_mode_of(reorder.reorder_hungarian) == 1

    def reorder_hungarian(a, b, c, d):  # e.g. rmsd.reorder_hungarian held by a user's ClustOptions
        return None

    assert reorder.reorder_mode_of(reorder_hungarian) == 1
    with pytest.raises(NotImplementedError):
        reorder.reorder_mode_of(lambda *a: None)
    with pytest.raises(NotImplementedError):
        reorder.reorder_hungarian(None, None, None, None)


def test_classify_fails_loudly_without_a_gpu(tmp_path, ref_golden_distmat):
    """The clustering step runs its linkage on the GPU (clusttraj_b200/classify.py): without one it raises instead of
    quietly calling SciPy.  (With a GPU, the reference's test_classify.py:30-52 case is test_gpu_parity.py's
    test_classify_on_golden_vector.)  centroid / median are documented SciPy pass-throughs."""
    import scipy.cluster.hierarchy as hcl

    from clusttraj_b200 import _engine
    from clusttraj_b200.classify import linkage
    from clusttraj_b200.io import ClustOptions
    from clusttraj_b200.main import classify

    opt = ClustOptions(method="ward", min_rmsd=1.0, opt_order=False, out_clust_name=str(tmp_path / "c.dat"))
    if _engine.device_count() == 0:
        with pytest.raises(_engine.EngineError):
            classify(opt, ref_golden_distmat)
    assert linkage(ref_golden_distmat, "median").tobytes() == hcl.linkage(ref_golden_distmat, "median").tobytes()
    with pytest.raises(ValueError, match="condensed"):
        linkage(np.zeros(4), "average")


def test_classify_consumers_check_arguments_and_fail_loudly_without_a_gpu():
    """clusttraj_b200.classify mirrors classify.py:146-177; shape errors are raised on the host like squareform's,
    and without a GPU the functions raise instead of falling back to a CPU path."""
    from clusttraj_b200 import _engine
    from clusttraj_b200.classify import find_medoids_from_clusters, sum_distmat

    with pytest.raises(ValueError, match="condensed"):
        sum_distmat(np.zeros(4))                      # 4 is not M (M - 1) / 2
    with pytest.raises(ValueError, match="one label per frame"):
        find_medoids_from_clusters(np.zeros(3), np.array([1, 2]))
    if _engine.device_count() == 0:
        with pytest.raises(_engine.EngineError):
            sum_distmat(np.zeros(3))
        with pytest.raises(_engine.EngineError):
            find_medoids_from_clusters(np.zeros(3), np.array([1, 1, 2]))


def test_native_readers_xyz_fast_path_pdb_gro(tmp_path, golden_dir):
    """The bulk XYZ parse equals the line-by-line one (also on the reference's fixture); PDB (MODEL frames, element
    column or atom-name guess) and GRO (nm -> Angstrom) give the same arrays as the XYZ of the same frames."""
    from clusttraj_b200 import xyz
    from clusttraj_b200.synth import make_trajectory

    z, x = make_trajectory(7, 9, seed=3, n_duplicates=0, mirrored=False, decimals=3)
    p = str(tmp_path / "t.xyz")
    xyz.write_xyz(p, z, x, decimals=3)
    a_fast, c_fast = xyz.read_xyz(p)
    a_slow, c_slow = xyz._read_xyz_slow(open(p).read().split("\n"), p)
    assert (a_fast == a_slow).all() and np.array_equal(c_fast, c_slow) and np.abs(c_fast - x).max() < 1e-12
    ref = os.path.join(golden_dir, "ref_testtraj.xyz")
    a1, c1 = xyz.read_xyz(ref)
    a2, c2 = xyz._read_xyz_slow(open(ref).read().split("\n"), ref)
    assert (a1 == a2).all() and np.array_equal(c1, c2) and c1.shape == (3, 41, 3)
    # irregular file (blank line between frames): the fall-back takes over
    irregular = str(tmp_path / "i.xyz")
    text = open(p).read().split("\n")
    open(irregular, "w").write("\n".join(text[:11]) + "\n\n" + "\n".join(text[11:]))
    a3, c3 = xyz.read_xyz(irregular)
    assert np.array_equal(c3, c_fast) and (a3 == a_fast).all()
    # PDB with and without the element column, GRO
    pdb = str(tmp_path / "t.pdb")
    with open(pdb, "w") as fh:
        for m in range(7):
            fh.write(f"MODEL     {m + 1:4d}\n")
            for k in range(9):
                sym = xyz.SYMBOLS[int(z[k])]
                el = f"{sym:>2s}" if m % 2 == 0 else "  "
                fh.write(f"ATOM  {k + 1:5d} {sym + str(k):<4s} MOL A   1    {x[m, k, 0]:8.3f}{x[m, k, 1]:8.3f}{x[m, k, 2]:8.3f}"
                         f"  1.00  0.00          {el}\n")
            fh.write("ENDMDL\n")
        fh.write("END\n")
    ap, cp = xyz.read_trajectory(pdb)
    assert (ap == a_fast).all() and np.abs(cp - x).max() < 1e-9
    gro = str(tmp_path / "t.gro")
    with open(gro, "w") as fh:
        for m in range(7):
            fh.write(f"frame {m}\n{9:5d}\n")
            for k in range(9):
                fh.write(f"{1:5d}{'MOL':<5s}{xyz.SYMBOLS[int(z[k])] + str(k):>5s}{k + 1:5d}"
                         f"{x[m, k, 0] / 10:8.4f}{x[m, k, 1] / 10:8.4f}{x[m, k, 2] / 10:8.4f}\n")
            fh.write("   5.00000   5.00000   5.00000\n")
    ag, cg = xyz.read_trajectory(gro)
    assert (ag == a_fast).all() and np.abs(cg - x).max() < 1e-3 + 1e-9   # .gro keeps 3-4 decimals of nm


def test_structure_output_host_side(tmp_path, golden_dir):
    """Host pieces of the structure writers (io.py:582-1015 mirror): the per-frame centre equals the one the oracle's
    restatement of io.py:823-856 subtracts, the .xyz writer refuses to overwrite, other formats are refused loudly."""
    from clusttraj_b200 import io as bio
    from clusttraj_b200.xyz import read_xyz
    from oracle import io_restated as IO

    atoms, coords = read_xyz(os.path.join(golden_dir, "ref_testtraj.xyz"))
    for noh, ns in ((False, None), (True, None), (False, 17), (True, 17)):
        c = bio._frame_centres(atoms[0], coords, noh, ns)
        for k in range(len(coords)):
            _, _, centred = IO.reference_frame(atoms[k], coords[k], noh, ns)
            assert np.abs((coords[k] - c[k]) - centred).max() < 1e-12
    path = str(tmp_path / "o.xyz")
    w = bio._XyzWriter(path, overwrite=False)
    w.write(atoms[0], coords[0], "frame 0")
    w.close()
    a, x = read_xyz(path)
    assert (a == atoms[0]).all() and np.abs(x[0] - coords[0]).max() < 1e-5
    with pytest.raises(IOError):
        bio._XyzWriter(path, overwrite=False)
    bio._XyzWriter(path, overwrite=True).close()
    with pytest.raises(NotImplementedError):
        bio._writer("pdb", str(tmp_path / "o.pdb"), True)
    assert bio._xyz_titles(os.path.join(golden_dir, "ref_testtraj.xyz")).__len__() == 3


def test_reorder_alg_distance_binds_its_handle(tmp_path):
    """--reorder-alg distance (rmsd.reorder_distance in the reference, io.py:562-563) selects assignment mode 2; brute / qml
    are refused."""
    from clusttraj_b200 import reorder
    from clusttraj_b200.io import configure_runtime

    traj = tmp_path / "t.xyz"
    traj.write_text("1\n\nC 0 0 0\n")
    base = [str(traj), "-rmsd", "1.0", "-oc", str(tmp_path / "c.dat"), "-od", str(tmp_path / "d.npy")]
    opt = configure_runtime(base + ["-e", "--reorder-alg", "distance"])
    assert opt.reorder_alg is reorder.reorder_distance and reorder.reorder_mode_of(opt.reorder_alg) == 2
    assert reorder.reorder_mode_of(configure_runtime(base + ["-e"]).reorder_alg) == 1
    with pytest.raises(SystemExit):
        configure_runtime(base + ["-e", "--reorder-alg", "brute"])
    with pytest.raises(NotImplementedError):
        reorder.reorder_distance(None, None, None, None)


def test_silhouette_and_metrics_match_the_unmodified_reference(tmp_path, monkeypatch):
    """classify_structures_silhouette / compute_metrics (host side: scikit-learn, SciPy) against the unmodified
    clusttraj/classify.py:11-81 and clusttraj/metrics.py:14-45, imported from /root/reference over the oracle's shims.
    The linkage itself is the GPU's job (identical Z, test_gpu_parity.py); here it is stubbed with SciPy's."""
    import logging

    import scipy.cluster.hierarchy as hcl
    from scipy.spatial.distance import pdist

    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("the reference is only present in the build container")
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_shims")
    monkeypatch.syspath_prepend(ref)
    monkeypatch.syspath_prepend(shims)
    import clusttraj.classify as RC
    import clusttraj.metrics as RM
    from clusttraj.io import ClustOptions as RefOptions
    from clusttraj.io import Logger as RefLogger

    from clusttraj_b200 import classify as C
    from clusttraj_b200.io import ClustOptions

    RefLogger.logger = logging.getLogger("reference-under-test")
    monkeypatch.setattr(C, "linkage", lambda d, method, optimal_ordering=False: hcl.linkage(d, method, optimal_ordering=bool(optimal_ordering)))
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(c, 0.3, size=(12, 3)) for c in (0, 3, 6)])
    d = pdist(x)
    ro = RefOptions(method="average", opt_order=False, out_clust_name=str(tmp_path / "r.dat"), silhouette_score=True)
    Zr, cr = RC.classify_structures_silhouette(ro, d)
    mo = ClustOptions(method="average", opt_order=False, out_clust_name=str(tmp_path / "m.dat"), silhouette_score=True)
    Zm, cm = C.classify_structures_silhouette(mo, d)
    assert Zm.tobytes() == Zr.tobytes() and (cm == cr).all()
    assert np.array_equal(np.asarray(mo.optimal_cut), np.asarray(ro.optimal_cut))
    assert (np.loadtxt(tmp_path / "m.dat", dtype=int) == np.loadtxt(tmp_path / "r.dat", dtype=int)).all()
    assert np.allclose(C.compute_metrics(d, Zm, cm), RM.compute_metrics(d, Zr, cr), rtol=0, atol=0)


def test_folded_level_recombination_is_exact():
    """The int8 Gram epilogue (csrc/ozaki.cu) takes the bias of the integer->double trick out inside its FMAs, level pairs
    in the order the MMA warp issues them (middle, low, high):
      acc = fma(M+v_mid, s_mid, -M(s_mid+s_lo)); acc = fma(M+v_lo, s_lo, acc); acc = acc + fma(M+v_hi, s_hi, -M s_hi).
    Exact rational arithmetic: the two intermediate values and the high product are exact and the result is the correctly
    rounded element, for both bias constants (32-bit and 34-bit level sums) and every fixed-point exponent the engine accepts."""
    import random
    from fractions import Fraction as F

    def fma(a, b, c):
        return float(F(a) * F(b) + F(c))

    rnd = random.Random(7)
    for magic, bits in ((4503601774854144.0, 31), (6755399441055744.0, 34)):
        for f in (22, 26, 30):
            s2 = 2.0 ** (-2 * f)
            s_hi, s_mid, s_lo = s2 * 2.0 ** 40, s2 * 2.0 ** 24, s2 * 2.0 ** 8
            k_mid, k_hi = -magic * (s_mid + s_lo), -magic * s_hi
            assert F(k_mid) == -F(magic) * (F(s_mid) + F(s_lo)) and F(k_hi) == -F(magic) * F(s_hi)
            lim = 2 ** bits - 1
            for t in range(1500):
                v_hi, v_mid, v_lo = (rnd.choice([-lim, lim, 0, 1, -1]) if t < 30 else rnd.randint(-lim, lim) for _ in range(3))
                xm = [magic + vi for vi in (v_hi, v_mid, v_lo)]
                assert all(F(x) == F(magic) + vi for x, vi in zip(xm, (v_hi, v_mid, v_lo)))
                a = fma(xm[1], s_mid, k_mid)
                assert F(a) == v_mid * F(s_mid) - F(magic) * F(s_lo)
                a = fma(xm[2], s_lo, a)
                assert F(a) == v_mid * F(s_mid) + v_lo * F(s_lo)
                h = fma(xm[0], s_hi, k_hi)
                assert F(h) == v_hi * F(s_hi)
                assert a + h == float(v_hi * F(s_hi) + v_mid * F(s_mid) + v_lo * F(s_lo))


def test_native_xyz_parser_matches_the_line_parser_bit_for_bit(tmp_path, golden_dir):
    """ct_read_xyz (parallel native parser in the engine library, host-only) against the general Python line parser:
    identical atomic numbers and identical float64 bits on the reference's trajectory, on odd but regular formatting
    (tabs, CRLF, exponents, extra columns, numeric / lower-case / decorated element tokens, no trailing newline), and a
    clean refusal (None -> fall-back) of irregular files."""
    from clusttraj_b200 import _engine, xyz

    def slow(path):
        with open(path) as fh:
            lines = fh.read().split("\n")
        while lines and not lines[-1].strip():
            lines.pop()
        return xyz._read_xyz_slow(lines, path)

    ref = os.path.join(golden_dir, "ref_testtraj.xyz")
    a0, c0 = slow(ref)
    a1, c1 = _engine.read_xyz_native(ref)
    assert (a0 == a1).all() and c0.tobytes() == c1.tobytes() and a1.dtype == np.int32 and c1.dtype == np.float64
    a2, c2 = xyz.read_xyz(ref)
    assert (a0 == a2).all() and c0.tobytes() == c2.tobytes()
    odd = tmp_path / "odd.xyz"
    odd.write_bytes(b"3\r\ncomment 1\r\nC\t1.0e-3 -2.5E+01  3 extra\r\n 8 0.1 0.2 0.30000000000000004\r\nhe1 1 2 3\r\n"
                    b"3\nsecond frame\ncl 123456789.123456789 -0.0 1e-320\nXx 1 2 3\nH 4 5 6")
    a3, c3 = slow(str(odd))
    a4, c4 = _engine.read_xyz_native(str(odd))
    assert (a3 == a4).all() and c3.tobytes() == c4.tobytes()
    assert a4.tolist() == [[6, 8, 2], [17, 0, 1]]
    rng = np.random.default_rng(0)
    big = tmp_path / "big.xyz"
    coords = rng.normal(scale=20.0, size=(300, 17, 3))
    xyz.write_xyz(str(big), rng.integers(1, 90, size=17), coords, decimals=9)
    a5, c5 = slow(str(big))
    for threads in (0, 1, 3):
        a6, c6 = _engine.read_xyz_native(str(big), n_threads=threads)
        assert (a5 == a6).all() and c5.tobytes() == c6.tobytes()
    for text in ("2\nc\nC 0 0 0\nC 1 1 1\n\n2\nc\nC 0 0 0\nC 1 1 1\n",      # blank separator line
                 "2\nc\nC 0 0 0\nC 1 1 1\n1\nc\nC 0 0 0\n",                  # frames of different sizes
                 "1\nc\nC 0 0\n", "1\nc\nQq 0 0 0\n", "1\nc\nC 0 0 1x\n", "x\n"):
        bad = tmp_path / "bad.xyz"
        bad.write_text(text)
        assert _engine.read_xyz_native(str(bad)) is None
    bad.write_text("1\nc\nQq 0 0 0\n")
    with pytest.raises(ValueError, match="unknown element"):
        xyz.read_xyz(str(bad))                                              # the fall-back words the error
    bad.write_text("2\nc\nC 0 0 0\nC 1 1 1\n\n2\nc\nC 0 0 0\nC 1 1 1\n")
    assert xyz.read_xyz(str(bad))[1].shape == (2, 2, 3)                     # ... and accepts what the fast path refuses


def test_native_xyz_formatter_writes_what_the_python_writer_writes():
    """ct_format_xyz (the layout OpenBabel's XYZ writer gives the reference, io.py:870-899) against Python's own
    formatting of the same frames: identical bytes, including -0.0, values that round at the fifth decimal, wide values,
    non-ASCII titles and thread counts that do not divide the frame count."""
    from clusttraj_b200 import _engine
    from clusttraj_b200.xyz import SYMBOLS

    rng = np.random.default_rng(1)
    z = rng.integers(0, 118, size=23)
    x = rng.normal(scale=30, size=(37, 23, 3))
    x[0, 0] = [0.0, -0.0, 123456789.987654321]
    x[1, 1] = [1e-7, -1e-7, 0.000005]
    x[2, 2] = [2.5e-6, 0.0000149999, -99999.999995]
    titles = [f"frame {i} é energy = {i * 0.1:.3f}" for i in range(37)]
    want = []
    for c, t in zip(x, titles):
        want.append(f"{len(z)}\n{t}\n")
        for a, (p, q, w) in zip(z, c):
            want.append(f"{SYMBOLS[int(a)]:<3s}{p:15.5f}{q:15.5f}{w:15.5f}\n")
    want = "".join(want).encode()
    for threads in (0, 1, 5):
        assert _engine.format_xyz_native(z, x, titles, n_threads=threads) == want
    assert _engine.format_xyz_native(z, x[:0], []) == b""
    assert _engine.format_xyz_native(z, x[:1], None).startswith(b"23\n\n")
    with pytest.raises(_engine.EngineError, match="atomic number"):
        _engine.format_xyz_native(np.array([200]), np.zeros((1, 1, 3)))
    with pytest.raises(ValueError):
        _engine.format_xyz_native(z, x[:, :5], titles)


def test_native_pdb_and_gro_parsers_match_the_python_readers(tmp_path):
    """ct_read_pdb / ct_read_gro (parallel, in the engine library) against clusttraj_b200.xyz's own line readers: identical
    element guesses and float64 bits on multi-frame files with element columns present and absent, HETATM records, TER /
    ENDMDL / END records, CRLF line ends; irregular files are refused (None) and left to the Python readers."""
    from clusttraj_b200 import _engine, xyz

    rng = np.random.default_rng(2)
    M, N = 23, 31
    names = ["CA", "N", "O", "HB1", "CL1", "FE", "1HG2", "OW", "HW1", "Na", "SOD", "C12", "Br", "X"]
    elems = ["C", "N", "O", "", "", "FE", "", "", "", "", "", "", "", ""]
    coords = rng.normal(scale=25, size=(M, N, 3))
    for eol in ("\n", "\r\n"):
        pdb = []
        for m in range(M):
            pdb.append(f"MODEL     {m + 1:4d}")
            for k in range(N):
                x, y, z = coords[m, k]
                rec = "HETATM" if k % 7 == 0 else "ATOM  "
                pdb.append(f"{rec}{k + 1:5d} {names[k % 14]:<4s} RES A{1:4d}    {x:8.3f}{y:8.3f}{z:8.3f}{1.0:6.2f}{0.0:6.2f}"
                           f"          {elems[k % 14]:>2s}")
            pdb += ["TER", "ENDMDL"]
        pdb.append("END")
        p = tmp_path / "t.pdb"
        p.write_bytes((eol.join(pdb) + eol).encode())
        a0, c0 = xyz.read_pdb(str(p), native=False)
        a1, c1 = _engine.read_xyz_native(str(p), fmt="pdb")
        assert a0.shape == (M, N) and (a0 == a1).all() and c0.tobytes() == c1.tobytes()
        a2, c2 = xyz.read_pdb(str(p))
        assert (a0 == a2).all() and c0.tobytes() == c2.tobytes()
        gro = []
        for m in range(M):
            gro += [f"frame t= {m}.0", f"{N:5d}"]
            for k in range(N):
                x, y, z = coords[m, k] / 10
                gro.append(f"{1:5d}{'RES':<5s}{names[k % 14]:>5s}{k + 1:5d}{x:8.3f}{y:8.3f}{z:8.3f}{0.1:8.4f}{0.2:8.4f}{0.3:8.4f}")
            gro.append(f"{3.0:10.5f}{3.0:10.5f}{3.0:10.5f}")
        g = tmp_path / "t.gro"
        g.write_bytes((eol.join(gro) + eol).encode())
        a0, c0 = xyz.read_gro(str(g), native=False)
        a1, c1 = _engine.read_xyz_native(str(g), fmt="gro")
        assert a0.shape == (M, N) and (a0 == a1).all() and c0.tobytes() == c1.tobytes()
        assert xyz.read_trajectory(str(g))[1].tobytes() == c0.tobytes()
    single = tmp_path / "one.pdb"
    single.write_text("\n".join(pdb[1:1 + N]) + "\n")                       # no MODEL / END records at all
    a0, c0 = xyz.read_pdb(str(single), native=False)
    a1, c1 = _engine.read_xyz_native(str(single), fmt="pdb")
    assert a0.shape == (1, N) and (a0 == a1).all() and c0.tobytes() == c1.tobytes()
    bad = tmp_path / "bad.pdb"
    bad.write_text("\n".join(pdb[:N + 3] + pdb[N + 3:2 * N + 3]) + "\n")    # second frame one atom short
    assert _engine.read_xyz_native(str(bad), fmt="pdb") is None
    with pytest.raises(ValueError):
        xyz.read_pdb(str(bad))
    bad.write_text("ATOM      1  CA  RES A   1      1.000   2.000\n")         # line too short for z
    assert _engine.read_xyz_native(str(bad), fmt="pdb") is None
    badg = tmp_path / "bad.gro"
    badg.write_text("\n".join(gro[:N + 1]) + "\n")                          # truncated frame
    assert _engine.read_xyz_native(str(badg), fmt="gro") is None
    with pytest.raises(ValueError):
        xyz.read_gro(str(badg))


def test_hand_made_pdb_and_gro_fixtures(golden_dir):
    """SURVEY 8f-3 pinned from OUTSIDE the repo's readers: three small files laid out by hand after the format definitions
    (tests/golden/hand_elements.pdb, hand_noelements.pdb, hand_water.gro) with the expected atomic numbers and Angstrom
    coordinates typed here as literals.  Both the Python line readers and the native parallel readers (ct_read_pdb /
    ct_read_gro) must return exactly these."""
    from clusttraj_b200 import _engine, xyz

    # 1. element columns present (77-78), two MODELs, ATOM + HETATM, REMARK / CRYST1 / TER / ENDMDL / END records
    z_want = [7, 6, 6, 8, 1, 30, 8]
    x_want = np.array([
        [[-1.195, 0.201, 5.124], [0.234, 0.407, 4.870], [0.986, -0.873, 4.563], [0.407, -1.952, 4.441],
         [0.657, 0.890, 5.752], [12.500, -10.250, 100.125], [-5.000, 3.250, -7.875]],
        [[-1.000, 0.250, 5.000], [0.250, 0.500, 4.750], [1.000, -0.750, 4.500], [0.500, -2.000, 4.250],
         [0.750, 1.000, 5.750], [12.625, -10.125, 99.875], [-4.500, 3.125, -8.000]]])
    path = os.path.join(golden_dir, "hand_elements.pdb")
    for atoms, coords in (xyz.read_pdb(path, native=False), _engine.read_xyz_native(path, fmt="pdb"), xyz.read_trajectory(path)):
        assert atoms.shape == (2, 7) and atoms.dtype == np.int32 and (atoms == z_want).all()
        assert coords.shape == (2, 7, 3) and coords.dtype == np.float64 and np.array_equal(coords, x_want)
    # 2. NO element columns: the alignment of the atom name decides (column 14 = one-letter element, column 13 = two letters,
    #    four-character hydrogen names): N, C-alpha (not calcium), C, H (HG12), H (1HG2), Ca ion, Fe, Cl, O
    z_want = [7, 6, 6, 1, 1, 20, 26, 17, 8]
    first = np.array([[1.000, 2.000, 3.000], [2.458, 2.000, 3.000], [3.000, 3.400, 3.100], [3.500, 3.900, 4.000],
                      [2.200, 4.100, 2.600], [10.000, -20.000, 30.000], [-1.500, 0.000, 0.500], [7.250, 7.500, 7.750],
                      [-9.999, 99.999, -100.001]])
    second = first.copy()
    second[0] = [1.125, 2.250, 3.375]
    second[8] = [-9.500, 99.500, -99.500]
    path = os.path.join(golden_dir, "hand_noelements.pdb")
    for atoms, coords in (xyz.read_pdb(path, native=False), _engine.read_xyz_native(path, fmt="pdb")):
        assert atoms.shape == (2, 9) and (atoms == z_want).all()
        assert np.array_equal(coords, np.stack([first, second]))
    for field, want in ((" CA ", "C"), ("CA  ", "CA"), ("HG12", "H"), ("1HG2", "H"), (" OW ", "O"), ("FE  ", "FE"), ("C12 ", "C"),
                        ("ZN", "ZN"), (" N", "N"), ("HE  ", "HE")):
        assert xyz.pdb_element_from_name(field) == want, field
    # 3. GROMACS .gro: positions in nm -> Angstrom, element from the atom name (OW, HW1, HW2, C1)
    z_want = [8, 1, 1, 8, 1, 1, 6]
    nm = np.array([
        [[0.126, 1.624, 1.679], [0.190, 1.661, 1.747], [0.177, 1.568, 1.613], [1.275, 0.053, 0.622],
         [1.337, 0.002, 0.680], [1.326, 0.120, 0.568], [-0.250, 2.500, 0.001]],
        [[0.130, 1.620, 1.680], [0.195, 1.655, 1.750], [0.180, 1.565, 1.610], [1.270, 0.050, 0.625],
         [1.335, 0.005, 0.685], [1.325, 0.125, 0.565], [-0.255, 2.505, 0.000]]])
    path = os.path.join(golden_dir, "hand_water.gro")
    for atoms, coords in (xyz.read_gro(path, native=False), _engine.read_xyz_native(path, fmt="gro"), xyz.read_trajectory(path)):
        assert atoms.shape == (2, 7) and (atoms == z_want).all()
        assert coords.shape == (2, 7, 3) and np.abs(coords - 10.0 * nm).max() < 1e-12
    assert abs(coords[0, 0, 0] - 1.26) < 1e-12 and abs(coords[1, 6, 1] - 25.05) < 1e-12


def test_cli_accepts_native_formats_and_xyz_structure_output(tmp_path, golden_dir):
    """-cc xyz / -mc xyz reach the native writers, .pdb / .ent / .gro trajectories are accepted without
    OpenBabel (case-insensitively), other output formats and -p are refused with argparse's exit."""
    import shutil

    from clusttraj_b200.io import configure_runtime

    common = ["--log", str(tmp_path / "c.log"), "-oc", str(tmp_path / "clusters.dat"), "-od", str(tmp_path / "d.npy")]
    traj = os.path.join(golden_dir, "ref_testtraj.xyz")
    opt = configure_runtime([traj, "-rmsd", "1.0", "-cc", "xyz", "-mc", "xyz"] + common)
    assert opt.save_confs and opt.save_medoids and opt.out_conf_fmt == "xyz" and opt.out_medoids_fmt == "xyz"
    assert opt.out_conf_name == str(tmp_path / "clusters_confs") and opt.out_medoids_name == str(tmp_path / "clusters_medoids")
    for bad in (["-cc", "pdb"], ["-mc", "mol2"], ["-p"]):
        with pytest.raises(SystemExit):
            configure_runtime([traj, "-rmsd", "1.0"] + bad + common)
    for name in ("hand_elements.pdb", "hand_water.gro"):
        assert configure_runtime([os.path.join(golden_dir, name), "-rmsd", "1.0"] + common).trajfile.endswith(name)
    upper = tmp_path / "UPPER.PDB"
    shutil.copy(os.path.join(golden_dir, "hand_elements.pdb"), upper)
    ent = tmp_path / "pdb1abc.ent"
    shutil.copy(os.path.join(golden_dir, "hand_elements.pdb"), ent)
    for p in (upper, ent):
        configure_runtime([str(p), "-rmsd", "1.0"] + common)
    weird = tmp_path / "traj.mol2"
    weird.write_text("x\n")
    with pytest.raises(SystemExit):
        configure_runtime([str(weird), "-rmsd", "1.0"] + common)


def test_frame_invariant_view_of_per_frame_element_lists():
    """distmat.frame_invariant_view (reference distmat.py:131-141 reads every frame's own atoms): the per-frame hydrogen
    masks are applied on the host, what cannot be described once for all frames is refused loudly."""
    from clusttraj_b200.distmat import frame_invariant_view
    from clusttraj_b200.reorder import reorder_hungarian

    rng = np.random.default_rng(3)
    coords = rng.normal(size=(3, 5, 3))
    same = np.array([[6, 1, 8, 1, 7]] * 3, dtype=np.int32)
    a, c, noh, ns, fz = frame_invariant_view(same, coords, True, None, 3)
    assert a.tolist() == [6, 1, 8, 1, 7] and c is coords and noh is True and ns == 3 and fz is None      # nothing to do
    # the hydrogens sit at different places, two kept solute atoms among the first three everywhere
    atoms = np.array([[6, 1, 8, 1, 7], [1, 6, 8, 7, 1], [6, 8, 1, 1, 7]], dtype=np.int32)
    a, c, noh, ns, fz = frame_invariant_view(atoms, coords, True, None, 3)
    assert noh is False and ns == 2 and a.tolist() == [6, 8, 7] and c.shape == (3, 3, 3) and fz is None
    assert np.array_equal(c[1], coords[1][[1, 2, 3]]) and np.array_equal(c[2], coords[2][[0, 1, 4]])
    # without -n nothing is masked: the elements only matter to a reorder
    a, c, noh, ns, fz = frame_invariant_view(atoms, coords, False, None, None)
    assert noh is False and ns is None and c.shape == (3, 5, 3) and np.array_equal(c, coords) and fz is None
    # a reorder gets every frame's own list (the engine builds one set of block index lists per frame) ...
    a, c, noh, ns, fz = frame_invariant_view(atoms, coords, False, reorder_hungarian, None)
    assert fz is not None and fz.dtype == np.int32 and np.array_equal(fz, atoms) and a.tolist() == atoms[0].tolist()
    # ... unless the kept elements agree after the mask
    a, c, noh, ns, fz = frame_invariant_view(atoms, coords, True, reorder_hungarian, None)
    assert a.tolist() == [6, 8, 7] and fz is None
    with pytest.raises(ValueError):
        frame_invariant_view(np.array([[6, 1, 8], [6, 8, 7]], dtype=np.int32), coords[:2, :3], True, None, None)
    with pytest.raises(NotImplementedError):   # -ns -n: the number of kept solute atoms changes from frame to frame
        frame_invariant_view(np.array([[6, 1, 8, 7], [6, 8, 1, 7]], dtype=np.int32), coords[:2, :4], True, None, 2)
