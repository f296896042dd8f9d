"""Drop-in for clusttraj/distmat.py: the same three entry points, the same condensed float64 vector,
computed by the B200 engine.

  get_distmat(clust_opt)            distmat.py:15-41   (-i load, or build + np.save)
  build_distance_matrix(clust_opt)  distmat.py:44-78   (whole matrix; the Pool over rows becomes
                                                        equal-area row slabs over the visible GPUs)
  compute_distmat_line(...)         distmat.py:81-277  (one row)

Differences a caller can observe: the trajectory is parsed ONCE (the reference re-parses it for every
pair), ``n_workers`` does not fork processes (CUDA and fork do not mix) and nothing runs on the CPU:
without a GPU these functions raise EngineError.
"""
import os
import threading

import numpy as np

from . import _engine
from .io import ClustOptions, Logger  # noqa: F401  (re-exported like the reference module)
from .reorder import reorder_mode_of
from .xyz import read_trajectory

_engines = {}
_engines_lock = threading.Lock()
_traj_cache = {}


def _engine_for(device):
    with _engines_lock:
        eng = _engines.get(device)
        if eng is None:
            eng = _engines[device] = _engine.Engine(device)
        return eng


def visible_devices():
    """GPUs this process drives: CLUSTTRAJ_B200_DEVICES="0,1,..." or all visible; under torchrun
    (LOCAL_RANK set) only the rank's own GPU."""
    env = os.environ.get("CLUSTTRAJ_B200_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    n = _engine.device_count()
    if n == 0:
        raise _engine.EngineError("clusttraj_b200: no CUDA device visible (there is no CPU fallback)")
    if "LOCAL_RANK" in os.environ:
        return [int(os.environ["LOCAL_RANK"]) % n]
    return list(range(n))


def _load_trajectory(trajfile):
    """(atomic numbers, coords float64[M,N,3]); the atomic numbers are int32[N] when every frame lists the same atoms
    in the same order (the usual trajectory) and int32[M,N] otherwise (see frame_invariant_view)."""
    key = (os.path.abspath(trajfile), os.path.getmtime(trajfile), os.path.getsize(trajfile))
    hit = _traj_cache.get("last")
    if hit and hit[0] == key:
        return hit[1], hit[2]
    atoms, coords = read_trajectory(trajfile)
    atoms = atoms[0].copy() if not (atoms != atoms[0]).any() else atoms
    _traj_cache["last"] = (key, atoms, coords)
    return atoms, coords


def frame_invariant_view(atoms, coords, noh, reorder, nsatoms):
    """Per-frame element lists (the reference reads every frame's own atoms, distmat.py:131-141) brought to the form the
    engine takes: returns (atoms int32[K], coords, noh, nsatoms, frame_elements) to compute with; ``frame_elements`` is
    None or int32[M, K] for ``Engine.load_frames(..., frame_elements=...)``.

    The reference uses a frame's elements for two things only: the ``-n`` mask (P[p_atoms != 1], Q[q_atoms != 1]; what is
    left is paired by POSITION) and the element blocks of a reorder.  Frames whose lists differ are handled by applying each
    frame's own hydrogen mask here, on the host, and handing the engine the compacted frames with the mask off; a reorder
    then gets every frame's own (kept) element list (``ct_set_frame_elements``: one set of block index lists per frame).
    Refused loudly, because the reference has no defined result or the engine no description for it:
      * frames that keep different numbers of atoms (the reference's Kabsch needs equal shapes too: ValueError);
      * with ``-ns -n`` a number of kept solute atoms that changes from frame to frame (the reference takes it from the
        row frame of each pair, :116-118);
      * a reorder over frames whose element COUNTS differ among the atoms it reorders (refused by the engine: the
        reference's per-element assignment fails on such a pair).
    """
    atoms = np.asarray(atoms)
    if atoms.ndim == 1 or not (atoms != atoms[0]).any():
        return np.ascontiguousarray(atoms if atoms.ndim == 1 else atoms[0], dtype=np.int32), coords, noh, nsatoms, None
    keep = (atoms != 1) if noh else np.ones(atoms.shape, dtype=bool)
    counts = keep.sum(axis=1)
    if (counts != counts[0]).any():
        k = int(np.argmax(counts != counts[0]))
        raise ValueError(f"frames keep different numbers of atoms ({int(counts[0])} in frame 0, {int(counts[k])} in frame {k}): "
                         "the RMSD of structures of different sizes is not defined")
    if nsatoms and noh:
        nat = (atoms[:, :nsatoms] != 1).sum(axis=1)
        if (nat != nat[0]).any():
            raise NotImplementedError("clusttraj_b200: -ns with -n needs the same number of non-hydrogen solute atoms in every "
                                      "frame (the reference takes it from the row frame of each pair)")
        nsatoms = int(nat[0])
    K = int(counts[0])
    kept = np.ascontiguousarray(atoms[keep].reshape(len(atoms), K), dtype=np.int32)
    frame_elements = kept if (reorder_mode_of(reorder) != 0 and (kept != kept[0]).any()) else None
    coords = np.ascontiguousarray(np.asarray(coords, dtype=np.float64)[keep].reshape(len(atoms), K, 3))
    return np.ascontiguousarray(kept[0], dtype=np.int32), coords, False, nsatoms, frame_elements


def _opts_from(noh, reorder, reorder_solvent_only, nsatoms, weight_solute, reorderexcl, final_kabsch):
    mode = reorder_mode_of(reorder)
    excl = np.asarray(reorderexcl if reorderexcl is not None else [], dtype=np.int32).ravel()
    if mode == 0:
        excl = excl[:0]  # the reference only reads the exclusions inside `if reorder` blocks
    return _engine.make_opts(no_hydrogen=bool(noh), solute_natoms=nsatoms, reorder=mode,
                             reorder_solvent_only=bool(reorder_solvent_only), final_kabsch=bool(final_kabsch),
                             weight_solute=weight_solute, reorder_excl=excl)


def slab_bounds(M, parts):
    """Row boundaries of `parts` contiguous slabs with (nearly) equal numbers of matrix entries
    (SURVEY.md §8e): the reference's consecutive-row chunks are badly unbalanced on a triangle."""
    total = M * (M - 1) // 2
    cuts = [0]
    for p in range(1, parts):
        target = total * p / parts
        # smallest r with r*M - r(r+1)/2 >= target
        b = 2 * M - 1
        r = int((b - np.sqrt(max(b * b - 8.0 * target, 0.0))) / 2.0)
        r = min(max(r, cuts[-1]), M)
        while r < M and r * M - r * (r + 1) // 2 < target:
            r += 1
        cuts.append(r)
    cuts.append(M)
    return cuts


def condensed_matrix(atomic_nums, coords, opts, devices=None, out=None, return_stats=False, frame_elements=None):
    """All-pairs matrix of in-memory frames on one or several GPUs (one engine + one host thread per
    GPU, no collective: every GPU owns a contiguous slab of the condensed vector)."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    M = coords.shape[0]
    total = M * (M - 1) // 2
    devices = list(devices) if devices is not None else visible_devices()
    if total == 0:
        res = np.empty(0, dtype=np.float64)
        return (res, []) if return_stats else res
    devices = devices[: max(1, min(len(devices), M - 1))]
    if out is None and len(devices) == 1:
        # One GPU, one shot: the matrix is computed into HBM, stays resident there for the consumers that follow
        # (clusttraj_b200.classify) and is copied into a plain numpy array.  Page-locking 1.6 GB for a single use costs
        # 0.6-0.7 s, the driver's staged copy into pageable memory 0.35-0.4 s in all (measured at 20k frames).  Matrices
        # that would not leave room for the clustering's working copy are streamed in chunks instead (below).
        eng = _engine_for(devices[0])
        if 8 * total * 4 < eng.free_memory():
            eng.load_frames(coords, atomic_nums, opts, frame_elements)
            _, n, st = eng.rmsd_rows_device(0, M)
            out = eng.copy_slab(0, n)
            eng._note_resident(out)
            stats = [dict(st, device=devices[0], rows=(0, M))]
            return (out, stats) if return_stats else out
    if out is None:
        # several GPUs (or one GPU whose memory cannot hold the whole matrix) stream chunks concurrently with the compute
        # into one vector: that needs pinned memory to run at full PCIe rate
        out = _engine.pinned_empty(total)
    cuts = slab_bounds(M, len(devices))
    stats = [None] * len(devices)
    errors = []
    # Several GPUs: the slabs also land in the first GPU's memory, chunk by chunk over NVLink while the next chunk is
    # computed (ct_rmsd_rows_gather), so that the clustering and the matrix consumers that follow
    # (clusttraj_b200.classify) find the whole matrix resident instead of uploading 8 L bytes again.  Only when it fits
    # with room to spare for the clustering's working copy.
    root = None
    if len(devices) > 1 and os.environ.get("CLUSTTRAJ_B200_GATHER", "1") != "0":
        cand = _engine_for(devices[0])
        if 8 * total * 2.5 < cand.free_memory():
            cand.matrix_reserve(M)
            root = cand

    def work(k):
        try:
            eng = _engine_for(devices[k])
            eng.load_frames(coords, atomic_nums, opts, frame_elements)
            lo = cuts[k] * M - cuts[k] * (cuts[k] + 1) // 2
            n = eng.rows_pairs(cuts[k], cuts[k + 1])
            if root is not None:
                _, st = eng.rmsd_rows_gather(cuts[k], cuts[k + 1], root, out[lo: lo + n])
            else:
                _, st = eng.rmsd_rows(cuts[k], cuts[k + 1], out[lo: lo + n])
            stats[k] = dict(st, device=devices[k], rows=(cuts[k], cuts[k + 1]), gathered_on=devices[0] if root is not None else None)
        except Exception as exc:  # surfaced below, on the caller's thread
            errors.append(exc)

    if len(devices) == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    if root is not None:
        root.matrix_commit(M, out)
    return (out, stats) if return_stats else out


def get_distmat(clust_opt):
    """Read (-i) or compute and save the condensed RMSD matrix (distmat.py:15-41)."""
    distmat, wait = get_distmat_saving_in_background(clust_opt)
    wait()
    return distmat


def get_distmat_saving_in_background(clust_opt):
    """get_distmat for a caller that has GPU work to do next (clusttraj_b200.main): returns (distmat, wait) as soon as the
    matrix is in host memory, while the 8 L bytes of the .npy file (0.3 s at 20k frames) are written by a background
    thread; `wait()` joins it and re-raises what it raised.  The array must not be written to before `wait()` returns."""
    if clust_opt.input_distmat:
        Logger.logger.info(f"Reading condensed RMSD matrix from {clust_opt.distmat_name}\n")
        return np.load(clust_opt.distmat_name), (lambda: None)
    Logger.logger.info(f"Calculating RMSD matrix on {len(visible_devices())} GPU(s)\n")
    distmat = build_distance_matrix(clust_opt)
    Logger.logger.info(f"Saving condensed RMSD matrix to {clust_opt.distmat_name}\n")
    failure = []

    def save():
        try:
            np.save(clust_opt.distmat_name, distmat)
        except BaseException as exc:   # surfaced by wait(), on the caller's thread
            failure.append(exc)

    worker = threading.Thread(target=save, name="clusttraj-b200-npy-writer")
    worker.start()

    def wait():
        worker.join()
        if failure:
            raise failure[0]

    return distmat, wait


def build_distance_matrix(clust_opt):
    """Whole condensed matrix for the options in ``clust_opt`` (distmat.py:44-78)."""
    atoms, coords = _load_trajectory(clust_opt.trajfile)
    atoms, coords, noh, nsatoms, frame_z = frame_invariant_view(atoms, coords, clust_opt.no_hydrogen, clust_opt.reorder_alg,
                                                                clust_opt.solute_natoms)
    opts = _opts_from(noh, clust_opt.reorder_alg, clust_opt.reorder_solvent_only, nsatoms, clust_opt.weight_solute,
                      clust_opt.reorder_excl, clust_opt.final_kabsch)
    return condensed_matrix(atoms, coords, opts, frame_elements=frame_z)


def compute_distmat_line(idx1, q_info, trajfile, noh, reorder, reorder_solvent_only, nsatoms, weight_solute,
                         reorderexcl, final_kabsch):
    """RMSD between frame ``idx1`` and every later frame (distmat.py:81-277).  ``q_info`` (the row
    frame's atoms and coordinates) is accepted for signature compatibility and checked against the file."""
    atoms, coords = _load_trajectory(trajfile)
    if q_info is not None:
        q_atoms = np.asarray(q_info[0])
        row_atoms = atoms if atoms.ndim == 1 else atoms[int(idx1)]
        if q_atoms.shape != row_atoms.shape or (q_atoms != row_atoms).any():
            raise ValueError("q_info atoms do not match the trajectory file")
    atoms, coords, noh, nsatoms, frame_z = frame_invariant_view(atoms, coords, noh, reorder, nsatoms)
    opts = _opts_from(noh, reorder, reorder_solvent_only, nsatoms, weight_solute, reorderexcl, final_kabsch)
    eng = _engine_for(visible_devices()[0])
    eng.load_frames(coords, atoms, opts, frame_z)
    line, _ = eng.rmsd_rows(int(idx1), int(idx1) + 1)
    return [np.float64(x) for x in line]
