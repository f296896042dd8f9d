"""ctypes binding of the C ABI in include/clusttraj_b200.h — the only way the Python side reaches the
GPU.  There is no CPU implementation behind it: a missing library or a missing device raises."""
import ctypes as C
import os
import threading
import weakref

import numpy as np

from . import _build


class EngineError(RuntimeError):
    pass


class ct_opts(C.Structure):
    _fields_ = [
        ("no_hydrogen", C.c_int32),
        ("solute_natoms", C.c_int32),
        ("reorder_mode", C.c_int32),
        ("reorder_solvent_only", C.c_int32),
        ("final_kabsch", C.c_int32),
        ("n_excl", C.c_int32),
        ("weight_solute", C.c_double),
        ("reorder_excl", C.POINTER(C.c_int32)),
    ]


class ct_stats(C.Structure):
    _fields_ = [
        ("pack_ms", C.c_double), ("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("gram_ms", C.c_double),
        ("pair_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
        ("pairs", C.c_int64), ("flagged", C.c_int64), ("launches", C.c_int64), ("kept_atoms", C.c_int64),
        ("lsap_steps", C.c_int64), ("gram_path", C.c_int64), ("quant_step", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "ct_version", "ct_device_count", "ct_last_error", "ct_engine_create", "ct_engine_destroy", "ct_host_alloc",
    "ct_host_free", "ct_load_frames", "ct_rows_pairs", "ct_rmsd_rows", "ct_rmsd_rows_device", "ct_copy_slab",
    "ct_rmsd_condensed", "ct_pair_values", "ct_get_stats", "ct_pack_resident", "ct_matrix_consumers",
    "ct_pair_transforms", "ct_linkage", "ct_matrix_reserve", "ct_rmsd_rows_gather", "ct_matrix_commit", "ct_read_xyz", "ct_read_pdb", "ct_read_gro", "ct_format_xyz", "ct_buffer_free", "ct_device_memory", "ct_set_frame_elements",
]

_lib = None
_lib_lock = threading.Lock()


def load_library(build_if_missing=True):
    """dlopen the in-tree engine (building it with nvcc first when it is missing or stale)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        if build_if_missing and _build.is_stale():
            try:
                _build.build()
            except Exception as exc:
                if not os.path.exists(path):
                    raise EngineError(f"clusttraj_b200: CUDA engine is not built and cannot be built: {exc}") from exc
                # a prebuilt library older than its sources (e.g. a box without nvcc): usable only if it still exports
                # the whole ABI, which is checked below; say so instead of running outdated kernels silently
                import warnings
                warnings.warn(f"clusttraj_b200: {path} is older than its sources and could not be rebuilt ({exc}); "
                              "loading the existing library", RuntimeWarning)
        if not os.path.exists(path):
            raise EngineError("clusttraj_b200: CUDA engine library missing (run `python -m clusttraj_b200._build`)")
        lib = C.CDLL(path)
        missing = [name for name in EXPORTS if not hasattr(lib, name)]
        if missing:
            raise EngineError(f"clusttraj_b200: {path} is out of date (missing C-ABI symbols {missing}); rebuild it with "
                              "`python -m clusttraj_b200._build`")
        P = C.POINTER
        lib.ct_version.restype = C.c_char_p
        lib.ct_device_count.restype = C.c_int
        lib.ct_last_error.restype = C.c_char_p
        lib.ct_last_error.argtypes = [C.c_void_p]
        lib.ct_engine_create.argtypes = [C.c_int, P(C.c_void_p)]
        lib.ct_engine_destroy.argtypes = [C.c_void_p]
        lib.ct_engine_destroy.restype = None
        lib.ct_host_alloc.argtypes = [P(C.c_void_p), C.c_uint64]
        lib.ct_host_free.argtypes = [C.c_void_p]
        lib.ct_load_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, P(ct_opts)]
        lib.ct_rows_pairs.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        lib.ct_rows_pairs.restype = C.c_int64
        lib.ct_rmsd_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, P(ct_stats)]
        lib.ct_rmsd_rows_device.argtypes = [C.c_void_p, C.c_int64, C.c_int64, P(C.c_void_p), P(ct_stats)]
        lib.ct_copy_slab.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        lib.ct_set_frame_elements.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        lib.ct_matrix_reserve.argtypes = [C.c_void_p, C.c_int64]
        lib.ct_rmsd_rows_gather.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, P(ct_stats)]
        lib.ct_matrix_commit.argtypes = [C.c_void_p, C.c_int64]
        lib.ct_rmsd_condensed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, P(ct_opts),
                                          C.c_void_p, P(ct_stats)]
        lib.ct_pair_values.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, P(ct_stats)]
        lib.ct_get_stats.argtypes = [C.c_void_p, P(ct_stats)]
        lib.ct_pack_resident.argtypes = [C.c_void_p, P(ct_stats)]
        lib.ct_pair_transforms.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, P(ct_stats)]
        lib.ct_device_memory.argtypes = [C.c_void_p, P(C.c_int64), P(C.c_int64)]
        lib.ct_format_xyz.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_char_p, P(C.c_void_p), P(C.c_int64), C.c_int32]
        lib.ct_buffer_free.argtypes = [C.c_void_p]
        lib.ct_buffer_free.restype = None
        for _name in ("ct_read_xyz", "ct_read_pdb", "ct_read_gro"):
            getattr(lib, _name).argtypes = [C.c_char_p, P(C.c_int64), P(C.c_int32), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        lib.ct_linkage.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, P(ct_stats)]
        lib.ct_matrix_consumers.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                            P(C.c_double), P(ct_stats)]
        _lib = lib
        return lib


def device_count():
    return int(load_library().ct_device_count())


def pinned_empty(n, dtype=np.float64):
    """A writable C-contiguous ndarray over page-locked host memory (freed when the array dies)."""
    lib = load_library()
    dtype = np.dtype(dtype)
    nbytes = max(int(n) * dtype.itemsize, 1)
    ptr = C.c_void_p()
    rc = lib.ct_host_alloc(C.byref(ptr), nbytes)
    if rc != 0:
        raise EngineError("ct_host_alloc failed: " + (lib.ct_last_error(None) or b"").decode())
    buf = (C.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))
    weakref.finalize(buf, lib.ct_host_free, ptr.value)
    return arr


def read_xyz_native(path, n_threads=0, fmt="xyz"):
    """(atomic numbers int32[M, N], coords float64[M, N, 3]) of a regular multi-frame XYZ (or, with fmt, PDB / GRO) file
    through the library's parallel parser (ct_read_xyz / ct_read_pdb / ct_read_gro; host-only, no GPU needed), or None
    when the file is not regular enough for it."""
    lib = load_library()
    reader = getattr(lib, "ct_read_" + fmt)
    M, N = C.c_int64(), C.c_int32()
    raw = os.fsencode(path)
    if reader(raw, C.byref(M), C.byref(N), None, None, 0, 0) != 0:
        return None
    atoms = np.empty((M.value, N.value), dtype=np.int32)
    coords = np.empty((M.value, N.value, 3), dtype=np.float64)
    if reader(raw, C.byref(M), C.byref(N), coords.ctypes.data, atoms.ctypes.data, M.value, int(n_threads)) != 0:
        return None
    return atoms, coords


def format_xyz_native(atomic_numbers, coords, titles=None, n_threads=0):
    """bytes: the XYZ text of the frames coords[F, N, 3] (same atoms in every frame) as the reference's OpenBabel writer
    lays it out, formatted by the library's host threads (ct_format_xyz)."""
    lib = load_library()
    z = np.ascontiguousarray(atomic_numbers, dtype=np.int32)
    x = np.ascontiguousarray(coords, dtype=np.float64)
    if x.ndim != 3 or x.shape[1:] != (z.size, 3):
        raise ValueError("coords must be [frames, atoms, 3] for the given atoms")
    blob = None
    if titles is not None:
        if len(titles) != x.shape[0]:
            raise ValueError("one title per frame")
        blob = b"".join(t.replace("\x00", "").encode() + b"\x00" for t in titles)
    ptr, n = C.c_void_p(), C.c_int64()
    rc = lib.ct_format_xyz(z.ctypes.data, x.ctypes.data, x.shape[0], z.size, blob, C.byref(ptr), C.byref(n), int(n_threads))
    if rc != 0:
        raise EngineError("ct_format_xyz: " + (lib.ct_last_error(None) or b"").decode())
    try:
        return C.string_at(ptr.value, n.value)
    finally:
        lib.ct_buffer_free(ptr)


def make_opts(no_hydrogen=False, solute_natoms=None, reorder=False, reorder_solvent_only=False, final_kabsch=False,
              weight_solute=None, reorder_excl=()):
    excl = np.ascontiguousarray(np.asarray(reorder_excl, dtype=np.int32).ravel())
    o = ct_opts()
    o.no_hydrogen = 1 if no_hydrogen else 0
    o.solute_natoms = int(solute_natoms) if solute_natoms else 0
    o.reorder_mode = (2 if reorder in (2, "distance") else 1) if reorder else 0   # 1 hungarian, 2 distance
    o.reorder_solvent_only = 1 if reorder_solvent_only else 0
    o.final_kabsch = 1 if final_kabsch else 0
    o.weight_solute = float(weight_solute) if weight_solute else 0.0
    o.n_excl = int(excl.size)
    o.reorder_excl = excl.ctypes.data_as(C.POINTER(C.c_int32)) if excl.size else None
    o._keepalive = excl
    return o


class Engine:
    """One engine = one GPU.  Not re-entrant: use one host thread per engine."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.ct_engine_create(int(device), C.byref(h))
        if rc != 0:
            raise EngineError("ct_engine_create failed: " + (self._lib.ct_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)
        self.M = 0
        self.N = 0
        self.K = 0
        self._finalizer = weakref.finalize(self, self._lib.ct_engine_destroy, h)

    def close(self):
        self._finalizer()

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): " + (self._lib.ct_last_error(self._h) or b"").decode())

    def load_frames(self, coords, atomic_nums, opts, frame_elements=None):
        """coords [M, N, 3], atomic_nums [N] (frame 0's when `frame_elements` [M, N] gives every frame its own list:
        ct_set_frame_elements; those frames must come already masked, with no_hydrogen off)."""
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim != 3 or coords.shape[2] != 3:
            raise ValueError("coords must be [M, N, 3]")
        z = np.ascontiguousarray(atomic_nums, dtype=np.int32)
        if z.shape != (coords.shape[1],):
            raise ValueError("atomic_nums must be [N] (frame-invariant)")
        self._check(self._lib.ct_load_frames(self._h, coords.ctypes.data, z.ctypes.data, coords.shape[0],
                                             coords.shape[1], C.byref(opts)), "ct_load_frames")
        self.M, self.N = coords.shape[0], coords.shape[1]
        self.K = int(self.stats()["kept_atoms"])
        if frame_elements is not None:
            fz = np.ascontiguousarray(frame_elements, dtype=np.int32)
            if fz.shape != (self.M, self.K):
                raise ValueError("frame_elements must be [M, K] (kept atoms of every frame)")
            self._check(self._lib.ct_set_frame_elements(self._h, fz.ctypes.data, self.M, self.K), "ct_set_frame_elements")

    def pack_resident(self):
        """Re-run K1 (centre + pack) on the coordinates already on the GPU; returns its device time in ms."""
        st = ct_stats()
        self._check(self._lib.ct_pack_resident(self._h, C.byref(st)), "ct_pack_resident")
        return st.pack_ms

    def rows_pairs(self, row_begin, row_end):
        return int(self._lib.ct_rows_pairs(self.M, row_begin, row_end))

    def rmsd_rows(self, row_begin, row_end, out=None):
        n = self.rows_pairs(row_begin, row_end)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size < n:
            raise ValueError("out must be a C-contiguous float64 array with room for the slab")
        st = ct_stats()
        self._check(self._lib.ct_rmsd_rows(self._h, row_begin, row_end, out.ctypes.data, C.byref(st)), "ct_rmsd_rows")
        return out[:n], st.as_dict()

    def matrix_reserve(self, n_frames):
        """Room for the whole condensed matrix of n_frames frames on this engine's GPU (target of rmsd_rows_gather)."""
        self._resident_ref = None
        self._check(self._lib.ct_matrix_reserve(self._h, int(n_frames)), "ct_matrix_reserve")

    def rmsd_rows_gather(self, row_begin, row_end, root, out=None, to_host=True):
        """rmsd_rows that also lands the rows in `root`'s reserved matrix (GPU -> GPU over NVLink, or in place when root is
        this engine); to_host=False skips the host copy.  Returns (host slab or None, stats)."""
        n = self.rows_pairs(row_begin, row_end)
        if to_host:
            if out is None:
                out = np.empty(n, dtype=np.float64)
            if out.dtype != np.float64 or not out.flags.c_contiguous or out.size < n:
                raise ValueError("out must be a C-contiguous float64 array with room for the slab")
        st = ct_stats()
        self._check(self._lib.ct_rmsd_rows_gather(self._h, row_begin, row_end, out.ctypes.data if to_host else None, root._h,
                                                  C.byref(st)), "ct_rmsd_rows_gather")
        return (out[:n] if to_host else None), st.as_dict()

    def matrix_commit(self, n_frames, array=None):
        """Every gathering call has returned: the reserved matrix is complete; `array` = the host vector it equals (for
        trusted callers, see matrix_consumers)."""
        self._check(self._lib.ct_matrix_commit(self._h, int(n_frames)), "ct_matrix_commit")
        self._note_resident(array)

    def free_memory(self):
        """Free bytes on this engine's GPU."""
        f, t = C.c_int64(), C.c_int64()
        self._check(self._lib.ct_device_memory(self._h, C.byref(f), C.byref(t)), "ct_device_memory")
        return int(f.value)

    # ---- which host array the matrix resident on the GPU corresponds to (so that trusted callers can skip re-uploads) ----
    def _note_resident(self, array):
        self._resident_ref = weakref.ref(array) if isinstance(array, np.ndarray) else None

    def _resident_is(self, array):
        ref = getattr(self, "_resident_ref", None)
        return ref is not None and ref() is array

    def rmsd_rows_device(self, row_begin, row_end):
        """Leaves the slab on the GPU; returns (device pointer, entries, stats)."""
        self._resident_ref = None
        st = ct_stats()
        ptr = C.c_void_p()
        self._check(self._lib.ct_rmsd_rows_device(self._h, row_begin, row_end, C.byref(ptr), C.byref(st)),
                    "ct_rmsd_rows_device")
        return ptr.value, self.rows_pairs(row_begin, row_end), st.as_dict()

    def copy_slab(self, offset, count, out=None):
        if out is None:
            out = np.empty(count, dtype=np.float64)
        self._check(self._lib.ct_copy_slab(self._h, offset, count, out.ctypes.data), "ct_copy_slab")
        return out

    def pair_values(self, pairs, want_perms=False):
        pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        vals = np.empty(len(pairs), dtype=np.float64)
        perms = np.empty((len(pairs), self.K), dtype=np.int32) if want_perms else None
        st = ct_stats()
        self._check(self._lib.ct_pair_values(self._h, pairs.ctypes.data, len(pairs), vals.ctypes.data,
                                             perms.ctypes.data if want_perms else None, C.byref(st)), "ct_pair_values")
        return (vals, perms, st.as_dict()) if want_perms else (vals, st.as_dict())

    def pair_transforms(self, pairs):
        """(values, R [n,3,3], T [n,3]) for (reference frame, moving frame) pairs: aligned = (raw - centre) @ R + T is
        what clusttraj/io.py:align_mol writes for the moving frame (all atoms)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        vals = np.empty(len(pairs), dtype=np.float64)
        xf = np.empty((len(pairs), 12), dtype=np.float64)
        st = ct_stats()
        self._check(self._lib.ct_pair_transforms(self._h, pairs.ctypes.data, len(pairs), vals.ctypes.data, xf.ctypes.data,
                                                 C.byref(st)), "ct_pair_transforms")
        return vals, xf[:, :9].reshape(-1, 3, 3).copy(), xf[:, 9:].copy()

    def matrix_consumers(self, n_frames, clusters=None, distmat=None, trusted=False):
        """(total sum, medoid frame index per cluster, stats) of a condensed matrix: `distmat` (host array) or, with
        distmat=None, the whole matrix the last rmsd_rows_device(0, M) left on the GPU.  `clusters` = 1-based labels.
        trusted=True: the caller vouches that `distmat` has not been written to since this engine produced / uploaded
        it, so the copy still resident on the GPU is used instead of another upload."""
        M = int(n_frames)
        ptr = None
        if distmat is not None:
            given = distmat
            distmat = np.ascontiguousarray(distmat, dtype=np.float64)
            if distmat.shape != (M * (M - 1) // 2,):
                raise ValueError("distmat must be the condensed vector of n_frames frames")
            if not (trusted and self._resident_is(given)):
                ptr = distmat.ctypes.data
                self._note_resident(given if given is distmat else None)
        if clusters is None:
            labels, n_clusters, med = None, 0, None
        else:
            labels = np.ascontiguousarray(clusters, dtype=np.int32)
            if labels.shape != (M,):
                raise ValueError("clusters must hold one label per frame")
            n_clusters = int(len(np.unique(labels)))   # classify.py:157
            med = np.empty(n_clusters, dtype=np.int64)
        total = C.c_double()
        st = ct_stats()
        self._check(self._lib.ct_matrix_consumers(self._h, ptr, M, labels.ctypes.data if labels is not None else None,
                                                  n_clusters, med.ctypes.data if med is not None else None,
                                                  C.byref(total), C.byref(st)), "ct_matrix_consumers")
        return float(total.value), med, st.as_dict()

    LINKAGE_METHODS = {"single": 0, "complete": 1, "average": 2, "ward": 5, "weighted": 6}   # SciPy's codes

    def linkage(self, n_frames, method="average", distmat=None, trusted=False):
        """(Z float64[M-1, 4], stats): scipy.cluster.hierarchy.linkage(distmat, method) on the GPU, for `distmat` (host
        array) or, with distmat=None, the whole matrix the last rmsd_rows_device(0, M) left there.  trusted: as in
        matrix_consumers."""
        M = int(n_frames)
        if method not in self.LINKAGE_METHODS:
            raise ValueError(f"linkage method {method!r} does not run on the GPU (single, complete, average, weighted, ward do)")
        ptr = None
        if distmat is not None:
            given = distmat
            distmat = np.ascontiguousarray(distmat, dtype=np.float64)
            if distmat.shape != (M * (M - 1) // 2,):
                raise ValueError("distmat must be the condensed vector of n_frames frames")
            if not (trusted and self._resident_is(given)):
                ptr = distmat.ctypes.data
                self._note_resident(given if given is distmat else None)
        Z = np.empty((max(M - 1, 0), 4), dtype=np.float64)
        st = ct_stats()
        self._check(self._lib.ct_linkage(self._h, ptr, M, self.LINKAGE_METHODS[method], Z.ctypes.data, C.byref(st)),
                    "ct_linkage")
        return Z, st.as_dict()

    def stats(self):
        st = ct_stats()
        self._lib.ct_get_stats(self._h, C.byref(st))
        return st.as_dict()
