"""Reorder-algorithm handles.

In the reference ``ClustOptions.reorder_alg`` holds a function of the third-party ``rmsd`` package
(``rmsd.reorder_hungarian`` etc., bound at clusttraj/io.py:559-567) and compute_distmat_line calls it on
the CPU for every pair (distmat.py:192,243).  Here the attribute stays a callable so that user code and
tests that compare ``clust_opt.reorder_alg`` by identity keep working, but the callable is only a
HANDLE: its identity selects the assignment kernel that runs inside the GPU per-pair kernel
(csrc/stepwise.cu).  There is deliberately no CPU implementation behind it.
"""


def reorder_hungarian(ref_atoms, mov_atoms, ref_coords, mov_coords):
    """Handle for the per-element Hungarian reordering (``--reorder-alg hungarian``).

    Calling convention of the reference: ``reorder(ref_atoms, mov_atoms, ref_coords, mov_coords)``
    returns ``perm`` with ``mov[perm]`` aligned to ``ref`` (distmat.py:243-244)."""
    raise NotImplementedError(
        "clusttraj_b200 runs the Hungarian reordering inside its CUDA per-pair kernel; this handle is not a "
        "CPU implementation.  Use clusttraj_b200.distmat (or Engine.pair_values(..., want_perms=True)).")


def reorder_distance(ref_atoms, mov_atoms, ref_coords, mov_coords):
    """Handle for the per-element reordering by distance from the centroid (``--reorder-alg distance``,
    rmsd.reorder_distance, bound at io.py:562-563): same calling convention, same GPU-only policy."""
    raise NotImplementedError(
        "clusttraj_b200 runs the distance reordering inside its CUDA per-pair kernel; this handle is not a "
        "CPU implementation.  Use clusttraj_b200.distmat (or Engine.pair_values(..., want_perms=True)).")


GPU_REORDER_MODES = {reorder_hungarian: 1, reorder_distance: 2}


def reorder_mode_of(callable_or_none):
    """0 = no reordering, 1 = Hungarian, 2 = distance.  Any other callable is rejected loudly (the reference's
    brute/qml reorders are outside the accelerated path)."""
    if not callable_or_none:
        return 0
    name = getattr(callable_or_none, "__name__", "")
    if callable_or_none in GPU_REORDER_MODES:
        return GPU_REORDER_MODES[callable_or_none]
    if name == "reorder_hungarian":  # e.g. rmsd.reorder_hungarian from a user's own ClustOptions
        return 1
    if name == "reorder_distance":
        return 2
    raise NotImplementedError(
        f"reorder algorithm {name or callable_or_none!r} is not implemented by the B200 engine "
        "(only --reorder-alg hungarian and distance)")
