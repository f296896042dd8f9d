"""CPU oracle for the clusttraj all-pairs minimum-RMSD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product path (``clusttraj_b200``) never imports it
and fails loudly when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * ``-n -ns 17`` without reordering is pinned bit-for-bit by the reference's
    own golden vector ``test/ref/test_distmat.npy`` (copied to
    ``tests/golden/ref_test_distmat.npy``).
  * Every other mode is pinned against the UNMODIFIED reference file
    ``/root/reference/clusttraj/distmat.py`` executed in the authoring
    container over import shims (``oracle/ref_shims``) for its two
    uninstallable third-party imports (``rmsd``, ``openbabel``); the vectors
    and the generating script (``oracle/make_golden.py``) are committed under
    ``tests/golden/``.  The arithmetic inside the ``rmsd`` shim is this
    package's restatement of charnley/rmsd (``rmsd>=1.6.3``, not vendored in
    the reference, not installable offline), so for Hungarian reordering /
    ``-ws`` / ``--final-kabsch`` the third-party arithmetic itself is
    "parity unpinned" beyond SciPy's live ``linear_sum_assignment``.
"""
