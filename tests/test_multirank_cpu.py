"""World-size-2 gloo test (CPU) of the N > 1 path's host logic: every rank takes the row slab bench.py /
clusttraj_b200.distmat give it, the slabs are disjoint, cover the condensed vector in order, and the
rank-reduced totals are what bench.py reports.  The per-slab values come from the oracle here (no GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from clusttraj_b200 import _engine
    from clusttraj_b200.distmat import slab_bounds
    from clusttraj_b200.synth import make_trajectory
    from oracle import distmat_restated as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = 40
    atoms, coords = make_trajectory(M, 12, seed=5)
    cuts = slab_bounds(M, world)
    r0, r1 = cuts[rank], cuts[rank + 1]
    lib = _engine.load_library()
    n_mine = lib.ct_rows_pairs(M, r0, r1)
    slab = O.build_distance_matrix(atoms, coords, rows=range(r0, r1))
    assert len(slab) == n_mine
    t = torch.tensor([float(n_mine), 1.0 + rank], dtype=torch.float64)
    total = t.clone(); dist.all_reduce(total, op=dist.ReduceOp.SUM)
    worst = t.clone(); dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    assert int(total[0].item()) == M * (M - 1) // 2
    assert worst[1].item() == float(world)  # max over ranks, as bench.py takes for the time
    np.save(os.path.join(tmp, f"slab{rank}.npy"), slab)
    dist.barrier()
    if rank == 0:
        full = O.build_distance_matrix(atoms, coords)
        parts = np.concatenate([np.load(os.path.join(tmp, f"slab{r}.npy")) for r in range(world)])
        assert np.array_equal(parts, full)
    dist.destroy_process_group()


def test_two_rank_slabs_cover_the_matrix(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
