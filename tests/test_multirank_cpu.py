"""World-size-2 gloo tests (CPU) of the N > 1 path's host logic.

1. Every rank takes the row slab bench.py / clusttraj_b200.distmat give it, the slabs are disjoint, cover the
   condensed vector in order, and the rank-reduced totals are what bench.py reports.
2. bench.py's own b200 arm end to end (headline + extra.workloads + per-rank parity reduced with MAX over the ranks)
   with the CUDA engine replaced by a stand-in that answers from the oracle: the JSON line has every key of the
   contract, at N = 1 and N = 2, and a corrupted slab on ONE rank shows up in the printed parity.
The per-slab values come from the oracle here (no GPU)."""
import json
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


# ---- bench.py's b200 arm over a stand-in engine ---------------------------------------------------------------------
class _StandInEngine:
    """Answers the calls bench.measure makes from the oracle (tests only: the product has no CPU path)."""
    corrupt_rank = None

    def __init__(self, device=0):
        from clusttraj_b200 import _engine

        self._lib = _engine.load_library()   # ct_rows_pairs is host arithmetic
        self.device = device

    def load_frames(self, coords, atoms, opts):
        self.coords, self.atoms, self.opts = np.array(coords), np.array(atoms), opts
        self.M = len(coords)

    def pack_resident(self):
        return 0.01

    def _rows(self, r0, r1):
        import bench
        from oracle import distmat_restated as O

        kw = dict(noh=bool(self.opts.no_hydrogen), ns=self.opts.solute_natoms or None, reorder=bool(self.opts.reorder_mode))
        return O.build_distance_matrix(self.atoms, self.coords, rows=range(r0, min(r1, self.M)), **bench.oracle_kwargs(kw))

    def _stats(self, n):
        path = 0 if self.opts.reorder_mode else 2
        return dict(kernel_ms=1.0, gram_ms=1.0 if path else 0.0, pair_ms=0.0 if path else 1.0, launches=2 if path else 1,
                    flagged=0, gram_path=path, quant_step=2.0 ** -25, pairs=n)

    def rmsd_rows_device(self, r0, r1):
        n = int(self._lib.ct_rows_pairs(self.M, r0, r1))
        return 0, n, self._stats(n)

    def rmsd_rows(self, r0, r1, out=None):
        vals = self._rows(r0, r1)
        if out is None:
            out = np.empty(len(vals))
        out[: len(vals)] = vals
        if self.corrupt_rank is not None and int(os.environ.get("RANK", "0")) == self.corrupt_rank and r1 - r0 > 1:
            out[: len(vals)] += 1e-3
        return out[: len(vals)], self._stats(len(vals))

    def close(self):
        pass


def _bench_line(rank, world, port, tmp, corrupt_rank):
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    from clusttraj_b200 import _engine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    bench.WORKLOADS = {
        "cfg2": (26, 10, dict(), "tiny plain"),
        "cfg3": (20, 12, dict(noh=True), "tiny -n"),
        "cfg5": (20, 12, dict(), "tiny plain 2"),
        "cfg4": (6, 360, dict(ns=60, reorder=True), "tiny hungarian"),
    }
    _StandInEngine.corrupt_rank = corrupt_rank
    _engine.Engine = _StandInEngine
    _engine.pinned_empty = lambda n, dtype=np.float64: np.empty(int(n), dtype=dtype)
    bench.ClockSampler = lambda idx: type("S", (), {"stop": lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}})()
    args = argparse.Namespace(workload="cfg2", scaling="weak", steps=2, warmup=3, no_cpu_baseline=False, no_extras=False)
    line = bench.run_b200_arm(args, rank, world, 0, backend="gloo")
    if rank == 0:
        with open(os.path.join(tmp, "line.json"), "w") as fh:
            json.dump(line, fh)


def _check_line(line, world):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "parity", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == world and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert line["roofline"]["frac"] == pytest.approx(line["roofline"]["achieved"] / line["roofline"]["peak"])
    assert "fp64_equiv" in line["roofline"]
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["config"]["pairs_total"] == line["config"]["frames"] * (line["config"]["frames"] - 1) // 2
    extras = line["extra"]["workloads"]
    assert "cfg3" in extras and "error" not in extras["cfg3"], extras
    assert extras["cfg3"]["scaling"] == "strong" and extras["cfg3"]["config"]["frames"] == 20
    for e in extras.values():
        assert e["parity"]["pairs_checked"] > 0 and "roofline" in e and "e2e" in e
    return extras


def test_bench_line_single_rank(tmp_path):
    """N = 1: headline with cpu_baseline + parity, extras cfg3 (strong) and cfg4 (with its own CPU baseline)."""
    import torch.multiprocessing as mp

    mp.spawn(_bench_line, args=(1, 29900 + (os.getpid() % 90), str(tmp_path), None), nprocs=1, join=True)
    line = json.load(open(tmp_path / "line.json"))
    extras = _check_line(line, 1)
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["max_abs_diff_vs_gpu"] < 1e-9
    assert line["parity"]["ok"] and line["parity"]["max_abs_diff"] < 1e-9
    assert "cfg4" in extras and extras["cfg4"]["cpu_baseline"]["value"] > 0 and extras["cfg4"]["parity"]["ok"]
    assert extras["cfg4"]["roofline"]["bound"] == "hbm"
    assert "cfg5" not in extras


@pytest.mark.parametrize("corrupt_rank", [None, 1])
def test_bench_line_two_ranks_parity_is_reduced_over_ranks(tmp_path, corrupt_rank):
    """N = 2 over gloo: the printed parity is the MAX over the ranks (an error on rank 1 only must surface on rank 0's line)."""
    import torch.multiprocessing as mp

    mp.spawn(_bench_line, args=(2, 29700 + (os.getpid() % 90) + (7 if corrupt_rank else 0), str(tmp_path), corrupt_rank),
             nprocs=2, join=True)
    line = json.load(open(tmp_path / "line.json"))
    extras = _check_line(line, 2)
    assert line["cpu_baseline"] is None and "cfg4" not in extras
    assert line["config"]["frames"] == int(round(26 * np.sqrt(2)))
    if corrupt_rank is None:
        assert line["parity"]["ok"] and line["parity"]["max_abs_diff"] < 1e-9
    else:
        assert not line["parity"]["ok"] and line["parity"]["max_abs_diff"] == pytest.approx(1e-3, rel=1e-6)
