#!/usr/bin/env python
"""Benchmark of the all-pairs minimum-RMSD hot path (BASELINE.json metric: RMSD pairs/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg4|cfg5]
                    [--scaling weak|strong] [--no-extras] [--no-cpu-baseline]

One "step" = one pass of the hot path over one synthetic trajectory: centre/pack (K1), the Gram kernel with
its fused RMSD epilogue (K2: the int8 tcgen05 kernel by default, the FP64 DMMA kernel with
CLUSTTRAJ_B200_GRAM_VARIANT=-2) and the fix-up of flagged pairs, producing the rank's slab of the condensed
matrix.  N = 1 runs BASELINE config 2 (20k frames x 100 atoms, plain Kabsch).  N > 1 (torchrun,
one rank per GPU) shards ONE matrix over the ranks by equal-area row slabs with no collective on the data
path; the frame count grows with sqrt(N) so that every GPU keeps config 2's pair count ("weak" scaling).

  value : whole-job pairs/s with the coordinates already resident in HBM, output left in HBM.
  e2e   : the same through the C ABI with HOST buffers: pinned host coordinates -> GPU -> condensed slab
          streamed to pinned host memory, all inside the timed region.
  roofline : against the pipe the dominant kernel runs on -- the algorithmic int8 work of the 15 exact digit GEMMs over the
          measured tcgen05 int8 peak (`frac`; `issued_frac` with the padding of the MMA shape; `fp64_equiv` = 18 K flop per
          pair over the measured DMMA peak); `shared_pipe` = how busy the pipe that the int8 MMAs and the FP64 unit share
          is (tensor + FP64, from the committed ncu capture); `traffic` = DRAM bytes of that capture.  The per-pair
          (Hungarian) kernel reports the HBM figure SURVEY 8d asks for plus `pipes` (FP64 pipe, issue slots, occupancy).
  parity: at EVERY N each rank checks sampled rows of its own slab (as they arrived in pinned host memory in the
          end-to-end leg) against the CPU oracle; the MAX over ranks is printed (tolerance of the north star: 1e-6 A).
  extra.workloads : the other BASELINE configs measured by the same run at the same N, each with value / e2e / roofline /
          parity: config 3 itself (50k x 300, -n) sharded over the N GPUs ("strong"), config 5 itself (100k x 300) when
          N = 8, and config 4 (-ns 60 -e, Hungarian) with its own CPU baseline when N = 1.
  --impl reference : the CPU oracle (restatement of the reference's rmsd-package path) on all host cores,
          on a bounded sample of rows of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (frames, atoms, engine/oracle options, description)
    "cfg2": (20000, 100, dict(), "synthetic 20k frames x 100 atoms, plain Kabsch"),
    "cfg3": (50000, 300, dict(noh=True), "synthetic 50k frames x 300 atoms, -n hydrogen mask"),
    "cfg5": (100000, 300, dict(), "synthetic 100k frames x 300 atoms, plain Kabsch"),
    "cfg4": (10000, 360, dict(ns=60, reorder=True), "synthetic 10k frames, 60 solute + 300 solvent atoms, -ns 60 -e hungarian"),
}
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_fp64_peak_microbench.json")
I8_PEAK_FILE = os.path.join(ROOT, "profiles", "r02_i8_peak_microbench.json")
PEAKS_FILE = os.path.join(ROOT, "MEASURED_PEAKS.json")


def frames_of(name, scaling, world):
    M0 = WORKLOADS[name][0]
    return M0 if scaling == "strong" else int(round(M0 * np.sqrt(world)))


def workload_config(name, scaling, world, K):
    """The `config` object of a JSON line: identical for the b200 arm and the reference arm of the same command."""
    M0, N, kw, desc = WORKLOADS[name]
    M = frames_of(name, scaling, world)
    how = "(the named size, sharded over the GPUs)" if scaling == "strong" else f"(= {M0} x sqrt(n_gpus))"
    return {"workload": f"{name}: {desc}; M={M} {how}, N={N}, K={K}", "frames": M, "atoms": N, "kept_atoms": K, "mode": kw,
            "pairs_total": M * (M - 1) // 2, "sharding": "equal-area row slabs, no collective on the data path",
            "l2": "256 MiB buffer zeroed between steps (L2 flush); every output slab also exceeds the 126 MB L2"}


def make_workload(name, n_frames):
    from clusttraj_b200.synth import make_solute_solvent, make_trajectory

    _, n_atoms, kw, _ = WORKLOADS[name]
    if name == "cfg4":
        return make_solute_solvent(n_frames, n_solute=60, n_solvent_mol=100, seed=20250604)
    return make_trajectory(n_frames, n_atoms, seed=20250600 + int(name[-1]))


def kept_atoms(atoms, kw):
    return int((atoms != 1).sum()) if kw.get("noh") else len(atoms)


def engine_opts(kw):
    from clusttraj_b200 import _engine

    return _engine.make_opts(no_hydrogen=kw.get("noh", False), solute_natoms=kw.get("ns"),
                             reorder=kw.get("reorder", False), reorder_solvent_only=kw.get("rs", False),
                             final_kabsch=kw.get("fk", False), weight_solute=kw.get("ws"))


def load_json(path, default):
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return default


# --------------------------------------------------------------------------------------------- CPU oracle
def oracle_kwargs(kw):
    from oracle import rmsd_restated as R

    return dict(noh=kw.get("noh", False), reorder=R.reorder_hungarian if kw.get("reorder") else None,
                reorder_solvent_only=kw.get("rs", False), nsatoms=kw.get("ns"), weight_solute=kw.get("ws"),
                final_kabsch=kw.get("fk", False))


def cpu_oracle_rate(atoms, coords, kw, rows, cores):
    """pairs/s of the oracle (oracle/distmat_restated.py) over `rows` x all later columns on `cores` processes."""
    from oracle import distmat_restated as O

    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"  # README.md:126-134 of the reference recommends single-threaded BLAS per worker
    t0 = time.perf_counter()
    vec = O.build_distance_matrix(atoms, coords, rows=rows, n_workers=cores, **oracle_kwargs(kw))
    dt = time.perf_counter() - t0
    return len(vec) / dt, len(vec), dt, vec


def sample_rows(M, n_rows, budget_pairs, seed=0):
    """Random rows whose total length stays near the pair budget (bounded CPU sample)."""
    rng = np.random.default_rng(seed)
    rows, total = [], 0
    for r in rng.permutation(M - 1):
        rows.append(int(r))
        total += M - 1 - int(r)
        if len(rows) >= n_rows or total >= budget_pairs:
            break
    return sorted(rows)


def slab_parity(atoms, coords, kw, r0, r1, slab, rank, n_rows=2, n_cols=1536):
    """max |GPU - oracle| over `n_rows` random rows of this rank's slab x up to `n_cols` random later columns each; `slab`
    is the rank's part of the condensed vector as the end-to-end leg left it in pinned host memory.  One process, one
    core per rank (the ranks run side by side)."""
    from oracle import distmat_restated as O

    M = len(coords)
    last = min(r1, M - 1)
    if last <= r0:
        return 0.0, 0
    rng = np.random.default_rng(977 + rank)
    base = r0 * M - r0 * (r0 + 1) // 2
    okw = oracle_kwargs(kw)
    worst, n = 0.0, 0
    for r in sorted(set(int(x) for x in rng.integers(r0, last, size=n_rows))):
        later = np.arange(r + 1, M)
        cols = later if len(later) <= n_cols else np.sort(rng.choice(later, size=n_cols, replace=False))
        want = np.asarray(O.distmat_line(r, atoms, coords, okw["noh"], okw["reorder"], okw["reorder_solvent_only"], okw["nsatoms"],
                                         okw["weight_solute"], np.zeros(0, np.int32), okw["final_kabsch"], columns=cols.tolist()))
        row_off = r * M - r * (r + 1) // 2 - base
        got = slab[row_off + (cols - r - 1)]
        worst = max(worst, float(np.abs(got - want).max()))
        n += len(cols)
    return worst, n


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, power, reasons, smmax = [], [], set(), None
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            # "under load" = samples in the upper half of the power range seen during the timed region
            p = np.array(power); s = np.array(sm)
            busy = s[p >= (p.min() + p.max()) / 2] if p.max() > p.min() else s
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=smmax, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(p.max()))
        return out


# --------------------------------------------------------------------------------------------- reference arm
def run_reference_arm(args, rank, world):
    """The reference's CPU implementation of the path (oracle port; the as-shipped package cannot run here:
    its rmsd / openbabel dependencies are not installable offline) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import distmat_restated as O

    name = args.workload
    kw = WORKLOADS[name][2]
    M = frames_of(name, args.scaling, world)
    cores = O.cpu_count()
    atoms, coords = make_workload(name, M)
    budget = 250_000 if not kw.get("reorder") else 12_000
    times, pairs = [], 0
    for step in range(args.warmup + args.steps):
        rows = sample_rows(M, 4 * cores, budget, seed=step)
        rate, n, dt, _ = cpu_oracle_rate(atoms, coords, kw, rows, cores)
        if step >= args.warmup:
            times.append(dt); pairs += n
    value = pairs / sum(times)
    line = {
        "impl": "reference", "metric": "rmsd_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(name, args.scaling, world, kept_atoms(atoms, kw)),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"each step = {4 * cores} random rows x all later columns of the workload's matrix "
                                   f"(<= {budget} pairs); {pairs} pairs over {args.steps} steps; frames held in memory "
                                   "(the reference re-parses the file per pair)"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- b200 arm
class Job:
    """Process-wide state of the b200 arm: device, process group, rank reductions.  `backend="gloo"` is the test hook of
    tests/test_multirank_cpu.py (host logic of the N > 1 path on CPU, engine replaced by a stand-in); the bench itself always
    runs NCCL on CUDA devices."""

    def __init__(self, rank, world, local_rank, backend="nccl"):
        import torch

        self.torch = torch
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dist = None
        self.cuda = backend == "nccl"
        if self.cuda:
            torch.cuda.set_device(local_rank)
            self.dev = torch.device("cuda", local_rank)
        else:
            self.dev = torch.device("cpu")
        if world > 1:
            import torch.distributed as dist

            if self.cuda:
                dist.init_process_group("nccl", device_id=self.dev)
            else:
                dist.init_process_group("gloo", rank=rank, world_size=world)
            self.dist = dist
        self.flush = torch.empty((256 << 20) if self.cuda else 1, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.cuda:
            self.torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()
        if self.cuda:
            self.torch.cuda.synchronize(self.dev)

    def _reduce(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist else None)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def ncu_pipe_busy(summary_path):
    """tensor / FP64 / shared pipe busy fractions from a committed tools/make_profiles.py summary (None if absent)."""
    return ncu_metrics(summary_path, {"sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed": "busy",
                                      "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor",
                                      "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64"})


def ncu_metrics(summary_path, want):
    """{short name: fraction} for the percentage metrics `want` names, from a tools/make_profiles.py summary."""
    got = {}
    try:
        with open(summary_path) as fh:
            for line in fh:
                parts = line.split()
                if len(parts) >= 2 and parts[0] in want and want[parts[0]] not in got:
                    got[want[parts[0]]] = float(parts[1]) / 100.0
    except (OSError, ValueError):
        return None
    return got if len(got) == len(want) else None


def roofline_of(name, kw, M, K, r0, r1, my_pairs, mean_gram_ms, st):
    """Roofline of the dominant kernel of one launch (this rank's slab).  Pure-Kabsch workloads: the algorithmic work is
    18 K FP64 flop per pair; on the int8 tensor pipe the same contraction is 15 exact digit GEMMs, i.e.
    15 x 18 K int8 operations per pair (csrc/ozaki.cu) -- `frac` is taken against the pipe the kernel runs on, the
    FP64-equivalent figure is kept as `fp64_equiv`."""
    peaks = load_json(PEAKS_FILE, {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0})
    if kw.get("reorder"):
        # per-pair kernel: HBM traffic 2 frames in + 8 B out per pair, latency bound
        bytes_per_launch = my_pairs * (2 * K * 24 + 8)
        achieved = bytes_per_launch / (mean_gram_ms * 1e-3) / 1e9
        out = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
               "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": "pair_kernel",
               "bytes_per_pair": 2 * K * 24 + 8, "pairs_per_launch": my_pairs, "launch_ms": mean_gram_ms,
               "peak_source": "MEASURED_PEAKS.json hbm_gbs" if os.path.exists(PEAKS_FILE) else "fallback 6650 GB/s",
               "note": "latency / issue bound assignment solver (one warp per pair, costs recomputed per scan); the HBM "
                       "figure is reported because SURVEY 8d asks for it; what the kernel is bound by is in `pipes`"}
        cap = os.path.join(ROOT, "profiles", "r02_pair_ncu_summary.txt")
        pipes = ncu_metrics(cap, {"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_busy",
                                  "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active",
                                  "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active"})
        if pipes:
            out["pipes"] = dict(pipes, source=os.path.relpath(cap, ROOT),
                                note="static: ncu --set full capture of pair_kernel on a 300-frame sample of the cfg4 shape")
        return out
    fp64 = load_json(FP64_PEAK_FILE, None)
    fp64_peak = float(fp64["dmma884_sustained_tflops"]) if fp64 else 37.1
    fp64_src = ("measured DMMA m8n8k4 microbenchmark on this pool (profiles/r01_fp64_peak_microbench.json); "
                "MEASURED_PEAKS.json has no FP64 entry") if fp64 else "fallback: nominal B200 FP64"
    fp64_tf = 18.0 * K * my_pairs / (mean_gram_ms * 1e-3) / 1e12
    fp64_equiv = {"achieved": fp64_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_tf / fp64_peak,
                  "flop_per_pair": 18 * K, "peak_source": fp64_src}
    if int(st["gram_path"]) != 2:
        out = {"bound": "tensor", **fp64_equiv, "traffic": None, "pairs_per_launch": my_pairs, "launch_ms": mean_gram_ms,
               "kernel": "gram_rmsd_ws_kernel" if K > 64 else "gram_rmsd_kernel",
               "note": "FP64 tensor pipe (DMMA.8x8x4); algorithmic flops = 18*K per output pair, epilogue and "
                       "padded/diagonal-block work not counted"}
        traffic_file = os.path.join(ROOT, "profiles", "r01_gram_traffic.json")
    else:
        chunks = (K + 31) // 32
        fj = 32 if chunks <= 4 else (16 if chunks <= 10 else 26)   # column-tile frames of the kernel variant (csrc/ozaki.cu)
        n_mma = 80 if chunks > 10 else 3 * fj
        ntj = (M + fj - 1) // fj
        last = min(r1, M - 1)
        blocks = sum(ntj - (40 * ti + 1) // fj for ti in range(r0 // 40, (last + 39) // 40))
        issued = blocks * 15 * chunks * 2.0 * 128 * n_mma * 32 / (mean_gram_ms * 1e-3) / 1e12
        i8 = load_json(I8_PEAK_FILE, None)
        if i8:
            i8_peak = float(i8["i8_dense_tops"])
            i8_src = "measured tcgen05.mma kind::i8 microbenchmark on this pool (profiles/r02_i8_peak_microbench.json)"
        else:
            i8_peak = 2.0 * float(peaks["bf16_tflops"])
            i8_src = "2 x the measured dense bf16 rate of MEASURED_PEAKS.json (no int8 measurement file present)"
        i8_tops = 15 * 18.0 * K * my_pairs / (mean_gram_ms * 1e-3) / 1e12
        out = {"bound": "tensor", "achieved": i8_tops, "peak": i8_peak, "unit": "TFLOP/s", "frac": i8_tops / i8_peak,
               "traffic": None, "pairs_per_launch": my_pairs, "launch_ms": mean_gram_ms,
               "ops_per_pair": 15 * 18 * K, "peak_source": i8_src,
               "kernel": "gram_i8_ts_kernel<8>" if chunks <= 4 else ("gram_i8_ts_kernel<4>" if chunks <= 10 else "gram_i8_kernel"),
               "pipe": f"int8 tcgen05.mma (SASS UTCIMMA), M128 N{n_mma} K32; unit = 1e12 int8 multiply-add operations x 2 per s",
               "issued_tops": issued, "issued_frac": issued / i8_peak,
               "fp64_equiv": fp64_equiv, "quant_step_angstrom": float(st["quant_step"]),
               "note": "algorithmic work = the 15 exact int8 digit GEMMs that rebuild the FP64 Gram product (15 x 18 K operations "
                       "per output pair) against the measured int8 tensor peak; `issued_*` adds the padding the MMA shape "
                       "forces (atoms to 32, 40-frame row tiles in 128 lanes, diagonal blocks); `fp64_equiv` is the same launch "
                       "counted as 18 K FP64 flop per pair against the FP64 DMMA peak (not bounded by 1: another pipe does "
                       "the work).  The kernel is bound by its FP64 recombination + 3x3 solve epilogue, see profiles/"}
        # what the kernel is bound by: on B200 the int8 tcgen05 MMAs and the FP64 unit are one pipe (profiles/r02_shared_pipe_probe.json),
        # so its busy fraction = tensor + FP64 (recombination and 3x3 solve) -- from the committed ncu capture of the same kernel
        if chunks <= 10:
            cap = os.path.join(ROOT, "profiles", "r02_gram_i8_ncu_summary.txt" if chunks <= 4 else "r02_gram_i8_k300_ncu_summary.txt")
            busy = ncu_pipe_busy(cap)
            if busy:
                out["shared_pipe"] = dict(busy, source=os.path.relpath(cap, ROOT),
                                          note="static: ncu --set full capture of this kernel "
                                               + ("on cfg2" if chunks <= 4 else "on 16000 frames x 300 atoms (the same kernel at this run's atom count differs in the MMA share)")
                                               + "; busy = sm__pipe_shared_cycles_active = tensor + fp64: the FP64 unit and the int8 MMAs exclude each other")
        traffic_file = os.path.join(ROOT, "profiles", "r02_gram_i8_traffic.json" if chunks <= 4 else "r02_gram_i8_k300_traffic.json")
        if not os.path.exists(traffic_file):
            traffic_file = traffic_file.replace("r02_", "r01_")
    tr = load_json(traffic_file, None)
    if tr and (name == "cfg2" or tr.get("per_pair")):
        # committed `ncu --set full` capture of the same kernel: bytes per launch on the capture's workload; scaled per pair
        # for the other workloads of the same kernel shape
        if name == "cfg2" and not tr.get("per_pair"):
            out["traffic"] = tr.get("dram_bytes_per_launch")
        else:
            out["traffic"] = float(tr["per_pair"]) * my_pairs
        out["traffic_source"] = os.path.relpath(traffic_file, ROOT)
    return out


def measure(job, name, scaling, steps, warmup, sample_clocks=False, light_warmup=False):
    """Kernel-only leg, end-to-end leg and per-rank parity of one workload at the job's N.  Returns a dict on every rank
    (rank-reduced numbers are identical on all ranks)."""
    from clusttraj_b200 import _engine
    from clusttraj_b200.distmat import slab_bounds

    rank, world = job.rank, job.world
    kw = WORKLOADS[name][2]
    M = frames_of(name, scaling, world)
    atoms, coords = make_workload(name, M)
    N = coords.shape[1]
    K = kept_atoms(atoms, kw)
    cuts = slab_bounds(M, world)
    r0, r1 = cuts[rank], cuts[rank + 1]
    eng = _engine.Engine(job.local_rank)
    opts = engine_opts(kw)
    my_pairs = int(eng._lib.ct_rows_pairs(M, r0, r1))

    pinned_in = _engine.pinned_empty(coords.size).reshape(coords.shape)
    pinned_in[...] = coords
    pinned_out = _engine.pinned_empty(my_pairs)

    # ---- kernel-only leg: coordinates resident in HBM, slab left in HBM -------------------------------
    eng.load_frames(pinned_in, atoms, opts)
    launches = 0
    gram_ms, step_ms = [], []
    sampler, clocks, st = None, None, None
    wall0 = time.perf_counter()
    for step in range(warmup + steps):
        job.flush.zero_()
        if step == warmup:
            job.barrier()
            sampler = ClockSampler(job.local_rank) if (rank == 0 and sample_clocks) else None
            wall0 = time.perf_counter()
        pack_ms = eng.pack_resident()
        if light_warmup and step < warmup:
            # a pass of this workload takes seconds (per-pair Hungarian kernel): warm up on the last rows of the slab
            lo = max(r0, min(r1, M - 1) - 48)
            _, _, st = eng.rmsd_rows_device(lo, r1)
        else:
            _, _, st = eng.rmsd_rows_device(r0, r1)
        if step >= warmup:
            step_ms.append(pack_ms + st["kernel_ms"])
            gram_ms.append(st["gram_ms"] if st["gram_ms"] > 0 else st["pair_ms"])
            # kernels of ct_pack_resident: centre (+ tile pack for the DMMA kernel | + max, exponent, quantise/slice, G for int8)
            launches += int(st["launches"]) + {0: 1, 1: 2, 2: 5}[int(st["gram_path"])]
    job.barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    flagged = int(st["flagged"])
    dev_ms = job.max(sum(step_ms))          # device time of the K steps, max over ranks
    total_pairs = job.sum(float(my_pairs))
    value = total_pairs * steps / (dev_ms * 1e-3)

    # ---- end-to-end leg: host buffers in, host buffers out, through the C ABI -------------------------
    e2e_ms = []
    for step in range((1 if light_warmup else warmup) + steps):
        job.flush.zero_()
        job.barrier()
        t0 = time.perf_counter()
        eng.load_frames(pinned_in, atoms, opts)           # H2D of the step's input + K1
        _, st2 = eng.rmsd_rows(r0, r1, pinned_out)        # K2 + fix-up, slab streamed to pinned host memory
        dt = 1e3 * (time.perf_counter() - t0)
        if step >= (1 if light_warmup else warmup):
            e2e_ms.append(dt)  # host wall clock around the two synchronous C-ABI calls
    job.barrier()
    if sampler is not None:
        clocks = sampler.stop()  # sampled over both timed legs (kernel-only and end-to-end)
    e2e_total = job.max(sum(e2e_ms))
    e2e_value = total_pairs * steps / (e2e_total * 1e-3)

    # ---- parity of what the end-to-end leg delivered, every rank, MAX over ranks -------------------------
    worst, checked = slab_parity(atoms, coords, kw, r0, r1, pinned_out, rank)
    parity = {"max_abs_diff": job.max(worst), "pairs_checked": int(job.sum(float(checked))), "tolerance": 1e-6,
              "how": "every rank: 2 random rows of its own slab x up to 1536 random later columns, CPU oracle "
                     "(oracle/distmat_restated.py) vs the slab the end-to-end leg streamed into pinned host memory; MAX over ranks"}
    parity["ok"] = bool(parity["max_abs_diff"] <= parity["tolerance"])
    checksum = job.sum(float(pinned_out[: min(my_pairs, 1 << 20)].sum()))

    res = {
        "name": name, "scaling": scaling, "M": M, "N": int(N), "K": K, "kw": kw, "rows": (r0, r1), "my_pairs": my_pairs,
        "total_pairs": int(total_pairs), "value": value, "ms_per_step": dev_ms / steps, "steps": steps, "warmup": warmup,
        "e2e_value": e2e_value, "e2e_ms_per_step": e2e_total / steps, "h2d": int(coords.nbytes + atoms.nbytes),
        "d2h": int(8 * my_pairs), "launches": int(job.sum(float(launches))), "clocks": clocks, "wall_ms": wall_ms, "flagged": flagged,
        "checksum": checksum, "parity": parity,
        "roofline": roofline_of(name, kw, M, K, r0, r1, my_pairs, float(np.mean(gram_ms)), st),
        "_state": (eng, atoms, coords, pinned_in, opts),
    }
    return res


def cpu_baseline_of(res, budget, n_rows):
    """Bounded CPU sample of the same workload on all host cores (rank 0, N = 1): the oracle port timed, and its rows
    compared with the GPU's."""
    from oracle import distmat_restated as O

    eng, atoms, coords, pinned_in, opts = res["_state"]
    cores = O.cpu_count()
    rows = sample_rows(res["M"], n_rows * cores, budget)
    rate, n, dt, vec = cpu_oracle_rate(atoms, coords, res["kw"], rows, cores)
    eng.load_frames(pinned_in, atoms, opts)
    worst, pos = 0.0, 0
    for r in rows:
        line, _ = eng.rmsd_rows(r, r + 1)
        worst = max(worst, float(np.abs(line - vec[pos: pos + len(line)]).max()))
        pos += len(line)
    return {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{len(rows)} random rows x all later columns = {n} pairs in {dt:.1f} s; oracle port with "
                      "frames in memory, multiprocessing.Pool over rows like distmat.py:73-76",
            "max_abs_diff_vs_gpu": worst}


def extra_entry(res, world, cpu=None):
    out = {"config": workload_config(res["name"], res["scaling"], world, res["K"]), "scaling": res["scaling"], "n_gpus": world,
           "steps": res["steps"], "warmup": res["warmup"], "value": res["value"], "unit": "pairs/s",
           "ms_per_step": res["ms_per_step"],
           "e2e": {"value": res["e2e_value"], "unit": "pairs/s", "ms_per_step": res["e2e_ms_per_step"],
                   "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"]},
           "roofline": res["roofline"], "parity": res["parity"], "gpu_launches": res["launches"],
           "flagged_pairs_per_step": res["flagged"]}
    if cpu is not None:
        out["cpu_baseline"] = cpu
    return out


def run_b200_arm(args, rank, world, local_rank, backend="nccl"):
    job = Job(rank, world, local_rank, backend)
    name = args.workload
    head = measure(job, name, args.scaling, args.steps, args.warmup, sample_clocks=True, light_warmup=(name == "cfg4"))

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        reorder = bool(WORKLOADS[name][2].get("reorder"))
        cpu = cpu_baseline_of(head, 16_000 if reorder else 1_500_000, 8)
    head.pop("_state")[0].close()

    # ---- the other BASELINE configs at the same N (the headline above stays comparable between rounds) -----
    extras = {}
    if not args.no_extras and name == "cfg2" and args.scaling == "weak":
        plan = [("cfg3", "strong", min(args.steps, 3), 3, False)]
        if world == 8:
            plan.append(("cfg5", "strong", min(args.steps, 3), 3, False))
        if world == 1:
            plan.append(("cfg4", "strong", 1, 3, True))
        for wname, scaling, steps, warmup, light in plan:
            try:
                r = measure(job, wname, scaling, steps, warmup, light_warmup=light)
                wcpu = None
                if wname == "cfg4" and rank == 0 and not args.no_cpu_baseline:
                    wcpu = cpu_baseline_of(r, 16_000, 8)
                r.pop("_state")[0].close()
                entry = extra_entry(r, world, wcpu)
                if light:
                    entry["warmup_note"] = ("one pass takes seconds: the 3 kernel warm-up launches run the last 48 rows of the "
                                            "matrix, the end-to-end leg has 1 full warm-up pass")
                extras[wname] = entry
            except Exception as exc:   # an extra must never take the headline down with it
                extras[wname] = {"error": f"{type(exc).__name__}: {exc}"}
            job.barrier()

    if rank == 0:
        cfg = workload_config(name, args.scaling, world, head["K"])
        line = {
            "metric": "rmsd_pairs_per_s", "value": head["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": head["roofline"],
            "cpu_baseline": cpu,
            "e2e": {"value": head["e2e_value"], "unit": "pairs/s", "h2d_bytes_per_step": head["h2d"],
                    "d2h_bytes_per_step": head["d2h"], "ms_per_step": head["e2e_ms_per_step"],
                    "h2d_bytes_per_step_all_ranks": head["h2d"] * world, "d2h_bytes_per_step_all_ranks": 8 * head["total_pairs"],
                    "note": "ct_load_frames (pinned H2D + K1) + ct_rmsd_rows (K2, fix-up, chunked D2H into pinned "
                            "host memory, overlapped); bytes are per rank"},
            "parity": head["parity"],
            "gpu_launches": head["launches"],
            "clocks": head["clocks"],
            "wall_ms_timed_region": head["wall_ms"],
            "flagged_pairs_per_step": head["flagged"],
            "checksum": head["checksum"],
            "extra": {"workloads": extras},
        }
        print(json.dumps(line), flush=True)
    job.close()
    return line if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra.workloads block (config 3 / 5 / 4 at the same N)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = frames x sqrt(N) (fixed pairs per GPU, default); strong = the named matrix sharded over the GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
