"""Option carrier, logger and command-line parsing with the reference's names and semantics
(clusttraj/io.py:42-579), re-hosted without its hard imports of ``rmsd`` / ``openbabel`` (neither is
needed: reordering runs inside the CUDA engine and XYZ / PDB / GRO are read natively), and the structure writers
(io.py:582-1015: align_mol, save_clusters_config, save_medoids_config) with their arithmetic on the GPU
(ct_pair_transforms) and the .xyz text written natively (ct_format_xyz); other output formats need OpenBabel and are
refused loudly, as are the matplotlib plots (-p)."""
import argparse
import logging
import os
import sys
from dataclasses import dataclass, fields
from typing import Callable, Optional

import numpy as np

from . import reorder as _reorder

LINKAGE_METHODS = ("single", "complete", "average", "weighted", "centroid", "median", "ward")
REORDER_ALGS = ("inertia-hungarian", "hungarian", "brute", "distance", "qml")  # names the reference accepts
NATIVE_INPUT_FORMATS = ("xyz", "pdb", "ent", "gro")   # read by clusttraj_b200/xyz.py (ct_read_xyz / _pdb / _gro), case-insensitive
NATIVE_OUTPUT_FORMATS = ("xyz",)                      # written by ct_format_xyz

header_string = """
clusttraj (B200 engine) - a solvent-informed tool for clustering trajectories.
All-pairs minimum-RMSD matrix computed on NVIDIA B200 GPUs; clustering by SciPy.

Reference implementation and method: Ribeiro, R. B. & Cezar, H. M., J. Chem. Theory Comput. 21, 6759-6768 (2025).
"""


@dataclass
class ClustOptions:
    """Same fields as the reference's ClustOptions (io.py:42-83)."""

    trajfile: Optional[str] = None
    min_rmsd: Optional[float] = None
    n_workers: Optional[int] = None
    method: Optional[str] = None
    reorder_alg_name: Optional[str] = None
    reorder_alg: Optional[Callable] = None
    out_conf_fmt: Optional[str] = None
    reorder: Optional[bool] = None
    reorder_solvent_only: Optional[bool] = None
    exclusions: Optional[bool] = None
    no_hydrogen: Optional[bool] = None
    input_distmat: Optional[bool] = None
    save_confs: Optional[bool] = None
    save_medoids: Optional[bool] = None
    plot: Optional[bool] = None
    opt_order: Optional[bool] = None
    overwrite: Optional[bool] = None
    final_kabsch: Optional[bool] = None
    silhouette_score: Optional[bool] = None
    n_clusters: Optional[int] = None
    metrics: Optional[bool] = None
    distmat_name: Optional[str] = None
    out_clust_name: Optional[str] = None
    evo_name: Optional[str] = None
    mds_name: Optional[str] = None
    dendrogram_name: Optional[str] = None
    out_conf_name: Optional[str] = None
    out_medoids_name: Optional[str] = None
    out_medoids_fmt: Optional[str] = None
    summary_name: Optional[str] = None
    solute_natoms: Optional[int] = None
    weight_solute: Optional[float] = None
    reorder_excl: Optional[np.ndarray] = None
    optimal_cut: Optional[np.ndarray] = None
    verbose: Optional[bool] = None

    def update(self, new: dict) -> None:
        known = {f.name for f in fields(self)}
        for key, value in new.items():
            if key in known:
                setattr(self, key, value)

    def __str__(self) -> str:
        out = [header_string, "Full command: " + " ".join(sys.argv), "",
               f"Clusterized from trajectory file: {self.trajfile}", f"Method: {self.method}"]
        if self.silhouette_score:
            cut = self.optimal_cut
            if isinstance(cut, (np.ndarray, list)):
                cut = cut[0]
            elif not isinstance(cut, (float, np.floating)):
                raise ValueError("optimal_cut must be a float or np.ndarray")
            out += ["", "Using silhouette score", f"RMSD criterion found by silhouette: {cut}"]
        elif self.n_clusters is not None:
            out.append(f"Number of clusters criterion: {self.n_clusters}")
        else:
            out.append(f"RMSD criterion: {self.min_rmsd}")
        out.append(f"Ignoring hydrogens?: {self.no_hydrogen}")
        if self.reorder:
            out.append("")
            if self.solute_natoms:
                out.append("Using solute-solvent adjustment/reordering")
                if self.final_kabsch:
                    out.append("Using final Kabsch rotation before computing RMSD")
                out.append(f"Number of solute atoms: {self.solute_natoms}")
                out.append(f"Weight of the solute atoms: {self.weight_solute}" if self.weight_solute
                           else "Unweighted RMSD according to solute/solvent.")
                if self.reorder_solvent_only:
                    out.append("Reordering only solvent atoms")
            else:
                out.append("Reordering all atom at the same time")
            out.append(f"Reordering algorithm: {self.reorder_alg_name}")
            if self.exclusions:
                out.append("Atoms that weren't considered in the reordering: "
                           + " ".join(str(int(v)) for v in self.reorder_excl))
        out.append("")
        verb = "read from" if self.input_distmat else "written in"
        out.append(f"RMSD matrix was {verb}: {self.distmat_name}")
        out.append(f"The classification of each configuration was written in: {self.out_clust_name}")
        return "\n".join(out) + "\n"


class Logger:
    """Process-wide logger: file + stdout, INFO (reference: io.py:169-196)."""

    logformat = "%(asctime)s %(levelname)-8s [%(filename)s:%(lineno)d] <%(funcName)s> %(message)s"
    formatter = logging.Formatter(fmt=logformat)
    logger = logging.getLogger("clusttraj_b200")

    @classmethod
    def setup(cls, logfile: str) -> None:
        cls.logger.setLevel(logging.INFO)
        for h in list(cls.logger.handlers):
            cls.logger.removeHandler(h)
        cls.fh = logging.FileHandler(logfile)
        cls.fh.setLevel(logging.DEBUG)
        cls.fh.setFormatter(cls.formatter)
        cls.ch = logging.StreamHandler(sys.stdout)
        cls.ch.setLevel(logging.INFO)
        cls.ch.setFormatter(cls.formatter)
        cls.logger.addHandler(cls.fh)
        cls.logger.addHandler(cls.ch)
        cls.logger.info(header_string)


def check_positive(value: str) -> int:
    number = int(value)
    if number <= 0:
        raise argparse.ArgumentTypeError(f"{value} is an invalid positive int value")
    return number


def extant_file(x: str) -> str:
    if not os.path.exists(x):
        raise argparse.ArgumentTypeError(f"{x} does not exist")
    return x


def _parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(
        prog="clusttraj_b200",
        description="Hierarchical clustering of a trajectory on the minimum-RMSD matrix (computed on B200 GPUs).")
    p.add_argument("trajectory_file", type=extant_file, help="trajectory file (.xyz, .pdb / .ent, .gro natively)")
    p.add_argument("-f", "--force-overwrite", dest="force", action="store_true")
    p.add_argument("-np", "--nprocesses", type=check_positive, default=4,
                   help="kept for compatibility; the engine uses GPUs, not worker processes")
    p.add_argument("-n", "--no-hydrogen", action="store_true")
    p.add_argument("-p", "--plot", action="store_true")
    p.add_argument("-m", "--method", default="average")
    p.add_argument("-cc", "--clusters-configurations")
    p.add_argument("-mc", "--medoids-configurations")
    p.add_argument("-oc", "--outputclusters", default="clusters.dat")
    p.add_argument("-e", "--reorder", action="store_true")
    p.add_argument("-eex", "--reorder-exclusions", nargs="+", type=check_positive)
    p.add_argument("-rs", "--reorder-solvent-only", action="store_true")
    p.add_argument("--reorder-alg", action="store", default="hungarian")
    p.add_argument("-ns", "--natoms-solute", type=check_positive)
    p.add_argument("-ws", "--weight-solute", type=float)
    p.add_argument("-odl", "--optimal-ordering", action="store_true")
    p.add_argument("--final-kabsch", action="store_true")
    p.add_argument("--log", type=str, default="clusttraj.log")
    p.add_argument("--metrics", action="store_true")
    p.add_argument("-v", "--verbose", action="store_true")
    crit = p.add_mutually_exclusive_group(required=True)
    crit.add_argument("-ss", "--silhouette-score", action="store_true")
    crit.add_argument("-rmsd", "--min-rmsd", type=float)
    crit.add_argument("-nc", "--n-clusters", type=check_positive)
    files = p.add_mutually_exclusive_group()
    files.add_argument("-i", "--input", type=extant_file)
    files.add_argument("-od", "--outputdistmat", default="distmat.npy")
    return p


def configure_runtime(args_in) -> ClustOptions:
    """argv -> validated ClustOptions (reference: io.py:234-493)."""
    parser = _parser()
    args = parser.parse_args(args_in)
    Logger.setup(args.log)
    ext = os.path.splitext(args.trajectory_file)[1][1:]
    if ext.lower() not in NATIVE_INPUT_FORMATS:
        # formats without a native reader go through OpenBabel when it is installed (the reference's check, io.py:424-427)
        try:
            from openbabel import pybel
            ok = ext in pybel.informats
        except Exception:
            ok = False
        if not ok:
            parser.error(f"The trajectory file format ({ext}) is not supported.")
    if args.method not in LINKAGE_METHODS:
        parser.error(f"The method you selected with -m ({args.method}) is not valid.")
    if args.reorder_alg not in REORDER_ALGS:
        parser.error(f"The reorder method you selected with --reorder-method ({args.reorder_alg}) is not valid.")
    if args.reorder and args.reorder_alg in ("brute", "qml"):
        parser.error(f"--reorder-alg {args.reorder_alg} is not implemented by the B200 engine (use hungarian or distance).")
    if args.clusters_configurations and args.clusters_configurations not in NATIVE_OUTPUT_FORMATS:
        parser.error("The format you selected to save the clustered superposed configurations "
                     f"({args.clusters_configurations}) is not valid: clusttraj_b200 writes xyz (other formats need OpenBabel).")
    if args.medoids_configurations and args.medoids_configurations not in NATIVE_OUTPUT_FORMATS:
        parser.error(f"The format you selected to save the superposed medoids ({args.medoids_configurations}) is not valid: "
                     "clusttraj_b200 writes xyz (other formats need OpenBabel).")
    if args.plot:
        parser.error("-p needs matplotlib plots, which are outside the B200 engine (SURVEY.md: out of scope).")
    if args.reorder_exclusions and args.no_hydrogen:
        parser.error("You cannot use --no-hydrogen and reorder exclusions with -eex at the same time")
    if not args.reorder and args.reorder_exclusions:
        logging.warning("The list of atoms to exclude for reordering only makes sense if reordering is enabled. "
                        "Ignoring the list.")
    if args.reorder_solvent_only and not args.natoms_solute:
        parser.error("You must specify the number of solute atoms with -ns to use the -rs option.")
    if args.weight_solute and not args.natoms_solute:
        parser.error("You must specify the number of solute atoms with -ns to use the -ws option.")
    if args.weight_solute and not 0.0 <= args.weight_solute <= 1.0:
        parser.error("The weight of the solute atoms must be between 0 and 1.")
    return parse_args(args)


def parse_args(args: argparse.Namespace) -> ClustOptions:
    """Namespace -> ClustOptions (reference: io.py:496-579): derives the output names, turns the 1-based
    -eex list into 0-based int32 and binds the reorder handle."""
    base = os.path.splitext(args.outputclusters)[0]
    excl = args.reorder_exclusions or []
    alg = None
    if args.reorder and args.reorder_alg == "hungarian":
        alg = _reorder.reorder_hungarian  # "inertia-hungarian" binds nothing in the reference either (io.py:559-567)
    elif args.reorder and args.reorder_alg == "distance":
        alg = _reorder.reorder_distance   # io.py:562-563
    if not args.input and os.path.exists(args.outputdistmat) and not args.force:
        raise FileExistsError(
            f"File {args.outputdistmat} already exists, specify a new filename with the -od command option or use "
            "the -f option. If you are trying to read the RMSD matrix from a file, use the -i option.")
    if os.path.exists(args.outputclusters) and not args.force:
        raise FileExistsError(
            f"File {args.outputclusters} already exists, specify a new filename with the -oc command option or use "
            "the -f option.")
    return ClustOptions(
        trajfile=args.trajectory_file, min_rmsd=args.min_rmsd, n_workers=args.nprocesses, method=args.method,
        reorder_alg_name=args.reorder_alg, reorder_alg=alg, out_conf_fmt=args.clusters_configurations or None,
        reorder=bool(args.reorder), reorder_solvent_only=bool(args.reorder_solvent_only), exclusions=bool(excl),
        no_hydrogen=args.no_hydrogen, input_distmat=bool(args.input), save_confs=bool(args.clusters_configurations),
        save_medoids=bool(args.medoids_configurations), plot=bool(args.plot), opt_order=args.optimal_ordering,
        overwrite=args.force, final_kabsch=args.final_kabsch, silhouette_score=args.silhouette_score,
        n_clusters=args.n_clusters, metrics=args.metrics,
        distmat_name=args.input if args.input else args.outputdistmat, out_clust_name=args.outputclusters,
        evo_name=base + "_evo.pdf" if args.plot else None, mds_name=base + "_mds.pdf" if args.plot else None,
        dendrogram_name=base + "_dendrogram.pdf" if args.plot else None,
        out_conf_name=base + "_confs" if args.clusters_configurations else None,
        out_medoids_name=base + "_medoids" if args.medoids_configurations else None,
        out_medoids_fmt=args.medoids_configurations or None, summary_name=base + ".out",
        solute_natoms=args.natoms_solute, weight_solute=args.weight_solute,
        reorder_excl=np.asarray([x - 1 for x in excl], np.int32), verbose=args.verbose)


# ---- structure output (reference: clusttraj/io.py:582-1015) -------------------------------------
# align_mol's arithmetic (centre, Kabsch rotations, Hungarian reorders, the optional final rotation) runs on the GPU
# through ct_pair_transforms, which returns the rigid transform of every (medoid, member) pair; the host applies it
# to ALL atoms of the member (hydrogens included, original atom order: what align_mol prints) and writes the files.

def _frame_centres(atoms, coords, noh, nsatoms):
    """Centre align_mol subtracts from every atom of a frame (io.py:624-647): centroid of the kept atoms, or of the
    kept solute atoms with -ns."""
    keep = atoms != 1 if noh else np.ones(len(atoms), bool)
    if nsatoms:
        keep = keep & (np.arange(len(atoms)) < nsatoms)
    return coords[:, keep, :].mean(axis=1)


def superposed_frames(trajfile, reference, members, noh, reorder, reorder_solvent_only, nsatoms, weight_solute,
                      reorderexcl, final_kabsch):
    """(atomic numbers [N], reference frame [N,3] centred as io.py:823-858 writes it, members [len(members), N, 3]
    superposed on it as io.py:align_mol returns them)."""
    from . import _engine
    from .distmat import _engine_for, _load_trajectory, _opts_from, visible_devices

    atoms, coords = _load_trajectory(trajfile)
    if atoms.ndim == 2:
        raise NotImplementedError("clusttraj_b200: superposed-structure output needs the same atoms in every frame "
                                  "(the matrix itself accepts per-frame element lists, see distmat.frame_invariant_view)")
    opts = _opts_from(noh, reorder, reorder_solvent_only, nsatoms, weight_solute, reorderexcl, final_kabsch)
    centres = _frame_centres(atoms, coords, bool(noh), nsatoms)
    ref = coords[reference] - centres[reference]
    members = [int(m) for m in members]
    if not members:
        return atoms, ref, np.empty((0,) + ref.shape)
    eng = _engine_for(visible_devices()[0])
    eng.load_frames(coords, atoms, opts)
    _, R, T = eng.pair_transforms([(reference, m) for m in members])
    moved = np.einsum("mnk,mkj->mnj", coords[members] - centres[members][:, None, :], R) + T[:, None, :]
    return atoms, ref, moved


def _xyz_titles(trajfile):
    titles = []
    with open(trajfile) as fh:
        lines = fh.read().split("\n")
    pos = 0
    while pos < len(lines):
        if not lines[pos].strip():
            pos += 1
            continue
        n = int(lines[pos].split()[0])
        titles.append(lines[pos + 1].rstrip() if pos + 1 < len(lines) else "")
        pos += n + 2
    return titles


class _XyzWriter:
    """Minimal stand-in for pybel.Outputfile("xyz", ...): refuses to overwrite unless asked (like OpenBabel's).  The text is
    formatted by the engine library's host threads (ct_format_xyz), a frame loop in Python only if it cannot be loaded."""

    def __init__(self, path, overwrite):
        if os.path.exists(path) and not overwrite:
            raise IOError(f"{path} already exists. Use 'overwrite=True' to overwrite it.")
        self.fh = open(path, "wb")

    def write_many(self, atoms, frames, titles):
        frames = np.asarray(frames, dtype=np.float64)
        if frames.shape[0] == 0:
            return
        try:
            from . import _engine
            self.fh.write(_engine.format_xyz_native(atoms, frames, list(titles)))
            return
        except (ImportError, OSError):
            pass
        from .xyz import SYMBOLS

        for coords, title in zip(frames, titles):
            self.fh.write(f"{len(atoms)}\n{title}\n".encode())
            for z, (x, y, w) in zip(atoms, coords):
                self.fh.write(f"{SYMBOLS[int(z)]:<3s}{x:15.5f}{y:15.5f}{w:15.5f}\n".encode())

    def write(self, atoms, coords, title):
        self.write_many(atoms, np.asarray(coords)[None], [title])

    def close(self):
        self.fh.close()


def _writer(outfmt, path, overwrite):
    if outfmt != "xyz":
        raise NotImplementedError(f"structure output in .{outfmt} needs OpenBabel; clusttraj_b200 writes .xyz natively")
    return _XyzWriter(path, overwrite)


def save_clusters_config(trajfile, clusters, distmat, noh, reorder, reorder_solvent_only, nsatoms, weight_solute,
                         outbasename, outfmt, reorderexcl, final_kabsch, overwrite):
    """Save the superposed configurations of every cluster, medoid first (io.py:758-899): one file
    ``<outbasename>_<cluster>.<outfmt>`` per cluster."""
    from .classify import find_medoids_from_clusters

    clusters = np.asarray(clusters)
    medoids = find_medoids_from_clusters(distmat, clusters)      # io.py:810-816 recomputes them from a squareform
    titles = _xyz_titles(trajfile) if trajfile.lower().endswith(".xyz") else None
    for cnum in range(1, int(clusters.max()) + 1):
        medoid = int(medoids[cnum - 1])
        members = [int(i) for i in np.where(clusters == cnum)[0] if i != medoid]
        atoms, ref, moved = superposed_frames(trajfile, medoid, members, noh, reorder, reorder_solvent_only, nsatoms,
                                              weight_solute, reorderexcl, final_kabsch)
        out = _writer(outfmt, f"{outbasename}_{cnum}.{outfmt}", overwrite)
        out.write(atoms, ref, titles[medoid] if titles else "")
        out.write_many(atoms, moved, [titles[m] if titles else "" for m in members])
        out.close()


def save_medoids_config(trajfile, medoids, noh, reorder, reorder_solvent_only, nsatoms, weight_solute, outbasename,
                        outfmt, reorderexcl, final_kabsch, overwrite):
    """Save the medoids superposed on the first one to a single file (io.py:902-1015)."""
    medoids = [int(m) for m in np.asarray(medoids)]
    titles = _xyz_titles(trajfile) if trajfile.lower().endswith(".xyz") else None
    atoms, ref, moved = superposed_frames(trajfile, medoids[0], medoids[1:], noh, reorder, reorder_solvent_only, nsatoms,
                                          weight_solute, reorderexcl, final_kabsch)
    out = _writer(outfmt, f"{outbasename}.{outfmt}", overwrite)
    out.write(atoms, ref, titles[medoids[0]] if titles else "")
    out.write_many(atoms, moved, [titles[m] if titles else "" for m in medoids[1:]])
    out.close()
