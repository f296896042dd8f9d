"""CLI entry (reference: clusttraj/main.py:24-174): options -> RMSD matrix on the GPU -> the unchanged
linkage on the GPU (SciPy's algorithm, identical Z) / fcluster -> clusters file and summary."""
import sys
import time

import numpy as np

from .classify import (classify_structures, classify_structures_nclusters, classify_structures_silhouette,
                       compute_metrics, find_medoids_from_clusters, sum_distmat, trust_resident)
from .distmat import get_distmat_saving_in_background
from .io import Logger, configure_runtime, save_clusters_config, save_medoids_config


def classify(clust_opt, distmat):
    """The clustering step as the reference's main dispatches it (main.py:47-57): -nc -> classify_structures_nclusters,
    else classify_structures; the linkage runs on the GPU (clusttraj_b200/classify.py)."""
    if clust_opt.silhouette_score:
        return classify_structures_silhouette(clust_opt, distmat)
    if clust_opt.n_clusters is not None:
        return classify_structures_nclusters(clust_opt, distmat)
    return classify_structures(clust_opt, distmat)


def main(args=None):
    t0 = time.monotonic()
    clust_opt = configure_runtime(sys.argv[1:] if args is None else args)
    t1 = time.monotonic()
    distmat, matrix_saved = get_distmat_saving_in_background(clust_opt)   # the .npy is written while the GPU clusters
    try:
        _run(clust_opt, distmat, t1)
    finally:
        matrix_saved()   # joins the writer and re-raises what it raised, whatever happened above
    Logger.logger.info(f"Total wall time: {time.monotonic() - t0:.6f} s\n")
    return 0


def _run(clust_opt, distmat, t_start):
    t2 = time.monotonic()
    if clust_opt.verbose:
        Logger.logger.info(f"Time spent computing (or loading) RMSD matrix: {t2 - t_start:.6f} s "
                           f"({distmat.size / max(t2 - t_start, 1e-12):.3e} pairs/s)\n")
    trust_resident(distmat)   # nothing below writes to the matrix: one upload serves linkage, sum and medoids
    try:
        Z, clusters = classify(clust_opt, distmat)
    except ValueError as exc:   # main.py:53-66: e.g. -nc larger than the number of snapshots
        Logger.logger.error(f"{exc}\n")
        raise SystemExit(str(exc)) from exc
    t3 = time.monotonic()
    if clust_opt.verbose:
        Logger.logger.info(f"Time spent clustering: {t3 - t2:.6f} s\n")
    # superposed structures per cluster / medoids (main.py:67-92,116-137): transforms from the GPU, .xyz written natively
    if clust_opt.save_confs:
        Logger.logger.info(f"Writing superposed configurations per cluster to files {clust_opt.out_conf_name}_*."
                           f"{clust_opt.out_conf_fmt}\n")
        save_clusters_config(clust_opt.trajfile, clusters, distmat, clust_opt.no_hydrogen, clust_opt.reorder_alg,
                             clust_opt.reorder_solvent_only, clust_opt.solute_natoms, clust_opt.weight_solute,
                             clust_opt.out_conf_name, clust_opt.out_conf_fmt, clust_opt.reorder_excl, clust_opt.final_kabsch,
                             clust_opt.overwrite)
        if clust_opt.verbose:
            Logger.logger.info(f"Time spent saving configurations: {time.monotonic() - t3:.6f} s\n")
    # the summary the reference writes and logs (main.py:107-174), same text: matrix sum, medoids, cluster sizes — both
    # reductions run on the GPU (clusttraj_b200/classify.py) instead of over a host-side squareform
    out = f"A total {len(clusters)} snapshots were read and {max(clusters)} cluster(s) was(were) found.\n"
    out += f"\nThe total sum of the RMSD matrix is {sum_distmat(distmat)}\n"
    medoids = find_medoids_from_clusters(distmat, clusters)
    if clust_opt.save_medoids:
        t4 = time.monotonic()
        Logger.logger.info(f"Writing superposed medoids to file {clust_opt.out_medoids_name}.{clust_opt.out_medoids_fmt}\n")
        save_medoids_config(clust_opt.trajfile, medoids, clust_opt.no_hydrogen, clust_opt.reorder_alg,
                            clust_opt.reorder_solvent_only, clust_opt.solute_natoms, clust_opt.weight_solute,
                            clust_opt.out_medoids_name, clust_opt.out_medoids_fmt, clust_opt.reorder_excl,
                            clust_opt.final_kabsch, clust_opt.overwrite)
        if clust_opt.verbose:
            Logger.logger.info(f"Time spent saving medoids: {time.monotonic() - t4:.6f} s\n")
    out += "The index for the medoids are:\n"
    for medoid in medoids:
        out += f"{medoid} "
    out += "\n\nThe cluster sizes are:\nCluster\tSize\n"
    labels, sizes = np.unique(clusters, return_counts=True)
    for label, size in zip(labels, sizes):
        out += f"{label}\t{size}\n"
    Logger.logger.info(out)
    if clust_opt.metrics:   # main.py:140-160
        t5 = time.monotonic()
        ss, ch, db, cpcc = compute_metrics(distmat, Z, clusters)
        out += (f"\nSilhouette score: {ss:.3f}\nCalinski Harabsz score: {ch:.3f}\n"
                f"Davies-Bouldin score: {db:.3f}\nCophenetic correlation coefficient: {cpcc:.3f}\n\n")
        if clust_opt.verbose:
            Logger.logger.info(f"Time spent computing metrics: {time.monotonic() - t5:.6f} s\n")
    with open(clust_opt.summary_name, "w") as fh:
        fh.write(str(clust_opt))
        fh.write(out)
