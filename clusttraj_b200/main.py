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
    t2 = time.monotonic()
    n = distmat.size
    Logger.logger.info(f"Time spent computing (or loading) RMSD matrix: {t2 - t1:.6f} s "
                       f"({n / max(t2 - t1, 1e-12):.3e} pairs/s)\n")
    trust_resident(distmat)   # nothing below writes to the matrix: one upload serves linkage, sum and medoids
    Z, clusters = classify(clust_opt, distmat)
    Logger.logger.info(f"A total {len(clusters)} snapshots were read and {clusters.max()} cluster(s) was(were) found.\n")
    # superposed structures per cluster / medoids (main.py:67-92,116-137): transforms from the GPU, .xyz written natively
    if clust_opt.save_confs:
        save_clusters_config(clust_opt.trajfile, clusters, distmat, clust_opt.no_hydrogen, clust_opt.reorder_alg,
                             clust_opt.reorder_solvent_only, clust_opt.solute_natoms, clust_opt.weight_solute,
                             clust_opt.out_conf_name, clust_opt.out_conf_fmt, clust_opt.reorder_excl, clust_opt.final_kabsch,
                             clust_opt.overwrite)
    # the summary the reference writes (main.py:107-135): matrix sum, cluster sizes and medoids — both reductions
    # run on the GPU (clusttraj_b200/classify.py) instead of over a host-side squareform
    total = sum_distmat(distmat)
    medoids = find_medoids_from_clusters(distmat, clusters)
    with open(clust_opt.summary_name, "w") as fh:
        fh.write(str(clust_opt))
        fh.write(f"A total {len(clusters)} snapshots were read and {clusters.max()} cluster(s) was(were) found.\n")
        fh.write(f"\nThe total sum of the RMSD matrix is {total}\n")
        fh.write("\nCluster\tSize\tMedoid\n")
        for c in range(1, int(clusters.max()) + 1):
            fh.write(f"{c}\t{int((clusters == c).sum())}\t{int(medoids[c - 1])}\n")
        if clust_opt.metrics:   # main.py:140-153
            ss, ch, db, cpcc = compute_metrics(distmat, Z, clusters)
            fh.write(f"\nSilhouette score: {ss:.3f}\nCalinski Harabsz score: {ch:.3f}\n"
                     f"Davies-Bouldin score: {db:.3f}\nCophenetic correlation coefficient: {cpcc:.3f}\n\n")
    if clust_opt.save_medoids:
        save_medoids_config(clust_opt.trajfile, medoids, clust_opt.no_hydrogen, clust_opt.reorder_alg,
                            clust_opt.reorder_solvent_only, clust_opt.solute_natoms, clust_opt.weight_solute,
                            clust_opt.out_medoids_name, clust_opt.out_medoids_fmt, clust_opt.reorder_excl,
                            clust_opt.final_kabsch, clust_opt.overwrite)
    matrix_saved()
    Logger.logger.info(f"Total wall time: {time.monotonic() - t0:.6f} s\n")
    return 0
