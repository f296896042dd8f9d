/* clusttraj_b200 — C ABI of the B200 all-pairs minimum-RMSD engine.
 *
 * The reference (hmcezar/clusttraj v1.2.0) has NO plugin / FFI interface: its seam for this path is
 * three Python callables in clusttraj/distmat.py plus the ClustOptions dataclass (io.py:42-83).
 * This header is the boundary a maintainer would bind from those callables (ctypes stub in
 * INTEGRATION.md).  Each entry point names the reference code it replaces.
 *
 * Conventions: plain pointers and sizes, no exceptions across the ABI.  Every function returns 0 on
 * success or a negative ct_status; ct_last_error(engine) gives the message.  The caller owns all host
 * buffers.  One host thread per engine handle at a time; one engine drives one GPU.
 */
#ifndef CLUSTTRAJ_B200_H
#define CLUSTTRAJ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ct_engine ct_engine;

enum ct_status {
    CT_OK = 0,
    CT_ERR_CUDA = -1,       /* a CUDA runtime call failed (message in ct_last_error) */
    CT_ERR_ARG = -2,        /* bad argument */
    CT_ERR_STATE = -3,      /* call order (e.g. rows requested before frames were loaded) */
    CT_ERR_UNSUPPORTED = -4 /* option combination the engine does not implement */
};

/* Per-pair options: the hot-path fields of ClustOptions (io.py:42-83) as read by
 * build_distance_matrix (distmat.py:54-70) and compute_distmat_line (distmat.py:81-94).
 * Truthiness follows the reference (distmat.py:134,172,231,259): 0 / 0.0 mean "off". */
typedef struct ct_opts {
    int32_t no_hydrogen;          /* -n   : drop atoms with atomic number 1 (distmat.py:135-141,150-156) */
    int32_t solute_natoms;        /* -ns  : solute size before the H mask; 0 = off (distmat.py:116-118,134) */
    int32_t reorder_mode;         /* -e   : 0 = none, 1 = hungarian (io.py:559-561 -> rmsd.reorder_hungarian),
                                     2 = distance (io.py:562-563 -> rmsd.reorder_distance) */
    int32_t reorder_solvent_only; /* -rs  : skip the solute reorder block (distmat.py:184) */
    int32_t final_kabsch;         /* --final-kabsch (distmat.py:265,271-275) */
    int32_t n_excl;               /* length of reorder_excl */
    double weight_solute;         /* -ws ; 0.0 = off (distmat.py:259-262) */
    const int32_t* reorder_excl;  /* -eex, 0-based kept-atom positions as in ClustOptions.reorder_excl (io.py:513-517) */
} ct_opts;

/* Timing / counters of the last ct_rmsd_rows* call (device times from CUDA events). */
typedef struct ct_stats {
    double pack_ms;      /* K1 centre/pack kernels of the last ct_load_frames */
    double h2d_ms;       /* host->device copy of the coordinates in the last ct_load_frames */
    double kernel_ms;    /* sum of the hot-path kernel launches (Gram+epilogue, fix-up, per-pair) */
    double gram_ms;      /* the Gram + fused epilogue kernel only (either path) */
    double pair_ms;      /* per-pair kernel (stepwise / Hungarian modes, fix-up of flagged pairs) */
    double d2h_ms;       /* device->host streaming of the condensed slab (overlapped with compute) */
    double total_ms;     /* first launch to last byte on the host, device clock */
    int64_t pairs;       /* matrix entries produced */
    int64_t flagged;     /* pairs re-done by the explicit-rotation path (near-duplicates, degenerate 3x3) */
    int64_t launches;    /* kernel launches issued by the engine in this call */
    int64_t kept_atoms;  /* K: atoms entering the contraction after the mask */
    int64_t lsap_steps;  /* shortest-augmenting-path frontier scans + rows assigned directly to their nearest column (Hungarian modes) */
    int64_t gram_path;   /* kernel behind the pure-Kabsch modes: 0 = none (per-pair kernel), 1 = FP64 DMMA Gram kernel,
                            2 = int8 tcgen05 Gram kernel (exact digit split of 31-bit fixed-point coordinates) */
    double quant_step;   /* path 2: the fixed-point quantum 2^-f in Angstrom (RMSD moves by at most ~1.8 quanta) */
} ct_stats;

/* Lifetime.  device = CUDA ordinal.  Fails (never falls back to the CPU) when no device exists. */
int ct_engine_create(int device, ct_engine** out);
void ct_engine_destroy(ct_engine* e);
const char* ct_last_error(const ct_engine* e); /* valid until the next call on e; e may be NULL */
int ct_device_count(void);

/* Pinned host memory for inputs/outputs (so streaming runs at PCIe rate).  Optional: pageable
 * pointers are accepted everywhere, only slower. */
int ct_host_alloc(void** ptr, uint64_t bytes);
int ct_host_free(void* ptr);

/* Replaces the per-pair get_mol_info + mask + centring of distmat.py:131-169 (done ONCE per frame
 * here): copies coords [M][N][3] float64 (row-major) and the frame-invariant atomic numbers [N] to the
 * GPU, applies the -n mask, centres every frame on the centroid the mode prescribes (all kept atoms, or
 * the first `natoms` kept atoms with -ns) and writes the packed layouts the kernels read. */
int ct_load_frames(ct_engine* e, const double* coords, const int32_t* atomic_nums, int64_t M, int32_t N,
                   const ct_opts* opts);

/* Per-frame element lists.  The reference reads every frame's own atoms (distmat.py:123-141) and builds the element blocks of
 * a reorder from them per pair (rmsd.reorder_hungarian / reorder_distance over Qa[view], Pa[view]: distmat.py:192,243), so
 * frames may list their atoms in different orders.  elements = [M][K] atomic numbers of the KEPT atoms of every frame of the
 * last ct_load_frames (whose own atomic_nums then only describe frame 0); load the frames with no_hydrogen = 0 after applying
 * each frame's own -n mask on the host (clusttraj_b200.distmat.frame_invariant_view).  Every frame must have the element
 * counts of frame 0 among the atoms a phase reorders (the reference fails on a pair whose counts differ): CT_ERR_ARG
 * otherwise.  Without a reorder the elements are not used and the call is a no-op. */
int ct_set_frame_elements(ct_engine* e, const int32_t* elements, int64_t M, int32_t K);

/* Re-run the centre/pack kernels (K1) on the coordinates already resident on the GPU (kernel-only
 * timing of the whole hot path without the host->device copy). */
int ct_pack_resident(ct_engine* e, ct_stats* stats);

/* Number of matrix entries in rows [row_begin,row_end): sum over i of (M-1-i). */
int64_t ct_rows_pairs(int64_t M, int64_t row_begin, int64_t row_end);

/* Replaces compute_distmat_line (distmat.py:81-277) for rows row_begin..row_end-1 and the row
 * flattening of build_distance_matrix (distmat.py:78): writes the contiguous slab of the SciPy-condensed
 * vector that those rows own, out[k] for k = idx(row_begin,row_begin+1) .. , to HOST memory `out`
 * (ct_rows_pairs entries).  Compute and the device->host stream are overlapped chunk by chunk. */
int ct_rmsd_rows(ct_engine* e, int64_t row_begin, int64_t row_end, double* out, ct_stats* stats);

/* Same, but leaves the slab in an engine-owned DEVICE buffer (kernel-only timing, or a consumer that
 * stays on the GPU).  *out_dev is valid until the next call on e. */
int ct_rmsd_rows_device(ct_engine* e, int64_t row_begin, int64_t row_end, const double** out_dev,
                        ct_stats* stats);
/* Copy entries [offset, offset+count) of the last device slab to host memory. */
int ct_copy_slab(ct_engine* e, int64_t offset, int64_t count, double* out);

/* Multi-GPU hand-off of the matrix to the consumers that stay on the GPU ("device-resident full
 * matrix for GPU linkage"; consumers at clusttraj/classify.py:100,146-177).  The reference's Pool over rows
 * (distmat.py:73-78) becomes one engine + one host thread per GPU, each owning a contiguous slab of rows; instead of
 * sending every slab to the host and uploading the whole vector again for the clustering, the slabs also land, chunk by
 * chunk while the next chunk is computed, in ONE GPU's memory:
 *   ct_matrix_reserve(root, M)        room for the whole condensed matrix of M frames on root's GPU;
 *   ct_rmsd_rows_gather(e, r0, r1, out, root, stats)
 *                                     ct_rmsd_rows(e, r0, r1, out, stats) that additionally copies the rows into root's
 *                                     matrix at their own place in the condensed vector: in place when e == root, else
 *                                     GPU -> GPU with cudaMemcpyPeerAsync (NVLink when peer access is available).
 *                                     `out` (host) may be NULL when only the device-resident matrix is wanted.  Engines
 *                                     with disjoint row ranges may gather into the same root from different host threads;
 *   ct_matrix_commit(root, M)         after every gathering call has returned: the matrix is complete, and
 *                                     ct_linkage / ct_matrix_consumers(root, NULL, M, ...) run on it without an upload. */
int ct_matrix_reserve(ct_engine* root, int64_t M);
int ct_rmsd_rows_gather(ct_engine* e, int64_t row_begin, int64_t row_end, double* out, ct_engine* root, ct_stats* stats);
int ct_matrix_commit(ct_engine* root, int64_t M);

/* Whole matrix in one call = build_distance_matrix (distmat.py:44-78): load + all rows. */
int ct_rmsd_condensed(ct_engine* e, const double* coords, const int32_t* atomic_nums, int64_t M, int32_t N,
                      const ct_opts* opts, double* out, ct_stats* stats);

/* Debug / parity aid (not in the reference, which never returns the permutation): value and composite
 * atom permutation (perm[k] = kept-atom index of frame j that ends at position k) for explicit pairs,
 * through the per-pair kernel whatever the mode.  pairs = [n_pairs][2] (i, j), i < j. */
int ct_pair_values(ct_engine* e, const int64_t* pairs, int64_t n_pairs, double* values, int32_t* perms /* [n_pairs][K] or NULL */,
                   ct_stats* stats);

/* Superposition transforms for structure output: what clusttraj/io.py:align_mol (582-756) applies to
 * ALL atoms of a frame when save_clusters_config / save_medoids_config (io.py:758-1015) write the members of a cluster
 * superposed on its medoid.  pairs = [n_pairs][2] (reference frame, moving frame), in either order of index; for each:
 * values[t] = the pair's RMSD as the matrix holds it, transforms[t] = 12 doubles R (3x3 row-major) and T such that
 *   aligned = (raw_moving - centre_moving) R + T        (row vectors; centre = the centroid ct_load_frames subtracts:
 *                                                        all kept atoms, or the kept solute atoms with -ns)
 * R is the product of the rotations of the per-pair pipeline (solute / whole-system Kabsch, second solute Kabsch after the
 * solute reorder) and of the final rotation where align_mol applies one (-ns -e --final-kabsch: io.py:743-746;
 * -ws --final-kabsch: the weighted Kabsch with its translation, io.py:733-741).  Runs the per-pair kernel whatever the
 * mode (the frames must have been loaded with the same options as the matrix). */
int ct_pair_transforms(ct_engine* e, const int64_t* pairs, int64_t n_pairs, double* values, double* transforms /* [n_pairs][12] */,
                       ct_stats* stats);

/* Consumers of the condensed matrix at scale, replacing clusttraj/classify.py:146-177 without
 * the M x M squareform the reference builds on the host:
 *   *total      = sum_distmat(distmat)                                   (classify.py:168-177)
 *   medoids[c]  = find_medoids_from_clusters(distmat, clusters)[c]       (classify.py:146-165): the member of cluster
 *                 c+1 with the smallest sum of distances to the other members, lowest frame index on ties.
 * distmat: HOST pointer to the M(M-1)/2 condensed vector (copied to the GPU), or NULL to use the whole matrix that
 * ct_rmsd_rows_device(e, 0, M, ...) left there.  labels: [M] cluster ids 1..n_clusters (the fcluster / cut_tree+1
 * convention of classify.py:110,136); n_clusters = 0 computes the sum alone.  A cluster without members is an error
 * (the reference raises). */
int ct_matrix_consumers(ct_engine* e, const double* distmat, int64_t M, const int32_t* labels, int32_t n_clusters,
                        int64_t* medoids, double* total, ct_stats* stats);

/* Hierarchical clustering of the condensed matrix on the GPU: scipy.cluster.hierarchy.linkage(distmat,
 * method) as clusttraj calls it at clusttraj/classify.py:30,99,127.  method uses SciPy's codes: 0 single, 1 complete,
 * 2 average, 5 ward, 6 weighted (centroid and median are not offered: CT_ERR_UNSUPPORTED, use SciPy).  Z: [M-1][4]
 * float64, the linkage matrix exactly as SciPy returns it (same merges, same order, same tie rules, bit-identical
 * heights).  distmat: HOST pointer to the condensed vector, or NULL to cluster the whole matrix that
 * ct_rmsd_rows_device(e, 0, M, ...) left on the GPU; the matrix on the GPU is left unchanged (the chain methods work on
 * a copy), so ct_matrix_consumers(e, NULL, ...) can follow.  stats: kernel_ms (clustering), h2d_ms (upload),
 * launches (step kernels). */
int ct_linkage(ct_engine* e, const double* distmat, int64_t M, int32_t method, double* Z, ct_stats* stats);

/* One-shot trajectory ingest, replacing the per-pair OpenBabel parse of clusttraj/distmat.py:56-61,
 * 123-131 + utils.get_mol_info (utils.py:20-31): a regular multi-frame XYZ file -> atomic_numbers[M][N] and
 * coords[M][N][3], parsed once by n_threads host threads (0 = all cores).  Call with coords = NULL to learn
 * *n_frames and *n_atoms, allocate, and call again with capacity_frames >= *n_frames.  Host-only: needs no GPU.
 * Irregular files (frames of different sizes, blank separator lines, short atom lines, unknown element symbols) are
 * refused with CT_ERR_ARG and a message in ct_last_error(NULL); clusttraj_b200.xyz then uses its general parser. */
int ct_read_xyz(const char* path, int64_t* n_frames, int32_t* n_atoms, double* coords, int32_t* atomic_numbers,
                int64_t capacity_frames, int32_t n_threads);
/* The same for multi-MODEL PDB files (ATOM / HETATM records, frames closed by END* records; element from columns 77-78,
 * else guessed from the atom name) and GROMACS .gro trajectories (positions converted from nm to Angstrom, element
 * guessed from the atom name): the rules of clusttraj_b200/xyz.py:read_pdb / read_gro, which stand in for OpenBabel's. */
int ct_read_pdb(const char* path, int64_t* n_frames, int32_t* n_atoms, double* coords, int32_t* atomic_numbers,
                int64_t capacity_frames, int32_t n_threads);
int ct_read_gro(const char* path, int64_t* n_frames, int32_t* n_atoms, double* coords, int32_t* atomic_numbers,
                int64_t capacity_frames, int32_t n_threads);

/* Text of n_frames consecutive XYZ frames of the same atoms, in the layout the reference writes through
 * pybel.Outputfile("xyz") in save_clusters_config / save_medoids_config (io.py:870-899,1000-1015): "<N>\n<title>\n" and
 * "%-3s%15.5f%15.5f%15.5f\n" per atom.  atomic_numbers[N]; coords[n_frames][N][3]; titles: n_frames NUL-terminated
 * strings back to back, or NULL.  The buffer is malloc'ed: *out, *length bytes, release with ct_buffer_free.  Host-only. */
int ct_format_xyz(const int32_t* atomic_numbers, const double* coords, int64_t n_frames, int32_t n_atoms, const char* titles,
                  char** out, int64_t* length, int32_t n_threads);
void ct_buffer_free(void* p);

/* Free / total memory of the engine's GPU (how the host side decides between a device-resident matrix and streaming). */
int ct_device_memory(ct_engine* e, int64_t* free_bytes, int64_t* total_bytes);

int ct_get_stats(const ct_engine* e, ct_stats* stats);
const char* ct_version(void);

#ifdef __cplusplus
}
#endif
#endif
