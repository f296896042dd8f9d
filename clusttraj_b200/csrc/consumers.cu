// Consumers of the condensed matrix at scale: the two reductions clusttraj runs right
// after the clustering —
//   sum_distmat(distmat)                         clusttraj/classify.py:168-177  (np.sum of the condensed vector)
//   find_medoids_from_clusters(distmat, clusters) clusttraj/classify.py:146-165  (squareform, then per cluster the
//                                                 member with the smallest sum of distances to the other members)
// — without the M x M squareform the reference builds on the host (3.2 GB at 20k frames, 80 GB at 100k).
// Both are single passes over the 8 L bytes of the condensed vector: HBM bound.
//
//   member_sum_kernel : one warp per frame i.  S[i] = sum of d(i, j) over the frames j of i's cluster: the part
//                       j > i is the contiguous condensed row i, the part j < i walks column i (one element per earlier
//                       row).  Fixed summation order (lane-strided, then a shuffle tree): bit-reproducible.
//   medoid kernels    : per cluster, the smallest S and, among the frames that attain it, the lowest frame index
//                       (np.argmin returns the first minimum, classify.py:163).
//   total_sum kernels : two-stage tree sum of the whole vector.
#include "common.cuh"

namespace ct {

__global__ void __launch_bounds__(256) member_sum_kernel(const double* __restrict__ d, const int* __restrict__ labels, int M,
                                                         double* __restrict__ S) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long i = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < M; i += (long long)gridDim.x * warps_per_block) {
        const int li = labels[i];
        double acc = 0.0;
        // column part: d(j, i) for j < i lives at row_base(j) + (i - j - 1)
        for (long long j = lane; j < i; j += 32)
            if (labels[j] == li) acc += d[condensed_row_base(M, j) + (i - j - 1)];
        // row part: d(i, j) for j > i, contiguous
        const double* row = d + condensed_row_base(M, i) - i - 1;
        for (long long j = i + 1 + lane; j < M; j += 32)
            if (labels[j] == li) acc += row[j];
        acc = warp_sum(acc);
        if (lane == 0) S[i] = acc;
    }
}

__global__ void cluster_min_kernel(const double* __restrict__ S, const int* __restrict__ labels, int M,
                                   unsigned long long* __restrict__ best_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    // S >= 0: the bit pattern of a non-negative double orders like the value
    atomicMin(&best_bits[labels[i] - 1], (unsigned long long)__double_as_longlong(S[i]));
}

__global__ void cluster_argmin_kernel(const double* __restrict__ S, const int* __restrict__ labels, int M,
                                      const unsigned long long* __restrict__ best_bits, long long* __restrict__ medoid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int c = labels[i] - 1;
    if ((unsigned long long)__double_as_longlong(S[i]) == best_bits[c]) atomicMin((unsigned long long*)&medoid[c], (unsigned long long)i);
}

__global__ void __launch_bounds__(256) partial_sum_kernel(const double* __restrict__ d, long long n, double* __restrict__ part) {
    __shared__ double red[8];
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) acc += d[k];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        part[blockIdx.x] = t;
    }
}

__global__ void final_sum_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    double acc = 0.0;
    for (int k = threadIdx.x; k < n; k += 32) acc += part[k];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) *out = acc;
}

constexpr int SUM_BLOCKS = 1184;  // 8 x 148

// d: condensed matrix on the device (L = M (M - 1) / 2 entries); labels: 1-based cluster ids on the device;
// scratch: M + SUM_BLOCKS + 1 doubles; best_bits / medoid: n_clusters entries each.
cudaError_t launch_consumers(const double* d, long long L, int M, const int* labels, int n_clusters, double* scratch,
                             unsigned long long* best_bits, long long* medoid, int sm_count, cudaStream_t stream) {
    double* S = scratch;
    double* part = scratch + M;
    double* total = part + SUM_BLOCKS;
    partial_sum_kernel<<<SUM_BLOCKS, 256, 0, stream>>>(d, L, part);
    final_sum_kernel<<<1, 32, 0, stream>>>(part, SUM_BLOCKS, total);
    if (n_clusters > 0) {
        cudaError_t e = cudaMemsetAsync(best_bits, 0xff, sizeof(unsigned long long) * (size_t)n_clusters, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(medoid, 0xff, sizeof(long long) * (size_t)n_clusters, stream);
        if (e != cudaSuccess) return e;
        int blocks = (M + 7) / 8;
        if (blocks > sm_count * 8) blocks = sm_count * 8;
        member_sum_kernel<<<blocks, 256, 0, stream>>>(d, labels, M, S);
        cluster_min_kernel<<<(M + 255) / 256, 256, 0, stream>>>(S, labels, M, best_bits);
        cluster_argmin_kernel<<<(M + 255) / 256, 256, 0, stream>>>(S, labels, M, best_bits, medoid);
    }
    return cudaGetLastError();
}

int consumers_scratch_doubles(int M) { return M + SUM_BLOCKS + 1; }
int consumers_sum_offset(int M) { return M + SUM_BLOCKS; }

}  // namespace ct
