// Explicit-rotation path, one warp per pair, on the centred [K][3] frames.
//
// fixup_kernel re-does the pairs the Gram epilogue flagged (near-duplicates, where
// G_i + G_j - 2 lambda cancels, and degenerate 3x3 problems) the way the reference itself evaluates
// every pair: rotate explicitly and sum |p U - q|^2 (rmsd.kabsch_rmsd as called at
// clusttraj/distmat.py:275), which has no cancellation.
#include "common.cuh"
#include "solve3x3.cuh"

namespace ct {

// covariance S[a][b] = sum_k p_k[a] q_k[b] over atoms [k0, k1), reduced over the warp
__device__ __forceinline__ void warp_covariance(const double* __restrict__ P, const double* __restrict__ Q, int k0,
                                                int k1, int lane, double* s) {
#pragma unroll
    for (int t = 0; t < 9; ++t) s[t] = 0.0;
    for (int k = k0 + lane; k < k1; k += 32) {
        const double px = P[3 * k], py = P[3 * k + 1], pz = P[3 * k + 2];
        const double qx = Q[3 * k], qy = Q[3 * k + 1], qz = Q[3 * k + 2];
        s[0] = fma(px, qx, s[0]); s[1] = fma(px, qy, s[1]); s[2] = fma(px, qz, s[2]);
        s[3] = fma(py, qx, s[3]); s[4] = fma(py, qy, s[4]); s[5] = fma(py, qz, s[5]);
        s[6] = fma(pz, qx, s[6]); s[7] = fma(pz, qy, s[7]); s[8] = fma(pz, qz, s[8]);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s[t] = warp_sum(s[t]);
}

__device__ __forceinline__ double warp_direct_rmsd(const double* __restrict__ P, const double* __restrict__ Q, int K,
                                                   int lane) {
    double s[9], U[9];
    warp_covariance(P, Q, 0, K, lane, s);
    kabsch_rotation(s, U);
    double res = 0.0;
    for (int k = lane; k < K; k += 32) {
        const double px = P[3 * k], py = P[3 * k + 1], pz = P[3 * k + 2];
        const double dx = fma(px, U[0], fma(py, U[3], pz * U[6])) - Q[3 * k];
        const double dy = fma(px, U[1], fma(py, U[4], pz * U[7])) - Q[3 * k + 1];
        const double dz = fma(px, U[2], fma(py, U[5], pz * U[8])) - Q[3 * k + 2];
        res = fma(dx, dx, fma(dy, dy, fma(dz, dz, res)));
    }
    res = warp_sum(res);
    return sqrt(res / (double)K);
}

__device__ __forceinline__ void condensed_decode(long long M, long long idx, int& i, int& j) {
    // largest i with row_base(i) <= idx
    const double b = 2.0 * (double)M - 1.0;
    long long r = (long long)floor((b - sqrt(b * b - 8.0 * (double)idx)) * 0.5);
    if (r < 0) r = 0;
    while (r > 0 && condensed_row_base(M, r) > idx) --r;
    while (condensed_row_base(M, r + 1) <= idx) ++r;
    i = (int)r;
    j = (int)(idx - condensed_row_base(M, r) + r + 1);
}

__global__ void __launch_bounds__(128) fixup_kernel(const double* __restrict__ centred, int M, int K, FixList fix,
                                                    double* __restrict__ out, long long idx_base, long long n_entries,
                                                    unsigned long long* __restrict__ flagged_total) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const unsigned int cnt = *fix.count;
    if (cnt == 0) return;
    if (gw == 0 && lane == 0) atomicAdd(flagged_total, (unsigned long long)cnt);
    const unsigned int n = cnt < fix.capacity ? cnt : fix.capacity;
    for (long long t = gw; t < n; t += nw) {
        const int2 ij = fix.pairs[t];
        const double v = warp_direct_rmsd(centred + (size_t)ij.y * K * 3, centred + (size_t)ij.x * K * 3, K, lane);
        if (lane == 0) out[condensed_row_base(M, ij.x) + (ij.y - ij.x - 1) - idx_base] = v;
    }
    if (cnt > fix.capacity) {  // list overflowed: the remaining flagged entries carry the marker -1
        for (long long e0 = gw * 32; e0 < n_entries; e0 += nw * 32) {
            const long long e = e0 + lane;
            const bool hit = e < n_entries && out[e] < 0.0;
            unsigned int mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                int i, j;
                condensed_decode(M, idx_base + e0 + src, i, j);
                const double v = warp_direct_rmsd(centred + (size_t)j * K * 3, centred + (size_t)i * K * 3, K, lane);
                if (lane == 0) out[e0 + src] = v;
            }
        }
    }
}

cudaError_t launch_fixup(const double* centred, int M, int K, const FixList& fix, double* out, long long idx_base,
                         long long n_entries, unsigned long long* flagged_total, int sm_count, cudaStream_t stream) {
    fixup_kernel<<<sm_count * 2, 128, 0, stream>>>(centred, M, K, fix, out, idx_base, n_entries, flagged_total);
    return cudaGetLastError();
}

}  // namespace ct
