// K1 — mask, centre and pack.  Replaces, once per FRAME instead of once per PAIR, the reference's
// get_mol_info (clusttraj/utils.py:20-31), the -n hydrogen mask (distmat.py:135-141,150-156) and the
// centring on rmsd.centroid (distmat.py:148-149,157-158,164-165,168-169).
//
//   centre_kernel : one CTA per frame.  Gathers the kept atoms, subtracts the centroid of the first
//                   `ncen` kept atoms (all of them, or the solute with -ns), writes the centred frame
//                   as [K][3] float64 (read by the per-pair kernels) and G[m] = 0.5 * sum |x|^2.
//   tile_pack_kernel : re-lays the centred frames into the tile-major layout of common.cuh (one thread
//                   per output element, fully coalesced stores, zero padding in atoms and frames).
#include "common.cuh"

namespace ct {

__global__ void __launch_bounds__(128) centre_kernel(const double* __restrict__ raw, const int* __restrict__ keep,
                                                     int M, int N, int K, int ncen, double* __restrict__ centred,
                                                     double* __restrict__ G) {
    const int m = blockIdx.x;
    const double* src = raw + (size_t)m * N * 3;
    double* dst = centred + (size_t)m * K * 3;
    __shared__ double red[4][4];
    __shared__ double cen[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int k = tid; k < ncen; k += blockDim.x) {
        const double* a = src + (size_t)keep[k] * 3;
        sx += a[0]; sy += a[1]; sz += a[2];
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) { red[warp][0] = sx; red[warp][1] = sy; red[warp][2] = sz; }
    __syncthreads();
    if (tid == 0) {
        const double inv = 1.0 / (double)ncen;
        cen[0] = (red[0][0] + red[1][0] + red[2][0] + red[3][0]) * inv;
        cen[1] = (red[0][1] + red[1][1] + red[2][1] + red[3][1]) * inv;
        cen[2] = (red[0][2] + red[1][2] + red[2][2] + red[3][2]) * inv;
    }
    __syncthreads();
    const double cx = cen[0], cy = cen[1], cz = cen[2];
    double g = 0.0;
    for (int k = tid; k < K; k += blockDim.x) {
        const double* a = src + (size_t)keep[k] * 3;
        const double x = a[0] - cx, y = a[1] - cy, z = a[2] - cz;
        dst[k * 3 + 0] = x; dst[k * 3 + 1] = y; dst[k * 3 + 2] = z;
        g = fma(x, x, fma(y, y, fma(z, z, g)));
    }
    g = warp_sum(g);
    __syncthreads();
    if (lane == 0) red[warp][3] = g;
    __syncthreads();
    if (tid == 0) G[m] = 0.5 * (red[0][3] + red[1][3] + red[2][3] + red[3][3]);  // the Gram epilogue wants G/2
}

__global__ void __launch_bounds__(256) tile_pack_kernel(const double* __restrict__ centred, int M, int K,
                                                        int kq_total, long long total, double* __restrict__ packed) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        const int a = (int)(o & 3);
        const int f = (int)((o >> 2) & (FT - 1));
        long long rest = o >> 7;  // (tile * kq_total + kq) * 3 + c
        const int c = (int)(rest % 3);
        rest /= 3;
        const int kq = (int)(rest % kq_total);
        const long long tile = rest / kq_total;
        const long long m = tile * FT + f;
        const int k = kq * 4 + a;
        double v = 0.0;
        if (m < M && k < K) v = centred[((size_t)m * K + k) * 3 + c];
        packed[o] = v;
    }
}

cudaError_t launch_pack(const double* raw, const int* keep, int M, int N, int K, int ncen, int m_pad, int kq_total,
                        double* centred, double* G, double* packed, int sm_count, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(G, 0, sizeof(double) * (size_t)m_pad, stream);
    if (e != cudaSuccess) return e;
    centre_kernel<<<M, 128, 0, stream>>>(raw, keep, M, N, K, ncen, centred, G);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (packed) {
        const long long total = (long long)(m_pad / FT) * kq_total * 3 * FT * 4;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)sm_count * 16;
        if (blocks > cap) blocks = cap;
        tile_pack_kernel<<<(unsigned)blocks, 256, 0, stream>>>(centred, M, K, kq_total, total, packed);
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace ct
