// Probe: how fast can the Gram inner loop (9 LDS.64 + 18 DMMA per k4 step, fragments from shared memory)
// run without TMA, barriers or epilogue?  Variants change the warp tile and the DMMA order.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MA x MB MMA-row/col blocks per warp, KQ k4 steps per "stage" held in smem, ORDER: 0 = (ma,mb,ca,cb), 1 = cb fastest inside ca with A reuse
template <int MA, int MB, int ORDER, int DB>
__global__ void __launch_bounds__(256) loop_kernel(double* out, int iters) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4 * 3 * 128 * 3; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[MA][MB][3][3][2];
#pragma unroll
    for (int a = 0; a < MA; a++)
#pragma unroll
        for (int b = 0; b < MB; b++)
#pragma unroll
            for (int x = 0; x < 3; x++)
#pragma unroll
                for (int y = 0; y < 3; y++) { acc[a][b][x][y][0] = 0; acc[a][b][x][y][1] = 0; }
    const double* sa = sm + (warp & 3) * 32 + lane;
    const double* sb = sm + 4 * 3 * 128 + (warp >> 2) * 64 + lane;
    if (DB == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int kq = 0; kq < 4; kq++) {
                double fa[MA][3], fb[MB][3];
#pragma unroll
                for (int a = 0; a < MA; a++)
#pragma unroll
                    for (int c = 0; c < 3; c++) fa[a][c] = sa[(kq * 3 + c) * 128 + a * 32 * 0 + a];
#pragma unroll
                for (int b = 0; b < MB; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++) fb[b][c] = sb[(kq * 3 + c) * 128 + b * 32];
                if (ORDER == 0) {
#pragma unroll
                    for (int a = 0; a < MA; a++)
#pragma unroll
                        for (int b = 0; b < MB; b++)
#pragma unroll
                            for (int x = 0; x < 3; x++)
#pragma unroll
                                for (int y = 0; y < 3; y++) dmma(acc[a][b][x][y][0], acc[a][b][x][y][1], fa[a][x], fb[b][y]);
                } else {
#pragma unroll
                    for (int x = 0; x < 3; x++)
#pragma unroll
                        for (int a = 0; a < MA; a++)
#pragma unroll
                            for (int b = 0; b < MB; b++)
#pragma unroll
                                for (int y = 0; y < 3; y++) dmma(acc[a][b][x][y][0], acc[a][b][x][y][1], fa[a][x], fb[b][y]);
                }
            }
        }
    } else {
        // register double buffering of the fragments: loads for step k+1 are issued before the DMMAs of step k
        double fa[2][MA][3], fb[2][MB][3];
#pragma unroll
        for (int a = 0; a < MA; a++)
#pragma unroll
            for (int c = 0; c < 3; c++) fa[0][a][c] = sa[c * 128 + a];
#pragma unroll
        for (int b = 0; b < MB; b++)
#pragma unroll
            for (int c = 0; c < 3; c++) fb[0][b][c] = sb[c * 128 + b * 32];
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int kq = 0; kq < 4; kq++) {
                const int cur = kq & 1, nxt = cur ^ 1, kn = (kq + 1) & 3;
#pragma unroll
                for (int a = 0; a < MA; a++)
#pragma unroll
                    for (int c = 0; c < 3; c++) fa[nxt][a][c] = sa[(kn * 3 + c) * 128 + a];
#pragma unroll
                for (int b = 0; b < MB; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++) fb[nxt][b][c] = sb[(kn * 3 + c) * 128 + b * 32];
#pragma unroll
                for (int a = 0; a < MA; a++)
#pragma unroll
                    for (int b = 0; b < MB; b++)
#pragma unroll
                        for (int x = 0; x < 3; x++)
#pragma unroll
                            for (int y = 0; y < 3; y++) dmma(acc[a][b][x][y][0], acc[a][b][x][y][1], fa[cur][a][x], fb[cur][b][y]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int a = 0; a < MA; a++)
#pragma unroll
        for (int b = 0; b < MB; b++)
#pragma unroll
            for (int x = 0; x < 3; x++)
#pragma unroll
                for (int y = 0; y < 3; y++) s += acc[a][b][x][y][0] + acc[a][b][x][y][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MA, int MB, int ORDER, int DB>
void run(const char* name, int sms, double* out, int ctas_per_sm) {
    const int iters = 4000;
    const int smem = 4 * 3 * 128 * 3 * 8;
    CK(cudaFuncSetAttribute(loop_kernel<MA, MB, ORDER, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        CK(cudaEventRecord(e0));
        loop_kernel<MA, MB, ORDER, DB><<<sms * ctas_per_sm, 256, smem>>>(out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double flop = 2.0 * 256 * 9.0 * MA * MB * 4 * iters * 8.0 * sms * ctas_per_sm;
    printf("%-34s ctas/sm %d: %.2f TFLOP/s\n", name, ctas_per_sm, flop / best * 1e-9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 4));
    for (int c = 1; c <= 2; c++) {
        run<1, 2, 0, 0>("1x2 order(ma,mb,ca,cb)", sms, out, c);
        run<1, 2, 1, 0>("1x2 order(ca,ma,mb,cb)", sms, out, c);
        run<1, 2, 0, 1>("1x2 double-buffered frags", sms, out, c);
    }
    run<2, 2, 0, 0>("2x2 order(ma,mb,ca,cb)", sms, out, 1);
    run<2, 2, 0, 1>("2x2 double-buffered frags", sms, out, 1);
    run<1, 1, 0, 0>("1x1", sms, out, 2);
    return 0;
}
