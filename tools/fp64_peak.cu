// Microbenchmark: FP64 DMMA (mma.sync m8n8k4 / m16n8k8) and DFMA peak on sm_100a, pinned D2H/H2D
// bandwidth.  Used once per pool to obtain the FP64 roofline denominator (MEASURED_PEAKS.json has
// no FP64 entry).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma884(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i][0] = 0; acc[i][1] = 0; }
    double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma1688(double* out, int iters, double a0, double b0) {
    double acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i][0] = 0; acc[i][1] = 0; acc[i][2] = 0; acc[i][3] = 0; }
    double a[4], b[2];
    for (int i = 0; i < 4; i++) a[i] = a0 + threadIdx.x + i;
    for (int i = 0; i < 2; i++) b[i] = b0 - threadIdx.x - i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) ffma(float* out, int iters, float a0, float b0) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = i;
    float a = a0 + threadIdx.x * 1e-9f, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\",\n", p.name, p.multiProcessorCount, p.major, p.minor);
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 16));
    const int iters = 20000;
    for (int warps_per_sm_x : {1, 2, 4, 8}) {  // CTAs of 256 threads per SM
        int grid = sms * warps_per_sm_x;
        float ms = time_ms([&] { dmma884<16><<<grid, 256>>>(out, iters, 1.0, 2.0); });
        double flop = 2.0 * 256 * 16 * (double)iters * 8 * grid;
        printf(" \"dmma884_ctas%d_tflops\": %.3f,\n", warps_per_sm_x, flop / ms * 1e-9);
        ms = time_ms([&] { dmma1688<8><<<grid, 256>>>(out, iters, 1.0, 2.0); });
        flop = 2.0 * 1024 * 8 * (double)iters * 8 * grid;
        printf(" \"dmma1688_ctas%d_tflops\": %.3f,\n", warps_per_sm_x, flop / ms * 1e-9);
        ms = time_ms([&] { dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
        flop = 2.0 * 16 * (double)iters * 256 * grid;
        printf(" \"dfma_ctas%d_tflops\": %.3f,\n", warps_per_sm_x, flop / ms * 1e-9);
    }
    {
        int grid = sms * 8;
        float ms = time_ms([&] { ffma<16><<<grid, 256>>>((float*)out, iters, 1.0000001f, 1e-9f); });
        double flop = 2.0 * 16 * (double)iters * 256 * grid;
        printf(" \"ffma_tflops\": %.3f,\n", flop / ms * 1e-9);
    }
    // single-warp latency of dependent DMMA chain (cycles-ish via time, 1 block of 32 threads)
    {
        float ms1 = time_ms([&] { dmma884<1><<<1, 32>>>(out, 200000, 1.0, 2.0); });
        float ms8 = time_ms([&] { dmma884<8><<<1, 32>>>(out, 200000, 1.0, 2.0); });
        printf(" \"dmma884_dep_chain_ns\": %.2f, \"dmma884_1warp_8acc_ns_per_mma\": %.2f,\n", ms1 * 1e6 / 200000, ms8 * 1e6 / 200000 / 8);
    }
    // sustained run (~3 s) for power-capped rate
    {
        int grid = sms * 8;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        for (; n < 60; n++) dmma884<16><<<grid, 256>>>(out, iters, 1.0, 2.0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flop = 2.0 * 256 * 16 * (double)iters * 8 * grid * n;
        printf(" \"dmma884_sustained_tflops\": %.3f, \"dmma884_sustained_s\": %.2f,\n", flop / ms * 1e-9, ms * 1e-3);
    }
    // pinned host bandwidth
    {
        size_t bytes = 1ull << 30;
        void *h, *d; CK(cudaMallocHost(&h, bytes)); CK(cudaMalloc(&d, bytes));
        float ms = time_ms([&] { CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost)); }, 3);
        printf(" \"d2h_pinned_gbs\": %.2f,\n", bytes / ms * 1e-6);
        ms = time_ms([&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); }, 3);
        printf(" \"h2d_pinned_gbs\": %.2f,\n", bytes / ms * 1e-6);
        void* d2; CK(cudaMalloc(&d2, bytes));
        ms = time_ms([&] { CK(cudaMemcpyAsync(d2, d, bytes, cudaMemcpyDeviceToDevice)); }, 5);
        printf(" \"d2d_copy_gbs_rw\": %.2f,\n", 2.0 * bytes / ms * 1e-6);
    }
    printf(" \"done\": true}\n");
    return 0;
}
