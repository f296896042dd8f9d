// Probe 2: add the per-stage bookkeeping of the Gram kernel to the bare LDS+DMMA loop, one piece at a time.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok != 0;
}
// FLAGS: 1 = dynamic ring slot base, 2 = syncwarp + arrive per stage, 4 = test_wait probe per stage, 8 = KQS 2 instead of 4
template <int FLAGS>
__global__ void __launch_bounds__(256) k(double* out, int stages, int nslots) {
    extern __shared__ double sm[];
    __shared__ uint64_t bars[8];
    constexpr int KQS = (FLAGS & 8) ? 2 : 4;
    constexpr int STAGE_D = 2 * KQS * 384;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < STAGE_D * nslots; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    if (threadIdx.x == 0) for (int i = 0; i < 8; i++) mbar_init(&bars[i], 8);
    __syncthreads();
    double acc[2][3][3][2];
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
        for (int x = 0; x < 3; x++)
#pragma unroll
            for (int y = 0; y < 3; y++) { acc[b][x][y][0] = 0; acc[b][x][y][1] = 0; }
    const int offA = (warp & 3) * 32 + lane, offB = KQS * 384 + (warp >> 2) * 64 + lane;
    unsigned ready_acc = 0;
#pragma unroll 1
    for (int g = 0; g < stages; g++) {
        const int slot = (FLAGS & 1) ? (g & 3) : 0;
        const double* sb = sm + slot * STAGE_D;
        if (FLAGS & 4) ready_acc += mbar_test(&bars[(g + 1) & 3], ((g + 1) >> 2) & 1);
#pragma unroll
        for (int kq = 0; kq < KQS; kq++) {
            double fa[3], fb[2][3];
#pragma unroll
            for (int c = 0; c < 3; c++) fa[c] = sb[offA + (kq * 3 + c) * 128];
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int c = 0; c < 3; c++) fb[b][c] = sb[offB + b * 32 + (kq * 3 + c) * 128];
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int x = 0; x < 3; x++)
#pragma unroll
                    for (int y = 0; y < 3; y++) dmma(acc[b][x][y][0], acc[b][x][y][1], fa[x], fb[b][y]);
        }
        if (FLAGS & 2) { __syncwarp(); if (lane == 0) mbar_arrive(&bars[slot]); }
    }
    double s = ready_acc;
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
        for (int x = 0; x < 3; x++)
#pragma unroll
            for (int y = 0; y < 3; y++) s += acc[b][x][y][0] + acc[b][x][y][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int FLAGS>
void run(const char* name, int sms, double* out, int ctas) {
    constexpr int KQS = (FLAGS & 8) ? 2 : 4;
    const int nslots = 4, stages = 16000 * 4 / KQS;
    const int smem = 2 * KQS * 384 * 8 * nslots;
    CK(cudaFuncSetAttribute(k<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0)); k<FLAGS><<<sms * ctas, 256, smem>>>(out, stages, nslots); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double flop = 2.0 * 256 * 18.0 * KQS * stages * 8.0 * sms * ctas;
    printf("%-52s ctas/sm %d: %.2f TFLOP/s\n", name, ctas, flop / best * 1e-9);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 4));
    for (int c = 1; c <= 2; c++) {
        run<0>("bare", sms, out, c);
        run<1>("+dynamic slot", sms, out, c);
        run<3>("+dynamic slot +arrive", sms, out, c);
        run<7>("+dynamic slot +arrive +test_wait", sms, out, c);
        run<15>("+dynamic slot +arrive +test_wait, 8 atoms/stage", sms, out, c);
    }
    return 0;
}
