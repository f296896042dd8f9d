// Probe: do FP64 instructions and tcgen05 int8 MMAs share an execution pipe on sm_100a?  (ncu reports
// sm__pipe_shared_cycles_active = tensor + fp64 for the int8 Gram kernels.)  One CTA per SM: warp 1 issues a long run of
// tcgen05.mma.kind::i8 (A in TMEM, N = 96), warps 4..19 run independent DFMA chains (the shape of the Gram epilogue).  Three
// launches: MMAs alone, DFMAs alone, both; if the pipes are independent the third takes max(first, second), if they exclude
// each other it takes their sum.  Output: JSON on stdout (profiles/r02_shared_pipe_probe.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/shared_pipe_probe tools/shared_pipe_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int FKIND>   // FKIND 0: DFMA, 1: FFMA, 2: IMAD
__global__ void __launch_bounds__(640, 1) probe(int mma_iters, int f_iters, double* sink, unsigned long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 4 * N * 32; i += blockDim.x) smem[i] = (unsigned char)(i * 37 + 11);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const unsigned long long t0 = clock64();
    if (warp == 1 && mma_iters > 0) {
        if (elect_one()) {
            uint64_t bd0 = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint64_t)(((N / 8 * 128) >> 4) & 0x3FFFu) << 16) |
                           ((uint64_t)((128 >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
#pragma unroll 1
            for (int it = 0; it < mma_iters; it += 16) {
#pragma unroll
                for (int acc = 0; acc < 4; ++acc)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma_ts(tmem + 32 + acc * N, tmem + ks * 8, bd0 + (uint64_t)((ks * N * 32) >> 4), IDESC, ks > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        if ((threadIdx.x & 31) == 0) cyc[2 * blockIdx.x] = clock64() - t0;
    }
    if (warp >= 4 && f_iters > 0) {
        if (FKIND == 0) {
            double a[8], m = 1.0000001, c = 1e-9 * threadIdx.x;
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = k;
#pragma unroll 1
            for (int it = 0; it < f_iters; ++it) {
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
            }
            double s = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += a[k];
            sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
        } else if (FKIND == 1) {
            float a[8], m = 1.0000001f, c = 1e-9f * threadIdx.x;
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = k;
#pragma unroll 1
            for (int it = 0; it < f_iters; ++it) {
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
            }
            float s = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += a[k];
            sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
        } else {
            int a[8], m = 3 + threadIdx.x, c = 7;
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = k;
#pragma unroll 1
            for (int it = 0; it < f_iters; ++it) {
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = a[k] * m + c;
            }
            int s = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += a[k];
            sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
        }
        if (warp == 4 && (threadIdx.x & 31) == 0) cyc[2 * blockIdx.x + 1] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int FKIND>
static double run(int sms, int mi, int fi, double* sink, unsigned long long* d_cyc, double* mma_cyc, double* f_cyc) {
    const int smem = 4 * N * 32 + 1024;
    CK(cudaFuncSetAttribute(probe<N, FKIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaMemset(d_cyc, 0, sizeof(unsigned long long) * 512));
    probe<N, FKIND><<<sms, 640, smem>>>(mi ? 256 : 0, fi ? 64 : 0, sink, d_cyc);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        probe<N, FKIND><<<sms, 640, smem>>>(mi, fi, sink, d_cyc);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    unsigned long long cyc[512];
    CK(cudaMemcpy(cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
    double m = 0, f = 0;
    for (int i = 0; i < sms; ++i) { m += (double)cyc[2 * i]; f += (double)cyc[2 * i + 1]; }
    *mma_cyc = m / sms; *f_cyc = f / sms;
    return best;
}

template <int FKIND>
static void family(const char* name, int sms, double* sink, unsigned long long* d_cyc, bool last) {
    const int mi = 1 << 16, fi = 1 << 13;   // 65536 MMAs (~3.1M cycles at 48 each); 16 warps x 8 x 8192 FMAs
    double mc, fc;
    const double t_m = run<96, FKIND>(sms, mi, 0, sink, d_cyc, &mc, &fc); const double mma_alone = mc;
    const double t_f = run<96, FKIND>(sms, 0, fi, sink, d_cyc, &mc, &fc); const double f_alone = fc;
    const double t_b = run<96, FKIND>(sms, mi, fi, sink, d_cyc, &mc, &fc);
    printf(" \"%s\": {\"mma_alone_ms\": %.3f, \"other_alone_ms\": %.3f, \"both_ms\": %.3f, \"sum_ms\": %.3f, \"max_ms\": %.3f,\n"
           "   \"mma_cycles_alone\": %.0f, \"mma_cycles_with_other\": %.0f, \"other_cycles_alone\": %.0f, \"other_cycles_with_mma\": %.0f}%s\n",
           name, t_m, t_f, t_b, t_m + t_f, t_m > t_f ? t_m : t_f, mma_alone, mc, f_alone, fc, last ? "" : ",");
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* sink; unsigned long long* d_cyc;
    CK(cudaMalloc(&sink, sizeof(double) * sms * 640));
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 512));
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n \"what\": \"per SM: one warp issues 65536 tcgen05.mma.kind::i8 (M128 N96 K32, A in TMEM) while 16 warps run 8 independent "
           "FMA chains of 8192 steps each, in FP64 / FP32 / int32; both_ms = max -> independent pipes, = sum -> one shared pipe\",\n", prop.name, sms);
    family<0>("dfma", sms, sink, d_cyc, false);
    family<1>("ffma", sms, sink, d_cyc, false);
    family<2>("imad", sms, sink, d_cyc, true);
    printf("}\n");
    return 0;
}
