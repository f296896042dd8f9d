// Microbenchmark: dense int8 tensor-core peak of tcgen05.mma.kind::i8 on sm_100a (B200), the roofline denominator of the
// int8 Gram kernels (csrc/ozaki.cu).  MEASURED_PEAKS.json only holds the bf16 rate.  One CTA per SM, one elected thread
// issues a long run of MMAs on operands that never change (A and B tiles resident in shared memory, or A resident in
// TMEM for the ".ts" form the Gram kernels use), no epilogue: the figure is what the tensor pipe can sustain for each
// instruction shape, M = 128, K = 32, cta_group::1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8_peak tools/i8_peak.cu && tools/i8_peak > profiles/r02_i8_peak_microbench.json
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
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// MODE 0: A and B from shared memory; MODE 1: A from TMEM (".ts").  NACC accumulators of N columns are cycled through.
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) i8_loop(int iters, unsigned long long* cycles_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    constexpr int A_BYTES = 128 * 32, B_BYTES = N * 32;
    constexpr int KSTEPS = 4;   // operand tiles hold 4 K-steps of 32: the descriptors walk through them like a real main loop
    unsigned char* sA = smem;
    unsigned char* sB = smem + KSTEPS * A_BYTES;
    for (int i = threadIdx.x; i < KSTEPS * (A_BYTES + B_BYTES); i += blockDim.x) smem[i] = (unsigned char)((i * 37 + 11) ^ (i >> 5));
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
    constexpr int ACOLS = MODE == 1 ? 32 : 0;                       // the A operand of 4 K-steps: 8 TMEM columns each
    constexpr int NACC = (512 - ACOLS) / N >= 4 ? 4 : (512 - ACOLS) / N;
    if (warp == 1) {
        const unsigned long long t0 = clock64();
        if (elect_one()) {
            // K-major, no swizzle: [K half (2)][row group][8 rows][16 B]; LBO = distance of the K halves, SBO = 128 B
            const uint64_t ad0 = smem_desc(smem_u32(sA), 128 / 8 * 128, 128);
            const uint64_t bd0 = smem_desc(smem_u32(sB), N / 8 * 128, 128);
#pragma unroll 1
            for (int it = 0; it < iters; it += NACC * KSTEPS) {
#pragma unroll
                for (int acc = 0; acc < NACC; ++acc) {
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        const uint32_t d = tmem + (uint32_t)(ACOLS + acc * N);
                        const uint64_t bd = bd0 + (uint64_t)((ks * B_BYTES) >> 4);
                        if (MODE == 0) mma_ss(d, ad0 + (uint64_t)((ks * A_BYTES) >> 4), bd, IDESC, ks > 0 ? 1u : 0u);
                        else mma_ts(d, tmem + (uint32_t)(ks * 8), bd, IDESC, ks > 0 ? 1u : 0u);
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        asm volatile(
            "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
                smem_u32(&bar))
            : "memory");
        const unsigned long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles_out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int MODE>
static void run(const char* name, int sms, unsigned long long* d_cyc, double* best, bool last) {
    const int smem = 4 * (128 * 32 + N * 32) + 1024;
    CK(cudaFuncSetAttribute(i8_loop<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 1 << 18;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    i8_loop<N, MODE><<<sms, 128, smem>>>(1 << 12, d_cyc);   // warm-up
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        i8_loop<N, MODE><<<sms, 128, smem>>>(iters, d_cyc);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) best_ms = ms;
    }
    unsigned long long cyc[256];
    CK(cudaMemcpy(cyc, d_cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < sms; ++i) mean += (double)cyc[i]; mean /= sms;
    const double ops = (double)sms * iters * 2.0 * 128 * N * 32;
    const double tops = ops / (best_ms * 1e-3) / 1e12;
    printf(" \"%s\": {\"n\": %d, \"a_operand\": \"%s\", \"tops\": %.1f, \"cycles_per_mma\": %.2f, \"ms\": %.3f}%s\n", name, N,
           MODE ? "tmem" : "smem", tops, mean / iters, best_ms, last ? "" : ",");
    if (tops > *best) *best = tops;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned long long* d_cyc;
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 256));
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"sm_clock_khz\": %d,\n \"instruction\": \"tcgen05.mma.cta_group::1.kind::i8, M = 128, K = 32, "
           "s32 accumulators in TMEM, operands resident (no loads, no epilogue), 2^18 MMAs per SM\",\n \"shapes\": {\n",
           prop.name, sms, prop.major, prop.minor, clk);
    double best = 0.0, best_kernel_shape = 0.0, dummy = 0.0;
    run<256, 0>("ss_n256", sms, d_cyc, &best, false);
    run<128, 0>("ss_n128", sms, d_cyc, &best, false);
    run<96, 0>("ss_n96", sms, d_cyc, &best, false);
    run<48, 0>("ss_n48", sms, d_cyc, &best, false);
    run<256, 1>("ts_n256", sms, d_cyc, &best, false);
    run<96, 1>("ts_n96", sms, d_cyc, &best_kernel_shape, false);
    run<80, 0>("ss_n80", sms, d_cyc, &dummy, false);
    run<48, 1>("ts_n48", sms, d_cyc, &dummy, true);
    if (best_kernel_shape > best) best = best_kernel_shape;
    printf(" },\n \"i8_dense_tops\": %.1f,\n \"ts_n96_tops\": %.1f,\n \"note\": \"i8_dense_tops = the best shape (the roofline denominator bench.py uses); "
           "ts_n96 / ts_n48 / ss_n80 are the shapes gram_i8_ts_kernel<4>, <2> and gram_i8_kernel issue\"\n}\n", best, best_kernel_shape);
    return 0;
}
