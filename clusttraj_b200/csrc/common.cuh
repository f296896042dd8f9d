// Shared declarations of the clusttraj_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ct {

// ---- packed ("tile-major") layout read by the Gram kernel --------------------------------------
// packed[tile][kq][c][f][a]  (float64)   tile = frame / 32, f = frame % 32, kq = atom / 4, a = atom % 4,
// c = x,y,z.  One (tile, kq, c) block is 32 frames x 4 atoms = 1 KiB and is exactly the pair of
// operand fragments four m8n8k4 DMMAs read (lane l <-> frame l/4, atom l%4), so fragment loads are one
// conflict-free 256 B LDS.64 per warp, and one pipeline stage of a tile (KQS k4 groups) is ONE contiguous
// run in HBM (KQS x 3 KiB) that a single cp.async.bulk (TMA) moves.
constexpr int FT = 32;                          // frames per packed tile
constexpr int KQ_DOUBLES = 3 * FT * 4;          // one k4 group of one tile: 384 doubles = 3 KiB

struct FixList {
    int2* pairs;            // flagged (i, j)
    unsigned int* count;    // tickets handed out in the current chunk (may exceed capacity)
    unsigned int capacity;
};

struct GramParams {
    const double* packed;   // tile-major centred coordinates
    const double* G;        // per-frame HALF sum of squares (0.5 * sum |x|^2), padded to whole tiles
    double* out;            // condensed slab; element idx(i,j) lives at out[idx(i,j) - idx_base]
    long long idx_base;
    int M;                  // frames
    int kq_total;           // padded atoms / 4
    int row_begin, row_end; // rows of the matrix to produce
    double inv_k2;          // 2 / kept atoms
    double flag_rel;        // pairs with E < flag_rel * (G_i + G_j) are re-done by the explicit-rotation path
    FixList fix;
};

// parameters of the int8 tensor-core (tcgen05) Gram kernel, ozaki.cu
struct OzParams {
    const uint8_t* packA;   // [ti][kb][slice][kc16 (8)][row group (16)][8 rows][16 B]
    const uint8_t* packB;
    const uint8_t* packT;   // [ti][128 rows][slice][128 B] (single K-block only): the A tile as TMEM lanes read it   // [tj][kb][slice][kc16 (8)][row group (12)][8 rows][16 B]
    const double* Gq;       // half sums of squares of the QUANTISED coordinates, real units
    const double* qinfo;    // [0] = 2^(-2f), [1] = f, [2] = max |x|
    double* out;
    long long idx_base;
    long long total_blocks;
    int M, nkb, last_chunks;   // K-blocks; 32-atom chunks in the last K-block (1..4)
    int row_begin, row_end;
    double inv_k2, flag_rel;
    FixList fix;
};

__host__ __device__ __forceinline__ long long condensed_row_base(long long M, long long i) {
    // index of (i, i+1) in SciPy's condensed order (clusttraj/distmat.py:78 flattens rows in order)
    return M * i - (i * (i + 1)) / 2;
}

// ---- mbarrier / bulk-copy (TMA) PTX ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// Same, with a suspend-time hint: the waiting warp sleeps in hardware (up to `ns`) instead of re-issuing the
// probe every few cycles, which would take issue slots from the warps it is waiting for.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns = 20000u) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared-memory sequence counters (absolute block numbers, so a waiter may be any number of phases ahead)
__device__ __forceinline__ void seq_publish(unsigned int* ctr, unsigned int v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(ctr)), "r"(v) : "memory");
}
__device__ __forceinline__ void seq_wait_ge(const unsigned int* ctr, unsigned int v) {
    unsigned int cur;
    do {
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(cur) : "r"(smem_u32(ctr)) : "memory");
        if (cur < v) __nanosleep(32);
    } while (cur < v);
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ct
