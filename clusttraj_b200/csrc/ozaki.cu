// K2' — the same all-pairs cross-covariance + fused RMSD epilogue as gram.cu, but with the contraction on
// the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM) instead of the FP64 DMMA
// pipe.  B200 has no FP64 tcgen05 kind, so the FP64 Gram product is rebuilt EXACTLY from integer pieces
// (the "Ozaki scheme" for matrix products on integer tensor cores):
//
//   * every centred coordinate is rounded once to a 31-bit fixed-point number X = rint(x * 2^f), f chosen
//     per run from max|x| (quantum 2^-f ~ 6e-8 A for a 40 A molecule).  That is a rigid perturbation of the
//     INPUT by at most half a quantum per coordinate, so the RMSD moves by at most ~2^-f whatever the
//     cancellation in G_i + G_j - 2 lambda (RMSD is 1-Lipschitz in the rms displacement);
//   * X is cut into four balanced base-256 digits (int8 slices, X = sum_s d_s 256^(3-s));
//   * sum_k X_i X_j = sum_{s,t} 256^(6-s-t) sum_k d_s d_t: every digit product sum is an int8 GEMM whose int32
//     accumulator is exact (|.| < 2^25), the products with the same s+t ("level") share one accumulator;
//     level 6 (d_3 d_3, weight 1) is below the FP64 rounding of the result and is dropped: 15 GEMMs;
//   * the epilogue recombines the six levels in FP64 (exact int -> double conversions, weights 2^(8(6-q)-2f)),
//     takes G from the SAME quantised coordinates, and solves the 3x3 problem as gram.cu does.
//
// Shape on the tensor core: one tcgen05.mma is M=128 x N=96 (or 48) x K=32 (int8).  The 128 A rows are 40 frames:
// TMEM quarter w, lane 3g+a  <->  frame 10w+g, coordinate a (two zero rows per quarter), so the three rows of a
// frame's covariance sit in three adjacent TMEM lanes of ONE epilogue warp; the B rows are 32 (16) frames x 3
// coordinates.  A CTA owns a contiguous run of blocks of the upper triangle (row-major).  Operands are stored in HBM
// already in the order the hardware wants (ct_load_frames -> quant_pack_kernel), so a pipeline stage is a handful of
// 1-D cp.async.bulk copies and the shared-memory descriptors need no swizzle.
//
// Three kernels (launch_gram_i8 picks by the padded atom count kp):
//   gram_i8_ts_kernel<8>  kp <= 128: A tile resident in TMEM (".ts" MMA), B streams, N = 96, 16 epilogue warps;
//   gram_i8_ts_kernel<4>  kp <= 320: the same with N = 48 (A takes up to 320 of the 512 TMEM columns), 16 epilogue warps;
//   gram_i8_kernel        larger kp: A and B both stream through shared memory in K-blocks of 128 atoms; N = 80 so that ALL
//                         six level accumulators (480 TMEM columns) stay resident over the K-blocks of a block and are
//                         converted once; 8 epilogue warps that transpose with shuffles (shared memory is full).  Bound
//                         by the L2 -> shared-memory operand traffic (104 KB per K-block and SM).
// Roles in every kernel: warp 0 = bulk-copy producer (one lane), warp 1 = MMA issuer (one elected lane), warp 2 =
// TMEM allocator, warps 4.. = epilogue.  TMEM (ts kernels): two buffers of 2 levels x N columns; the MMA warp fills one
// "level pair" while the epilogue drains the other.
#include "common.cuh"
#include "solve3x3.cuh"

#include <cstdlib>

namespace ct {

constexpr int OZ_FI = 40;                    // frames per A (row) tile
constexpr int OZ_FJ = 26;                    // frames per B (column) tile of gram_i8_kernel: 78 rows + 2 zero rows
constexpr int OZ_M = 128, OZ_N = 80;         // MMA tile of gram_i8_kernel (six levels x 80 columns = 480 TMEM columns)
constexpr int OZ_KB = 128;                   // atoms per K-block (= bytes per operand row per slice per stage)
constexpr int OZ_SLICES = 4;
constexpr int OZ_A_SLICE = OZ_M * OZ_KB;     // 16 KiB
constexpr int OZ_B_SLICE = OZ_N * OZ_KB;     // 12 KiB
constexpr int OZ_A_BYTES = OZ_SLICES * OZ_A_SLICE;
constexpr int OZ_B_BYTES = OZ_SLICES * OZ_B_SLICE;
constexpr int OZ_STAGE_BYTES = OZ_A_BYTES + OZ_B_BYTES;   // 112 KiB
constexpr int OZ_STAGES = 2;
constexpr int OZ_THREADS = 384;
constexpr int OZ_SMEM_BYTES = OZ_STAGES * OZ_STAGE_BYTES + 128;
constexpr int OZ_TMEM_COLS = 512;
constexpr int OZ_NPAIR = 3;                  // level pairs (levels 0..5): the A-in-TMEM kernels issue them from the most
                                             // significant down (their recombination needs that order), the shared-memory
                                             // kernel in the order (2,3), (4,5), (0,1)


// ---- tcgen05 / TMEM PTX ------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// One lane of a converged warp (elect.sync): with warp-uniform operands the compiler keeps tcgen05 descriptors in
// uniform registers instead of broadcasting them lane by lane (a ~18-instruction loop per MMA otherwise).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// A operand in TMEM (".ts"): row m of A in TMEM lane m, 32 int8 of K per 8 columns
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous (128 B);
// LBO = byte distance between the two 16-byte K halves of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;
}
// Instruction descriptor of kind::i8: D = s32, A = B = signed 8 bit, both K-major, dense.
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_N >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);

// ---- block walk ------------------------------------------------------------------------------------------------
// Blocks (ti, tj) of the upper triangle: rows [40 ti, 40 ti + 40) x columns [32 tj, 32 tj + 32), needed when some
// j > i falls inside, i.e. tj >= (40 ti + 1) / 32.  Row-major enumeration; a CTA takes a contiguous run.
struct OzWalker {
    int ti, tj, ntj;
    int left;
    __host__ __device__ static int first_tj(int ti) { return (OZ_FI * ti + 1) / OZ_FJ; }
    __device__ void init(int ti_begin, int ntj_, long long start, long long count) {
        ntj = ntj_; left = (int)count; ti = ti_begin;
        long long s = start;
        for (;;) {
            const long long n = ntj - first_tj(ti);
            if (s < n) break;
            s -= n; ++ti;
        }
        tj = first_tj(ti) + (int)s;
    }
    __device__ bool valid() const { return left > 0; }
    __device__ void next() {
        --left;
        if (++tj >= ntj) { ++ti; tj = first_tj(ti); }
    }
};

// exact  e * 256 + o  ->  double  without an I2F.F64: add the integer to the bit pattern of 2^52 + 2^51
__device__ __forceinline__ double level_pair(int e, int o) {
    const long long v = (long long)e * 256 + (long long)o;  // |v| < 2^34
    const int hi = (int)(v >> 32) + 0x43380000;
    return __hiloint2double(hi, (int)(unsigned)(v & 0xffffffffLL)) - 6755399441055744.0;
}

// The same for kp <= 128 atoms: every level sum is below 4 * 2^14 * 128 = 2^23 there, so even * 256 + odd still fits
// 32 bits (|v| < 1.62e9) and the 64-bit multiply-add and sign extension are not needed.
// the same two conversions WITHOUT the final subtraction: MAGIC + v, exactly (see the folded recombination in the epilogue)
__device__ __forceinline__ double level_pair_biased(int e, int o) {
    const long long v = (long long)e * 256 + (long long)o;  // |v| < 2^34
    return __hiloint2double((int)(v >> 32) + 0x43380000, (int)(unsigned)(v & 0xffffffffLL));   // 2^52 + 2^51 + v
}
__device__ __forceinline__ double level_pair32_biased(int e, int o) {
    return __hiloint2double(0x43300000, (e * 256 + o) ^ (int)0x80000000);   // 2^52 + 2^31 + v
}
__device__ __forceinline__ double level_pair32(int e, int o) {
    const int v = e * 256 + o;
    return __hiloint2double(0x43300000, v ^ (int)0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}

__device__ __forceinline__ double sel3(int k, double x0, double x1, double x2) { return k == 0 ? x0 : (k == 1 ? x1 : x2); }

__global__ void __launch_bounds__(OZ_THREADS, 1) gram_i8_kernel(const OzParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OZ_STAGES * OZ_STAGE_BYTES);
    uint64_t* full = bars;            // [2] stage loaded (tx bytes)
    uint64_t* empty = bars + 2;       // [2] stage consumed (tcgen05.commit)
    uint64_t* tfull = bars + 4;       // [3] level pair complete over ALL K-blocks of the block (tcgen05.commit)
    uint64_t* tempty = bars + 7;      // [3] level pair drained (8 epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ti_begin = p.row_begin / OZ_FI;
    const int ntj = (p.M + OZ_FJ - 1) / OZ_FJ;
    const long long blk_start = p.total_blocks * blockIdx.x / gridDim.x;
    const long long blk_count = p.total_blocks * (blockIdx.x + 1) / gridDim.x - blk_start;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
#pragma unroll
        for (int s = 0; s < OZ_NPAIR; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, OZ_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0 && lane == 0) {
            // ---- producer: 1-D bulk copies of pre-laid-out operand slices -------------------------------------
            OzWalker w;
            w.init(ti_begin, ntj, blk_start, blk_count);
            int stage = 0; uint32_t phase = 0;
            long long tagA[2] = {-1, -1};
            for (; w.valid(); w.next()) {
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    const int nch = (kb == p.nkb - 1) ? p.last_chunks : 4;
                    const long long tag = (long long)w.ti * p.nkb + kb;
                    const bool loadA = tagA[stage] != tag;
                    tagA[stage] = tag;
                    const uint32_t a_bytes = (uint32_t)nch * (OZ_A_SLICE / 4), b_bytes = (uint32_t)nch * (OZ_B_SLICE / 4);
                    mbar_expect_tx(&full[stage], OZ_SLICES * ((loadA ? a_bytes : 0u) + b_bytes));
                    unsigned char* sA = smem + stage * OZ_STAGE_BYTES;
                    unsigned char* sB = sA + OZ_A_BYTES;
                    const uint8_t* gA = p.packA + ((size_t)w.ti * p.nkb + kb) * OZ_A_BYTES;
                    const uint8_t* gB = p.packB + ((size_t)w.tj * p.nkb + kb) * OZ_B_BYTES;
#pragma unroll
                    for (int s = 0; s < OZ_SLICES; ++s) {
                        if (loadA) bulk_g2s(sA + s * OZ_A_SLICE, gA + s * OZ_A_SLICE, a_bytes, &full[stage]);
                        bulk_g2s(sB + s * OZ_B_SLICE, gB + s * OZ_B_SLICE, b_bytes, &full[stage]);
                    }
                    stage ^= 1; if (stage == 0) phase ^= 1;
                }
            }
        } else if (warp == 1) {
            // ---- MMA issuer ------------------------------------------------------------------------------------
            // All six level accumulators of a block stay in TMEM while its K-blocks stream through shared memory: the
            // int32 sums run over the whole contraction (|.| < 4 * 2^14 * K, no overflow below 32k atoms) and the
            // epilogue converts each level pair ONCE per block instead of once per K-block.
            int stage = 0; uint32_t phase = 0;
            constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);                 // SBO = 128 B, version 1
            for (long long blk = 0; blk < blk_count; ++blk) {
                for (int kb = 0; kb < p.nkb; ++kb) {
                    const int nch = (kb == p.nkb - 1) ? p.last_chunks : 4;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t aBase = smem_u32(smem + stage * OZ_STAGE_BYTES);
                    const uint32_t bBase = aBase + OZ_A_BYTES;
                    // descriptor words: only the 14-bit address field changes from one MMA to the next
                    const uint32_t a_lo = ((uint32_t)(OZ_M * 16 >> 4) << 16) | (aBase >> 4);  // LBO = one K half of all rows
                    const uint32_t b_lo = ((uint32_t)(OZ_N * 16 >> 4) << 16) | (bBase >> 4);
#pragma unroll
                    for (int pr = 0; pr < OZ_NPAIR; ++pr) {
                        if (kb == 0) {   // the epilogue must have drained this pair of the previous block
                            mbar_wait(&tempty[pr], (uint32_t)(blk & 1) ^ 1u);
                            tc_fence_after();
                        }
                        const bool leader = elect_one();
                        if (leader) {
#pragma unroll
                            for (int lv = 0; lv < 2; ++lv) {
                                const int q = 2 * ((pr + 1) % OZ_NPAIR) + lv;   // level pairs in the order (2,3), (4,5), (0,1)
                                const uint32_t d = tmem_base + (uint32_t)(pr * 2 * OZ_N + lv * OZ_N);
                                const int s_lo = q > 3 ? q - 3 : 0, s_hi = q < 3 ? q : 3;
#pragma unroll
                                for (int s = s_lo; s <= s_hi; ++s) {
#pragma unroll
                                    for (int kc = 0; kc < 4; ++kc) {
                                        if (kc < nch) {
                                            // one MMA = 32 atoms = two 16-byte K halves of (rows / 8) core matrices each
                                            const uint64_t ad = ((uint64_t)DESC_HI << 32) |
                                                                (a_lo + (uint32_t)((s * OZ_A_SLICE + kc * 2 * OZ_M * 16) >> 4));
                                            const uint64_t bd = ((uint64_t)DESC_HI << 32) |
                                                                (b_lo + (uint32_t)(((q - s) * OZ_B_SLICE + kc * 2 * OZ_N * 16) >> 4));
                                            umma_i8(d, ad, bd, OZ_IDESC, (kb > 0 || s > s_lo || kc > 0) ? 1u : 0u);
                                        }
                                    }
                                }
                            }
                            if (kb == p.nkb - 1) umma_commit(&tfull[pr]);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(&empty[stage]);
                    __syncwarp();
                    stage ^= 1; if (stage == 0) phase ^= 1;
                }
            }
        }
    } else {
        // ---- epilogue warps ------------------------------------------------------------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int quarter = warp & 3;            // TMEM lanes 32 * quarter .. + 31
        const int half = (warp - 4) >> 2;        // accumulator columns 39 * half .. + 38 = column frames 13 * half .. + 12
        const int g = lane / 3, a = lane - 3 * g;
        const bool row_ok = lane < 30;
        const int src1 = (lane - a + (a + 1) % 3) & 31, src2 = (lane - a + (a + 2) % 3) & 31;
        const int u1 = (a + 2) % 3, u2 = (a + 1) % 3;   // which of the three pairs this lane SENDS in shuffle 1 / 2
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 39);
        const double s2 = p.qinfo[0];
        const double sc[OZ_NPAIR] = {s2 * 16777216.0, s2 * 256.0, s2 * 1099511627776.0};   // 2^24, 2^8, 2^40 (issue order)
        const int last_row = min(p.row_end, p.M - 1);
        OzWalker w;
        w.init(ti_begin, ntj, blk_start, blk_count);
        uint32_t tphase = 0;
        for (; w.valid(); w.next()) {
            double acc[39];
#pragma unroll
            for (int pr = 0; pr < OZ_NPAIR; ++pr) {
                mbar_wait(&tfull[pr], tphase);
                tc_fence_after();
                const uint32_t t0 = tlane + (uint32_t)(pr * 2 * OZ_N);
                // 39 columns of this half as 16 + 16 + 8 (the 40th belongs to the neighbour half / the padding)
                uint32_t e0[16], e1[16], e2[8], o0[16], o1[16], o2[8];
                tmem_ld16(t0, e0); tmem_ld16(t0 + 16, e1); tmem_ld8(t0 + 32, e2);
                tmem_ld16(t0 + OZ_N, o0); tmem_ld16(t0 + OZ_N + 16, o1); tmem_ld8(t0 + OZ_N + 32, o2);
                tmem_wait_ld();
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                    const double v0 = level_pair((int)e0[n], (int)o0[n]) * sc[pr], v1 = level_pair((int)e1[n], (int)o1[n]) * sc[pr];
                    acc[n] = pr == 0 ? v0 : acc[n] + v0;
                    acc[16 + n] = pr == 0 ? v1 : acc[16 + n] + v1;
                }
#pragma unroll
                for (int n = 0; n < 7; ++n) {
                    const double v2 = level_pair((int)e2[n], (int)o2[n]) * sc[pr];
                    acc[32 + n] = pr == 0 ? v2 : acc[32 + n] + v2;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[pr]);
            }
            tphase ^= 1;
            // ---- 3x3 solve: lanes 3g, 3g+1, 3g+2 hold the three rows of the covariances of frame i with 13 frames j;
            // lane a takes the pairs jj = 3m + a and collects the two other rows with shuffles.  The rows arrive in
            // cyclic order (a, a+1, a+2), which leaves |S|^2, |cof S|^2 and det S unchanged.
            const int i = w.ti * OZ_FI + quarter * 10 + g;
            const int j0 = w.tj * OZ_FJ + half * 13;
            const double gi = p.Gq[i];
            const bool row_mine = row_ok && i >= p.row_begin && i < last_row;
            const long long row_off = condensed_row_base(p.M, i) - p.idx_base - i - 1;
            double qa[5], qb[5], qd[5], hg[5], lam[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                double s9[9];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const double r0 = acc[9 * m + b];
                    const double r1 = (9 * m + 3 + b < 39) ? acc[(9 * m + 3 + b) % 39] : 0.0;
                    const double r2 = (9 * m + 6 + b < 39) ? acc[(9 * m + 6 + b) % 39] : 0.0;
                    s9[b] = sel3(a, r0, r1, r2);
                    s9[3 + b] = __shfl_sync(0xffffffffu, sel3(u1, r0, r1, r2), src1);
                    s9[6 + b] = __shfl_sync(0xffffffffu, sel3(u2, r0, r1, r2), src2);
                }
                const int jj = 3 * m + a;
                const int j = j0 + jj;
                const bool live = row_ok && jj < 13 && j < p.M;
                if (!live) {
                    // idle lanes / padding: a well-conditioned stand-in (identity) so that no lane drags its warp
                    // through the solver's slow path; the result is never stored
#pragma unroll
                    for (int k = 0; k < 9; ++k) s9[k] = (k % 4 == 0) ? 1.0 : 0.0;
                }
                quartic_invariants(s9, qa[m], qb[m], qd[m]);
                hg[m] = live ? gi + p.Gq[j] : 3.0;  // half sums: G_i = G_j = 3 for the identity stand-in
            }
            const unsigned okmask = largest_root_f64<5>(qa, qb, qd, hg, lam);
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const int jj = 3 * m + a;
                const int j = j0 + jj;
                const double eh = hg[m] - lam[m];
                const double v = (eh > 0.0 ? eh : 0.0) * p.inv_k2;
                const double y = rsqrt_approx_f64(v + 1e-300);        // sqrt(v) = v / sqrt(v), one Newton step
                double r = v * y;
                r = fma(fma(-r, r, v), 0.5 * y, r);
                const bool ok = (okmask >> m) & 1u;
                const bool mine = row_mine && jj < 13 && i < j && j < p.M;
                if ((!ok || eh < p.flag_rel * hg[m]) && mine) {
                    const unsigned int t = atomicAdd(p.fix.count, 1u);
                    if (t < p.fix.capacity) p.fix.pairs[t] = make_int2(i, j);
                    else r = -1.0;
                }
                if (mine) __stcs(p.out + row_off + j, r);   // streaming store: the 8 B/pair output must not evict the operand tiles from L2
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, OZ_TMEM_COLS);
}


// ---- K <= 320 atoms: A operand resident in TMEM, sixteen epilogue warps -------------------------------------------
// With the whole contraction in one K-block the A tile (40 frames x 4 slices x kp bytes) is kp TMEM columns next to the
// accumulators.  It is written there once per row of blocks (tcgen05.st, straight from global memory, by the warps that
// own the TMEM lanes) and every MMA takes it from TMEM (".ts"), so shared memory only carries the streaming B tiles: the
// MMA no longer reads more than the 128 B/clk shared memory can deliver, and the freed shared memory holds the hand-off
// buffer in which the row-per-lane accumulator layout is transposed: a lane writes its (frame i, coordinate a) row, and
// reads back the whole 3x3 of one pair, lanes running along j so that the RMSD stores are contiguous runs.
// Two shapes, chosen by the padded atom count kp (a multiple of 32, all of the contraction in ONE shared-memory stage):
//   PW = 8: column tiles of 32 frames (N = 96), two level-pair buffers (384 TMEM columns), A up to 128 columns -> kp <= 128
//   PW = 4: column tiles of 16 frames (N = 48), A up to 320 columns -> kp <= 320; THREE level-pair buffers (288 columns) while
//           kp <= 224, so that the MMA warp runs a whole block ahead of the epilogue, else two
// Both have FOUR epilogue warpgroups (16 warps, 4 per scheduler): warpgroup wg owns 3 PW accumulator columns = PW column
// frames of every row.  (The N = 48 shape used to run with two warpgroups of 24 columns: two warps per scheduler left its
// epilogue latency-bound at 30 % issue utilisation.)
template <int PW>
struct TsCfg {
    static constexpr int WGS = 4;                      // epilogue warpgroups
    static constexpr int FJ = WGS * PW;                // frames per column tile
    static constexpr int N = 3 * FJ;                   // MMA N
    static constexpr int CW = 3 * PW;                  // accumulator columns per warpgroup
    static constexpr int THREADS = 128 + 128 * WGS;    // 4 control warps + 16 epilogue warps
    static constexpr int KP_MAX = PW == 8 ? 128 : 320;
    static constexpr int PITCH = PW == 8 ? 73 : 41;    // doubles per frame row of a warpgroup's hand-off region: 9 PW padded to
                                                       // 9 (mod 16) 8-byte bank units
    static constexpr int NR = PW == 8 ? 3 : 2;         // solve passes of the busiest warps of a warpgroup (the others: NR - 1)
    static constexpr int B_STAGE = OZ_SLICES * N * KP_MAX;
    static constexpr int SCRATCH = WGS * OZ_FI * PITCH * 8;
    static constexpr int TABLES = WGS * 2 * OZ_FI * 8; // per warpgroup: output offset and G of the 40 rows of the block row
    static constexpr int SMEM = 2 * B_STAGE + SCRATCH + 256 + TABLES;
    static constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);
    // TMEM: nbuf level-pair buffers of 2 N columns, then the A operand (kp columns)
    __host__ __device__ static int nbuf(int kp) { return (OZ_TMEM_COLS - kp) / (2 * N) >= 3 ? 3 : 2; }
    static_assert(4 * N + KP_MAX <= OZ_TMEM_COLS, "two level-pair buffers and the A operand must fit the 512 TMEM columns");
};
#ifndef TS_CTRL_REGS
#define TS_CTRL_REGS 32    // PW = 8: registers the four control warps keep (96 x 640 are the CTA's allocation) ...
#define TS_EPI_REGS 112    // ... and what that leaves for each of the 512 epilogue threads: (61440 - 128 x 32) / 512
#define TS_CTRL_REGS_N48 48   // PW = 4: the unrolled MMA issue wants more uniform-register feeding, the epilogue (12 accumulator
#define TS_EPI_REGS_N48 104   // columns, two solve passes) less
#endif

// upper-triangle block walk for column tiles of FJ frames (see OzWalker)
template <int FJ>
struct TsWalker {
    int ti, tj, ntj;
    int left;
    __host__ __device__ static int first_tj(int ti) { return (OZ_FI * ti + 1) / FJ; }
    __device__ void init(int ti_begin, int ntj_, long long start, long long count) {
        ntj = ntj_; left = (int)count; ti = ti_begin;
        long long s = start;
        for (;;) {
            const long long n = ntj - first_tj(ti);
            if (s < n) break;
            s -= n; ++ti;
        }
        tj = first_tj(ti) + (int)s;
    }
    __device__ bool valid() const { return left > 0; }
    __device__ void next() {
        --left;
        if (++tj >= ntj) { ++ti; tj = first_tj(ti); }
    }
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
// L2 eviction priorities: the B tiles are re-read by every CTA of the grid (keep), an A tile by one CTA once (do not keep)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
// 32 bytes (one 32-atom chunk of one A row) in one 256-bit load that is not kept in L1 and leaves L2 first
__device__ __forceinline__ void ldg256_evict_first(const void* ptr, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(ptr));
}

// The MMAs of one level pair of one block (issued by one elected thread): levels q = 2 ((pr + 1) % 3) + lv, lv = 0, 1, each the sum
// over the digit-slice pairs (s, q - s) and over the 32-atom chunks kc of  A[slice s, chunk kc] (TMEM) x B[slice q - s, chunk kc]
// (shared memory).  NCH > 0: chunk count known at compile time, every operand address is the base plus an immediate;
// NCH = 0: run-time chunk count `nch`.
template <int N, int NCH>
__device__ __forceinline__ void ts_issue_pair(int pr, uint32_t d0, uint32_t a0, uint32_t b_lo, int nch = NCH) {
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version of sm_100
    constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);
    const int kp = 32 * nch;
    const uint32_t b_slice16 = (uint32_t)(N * kp) >> 4, a_slice_cols = (uint32_t)kp >> 2;
#pragma unroll
    for (int lv = 0; lv < 2; ++lv) {
        const int q = 2 * ((pr + 1) % OZ_NPAIR) + lv;
        const uint32_t d = d0 + (uint32_t)(lv * N);
        const int s_lo = q > 3 ? q - 3 : 0, s_hi = q < 3 ? q : 3;
#pragma unroll
        for (int s = s_lo; s <= s_hi; ++s) {
            if constexpr (NCH > 0) {
#pragma unroll
                for (int kc = 0; kc < NCH; ++kc) {
                    // one MMA = 32 atoms: A from 8 TMEM columns, B = two 16-byte K halves of N / 8 core matrices
                    const uint64_t bd = ((uint64_t)DESC_HI << 32) | (b_lo + (uint32_t)(q - s) * b_slice16 + (uint32_t)(kc * 2 * N));
                    umma_i8_ts(d, a0 + (uint32_t)s * a_slice_cols + (uint32_t)(kc * 8), bd, IDESC, (s > s_lo || kc > 0) ? 1u : 0u);
                }
            } else {
#pragma unroll 2
                for (int kc = 0; kc < nch; ++kc) {
                    const uint64_t bd = ((uint64_t)DESC_HI << 32) | (b_lo + (uint32_t)(q - s) * b_slice16 + (uint32_t)(kc * 2 * N));
                    umma_i8_ts(d, a0 + (uint32_t)s * a_slice_cols + (uint32_t)(kc * 8), bd, IDESC, (s > s_lo || kc > 0) ? 1u : 0u);
                }
            }
        }
    }
}

template <int PW>
__global__ void __launch_bounds__(TsCfg<PW>::THREADS, 1) gram_i8_ts_kernel(const OzParams p) {
    using Cfg = TsCfg<PW>;
    constexpr int FJ = Cfg::FJ, N = Cfg::N, CW = Cfg::CW, PITCH = Cfg::PITCH, NR = Cfg::NR;
    extern __shared__ __align__(1024) unsigned char smem[];
    double* scratch = reinterpret_cast<double*>(smem + 2 * Cfg::B_STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * Cfg::B_STAGE + Cfg::SCRATCH);
    uint64_t* bfull = bars;           // [2] B stage loaded
    uint64_t* bempty = bars + 2;      // [2] B stage consumed
    uint64_t* tfull = bars + 4;       // [3] level pair complete in TMEM
    uint64_t* tempty = bars + 7;      // [3] level pair in registers (buffer free)
    uint64_t* a_ready = bars + 10;    // A tile of the current block row is in TMEM (4 warps)
    uint64_t* a_free = bars + 11;     // every MMA of the previous block row has completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ti_begin = p.row_begin / OZ_FI;
    const int ntj = (p.M + FJ - 1) / FJ;
    const long long blk_start = p.total_blocks * blockIdx.x / gridDim.x;
    const long long blk_count = p.total_blocks * (blockIdx.x + 1) / gridDim.x - blk_start;
    const int nch = p.last_chunks;    // 32-atom chunks of the contraction (all in one stage)
    const int kp = 32 * nch;          // padded atoms = bytes per operand row per slice
    const int nbuf = Cfg::nbuf(kp);   // level-pair buffers in TMEM
    const uint32_t acol = (uint32_t)(nbuf * 2 * N);   // first TMEM column of the A operand

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
#pragma unroll
        for (int s = 0; s < 3; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4 * Cfg::WGS); }
        mbar_init(a_ready, 4); mbar_init(a_free, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, OZ_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform

    if (warp < 4) {
        if constexpr (PW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TS_CTRL_REGS));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TS_CTRL_REGS_N48));
        if (warp == 0 && lane == 0) {
            // ---- producer: the B tile of every block, four slices ------------------------------------------------
            TsWalker<FJ> w;
            w.init(ti_begin, ntj, blk_start, blk_count);
            int stage = 0; uint32_t phase = 0;
            const uint32_t b_slice = (uint32_t)(N * kp);   // one digit slice of a B tile
            const uint64_t keep = l2_policy_evict_last();
            for (; w.valid(); w.next()) {
                mbar_wait_sleep(&bempty[stage], phase ^ 1, 2000u);
                mbar_expect_tx(&bfull[stage], OZ_SLICES * b_slice);
                unsigned char* dst = smem + stage * Cfg::B_STAGE;
                const uint8_t* src = p.packB + (size_t)w.tj * (OZ_SLICES * b_slice);
#pragma unroll
                for (int s = 0; s < OZ_SLICES; ++s) bulk_g2s_hint(dst + s * b_slice, src + s * b_slice, b_slice, &bfull[stage], keep);
                stage ^= 1; if (stage == 0) phase ^= 1;
            }
        } else if (warp == 1) {
            // ---- MMA issuer -----------------------------------------------------------------------------------------
            TsWalker<FJ> w;
            w.init(ti_begin, ntj, blk_start, blk_count);
            int stage = 0; uint32_t phase = 0;
            int tb = 0; uint32_t tphase = 0;
            int cur_ti = -1; uint32_t a_phase = 0;
            for (; w.valid(); w.next()) {
                if (w.ti != cur_ti) {
                    // new block row: the accumulate warps replace the A tile once the MMAs that read the old one are done
                    if (cur_ti >= 0 && elect_one()) umma_commit(a_free);
                    __syncwarp();
                    mbar_wait(a_ready, a_phase); a_phase ^= 1;
                    tc_fence_after();
                    cur_ti = w.ti;
                }
                mbar_wait(&bfull[stage], phase);
                tc_fence_after();
                const uint32_t b_lo = ((uint32_t)(N * 16 >> 4) << 16) | (smem_u32(smem + stage * Cfg::B_STAGE) >> 4);
#pragma unroll
                for (int pr = 0; pr < OZ_NPAIR; ++pr) {
                    mbar_wait(&tempty[tb], tphase ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
                        // Level pairs in the order (2,3), (4,5), (0,1) = 28, 20 and 12 MMAs per 128 atoms: with two TMEM
                        // buffers the third pair of a block can only be issued once the epilogue has taken the first one
                        // out of its buffer; it is the cheapest, so the epilogue waits least for the tensor pipe.
                        // (A variant with the chunk loop unrolled per contraction length and compile-time operand offsets
                        // issued an MMA in ~26 instead of ~57 cycles and made the kernels SLOWER -- cfg2 40.0 -> 35.5, K = 300
                        // 21.0 -> 20.6 Gpairs/s: issue is not the limiter, and the larger code of five instantiations costs
                        // instruction-cache misses in the epilogue warps.)
                        ts_issue_pair<N, 0>(pr, tmem_base + (uint32_t)(tb * 2 * N), tmem_base + acol, b_lo, nch);
                        umma_commit(&tfull[tb]);
                    }
                    __syncwarp();
                    if (++tb == nbuf) { tb = 0; tphase ^= 1; }
                }
                if (elect_one()) umma_commit(&bempty[stage]);
                __syncwarp();
                stage ^= 1; if (stage == 0) phase ^= 1;
            }
        }
    } else {
        // ---- epilogue warps (4 warpgroups): warpgroup wg owns accumulator columns CW wg .. CW wg + CW - 1 = PW column frames ----
        // 96 registers x 640 threads are the CTA's allocation: the control warps keep 32, which leaves 112 for each epilogue thread
        if constexpr (PW == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TS_EPI_REGS));
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TS_EPI_REGS_N48));
        const int quarter = warp & 3;                 // TMEM lanes 32 * quarter .. + 31
        const int wg = (warp - 4) >> 2;               // 0..3
        const int wt = tid - 128 - wg * 128;          // thread index inside the warpgroup
        const int g = lane / 3, a = lane - 3 * g;
        const bool row_ok = lane < 30;
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const double s2 = p.qinfo[0];
        // weights of the level pairs in ISSUE order (2,3), (4,5), (0,1): 2^24, 2^8, 2^40 (times 2^-2f)
        const double s_mid = s2 * 16777216.0, s_lo = s2 * 256.0, s_hi = s2 * 1099511627776.0;
        // The int32 level sums enter the FP64 recombination still carrying the bias of the integer -> double trick
        // (xm = MAGIC + v, exact); the bias is taken out inside the FMAs instead of by a subtraction per value:
        //   acc = fma(xm_mid, s_mid, -MAGIC (s_mid + s_lo))  = v_mid s_mid - MAGIC s_lo            (exact: integers < 2^53 in units of s_lo)
        //   acc = fma(xm_lo, s_lo, acc)                       = v_mid s_mid + v_lo s_lo             (exact: < 2^51 units of s_lo)
        //   acc = acc + fma(xm_hi, s_hi, -MAGIC s_hi)         = the whole element, rounded ONCE     (the fma is exact: v_hi s_hi)
        // (scales are powers of two 2^16 apart, |v| < 2^34.)  4 FP64 operations per value instead of 6, on the pipe the
        // epilogue is throttled by, and a correctly rounded sum.
        constexpr double MAGIC = PW == 8 ? 4503601774854144.0 : 6755399441055744.0;   // 2^52 + 2^31 | 2^52 + 2^51
        const double kfold_mid = -MAGIC * (s_mid + s_lo), kfold_hi = -MAGIC * s_hi;
        double* wg_scratch = scratch + wg * (OZ_FI * PITCH);
        // per-warpgroup tables of the current block row, refreshed when the row changes: element offset of every row's
        // start in the output slab (shifted so that column j indexes it directly) and G of the row's frame
        long long* row_off_s = reinterpret_cast<long long*>(smem + 2 * Cfg::B_STAGE + Cfg::SCRATCH + 256) + wg * (2 * OZ_FI);
        double* row_g_s = reinterpret_cast<double*>(row_off_s + OZ_FI);
        // Pairs of this warpgroup: 40 rows x PW columns, solved in warp passes of 32 lanes.  Warp ww of the warpgroup takes
        // the passes q, q + 4, ... with q = (ww - wg) mod 4: the warps with one pass more sit on different schedulers in
        // different warpgroups (warp ww of every warpgroup shares scheduler ww).
        //   PW = 8: a pass = four rows 8 apart x 8 columns (lane = 8 sub + jl): 40 (row, sub) slots, ten passes -> 3,3,2,2.
        //   PW = 4: a pass = two half-warps of four rows 4 apart x 4 columns (lane = 16 h + 4 sub + jl): twelve half-warps
        //           (rows 32-39 fill theirs half), six passes -> 2,2,1,1.
        // With the row pitch = 9 (mod 16) doubles the 16 lanes of a shared-memory wavefront then read 16 different 8-byte
        // bank units (PW = 8: except where a pass wraps from row 32+ to row 1+), and the writers -- lane 3 g + a holds
        // the row of frame g, coordinate a -- write conflict-free; every row's PW results are one run of the output.
        const int q_pass = ((wt >> 5) - wg) & 3;
        // Thread-invariant shared-memory addresses, made opaque so that the compiler keeps them in registers instead of
        // re-deriving them from the thread index in every block (~75 instructions per block otherwise)
        uint32_t my_row_s = smem_u32(wg_scratch + (quarter * 10 + g) * PITCH + 3 * a);
        uint32_t rd_s[NR], tab_s[NR];
        int il_r[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int pass = q_pass + 4 * r;
            int il, jl;
            bool valid;
            if constexpr (PW == 8) {
                const int slot = 4 * pass + (lane >> 3);          // 0..39 (valid) .. 47
                valid = slot < OZ_FI;
                il = (slot / 5) + 8 * (slot % 5);
                jl = lane & 7;
            } else {
                const int hw = 2 * pass + (lane >> 4);            // half-warp 0..11 (valid) .. 15
                il = 16 * (hw >> 2) + (hw & 3) + 4 * ((lane >> 2) & 3);
                valid = hw < 12 && il < OZ_FI;
                jl = lane & 3;
            }
            if (!valid) il = 0;
            il_r[r] = valid ? il : -1;
            rd_s[r] = smem_u32(wg_scratch + il * PITCH + jl * 9);
            tab_s[r] = smem_u32(row_off_s + il);
            asm volatile("" : "+r"(rd_s[r]), "+r"(tab_s[r]), "+r"(il_r[r]));
        }
        asm volatile("" : "+r"(my_row_s));
        const int jl_mine = PW == 8 ? (lane & 7) : (lane & 3);
        const int last_row = min(p.row_end, p.M - 1);
        TsWalker<FJ> w;
        w.init(ti_begin, ntj, blk_start, blk_count);
        int tb = 0; uint32_t tphase = 0;
        int cur_ti = -1; uint32_t afree_phase = 0;
        // register pairs that hold the high word of the magic constant (PW = 8: the 32-bit form of the conversion)
        double magic_pair[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) magic_pair[k] = p.qinfo[4 + k];   // 2^52 + 2^31, opaque to the compiler (see quant_info_kernel)
        for (; w.valid(); w.next()) {
            const bool new_row = w.ti != cur_ti;
            if (wg == 0 && new_row) {
                // this warp's 32 A rows of the new block row: global -> registers -> TMEM
                if (cur_ti >= 0) { mbar_wait(a_free, afree_phase); afree_phase ^= 1; tc_fence_after(); }
                const uint8_t* src = p.packT + ((size_t)w.ti * OZ_M + quarter * 32 + lane) * (size_t)(OZ_SLICES * kp);
#pragma unroll
                for (int s = 0; s < OZ_SLICES; ++s) {
                    for (int kc = 0; kc < nch; ++kc) {   // 32 atoms = 32 bytes = 8 TMEM columns
                        uint32_t v[8];
                        ldg256_evict_first(src + s * kp + 32 * kc, v);
                        tmem_st8(tlane + acol + (uint32_t)(s * (kp >> 2) + kc * 8), v);
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
            }
            cur_ti = w.ti;
            // G of this thread's column frame: fetched now, used after the accumulate phase
            const int j = w.tj * FJ + wg * PW + jl_mine;
            const double gj = p.Gq[j];
            // Level pairs -> FP64 covariance rows.  A pair's TMEM buffer goes back to the MMA warp the moment its columns are
            // in registers, BEFORE they are converted: with two buffers the third pair of a block can only be issued into the
            // buffer of the first, and every cycle between that release and the point where the epilogue needs the third pair
            // is one it does not wait for the tensor pipe.
            double acc[CW];
            auto convert = [&](int pr, int n, uint32_t e_raw, uint32_t o_raw) {
                const int e = (int)e_raw, o = (int)o_raw;
                double xm;
                if constexpr (PW == 8) {
                    // kp <= 128: |256 e + o| < 2^31.  2^52 + 2^31 + v: only the low word of a resident register pair changes
                    const int lo = (e * 256 + o) ^ (int)0x80000000;
                    asm("{\n.reg .b32 l, h;\nmov.b64 {l, h}, %0;\nmov.b64 %0, {%1, h};\n}" : "+d"(magic_pair[n & 7]) : "r"(lo));
                    xm = magic_pair[n & 7];
                } else {
                    xm = level_pair_biased(e, o);
                }
                if (pr == 0) acc[n] = fma(xm, s_mid, kfold_mid);
                else if (pr == 1) acc[n] = fma(xm, s_lo, acc[n]);
                else acc[n] += fma(xm, s_hi, kfold_hi);
            };
#pragma unroll
            for (int pr = 0; pr < OZ_NPAIR; ++pr) {
                mbar_wait(&tfull[tb], tphase);
                tc_fence_after();
                const uint32_t t0 = tlane + (uint32_t)(tb * 2 * N + wg * CW);
                uint32_t ev[2 * PW], od[2 * PW], ev2[PW], od2[PW];   // CW columns = 2 PW + PW
                if constexpr (PW == 8) {
                    tmem_ld16(t0, ev); tmem_ld16(t0 + N, od); tmem_ld8(t0 + 16, ev2); tmem_ld8(t0 + N + 16, od2);
                } else {
                    tmem_ld8(t0, ev); tmem_ld8(t0 + N, od); tmem_ld4(t0 + 8, ev2); tmem_ld4(t0 + N + 8, od2);
                }
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[tb]);
                if (++tb == nbuf) { tb = 0; tphase ^= 1; }
#pragma unroll
                for (int n = 0; n < 2 * PW; ++n) convert(pr, n, ev[n], od[n]);
#pragma unroll
                for (int n = 0; n < PW; ++n) convert(pr, 2 * PW + n, ev2[n], od2[n]);
            }
            // ---- transposition through shared memory: rows (frame i, coordinate a) in, whole 3x3 of one pair out ----
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");   // everyone has read the previous block (and its tables)
            if (row_ok) {
#pragma unroll
                for (int jj = 0; jj < PW; ++jj) {
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(my_row_s + (uint32_t)(jj * 72)), "d"(acc[3 * jj + 0]) : "memory");
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(my_row_s + (uint32_t)(jj * 72 + 8)), "d"(acc[3 * jj + 1]) : "memory");
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(my_row_s + (uint32_t)(jj * 72 + 16)), "d"(acc[3 * jj + 2]) : "memory");
                }
            }
            if (new_row && wt < OZ_FI) {
                const int i = w.ti * OZ_FI + wt;
                row_off_s[wt] = condensed_row_base(p.M, i) - p.idx_base - i - 1;
                row_g_s[wt] = p.Gq[i];
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
            double qa[NR], qb[NR], qd[NR], hg[NR], lam[NR];
            long long out_off[NR];   // offset of the pair's entry in the output slab
            bool pmine[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int i = w.ti * OZ_FI + il_r[r];
                const bool live = il_r[r] >= 0 && i < j && j < p.M;
                pmine[r] = live && i >= p.row_begin && i < last_row;
                double s9[9];
                if (live) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(s9[k]) : "r"(rd_s[r] + (uint32_t)(8 * k)) : "memory");
                    double gi;
                    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(out_off[r]) : "r"(tab_s[r]) : "memory");
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(gi) : "r"(tab_s[r] + (uint32_t)(8 * OZ_FI)) : "memory");
                    hg[r] = gi + gj;
                } else {
                    // lower triangle / padding: a well-conditioned stand-in (identity) so that no lane drags its warp
                    // through the solver's slow path; the result is never stored
#pragma unroll
                    for (int k = 0; k < 9; ++k) s9[k] = (k % 4 == 0) ? 1.0 : 0.0;
                    hg[r] = 3.0;
                    out_off[r] = 0;
                }
                quartic_invariants(s9, qa[r], qb[r], qd[r]);
            }
            unsigned okmask;
            if (q_pass < 2) {
                okmask = largest_root_fast<NR>(qa, qb, qd, hg, lam);
            } else {  // the last pass does not exist for these warps
                double a2[NR - 1], b2[NR - 1], d2[NR - 1], h2[NR - 1], l2[NR - 1];
#pragma unroll
                for (int r = 0; r < NR - 1; ++r) { a2[r] = qa[r]; b2[r] = qb[r]; d2[r] = qd[r]; h2[r] = hg[r]; }
                okmask = largest_root_fast<NR - 1>(a2, b2, d2, h2, l2);
#pragma unroll
                for (int r = 0; r < NR - 1; ++r) lam[r] = l2[r];
                lam[NR - 1] = 0.0;
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double eh = hg[r] - lam[r];
                const double v = (eh > 0.0 ? eh : 0.0) * p.inv_k2;
                const double y = rsqrt_approx_f64(v + 1e-300);        // sqrt(v) = v / sqrt(v), one Newton step
                double rm = v * y;
                rm = fma(fma(-rm, rm, v), 0.5 * y, rm);
                const bool ok = (okmask >> r) & 1u;
                if ((!ok || eh < p.flag_rel * hg[r]) && pmine[r]) {
                    const unsigned int t = atomicAdd(p.fix.count, 1u);
                    if (t < p.fix.capacity) p.fix.pairs[t] = make_int2(w.ti * OZ_FI + il_r[r], j);
                    else rm = -1.0;
                }
                if (pmine[r]) __stcs(p.out + out_off[r] + j, rm);   // streaming store: the output must not evict the operand tiles from L2
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, OZ_TMEM_COLS);
}

// ---- quantise + slice + lay out ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxabs_kernel(const double* __restrict__ x, long long n, unsigned long long* out) {
    double m = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    bool bad = false;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const double v = fabs(x[k]);
        bad |= !(v <= 1.7976931348623157e308);   // NaN or infinity: no fixed-point image
        m = fmax(m, v);
    }
    if (bad) m = __longlong_as_double(0x7ff0000000000000LL);   // +inf: quant_exponent() turns it into "do not quantise"
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));  // m >= 0: bit order = value order
}

// f such that |X| = |rint(x 2^f)| <= 2^30 for every coordinate
__device__ __forceinline__ int quant_exponent(unsigned long long maxbits) {
    const double m = __longlong_as_double((long long)maxbits);
    if (!(m > 0.0)) return 30;
    if (!(m <= 1.7976931348623157e308)) return -1000;   // non-finite coordinates: the engine takes the FP64 kernel
    int e;
    frexp(m, &e);  // m = frac * 2^e, frac in [0.5, 1)  =>  m < 2^e
    return 30 - e;
}

__global__ void quant_info_kernel(const unsigned long long* maxbits, double* qinfo) {
    const int f = quant_exponent(*maxbits);
    qinfo[0] = ldexp(1.0, -2 * f);
    qinfo[1] = (double)f;
    qinfo[2] = __longlong_as_double((long long)*maxbits);
    // 2^52 + 2^31: the epilogue of gram_i8_ts_kernel<8> reads these eight copies at run time so that the compiler sees eight
    // unrelated values and keeps each in its own register pair, of which a conversion only replaces the low word
    for (int k = 0; k < 8; ++k) qinfo[4 + k] = 4503601774854144.0;
}

// Which tcgen05 kernel serves K kept atoms, and the operand layouts it reads.
//   mode 4 / 2: gram_i8_ts_kernel<8> / <4> (A in TMEM; column tiles of 32 / 16 frames), whole contraction in one stage:
//               packT [ti][128 rows][slice][kp B], packB [tj][slice][kp / 16][N / 8][8 rows][16 B]
//   mode 0    : gram_i8_kernel (A and B through shared memory, K-blocks of 128 atoms): packA / packB as in OzParams
struct OzPlan {
    int mode, kp, fj, n;   // kp = padded atoms per stage row, fj = frames per column tile, n = 3 fj
    int groups;            // 16-atom groups per row
};
static OzPlan oz_plan(int K) {
    OzPlan pl;
    const int kp = (K + 31) / 32 * 32;
    const char* force = getenv("CLUSTTRAJ_B200_OZ_MODE");   // development: 2 = prefer the N = 48 shape, 0 = shared-memory kernel
    const int fm = force ? atoi(force) : 4;
    if (kp <= TsCfg<8>::KP_MAX && fm == 4) pl = {4, kp, TsCfg<8>::FJ, TsCfg<8>::N, kp / 16};
    else if (kp <= TsCfg<4>::KP_MAX && fm != 0) pl = {2, kp, TsCfg<4>::FJ, TsCfg<4>::N, kp / 16};
    else pl = {0, OZ_KB, OZ_FJ, OZ_N, (K + OZ_KB - 1) / OZ_KB * 8};
    return pl;
}

// one thread = 16 consecutive atoms of one coordinate of one frame -> one 16-byte run per slice in each layout
__global__ void __launch_bounds__(256) quant_pack_kernel(const double* __restrict__ centred, int M, int K, OzPlan pl,
                                                         const unsigned long long* __restrict__ maxbits,
                                                         uint8_t* __restrict__ packA, uint8_t* __restrict__ packB,
                                                         uint8_t* __restrict__ packT) {
    const int groups = pl.groups;
    const long long total = (long long)M * 3 * groups;
    const double scale = ldexp(1.0, quant_exponent(*maxbits));
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        const int g16 = (int)(o % groups);
        const long long rest = o / groups;
        const int c = (int)(rest % 3);
        const int fr = (int)(rest / 3);
        uint32_t w[4][4];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int q = 0; q < 4; ++q) w[s][q] = 0u;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int k = g16 * 16 + t;
            int X = 0;
            if (k < K) X = (int)rint(centred[((size_t)fr * K + k) * 3 + c] * scale);
            // balanced base-256 digits, most significant first in slice 0
            const int d3 = (int)(signed char)(X & 0xff); X = (X - d3) >> 8;
            const int d2 = (int)(signed char)(X & 0xff); X = (X - d2) >> 8;
            const int d1 = (int)(signed char)(X & 0xff); X = (X - d1) >> 8;
            const int d0 = X;
            const int sh = (t & 3) * 8;
            w[0][t >> 2] |= (uint32_t)(d0 & 0xff) << sh;
            w[1][t >> 2] |= (uint32_t)(d1 & 0xff) << sh;
            w[2][t >> 2] |= (uint32_t)(d2 & 0xff) << sh;
            w[3][t >> 2] |= (uint32_t)(d3 & 0xff) << sh;
        }
        const int ti = fr / OZ_FI, fi = fr - ti * OZ_FI;
        const int rowA = (fi / 10) * 32 + (fi % 10) * 3 + c;
        const int tj = fr / pl.fj, fj = fr - tj * pl.fj;
        const int rowB = fj * 3 + c;
        if (pl.mode) {
            // A: one contiguous [slice][kp] run per TMEM lane; B: canonical no-swizzle K-major core matrices, all of K
            uint8_t* dt = packT + ((size_t)ti * OZ_M + rowA) * (size_t)(OZ_SLICES * pl.kp) + g16 * 16;
            uint8_t* db = packB + (size_t)tj * (size_t)(OZ_SLICES * pl.n * pl.kp) + (size_t)g16 * (pl.n * 16) + (rowB >> 3) * 128 +
                          (rowB & 7) * 16;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                *reinterpret_cast<uint4*>(dt + s * pl.kp) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
                *reinterpret_cast<uint4*>(db + (size_t)s * (pl.n * pl.kp)) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
            }
        } else {
            const int nkb = groups >> 3, kb = g16 >> 3, kc16 = g16 & 7;
            uint8_t* da = packA + ((size_t)ti * nkb + kb) * OZ_A_BYTES + (size_t)kc16 * (OZ_M * 16) + (rowA >> 3) * 128 + (rowA & 7) * 16;
            uint8_t* db = packB + ((size_t)tj * nkb + kb) * OZ_B_BYTES + (size_t)kc16 * (OZ_N * 16) + (rowB >> 3) * 128 + (rowB & 7) * 16;
            static_assert(3 * OZ_FJ <= OZ_N && OZ_N % 16 == 0, "column tile: 3 rows per frame, MMA N a multiple of 16");
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                *reinterpret_cast<uint4*>(da + s * OZ_A_SLICE) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
                *reinterpret_cast<uint4*>(db + s * OZ_B_SLICE) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
            }
        }
    }
}

// half sum of squares of the quantised coordinates (what the integer Gram product is consistent with)
__global__ void __launch_bounds__(128) quant_g_kernel(const double* __restrict__ centred, int M, int K,
                                                      const unsigned long long* __restrict__ maxbits, double* __restrict__ Gq) {
    const int fr = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (fr >= M) return;
    const int f = quant_exponent(*maxbits);
    const double scale = ldexp(1.0, f);
    const double* src = centred + (size_t)fr * K * 3;
    double g = 0.0;
    for (int k = lane; k < 3 * K; k += 32) {
        const double X = rint(src[k] * scale);
        g = fma(X, X, g);
    }
    g = warp_sum(g);
    if (lane == 0) Gq[fr] = 0.5 * g * ldexp(1.0, -2 * f);
}

// operand buffer sizes (packA doubles as the size of packT: both hold ceil(M / 40) row tiles of 128 rows)
size_t oz_packA_bytes(int M, int K) {
    const OzPlan pl = oz_plan(K);
    const size_t tiles = (size_t)((M + OZ_FI - 1) / OZ_FI);
    return pl.mode ? tiles * OZ_M * OZ_SLICES * pl.kp : tiles * (size_t)(pl.groups >> 3) * OZ_A_BYTES;
}
size_t oz_packB_bytes(int M, int K) {
    const OzPlan pl = oz_plan(K);
    const size_t tiles = (size_t)((M + pl.fj - 1) / pl.fj);
    return pl.mode ? tiles * OZ_SLICES * pl.n * pl.kp : tiles * (size_t)(pl.groups >> 3) * OZ_B_BYTES;
}
size_t oz_g_entries(int M) { return (size_t)((M + OZ_FI - 1) / OZ_FI) * OZ_FI + 128; }
int oz_row_align() { return OZ_FI; }

cudaError_t launch_quant_pack(const double* centred, int M, int K, uint8_t* packA, uint8_t* packB, uint8_t* packT, double* Gq,
                              double* qinfo, unsigned long long* maxbits, int sm_count, cudaStream_t stream) {
    const OzPlan pl = oz_plan(K);
    cudaError_t e;
    if ((e = cudaMemsetAsync(maxbits, 0, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(pl.mode ? packT : packA, 0, oz_packA_bytes(M, K), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(packB, 0, oz_packB_bytes(M, K), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(Gq, 0, sizeof(double) * oz_g_entries(M), stream)) != cudaSuccess) return e;
    const long long n = (long long)M * K * 3;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
    maxabs_kernel<<<(unsigned)blocks, 256, 0, stream>>>(centred, n, maxbits);
    quant_info_kernel<<<1, 1, 0, stream>>>(maxbits, qinfo);
    const long long total = (long long)M * 3 * pl.groups;
    blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    quant_pack_kernel<<<(unsigned)blocks, 256, 0, stream>>>(centred, M, K, pl, maxbits, packA, packB, packT);
    quant_g_kernel<<<(M + 3) / 4, 128, 0, stream>>>(centred, M, K, maxbits, Gq);
    return cudaGetLastError();
}

template <int PW>
static cudaError_t launch_ts(OzParams p, int sm_count, size_t b_bytes, cudaStream_t stream) {
    using Cfg = TsCfg<PW>;
    const int ti_begin = p.row_begin / OZ_FI;
    const int last_row = p.row_end < p.M - 1 ? p.row_end : p.M - 1;
    const int ti_end = (last_row + OZ_FI - 1) / OZ_FI;
    const long long ntj = (p.M + Cfg::FJ - 1) / Cfg::FJ;
    long long blocks = 0;
    for (int ti = ti_begin; ti < ti_end; ++ti) blocks += ntj - TsWalker<Cfg::FJ>::first_tj(ti);
    if (blocks <= 0) return cudaSuccess;
    p.total_blocks = blocks;
    const long long grid = blocks < sm_count ? blocks : sm_count;
    cudaError_t e = cudaFuncSetAttribute(gram_i8_ts_kernel<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    // The B operand is what every CTA re-reads (one 16- or 32-frame tile per block, the whole array once per row of blocks)
    // while 8 bytes per pair stream out through the same L2: mark it persisting for this launch (an access policy window over
    // the array; a hit ratio below 1 when it exceeds the set-aside part of the L2), everything else streaming.
    static const bool l2_window = !(getenv("CLUSTTRAJ_B200_L2_WINDOW") && atoi(getenv("CLUSTTRAJ_B200_L2_WINDOW")) == 0);
    bool window_set = false;
    if (l2_window && b_bytes > 0) {
        int dev = 0, max_persist = 0, max_window = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        if (max_persist > 0 && max_window > 0) {
            size_t lim = 0;
            if (cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize) != cudaSuccess || lim < (size_t)max_persist)
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
            cudaStreamAttrValue av{};
            const size_t bytes = b_bytes < (size_t)max_window ? b_bytes : (size_t)max_window;
            av.accessPolicyWindow.base_ptr = const_cast<uint8_t*>(p.packB);
            av.accessPolicyWindow.num_bytes = bytes;
            av.accessPolicyWindow.hitRatio = bytes <= (size_t)max_persist ? 1.0f : (float)((double)max_persist / (double)bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            window_set = cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
        }
        cudaGetLastError();   // the hint is optional: never fail the launch over it
    }
    gram_i8_ts_kernel<PW><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, stream>>>(p);
    e = cudaGetLastError();
    if (window_set) {
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    return e;
}

cudaError_t launch_gram_i8(OzParams p, int K, int sm_count, cudaStream_t stream) {
    const OzPlan pl = oz_plan(K);
    if (pl.mode) {
        p.nkb = 1;
        p.last_chunks = pl.kp / 32;
        const size_t b_bytes = oz_packB_bytes(p.M, K);
        return pl.mode == 4 ? launch_ts<8>(p, sm_count, b_bytes, stream) : launch_ts<4>(p, sm_count, b_bytes, stream);
    }
    p.nkb = (K + OZ_KB - 1) / OZ_KB;
    const int tail = K - (p.nkb - 1) * OZ_KB;
    p.last_chunks = (tail + 31) / 32;
    const int ti_begin = p.row_begin / OZ_FI;
    const int last_row = p.row_end < p.M - 1 ? p.row_end : p.M - 1;
    const int ti_end = (last_row + OZ_FI - 1) / OZ_FI;
    const long long ntj = (p.M + OZ_FJ - 1) / OZ_FJ;
    long long blocks = 0;
    for (int ti = ti_begin; ti < ti_end; ++ti) blocks += ntj - OzWalker::first_tj(ti);
    if (blocks <= 0) return cudaSuccess;
    p.total_blocks = blocks;
    const long long grid = blocks < sm_count ? blocks : sm_count;
    cudaError_t e = cudaFuncSetAttribute(gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    gram_i8_kernel<<<(unsigned)grid, OZ_THREADS, OZ_SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

int oz_kernel_mode(int K) { return oz_plan(K).mode; }

}  // namespace ct
