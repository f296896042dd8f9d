// K2 — all-pairs cross-covariance (Gram contraction) on the FP64 tensor pipe with the RMSD solved in
// the epilogue, for the modes of clusttraj/distmat.py whose value is a pure Kabsch RMSD (no reorder):
// plain, -n, and -ns without -e (SURVEY.md §8a dispatch rows 1-2).  Replaces, for every pair, the two
// rmsd.kabsch calls + np.dot + rmsd.rmsd of distmat.py:227-228,275 (or :178-181,275 with -ns).
//
// Shape of the work: the 3M x K matrix of centred coordinates times its transpose, restricted to
// the upper triangle of frame pairs.  A CTA owns a TI x TJ block of frame pairs and runs a persistent
// loop over such blocks.  Operand rows are ordered (coordinate, frame) inside an MMA tile, so the nine
// m8n8k4 accumulator tiles of one 8x8 block of frame pairs leave the complete 3x3 covariance of two
// pairs in each lane's own registers: the epilogue needs no shuffles.
//
// Data movement: cp.async.bulk (TMA, 1-D) brings one KC-atom stage of each 32-frame tile as a single
// contiguous 6 KiB run into a STAGES-deep shared-memory ring guarded by mbarriers; loads run STAGES-1
// stages ahead across block boundaries, so the epilogue of one block overlaps the first loads of the
// next.  Warps never meet at a block-wide barrier in the steady state: each warp releases a ring slot
// on an "empty" mbarrier and thread 0 refills it PREFETCH stages ahead.  Every lane stores its own
// pairs straight into the condensed vector (the row segments of neighbouring lanes/warps merge in L2).
#include "common.cuh"
#include "solve3x3.cuh"

namespace ct {

template <int WI_, int WJ_, int MA_, int MB_, int STAGES_, int KQS_>
struct GramCfg {
    static constexpr int WI = WI_, WJ = WJ_, MA = MA_, MB = MB_, STAGES = STAGES_;
    static constexpr int KQS = KQS_;                               // k4 groups (4 atoms each) per pipeline stage
    static constexpr int TILE_STAGE_DOUBLES = KQS * KQ_DOUBLES;    // one tile, one stage
    static constexpr int TILE_STAGE_BYTES = TILE_STAGE_DOUBLES * 8;
    static constexpr int TI = 8 * MA * WI;  // frames per block, row side
    static constexpr int TJ = 8 * MB * WJ;  // frames per block, column side
    static constexpr int NTI = TI / FT, NTJ = TJ / FT;
    static constexpr int THREADS = 32 * WI * WJ;
    static constexpr int STAGE_DOUBLES = (NTI + NTJ) * TILE_STAGE_DOUBLES;
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int PREFETCH = STAGES - 2;  // a slot is refilled two consumer iterations after its use
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16;
    static_assert(STAGES >= 3, "need at least three ring slots");
    static_assert(TI % FT == 0 && TJ % FT == 0, "block sides must be whole packed tiles");
    static_assert(TJ % TI == 0, "column side must be a multiple of the row side");
};

// Walks the upper-triangle blocks of rows [ti_begin, ti_end) in row-major order with a fixed stride.
template <class Cfg>
struct BlockWalker {
    int ti, tj, ti_end, ntj;
    __device__ __forceinline__ static int first_tj(int ti) { return (ti * Cfg::TI) / Cfg::TJ; }
    __device__ __forceinline__ void settle() {
        while (ti < ti_end && tj >= ntj) {
            ++ti;
            tj = tj - ntj + first_tj(ti);
        }
    }
    __device__ __forceinline__ void init(int ti_begin, int ti_end_, int ntj_, int offset) {
        ti = ti_begin; ti_end = ti_end_; ntj = ntj_;
        tj = first_tj(ti) + offset;
        settle();
    }
    __device__ __forceinline__ bool valid() const { return ti < ti_end; }
    __device__ __forceinline__ void advance(int stride) { tj += stride; settle(); }
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, (Cfg::THREADS <= 256) ? 2 : 1) gram_rmsd_kernel(const GramParams p) {
    constexpr int MA = Cfg::MA, MB = Cfg::MB, STAGES = Cfg::STAGES, TI = Cfg::TI, TJ = Cfg::TJ, KQS = Cfg::KQS;
    constexpr int TILE_STAGE_DOUBLES = Cfg::TILE_STAGE_DOUBLES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_buf = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(stage_buf + STAGES * Cfg::STAGE_DOUBLES);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = warp % Cfg::WI, wj = warp / Cfg::WI;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], Cfg::THREADS / 32); }
        fence_barrier_init();
    }
    __syncthreads();

    const int ti_begin = p.row_begin / TI;
    const int last_row = min(p.row_end, p.M - 1);  // row M-1 owns no entries
    const int ti_end = (last_row + TI - 1) / TI;
    const int ntj = (p.M + TJ - 1) / TJ;
    const int n_stage = (p.kq_total + KQS - 1) / KQS;
    const int tail_kq = p.kq_total - (n_stage - 1) * KQS;  // k4 groups in the last stage of a block (1..KQS)
    const size_t tile_doubles = (size_t)p.kq_total * 3 * FT * 4;  // one 32-frame tile, all atoms

    // ---- producer state (thread 0 only): runs STAGES-1 pipeline stages ahead of the consumers -----
    BlockWalker<Cfg> lw;
    lw.init(ti_begin, ti_end, ntj, blockIdx.x);
    int lk = 0;                // next stage of the producer's block
    unsigned int lg = 0;       // stages issued so far
    auto issue_load = [&]() {
        if (!lw.valid()) return;
        const int slot = lg % STAGES;
        if (lg >= STAGES) mbar_wait(&empty[slot], ((lg / STAGES) - 1) & 1);  // every warp released the slot
        double* dst = stage_buf + slot * Cfg::STAGE_DOUBLES;
        // the last stage of a block may be partly filled (atoms are padded to 4, not to a whole stage)
        const uint32_t bytes = (lk == n_stage - 1 ? tail_kq : KQS) * (KQ_DOUBLES * 8);
        mbar_expect_tx(&full[slot], bytes * (Cfg::NTI + Cfg::NTJ));
        const size_t koff = (size_t)lk * TILE_STAGE_DOUBLES;
#pragma unroll
        for (int t = 0; t < Cfg::NTI; ++t)
            bulk_g2s(dst + t * TILE_STAGE_DOUBLES, p.packed + (size_t)(lw.ti * Cfg::NTI + t) * tile_doubles + koff,
                     bytes, &full[slot]);
#pragma unroll
        for (int t = 0; t < Cfg::NTJ; ++t)
            bulk_g2s(dst + (Cfg::NTI + t) * TILE_STAGE_DOUBLES,
                     p.packed + (size_t)(lw.tj * Cfg::NTJ + t) * tile_doubles + koff, bytes, &full[slot]);
        ++lg;
        if (++lk == n_stage) { lk = 0; lw.advance(gridDim.x); }
    };
    if (tid == 0) {
        for (int s = 0; s < Cfg::PREFETCH; ++s) issue_load();
    }

    // ---- consumers ------------------------------------------------------------------------------------
    // fragment offsets inside a stage (doubles): tile*TILE_STAGE_DOUBLES + (kq*3+c)*128 + (8-frame block)*32 + lane
    int offA[MA], offB[MB];
#pragma unroll
    for (int ma = 0; ma < MA; ++ma) {
        const int fb = wi * MA + ma;
        offA[ma] = (fb >> 2) * TILE_STAGE_DOUBLES + (fb & 3) * 32 + lane;
    }
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        const int fb = wj * MB + mb;
        offB[mb] = (Cfg::NTI + (fb >> 2)) * TILE_STAGE_DOUBLES + (fb & 3) * 32 + lane;
    }

    BlockWalker<Cfg> cw;
    cw.init(ti_begin, ti_end, ntj, blockIdx.x);
    unsigned int g = 0;  // stages consumed so far
    bool ready = false;  // outcome of the early probe of the next stage's "full" barrier
    for (; cw.valid(); cw.advance(gridDim.x)) {
        double acc[MA][MB][3][3][2];
#pragma unroll
        for (int ma = 0; ma < MA; ++ma)
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                    for (int cb = 0; cb < 3; ++cb) { acc[ma][mb][ca][cb][0] = 0.0; acc[ma][mb][ca][cb][1] = 0.0; }

#pragma unroll 1
        for (int k = 0; k < n_stage; ++k, ++g) {
            const int slot = g % STAGES;
            if (tid == 0) issue_load();  // stage g + PREFETCH
            if (!ready) mbar_wait(&full[slot], (g / STAGES) & 1);
            // probe the next stage now; the answer is consumed after this stage's DMMAs (hides the
            // ~90-cycle mbarrier round trip)
            ready = mbar_test(&full[(g + 1) % STAGES], ((g + 1) / STAGES) & 1);
            const double* sb = stage_buf + slot * Cfg::STAGE_DOUBLES;
            const int nkq = (k == n_stage - 1) ? tail_kq : KQS;
#pragma unroll
            for (int kq = 0; kq < KQS; ++kq) {
                if (kq >= nkq) break;
                double fa[MA][3], fb[MB][3];
#pragma unroll
                for (int ma = 0; ma < MA; ++ma)
#pragma unroll
                    for (int c = 0; c < 3; ++c) fa[ma][c] = sb[offA[ma] + (kq * 3 + c) * 128];
#pragma unroll
                for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                    for (int c = 0; c < 3; ++c) fb[mb][c] = sb[offB[mb] + (kq * 3 + c) * 128];
#pragma unroll
                for (int ma = 0; ma < MA; ++ma)
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                        for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                            for (int cb = 0; cb < 3; ++cb)
                                dmma884(acc[ma][mb][ca][cb][0], acc[ma][mb][ca][cb][1], fa[ma][ca], fb[mb][cb]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);  // this warp no longer reads the slot
        }

        // ---- epilogue: each lane owns the 3x3 covariance of NP = 2*MA*MB frame pairs; all of them are
        // solved in lockstep (see largest_root_lockstep) ---------------------------------------------------
        const int i0 = cw.ti * TI, j0 = cw.tj * TJ;
        constexpr int NP = 2 * MA * MB;
        double qa[NP], qb[NP], qd[NP], hg[NP], lam[NP];
#pragma unroll
        for (int ma = 0; ma < MA; ++ma) {
            const double gi = p.G[i0 + (wi * MA + ma) * 8 + (lane >> 2)];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const double2 gj = *reinterpret_cast<const double2*>(p.G + j0 + (wj * MB + mb) * 8 + 2 * (lane & 3));
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int q = (ma * MB + mb) * 2 + e;
                    double s[9];
#pragma unroll
                    for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                        for (int cb = 0; cb < 3; ++cb) s[ca * 3 + cb] = acc[ma][mb][ca][cb][e];
                    quartic_invariants(s, qa[q], qb[q], qd[q]);
                    hg[q] = 0.5 * (gi + (e ? gj.y : gj.x));
                }
            }
        }
        const unsigned okmask = largest_root_lockstep<NP>(qa, qb, qd, hg, lam);
#pragma unroll
        for (int ma = 0; ma < MA; ++ma) {
            const int il = (wi * MA + ma) * 8 + (lane >> 2);
            const int i = i0 + il;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                const int jl = (wj * MB + mb) * 8 + 2 * (lane & 3);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int q = (ma * MB + mb) * 2 + e;
                    const double E = 2.0 * (hg[q] - lam[q]);
                    // sqrt(E/K): FP32 rsqrt seed + one FP64 Newton step (relative error ~1e-13)
                    const double v = fmax(E, 0.0) * p.inv_k;
                    const float y = rsqrtf(fmaxf((float)v, 1e-37f));
                    double r = v * (double)y;
                    r = fma(0.5 * (double)y, fma(-r, r, v), r);
                    const int j = j0 + jl + e;
                    const bool ok = (okmask >> q) & 1u;
                    if ((!ok || E < 2.0 * p.flag_rel * hg[q]) && i < j && j < p.M && i >= p.row_begin && i < p.row_end) {
                        const unsigned int t = atomicAdd(p.fix.count, 1u);
                        if (t < p.fix.capacity) p.fix.pairs[t] = make_int2(i, j);
                        else r = -1.0;  // list full: mark the entry, the fix-up kernel scans for it
                    }
                    if (i < j && j < p.M && i >= p.row_begin && i < last_row)
                        p.out[condensed_row_base(p.M, i) - p.idx_base + (j - i - 1)] = r;
                }
            }
        }
    }
}

// ---- host-side launcher ---------------------------------------------------------------------------
template <class Cfg>
static cudaError_t launch_gram_cfg(const GramParams& p, int sm_count, cudaStream_t stream) {
    // per-device attribute: set on every launch (several engines / devices may live in one process)
    cudaError_t e = cudaFuncSetAttribute(gram_rmsd_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gram_rmsd_kernel<Cfg>, Cfg::THREADS,
                                                                  Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    // number of blocks to do, to avoid launching idle CTAs on tiny problems
    const int ti_begin = p.row_begin / Cfg::TI;
    const int last_row = p.row_end < p.M - 1 ? p.row_end : p.M - 1;
    const int ti_end = (last_row + Cfg::TI - 1) / Cfg::TI;
    const long long ntj = (p.M + Cfg::TJ - 1) / Cfg::TJ;
    long long blocks = 0;
    for (int ti = ti_begin; ti < ti_end; ++ti) blocks += ntj - (long long)(ti * Cfg::TI) / Cfg::TJ;
    if (blocks <= 0) return cudaSuccess;
    long long grid = (long long)sm_count * per_sm;  // persistent: one wave of resident CTAs
    if (grid > blocks) grid = blocks;
    gram_rmsd_kernel<Cfg><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

using GramDefault = GramCfg<4, 2, 1, 2, 4, 4>;  // 32 x 32 frame pairs, 8 warps, 2 CTAs / SM, 16 atoms / stage
using GramWide = GramCfg<4, 4, 1, 2, 4, 4>;     // 32 x 64 frame pairs, 16 warps, 1 CTA / SM
using GramK8 = GramCfg<4, 2, 1, 2, 6, 2>;       // as default with 8 atoms / stage, deeper ring
using GramK24 = GramCfg<4, 2, 1, 2, 3, 6>;      // 24 atoms / stage

cudaError_t launch_gram(const GramParams& p, int variant, int sm_count, cudaStream_t stream) {
    if (variant == 1) return launch_gram_cfg<GramWide>(p, sm_count, stream);
    if (variant == 2) return launch_gram_cfg<GramK8>(p, sm_count, stream);
    if (variant == 3) return launch_gram_cfg<GramK24>(p, sm_count, stream);
    return launch_gram_cfg<GramDefault>(p, sm_count, stream);
}

int gram_row_align(int variant) { return variant == 1 ? GramWide::TI : GramDefault::TI; }

}  // namespace ct
