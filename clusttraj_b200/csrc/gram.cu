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
//
// Warp specialisation (WS = 1, the default): the CTA has as many epilogue warps as MMA warps.  An MMA
// warp hands the 36 accumulators of each lane to its partner through shared memory (mbarrier full /
// empty pair per warp) and goes straight back to the DMMA loop of the next block; the partner solves
// the 3x3 problems.  The epilogue's FP64 instructions still share the pipe with the DMMAs (about 7 % of
// it at K = 100), but no DMMA-issuing warp ever waits behind them.
#include "common.cuh"
#include "solve3x3.cuh"

namespace ct {

template <int WI_, int WJ_, int MA_, int MB_, int STAGES_, int KQS_, int WS_, int PREFETCH_ = STAGES_ - 2>
struct GramCfg {
    static constexpr int WI = WI_, WJ = WJ_, MA = MA_, MB = MB_, STAGES = STAGES_;
    static constexpr int KQS = KQS_;                               // k4 groups (4 atoms each) per pipeline stage
    static constexpr bool WS = WS_ != 0;                           // separate epilogue warps
    static constexpr int TILE_STAGE_DOUBLES = KQS * KQ_DOUBLES;    // one tile, one stage
    static constexpr int TILE_STAGE_BYTES = TILE_STAGE_DOUBLES * 8;
    static constexpr int TI = 8 * MA * WI;  // frames per block, row side
    static constexpr int TJ = 8 * MB * WJ;  // frames per block, column side
    static constexpr int NTI = TI / FT, NTJ = TJ / FT;
    static constexpr int MMA_WARPS = WI * WJ;
    // WS: MMA warps + as many epilogue warps + one producer warp, rounded up to whole warpgroups (setmaxnreg)
    static constexpr int THREADS = WS ? 128 * ((2 * MMA_WARPS + 1 + 3) / 4) : 32 * MMA_WARPS;
    static constexpr int CTAS_PER_SM = (THREADS <= 256) ? 2 : 1;
    static constexpr int STAGE_DOUBLES = (NTI + NTJ) * TILE_STAGE_DOUBLES;
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int NACC = 18 * MA * MB;                      // accumulator doubles per lane
    static constexpr int HANDOFF_DOUBLES = WS ? MMA_WARPS * NACC * 32 : 0;
    static constexpr int PREFETCH = PREFETCH_;   // loads run this many stages ahead; a slot is refilled
                                                 // STAGES - PREFETCH consumer iterations after its use
    static constexpr int SMEM_BYTES = (STAGES * STAGE_DOUBLES + HANDOFF_DOUBLES) * 8 + (2 * STAGES + 2 * MMA_WARPS) * 8 + 32;
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

// Fused epilogue: each lane owns the 3x3 covariance of NP = 2*MA*MB frame pairs (acc[ma][mb][ca][cb][e]);
// all of them are solved in lockstep (largest_root_f64) and stored into the condensed vector.
template <class Cfg>
__device__ __forceinline__ void epilogue_store(const double (&acc)[Cfg::MA][Cfg::MB][3][3][2], const GramParams& p, int i0,
                                               int j0, int wi, int wj, int lane, int last_row) {
    constexpr int MA = Cfg::MA, MB = Cfg::MB, NP = 2 * MA * MB;
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
                hg[q] = gi + (e ? gj.y : gj.x);  // G holds half sums of squares
            }
        }
    }
    const unsigned okmask = largest_root_f64<NP>(qa, qb, qd, hg, lam);
#pragma unroll
    for (int ma = 0; ma < MA; ++ma) {
        const int i = i0 + (wi * MA + ma) * 8 + (lane >> 2);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
            const int jl = (wj * MB + mb) * 8 + 2 * (lane & 3);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = (ma * MB + mb) * 2 + e;
                // E/2 = (G_i + G_j)/2 - lambda;  rmsd = sqrt(max(E, 0) / K): FP32 rsqrt seed + one FP64 Newton step
                const double eh = hg[q] - lam[q];
                const double v = fmax(eh, 0.0) * p.inv_k2;
                const float y = rsqrtf(fmaxf((float)v, 1e-37f));
                double r = v * (double)y;
                r = fma(fma(-r, r, v), (double)(0.5f * y), r);
                const int j = j0 + jl + e;
                const bool ok = (okmask >> q) & 1u;
                const bool mine = i < j && j < p.M && i >= p.row_begin && i < last_row;
                if ((!ok || eh < p.flag_rel * hg[q]) && mine) {
                    const unsigned int t = atomicAdd(p.fix.count, 1u);
                    if (t < p.fix.capacity) p.fix.pairs[t] = make_int2(i, j);
                    else r = -1.0;  // list full: mark the entry, the fix-up kernel scans for it
                }
                if (mine) __stcs(p.out + (condensed_row_base(p.M, i) - p.idx_base + (j - i - 1)), r);   // streaming store
            }
        }
    }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::CTAS_PER_SM) gram_rmsd_kernel(const GramParams p) {
    constexpr int MA = Cfg::MA, MB = Cfg::MB, STAGES = Cfg::STAGES, TI = Cfg::TI, TJ = Cfg::TJ, KQS = Cfg::KQS;
    constexpr int TILE_STAGE_DOUBLES = Cfg::TILE_STAGE_DOUBLES, NW = Cfg::MMA_WARPS, NACC = Cfg::NACC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_buf = reinterpret_cast<double*>(smem_raw);
    double* handoff = stage_buf + STAGES * Cfg::STAGE_DOUBLES;                 // [MMA warp][NACC][32 lanes]
    uint64_t* full = reinterpret_cast<uint64_t*>(handoff + Cfg::HANDOFF_DOUBLES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;   // per MMA warp: accumulators are in the hand-off buffer
    uint64_t* acc_empty = acc_full + NW;   // per MMA warp: the partner has read them

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_mma = warp < NW;
    const int w = is_mma ? warp : warp - NW;
    const int wi = w % Cfg::WI, wj = w / Cfg::WI;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
#pragma unroll
        for (int s = 0; s < NW; ++s) { mbar_init(&acc_full[s], 32); mbar_init(&acc_empty[s], 32); }
        fence_barrier_init();
    }
    __syncthreads();

    const int ti_begin = p.row_begin / TI;
    const int last_row = min(p.row_end, p.M - 1);  // row M-1 owns no entries
    const int ti_end = (last_row + TI - 1) / TI;
    const int ntj = (p.M + TJ - 1) / TJ;

    if (Cfg::WS && !is_mma) {
        // ---- epilogue warps: take each block's accumulators from the partner MMA warp ----------------------
        BlockWalker<Cfg> ew;
        ew.init(ti_begin, ti_end, ntj, blockIdx.x);
        const double* src = handoff + (size_t)w * NACC * 32 + lane;
        for (unsigned int t = 0; ew.valid(); ew.advance(gridDim.x), ++t) {
            double acc[MA][MB][3][3][2];
            mbar_wait(&acc_full[w], t & 1);
#pragma unroll
            for (int ma = 0; ma < MA; ++ma)
#pragma unroll
                for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                    for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                        for (int cb = 0; cb < 3; ++cb)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                acc[ma][mb][ca][cb][e] = src[((((ma * MB + mb) * 3 + ca) * 3 + cb) * 2 + e) * 32];
            mbar_arrive(&acc_empty[w]);  // the values are in registers: the MMA warp may overwrite the buffer
            epilogue_store<Cfg>(acc, p, ew.ti * TI, ew.tj * TJ, wi, wj, lane, last_row);
        }
        return;
    }

    const int n_stage = (p.kq_total + KQS - 1) / KQS;
    const int tail_kq = p.kq_total - (n_stage - 1) * KQS;  // k4 groups in the last stage of a block (1..KQS)
    const size_t tile_doubles = (size_t)p.kq_total * 3 * FT * 4;  // one 32-frame tile, all atoms

    // ---- producer state (thread 0 only): runs PREFETCH pipeline stages ahead of the consumers ---------
    BlockWalker<Cfg> lw;
    lw.init(ti_begin, ti_end, ntj, blockIdx.x);
    int lk = 0;                // next stage of the producer's block
    unsigned int lg = 0;       // stages issued so far
    auto issue_load = [&]() {
        if (!lw.valid()) return;
        const int slot = lg % STAGES;
        if (lg >= STAGES) mbar_wait(&empty[slot], ((lg / STAGES) - 1) & 1);  // every MMA warp released the slot
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

    // ---- MMA warps ------------------------------------------------------------------------------------------
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
    unsigned int g = 0;     // stages consumed so far
    unsigned int nblk = 0;  // blocks finished by this warp

    bool ready = false;  // outcome of the early probe of the next slot's "full" barrier
    for (; cw.valid(); ++nblk) {
        double acc[MA][MB][3][3][2];
#pragma unroll
        for (int ma = 0; ma < MA; ++ma)
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                    for (int cb = 0; cb < 3; ++cb) { acc[ma][mb][ca][cb][0] = 0.0; acc[ma][mb][ca][cb][1] = 0.0; }
        const int blk_ti = cw.ti, blk_tj = cw.tj;
        cw.advance(gridDim.x);
        auto k4_step = [&](const double* sb, int kq) {
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
        };
#pragma unroll 1
        for (int k = 0; k < n_stage; ++k, ++g) {
            const int slot = g % STAGES;
            if (tid == 0) issue_load();  // stage g + PREFETCH
            if (!ready) mbar_wait(&full[slot], (g / STAGES) & 1);
            ready = mbar_test(&full[(g + 1) % STAGES], ((g + 1) / STAGES) & 1);
            const double* sb = stage_buf + slot * Cfg::STAGE_DOUBLES;
            if (k + 1 < n_stage || tail_kq == KQS) {
#pragma unroll
                for (int kq = 0; kq < KQS; ++kq) k4_step(sb, kq);
            } else {
#pragma unroll
                for (int kq = 0; kq < KQS; ++kq) {
                    if (kq >= tail_kq) break;
                    k4_step(sb, kq);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);  // this warp no longer reads the slot
        }

        if (Cfg::WS) {
            // hand the accumulators to the partner epilogue warp and return to the DMMA loop
            if (nblk > 0) mbar_wait(&acc_empty[w], (nblk - 1) & 1);
            double* dst = handoff + (size_t)w * NACC * 32 + lane;
#pragma unroll
            for (int ma = 0; ma < MA; ++ma)
#pragma unroll
                for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                    for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                        for (int cb = 0; cb < 3; ++cb)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                dst[((((ma * MB + mb) * 3 + ca) * 3 + cb) * 2 + e) * 32] = acc[ma][mb][ca][cb][e];
            mbar_arrive(&acc_full[w]);
        } else {
            epilogue_store<Cfg>(acc, p, blk_ti * TI, blk_tj * TJ, wi, wj, lane, last_row);
        }
    }
}

// ---- warp-specialised kernel ----------------------------------------------------------------------------
// Same data flow as above with the roles split over the 16 warps of a CTA (one CTA per SM):
//   warps 0..7   MMA: stream k4 steps out of the ring, nothing else — they do not even know which block
//                they are on; each finished block's accumulators go to a hand-off buffer;
//   warps 8..15  epilogue: one partner per MMA warp, picks up the hand-off buffer and solves the 3x3 problems;
//   warp 16      TMA producer (warps 17..19 only exist to complete its warpgroup and exit at once);
// registers are re-balanced per warpgroup with setmaxnreg: the kernel is launched at 96 registers per thread
// (640 threads = 61440 registers, which is also the pool setmaxnreg re-distributes), the producer group shrinks
// to 24, the MMA groups grow to 120, the epilogue groups to 104 (30720 + 26624 + 3072 <= 61440).
//   producer (one lane): refills ring slots as soon as every MMA warp has released them,
//                so loads run a full ring ahead and never wait for a DMMA-issuing warp.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) gram_rmsd_ws_kernel(const GramParams p) {
    constexpr int MA = Cfg::MA, MB = Cfg::MB, STAGES = Cfg::STAGES, TI = Cfg::TI, TJ = Cfg::TJ, KQS = Cfg::KQS;
    constexpr int TILE_STAGE_DOUBLES = Cfg::TILE_STAGE_DOUBLES, NW = Cfg::MMA_WARPS, NACC = Cfg::NACC;
    constexpr int NEPI = NW;      // epilogue warps; the last warp of the CTA is the producer
    static_assert(Cfg::WS, "warp-specialised configuration expected");
    static_assert(NW % 4 == 0, "roles must fill whole warpgroups");
    static_assert(MA == 1 && (MB == 1 || MB == 2 || MB == 4), "fragment addressing assumes MA = 1 and MB | 4");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_buf = reinterpret_cast<double*>(smem_raw);
    double* handoff = stage_buf + STAGES * Cfg::STAGE_DOUBLES;                 // [MMA warp][NACC][32 lanes]
    uint64_t* full = reinterpret_cast<uint64_t*>(handoff + Cfg::HANDOFF_DOUBLES);
    uint64_t* empty = full + STAGES;
    unsigned int* acc_written = reinterpret_cast<unsigned int*>(empty + STAGES);  // per MMA warp: blocks handed off
    unsigned int* acc_taken = acc_written + NW;                                   // per MMA warp: blocks picked up

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti_begin = p.row_begin / TI;
    const int last_row = min(p.row_end, p.M - 1);
    const int ti_end = (last_row + TI - 1) / TI;
    const int ntj = (p.M + TJ - 1) / TJ;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
#pragma unroll
        for (int s = 0; s < NW; ++s) { acc_written[s] = 0; acc_taken[s] = 0; }
        fence_barrier_init();
    }
    __syncthreads();
    const int n_stage = (p.kq_total + KQS - 1) / KQS;
    const int tail_kq = p.kq_total - (n_stage - 1) * KQS;

    if (warp >= 2 * NW) {
        // ---- TMA producer ----------------------------------------------------------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp != 2 * NW || lane != 0) return;
        const size_t tile_doubles = (size_t)p.kq_total * KQ_DOUBLES;
        BlockWalker<Cfg> lw;
        unsigned int lg = 0;
        for (lw.init(ti_begin, ti_end, ntj, blockIdx.x); lw.valid(); lw.advance(gridDim.x)) {
            const double* src_i = p.packed + (size_t)(lw.ti * Cfg::NTI) * tile_doubles;
            const double* src_j = p.packed + (size_t)(lw.tj * Cfg::NTJ) * tile_doubles;
            for (int lk = 0; lk < n_stage; ++lk, ++lg) {
                const int slot = lg % STAGES;
                if (lg >= STAGES) mbar_wait(&empty[slot], ((lg / STAGES) - 1) & 1);  // every MMA warp released it
                double* dst = stage_buf + slot * Cfg::STAGE_DOUBLES;
                // the last stage of a block may be partly filled (atoms are padded to 4, not to a whole stage)
                const uint32_t bytes = (lk == n_stage - 1 ? tail_kq : KQS) * (KQ_DOUBLES * 8);
                mbar_expect_tx(&full[slot], bytes * (Cfg::NTI + Cfg::NTJ));
                const size_t koff = (size_t)lk * TILE_STAGE_DOUBLES;
#pragma unroll
                for (int t = 0; t < Cfg::NTI; ++t)
                    bulk_g2s(dst + t * TILE_STAGE_DOUBLES, src_i + t * tile_doubles + koff, bytes, &full[slot]);
#pragma unroll
                for (int t = 0; t < Cfg::NTJ; ++t)
                    bulk_g2s(dst + (Cfg::NTI + t) * TILE_STAGE_DOUBLES, src_j + t * tile_doubles + koff, bytes,
                             &full[slot]);
            }
        }
        return;
    }

    if (warp >= NW) {
        // ---- epilogue warps ----------------------------------------------------------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int e = warp - NW;  // 0 .. NEPI-1
        BlockWalker<Cfg> ew;
        ew.init(ti_begin, ti_end, ntj, blockIdx.x);
        unsigned int blk = 0;     // block the walker points at
        for (unsigned int n = e; ew.valid(); n += NEPI) {
            const unsigned int nb = n / NW;
            const int w = n % NW;
            while (blk < nb && ew.valid()) { ew.advance(gridDim.x); ++blk; }
            if (!ew.valid()) break;
            const int wi = w % Cfg::WI, wj = w / Cfg::WI;
            const double* src = handoff + (size_t)w * NACC * 32 + lane;
            double acc[MA][MB][3][3][2];
            seq_wait_ge(&acc_written[w], nb + 1);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                    for (int cb = 0; cb < 3; ++cb)
#pragma unroll
                        for (int ee = 0; ee < 2; ++ee)
                            acc[0][mb][ca][cb][ee] = src[(((mb * 3 + ca) * 3 + cb) * 2 + ee) * 32];
            __syncwarp();
            if (lane == 0) seq_publish(&acc_taken[w], nb + 1);  // values are in registers: the buffer is free
            epilogue_store<Cfg>(acc, p, ew.ti * TI, ew.tj * TJ, wi, wj, lane, last_row);
        }
        return;
    }

    // ---- MMA warps ---------------------------------------------------------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    const int w = warp, wi = w % Cfg::WI, wj = w / Cfg::WI;
    int nblk_total = 0;
    {
        BlockWalker<Cfg> cnt;
        for (cnt.init(ti_begin, ti_end, ntj, blockIdx.x); cnt.valid(); cnt.advance(gridDim.x)) ++nblk_total;
    }
    // fragment addresses inside a stage: + (kq*3+c)*128 doubles; the second column block is +32 doubles
    const int fragA = (wi >> 2) * TILE_STAGE_DOUBLES + (wi & 3) * 32 + lane;
    const int fragB = (Cfg::NTI + ((wj * MB) >> 2)) * TILE_STAGE_DOUBLES + ((wj * MB) & 3) * 32 + lane;

    unsigned int g = 0;
    bool ready = false;  // outcome of the early probe of the next slot's "full" barrier
    double* hdst = handoff + (size_t)w * NACC * 32 + lane;
#pragma unroll 1
    for (int blk = 0; blk < nblk_total; ++blk) {
        double acc[MB][3][3][2];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                for (int cb = 0; cb < 3; ++cb) { acc[mb][ca][cb][0] = 0.0; acc[mb][ca][cb][1] = 0.0; }
        // one k4 step: 9 fragment loads (one LDS.64 each) and 18 DMMAs.  In a full stage the KQS steps are
        // straight-line code, so the compiler hoists the loads of step s+1 over the DMMAs of step s.
        auto k4_step = [&](const double* sb, int kq) {
            double fa[3], fb[MB][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) fa[c] = sb[fragA + (kq * 3 + c) * 128];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int c = 0; c < 3; ++c) fb[mb][c] = sb[fragB + mb * 32 + (kq * 3 + c) * 128];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                    for (int cb = 0; cb < 3; ++cb) dmma884(acc[mb][ca][cb][0], acc[mb][ca][cb][1], fa[ca], fb[mb][cb]);
        };
#pragma unroll 1
        for (int k = 0; k < n_stage; ++k, ++g) {
            const int slot = g % STAGES;
            if (!ready) mbar_wait(&full[slot], (g / STAGES) & 1);
            // probe the next slot now; the answer is used after this stage's DMMAs (hides the mbarrier round trip)
            ready = mbar_test(&full[(g + 1) % STAGES], ((g + 1) / STAGES) & 1);
            const double* sb = stage_buf + slot * Cfg::STAGE_DOUBLES;
            if (k + 1 < n_stage || tail_kq == KQS) {
#pragma unroll
                for (int kq = 0; kq < KQS; ++kq) k4_step(sb, kq);
            } else {
                // partly filled last stage of a block
#pragma unroll
                for (int kq = 0; kq < KQS; ++kq) {
                    if (kq >= tail_kq) break;
                    k4_step(sb, kq);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        // hand the accumulators to an epilogue warp and go straight to the next block
        if (blk > 0) seq_wait_ge(&acc_taken[w], (unsigned int)blk);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int ca = 0; ca < 3; ++ca)
#pragma unroll
                for (int cb = 0; cb < 3; ++cb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) hdst[(((mb * 3 + ca) * 3 + cb) * 2 + e) * 32] = acc[mb][ca][cb][e];
        __syncwarp();
        if (lane == 0) seq_publish(&acc_written[w], (unsigned int)blk + 1);
    }
}

// ---- host-side launcher ---------------------------------------------------------------------------
template <class Cfg>
static auto kernel_of() {
    if constexpr (Cfg::WS) return &gram_rmsd_ws_kernel<Cfg>;
    else return &gram_rmsd_kernel<Cfg>;
}

template <class Cfg>
static cudaError_t launch_gram_cfg(const GramParams& p, int sm_count, cudaStream_t stream) {
    auto kernel = kernel_of<Cfg>();
    // per-device attribute: set on every launch (several engines / devices may live in one process)
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::THREADS,
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
    kernel<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

using GramUnified = GramCfg<4, 2, 1, 2, 4, 4, 0>;  // 32 x 32 frame pairs, 8 warps do MMA + epilogue, 2 CTAs / SM
using GramWS = GramCfg<4, 2, 1, 2, 4, 4, 1>;       // 8 MMA warps + 8 epilogue warps, 1 CTA / SM
using GramWS6 = GramCfg<4, 2, 1, 2, 4, 6, 1, 2>;   // same, 24 atoms per stage
using GramWSk2 = GramCfg<4, 2, 1, 2, 3, 8, 1, 1>;  // same, 32 atoms per stage, three slots

cudaError_t launch_gram(const GramParams& p, int variant, int sm_count, cudaStream_t stream) {
    if (variant < 0) {
        // measured on B200 (profiles/): with few atoms the epilogue dominates and the unified kernel (every warp
        // does both jobs, 2 CTAs/SM) wins; from ~200 atoms the longer stage amortises the per-stage bookkeeping
        const int atoms = p.kq_total * 4;
        variant = atoms <= 64 ? 1 : (atoms < 200 ? 0 : 2);
    }
    if (variant == 1) return launch_gram_cfg<GramUnified>(p, sm_count, stream);
    if (variant == 2) return launch_gram_cfg<GramWS6>(p, sm_count, stream);
    if (variant == 3) return launch_gram_cfg<GramWSk2>(p, sm_count, stream);
    return launch_gram_cfg<GramWS>(p, sm_count, stream);
}

int gram_row_align(int variant) { (void)variant; return GramWS::TI; }

}  // namespace ct
