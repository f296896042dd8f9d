// K3 / K4 — the general per-pair path of clusttraj/distmat.py:133-275, one warp per frame pair:
// stepwise solute superposition (-ns), per-element Hungarian reordering (-e, rmsd.reorder_hungarian =
// scipy cdist("euclidean") + linear_sum_assignment per element, bound at io.py:559-561 and called at
// distmat.py:192,243), --final-kabsch, -ws weights and -eex exclusions.  The control flow per pair
// follows the reference's dispatch (SURVEY.md §8a mode table); the arithmetic is new:
//   * rotations come from the quaternion eigenproblem (solve3x3.cuh), not an SVD;
//   * the assignment is a warp-parallel shortest-augmenting-path solver (the algorithm family of
//     SciPy's linear_sum_assignment, Crouse 2016): lanes own columns, the frontier minimum is a
//     warp-shuffle argmin, duals and path state live in shared memory, and EUCLIDEAN (not squared)
//     costs are evaluated on the fly in FP64 from coordinates staged in shared memory — no cost matrix
//     is ever materialised.
// With continuous coordinates the optimal assignment is unique, so the permutation equals SciPy's.
#include "common.cuh"
#include "solve3x3.cuh"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/clusttraj_b200.h"

namespace ct {

struct PairPlan {
    int K;            // kept atoms
    int natoms;       // kept solute atoms (0 without -ns)
    int has_ns, do_reorder, solute_phase, final_direct, weighted, final_kabsch;
    int by_distance;  // reorder_mode 2: rank matching by distance from the origin instead of the assignment problem
    double w_solute, w_solvent;  // per-atom weights (distmat.py:259-262)
    // element blocks: phase 0 = solute reorder (distmat.py:184-223), phase 1 = main reorder (:231-256)
    int nblocks[2];
    const int* block_off[2];  // [nblocks+1] offsets into idx
    const int* idx[2];        // kept-atom positions, ascending inside a block
    long long idx_stride[2];  // 0: the same index lists for every frame; > 0: one list per frame (per-frame element lists,
                              // distmat.py:131-141: the reference builds the element blocks from each frame's own atoms)
    int nmax;                 // largest block
    int credit0;              // initial credit of the two-row collision shortcut (0 switches it off)
};

struct PairPlanHost {
    PairPlan plan{};
    std::vector<int> view[2], off[2];   // host copies: the atoms each phase reorders (ascending) and the block boundaries
    std::vector<int> block_z[2];        // element of every block
    int* d_frame_idx = nullptr;   // per-frame index lists (pair_plan_frame_elements)
    int* d_ints = nullptr;        // all index arrays in one allocation
    double* d_work = nullptr;     // per-resident-warp working copy of the moving frame [K][3]
    int* d_perm = nullptr;        // per-resident-warp composite permutation [K]
    size_t work_warps = 0;
    int smem_per_warp = 0;
};

static constexpr int PAIR_WARPS = 4;
#ifndef CT_NN_UNROLL
#define CT_NN_UNROLL 1    // columns per iteration of the nearest-column pass (2 and 4 measured slower: 8.6 -> 8.0 -> 7.1e6 pairs/s)
#endif

__host__ __device__ inline int pair_smem_per_warp(int nmax) {
    // Qe, Pe (3 doubles per row/column), u, v, shortest (doubles), pred, row4col, col4row, old perm (ints),
    // scanned (bytes)
    int b = nmax * (24 + 24 + 24 + 16) + ((nmax + 7) / 8) * 8;
    return (b + 15) / 16 * 16;
}

// sqrt for the Euclidean cost: rsqrt.approx.f64 seed (one XU operation, ~22 bits; the FP32 rsqrt it replaces needed
// two F2F conversions around it, and the quarter-rate XU pipe was as loaded as the FP64 pipe in the frontier scans)
// + two FP64 Newton steps on the residual: within 1 ulp of the correctly rounded value, ~3x cheaper than the IEEE
// sequence and free of long dependent chains.  (One step, 6e-14 relative, would do for the assignment -- but measured
// SLOWER: with the shorter chain ptxas interleaves the eight columns of a scan less: cfg4 shape 8.3 -> 7.6e6 pairs/s.)
__device__ __forceinline__ double cost_sqrt(double d2) {
    // + 1e-300: keeps d2 = 0 away from rsqrt's infinity and changes no other input by a bit (one DADD where fmax's NaN
    // handling costs eight instructions per entry)
    const double yd = rsqrt_approx_f64(d2 + 1e-300), h = 0.5 * yd;
    double s = d2 * yd;
    s = fma(fma(-s, s, d2), h, s);
    s = fma(fma(-s, s, d2), h, s);
    return s;
}

// squared distance in one fixed operation order: the nearest-column pass and the frontier scans must see the same bits
__device__ __forceinline__ double dist2(double dx, double dy, double dz) { return fma(dx, dx, fma(dy, dy, dz * dz)); }

// ---- warp-parallel shortest augmenting path, column state in registers --------------------------------
// For blocks of up to 32*C atoms.  Lane l owns columns l, l+32, ...: their coordinates, dual v and
// tentative path cost stay in registers for the whole solve, so a frontier scan is C independent
// distance evaluations per lane (ILP) + one warp argmin.  Row duals, predecessors and the matching live
// in shared memory (touched once per scan / per augmentation).
//
// Nearest-column pass.  Column duals only ever decrease from 0 and a row's dual is 0 until the row is inserted, so
// the first frontier scan of row i returns min_j (c_ij - v_j) >= min_j c_ij = m_i, with equality at the nearest
// column a_i whenever v[a_i] is still 0.  When a_i is also unmatched that scan ends the augmentation at once:
// row i takes a_i, u_i = m_i, no other dual moves.  So all m_i, a_i are computed up front, lanes owning ROWS, over
// SQUARED distances (7 FP64 operations per entry instead of the ~18 of a scan entry, no cross-lane reduction, one
// square root per row), and the sequential insertion loop runs the Dijkstra scans only for the rows whose nearest
// column is taken or has a moved dual (~10 % of the rows on solvent-shell data).  The result is the one the plain
// insertion order produces, step for step.
// SHORTCUT: with the two-row collision shortcut below.  A separate instantiation because the mere presence of that code
// changes how ptxas schedules the frontier scans (-e without -ns, where the shortcut almost never applies and ~3800
// scans per pair remain: 5.8e5 pairs/s without the code, 4.4e5 with it); the host picks per run (PairPlan::credit0).
template <int C, bool SHORTCUT>
__device__ int lsap_warp_reg(int n, const double* Qe, const double* Pe, double* u, int* pred, int* r4c, int* c4r,
                             double* m_row, int* a_row, unsigned char* moved, int lane, int credit) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    {
        double qx[C], qy[C], qz[C], bd[C];
        int bj[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int i = lane + 32 * c;
            const bool in = i < n;
            qx[c] = in ? Qe[3 * i] : 0.0; qy[c] = in ? Qe[3 * i + 1] : 0.0; qz[c] = in ? Qe[3 * i + 2] : 0.0;
            bd[c] = INF; bj[c] = -1;
        }
        constexpr int NN_UNROLL = CT_NN_UNROLL;
#pragma unroll NN_UNROLL
        for (int j = 0; j < n; ++j) {
            const double x = Pe[3 * j], y = Pe[3 * j + 1], z = Pe[3 * j + 2];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double d2 = dist2(qx[c] - x, qy[c] - y, qz[c] - z);
                const bool lt = d2 < bd[c];
                bd[c] = lt ? d2 : bd[c];
                bj[c] = lt ? j : bj[c];
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int i = lane + 32 * c;
            if (i < n) { m_row[i] = cost_sqrt(bd[c]); a_row[i] = bj[c]; moved[i] = 0; }
        }
    }
    double px[C], py[C], pz[C], v[C], sh[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = lane + 32 * c;
        const bool in = j < n;
        px[c] = in ? Pe[3 * j] : 0.0; py[c] = in ? Pe[3 * j + 1] : 0.0; pz[c] = in ? Pe[3 * j + 2] : 0.0;
        v[c] = 0.0;
    }
    for (int j = lane; j < n; j += 32) { u[j] = 0.0; r4c[j] = -1; c4r[j] = -1; }
    __syncwarp();
    unsigned valid = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) if (lane + 32 * c < n) valid |= 1u << c;
    int steps = 0;   // frontier scans + direct assignments (each stands for the first scan of its row)
    int cur = 0;
    int b_base = 0, b_a = -1;       // the current batch of direct assignments: first row, this lane's nearest column,
    bool b_ok = false, b_valid = false;   // whether this lane's row may take it, whether the evaluation still holds
    // credit: two-row collision shortcut: +1 per success (up to 8), -2 per miss, not attempted at <= 0
#pragma unroll 1
    while (cur < n) {
        // direct assignments, 32 consecutive rows at a time: a row takes its nearest column when that column is
        // unmatched, its dual has not moved and no earlier row of the batch wants it too; the rows before the first
        // failing one are committed together, the failing one goes through the shortcut or the scans below.
        // The evaluation of a batch (b_base, b_a, b_ok) stays valid across collision shortcuts: a shortcut matches exactly
        // one new column jw and moves no dual of an unmatched column, so the rows after the failing one only have to drop
        // out if they wanted jw; after a frontier search the batch is evaluated afresh.
        {
            if (!b_valid) {
                b_base = cur;
                const int row = b_base + lane;
                b_a = -1;
                b_ok = false;
                if (row < n) {
                    b_a = a_row[row];
                    b_ok = b_a >= 0 && r4c[b_a] < 0 && !moved[b_a];
                }
                // "no earlier row of the batch wants the same column": the lowest row claims the column in shared memory
                // (pred[] is scratch between two searches).  MATCH.ANY did the same in ~300 cycles, 3.4 % of the kernel.
                if (b_a >= 0) pred[b_a] = 0x7fffffff;
                __syncwarp();
                if (b_a >= 0) atomicMin(&pred[b_a], row);
                __syncwarp();
                b_ok = b_ok && pred[b_a] == row;
                b_valid = SHORTCUT;
            }
            const int off = cur - b_base;   // rows of the batch already dealt with (< 32)
            const unsigned fail = ~__ballot_sync(0xffffffffu, b_ok) & (0xffffffffu << off);
            const int first = fail ? __ffs(fail) - 1 : 32;
            if (lane >= off && lane < first) { r4c[b_a] = b_base + lane; c4r[b_base + lane] = b_a; u[b_base + lane] = m_row[b_base + lane]; }
            __syncwarp();
            steps += first - off;
            cur = b_base + first;
            if (first == 32 || cur >= n) { b_valid = false; continue; }
        }
        // Two-row collision, resolved without frontier scans.  Row `cur` lost its nearest column a to an earlier row i1 and a's
        // dual has not moved.  The scans would then go: scan 1 (from cur) pops a at m_cur; scan 2 (from i1, offset
        // m_cur - u[i1]) leaves every other column j at min(c(cur,j), c(i1,j) + m_cur - u[i1]) - v_j, whose minimum lies at the
        // nearest other column of cur (b2) or of i1 (b1), provided their duals are still 0 (v <= 0 always, so a moved column
        // can only look worse than its distance).  If that column is unmatched the augmentation ends there: one fused pass over
        // SQUARED distances for both rows (12 FP64 operations per column instead of two scans' 2 x 17 + bookkeeping, two square
        // roots in all) gives the same assignment, duals and tie decisions, bit for bit.  Anything else takes the scans.
        {
            const int a = a_row[cur];
            const int i1 = a >= 0 ? r4c[a] : -1;
            bool fast = SHORTCUT && i1 >= 0 && !moved[a] && credit > 0;   // (credit: see below)
            if (fast) {
                // nearest column other than a, for row cur (t = 0) and for row i1 (t = 1).  Unrolled: the two chains of three warp
                // reductions overlap (rolled, to save code: 9.74 against 9.96e6 pairs/s, -rs 11.4 against 12.2e6)
                unsigned long long kmin0 = 0, kmin1 = 0;
                int b2 = 0x7fffffff, b1 = 0x7fffffff;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int row = t ? i1 : cur;
                    const double qx = Qe[3 * row], qy = Qe[3 * row + 1], qz = Qe[3 * row + 2];
                    double dm = INF;
                    int jm = 0x7fffffff;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = lane + 32 * c;
                        const double e = dist2(qx - px[c], qy - py[c], qz - pz[c]);
                        const bool take = ((valid >> c) & 1u) && j != a && e < dm;
                        dm = take ? e : dm; jm = take ? j : jm;
                    }
                    // warp argmin (d >= 0: the bit pattern orders like the value; lowest column on a tie)
                    const unsigned long long k = (unsigned long long)__double_as_longlong(dm);
                    const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32));
                    const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32) == hi ? (unsigned)k : 0xffffffffu);
                    const unsigned long long km = ((unsigned long long)hi << 32) | lo;
                    const int bc = (int)__reduce_min_sync(0xffffffffu, k == km ? (unsigned)jm : 0x7fffffffu);
                    if (t == 0) { kmin0 = km; b2 = bc; } else { kmin1 = km; b1 = bc; }
                }
                fast = b2 != 0x7fffffff && b1 != 0x7fffffff && !moved[b2] && !moved[b1];
                if (fast) {
                    const double m_cur = m_row[cur];
                    // what the scans would compute: r = cost + (minVal - u[i]) - v[j], with v = 0 here
                    const double A = cost_sqrt(__longlong_as_double((long long)kmin0)) + (0.0 - u[cur]);
                    const double B = cost_sqrt(__longlong_as_double((long long)kmin1)) + (m_cur - u[i1]);
                    // column b (both rows' candidate when b1 == b2: the later scan only replaces on a strictly lower value)
                    const bool via_i1 = (b1 == b2) ? (B < A) : (B < A || (B == A && b1 < b2));
                    const int jw = via_i1 ? b1 : b2;
                    const double mu = via_i1 ? B : A;
                    fast = r4c[jw] < 0 && A == A && B == B;
                    if (fast) {
                        const double delta = mu - m_cur;                     // minVal - sh[a]
                        if (delta != 0.0) {
                            // branch-free: an `if ((a >> 5) == c)` inside the owner lane turns v[] into a local-memory array
                            // behind a jump table (LDL / STL: 8 % of the kernel's stall samples)
#pragma unroll
                            for (int c = 0; c < C; ++c) v[c] -= (a == lane + 32 * c) ? delta : 0.0;
                            if (lane == 0) { moved[a] = 1; u[i1] += delta; }
                        }
                        if (lane == 0) {
                            u[cur] += mu;
                            if (via_i1) { r4c[jw] = i1; c4r[i1] = jw; r4c[a] = cur; c4r[cur] = a; }
                            else { r4c[jw] = cur; c4r[cur] = jw; }
                        }
                        __syncwarp();
                        steps += 2;
                        ++cur;
                        credit = credit < 8 ? credit + 1 : 8;
                        if (b_a == jw) b_ok = false;                  // the one column that is no longer free
                        if (cur - b_base >= 32) b_valid = false;
                        continue;
                    }
                }
                // The fused pass was for nothing.  Where the frames are poorly superposed (no -ns: the whole system is
                // rotated at once) most collisions need long augmenting paths: after a few misses in a row the attempt is
                // skipped for the rest of the block.
                credit -= 2;
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) sh[c] = INF;
        unsigned open = valid;  // owned columns not yet scanned
        int i = cur, sink = -1;
        double minVal = 0.0;
#pragma unroll 1
        while (sink < 0) {
            const double qx = Qe[3 * i], qy = Qe[3 * i + 1], qz = Qe[3 * i + 2];
            const double mu = minVal - u[i];
            double best = INF;
            int bestj = 0x7fffffff;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double r = cost_sqrt(dist2(qx - px[c], qy - py[c], qz - pz[c])) + mu - v[c];
                const bool live = (open >> c) & 1u;
                const bool upd = live && r < sh[c];
                sh[c] = upd ? r : sh[c];
                if (upd) pred[lane + 32 * c] = i;
                const bool better = live && sh[c] < best;
                best = better ? sh[c] : best;
                bestj = better ? lane + 32 * c : bestj;
            }
            // warp argmin with three REDUX passes over an order-preserving 64-bit key (high word, low
            // word, then the column index: the lowest column wins an exact tie)
            const unsigned long long bits = (unsigned long long)__double_as_longlong(best);
            const unsigned long long key = bits ^ ((bits >> 63) ? ~0ULL : 0x8000000000000000ULL);
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
            const unsigned mj = __reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? (unsigned)bestj : 0x7fffffffu);
            ++steps;
            if (mj == 0x7fffffffu) return -steps;  // no finite cost left (NaN coordinates)
            const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
            minVal = __longlong_as_double((long long)(mkey ^ ((mkey >> 63) ? 0x8000000000000000ULL : ~0ULL)));
            const int bj = (int)mj;
            if ((bj & 31) == lane) open &= ~(1u << (bj >> 5));
            __syncwarp();  // pred[] writes of this scan are visible before a possible augmentation
            const int owner = r4c[bj];
            if (owner < 0) sink = bj; else i = owner;
        }
        // dual update through the scanned columns (their rows are exactly the scanned rows)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (((valid & ~open) >> c) & 1u) {
                const double delta = minVal - sh[c];
                if (delta != 0.0) {
                    v[c] -= delta;
                    moved[lane + 32 * c] = 1;
                    const int r = r4c[lane + 32 * c];
                    if (r >= 0) u[r] += delta;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            u[cur] += minVal;
            int j = sink;
            while (true) {
                const int r = pred[j];
                r4c[j] = r;
                const int pj = c4r[r];
                c4r[r] = j;
                j = pj;
                if (r == cur) break;
            }
        }
        __syncwarp();
        ++cur;
        b_valid = false;   // duals and matches of many columns may have changed
    }
    return steps;
}

template <int C>
__device__ __forceinline__ int lsap_reg_dispatch(int credit0, int n, const double* Qe, const double* Pe, double* u, int* pred, int* r4c,
                                                 int* c4r, double* m_row, int* a_row, unsigned char* moved, int lane) {
    if (credit0 > 0) return lsap_warp_reg<C, true>(n, Qe, Pe, u, pred, r4c, c4r, m_row, a_row, moved, lane, credit0);
    return lsap_warp_reg<C, false>(n, Qe, Pe, u, pred, r4c, c4r, m_row, a_row, moved, lane, 0);
}

// ---- warp-parallel shortest augmenting path ------------------------------------------------------------
// rows = atoms of the reference frame (Qe), columns = atoms of the moving frame (Pe); on return
// c4r[r] is the column matched to row r.  Returns the number of frontier scans.
__device__ int lsap_warp(int n, const double* Qe, const double* Pe, double* u, double* v, double* sh, int* pred,
                         int* r4c, int* c4r, unsigned char* scanned, int lane) {
    for (int j = lane; j < n; j += 32) { u[j] = 0.0; v[j] = 0.0; r4c[j] = -1; c4r[j] = -1; }
    __syncwarp();
    int steps = 0;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    for (int cur = 0; cur < n; ++cur) {
        for (int j = lane; j < n; j += 32) { sh[j] = INF; scanned[j] = 0; }
        __syncwarp();
        int i = cur, sink = -1;
        double minVal = 0.0;
        while (sink < 0) {
            const double qx = Qe[3 * i], qy = Qe[3 * i + 1], qz = Qe[3 * i + 2], ui = u[i];
            double best = INF;
            int bestj = 0x7fffffff, bestfree = 0;
            for (int j = lane; j < n; j += 32) {
                if (scanned[j]) continue;
                const double dx = qx - Pe[3 * j], dy = qy - Pe[3 * j + 1], dz = qz - Pe[3 * j + 2];
                const double cost = cost_sqrt(dx * dx + dy * dy + dz * dz);
                const double r = minVal + cost - ui - v[j];
                double s = sh[j];
                if (r < s) { s = r; sh[j] = r; pred[j] = i; }
                const int fr = r4c[j] < 0;
                if (s < best || (s == best && fr > bestfree)) { best = s; bestj = j; bestfree = fr; }
            }
            // warp argmin: lower value, then unassigned column, then lower index
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
                const int of = __shfl_xor_sync(0xffffffffu, bestfree, o);
                const bool take = ob < best || (ob == best && (of > bestfree || (of == bestfree && oj < bestj)));
                if (take) { best = ob; bestj = oj; bestfree = of; }
            }
            ++steps;
            if (bestj == 0x7fffffff) return -steps;  // no finite cost left (NaN coordinates)
            minVal = best;
            if (lane == 0) scanned[bestj] = 1;
            __syncwarp();
            const int owner = r4c[bestj];
            if (owner < 0) sink = bestj; else i = owner;
        }
        // dual update through the scanned columns (their rows are exactly the scanned rows)
        for (int j = lane; j < n; j += 32) {
            if (scanned[j]) {
                const double delta = minVal - sh[j];
                v[j] -= delta;
                const int r = r4c[j];
                if (r >= 0) u[r] += delta;
            }
        }
        __syncwarp();
        if (lane == 0) {
            u[cur] += minVal;
            int j = sink;
            while (true) {
                const int r = pred[j];
                r4c[j] = r;
                const int pj = c4r[r];
                c4r[r] = j;
                j = pj;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
    return steps;
}

// One copy of the 4x4 Jacobi instead of one per call site: the per-pair kernel is bound by instruction fetch before
// anything else (ncu: "no instruction" is its largest stall), so code size is a first-order quantity here.
__device__ __noinline__ double kabsch_rotation_jacobi(const double* s, double* U) { return kabsch_rotation(s, U); }
__device__ __noinline__ double kabsch_rotation_shared(const double* s, double* U) {
    double lam;
    if (kabsch_rotation_adjugate(s, U, lam)) return lam;   // Newton root + adjugate null vector: ~20x less work
    return kabsch_rotation_jacobi(s, U);                   // (near-)degenerate largest eigenvalue
}

// rotate the working frame in place: W[k] = W[k] * U (row vector times matrix)
__device__ __noinline__ void warp_rotate(double* W, int K, const double* U, int lane) {
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
        const double x = W[3 * k], y = W[3 * k + 1], z = W[3 * k + 2];
        W[3 * k] = fma(x, U[0], fma(y, U[3], z * U[6]));
        W[3 * k + 1] = fma(x, U[1], fma(y, U[4], z * U[7]));
        W[3 * k + 2] = fma(x, U[2], fma(y, U[5], z * U[8]));
    }
    __syncwarp();
}

__device__ __noinline__ void warp_cov(const double* P, const double* Q, int k1, int lane, double* s) {
#pragma unroll
    for (int t = 0; t < 9; ++t) s[t] = 0.0;
#pragma unroll 1
    for (int k = lane; k < k1; k += 32) {
        const double px = P[3 * k], py = P[3 * k + 1], pz = P[3 * k + 2];
        const double qx = Q[3 * k], qy = Q[3 * k + 1], qz = Q[3 * k + 2];
        s[0] = fma(px, qx, s[0]); s[1] = fma(px, qy, s[1]); s[2] = fma(px, qz, s[2]);
        s[3] = fma(py, qx, s[3]); s[4] = fma(py, qy, s[4]); s[5] = fma(py, qz, s[5]);
        s[6] = fma(pz, qx, s[6]); s[7] = fma(pz, qy, s[7]); s[8] = fma(pz, qz, s[8]);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s[t] = warp_sum(s[t]);
}

// R <- R * U (row-major 3x3, both act on row vectors from the right)
__device__ __forceinline__ void mat3_rmul(double* R, const double* U) {
    double T[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[3 * r + c] = fma(R[3 * r], U[c], fma(R[3 * r + 1], U[3 + c], R[3 * r + 2] * U[6 + c]));
#pragma unroll
    for (int t = 0; t < 9; ++t) R[t] = T[t];
}

// ---- rmsd.reorder_distance (io.py:562-563): within an element block the k-th closest atom to the origin of the
// reference frame is paired with the k-th closest atom of the moving frame.  Ranks by counting (lanes own atoms, the
// norms of the block are broadcast from shared memory), equal norms ordered by index; on return c4r[r] is the column
// matched to row r.  Norms as numpy computes them: sqrt((x*x + y*y) + z*z), every operation rounded on its own.
__device__ void match_by_norm(int n, const double* Qe, const double* Pe, double* nq, double* np_, int* rank_q, int* inv_p,
                              int* c4r, int lane) {
    for (int t = lane; t < n; t += 32) {
        nq[t] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(Qe[3 * t], Qe[3 * t]), __dmul_rn(Qe[3 * t + 1], Qe[3 * t + 1])),
                                     __dmul_rn(Qe[3 * t + 2], Qe[3 * t + 2])));
        np_[t] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(Pe[3 * t], Pe[3 * t]), __dmul_rn(Pe[3 * t + 1], Pe[3 * t + 1])),
                                      __dmul_rn(Pe[3 * t + 2], Pe[3 * t + 2])));
    }
    __syncwarp();
    for (int t = lane; t < n; t += 32) {
        const double a = nq[t], b = np_[t];
        int ra = 0, rb = 0;
        for (int o = 0; o < n; ++o) {
            const double oa = nq[o], ob = np_[o];
            ra += (oa < a || (oa == a && o < t)) ? 1 : 0;
            rb += (ob < b || (ob == b && o < t)) ? 1 : 0;
        }
        rank_q[t] = ra;
        inv_p[rb] = t;
    }
    __syncwarp();
    for (int t = lane; t < n; t += 32) c4r[t] = inv_p[rank_q[t]];
    __syncwarp();
}

// Per-frame element lists: the moving frame keeps its atoms of element X at other positions than the reference frame.
// Relabel it first -- its t-th atom of the phase's element-sorted list moves to the position of the reference frame's t-th --
// through the second half of the warp's working buffers (block by block in place would overwrite atoms that a later block
// still has to read).  The columns of every block keep their order, so the assignment problems, and with them the reordered
// frame and the composite permutation, are the reference's; from here on both frames use the reference frame's lists.
// (Out of line: cold code inside reorder_phase costs the frame-invariant runs instruction-cache space.)
__device__ __noinline__ void relabel_moving_frame(const PairPlan& pl, int phase, double* W, int* perm, int lane, int frame_q,
                                                  int frame_p) {
    const int nv = (int)pl.idx_stride[phase];
    const int* iq = pl.idx[phase] + (size_t)frame_q * nv;
    const int* ip = pl.idx[phase] + (size_t)frame_p * nv;
    double* T = W + 3 * (size_t)pl.K;
    int* tp = perm + pl.K;
#pragma unroll 1
    for (int t = lane; t < nv; t += 32) {
        const int k = ip[t];
        T[3 * t] = W[3 * k]; T[3 * t + 1] = W[3 * k + 1]; T[3 * t + 2] = W[3 * k + 2];
        tp[t] = perm[k];
    }
    __syncwarp();
#pragma unroll 1
    for (int t = lane; t < nv; t += 32) {
        const int k = iq[t];
        W[3 * k] = T[3 * t]; W[3 * k + 1] = T[3 * t + 1]; W[3 * k + 2] = T[3 * t + 2];
        perm[k] = tp[t];
    }
    __syncwarp();
}

__device__ __noinline__ int reorder_phase(const PairPlan& pl, int phase, double* W, int* perm, const double* Q, unsigned char* sm,
                             int lane, int frame_q, int frame_p) {
    const int nmax = pl.nmax;
    double* Qe = reinterpret_cast<double*>(sm);
    double* Pe = Qe + 3 * nmax;
    double* u = Pe + 3 * nmax;
    double* v = u + nmax;
    double* sh = v + nmax;
    int* pred = reinterpret_cast<int*>(sh + nmax);
    int* r4c = pred + nmax;
    int* c4r = r4c + nmax;
    int* oldp = c4r + nmax;
    unsigned char* scanned = reinterpret_cast<unsigned char*>(oldp + nmax);
    int steps = 0;
    if (pl.idx_stride[phase]) relabel_moving_frame(pl, phase, W, perm, lane, frame_q, frame_p);   // per-frame element lists
    for (int b = 0; b < pl.nblocks[phase]; ++b) {
        const int o0 = pl.block_off[phase][b], n = pl.block_off[phase][b + 1] - o0;
        if (n < 2) continue;
        const int* idx = pl.idx[phase] + (size_t)frame_q * pl.idx_stride[phase] + o0;   // the reference frame's list
        // index -> coordinate is a chain of two dependent global loads: two atoms per lane in flight, stores afterwards
#pragma unroll 1
        for (int t0 = lane; t0 < n; t0 += 64) {
            const int t1 = t0 + 32 < n ? t0 + 32 : t0;
            const int k0 = idx[t0], k1 = idx[t1];
            const double qa = Q[3 * k0], qb = Q[3 * k0 + 1], qc = Q[3 * k0 + 2], qd = Q[3 * k1], qe = Q[3 * k1 + 1], qf = Q[3 * k1 + 2];
            const double pa = W[3 * k0], pb = W[3 * k0 + 1], pc = W[3 * k0 + 2], pd = W[3 * k1], pe = W[3 * k1 + 1], pf = W[3 * k1 + 2];
            const int o0p = perm[k0], o1p = perm[k1];
            Qe[3 * t0] = qa; Qe[3 * t0 + 1] = qb; Qe[3 * t0 + 2] = qc;
            Pe[3 * t0] = pa; Pe[3 * t0 + 1] = pb; Pe[3 * t0 + 2] = pc;
            oldp[t0] = o0p;
            Qe[3 * t1] = qd; Qe[3 * t1 + 1] = qe; Qe[3 * t1 + 2] = qf;
            Pe[3 * t1] = pd; Pe[3 * t1 + 1] = pe; Pe[3 * t1 + 2] = pf;
            oldp[t1] = o1p;
        }
        __syncwarp();
        int st;
        if (pl.by_distance) {
            match_by_norm(n, Qe, Pe, u, v, pred, r4c, c4r, lane);
            st = 0;
        } else
        switch ((n + 31) >> 5) {
            // two register-resident instances (4 and 8 columns per lane) rather than one per block size: code size
            case 1: case 2: case 3: case 4: st = lsap_reg_dispatch<4>(pl.credit0, n, Qe, Pe, u, pred, r4c, c4r, v, reinterpret_cast<int*>(sh), scanned, lane); break;
            case 5: case 6: case 7: case 8:
                st = lsap_reg_dispatch<8>(pl.credit0, n, Qe, Pe, u, pred, r4c, c4r, v, reinterpret_cast<int*>(sh), scanned, lane); break;
            default: st = lsap_warp(n, Qe, Pe, u, v, sh, pred, r4c, c4r, scanned, lane); break;
        }
        steps += st < 0 ? -st : st;
        if (st >= 0) {
            // rmsd.reorder_hungarian: the atom matched to reference row t moves to position idx[t]
#pragma unroll 1
            for (int t = lane; t < n; t += 32) {
                const int c = c4r[t], k = idx[t];
                W[3 * k] = Pe[3 * c]; W[3 * k + 1] = Pe[3 * c + 1]; W[3 * k + 2] = Pe[3 * c + 2];
                perm[k] = oldp[c];
            }
        }
        __syncwarp();
    }
    return steps;
}

// XF: also return the rigid transform clusttraj/io.py:align_mol (582-756) applies to ALL atoms of the moving frame
// (hydrogens and excluded atoms included) when it writes superposed structures: xf[0..8] = R (row-major, acts on row
// vectors: out = (x - centre) R + T), xf[9..11] = T.  It is the product of the rotations of the value pipeline, except
// for the last one, which align_mol only applies for -ns -e --final-kabsch and for -ws --final-kabsch (io.py:733-746).
template <bool XF>
__device__ double pair_value_warp(const PairPlan& pl, const double* __restrict__ Pg, const double* __restrict__ Qg,
                                  double* W, int* perm, unsigned char* sm, int lane, int& steps, double* xf, int frame_q,
                                  int frame_p) {
    const int K = pl.K, ns = pl.natoms;
    double s[9], U[9];
    double Rt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Tt[3] = {0, 0, 0};
    steps = 0;
    // working copy of the moving frame; with -ns the reference subtracts centroid(Q[:natoms]) from P
    // once more (distmat.py:174) although Q is already centred on it
    double qc0 = 0.0, qc1 = 0.0, qc2 = 0.0;
    if (pl.has_ns) {
        for (int k = lane; k < ns; k += 32) { qc0 += Qg[3 * k]; qc1 += Qg[3 * k + 1]; qc2 += Qg[3 * k + 2]; }
        qc0 = warp_sum(qc0) / ns; qc1 = warp_sum(qc1) / ns; qc2 = warp_sum(qc2) / ns;
    }
#pragma unroll 1
    for (int k = lane; k < K; k += 32) {
        W[3 * k] = Pg[3 * k] - qc0; W[3 * k + 1] = Pg[3 * k + 1] - qc1; W[3 * k + 2] = Pg[3 * k + 2] - qc2;
        perm[k] = k;
    }
    __syncwarp();
    // rotation(s) before the main reorder (distmat.py:172-228)
    warp_cov(W, Qg, pl.has_ns ? ns : K, lane, s);
    kabsch_rotation_shared(s, U);
    warp_rotate(W, K, U, lane);
    if (XF) mat3_rmul(Rt, U);
    // phase 0: solute reorder + second solute rotation (distmat.py:184-223); phase 1: main reorder (:231-256).  One call
    // site in a loop, so that the compiler keeps ONE copy of the solver instead of one clone per constant phase
    for (int phase = pl.solute_phase ? 0 : 1; phase < 2 && pl.do_reorder; ++phase) {
        steps += reorder_phase(pl, phase, W, perm, Qg, sm, lane, frame_q, frame_p);
        if (phase == 0) {
            warp_cov(W, Qg, ns, lane, s);
            kabsch_rotation_shared(s, U);
            warp_rotate(W, K, U, lane);
            if (XF) mat3_rmul(Rt, U);
        }
    }

    // final value (distmat.py:259-275)
    if (pl.final_direct) {
        double res = 0.0;
#pragma unroll 1
        for (int k = lane; k < K; k += 32) {
            const double dx = W[3 * k] - Qg[3 * k], dy = W[3 * k + 1] - Qg[3 * k + 1], dz = W[3 * k + 2] - Qg[3 * k + 2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            res += pl.weighted ? d2 * (k < ns ? pl.w_solute : pl.w_solvent) : d2;
        }
        res = warp_sum(res);
        if (XF && lane == 0) {
#pragma unroll
            for (int t = 0; t < 9; ++t) xf[t] = Rt[t];
            xf[9] = 0.0; xf[10] = 0.0; xf[11] = 0.0;
        }
        return pl.weighted ? sqrt(res) : sqrt(res / (double)K);
    }
    if (!pl.weighted) {
        // rmsd.kabsch_rmsd: rotate again (no re-centring) and measure directly
        warp_cov(W, Qg, K, lane, s);
        kabsch_rotation_shared(s, U);
        if (XF) {
            if (pl.has_ns && pl.do_reorder && pl.final_kabsch) mat3_rmul(Rt, U);   // io.py:743-746
            if (lane == 0) {
#pragma unroll
                for (int t = 0; t < 9; ++t) xf[t] = Rt[t];
                xf[9] = 0.0; xf[10] = 0.0; xf[11] = 0.0;
            }
        }
        double res = 0.0;
        for (int k = lane; k < K; k += 32) {
            const double x = W[3 * k], y = W[3 * k + 1], z = W[3 * k + 2];
            const double dx = fma(x, U[0], fma(y, U[3], z * U[6])) - Qg[3 * k];
            const double dy = fma(x, U[1], fma(y, U[4], z * U[7])) - Qg[3 * k + 1];
            const double dz = fma(x, U[2], fma(y, U[5], z * U[8])) - Qg[3 * k + 2];
            res = fma(dx, dx, fma(dy, dy, fma(dz, dz, res)));
        }
        res = warp_sum(res);
        return sqrt(res / (double)K);
    }
    // rmsd.kabsch_weighted_rmsd (distmat.py:273): weighted, weighted-centred covariance; the value is
    // sqrt(max(psq + qsq - 2 lambda, 0)) with sum(w) = 1
    double cp[3] = {0, 0, 0}, cq[3] = {0, 0, 0}, pq = 0.0;
#pragma unroll
    for (int t = 0; t < 9; ++t) s[t] = 0.0;
    double wsum = 0.0;
    for (int k = lane; k < K; k += 32) {
        const double w = k < ns ? pl.w_solute : pl.w_solvent;
        const double px = W[3 * k], py = W[3 * k + 1], pz = W[3 * k + 2];
        const double qx = Qg[3 * k], qy = Qg[3 * k + 1], qz = Qg[3 * k + 2];
        wsum += w;
        cp[0] += w * px; cp[1] += w * py; cp[2] += w * pz;
        cq[0] += w * qx; cq[1] += w * qy; cq[2] += w * qz;
        pq += w * (px * px + py * py + pz * pz + qx * qx + qy * qy + qz * qz);
        s[0] += w * px * qx; s[1] += w * px * qy; s[2] += w * px * qz;
        s[3] += w * py * qx; s[4] += w * py * qy; s[5] += w * py * qz;
        s[6] += w * pz * qx; s[7] += w * pz * qy; s[8] += w * pz * qz;
    }
    wsum = warp_sum(wsum); pq = warp_sum(pq);
#pragma unroll
    for (int t = 0; t < 3; ++t) { cp[t] = warp_sum(cp[t]); cq[t] = warp_sum(cq[t]); }
#pragma unroll
    for (int t = 0; t < 9; ++t) s[t] = warp_sum(s[t]);
    const double iw = 1.0 / wsum;
    double c2 = 0.0;
#pragma unroll
    for (int t = 0; t < 3; ++t) c2 += cp[t] * cp[t] + cq[t] * cq[t];
    const double pqs = pq - c2 * iw;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) s[a * 3 + b] = (s[a * 3 + b] - cp[a] * cq[b] * iw) * iw;
    const double lam = kabsch_rotation_shared(s, U);
    const double msd = pqs * iw - 2.0 * lam;
    if (XF) {
        // rmsd.kabsch_weighted(Q, Pr, W) as used at io.py:733-741: x_ref ~ x_mov U + (c_ref - c_mov U), weighted centres
        if (pl.final_kabsch) {
            mat3_rmul(Rt, U);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                Tt[c] = (cq[c] - fma(cp[0], U[c], fma(cp[1], U[3 + c], cp[2] * U[6 + c]))) * iw;
        }
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < 9; ++t) xf[t] = Rt[t];
            xf[9] = Tt[0]; xf[10] = Tt[1]; xf[11] = Tt[2];
        }
    }
    return sqrt(fmax(msd, 0.0));
}

// pairs given by the condensed slab of rows [row_begin,row_end) (pairs_list == nullptr) or explicitly
template <bool XF>
__global__ void __launch_bounds__(PAIR_WARPS * 32) pair_kernel(const PairPlan pl, const double* __restrict__ centred, int M,
                                                               long long idx_base, long long n_entries,
                                                               const long long* __restrict__ pairs_list,
                                                               double* __restrict__ out, int* __restrict__ perms_out,
                                                               double* __restrict__ work, int* __restrict__ perm_work,
                                                               int smem_per_warp, unsigned long long* counters,
                                                               double* __restrict__ xf_out) {
    extern __shared__ __align__(16) unsigned char pair_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * PAIR_WARPS + warp;
    const long long nw = (long long)gridDim.x * PAIR_WARPS;
    unsigned char* sm = pair_smem + (size_t)warp * smem_per_warp;
    const int wmul = (pl.idx_stride[0] | pl.idx_stride[1]) ? 2 : 1;   // per-frame element lists: room to relabel the moving frame
    double* W = work + (size_t)gw * pl.K * 3 * wmul;
    int* perm = perm_work + (size_t)gw * pl.K * wmul;
    unsigned long long my_steps = 0;
    for (long long e = gw; e < n_entries; e += nw) {
        int i, j;
        if (pairs_list) { i = (int)pairs_list[2 * e]; j = (int)pairs_list[2 * e + 1]; }
        else {
            const long long idx = idx_base + e;
            const double b = 2.0 * (double)M - 1.0;
            long long r = (long long)floor((b - sqrt(b * b - 8.0 * (double)idx)) * 0.5);
            if (r < 0) r = 0;
            while (r > 0 && condensed_row_base(M, r) > idx) --r;
            while (condensed_row_base(M, r + 1) <= idx) ++r;
            i = (int)r;
            j = (int)(idx - condensed_row_base(M, r) + r + 1);
        }
        int steps;
        const double val = pair_value_warp<XF>(pl, centred + (size_t)j * pl.K * 3, centred + (size_t)i * pl.K * 3, W, perm, sm,
                                               lane, steps, XF ? xf_out + (size_t)e * 12 : nullptr, i, j);
        my_steps += (unsigned long long)steps;
        if (lane == 0) out[e] = val;
        if (perms_out) {
            __syncwarp();
            for (int k = lane; k < pl.K; k += 32) perms_out[(size_t)e * pl.K + k] = perm[k];
        }
        __syncwarp();
    }
    if (lane == 0 && my_steps) atomicAdd(counters + 1, my_steps);
}

// ---- host side: build the frame-invariant plan ---------------------------------------------------------
static void element_blocks(const std::vector<int>& view, const int32_t* kept_z, std::vector<int>& off, std::vector<int>& idx,
                           std::vector<int>* block_z = nullptr) {
    // np.unique(ref atoms) order = ascending atomic number; positions ascending inside an element
    std::map<int, std::vector<int>> by_z;
    for (int k : view) by_z[kept_z[k]].push_back(k);
    off.assign(1, 0);
    idx.clear();
    if (block_z) block_z->clear();
    for (auto& kv : by_z) {
        idx.insert(idx.end(), kv.second.begin(), kv.second.end());
        off.push_back((int)idx.size());
        if (block_z) block_z->push_back(kv.first);
    }
}

int pair_plan_build(PairPlanHost** out, const int32_t* kept_z, int K, int natoms, const ct_opts* o, std::string& err) {
    for (int t = 1; t < o->n_excl; ++t)
        if (o->reorder_excl[t] <= o->reorder_excl[t - 1]) {
            err = "reorder_excl must be strictly ascending (the reference re-inserts excluded atoms in list order, "
                  "distmat.py:200-214,247-253; other orders are not implemented)";
            return CT_ERR_UNSUPPORTED;
        }
    PairPlanHost* h = new PairPlanHost();
    PairPlan& p = h->plan;
    p.K = K;
    p.has_ns = o->solute_natoms != 0;
    p.natoms = p.has_ns ? natoms : 0;
    p.do_reorder = o->reorder_mode != 0;
    p.by_distance = o->reorder_mode == 2;
    p.solute_phase = p.has_ns && p.do_reorder && !o->reorder_solvent_only;
    p.final_direct = p.has_ns && p.do_reorder && !o->final_kabsch;
    p.weighted = o->weight_solute != 0.0;
    p.final_kabsch = o->final_kabsch != 0;
    {
        // the two-row collision shortcut pays where the frames are superposed on a solute first (-ns): the solvent atoms
        // then sit near their partners and nearly every collision is a two-row affair
        const char* ce = getenv("CLUSTTRAJ_B200_LSAP_CREDIT");   // development: 0 = plain scans only
        p.credit0 = ce ? atoi(ce) : (p.has_ns ? 6 : 0);
    }
    if (p.weighted) {
        if (!p.has_ns || K == natoms) { delete h; err = "weight_solute needs solute_natoms and at least one solvent atom"; return CT_ERR_UNSUPPORTED; }
        p.w_solute = o->weight_solute / natoms;
        p.w_solvent = (1.0 - o->weight_solute) / (K - natoms);
    }
    std::vector<bool> excl(K, false);
    for (int t = 0; t < o->n_excl; ++t) excl[o->reorder_excl[t]] = true;
    std::vector<int> off[2], idx[2];
    off[0].assign(1, 0); off[1].assign(1, 0);
    if (p.solute_phase) {
        std::vector<int> view;
        for (int k = 0; k < natoms; ++k) if (!excl[k]) view.push_back(k);
        element_blocks(view, kept_z, off[0], idx[0], &h->block_z[0]);
        h->view[0] = view;
    }
    if (p.do_reorder) {
        std::vector<int> view;
        for (int k = p.has_ns ? natoms : 0; k < K; ++k) if (!excl[k]) view.push_back(k);
        element_blocks(view, kept_z, off[1], idx[1], &h->block_z[1]);
        h->view[1] = view;
    }
    h->off[0] = off[0]; h->off[1] = off[1];
    p.nmax = 1;
    for (int ph = 0; ph < 2; ++ph) {
        p.nblocks[ph] = (int)off[ph].size() - 1;
        for (int b = 0; b + 1 < (int)off[ph].size(); ++b) p.nmax = std::max(p.nmax, off[ph][b + 1] - off[ph][b]);
    }
    h->smem_per_warp = pair_smem_per_warp(p.nmax);
    if (const char* pad = getenv("CLUSTTRAJ_B200_PAIR_SMEM_PAD")) h->smem_per_warp += atoi(pad) / 16 * 16;   // development: occupancy probe
    if ((size_t)h->smem_per_warp * PAIR_WARPS > 227 * 1024) {
        // fall back to fewer warps is not implemented; blocks beyond ~680 atoms per element do not fit
        delete h;
        err = "an element block is too large for the shared-memory assignment solver";
        return CT_ERR_UNSUPPORTED;
    }
    std::vector<int> all;
    size_t pos[4];
    pos[0] = all.size(); all.insert(all.end(), off[0].begin(), off[0].end());
    pos[1] = all.size(); all.insert(all.end(), idx[0].begin(), idx[0].end());
    pos[2] = all.size(); all.insert(all.end(), off[1].begin(), off[1].end());
    pos[3] = all.size(); all.insert(all.end(), idx[1].begin(), idx[1].end());
    if (cudaMalloc(&h->d_ints, sizeof(int) * std::max<size_t>(all.size(), 1)) != cudaSuccess ||
        cudaMemcpy(h->d_ints, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        err = std::string("pair plan upload failed: ") + cudaGetErrorString(cudaGetLastError());
        delete h;
        return CT_ERR_CUDA;
    }
    p.block_off[0] = h->d_ints + pos[0]; p.idx[0] = h->d_ints + pos[1];
    p.block_off[1] = h->d_ints + pos[2]; p.idx[1] = h->d_ints + pos[3];
    *out = h;
    return CT_OK;
}

// Per-frame element lists (distmat.py:131-141: every frame's own atoms decide its element blocks).  z = [M][K] atomic numbers
// of the kept atoms of every frame.  The blocks of a pair are (reference frame's atoms of element X) x (moving frame's atoms
// of element X): the same counts in every frame are required (the reference fails on a pair whose counts differ), the
// positions may differ from frame to frame.  Replaces the plan's index lists by one list per frame.
int pair_plan_frame_elements(PairPlanHost* h, const int32_t* z, int M, int K, std::string& err) {
    PairPlan& p = h->plan;
    if (K != p.K) { err = "frame elements: K does not match the loaded frames"; return CT_ERR_ARG; }
    const size_t n0 = h->view[0].size(), n1 = h->view[1].size();
    std::vector<int> all((size_t)M * (n0 + n1));
    std::vector<int> off, idx, bz;
    for (int ph = 0; ph < 2; ++ph) {
        const size_t nv = ph ? n1 : n0, base = ph ? (size_t)M * n0 : 0;
        if (!nv) continue;
        for (int f = 0; f < M; ++f) {
            element_blocks(h->view[ph], z + (size_t)f * K, off, idx, &bz);
            if (off != h->off[ph] || bz != h->block_z[ph]) {
                err = "frame " + std::to_string(f) + " does not have the element counts of frame 0 among the atoms to reorder "
                      "(the reference's per-element assignment needs equal counts in both frames of a pair)";
                return CT_ERR_ARG;
            }
            std::copy(idx.begin(), idx.end(), all.begin() + base + (size_t)f * nv);
        }
    }
    if (h->d_frame_idx) { cudaFree(h->d_frame_idx); h->d_frame_idx = nullptr; }
    if (cudaMalloc(&h->d_frame_idx, sizeof(int) * std::max<size_t>(all.size(), 1)) != cudaSuccess ||
        cudaMemcpy(h->d_frame_idx, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        err = std::string("frame elements upload failed: ") + cudaGetErrorString(cudaGetLastError());
        return CT_ERR_CUDA;
    }
    p.idx[0] = h->d_frame_idx; p.idx_stride[0] = (long long)n0;
    p.idx[1] = h->d_frame_idx + (size_t)M * n0; p.idx_stride[1] = (long long)n1;
    return CT_OK;
}

void pair_plan_free(PairPlanHost* p) {
    if (!p) return;
    cudaFree(p->d_frame_idx);
    cudaFree(p->d_ints); cudaFree(p->d_work); cudaFree(p->d_perm);
    delete p;
}

static cudaError_t launch_pairs(PairPlanHost* h, const double* centred, int M, long long idx_base, long long n_entries,
                                const long long* pairs_list, double* out, int* perms_out, unsigned long long* counters,
                                int sm_count, cudaStream_t stream, double* xf_out = nullptr) {
    if (n_entries <= 0) return cudaSuccess;
    const int smem = h->smem_per_warp * PAIR_WARPS;
    auto kernel = xf_out ? pair_kernel<true> : pair_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, PAIR_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count * per_sm;
    const long long need = (n_entries + PAIR_WARPS - 1) / PAIR_WARPS;
    if (grid > need) grid = need;
    const size_t warps = (size_t)grid * PAIR_WARPS * ((h->plan.idx_stride[0] | h->plan.idx_stride[1]) ? 2 : 1);
    if (warps > h->work_warps) {
        cudaFree(h->d_work); cudaFree(h->d_perm);
        h->d_work = nullptr; h->d_perm = nullptr; h->work_warps = 0;
        if ((e = cudaMalloc(&h->d_work, sizeof(double) * warps * h->plan.K * 3)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&h->d_perm, sizeof(int) * warps * h->plan.K)) != cudaSuccess) return e;
        h->work_warps = warps;
    }
    kernel<<<(unsigned)grid, PAIR_WARPS * 32, smem, stream>>>(h->plan, centred, M, idx_base, n_entries, pairs_list, out,
                                                              perms_out, h->d_work, h->d_perm, h->smem_per_warp, counters,
                                                              xf_out);
    return cudaGetLastError();
}

cudaError_t launch_pairs_rows(const PairPlanHost* plan, const double* centred, int M, int K, int row_begin, int row_end,
                              double* out, long long idx_base, unsigned long long* counters, int sm_count,
                              cudaStream_t stream) {
    (void)K;
    const long long last = row_end < M - 1 ? row_end : M - 1;
    const long long n = last > row_begin ? condensed_row_base(M, last) - condensed_row_base(M, row_begin) : 0;
    return launch_pairs(const_cast<PairPlanHost*>(plan), centred, M, idx_base, n, nullptr, out, nullptr, counters, sm_count,
                        stream);
}

cudaError_t launch_pairs_list(const PairPlanHost* plan, const double* centred, int M, int K, const long long* pairs_dev,
                              long long n_pairs, double* values_dev, int* perms_dev, unsigned long long* counters,
                              int sm_count, cudaStream_t stream, double* xf_dev) {
    (void)K;
    return launch_pairs(const_cast<PairPlanHost*>(plan), centred, M, 0, n_pairs, pairs_dev, values_dev, perms_dev, counters,
                        sm_count, stream, xf_dev);
}

}  // namespace ct
