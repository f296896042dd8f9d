// K6 — hierarchical clustering of the condensed RMSD matrix in HBM ("linkage on GPU"):
// scipy.cluster.hierarchy.linkage(distmat, method) as clusttraj calls it at clusttraj/classify.py:30,99,127, for the
// methods SciPy serves with its nearest-neighbour-chain algorithm (complete, average, weighted, ward) and with Prim's
// minimum spanning tree (single).  SciPy is a third-party dependency of the reference (requirements.txt), importable on
// the GPU box: the parity tests compare the linkage matrix Z bit for bit with the live module; the CPU restatement this
// file follows step for step is oracle/linkage_restated.py (pinned against the same module, ties and duplicates
// included).  centroid / median (SciPy's generic priority-queue algorithm) stay on the host.
//
// The algorithm is a chain of dependent steps, each a pass over ONE row of the matrix (M entries): "nearest live cluster
// of the chain tip" (<= 3 M of them; the Lance-Williams update of the row of a cluster merged in the step before rides
// in the same pass, the one entry the two share being handed over in registers), or one Prim step.  8 M bytes per
// step: a step is bound by memory latency and synchronisation, never by bandwidth.  Two drivers of the same state
// machine (LinkState, in device memory):
//   * linkage_cluster_kernel (default): ONE persistent launch on a single thread-block cluster over the square form
//     of the matrix; one hardware cluster barrier and one distributed-shared-memory fold per step (see its comment);
//   * linkage_step_kernel (fall-back when the square form does not fit or the cluster cannot be placed): one launch per
//     step over up to 148 CTAs on the condensed layout; the last CTA to finish (ticket counter) folds the per-CTA
//     partial results and advances the state machine, so the host never reads anything back between steps: it replays
//     a CUDA graph of identical step launches until the state says "done" (steps after that are no-ops).
//
// Exactness: the step order, the tie rules (previous chain element first, then the lowest index; the merged cluster
// lives in the higher slot) and the floating-point operation order of the update formulas are SciPy's; the formulas
// use explicitly rounded operations (__dmul_rn ...) so that the compiler cannot contract them into FMAs.
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <vector>

namespace ct {

enum { LK_SINGLE = 0, LK_COMPLETE = 1, LK_AVERAGE = 2, LK_WARD = 5, LK_WEIGHTED = 6 };  // SciPy's method codes
enum { ACT_SCAN = 0, ACT_DONE = 2, ACT_PRIM = 3 };
constexpr int LK_THREADS = 256;
constexpr int LK_NONE = 0x7fffffff;

struct LinkState {
    int action;
    int x, y;        // x = chain tip to scan from, y = SciPy's loop variable y (persists)
    int prev;        // chain element under the tip, or -1
    int clen, k;     // chain length, merges recorded
    int upd;         // 1: the merge (a, b) decided in the step before still has to be applied to column b
    int a, b, na, nb;
    int first_live;  // lowest live slot (where SciPy restarts an emptied chain)
    double dxy;      // distance of that merge
    unsigned int ticket;
    int steps;       // step launches that did work
};

__device__ __forceinline__ long long cidx(long long n, long long i, long long j) {  // i != j
    const long long a = i < j ? i : j, b = i < j ? j : i;
    return condensed_row_base(n, a) + (b - a - 1);
}

__device__ __forceinline__ double lw_update(int method, double dxi, double dyi, double dxy, int nx, int ny, int ni) {
    switch (method) {
        case LK_COMPLETE: return dxi > dyi ? dxi : dyi;
        case LK_AVERAGE:
            return __ddiv_rn(__dadd_rn(__dmul_rn((double)nx, dxi), __dmul_rn((double)ny, dyi)), (double)(nx + ny));
        case LK_WEIGHTED: return __dmul_rn(0.5, __dadd_rn(dxi, dyi));
        default: {  // ward
            const double t = __ddiv_rn(1.0, (double)(nx + ny + ni));
            const double a = __dmul_rn(__dmul_rn(__dmul_rn((double)(ni + nx), t), dxi), dxi);
            const double b = __dmul_rn(__dmul_rn(__dmul_rn((double)(ni + ny), t), dyi), dyi);
            const double c = __dmul_rn(__dmul_rn(__dmul_rn((double)ni, t), dxy), dxy);
            return __dsqrt_rn(__dsub_rn(__dadd_rn(a, b), c));
        }
    }
}

// lexicographic minimum of (value, index) over a block of NT threads; the result is valid in thread 0
template <int NT>
__device__ __forceinline__ void block_min_pair(double& d, int& i, double* sd, int* si) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (od < d || (od == d && oi < i)) { d = od; i = oi; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sd[w] = d; si[w] = i; }
    __syncthreads();
    if (w == 0) {
        d = l < NT / 32 ? sd[l] : __longlong_as_double(0x7ff0000000000000LL);
        i = l < NT / 32 ? si[l] : LK_NONE;
#pragma unroll
        for (int o = NT / 64; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, d, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (od < d || (od == d && oi < i)) { d = od; i = oi; }
        }
    }
    __syncthreads();
}

// order-preserving 64-bit key of a double (no NaNs reach the reductions: a NaN distance never wins a '<')
__device__ __forceinline__ unsigned long long dkey(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return b ^ ((b >> 63) ? ~0ULL : 0x8000000000000000ULL);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ULL : ~0ULL)));
}
// lexicographic warp minimum of (key, index) with three REDUX operations; every lane gets the result
__device__ __forceinline__ void warp_min_key(unsigned long long& k, int& i) {
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const unsigned mi = __reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? (unsigned)i : 0xffffffffu);
    k = ((unsigned long long)mhi << 32) | mlo;
    i = (int)mi;
}

// D: working copy of the condensed matrix (NN-chain methods: updated in place; single: read only)
__global__ void __launch_bounds__(LK_THREADS) linkage_step_kernel(LinkState* __restrict__ st, double* __restrict__ D,
                                                                  int* __restrict__ size, int* __restrict__ chain,
                                                                  double* __restrict__ Zraw, double* __restrict__ dmin,
                                                                  double* __restrict__ part_d, int* __restrict__ part_i,
                                                                  int M, int method) {
    __shared__ LinkState s;
    __shared__ double sd[LK_THREADS / 32];
    __shared__ int si[LK_THREADS / 32];
    __shared__ bool last;
    const int gtid = blockIdx.x * LK_THREADS + threadIdx.x, gthreads = gridDim.x * LK_THREADS;
    int sz = gtid < M ? size[gtid] : 0;   // does not depend on the state: in flight while the state arrives
    if (threadIdx.x == 0) s = *st;
    __syncthreads();
    const int action = s.action;
    if (action == ACT_DONE) return;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int x = s.x;
    // what the state machine will need if this scan ends in a merge (the partner can only be the previous chain
    // element): fetched now, alongside the row, instead of as dependent loads after the reduction
    double t_dprev = INF;
    int t_sx = 0, t_sp = 0, t_c3 = -1, t_c4 = -1;
    if (threadIdx.x == 0 && action == ACT_SCAN && s.clen > 1) {
        t_dprev = D[cidx(M, x, s.prev)];   // never in the column an update of this step rewrites
        t_sx = size[x]; t_sp = size[s.prev];
        if (s.clen >= 3) t_c3 = chain[s.clen - 3];
        if (s.clen >= 4) t_c4 = chain[s.clen - 4];
    }
    double bd = INF;
    int bi = LK_NONE;
    if (action == ACT_SCAN) {
        const int upd = s.upd, a = s.a, b = s.b, na = s.na, nb = s.nb;
        const double dxy = s.dxy;
        for (int i = gtid; i < M; i += gthreads) {
            if (i != gtid) sz = size[i];
            const bool live = sz != 0;
            double cd = INF;
            int ci = i;
            bool has = false;
            if (upd) {
                if (live && i != b) {
                    // Lance-Williams update of D[i, b]; when the tip is b itself these ARE the scanned values, and the
                    // tip's own entry is the candidate "b" of its scan
                    const long long ib = cidx(M, i, b);
                    const double dn = lw_update(method, D[cidx(M, i, a)], D[ib], dxy, na, nb, sz);
                    D[ib] = dn;
                    if (x == b) { cd = dn; has = true; }
                    else if (i == x) { cd = dn; ci = b; has = true; }
                }
                if (x != b && live && i != x && i != b) { cd = D[cidx(M, x, i)]; has = true; }
            } else if (live && i != x) {
                cd = D[cidx(M, x, i)];
                has = true;
            }
            if (has && (cd < bd || (cd == bd && ci < bi))) { bd = cd; bi = ci; }
        }
    } else {  // ACT_PRIM
        for (int i = gtid; i < M; i += gthreads) {
            if (i != gtid) sz = size[i];
            if (sz == 0) continue;  // already in the tree
            const double d = D[cidx(M, x, i)];
            double dm = dmin[i];
            if (dm > d) { dm = d; dmin[i] = d; }
            if (dm < bd) { bd = dm; bi = i; }
        }
    }
    block_min_pair<LK_THREADS>(bd, bi, sd, si);
    if (threadIdx.x == 0) {
        part_d[blockIdx.x] = bd;
        part_i[blockIdx.x] = bi;
        __threadfence();
        last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    bd = INF; bi = LK_NONE;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += LK_THREADS) {
        const double od = __ldcg(part_d + b);
        const int oi = __ldcg(part_i + b);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    block_min_pair<LK_THREADS>(bd, bi, sd, si);
    if (threadIdx.x != 0) return;
    // ---- state machine, one thread ----
    LinkState n = s;
    n.ticket = 0;
    n.steps = s.steps + 1;
    n.upd = 0;
    if (action == ACT_SCAN) {
        int y = s.y;
        double cur = INF;
        if (s.clen > 1) { y = s.prev; cur = t_dprev; }
        if (bi != LK_NONE && bd < cur) { cur = bd; y = bi; }
        n.y = y;
        if (s.clen > 1 && y == s.prev) {
            // reciprocal nearest neighbours: merge, the new cluster lives in the higher slot; its row is updated by the
            // next step, which also scans from the new chain tip
            const int a = x < y ? x : y, b = x < y ? y : x;
            const int na = x < y ? t_sx : t_sp, nb = x < y ? t_sp : t_sx;
            double* z = Zraw + 4 * (size_t)s.k;
            z[0] = (double)a; z[1] = (double)b; z[2] = cur; z[3] = (double)(na + nb);
            size[a] = 0; size[b] = na + nb;
            n.k = s.k + 1;
            if (n.k == M - 1) n.action = ACT_DONE;
            else {
                n.upd = 1; n.a = a; n.b = b; n.na = na; n.nb = nb; n.dxy = cur;
                int f = s.first_live;
                if (a == f) { do ++f; while (size[f] == 0); }
                n.first_live = f;
                n.clen = s.clen - 2;
                if (n.clen == 0) { chain[0] = f; n.clen = 1; n.x = f; n.prev = -1; }
                else { n.x = t_c3; n.prev = t_c4; }
            }
        } else {
            chain[s.clen] = y;
            n.clen = s.clen + 1;
            n.prev = x; n.x = y;
        }
    } else {  // ACT_PRIM
        const int y = bi != LK_NONE ? bi : s.y;
        double* z = Zraw + 4 * (size_t)s.k;
        z[0] = (double)x; z[1] = (double)y; z[2] = bd; z[3] = 0.0;
        size[y] = 0;
        n.x = y; n.y = y; n.k = s.k + 1;
        if (n.k == M - 1) n.action = ACT_DONE;
    }
    *st = n;
}


// ---- the same state machine as ONE persistent kernel on a single thread-block cluster ------------------------------
// For the sizes clusttraj meets (up to ~10^5 frames) a step is bound by a handful of dependent memory round trips, not
// by bandwidth, so the fastest layout keeps everything that is not the matrix on chip: a cluster of 8 or 16 CTAs x 1024
// threads covers the row, each CTA reduces its part, ONE hardware cluster barrier per step publishes the partial
// results (read through distributed shared memory) together with the global writes of the step, and every CTA
// advances its own replica of the state machine (identical inputs -> identical decisions; CTA 0 alone writes Z, the
// chain and the cluster sizes).  No launches, tickets or state reloads between steps.  What CTA 0 wrote in step n is
// only guaranteed visible to the others after the barrier of step n + 1; the replicas bridge that one step from their
// own state (the merge just decided patches size[a], size[b]; chain entries are first read two steps after they are
// written).  The matrix, the sizes and the chain are read with ld.global.cg (L1 is not coherent across the CTAs).
// This kernel works on the SQUARE form of the matrix (M x M, expanded once from the condensed vector by
// expand_square_kernel): every row it reads — the tip's, and the two rows of a merge — is then one contiguous,
// coalesced stream with a one-add address per entry, where the condensed layout costs a 32-byte DRAM sector and two
// 64-bit multiplies per entry of the column part; an update writes the merged cluster's row and (posted, off the
// critical path) its column, so the square stays symmetric.
constexpr int LC_THREADS = 1024;
constexpr int LC_BATCH = 4;   // entries a thread has in flight at once

// One pass over the row of the current step for the entries i = gtid, gtid + gthreads, ...: B entries per thread are
// fetched together (all their loads in flight at once), then processed.
template <int B>
__device__ __forceinline__ void row_pass(const LinkState& s, double* __restrict__ D, const int* __restrict__ size,
                                         double* __restrict__ dmin, int M, int method, int gtid, int gthreads, double& bd,
                                         int& bi) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int action = s.action, x = s.x, upd = s.upd, a = s.a, b = s.b, na = s.na, nb = s.nb;
    const double* __restrict__ rowx = D + (size_t)x * M;
    if (action == ACT_SCAN && !upd) {
        // plain scan (two steps out of three): nearest live cluster of the tip
        for (int i0 = gtid; i0 < M; i0 += B * gthreads) {
            int sz[B];
            double dxi[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int i = i0 + u * gthreads;
                sz[u] = 0; dxi[u] = INF;
                if (i < M && i != x) { sz[u] = __ldcg(size + i); dxi[u] = __ldcg(rowx + i); }
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int i = i0 + u * gthreads;
                if (sz[u] != 0 && dxi[u] < bd) { bd = dxi[u]; bi = i; }   // ascending i per thread: '<' keeps the lowest index
            }
        }
    } else if (action == ACT_SCAN) {
        const double dxy = s.dxy;
        const double* __restrict__ rowa = D + (size_t)a * M;
        double* __restrict__ rowb = D + (size_t)b * M;
        for (int i0 = gtid; i0 < M; i0 += B * gthreads) {
            // all loads of a batch of entries are issued together (addresses are valid whatever the slot's state):
            // one memory round trip per step instead of "size, then distance", entry after entry
            int sz[B];
            double dxi[B], dai[B], dbi[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int i = i0 + u * gthreads;
                sz[u] = 0; dxi[u] = INF; dai[u] = 0.0; dbi[u] = 0.0;
                if (i < M) {
                    sz[u] = __ldcg(size + i);
                    if (i != x) dxi[u] = __ldcg(rowx + i);
                    if (i != a && i != b) { dai[u] = __ldcg(rowa + i); dbi[u] = __ldcg(rowb + i); }
                }
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int i = i0 + u * gthreads;
                if (i >= M) break;
                int szi = sz[u];
                if (i == a) szi = 0; else if (i == b) szi = na + nb;
                const bool live = szi != 0;
                double cd = INF;
                int ci = i;
                bool has = false;
                if (live && i != b) {
                    const double dn = lw_update(method, dai[u], dbi[u], dxy, na, nb, szi);
                    rowb[i] = dn;
                    D[(size_t)i * M + b] = dn;
                    if (x == b) { cd = dn; has = true; }
                    else if (i == x) { cd = dn; ci = b; has = true; }
                }
                if (x != b && live && i != x && i != b) { cd = dxi[u]; has = true; }
                if (has && (cd < bd || (cd == bd && ci < bi))) { bd = cd; bi = ci; }
            }
        }
    } else {  // ACT_PRIM: x joined the tree in the step before (its flag may not have arrived yet)
        for (int i0 = gtid; i0 < M; i0 += B * gthreads) {
            int sz[B];
            double d[B], dm[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int i = i0 + u * gthreads;
                sz[u] = 0; d[u] = INF; dm[u] = INF;
                if (i < M && i != x) { sz[u] = __ldcg(size + i); d[u] = __ldcg(rowx + i); dm[u] = dmin[i]; }
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int i = i0 + u * gthreads;
                if (i >= M) break;
                if (sz[u] == 0) continue;
                double m = dm[u];
                if (m > d[u]) { m = d[u]; dmin[i] = m; }
                if (m < bd) { bd = m; bi = i; }
            }
        }
    }
}

__global__ void __launch_bounds__(LC_THREADS, 1) linkage_cluster_kernel(LinkState* __restrict__ st_out, double* __restrict__ D,
                                                                        int* __restrict__ size, int* __restrict__ chain,
                                                                        double* __restrict__ Zraw, double* __restrict__ dmin,
                                                                        int M, int method, long long max_steps) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), ncta = (int)cluster.num_blocks();
    __shared__ LinkState s;
    __shared__ unsigned long long sk[LC_THREADS / 32];
    __shared__ int si[LC_THREADS / 32];
    __shared__ __align__(16) ulonglong2 cpart[2];   // this CTA's partial result {key, index}, double-buffered by step parity
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if (threadIdx.x == 0) s = *st_out;
    __syncthreads();
    const int gtid = rank * LC_THREADS + threadIdx.x, gthreads = ncta * LC_THREADS;
    int par = 0;
    for (long long step = 0; step < max_steps; ++step) {
        const int action = s.action;
        if (action == ACT_DONE) break;
        const int x = s.x, upd = s.upd, a = s.a, b = s.b, na = s.na, nb = s.nb;
        double t_dprev = INF;
        int t_sx = 0, t_sp = 0, t_c3 = -1, t_c4 = -1;
        const double* __restrict__ rowx = D + (size_t)x * M;
        if (threadIdx.x == 0 && action == ACT_SCAN && s.clen > 1) {
            const int prev = s.prev;
            t_dprev = __ldcg(rowx + prev);
            t_sx = __ldcg(size + x); t_sp = __ldcg(size + prev);
            if (s.clen >= 3) t_c3 = __ldcg(chain + s.clen - 3);
            if (s.clen >= 4) t_c4 = __ldcg(chain + s.clen - 4);
        }
        double bd = INF;
        int bi = LK_NONE;
        if (M <= 2 * gthreads) row_pass<1>(s, D, size, dmin, M, method, gtid, gthreads, bd, bi);
        else row_pass<LC_BATCH>(s, D, size, dmin, M, method, gtid, gthreads, bd, bi);
        // block minimum (REDUX per warp, then over the 32 warp results), published in this CTA's shared memory
        unsigned long long bk = dkey(bd);
        warp_min_key(bk, bi);
        if ((threadIdx.x & 31) == 0) { sk[threadIdx.x >> 5] = bk; si[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x < 32) {
            bk = threadIdx.x < LC_THREADS / 32 ? sk[threadIdx.x] : ~0ULL;
            bi = threadIdx.x < LC_THREADS / 32 ? si[threadIdx.x] : LK_NONE;
            warp_min_key(bk, bi);
            if (threadIdx.x == 0) cpart[par] = make_ulonglong2(bk, (unsigned long long)(unsigned)bi);
        }
        cluster.sync();   // partials (shared) and this step's matrix / bookkeeping writes (global) are now visible cluster-wide
        if (threadIdx.x < 32) {
            bk = ~0ULL; bi = LK_NONE;
            if ((int)threadIdx.x < ncta) {
                const ulonglong2 r = *cluster.map_shared_rank(&cpart[par], threadIdx.x);
                bk = r.x; bi = (int)(unsigned)r.y;
            }
            warp_min_key(bk, bi);
            bd = dunkey(bk);
        }
        if (threadIdx.x == 0) {
            const bool writer = rank == 0;
            LinkState n = s;
            n.steps = s.steps + 1;
            n.upd = 0;
            if (action == ACT_SCAN) {
                int y = s.y;
                double cur = INF;
                if (s.clen > 1) { y = s.prev; cur = t_dprev; }
                if (bi != LK_NONE && bd < cur) { cur = bd; y = bi; }
                n.y = y;
                if (s.clen > 1 && y == s.prev) {
                    if (upd) {   // the sizes of the merge decided one step ago may not have arrived in memory yet
                        if (x == a) t_sx = 0; else if (x == b) t_sx = na + nb;
                        if (y == a) t_sp = 0; else if (y == b) t_sp = na + nb;
                    }
                    const int ma = x < y ? x : y, mb = x < y ? y : x;
                    const int mna = x < y ? t_sx : t_sp, mnb = x < y ? t_sp : t_sx;
                    if (writer) {
                        double* z = Zraw + 4 * (size_t)s.k;
                        z[0] = (double)ma; z[1] = (double)mb; z[2] = cur; z[3] = (double)(mna + mnb);
                        size[ma] = 0; size[mb] = mna + mnb;
                    }
                    n.k = s.k + 1;
                    if (n.k == M - 1) n.action = ACT_DONE;
                    else {
                        n.upd = 1; n.a = ma; n.b = mb; n.na = mna; n.nb = mnb; n.dxy = cur;
                        int f = s.first_live;
                        if (ma == f) { do ++f; while (f != mb && __ldcg(size + f) == 0); }   // slots in (ma, mb) are stable
                        n.first_live = f;
                        n.clen = s.clen - 2;
                        if (n.clen == 0) { if (writer) chain[0] = f; n.clen = 1; n.x = f; n.prev = -1; }
                        else { n.x = t_c3; n.prev = t_c4; }
                    }
                } else {
                    if (writer) chain[s.clen] = y;
                    n.clen = s.clen + 1;
                    n.prev = x; n.x = y;
                }
            } else {
                const int y = bi != LK_NONE ? bi : s.y;
                if (writer) {
                    double* z = Zraw + 4 * (size_t)s.k;
                    z[0] = (double)x; z[1] = (double)y; z[2] = bd; z[3] = 0.0;
                    size[y] = 0;
                }
                n.x = y; n.y = y; n.k = s.k + 1;
                if (n.k == M - 1) n.action = ACT_DONE;
            }
            s = n;
        }
        __syncthreads();
        par ^= 1;
    }
    cluster.sync();   // nobody leaves while a neighbour may still read its shared memory
    if (rank == 0 && threadIdx.x == 0) *st_out = s;
}

// square form of the condensed vector (zero diagonal); one block row per matrix row
__global__ void __launch_bounds__(256) expand_square_kernel(const double* __restrict__ cond, double* __restrict__ sq, int M) {
    const int r = blockIdx.x;
    const int c = blockIdx.y * 256 + threadIdx.x;
    if (c >= M) return;
    sq[(size_t)r * M + c] = r == c ? 0.0 : cond[cidx(M, r, c)];
}

__global__ void linkage_init_kernel(LinkState* st, int* size, int* chain, double* dmin, int M, int method) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
        size[i] = (method == LK_SINGLE && i == 0) ? 0 : 1;
        dmin[i] = __longlong_as_double(0x7ff0000000000000LL);
    }
    if (i == 0) {
        LinkState n{};
        n.action = M < 2 ? ACT_DONE : (method == LK_SINGLE ? ACT_PRIM : ACT_SCAN);
        n.x = 0; n.y = 0; n.prev = -1; n.clen = 1; n.k = 0; n.first_live = 0;
        chain[0] = 0;
        *st = n;
    }
}

// scratch layout (bytes): state | size[M] | chain[M] | dmin[M] | part_d[grid] | part_i[grid] | Zraw[4 (M - 1)]
size_t linkage_scratch_bytes(int M, int sm_count) {
    return 256 + sizeof(int) * 2 * (size_t)M + sizeof(double) * (size_t)M + 12 * (size_t)sm_count + 64 +
           sizeof(double) * 4 * (size_t)M;
}

// Clusters the condensed matrix `cond` (left untouched) and leaves the raw merges in `zraw_host` (pinned or pageable,
// 4 (M - 1) doubles).  `work` is scratch for the matrix the algorithm updates: M * M doubles select the persistent
// cluster kernel on the square form, M (M - 1) / 2 the step-launch variant on a condensed copy.  `done_flag` is a pinned
// int the state's action is copied to between graph replays.
cudaError_t run_linkage(const double* cond, double* work, size_t work_doubles, int M, int method, unsigned char* scratch,
                        int sm_count, double* zraw_host, int* done_flag, long long* steps_out, cudaStream_t stream) {
    const size_t L = (size_t)M * (M - 1) / 2;
    double* D = nullptr;
    LinkState* st = reinterpret_cast<LinkState*>(scratch);
    int* size = reinterpret_cast<int*>(scratch + 256);
    int* chain = size + M;
    double* dmin = reinterpret_cast<double*>(scratch + 256 + sizeof(int) * 2 * (size_t)M);
    double* part_d = dmin + M;
    int* part_i = reinterpret_cast<int*>(part_d + sm_count);
    double* Zraw = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(part_i) + ((sizeof(int) * sm_count + 7) / 8) * 8);
    int grid = (M + LK_THREADS - 1) / LK_THREADS;
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    linkage_init_kernel<<<(M + 255) / 256, 256, 0, stream>>>(st, size, chain, dmin, M, method);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const long long max_steps = 3LL * M + 8;
    const char* force = std::getenv("CLUSTTRAJ_B200_LINKAGE");   // "graph" / "cluster": pick a variant (tests, timing)
    if (!(force && force[0] == 'g') && work_doubles >= (size_t)M * M) {
        // preferred: one persistent kernel on a single cluster (up to 16 CTAs where the GPU allows it, else 8)
        D = work;
        cudaFuncSetAttribute(linkage_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int want = 1;   // smallest power of two of CTAs that gives every entry of a row its own thread, at most 16
        while (want < 16 && want * LC_THREADS < M) want <<= 1;
        for (int ncta = want; ncta >= (want == 16 ? 8 : want); ncta >>= 1) {
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)ncta); cfg.blockDim = dim3(LC_THREADS); cfg.stream = stream;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, linkage_cluster_kernel, &cfg) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                continue;
            }
            expand_square_kernel<<<dim3(M, (M + 255) / 256), 256, 0, stream>>>(cond, D, M);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaLaunchKernelEx(&cfg, linkage_cluster_kernel, st, D, size, chain, Zraw, dmin, M, method, max_steps)) != cudaSuccess)
                return e;
            LinkState fin;
            if ((e = cudaMemcpyAsync(&fin, st, sizeof fin, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if (M > 1 && (e = cudaMemcpyAsync(zraw_host, Zraw, sizeof(double) * 4 * (size_t)(M - 1), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
                return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            if (fin.action != ACT_DONE) return cudaErrorUnknown;
            *steps_out = fin.steps;
            return cudaSuccess;
        }
        if (force && force[0] == 'c') return cudaErrorInvalidConfiguration;
    }
    if (force && force[0] == 'c') return cudaErrorInvalidConfiguration;
    if (method == LK_SINGLE) D = const_cast<double*>(cond);   // Prim only reads
    else {
        if (work_doubles < L) return cudaErrorInvalidValue;
        if ((e = cudaMemcpyAsync(work, cond, sizeof(double) * L, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
        D = work;
    }
    // fall-back: a graph of identical step launches over the whole GPU, replayed until the state machine reports "done"
    const int batch = (int)std::min<long long>(max_steps, 2048);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if ((e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return e;
    for (int b = 0; b < batch; ++b)
        linkage_step_kernel<<<grid, LK_THREADS, 0, stream>>>(st, D, size, chain, Zraw, dmin, part_d, part_i, M, method);
    e = cudaStreamEndCapture(stream, &graph);
    if (e != cudaSuccess) return e;
    if ((e = cudaGraphInstantiate(&exec, graph, 0)) != cudaSuccess) { cudaGraphDestroy(graph); return e; }
    cudaError_t rc = cudaSuccess;
    bool done = false;
    for (long long launched = 0; launched < max_steps + batch && !done; launched += batch) {
        if ((rc = cudaGraphLaunch(exec, stream)) != cudaSuccess) break;
        if ((rc = cudaMemcpyAsync(done_flag, &st->action, sizeof(int), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) break;
        if ((rc = cudaStreamSynchronize(stream)) != cudaSuccess) break;
        done = *done_flag == ACT_DONE;
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    if (rc != cudaSuccess) return rc;
    if (!done) return cudaErrorUnknown;  // cannot happen: the chain algorithm needs at most 3 M scans
    int steps = 0;
    if ((rc = cudaMemcpyAsync(&steps, &st->steps, sizeof(int), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return rc;
    if (M > 1 && (rc = cudaMemcpyAsync(zraw_host, Zraw, sizeof(double) * 4 * (size_t)(M - 1), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
        return rc;
    if ((rc = cudaStreamSynchronize(stream)) != cudaSuccess) return rc;
    *steps_out = steps;
    return cudaSuccess;
}

// what SciPy does after either algorithm: stable sort of the merges by distance, then union-find relabelling
// (new clusters numbered M, M + 1, ... in sorted order; the two members of a merge in ascending order)
void linkage_finish(const double* zraw, int M, double* Z) {
    const int n = M - 1;
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    // numpy's order: ascending, NaNs (a ward radicand rounded below zero) last, equal keys in input order
    std::stable_sort(order.begin(), order.end(), [zraw](int a, int b) {
        const double za = zraw[4 * (size_t)a + 2], zb = zraw[4 * (size_t)b + 2];
        return za < zb || (za == za && zb != zb);
    });
    std::vector<int> parent(2 * (size_t)M - 1), csize(2 * (size_t)M - 1, 0);
    std::iota(parent.begin(), parent.end(), 0);
    std::fill(csize.begin(), csize.begin() + M, 1);
    auto find = [&parent](int v) {
        int r = v;
        while (parent[r] != r) r = parent[r];
        while (parent[v] != r) { const int nx = parent[v]; parent[v] = r; v = nx; }
        return r;
    };
    int next = M;
    for (int k = 0; k < n; ++k) {
        const double* z = zraw + 4 * (size_t)order[k];
        const int xr = find((int)z[0]), yr = find((int)z[1]);
        double* o = Z + 4 * (size_t)k;
        o[0] = (double)(xr < yr ? xr : yr);
        o[1] = (double)(xr < yr ? yr : xr);
        o[2] = z[2];
        parent[xr] = next; parent[yr] = next;
        csize[next] = csize[xr] + csize[yr];
        o[3] = (double)csize[next];
        ++next;
    }
}

}  // namespace ct
