// 3x3 superposition solvers used by the Gram epilogue and the per-pair kernels.
//
// What they replace in the reference: rmsd.kabsch (numpy.linalg.svd of the 3x3 covariance + the
// det(V)det(W) reflection fix) as reached from clusttraj/distmat.py:178,220,227 and, through
// rmsd.kabsch_rmsd, from distmat.py:275.  Neither function is a translation of that code: the
// RMSD only needs  lambda = s1 + s2 + sign(det S) s3  (S = covariance, s = singular values), which is
// the largest eigenvalue of Horn's symmetric 4x4 quaternion matrix; its characteristic polynomial in
// terms of three invariants of S is
//     f(x) = (x^2 - a)^2 - 8 d x - 4 b,   a = |S|_F^2,  b = |cof S|_F^2,  d = det S,
// and lambda is its largest root.  The rotation itself (needed before a Hungarian reorder) is the
// eigenvector of that root.
#pragma once
#include <cuda_runtime.h>

namespace ct {

// Invariants of the 3x3 covariance S (row-major s[9]).  39 FP64 operations.
__device__ __forceinline__ void quartic_invariants(const double* s, double& a, double& b, double& d) {
    const double c00 = s[4] * s[8] - s[5] * s[7];
    const double c01 = s[5] * s[6] - s[3] * s[8];
    const double c02 = s[3] * s[7] - s[4] * s[6];
    const double c10 = s[2] * s[7] - s[1] * s[8];
    const double c11 = s[0] * s[8] - s[2] * s[6];
    const double c12 = s[1] * s[6] - s[0] * s[7];
    const double c20 = s[1] * s[5] - s[2] * s[4];
    const double c21 = s[2] * s[3] - s[0] * s[5];
    const double c22 = s[0] * s[4] - s[1] * s[3];
    a = s[0] * s[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) a = fma(s[k], s[k], a);
    b = c00 * c00;
    b = fma(c01, c01, b); b = fma(c02, c02, b);
    b = fma(c10, c10, b); b = fma(c11, c11, b); b = fma(c12, c12, b);
    b = fma(c20, c20, b); b = fma(c21, c21, b); b = fma(c22, c22, b);
    d = fma(s[0], c00, fma(s[1], c01, s[2] * c02));
}

// Largest root of f.  `half_g` = (G_i + G_j)/2 is an upper bound of the root (E >= 0), sqrt(3a) is
// another; Newton from above on a polynomial with only real roots is monotone, so the coarse phase
// runs in FP32 on the normalised polynomial (separate pipe from the FP64 MMA) and FP64 only polishes.
// Returns false when the iteration did not settle (near-double root: mirror images, collinear atoms,
// or non-finite intermediates); the caller then re-does the pair with the explicit-rotation path.
__device__ __forceinline__ bool largest_root(double a, double b, double d, double half_g, double& lam_out) {
    const float af0 = (float)a;
    float l0 = fminf((float)half_g, sqrtf(3.0f * af0)) * 1.000001f;
    if (!(l0 > 0.0f)) {  // S == 0 (e.g. zero padding): all roots are 0
        lam_out = 0.0;
        return a == 0.0;
    }
    const float sc = 1.0f / l0;
    const float sc2 = sc * sc;
    const float af = af0 * sc2;
    const float df2 = 2.0f * ((float)d * sc2 * sc);
    const float bf4 = 4.0f * ((float)b * sc2 * sc2);
    float x = 1.0f;
#pragma unroll 1
    for (int it = 0; it < 14; ++it) {
        const float t = fmaf(x, x, -af);
        const float f = fmaf(t, t, fmaf(-4.0f * df2, x, -bf4));
        const float fq = fmaf(x, t, -df2);
        if (!(fq > 0.0f)) break;
        const float step = 0.25f * __fdividef(f, fq);
        x -= step;
        if (fabsf(step) <= 3e-7f * x) break;
    }
    double lam = (double)x * (double)l0;
    const double m2d = -2.0 * d, m4b = -4.0 * b;
    double step = 0.0;
    bool ok = false;
#pragma unroll 1
    for (int it = 0; it < 24; ++it) {
        const double t = fma(lam, lam, -a);
        const double f = fma(t, t, fma(4.0 * m2d, lam, m4b));
        const double fq = fma(lam, t, m2d);
        const float r = 0.25f * __frcp_rn((float)fq);
        step = f * (double)r;
        lam -= step;
        if (it >= 1 && fabs(step) <= 1e-13 * lam) { ok = true; break; }
    }
    lam_out = lam;
    return ok && (lam == lam);
}

// The same solve for NP independent pairs held by one thread, advanced in lockstep.  The FP64 pipe is shared
// with the DMMAs of the MMA warps, so every FP64 instruction issued here queues behind 16-cycle DMMAs: the
// FP64 instruction count per pair is what bounds the epilogue, and it is kept minimal — one polishing
// Newton step (quadratic: an FP32-converged start with relative error ~1e-7 lands at ~1e-13) accepted when
// the step itself is below 1e-6 * lambda; only lanes that fail the test iterate further.
// ok_mask bit p is set when pair p settled.
template <int NP>
__device__ __forceinline__ unsigned largest_root_lockstep(const double (&a)[NP], const double (&b)[NP],
                                                          const double (&d)[NP], const double (&half_g)[NP],
                                                          double (&lam)[NP]) {
    float l0[NP], af[NP], df2[NP], bf4[NP], x[NP];
    unsigned live = 0;  // pairs still iterating
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const float af0 = (float)a[p];
        float l = fminf((float)half_g[p], sqrtf(3.0f * af0)) * 1.000001f;
        if (!(l > 0.0f)) l = 1.0f;  // S == 0: every root is 0; handled by the a == 0 test below
        l0[p] = l;
        const float sc = __frcp_rn(l);
        const float sc2 = sc * sc;
        af[p] = af0 * sc2;
        df2[p] = 2.0f * ((float)d[p] * sc2 * sc);
        bf4[p] = 4.0f * ((float)b[p] * sc2 * sc2);
        x[p] = 1.0f;
        live |= 1u << p;
    }
#pragma unroll 1
    for (int it = 0; it < 14 && live; ++it) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float t = fmaf(x[p], x[p], -af[p]);
            const float f = fmaf(t, t, fmaf(-4.0f * df2[p], x[p], -bf4[p]));
            const float fq = fmaf(x[p], t, -df2[p]);
            const float step = 0.25f * __fdividef(f, fq);
            const bool act = (live >> p) & 1u;
            const bool good = fq > 0.0f;
            if (act && good) x[p] -= step;
            // Stop once the step is small enough for the quadratic convergence of the FP64 step below to reach
            // round-off (a relative error e becomes ~e^2); testing against the FP32 noise floor (3e-7) let one noisy
            // lane keep its whole warp iterating.
            if (!good || fabsf(step) <= 2e-5f * x[p]) live &= ~(1u << p);
        }
    }
    unsigned ok = 0;
    double d2[NP], d8[NP], b4[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        d2[p] = 2.0 * d[p];
        d8[p] = 8.0 * d[p];
        b4[p] = 4.0 * b[p];
        double l = (double)(x[p] * l0[p]);  // FP32 product: only the starting point of the FP64 step
        const double t = fma(l, l, -a[p]);
        const double f = fma(t, t, -fma(d8[p], l, b4[p]));
        const double fq = fma(l, t, -d2[p]);
        const double step = f * (double)(0.25f * __frcp_rn((float)fq));
        l -= step;
        lam[p] = l;
        if (fabs(step) <= 1e-6 * l) ok |= 1u << p;
    }
    if (ok != (1u << NP) - 1u) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if ((ok >> p) & 1u) continue;
            if (a[p] == 0.0) { lam[p] = 0.0; ok |= 1u << p; continue; }
#pragma unroll 1
            for (int it = 0; it < 24; ++it) {
                const double t = fma(lam[p], lam[p], -a[p]);
                const double f = fma(t, t, -fma(d8[p], lam[p], b4[p]));
                const double fq = fma(lam[p], t, -d2[p]);
                const double st = f * (double)(0.25f * __frcp_rn((float)fq));
                lam[p] -= st;
                if (fabs(st) <= 1e-13 * lam[p]) { ok |= 1u << p; break; }
            }
        }
    }
    return ok;
}

// FP64 reciprocal / reciprocal square root approximations (about 23 bits; SASS MUFU.RCP64H / MUFU.RSQ64H): ONE
// special-function-unit operation instead of the F2F + MUFU + F2F of going through FP32, which matters where the
// quarter-rate XU pipe is shared by many warps doing the same solve (ozaki.cu).
__device__ __forceinline__ double rcp_approx_f64(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double rsqrt_approx_f64(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// All-FP64 variant of the lockstep root solve: Newton from the upper bound min((G_i+G_j)/2, sqrt(3a)) with the
// division replaced by rcp.approx.f64 (a step only needs ~1e-6 relative accuracy: the iteration is self-correcting),
// no FP32 phase, no conversions.  Typical trajectories start within a few per cent of the root: 4 steps.
// Acceptance must tell quadratic from linear convergence: at a (near-)double root — planar or collinear structures
// (sigma_3 = 0 makes sigma_1+sigma_2+-sigma_3 coincide), mirror images — the steps only halve, a step of 1e-7 then
// means an ERROR of 1e-7, and once the halving reaches the sqrt(eps) noise floor of a double root single steps
// collapse at random.  A pair settles when its step is below 1e-7 of the root AND quadratically smaller than the step
// before (step * x <= 16 * previous^2; on the very first step: below 1e-13 outright).  Pairs that never get there
// within 20 steps are reported through ok_mask and re-done by the explicit-rotation fix-up path.
template <int NP>
__device__ __forceinline__ unsigned largest_root_f64(const double (&a)[NP], const double (&b)[NP], const double (&d)[NP],
                                                     const double (&half_g)[NP], double (&lam)[NP]) {
    double x[NP], m8d[NP], m4b[NP], d2[NP];
    int e_last[NP];                   // biased exponent of the previous step (the tests below are integer-only: the
                                      // FP64 pipe is the busiest one in these epilogues)
    unsigned ok = 0;
    constexpr unsigned ALL = (1u << NP) - 1u;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double a3 = 3.0 * a[p];
        // '+ 1e-300' instead of fmax, '<' instead of fmin: same values on every input that matters (a3 = 0 stays 0, any
        // a3 >= 1e-284 is unchanged to the bit) without fmax/fmin's NaN-handling instruction sequences
        const double s = a3 * rsqrt_approx_f64(a3 + 1e-300) * 1.000001;   // >= sqrt(3a)
        const double x0 = half_g[p] < s ? half_g[p] : s;
        x[p] = x0 > 0.0 ? x0 : 0.0;   // S == 0 (zero padding): every root is 0 and the iteration stays there
        m8d[p] = -8.0 * d[p];
        m4b[p] = -4.0 * b[p];
        d2[p] = 2.0 * d[p];
        // "previous step" of the first step: 2^-22 x, so that only a first step below ~1e-12 x settles
        e_last[p] = ((__double2hiint(x[p]) >> 20) & 0x7ff) - 22;
        if (a[p] == 0.0) ok |= 1u << p;   // (a NaN covariance settles nowhere and goes to the fix-up path, which propagates it)
    }
#pragma unroll 1
    for (int it = 0; it < 20 && ok != ALL; ++it) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const double t = fma(x[p], x[p], -a[p]);
            const double f = fma(t, t, fma(m8d[p], x[p], m4b[p]));
            const double fq = fma(x[p], t, -d2[p]);                      // f' / 4
            const bool up = fq > 0.0;
            const double step = up ? f * (0.25 * rcp_approx_f64(fq)) : 0.0;
            x[p] -= step;                                                // a settled pair barely moves
            // |step| <= 2^-23 x  and  |step| x <= 16 previous^2, on the exponents (each within a factor 2)
            const int e_as = (__double2hiint(step) >> 20) & 0x7ff, e_x = (__double2hiint(x[p]) >> 20) & 0x7ff;
            if (up && e_as + 23 <= e_x && e_as + e_x <= 2 * e_last[p] + 4) ok |= 1u << p;
            e_last[p] = e_as;
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        lam[p] = x[p];
        if (!(x[p] == x[p])) ok &= ~(1u << p);
    }
    return ok;
}

// Fixed-trip variant for the epilogues that are bound by instruction issue (ozaki.cu): FOUR Newton steps from the same
// upper bound with no test inside the iteration (7 instructions per step and pair instead of ~25: the acceptance test of
// largest_root_f64 is a dozen integer operations per step), then ONE acceptance test per pair:
//   * the fourth step is below 2^-30 of the root (typical trajectories start within a per cent of it and reach round-off
//     in three steps), AND
//   * the root is well separated: f'(x) / 4 >= 2^-10 x^3.  f' = 8 (s2+s3)(s1+s3)(s1+s2) at the largest root, so this
//     refuses exactly the ill-conditioned cases -- near-collinear structures, mirror images of near-symmetric ones --
//     where Newton converges only linearly and a small step does not mean a small error.  With the separation bound the
//     error recurrence is e' <= 1250 e^2 / x, i.e. a step below 2^-30 x leaves an error below 1e-15 x.
// Pairs that fail the test continue in the tested loop of largest_root_f64 (same state); what that loop cannot settle is
// reported through ok_mask and re-done by the explicit-rotation fix-up path, as before.
template <int NP>
__device__ __forceinline__ unsigned largest_root_fast(const double (&a)[NP], const double (&b)[NP], const double (&d)[NP],
                                                      const double (&half_g)[NP], double (&lam)[NP]) {
    double x[NP], m8d[NP], m4b[NP], d2[NP], step[NP], fq[NP];
    constexpr unsigned ALL = (1u << NP) - 1u;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double a3 = 3.0 * a[p];
        const double s = a3 * rsqrt_approx_f64(a3 + 1e-300) * 1.000001;   // >= sqrt(3a)
        const double x0 = half_g[p] < s ? half_g[p] : s;
        x[p] = x0 > 0.0 ? x0 : 0.0;
        m8d[p] = -8.0 * d[p];
        m4b[p] = -4.0 * b[p];
        d2[p] = 2.0 * d[p];
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const double t = fma(x[p], x[p], -a[p]);
            const double f = fma(t, t, fma(m8d[p], x[p], m4b[p]));
            fq[p] = fma(x[p], t, -d2[p]);                                 // f' / 4
            const double q = 0.25 * rcp_approx_f64(fq[p]);
            if (it == 3) step[p] = f * q;
            x[p] = fma(-f, q, x[p]);
        }
    }
    unsigned ok = 0;
    int e_last[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int e_s = (__double2hiint(step[p]) >> 20) & 0x7ff, e_x = (__double2hiint(x[p]) >> 20) & 0x7ff;
        const double x3 = x[p] * x[p] * x[p];
        // (x = NaN or infinity -- S == 0 makes the very first quotient 0 / 0 -- has exponent 0x7ff and fails the first test)
        if (e_s + 30 <= e_x && e_x != 0x7ff && fq[p] >= 0.0009765625 * x3) ok |= 1u << p;
        e_last[p] = e_s;
        if (a[p] == 0.0) { x[p] = 0.0; ok |= 1u << p; }                   // zero covariance (1-atom frames): every root is 0
    }
    if (ok != ALL) {
        // the tested iteration of largest_root_f64, continued from where the four steps ended
        unsigned bad = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (!(x[p] == x[p]) || !(x[p] <= 1.7976931348623157e308)) {     // restart a pair that blew up from its upper bound
                const double a3 = 3.0 * a[p];
                const double s = a3 * rsqrt_approx_f64(a3 + 1e-300) * 1.000001;
                const double x0 = half_g[p] < s ? half_g[p] : s;
                x[p] = x0 > 0.0 ? x0 : 0.0;
                e_last[p] = ((__double2hiint(x[p]) >> 20) & 0x7ff) - 22;
                bad |= 1u << p;
            }
        (void)bad;
#pragma unroll 1
        for (int it = 0; it < 20 && ok != ALL; ++it) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const double t = fma(x[p], x[p], -a[p]);
                const double f = fma(t, t, fma(m8d[p], x[p], m4b[p]));
                const double fq2 = fma(x[p], t, -d2[p]);
                const bool up = fq2 > 0.0;
                const double st = (up && !((ok >> p) & 1u)) ? f * (0.25 * rcp_approx_f64(fq2)) : 0.0;   // accepted pairs stay put
                x[p] -= st;
                const int e_as = (__double2hiint(st) >> 20) & 0x7ff, e_x = (__double2hiint(x[p]) >> 20) & 0x7ff;
                if (up && e_as + 23 <= e_x && e_as + e_x <= 2 * e_last[p] + 4) ok |= 1u << p;
                e_last[p] = e_as;
            }
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        lam[p] = x[p];
        if (!(x[p] == x[p])) ok &= ~(1u << p);
    }
    return ok;
}

// Cyclic Jacobi on a symmetric 4x4 matrix (in registers, fully unrolled indices).  On return A holds the
// eigenvalues on its diagonal and the columns of V the eigenvectors.
__device__ __forceinline__ void jacobi4(double (&A)[4][4], double (&V)[4][4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[0][3] * A[0][3] + A[1][2] * A[1][2] +
                           A[1][3] * A[1][3] + A[2][3] * A[2][3];
        const double dia = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2] + A[3][3] * A[3][3];
        if (!(off > 1e-34 * dia)) break;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[p][q];
                if (apq != 0.0) {
                    const double theta = 0.5 * (A[q][q] - A[p][p]) / apq;
                    const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                    const double c = rsqrt(fma(t, t, 1.0));
                    const double s = t * c;
                    A[p][p] -= t * apq;
                    A[q][q] += t * apq;
                    A[p][q] = 0.0;
                    A[q][p] = 0.0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r != p && r != q) {
                            const double arp = A[r][p], arq = A[r][q];
                            const double np_ = c * arp - s * arq, nq_ = s * arp + c * arq;
                            A[r][p] = np_; A[p][r] = np_;
                            A[r][q] = nq_; A[q][r] = nq_;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const double vrp = V[r][p], vrq = V[r][q];
                        V[r][p] = c * vrp - s * vrq;
                        V[r][q] = s * vrp + c * vrq;
                    }
                }
            }
        }
    }
}

// Best proper rotation for covariance S[a][b] = sum_k p_k[a] q_k[b] (p = moving frame, q = reference
// frame).  Writes U (row-major, used as p_row * U like the reference's np.dot(P, U), distmat.py:181,223,
// 228) and returns lambda = s1 + s2 + sign(det S) s3.
__device__ __forceinline__ double kabsch_rotation(const double* s, double* U) {
    double A[4][4], V[4][4];
    const double Sxx = s[0], Sxy = s[1], Sxz = s[2], Syx = s[3], Syy = s[4], Syz = s[5], Szx = s[6], Szy = s[7],
                 Szz = s[8];
    A[0][0] = Sxx + Syy + Szz;
    A[1][1] = Sxx - Syy - Szz;
    A[2][2] = -Sxx + Syy - Szz;
    A[3][3] = -Sxx - Syy + Szz;
    A[0][1] = A[1][0] = Syz - Szy;
    A[0][2] = A[2][0] = Szx - Sxz;
    A[0][3] = A[3][0] = Sxy - Syx;
    A[1][2] = A[2][1] = Sxy + Syx;
    A[1][3] = A[3][1] = Szx + Sxz;
    A[2][3] = A[3][2] = Syz + Szy;
    jacobi4(A, V);
    double lam = A[0][0];
    double q0 = V[0][0], q1 = V[1][0], q2 = V[2][0], q3 = V[3][0];
#pragma unroll
    for (int c = 1; c < 4; ++c) {
        if (A[c][c] > lam) {
            lam = A[c][c];
            q0 = V[0][c]; q1 = V[1][c]; q2 = V[2][c]; q3 = V[3][c];
        }
    }
    const double nrm = rsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 *= nrm; q1 *= nrm; q2 *= nrm; q3 *= nrm;
    // R rotates column vectors p -> q; U = R^T acts on row vectors.
    const double r00 = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, r01 = 2.0 * (q1 * q2 - q0 * q3),
                 r02 = 2.0 * (q1 * q3 + q0 * q2);
    const double r10 = 2.0 * (q2 * q1 + q0 * q3), r11 = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3,
                 r12 = 2.0 * (q2 * q3 - q0 * q1);
    const double r20 = 2.0 * (q3 * q1 - q0 * q2), r21 = 2.0 * (q3 * q2 + q0 * q1),
                 r22 = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
    U[0] = r00; U[1] = r10; U[2] = r20;
    U[3] = r01; U[4] = r11; U[5] = r21;
    U[6] = r02; U[7] = r12; U[8] = r22;
    return lam;
}

// The same rotation without the Jacobi sweeps (the per-pair kernel computes three of them per pair, redundantly in all 32
// lanes: 13 % of its time).  lambda is the largest root of the quartic, by Newton from the upper bound sqrt(3a) in full
// FP64; the quaternion is the null vector of M = K - lambda I, taken as the column of adj(M) with the largest diagonal
// cofactor (adj(M) = q q^T prod(other eigenvalue gaps) for a simple eigenvalue).  Accepted only when the residual |M q| is
// at round-off level relative to |K| |q|; anything else -- a (near-)double largest eigenvalue: collinear structures, mirror
// images of near-symmetric ones, a stalled Newton iteration -- returns false and the caller runs the Jacobi solver.
__device__ __forceinline__ bool kabsch_rotation_adjugate(const double* s, double* U, double& lam_out) {
    double a, b, d;
    quartic_invariants(s, a, b, d);
    if (!(a > 0.0) || !(a <= 1.7976931348623157e308)) return false;
    double x = sqrt(3.0 * a) * 1.0000001;
    bool settled = false;
#pragma unroll 1
    for (int it = 0; it < 48; ++it) {
        const double t = fma(x, x, -a);
        const double f = fma(t, t, -fma(8.0 * d, x, 4.0 * b));
        const double fq = fma(x, t, -2.0 * d);
        if (!(fq > 0.0)) break;
        const double step = f * (0.25 * rcp_approx_f64(fq));   // 23-bit reciprocal: a Newton step is self-correcting
        x -= step;
        // a step below 1e-9 x leaves a quadratically smaller error (the root is simple, or the residual test below refuses)
        if (fabs(step) <= 1e-9 * x) {
            const double t2 = fma(x, x, -a);
            x -= fma(t2, t2, -fma(8.0 * d, x, 4.0 * b)) * (0.25 * rcp_approx_f64(fma(x, t2, -2.0 * d)));   // one more, for the last bits
            settled = true;
            break;
        }
    }
    if (!settled) return false;
    const double Sxx = s[0], Sxy = s[1], Sxz = s[2], Syx = s[3], Syy = s[4], Syz = s[5], Szx = s[6], Szy = s[7], Szz = s[8];
    // M = K - lambda I, symmetric: m00 m01 m02 m03 / m11 m12 m13 / m22 m23 / m33
    const double m00 = Sxx + Syy + Szz - x, m11 = Sxx - Syy - Szz - x, m22 = -Sxx + Syy - Szz - x, m33 = -Sxx - Syy + Szz - x;
    const double m01 = Syz - Szy, m02 = Szx - Sxz, m03 = Sxy - Syx, m12 = Sxy + Syx, m13 = Szx + Sxz, m23 = Syz + Szy;
    // 2x2 minors of rows (2,3) and of rows (0,1), shared by the cofactors
    const double r01 = m02 * m13 - m03 * m12, r02 = m02 * m23 - m03 * m22, r03 = m02 * m33 - m03 * m23;
    const double r12 = m12 * m23 - m13 * m22, r13 = m12 * m33 - m13 * m23, r23 = m22 * m33 - m23 * m23;
    const double t01 = m00 * m11 - m01 * m01, t02 = m00 * m12 - m01 * m02, t03 = m00 * m13 - m01 * m03;
    const double t12 = m01 * m12 - m11 * m02, t13 = m01 * m13 - m11 * m03, t23 = m02 * m13 - m12 * m03;
    // adjugate of the symmetric 4x4 (Laplace expansion over row pairs (0,1) | (2,3))
    const double c00 = m11 * r23 - m12 * r13 + m13 * r12;
    const double c01 = -(m01 * r23 - m12 * r03 + m13 * r02);
    const double c02 = m01 * r13 - m11 * r03 + m13 * r01;
    const double c03 = -(m01 * r12 - m11 * r02 + m12 * r01);
    const double c11 = m00 * r23 - m02 * r03 + m03 * r02;
    const double c12 = -(m00 * r13 - m01 * r03 + m03 * r01);
    const double c13 = m00 * r12 - m01 * r02 + m02 * r01;
    const double c22 = m33 * t01 - m13 * t03 + m03 * t13;
    const double c23 = -(m23 * t01 - m13 * t02 + m03 * t12);
    const double c33 = m22 * t01 - m12 * t02 + m02 * t12;
    (void)t23;
    // column with the largest diagonal cofactor
    double q0 = c00, q1 = c01, q2 = c02, q3 = c03, best = fabs(c00);
    if (fabs(c11) > best) { best = fabs(c11); q0 = c01; q1 = c11; q2 = c12; q3 = c13; }
    if (fabs(c22) > best) { best = fabs(c22); q0 = c02; q1 = c12; q2 = c22; q3 = c23; }
    if (fabs(c33) > best) { best = fabs(c33); q0 = c03; q1 = c13; q2 = c23; q3 = c33; }
    const double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
    if (!(n2 > 0.0) || !(n2 <= 1.7976931348623157e308)) return false;
    const double nrm = rsqrt(n2);
    q0 *= nrm; q1 *= nrm; q2 *= nrm; q3 *= nrm;
    // residual of the eigen-equation: round-off level or the Jacobi solver takes over
    const double e0 = m00 * q0 + m01 * q1 + m02 * q2 + m03 * q3, e1 = m01 * q0 + m11 * q1 + m12 * q2 + m13 * q3;
    const double e2 = m02 * q0 + m12 * q1 + m22 * q2 + m23 * q3, e3 = m03 * q0 + m13 * q1 + m23 * q2 + m33 * q3;
    const double res2 = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    if (!(res2 <= 1e-29 * a)) return false;     // |M q| <= 3e-15 |S|_F  (an exact eigenvector leaves ~1e-15 through lambda's last bit)
    // ... and the eigenvalue must be well separated (forward error ~ residual / gap): adj(M) = q q^T x (product of the
    // three gaps), so the largest diagonal cofactor measures the gaps against |S|^3
    if (!(best >= 1e-5 * a * sqrt(a))) return false;
    const double r00 = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, r01_ = 2.0 * (q1 * q2 - q0 * q3), r02_ = 2.0 * (q1 * q3 + q0 * q2);
    const double r10 = 2.0 * (q2 * q1 + q0 * q3), r11 = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, r12_ = 2.0 * (q2 * q3 - q0 * q1);
    const double r20 = 2.0 * (q3 * q1 - q0 * q2), r21 = 2.0 * (q3 * q2 + q0 * q1), r22 = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
    U[0] = r00; U[1] = r10; U[2] = r20;
    U[3] = r01_; U[4] = r11; U[5] = r21;
    U[6] = r02_; U[7] = r12_; U[8] = r22;
    lam_out = x;
    return true;
}

}  // namespace ct
