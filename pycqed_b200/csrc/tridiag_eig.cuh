// tridiag_eig.cuh -- lowest-k eigenpairs of a real symmetric tridiagonal matrix (d, e), executed
// cooperatively by one CTA.  This is the stage LAPACK's zheevr reaches through dstebz/dstein when
// the reference calls scipy.linalg.eigh(..., subset_by_index=(0, k-1)) (pyscqed/util.py:153).
//
//   eigenvalues : Sturm-count multisection, one warp per eigenvalue index, 32 trial shifts per
//                 round (interval shrinks 33x per round; ~11-13 rounds to 2 ulp)
//   eigenvectors: inverse iteration on T - lambda I with partial-pivoting tridiagonal LU
//                 (one thread per vector), then modified Gram-Schmidt across clustered
//                 eigenvalues (one warp), a purification pass and a second MGS so that exactly
//                 degenerate levels (floating CSFQ doublets) come out as
//                 an orthonormal basis of their eigenspace.
#pragma once
#include "common.cuh"

namespace scqed {


// 1/t to ~1 ulp: hardware seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps.  |t| >= pivmin
// (a normal number) is guaranteed by the caller, so neither the seed nor the refinement overflows.
// An IEEE division expands to ~25 instructions; the Sturm recurrence is nothing but divisions.
__device__ __forceinline__ double fast_rcp(double t) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));
    double e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// number of eigenvalues of T that are < x  (LAPACK dlaebz / dstebz recurrence)
__device__ __forceinline__ int sturm_count(const double *__restrict__ d, const double *__restrict__ e2,
                                           int n, double x, double pivmin) {
    double t = d[0] - x;
    if (fabs(t) < pivmin) t = -pivmin;
    int c = t <= 0.0;
    for (int j = 1; j < n; ++j) {
        const double q = e2[j - 1];
        t = d[j] - (q == 0.0 ? 0.0 : q * fast_rcp(t)) - x;
        if (fabs(t) < pivmin) t = -pivmin;
        c += (t <= 0.0);
    }
    return c;
}

// Same count from the three-term recurrence of the leading principal minors,
//     p_0 = 1,  p_1 = d_0 - x,  p_{j+1} = (d_j - x) p_j - e_{j-1}^2 p_{j-1},
// as the number of sign changes in (p_0 ... p_n), a zero taking the sign opposite to its predecessor (the same
// convention as t <= 0 above).  No division: 3 FP64 operations per step instead of ~10, which is what the
// one-thread-per-eigenvalue bisection is bound by (FP64 pipe 54 % active).  The pair (p_j, p_{j-1}) is rescaled by
// 2^-+332 whenever |p_j| leaves [2^-332, 2^332] (checked every 4 steps; a step changes the magnitude by at most
// ~|d - x| + e^2, far below 2^83 for any matrix whose entries are below ~1e12).
__device__ __forceinline__ int sturm_count_prod(const double *__restrict__ d, const double *__restrict__ e2,
                                                int n, double x) {
    const double BIG = 8.749002899132047e+99, SMALL = 1.1429873912822749e-100;    // 2^332, 2^-332
    double pm = 1.0, p = d[0] - x;
    int neg = p < 0.0 || p == 0.0;          // "sign" of p_1 relative to p_0 = +1 (zero counts as a change)
    int c = neg;
    int j = 1;
    for (; j + 4 <= n; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double pn = fma(d[j + u] - x, p, -(e2[j + u - 1] * pm));
            pm = p; p = pn;
            const int ng = pn < 0.0 || (pn == 0.0 && !neg);
            c += ng != neg;
            neg = ng;
        }
        const double ap = fabs(p);
        if (ap > BIG) { p *= SMALL; pm *= SMALL; }
        else if (ap < SMALL && ap > 0.0) { p *= BIG; pm *= BIG; }
    }
    for (; j < n; ++j) {
        const double pn = fma(d[j] - x, p, -(e2[j - 1] * pm));
        pm = p; p = pn;
        const int ng = pn < 0.0 || (pn == 0.0 && !neg);
        c += ng != neg;
        neg = ng;
    }
    return c;
}

// Block-wide: Gershgorin interval and pivmin of T.  out[0]=gl, out[1]=gu, out[2]=pivmin,
// out[3]=||T||_1-ish (max abs row sum).  `red` is scratch for 3*32 doubles.
static __device__ void tridiag_bounds(const double *d, const double *e, int n, double *out, double *red) {
    double lo = 1e300, hi = -1e300, emax = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < n - 1 ? fabs(e[i]) : 0.0);
        lo = fmin(lo, d[i] - r);
        hi = fmax(hi, d[i] + r);
        if (i < n - 1) emax = fmax(emax, e[i] * e[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) { red[w] = lo; red[32 + w] = hi; red[64 + w] = emax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < nw; ++i) {
            lo = fmin(lo, red[i]); hi = fmax(hi, red[32 + i]); emax = fmax(emax, red[64 + i]);
        }
        double tn = fmax(fabs(lo), fabs(hi));
        double pad = 2.0 * tn * SCQED_EPS * n + 2.0 * SCQED_SAFMIN;
        out[0] = lo - pad;
        out[1] = hi + pad;
        out[2] = SCQED_SAFMIN * fmax(1.0, emax);
        out[3] = tn;
    }
    __syncthreads();
}

// Block-wide multisection for eigenvalue indices [0, k).  e2[i] = e[i]^2.  lam[k] out.
// Returns (per thread) nonzero if some interval failed to converge.
static __device__ int tridiag_bisect(const double *d, const double *e2, int n, int k, const double *bounds,
                              double *lam) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    const double gl = bounds[0], gu = bounds[1], pivmin = bounds[2];
    // absolute floor like LAPACK dstebz (ABSTOL = ulp * ||T||), a little tighter
    const double atol = 0.0625 * SCQED_EPS * bounds[3] + 2.0 * pivmin;
    int bad = 0;
    for (int j = w; j < k; j += nw) {
        double lo = gl, hi = gu;
        int it = 0;
        for (; it < 64; ++it) {
            double width = hi - lo;
            double tol = 2.0 * SCQED_EPS * fmax(fabs(lo), fabs(hi)) + atol;
            if (width <= tol) break;
            double x = lo + width * ((double)(l + 1) / 33.0);
            int c = sturm_count(d, e2, n, x, pivmin);
            unsigned above = __ballot_sync(0xffffffffu, c >= j + 1);   // lambda_j < x
            double nlo, nhi;
            if (above == 0u) {
                nlo = __shfl_sync(0xffffffffu, x, 31); nhi = hi;
            } else {
                int first = __ffs(above) - 1;
                nhi = __shfl_sync(0xffffffffu, x, first);
                double xl = __shfl_sync(0xffffffffu, x, first > 0 ? first - 1 : 0);
                nlo = first > 0 ? xl : lo;
            }
            if (!(nhi - nlo < width)) { lo = nlo; hi = nhi; break; }   // no progress: at resolution
            lo = nlo; hi = nhi;
        }
        if (it >= 64) bad = 1;
        if (l == 0) lam[j] = 0.5 * (lo + hi);
    }
    return bad;
}

// Workspace of one inverse iteration (per vector): 4 n doubles + n flags.
struct TriLU {
    double *u0, *u1, *u2, *mul;
    unsigned char *swp;
    int st;      // element stride (1: private contiguous arrays; G: arrays interleaved over G threads)
};

// The three routines below walk long arrays that may live in global memory (n up to ~10^4 on the
// HBM dense path); every loop loads the operands of iteration i+1 before the dependent arithmetic
// of iteration i, so the critical path is the division chain, not the memory latency.
__device__ __forceinline__ void trilu_factor(const double *__restrict__ d, const double *__restrict__ e,
                                             int n, double lambda, double eps3, TriLU W) {
    double u = d[0] - lambda;
    double v = n > 1 ? e[0] : 0.0;
    double ei_n = n > 1 ? e[0] : 0.0, di_n = n > 1 ? d[1] : 0.0, en_n = n > 2 ? e[1] : 0.0;
    for (int i = 1; i < n; ++i) {
        const double ei = ei_n, di = di_n - lambda, en = en_n;
        if (i + 1 < n) { ei_n = e[i]; di_n = d[i + 1]; en_n = i + 2 < n ? e[i + 1] : 0.0; }
        if (u == 0.0 && ei == 0.0) u = eps3;
        if (fabs(ei) >= fabs(u)) {          // pivot on row i
            double xu = u / ei;
            W.mul[(size_t)(i) * W.st] = xu; W.swp[(size_t)(i) * W.st] = 1;
            W.u0[(size_t)(i - 1) * W.st] = ei; W.u1[(size_t)(i - 1) * W.st] = di; W.u2[(size_t)(i - 1) * W.st] = en;
            u = v - xu * di;
            v = -xu * en;
        } else {
            double xu = ei / u;
            W.mul[(size_t)(i) * W.st] = xu; W.swp[(size_t)(i) * W.st] = 0;
            W.u0[(size_t)(i - 1) * W.st] = u; W.u1[(size_t)(i - 1) * W.st] = v; W.u2[(size_t)(i - 1) * W.st] = 0.0;
            u = di - xu * v;
            v = en;
        }
    }
    if (u == 0.0) u = eps3;
    W.u0[(size_t)(n - 1) * W.st] = u; W.u1[(size_t)(n - 1) * W.st] = 0.0; W.u2[(size_t)(n - 1) * W.st] = 0.0;
}

// z <- U^-1 L^-1 P z   (apply_l = 0 skips the L^-1 P part: first iteration from a random start)
__device__ __forceinline__ void trilu_solve(int n, TriLU W, double *z, int zs, double eps3, bool apply_l) {
    if (apply_l && n > 1) {
        double cur = z[0];
        double m_n = W.mul[(size_t)(1) * W.st], z_n = z[zs];
        unsigned char s_n = W.swp[(size_t)(1) * W.st];
        for (int i = 1; i < n; ++i) {
            const double m_i = m_n, b = z_n;
            const unsigned char s_i = s_n;
            if (i + 1 < n) { m_n = W.mul[(size_t)(i + 1) * W.st]; s_n = W.swp[(size_t)(i + 1) * W.st]; z_n = z[(size_t)(i + 1) * zs]; }
            if (s_i) { z[(size_t)(i - 1) * zs] = b; cur = cur - m_i * b; }
            else     { z[(size_t)(i - 1) * zs] = cur; cur = b - m_i * cur; }
        }
        z[(size_t)(n - 1) * zs] = cur;
    }
    double z1 = 0.0, z2 = 0.0;
    double p_n = W.u0[(size_t)(n - 1) * W.st], u1_n = W.u1[(size_t)(n - 1) * W.st], u2_n = W.u2[(size_t)(n - 1) * W.st], r_n = z[(size_t)(n - 1) * zs];
    for (int i = n - 1; i >= 0; --i) {
        double p = p_n;
        const double u1 = u1_n, u2 = u2_n, r = r_n;
        if (i > 0) { p_n = W.u0[(size_t)(i - 1) * W.st]; u1_n = W.u1[(size_t)(i - 1) * W.st]; u2_n = W.u2[(size_t)(i - 1) * W.st]; r_n = z[(size_t)(i - 1) * zs]; }
        if (fabs(p) < eps3 * 1e-3) p = copysign(eps3, p == 0.0 ? 1.0 : p);
        double zi = (r - u1 * z1 - u2 * z2) / p;
        z[(size_t)i * zs] = zi;
        z2 = z1; z1 = zi;
    }
}

__device__ __forceinline__ double lcg_uniform(unsigned long long &s) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((double)(s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

// scale z (stride zs) to unit 2-norm; returns the norm before scaling
__device__ __forceinline__ double normalize_serial(int n, double *z, int zs) {
    double big = 0.0;
    for (int i = 0; i < n; ++i) big = fmax(big, fabs(z[i * zs]));
    if (big == 0.0) return 0.0;
    double s = 0.0, inv = 1.0 / big;
    for (int i = 0; i < n; ++i) { double t = z[i * zs] * inv; s += t * t; }
    double nrm = big * sqrt(s);
    double sc = 1.0 / nrm;
    for (int i = 0; i < n; ++i) z[i * zs] *= sc;
    return nrm;
}

// One warp: modified Gram-Schmidt of the vectors of each eigenvalue cluster.
// Z: vector j element i at Z[j * zld + i * zs].
static __device__ void cluster_mgs_warp(int n, int k, const double *lam, double ortol, double *Z, int zld, int zs) {
    const int l = threadIdx.x & 31;
    int start = 0;
    for (int j = 0; j < k; ++j) {
        if (j > 0 && fabs(lam[j] - lam[j - 1]) > ortol) start = j;
        double *zj = Z + (size_t)j * zld;
        for (int q = start; q < j; ++q) {
            const double *zq = Z + (size_t)q * zld;
            double dot = 0.0;
            for (int i = l; i < n; i += 32) dot += zq[i * zs] * zj[i * zs];
            dot = warp_sum(dot);
            for (int i = l; i < n; i += 32) zj[i * zs] -= dot * zq[i * zs];
            __syncwarp();
        }
        double s = 0.0;
        for (int i = l; i < n; i += 32) s += zj[i * zs] * zj[i * zs];
        s = warp_sum(s);
        double sc = s > 0.0 ? rsqrt(s) : 0.0;
        // one Newton step: rsqrt is not correctly rounded
        sc = sc * (1.5 - 0.5 * s * sc * sc);
        for (int i = l; i < n; i += 32) zj[i * zs] *= sc;
        __syncwarp();
    }
}

// Block-wide inverse iteration for the k eigenvalues lam[] of T.  Vector j is written (real) to
// Z[j*zld + i*zs].  lu_ws: workspace for `vb` simultaneous vectors, each 4n doubles + n bytes
// (n bytes padded to 8).  Returns nonzero on a vector that failed to grow.
static __device__ int tridiag_inverse_iteration(const double *d, const double *e, int n, int k, const double *lam,
                                         double tnorm, double *Z, int zld, int zs, double *lu_ws, int vb,
                                         unsigned long long seed) {
    const double eps3 = fmax(SCQED_EPS * tnorm, SCQED_SAFMIN * 1e16);
    const size_t per = (size_t)4 * n + (size_t)((n + 7) / 8);
    const double ortol = 1e-3 * tnorm;   // LAPACK dstein
    int bad = 0;
    // A purification pass + Gram-Schmidt is only needed when two of the wanted eigenvalues are
    // closer than LAPACK dstein's ortol; well separated spectra take the short path.
    bool clustered = false;
    for (int j = 1; j < k; ++j) clustered |= (fabs(lam[j] - lam[j - 1]) <= ortol);
    const int npass = clustered ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
        for (int base = 0; base < k; base += vb) {
            int j = base + (int)threadIdx.x;
            if ((int)threadIdx.x < vb && j < k) {
                double *w = lu_ws + per * threadIdx.x;
                TriLU W = { w, w + n, w + 2 * n, w + 3 * n, (unsigned char *)(w + 4 * n), 1 };
                double *z = Z + (size_t)j * zld;
                // separate numerically coincident shifts a little (dstein's pertol idea)
                double lj = lam[j];
                if (j > 0 && fabs(lj - lam[j - 1]) <= 10.0 * SCQED_EPS * fabs(lj)) lj += 10.0 * eps3 * (j & 1 ? 1.0 : 0.5);
                trilu_factor(d, e, n, lj, eps3, W);
                if (pass == 0) {
                    unsigned long long s = seed * 0x9E3779B97F4A7C15ULL + (unsigned long long)(j + 1) * 0xD1B54A32D192ED03ULL;
                    for (int i = 0; i < n; ++i) z[i * zs] = lcg_uniform(s);
                    trilu_solve(n, W, z, zs, eps3, false);
                    double g = normalize_serial(n, z, zs);
                    for (int it = 0; it < 2; ++it) {
                        trilu_solve(n, W, z, zs, eps3, true);
                        g = normalize_serial(n, z, zs);
                    }
                    // growth of a unit vector under (T - lambda)^-1 should be ~ 1/eps3-ish
                    if (!(g * eps3 * n > 1e-6)) bad = 1;
                } else {
                    trilu_solve(n, W, z, zs, eps3, true);
                    normalize_serial(n, z, zs);
                }
            }
            __syncthreads();
        }
        if (clustered) {
            if (threadIdx.x < 32) cluster_mgs_warp(n, k, lam, ortol, Z, zld, zs);
            __syncthreads();
        }
    }
    return bad;
}

}  // namespace scqed
