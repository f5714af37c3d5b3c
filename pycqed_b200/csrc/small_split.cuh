// small_split.cuh -- S1 (n <= ~110) as a pipeline of phase kernels, each launched with the
// parallelism its phase has.  H(p) still lives only in shared memory; what crosses HBM between the
// phases is O(n) per point (d, e, tau, eigenvalues) plus the Householder reflectors when
// eigenvectors are requested (n^2 values per point, written once, read once).
//
//   s1_tridiag_kernel   1 CTA per point (2 per SM on the real-symmetric path): assemble (K2) +
//                       Householder tridiagonalisation (small_tridiag, small_eigh.cuh)
//   s1_bisect_kernel    1 warp per (point, eigenvalue): Sturm multisection
//   s1_invit_kernel     1 thread per (point, vector): tridiagonal inverse iteration, workspace
//                       interleaved over threads so that every access is coalesced
//   s1_backtr_kernel    1 warp per (point, group of 8 vectors): x = Q z, vectors held in
//                       registers, each reflector read once per group
//   s1_resonator_kernel 1 CTA per point: RWA strips (K7)
//
// The monolithic small_eigh_kernel did all of this in one CTA per point and spent most of its time
// in phases that keep 10 threads or 10 warps of 512 threads busy (ncu: 37 % barrier stalls).
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "resonator.cuh"
#include "small_eigh.cuh"

namespace scqed {

struct S1Ws {                 // device workspace for a chunk of P points
    double *dd, *ee, *e2;     // [P][n]
    double *bnd;              // [P][4]
    cplx *tau;                // [P][n]   (real mode: doubles in the same slots)
    void *Vst;                // [P][n*n] T   reflectors (column-major copy of A after the reduction)
    double *lam;              // [P][k]
    double *lu;               // [4][n][G] interleaved
    unsigned char *swp;       // [n][G]
    double *zw;               // [n][G]    interleaved tridiagonal eigenvectors
    int *flags;               // [P] 1 = point is not real (real mode only)
};

struct S1Args {
    PlanDev P;
    const cplx *coef;
    const double *dcoef;       // [n_pts][2 * n_dense] dense-table coefficients (alpha | beta) or NULL
    const cplx *Hin;
    long long n_pts;
    int n, k, tidyup, want_vec;
    S1Ws W;
    int *info;
};

// ---- phase A -----------------------------------------------------------------------------------------
template <typename T> struct S1Smem {
    static size_t bytes(int n, int n_terms, int nthreads) {
        size_t o = (size_t)n * n * sizeof(T);
        o = (o + 15) & ~(size_t)15;
        o += (size_t)n * 8 * 2 + 16;                    // dv, ev
        o += (size_t)n * sizeof(T) + 16;                // tau
        o += (size_t)4 * n * sizeof(T) + 32;            // vbuf, wbuf
        (void)nthreads;
        size_t part = (size_t)SMALL_SEGS * n * sizeof(T);
        size_t cf = (size_t)n_terms * 16;
        o += (part > cf ? part : cf) + 16;
        o += 96 * 8 + 4 * 8 + 64 + 1024;                // red, bounds, static shared of small_tridiag
        return o;
    }
};

template <typename T>
static __global__ void __launch_bounds__(sizeof(T) == sizeof(double) ? 256 : 512)
s1_tridiag_kernel(S1Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, tid = threadIdx.x, NT = blockDim.x;
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char *r = smem + o; o += (bytes + 15) & ~(size_t)15; return r; };
    T *A = (T *)take((size_t)n * n * sizeof(T));
    double *dv = (double *)take((size_t)n * 8);
    double *ev = (double *)take((size_t)n * 8);
    T *tau = (T *)take((size_t)n * sizeof(T));
    T *vbuf = (T *)take((size_t)2 * n * sizeof(T));
    T *wbuf = (T *)take((size_t)2 * n * sizeof(T));
    size_t partb = (size_t)SMALL_SEGS * n * sizeof(T), cfb = (size_t)(a.P.n_terms > a.P.n_dense ? a.P.n_terms : a.P.n_dense) * 16;
    T *part = (T *)take(partb > cfb ? partb : cfb);
    double *red = (double *)take(96 * 8);
    double *bounds = (double *)take(4 * 8);
    constexpr bool REAL = sizeof(T) == sizeof(double);

    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        int not_real = 0;
        if (a.Hin) {
            const cplx *Hp = a.Hin + (size_t)p * n * n;
            for (int idx = tid; idx < n * n; idx += NT) {
                int i = idx / n, j = idx - i * n;              // row-major source, coalesced
                cplx v = Hp[idx];
                if constexpr (REAL) { not_real |= (v.y != 0.0); ((double *)A)[i + j * n] = v.x; }
                else ((cplx *)A)[i + j * n] = v;
            }
        } else if (a.dcoef) {
            // dense real term table: H[i][j] = sum_g (alpha_g + i beta_g) G_g[i][j] on the lower
            // triangle, mirrored by Hermiticity.  One warp per column, lanes down the column:
            // every G_g load is a coalesced 8-byte read of the (L2-resident) table.
            const int NG = a.P.n_dense;
            double *s_dc = (double *)part;
            for (int t = tid; t < 2 * NG; t += NT) s_dc[t] = a.dcoef[(size_t)p * 2 * NG + t];
            __syncthreads();
            double imw = 0.0;
            // (diagonal matrices cannot make a Hermitian H complex: only the full ones count)
            for (int g = 0; g < NG - a.P.n_dense_diag; ++g) imw += fabs(s_dc[NG + g]) * __ldg(a.P.dense_gmax + g);
            // tidyup zeroes |Im| < 1e-12 (numerical_system.py:19,594): below that the point is real
            const bool pt_real = a.tidyup ? imw < 1e-12 : imw == 0.0;
            if (REAL && !pt_real) not_real = 1;
            else {
                const size_t nn = (size_t)n * n;
                const int wp = tid >> 5, lane = tid & 31, NW = NT >> 5;
                const int NF = NG - a.P.n_dense_diag;                  // full matrices, then diagonal ones
                const double *dg = a.P.dense_tab + (size_t)NF * nn;
                for (int j = wp; j < n; j += NW) {
                    const double *col = a.P.dense_tab + (size_t)j * n;
                    for (int i = j + lane; i < n; i += 32) {
                        double re = 0.0, im = 0.0;
                        for (int g = 0; g < NF; ++g) {
                            const double gv = __ldg(col + g * nn + i);
                            re = fma(s_dc[g], gv, re);
                            if (!REAL && !pt_real) im = fma(s_dc[NG + g], gv, im);
                        }
                        if (i == j)
                            for (int g = NF; g < NG; ++g) re = fma(s_dc[g], __ldg(dg + (size_t)(g - NF) * n + i), re);
                        if (a.tidyup) { re = fabs(re) < 1e-12 ? 0.0 : re; im = fabs(im) < 1e-12 ? 0.0 : im; }
                        if (i == j) im = 0.0;
                        if constexpr (REAL) { ((double *)A)[i + j * n] = re; ((double *)A)[j + i * n] = re; }
                        else { ((cplx *)A)[i + j * n] = cmake(re, im); ((cplx *)A)[j + i * n] = cmake(re, -im); }
                    }
                }
            }
        } else {
            cplx *s_coef = (cplx *)part;
            for (int t = tid; t < a.P.n_terms; t += NT) s_coef[t] = a.coef[p * a.P.n_terms + t];
            __syncthreads();
            for (int idx = tid; idx < n * n; idx += NT) {
                int j = idx / n, i = idx - j * n;               // column-major destination
                int id[SCQED_MAX_NODES], jd[SCQED_MAX_NODES];
                digits(a.P, i, id);
                digits(a.P, j, jd);
                // H is Hermitian: H[i][j] = conj(H[j][i]).  Evaluating element (j, i) makes the
                // factor-table reads of consecutive threads (consecutive i) contiguous.
                cplx v = cconj(h_element(a.P, s_coef, jd, id));
                if (a.tidyup) v = tidy(v, 1e-12);
                if constexpr (REAL) { not_real |= (v.y != 0.0); ((double *)A)[idx] = v.x; }
                else ((cplx *)A)[idx] = v;
            }
        }
        not_real = __syncthreads_or(not_real);
        if (REAL) {
            if (tid == 0) { a.W.flags[p] = not_real; if (not_real) a.info[p] = -2; }
            if (not_real) continue;                              // info = -2: caller re-runs in complex mode
        }
        small_tridiag<T>(A, n, dv, ev, tau, vbuf, wbuf, part);
        if (tid == 0) ev[n - 1] = 0.0;
        __syncthreads();
        tridiag_bounds(dv, ev, n, bounds, red);
        for (int i = tid; i < n; i += NT) {
            double e = ev[i];
            a.W.dd[p * n + i] = dv[i];
            a.W.ee[p * n + i] = e;
            a.W.e2[p * n + i] = e * e;
            ((T *)a.W.tau)[p * n + i] = tau[i];
        }
        if (tid < 4) a.W.bnd[p * 4 + tid] = bounds[tid];
        if (a.want_vec) {
            T *dst = (T *)a.W.Vst + (size_t)p * n * n;
            for (int idx = tid; idx < n * n; idx += NT) dst[idx] = A[idx];
        }
        __syncthreads();
    }
}

// ---- phase A': H(p) is already Hermitian tridiagonal (single-node charge basis: transmon, Averin coupler) --------
// No Householder stage at all.  From the dense term table: d_i = H[i][i], s_i = H[i+1][i]; the unitary diagonal
// similarity D = diag(ph), ph_0 = 1, ph_{i+1} = ph_i s_i / |s_i| turns H into the REAL symmetric tridiagonal
// T = D^H H D with off-diagonals |s_i| (same spectrum; eigenvectors v = D z).  One warp per point: lanes
// stride over i, Gershgorin bounds by shuffles, the phase prefix product serially per 32-chunk.
static __global__ void __launch_bounds__(128)
s1_direct_tridiag_kernel(S1Args a) {
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = threadIdx.x & 31, n = a.n;
    if (p >= a.n_pts) return;
    const PlanDev &P = a.P;
    const int NG = P.n_dense, NF = NG - P.n_dense_diag;
    const double *dc = a.dcoef + (size_t)p * 2 * NG;
    const size_t nn = (size_t)n * n;
    const double *dg = P.dense_tab + (size_t)NF * nn;
    double lo = 1e300, hi = -1e300, emax = 0.0;
    double prev_e = 0.0;                    // |s_{i-1}| of the lane's first element is fetched by shuffle
    cplx carry = cmake(1.0, 0.0);           // phase of element (32 q) of this chunk
    cplx *ph = a.W.tau + p * n;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + l;
        double d = 0.0, e = 0.0;
        cplx u = cmake(1.0, 0.0);           // s_i / |s_i|
        if (i < n) {
            double sre = 0.0, sim = 0.0;
            for (int g = 0; g < NF; ++g) {
                d = fma(dc[g], __ldg(P.dense_tab + g * nn + (size_t)i * n + i), d);
                if (i + 1 < n) {
                    const double gv = __ldg(P.dense_tab + g * nn + (size_t)i * n + i + 1);      // G[i+1][i]
                    sre = fma(dc[g], gv, sre);
                    sim = fma(dc[NG + g], gv, sim);
                }
            }
            for (int g = NF; g < NG; ++g) d = fma(dc[g], __ldg(dg + (size_t)(g - NF) * n + i), d);
            if (a.tidyup) {
                d = fabs(d) < 1e-12 ? 0.0 : d;
                sre = fabs(sre) < 1e-12 ? 0.0 : sre;
                sim = fabs(sim) < 1e-12 ? 0.0 : sim;
            }
            e = sqrt(sre * sre + sim * sim);
            if (e > 0.0) u = cmake(sre / e, sim / e);
            a.W.dd[p * n + i] = d;
            a.W.ee[p * n + i] = e;
            a.W.e2[p * n + i] = e * e;
        }
        // Gershgorin: |e_{i-1}| + |e_i|
        double eprev = __shfl_up_sync(0xffffffffu, e, 1);
        if (l == 0) eprev = prev_e;
        prev_e = __shfl_sync(0xffffffffu, e, 31);
        if (i < n) {
            const double r = (i > 0 ? eprev : 0.0) + (i < n - 1 ? e : 0.0);
            lo = fmin(lo, d - r); hi = fmax(hi, d + r);
            emax = fmax(emax, e * e);
        }
        // inclusive prefix product of the unit phases inside the chunk (Hillis-Steele), then apply the carry:
        // ph_{i+1} = carry * u_{i0} ... u_i
        cplx pref = u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            cplx other = cmake(__shfl_up_sync(0xffffffffu, pref.x, o), __shfl_up_sync(0xffffffffu, pref.y, o));
            if (l >= o) pref = cmul(pref, other);
        }
        cplx excl = cmake(__shfl_up_sync(0xffffffffu, pref.x, 1), __shfl_up_sync(0xffffffffu, pref.y, 1));
        if (l == 0) excl = cmake(1.0, 0.0);
        if (i < n) {
            cplx v = cmul(carry, excl);
            const double nv = 1.0 / sqrt(v.x * v.x + v.y * v.y);      // keep |ph| = 1 against rounding drift
            ph[i] = cmake(v.x * nv, v.y * nv);
        }
        cplx last = cmake(__shfl_sync(0xffffffffu, pref.x, 31), __shfl_sync(0xffffffffu, pref.y, 31));
        carry = cmul(carry, last);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    if (l == 0) {                           // same conventions as tridiag_bounds
        double tn = fmax(fabs(lo), fabs(hi));
        double pad = 2.0 * tn * SCQED_EPS * n + 2.0 * SCQED_SAFMIN;
        a.W.bnd[p * 4 + 0] = lo - pad;
        a.W.bnd[p * 4 + 1] = hi + pad;
        a.W.bnd[p * 4 + 2] = fmax(SCQED_SAFMIN * fmax(1.0, emax), SCQED_SAFMIN);
        a.W.bnd[p * 4 + 3] = tn;
        a.W.flags[p] = 0;
    }
}

// v = D z: eigenvectors of the tridiagonal H from those of T.  One thread per (point, vector, i).
static __global__ void __launch_bounds__(256)
s1_phase_backtr_kernel(S1Args a, long long G, cplx *__restrict__ evecs) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.n, k = a.k;
    if (idx >= a.n_pts * k * n) return;
    const long long pv = idx / n;                 // p * k + j
    const int i = (int)(idx - pv * n);
    const long long p = pv / k;
    const double z = a.W.zw[(size_t)i * G + pv];
    const cplx ph = a.W.tau[p * n + i];
    evecs[idx] = cmake(z * ph.x, z * ph.y);
}

// ---- phase B: one warp per (point, eigenvalue) -----------------------------------------------------------
static __global__ void __launch_bounds__(256)
s1_bisect_kernel(S1Args a, const int *__restrict__ skip, double *__restrict__ evals) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = threadIdx.x & 31, n = a.n, k = a.k;
    if (gw >= a.n_pts * k) return;
    const long long p = gw / k;
    const int j = (int)(gw - p * k);
    if (skip && skip[p]) return;
    const double *d = a.W.dd + p * n, *e2 = a.W.e2 + p * n, *bnd = a.W.bnd + p * 4;
    const double gl = bnd[0], gu = bnd[1], pivmin = bnd[2];
    const double atol = 0.0625 * SCQED_EPS * bnd[3] + 2.0 * pivmin;
    double lo = gl, hi = gu;
    int it = 0;
    for (; it < 64; ++it) {
        double width = hi - lo;
        double tol = 2.0 * SCQED_EPS * fmax(fabs(lo), fabs(hi)) + atol;
        if (width <= tol) break;
        double x = lo + width * ((double)(l + 1) / 33.0);
        int c = sturm_count(d, e2, n, x, pivmin);
        unsigned above = __ballot_sync(0xffffffffu, c >= j + 1);
        double nlo, nhi;
        if (above == 0u) { nlo = __shfl_sync(0xffffffffu, x, 31); nhi = hi; }
        else {
            int first = __ffs(above) - 1;
            nhi = __shfl_sync(0xffffffffu, x, first);
            double xl = __shfl_sync(0xffffffffu, x, first > 0 ? first - 1 : 0);
            nlo = first > 0 ? xl : lo;
        }
        if (!(nhi - nlo < width)) { lo = nlo; hi = nhi; break; }
        lo = nlo; hi = nhi;
    }
    if (l == 0) {
        double v = 0.5 * (lo + hi);
        a.W.lam[p * k + j] = v;
        evals[p * k + j] = v;
        if (it >= 64) atomicOr(a.info + p, 1);
    }
}

// ---- phase B': one THREAD per (point, eigenvalue) -- plain Sturm bisection ------------------------------
// When the sweep has tens of thousands of (point, eigenvalue) pairs there is no need to spend 32
// lanes per pair on trial shifts: binary bisection needs ~53 Sturm counts per eigenvalue against
// 32 x 11 for the 33-way multisection (6.6x fewer FP64 operations); the longer dependent chain is
// hidden by the other warps.  The k eigenvalues of a point sit in adjacent lanes so that the loads
// of d and e^2 are broadcasts.  Same count recurrence, tolerances and stopping rule as dstebz.
static __global__ void __launch_bounds__(128)
s1_bisect_thread_kernel(S1Args a, const int *__restrict__ skip, double *__restrict__ evals, int prod) {
    const int k = a.k, n = a.n, l = threadIdx.x & 31;
    const int ppw = 32 / k;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int slot = l / k, j = l - slot * k;
    const long long p = warp * ppw + slot;
    if (!(slot < ppw && p < a.n_pts) || (skip && skip[p])) return;
    const double *__restrict__ d = a.W.dd + p * n, *__restrict__ e2 = a.W.e2 + p * n, *bnd = a.W.bnd + p * 4;
    const double pivmin = bnd[2];
    const double atol = 0.0625 * SCQED_EPS * bnd[3] + 2.0 * pivmin;
    double lo = bnd[0], hi = bnd[1];
    int it = 0;
    for (; it < 128; ++it) {
        const double width = hi - lo;
        const double tol = 2.0 * SCQED_EPS * fmax(fabs(lo), fabs(hi)) + atol;
        if (width <= tol) break;
        const double x = lo + 0.5 * width;
        if (!(x > lo && x < hi)) break;                       // at resolution
        if ((prod ? sturm_count_prod(d, e2, n, x) : sturm_count(d, e2, n, x, pivmin)) >= j + 1) hi = x; else lo = x;
    }
    const double v = 0.5 * (lo + hi);
    a.W.lam[p * k + j] = v;
    evals[p * k + j] = v;
    if (it >= 128) atomicOr(a.info + p, 1);
}

// ---- phase C: one thread per (point, vector); the k vectors of a point sit in one warp -------------------
static __global__ void __launch_bounds__(128)
s1_invit_kernel(S1Args a, const int *__restrict__ skip, long long G) {
    const int k = a.k, n = a.n, l = threadIdx.x & 31;
    const int ppw = 32 / k;                                     // points per warp
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int slot = l / k, j = l - slot * k;
    const long long p = warp * ppw + slot;
    const bool act = slot < ppw && p < a.n_pts && !(skip && skip[p]);
    const long long g = act ? p * k + j : 0;
    const double *d = a.W.dd + (act ? p : 0) * n, *e = a.W.ee + (act ? p : 0) * n;
    const double *lam = a.W.lam + (act ? p : 0) * k;
    const double tnorm = a.W.bnd[(act ? p : 0) * 4 + 3];
    const double eps3 = fmax(SCQED_EPS * tnorm, SCQED_SAFMIN * 1e16);
    const double ortol = 1e-3 * tnorm;
    TriLU W = { a.W.lu + g, a.W.lu + (size_t)n * G + g, a.W.lu + (size_t)2 * n * G + g, a.W.lu + (size_t)3 * n * G + g,
                a.W.swp + g, (int)G };
    double *z = a.W.zw + g;
    const int zs = (int)G;
    int bad = 0;
    bool clustered = false;
    if (act) for (int q = 1; q < k; ++q) clustered |= (fabs(lam[q] - lam[q - 1]) <= ortol);
    double lj = act ? lam[j] : 0.0;
    if (act && j > 0 && fabs(lj - lam[j - 1]) <= 10.0 * SCQED_EPS * fabs(lj)) lj += 10.0 * eps3 * (j & 1 ? 1.0 : 0.5);
    const unsigned any_clustered = __ballot_sync(0xffffffffu, act && clustered);
    const int npass = any_clustered ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
        if (act && (pass == 0 || clustered)) {
            trilu_factor(d, e, n, lj, eps3, W);
            if (pass == 0) {
                // the start vector depends on the level only, not on the position of the point inside the launch: a point's
                // eigenvectors (signs included) must not change with the way a sweep is chunked or sharded
                unsigned long long s = 0x9E3779B97F4A7C15ULL + (unsigned long long)(j + 1) * 0xD1B54A32D192ED03ULL;
                for (int i = 0; i < n; ++i) z[(size_t)i * zs] = lcg_uniform(s);
                trilu_solve(n, W, z, zs, eps3, false);
                double gr = normalize_serial(n, z, zs);
                for (int it = 0; it < 2; ++it) {
                    trilu_solve(n, W, z, zs, eps3, true);
                    gr = normalize_serial(n, z, zs);
                }
                if (!(gr * eps3 * n > 1e-6)) bad = 1;
            } else {
                trilu_solve(n, W, z, zs, eps3, true);
                normalize_serial(n, z, zs);
            }
        }
        if (any_clustered) {
            // modified Gram-Schmidt inside each eigenvalue cluster, sequential in j
            for (int jj = 1; jj < k; ++jj) {
                __syncwarp();
                if (act && clustered && j == jj) {
                    int start = jj;
                    while (start > 0 && fabs(lam[start] - lam[start - 1]) <= ortol) --start;
                    for (int q = start; q < jj; ++q) {
                        const double *zq = a.W.zw + (p * k + q);
                        double dot = 0.0;
                        for (int i = 0; i < n; ++i) dot = fma(zq[(size_t)i * zs], z[(size_t)i * zs], dot);
                        for (int i = 0; i < n; ++i) z[(size_t)i * zs] = fma(-dot, zq[(size_t)i * zs], z[(size_t)i * zs]);
                    }
                    if (start < jj) normalize_serial(n, z, zs);
                }
            }
            __syncwarp();
        }
    }
    if (act && bad) atomicOr(a.info + p, 1);
}

// ---- phase D: x = Q z.  One warp per (point, group of KV vectors) ------------------------------------------
template <typename T, int S1_KV, int RQ = SMALL_RQ>
static __global__ void __launch_bounds__(128)
s1_backtr_kernel(S1Args a, const int *__restrict__ skip, long long G, cplx *__restrict__ evecs) {
    const int n = a.n, k = a.k, l = threadIdx.x & 31;
    const int groups = (k + S1_KV - 1) / S1_KV;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= a.n_pts * groups) return;
    const long long p = warp / groups;
    const int g0 = (int)(warp - p * groups) * S1_KV;
    if (skip && skip[p]) return;
    const int kv = min(S1_KV, k - g0);
    T x[RQ][S1_KV];
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        int i = l + 32 * q;
#pragma unroll
        for (int u = 0; u < S1_KV; ++u) {
            double zv = (i < n && u < kv) ? a.W.zw[(size_t)i * G + (p * k + g0 + u)] : 0.0;
            if constexpr (sizeof(T) == sizeof(double)) x[q][u] = zv; else x[q][u] = cmake(zv, 0.0);
        }
    }
    const T *V = (const T *)a.W.Vst + (size_t)p * n * n;
    const T *tau = (const T *)a.W.tau + p * n;
    // reflector j lives in column j, rows j+1..n-1 (v[j+1] = 1 stored explicitly)
    T vn[RQ];
    auto load_col = [&](int j, T *dst) {
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            int i = l + 32 * q;
            dst[q] = (j >= 0 && i > j && i < n) ? V[(size_t)j * n + i] : nzero(T());
        }
    };
    load_col(n - 2, vn);
    for (int j = n - 2; j >= 0; --j) {
        T v[RQ];
#pragma unroll
        for (int q = 0; q < RQ; ++q) v[q] = vn[q];
        const T tj = tau[j];
        load_col(j - 1, vn);                                   // prefetch the next reflector
        if (niszero(tj)) continue;
        T dot[S1_KV];
#pragma unroll
        for (int u = 0; u < S1_KV; ++u) {
            dot[u] = nzero(T());
#pragma unroll
            for (int q = 0; q < RQ; ++q) nfmac(dot[u], v[q], x[q][u]);      // v^H x
        }
        if constexpr (sizeof(T) == sizeof(double) && S1_KV == 8) {
            // transpose-reduce: 8 partial dots over 32 lanes in 4 + 2 + 1 + 1 + 1 exchanges (each stage halves the
            // number of live values), then 8 broadcasts -- 17 shuffles instead of 40
            double e4[4], e2[2], g;
            const bool b16 = l & 16, b8 = l & 8, b4 = l & 4;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double send = b16 ? dot[u] : dot[u + 4];
                const double keep = b16 ? dot[u + 4] : dot[u];
                e4[u] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double send = b8 ? e4[u] : e4[u + 2];
                const double keep = b8 ? e4[u + 2] : e4[u];
                e2[u] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            {
                const double send = b4 ? e2[0] : e2[1];
                const double keep = b4 ? e2[1] : e2[0];
                g = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            g += __shfl_xor_sync(0xffffffffu, g, 2);
            g += __shfl_xor_sync(0xffffffffu, g, 1);
            // lane with bits (16, 8, 4) = (u2, u1, u0) holds dot u = 4 u2 + 2 u1 + u0
#pragma unroll
            for (int u = 0; u < 8; ++u) dot[u] = __shfl_sync(0xffffffffu, g, ((u & 4) ? 16 : 0) | ((u & 2) ? 8 : 0) | ((u & 1) ? 4 : 0));
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int u = 0; u < S1_KV; ++u) {
                    if constexpr (sizeof(T) == sizeof(double)) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
                    else {
                        cplx *dc = (cplx *)&dot[u];
                        dc->x += __shfl_xor_sync(0xffffffffu, dc->x, o);
                        dc->y += __shfl_xor_sync(0xffffffffu, dc->y, o);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < S1_KV; ++u) {
            const T f = nmul(tj, dot[u]);
#pragma unroll
            for (int q = 0; q < RQ; ++q) x[q][u] = nsub(x[q][u], nmul(f, v[q]));
        }
    }
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
        int i = l + 32 * q;
        if (i >= n) continue;
#pragma unroll
        for (int u = 0; u < S1_KV; ++u) {
            if (u >= kv) continue;
            cplx out;
            if constexpr (sizeof(T) == sizeof(double)) out = cmake(x[q][u], 0.0); else out = x[q][u];
            evecs[((size_t)p * k + g0 + u) * n + i] = out;
        }
    }
}

// ---- phase E: resonator strips, one CTA per point -------------------------------------------------------------
static __global__ void __launch_bounds__(128)
s1_resonator_kernel(S1Args a, ResonatorArgs R, const int *__restrict__ skip, const double *__restrict__ scalars,
                    const cplx *__restrict__ evecs, double *__restrict__ post) {
    __shared__ double scratch[8 * SCQED_RES_KMAX];
    const long long p = blockIdx.x;
    if (skip && skip[p]) return;
    resonator_strips(a.P, R, scalars + p * a.P.n_scalar, a.W.lam + p * a.k, evecs + (size_t)p * a.k * a.n, a.n, a.k,
                     scratch, post + (size_t)p * R.nstrips * a.k, nullptr, a.info + p);   // QL on private (L1-resident) arrays: a shared tile costs occupancy
}

// The same post-op behind the HBM dense (S2) and matrix-free (S3) solvers: eigenpairs come from
// global memory, nothing else is needed.
static __global__ void __launch_bounds__(128)
resonator_generic_kernel(PlanDev P, ResonatorArgs R, const double *__restrict__ scalars, const double *__restrict__ evals,
                         const cplx *__restrict__ evecs, int n, int k, double *__restrict__ post, int *__restrict__ info) {
    __shared__ double scratch[8 * SCQED_RES_KMAX];
    const long long p = blockIdx.x;
    resonator_strips(P, R, scalars + p * P.n_scalar, evals + p * k, evecs + (size_t)p * k * n, n, k, scratch,
                     post + (size_t)p * R.nstrips * k, nullptr, info + p);
}

// complement of the flags (second, complex pass only handles the flagged points)
static __global__ void s1_invert_flags(const int *__restrict__ flags, int *__restrict__ skip, long long n, int *__restrict__ count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = flags[i];
    skip[i] = !f;
    if (f) atomicAdd(count, 1);
}

}  // namespace scqed
