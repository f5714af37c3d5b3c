// small_packed.cuh -- S1 phase A on LOWER-TRIANGLE PACKED storage.
//
// s1_tridiag_kernel keeps the full n x n matrix in shared memory (both triangles consistent), which caps
// the real-symmetric path at 2 CTAs per SM for n = 101 (95.7 KB each) and at n <= ~165 overall; the
// kernel is latency bound, so the number of matrices in flight per SM is what limits it
// (profiles/r01_s1_tridiag_*).  Here only the lower triangle lives on chip:
//
//   packed, column-major:  L[r][c] (r >= c) at  P[c*n - c(c-1)/2 + (r - c)]
//   n = 101 real: 41 KB of matrix + 13 KB of vectors  ->  4 CTAs per SM;  n <= 210 real fits one SM.
//
// The trailing block of step j is itself a packed lower triangle (a suffix of P), so the same code
// serves every step.  Fused pass (update of step j-1 + mat-vec of step j), warp w owns columns
// c = w, w + NW, ...; lanes own rows by residue (r = lane + 32 i), so
//   * the row contributions  y_r += x v_c            accumulate in registers,
//   * the column contribution y_c += conj(x) v_r     is one warp reduction per column, four columns
//     batched so the shuffle chains overlap,
//   * column c is contiguous in shared memory: no bank conflicts.
// LAPACK z/d-hetd2 'L' arithmetic; outputs identical in meaning to s1_tridiag_kernel (d, e, e^2, tau,
// bounds, and the reflectors in the full column-major layout the back-transformation reads).
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"

namespace scqed {

#define S1P_CB 4      // columns per reduction batch

__device__ __forceinline__ void pser_barrier(int nthr) { asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); }

__device__ __forceinline__ double pser_sum(double v, double *slot, int nw, int nthr) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = v;
    pser_barrier(nthr);
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += slot[w];
    return s;
}

// threads 0 .. nthr-1 (row r = tid), nthr = 32 * ceil(m / 32): see ser_next_reflector
template <typename T>
__device__ __forceinline__ void pser_next_reflector(T *col, int m, bool upd, const T *vc, T w0, T wr, T *vnext,
                                                    double *dv, double *ev, T *tau, int jn, double *sm, T *sm_t,
                                                    int nw, int nthr) {
    const int r = threadIdx.x;
    T c = nzero(T());
    double xn = 0.0;
    if (r < m) {
        c = col[r];
        if (upd) { nfnmac(c, vc[r], w0); c = nsub(c, wr); }
        if (r >= 2) xn = nabs2(c);
        if (r < 2) sm_t[r] = c;
    }
    xn = pser_sum(xn, sm + 16, nw, nthr);
    const T diag = sm_t[0];
    if (m < 2) {
        if (r == 0) { dv[jn] = nre(diag); ev[jn] = 0.0; }
        return;
    }
    const T alpha = sm_t[1];
    T tj, sc;
    double beta;
    reflector_scalars(alpha, xn, tj, beta, sc);
    const bool zero = niszero(tj);
    if (r >= 1 && r < m) {
        T v = r == 1 ? none(T()) : (zero ? nzero(T()) : nmul(c, sc));
        vnext[r - 1] = v;
        col[r] = v;
    }
    if (r == 0) { col[0] = diag; dv[jn] = nre(diag); ev[jn] = beta; tau[jn] = tj; }
}

template <typename T> struct S1PSmem {
    static size_t bytes(int n, int n_coef, int nthreads) {
        const int nw = nthreads / 32;
        size_t o = ((size_t)n * (n + 1) / 2 * sizeof(T) + 15) & ~(size_t)15;
        o += (size_t)n * 8 * 2 + 32;                    // dv, ev
        o += (size_t)n * sizeof(T) + 16;                // tau
        o += (size_t)4 * n * sizeof(T) + 32;            // vbuf, wbuf
        o += (size_t)n * sizeof(T) + 16;                // ycol
        size_t part = (size_t)nw * n * sizeof(T), cf = (size_t)n_coef * 16;
        o += (part > cf ? part : cf) + 16;
        o += 96 * 8 + 4 * 8 + 40 * 8 + 64 + 256;        // red, bounds, ser_sm, ser_t, slack
        return o;
    }
};

// NQ = row tiles per lane (ceil(n / 32) <= NQ)
template <typename T, int NQ>
static __global__ void __launch_bounds__(NQ <= 4 ? 256 : 512)
s1p_tridiag_kernel(S1Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, tid = threadIdx.x, NT = blockDim.x;
    const int wp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    constexpr bool REAL = sizeof(T) == sizeof(double);
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char *r = smem + o; o += (bytes + 15) & ~(size_t)15; return r; };
    T *P = (T *)take((size_t)n * (n + 1) / 2 * sizeof(T));
    double *dv = (double *)take((size_t)n * 8);
    double *ev = (double *)take((size_t)n * 8);
    T *tau = (T *)take((size_t)n * sizeof(T));
    T *vbuf = (T *)take((size_t)2 * n * sizeof(T));
    T *wbuf = (T *)take((size_t)2 * n * sizeof(T));
    T *ycol = (T *)take((size_t)n * sizeof(T));
    const int ncf = a.P.n_terms > a.P.n_dense ? a.P.n_terms : a.P.n_dense;
    size_t partb = (size_t)NW * n * sizeof(T), cfb = (size_t)ncf * 16;
    T *part = (T *)take(partb > cfb ? partb : cfb);
    double *red = (double *)take(96 * 8);
    double *bounds = (double *)take(4 * 8);
    double *ser_sm = (double *)take(40 * 8);
    T *ser_t = (T *)take(2 * 16);
    auto off = [n](int c) { return (size_t)c * n - (size_t)c * (c - 1) / 2; };

    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        int not_real = 0;
        // ---- assemble the lower triangle ------------------------------------------------------------
        if (a.Hin) {
            // row-major Hermitian source: H[i][j] = conj(H[j][i]); row j is contiguous
            const cplx *Hp = a.Hin + (size_t)p * n * n;
            for (int j = wp; j < n; j += NW) {
                T *col = P + off(j) - j;
                for (int i = j + lane; i < n; i += 32) {
                    cplx v = cconj(Hp[(size_t)j * n + i]);
                    if constexpr (REAL) { not_real |= (v.y != 0.0); col[i] = v.x; } else col[i] = v;
                }
            }
        } else if (a.dcoef) {
            const int NG = a.P.n_dense;
            double *s_dc = (double *)part;
            for (int t = tid; t < 2 * NG; t += NT) s_dc[t] = a.dcoef[(size_t)p * 2 * NG + t];
            __syncthreads();
            const int NF = NG - a.P.n_dense_diag;
            double imw = 0.0;
            for (int g = 0; g < NF; ++g) imw += fabs(s_dc[NG + g]) * __ldg(a.P.dense_gmax + g);
            const bool pt_real = a.tidyup ? imw < 1e-12 : imw == 0.0;
            if (REAL && !pt_real) not_real = 1;
            else {
                const size_t nn = (size_t)n * n;
                const double *dg = a.P.dense_tab + (size_t)NF * nn;
                for (int j = wp; j < n; j += NW) {
                    const double *gcol = a.P.dense_tab + (size_t)j * n;
                    T *col = P + off(j) - j;
                    for (int i = j + lane; i < n; i += 32) {
                        double re = 0.0, im = 0.0;
                        for (int g = 0; g < NF; ++g) {
                            const double gv = __ldg(gcol + g * nn + i);
                            re = fma(s_dc[g], gv, re);
                            if (!REAL && !pt_real) im = fma(s_dc[NG + g], gv, im);
                        }
                        if (i == j)
                            for (int g = NF; g < NG; ++g) re = fma(s_dc[g], __ldg(dg + (size_t)(g - NF) * n + i), re);
                        if (a.tidyup) { re = fabs(re) < 1e-12 ? 0.0 : re; im = fabs(im) < 1e-12 ? 0.0 : im; }
                        if (i == j) im = 0.0;
                        if constexpr (REAL) col[i] = re; else col[i] = cmake(re, im);
                    }
                }
            }
        } else {
            cplx *s_coef = (cplx *)part;
            for (int t = tid; t < a.P.n_terms; t += NT) s_coef[t] = a.coef[p * a.P.n_terms + t];
            __syncthreads();
            for (int j = wp; j < n; j += NW) {
                int jd[SCQED_MAX_NODES], id[SCQED_MAX_NODES];
                digits(a.P, j, jd);
                T *col = P + off(j) - j;
                for (int i = j + lane; i < n; i += 32) {
                    digits(a.P, i, id);
                    cplx v = h_element(a.P, s_coef, id, jd);
                    if (a.tidyup) v = tidy(v, 1e-12);
                    if constexpr (REAL) { not_real |= (v.y != 0.0); col[i] = v.x; } else col[i] = v;
                }
            }
        }
        not_real = __syncthreads_or(not_real);
        if (REAL) {
            if (tid == 0) { a.W.flags[p] = not_real; if (not_real) a.info[p] = -2; }
            if (not_real) continue;
        }

        // ---- Householder tridiagonalisation --------------------------------------------------------------
        T *vcur = vbuf, *vprev = vbuf + n, *wprev = wbuf, *wcur = wbuf + n;
        {
            const int nw0 = (n + 31) >> 5, nthr0 = nw0 * 32;
            if (tid < nthr0)
                pser_next_reflector<T>(P, n, false, vcur, nzero(T()), nzero(T()), vcur, dv, ev, tau, 0, ser_sm,
                                       ser_t, nw0, nthr0);
        }
        __syncthreads();
        for (int j = 0; j < n - 1; ++j) {
            const int m = n - j - 1;
            T *Pj = P + off(j + 1);                              // packed lower triangle of size m
            const bool upd = j > 0;
            const int nq = (m + 31) >> 5;
            // this lane's rows: current v, previous v and w
            T acc[NQ], vrow[NQ], vr[NQ], wr[NQ];
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const int r = lane + 32 * i;
                const bool on = i < nq && r < m;
                acc[i] = nzero(T());
                vrow[i] = on ? vcur[r] : nzero(T());
                vr[i] = (on && upd) ? vprev[r + 1] : nzero(T());
                wr[i] = (on && upd) ? wprev[r + 1] : nzero(T());
            }
            for (int c0 = wp; c0 < m; c0 += S1P_CB * NW) {
                T cp[S1P_CB];
#pragma unroll
                for (int u = 0; u < S1P_CB; ++u) {
                    cp[u] = nzero(T());
                    const int c = c0 + u * NW;
                    if (c < m) {
                        const T vc = vcur[c];
                        const T wpc = upd ? wprev[c + 1] : nzero(T());
                        const T vpc = upd ? vprev[c + 1] : nzero(T());
                        T *col = Pj + ((size_t)c * m - (size_t)c * (c - 1) / 2) - c;   // col[r] = L[r][c]
                        const int ilo = c >> 5;
#pragma unroll
                        for (int i = 0; i < NQ; ++i) {
                            const int r = lane + 32 * i;
                            if (i >= ilo && i < nq && r >= c && r < m) {
                                T x = col[r];
                                if (upd) {
                                    nfnmac(x, vr[i], wpc);
                                    nfnmac(x, wr[i], vpc);
                                    col[r] = x;
                                }
                                nfma(acc[i], x, vc);                       // y_r += L[r][c] v_c
                                if (r > c) nfmac(cp[u], x, vrow[i]);       // y_c += conj(L[r][c]) v_r
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < S1P_CB; ++u) cp[u] = warp_sum(cp[u]);
                if (lane < S1P_CB) {
                    T mine = cp[0];
#pragma unroll
                    for (int u = 1; u < S1P_CB; ++u) if (lane == u) mine = cp[u];
                    const int c = c0 + lane * NW;
                    if (c < m) ycol[c] = mine;
                }
            }
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                const int r = lane + 32 * i;
                if (i < nq && r < m) part[wp * m + r] = acc[i];
            }
            __syncthreads();
            const int nthr = nq * 32;
            if (tid < nthr) {
                const T tj = tau[j];
                T pr = nzero(T());
                double dx = 0.0, dy = 0.0;
                if (tid < m) {
                    T sum = ycol[tid];
                    for (int u = 0; u < NW; ++u) sum = nadd(sum, part[u * m + tid]);
                    pr = nmul(tj, sum);
                    T d1 = nzero(T());
                    nfmac(d1, pr, vcur[tid]);                    // p^H v
                    dx = nre(d1);
                    if constexpr (!REAL) dy = ((cplx *)&d1)->y;
                }
                dx = warp_sum(dx);
                if constexpr (!REAL) dy = warp_sum(dy);
                if (lane == 0) { ser_sm[wp] = dx; ser_sm[8 + wp] = dy; }
                pser_barrier(nthr);
                dx = 0.0; dy = 0.0;
                for (int w = 0; w < nq; ++w) { dx += ser_sm[w]; dy += ser_sm[8 + w]; }
                T dot;
                if constexpr (REAL) dot = dx; else dot = cmake(dx, dy);
                const T a2 = nmul(nscale(-0.5, tj), dot);
                T wrr = nzero(T());
                if (tid < m) { wrr = nadd(pr, nmul(a2, vcur[tid])); wcur[tid] = wrr; }
                if (tid == 0) ser_t[0] = wrr;
                pser_barrier(nthr);
                const T w0 = ser_t[0];
                pser_barrier(nthr);
                pser_next_reflector<T>(Pj, m, true, vcur, w0, wrr, vprev, dv, ev, tau, j + 1, ser_sm, ser_t, nq, nthr);
            }
            __syncthreads();
            T *t = vcur; vcur = vprev; vprev = t;
            t = wcur; wcur = wprev; wprev = t;
        }
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
            // reflector j: rows j+1.. of column j, written where s1_backtr_kernel reads them
            T *dst = (T *)a.W.Vst + (size_t)p * n * n;
            for (int j = wp; j < n - 1; j += NW) {
                const T *col = P + off(j) - j;
                for (int i = j + 1 + lane; i < n; i += 32) dst[(size_t)j * n + i] = col[i];
            }
        }
        __syncthreads();
    }
}

}  // namespace scqed
