// small_ladder.cuh -- S1 phase A, real symmetric, as an OCCUPANCY LADDER of partial reductions.
//
// Measured (tools/s1_scaling.py, profiles/r02_s1_scaling.md): the shared-memory tridiagonalisation is bound by the
// dependent chain of a Householder step (~3300 clk per column for ONE resident CTA) and its throughput grows with the
// number of matrices resident per SM until the per-column instruction count takes over: 1876 SM-clk per (matrix,
// column) at n = 101 (2 CTAs per SM: 94 KB of shared memory each), 1216 at n = 72 (4 CTAs), 1030 at n = 48 (8 CTAs).
// The trailing matrix SHRINKS during the reduction, so the reduction is cut into stages: stage s reduces the columns
// down to the next size, applies every pending update, and hands the (exact) trailing block to the next stage through
// global memory (n2^2 doubles per point, L2-sized); the next stage is an independent, smaller tridiagonalisation
// that runs with more CTAs per SM.  Same arithmetic as the single-kernel path (small_tridiag), bit-identical
// reflectors up to the order of the hand-off update.
#pragma once
#include "common.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"
#include "small_reg.cuh"

namespace scqed {

struct S1LStage {
    int n;                 // dimension of the matrix this stage works on
    int off;               // global index of its first row / column
    int j_stop;            // local columns 0 .. j_stop-1 are reduced here; n - 1: the reduction is completed
    const double *Tin;     // [n_pts][n * n] column-major trailing block of the previous stage (NULL: assemble H)
    double *Tout;          // [n_pts][(n - j_stop)^2] trailing block for the next stage (NULL when completed)
};

struct S1LSmem {
    static size_t bytes(int n_stage, int n_full, int n_coef) {
        size_t o = ((size_t)n_stage * n_stage * 8 + 15) & ~(size_t)15;
        o += (size_t)2 * n_full * 8 + 32;                       // dv, ev (full length: the last stage forms the bounds)
        o += (size_t)n_stage * 8 + 16;                          // tau
        o += (size_t)4 * n_stage * 8 + 32;                      // vbuf, wbuf
        size_t part = (size_t)SMALL_SEGS * n_stage * 8, cf = (size_t)n_coef * 16;
        o += (part > cf ? part : cf) + 16;
        o += 96 * 8 + 4 * 8 + 64 + 1024;                        // red, bounds, static shared of small_tridiag
        return o;
    }
};

static __global__ void __launch_bounds__(256)
s1l_tridiag_kernel(S1Args a, S1LStage s) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nf = a.n, n = s.n, tid = threadIdx.x, NT = blockDim.x;
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char *r = smem + o; o += (bytes + 15) & ~(size_t)15; return r; };
    double *A = (double *)take((size_t)n * n * 8);
    double *dv = (double *)take((size_t)nf * 8);
    double *ev = (double *)take((size_t)nf * 8);
    double *tau = (double *)take((size_t)n * 8);
    double *vbuf = (double *)take((size_t)2 * n * 8);
    double *wbuf = (double *)take((size_t)2 * n * 8);
    const int ncf = a.P.n_terms > a.P.n_dense ? a.P.n_terms : a.P.n_dense;
    size_t partb = (size_t)SMALL_SEGS * n * 8, cfb = (size_t)ncf * 16;
    double *part = (double *)take(partb > cfb ? partb : cfb);
    double *red = (double *)take(96 * 8);
    double *bounds = (double *)take(4 * 8);
    const bool complete = s.j_stop >= n - 1;
    const int n_out = complete ? n : s.j_stop;                   // columns whose d / e / tau are final after this stage
    const int n_refl = complete ? n - 1 : s.j_stop;              // reflector columns formed here

    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        if (s.Tin) {
            if (a.W.flags[p]) continue;                          // stage 1 found the point complex (info = -2)
            const double *Tp = s.Tin + (size_t)p * n * n;
            for (int idx = tid; idx < n * n; idx += NT) A[idx] = Tp[idx];
            __syncthreads();
        } else {
            int not_real = s1g_assemble(a, p, A, part, n);
            not_real = __syncthreads_or(not_real);
            if (tid == 0) { a.W.flags[p] = not_real; if (not_real) a.info[p] = -2; }
            if (not_real) continue;
        }
        if (complete) small_tridiag<double, false>(A, n, dv + s.off, ev + s.off, tau, vbuf, wbuf, part);
        else small_tridiag<double, true>(A, n, dv + s.off, ev + s.off, tau, vbuf, wbuf, part, s.j_stop);
        if (complete && tid == 0) ev[s.off + n - 1] = 0.0;
        __syncthreads();
        for (int i = tid; i < n_out; i += NT) {
            const double e = ev[s.off + i];
            a.W.dd[p * nf + s.off + i] = dv[s.off + i];
            a.W.ee[p * nf + s.off + i] = e;
            a.W.e2[p * nf + s.off + i] = e * e;
            ((double *)a.W.tau)[p * nf + s.off + i] = i < n_refl ? tau[i] : 0.0;
        }
        if (a.want_vec) {
            // reflector j (global column off + j): rows off + j + 1 .. nf - 1
            double *dst = (double *)a.W.Vst + (size_t)p * nf * nf;
            const int wp = tid >> 5, lane = tid & 31, NW = NT >> 5;
            for (int j = wp; j < n_refl; j += NW)
                for (int i = j + 1 + lane; i < n; i += 32) dst[(size_t)(s.off + j) * nf + s.off + i] = A[i + (size_t)j * n];
        }
        if (!complete) {
            const int m = n - s.j_stop;
            double *Tp = s.Tout + (size_t)p * m * m;
            for (int idx = tid; idx < m * m; idx += NT) {
                const int c = idx / m, r = idx - c * m;
                Tp[idx] = A[(s.j_stop + r) + (size_t)(s.j_stop + c) * n];
            }
        } else {
            // Gershgorin bounds need the whole tridiagonal matrix: the columns of the earlier stages come from HBM
            for (int i = tid; i < s.off; i += NT) { dv[i] = a.W.dd[p * nf + i]; ev[i] = a.W.ee[p * nf + i]; }
            __syncthreads();
            tridiag_bounds(dv, ev, nf, bounds, red);
            if (tid < 4) a.W.bnd[p * 4 + tid] = bounds[tid];
        }
        __syncthreads();
    }
}

}  // namespace scqed
