// small_eigh.cuh -- S1: the fused small-n path.  One CTA per sweep point; H(p) is assembled
// straight into shared memory and never touches HBM.
//
//   assemble (K2)  -> Householder tridiagonalisation (zhetd2, lower)  -> Sturm multisection
//   -> inverse iteration -> back-transformation -> (optional) resonator RWA strips (K7)
//
// Replaces, per sweep point, pyscqed/numerical_system.py:509-594 (getHamiltonian),
// pyscqed/util.py:127-176 (diagDenseH -> LAPACK zheevr) and numerical_system.py:736-784
// (getResonatorResponse).  Only c_t(p) comes in and only E, V, Erwa go out, so the stage is
// FP64-pipe / latency bound, not HBM bound.
//
// Shared-memory layout (host: small_smem_layout):  A[n*n] complex, column-major with the LOWER
// triangle authoritative during the reduction (both triangles are kept consistent so the
// matrix-vector product can stream contiguous columns: A[i + j*n] == H[i][j]).
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "resonator.cuh"

namespace scqed {

#define SMALL_THREADS 512
#define SMALL_SMAX 16

struct SmallLayout {
    int off_A, off_d, off_e, off_e2, off_tau, off_v, off_w, off_part, off_red, off_lam, off_bounds,
        off_X, off_lu, total;
    int vb;
};

struct SmallArgs {
    PlanDev P;
    const cplx *coef;       // [n_pts, n_terms]   (NULL when Hin is given)
    const cplx *Hin;        // optional [n_pts, n, n] row-major pre-assembled matrices
    const double *scalars;  // [n_pts, n_scalar]
    long long n_pts;
    int n, k, want_vec, tidyup;
    double *evals;          // [n_pts, k]
    cplx *evecs;            // [n_pts, k, n] or NULL
    int *info;              // [n_pts]
    ResonatorArgs res;
    double *post;           // [n_pts, res.nstrips, k] or NULL
    SmallLayout L;
};

// two-stage shuffle reductions (blockDim.x <= 1024): warp sums -> smem -> every warp re-reduces
__device__ __forceinline__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = l < nw ? red[l] : 0.0;
    return warp_sum(s);
}
__device__ __forceinline__ cplx block_sum(cplx v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) { red[2 * w] = v.x; red[2 * w + 1] = v.y; }
    __syncthreads();
    cplx s = l < nw ? cmake(red[2 * l], red[2 * l + 1]) : cmake(0.0, 0.0);
    return warp_sum(s);
}

__global__ void __launch_bounds__(SMALL_THREADS, 1)
small_eigh_kernel(SmallArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmallLayout &L = a.L;
    cplx *A = (cplx *)(smem + L.off_A);
    double *dv = (double *)(smem + L.off_d);
    double *ev = (double *)(smem + L.off_e);
    double *e2 = (double *)(smem + L.off_e2);
    cplx *tau = (cplx *)(smem + L.off_tau);
    cplx *vv = (cplx *)(smem + L.off_v);
    cplx *ww = (cplx *)(smem + L.off_w);
    cplx *part = (cplx *)(smem + L.off_part);
    double *red = (double *)(smem + L.off_red);
    double *lam = (double *)(smem + L.off_lam);
    double *bounds = (double *)(smem + L.off_bounds);
    cplx *X = (cplx *)(smem + L.off_X);
    double *lu = (double *)(smem + L.off_lu);

    const int n = a.n, k = a.k, tid = threadIdx.x, T = blockDim.x;
    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        // ---- assemble H(p) into shared memory -------------------------------------------------
        if (a.Hin) {
            const cplx *Hp = a.Hin + (size_t)p * n * n;
            for (int idx = tid; idx < n * n; idx += T) {
                int i = idx / n, j = idx - i * n;          // row-major source, coalesced
                A[i + j * n] = Hp[idx];
            }
        } else {
            cplx *s_coef = part;                            // scratch, free at this point
            for (int t = tid; t < a.P.n_terms; t += T) s_coef[t] = a.coef[p * a.P.n_terms + t];
            __syncthreads();
            for (int idx = tid; idx < n * n; idx += T) {
                int j = idx / n, i = idx - j * n;           // column-major destination
                int id[SCQED_MAX_NODES], jd[SCQED_MAX_NODES];
                digits(a.P, i, id);
                digits(a.P, j, jd);
                cplx v = h_element(a.P, s_coef, id, jd);
                if (a.tidyup) v = tidy(v, 1e-12);
                A[idx] = v;
            }
        }
        __syncthreads();

        // ---- Householder tridiagonalisation (LAPACK zhetd2, UPLO='L') ---------------------------
        for (int j = 0; j < n - 1; ++j) {
            const int m = n - j - 1;                        // trailing dimension; rows j+1 .. n-1
            cplx *col = A + (j + 1) + (size_t)j * n;        // x = A[j+1:n, j]
            double xn = 0.0;
            for (int r = 1 + tid; r < m; r += T) xn += cabs2(col[r]);
            xn = block_sum(xn, red);
            cplx alpha = col[0];
            cplx tj;
            double beta;
            if (xn == 0.0 && alpha.y == 0.0) {
                tj = cmake(0.0, 0.0);
                beta = alpha.x;
            } else {
                beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn), alpha.x);
                tj = cmake((beta - alpha.x) / beta, -alpha.y / beta);
            }
            // v = x / (alpha - beta), v[0] = 1
            cplx den = cmake(alpha.x - beta, alpha.y);
            double dn = 1.0 / cabs2(den);
            cplx sc = cmake(den.x * dn, -den.y * dn);
            __syncthreads();                                // everyone has read col[0]
            for (int r = tid; r < m; r += T) {
                cplx v = r == 0 ? cmake(1.0, 0.0) : ((tj.x == 0.0 && tj.y == 0.0) ? cmake(0.0, 0.0) : cmul(col[r], sc));
                vv[r] = v;
                col[r] = v;                                 // keep the reflector for the back-transform
            }
            if (tid == 0) {
                dv[j] = A[j + (size_t)j * n].x;
                ev[j] = beta;
                tau[j] = tj;
            }
            __syncthreads();
            if (tj.x == 0.0 && tj.y == 0.0) continue;

            // p = tau * A22 v : thread (r, s) sums a column segment of row r
            cplx *A22 = A + (j + 1) + (size_t)(j + 1) * n;
            int S = T / m; if (S > SMALL_SMAX) S = SMALL_SMAX; if (S < 1) S = 1;
            const int seg = (m + S - 1) / S;
            const int r = tid % m, s = tid / m;
            if (s < S) {
                int c0 = s * seg, c1 = min(m, c0 + seg);
                cplx acc = cmake(0.0, 0.0);
                for (int c = c0; c < c1; ++c) cfma(acc, A22[r + (size_t)c * n], vv[c]);
                part[s * m + r] = acc;
            }
            __syncthreads();
            cplx pr = cmake(0.0, 0.0), dot = cmake(0.0, 0.0);
            if (tid < m) {
                for (int q = 0; q < S; ++q) pr = cadd(pr, part[q * m + tid]);
                pr = cmul(tj, pr);
                cfmac(dot, pr, vv[tid]);                   // p^H v
            }
            dot = block_sum(dot, red);
            // alpha2 = -1/2 tau (p^H v);  w = p + alpha2 v
            cplx a2 = cmul(cmake(-0.5 * tj.x, -0.5 * tj.y), dot);
            if (tid < m) ww[tid] = cadd(pr, cmul(a2, vv[tid]));
            __syncthreads();
            // A22 -= v w^H + w v^H
            if (s < S) {
                int c0 = s * seg, c1 = min(m, c0 + seg);
                cplx vr = vv[r], wr = ww[r];
                for (int c = c0; c < c1; ++c) {
                    cplx x = A22[r + (size_t)c * n];
                    cfnmac(x, vr, ww[c]);
                    cfnmac(x, wr, vv[c]);
                    A22[r + (size_t)c * n] = x;
                }
            }
            __syncthreads();
        }
        if (tid == 0) { dv[n - 1] = A[(n - 1) + (size_t)(n - 1) * n].x; ev[n - 1] = 0.0; }
        __syncthreads();
        for (int i = tid; i < n; i += T) e2[i] = ev[i] * ev[i];
        __syncthreads();

        // ---- lowest-k eigenvalues of T ----------------------------------------------------------
        tridiag_bounds(dv, ev, n, bounds, red);
        int bad = tridiag_bisect(dv, e2, n, k, bounds, lam);
        __syncthreads();
        for (int i = tid; i < k; i += T) a.evals[p * k + i] = lam[i];

        // ---- eigenvectors: inverse iteration on T, then x = Q z -----------------------------------
        if (a.want_vec) {
            for (int idx = tid; idx < k * n; idx += T) X[idx] = cmake(0.0, 0.0);
            __syncthreads();
            bad |= tridiag_inverse_iteration(dv, ev, n, k, lam, bounds[3], (double *)X, 2 * n, 2, lu, L.vb,
                                             (unsigned long long)p + 1ULL);
            // Q = H(0) H(1) ... H(n-2): apply H(n-2) first.  One warp per vector, no block syncs.
            const int w = tid >> 5, l = tid & 31, nw = T >> 5;
            for (int jv = w; jv < k; jv += nw) {
                cplx *x = X + (size_t)jv * n;
                for (int j = n - 2; j >= 0; --j) {
                    cplx tj = tau[j];
                    if (tj.x == 0.0 && tj.y == 0.0) continue;
                    const int m = n - j - 1;
                    const cplx *v = A + (j + 1) + (size_t)j * n;
                    cplx *xs = x + j + 1;
                    cplx dot = cmake(0.0, 0.0);
                    for (int r = l; r < m; r += 32) cfmac(dot, v[r], xs[r]);   // v^H x
                    dot = warp_sum(dot);
                    cplx f = cmul(tj, dot);
                    for (int r = l; r < m; r += 32) {
                        cplx t = xs[r];
                        cplx vr = v[r];
                        t.x -= f.x * vr.x - f.y * vr.y;
                        t.y -= f.x * vr.y + f.y * vr.x;
                        xs[r] = t;
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            if (a.evecs) {
                cplx *out = a.evecs + (size_t)p * k * n;
                for (int idx = tid; idx < k * n; idx += T) out[idx] = X[idx];
            }
            if (a.res.enabled) {
                __syncthreads();
                // scratch: the Householder matrix A is dead now -> reuse it
                resonator_strips(a.P, a.res, a.scalars + p * a.P.n_scalar, lam, X, n, k,
                                 (double *)A, a.post + (size_t)p * a.res.nstrips * k);
            }
        }
        bad = __syncthreads_or(bad);
        if (tid == 0) a.info[p] = bad;
        __syncthreads();
    }
}

// Host: carve shared memory.  Returns total bytes; vb = vectors per inverse-iteration batch.
static inline SmallLayout small_smem_layout(int n, int k, int want_vec, int n_terms, size_t max_bytes) {
    SmallLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return (int)r; };
    L.off_A = take((size_t)n * n * 16);
    L.off_d = take((size_t)n * 8);
    L.off_e = take((size_t)n * 8);
    L.off_e2 = take((size_t)n * 8);
    L.off_tau = take((size_t)n * 16);
    L.off_v = take((size_t)n * 16);
    L.off_w = take((size_t)n * 16);
    size_t part = (size_t)SMALL_THREADS;
    if (part < (size_t)n_terms) part = n_terms;
    L.off_part = take(part * 16);
    L.off_red = take(96 * 8);
    L.off_lam = take((size_t)(k > 0 ? k : 1) * 8);
    L.off_bounds = take(4 * 8);
    L.off_X = take(want_vec ? (size_t)k * n * 16 : 0);
    L.off_lu = (int)o;
    L.vb = 0;
    if (want_vec) {
        size_t per = ((size_t)4 * n + (size_t)((n + 7) / 8)) * 8;
        long long room = (long long)max_bytes - (long long)o;
        int vb = room > 0 ? (int)(room / (long long)per) : 0;
        if (vb > k) vb = k;
        if (vb > SMALL_THREADS) vb = SMALL_THREADS;
        L.vb = vb;
        o += (size_t)(vb > 0 ? vb : 0) * per;
    }
    L.total = (int)o;
    return L;
}

}  // namespace scqed
