// dense_eigh.cuh -- S2: batched dense Hermitian eigensolver for matrices that live in HBM
// (n above the shared-memory limit of the fused path).  Replaces pyscqed/util.py:127-176
// (diagDenseH -> scipy.linalg.eigh -> LAPACK zheevr) for a whole batch of sweep points.
//
// Storage: A[b] is n x n complex, COLUMN-major (A[i + j*n] == H[i][j]); both triangles are kept
// consistent so that the Hermitian matrix-vector product streams contiguous columns.
//
// Stage 1 (this file, unblocked zhetd2 form): per column j
//     hh_gen   : Householder vector of A[j+1:, j]                      (1 CTA per matrix)
//     hemv     : partial p = A22 v over column segments                (HBM/L2 bound: 16 m^2 B)
//     hh_w     : w = tau p - 1/2 tau (p^H v) tau... (zhetd2)           (1 CTA per matrix)
//     rank2    : A22 -= v w^H + w v^H                                  (32 m^2 B)
// Stage 2: tridiag_eig_kernel (Sturm multisection + inverse iteration, 1 CTA per matrix)
// Stage 3: back-transformation x = Q z, one CTA per (matrix, vector).
// The blocked zlatrd/her2k form with FP64 tensor-core (DMMA) trailing updates is in
// dense_blocked.cuh and supersedes stage 1 when enabled.
#pragma once
#include "common.cuh"
#include "tridiag_eig.cuh"

namespace scqed {

#define D2_THREADS 256

__device__ __forceinline__ double block_sum256(double v, double *red) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}

// grid (1, B).  Generates reflector j; writes v to vbuf[b][0..m) (v[0] = 1) and into the column.
static __global__ void __launch_bounds__(D2_THREADS)
d2_hh_gen(cplx *A, int n, int j, cplx *vbuf, double *dd, double *ee, cplx *tau) {
    __shared__ double red[32];
    const int b = blockIdx.y;
    cplx *Ab = A + (size_t)b * n * n;
    const int m = n - j - 1;
    cplx *col = Ab + (j + 1) + (size_t)j * n;
    double xn = 0.0;
    for (int r = 1 + threadIdx.x; r < m; r += blockDim.x) xn += cabs2(col[r]);
    xn = block_sum256(xn, red);
    cplx alpha = col[0];
    cplx tj;
    double beta;
    if (xn == 0.0 && alpha.y == 0.0) { tj = cmake(0.0, 0.0); beta = alpha.x; }
    else {
        beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn), alpha.x);
        tj = cmake((beta - alpha.x) / beta, -alpha.y / beta);
    }
    cplx den = cmake(alpha.x - beta, alpha.y);
    double dn = 1.0 / cabs2(den);
    cplx sc = cmake(den.x * dn, -den.y * dn);
    bool zero = (tj.x == 0.0 && tj.y == 0.0);
    __syncthreads();
    cplx *v = vbuf + (size_t)b * n;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        cplx x = r == 0 ? cmake(1.0, 0.0) : (zero ? cmake(0.0, 0.0) : cmul(col[r], sc));
        v[r] = x;
        col[r] = x;
    }
    if (threadIdx.x == 0) {
        dd[(size_t)b * n + j] = Ab[j + (size_t)j * n].x;
        ee[(size_t)b * n + j] = beta;
        tau[(size_t)b * n + j] = tj;
        if (j == n - 2) {
            // last diagonal entry is final only after this step's rank-2 update; d2_hh_w fixes it
            ee[(size_t)b * n + n - 1] = 0.0;
        }
    }
}

// grid (row_blocks, nseg, B), block 128: part[b][seg][r] = sum_{c in seg} A22[r][c] v[c]
#define D2_HEMV_ROWS 128
static __global__ void __launch_bounds__(D2_HEMV_ROWS)
d2_hemv(const cplx *__restrict__ A, int n, int j, const cplx *__restrict__ vbuf, cplx *__restrict__ part, int nseg) {
    extern __shared__ cplx s_v[];
    const int b = blockIdx.z, m = n - j - 1;
    const int seg = (m + nseg - 1) / nseg;
    const int c0 = blockIdx.y * seg, c1 = min(m, c0 + seg);
    const int r = blockIdx.x * D2_HEMV_ROWS + threadIdx.x;
    const cplx *v = vbuf + (size_t)b * n;
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) s_v[c - c0] = v[c];
    __syncthreads();
    if (r >= m) return;
    const cplx *A22 = A + (size_t)b * n * n + (j + 1) + (size_t)(j + 1) * n;
    cplx acc0 = cmake(0.0, 0.0), acc1 = cmake(0.0, 0.0);
    int c = c0;
    for (; c + 1 < c1; c += 2) {
        cplx a0 = A22[r + (size_t)c * n], a1 = A22[r + (size_t)(c + 1) * n];
        cfma(acc0, a0, s_v[c - c0]);
        cfma(acc1, a1, s_v[c + 1 - c0]);
    }
    if (c < c1) cfma(acc0, A22[r + (size_t)c * n], s_v[c - c0]);
    part[((size_t)b * nseg + blockIdx.y) * n + r] = cadd(acc0, acc1);
}

// grid (1, B): p = tau * sum_seg part ; w = p - 1/2 tau (p^H v) v
static __global__ void __launch_bounds__(D2_THREADS)
d2_hh_w(int n, int j, const cplx *vbuf, const cplx *part, int nseg, const cplx *tau, cplx *wbuf) {
    __shared__ double red[64];
    const int b = blockIdx.y, m = n - j - 1;
    const cplx tj = tau[(size_t)b * n + j];
    const cplx *v = vbuf + (size_t)b * n;
    cplx *w = wbuf + (size_t)b * n;
    cplx dot = cmake(0.0, 0.0);
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        cplx p = cmake(0.0, 0.0);
        for (int s = 0; s < nseg; ++s) p = cadd(p, part[((size_t)b * nseg + s) * n + r]);
        p = cmul(tj, p);
        w[r] = p;
        cfmac(dot, p, v[r]);
    }
    double dx = block_sum256(dot.x, red);
    double dy = block_sum256(dot.y, red);
    cplx a2 = cmul(cmake(-0.5 * tj.x, -0.5 * tj.y), cmake(dx, dy));
    for (int r = threadIdx.x; r < m; r += blockDim.x) w[r] = cadd(w[r], cmul(a2, v[r]));
}

// grid (row_blocks, col_blocks, B), block 128 rows x 16 columns per block
#define D2_R2_COLS 16
static __global__ void __launch_bounds__(D2_HEMV_ROWS)
d2_rank2(cplx *A, int n, int j, const cplx *__restrict__ vbuf, const cplx *__restrict__ wbuf,
         const cplx *__restrict__ tau, double *dd) {
    const int b = blockIdx.z, m = n - j - 1;
    const cplx tj = tau[(size_t)b * n + j];
    if (tj.x == 0.0 && tj.y == 0.0) {
        if (j == n - 2 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
            dd[(size_t)b * n + n - 1] = A[(size_t)b * n * n + (n - 1) + (size_t)(n - 1) * n].x;
        return;
    }
    const int r = blockIdx.x * D2_HEMV_ROWS + threadIdx.x;
    const int c0 = blockIdx.y * D2_R2_COLS, c1 = min(m, c0 + D2_R2_COLS);
    if (r >= m) return;
    const cplx *v = vbuf + (size_t)b * n, *w = wbuf + (size_t)b * n;
    cplx *A22 = A + (size_t)b * n * n + (j + 1) + (size_t)(j + 1) * n;
    const cplx vr = v[r], wr = w[r];
    for (int c = c0; c < c1; ++c) {
        cplx x = A22[r + (size_t)c * n];
        cfnmac(x, vr, __ldg(w + c));
        cfnmac(x, wr, __ldg(v + c));
        A22[r + (size_t)c * n] = x;
        if (j == n - 2) dd[(size_t)b * n + n - 1] = x.x;   // m == 1: the final diagonal entry
    }
}

// Stage 2: one CTA per matrix.  d/e [B][n]; lam [B][k]; Z real [B][k][n]; lu_ws [B][vb*per].
static __global__ void __launch_bounds__(256)
tridiag_eig_kernel(const double *dd, const double *ee, int n, int k, int want_vec, double *evals, double *Z,
                   double *lu_ws, int vb, double *e2buf, int *info) {
    __shared__ double red[96];
    __shared__ double bounds[4];
    extern __shared__ double s_lam[];
    const int b = blockIdx.x;
    const double *d = dd + (size_t)b * n, *e = ee + (size_t)b * n;
    double *e2 = e2buf + (size_t)b * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) e2[i] = e[i] * e[i];
    __syncthreads();
    tridiag_bounds(d, e, n, bounds, red);
    int bad = tridiag_bisect(d, e2, n, k, bounds, s_lam);
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) evals[(size_t)b * k + i] = s_lam[i];
    if (want_vec) {
        const size_t per = (size_t)4 * n + (size_t)((n + 7) / 8);
        bad |= tridiag_inverse_iteration(d, e, n, k, s_lam, bounds[3], Z + (size_t)b * k * n, n, 1,
                                         lu_ws + (size_t)b * vb * per, vb, (unsigned long long)b + 1ULL);
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) info[b] = bad;
}

// Stage 3: grid (k, B): x = H(0) ... H(n-2) z for one vector.  conj_out: write conj(x)
// (used when the caller's matrix was row-major, i.e. the transposed = conjugated problem).
static __global__ void __launch_bounds__(D2_THREADS)
d2_backtransform(const cplx *__restrict__ A, int n, int k, const cplx *__restrict__ tau,
                 const double *__restrict__ Z, cplx *__restrict__ X, int conj_out) {
    extern __shared__ cplx s_x[];
    __shared__ double red[64];
    const int jv = blockIdx.x, b = blockIdx.y;
    const cplx *Ab = A + (size_t)b * n * n;
    const double *z = Z + ((size_t)b * k + jv) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_x[i] = cmake(z[i], 0.0);
    __syncthreads();
    for (int j = n - 2; j >= 0; --j) {
        cplx tj = tau[(size_t)b * n + j];
        if (tj.x == 0.0 && tj.y == 0.0) continue;
        const int m = n - j - 1;
        const cplx *v = Ab + (j + 1) + (size_t)j * n;
        cplx *xs = s_x + j + 1;
        cplx dot = cmake(0.0, 0.0);
        for (int r = threadIdx.x; r < m; r += blockDim.x) cfmac(dot, v[r], xs[r]);
        double dx = block_sum256(dot.x, red);
        double dy = block_sum256(dot.y, red);
        cplx f = cmul(tj, cmake(dx, dy));
        for (int r = threadIdx.x; r < m; r += blockDim.x) {
            cplx t = xs[r], vr = v[r];
            t.x -= f.x * vr.x - f.y * vr.y;
            t.y -= f.x * vr.y + f.y * vr.x;
            xs[r] = t;
        }
        __syncthreads();
    }
    cplx *out = X + ((size_t)b * k + jv) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        cplx t = s_x[i];
        out[i] = conj_out ? cconj(t) : t;
    }
}

}  // namespace scqed
