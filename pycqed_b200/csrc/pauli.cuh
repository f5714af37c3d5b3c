// pauli.cuh -- K8: local-basis reduction of the two lowest levels to Pauli coefficients.
//
// Replaces the per-point Python loop of util.pauliCoefficients (pyscqed/util.py:300-391): for each
// sweep point, diagonalise the 2x2 computational-basis operator Op (np.linalg.eigh, lower
// triangle), fix the gauge of both eigenvectors so their first component is real and positive
// (util.py:369), U = [v_0 v_1], Hq = U^H diag(E0, E1) U, and project on sigma_x,y,z:
//     hx = Re Hq01,  hy = -Im Hq01,  hz = (Hq00 - Hq11) / 2.
// The reference does this arithmetic in complex64; here it is float64 throughout.
// One thread per point; 48 bytes in, 24 bytes out.
#pragma once
#include "common.cuh"

namespace scqed {

static __global__ void __launch_bounds__(128)
pauli_coef_kernel(const double *__restrict__ e01, const cplx *__restrict__ op, long long n_pts,
                  double *__restrict__ h) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const double E0 = e01[2 * p], E1 = e01[2 * p + 1];
    // numpy.linalg.eigh reads the lower triangle: a = Re Op00, d = Re Op11, Op10 = conj(b)
    const double a = op[4 * p].x, d = op[4 * p + 3].x;
    const cplx b = cconj(op[4 * p + 2]);
    const double delta = 0.5 * (a - d);
    const double bb = b.x * b.x + b.y * b.y;
    const double r = sqrt(delta * delta + bb);
    // eigenvectors of [[a, b], [conj b, d]] for lambda_-/+ = (a+d)/2 -/+ r, cancellation-free forms
    cplx v0[2], v1[2];
    if (r == 0.0) {
        v0[0] = cmake(1.0, 0.0); v0[1] = cmake(0.0, 0.0);
        v1[0] = cmake(0.0, 0.0); v1[1] = cmake(1.0, 0.0);
    } else {
        if (delta > 0.0) { v0[0] = b; v0[1] = cmake(-delta - r, 0.0); v1[0] = cmake(delta + r, 0.0); v1[1] = cconj(b); }
        else             { v0[0] = cmake(delta - r, 0.0); v0[1] = cconj(b); v1[0] = b; v1[1] = cmake(r - delta, 0.0); }
    }
    cplx *vs[2] = {v0, v1};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        cplx *v = vs[q];
        const double nrm = sqrt(v[0].x * v[0].x + v[0].y * v[0].y + v[1].x * v[1].x + v[1].y * v[1].y);
        const double m0 = sqrt(v[0].x * v[0].x + v[0].y * v[0].y);
        // gauge: multiply by |v_0| / v_0 (util.py:369); a vanishing first component (NaN in the
        // reference) keeps its phase
        cplx ph = m0 > 0.0 ? cmake(v[0].x / m0, -v[0].y / m0) : cmake(1.0, 0.0);
        v[0] = cmul(v[0], ph); v[1] = cmul(v[1], ph);
        v[0].x /= nrm; v[0].y /= nrm; v[1].x /= nrm; v[1].y /= nrm;
    }
    // Hq_ab = sum_m conj(U_ma) E_m U_mb with U_mk = (v_k)_m
    const double hq00 = E0 * (v0[0].x * v0[0].x + v0[0].y * v0[0].y) + E1 * (v0[1].x * v0[1].x + v0[1].y * v0[1].y);
    const double hq11 = E0 * (v1[0].x * v1[0].x + v1[0].y * v1[0].y) + E1 * (v1[1].x * v1[1].x + v1[1].y * v1[1].y);
    const cplx t0 = cmul(cconj(v0[0]), v1[0]), t1 = cmul(cconj(v0[1]), v1[1]);
    const double hq01_re = E0 * t0.x + E1 * t1.x, hq01_im = E0 * t0.y + E1 * t1.y;
    h[3 * p] = hq01_re;
    h[3 * p + 1] = -hq01_im;
    h[3 * p + 2] = 0.5 * (hq00 - hq11);
}

// <V_a| A |V_b> of a FIXED dense operator A (n x n, row-major) between m vectors of every point:
// out[p][a][b], one CTA per point.  The m vectors are staged in shared memory; warp w takes rows
// i = w, w + nw, ...: lanes stride over the columns of row i (coalesced reads of A, which stays in L2 across
// the points), y_b = sum_j A[i][j] V_b[j] for every b, then acc[a][b] += conj(V_a[i]) y_b.
// util.pauliCoefficients' fixed-operator branch (util.py:320-342) loops this over points in Python.
#define DMEL_MMAX 4
static __global__ void __launch_bounds__(256)
dense_matel_kernel(const cplx *__restrict__ A, const cplx *__restrict__ V, int n, int m, long long n_pts,
                   cplx *__restrict__ out) {
    extern __shared__ cplx dm_s[];
    cplx *sv = dm_s;                                   // [m][n]
    cplx *sacc = dm_s + (size_t)m * n;                 // [nw][m*m]
    const long long p = blockIdx.x;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    const cplx *Vp = V + (size_t)p * m * n;
    for (int idx = threadIdx.x; idx < m * n; idx += blockDim.x) sv[idx] = Vp[idx];
    __syncthreads();
    cplx acc[DMEL_MMAX][DMEL_MMAX];
#pragma unroll
    for (int a = 0; a < DMEL_MMAX; ++a)
#pragma unroll
        for (int b = 0; b < DMEL_MMAX; ++b) acc[a][b] = cmake(0.0, 0.0);
    for (int i = w; i < n; i += nw) {
        cplx y[DMEL_MMAX];
#pragma unroll
        for (int b = 0; b < DMEL_MMAX; ++b) y[b] = cmake(0.0, 0.0);
        for (int j = l; j < n; j += 32) {
            const cplx aij = __ldg(A + (size_t)i * n + j);
#pragma unroll
            for (int b = 0; b < DMEL_MMAX; ++b) if (b < m) cfma(y[b], aij, sv[b * n + j]);
        }
#pragma unroll
        for (int b = 0; b < DMEL_MMAX; ++b) y[b] = warp_sum(y[b]);
#pragma unroll
        for (int a = 0; a < DMEL_MMAX; ++a)
#pragma unroll
            for (int b = 0; b < DMEL_MMAX; ++b) if (a < m && b < m) cfmac(acc[a][b], sv[a * n + i], y[b]);
    }
    if (l == 0)
        for (int a = 0; a < m; ++a)
            for (int b = 0; b < m; ++b) sacc[(size_t)w * m * m + a * m + b] = acc[a][b];
    __syncthreads();
    if ((int)threadIdx.x < m * m) {
        cplx s = cmake(0.0, 0.0);
        for (int ww = 0; ww < nw; ++ww) s = cadd(s, sacc[(size_t)ww * m * m + threadIdx.x]);
        out[(size_t)p * m * m + threadIdx.x] = s;
    }
}

}  // namespace scqed
