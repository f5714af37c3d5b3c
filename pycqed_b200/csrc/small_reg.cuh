// small_reg.cuh -- S1 phase A, real symmetric, with the WHOLE matrix in REGISTERS.
//
// Why: s1_tridiag_kernel streams the trailing block through shared memory once per Householder step and is bound by
// shared-memory wavefronts (ncu r01: 1.13e9 per 10^4-point launch, 60 % of its time); keeping only a packed triangle
// on chip (small_packed.cuh, small_rows.cuh) did not change that count.  Here shared memory carries only O(n) operands:
//
//   * TPR threads share a row (TPR = 2: lanes 2 rl, 2 rl + 1 of a warp), thread (r, q) owns the column pairs
//     c = 2 (TPR i + q) + {0, 1}, i = 0 .. NPAIR-1, of row r in 2 NPAIR double registers.  Both triangles are kept:
//     row j of the trailing block IS column j, so nothing is ever transposed.
//   * step j, unnormalised reflector H = I - a x~ x~^T with x~ = (alpha - beta, x_2, ..., x_m), a = tau / (alpha - beta)^2:
//       select    x_r = A[r][j] out of the registers (jump table on the pair index), publish it, warp-partial sums of
//                 squares                                                                   -> block barrier 1
//       mat-vec   y_r = sum_c A[r][c] x~_c : register FMAs against 16-byte broadcast loads of x~ (one wavefront per
//                 warp and pair group), combined over the TPR lanes by shuffles; y and the warp-partial dots y.x~ are
//                 published                                                                 -> block barrier 2
//       update    z = a y - b x~ (b = a^2 (y.x~) / 2) is rebuilt by every warp privately (m FMAs per warp), then
//                 A[r][c] -= x~_r z_c + z_r x~_c in registers.
//     Dead columns (c <= j) hold zeros in x~ and z, so the unrolled passes need no masks; whole dead pair groups are
//     skipped by entering the unrolled pass through a jump table (Duff's device).
//   Shared-memory wavefronts per step: ~3 per live pair group and warp (full-storage kernel: the whole trailing block
//   twice), 3 FMAs per element and step, two block barriers per step.
// Arithmetic: LAPACK dsytd2 'L' (beta = -sign(alpha) norm, tau = (beta - alpha) / beta, v = x~ / (alpha - beta)); the
// normalised v and tau are what leaves the kernel, exactly like s1_tridiag_kernel<double>.
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"

namespace scqed {

#define S1G_CASES(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) \
    M(16) M(17) M(18) M(19) M(20) M(21) M(22) M(23) M(24) M(25) M(26) M(27) M(28) M(29) M(30) M(31)

struct S1GSmem {
    // X (2 buffers), Y, per-warp Z: rows / columns padded to NC; the staging copy of H for the assembly
    static size_t bytes(int n, int n_coef, int NC, int NW) {
        size_t o = ((size_t)n * n * 8 + 15) & ~(size_t)15;      // As
        o += (size_t)(3 + NW) * NC * 8 + 64;                    // X[2], Y, Zw[NW]
        o += (size_t)3 * n * 8 + 48;                            // dv, ev, tauv
        o += (size_t)n_coef * 16 + 16;                          // coefficients during the assembly
        o += (96 + 4 + 80) * 8 + 64;                            // red, bounds, ps / pd / scalar slots
        return o;
    }
};

// Assemble the real symmetric H(p) into shared memory (column-major with leading dimension lda, both triangles); returns non-zero when the point
// is not real.  Same three sources as s1_tridiag_kernel: a caller matrix, the dense real term table, the term walk.
static __device__ int s1g_assemble(const S1Args &a, long long p, double *A, double *cf, int lda) {
    const int n = a.n, tid = threadIdx.x, NT = blockDim.x;
    int not_real = 0;
    if (a.Hin) {
        const cplx *Hp = a.Hin + (size_t)p * n * n;
        for (int idx = tid; idx < n * n; idx += NT) {
            const int i = idx / n, j = idx - i * n;              // row-major source, coalesced
            const cplx v = Hp[idx];
            not_real |= (v.y != 0.0);
            A[i + j * lda] = v.x;
        }
    } else if (a.dcoef) {
        const int NG = a.P.n_dense;
        for (int t = tid; t < 2 * NG; t += NT) cf[t] = a.dcoef[(size_t)p * 2 * NG + t];
        __syncthreads();
        const int NF = NG - a.P.n_dense_diag;
        double imw = 0.0;
        for (int g = 0; g < NF; ++g) imw += fabs(cf[NG + g]) * __ldg(a.P.dense_gmax + g);
        const bool pt_real = a.tidyup ? imw < 1e-12 : imw == 0.0;
        if (!pt_real) not_real = 1;
        else {
            const size_t nn = (size_t)n * n;
            const int wp = tid >> 5, lane = tid & 31, NW = NT >> 5;
            const double *dg = a.P.dense_tab + (size_t)NF * nn;
            for (int j = wp; j < n; j += NW) {
                const double *col = a.P.dense_tab + (size_t)j * n;
                // two rows per lane and pass, the table loads of both issued back to back (the assembly was stalled on
                // one L2 load at a time: ncu r02, 11 % of the first ladder stage)
                for (int i = j + lane; i < n; i += 64) {
                    const int i2 = i + 32;
                    const bool two = i2 < n;
                    double re = 0.0, re2 = 0.0;
#pragma unroll 4
                    for (int g = 0; g < NF; ++g) {
                        const double t0 = __ldg(col + (size_t)g * nn + i);
                        const double t1 = two ? __ldg(col + (size_t)g * nn + i2) : 0.0;
                        re = fma(cf[g], t0, re);
                        re2 = fma(cf[g], t1, re2);
                    }
                    if (i == j)
                        for (int g = NF; g < NG; ++g) re = fma(cf[g], __ldg(dg + (size_t)(g - NF) * n + i), re);
                    if (a.tidyup) { re = fabs(re) < 1e-12 ? 0.0 : re; re2 = fabs(re2) < 1e-12 ? 0.0 : re2; }
                    A[i + j * lda] = re;
                    A[j + i * lda] = re;
                    if (two) { A[i2 + j * lda] = re2; A[j + i2 * lda] = re2; }
                }
            }
        }
    } else {
        cplx *s_coef = (cplx *)cf;
        for (int t = tid; t < a.P.n_terms; t += NT) s_coef[t] = a.coef[p * a.P.n_terms + t];
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += NT) {
            const int j = idx / n, i = idx - j * n;               // column-major destination
            int id[SCQED_MAX_NODES], jd[SCQED_MAX_NODES];
            digits(a.P, i, id);
            digits(a.P, j, jd);
            cplx v = cconj(h_element(a.P, s_coef, jd, id));
            if (a.tidyup) v = tidy(v, 1e-12);
            not_real |= (v.y != 0.0);
            A[i + j * lda] = v.x;
        }
    }
    return not_real;
}

template <int TPR, int NPAIR, int MAXT, int MAXR>
static __global__ void __launch_bounds__(MAXT) __maxnreg__(MAXR)
s1g_tridiag_kernel(S1Args a) {
    static_assert(TPR == 1 || TPR == 2 || TPR == 4, "threads per row");
    static_assert(NPAIR <= 32, "at most 32 column pairs per thread");
    constexpr int R = 32 / TPR;                 // rows per warp
    constexpr int NC = 2 * TPR * NPAIR;         // padded matrix dimension
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, tid = threadIdx.x, NT = blockDim.x;
    const int wp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    const int q = lane & (TPR - 1), r = wp * R + lane / TPR;
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char *p_ = smem + o; o += (bytes + 15) & ~(size_t)15; return p_; };
    double *As = (double *)take((size_t)n * n * 8);
    double *X = (double *)take((size_t)2 * NC * 8);
    double *Y = (double *)take((size_t)NC * 8);
    double *Zall = (double *)take((size_t)NW * NC * 8);
    double *dv = (double *)take((size_t)n * 8);
    double *ev = (double *)take((size_t)n * 8);
    double *tauv = (double *)take((size_t)n * 8);
    const int ncf = a.P.n_terms > a.P.n_dense ? a.P.n_terms : a.P.n_dense;
    double *cf = (double *)take((size_t)ncf * 16);
    double *red = (double *)take(96 * 8);
    double *bounds = (double *)take(4 * 8);
    double *ps = (double *)take(32 * 8);         // warp partials: sum of squares
    double *pd = (double *)take(32 * 8);         // warp partials: y . x~
    double *slot = (double *)take(16 * 8);       // 0 / 1: alpha of even / odd steps
    double *Zw = Zall + (size_t)wp * NC;
    const int last_row = wp * R + R - 1;         // the warp is idle once j >= last_row

    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        int not_real = s1g_assemble(a, p, As, cf, n);
        for (int t = tid; t < 3 * NC; t += NT) X[t] = 0.0;          // X[2] and Y are contiguous
        for (int t = lane; t < NC; t += 32) Zw[t] = 0.0;
        not_real = __syncthreads_or(not_real);
        if (tid == 0) { a.W.flags[p] = not_real; if (not_real) a.info[p] = -2; }
        if (not_real) continue;                              // info = -2: the caller re-runs in complex mode

        // ---- rows into registers ----
        double A_[NPAIR][2];
#pragma unroll
        for (int i = 0; i < NPAIR; ++i) {
            const int c0 = 2 * (TPR * i + q);
            A_[i][0] = (r < n && c0 < n) ? As[r + (size_t)c0 * n] : 0.0;
            A_[i][1] = (r < n && c0 + 1 < n) ? As[r + (size_t)(c0 + 1) * n] : 0.0;
        }
        double *Vout = (double *)a.W.Vst + (size_t)p * n * n;

        for (int j = 0; j < n; ++j) {
            double *Xc = X + (j & 1) * NC;
            // ---- select column j of this row out of the registers ----
            const int qj = (j >> 1) & (TPR - 1), ij = j / (2 * TPR), sj = j & 1;
            double xsel = 0.0;
            if (q == qj && r >= j) {
                switch (ij) {
#define S1G_SEL(I) case I: if constexpr (I < NPAIR) xsel = sj ? A_[I < NPAIR ? I : 0][1] : A_[I < NPAIR ? I : 0][0]; break;
                    S1G_CASES(S1G_SEL)
#undef S1G_SEL
                    default: break;
                }
            }
            if (q == qj) {
                if (r == j) dv[j] = xsel;
                if (r >= j - 1 && r < NC) Xc[r] = r > j ? xsel : 0.0;         // rows j-1, j: dead entries of this buffer
                if (r == j + 1) slot[j & 1] = xsel;
            }
            if (j == n - 1) break;
            {
                double s2 = (r > j + 1) ? xsel * xsel : 0.0;                  // xsel = 0 on the non-selecting lanes
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                if (lane == 0) ps[wp] = s2;
            }
            __syncthreads();                                                   // barrier 1
            double sigma = 0.0;
            for (int w = 0; w < NW; ++w) sigma += ps[w];
            const double alpha = slot[j & 1];
            if (sigma == 0.0) {                                                // H = I (dlarfg: tau = 0, beta = alpha)
                if (tid == 0) { ev[j] = alpha; tauv[j] = 0.0; }
                if (r == j && q == 0) Y[j] = 0.0;
                __syncthreads();                                               // ps / slot are rewritten by the next step
                continue;
            }
            const double beta = -copysign(sqrt(fma(alpha, alpha, sigma)), alpha);
            const double piv = alpha - beta;
            const double sc = 1.0 / piv;
            const double tau = (beta - alpha) / beta;
            const double ac = tau * sc * sc;
            if (tid == 0) { ev[j] = beta; tauv[j] = tau; }
            if (lane == 0) Xc[j + 1] = piv;                                    // every warp, same value: no barrier
            __syncwarp();
            const double xr = (r > j && r < NC) ? Xc[r] : 0.0;
            if (a.want_vec && q == 0 && r > j && r < n) Vout[(size_t)j * n + r] = (r == j + 1) ? 1.0 : xr * sc;
            const int imin = (j + 1) / (2 * TPR);
            const double *Xq = Xc + 2 * q;
            // ---- mat-vec ----
            double yr = 0.0;
            if (last_row > j) {
                double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
                switch (imin) {
#define S1G_MV(I) case I: if constexpr (I < NPAIR) {                                                        \
                        const double2 u = *reinterpret_cast<const double2 *>(Xq + 2 * TPR * I);            \
                        if constexpr ((I & 1) == 0) { y0 = fma(A_[I < NPAIR ? I : 0][0], u.x, y0); y1 = fma(A_[I < NPAIR ? I : 0][1], u.y, y1); } \
                        else { y2 = fma(A_[I < NPAIR ? I : 0][0], u.x, y2); y3 = fma(A_[I < NPAIR ? I : 0][1], u.y, y3); } }
                    S1G_CASES(S1G_MV)
#undef S1G_MV
                    default: break;
                }
                yr = (y0 + y1) + (y2 + y3);
#pragma unroll
                for (int off = 1; off < TPR; off <<= 1) yr += __shfl_xor_sync(0xffffffffu, yr, off);
                if (r <= j) yr = 0.0;
                double dt = q == 0 ? yr * xr : 0.0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) dt += __shfl_xor_sync(0xffffffffu, dt, off);
                if (lane == 0) pd[wp] = dt;
                if (q == 0 && r >= j && r < NC) Y[r] = yr;                     // Y[j] = 0: the row died
            } else if (lane == 0) pd[wp] = 0.0;
            __syncthreads();                                                   // barrier 2
            if (last_row > j) {
                double dot = 0.0;
                for (int w = 0; w < NW; ++w) dot += pd[w];
                const double bc = 0.5 * ac * ac * dot;
                for (int c = imin * 2 * TPR + lane; c < NC; c += 32) Zw[c] = fma(ac, Y[c], -bc * Xc[c]);
                __syncwarp();
                const double nzr = -fma(ac, yr, -bc * xr), nxr = -xr;
                const double *Zq = Zw + 2 * q;
                switch (imin) {
#define S1G_UP(I) case I: if constexpr (I < NPAIR) {                                                        \
                        const double2 u = *reinterpret_cast<const double2 *>(Xq + 2 * TPR * I);            \
                        const double2 z = *reinterpret_cast<const double2 *>(Zq + 2 * TPR * I);            \
                        A_[I < NPAIR ? I : 0][0] = fma(nzr, u.x, fma(nxr, z.x, A_[I < NPAIR ? I : 0][0]));  \
                        A_[I < NPAIR ? I : 0][1] = fma(nzr, u.y, fma(nxr, z.y, A_[I < NPAIR ? I : 0][1])); }
                    S1G_CASES(S1G_UP)
#undef S1G_UP
                    default: break;
                }
            }
        }
        __syncthreads();
        if (tid == 0) ev[n - 1] = 0.0;
        __syncthreads();
        tridiag_bounds(dv, ev, n, bounds, red);
        for (int i = tid; i < n; i += NT) {
            const double e = ev[i];
            a.W.dd[p * n + i] = dv[i];
            a.W.ee[p * n + i] = e;
            a.W.e2[p * n + i] = e * e;
            ((double *)a.W.tau)[p * n + i] = i < n - 1 ? tauv[i] : 0.0;
        }
        if (tid < 4) a.W.bnd[p * 4 + tid] = bounds[tid];
        __syncthreads();
    }
}

}  // namespace scqed
