// small_blocked.cuh -- S1 phase A, real symmetric, BLOCKED (LAPACK dlatrd / dsytrd form) with the trailing
// rank-2k update on the FP64 tensor cores.
//
// Why: the unblocked kernel (s1_tridiag_kernel, small_eigh.cuh::small_tridiag) reads AND writes the whole trailing
// block of H in shared memory at every Householder step (fused "apply the pending rank-2 update + next mat-vec" pass)
// and is bound by shared-memory wavefronts (ncu r01: 1.13e9 per 10^4-point launch, ~60 % of its time, 3.3e9 warp
// instructions).  In the blocked form the trailing block is only READ per step (y = A v: one 16-byte load and two FMAs
// per two elements) and written once per panel of NB = 8 columns, by DMMA:
//
//   panel columns j = j0 .. j0+7 (A22 = A[j0+8:, j0+8:] stays stale, true A = A - V W^T - W V^T):
//     S1/S2 (ONE warp, lane owns rows lane + 32 q: every reduction is a warp shuffle, no barrier, four independent rows
//           per lane hide the FP64 latency) column j made current: a_r -= sum_k V[r][k] W[j][k] + W[r][k] V[j][k];
//           d_j, reflector v_j (dlarfg), tau_j
//     P     (all 8 warps) y = A_stale v_j: warps split the live columns, a thread owns two row pairs (LDS.128);
//           warps k < jj also form t1[k] = W[:,k].v, t2[k] = V[:,k].v for the corrections
//     S3    y -= V t1 + W t2; p = tau y; w = p - (tau/2)(p.v) v -> column jj of the W panel
//   after the panel: A22 -= V W^T + W V^T as 8x8 DMMA tiles (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4): accumulators
//           are the tiles of A22 (loaded from / stored to shared memory once per panel), A/B operands are 8x4 slices of
//           the panels.
//
// Shared-memory traffic per matrix: ~8 B per element and step (the mat-vec) + 16 B per element and PANEL, against
// 16 B per element and step before.  Same arithmetic as LAPACK dsytrd 'L' with nb = 8; outputs (d, e, tau, normalised
// reflectors with v[j+1] = 1, Gershgorin bounds) exactly like s1_tridiag_kernel<double>.
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"
#include "small_reg.cuh"

namespace scqed {

#define S1B_NB 8
#define S1B_THREADS 256

__device__ __forceinline__ void s1b_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct S1BSmem {
    __host__ __device__ static int lda(int n) { return (n + 1) & ~1; }          // even: 16-byte row pairs
    __host__ __device__ static int ps(int n) { return (n + 7) & ~7; }           // panel / partial-sum row stride
    static size_t bytes(int n, int n_coef) {
        size_t o = ((size_t)n * lda(n) * 8 + 15) & ~(size_t)15;                 // A
        o += (size_t)2 * S1B_NB * ps(n) * 8;                                    // V and W panels
        size_t part = (size_t)8 * ps(n) * 8, cf = (size_t)n_coef * 16;
        o += ((part > cf ? part : cf) + 15) & ~(size_t)15;                      // partial sums | coefficients
        o += (size_t)3 * n * 8 + 48;                                            // dv, ev, tauv
        o += (96 + 4 + 2 * S1B_NB + 24) * 8 + 64;                               // red, bounds, t1|t2, reduction slots
        return o;
    }
};

// j_stop (a multiple of S1B_NB, or >= n for the complete reduction): stop after the panels below j_stop, write the exact
// trailing block A[j_stop:, j_stop:] to Tout [n_pts][(n - j_stop)^2] (column-major) for a later stage (small_ladder.cuh).
static __global__ void __launch_bounds__(S1B_THREADS)
s1b_tridiag_kernel(S1Args a, int j_stop, double *Tout) {
    constexpr int NB = S1B_NB;
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, tid = threadIdx.x;
    const int wp = tid >> 5, lane = tid & 31;
    const int lda = S1BSmem::lda(n), PS = S1BSmem::ps(n);
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char *p_ = smem + o; o += (bytes + 15) & ~(size_t)15; return p_; };
    double *A = (double *)take((size_t)n * lda * 8);
    double *Vs = (double *)take((size_t)NB * PS * 8);
    double *Ws = (double *)take((size_t)NB * PS * 8);
    const int ncf = a.P.n_terms > a.P.n_dense ? a.P.n_terms : a.P.n_dense;
    size_t partb = (size_t)8 * PS * 8, cfb = (size_t)ncf * 16;
    double *part = (double *)take(partb > cfb ? partb : cfb);
    double *dv = (double *)take((size_t)n * 8);
    double *ev = (double *)take((size_t)n * 8);
    double *tauv = (double *)take((size_t)n * 8);
    double *red = (double *)take(96 * 8);
    double *bounds = (double *)take(4 * 8);
    double *ts = (double *)take((size_t)2 * NB * 8);      // t1 = W^T v | t2 = V^T v
    (void)take(24 * 8);

    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        int not_real = s1g_assemble(a, p, A, part, lda);
        not_real = __syncthreads_or(not_real);
        if (tid == 0) { a.W.flags[p] = not_real; if (not_real) a.info[p] = -2; }
        if (not_real) continue;                              // info = -2: the caller re-runs in complex mode

        const bool complete = j_stop >= n;
        for (int j0 = 0; j0 < n && j0 < j_stop; j0 += NB) {
            const int pn = min(NB, n - j0);
            for (int jj = 0; jj < pn; ++jj) {
                const int j = j0 + jj;
                double vq[4] = {0.0, 0.0, 0.0, 0.0}, tau = 0.0;
                // ---- S1 / S2 (warp 0, lane owns rows lane + 32 q): column j made current, reflector j ----
                if (wp == 0) {
                    double c[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = lane + 32 * q;
                        c[q] = (r >= j && r < n) ? A[r + (size_t)j * lda] : 0.0;
                    }
                    for (int k = 0; k < jj; ++k) {
                        const double wjk = Ws[k * PS + j], vjk = Vs[k * PS + j];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = lane + 32 * q;
                            if (r < PS) c[q] = fma(-Vs[k * PS + r], wjk, fma(-Ws[k * PS + r], vjk, c[q]));
                        }
                    }
                    {
                        const int qd = j >> 5;
                        const double cd = qd == 0 ? c[0] : (qd == 1 ? c[1] : (qd == 2 ? c[2] : c[3]));
                        if (lane == (j & 31)) dv[j] = cd;
                    }
                    if (j < n - 1) {
                        double xn = 0.0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = lane + 32 * q;
                            if (r >= j + 2 && r < n) xn = fma(c[q], c[q], xn);
                        }
                        xn = warp_sum(xn);
                        const int qa = (j + 1) >> 5;
                        const double ca = qa == 0 ? c[0] : (qa == 1 ? c[1] : (qa == 2 ? c[2] : c[3]));
                        const double alpha = __shfl_sync(0xffffffffu, ca, (j + 1) & 31);
                        double beta, sc;
                        reflector_scalars(alpha, xn, tau, beta, sc);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = lane + 32 * q;
                            vq[q] = (r == j + 1) ? 1.0 : ((r > j + 1 && r < n) ? c[q] * sc : 0.0);
                            if (r < PS) Vs[jj * PS + r] = vq[q];
                            if (r > j && r < n) A[r + (size_t)j * lda] = vq[q];       // kept for the back-transformation
                        }
                        if (lane == 0) { ev[j] = beta; tauv[j] = tau; }
                    }
                }
                if (j == n - 1) break;                                            // only the last diagonal was left
                __syncthreads();
                // ---- P: y = A_stale v over the live columns; t1 / t2 for the corrections ----
                {
                    const double *vj = Vs + jj * PS;
                    const int m = n - 1 - j, seg = (m + 7) >> 3;
                    const int c0 = j + 1 + wp * seg, c1 = min(n, c0 + seg);
                    const bool q0 = j < 63, q1 = n > 64;                          // row blocks 0..63 / 64..127 alive
                    double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
                    const double *ap = A + 2 * lane + (size_t)c0 * lda;
                    const double *vp = vj + c0;
                    int cnt = c1 - c0;
                    if (q0 && q1) {
#pragma unroll 4
                        for (; cnt > 0; --cnt, ap += lda, ++vp) {
                            const double vc = *vp;
                            const double2 x0 = *reinterpret_cast<const double2 *>(ap);
                            const double2 x1 = *reinterpret_cast<const double2 *>(ap + 64);
                            acc0.x = fma(x0.x, vc, acc0.x); acc0.y = fma(x0.y, vc, acc0.y);
                            acc1.x = fma(x1.x, vc, acc1.x); acc1.y = fma(x1.y, vc, acc1.y);
                        }
                    } else if (q0) {
#pragma unroll 4
                        for (; cnt > 0; --cnt, ap += lda, ++vp) {
                            const double vc = *vp;
                            const double2 x0 = *reinterpret_cast<const double2 *>(ap);
                            acc0.x = fma(x0.x, vc, acc0.x); acc0.y = fma(x0.y, vc, acc0.y);
                        }
                    } else {
#pragma unroll 4
                        for (; cnt > 0; --cnt, ap += lda, ++vp) {
                            const double vc = *vp;
                            const double2 x1 = *reinterpret_cast<const double2 *>(ap + 64);
                            acc1.x = fma(x1.x, vc, acc1.x); acc1.y = fma(x1.y, vc, acc1.y);
                        }
                    }
                    if (2 * lane < PS) *reinterpret_cast<double2 *>(part + wp * PS + 2 * lane) = acc0;
                    if (2 * lane + 64 < PS) *reinterpret_cast<double2 *>(part + wp * PS + 2 * lane + 64) = acc1;
                    const int kt = 7 - wp;                                        // warp 0 (the serial warp) goes last
                    if (kt < jj) {
                        double t1 = 0.0, t2 = 0.0;
                        for (int i = j + 1 + lane; i < n; i += 32) {
                            const double vi = vj[i];
                            t1 = fma(Ws[kt * PS + i], vi, t1);
                            t2 = fma(Vs[kt * PS + i], vi, t2);
                        }
                        t1 = warp_sum(t1); t2 = warp_sum(t2);
                        if (lane == 0) { ts[kt] = t1; ts[NB + kt] = t2; }
                    }
                }
                __syncthreads();
                // ---- S3 (warp 0): corrections, w ----
                if (wp == 0) {
                    double y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = lane + 32 * q;
                        y[q] = 0.0;
                        if (r < PS)
                            y[q] = ((part[r] + part[PS + r]) + (part[2 * PS + r] + part[3 * PS + r])) +
                                   ((part[4 * PS + r] + part[5 * PS + r]) + (part[6 * PS + r] + part[7 * PS + r]));
                    }
                    for (int k = 0; k < jj; ++k) {
                        const double t1 = ts[k], t2 = ts[NB + k];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = lane + 32 * q;
                            if (r < PS) y[q] = fma(-Vs[k * PS + r], t1, fma(-Ws[k * PS + r], t2, y[q]));
                        }
                    }
                    double dt = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = lane + 32 * q;
                        y[q] = (r > j && r < n) ? tau * y[q] : 0.0;               // p = tau y on the live rows
                        dt = fma(y[q], vq[q], dt);
                    }
                    dt = warp_sum(dt);
                    const double hc = -0.5 * tau * dt;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = lane + 32 * q;
                        if (r < PS) Ws[jj * PS + r] = fma(hc, vq[q], y[q]);
                    }
                    __syncwarp();
                }
            }
            // ---- trailing update on the FP64 tensor cores: A22 -= V W^T + W V^T ----
            const int b0 = j0 + NB;
            if (pn == NB && b0 < n) {
                __syncthreads();
                const int nt = (n - b0 + 7) >> 3;
                const int g = lane >> 2, t = lane & 3;
                for (int tr = wp; tr < nt; tr += S1B_THREADS / 32) {
                    const int R0 = b0 + 8 * tr;
                    const double va0 = -Vs[t * PS + R0 + g], va1 = -Vs[(4 + t) * PS + R0 + g];
                    const double wa0 = -Ws[t * PS + R0 + g], wa1 = -Ws[(4 + t) * PS + R0 + g];
                    const int row = R0 + g;
                    for (int tc = 0; tc < nt; ++tc) {
                        const int C0 = b0 + 8 * tc, col = C0 + 2 * t;
                        const double vb0 = Vs[t * PS + C0 + g], vb1 = Vs[(4 + t) * PS + C0 + g];
                        const double wb0 = Ws[t * PS + C0 + g], wb1 = Ws[(4 + t) * PS + C0 + g];
                        const bool ok0 = row < n && col < n, ok1 = row < n && col + 1 < n;
                        double d0 = ok0 ? A[row + (size_t)col * lda] : 0.0;
                        double d1 = ok1 ? A[row + (size_t)(col + 1) * lda] : 0.0;
                        s1b_dmma(d0, d1, va0, wb0);
                        s1b_dmma(d0, d1, va1, wb1);
                        s1b_dmma(d0, d1, wa0, vb0);
                        s1b_dmma(d0, d1, wa1, vb1);
                        if (ok0) A[row + (size_t)col * lda] = d0;
                        if (ok1) A[row + (size_t)(col + 1) * lda] = d1;
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (complete && tid == 0) { ev[n - 1] = 0.0; tauv[n - 1] = 0.0; }
        __syncthreads();
        const int n_out = complete ? n : j_stop;
        for (int i = tid; i < n_out; i += S1B_THREADS) {
            const double e = ev[i];
            a.W.dd[p * n + i] = dv[i];
            a.W.ee[p * n + i] = e;
            a.W.e2[p * n + i] = e * e;
            ((double *)a.W.tau)[p * n + i] = tauv[i];
        }
        if (complete) {
            tridiag_bounds(dv, ev, n, bounds, red);
            if (tid < 4) a.W.bnd[p * 4 + tid] = bounds[tid];
        } else {
            const int m = n - j_stop;
            double *Tp = Tout + (size_t)p * m * m;
            for (int idx = tid; idx < m * m; idx += S1B_THREADS) {
                const int c = idx / m, r = idx - c * m;
                Tp[idx] = A[(j_stop + r) + (size_t)(j_stop + c) * lda];
            }
        }
        if (a.want_vec) {
            // reflector j: column j, rows j+1 .. n-1 (s1_backtr_kernel reads nothing else)
            double *dst = (double *)a.W.Vst + (size_t)p * n * n;
            const int n_refl = complete ? n - 1 : j_stop;
            for (int j = wp; j < n_refl; j += S1B_THREADS / 32)
                for (int i = j + 1 + lane; i < n; i += 32) dst[(size_t)j * n + i] = A[i + (size_t)j * lda];
        }
        __syncthreads();
    }
}

}  // namespace scqed
