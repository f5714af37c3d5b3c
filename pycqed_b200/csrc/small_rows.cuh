// small_rows.cuh -- S1 phase A, real symmetric, ONE THREAD PER ROW on a packed lower triangle.
//
// Why: s1_tridiag_kernel (full n x n in shared memory, 256 threads) is bound by shared-memory wavefronts
// (ncu r01: 1.13e9 per 10^4-point launch = 60 % of its time) and by latency at 2 CTAs per SM (95.7 KB each).
// Here
//   * only the lower triangle lives on chip, column-major, every column padded to a length == 1 (mod 16):
//         L[i][c] (i >= c)  at  P[cst[c] + (i - c)],   cst[c] == c (mod 16)
//     n = 101: 47 KB of matrix + 6 KB of vectors -> 4 CTAs per SM;
//   * thread i owns ROW i of the trailing block.  Per Householder step the symmetric product y = A v is
//       phase R  (row part)     y_i += L[i][c] v_c,   c < i : lanes read one column at consecutive rows
//                               (contiguous), the three per-column operands v_c, v'_c, w'_c are broadcasts;
//                               the pending rank-2 update of the previous step is applied and stored here, so
//                               every element is loaded and stored ONCE per step;
//       phase C  (column part)  y_i += L[i'][i] v_i', i' > i : every lane walks down its own column; at a given
//                               offset the lanes touch addresses cst[i] + t == i + t (mod 16): conflict free,
//                               and v_{i+t} is a plain contiguous load.
//     y_i is complete inside thread i: no partial sums through shared memory, no warp reductions per column
//     (what made the column-per-warp packed kernel, small_packed.cuh, slower than full storage).
//   * the O(m) work between passes needs two block reductions (p.v and the norm of the next column); the
//     scalars that used to cost extra barriers (w_{j+1}, alpha) travel with those reductions: 4 barriers per step.
// Shared-memory traffic per step: ~0.19 m^2 wavefronts (full-storage kernel: ~0.38 m^2), 16 warps per SM as
// before but with 4 independent matrices instead of 2.
// Arithmetic is LAPACK dsytd2 'L' (same as small_tridiag); outputs as s1_tridiag_kernel<double>.
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"

namespace scqed {

// allocated length of packed column c: >= n - c and == 1 (mod 16), so that cst[c] == c (mod 16)
__host__ __device__ __forceinline__ int s1r_collen(int n, int c) { return ((n - c - 1 + 15) & ~15) + 1; }

struct S1RSmem {
    __host__ __device__ static size_t tri_elems(int n) {
        size_t s = 0;
        for (int c = 0; c < n; ++c) s += (size_t)s1r_collen(n, c);
        return s;
    }
    static size_t bytes(int n, int n_coef) {
        size_t o = (tri_elems(n) * 8 + 15) & ~(size_t)15;
        o += ((size_t)n * 4 + 15) & ~(size_t)15;          // cst
        o += (size_t)(4 * n + 8) * 8 + 64;                // v (2 buffers), w (2 buffers)
        o += (size_t)3 * n * 8 + 48;                      // d, e, tau
        o += (size_t)n_coef * 16 + 16;                    // coefficients during assembly
        o += (96 + 4 + 32) * 8 + 64;                      // red (tridiag_bounds), bounds, reduction slots
        return o;
    }
};

static __global__ void __launch_bounds__(256)
s1r_tridiag_kernel(S1Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, tid = threadIdx.x, NT = blockDim.x;
    const int wp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char *r = smem + o; o += (bytes + 15) & ~(size_t)15; return r; };
    double *P = (double *)take(S1RSmem::tri_elems(n) * 8);
    int *cst = (int *)take((size_t)n * 4);
    double *vbuf = (double *)take((size_t)(2 * n + 4) * 8);
    double *wbuf = (double *)take((size_t)(2 * n + 4) * 8);
    double *dv = (double *)take((size_t)n * 8);
    double *ev = (double *)take((size_t)n * 8);
    double *tauv = (double *)take((size_t)n * 8);
    const int ncf = a.P.n_terms > a.P.n_dense ? a.P.n_terms : a.P.n_dense;
    double *cf = (double *)take((size_t)ncf * 16);
    double *red = (double *)take(96 * 8);
    double *bounds = (double *)take(4 * 8);
    double *slot = (double *)take(32 * 8);              // [0..8) dot partials, [8..16) norm partials, 16: p_{j+1}, 17: alpha, 18: diag
    const int nh = (n + 1) & ~1;                        // vector buffers: second copy at an even offset (16-byte pairs)

    if (tid == 0) {
        int s = 0;
        for (int c = 0; c < n; ++c) { cst[c] = s; s += s1r_collen(n, c); }
    }
    __syncthreads();
    const int i = tid;                                  // this thread's row (and column, in phase C)
    const int my_cst = i < n ? cst[i] : 0;

    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        int not_real = 0;
        // ---- assemble the lower triangle: warps over columns, lanes down the column (coalesced table reads) ----
        if (a.Hin) {
            const cplx *Hp = a.Hin + (size_t)p * n * n;           // row-major Hermitian: H[r][c] = conj(H[c][r])
            for (int c = wp; c < n; c += NW) {
                double *col = P + cst[c] - c;
                for (int r = c + lane; r < n; r += 32) {
                    const cplx v = Hp[(size_t)c * n + r];
                    not_real |= (v.y != 0.0);
                    col[r] = v.x;
                }
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
                const double *dg = a.P.dense_tab + (size_t)NF * nn;
                for (int c = wp; c < n; c += NW) {
                    const double *gcol = a.P.dense_tab + (size_t)c * n;
                    double *col = P + cst[c] - c;
                    for (int r = c + lane; r < n; r += 32) {
                        double re = 0.0;
                        for (int g = 0; g < NF; ++g) re = fma(cf[g], __ldg(gcol + g * nn + r), re);
                        if (r == c)
                            for (int g = NF; g < NG; ++g) re = fma(cf[g], __ldg(dg + (size_t)(g - NF) * n + r), re);
                        if (a.tidyup) re = fabs(re) < 1e-12 ? 0.0 : re;
                        col[r] = re;
                    }
                }
            }
        } else {
            cplx *s_coef = (cplx *)cf;
            for (int t = tid; t < a.P.n_terms; t += NT) s_coef[t] = a.coef[p * a.P.n_terms + t];
            __syncthreads();
            for (int c = wp; c < n; c += NW) {
                int jd[SCQED_MAX_NODES], id[SCQED_MAX_NODES];
                digits(a.P, c, jd);
                double *col = P + cst[c] - c;
                for (int r = c + lane; r < n; r += 32) {
                    digits(a.P, r, id);
                    cplx v = h_element(a.P, s_coef, id, jd);
                    if (a.tidyup) v = tidy(v, 1e-12);
                    not_real |= (v.y != 0.0);
                    col[r] = v.x;
                }
            }
        }
        not_real = __syncthreads_or(not_real);
        if (tid == 0) { a.W.flags[p] = not_real; if (not_real) a.info[p] = -2; }
        if (not_real) continue;                                   // info = -2: the caller re-runs in complex mode

        double *Vout = a.want_vec ? (double *)a.W.Vst + (size_t)p * n * n : nullptr;
        double *vcur = vbuf, *vprev = vbuf + nh, *wcur = wbuf, *wprev = wbuf + nh;

        // Turn the next column (values c_i held by the threads, i > jn; thread jn holds the diagonal) into
        // reflector jn: publishes v into vdst[] (global row index, v[jn+1] = 1), d / e / tau, and the
        // reflector for the back-transformation.  Returns tau_jn.  One block barrier inside.
        auto make_reflector = [&](int jn, double ci, double *vdst) -> double {
            const int mm = n - jn;                                // rows jn .. n-1
            double xn = (i >= jn + 2 && i < n) ? ci * ci : 0.0;
            xn = warp_sum(xn);
            if (lane == 0) slot[8 + wp] = xn;
            if (i == jn + 1) slot[17] = ci;                       // alpha
            if (i == jn) slot[18] = ci;                           // diagonal
            __syncthreads();
            xn = 0.0;
            for (int w = 0; w < NW; ++w) xn += slot[8 + w];
            double tj = 0.0, beta = 0.0, sc = 0.0;
            if (mm >= 2) reflector_scalars(slot[17], xn, tj, beta, sc);
            if (i > jn && i < n) {
                const double v = (i == jn + 1) ? 1.0 : (tj == 0.0 ? 0.0 : ci * sc);
                vdst[i] = v;
                if (Vout) Vout[(size_t)jn * n + i] = v;
            }
            if (i == jn) { dv[jn] = slot[18]; ev[jn] = mm >= 2 ? beta : 0.0; tauv[jn] = tj; }
            return tj;
        };

        double tj;
        {
            const double c0 = i < n ? P[i] : 0.0;                 // column 0 (cst[0] = 0)
            tj = make_reflector(0, c0, vcur);
        }
        __syncthreads();

        for (int j = 0; j < n - 1; ++j) {
            const bool upd = j > 0;
            const bool on = i > j && i < n;
            double acc0 = 0.0, acc1 = 0.0;
            // ---- phase R: row part, with the pending update of step j-1 applied and stored ----------------------
            if ((wp << 5) + 31 > j) {                             // some row of this warp is in the trailing block
                const double vpi = (on && upd) ? vprev[i] : 0.0, wpi = (on && upd) ? wprev[i] : 0.0;
                const int cend = min(n, (wp << 5) + 32) - 1;      // columns j+1 .. cend-1 matter to some lane
                int addr = on ? cst[j + 1] + i - (j + 1) : 0;     // &L[i][c] at c = j+1
                int c = j + 1;
                if (upd) {
                    // four columns per trip: all loads are issued before the first store (the compiler cannot prove that
                    // the stores into P do not alias the vectors, so a one-column loop serialises on shared-memory latency)
                    for (; c + 4 <= cend; c += 4) {
                        double vc[4], vpc[4], wpc[4], x[4];
                        int ad[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { vc[u] = vcur[c + u]; vpc[u] = vprev[c + u]; wpc[u] = wprev[c + u]; }
#pragma unroll
                        for (int u = 0; u < 4; ++u) { ad[u] = addr; addr += (n + 14 - (c + u)) & ~15; }
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = (on && c + u < i) ? P[ad[u]] : 0.0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) { x[u] = fma(-vpi, wpc[u], x[u]); x[u] = fma(-wpi, vpc[u], x[u]); }
#pragma unroll
                        for (int u = 0; u < 4; ++u) if (on && c + u < i) P[ad[u]] = x[u];
                        if (on) {
                            acc0 = fma(c < i ? x[0] : 0.0, vc[0], acc0);
                            acc1 = fma(c + 1 < i ? x[1] : 0.0, vc[1], acc1);
                            acc0 = fma(c + 2 < i ? x[2] : 0.0, vc[2], acc0);
                            acc1 = fma(c + 3 < i ? x[3] : 0.0, vc[3], acc1);
                        }
                    }
                    for (; c < cend; ++c) {
                        const double vc = vcur[c], vpc = vprev[c], wpc = wprev[c];
                        if (on && c < i) {
                            double x = P[addr];
                            x = fma(-vpi, wpc, x);
                            x = fma(-wpi, vpc, x);
                            P[addr] = x;
                            acc0 = fma(x, vc, acc0);
                        }
                        addr += (n + 14 - c) & ~15;               // collen(c) - 1
                    }
                    if (on) {
                        double x = P[my_cst];
                        x = fma(-2.0 * vpi, wpi, x);
                        P[my_cst] = x;
                        acc1 = fma(x, vcur[i], acc1);
                    }
                } else {
                    for (; c < cend; ++c) {
                        const double vc = vcur[c];
                        if (on && c < i) acc0 = fma(P[addr], vc, acc0);
                        addr += (n + 14 - c) & ~15;
                    }
                    if (on) acc1 = fma(P[my_cst], vcur[i], acc1);
                }
            }
            __syncthreads();
            // ---- phase C: column part, every lane walks down its own (now updated) column ------------------------
            if (on) {
                const double *col = P + my_cst;
                const double *vv = vcur + i;
                const int len = n - i;                            // offsets 1 .. len-1
                int t = 1;
                double acc2 = 0.0, acc3 = 0.0;
                for (; t + 8 <= len; t += 8) {
                    double xa[8], va[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { xa[u] = col[t + u]; va[u] = vv[t + u]; }
                    acc0 = fma(xa[0], va[0], acc0); acc1 = fma(xa[1], va[1], acc1);
                    acc2 = fma(xa[2], va[2], acc2); acc3 = fma(xa[3], va[3], acc3);
                    acc0 = fma(xa[4], va[4], acc0); acc1 = fma(xa[5], va[5], acc1);
                    acc2 = fma(xa[6], va[6], acc2); acc3 = fma(xa[7], va[7], acc3);
                }
                acc0 += acc2; acc1 += acc3;
                for (; t < len; ++t) acc0 = fma(col[t], vv[t], acc0);
            }
            // ---- w = tau y - 1/2 tau (tau y . v) v --------------------------------------------------------------
            const double vi = on ? vcur[i] : 0.0;
            const double pi = on ? tj * (acc0 + acc1) : 0.0;
            double dot = warp_sum(pi * vi);
            if (lane == 0) slot[wp] = dot;
            if (i == j + 1) slot[16] = pi;                        // p_{j+1}: w_{j+1} = p_{j+1} + a2 (v_{j+1} = 1)
            __syncthreads();
            dot = 0.0;
            for (int w = 0; w < NW; ++w) dot += slot[w];
            const double a2 = -0.5 * tj * dot;
            const double wi = on ? fma(a2, vi, pi) : 0.0;
            if (on) wcur[i] = wi;
            const double w0 = slot[16] + a2;
            // ---- column j+1 with this step's update applied -> reflector j+1 --------------------------------------
            const int jn = j + 1;
            double ci = 0.0;
            if (i == jn) ci = fma(-2.0, w0, P[my_cst]);           // L[jn][jn] - 2 v_jn w_jn, v_jn = 1
            else if (i > jn && i < n) {
                ci = P[cst[jn] + i - jn];
                ci = fma(-vi, w0, ci);
                ci -= wi;                                         // - w_i v_jn
            }
            tj = make_reflector(jn, ci, vprev);                   // (its barrier also publishes wcur)
            __syncthreads();
            double *t_ = vcur; vcur = vprev; vprev = t_;
            t_ = wcur; wcur = wprev; wprev = t_;
        }
        if (tid == 0) ev[n - 1] = 0.0;
        __syncthreads();
        tridiag_bounds(dv, ev, n, bounds, red);
        for (int q = tid; q < n; q += NT) {
            const double e = ev[q];
            a.W.dd[p * n + q] = dv[q];
            a.W.ee[p * n + q] = e;
            a.W.e2[p * n + q] = e * e;
            ((double *)a.W.tau)[p * n + q] = tauv[q];
        }
        if (tid < 4) a.W.bnd[p * 4 + tid] = bounds[tid];
        __syncthreads();
    }
}

}  // namespace scqed
