// small_tiled.cuh -- S1 phase A, real symmetric: blocked tridiagonalisation on a TILE-PACKED LOWER TRIANGLE.
//
// Measured (profiles/r02_s1_scaling.md): the shared-memory tridiagonalisation is bound by the dependent chain of a
// Householder step times the number of matrices one SM can hold -- 2 at n = 101 in full storage (94 KB per CTA).  Here
// only the lower triangle lives on chip, as 8 x 8 row-major tiles (R, C), R >= C (diagonal tiles complete): 91 tiles =
// 46.6 KB at n = 101, and the whole CTA stays under 55 KB => FOUR CTAs per SM.  The layout is chosen for the three
// accesses of the blocked (dlatrd) form:
//   * trailing update A22 -= V W^T + W V^T once per panel of NB = 4 columns: a tile IS the C fragment of
//     mma.sync.m8n8k4.f64 (lane (g, t) <-> T[g][2t], T[g][2t+1]): one LDS.128, two DMMA, one STS.128 per tile;
//   * symmetric mat-vec y = A v per column from the stored half: work item R = "tile row R (T v_C, C <= R) + tile column
//     R (T^T v_R', R' > R)".  Both parts feed the SAME eight rows 8R..8R+7, every item touches NT - T0 tiles (perfectly
//     balanced over the warps), partial sums never leave the warp: two shuffles per item, no partial-sum arrays;
//   * column j for the reflector: a strided walk over at most NT tiles (O(m), one warp).
// The O(m) sections run on ONE warp (lane owns rows lane + 32 q): every reduction is a warp shuffle, ~700 instructions
// per column instead of ~2000 on four warps -- at four CTAs per SM the issue slots, not the chain, are what is scarce.
// Arithmetic: LAPACK dsytrd 'L' with nb = 4; outputs exactly like s1_tridiag_kernel<double>.
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"
#include "small_blocked.cuh"
#include "tiled_common.cuh"

namespace scqed {

#define S1T_THREADS 128

struct S1TSmem {
    __host__ __device__ static int nt(int n) { return (n + 7) >> 3; }
    __host__ __device__ static int tile_doubles(int n) { return nt(n) * (nt(n) + 1) / 2 * 64; }
    static size_t bytes(int n) {
        const int ps = 8 * nt(n) + 4;
        return (size_t)tile_doubles(n) * 8 + (size_t)2 * S1T_NB * ps * 8 + (size_t)ps * 8 + 2 * S1T_NB * 8 + 64;
    }
    // the coefficients of the assembly (and, after the reduction, d / e / bounds scratch) alias the panels
    static bool scratch_fits(int n, int n_coef) {
        const size_t panels = (size_t)2 * S1T_NB * (8 * nt(n) + 4) * 8;
        return (size_t)n_coef * 16 <= panels && (size_t)(2 * n + 96 + 4 + 8) * 8 <= panels;
    }
};

// element (i, j), i >= j, of the tile-packed lower triangle (diagonal tiles hold both triangles)
__device__ __forceinline__ void s1t_store(double *tiles, int i, int j, double v) {
    const int R = i >> 3, C = j >> 3;
    double *t = tiles + s1t_tile(R, C);
    t[(i & 7) * 8 + (j & 7)] = v;
    if (R == C) t[(j & 7) * 8 + (i & 7)] = v;
}

// PSER: the O(m) sections on ALL warps (thread r = row r, n <= NTHR) instead of one warp: shorter dependent chain
// per column (one row per lane), four block barriers per column instead of two.
template <bool PSER, int NTHR>
static __global__ void __launch_bounds__(NTHR, NTHR <= 128 ? 4 : 1)
s1t_tridiag_kernel(S1Args a, int *sm_slots, int rotate) {
    constexpr int NB = S1T_NB;
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = a.n, tid = threadIdx.x;
    const int wp = tid >> 5, lane = tid & 31;
    const int NT = S1TSmem::nt(n), PS = 8 * NT + 4, NTD = S1TSmem::tile_doubles(n);   // PS = 4 (mod 8): conflict-free fragments
    double *tiles = (double *)smem;
    double *Vs = tiles + NTD;
    double *Ws = Vs + NB * PS;
    double *Y = Ws + NB * PS;
    double *ts = Y + PS;                                   // t1 = W^T v | t2 = V^T v
    double *cf = Vs;                                       // assembly coefficients (before the first panel)
    const int g8 = lane >> 2, t4 = lane & 3;             // row part / DMMA fragments: row g8, column pair 2 t4
    const int c8 = lane & 7, r4 = lane >> 3;             // column part: column c8, rows r4 and r4 + 4 (conflict-free: 8 r4 + c8)
    // !PSER: the O(m) sections run on ONE warp per CTA.  Warp w of every CTA lives on scheduler w % 4, so with warp 0 in
    // all four co-resident CTAs the serial sections share one scheduler.  Spreading them (rotate = 1: serial warp = arrival
    // order of the CTA on its SM) or keeping scheduler 0 free of mat-vec work (rotate = 2) both measured SLOWER (6.2-6.5 vs
    // 5.95 ms at 256 threads): the serial warps are latency bound and interleave well on one scheduler, the mat-vec warps
    // are what needs the other three.  Kept as experiment switches (SCQED_S1_TILED_ROTATE).
    __shared__ int s_sw;
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_sw = rotate == 1 ? (atomicAdd(sm_slots + (smid & 1023), 1) & 3) : 0;
    }
    __syncthreads();
    const int sw = s_sw;

    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        // ---- assemble the lower triangle into the tiles ----
        int not_real = 0;
        for (int q = tid; q < NTD; q += NTHR) tiles[q] = 0.0;
        if (a.Hin) {
            __syncthreads();
            const cplx *Hp = a.Hin + (size_t)p * n * n;
            for (int idx = tid; idx < n * n; idx += NTHR) {
                const int i = idx / n, j = idx - i * n;          // row-major source, coalesced
                if (j <= i) {
                    const cplx v = Hp[idx];
                    not_real |= (v.y != 0.0);
                    s1t_store(tiles, i, j, v.x);
                }
            }
        } else if (a.dcoef) {
            const int NG = a.P.n_dense;
            for (int t = tid; t < 2 * NG; t += NTHR) cf[t] = a.dcoef[(size_t)p * 2 * NG + t];
            __syncthreads();
            const int NF = NG - a.P.n_dense_diag;
            double imw = 0.0;
            for (int g = 0; g < NF; ++g) imw += fabs(cf[NG + g]) * __ldg(a.P.dense_gmax + g);
            const bool pt_real = a.tidyup ? imw < 1e-12 : imw == 0.0;
            if (!pt_real) not_real = 1;
            else {
                const size_t nn = (size_t)n * n;
                const double *dg = a.P.dense_tab + (size_t)NF * nn;
                for (int j = wp; j < n; j += NTHR / 32) {
                    const double *col = a.P.dense_tab + (size_t)j * n;
                    for (int i = j + lane; i < n; i += 128) {       // four rows per lane, their table loads back to back
                        double re[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
                        for (int g = 0; g < NF; ++g) {
                            const double cg = cf[g];
                            double t[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) t[u] = i + 32 * u < n ? __ldg(col + (size_t)g * nn + i + 32 * u) : 0.0;
#pragma unroll
                            for (int u = 0; u < 4; ++u) re[u] = fma(cg, t[u], re[u]);
                        }
                        if (i == j)
                            for (int g = NF; g < NG; ++g) re[0] = fma(cf[g], __ldg(dg + (size_t)(g - NF) * n + i), re[0]);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (i + 32 * u >= n) break;
                            double v = re[u];
                            if (a.tidyup) v = fabs(v) < 1e-12 ? 0.0 : v;
                            s1t_store(tiles, i + 32 * u, j, v);
                        }
                    }
                }
            }
        } else {
            cplx *s_coef = (cplx *)cf;
            for (int t = tid; t < a.P.n_terms; t += NTHR) s_coef[t] = a.coef[p * a.P.n_terms + t];
            __syncthreads();
            for (int idx = tid; idx < n * n; idx += NTHR) {
                const int j = idx / n, i = idx - j * n;
                if (i < j) continue;
                int id[SCQED_MAX_NODES], jd[SCQED_MAX_NODES];
                digits(a.P, i, id);
                digits(a.P, j, jd);
                cplx v = cconj(h_element(a.P, s_coef, jd, id));
                if (a.tidyup) v = tidy(v, 1e-12);
                not_real |= (v.y != 0.0);
                s1t_store(tiles, i, j, v.x);
            }
        }
        not_real = __syncthreads_or(not_real);
        if (tid == 0) { a.W.flags[p] = not_real; if (not_real) a.info[p] = -2; }
        if (not_real) continue;                              // info = -2: the caller re-runs in complex mode
        for (int q = tid; q < 2 * NB * PS + PS; q += NTHR) Vs[q] = 0.0;      // panels and Y (cf is dead now)
        __syncthreads();

        double *dd = a.W.dd + p * n, *ee = a.W.ee + p * n, *e2 = a.W.e2 + p * n, *tauo = (double *)a.W.tau + p * n;
        __shared__ double s_red[2][8], s_scal[2];
        for (int j0 = 0; j0 < n; j0 += NB) {
            const int pn = min(NB, n - j0);
            double v_r = 0.0, w_r = 0.0, wpiv = 0.0;          // PSER: this row's entries of the newest panel column
            for (int jj = 0; jj < pn; ++jj) {
                const int j = j0 + jj;
                if constexpr (PSER) {
                    const int r = tid;
                    constexpr int NW = NTHR / 32;
                    // ---- S1: column j made current (newest panel column from registers, older ones from the panels) ----
                    double c = 0.0;
                    if (r >= j && r < n) {
                        c = tiles[s1t_tile(r >> 3, j >> 3) + (r & 7) * 8 + (j & 7)];
                        for (int k = 0; k + 1 < jj; ++k)
                            c = fma(-Vs[k * PS + r], Ws[k * PS + j], fma(-Ws[k * PS + r], Vs[k * PS + j], c));
                        if (jj > 0) c = fma(-v_r, wpiv, c - w_r);                 // V[j][jj-1] = 1 (pivot row of column j-1)
                    }
                    if (r == j) dd[j] = c;
                    if (j == n - 1) {
                        if (tid == 0) { ee[j] = 0.0; e2[j] = 0.0; tauo[j] = 0.0; }
                        break;
                    }
                    if (r == j + 1) s_scal[0] = c;
                    double xn = (r >= j + 2 && r < n) ? c * c : 0.0;
                    xn = warp_sum(xn);
                    if (lane == 0) s_red[0][wp] = xn;
                    __syncthreads();
                    xn = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) xn += s_red[0][w];
                    double tau, beta, sc;
                    s1t_reflector_scalars(s_scal[0], xn, tau, beta, sc);
                    v_r = (r == j + 1) ? 1.0 : ((r > j + 1 && r < n) ? c * sc : 0.0);
                    if (r < PS) Vs[jj * PS + r] = v_r;
                    if (r > j && r < n) tiles[s1t_tile(r >> 3, j >> 3) + (r & 7) * 8 + (j & 7)] = v_r;
                    if (tid == 0) { ee[j] = beta; e2[j] = beta * beta; tauo[j] = tau; }
                    __syncthreads();
                    s1t_pass<NTHR>(tiles, Vs, Ws, Y, ts, n, NT, PS, j, jj, wp, lane, 0);
                    __syncthreads();
                    // ---- S3: corrections, w ----
                    double pr = 0.0;
                    if (r > j && r < n) {
                        double y = Y[r];
                        for (int k = 0; k < jj; ++k) y = fma(-Vs[k * PS + r], ts[k], fma(-Ws[k * PS + r], ts[NB + k], y));
                        pr = tau * y;
                    }
                    if (r == j + 1) s_scal[1] = pr;
                    double dt = warp_sum(pr * v_r);
                    if (lane == 0) s_red[1][wp] = dt;
                    __syncthreads();
                    dt = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) dt += s_red[1][w];
                    const double hc = -0.5 * tau * dt;
                    w_r = (r > j && r < n) ? fma(hc, v_r, pr) : 0.0;
                    wpiv = s_scal[1] + hc;                                        // w of the pivot row j + 1 (v = 1 there)
                    if (r < PS) Ws[jj * PS + r] = w_r;
                    continue;
                }
                double vq[4] = {0.0, 0.0, 0.0, 0.0}, tau = 0.0;
                // ---- S1 / S2 (warp 0, lane owns rows lane + 32 q): column j made current, reflector j ----
                if (wp == sw) {
                    const int Cj = j >> 3, cj = j & 7;
                    double c[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = lane + 32 * q;
                        c[q] = (r >= j && r < n) ? tiles[s1t_tile(r >> 3, Cj) + (r & 7) * 8 + cj] : 0.0;
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
                        if (lane == (j & 31)) dd[j] = cd;
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
                            if (r > j && r < n) tiles[s1t_tile(r >> 3, Cj) + (r & 7) * 8 + cj] = vq[q];     // kept for the back-transformation
                        }
                        if (lane == 0) { ee[j] = beta; e2[j] = beta * beta; tauo[j] = tau; }
                    } else if (lane == 0) { ee[j] = 0.0; e2[j] = 0.0; tauo[j] = 0.0; }
                }
                if (j == n - 1) break;                                            // only the last diagonal was left
                __syncthreads();
                s1t_pass<NTHR>(tiles, Vs, Ws, Y, ts, n, NT, PS, j, jj, wp, lane, rotate);
                __syncthreads();
                // ---- S3 (warp 0): corrections, w ----
                if (wp == sw) {
                    double y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = lane + 32 * q;
                        y[q] = (r < PS) ? Y[r] : 0.0;
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
            // ---- trailing update on the FP64 tensor cores: A22 -= V W^T + W V^T on the stored tiles ----
            const int b0 = j0 + NB;
            if (pn == NB && b0 < n) {
                __syncthreads();
                s1t_update<NTHR>(tiles, Vs, Ws, NT, PS, b0, wp, lane);
                __syncthreads();
            }
        }
        __syncthreads();
        // ---- Gershgorin bounds over the whole tridiagonal matrix (d, e come back from this CTA's own global writes) ----
        {
            double *dv = Vs, *ev = Vs + n, *red = Vs + 2 * n, *bounds = red + 96;
            for (int i = tid; i < n; i += NTHR) { dv[i] = dd[i]; ev[i] = ee[i]; }
            __syncthreads();
            tridiag_bounds(dv, ev, n, bounds, red);
            if (tid < 4) a.W.bnd[p * 4 + tid] = bounds[tid];
        }
        if (a.want_vec) {
            // reflector j: column j, rows j+1 .. n-1 (s1_backtr_kernel reads nothing else)
            double *dst = (double *)a.W.Vst + (size_t)p * n * n;
            for (int j = wp; j < n - 1; j += NTHR / 32) {
                const int Cj = j >> 3, cj = j & 7;
                for (int i = j + 1 + lane; i < n; i += 32) dst[(size_t)j * n + i] = tiles[s1t_tile(i >> 3, Cj) + (i & 7) * 8 + cj];
            }
        }
        __syncthreads();
    }
}

}  // namespace scqed
