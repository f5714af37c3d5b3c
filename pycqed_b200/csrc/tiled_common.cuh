// tiled_common.cuh -- device functions shared by the tile-packed tridiagonalisations: small_tiled.cuh (matrix in shared
// memory, n <= 224) and dense_tiled.cuh (matrix in global memory / L2, 224 < n <= 2048).  `tiles` may point to shared or
// global memory (generic loads): 8 x 8 row-major tiles (R, C), R >= C, of the lower triangle, diagonal tiles complete.
#pragma once
#include "common.cuh"

namespace scqed {

#define S1T_NB 4

__device__ __forceinline__ void tc_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// dlarfg scalars without IEEE sqrt / divisions (they were 6.5 % of the samples of this kernel): rsqrt + one Newton step for
// the norm, MUFU reciprocal seed + two Newton steps for 1 / (alpha - beta); results within ~2 ulp, which is inside the
// backward error of the reduction.
__device__ __forceinline__ void s1t_reflector_scalars(double alpha, double xn, double &tau, double &beta, double &sc) {
    if (xn == 0.0) { tau = 0.0; beta = alpha; sc = 0.0; return; }
    const double s2 = fma(alpha, alpha, xn);
    const double rs = rsqrt(s2);
    double nrm = s2 * rs;
    nrm = fma(0.5 * rs, fma(-nrm, nrm, s2), nrm);
    beta = -copysign(nrm, alpha);
    tau = fma(fabs(alpha), rs, 1.0);                     // (beta - alpha) / beta = 1 + |alpha| / norm
    const double d = alpha - beta;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    sc = fma(y, e, y);
}

__device__ __forceinline__ int s1t_tile(int R, int C) { return ((R * (R + 1) >> 1) + C) << 6; }

// P: y = A_stale v from the stored half (item R = tile row R + tile column R, both feed rows 8R..8R+7), and the panel dot
// products t1[k] = W[:,k].v, t2[k] = V[:,k].v for the corrections.
template <int NTHR, int UNR = 4>
__device__ __forceinline__ void s1t_pass(const double *tiles, const double *Vs, const double *Ws, double *Y, double *ts,
                                         int n, int NT, int PS, int j, int jj, int wp, int lane, int rotate) {
    constexpr int NB = S1T_NB;
    const int g8 = lane >> 2, t4 = lane & 3, c8 = lane & 7, r4 = lane >> 3;

                    const double *vj = Vs + jj * PS;
                    const int T0 = (j + 1) >> 3;
                    // rotate == 2: scheduler 0 (warps 0 and 4) is left to the serial warps of the co-resident CTAs; the six
                    // other warps share the mat-vec items, warp 4 forms the panel dot products
                    const int pw = rotate == 2 ? (wp & 3) - 1 + 3 * (wp >> 2) : wp;          // 1,2,3,5,6,7 -> 0..5
                    const int npw = rotate == 2 ? 6 : NTHR / 32;
                    const bool in_pass = rotate != 2 || (wp & 3) != 0;
                    for (int R = T0 + pw; in_pass && R < NT; R += npw) {
                        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
                        const double *trow = tiles + s1t_tile(R, T0) + g8 * 8 + 2 * t4;
                        const double *vr = vj + 8 * T0 + 2 * t4;
#pragma unroll UNR
                        for (int C = T0; C <= R; ++C, trow += 64, vr += 8) {
                            const double2 x = *reinterpret_cast<const double2 *>(trow);
                            const double2 vv = *reinterpret_cast<const double2 *>(vr);
                            a0 = fma(x.x, vv.x, a0);
                            a1 = fma(x.y, vv.y, a1);
                        }
                        // tile column R: lane (r4, c8) reads T[r4][c8] and T[r4 + 4][c8] of tile (R2, R)
                        const double *tt = tiles + s1t_tile(R + 1, R) + 8 * r4 + c8;
                        const double *vc = vj + 8 * (R + 1) + r4;
#pragma unroll UNR
                        for (int R2 = R + 1; R2 < NT; ++R2) {
                            const double x0 = tt[0], x1 = tt[32];
                            b0 = fma(x0, vc[0], b0);
                            b1 = fma(x1, vc[4], b1);
                            tt += (R2 + 1) << 6;
                            vc += 8;
                        }
                        double yr = a0 + a1, yc = b0 + b1;
                        yr += __shfl_xor_sync(0xffffffffu, yr, 1);
                        yr += __shfl_xor_sync(0xffffffffu, yr, 2);               // row 8R + g8 in every lane of the quad
                        yc += __shfl_xor_sync(0xffffffffu, yc, 8);
                        yc += __shfl_xor_sync(0xffffffffu, yc, 16);              // row 8R + c8 in the lanes with that c8
                        yr = __shfl_sync(0xffffffffu, yr, 4 * c8);               // ... and the row part of the same row
                        if (lane < 8) Y[8 * R + lane] = yr + yc;
                    }
                    for (int kt = 0; kt < jj; ++kt) {
                        // panel column kt: warp 4 takes them all (rotate == 2), else warps 7, 6, 5 one each
                        if (rotate == 2 ? wp != 4 : wp != NTHR / 32 - 1 - kt % (NTHR / 32)) continue;
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

// Trailing update on the FP64 tensor cores: A22 -= V W^T + W V^T on the stored tiles (R, C), R >= C >= b0 / 8.  A tile IS
// the C fragment of mma.sync.m8n8k4.f64 (lane (g, t) <-> T[g][2t], T[g][2t+1]); A / B operands are 8 x 4 slices of the panels.
template <int NTHR, int UPD = 4>
__device__ __forceinline__ void s1t_update(double *tiles, const double *Vs, const double *Ws, int NT, int PS, int b0, int wp, int lane) {
    const int g8 = lane >> 2, t4 = lane & 3;
                const int B0 = b0 >> 3;                       // first tile row / column that holds trailing elements
                const int K = NT - B0;
                // tile rows in balanced pairs (r, K-1-r): r + 1 and K - r tiles; the A-operand fragments are per row
                for (int it = wp; it < (K + 1) / 2; it += NTHR / 32) {
                    for (int half = 0; half < 2; ++half) {
                        const int r = half == 0 ? it : K - 1 - it;
                        if (half == 1 && r == it) break;
                        const int R = B0 + r;
                        // rows / columns below b0 inside the first tile row / column are not part of the trailing block
                        // (they hold reflectors of this panel): their operand fragments are zeroed
                        const bool rok = 8 * R + g8 >= b0;
                        const double va = rok ? -Vs[t4 * PS + 8 * R + g8] : 0.0, wa = rok ? -Ws[t4 * PS + 8 * R + g8] : 0.0;
                        double2 *tp = reinterpret_cast<double2 *>(tiles + s1t_tile(R, B0) + g8 * 8 + 2 * t4);
                        const double *vbp = Vs + t4 * PS + 8 * B0 + g8, *wbp = Ws + t4 * PS + 8 * B0 + g8;
                        int C = B0;
                        // UPD tiles per iteration, all accumulator loads issued before the first DMMA (global tiles: one
                        // tile in flight per warp left the update latency bound)
                        for (; C + UPD <= R + 1; C += UPD, tp += 32 * UPD, vbp += 8 * UPD, wbp += 8 * UPD) {
                            double2 d[UPD];
                            double vb[UPD], wb[UPD];
#pragma unroll
                            for (int u = 0; u < UPD; ++u) {
                                d[u] = tp[32 * u];
                                const bool cok = 8 * (C + u) + g8 >= b0;
                                vb[u] = cok ? vbp[8 * u] : 0.0;
                                wb[u] = cok ? wbp[8 * u] : 0.0;
                            }
#pragma unroll
                            for (int u = 0; u < UPD; ++u) {
                                tc_dmma(d[u].x, d[u].y, va, wb[u]);
                                tc_dmma(d[u].x, d[u].y, wa, vb[u]);
                                tp[32 * u] = d[u];
                            }
                        }
                        for (; C <= R; ++C, tp += 32, vbp += 8, wbp += 8) {
                            const bool cok = 8 * C + g8 >= b0;
                            const double vb = cok ? *vbp : 0.0, wb = cok ? *wbp : 0.0;
                            double2 d = *tp;
                            tc_dmma(d.x, d.y, va, wb);
                            tc_dmma(d.x, d.y, wa, vb);
                            *tp = d;
                        }
                    }
                }
}

}  // namespace scqed
