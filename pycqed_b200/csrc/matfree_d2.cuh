// matfree_d2.cuh -- Kronecker mat-vec for TWO-node plans with dense single-node operators
// (inductively / capacitively coupled oscillator-basis nodes: the fluxonium atom, C5).
//
// mf_apply walks every term per output element; with the oscillator displacement operators D, D^dagger
// (dense d x d) on both nodes that is four dense mode products of complex arithmetic per mat-vec.
// Here all single-node terms of node k (charge^2, flux^2, D, D^dagger, constants) are first COMBINED per
// sweep point into one d_k x d_k matrix
//        H_k(p) = sum_{t on node k only} c_t(p) F_t            (mf_d2_build, once per point)
// so that, with x viewed as a d0 x d1 matrix X,
//        H x  =  H_0 X  +  X H_1^T  +  (coupling terms, sparse factors on both nodes).
// H_k(p) is Hermitian and for these circuits REAL (c D + conj(c) D^dagger = 2 Re(c D) for symmetric D):
// two real x complex mode products instead of four complex ones, 4x fewer flops, and the products run
// as register-tiled GEMM slabs: a CTA owns RB rows of X for one (point, vector); thread = column i1;
// the H_0 row block and the X row block are staged in shared memory and broadcast, X / H_1 stream
// coalesced from L2.  Algorithmic work per mat-vec and vector: 2 n (d0 + d1) real-complex MACs
// (n = d0 d1), traffic 32 n bytes.
#pragma once
#include <type_traits>
#include "matfree.cuh"

namespace scqed {

#define D2_RB 16

// H_k(p) for k = 0, 1: re / im planes [n_pts][d0*d0 + d1*d1]; max |Im| accumulated into imax[0].
static __global__ void __launch_bounds__(256)
mf_d2_build(MfArgs a, double *__restrict__ Hre, double *__restrict__ Him, double *__restrict__ imax) {
    const long long p = blockIdx.y;
    const PlanDev &P = a.P;
    const int d0 = P.dims[0], d1 = P.dims[1];
    const int tot = d0 * d0 + d1 * d1;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double my = 0.0;
    if (e < tot) {
        const int k = e < d0 * d0 ? 0 : 1;
        const int dk = k == 0 ? d0 : d1;
        const int q = k == 0 ? e : e - d0 * d0;
        const int r = q / dk, c = q - r * dk;
        const cplx *coef = a.coef + p * P.n_terms;
        cplx acc = cmake(0.0, 0.0);
        for (int t = 0; t < P.n_terms; ++t) {
            const int *tm = P.term_meta + t * 7;
            const int nn = __ldg(tm);
            if (nn == 0) { if (k == 0 && r == c) acc = cadd(acc, coef[t]); continue; }
            if (nn != 1 || __ldg(tm + 1) != k) continue;
            const cplx f = __ldg(P.factor_data + __ldg(P.factor_offset + __ldg(tm + 4)) + (size_t)r * dk + c);
            cfma(acc, coef[t], f);
        }
        Hre[p * tot + e] = acc.x;
        Him[p * tot + e] = acc.y;
        my = fabs(acc.y);
    }
    my = fmax(my, __shfl_xor_sync(0xffffffffu, my, 16));
    my = fmax(my, __shfl_xor_sync(0xffffffffu, my, 8));
    my = fmax(my, __shfl_xor_sync(0xffffffffu, my, 4));
    my = fmax(my, __shfl_xor_sync(0xffffffffu, my, 2));
    my = fmax(my, __shfl_xor_sync(0xffffffffu, my, 1));
    if ((threadIdx.x & 31) == 0 && my > 0.0) atomicMax((unsigned long long *)imax, (unsigned long long)__double_as_longlong(my));
    // coupling terms: real iff (factor-entry product real and c real) or (product imaginary and c imaginary);
    // d2_phase: 0 real, 1 imaginary, 2 mixed.  Violations raise imax[1].
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const cplx *coef = a.coef + p * P.n_terms;
        double viol = 0.0;
        for (int q = 0; q < P.d2_nmulti; ++q) {
            const cplx c = coef[__ldg(P.d2_cidx + q)];
            const int ph = __ldg(P.d2_phase + q);
            const double bad = ph == 0 ? fabs(c.y) : (ph == 1 ? fabs(c.x) : fabs(c.x) + fabs(c.y));
            viol = fmax(viol, bad);
        }
        if (viol > 0.0) atomicMax((unsigned long long *)(imax + 1), (unsigned long long)__double_as_longlong(viol));
    }
}

// Tighter spectral upper bound from the combined node Hamiltonians: ||H|| <= ||H0||_inf + ||H1||_inf + sum over
// coupling terms |c| ||F_a||_inf ||F_b||_inf.  The term-by-term bound of mf_prepare counts charge^2 and flux^2
// separately although their off-diagonals cancel in H_k (4303 vs 2566 for the 180 x 180 atom; the true top is
// 2499): a narrower filter interval means fewer filter applications.  One block per point.
static __global__ void __launch_bounds__(256)
mf_d2_bound(MfArgs a, const double *__restrict__ Hre, const double *__restrict__ Him) {
    __shared__ double smax[2][8];
    const long long p = blockIdx.x;
    const PlanDev &P = a.P;
    const int d0 = P.dims[0], d1 = P.dims[1], tid = threadIdx.x;
    const size_t tot = (size_t)d0 * d0 + (size_t)d1 * d1;
    const double *re = Hre + p * tot, *im = Him + p * tot;
    double m0 = 0.0, m1 = 0.0;
    for (int r = tid; r < d0 + d1; r += blockDim.x) {
        const int k = r < d0 ? 0 : 1, dk = k ? d1 : d0, rr = k ? r - d0 : r;
        const size_t base = (k ? (size_t)d0 * d0 : 0) + (size_t)rr * dk;
        double sum = 0.0;
        for (int c = 0; c < dk; ++c) sum += sqrt(re[base + c] * re[base + c] + im[base + c] * im[base + c]);
        if (k) m1 = fmax(m1, sum); else m0 = fmax(m0, sum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    }
    if ((tid & 31) == 0) { smax[0][tid >> 5] = m0; smax[1][tid >> 5] = m1; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { m0 = fmax(m0, smax[0][w]); m1 = fmax(m1, smax[1][w]); }
        double ub = m0 + m1;
        const cplx *coef = a.coef + p * P.n_terms;
        for (int q = 0; q < P.d2_nmulti; ++q) {
            const int *tm = P.d2_meta + q * 7;
            ub += sqrt(cabs2(coef[__ldg(P.d2_cidx + q)])) * __ldg(P.fac_rowsum + __ldg(tm + 4)) * __ldg(P.fac_rowsum + __ldg(tm + 5));
        }
        ub *= 1.0 + 1e-10;
        if (ub < a.fpar[p * 4 + 1]) a.fpar[p * 4 + 1] = ub;
    }
}

// max |Im X| over the block of every active point -> imax[2]  (decides whether the real-vector kernels apply)
static __global__ void __launch_bounds__(256)
mf_d2_imag_max(MfArgs a, const cplx *__restrict__ X, double *__restrict__ imax) {
    const long long p = blockIdx.x;
    if (!a.active[p]) return;
    const size_t tot = (size_t)a.b * a.n;
    const cplx *x = X + p * tot;
    double my = 0.0;
    for (size_t i = threadIdx.x; i < tot; i += blockDim.x) my = fmax(my, fabs(x[i].y));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my = fmax(my, __shfl_xor_sync(0xffffffffu, my, o));
    if ((threadIdx.x & 31) == 0 && my > 0.0) atomicMax((unsigned long long *)(imax + 2), (unsigned long long)__double_as_longlong(my));
}

// MODE 0: HX = H X.  MODE 1: Chebyshev step m (same recurrence as mf_cheb_step), `out` holds y_{m-2} and
// receives y_m.  grid (ceil(d0 / D2_RB), b, n_pts), block = d1 rounded up to a warp multiple (<= 256).
// REALX: H real AND the vectors real (exactly zero imaginary parts, checked by mf_d2_imag_max before every
// use): the imaginary half of the arithmetic is skipped.
template <int MODE, bool REALH, bool REALX>
static __global__ void __launch_bounds__(256)
mf_d2_kernel(MfArgs a, const cplx *__restrict__ X, cplx *out, int m, const double *__restrict__ Hre,
             const double *__restrict__ Him) {
    extern __shared__ __align__(16) unsigned char d2smem[];
    const long long p = blockIdx.z;
    if (!a.active[p]) return;
    const PlanDev &P = a.P;
    const int d0 = P.dims[0], d1 = P.dims[1], n = a.n;
    const int j = blockIdx.y, R0 = blockIdx.x * D2_RB, tid = threadIdx.x;
    const int rb = min(D2_RB, d0 - R0);
    const int tot = d0 * d0 + d1 * d1;
    const double *H0re = Hre + p * tot, *H0im = Him + p * tot;
    const double *H1re = H0re + (size_t)d0 * d0, *H1im = H0im + (size_t)d0 * d0;
    // shared: H0 rows [D2_RB][d0] (re, then im when complex), X rows [D2_RB][d1], compact coupling coefficients
    double *sHre = (double *)d2smem;
    double *sHim = sHre + (size_t)D2_RB * d0;
    cplx *sX = (cplx *)(sHim + (REALH ? 0 : (size_t)D2_RB * d0));
    cplx *sC = sX + (size_t)D2_RB * d1;
    double *sXr = (double *)(sC + P.d2_nmulti + 1);          // REALX: real parts of the X rows, pitch d1e (even)
    const int d1e = (d1 + 1) & ~1;
    const cplx *x = X + (p * a.b + j) * n;
    for (int q = tid; q < rb * d0; q += blockDim.x) {
        const int r = q / d0, c = q - r * d0;
        sHre[r * d0 + c] = H0re[(size_t)(R0 + r) * d0 + c];
        if (!REALH) sHim[r * d0 + c] = H0im[(size_t)(R0 + r) * d0 + c];
    }
    for (int q = tid; q < rb * d1; q += blockDim.x) {
        const cplx v = x[(size_t)R0 * d1 + q];
        sX[q] = v;
        if (REALX) { const int r = q / d1; sXr[r * d1e + (q - r * d1)] = v.x; }
    }
    for (int q = tid; q < P.d2_nmulti; q += blockDim.x) sC[q] = a.coef[p * P.n_terms + __ldg(P.d2_cidx + q)];
    __syncthreads();
    const int i1 = tid;
    if (i1 >= d1) return;
    cplx acc[D2_RB];
#pragma unroll
    for (int r = 0; r < D2_RB; ++r) acc[r] = cmake(0.0, 0.0);
    int bA = 0, bB = 0;
    if (REALX && (d0 & 1) == 0) {
        // real vectors: four independent global loads per step (the kernel is bound by the latency of this L2
        // stream), H0 through 16-byte shared-memory broadcasts (two columns per load)
        for (; bA + 4 <= d0; bA += 4) {
            const double x0 = x[(size_t)bA * d1 + i1].x, x1 = x[(size_t)(bA + 1) * d1 + i1].x;
            const double x2 = x[(size_t)(bA + 2) * d1 + i1].x, x3 = x[(size_t)(bA + 3) * d1 + i1].x;
#pragma unroll
            for (int r = 0; r < D2_RB; ++r) {
                const double2 h01 = *(const double2 *)(sHre + r * d0 + bA);
                const double2 h23 = *(const double2 *)(sHre + r * d0 + bA + 2);
                double t = acc[r].x;
                t = fma(h01.x, x0, t); t = fma(h01.y, x1, t); t = fma(h23.x, x2, t); t = fma(h23.y, x3, t);
                acc[r].x = t;
            }
        }
        for (; bB + 4 <= d1; bB += 4) {
            const double h0 = H1re[(size_t)bB * d1 + i1], h1 = H1re[(size_t)(bB + 1) * d1 + i1];
            const double h2 = H1re[(size_t)(bB + 2) * d1 + i1], h3 = H1re[(size_t)(bB + 3) * d1 + i1];
#pragma unroll
            for (int r = 0; r < D2_RB; ++r) {
                const double2 v01 = *(const double2 *)(sXr + r * d1e + bB);
                const double2 v23 = *(const double2 *)(sXr + r * d1e + bB + 2);
                double t = acc[r].x;
                t = fma(h0, v01.x, t); t = fma(h1, v01.y, t); t = fma(h2, v23.x, t); t = fma(h3, v23.y, t);
                acc[r].x = t;
            }
        }
    }
    // Y[R0 + r][i1] += sum_b H0[R0 + r][b] X[b][i1]   (X streams coalesced from L2, H0 broadcast from shared memory)
#pragma unroll 2
    for (int b = bA; b < d0; ++b) {
        const cplx xb = x[(size_t)b * d1 + i1];
#pragma unroll
        for (int r = 0; r < D2_RB; ++r) {
            const double hr = sHre[r * d0 + b];
            acc[r].x = fma(hr, xb.x, acc[r].x);
            if (!REALX) acc[r].y = fma(hr, xb.y, acc[r].y);
            if (!REALH) {
                const double hi = sHim[r * d0 + b];
                acc[r].x = fma(-hi, xb.y, acc[r].x);
                acc[r].y = fma(hi, xb.x, acc[r].y);
            }
        }
    }
    // Y[R0 + r][i1] += sum_b X[R0 + r][b] H1[i1][b];  H1 Hermitian: H1[i1][b] = conj(H1[b][i1]) (row b is contiguous)
#pragma unroll 2
    for (int b = bB; b < d1; ++b) {
        const double hr = H1re[(size_t)b * d1 + i1];
        const double hi = REALH ? 0.0 : -H1im[(size_t)b * d1 + i1];
#pragma unroll
        for (int r = 0; r < D2_RB; ++r) {
            const cplx xv = sX[r * d1 + b];
            acc[r].x = fma(hr, xv.x, acc[r].x);
            if (!REALX) acc[r].y = fma(hr, xv.y, acc[r].y);
            if (!REALH) {
                acc[r].x = fma(-hi, xv.y, acc[r].x);
                acc[r].y = fma(hi, xv.x, acc[r].y);
            }
        }
    }
    ChebPar q;
    double s = 0.0, t = 0.0;
    if (MODE == 1) {
        q = cheb_par(a.fpar + p * 4);
        if (m == 1) s = q.sigma1 / q.e;
        else {
            double sigma = q.sigma1;
            for (int u = 2; u < m; ++u) sigma = 1.0 / (2.0 / q.sigma1 - sigma);
            const double sigma2 = 1.0 / (2.0 / q.sigma1 - sigma);
            s = 2.0 * sigma2 / q.e;
            t = sigma * sigma2;
        }
    }
    cplx *o = out + (p * a.b + j) * n;
#pragma unroll
    for (int r = 0; r < D2_RB; ++r) {
        if (r >= rb) break;
        const int i = (R0 + r) * d1 + i1;
        cplx h = acc[r];
        if (P.d2_nmulti > 0) h = cadd(h, mf_apply_terms(P, P.d2_meta, 0, P.d2_nmulti, sC, x, i));
        if (REALX) h.y = 0.0;
        if (MODE == 0) o[i] = h;
        else {
            const cplx yi = sX[r * d1 + i1];
            if (m == 1) o[i] = cmake(s * (h.x - q.c * yi.x), s * (h.y - q.c * yi.y));
            else {
                const cplx pi = o[i];
                o[i] = cmake(s * (h.x - q.c * yi.x) - t * pi.x, s * (h.y - q.c * yi.y) - t * pi.y);
            }
        }
    }
}

// ---- FP64 tensor-core variant for real H and real vectors --------------------------------------------------------
// Both mode products as DMMA.8x8x4 tiles (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind).  A CTA owns D2_RB = 16
// rows of Y (two 8-row tiles) and all d1 columns; warp w owns the 8-column tiles w, w + 8, w + 16, ...  The A operands
// (H0 rows, then the X rows) come from shared memory, the B operands (X rows, then H1 rows) straight from L2 with the
// next k-step prefetched into registers.  One DMMA carries 256 multiply-adds: 8x fewer issue slots than the DFMA
// kernel and 12 accumulator registers per lane instead of 16-32, so four CTAs fit an SM.
#define D2_NT 4      // column tiles per warp (8 warps x 4 x 8 = 256 columns max)

__device__ __forceinline__ void d2_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
static __global__ void __launch_bounds__(256)
mf_d2_dmma_kernel(MfArgs a, const cplx *__restrict__ X, cplx *out, int m, const double *__restrict__ Hre) {
    extern __shared__ __align__(16) unsigned char d2smem[];
    const long long p = blockIdx.z;
    if (!a.active[p]) return;
    const PlanDev &P = a.P;
    const int d0 = P.dims[0], d1 = P.dims[1], n = a.n;
    const int j = blockIdx.y, R0 = blockIdx.x * D2_RB, tid = threadIdx.x;
    const int rb = min(D2_RB, d0 - R0);
    const size_t tot = (size_t)d0 * d0 + (size_t)d1 * d1;
    const double *H0re = Hre + p * tot, *H1re = H0re + (size_t)d0 * d0;
    // pitches = 4 (mod 16) doubles: the 8 rows x 4 k of an A fragment cover the 16 64-bit banks exactly twice
    const int d0p = d0 + ((20 - (d0 & 15)) & 15), d1p = d1 + ((20 - (d1 & 15)) & 15);
    double *sH = (double *)d2smem;                            // [D2_RB][d0p]  H0 rows (zero beyond rb)
    double *sXr = sH + (size_t)D2_RB * d0p;                   // [D2_RB][d1p]  real parts of the X rows
    cplx *sC = (cplx *)(sXr + (size_t)D2_RB * d1p);
    const cplx *x = X + (p * a.b + j) * n;
    for (int q = tid; q < D2_RB * d0; q += blockDim.x) {
        const int r = q / d0, c = q - r * d0;
        sH[r * d0p + c] = r < rb ? H0re[(size_t)(R0 + r) * d0 + c] : 0.0;
    }
    for (int q = tid; q < D2_RB * d1; q += blockDim.x) {
        const int r = q / d1, c = q - r * d1;
        sXr[r * d1p + c] = r < rb ? x[(size_t)(R0 + r) * d1 + c].x : 0.0;
    }
    for (int q = tid; q < P.d2_nmulti; q += blockDim.x) sC[q] = a.coef[p * P.n_terms + __ldg(P.d2_cidx + q)];
    __syncthreads();
    const int warp = tid >> 5, l = tid & 31, ar = l >> 2, ak = l & 3;
    const int ntiles = (d1 + 7) >> 3;
    double c[2][D2_NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < D2_NT; ++nt) c[mt][nt][0] = c[mt][nt][1] = 0.0;
    int col[D2_NT];
#pragma unroll
    for (int nt = 0; nt < D2_NT; ++nt) col[nt] = (warp + 8 * nt) < ntiles ? (warp + 8 * nt) * 8 + ar : -1;
    // ---- Y[R0 + r][i1] += sum_b H0[R0 + r][b] X[b][i1] ----------------------------------------------------------------
    {
        double bn[D2_NT];
#pragma unroll
        for (int nt = 0; nt < D2_NT; ++nt) bn[nt] = (col[nt] >= 0 && col[nt] < d1 && ak < d0) ? x[(size_t)ak * d1 + col[nt]].x : 0.0;
        for (int k0 = 0; k0 < d0; k0 += 4) {
            double bc[D2_NT];
#pragma unroll
            for (int nt = 0; nt < D2_NT; ++nt) bc[nt] = bn[nt];
            const int kn = k0 + 4 + ak;
#pragma unroll
            for (int nt = 0; nt < D2_NT; ++nt) bn[nt] = (col[nt] >= 0 && col[nt] < d1 && kn < d0) ? x[(size_t)kn * d1 + col[nt]].x : 0.0;
            const int kk = k0 + ak;
            const double a0 = kk < d0 ? sH[ar * d0p + kk] : 0.0;
            const double a1 = kk < d0 ? sH[(8 + ar) * d0p + kk] : 0.0;
#pragma unroll
            for (int nt = 0; nt < D2_NT; ++nt) {
                if (col[nt] >= 0) {
                    d2_dmma(c[0][nt][0], c[0][nt][1], a0, bc[nt]);
                    d2_dmma(c[1][nt][0], c[1][nt][1], a1, bc[nt]);
                }
            }
        }
    }
    // ---- Y[R0 + r][i1] += sum_b X[R0 + r][b] H1[b][i1]   (H1 real symmetric) -------------------------------------------
    {
        double bn[D2_NT];
#pragma unroll
        for (int nt = 0; nt < D2_NT; ++nt) bn[nt] = (col[nt] >= 0 && col[nt] < d1 && ak < d1) ? H1re[(size_t)ak * d1 + col[nt]] : 0.0;
        for (int k0 = 0; k0 < d1; k0 += 4) {
            double bc[D2_NT];
#pragma unroll
            for (int nt = 0; nt < D2_NT; ++nt) bc[nt] = bn[nt];
            const int kn = k0 + 4 + ak;
#pragma unroll
            for (int nt = 0; nt < D2_NT; ++nt) bn[nt] = (col[nt] >= 0 && col[nt] < d1 && kn < d1) ? H1re[(size_t)kn * d1 + col[nt]] : 0.0;
            const int kk = k0 + ak;
            const double a0 = kk < d1 ? sXr[ar * d1p + kk] : 0.0;
            const double a1 = kk < d1 ? sXr[(8 + ar) * d1p + kk] : 0.0;
#pragma unroll
            for (int nt = 0; nt < D2_NT; ++nt) {
                if (col[nt] >= 0) {
                    d2_dmma(c[0][nt][0], c[0][nt][1], a0, bc[nt]);
                    d2_dmma(c[1][nt][0], c[1][nt][1], a1, bc[nt]);
                }
            }
        }
    }
    ChebPar q;
    double s = 0.0, t = 0.0;
    if (MODE == 1) {
        q = cheb_par(a.fpar + p * 4);
        if (m == 1) s = q.sigma1 / q.e;
        else {
            double sigma = q.sigma1;
            for (int u = 2; u < m; ++u) sigma = 1.0 / (2.0 / q.sigma1 - sigma);
            const double sigma2 = 1.0 / (2.0 / q.sigma1 - sigma);
            s = 2.0 * sigma2 / q.e;
            t = sigma * sigma2;
        }
    }
    cplx *o = out + (p * a.b + j) * n;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int r = mt * 8 + ar;
        if (r >= rb) continue;
#pragma unroll
        for (int nt = 0; nt < D2_NT; ++nt) {
            if ((warp + 8 * nt) >= ntiles) continue;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int cc = (warp + 8 * nt) * 8 + 2 * ak + h2;
                if (cc >= d1) continue;
                const int i = (R0 + r) * d1 + cc;
                double h = c[mt][nt][h2];
                if (P.d2_nmulti > 0) h += mf_apply_terms(P, P.d2_meta, 0, P.d2_nmulti, sC, x, i).x;
                if (MODE == 0) o[i] = cmake(h, 0.0);
                else {
                    const double yi = sXr[r * d1p + cc];
                    if (m == 1) o[i] = cmake(s * (h - q.c * yi), 0.0);
                    else o[i] = cmake(s * (h - q.c * yi) - t * o[i].x, 0.0);
                }
            }
        }
    }
}

// ---- TMA-pipelined variant for real H and real vectors (the production path of the 180 x 180 atom) --------------
// The plain kernel above is bound by the latency of its L2 streams (X rows in the first mode product, H1 rows in
// the second: long-scoreboard stalls at 18 % occupancy).  Here those streams go through the TMA: one thread issues
// 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) of D2_BK rows at a time into a D2_ST-stage shared-memory
// ring; a "full" mbarrier per stage is completed by the copy engine, an "empty" mbarrier per stage by one arrival
// per warp, so there is no block-wide barrier inside the mode products.  Same arithmetic, same results.
#define D2_BK 4
#define D2_ST 4

__device__ __forceinline__ unsigned d2_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void d2_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(d2_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void d2_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(d2_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void d2_tma_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    const unsigned b = d2_smem_addr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d2_smem_addr(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void d2_mbar_wait(unsigned long long *bar, int parity) {
    const unsigned b = d2_smem_addr(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "D2_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D2_DONE_%=;\n"
        "bra D2_WAIT_%=;\n"
        "D2_DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity) : "memory");
}

// One pass of the ring over `total` rows of `src` (row_bytes each): calls body(chunk rows pointer, b0, rows).
// `it` is the running chunk counter shared by both mode products (stage = it % D2_ST, parity from it / D2_ST).
template <typename Body>
__device__ __forceinline__ void d2_ring(const unsigned char *src, int total, int row_bytes, unsigned char *stages,
                                        size_t stage_bytes, unsigned long long *full, unsigned long long *empty,
                                        int &it, int tid, Body body) {
    const int nch = (total + D2_BK - 1) / D2_BK;
    auto issue = [&](int c, int gi) {                      // chunk c of this pass, global chunk index gi
        const int s = gi % D2_ST;
        if (gi >= D2_ST) d2_mbar_wait(&empty[s], ((gi / D2_ST) - 1) & 1);      // all warps released the stage
        const int b0 = c * D2_BK, rows = min(D2_BK, total - b0);
        d2_tma_load(stages + s * stage_bytes, src + (size_t)b0 * row_bytes, (unsigned)(rows * row_bytes), &full[s]);
    };
    if (tid == 0) for (int c = 0; c < D2_ST - 1 && c < nch; ++c) issue(c, it + c);
    for (int c = 0; c < nch; ++c) {
        const int gi = it + c, s = gi % D2_ST;
        if (tid == 0 && c + D2_ST - 1 < nch) issue(c + D2_ST - 1, gi + D2_ST - 1);
        d2_mbar_wait(&full[s], (gi / D2_ST) & 1);
        body(stages + s * stage_bytes, c * D2_BK, min(D2_BK, total - c * D2_BK));
        __syncwarp();
        if ((tid & 31) == 0) d2_mbar_arrive(&empty[s]);
    }
    it += nch;
}

// shared: ring of D2_ST stages (D2_BK * d1 * 16 B each), H0 rows [D2_RB][d0], X rows (real parts) [D2_RB][d1e],
// coupling coefficients
template <int MODE>
static __global__ void __launch_bounds__(256)
mf_d2_tma_kernel(MfArgs a, const cplx *__restrict__ X, cplx *out, int m, const double *__restrict__ Hre) {
    extern __shared__ __align__(128) unsigned char d2smem[];
    __shared__ __align__(8) unsigned long long full[D2_ST], empty[D2_ST];
    const long long p = blockIdx.z;
    if (!a.active[p]) return;
    const PlanDev &P = a.P;
    const int d0 = P.dims[0], d1 = P.dims[1], n = a.n;
    const int j = blockIdx.y, R0 = blockIdx.x * D2_RB, tid = threadIdx.x;
    const int rb = min(D2_RB, d0 - R0);
    const size_t tot = (size_t)d0 * d0 + (size_t)d1 * d1;
    const double *H0re = Hre + p * tot, *H1re = H0re + (size_t)d0 * d0;
    const int d1e = (d1 + 1) & ~1;
    const size_t stage_bytes = ((size_t)D2_BK * d1 * 16 + 127) & ~(size_t)127;
    double *sHre = (double *)(d2smem + D2_ST * stage_bytes);
    double *sXr = sHre + (size_t)D2_RB * d0;
    cplx *sC = (cplx *)(sXr + (size_t)D2_RB * d1e);
    const cplx *x = X + (p * a.b + j) * n;
    if (tid == 0) {
        for (int s = 0; s < D2_ST; ++s) { d2_mbar_init(&full[s], 1); d2_mbar_init(&empty[s], (int)(blockDim.x >> 5)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int q = tid; q < rb * d0; q += blockDim.x) { const int r = q / d0, c = q - r * d0; sHre[r * d0 + c] = H0re[(size_t)(R0 + r) * d0 + c]; }
    for (int q = tid; q < rb * d1; q += blockDim.x) { const int r = q / d1; sXr[r * d1e + (q - r * d1)] = x[(size_t)R0 * d1 + q].x; }
    for (int q = tid; q < P.d2_nmulti; q += blockDim.x) sC[q] = a.coef[p * P.n_terms + __ldg(P.d2_cidx + q)];
    __syncthreads();
    const int i1 = tid;
    const bool on = i1 < d1;
    double acc[D2_RB];
#pragma unroll
    for (int r = 0; r < D2_RB; ++r) acc[r] = 0.0;
    int it = 0;
    // Y[R0 + r][i1] += sum_b H0[R0 + r][b] X[b][i1]
    d2_ring((const unsigned char *)x, d0, d1 * 16, d2smem, stage_bytes, full, empty, it, tid,
            [&](const unsigned char *buf, int b0, int rows) {
                if (!on) return;
                const cplx *xs = (const cplx *)buf;
                if (rows == D2_BK && (d0 & 1) == 0) {
                    double xv[D2_BK];
#pragma unroll
                    for (int u = 0; u < D2_BK; ++u) xv[u] = xs[u * d1 + i1].x;
#pragma unroll
                    for (int r = 0; r < D2_RB; ++r) {
                        double t = acc[r];
#pragma unroll
                        for (int u = 0; u < D2_BK; u += 2) {
                            const double2 h = *(const double2 *)(sHre + r * d0 + b0 + u);
                            t = fma(h.x, xv[u], t); t = fma(h.y, xv[u + 1], t);
                        }
                        acc[r] = t;
                    }
                } else {
                    for (int u = 0; u < rows; ++u) {
                        const double xv = xs[u * d1 + i1].x;
#pragma unroll
                        for (int r = 0; r < D2_RB; ++r) acc[r] = fma(sHre[r * d0 + b0 + u], xv, acc[r]);
                    }
                }
            });
    // Y[R0 + r][i1] += sum_b X[R0 + r][b] H1[b][i1]   (H1 real symmetric)
    d2_ring((const unsigned char *)H1re, d1, d1 * 8, d2smem, stage_bytes, full, empty, it, tid,
            [&](const unsigned char *buf, int b0, int rows) {
                if (!on) return;
                const double *hs = (const double *)buf;
                if (rows == D2_BK) {
                    double hv[D2_BK];
#pragma unroll
                    for (int u = 0; u < D2_BK; ++u) hv[u] = hs[u * d1 + i1];
#pragma unroll
                    for (int r = 0; r < D2_RB; ++r) {
                        double t = acc[r];
#pragma unroll
                        for (int u = 0; u < D2_BK; u += 2) {
                            const double2 v = *(const double2 *)(sXr + r * d1e + b0 + u);
                            t = fma(hv[u], v.x, t); t = fma(hv[u + 1], v.y, t);
                        }
                        acc[r] = t;
                    }
                } else {
                    for (int u = 0; u < rows; ++u) {
                        const double hv = hs[u * d1 + i1];
#pragma unroll
                        for (int r = 0; r < D2_RB; ++r) acc[r] = fma(hv, sXr[r * d1e + b0 + u], acc[r]);
                    }
                }
            });
    if (!on) return;
    ChebPar q;
    double s = 0.0, t = 0.0;
    if (MODE == 1) {
        q = cheb_par(a.fpar + p * 4);
        if (m == 1) s = q.sigma1 / q.e;
        else {
            double sigma = q.sigma1;
            for (int u = 2; u < m; ++u) sigma = 1.0 / (2.0 / q.sigma1 - sigma);
            const double sigma2 = 1.0 / (2.0 / q.sigma1 - sigma);
            s = 2.0 * sigma2 / q.e;
            t = sigma * sigma2;
        }
    }
    cplx *o = out + (p * a.b + j) * n;
#pragma unroll
    for (int r = 0; r < D2_RB; ++r) {
        if (r >= rb) break;
        const int i = (R0 + r) * d1 + i1;
        double h = acc[r];
        if (P.d2_nmulti > 0) h += mf_apply_terms(P, P.d2_meta, 0, P.d2_nmulti, sC, x, i).x;
        if (MODE == 0) o[i] = cmake(h, 0.0);
        else {
            const double yi = sXr[r * d1e + i1];
            if (m == 1) o[i] = cmake(s * (h - q.c * yi), 0.0);
            else o[i] = cmake(s * (h - q.c * yi) - t * o[i].x, 0.0);
        }
    }
}

// The whole degree-`deg` Chebyshev filter of one vector out of shared memory (n <= ~7000), two-node form with
// REAL combined node Hamiltonians: H0, H1 and both n-vectors live on chip; y = H0 X + X H1^T + coupling costs
// (d0 + d1) real MACs per element instead of four complex dense mode products through the ELL tables.
// grid (b, n_pts), block 512.  REALX: the vector is exactly real (imaginary half skipped, half the footprint).
template <bool REALX>
static __global__ void __launch_bounds__(512)
mf_cheb_smem_d2(MfArgs a, const cplx *__restrict__ X, cplx *__restrict__ Y, int deg, const double *__restrict__ Hre) {
    extern __shared__ __align__(16) unsigned char d2f[];
    typedef typename std::conditional<REALX, double, cplx>::type V;
    const long long p = blockIdx.y;
    if (!a.active[p]) return;
    const PlanDev &P = a.P;
    const int j = blockIdx.x, n = a.n, tid = threadIdx.x, T = blockDim.x;
    const int d0 = P.dims[0], d1 = P.dims[1], d1p = d1 | 1;             // odd pitch: conflict-free H1 rows
    // 16-byte objects first: coupling coefficients, (REALX) complex copy of the vector for the coupling terms,
    // the two vectors; then the real node Hamiltonians
    cplx *sC = (cplx *)d2f;
    cplx *sx = sC + P.d2_nmulti + 1;
    V *prev = (V *)(sx + ((REALX && P.d2_nmulti > 0) ? n : 0));
    V *cur = prev + n + (REALX ? (n & 1) : 0);                          // keep 16-byte alignment
    double *sH0 = (double *)(cur + n + (REALX ? (n & 1) : 0));          // [d0][d0]
    double *sH1 = sH0 + (size_t)d0 * d0;                                // [d1][d1p]
    const double *Hp = Hre + p * ((size_t)d0 * d0 + (size_t)d1 * d1);
    for (int q = tid; q < d0 * d0; q += T) sH0[q] = Hp[q];
    for (int q = tid; q < d1 * d1; q += T) sH1[(q / d1) * d1p + (q % d1)] = Hp[(size_t)d0 * d0 + q];
    for (int q = tid; q < P.d2_nmulti; q += T) sC[q] = a.coef[p * P.n_terms + __ldg(P.d2_cidx + q)];
    const cplx *x = X + (p * a.b + j) * n;
    for (int i = tid; i < n; i += T) {
        if constexpr (REALX) prev[i] = x[i].x; else prev[i] = x[i];
    }
    __syncthreads();
    const ChebPar q = cheb_par(a.fpar + p * 4);
    double sigma = q.sigma1;
    for (int m = 1; m <= deg; ++m) {
        // m == 1: cur = s1 (H - c) prev;   m >= 2: prev <- s (H - c) cur - t prev, then swap
        const V *src = m == 1 ? prev : cur;
        V *dst = m == 1 ? cur : prev;
        double s, t = 0.0, sigma2 = sigma;
        if (m == 1) s = q.sigma1 / q.e;
        else { sigma2 = 1.0 / (2.0 / q.sigma1 - sigma); s = 2.0 * sigma2 / q.e; t = sigma * sigma2; }
        if (REALX && P.d2_nmulti > 0) {
            for (int i = tid; i < n; i += T) sx[i] = cmake(((const double *)src)[i], 0.0);
            __syncthreads();
        }
        for (int i = tid; i < n; i += T) {
            const int i0 = i / d1, i1 = i - i0 * d1;
            double hr = 0.0, hi = 0.0;
            const double *h0 = sH0 + (size_t)i0 * d0;
            for (int b = 0; b < d0; ++b) {
                if constexpr (REALX) hr = fma(h0[b], src[b * d1 + i1], hr);
                else { const cplx v = src[b * d1 + i1]; hr = fma(h0[b], v.x, hr); hi = fma(h0[b], v.y, hi); }
            }
            const double *h1 = sH1 + (size_t)i1 * d1p;
            const V *row = src + (size_t)i0 * d1;
            for (int b = 0; b < d1; ++b) {
                if constexpr (REALX) hr = fma(h1[b], row[b], hr);
                else { const cplx v = row[b]; hr = fma(h1[b], v.x, hr); hi = fma(h1[b], v.y, hi); }
            }
            if (P.d2_nmulti > 0) {
                cplx c;
                if constexpr (REALX) c = mf_apply_terms(P, P.d2_meta, 0, P.d2_nmulti, sC, sx, i);
                else c = mf_apply_terms(P, P.d2_meta, 0, P.d2_nmulti, sC, (const cplx *)src, i);
                hr += c.x; hi += c.y;
            }
            if constexpr (REALX) {
                const double yi = src[i];
                dst[i] = m == 1 ? s * (hr - q.c * yi) : s * (hr - q.c * yi) - t * dst[i];
            } else {
                const cplx yi = src[i];
                if (m == 1) dst[i] = cmake(s * (hr - q.c * yi.x), s * (hi - q.c * yi.y));
                else { const cplx pi = dst[i]; dst[i] = cmake(s * (hr - q.c * yi.x) - t * pi.x, s * (hi - q.c * yi.y) - t * pi.y); }
            }
        }
        __syncthreads();
        if (m >= 2) { V *tmp = prev; prev = cur; cur = tmp; }
        sigma = sigma2;
    }
    cplx *y = Y + (p * a.b + j) * n;
    for (int i = tid; i < n; i += T) {
        if constexpr (REALX) y[i] = cmake(cur[i], 0.0); else y[i] = cur[i];
    }
}

}  // namespace scqed
