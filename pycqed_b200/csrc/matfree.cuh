// matfree.cuh -- S3: matrix-free lowest-k eigensolver on the Kronecker structure
//     H(p) = sum_t c_t(p) * kron_k F_{t,k}
// Chebyshev-filtered block subspace iteration (ChebFSI): no matrix is ever formed.
//
// Replaces the reference's sparse path, which materialises every Kronecker product as an n x n
// CSR matrix (pyscqed/numerical_system.py:226-251) and calls ARPACK, usually in shift-invert
// mode through a SuperLU factorisation (pyscqed/util.py:76-125, examples/csfq-example.py:509-553).
// Shift-invert has no matrix-free analogue; a polynomial filter does.
//
// Per outer iteration and sweep point (block of b = k + guard vectors):
//   mf_cheb_smem / mf_cheb_step : Y = p_m(H) X, degree-m Chebyshev polynomial that damps
//                                 [theta_b, ub] (K3: the Kronecker mat-vec; for diagonal / shift
//                                 factors the whole filter of one vector runs out of shared
//                                 memory, two n-vectors, zero HBM traffic)
//   mf_orth                     : CGS2 orthonormalisation of Y
//   mf_matvec, mf_gram          : HX = H X,  G = X^H HX  (b x b)
//   small_eigh_kernel           : eigen-decomposition of G (the fused small-n solver)
//   mf_rotate, mf_converge      : Ritz vectors, residual norms, per-point convergence flags
//
// Algorithmic traffic of one block mat-vec in HBM/L2: 32 n b bytes (read X, write Y); flops
// 8 n b sum_t prod_k w_{t,k} with w = non-zeros per row of each factor.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace scqed {

#define MF_BMAX 32
#define MF_THREADS 256

struct MfArgs {
    PlanDev P;
    const cplx *coef;      // [n_pts][n_terms]
    int n, b, k;
    long long n_pts;
    double *fpar;          // [n_pts][4]: lo, ub, a0, spare
    int *active;           // [n_pts]
    int *n_active;         // [1]
    double *theta;         // [n_pts][b]
    double *rnorm2;        // [n_pts][b]
};

// (H x)[i] for one sweep point; x may live in shared or global memory.
__device__ __forceinline__ cplx mf_apply_terms(const PlanDev &P, const int *__restrict__ meta, int t_begin, int t_end,
                                               const cplx *__restrict__ coef, const cplx *x, int i) {
    int id[SCQED_MAX_NODES];
#pragma unroll
    for (int k = 0; k < SCQED_MAX_NODES; ++k)
        if (k < P.n_nodes) id[k] = (i / P.strides[k]) % P.dims[k];
    cplx acc = cmake(0.0, 0.0);
    for (int t = t_begin; t < t_end; ++t) {
        const int *tm = meta + t * 7;
        const int nn = __ldg(tm);
        const cplx c = coef[t];
        if (nn == 0) { cfma(acc, c, x[i]); continue; }
        const int k0 = __ldg(tm + 1), f0 = __ldg(tm + 4);
        const int w0 = __ldg(P.ell_w + f0);
        const long long o0 = __ldg(P.ell_off + f0) + (long long)id[k0] * w0;
        const int s0 = P.strides[k0], b0 = i - id[k0] * s0;
        if (nn == 1) {
            cplx sum = cmake(0.0, 0.0);
            for (int a = 0; a < w0; ++a) {
                cplx v = __ldg(P.ell_vals + o0 + a);
                int col = __ldg(P.ell_cols + o0 + a);
                cfma(sum, v, x[b0 + col * s0]);
            }
            cfma(acc, c, sum);
            continue;
        }
        const int k1 = __ldg(tm + 2), f1 = __ldg(tm + 5);
        const int w1 = __ldg(P.ell_w + f1);
        const long long o1 = __ldg(P.ell_off + f1) + (long long)id[k1] * w1;
        const int s1 = P.strides[k1], b1 = b0 - id[k1] * s1;
        if (nn == 2) {
            cplx sum = cmake(0.0, 0.0);
            for (int a = 0; a < w0; ++a) {
                cplx va = __ldg(P.ell_vals + o0 + a);
                if (va.x == 0.0 && va.y == 0.0) continue;
                int ca = __ldg(P.ell_cols + o0 + a);
                cplx inner = cmake(0.0, 0.0);
                for (int q = 0; q < w1; ++q) {
                    cplx vq = __ldg(P.ell_vals + o1 + q);
                    int cq = __ldg(P.ell_cols + o1 + q);
                    cfma(inner, vq, x[b1 + ca * s0 + cq * s1]);
                }
                cfma(sum, va, inner);
            }
            cfma(acc, c, sum);
            continue;
        }
        const int k2 = __ldg(tm + 3), f2 = __ldg(tm + 6);
        const int w2 = __ldg(P.ell_w + f2);
        const long long o2 = __ldg(P.ell_off + f2) + (long long)id[k2] * w2;
        const int s2 = P.strides[k2], b2 = b1 - id[k2] * s2;
        cplx sum = cmake(0.0, 0.0);
        for (int a = 0; a < w0; ++a) {
            cplx va = __ldg(P.ell_vals + o0 + a);
            if (va.x == 0.0 && va.y == 0.0) continue;
            int ca = __ldg(P.ell_cols + o0 + a);
            for (int q = 0; q < w1; ++q) {
                cplx vq = __ldg(P.ell_vals + o1 + q);
                if (vq.x == 0.0 && vq.y == 0.0) continue;
                int cq = __ldg(P.ell_cols + o1 + q);
                cplx vaq = cmul(va, vq);
                cplx inner = cmake(0.0, 0.0);
                for (int r = 0; r < w2; ++r) {
                    cplx vr = __ldg(P.ell_vals + o2 + r);
                    int cr = __ldg(P.ell_cols + o2 + r);
                    cfma(inner, vr, x[b2 + ca * s0 + cq * s1 + cr * s2]);
                }
                cfma(sum, vaq, inner);
            }
        }
        cfma(acc, c, sum);
    }
    return acc;
}

__device__ __forceinline__ cplx mf_apply(const PlanDev &P, const cplx *__restrict__ coef, const cplx *x, int i) {
    return mf_apply_terms(P, P.term_meta, 0, P.n_terms, coef, x, i);
}

// K7b: matrix elements of auxiliary operators between eigenvectors (Current / Voltage evaluables,
// pyscqed/numerical_system.py:596-694; util.py:322-328): out[p][q] = <x_i| Op_q |x_j> (complex).  One warp per (point, pair).
static __global__ void __launch_bounds__(128)
aux_matel_kernel(PlanDev P, const cplx *__restrict__ auxcoef, const cplx *__restrict__ evecs, long long n_pts, int k,
                 cplx *__restrict__ out) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = threadIdx.x & 31;
    if (gw >= n_pts * P.n_pairs) return;
    const long long p = gw / P.n_pairs;
    const int q = (int)(gw - p * P.n_pairs);
    const int tb = __ldg(P.pairs + 4 * q), te = __ldg(P.pairs + 4 * q + 1);
    const int i1 = __ldg(P.pairs + 4 * q + 2), i2 = __ldg(P.pairs + 4 * q + 3);
    const cplx *xi = evecs + ((size_t)p * k + i1) * P.n, *xj = evecs + ((size_t)p * k + i2) * P.n;
    const cplx *cf = auxcoef + p * P.n_aux;
    cplx acc = cmake(0.0, 0.0);
    for (int a_ = l; a_ < P.n; a_ += 32) {
        cplx y = mf_apply_terms(P, P.aux_meta, tb, te, cf, xj, a_);
        cfmac(acc, xi[a_], y);
    }
    acc = warp_sum(acc);
    if (l == 0) out[p * P.n_pairs + q] = acc;
}

// diagonal element H[i][i]
__device__ __forceinline__ double mf_diag(const PlanDev &P, const cplx *__restrict__ coef, int i) {
    int id[SCQED_MAX_NODES];
#pragma unroll
    for (int k = 0; k < SCQED_MAX_NODES; ++k)
        if (k < P.n_nodes) id[k] = (i / P.strides[k]) % P.dims[k];
    double acc = 0.0;
    for (int t = 0; t < P.n_terms; ++t) {
        cplx prod = coef[t];
        for (int k = 0; k < P.n_nodes; ++k) {
            int f = __ldg(P.term_factors + t * P.n_nodes + k);
            if (f >= 0) {
                const cplx *F = P.factor_data + __ldg(P.factor_offset + f);
                prod = cmul(prod, __ldg(F + id[k] * P.dims[k] + id[k]));
            }
        }
        acc += prod.x;
    }
    return acc;
}

__device__ __forceinline__ double mf_hash_uniform(unsigned long long s) {
    s ^= s >> 33; s *= 0xff51afd7ed558ccdULL; s ^= s >> 33; s *= 0xc4ceb9fe1a85ec53ULL; s ^= s >> 33;
    return ((double)(s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

// ---- start block: unit vectors at the b smallest diagonal entries + small noise; Gershgorin-type
// upper bound ub = sum_t |c_t| prod_k ||F_tk||_inf.   grid (n_pts), block MF_THREADS.
// dscr: [n_pts][n] scratch for the diagonal.
static __global__ void __launch_bounds__(MF_THREADS)
mf_prepare(MfArgs a, cplx *X, double *dscr) {
    __shared__ double sval[MF_THREADS];
    __shared__ int sidx[MF_THREADS];
    __shared__ int chosen[MF_BMAX];
    const long long p = blockIdx.x;
    const int n = a.n, b = a.b, tid = threadIdx.x;
    const cplx *coef = a.coef + p * a.P.n_terms;
    double *d = dscr + p * n;
    double gmax = -1e300;
    for (int i = tid; i < n; i += blockDim.x) {
        double di = mf_diag(a.P, coef, i);
        d[i] = di;
        if (a.P.stencil_ok) {                       // exact Gershgorin row bound from the stencil form
            double rs = 0.0;
            for (int t = a.P.n_sdiag; t < a.P.n_sdiag + a.P.n_soff; ++t) {
                const int *r = a.P.st_terms + t * 9;
                double w = sqrt(cabs2(coef[r[8]]));
                for (int u = 0; u < r[0]; ++u) {
                    int kk = r[1 + u];
                    w *= sqrt(cabs2(a.P.st_val[r[4 + u] + (i / a.P.strides[kk]) % a.P.dims[kk]]));
                }
                rs += w;
            }
            gmax = fmax(gmax, di + rs);
        }
    }
    sval[tid] = gmax;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) sval[tid] = fmax(sval[tid], sval[tid + s]);
        __syncthreads();
    }
    gmax = sval[0];
    __syncthreads();
    if (tid == 0) {
        double ub = 0.0;
        for (int t = 0; t < a.P.n_terms; ++t) {
            double m = sqrt(cabs2(coef[t]));
            for (int k = 0; k < a.P.n_nodes; ++k) {
                int f = a.P.term_factors[t * a.P.n_nodes + k];
                if (f >= 0) m *= a.P.fac_rowsum[f];
            }
            ub += m;
        }
        if (a.P.stencil_ok) ub = gmax + 1e-9 * fabs(gmax);
        a.fpar[p * 4 + 1] = ub;
        a.active[p] = 1;
    }
    __syncthreads();
    for (int j = 0; j < b; ++j) {          // b rounds of block-wide argmin
        double best = 1e300; int bi = -1;
        for (int i = tid; i < n; i += blockDim.x) if (d[i] < best) { best = d[i]; bi = i; }
        sval[tid] = best; sidx[tid] = bi;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (tid < s && (sval[tid + s] < sval[tid] || (sval[tid + s] == sval[tid] && sidx[tid + s] >= 0 && sidx[tid + s] < sidx[tid]))) {
                sval[tid] = sval[tid + s]; sidx[tid] = sidx[tid + s];
            }
            __syncthreads();
        }
        if (tid == 0) { chosen[j] = sidx[0]; if (sidx[0] >= 0) d[sidx[0]] = 1e300; }
        __syncthreads();
    }
    for (int j = 0; j < b; ++j) {
        cplx *x = X + (p * b + j) * n;
        const int c = chosen[j];
        for (int i = tid; i < n; i += blockDim.x) {
            double r = 1e-2 * mf_hash_uniform(((unsigned long long)(p + 1) << 40) ^ ((unsigned long long)(j + 1) << 24) ^ (unsigned long long)i);
            x[i] = cmake(r + (i == c ? 1.0 : 0.0), 0.0);
        }
    }
}

// ---- HX = H X.  grid (ceil(n/MF_THREADS), b, n_pts) --------------------------------------------------
static __global__ void __launch_bounds__(MF_THREADS)
mf_matvec(MfArgs a, const cplx *__restrict__ X, cplx *__restrict__ HX) {
    const long long p = blockIdx.z;
    if (!a.active[p]) return;
    const int j = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const cplx *x = X + (p * a.b + j) * a.n;
    HX[(p * a.b + j) * a.n + i] = mf_apply(a.P, a.coef + p * a.P.n_terms, x, i);
}

// Chebyshev scalars: e, c from (lo, ub); sigma_1 = e / (a0 - c)
struct ChebPar { double e, c, sigma1; };
__device__ __forceinline__ ChebPar cheb_par(const double *fp) {
    ChebPar q;
    q.e = 0.5 * (fp[1] - fp[0]);
    q.c = 0.5 * (fp[1] + fp[0]);
    q.sigma1 = q.e / (fp[2] - q.c);
    return q;
}

// ---- whole filter of one vector in shared memory.  grid (b, n_pts), dyn smem 2*n*16 -----------------
static __global__ void __launch_bounds__(512)
mf_cheb_smem(MfArgs a, const cplx *__restrict__ X, cplx *__restrict__ Y, int deg) {
    extern __shared__ cplx sbuf[];
    const long long p = blockIdx.y;
    if (!a.active[p]) return;
    const int j = blockIdx.x, n = a.n, tid = threadIdx.x, T = blockDim.x;
    const cplx *coef = a.coef + p * a.P.n_terms;
    cplx *prev = sbuf, *cur = sbuf + n;
    const cplx *x = X + (p * a.b + j) * n;
    for (int i = tid; i < n; i += T) prev[i] = x[i];
    __syncthreads();
    const ChebPar q = cheb_par(a.fpar + p * 4);
    double sigma = q.sigma1;
    {
        const double s = q.sigma1 / q.e;
        for (int i = tid; i < n; i += T) {
            cplx h = mf_apply(a.P, coef, prev, i);
            cplx xi = prev[i];
            cur[i] = cmake(s * (h.x - q.c * xi.x), s * (h.y - q.c * xi.y));
        }
    }
    __syncthreads();
    for (int m = 2; m <= deg; ++m) {
        const double sigma2 = 1.0 / (2.0 / q.sigma1 - sigma);
        const double s = 2.0 * sigma2 / q.e, t = sigma * sigma2;
        for (int i = tid; i < n; i += T) {
            cplx h = mf_apply(a.P, coef, cur, i);
            cplx yi = cur[i], pi = prev[i];
            prev[i] = cmake(s * (h.x - q.c * yi.x) - t * pi.x, s * (h.y - q.c * yi.y) - t * pi.y);
        }
        __syncthreads();
        cplx *tmp = prev; prev = cur; cur = tmp;
        sigma = sigma2;
    }
    cplx *y = Y + (p * a.b + j) * n;
    for (int i = tid; i < n; i += T) y[i] = cur[i];
}

// ---- stencil filter: every factor has one non-zero diagonal (charge-basis diagonal / shift operators).
// (H x)[i] = D[i] x[i] + sum_off-terms c_t prod_k val_k[i_k] x[i + delta_t], with D and all tables in
// shared memory and the digits of a thread's elements in registers.  A thread-block CLUSTER of CS CTAs
// holds one vector pair: CTA r owns elements [r ns, (r+1) ns) of both Chebyshev vectors and reads the
// neighbours it needs from its peers' shared memory (DSMEM), so the whole degree-m filter stays on chip
// for n up to CS * 5500 (CS <= 8): one cluster barrier per Chebyshev step, zero HBM traffic.
#define MF_MAXE 11          // elements per thread: ns <= MF_MAXE * 512
namespace cg = cooperative_groups;

template <int CS>
static __global__ void __launch_bounds__(512, 1)
mf_cheb_stencil(MfArgs a, const cplx *__restrict__ X, cplx *__restrict__ Y, int deg, int ns) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ cplx *peer[2][CS];
    const int n = a.n, tid = threadIdx.x, T = blockDim.x;
    const long long p = blockIdx.y;
    if (!a.active[p]) return;                       // uniform over the cluster (same p)
    unsigned rank = 0;
    if constexpr (CS > 1) rank = cg::this_cluster().block_rank();
    const int j = blockIdx.x / CS;
    const int i0 = (int)rank * ns;
    const int nloc = max(0, min(ns, n - i0));
    cplx *bufA = (cplx *)smraw, *bufB = bufA + ns;
    double *Dg = (double *)(bufB + ns);
    cplx *tab = (cplx *)(Dg + ((ns + 1) & ~1));          // keep 16-byte alignment
    cplx *cf = tab + a.P.st_val_total;
    int *rec = (int *)(cf + a.P.n_terms);
    const int nst = a.P.n_sdiag + a.P.n_soff;
    const cplx *coef = a.coef + p * a.P.n_terms;
    for (int q = tid; q < a.P.st_val_total; q += T) tab[q] = a.P.st_val[q];
    for (int q = tid; q < a.P.n_terms; q += T) cf[q] = coef[q];
    for (int q = tid; q < nst * 9; q += T) rec[q] = a.P.st_terms[q];
    const cplx *x = X + (p * a.b + j) * n;
    for (int e = tid; e < nloc; e += T) bufA[e] = x[i0 + e];
    if (tid < CS) {
        if constexpr (CS > 1) {
            peer[0][tid] = cg::this_cluster().map_shared_rank(bufA, tid);
            peer[1][tid] = cg::this_cluster().map_shared_rank(bufB, tid);
        } else { peer[0][0] = bufA; peer[1][0] = bufB; }
    }
    __syncthreads();
    // digits of my elements + the diagonal of H
    int dig[MF_MAXE];
#pragma unroll
    for (int q = 0; q < MF_MAXE; ++q) {
        int e = tid + q * T;
        dig[q] = 0;
        if (e < nloc) {
            int i = i0 + e, packed = 0;
            for (int k = 0; k < a.P.n_nodes; ++k) packed |= ((i / a.P.strides[k]) % a.P.dims[k]) << (10 * k);
            dig[q] = packed;
            double dsum = 0.0;
            for (int t = 0; t < a.P.n_sdiag; ++t) {
                const int *r = rec + t * 9;
                cplx w = cf[r[8]];
                for (int u = 0; u < r[0]; ++u) w = cmul(w, tab[r[4 + u] + ((packed >> (10 * r[1 + u])) & 1023)]);
                dsum += w.x;
            }
            Dg[e] = dsum;
        }
    }
    // pure-shift plans (every off-diagonal factor entry is 0 or 1): the weight of term t at element i
    // is c_t or 0 for the whole filter -> one validity bit per (element, term), computed once
    const bool pure = a.P.st_pure_shift && a.P.n_soff <= 32;
    unsigned msk[MF_MAXE];
    if (pure) {
#pragma unroll
        for (int q = 0; q < MF_MAXE; ++q) {
            int e = tid + q * T;
            unsigned mbits = 0u;
            if (e < nloc) {
                const int packed = dig[q];
                for (int t = a.P.n_sdiag; t < nst; ++t) {
                    const int *r = rec + t * 9;
                    bool ok = true;
                    for (int u = 0; u < r[0]; ++u) {
                        cplx v = tab[r[4 + u] + ((packed >> (10 * r[1 + u])) & 1023)];
                        ok &= (v.x != 0.0 || v.y != 0.0);
                    }
                    if (ok) mbits |= 1u << (t - a.P.n_sdiag);
                }
            }
            msk[q] = mbits;
        }
    }
    const float inv_ns = 1.0f / (float)ns;
    auto fetch = [&](int which, int i) -> cplx {
        if constexpr (CS == 1) return peer[which][0][i];
        else { int r = __float2int_rz(((float)i + 0.5f) * inv_ns); return peer[which][r][i - r * ns]; }
    };
    auto apply = [&](int which, const cplx *loc, int q) -> cplx {       // (H v)[i0 + e], v in buffer `which`
        const int e = tid + q * T, i = i0 + e, packed = dig[q];
        cplx xi = loc[e];
        cplx acc = cmake(Dg[e] * xi.x, Dg[e] * xi.y);
        if (pure) {
            unsigned mbits = msk[q];
            while (mbits) {
                const int tb = __ffs(mbits) - 1;
                mbits &= mbits - 1;
                const int *r = rec + (a.P.n_sdiag + tb) * 9;
                cfma(acc, cf[r[8]], fetch(which, i + r[7]));
            }
            return acc;
        }
        for (int t = a.P.n_sdiag; t < nst; ++t) {
            const int *r = rec + t * 9;
            cplx w = cf[r[8]];
            for (int u = 0; u < r[0]; ++u) w = cmul(w, tab[r[4 + u] + ((packed >> (10 * r[1 + u])) & 1023)]);
            if (w.x != 0.0 || w.y != 0.0) cfma(acc, w, fetch(which, i + r[7]));
        }
        return acc;
    };
    auto csync = [&]() { if constexpr (CS > 1) cg::this_cluster().sync(); else __syncthreads(); };
    csync();
    const ChebPar cq = cheb_par(a.fpar + p * 4);
    double sigma = cq.sigma1;
    cplx *prev = bufA, *cur = bufB;
    int wprev = 0, wcur = 1;
    {
        const double s = cq.sigma1 / cq.e;
#pragma unroll
        for (int q = 0; q < MF_MAXE; ++q) {
            int e = tid + q * T;
            if (e < nloc) {
                cplx h = apply(wprev, prev, q), xi = prev[e];
                cur[e] = cmake(s * (h.x - cq.c * xi.x), s * (h.y - cq.c * xi.y));
            }
        }
    }
    csync();
    for (int m = 2; m <= deg; ++m) {
        const double sigma2 = 1.0 / (2.0 / cq.sigma1 - sigma);
        const double s = 2.0 * sigma2 / cq.e, tt = sigma * sigma2;
#pragma unroll
        for (int q = 0; q < MF_MAXE; ++q) {
            int e = tid + q * T;
            if (e < nloc) {
                cplx h = apply(wcur, cur, q), yi = cur[e], pi = prev[e];
                prev[e] = cmake(s * (h.x - cq.c * yi.x) - tt * pi.x, s * (h.y - cq.c * yi.y) - tt * pi.y);
            }
        }
        csync();
        cplx *tp = prev; prev = cur; cur = tp;
        int ti = wprev; wprev = wcur; wcur = ti;
        sigma = sigma2;
    }
    cplx *y = Y + (p * a.b + j) * n;
    for (int e = tid; e < nloc; e += T) y[i0 + e] = cur[e];
    csync();                                        // peers may still be reading this CTA's buffers
}

static inline size_t mf_stencil_smem(const PlanDev &P, int ns) {
    return (size_t)ns * 40 + 16 + (size_t)P.st_val_total * 16 + (size_t)P.n_terms * 16 + (size_t)(P.n_sdiag + P.n_soff) * 36 + 64;
}

// ---- one Chebyshev step through global memory (large n).  grid (ceil(n/T), b, n_pts).
// m == 1: out = s1 (H x - c x);  m >= 2: prev <- s (H cur - c cur) - t prev   (in place in prev)
static __global__ void __launch_bounds__(MF_THREADS)
mf_cheb_step(MfArgs a, const cplx *__restrict__ cur, cplx *prev, int m) {
    const long long p = blockIdx.z;
    if (!a.active[p]) return;
    const int j = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x, n = a.n;
    if (i >= n) return;
    const ChebPar q = cheb_par(a.fpar + p * 4);
    const cplx *y = cur + (p * a.b + j) * n;
    cplx h = mf_apply(a.P, a.coef + p * a.P.n_terms, y, i);
    cplx yi = y[i];
    cplx *o = prev + (p * a.b + j) * n;
    if (m == 1) {
        const double s = q.sigma1 / q.e;
        o[i] = cmake(s * (h.x - q.c * yi.x), s * (h.y - q.c * yi.y));
    } else {
        double sigma = q.sigma1;
        for (int u = 2; u < m; ++u) sigma = 1.0 / (2.0 / q.sigma1 - sigma);
        const double sigma2 = 1.0 / (2.0 / q.sigma1 - sigma);
        const double s = 2.0 * sigma2 / q.e, t = sigma * sigma2;
        cplx pi = o[i];
        o[i] = cmake(s * (h.x - q.c * yi.x) - t * pi.x, s * (h.y - q.c * yi.y) - t * pi.y);
    }
}

__device__ __forceinline__ double mf_block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = l < nw ? red[l] : 0.0;
    return warp_sum(s);
}

// ---- CGS2 orthonormalisation of the block (in place).  grid (n_pts), block 512 -----------------------
static __global__ void __launch_bounds__(512)
mf_orth(MfArgs a, cplx *Y, int *info, const int *__restrict__ only) {
    __shared__ double red[32];
    __shared__ double sdot[MF_BMAX][2];
    const long long p = blockIdx.x;
    if (!a.active[p]) return;
    if (only && !only[p]) return;
    const int n = a.n, b = a.b, tid = threadIdx.x, T = blockDim.x;
    const int w = tid >> 5, l = tid & 31, nw = T >> 5;
    __shared__ double swarp[16][MF_BMAX][2];
    for (int j = 0; j < b; ++j) {
        cplx *y = Y + (p * b + j) * n;
        double nrm0 = 0.0;
        for (int pass = 0; pass < 2 && j > 0; ++pass) {
            // all dots <y_q, y_j>, q < j, in one sweep (register accumulators, chunks of 8)
            for (int q0 = 0; q0 < j; q0 += 8) {
                cplx acc[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = cmake(0.0, 0.0);
                for (int i = tid; i < n; i += T) {
                    cplx yi = y[i];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (q0 + u < j) cfmac(acc[u], Y[(p * b + q0 + u) * n + i], yi);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    cplx s = warp_sum(acc[u]);
                    if (l == 0 && q0 + u < j) { swarp[w][q0 + u][0] = s.x; swarp[w][q0 + u][1] = s.y; }
                }
            }
            __syncthreads();
            if (tid < j) {
                double sx = 0.0, sy = 0.0;
                for (int ww = 0; ww < nw; ++ww) { sx += swarp[ww][tid][0]; sy += swarp[ww][tid][1]; }
                sdot[tid][0] = sx; sdot[tid][1] = sy;
            }
            __syncthreads();
            for (int i = tid; i < n; i += T) {
                cplx yi = y[i];
                for (int q = 0; q < j; ++q) {
                    cplx yq = Y[(p * b + q) * n + i];
                    cplx dq = cmake(sdot[q][0], sdot[q][1]);
                    yi.x -= yq.x * dq.x - yq.y * dq.y;
                    yi.y -= yq.x * dq.y + yq.y * dq.x;
                }
                y[i] = yi;
            }
            __syncthreads();
        }
        double s = 0.0;
        for (int i = tid; i < n; i += T) s += cabs2(y[i]);
        s = mf_block_sum(s, red);
        (void)nrm0;
        if (!(s > 1e-280)) {                      // rank deficient: should not happen; flag it
            if (tid == 0) info[p] = 2;
            s = 1.0;
        }
        const double sc = 1.0 / sqrt(s);
        for (int i = tid; i < n; i += T) { cplx yi = y[i]; y[i] = cmake(yi.x * sc, yi.y * sc); }
        __syncthreads();
    }
}

// ---- CholQR: Y <- Y T with T = S R^-1, where S = diag(M)^-1/2 and R^H R = S M S is the Cholesky factor of
// the column-scaled Gram matrix M = Y^H Y.  Two rounds (CholQR2) orthonormalise a filtered block to machine
// precision with 2 Gram sweeps + 2 row-parallel updates instead of the b sequential CGS2 steps of mf_orth
// (one CTA per point, 22 % of the n = 9261 solve).  A point whose scaled Gram matrix is numerically
// singular (pivot^2 < 1e-10) is flagged in `redo` and orthonormalised by mf_orth afterwards.
// One warp per point; lane = column.
static __global__ void __launch_bounds__(32)
mf_chol(MfArgs a, const cplx *__restrict__ M, cplx *__restrict__ Tout, int *__restrict__ redo, int first_round) {
    __shared__ cplx R[MF_BMAX][MF_BMAX + 1];
    __shared__ double sc[MF_BMAX];
    const long long p = blockIdx.x;
    const int b = a.b, l = threadIdx.x;
    if (first_round && l == 0) redo[p] = 0;
    if (!a.active[p]) return;
    if (!first_round && redo[p]) return;
    bool bad = false;
    if (l < b) {
        const double dg = M[(p * b + l) * b + l].x;
        bad = !(dg > 1e-280) || !isfinite(dg);
        sc[l] = bad ? 1.0 : 1.0 / sqrt(dg);
    }
    __syncwarp();
    for (int r = 0; r < b; ++r)
        if (l < b) { cplx m = M[(p * b + r) * b + l]; const double f = sc[r] * sc[l]; R[r][l] = cmake(m.x * f, m.y * f); }
    __syncwarp();
    // upper Cholesky, column l owned by lane l: R[j][l] for j <= l
    for (int j = 0; j < b; ++j) {
        if (l >= j && l < b) {
            cplx v = R[j][l];
            for (int q = 0; q < j; ++q) { cplx rq = R[q][j], rl = R[q][l]; v.x -= rq.x * rl.x + rq.y * rl.y; v.y -= rq.x * rl.y - rq.y * rl.x; }   // conj(R[q][j]) R[q][l]
            R[j][l] = v;
        }
        __syncwarp();
        const double piv = R[j][j].x;
        if (!(piv > 1e-10)) bad = true;
        const double inv = 1.0 / sqrt(piv > 1e-10 ? piv : 1.0);
        __syncwarp();
        if (l >= j && l < b) { cplx v = R[j][l]; R[j][l] = l == j ? cmake(sqrt(piv > 1e-10 ? piv : 1.0), 0.0) : cmake(v.x * inv, v.y * inv); }
        __syncwarp();
    }
    if (__any_sync(0xffffffffu, bad)) { if (l == 0) redo[p] = 1; return; }
    // X = R^-1 (upper), column l by back substitution; T = S X
    if (l < b) {
        cplx x[MF_BMAX];
#pragma unroll
        for (int i = 0; i < MF_BMAX; ++i) x[i] = cmake(0.0, 0.0);
#pragma unroll
        for (int i = MF_BMAX - 1; i >= 0; --i) {
            if (i <= l && i < b) {
                cplx v = i == l ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
#pragma unroll
                for (int q = i + 1; q < MF_BMAX; ++q)
                    if (q <= l) { cplx r = R[i][q]; v.x -= r.x * x[q].x - r.y * x[q].y; v.y -= r.x * x[q].y + r.y * x[q].x; }
                const double d = 1.0 / R[i][i].x;
                x[i] = cmake(v.x * d, v.y * d);
            }
        }
#pragma unroll
        for (int i = 0; i < MF_BMAX; ++i)
            if (i < b) Tout[(p * b + i) * b + l] = i <= l ? cmake(x[i].x * sc[i], x[i].y * sc[i]) : cmake(0.0, 0.0);
    }
}

// Y[:, j] <- sum_{q <= j} Y[:, q] T[q][j], in place; grid (ceil(n / 256), n_pts), one thread per row.
template <int BB>
static __global__ void __launch_bounds__(MF_THREADS)
mf_right_apply(MfArgs a, cplx *__restrict__ Y, const cplx *__restrict__ Tm, const int *__restrict__ redo) {
    __shared__ cplx sT[BB * BB];
    const long long p = blockIdx.y;
    if (!a.active[p] || redo[p]) return;
    const int n = a.n, b = a.b, tid = threadIdx.x;
    for (int q = tid; q < b * b; q += blockDim.x) sT[q] = Tm[p * b * b + q];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + tid;
    if (i >= n) return;
    cplx y[BB];
#pragma unroll
    for (int q = 0; q < BB; ++q) y[q] = q < b ? Y[(p * b + q) * n + i] : cmake(0.0, 0.0);
#pragma unroll
    for (int j = BB - 1; j >= 0; --j) {
        if (j < b) {
            cplx acc = cmake(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < BB; ++q)
                if (q <= j) cfma(acc, y[q], sT[q * b + j]);
            Y[(p * b + j) * n + i] = acc;
        }
    }
}

// ---- G[a_][c] = <x_a, hx_c>.  grid (b, n_pts), block 256.  G row-major [n_pts][b][b] -----------------
static __global__ void __launch_bounds__(MF_THREADS)
mf_gram(MfArgs a, const cplx *__restrict__ X, const cplx *__restrict__ HX, cplx *__restrict__ G) {
    __shared__ double swarp[MF_THREADS / 32][MF_BMAX][2];
    const long long p = blockIdx.y;
    const int ia = blockIdx.x, n = a.n, b = a.b, tid = threadIdx.x, T = blockDim.x;
    const int w = tid >> 5, l = tid & 31, nw = T >> 5;
    if (!a.active[p]) return;
    const cplx *xa = X + (p * b + ia) * n;
    for (int c0 = 0; c0 < b; c0 += 8) {
        cplx acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = cmake(0.0, 0.0);
        for (int i = tid; i < n; i += T) {
            cplx xi = xa[i];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < b) cfmac(acc[u], xi, HX[(p * b + c0 + u) * n + i]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            cplx s = warp_sum(acc[u]);
            if (l == 0 && c0 + u < b) { swarp[w][c0 + u][0] = s.x; swarp[w][c0 + u][1] = s.y; }
        }
    }
    __syncthreads();
    if (tid < b) {
        double sx = 0.0, sy = 0.0;
        for (int ww = 0; ww < nw; ++ww) { sx += swarp[ww][tid][0]; sy += swarp[ww][tid][1]; }
        G[(p * b + ia) * b + tid] = cmake(sx, sy);
    }
}

// symmetrise G in place (Hermitian part): grid (n_pts), block b*b threads (<= 1024)
static __global__ void mf_symmetrize(MfArgs a, cplx *G) {
    const long long p = blockIdx.x;
    const int b = a.b, r = threadIdx.x / b, c = threadIdx.x % b;
    if (r >= b) return;
    cplx x = G[(p * b + r) * b + c], y = G[(p * b + c) * b + r];
    __syncthreads();
    G[(p * b + r) * b + c] = cmake(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
}

// ---- Xn = X S, HXn = HX S; residual norms.  S[a_][j] = evec[j][a_].  grid (ceil(n/T), n_pts) --------
static __global__ void __launch_bounds__(MF_THREADS)
mf_rotate(MfArgs a, const cplx *__restrict__ X, const cplx *__restrict__ HX, const cplx *__restrict__ S,
          cplx *__restrict__ Xn, cplx *__restrict__ HXn) {
    __shared__ cplx sS[MF_BMAX * MF_BMAX];
    __shared__ double swarp[MF_THREADS / 32][MF_BMAX];
    const long long p = blockIdx.y;
    if (!a.active[p]) return;
    const int n = a.n, b = a.b, tid = threadIdx.x;
    const int i = blockIdx.x * blockDim.x + tid;
    const int w = tid >> 5, l = tid & 31, nw = blockDim.x >> 5;
    for (int q = tid; q < b * b; q += blockDim.x) sS[q] = S[p * b * b + q];     // [j][a]
    __syncthreads();
    const double *th = a.theta + p * b;
    for (int j0 = 0; j0 < b; j0 += 8) {
        cplx ax[8], ah[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { ax[u] = cmake(0.0, 0.0); ah[u] = cmake(0.0, 0.0); }
        if (i < n) {
            for (int q = 0; q < b; ++q) {
                cplx xq = X[(p * b + q) * n + i], hq = HX[(p * b + q) * n + i];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (j0 + u < b) { cplx s = sS[(j0 + u) * b + q]; cfma(ax[u], xq, s); cfma(ah[u], hq, s); }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double r2 = 0.0;
            if (j0 + u < b && i < n) {
                Xn[(p * b + j0 + u) * n + i] = ax[u];
                HXn[(p * b + j0 + u) * n + i] = ah[u];
                double t = th[j0 + u];
                double rx = ah[u].x - t * ax[u].x, ry = ah[u].y - t * ax[u].y;
                r2 = rx * rx + ry * ry;
            }
            r2 = warp_sum(r2);
            if (l == 0 && j0 + u < b) swarp[w][j0 + u] = r2;
        }
    }
    __syncthreads();
    if (tid < b) {
        double s = 0.0;
        for (int ww = 0; ww < nw; ++ww) s += swarp[ww][tid];
        atomicAdd(a.rnorm2 + p * b + tid, s);
    }
}

// ---- convergence + next filter window.  grid (ceil(n_pts/128)), block 128 ---------------------------
static __global__ void mf_converge(MfArgs a, double tol, int *info) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_pts || !a.active[p]) return;
    const int b = a.b, k = a.k;
    const double *th = a.theta + p * b;
    double *r2 = a.rnorm2 + p * b;
    bool conv = true;
    for (int j = 0; j < k; ++j) conv &= (sqrt(r2[j]) <= tol * fmax(1.0, fabs(th[j])));
    for (int j = 0; j < b; ++j) r2[j] = 0.0;
    double lo = th[b - 1], a0 = th[0], ub = a.fpar[p * 4 + 1];
    if (!(lo > a0)) lo = a0 + 1e-6 * (ub - a0);
    if (!(ub > lo)) ub = lo + 1.0;
    a.fpar[p * 4 + 0] = lo;
    a.fpar[p * 4 + 1] = ub;
    a.fpar[p * 4 + 2] = a0;
    if (conv) { a.active[p] = 0; info[p] = 0; }
    else atomicAdd(a.n_active, 1);
}

// ---- outputs: first k Ritz pairs.  grid (ceil(n*k/256), n_pts) -----------------------------------------
// `oidx` (optional): position of local point p in the caller's arrays (coarse-to-fine ordering, see run_matfree)
static __global__ void mf_output(MfArgs a, const cplx *__restrict__ X, double *__restrict__ evals, cplx *__restrict__ evecs,
                          const int *__restrict__ oidx) {
    const long long p = blockIdx.y;
    const long long po = oidx ? oidx[p] : p;
    const int q = blockIdx.x * blockDim.x + threadIdx.x, n = a.n, k = a.k, b = a.b;
    if (q < k) evals[po * k + q] = a.theta[p * b + q];
    if (evecs && q < n * k) {
        int j = q / n, i = q - j * n;
        evecs[(po * k + j) * n + i] = X[(p * b + j) * n + i];
    }
}

// coarse-to-fine start: X[p] = Xsrc[src[p]] (the converged Ritz block of the nearest already-solved sweep point)
static __global__ void __launch_bounds__(256)
mf_init_from(MfArgs a, cplx *__restrict__ X, const cplx *__restrict__ Xsrc, const int *__restrict__ src) {
    const long long p = blockIdx.y;
    const size_t tot = (size_t)a.b * a.n;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < tot) X[p * tot + i] = Xsrc[(size_t)src[p] * tot + i];
}

static __global__ void mf_gather_coef(const cplx *__restrict__ in, cplx *__restrict__ out, const int *__restrict__ perm, long long n_pts, int n_terms) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pts * n_terms) return;
    const long long p = i / n_terms;
    out[i] = in[(long long)perm[p] * n_terms + (i - p * n_terms)];
}

static __global__ void mf_scatter_info(const int *__restrict__ in, int *__restrict__ out, const int *__restrict__ perm, long long n_pts) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pts) out[perm[p]] = in[p];
}

}  // namespace scqed
