// resonator.cuh -- K7: eigenpair post-processing for the `Resonator` evaluable.
//
// Restates pyscqed/numerical_system.py:736-784 (getResonatorResponse) on the device, fused
// behind the eigensolver so that only the strip energies leave the SM:
//   g_i   = gC * | <i|Q_c + q_b|i+1> / <0|Q_c + q_b|1> |                      (:755-762)
//   order = position of each bare level -i*wrl + E_i in the sorted bare list   (:765-770)
//   for every photon-number strip n: eigenvalues of the k x k symmetric tridiagonal
//   diag (n-i) wrl + E_i, off-diag g_i sqrt(max(n-i,0)), re-ordered by `order`  (:776-782)
// The reference makes nmax+1+k separate LAPACK calls per sweep point for this.
#pragma once
#include "common.cuh"

namespace scqed {

#define SCQED_RES_KMAX 64

struct ResonatorArgs {
    int enabled;
    int nstrips;       // nmax + 1 + k
    int op_factor;     // factor id of the coupled node's charge operator
    int op_node;       // node index of that factor
    int s_gc, s_frl, s_qb;   // scalar slots
};

// all eigenvalues of a symmetric tridiagonal (d[0..n), e[0..n-1)) by implicit QL; ascending on exit.
// Elements are `st` doubles apart: st = 1 for private arrays, st = blockDim.x for per-thread
// columns of a shared-memory tile (conflict-free, and no local-memory traffic).
__device__ inline int tql_eigenvalues(double *d, double *e, int n, int st) {
#define D_(i) d[(i) * st]
#define E_(i) e[(i) * st]
    int fail = 0;
    E_(n - 1) = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                double dd = fabs(D_(m)) + fabs(D_(m + 1));
                if (fabs(E_(m)) <= SCQED_EPS * dd) break;
            }
            if (m != l) {
                if (iter++ == 80) { fail = 1; break; }
                double g = (D_(l + 1) - D_(l)) / (2.0 * E_(l));
                double r = sqrt(fma(g, g, 1.0));
                g = D_(m) - D_(l) + E_(l) / (g + copysign(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * E_(i), b = c * E_(i);
                    r = sqrt(fma(f, f, g * g));      // |f|, |g| are O(level spacing): no overflow guard needed
                    E_(i + 1) = r;
                    if (r == 0.0) { D_(i + 1) -= p; E_(m) = 0.0; break; }
                    s = f / r; c = g / r;
                    g = D_(i + 1) - p;
                    r = (D_(i) - g) * s + 2.0 * c * b;
                    p = s * r;
                    D_(i + 1) = g + p;
                    g = c * r - b;
                }
                if (r == 0.0 && i >= l) continue;
                D_(l) -= p; E_(l) = g; E_(m) = 0.0;
            }
        } while (m != l);
    }
    for (int i = 1; i < n; ++i) {            // insertion sort
        double x = D_(i);
        int j = i - 1;
        while (j >= 0 && D_(j) > x) { D_(j + 1) = D_(j); --j; }
        D_(j + 1) = x;
    }
#undef D_
#undef E_
    return fail;
}

// Block-wide.  X: k eigenvectors [k][n]; lam: k eigenvalues; scratch: >= 8*k doubles.
// out: [nstrips][k].
static __device__ void resonator_strips(const PlanDev &P, const ResonatorArgs &R, const double *scal,
                                 const double *lam, const cplx *X, int n, int k, double *scratch,
                                 double *out, double *qlws = nullptr, int *info_p = nullptr) {
    cplx *g = (cplx *)scratch;                   // [k]   matrix elements <i|Op|i+1>
    double *glist = scratch + 2 * k;             // [k]
    double *Erel = scratch + 3 * k;              // [k]
    int *order = (int *)(scratch + 4 * k);       // [k]
    const double gC = scal[R.s_gc], wrl = scal[R.s_frl], qb = scal[R.s_qb];
    const int c = R.op_node, dc = P.dims[c], sc = P.strides[c];
    // ELL rows of the coupled node's charge operator (2 non-zeros per row in the oscillator basis,
    // 1 in the charge basis): the dense row walk was 50x more loads
    const int fw = __ldg(P.ell_w + R.op_factor);
    const long long foff = __ldg(P.ell_off + R.op_factor);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = w; i < k - 1; i += nw) {
        const cplx *xi = X + (size_t)i * n, *xj = X + (size_t)(i + 1) * n;
        cplx acc = cmake(0.0, 0.0);
        for (int a = l; a < n; a += 32) {
            int ac = (a / sc) % dc;
            int base = a - ac * sc;
            cplx y = cscale(qb, xj[a]);
            for (int b = 0; b < fw; ++b) {
                const cplx f = __ldg(P.ell_vals + foff + (long long)ac * fw + b);
                const int col = __ldg(P.ell_cols + foff + (long long)ac * fw + b);
                cfma(y, f, xj[base + col * sc]);
            }
            cfmac(acc, xi[a], y);
        }
        acc = warp_sum(acc);
        if (l == 0) g[i] = acc;
    }
    __syncthreads();
    if ((int)threadIdx.x < k) {
        int i = threadIdx.x;
        Erel[i] = lam[i] - lam[0];
        if (i < k - 1) glist[i] = gC * sqrt(cabs2(g[i]) / cabs2(g[0]));
    }
    __syncthreads();
    if ((int)threadIdx.x < k) {
        int i = threadIdx.x;
        double bi = -(double)i * wrl + Erel[i];
        int rank = 0;
        for (int j = 0; j < k; ++j) rank += ((-(double)j * wrl + Erel[j]) < bi);
        order[i] = rank;
    }
    __syncthreads();
    if (qlws) {
        // per-thread columns of a shared tile: element i of this thread at qlws[i * T + tid]
        const int T = blockDim.x;
        double *d = qlws + threadIdx.x, *e = qlws + (size_t)k * T + threadIdx.x;
        for (int s = threadIdx.x; s < R.nstrips; s += T) {
            for (int i = 0; i < k; ++i) {
                d[i * T] = (double)(s - i) * wrl + Erel[i];
                int q = s - i;
                e[i * T] = i < k - 1 ? glist[i] * sqrt((double)(q > 0 ? q : 0)) : 0.0;
            }
            if (tql_eigenvalues(d, e, k, T) && info_p) atomicOr(info_p, 4);      // QL did not converge: bad strip
            for (int i = 0; i < k; ++i) out[(size_t)s * k + i] = d[order[i] * T];
        }
    } else {
        for (int s = threadIdx.x; s < R.nstrips; s += blockDim.x) {
            double d[SCQED_RES_KMAX], e[SCQED_RES_KMAX];
            for (int i = 0; i < k; ++i) {
                d[i] = (double)(s - i) * wrl + Erel[i];
                int q = s - i;
                e[i] = i < k - 1 ? glist[i] * sqrt((double)(q > 0 ? q : 0)) : 0.0;
            }
            if (tql_eigenvalues(d, e, k, 1) && info_p) atomicOr(info_p, 4);      // QL did not converge: bad strip
            for (int i = 0; i < k; ++i) out[(size_t)s * k + i] = d[order[i]];
        }
    }
    __syncthreads();
}

}  // namespace scqed
