// common.cuh -- shared helpers for the sm_100a sweep engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace scqed {

typedef double2 cplx;   // interleaved complex128, .x = re, .y = im

#define SCQED_EPS 2.220446049250313e-16
#define SCQED_SAFMIN 2.2250738585072014e-308

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(double s, cplx a) { return make_double2(s * a.x, s * a.y); }
// acc += a * b
__device__ __forceinline__ void cfma(cplx &acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfmac(cplx &acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
// acc -= a * conj(b)
__device__ __forceinline__ void cfnmac(cplx &acc, cplx a, cplx b) {
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(-a.y, b.x, acc.y); acc.y = fma(a.x, b.y, acc.y);
}
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// The reference thresholds H at 1e-12, real and imaginary parts separately
// (pyscqed/numerical_system.py:19,594  Qobj.tidyup).
__host__ __device__ __forceinline__ cplx tidy(cplx a, double atol) {
    return make_double2(fabs(a.x) < atol ? 0.0 : a.x, fabs(a.y) < atol ? 0.0 : a.y);
}

// Device-side copy of a plan's tables (all pointers are device pointers).
#define SCQED_MAX_NODES 8
struct PlanDev {
    int n_nodes;
    int n;                         // Hilbert dimension = prod dims
    int dims[SCQED_MAX_NODES];
    int strides[SCQED_MAX_NODES];  // row-major strides of the composite index
    int n_factors;
    const long long *factor_offset;
    const cplx *factor_data;
    int n_terms;
    const int *term_factors;       // [n_terms * n_nodes]
    // matrix-free view: every factor as ELL rows (ell_w[f] entries per row, padded with value 0)
    const int *ell_w;              // [n_factors]
    const long long *ell_off;      // [n_factors] offset into ell_cols / ell_vals
    const int *ell_cols;
    const cplx *ell_vals;
    const double *fac_rowsum;      // [n_factors] max_i sum_j |F[i][j]|  (infinity norm)
    // per term: number of non-identity nodes nn (<= 3), their node ids and factor ids: [n_terms * 7]
    const int *term_meta;
    int max_term_nn;
    // stencil view (every factor has a single non-zero diagonal: diagonal / shift operators of the
    // charge basis): term t = c_t * prod_k val_fk[i_k] * x[i + delta_t].  9 ints per term:
    // nn, k0,k1,k2, tab0,tab1,tab2 (offsets into st_val), delta, original term id; the n_sdiag
    // terms with delta == 0 come first.
    int stencil_ok, st_pure_shift, n_sdiag, n_soff, st_val_total;
    const int *st_terms;
    const cplx *st_val;
    // auxiliary operators (Current / Voltage evaluables): extra terms with their own coefficients,
    // and the list of requested matrix elements <i| sum_{t in [tb, te)} c_t kron F |j>
    int n_aux, n_pairs;
    const int *aux_meta;           // [n_aux * 7] like term_meta
    const int *aux_re, *aux_im;    // [n_aux] coefficient registers
    const int *pairs;              // [n_pairs * 4]: tb, te, i, j
    // dense real term table of the fused small-n solver: Re H = sum_g alpha_g G_g, Im H = sum_g beta_g G_g
    int dense_tridiag;             // H(p) is Hermitian tridiagonal for every p
    int n_dense, n_dense_diag;     // the last n_dense_diag are diagonal (n-vectors after the full matrices)
    const double *dense_tab;       // [n_dense - n_dense_diag][n*n] column-major, then [n_dense_diag][n]
    const int *dense_re, *dense_im;   // [n_dense] coefficient registers (-1 = zero)
    const double *dense_gmax;      // [n_dense]
    // entry list of the dense assembly (K2): distinct positions of H and their (term, value) contributions
    int n_entries, ent_dense;
    const long long *ent_key;      // [n_entries] i*n + j, sorted
    const int *ent_ptr;            // [n_entries + 1]
    const int *ent_term;           // [n_pairs_total]
    const cplx *ent_val;           // [n_pairs_total]
    const int *ent_map;            // dense patterns: [n*n] element -> entry (-1 = structural zero)
    // two-node plans with dense single-node operators (matfree_d2.cuh): the coupling terms (two non-identity
    // factors) as a compact term_meta-style list + their original term ids
    int d2_nmulti;
    const int *d2_meta;            // [d2_nmulti * 7]
    const int *d2_cidx;            // [d2_nmulti]
    const int *d2_phase;           // [d2_nmulti] product of the two factors' entries: 0 real, 1 imaginary, 2 mixed
    int n_swept, n_instr, n_scalar;
    const int *instr;              // [n_instr * 3]
    const double *consts;
    const int *coef_re, *coef_im, *scalar_regs;
};

}  // namespace scqed
