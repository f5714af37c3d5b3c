// assemble.cuh -- K2: H(p) = sum_t c_t(p) * kron_k F_{t,k} evaluated element-wise.
//
// Replaces pyscqed/numerical_system.py:509-594 (getHamiltonian): the reference multiplies and adds
// full-space CSR operators through Python operator overloading; here an element of H is a short
// sum of products of entries of the small per-node factors, so nothing but the factor table
// (KB-scale, cache resident) is ever read and H is written exactly once with coalesced 16-byte
// stores.  Algorithmic HBM traffic: 16 n^2 bytes per point (write only).
#pragma once
#include "common.cuh"

namespace scqed {

// One element H[i][j]; id[]/jd[] are the per-node digits of i and j, coef the point's c_t.
__device__ __forceinline__ cplx h_element(const PlanDev &P, const cplx *__restrict__ coef,
                                          const int *id, const int *jd) {
    cplx acc = cmake(0.0, 0.0);
    for (int t = 0; t < P.n_terms; ++t) {
        cplx prod = coef[t];
        bool nz = true;
        for (int k = 0; k < P.n_nodes; ++k) {
            int f = __ldg(P.term_factors + t * P.n_nodes + k);
            if (f < 0) {
                if (id[k] != jd[k]) { nz = false; break; }
            } else {
                const cplx *F = P.factor_data + __ldg(P.factor_offset + f);
                cplx e = __ldg(F + id[k] * P.dims[k] + jd[k]);
                if (e.x == 0.0 && e.y == 0.0) { nz = false; break; }
                prod = cmul(prod, e);
            }
        }
        if (nz) acc = cadd(acc, prod);
    }
    return acc;
}

__device__ __forceinline__ void digits(const PlanDev &P, int idx, int *d) {
#pragma unroll
    for (int k = 0; k < SCQED_MAX_NODES; ++k)
        if (k < P.n_nodes) d[k] = (idx / P.strides[k]) % P.dims[k];
}

#define ASM_TR 8
#define ASM_TC 128
// grid.x = n_pts * row_tiles * col_tiles ; block = 128 threads, each owns one column of the tile
// and walks ASM_TR rows, so a warp stores 512 contiguous bytes per row.
static __global__ void __launch_bounds__(ASM_TC)
assemble_dense_kernel(PlanDev P, const cplx *__restrict__ coef_all, long long n_pts, int tidyup,
                      cplx *__restrict__ H) {
    extern __shared__ cplx s_coef[];
    const int n = P.n;
    const int col_tiles = (n + ASM_TC - 1) / ASM_TC;
    const int row_tiles = (n + ASM_TR - 1) / ASM_TR;
    long long b = blockIdx.x;
    int ct = (int)(b % col_tiles); b /= col_tiles;
    int rt = (int)(b % row_tiles); b /= row_tiles;
    long long p = b;
    for (int t = threadIdx.x; t < P.n_terms; t += blockDim.x) s_coef[t] = coef_all[p * P.n_terms + t];
    __syncthreads();
    int j = ct * ASM_TC + threadIdx.x;
    if (j >= n) return;
    int jd[SCQED_MAX_NODES], id[SCQED_MAX_NODES];
    digits(P, j, jd);
    cplx *Hp = H + (size_t)p * n * n;
#pragma unroll 1
    for (int r = 0; r < ASM_TR; ++r) {
        int i = rt * ASM_TR + r;
        if (i >= n) break;
        digits(P, i, id);
        cplx v = h_element(P, s_coef, id, jd);
        if (tidyup) v = tidy(v, 1e-12);
        Hp[(size_t)i * n + j] = v;
    }
}

// ---- K2 through the entry list ---------------------------------------------------------------------
// The pattern of H is the same at every sweep point.  Sparse patterns (charge-basis circuits: 0.2-3 %
// non-zeros): the batch is zero-filled at memset speed and one thread per ENTRY writes
// sum_q c[term_q] * val_q -- every structural non-zero is written exactly once, 16 n^2 B/pt of HBM
// writes and no per-element term walk.  grid = (point groups, entry blocks).
#define ASM_PG 8      // points served by one block: the pattern tables are read once per ASM_PG points
static __global__ void __launch_bounds__(256)
assemble_entries_kernel(PlanDev P, const cplx *__restrict__ coef_all, long long n_pts, int tidyup, cplx *__restrict__ H) {
    extern __shared__ cplx s_coef[];                       // [ASM_PG][n_terms]
    const long long p0 = (long long)blockIdx.x * ASM_PG;
    const int npg = (int)(n_pts - p0 < ASM_PG ? n_pts - p0 : ASM_PG);
    const int nt = P.n_terms;
    for (int t = threadIdx.x; t < npg * nt; t += blockDim.x) s_coef[t] = coef_all[p0 * nt + t];
    __syncthreads();
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (e >= P.n_entries) return;
    const int q0 = __ldg(P.ent_ptr + e), q1 = __ldg(P.ent_ptr + e + 1);
    cplx acc[ASM_PG];
#pragma unroll
    for (int g = 0; g < ASM_PG; ++g) acc[g] = cmake(0.0, 0.0);
    for (int q = q0; q < q1; ++q) {
        const int t = __ldg(P.ent_term + q);
        const cplx v = __ldg(P.ent_val + q);
#pragma unroll
        for (int g = 0; g < ASM_PG; ++g)
            if (g < npg) cfma(acc[g], s_coef[g * nt + t], v);
    }
    const size_t nn = (size_t)P.n * P.n, key = (size_t)__ldg(P.ent_key + e);
#pragma unroll
    for (int g = 0; g < ASM_PG; ++g)
        if (g < npg) H[(size_t)(p0 + g) * nn + key] = tidyup ? tidy(acc[g], 1e-12) : acc[g];
}

// Dense patterns (oscillator displacement operators): one thread per ELEMENT, fully coalesced 16-byte
// stores, element -> entry through a 4-byte map; structural zeros are written directly.  A block serves
// ASM_PG consecutive points: the entry's (term, value) pairs are read once and applied to the ASM_PG
// coefficient rows staged in shared memory, so the table traffic per stored element drops ASM_PG-fold.
static __global__ void __launch_bounds__(256)
assemble_mapped_kernel(PlanDev P, const cplx *__restrict__ coef_all, long long n_pts, int tidyup, cplx *__restrict__ H) {
    extern __shared__ cplx s_coef[];                       // [ASM_PG][n_terms]
    const long long p0 = (long long)blockIdx.x * ASM_PG;
    const int npg = (int)(n_pts - p0 < ASM_PG ? n_pts - p0 : ASM_PG);
    const int nt = P.n_terms;
    for (int t = threadIdx.x; t < npg * nt; t += blockDim.x) s_coef[t] = coef_all[p0 * nt + t];
    __syncthreads();
    const size_t nn = (size_t)P.n * P.n;
    const size_t idx = (size_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= nn) return;
    const int e = __ldg(P.ent_map + idx);
    cplx acc[ASM_PG];
#pragma unroll
    for (int g = 0; g < ASM_PG; ++g) acc[g] = cmake(0.0, 0.0);
    if (e >= 0) {
        const int q0 = __ldg(P.ent_ptr + e), q1 = __ldg(P.ent_ptr + e + 1);
        for (int q = q0; q < q1; ++q) {
            const int t = __ldg(P.ent_term + q);
            const cplx v = __ldg(P.ent_val + q);
#pragma unroll
            for (int g = 0; g < ASM_PG; ++g)
                if (g < npg) cfma(acc[g], s_coef[g * nt + t], v);
        }
    }
#pragma unroll
    for (int g = 0; g < ASM_PG; ++g)
        if (g < npg) H[(size_t)(p0 + g) * nn + idx] = tidyup ? tidy(acc[g], 1e-12) : acc[g];
}

}  // namespace scqed
