// matfree.cu -- host side of S3: Chebyshev-filtered block subspace iteration on the Kronecker mat-vec
// (matfree.cuh, matfree_d2.cuh); the b x b Rayleigh-Ritz problems go through S1 (run_small_any).
#include "internal.h"
#include "matfree.cuh"
#include "matfree_d2.cuh"

static __global__ void __launch_bounds__(256)
fill_i32_kernel(int *__restrict__ dst, long long n, int val) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = val;
}


static int mf_block_size(int n, int k) {
    int b = k + 6;
    b = (b + 3) & ~3;
    if (b > MF_BMAX) b = MF_BMAX;
    if (b > n) b = n;
    return b;
}

// Warm start hooks of one run_matfree_core call (all optional):
//   oidx      [n_pts] where each local point goes in evals / evecs (info stays local)
//   keep_x    receives the converged Ritz blocks [n_pts][b][n]
//   init_x / init_src   start blocks: X[p] = init_x[init_src[p]]
struct MfWarm { const int *oidx; cplx *keep_x; const cplx *init_x; const int *init_src; };

static int run_matfree_core(scqed_plan *pl, const cplx *coef_all, long long n_pts, int k, int want_vec, double *evals,
                            cplx *evecs, int *info, cudaStream_t st, int deg, int maxit, double tol, const MfWarm &warm);

// Coarse-to-fine sweep: every MF_STRIDE-th point of the (flattened) grid is solved from scratch, the others
// start from the converged Ritz block of their nearest solved neighbour.  Neighbouring sweep points have
// nearly the same low-energy subspace, so the filtered subspace iteration needs 2-4 outer iterations
// instead of ~10; the accuracy criterion (residual <= tol) is unchanged.
#define MF_STRIDE 8
int run_matfree(scqed_plan *pl, const cplx *coef_all, long long n_pts, int k, int want_vec, double *evals,
                       cplx *evecs, int *info, cudaStream_t st, int deg, int maxit, double tol) {
    const PlanDev &P = pl->P;
    const MfWarm none = {nullptr, nullptr, nullptr, nullptr};
    // small sweeps: the two passes would each under-fill the GPU and pay their own iteration latencies
    if (n_pts < 16 * MF_STRIDE || n_pts > 2000000000LL || getenv("SCQED_NO_WARM_START"))
        return run_matfree_core(pl, coef_all, n_pts, k, want_vec, evals, evecs, info, st, deg, maxit, tol, none);
    const int b = mf_block_size(P.n, k);
    if (b < k) return fail(SCQED_EUNSUPPORTED, "matrix-free path supports k <= %d", MF_BMAX);
    const long long nc = (n_pts + MF_STRIDE - 1) / MF_STRIDE, nf = n_pts - nc;
    std::vector<int> perm((size_t)n_pts), src((size_t)(nf > 0 ? nf : 1));
    for (long long c = 0; c < nc; ++c) perm[(size_t)c] = (int)(c * MF_STRIDE);
    long long q = nc;
    for (long long p = 0; p < n_pts; ++p) {
        if (p % MF_STRIDE == 0) continue;
        long long lo = p / MF_STRIDE, hi = lo + 1;                       // nearest solved neighbour in grid order
        long long pick = (hi < nc && (hi * MF_STRIDE - p) < (p - lo * MF_STRIDE)) ? hi : lo;
        perm[(size_t)q] = (int)p;
        src[(size_t)(q - nc)] = (int)pick;
        ++q;
    }
    const size_t xblk = (size_t)b * P.n * 16;
    const size_t bytes = (size_t)n_pts * P.n_terms * 16 + (size_t)n_pts * 4 * 3 + (size_t)nc * xblk + 1024;
    if (pl->mfwarm.ensure(bytes)) return fail(SCQED_ENOMEM, "warm-start buffers (%zu MB)", bytes >> 20);
    unsigned char *wp = (unsigned char *)pl->mfwarm.p;
    auto take = [&](size_t nb) { unsigned char *r = wp; wp += (nb + 255) & ~(size_t)255; return r; };
    cplx *xkeep = (cplx *)take((size_t)nc * xblk);
    cplx *coefp = (cplx *)take((size_t)n_pts * P.n_terms * 16);
    int *d_perm = (int *)take((size_t)n_pts * 4), *d_src = (int *)take((size_t)(nf > 0 ? nf : 1) * 4), *infop = (int *)take((size_t)n_pts * 4);
    CK(cudaMemcpyAsync(d_perm, perm.data(), (size_t)n_pts * 4, cudaMemcpyHostToDevice, st));
    if (nf > 0) CK(cudaMemcpyAsync(d_src, src.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                                           // perm / src are stack-owned host vectors
    const long long tot = n_pts * P.n_terms;
    mf_gather_coef<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(coef_all, coefp, d_perm, n_pts, P.n_terms);
    LAUNCHED();
    MfWarm wc = {d_perm, xkeep, nullptr, nullptr};
    int rc = run_matfree_core(pl, coefp, nc, k, want_vec, evals, evecs, infop, st, deg, maxit, tol, wc);
    if (rc) return rc;
    if (nf > 0) {
        MfWarm wf = {d_perm + nc, nullptr, xkeep, d_src};
        rc = run_matfree_core(pl, coefp + (size_t)nc * P.n_terms, nf, k, want_vec, evals, evecs, infop + nc, st, deg, maxit, tol, wf);
        if (rc) return rc;
    }
    mf_scatter_info<<<(unsigned)((n_pts + 255) / 256), 256, 0, st>>>(infop, info, d_perm, n_pts);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

static int run_matfree_core(scqed_plan *pl, const cplx *coef_all, long long n_pts, int k, int want_vec, double *evals,
                            cplx *evecs, int *info, cudaStream_t st, int deg, int maxit, double tol, const MfWarm &warm) {
    const PlanDev &P = pl->P;
    const int n = P.n;
    if (P.max_term_nn > 3) return fail(SCQED_EUNSUPPORTED, "matrix-free path supports terms with at most 3 non-identity factors");
    const int b = mf_block_size(n, k);
    if (b < k) return fail(SCQED_EUNSUPPORTED, "matrix-free path supports k <= %d", MF_BMAX);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
    const size_t per_pt = (size_t)4 * b * n * 16 + (size_t)n * 8 + (size_t)2 * b * b * 16 + (size_t)2 * b * 8 + 64 + 8 +
                          (pl->d2_ok ? (size_t)(P.dims[0] * P.dims[0] + P.dims[1] * P.dims[1]) * 16 : 0);
    long long chunk = (long long)((free_b / 2) / per_pt);
    if (chunk < 1) return fail(SCQED_ENOMEM, "not enough device memory for one matrix-free block (n=%d, b=%d)", n, b);
    if (chunk > n_pts) chunk = n_pts;
    if (chunk > 16384) chunk = 16384;
    if (pl->work.ensure((size_t)chunk * per_pt + 16 * 256)) return fail(SCQED_ENOMEM, "matrix-free workspace");
    unsigned char *wp = (unsigned char *)pl->work.p;
    auto take = [&](size_t bytes) { unsigned char *r = wp; wp += (bytes + 255) & ~(size_t)255; return r; };
    cplx *bufX = (cplx *)take((size_t)chunk * b * n * 16), *bufY = (cplx *)take((size_t)chunk * b * n * 16);
    cplx *bufH = (cplx *)take((size_t)chunk * b * n * 16), *bufHn = (cplx *)take((size_t)chunk * b * n * 16);
    double *dscr = (double *)take((size_t)chunk * n * 8);
    cplx *G = (cplx *)take((size_t)chunk * b * b * 16), *S = (cplx *)take((size_t)chunk * b * b * 16);
    double *theta = (double *)take((size_t)chunk * b * 8), *rn2 = (double *)take((size_t)chunk * b * 8);
    double *fpar = (double *)take((size_t)chunk * 4 * 8);
    int *active = (int *)take((size_t)chunk * 4), *n_active = (int *)take(256), *info_s1 = (int *)take((size_t)chunk * 4);
    int *redo = (int *)take((size_t)chunk * 4);
    const bool use_cholqr = !getenv("SCQED_NO_CHOLQR");
    const bool smem_filter = (size_t)2 * n * 16 <= 220 * 1024 && !getenv("SCQED_MF_GLOBAL");     // env: force the global-memory filter (tests)
    // stencil filter (banded-monomial factors): smallest cluster whose slices fit shared memory
    int stencil_cs = 0, stencil_ns = 0;
    if (P.stencil_ok && !getenv("SCQED_NO_STENCIL")) {
        bool dims_ok = true;
        for (int q = 0; q < P.n_nodes; ++q) dims_ok &= P.dims[q] < 1024;
        for (int cs = 1; cs <= 8 && dims_ok; cs *= 2) {
            int ns = (n + cs - 1) / cs;
            if (ns <= MF_MAXE * 512 && mf_stencil_smem(P, ns) <= 220 * 1024) { stencil_cs = cs; stencil_ns = ns; break; }
        }
        if (stencil_cs) {
            CK(cudaFuncSetAttribute(mf_cheb_stencil<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
            CK(cudaFuncSetAttribute(mf_cheb_stencil<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
            CK(cudaFuncSetAttribute(mf_cheb_stencil<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
            CK(cudaFuncSetAttribute(mf_cheb_stencil<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        }
    }
    CK(cudaFuncSetAttribute(mf_cheb_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    // two-node plans with dense single-node operators on the global-memory filter: combined per-point
    // node Hamiltonians + real mode products (matfree_d2.cuh)
    const bool use_d2 = pl->d2_ok && stencil_cs == 0 && !getenv("SCQED_NO_D2");
    const int d2_tot = P.dims[0] * P.dims[0] + (P.n_nodes > 1 ? P.dims[1] * P.dims[1] : 0);
    double *d2re = nullptr, *d2im = nullptr, *d2imax = nullptr;
    size_t d2_smem_real = 0, d2_smem_cplx = 0;
    if (use_d2) {
        if (pl->d2h.ensure((size_t)chunk * d2_tot * 16 + 256)) return fail(SCQED_ENOMEM, "combined node Hamiltonians");
        d2imax = (double *)pl->d2h.p;
        d2re = (double *)((unsigned char *)pl->d2h.p + 256);
        d2im = d2re + (size_t)chunk * d2_tot;
        d2_smem_real = (size_t)D2_RB * P.dims[0] * 8 + (size_t)D2_RB * P.dims[1] * 16 + (size_t)(P.d2_nmulti + 1) * 16 +
                       (size_t)D2_RB * (P.dims[1] + 2) * 8 + 32;
        d2_smem_cplx = d2_smem_real + (size_t)D2_RB * P.dims[0] * 8;
        CK(set_max_dynamic_smem(mf_d2_kernel<0, true, true>, pl->smem_optin));
        CK(set_max_dynamic_smem(mf_d2_kernel<1, true, true>, pl->smem_optin));
        CK(set_max_dynamic_smem(mf_d2_kernel<0, true, false>, pl->smem_optin));
        CK(set_max_dynamic_smem(mf_d2_kernel<1, true, false>, pl->smem_optin));
        CK(set_max_dynamic_smem(mf_d2_kernel<0, false, false>, pl->smem_optin));
        CK(set_max_dynamic_smem(mf_d2_kernel<1, false, false>, pl->smem_optin));
    }
    bool d2_real = false, d2_cpl_real = false, d2_realx = false;
    // TMA-pipelined real kernel: 1-D bulk copies need 16-byte aligned sources and sizes (rows of X and of H1)
    size_t d2_smem_tma = 0, d2_smem_dmma = 0;
    bool d2_tma = false, d2_dmma = false;
    if (use_d2) {      // FP64 tensor-core kernel: the default for real H and real vectors
        d2_smem_dmma = (size_t)D2_RB * (P.dims[0] + 16) * 8 + (size_t)D2_RB * (P.dims[1] + 16) * 8 + (size_t)(P.d2_nmulti + 2) * 16 + 64;
        d2_dmma = P.dims[1] <= 8 * 8 * D2_NT && d2_smem_dmma <= 200 * 1024 && !getenv("SCQED_NO_D2_DMMA");
        if (d2_dmma) {
            CK(set_max_dynamic_smem(mf_d2_dmma_kernel<0>, pl->smem_optin));
            CK(set_max_dynamic_smem(mf_d2_dmma_kernel<1>, pl->smem_optin));
        }
    }
    if (use_d2) {
        const int d0_ = P.dims[0], d1_ = P.dims[1];
        const size_t stage = (((size_t)D2_BK * d1_ * 16) + 127) & ~(size_t)127;
        d2_smem_tma = D2_ST * stage + (size_t)D2_RB * d0_ * 8 + (size_t)D2_RB * ((d1_ + 1) & ~1) * 8 + (size_t)(P.d2_nmulti + 2) * 16 + 128;
        d2_tma = (d1_ % 2 == 0) && (d0_ % 2 == 0) && (n % 1 == 0) && d2_smem_tma <= 200 * 1024 &&
                 getenv("SCQED_D2_TMA") && atoi(getenv("SCQED_D2_TMA")) != 0;      // opt-in: measured 7 % slower than the plain kernel (35.6 vs 38.4 pts/s)
        if (d2_tma) {
            CK(set_max_dynamic_smem(mf_d2_tma_kernel<0>, pl->smem_optin));
            CK(set_max_dynamic_smem(mf_d2_tma_kernel<1>, pl->smem_optin));
        }
    }
    const dim3 d2_grid_base((unsigned)((P.dims[0] + D2_RB - 1) / D2_RB), (unsigned)b, 1);
    const int d2_threads = P.n_nodes > 1 ? ((P.dims[1] + 31) / 32) * 32 : 32;
    // filter degree: 80 for the stencil filter (fewer Rayleigh-Ritz rounds per mat-vec: +4-9 % on the CSFQ sweeps),
    // 60 otherwise (measured best for the fluxonium atom at 31 x 31 and 180 x 180)
    if (deg <= 0) deg = stencil_cs > 0 ? 80 : 60;
    ResonatorArgs nores;
    memset(&nores, 0, sizeof(nores));
    PlanDev Pg;                         // dummy plan for the b x b Rayleigh-Ritz solves (Hin path)
    memset(&Pg, 0, sizeof(Pg));
    const int rowblocks = (n + MF_THREADS - 1) / MF_THREADS;

    for (long long p0 = 0; p0 < n_pts; p0 += chunk) {
        const long long np = n_pts - p0 < chunk ? n_pts - p0 : chunk;
        MfArgs a;
        a.P = P; a.coef = coef_all + (size_t)p0 * P.n_terms; a.n = n; a.b = b; a.k = k; a.n_pts = np;
        a.fpar = fpar; a.active = active; a.n_active = n_active; a.theta = theta; a.rnorm2 = rn2;
        int *info_c = info + p0;
        CK(cudaMemsetAsync(rn2, 0, (size_t)np * b * 8, st));
        fill_i32_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(info_c, np, 1);     // 1 = not converged (yet); mf_converge clears it
        LAUNCHED();
        cplx *X = bufX, *Y = bufY, *H = bufH, *Hn = bufHn;
        mf_prepare<<<(unsigned)np, MF_THREADS, 0, st>>>(a, X, dscr);
        LAUNCHED();
        if (warm.init_x) {
            const size_t xt = (size_t)b * n;
            mf_init_from<<<dim3((unsigned)((xt + 255) / 256), (unsigned)np), 256, 0, st>>>(a, X, warm.init_x, warm.init_src + p0);
            LAUNCHED();
        }
        auto d2_launch = [&](int mode, const cplx *src, cplx *dst, int mstep) {
            dim3 g(d2_grid_base.x, d2_grid_base.y, (unsigned)np);
            if (d2_real && d2_realx && d2_dmma) {
                if (mode == 0) mf_d2_dmma_kernel<0><<<g, 256, d2_smem_dmma, st>>>(a, src, dst, mstep, d2re);
                else mf_d2_dmma_kernel<1><<<g, 256, d2_smem_dmma, st>>>(a, src, dst, mstep, d2re);
            } else if (d2_real && d2_realx && d2_tma) {
                if (mode == 0) mf_d2_tma_kernel<0><<<g, d2_threads, d2_smem_tma, st>>>(a, src, dst, mstep, d2re);
                else mf_d2_tma_kernel<1><<<g, d2_threads, d2_smem_tma, st>>>(a, src, dst, mstep, d2re);
            } else if (d2_real && d2_realx) {
                if (mode == 0) mf_d2_kernel<0, true, true><<<g, d2_threads, d2_smem_real, st>>>(a, src, dst, mstep, d2re, d2im);
                else mf_d2_kernel<1, true, true><<<g, d2_threads, d2_smem_real, st>>>(a, src, dst, mstep, d2re, d2im);
            } else if (d2_real) {
                if (mode == 0) mf_d2_kernel<0, true, false><<<g, d2_threads, d2_smem_real, st>>>(a, src, dst, mstep, d2re, d2im);
                else mf_d2_kernel<1, true, false><<<g, d2_threads, d2_smem_real, st>>>(a, src, dst, mstep, d2re, d2im);
            } else {
                if (mode == 0) mf_d2_kernel<0, false, false><<<g, d2_threads, d2_smem_cplx, st>>>(a, src, dst, mstep, d2re, d2im);
                else mf_d2_kernel<1, false, false><<<g, d2_threads, d2_smem_cplx, st>>>(a, src, dst, mstep, d2re, d2im);
            }
        };
        if (use_d2) {
            CK(cudaMemsetAsync(d2imax, 0, 32, st));
            mf_d2_build<<<dim3((unsigned)((d2_tot + 255) / 256), (unsigned)np), 256, 0, st>>>(a, d2re, d2im, d2imax);
            LAUNCHED();
            if (!getenv("SCQED_NO_D2_BOUND")) {
                mf_d2_bound<<<(unsigned)np, 256, 0, st>>>(a, d2re, d2im);      // after mf_prepare: tightens fpar[1]
                LAUNCHED();
            }
            double h_imax[2] = {0.0, 0.0};
            CK(cudaMemcpyAsync(h_imax, d2imax, 16, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            d2_real = h_imax[0] < 1e-12;        // the reference's tidyup threshold: such an imaginary part is dropped
            d2_cpl_real = h_imax[1] == 0.0;     // coupling terms exactly real
            d2_realx = false;                   // set per outer iteration from max |Im X|
        }
        // Rayleigh-Ritz on block R (orthonormalised in place): leaves Ritz vectors in X, H X in H
        auto rayleigh_ritz = [&](cplx *R) -> int {
            if (use_cholqr) {
                for (int round = 0; round < 2; ++round) {
                    mf_gram<<<dim3(b, (unsigned)np), MF_THREADS, 0, st>>>(a, R, R, G);
                    mf_chol<<<(unsigned)np, 32, 0, st>>>(a, G, S, redo, round == 0);
                    if (b <= 16) mf_right_apply<16><<<dim3(rowblocks, (unsigned)np), MF_THREADS, 0, st>>>(a, R, S, redo);
                    else mf_right_apply<32><<<dim3(rowblocks, (unsigned)np), MF_THREADS, 0, st>>>(a, R, S, redo);
                }
                mf_orth<<<(unsigned)np, 512, 0, st>>>(a, R, info_c, redo);        // only the flagged points do work
                g_launches.fetch_add(6, std::memory_order_relaxed);
            } else {
                mf_orth<<<(unsigned)np, 512, 0, st>>>(a, R, info_c, nullptr);
            }
            if (use_d2) d2_launch(0, R, H, 0);
            else mf_matvec<<<dim3(rowblocks, b, (unsigned)np), MF_THREADS, 0, st>>>(a, R, H);
            mf_gram<<<dim3(b, (unsigned)np), MF_THREADS, 0, st>>>(a, R, H, G);
            mf_symmetrize<<<(unsigned)np, b * b, 0, st>>>(a, G);
            g_launches.fetch_add(4, std::memory_order_relaxed);
            int rc = run_small_any(pl, Pg, nullptr, G, nullptr, np, b, b, 1, 0, nores, theta, S, nullptr, info_s1, st, false);
            if (rc) return rc;
            cplx *dst = (R == X) ? Y : X;
            mf_rotate<<<dim3(rowblocks, (unsigned)np), MF_THREADS, 0, st>>>(a, R, H, S, dst, Hn);
            cudaMemsetAsync(n_active, 0, 4, st);
            mf_converge<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(a, tol, info_c);
            g_launches.fetch_add(2, std::memory_order_relaxed);
            // after rotation the Ritz vectors live in dst; make X point at them
            if (dst != X) { Y = X; X = dst; }
            if (use_d2 && d2_real && d2_cpl_real) {       // are the Ritz vectors exactly real? (read at the next sync)
                cudaMemsetAsync(d2imax + 2, 0, 8, st);
                mf_d2_imag_max<<<(unsigned)np, 256, 0, st>>>(a, X, d2imax);
                LAUNCHED();
            }
            cplx *t = H; H = Hn; Hn = t;
            return 0;
        };
        int rc = rayleigh_ritz(X);
        if (rc) return rc;
        CK(cudaGetLastError());
        for (int it = 0; it < maxit; ++it) {
            int h_active = 0;
            double h_ximax = 1.0;
            CK(cudaMemcpyAsync(&h_active, n_active, 4, cudaMemcpyDeviceToHost, st));
            if (use_d2 && d2_real && d2_cpl_real) CK(cudaMemcpyAsync(&h_ximax, d2imax + 2, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (h_active == 0) break;
            d2_realx = use_d2 && d2_real && d2_cpl_real && h_ximax == 0.0 && !getenv("SCQED_NO_D2_REALX");
            cplx *R;
            // roofline accounting (bench.py): the whole filter of this outer iteration is one timed span; tag 5 = stencil
            // filter, 6 = shared-memory filter (generic or two-node form), 7 = global-memory filter (generic step kernel or
            // the two-node mode-product kernels); counter 1 += vector mat-vecs = active points x block size x degree
            TimingScope fscope_(st, stencil_cs > 0 ? 5 : (smem_filter ? 6 : 7));
            g_matvecs.fetch_add((long long)h_active * b * deg, std::memory_order_relaxed);
            {   // algorithmic flops of one vector mat-vec in the form used: term walk / stencil = 8 n (complex MACs per row);
                // two-node combined form = n (d0 + d1) real x complex MACs (4 flops), real x real (2 flops) for real vectors
                const double per_vec = use_d2 ? (double)n * (P.dims[0] + (P.n_nodes > 1 ? P.dims[1] : 0)) * (d2_real && d2_realx ? 2.0 : (d2_real ? 4.0 : 8.0))
                                              : 8.0 * (double)n * (double)pl->mf_work_per_row;
                g_mf_flops.fetch_add((long long)(per_vec * (double)h_active * b * deg), std::memory_order_relaxed);
            }
            if (stencil_cs > 0) {
                const size_t smem = mf_stencil_smem(P, stencil_ns);
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3((unsigned)(b * stencil_cs), (unsigned)np, 1);
                cfg.blockDim = dim3(512, 1, 1);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = stencil_cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                const cplx *Xc = X;
                if (stencil_cs == 1) CK(cudaLaunchKernelEx(&cfg, mf_cheb_stencil<1>, a, Xc, Y, deg, stencil_ns));
                else if (stencil_cs == 2) CK(cudaLaunchKernelEx(&cfg, mf_cheb_stencil<2>, a, Xc, Y, deg, stencil_ns));
                else if (stencil_cs == 4) CK(cudaLaunchKernelEx(&cfg, mf_cheb_stencil<4>, a, Xc, Y, deg, stencil_ns));
                else CK(cudaLaunchKernelEx(&cfg, mf_cheb_stencil<8>, a, Xc, Y, deg, stencil_ns));
                LAUNCHED();
                R = Y;
            } else if (smem_filter) {
                // two-node dense form with real node Hamiltonians: H0, H1 on chip next to the two vectors
                const size_t hbytes = ((size_t)P.dims[0] * P.dims[0] + (size_t)(P.n_nodes > 1 ? P.dims[1] * (P.dims[1] | 1) : 0) + 2) * 8 +
                                      (size_t)(P.d2_nmulti + 2) * 16;
                const size_t sm_rx = (size_t)2 * n * 8 + hbytes + (P.d2_nmulti > 0 ? (size_t)n * 16 : 0) + 64;
                const size_t sm_cx = (size_t)2 * n * 16 + hbytes + 64;
                if (use_d2 && d2_real && d2_realx && sm_rx <= 220 * 1024) {
                    CK(set_max_dynamic_smem(mf_cheb_smem_d2<true>, pl->smem_optin));
                    mf_cheb_smem_d2<true><<<dim3(b, (unsigned)np), 512, sm_rx, st>>>(a, X, Y, deg, d2re);
                } else if (use_d2 && d2_real && sm_cx <= 220 * 1024) {
                    CK(set_max_dynamic_smem(mf_cheb_smem_d2<false>, pl->smem_optin));
                    mf_cheb_smem_d2<false><<<dim3(b, (unsigned)np), 512, sm_cx, st>>>(a, X, Y, deg, d2re);
                } else {
                    mf_cheb_smem<<<dim3(b, (unsigned)np), 512, (size_t)2 * n * 16, st>>>(a, X, Y, deg);
                }
                LAUNCHED();
                R = Y;
            } else {
                // Y_0 = X, Y_1 -> Y, then alternate in place
                if (use_d2) d2_launch(1, X, Y, 1);
                else mf_cheb_step<<<dim3(rowblocks, b, (unsigned)np), MF_THREADS, 0, st>>>(a, X, Y, 1);
                cplx *cur = Y, *prev = X;
                for (int m = 2; m <= deg; ++m) {
                    if (use_d2) d2_launch(1, cur, prev, m);
                    else mf_cheb_step<<<dim3(rowblocks, b, (unsigned)np), MF_THREADS, 0, st>>>(a, cur, prev, m);
                    cplx *t = cur; cur = prev; prev = t;
                }
                g_launches.fetch_add(deg, std::memory_order_relaxed);
                R = cur;
            }
            fscope_.end();
            // make the filtered block the "X" buffer for the Rayleigh-Ritz bookkeeping
            if (R != X) { cplx *t = X; X = R; Y = t; }
            rc = rayleigh_ritz(X);
            if (rc) return rc;
            CK(cudaGetLastError());
        }
        unsigned ob = (unsigned)(((size_t)n * k + 255) / 256);
        if (warm.keep_x) CK(cudaMemcpyAsync(warm.keep_x + (size_t)p0 * b * n, X, (size_t)np * b * n * 16, cudaMemcpyDeviceToDevice, st));
        if (warm.oidx)
            mf_output<<<dim3(ob, (unsigned)np), 256, 0, st>>>(a, X, evals, (want_vec && evecs) ? evecs : nullptr, warm.oidx + p0);
        else
            mf_output<<<dim3(ob, (unsigned)np), 256, 0, st>>>(a, X, evals + (size_t)p0 * k,
                                                            (want_vec && evecs) ? evecs + (size_t)p0 * k * n : nullptr, nullptr);
        LAUNCHED();
        CK(cudaGetLastError());
    }
    return 0;
}


// <i|Op|j> of the auxiliary operators (Current / Voltage evaluables): one warp per (point, element)
int run_aux_matel(const PlanDev &P, const cplx *auxcoef, const cplx *evecs, long long n_pts, int k, cplx *out, cudaStream_t st) {
    const long long warps = n_pts * P.n_pairs;
    aux_matel_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(P, auxcoef, evecs, n_pts, k, out);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}
