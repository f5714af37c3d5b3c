// dense.cu -- host side of S2: batches of dense Hermitian matrices through HBM (dense_blocked.cuh: blocked
// tridiagonalisation with DMMA rank-2k updates; dense_eigh.cuh: the unblocked cross-check).
#include "internal.h"
#include "tridiag_eig.cuh"
#include "dense_eigh.cuh"
#include "dense_blocked.cuh"
#include "dense_tiled.cuh"

// S2 on a batch of B column-major matrices already resident in A.
struct DenseWs {
    cplx *vbuf, *wbuf, *part, *tau;
    double *dd, *ee, *e2, *Z, *lu;
    int nseg;
};

static size_t dense_ws_bytes(int n, int k, int B, int nseg) {
    size_t per = (size_t)4 * n + (size_t)((n + 7) / 8);
    size_t b = 0;
    b += (size_t)B * n * 16 * 3;                 // vbuf, wbuf, tau
    b += (size_t)B * nseg * n * 16;              // part
    b += (size_t)B * n * 8 * 3;                  // dd, ee, e2
    b += (size_t)B * k * n * 8;                  // Z
    b += (size_t)B * k * per * 8;                // lu
    return b + 4096;
}

static void dense_ws_carve(unsigned char *p, int n, int k, int B, int nseg, DenseWs *w) {
    size_t per = (size_t)4 * n + (size_t)((n + 7) / 8);
    auto take = [&](size_t bytes) { unsigned char *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    w->vbuf = (cplx *)take((size_t)B * n * 16);
    w->wbuf = (cplx *)take((size_t)B * n * 16);
    w->tau = (cplx *)take((size_t)B * n * 16);
    w->part = (cplx *)take((size_t)B * nseg * n * 16);
    w->dd = (double *)take((size_t)B * n * 8);
    w->ee = (double *)take((size_t)B * n * 8);
    w->e2 = (double *)take((size_t)B * n * 8);
    w->Z = (double *)take((size_t)B * k * n * 8);
    w->lu = (double *)take((size_t)B * k * per * 8);
    w->nseg = nseg;
}

static int dense_nseg(int n, int sm_count, int B) {
    int rows = (n + D2_HEMV_ROWS - 1) / D2_HEMV_ROWS;
    int want = (2 * sm_count + rows * B - 1) / (rows * B);
    int floor_ = (n + 383) / 384;          // keeps the per-CTA column segment (80 B/col of smem) under 32 KB
    if (want < floor_) want = floor_;
    if (want < 1) want = 1;
    if (const char *e = getenv("SCQED_DENSE_NSEG")) { int v = atoi(e); if (v >= 1) want = v; }
    if (want > 64) want = 64;
    return want;
}

static int run_dense_batch(scqed_plan *pl, int sm_count, cplx *A, int n, int B, int k, int want_vec, int conj_out,
                           const DenseWs &w, double *evals, cplx *evecs, int *info, cudaStream_t st) {
    (void)pl;
    for (int j = 0; j < n - 1; ++j) {
        int m = n - j - 1;
        d2_hh_gen<<<dim3(1, B), D2_THREADS, 0, st>>>(A, n, j, w.vbuf, w.dd, w.ee, w.tau);
        int rows = (m + D2_HEMV_ROWS - 1) / D2_HEMV_ROWS;
        int nseg = w.nseg;
        if (nseg > m) nseg = m;
        int seg = (m + nseg - 1) / nseg;
        d2_hemv<<<dim3(rows, nseg, B), D2_HEMV_ROWS, (size_t)seg * 16, st>>>(A, n, j, w.vbuf, w.part, nseg);
        d2_hh_w<<<dim3(1, B), D2_THREADS, 0, st>>>(n, j, w.vbuf, w.part, nseg, w.tau, w.wbuf);
        int colb = (m + D2_R2_COLS - 1) / D2_R2_COLS;
        d2_rank2<<<dim3(rows, colb, B), D2_HEMV_ROWS, 0, st>>>(A, n, j, w.vbuf, w.wbuf, w.tau, w.dd);
        g_launches.fetch_add(4, std::memory_order_relaxed);
    }
    CK(cudaGetLastError());
    tridiag_eig_kernel<<<B, 256, (size_t)k * 8, st>>>(w.dd, w.ee, n, k, want_vec, evals, w.Z, w.lu, k, w.e2, info);
    LAUNCHED();
    CK(cudaGetLastError());
    if (want_vec) {
        CK(cudaFuncSetAttribute(d2_backtransform, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if ((size_t)n * 16 > 200 * 1024) return fail(SCQED_EUNSUPPORTED, "n=%d too large for the dense back-transformation", n);
        d2_backtransform<<<dim3(k, B), D2_THREADS, (size_t)n * 16, st>>>(A, n, k, w.tau, w.Z, evecs, conj_out);
        LAUNCHED();
        CK(cudaGetLastError());
    }
    return 0;
}

// ---- S2b: blocked zlatrd + DMMA her2k -------------------------------------------------------------
struct BlkWs {
    BlkBufs K;
    double *e2, *Z, *lu;
    cplx *Ypart;
};

static size_t blk_ws_bytes(int n, int k, int B, int nseg) {
    size_t nrb = (n + BLK_ROWS - 1) / BLK_ROWS, np = (n - 1 + BLK_NB - 1) / BLK_NB;
    size_t per = (size_t)4 * n + (size_t)((n + 7) / 8);
    size_t b = 0;
    b += (size_t)B * n * BLK_NB * 16 * 3;          // Vp, Wp, G
    b += (size_t)B * n * 16 * 2;                   // acol, tau
    b += (size_t)B * nrb * 8;                      // normpart
    b += (size_t)B * nseg * n * 16;                // part
    b += (size_t)B * nrb * n * 16;                 // cpart
    b += (size_t)B * nrb * BLK_NB * 16 * 2;        // s1part, s2part
    b += (size_t)B * nrb * nseg * 8;               // vAvpart
    b += (size_t)B * np * BLK_NB * BLK_NB * 16;    // Tbuf
    b += (size_t)B * n * 8 * 3;                    // dd, ee, e2
    b += (size_t)B * k * n * 8;                    // Z
    b += (size_t)B * k * per * 8;                  // lu
    b += (size_t)B * nrb * BLK_NB * k * 16;        // Ypart
    return b + 24 * 256 + 4096;
}

static void blk_ws_carve(unsigned char *p, cplx *A, int n, int k, int B, int nseg, BlkWs *w) {
    size_t nrb = (n + BLK_ROWS - 1) / BLK_ROWS, np = (n - 1 + BLK_NB - 1) / BLK_NB;
    size_t per = (size_t)4 * n + (size_t)((n + 7) / 8);
    auto take = [&](size_t bytes) { unsigned char *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    BlkBufs &K = w->K;
    K.A = A;
    K.Vp = (cplx *)take((size_t)B * n * BLK_NB * 16);
    K.Wp = (cplx *)take((size_t)B * n * BLK_NB * 16);
    K.G = (cplx *)take((size_t)B * n * BLK_NB * 16);
    K.acol = (cplx *)take((size_t)B * n * 16);
    K.tau = (cplx *)take((size_t)B * n * 16);
    K.normpart = (double *)take((size_t)B * nrb * 8);
    K.part = (cplx *)take((size_t)B * nseg * n * 16);
    K.cpart = (cplx *)take((size_t)B * nrb * n * 16);
    K.s1part = (cplx *)take((size_t)B * nrb * BLK_NB * 16);
    K.s2part = (cplx *)take((size_t)B * nrb * BLK_NB * 16);
    K.vAvpart = (double *)take((size_t)B * nrb * nseg * 8);
    K.Tbuf = (cplx *)take((size_t)B * np * BLK_NB * BLK_NB * 16);
    K.dd = (double *)take((size_t)B * n * 8);
    K.ee = (double *)take((size_t)B * n * 8);
    w->e2 = (double *)take((size_t)B * n * 8);
    w->Z = (double *)take((size_t)B * k * n * 8);
    w->lu = (double *)take((size_t)B * k * per * 8);
    w->Ypart = (cplx *)take((size_t)B * nrb * BLK_NB * k * 16);
    K.n = n; K.nseg = nseg; K.nrb = (int)nrb; K.npanels = (int)np;
}

static int run_dense_blocked_batch(cplx *A, int n, int B, int k, int want_vec, int conj_out, BlkWs &w, double *evals,
                                   cplx *evecs, int *info, cudaStream_t st) {
    BlkBufs K = w.K;
    K.A = A;
    if (k > BLK_KMAX && want_vec) return fail(SCQED_EUNSUPPORTED, "blocked dense path supports k <= %d eigenvectors", BLK_KMAX);
    const int seg_max = (n + K.nseg - 1) / K.nseg;
    TimingScope tscope_(st, 8);            // tag 8: the whole dense -> tridiagonal reduction of this batch
    blk_w_nextcol<<<dim3(K.nrb, B), BLK_ROWS, 0, st>>>(K, -1, -1);
    LAUNCHED();
    for (int p = 0; p < K.npanels; ++p) {
        const int j0 = p * BLK_NB;
        const int nbp = (n - 1) - j0 < BLK_NB ? (n - 1) - j0 : BLK_NB;
        for (int i = 0; i < nbp; ++i) {
            blk_reflect_hemv<<<dim3(K.nrb, K.nseg, B), BLK_ROWS, (size_t)seg_max * 16 * 5, st>>>(K, j0 + i, i);
            blk_w_nextcol<<<dim3(K.nrb, B), BLK_ROWS, 0, st>>>(K, j0 + i, i);
        }
        g_launches.fetch_add(2 * nbp, std::memory_order_relaxed);
        const int j1 = j0 + nbp;
        if (p + 1 < K.npanels && j1 < n) {
            int tiles = (n - j1 + H2K_TILE - 1) / H2K_TILE;
            blk_her2k_dmma<<<dim3(tiles, tiles, B), 256, 0, st>>>(K, j1, nbp);
            LAUNCHED();
        }
    }
    CK(cudaGetLastError());
    tscope_.end();
    tridiag_eig_kernel<<<B, 256, (size_t)k * 8, st>>>(K.dd, K.ee, n, k, want_vec, evals, w.Z, w.lu, k, w.e2, info);
    LAUNCHED();
    CK(cudaGetLastError());
    if (want_vec) {
        blk_build_T<<<dim3(K.npanels, B), 32, 0, st>>>(K);
        size_t total = (size_t)B * k * n;
        blk_init_x<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w.Z, evecs, total);
        g_launches.fetch_add(2, std::memory_order_relaxed);
        for (int p = K.npanels - 1; p >= 0; --p) {
            blk_bt_dots<<<dim3(K.nrb, B), BLK_ROWS, 0, st>>>(K, p, k, evecs, w.Ypart);
            blk_bt_apply<<<dim3(K.nrb, B), BLK_ROWS, (size_t)2 * BLK_NB * k * 16, st>>>(K, p, k, evecs, w.Ypart);
        }
        g_launches.fetch_add(2 * K.npanels, std::memory_order_relaxed);
        if (conj_out) { blk_conj_x<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(evecs, total); LAUNCHED(); }
        CK(cudaGetLastError());
    }
    return 0;
}

// S2t (dense_tiled.cuh): real symmetric matrices beyond the shared-memory path, one persistent CTA per matrix on the
// tile-packed lower triangle in global memory.  Two kernels: dtg_tridiag_kernel (row-major tiles, register loads; the fastest
// below n ~ 400 where a step is latency bound) and the streamed dtc_tridiag_kernel (column-major tiles through a bulk-copy
// ring, every tile read once per column).  *done = false (and nothing written) when the batch is not real symmetric or does
// not fit: the caller then takes the complex blocked solver.  SCQED_DENSE_TILED = 0 disables the path, 1 forces the
// register kernel, 2 the streamed one; SCQED_DTC_NB / SCQED_DTC_NST override the panel width / ring depth.
// measured (reduction ms per 148 matrices, panels in shared memory vs global): n = 1000: 52 vs 58, 1200: 83.5 vs 90.9,
// 1500: 173.8 vs 166.6, 2000: 450 vs 383, 3375: 5970 (NB = 2, 32 KB ring) vs 2000 (NB = 8, one-slot row accumulators, 128 KB ring)
#define DTC_GP_MIN_N 1400
template <int NB, int YS, int CW, bool GP = false>
static int launch_dtc(scqed_plan *pl, double *tiles, int n, int B, double *dd, double *ee, double *tau, int ringb, int per_sm,
                      double *gpanels, cudaStream_t st) {
    const size_t sm = DtcGeom::fixed(n, GP ? 0 : NB, YS) + (size_t)ringb;
    CK((set_max_dynamic_smem(dtc_tridiag_kernel<NB, YS, CW, GP>, pl->smem_optin)));
    long long grid = (long long)per_sm * pl->sm_count;
    if (grid > B) grid = B;
    dtc_tridiag_kernel<NB, YS, CW, GP><<<(unsigned)grid, DTC_NTHREADS(CW), sm, st>>>(tiles, n, B, dd, ee, tau, ringb, GP ? gpanels : nullptr);
    return 0;
}

static int run_dense_tiled_batch(scqed_plan *pl, const cplx *A, int n, int B, int k, int want_vec, BlkWs &w, double *evals,
                                 cplx *evecs, int *info, cudaStream_t st, bool *done) {
    *done = false;
    int mode = -1;
    if (const char *e = getenv("SCQED_DENSE_TILED")) mode = atoi(e);
    if (mode == 0) return 0;
    const bool forced = mode > 0;
    if (n <= 224 || n > 3400 || k > 32) return 0;
    // one CTA per matrix needs a batch to fill the machine (measured against the blocked complex solver in DESIGN.md)
    if (!forced && B < 48) return 0;
    bool streamed = mode == 2 || (mode < 0 && n >= 560);        // measured crossover (tools/s2t_ab.sh): n ~ 550
    const size_t cap = pl->smem_optin > 4096 ? pl->smem_optin - 2048 : 0;
    int nb = 0, ys = 0, nst = 0, ringb = 0;
    bool two = false, gp = false;
    if (streamed) {
        // Panel width NB and ring size (nst units of 16 KB).  Measured at n = 1000 (ms per 296 matrices): NB = 8 with the 48 KB
        // ring that is left: 114-127 (latency bound: 3.7 TB/s); NB = 4 with 96 KB: 108-120 (5.2 TB/s, 33 % more traffic).  So:
        // the widest panel that leaves >= 64 KB in flight, else the widest that leaves 48 KB; the four-slot row accumulators
        // (no shuffles) before the one-slot ones.
        static const int cand[6][2] = {{8, 4}, {4, 4}, {8, 1}, {4, 1}, {2, 4}, {2, 1}};
        int fnb = 0, fys = 0, fst = 0;
        if (const char *e = getenv("SCQED_DTC_NB")) { int v = atoi(e); if (v == 8 || v == 4 || v == 2) fnb = v; }
        if (const char *e = getenv("SCQED_DTC_YS")) { int v = atoi(e); if (v == 4 || v == 1) fys = v; }
        if (const char *e = getenv("SCQED_DTC_NST")) { int v = atoi(e); if (v >= 4 && v <= 8) fst = v; }    // >= two 32 KB pieces
        for (int minst = 4; minst >= 4 && !nb; --minst)            // >= 64 KB: two pieces of DTC_PIECE tiles
            for (int c = 0; c < 6 && !nb; ++c) {
                if ((fnb && cand[c][0] != fnb) || (fys && cand[c][1] != fys)) continue;
                for (int s = fst ? fst : 6; s >= (fst ? fst : minst); --s)
                    if (DtcGeom::fixed(n, cand[c][0], cand[c][1]) + (size_t)s * 16384 <= cap) { nb = cand[c][0]; ys = cand[c][1]; nst = s; break; }
            }
        ringb = nst * 16384;
        // Panels in global memory (L2) and the whole shared memory for vectors + ring: NB = 8 whatever n.
        {
            int gmode = -1;
            if (const char *e = getenv("SCQED_DTC_GP")) gmode = atoi(e);
            const bool want_gp = gmode == 1 || (gmode < 0 && n >= DTC_GP_MIN_N && !fnb && !fys && !fst);
            if (want_gp || !nb) {
                for (int cys : {4, 1}) {
                    if (fys && cys != fys) continue;
                    const size_t fx = DtcGeom::fixed(n, 0, cys);
                    // the four-slot row accumulators only while they leave a ring of >= 96 KB (n = 3375: 38 matrices/s with
                    // four slots and 48 KB, 66 with one slot and 128 KB)
                    if (fx + (cys == 4 && !fys ? 98304 : 65536) > cap) continue;
                    gp = true; nb = 8; ys = cys;
                    size_t r = ((cap - fx) / 16384) * 16384;
                    if (r > 131072) r = 131072;
                    if (fst) r = (size_t)fst * 16384;
                    ringb = (int)r;
                    break;
                }
            }
        }
        if (!nb) {
            if (mode == 2) return 0;
            streamed = false;
        }
        // Opt-in (SCQED_DTC_TWO=1): two matrices per SM (two CTAs of 8 consumer warps, NB = 4, a ring of >= 32 KB each within
        // half of the SM's shared memory).  Measured SLOWER than one CTA of 16 consumer warps with the whole shared memory
        // (n = 600: 68 vs 58 ms per 592 matrices, 700: 122 vs 84, 300 / 400: equal): a second resident matrix does not
        // overlap anything the first one waits for -- the passes are bound by bytes in flight per SM, not by their
        // dependent sections.  Kept as a tested variant.
        if (streamed && !gp && B > pl->sm_count && getenv("SCQED_DTC_TWO") && atoi(getenv("SCQED_DTC_TWO")) == 1 && !fnb && !fys && !fst) {
            const size_t half = (size_t)113 * 1024;
            const size_t fx = DtcGeom::fixed(n, 4, 4);
            if (fx + 65536 <= half) {
                two = true; nb = 4; ys = 4;
                ringb = (int)(((half - fx) / 512) * 512);
                if (ringb > 98304) ringb = 98304;
            }
        }
    }
    if (!streamed && (n > 2048 || DtgGeom::tridiag_smem(n) + 2048 > pl->smem_optin)) return 0;
    const size_t sm_bt = DtgGeom::backtr_smem(n);
    if (want_vec && sm_bt + 1024 > cap) return 0;
    const size_t td = DtgGeom::tile_doubles(n);
    const size_t gp_bytes = gp ? (size_t)pl->sm_count * 2 * 8 * DtgGeom::ps(n) * 8 : 0;
    if (pl->dtg.ensure((size_t)B * td * 8 + 256 + gp_bytes)) return fail(SCQED_ENOMEM, "tile-packed matrices (%zu MB)", ((size_t)B * td * 8) >> 20);
    int *d_flag = (int *)pl->dtg.p;
    double *tiles = (double *)((unsigned char *)pl->dtg.p + 256);
    double *gpanels = gp ? tiles + (size_t)B * td : nullptr;
    CK(cudaMemsetAsync(d_flag, 0, 4, st));
    const long long warps = (long long)B * (long long)(DtgGeom::nt(n)) * (DtgGeom::nt(n) + 1) / 2;
    if (streamed) dtg_pack_kernel<true><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(A, n, B, tiles, d_flag);
    else dtg_pack_kernel<false><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(A, n, B, tiles, d_flag);
    LAUNCHED();
    int h_flag = 0;
    CK(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_flag) return 0;                                    // some imaginary part: complex solver
    BlkBufs &K = w.K;
    double *tau = (double *)K.tau;
    {
        TimingScope tscope_(st, 8);
        if (streamed) {
            int rc = gp ? (ys == 4 ? launch_dtc<8, 4, 16, true>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st)
                                   : launch_dtc<8, 1, 16, true>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st))
                   : two ? launch_dtc<4, 4, 8>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 2, gpanels, st)
                   : nb == 8 ? (ys == 4 ? launch_dtc<8, 4, 16>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st)
                                        : launch_dtc<8, 1, 16>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st))
                   : nb == 4 ? (ys == 4 ? launch_dtc<4, 4, 16>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st)
                                        : launch_dtc<4, 1, 16>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st))
                             : (ys == 4 ? launch_dtc<2, 4, 16>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st)
                                        : launch_dtc<2, 1, 16>(pl, tiles, n, B, K.dd, K.ee, tau, ringb, 1, gpanels, st));
            if (rc) return rc;
        } else {
            const size_t sm_tri = DtgGeom::tridiag_smem(n);
            CK(set_max_dynamic_smem(dtg_tridiag_kernel, pl->smem_optin));
            int occ = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dtg_tridiag_kernel, DTG_THREADS, sm_tri));
            if (occ < 1) return 0;
            long long grid = (long long)occ * pl->sm_count;
            if (grid > B) grid = B;
            dtg_tridiag_kernel<<<(unsigned)grid, DTG_THREADS, sm_tri, st>>>(tiles, n, B, K.dd, K.ee, tau);
        }
        tscope_.end();
    }
    LAUNCHED();
    CK(cudaGetLastError());
    const double *zw = nullptr;
    long long G = 0;
    {
        int rc = s1_tridiag_eig_batch(pl, K.dd, K.ee, n, k, B, want_vec, evals, &zw, &G, info, st);
        if (rc) return rc;
    }
    if (want_vec) {
        const long long pairs = (long long)B * k;
        const unsigned grid = (unsigned)((pairs + DTG_BT_WARPS - 1) / DTG_BT_WARPS);
        if (streamed) {
            CK(set_max_dynamic_smem(dtg_backtr_kernel<true>, pl->smem_optin));
            dtg_backtr_kernel<true><<<grid, 32 * DTG_BT_WARPS, sm_bt, st>>>(tiles, n, k, B, tau, zw, G, evecs);
        } else {
            CK(set_max_dynamic_smem(dtg_backtr_kernel<false>, pl->smem_optin));
            dtg_backtr_kernel<false><<<grid, 32 * DTG_BT_WARPS, sm_bt, st>>>(tiles, n, k, B, tau, zw, G, evecs);
        }
        LAUNCHED();
        CK(cudaGetLastError());
    }
    *done = true;
    return 0;
}

static int pick_dense_batch(int n, long long n_pts, int k, int sm_count) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
    size_t budget = free_b / 2;
    size_t per = (size_t)n * n * 16 + blk_ws_bytes(n, k, 1, 64) + dense_ws_bytes(n, k, 1, 64);
    long long B = (long long)(budget / per);
    // enough matrices in flight to fill the SMs during the skinny tail of the reduction
    // (the per-column kernels are launch / latency bound for small n: n = 256 runs 2.4 k matrices/s with 32 matrices
    // per launch and 21.8 k/s with 2048)
    // (one persistent CTA per matrix on the tile-packed path: whole waves of sm_count matrices)
    long long want = n <= 512 ? 28LL * sm_count : (n <= 1536 ? 4LL * sm_count : (n <= 4096 ? sm_count : 8));
    if (const char *e = getenv("SCQED_DENSE_BATCH")) { long long v = atoll(e); if (v >= 1) want = v; }
    if (B > want) B = want;
    if (B > n_pts) B = n_pts;
    if (B > 65535) B = 65535;
    if (B < 1) B = 0;
    return (int)B;
}


// scqed_eigh_batched, dense part: `batch` caller matrices (row-major = column-major storage of conj(H): same
// spectrum, conjugated eigenvectors) solved in batches sized from free memory.  `tmp` carries device
// properties and the workspace buffer.
int dense_eigh_resident(scqed_plan *tmp, cplx *H_dev, int n, long long batch, int k, int want_vec, double *evals,
                        cplx *evecs, int *info, int solver, cudaStream_t st) {
    long long B = pick_dense_batch(n, batch, k, tmp->sm_count);
    if (B < 1) return fail(SCQED_ENOMEM, "not enough device memory for one %d x %d matrix", n, n);
    int nseg = dense_nseg(n, tmp->sm_count, (int)B);
    const bool blocked = solver != 3;
    size_t wsb = blocked ? blk_ws_bytes(n, k, (int)B, nseg) : dense_ws_bytes(n, k, (int)B, nseg);
    if (tmp->work.ensure(wsb)) return fail(SCQED_ENOMEM, "dense workspace");
    DenseWs w;
    BlkWs bw;
    if (blocked) blk_ws_carve((unsigned char *)tmp->work.p, nullptr, n, k, (int)B, nseg, &bw);
    else dense_ws_carve((unsigned char *)tmp->work.p, n, k, (int)B, nseg, &w);
    int rc = 0;
    for (long long b0 = 0; b0 < batch && !rc; b0 += B) {
        int nb = (int)(batch - b0 < B ? batch - b0 : B);
        cplx *Ab = H_dev + (size_t)b0 * n * n;
        double *ev = evals + (size_t)b0 * k;
        cplx *vec = evecs ? evecs + (size_t)b0 * k * n : nullptr;
        bool done = false;
        if (blocked && solver != 2) rc = run_dense_tiled_batch(tmp, Ab, n, nb, k, want_vec, bw, ev, vec, info + b0, st, &done);
        if (rc || done) continue;
        rc = blocked ? run_dense_blocked_batch(Ab, n, nb, k, want_vec, 1, bw, ev, vec, info + b0, st)
                     : run_dense_batch(tmp, tmp->sm_count, Ab, n, nb, k, want_vec, 1, w, ev, vec, info + b0, st);
    }
    return rc;
}

// scqed_sweep_run, S2: assemble B matrices at a time into HBM (K2) and reduce them.
int dense_sweep(scqed_plan *pl, const cplx *coef, long long n_pts, int k, int want_vec, int tidyup, int solver, bool allow_tiled,
                double *evals, cplx *vecs, int *info, cudaStream_t st) {
    const PlanDev &P = pl->P;
    const int n = P.n;
    const bool blocked = solver != 3;
    int B = pick_dense_batch(n, n_pts, k, pl->sm_count);
    if (B < 1) return fail(SCQED_ENOMEM, "not enough device memory for one %d x %d matrix", n, n);
    int nseg = dense_nseg(n, pl->sm_count, B);
    size_t a_bytes = ((size_t)B * n * n * 16 + 255) & ~(size_t)255;
    size_t wsb = blocked ? blk_ws_bytes(n, k, B, nseg) : dense_ws_bytes(n, k, B, nseg);
    if (pl->work.ensure(a_bytes + wsb)) return fail(SCQED_ENOMEM, "dense workspace (B=%d)", B);
    cplx *A = (cplx *)pl->work.p;
    DenseWs w;
    BlkWs bw;
    if (blocked) blk_ws_carve((unsigned char *)pl->work.p + a_bytes, A, n, k, B, nseg, &bw);
    else dense_ws_carve((unsigned char *)pl->work.p + a_bytes, n, k, B, nseg, &w);
    for (long long p0 = 0; p0 < n_pts; p0 += B) {
        int nb = (int)(n_pts - p0 < B ? n_pts - p0 : B);
        // The assembly writes row-major H.  Read column-major, that buffer is H^T = conj(H): same
        // eigenvalues, conjugated eigenvectors -> conj_out = 1 restores them.
        int rc = run_assemble(pl, coef + (size_t)p0 * P.n_terms, nb, tidyup, A, st);
        if (rc) return rc;
        double *ev = evals + (size_t)p0 * k;
        cplx *vec = vecs ? vecs + (size_t)p0 * k * n : nullptr;
        bool done = false;
        if (blocked && allow_tiled) { rc = run_dense_tiled_batch(pl, A, n, nb, k, want_vec, bw, ev, vec, info + p0, st, &done); if (rc) return rc; }
        if (done) continue;
        rc = blocked ? run_dense_blocked_batch(A, n, nb, k, want_vec, 1, bw, ev, vec, info + p0, st)
                     : run_dense_batch(pl, pl->sm_count, A, n, nb, k, want_vec, 1, w, ev, vec, info + p0, st);
        if (rc) return rc;
    }
    return 0;
}
