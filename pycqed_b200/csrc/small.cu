// small.cu -- host side of S1: Hamiltonians that fit one SM's shared memory (small_eigh.cuh, small_split.cuh,
// small_packed.cuh).  See internal.h for the entry points.
#include "internal.h"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"
#include "small_packed.cuh"
#include "small_rows.cuh"
#include "small_reg.cuh"
#include "small_blocked.cuh"
#include "small_ladder.cuh"
#include "small_tiled.cuh"

static bool small_fits(const scqed_plan *pl, int n, int k, int want_vec, int n_terms, SmallLayout *Lout) {
    SmallLayout L = small_smem_layout(n, k, want_vec, n_terms, pl->smem_optin);
    if ((size_t)L.total + 1024 > pl->smem_optin) return false;
    if (want_vec && L.vb < 1) return false;
    if (Lout) *Lout = L;
    return true;
}

bool small_fits(const scqed_plan *pl, int n, int k, int want_vec, int n_terms) {
    return small_fits(pl, n, k, want_vec, n_terms, nullptr);
}

static int run_small(scqed_plan *pl, const PlanDev &P, const cplx *coef, const cplx *Hin, const double *scal,
                     long long n_pts, int n, int k, int want_vec, int tidyup, const ResonatorArgs &res,
                     double *evals, cplx *evecs, double *post, int *info, cudaStream_t st) {
    SmallArgs a;
    memset(&a, 0, sizeof(a));
    if (!small_fits(pl, n, k, want_vec, P.n_terms, &a.L)) return fail(SCQED_EUNSUPPORTED, "n=%d k=%d does not fit the fused shared-memory path", n, k);
    a.P = P; a.coef = coef; a.Hin = Hin; a.scalars = scal; a.n_pts = n_pts; a.n = n; a.k = k;
    a.want_vec = want_vec; a.tidyup = tidyup; a.evals = evals; a.evecs = evecs; a.info = info;
    a.res = res; a.post = post;
    // per DEVICE, not per process: set on every launch (microseconds) so that several GPUs driven by one process work
    CK(cudaFuncSetAttribute(small_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_optin - 1024));
    long long grid = n_pts < pl->sm_count ? n_pts : pl->sm_count;
    small_eigh_kernel<<<(unsigned)grid, SMALL_THREADS, a.L.total, st>>>(a);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// ---- S1 as a phase pipeline (small_split.cuh) ------------------------------------------------------------
// SCQED_S1_PACKED: 0 = full-storage kernel only (n <= 128), 1 = packed lower triangle for every n,
// unset = auto: full storage for n <= 128 (faster per matrix: 6.6 vs 9.9 ms on the bench), packed above
// (the only way 128 < n <= ~212 real symmetric stays on chip: 8x the HBM solver at n = 200)
// Every launch below sets MaxDynamicSharedMemorySize to the device's opt-in maximum, never to the size of the launch at
// hand: the attribute is per function and device, and several host threads (one per engine, EnginePool) launch the same
// kernel with different footprints -- a per-launch value set by one thread could undercut another thread's launch.
static int s1_packed_mode() {
    const char *e = getenv("SCQED_S1_PACKED");
    return e ? atoi(e) : 2;
}

// real_hint: -1 unknown, 1 known real, 0 known complex
bool split_fits(const scqed_plan *pl, int n, int k, int n_terms, int real_hint) {
    if (k > 32 || n < 2) return false;
    if (n <= 32 * SMALL_RQ)
        return S1Smem<cplx>::bytes(n, n_terms, 512) <= pl->smem_optin ||
               (s1_packed_mode() && S1PSmem<cplx>::bytes(n, n_terms, 256) <= pl->smem_optin);
    if (n > 224 || real_hint == 0 || !s1_packed_mode()) return false;
    // real symmetric only: tile-packed (small_tiled.cuh, up to n = 224) or column-packed lower triangle
    const bool tiled_off = getenv("SCQED_S1_TILED") && atoi(getenv("SCQED_S1_TILED")) == 0;
    if (!tiled_off && S1TSmem::scratch_fits(n, n_terms) && S1TSmem::bytes(n) + 1024 <= pl->smem_optin) return true;
    return S1PSmem<double>::bytes(n, n_terms, 512) <= pl->smem_optin;
}

template <typename T, int NQ>
static int s1p_launch(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    const int threads = NQ <= 4 ? 256 : 512;
    const size_t smem = S1PSmem<T>::bytes(a.n, n_terms, threads);
    *launched = false;
    if (smem > pl->smem_optin) return 0;
    CK(set_max_dynamic_smem(s1p_tridiag_kernel<T, NQ>, pl->smem_optin));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1p_tridiag_kernel<T, NQ>, threads, smem));
    if (occ < 1) return 0;
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    TimingScope tscope_(st, sizeof(T) == sizeof(double) ? 1 : 2);
    s1p_tridiag_kernel<T, NQ><<<(unsigned)grid, threads, smem, st>>>(a);
    tscope_.end();
    LAUNCHED();
    CK(cudaGetLastError());
    *launched = true;
    return 0;
}

// SCQED_S1_ROWS=1: use the row-per-thread packed kernel (small_rows.cuh) for real symmetric matrices whenever it
// fits.  Opt-in: measured SLOWER than the kernels it would replace (n = 101: 8.43 vs 6.61 ms per 10^4 points,
// n = 200: 77.9 vs 64.9 ms) -- its shared-memory wavefront count came out the same (1.34e9 vs 1.13e9, ncu
// profiles/r02_s1r_rows_ncu_summary.csv): on a packed triangle half of the lanes of the diagonal blocks are
// predicated off and every warp re-loads the per-column broadcast operands.  Kept as a parity-tested alternate.
static int s1_rows_mode() {
    const char *e = getenv("SCQED_S1_ROWS");
    return e ? atoi(e) : 0;
}

static int s1r_launch(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    *launched = false;
    if (a.n > 256 || a.n < 2) return 0;
    const int threads = a.n <= 128 ? 128 : 256;
    const size_t smem = S1RSmem::bytes(a.n, n_terms);
    if (smem > pl->smem_optin) return 0;
    CK(set_max_dynamic_smem(s1r_tridiag_kernel, pl->smem_optin));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1r_tridiag_kernel, threads, smem));
    if (occ < 1) return 0;
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    TimingScope tscope_(st, 1);
    s1r_tridiag_kernel<<<(unsigned)grid, threads, smem, st>>>(a);
    tscope_.end();
    LAUNCHED();
    CK(cudaGetLastError());
    *launched = true;
    return 0;
}

// Blocked tridiagonalisation with DMMA trailing updates (small_blocked.cuh) for real symmetric matrices whose full
// storage fits shared memory (n <= 128): SCQED_S1_BLK=1.
static int s1b_launch(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    *launched = false;
    // opt-in: correct and parity-tested, but not faster than the unblocked kernel on the bench workload (n = 101: 7.2 ms
    // with the O(m) sections on four warps, 9.1 ms on one warp, against 6.6 ms): every variant is bound by the dependent
    // chain of a Householder step (~4000 clk per column and CTA at 2 CTAs per SM), not by the shared-memory traffic the
    // blocked form removes (profiles/r02_s1b_blocked_ncu_summary.csv)
    const char *e = getenv("SCQED_S1_BLK");
    if (!e || atoi(e) == 0) return 0;
    if (a.n < 3 || a.n > 128) return 0;
    const size_t smem = S1BSmem::bytes(a.n, n_terms);
    if (smem > pl->smem_optin) return 0;
    CK(set_max_dynamic_smem(s1b_tridiag_kernel, pl->smem_optin));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1b_tridiag_kernel, S1B_THREADS, smem));
    if (occ < 1) return 0;
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    TimingScope tscope_(st, 1);
    s1b_tridiag_kernel<<<(unsigned)grid, S1B_THREADS, smem, st>>>(a, 1 << 30, nullptr);
    tscope_.end();
    LAUNCHED();
    CK(cudaGetLastError());
    *launched = true;
    return 0;
}

// Tile-packed blocked tridiagonalisation (small_tiled.cuh): four CTAs per SM at n = 101.
static int s1t_launch(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    *launched = false;
    // default for 80 <= n <= 128 (measured against the ladder / unblocked kernels, tools/s1_scaling.py, 4096 matrices:
    // n = 88: 1.78 vs 1.89 ms, 101: 2.34 vs 2.69, 118: 3.82 vs 5.29, 128: 5.83 vs 6.21; below 80 the others win) and for
    // 128 < n <= 224 with 256 threads and one CTA per SM (fluxonium at truncation 200, 10^4 points: 64.9 -> 39.3 ms);
    // SCQED_S1_TILED=1 forces it for every n >= 16, =0 disables it
    const char *e = getenv("SCQED_S1_TILED");
    const int mode = e ? atoi(e) : -1;
    if (mode == 0 || (mode < 0 && (a.n < 80 || getenv("SCQED_S1_PACKED") || s1_rows_mode()))) return 0;   // explicit storage opt-ins win
    if (a.n < 16 || a.n > 224 || !S1TSmem::scratch_fits(a.n, n_terms)) return 0;
    if (a.n > 128 && getenv("SCQED_S1_TILED_BIG") && atoi(getenv("SCQED_S1_TILED_BIG")) == 0) return 0;   // A/B switch: packed kernel instead
    const size_t smem = S1TSmem::bytes(a.n);
    if (smem + 1024 > pl->smem_optin) return 0;
    const bool big = a.n > 128;                            // one row per thread in the O(m) sections: 256 threads
    const int threads = big ? 256 : S1T_THREADS;
    const bool pser = !(getenv("SCQED_S1_TILED_PSER") && atoi(getenv("SCQED_S1_TILED_PSER")) == 0);
    if (big && !pser) return 0;                            // the one-warp O(m) sections cover 128 rows
    void (*kern)(S1Args, int *, int) = big ? (pser ? s1t_tridiag_kernel<true, 256> : s1t_tridiag_kernel<false, 256>)
                                           : (pser ? s1t_tridiag_kernel<true, S1T_THREADS> : s1t_tridiag_kernel<false, S1T_THREADS>);
    CK(set_max_dynamic_smem(kern, pl->smem_optin));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return 0;
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    if (pl->s1slots.ensure(1024 * sizeof(int))) return fail(SCQED_ENOMEM, "per-SM slot counters");
    CK(cudaMemsetAsync(pl->s1slots.p, 0, 1024 * sizeof(int), st));
    TimingScope tscope_(st, 1);
    const int rotate = getenv("SCQED_S1_TILED_ROTATE") ? atoi(getenv("SCQED_S1_TILED_ROTATE")) : 0;
    kern<<<(unsigned)grid, threads, smem, st>>>(a, (int *)pl->s1slots.p, rotate);
    tscope_.end();
    LAUNCHED();
    CK(cudaGetLastError());
    *launched = true;
    return 0;
}

// Occupancy ladder (small_ladder.cuh): the reduction of a real symmetric matrix in full storage is cut into stages of
// decreasing dimension, each launched with as many CTAs per SM as its shared-memory footprint allows.
// SCQED_S1_LADDER="n2,n3,..." fixes the stage dimensions, "0" disables the ladder.
static int s1l_launch(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    *launched = false;
    const int n = a.n;
    if (n < 40 || n > 32 * SMALL_RQ) return 0;
    if (S1LSmem::bytes(n, n, n_terms) > pl->smem_optin) return 0;
    std::vector<int> dims;                                   // dimensions of the stages after the first
    if (const char *e = getenv("SCQED_S1_LADDER")) {
        if (atoi(e) == 0) return 0;
        for (const char *q = e; *q;) {
            int v = atoi(q);
            if (v >= 8 && v < (dims.empty() ? n : dims.back()) - 3) dims.push_back(v);
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
    } else {
        // largest dimension at which 4 CTAs fit one SM (measured on the bench workload, n = 101: one hand-off at 72 gives
        // 6.60 -> 6.18 ms, a second one at 48 gives nothing: the large-m columns carry most of the work)
        const size_t cap = 227 * 1024;
        int cur = n;
        for (int occ_target : {4}) {
            int best = 0;
            for (int m = cur - 8; m >= 24; --m)
                if ((S1LSmem::bytes(m, n, n_terms) + 1024) * occ_target <= cap) { best = m; break; }
            if (best >= 24 && best <= cur - 12) { dims.push_back(best); cur = best; }
        }
    }
    if (dims.empty()) return 0;
    // hand-off buffers: stage i writes its trailing block for stage i + 1
    size_t need = 0;
    for (int m : dims) need += (size_t)a.n_pts * m * m * 8;
    if (pl->s1ladder.ensure(need)) return fail(SCQED_ENOMEM, "ladder hand-off buffers (%zu MB)", need >> 20);
    TimingScope tscope_(st, 1);
    double *buf = (double *)pl->s1ladder.p;
    const double *tin = nullptr;
    int cur = n, off = 0;
    const bool blocked_first = getenv("SCQED_S1_LADDER_BLK") && atoi(getenv("SCQED_S1_LADDER_BLK")) != 0 &&
                               (n - dims[0]) % S1B_NB == 0 && S1BSmem::bytes(n, n_terms) <= pl->smem_optin;
    for (size_t i = 0; i <= dims.size(); ++i) {
        if (i == 0 && blocked_first) {
            // first stage on the blocked kernel (read-only mat-vec pass + DMMA trailing update): the large-m columns
            const size_t smem = S1BSmem::bytes(n, n_terms);
            CK(set_max_dynamic_smem(s1b_tridiag_kernel, pl->smem_optin));
            int occ = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1b_tridiag_kernel, S1B_THREADS, smem));
            if (occ < 1) return fail(SCQED_EUNSUPPORTED, "blocked ladder stage does not fit (n=%d)", n);
            long long grid = (long long)occ * pl->sm_count;
            if (grid > a.n_pts) grid = a.n_pts;
            s1b_tridiag_kernel<<<(unsigned)grid, S1B_THREADS, smem, st>>>(a, n - dims[0], buf);
            LAUNCHED();
            CK(cudaGetLastError());
            tin = buf;
            buf += (size_t)a.n_pts * dims[0] * dims[0];
            off += cur - dims[0];
            cur = dims[0];
            continue;
        }
        S1LStage sg;
        sg.n = cur; sg.off = off; sg.Tin = tin;
        if (i < dims.size()) { sg.j_stop = cur - dims[i]; sg.Tout = buf; }
        else { sg.j_stop = cur - 1; sg.Tout = nullptr; }
        const size_t smem = S1LSmem::bytes(cur, n, n_terms);
        CK(set_max_dynamic_smem(s1l_tridiag_kernel, pl->smem_optin));
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1l_tridiag_kernel, 256, smem));
        if (occ < 1) return fail(SCQED_EUNSUPPORTED, "ladder stage does not fit (n=%d)", cur);
        long long grid = (long long)occ * pl->sm_count;
        if (grid > a.n_pts) grid = a.n_pts;
        s1l_tridiag_kernel<<<(unsigned)grid, 256, smem, st>>>(a, sg);
        LAUNCHED();
        CK(cudaGetLastError());
        if (i < dims.size()) {
            tin = buf;
            buf += (size_t)a.n_pts * dims[i] * dims[i];
            off += cur - dims[i];
            cur = dims[i];
        }
    }
    tscope_.end();
    *launched = true;
    return 0;
}

// Register-resident tridiagonalisation (small_reg.cuh), real symmetric matrices up to n = 128: SCQED_S1_REG=1.
template <int TPR, int NPAIR, int MAXT, int MAXR>
static int s1g_launch_t(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    constexpr int R = 32 / TPR, NC = 2 * TPR * NPAIR;
    const int NW = (a.n + R - 1) / R;
    const size_t smem = S1GSmem::bytes(a.n, n_terms, NC, NW);
    if (smem > pl->smem_optin || NW * 32 > MAXT) return 0;
    auto kern = s1g_tridiag_kernel<TPR, NPAIR, MAXT, MAXR>;
    CK(set_max_dynamic_smem(kern, pl->smem_optin));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, smem));
    if (occ < 1) return 0;
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    TimingScope tscope_(st, 1);
    kern<<<(unsigned)grid, NW * 32, smem, st>>>(a);
    tscope_.end();
    LAUNCHED();
    CK(cudaGetLastError());
    *launched = true;
    return 0;
}

static int s1g_launch(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    *launched = false;
    const char *e = getenv("SCQED_S1_REG");          // opt-in: measured slower than the shared-memory kernels (17.2 vs 6.4 ms)
    if (!e || atoi(e) == 0) return 0;
    const int n = a.n;
    if (n < 3 || n > 128) return 0;
    if (n <= 32) return s1g_launch_t<2, 8, 64, 128>(pl, a, n_terms, st, launched);
    if (n <= 48) return s1g_launch_t<2, 12, 96, 112>(pl, a, n_terms, st, launched);
    if (n <= 64) return s1g_launch_t<2, 16, 128, 128>(pl, a, n_terms, st, launched);
    if (n <= 80) return s1g_launch_t<2, 20, 160, 136>(pl, a, n_terms, st, launched);
    if (n <= 104) return s1g_launch_t<2, 26, 224, 144>(pl, a, n_terms, st, launched);
    return s1g_launch_t<4, 16, 512, 128>(pl, a, n_terms, st, launched);
}

template <typename T>
static int s1_launch_tridiag(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st) {
    if constexpr (sizeof(T) == sizeof(double)) {
        {
            bool launched = false;
            int rc = s1g_launch(pl, a, n_terms, st, &launched);
            if (rc) return rc;
            if (launched) return 0;
            rc = s1b_launch(pl, a, n_terms, st, &launched);
            if (rc) return rc;
            if (launched) return 0;
            rc = s1t_launch(pl, a, n_terms, st, &launched);
            if (rc) return rc;
            if (launched) return 0;
            if (!s1_rows_mode() && s1_packed_mode() != 1) {
                rc = s1l_launch(pl, a, n_terms, st, &launched);
                if (rc) return rc;
                if (launched) return 0;
            }
        }
        if (s1_rows_mode()) {
            bool launched = false;
            int rc = s1r_launch(pl, a, n_terms, st, &launched);
            if (rc) return rc;
            if (launched) return 0;
        }
    }
    const bool full_fits = a.n <= 32 * SMALL_RQ &&
        S1Smem<T>::bytes(a.n, n_terms, sizeof(T) == sizeof(double) ? 256 : 512) <= pl->smem_optin;
    if (s1_packed_mode() == 1 || (s1_packed_mode() == 2 && !full_fits)) {
        bool launched = false;
        int rc = 0;
        if (a.n <= 128) rc = s1p_launch<T, 4>(pl, a, n_terms, st, &launched);
        else if constexpr (sizeof(T) == sizeof(double)) rc = s1p_launch<T, 7>(pl, a, n_terms, st, &launched);
        if (rc) return rc;
        if (launched) return 0;
        if (a.n > 32 * SMALL_RQ) return fail(SCQED_EUNSUPPORTED, "n=%d does not fit the shared-memory tridiagonalisation", a.n);
    }
    const int threads = sizeof(T) == sizeof(double) ? 256 : 512;
    const size_t smem = S1Smem<T>::bytes(a.n, n_terms, threads);
    CK(set_max_dynamic_smem(s1_tridiag_kernel<T>, pl->smem_optin));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1_tridiag_kernel<T>, threads, smem));
    if (occ < 1) return fail(SCQED_EUNSUPPORTED, "small tridiagonalisation kernel does not fit (n=%d)", a.n);
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    TimingScope tscope_(st, sizeof(T) == sizeof(double) ? 1 : 2);
    s1_tridiag_kernel<T><<<(unsigned)grid, threads, smem, st>>>(a);
    tscope_.end();
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// number of entries of info[] equal to `val` (device-side reduction, one atomic per block that saw one)
static __global__ void __launch_bounds__(256)
count_info_kernel(const int *__restrict__ info, long long n, int val, int *__restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int hit = (i < n && info[i] == val) ? 1 : 0;
    const int c = __syncthreads_count(hit);
    if (threadIdx.x == 0 && c) atomicAdd(count, c);
}

static int run_small_split(scqed_plan *pl, const PlanDev &P, const cplx *coef, const cplx *Hin, const double *scal,
                           long long n_pts, int n, int k, int want_vec, int tidyup, const ResonatorArgs &res,
                           double *evals, cplx *evecs, double *post, int *info, cudaStream_t st) {
    const bool need_vec = want_vec || res.enabled;
    long long chunk = n_pts < 16384 ? n_pts : 16384;
    // a plan whose realness is unknown is probed on a small first chunk
    long long first = (pl->s1_real_mode < 0 && !Hin) ? (long long)2 * pl->sm_count : chunk;
    if (first > chunk) first = chunk;
    const long long G = chunk * k;
    size_t bytes = 0;
    auto need = [&](size_t b) { size_t r = bytes; bytes += (b + 255) & ~(size_t)255; return r; };
    size_t o_dd = need((size_t)chunk * n * 8), o_ee = need((size_t)chunk * n * 8), o_e2 = need((size_t)chunk * n * 8);
    size_t o_bnd = need((size_t)chunk * 4 * 8), o_tau = need((size_t)chunk * n * 16), o_lam = need((size_t)chunk * k * 8);
    size_t o_flags = need((size_t)chunk * 4), o_cnt = need(256);
    size_t o_V = 0, o_lu = 0, o_swp = 0, o_zw = 0;
    if (need_vec) {
        o_V = need((size_t)chunk * n * n * 16);
        o_lu = need((size_t)4 * n * G * 8);
        o_swp = need((size_t)n * G);
        o_zw = need((size_t)n * G * 8);
    }
    if (pl->s1work.ensure(bytes)) return fail(SCQED_ENOMEM, "small-path workspace (%zu MB)", bytes >> 20);
    unsigned char *base = (unsigned char *)pl->s1work.p;
    cplx *vec_tmp = nullptr;
    if (need_vec && !evecs) {
        if (pl->s1vec.ensure((size_t)chunk * k * n * 16)) return fail(SCQED_ENOMEM, "small-path eigenvector staging");
        vec_tmp = (cplx *)pl->s1vec.p;
    }
    S1Args a;
    memset(&a, 0, sizeof(a));
    a.P = P; a.n = n; a.k = k; a.tidyup = tidyup; a.want_vec = need_vec;
    a.W.dd = (double *)(base + o_dd); a.W.ee = (double *)(base + o_ee); a.W.e2 = (double *)(base + o_e2);
    a.W.bnd = (double *)(base + o_bnd); a.W.tau = (cplx *)(base + o_tau); a.W.lam = (double *)(base + o_lam);
    a.W.flags = (int *)(base + o_flags);
    int *d_cnt = (int *)(base + o_cnt);
    bool assumed_real = false;          // some chunk ran on the strength of pl->s1_real_mode == 1 without a host check
    if (need_vec) {
        a.W.Vst = base + o_V; a.W.lu = (double *)(base + o_lu); a.W.swp = base + o_swp; a.W.zw = (double *)(base + o_zw);
    }
    for (long long p0 = 0; p0 < n_pts;) {
        const long long np = (p0 == 0 ? first : chunk) < n_pts - p0 ? (p0 == 0 ? first : chunk) : n_pts - p0;
        a.n_pts = np;
        a.coef = coef ? coef + (size_t)p0 * P.n_terms : nullptr;
        a.dcoef = (P.n_dense > 0 && !Hin && pl->dcoef.p) ? (const double *)pl->dcoef.p + (size_t)p0 * 2 * P.n_dense : nullptr;
        a.Hin = Hin ? Hin + (size_t)p0 * n * n : nullptr;
        a.info = info + p0;
        const long long Gc = np * k;
        CK(cudaMemsetAsync(a.info, 0, (size_t)np * 4, st));
        bool real_mode = false;
        const bool direct = P.dense_tridiag && a.dcoef && !Hin;     // H(p) is already tridiagonal: no Householder stage
        if (direct) {
            TimingScope tscope_(st, 4);
            s1_direct_tridiag_kernel<<<(unsigned)((np * 32 + 127) / 128), 128, 0, st>>>(a);
            tscope_.end();
            LAUNCHED();
            real_mode = true;
        } else if (pl->s1_real_mode == 1 && !Hin) {
            // known real-symmetric plan: no host round trip; a point that turns out complex is
            // flagged info = -2 by the kernel and the host-level driver re-runs in complex mode
            int rc = s1_launch_tridiag<double>(pl, a, P.n_terms > P.n_dense ? P.n_terms : P.n_dense, st);
            if (rc) return rc;
            real_mode = true;
            assumed_real = true;
        } else if (pl->s1_real_mode != 0) {
            int rc = s1_launch_tridiag<double>(pl, a, P.n_terms > P.n_dense ? P.n_terms : P.n_dense, st);
            if (rc) return rc;
            int h_cnt = 0;
            std::vector<int> hflags((size_t)np);
            CK(cudaMemcpyAsync(hflags.data(), a.W.flags, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (long long q = 0; q < np; ++q) h_cnt += hflags[(size_t)q] != 0;
            real_mode = (h_cnt == 0);
            if (!Hin) pl->s1_real_mode = real_mode ? 1 : 0;
            if (!real_mode) CK(cudaMemsetAsync(a.info, 0, (size_t)np * 4, st));
        }
        if (!real_mode && n > 32 * SMALL_RQ) return S1_NOT_REAL;       // caller falls back to the HBM solvers
        if (!real_mode) {
            int rc = s1_launch_tridiag<cplx>(pl, a, P.n_terms > P.n_dense ? P.n_terms : P.n_dense, st);
            if (rc) return rc;
        }
        // enough (point, eigenvalue) pairs to fill the machine with one thread each: plain bisection
        // (6.6x less arithmetic); small sweeps keep the 33-way warp multisection (short dependent chain)
        const int bis_mode = getenv("SCQED_BISECT") ? atoi(getenv("SCQED_BISECT")) : 0;
        const bool thread_bisect = bis_mode == 2 || (bis_mode == 0 && n_pts * k >= 4096 && k <= 32);      // by the size of the whole call: the result of a point must not depend on how the call is chunked
        if (thread_bisect) {
            const int ppw = 32 / k;
            const long long bw = (np + ppw - 1) / ppw;
            // three-term (division-free) Sturm recurrence unless SCQED_STURM_RATIO is set
            s1_bisect_thread_kernel<<<(unsigned)((bw * 32 + 127) / 128), 128, 0, st>>>(a, nullptr, evals + (size_t)p0 * k, getenv("SCQED_STURM_RATIO") ? 0 : 1);
        } else {
            const long long warps = np * k;
            s1_bisect_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(a, nullptr, evals + (size_t)p0 * k);
        }
        LAUNCHED();
        if (need_vec) {
            const int ppw = 32 / k;
            const long long iw = (np + ppw - 1) / ppw;
            s1_invit_kernel<<<(unsigned)((iw * 32 + 127) / 128), 128, 0, st>>>(a, nullptr, Gc);
            cplx *vout = evecs ? evecs + (size_t)p0 * k * n : vec_tmp;
            if (direct) {
                const long long tot = np * k * n;
                s1_phase_backtr_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a, Gc, vout);
            } else if (real_mode && n > 32 * SMALL_RQ) {
                const long long bw = np * ((k + 7) / 8);
                s1_backtr_kernel<double, 8, 7><<<(unsigned)((bw * 32 + 127) / 128), 128, 0, st>>>(a, nullptr, Gc, vout);
            } else if (real_mode) {
                const long long bw = np * ((k + 7) / 8);
                s1_backtr_kernel<double, 8><<<(unsigned)((bw * 32 + 127) / 128), 128, 0, st>>>(a, nullptr, Gc, vout);
            } else {
                const long long bw = np * ((k + 3) / 4);
                s1_backtr_kernel<cplx, 4><<<(unsigned)((bw * 32 + 127) / 128), 128, 0, st>>>(a, nullptr, Gc, vout);
            }
            g_launches.fetch_add(2, std::memory_order_relaxed);
            if (res.enabled) {
                s1_resonator_kernel<<<(unsigned)np, 128, 0, st>>>(a, res, nullptr, scal + (size_t)p0 * P.n_scalar, vout,
                                                              post + (size_t)p0 * res.nstrips * k);
                LAUNCHED();
            }
        }
        CK(cudaGetLastError());
        p0 += np;
    }
    if (assumed_real && !pl->defer_real_check) {
        // A plan known to be real symmetric on its first points may turn complex later in the grid (e.g. a 2-D sweep
        // whose outer flux axis starts at 0): such points were flagged info = -2 and skipped by the real kernel.
        // Count them on the device; if there are any, the plan is complex Hermitian: redo this call in complex mode.
        int h_cnt = 0;
        CK(cudaMemsetAsync(d_cnt, 0, 4, st));
        count_info_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, st>>>(info, n_pts, -2, d_cnt);
        LAUNCHED();
        CK(cudaMemcpyAsync(&h_cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_cnt > 0) {
            pl->s1_real_mode = 0;
            return run_small_split(pl, P, coef, Hin, scal, n_pts, n, k, want_vec, tidyup, res, evals, evecs, post, info, st);
        }
    }
    return 0;
}

int run_small_any(scqed_plan *pl, const PlanDev &P, const cplx *coef, const cplx *Hin, const double *scal,
                         long long n_pts, int n, int k, int want_vec, int tidyup, const ResonatorArgs &res,
                         double *evals, cplx *evecs, double *post, int *info, cudaStream_t st, bool monolithic) {
    if (!monolithic && split_fits(pl, n, k, P.n_terms > P.n_dense ? P.n_terms : P.n_dense, Hin ? -1 : pl->s1_real_mode))
        return run_small_split(pl, P, coef, Hin, scal, n_pts, n, k, want_vec, tidyup, res, evals, evecs, post, info, st);
    return run_small(pl, P, coef, Hin, scal, n_pts, n, k, want_vec, tidyup, res, evals, evecs, post, info, st);
}


int run_resonator_generic(const PlanDev &P, const ResonatorArgs &res, const double *scal, const double *evals,
                          const cplx *vecs, long long n_pts, int n, int k, double *post, int *info, cudaStream_t st) {
    resonator_generic_kernel<<<(unsigned)n_pts, 128, 0, st>>>(P, res, scal, evals, vecs, n, k, post, info);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// ---- tridiagonal eigenpairs for the dense tile-packed solver (S2t, dense.cu) -----------------------------------------
// B real symmetric tridiagonal matrices (dd, ee [B][n], already on the device) through the sweep-sized kernels of the S1
// pipeline: one thread per (matrix, eigenvalue) Sturm bisection (or the warp multisection for small batches) and one
// thread per (matrix, vector) inverse iteration on workspaces interleaved over the threads.  tridiag_eig_kernel (one CTA
// per matrix: ten busy threads of 256 in the inverse iteration) took 5-6 ms per launch whatever the batch; these take
// the length of one dependent chain.  Eigenvectors come back INTERLEAVED: element i of vector (b, j) at zw[i * G + b * k + j].
static __global__ void __launch_bounds__(128)
s1_bounds_kernel(S1Args a) {
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int l = threadIdx.x & 31, n = a.n;
    if (p >= a.n_pts) return;
    const double *d = a.W.dd + p * n, *e = a.W.ee + p * n;
    double lo = 1e300, hi = -1e300, emax = 0.0;
    for (int i = l; i < n; i += 32) {
        const double ei = i < n - 1 ? e[i] : 0.0;
        const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + fabs(ei);
        lo = fmin(lo, d[i] - r);
        hi = fmax(hi, d[i] + r);
        emax = fmax(emax, ei * ei);
        a.W.e2[p * n + i] = ei * ei;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    if (l == 0) {                           // same conventions as tridiag_bounds
        const double tn = fmax(fabs(lo), fabs(hi));
        const double pad = 2.0 * tn * SCQED_EPS * n + 2.0 * SCQED_SAFMIN;
        a.W.bnd[p * 4 + 0] = lo - pad;
        a.W.bnd[p * 4 + 1] = hi + pad;
        a.W.bnd[p * 4 + 2] = fmax(SCQED_SAFMIN * fmax(1.0, emax), SCQED_SAFMIN);
        a.W.bnd[p * 4 + 3] = tn;
    }
}

int s1_tridiag_eig_batch(scqed_plan *pl, double *dd, double *ee, int n, int k, long long B, int want_vec, double *evals,
                         const double **zw_out, long long *G_out, int *info, cudaStream_t st) {
    if (k > 32) return fail(SCQED_EUNSUPPORTED, "tridiagonal batch solver: k <= 32");
    const long long G = B * k;
    size_t bytes = 0;
    auto need = [&](size_t b) { size_t r = bytes; bytes += (b + 255) & ~(size_t)255; return r; };
    const size_t o_e2 = need((size_t)B * n * 8), o_bnd = need((size_t)B * 4 * 8), o_lam = need((size_t)B * k * 8);
    size_t o_lu = 0, o_swp = 0, o_zw = 0;
    if (want_vec) {
        o_lu = need((size_t)4 * n * G * 8);
        o_swp = need((size_t)n * G);
        o_zw = need((size_t)n * G * 8);
    }
    if (pl->s1work.ensure(bytes)) return fail(SCQED_ENOMEM, "tridiagonal batch workspace (%zu MB)", bytes >> 20);
    unsigned char *base = (unsigned char *)pl->s1work.p;
    S1Args a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.k = k; a.n_pts = B; a.want_vec = want_vec; a.info = info;
    a.W.dd = dd; a.W.ee = ee; a.W.e2 = (double *)(base + o_e2); a.W.bnd = (double *)(base + o_bnd);
    a.W.lam = (double *)(base + o_lam);
    if (want_vec) { a.W.lu = (double *)(base + o_lu); a.W.swp = base + o_swp; a.W.zw = (double *)(base + o_zw); }
    CK(cudaMemsetAsync(info, 0, (size_t)B * 4, st));
    s1_bounds_kernel<<<(unsigned)((B * 32 + 127) / 128), 128, 0, st>>>(a);
    LAUNCHED();
    const int ppw = 32 / k;
    if (B * k >= 4096) {
        const long long bw = (B + ppw - 1) / ppw;
        s1_bisect_thread_kernel<<<(unsigned)((bw * 32 + 127) / 128), 128, 0, st>>>(a, nullptr, evals, 1);
    } else {
        s1_bisect_kernel<<<(unsigned)((G * 32 + 255) / 256), 256, 0, st>>>(a, nullptr, evals);
    }
    LAUNCHED();
    if (want_vec) {
        const long long iw = (B + ppw - 1) / ppw;
        s1_invit_kernel<<<(unsigned)((iw * 32 + 127) / 128), 128, 0, st>>>(a, nullptr, G);
        LAUNCHED();
    }
    CK(cudaGetLastError());
    if (zw_out) *zw_out = a.W.zw;
    if (G_out) *G_out = G;
    return 0;
}
