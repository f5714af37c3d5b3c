// small_eigh.cuh -- S1: the fused small-n path.  One CTA per sweep point; H(p) is assembled
// straight into shared memory and never touches HBM.
//
//   assemble (K2)  -> Householder tridiagonalisation (zhetd2, lower)  -> Sturm multisection
//   -> inverse iteration -> back-transformation -> (optional) resonator RWA strips (K7)
//
// Replaces, per sweep point, pyscqed/numerical_system.py:509-594 (getHamiltonian),
// pyscqed/util.py:127-176 (diagDenseH -> LAPACK zheevr) and numerical_system.py:736-784
// (getResonatorResponse).  Only c_t(p) comes in and only E, V, Erwa go out, so the stage is
// FP64-pipe / latency bound, not HBM bound.
//
// Shared-memory layout (host: small_smem_layout):  A[n*n] complex, column-major with the LOWER
// triangle authoritative during the reduction (both triangles are kept consistent so the
// matrix-vector product can stream contiguous columns: A[i + j*n] == H[i][j]).
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "resonator.cuh"

namespace scqed {

#define SMALL_THREADS 512
#define SMALL_SMAX 16

struct SmallLayout {
    int off_A, off_d, off_e, off_e2, off_tau, off_v, off_w, off_part, off_red, off_lam, off_bounds,
        off_X, off_lu, total;
    int vb;
};

struct SmallArgs {
    PlanDev P;
    const cplx *coef;       // [n_pts, n_terms]   (NULL when Hin is given)
    const cplx *Hin;        // optional [n_pts, n, n] row-major pre-assembled matrices
    const double *scalars;  // [n_pts, n_scalar]
    long long n_pts;
    int n, k, want_vec, tidyup;
    double *evals;          // [n_pts, k]
    cplx *evecs;            // [n_pts, k, n] or NULL
    int *info;              // [n_pts]
    ResonatorArgs res;
    double *post;           // [n_pts, res.nstrips, k] or NULL
    SmallLayout L;
};

// two-stage shuffle reductions (blockDim.x <= 1024): warp sums -> smem -> every warp re-reduces
__device__ __forceinline__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = l < nw ? red[l] : 0.0;
    return warp_sum(s);
}
__device__ __forceinline__ cplx block_sum(cplx v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) { red[2 * w] = v.x; red[2 * w + 1] = v.y; }
    __syncthreads();
    cplx s = l < nw ? cmake(red[2 * l], red[2 * l + 1]) : cmake(0.0, 0.0);
    return warp_sum(s);
}

// ---- scalar helpers shared by the complex-Hermitian and the real-symmetric code paths --------------
__device__ __forceinline__ double nzero(double) { return 0.0; }
__device__ __forceinline__ cplx nzero(cplx) { return cmake(0.0, 0.0); }
__device__ __forceinline__ double none(double) { return 1.0; }
__device__ __forceinline__ cplx none(cplx) { return cmake(1.0, 0.0); }
__device__ __forceinline__ double nconj(double a) { return a; }
__device__ __forceinline__ cplx nconj(cplx a) { return cconj(a); }
__device__ __forceinline__ double nmul(double a, double b) { return a * b; }
__device__ __forceinline__ cplx nmul(cplx a, cplx b) { return cmul(a, b); }
__device__ __forceinline__ double nadd(double a, double b) { return a + b; }
__device__ __forceinline__ cplx nadd(cplx a, cplx b) { return cadd(a, b); }
__device__ __forceinline__ double nsub(double a, double b) { return a - b; }
__device__ __forceinline__ cplx nsub(cplx a, cplx b) { return csub(a, b); }
__device__ __forceinline__ double nscale(double s, double a) { return s * a; }
__device__ __forceinline__ cplx nscale(double s, cplx a) { return cscale(s, a); }
__device__ __forceinline__ void nfma(double &acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void nfma(cplx &acc, cplx a, cplx b) { cfma(acc, a, b); }
__device__ __forceinline__ void nfmac(double &acc, double a, double b) { acc = fma(a, b, acc); }     // += conj(a) b
__device__ __forceinline__ void nfmac(cplx &acc, cplx a, cplx b) { cfmac(acc, a, b); }
__device__ __forceinline__ void nfnmac(double &acc, double a, double b) { acc = fma(-a, b, acc); }   // -= a conj(b)
__device__ __forceinline__ void nfnmac(cplx &acc, cplx a, cplx b) { cfnmac(acc, a, b); }
__device__ __forceinline__ double nabs2(double a) { return a * a; }
__device__ __forceinline__ double nabs2(cplx a) { return cabs2(a); }
__device__ __forceinline__ double nre(double a) { return a; }
__device__ __forceinline__ double nre(cplx a) { return a.x; }
__device__ __forceinline__ bool niszero(double a) { return a == 0.0; }
__device__ __forceinline__ bool niszero(cplx a) { return a.x == 0.0 && a.y == 0.0; }
__device__ __forceinline__ double nshfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx nshfl(cplx v, int src) {
    return cmake(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// zlarfg / dlarfg: reflector scalars from alpha and the squared norm of the tail
__device__ __forceinline__ void reflector_scalars(double alpha, double xn, double &tau, double &beta, double &sc) {
    if (xn == 0.0) { tau = 0.0; beta = alpha; sc = 0.0; return; }
    beta = -copysign(sqrt(alpha * alpha + xn), alpha);
    tau = (beta - alpha) / beta;
    sc = 1.0 / (alpha - beta);
}
__device__ __forceinline__ void reflector_scalars(cplx alpha, double xn, cplx &tau, double &beta, cplx &sc) {
    if (xn == 0.0 && alpha.y == 0.0) { tau = cmake(0.0, 0.0); beta = alpha.x; sc = cmake(0.0, 0.0); return; }
    beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn), alpha.x);
    tau = cmake((beta - alpha.x) / beta, -alpha.y / beta);
    cplx den = cmake(alpha.x - beta, alpha.y);
    double dn = 1.0 / cabs2(den);
    sc = cmake(den.x * dn, -den.y * dn);
}

#define SMALL_RQ 4      // n <= 32 * SMALL_RQ = 128 on the small path
#define SMALL_SER 128   // threads of the serial section (4 warps, one row each)
#define SMALL_SEGS 8    // column segments of the fused pass (partial sums per row)

__device__ __forceinline__ void ser_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// sum over the 128 threads of the serial section (two named-barrier phases); `slot` = 8 doubles
__device__ __forceinline__ double ser_sum(double v, double *slot) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = v;
    ser_barrier();
    return slot[0] + slot[1] + slot[2] + slot[3];
}

// Threads 0..127 (row r = tid): turn column `col` (col[0] = diagonal, col[1] = alpha, col[2..m) =
// tail) -- after subtracting the pending rank-2 update  vc[r] conj(w0) + wr conj(vc[0]=1)  when
// `upd` -- into the next reflector.  Writes vnext[0..m-1) (vnext[0] = 1), keeps it in col[1..m),
// and d/e/tau of column jn.  sm: 16 doubles of scratch + 2 T.
template <typename T>
__device__ __forceinline__ void ser_next_reflector(T *col, int m, bool upd, const T *vc, T w0, T wr,
                                                   T *vnext, double *dv, double *ev, T *tau, int jn,
                                                   double *sm, T *sm_t) {
    const int r = threadIdx.x;
    T c = nzero(T());
    double xn = 0.0;
    if (r < m) {
        c = col[r];
        if (upd) { nfnmac(c, vc[r], w0); c = nsub(c, wr); }
        if (r >= 2) xn = nabs2(c);
        if (r < 2) sm_t[r] = c;                     // diagonal and alpha
    }
    xn = ser_sum(xn, sm + 8);                       // (barrier also publishes sm_t)
    const T diag = sm_t[0];
    if (m < 2) {                                    // 1 x 1 trailing block: only the diagonal is left
        if (r == 0) { dv[jn] = nre(diag); ev[jn] = 0.0; }
        return;
    }
    const T alpha = sm_t[1];
    T tj, sc;
    double beta;
    reflector_scalars(alpha, xn, tj, beta, sc);
    const bool zero = niszero(tj);
    if (r >= 1 && r < m) {
        T v = r == 1 ? none(T()) : (zero ? nzero(T()) : nmul(c, sc));
        vnext[r - 1] = v;
        col[r] = v;                                 // reflector kept for the back-transformation
    }
    if (r == 0) { col[0] = diag; dv[jn] = nre(diag); ev[jn] = beta; tau[jn] = tj; }
}

// Householder tridiagonalisation of the full-storage column-major matrix A (n <= 128), LAPACK
// z/d-hetd2 'L' arithmetic, restructured so that one pass over the trailing block applies the
// PENDING rank-2 update of step j-1 and accumulates A v_j of step j, and the O(m) work between
// passes runs on four warps with named barriers: two block barriers per column instead of nine.
// j_stop < n - 1: PARTIAL reduction -- reflectors 0 .. j_stop-1 are formed and applied (d, e, tau of those columns are
// set, the reflectors stay in their columns), then A[j_stop:, j_stop:] holds the exact trailing matrix whose reduction a
// later stage continues as an independent (n - j_stop)-dimensional problem (small_ladder.cuh).
template <typename T, bool PARTIAL = false>
__device__ void small_tridiag(T *A, int n, double *dv, double *ev, T *tau, T *vbuf, T *wbuf, T *part, int j_stop = 1 << 30) {
    __shared__ double ser_sm[24];
    __shared__ cplx ser_t[2];
    const int tid = threadIdx.x, NT = blockDim.x;
    T *vcur = vbuf, *vprev = vbuf + n, *wprev = wbuf, *wcur = wbuf + n;
    if (tid < SMALL_SER)
        ser_next_reflector<T>(A, n, false, vcur, nzero(T()), nzero(T()), vcur, dv, ev, tau, 0, ser_sm, (T *)ser_t);
    __syncthreads();
    // fused pass mapping: warp w -> column segment s = w % 8, row group rg = w / 8; a thread owns
    // rows lane + 32 q for RPT consecutive q (register tile), so the three broadcast operands of a
    // column are loaded once per RPT elements and there is no integer division in the loop.
    constexpr int S = SMALL_SEGS;
    const int NW = NT >> 5, RG = NW / S > 0 ? NW / S : 1;
    const int RPT = SMALL_RQ / RG > 0 ? SMALL_RQ / RG : 1;
    const int wp_ = tid >> 5, lane = tid & 31;
    const int sidx = wp_ % S, q0 = (wp_ / S) * RPT;
    for (int j = 0; j < n - 1; ++j) {
        const int m = n - j - 1;
        T *A22 = A + (size_t)(j + 1) * (n + 1);
        const int seg = (m + S - 1) / S;
        if (wp_ < S * RG) {
            const int c0 = sidx * seg, c1 = min(m, c0 + seg);
            T acc[SMALL_RQ], vr[SMALL_RQ], wr[SMALL_RQ];
#pragma unroll
            for (int q = 0; q < SMALL_RQ; ++q) {
                acc[q] = nzero(T());
                int r = lane + 32 * (q0 + q);
                bool on = q < RPT && r < m && j > 0;
                vr[q] = on ? vprev[r + 1] : nzero(T());
                wr[q] = on ? wprev[r + 1] : nzero(T());
            }
            for (int c = c0; c < c1; ++c) {
                const T vc = vcur[c];
                const T wpc = j > 0 ? wprev[c + 1] : nzero(T());
                const T vpc = j > 0 ? vprev[c + 1] : nzero(T());
#pragma unroll
                for (int q = 0; q < SMALL_RQ; ++q) {
                    int r = lane + 32 * (q0 + q);
                    if (q < RPT && r < m) {
                        T x = A22[r + (size_t)c * n];
                        if (j > 0) {
                            nfnmac(x, vr[q], wpc);
                            nfnmac(x, wr[q], vpc);
                            A22[r + (size_t)c * n] = x;
                        }
                        nfma(acc[q], x, vc);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < SMALL_RQ; ++q) {
                int r = lane + 32 * (q0 + q);
                if (q < RPT && r < m) part[sidx * m + r] = acc[q];
            }
        }
        __syncthreads();
        if (tid < SMALL_SER) {
            const T tj = tau[j];
            T pr = nzero(T());
            double dx = 0.0, dy = 0.0;
            if (tid < m) {
                T sum = nzero(T());
                for (int u = 0; u < SMALL_SEGS; ++u) sum = nadd(sum, part[u * m + tid]);
                pr = nmul(tj, sum);
                T d1 = nzero(T());
                nfmac(d1, pr, vcur[tid]);                        // p^H v
                dx = nre(d1);
                if constexpr (sizeof(T) != sizeof(double)) dy = ((cplx *)&d1)->y;
            }
            // both components in one pass: pack (dx, dy) through two slots
            dx = warp_sum(dx); dy = warp_sum(dy);
            if ((tid & 31) == 0) { ser_sm[tid >> 5] = dx; ser_sm[4 + (tid >> 5)] = dy; }
            ser_barrier();
            dx = ser_sm[0] + ser_sm[1] + ser_sm[2] + ser_sm[3];
            dy = ser_sm[4] + ser_sm[5] + ser_sm[6] + ser_sm[7];
            T dot;
            if constexpr (sizeof(T) == sizeof(double)) dot = dx; else dot = cmake(dx, dy);
            const T a2 = nmul(nscale(-0.5, tj), dot);            // -1/2 tau (p^H v)
            T wr = nzero(T());
            if (tid < m) { wr = nadd(pr, nmul(a2, vcur[tid])); wcur[tid] = wr; }
            if (tid == 0) ((T *)ser_t)[0] = wr;                  // w[0]
            ser_barrier();
            const T w0 = ((T *)ser_t)[0];
            ser_barrier();                                       // ser_t is reused below
            // column j+1 with this step's update applied -> reflector j+1 (or the last diagonal)
            if (!PARTIAL || j + 1 != j_stop)
                ser_next_reflector<T>(A22, m, true, vcur, w0, wr, vprev /* becomes vcur */, dv, ev, tau, j + 1,
                                      ser_sm, (T *)ser_t);
        }
        __syncthreads();
        if (PARTIAL && j + 1 == j_stop) {
            // hand-off: the pending rank-2 update of this step goes into the whole trailing block, which then IS
            // the remaining problem
            for (int idx = tid; idx < m * m; idx += NT) {
                const int c = idx / m, r = idx - c * m;
                T x = A22[r + (size_t)c * n];
                nfnmac(x, vcur[r], wcur[c]);
                nfnmac(x, wcur[r], vcur[c]);
                A22[r + (size_t)c * n] = x;
            }
            __syncthreads();
            return;
        }
        T *t = vcur; vcur = vprev; vprev = t;
        t = wcur; wcur = wprev; wprev = t;
    }
}

// x = H(0) H(1) ... H(n-2) z for k vectors; one warp per vector, no block barriers.
// X is complex [k][n]; the real-symmetric path only touches the real parts.
template <typename T>
__device__ void small_backtransform(const T *A, int n, const T *tau, cplx *X, int k) {
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31, nw = blockDim.x >> 5;
    for (int jv = w; jv < k; jv += nw) {
        cplx *x = X + (size_t)jv * n;
        for (int j = n - 2; j >= 0; --j) {
            const T tj = tau[j];
            if (niszero(tj)) continue;
            const int m = n - j - 1;
            const T *v = A + (j + 1) + (size_t)j * n;
            cplx *xs = x + j + 1;
            if constexpr (sizeof(T) == sizeof(double)) {
                const double *vd = (const double *)v;
                const double td = nre(tj);
                double dot = 0.0;
                for (int r = l; r < m; r += 32) dot = fma(r == 0 ? 1.0 : vd[r], xs[r].x, dot);
                dot = warp_sum(dot) * td;
                for (int r = l; r < m; r += 32) xs[r].x = fma(-dot, r == 0 ? 1.0 : vd[r], xs[r].x);
            } else {
                const cplx *vc = (const cplx *)v;
                const cplx tc = *(const cplx *)&tj;
                cplx dot = cmake(0.0, 0.0);
                for (int r = l; r < m; r += 32) cfmac(dot, r == 0 ? cmake(1.0, 0.0) : vc[r], xs[r]);   // v^H x
                dot = warp_sum(dot);
                cplx f = cmul(tc, dot);
                for (int r = l; r < m; r += 32) {
                    cplx t = xs[r], vr = r == 0 ? cmake(1.0, 0.0) : vc[r];
                    t.x -= f.x * vr.x - f.y * vr.y;
                    t.y -= f.x * vr.y + f.y * vr.x;
                    xs[r] = t;
                }
            }
            __syncwarp();
        }
    }
}

static __global__ void __launch_bounds__(SMALL_THREADS, 1)
small_eigh_kernel(SmallArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmallLayout &L = a.L;
    cplx *A = (cplx *)(smem + L.off_A);
    double *dv = (double *)(smem + L.off_d);
    double *ev = (double *)(smem + L.off_e);
    double *e2 = (double *)(smem + L.off_e2);
    cplx *tau = (cplx *)(smem + L.off_tau);
    cplx *vv = (cplx *)(smem + L.off_v);
    cplx *ww = (cplx *)(smem + L.off_w);
    cplx *part = (cplx *)(smem + L.off_part);
    double *red = (double *)(smem + L.off_red);
    double *lam = (double *)(smem + L.off_lam);
    double *bounds = (double *)(smem + L.off_bounds);
    cplx *X = (cplx *)(smem + L.off_X);
    double *lu = (double *)(smem + L.off_lu);

    const int n = a.n, k = a.k, tid = threadIdx.x, T = blockDim.x;
    for (long long p = blockIdx.x; p < a.n_pts; p += gridDim.x) {
        // ---- assemble H(p) into shared memory -------------------------------------------------
        if (a.Hin) {
            const cplx *Hp = a.Hin + (size_t)p * n * n;
            for (int idx = tid; idx < n * n; idx += T) {
                int i = idx / n, j = idx - i * n;          // row-major source, coalesced
                A[i + j * n] = Hp[idx];
            }
        } else {
            cplx *s_coef = part;                            // scratch, free at this point
            for (int t = tid; t < a.P.n_terms; t += T) s_coef[t] = a.coef[p * a.P.n_terms + t];
            __syncthreads();
            for (int idx = tid; idx < n * n; idx += T) {
                int j = idx / n, i = idx - j * n;           // column-major destination
                int id[SCQED_MAX_NODES], jd[SCQED_MAX_NODES];
                digits(a.P, i, id);
                digits(a.P, j, jd);
                cplx v = h_element(a.P, s_coef, id, jd);
                if (a.tidyup) v = tidy(v, 1e-12);
                A[idx] = v;
            }
        }
        __syncthreads();

        // ---- real-symmetric fast path: H(p) exactly real (after the reference's 1e-12 tidy-up the
        // single-phase oscillator Hamiltonians are) -> compact to doubles, 4x fewer flops ------------
        int has_imag = 0;
        for (int idx = tid; idx < n * n; idx += T) has_imag |= (A[idx].y != 0.0);
        const bool is_real = !__syncthreads_or(has_imag);
        if (is_real) {
            double *Ar = (double *)A;
            for (int base = 0; base < n * n; base += T) {       // in-place compaction, ascending chunks
                int idx = base + tid;
                double x = idx < n * n ? A[idx].x : 0.0;
                __syncthreads();
                if (idx < n * n) Ar[idx] = x;
                __syncthreads();
            }
            small_tridiag<double>(Ar, n, dv, ev, (double *)tau, (double *)vv, (double *)ww, (double *)part);
        } else {
            small_tridiag<cplx>(A, n, dv, ev, tau, vv, ww, part);
        }
        if (tid == 0) ev[n - 1] = 0.0;
        __syncthreads();
        for (int i = tid; i < n; i += T) e2[i] = ev[i] * ev[i];
        __syncthreads();

        // ---- lowest-k eigenvalues of T ----------------------------------------------------------
        tridiag_bounds(dv, ev, n, bounds, red);
        int bad = tridiag_bisect(dv, e2, n, k, bounds, lam);
        __syncthreads();
        for (int i = tid; i < k; i += T) a.evals[p * k + i] = lam[i];

        // ---- eigenvectors: inverse iteration on T, then x = Q z -----------------------------------
        if (a.want_vec) {
            for (int idx = tid; idx < k * n; idx += T) X[idx] = cmake(0.0, 0.0);
            __syncthreads();
            bad |= tridiag_inverse_iteration(dv, ev, n, k, lam, bounds[3], (double *)X, 2 * n, 2, lu, L.vb,
                                             (unsigned long long)p + 1ULL);
            if (is_real) small_backtransform<double>((const double *)A, n, (const double *)tau, X, k);
            else small_backtransform<cplx>(A, n, tau, X, k);
            __syncthreads();
            if (a.evecs) {
                cplx *out = a.evecs + (size_t)p * k * n;
                for (int idx = tid; idx < k * n; idx += T) out[idx] = X[idx];
            }
            if (a.res.enabled) {
                __syncthreads();
                // scratch: the Householder matrix A is dead now -> reuse it
                resonator_strips(a.P, a.res, a.scalars + p * a.P.n_scalar, lam, X, n, k,
                                 (double *)A, a.post + (size_t)p * a.res.nstrips * k);
            }
        }
        bad = __syncthreads_or(bad);
        if (tid == 0) a.info[p] = bad;
        __syncthreads();
    }
}

// Host: carve shared memory.  Returns total bytes; vb = vectors per inverse-iteration batch.
static inline SmallLayout small_smem_layout(int n, int k, int want_vec, int n_terms, size_t max_bytes) {
    SmallLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return (int)r; };
    L.off_A = take((size_t)n * n * 16);
    L.off_d = take((size_t)n * 8);
    L.off_e = take((size_t)n * 8);
    L.off_e2 = take((size_t)n * 8);
    L.off_tau = take((size_t)n * 16);
    L.off_v = take((size_t)2 * n * 16);      // current + previous reflector
    L.off_w = take((size_t)2 * n * 16);
    size_t part = (size_t)SMALL_SEGS * n;
    if (part < (size_t)n_terms) part = n_terms;
    if (part < 64) part = 64;
    L.off_part = take(part * 16);
    L.off_red = take(96 * 8);
    L.off_lam = take((size_t)(k > 0 ? k : 1) * 8);
    L.off_bounds = take(4 * 8);
    L.off_X = take(want_vec ? (size_t)k * n * 16 : 0);
    L.off_lu = (int)o;
    L.vb = 0;
    if (want_vec) {
        size_t per = ((size_t)4 * n + (size_t)((n + 7) / 8)) * 8;
        long long room = (long long)max_bytes - 1024 - (long long)o;
        int vb = room > 0 ? (int)(room / (long long)per) : 0;
        if (vb > k) vb = k;
        if (vb > SMALL_THREADS) vb = SMALL_THREADS;
        L.vb = vb;
        o += (size_t)(vb > 0 ? vb : 0) * per;
    }
    L.total = (int)o;
    return L;
}

}  // namespace scqed
