// dense_tiled.cuh -- S2t: REAL SYMMETRIC dense matrices beyond the shared-memory path (224 < n <= 2048), ONE PERSISTENT CTA
// PER MATRIX on the tile-packed lower triangle in global memory (L2 resident for the later panels).
//
// Why: the blocked HBM solver (dense_blocked.cuh) is complex-only and re-reads the trailing matrix with two launches per
// column; for the real symmetric matrices that dominate the dense mid-size cases (single oscillator node: fluxonium at
// truncation 300 ... 1000) it moves 16 B per element where 8 would do, and pays launch / tail latency 2 n times.  Here
//   * the matrix is packed ONCE into 8 x 8 row-major tiles of the lower triangle (dtg_pack_kernel: half the bytes of a
//     real matrix, a quarter of the complex one), and one CTA owns the matrix for the whole reduction: no launches, no
//     grid-wide synchronisation, the panels V, W and every vector live in shared memory;
//   * per column the stored half is read once (symmetric mat-vec: work item R = tile row R + tile column R, see
//     tiled_common.cuh) -- 4 m^2 bytes per column, (4/3) n^3 B per matrix -- and every NB = 4 columns the trailing tiles
//     are updated in place by DMMA (the tile is the accumulator fragment): + 2 m^2 bytes per column;
//   * the back-transformation applies the reflectors straight from the tiles to the k tridiagonal eigenvectors held in
//     shared memory (dtg_backtr_kernel).
// The HBM floor of this one-stage form is 2 n^3 B per matrix (n = 1000: 2 GB => 3 k matrices/s at 6 TB/s; the complex
// blocked solver: (8/3 + 16/3/32...) n^3 ~ 5.5 GB).  Arithmetic: LAPACK dsytrd 'L' with nb = 4, as small_tiled.cuh.
#pragma once
#include "common.cuh"
#include "tiled_common.cuh"
#include "tridiag_eig.cuh"

namespace scqed {

#define DTG_THREADS 512

struct DtgGeom {
    __host__ __device__ static int nt(int n) { return (n + 7) >> 3; }
    __host__ __device__ static size_t tile_doubles(int n) { return (size_t)nt(n) * (nt(n) + 1) / 2 * 64; }
    __host__ __device__ static int ps(int n) { return 8 * nt(n) + 4; }
    static size_t tridiag_smem(int n) { return ((size_t)2 * S1T_NB * ps(n) + ps(n) + 2 * S1T_NB + 8) * 8; }
    static size_t backtr_smem(int n, int k) { return ((size_t)k * n + 64) * 8; }
};

__device__ __forceinline__ double dtg_block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < DTG_THREADS / 32; ++q) s += red[q];
    return s;
}

// A [B][n*n] complex row-major Hermitian -> tiles [B][tile_doubles] (real parts of the lower triangle, zero padded);
// flags[0] |= 1 when some imaginary part is not zero (the caller then takes the complex solver).  One warp per tile.
static __global__ void __launch_bounds__(256)
dtg_pack_kernel(const cplx *__restrict__ A, int n, long long B, double *__restrict__ tiles, int *__restrict__ flags) {
    const int NT = DtgGeom::nt(n);
    const long long ntile = (long long)NT * (NT + 1) / 2;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (wid >= ntile * B) return;
    const long long b = wid / ntile;
    const int idx = (int)(wid - b * ntile);
    int R = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
    if ((R + 1) * (R + 2) / 2 <= idx) ++R;
    if (R * (R + 1) / 2 > idx) --R;
    const int C = idx - R * (R + 1) / 2;
    const int i = 8 * R + g, j = 8 * C + 2 * t;
    const cplx *Ab = A + (size_t)b * n * n;
    double2 out = make_double2(0.0, 0.0);
    int bad = 0;
    if (i < n) {
        // diagonal tiles: element (i, j) with j > i comes from the stored (j, i) by symmetry
        if (j < n) { const cplx v = j <= i ? Ab[(size_t)i * n + j] : Ab[(size_t)j * n + i]; out.x = v.x; bad |= v.y != 0.0; }
        if (j + 1 < n) { const cplx v = j + 1 <= i ? Ab[(size_t)i * n + j + 1] : Ab[(size_t)(j + 1) * n + i]; out.y = v.x; bad |= v.y != 0.0; }
    }
    *reinterpret_cast<double2 *>(tiles + (size_t)b * DtgGeom::tile_doubles(n) + (size_t)idx * 64 + g * 8 + 2 * t) = out;
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flags, 1);
}

// One persistent CTA per matrix.  Outputs d, e, tau [B][n]; the reflectors stay in the tiles (column j, rows > j).
static __global__ void __launch_bounds__(DTG_THREADS)
dtg_tridiag_kernel(double *__restrict__ tiles_all, int n, long long B, double *__restrict__ dd_all, double *__restrict__ ee_all,
                   double *__restrict__ tau_all) {
    constexpr int NB = S1T_NB, NTHR = DTG_THREADS;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[NTHR / 32];
    __shared__ double s_scal[2];
    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31;
    const int NT = DtgGeom::nt(n), PS = DtgGeom::ps(n);
    double *Vs = (double *)smem;
    double *Ws = Vs + NB * PS;
    double *Y = Ws + NB * PS;
    double *ts = Y + PS;
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        double *tiles = tiles_all + (size_t)b * DtgGeom::tile_doubles(n);
        double *dd = dd_all + (size_t)b * n, *ee = ee_all + (size_t)b * n, *tauo = tau_all + (size_t)b * n;
        for (int q = tid; q < 2 * NB * PS + PS; q += NTHR) Vs[q] = 0.0;
        __syncthreads();
        for (int j0 = 0; j0 < n; j0 += NB) {
            const int pn = min(NB, n - j0);
            for (int jj = 0; jj < pn; ++jj) {
                const int j = j0 + jj, Cj = j >> 3, cj = j & 7;
                double *vj = Vs + jj * PS;
                // ---- S1: column j made current (kept in the new panel column until it is scaled into v) ----
                double xn = 0.0;
                for (int r = j + tid; r < n; r += NTHR) {
                    double c = tiles[s1t_tile(r >> 3, Cj) + (r & 7) * 8 + cj];
                    for (int k = 0; k < jj; ++k)
                        c = fma(-Vs[k * PS + r], Ws[k * PS + j], fma(-Ws[k * PS + r], Vs[k * PS + j], c));
                    vj[r] = c;
                    if (r == j) dd[j] = c;
                    if (r == j + 1) s_scal[0] = c;
                    if (r >= j + 2) xn = fma(c, c, xn);
                }
                if (j == n - 1) {
                    if (tid == 0) { ee[j] = 0.0; tauo[j] = 0.0; }
                    break;
                }
                xn = dtg_block_sum(xn, red);
                double tau, beta, sc;
                s1t_reflector_scalars(s_scal[0], xn, tau, beta, sc);
                for (int r = j + tid; r < n; r += NTHR) {
                    const double v = r == j ? 0.0 : (r == j + 1 ? 1.0 : vj[r] * sc);
                    vj[r] = v;
                    if (r > j) tiles[s1t_tile(r >> 3, Cj) + (r & 7) * 8 + cj] = v;       // kept for the back-transformation
                }
                if (tid == 0) { ee[j] = beta; tauo[j] = tau; }
                __syncthreads();
                s1t_pass<NTHR, 8>(tiles, Vs, Ws, Y, ts, n, NT, PS, j, jj, wp, lane, 0);
                __syncthreads();
                // ---- S3: corrections, w ----
                double dt = 0.0;
                for (int r = j + 1 + tid; r < n; r += NTHR) {
                    double y = Y[r];
                    for (int k = 0; k < jj; ++k) y = fma(-Vs[k * PS + r], ts[k], fma(-Ws[k * PS + r], ts[NB + k], y));
                    const double pr = tau * y;
                    Y[r] = pr;
                    dt = fma(pr, vj[r], dt);
                }
                dt = dtg_block_sum(dt, red);
                const double hc = -0.5 * tau * dt;
                for (int r = tid; r < PS; r += NTHR) Ws[jj * PS + r] = (r > j && r < n) ? fma(hc, vj[r], Y[r]) : 0.0;
                __syncthreads();
            }
            const int b0 = j0 + NB;
            if (pn == NB && b0 < n) {
                s1t_update<NTHR, 8>(tiles, Vs, Ws, NT, PS, b0, wp, lane);
                __syncthreads();
                // panel columns of the NEXT panel start from zero below their first live row (stale rows are never read
                // above j, but t1 / t2 and the update read whole columns)
                for (int q = tid; q < 2 * NB * PS; q += NTHR) Vs[q] = 0.0;
                __syncthreads();
            }
        }
        __syncthreads();
    }
}

// x = H(0) H(1) ... H(n-2) z for the k tridiagonal eigenvectors of one matrix (Z [B][k][n] real): one CTA per matrix,
// the vectors in shared memory, reflector j read straight from column j of the tiles.  X complex [B][k][n].
template <int KMAX>
static __global__ void __launch_bounds__(DTG_THREADS)
dtg_backtr_kernel(const double *__restrict__ tiles_all, int n, int k, const double *__restrict__ tau_all,
                  const double *__restrict__ Z, cplx *__restrict__ X) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[DTG_THREADS / 32][KMAX];
    double *xs = (double *)smem;                             // [k][n]
    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31;
    const long long b = blockIdx.x;
    const double *tiles = tiles_all + (size_t)b * DtgGeom::tile_doubles(n);
    const double *tau = tau_all + (size_t)b * n;
    for (int q = tid; q < k * n; q += DTG_THREADS) xs[q] = Z[(size_t)b * k * n + q];
    __syncthreads();
    for (int j = n - 2; j >= 0; --j) {
        const double tj = tau[j];
        if (tj == 0.0) continue;
        const int Cj = j >> 3, cj = j & 7;
        double dot[KMAX];
#pragma unroll
        for (int u = 0; u < KMAX; ++u) dot[u] = 0.0;
        // this thread's rows of the reflector stay in registers between the dot products and the update (n <= 2048: <= 8)
        double vr[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = j + 1 + tid + q * DTG_THREADS;
            vr[q] = r < n ? (r == j + 1 ? 1.0 : tiles[s1t_tile(r >> 3, Cj) + (r & 7) * 8 + cj]) : 0.0;
            if (r < n) {
#pragma unroll
                for (int u = 0; u < KMAX; ++u) if (u < k) dot[u] = fma(vr[q], xs[u * n + r], dot[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < KMAX; ++u) if (u < k) dot[u] = warp_sum(dot[u]);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < KMAX; ++u) if (u < k) red[wp][u] = dot[u];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < KMAX; ++u) {
            if (u < k) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < DTG_THREADS / 32; ++w) s += red[w][u];
                dot[u] = -tj * s;
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = j + 1 + tid + q * DTG_THREADS;
            if (r < n) {
#pragma unroll
                for (int u = 0; u < KMAX; ++u) if (u < k) xs[u * n + r] = fma(dot[u], vr[q], xs[u * n + r]);
            }
        }
        __syncthreads();
    }
    cplx *out = X + (size_t)b * k * n;
    for (int q = tid; q < k * n; q += DTG_THREADS) out[q] = cmake(xs[q], 0.0);
}

}  // namespace scqed
