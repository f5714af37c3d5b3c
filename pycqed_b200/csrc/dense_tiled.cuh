// dense_tiled.cuh -- S2t: REAL SYMMETRIC dense matrices beyond the shared-memory path (224 < n <= 3400), ONE PERSISTENT CTA
// PER MATRIX on the tile-packed lower triangle in global memory.  Two tridiagonalisation kernels: dtg_tridiag_kernel (first
// half of this file: row-major tiles, register loads, the fastest below n ~ 550) and the streamed dtc_tridiag_kernel (second
// half: column-major tiles through a bulk-copy ring, every tile read once per column, update fused into a mat-vec pass).
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
    static size_t backtr_smem(int n) { return (size_t)8 * n * 8; }            // DTG_BT_WARPS vectors
};

// Tile order in global memory.  CM = false: row-major over tiles (tiled_common.cuh, what dtg_tridiag_kernel walks); CM = true:
// COLUMN-major over tiles -- tile column C holds (C, C), (C + 1, C), ... (NT - 1, C) back to back, so the trailing block of
// every Householder step (tile columns >= T0) is ONE contiguous suffix of the array: the streamed kernel below reads it
// with linear bulk copies.
template <bool CM>
__device__ __forceinline__ int dtg_off(int R, int C, int NT) {
    return CM ? ((C * NT - ((C * (C - 1)) >> 1) + R - C) << 6) : s1t_tile(R, C);
}

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
template <bool CM>
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
    *reinterpret_cast<double2 *>(tiles + (size_t)b * DtgGeom::tile_doubles(n) + (size_t)dtg_off<CM>(R, C, NT) + g * 8 + 2 * t) = out;
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

// x = H(0) H(1) ... H(n-2) z for the tridiagonal eigenvectors (Z interleaved: element i of vector `pair` at Z[i * G + pair],
// as s1_tridiag_eig_batch leaves them): ONE WARP per (matrix, vector), the vector in
// shared memory with a fixed row -> lane mapping (row r belongs to lane r mod 32: no hazards between lanes, no barriers of
// any kind), reflector j read straight from column j of the tiles (the warps of one matrix share those lines through L1 /
// L2).  Per column: one warp reduction instead of the block-wide reduction of k dot products (2 block barriers + 10 warp
// reductions per thread) of the first version, which cost as much as the tridiagonalisation itself at n = 256
// (15.0 ms vs 15.5 ms per 2072 matrices).  X complex [B][k][n].
#define DTG_BT_WARPS 8
template <bool CM>
static __global__ void __launch_bounds__(32 * DTG_BT_WARPS)
dtg_backtr_kernel(const double *__restrict__ tiles_all, int n, int k, long long B, const double *__restrict__ tau_all,
                  const double *__restrict__ Z, long long G, cplx *__restrict__ X) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pair = (long long)blockIdx.x * DTG_BT_WARPS + wp;
    if (pair >= B * k) return;
    const long long b = pair / k;
    const int NT = DtgGeom::nt(n);
    double *x = (double *)smem + (size_t)wp * n;
    const double *tiles = tiles_all + (size_t)b * DtgGeom::tile_doubles(n);
    const double *tau = tau_all + (size_t)b * n;
    for (int i = lane; i < n; i += 32) x[i] = Z[(size_t)i * G + pair];
    for (int j = n - 2; j >= 0; --j) {
        const double tj = tau[j];
        if (tj == 0.0) continue;
        const int Cj = j >> 3, cj = j & 7;
        // rows r > j with r mod 32 == lane
        const int r0 = ((j + 1 - lane + 31) & ~31) + lane;
        double dot = 0.0;
        if (CM) {
            // column-major tiles: element (r, j) at (colstart(Cj) - Cj) * 64 + 8 r + cj
            const double *col = tiles + (size_t)dtg_off<true>(Cj, Cj, NT) - (size_t)Cj * 64 + cj;
            for (int r = r0; r < n; r += 32) dot = fma(r == j + 1 ? 1.0 : col[8 * r], x[r], dot);
            dot = -tj * warp_sum(dot);
            for (int r = r0; r < n; r += 32) x[r] = fma(dot, r == j + 1 ? 1.0 : col[8 * r], x[r]);
        } else {
            for (int r = r0; r < n; r += 32)
                dot = fma(r == j + 1 ? 1.0 : tiles[dtg_off<false>(r >> 3, Cj, NT) + (r & 7) * 8 + cj], x[r], dot);
            dot = -tj * warp_sum(dot);
            for (int r = r0; r < n; r += 32)
                x[r] = fma(dot, r == j + 1 ? 1.0 : tiles[dtg_off<false>(r >> 3, Cj, NT) + (r & 7) * 8 + cj], x[r]);
        }
    }
    cplx *out = X + (size_t)pair * n;
    for (int i = lane; i < n; i += 32) out[i] = cmake(x[i], 0.0);
}

// ======================================================================================================================
// S2t, streamed (dtc_*): the same reduction for the sizes where the matrix is read from HBM / L2 every column
// (n >~ 600).  dtg_tridiag_kernel reads every stored tile TWICE per column (item R = tile row R + tile column R) through
// register loads of which the compiler keeps two or three in flight per warp: ~10 GB/s per SM.  Here
//   * the tiles are stored column-major (dtg_off<true>): the trailing block is one contiguous suffix, which a PRODUCER warp
//     streams through a shared-memory BYTE ring with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx): pieces of at
//     most DTC_PIECE tiles = 32 KB of one tile column, up to 16 pieces / 64-128 KB in flight per SM;
//   * every tile is read ONCE per column: the consumer warp R mod 16 takes tile (R, C) out of the ring (one LDS.128 per
//     lane) and forms both y_R += T v_C (row part: four partial sums per row, one per lane of the quad, accumulated in
//     place in Yd4[32 R + lane]: no shuffles, rows owned by one warp) and y_C += T^T v_R (column part: two accumulators per
//     lane carried over the warp's tiles of the column, reduced over the lanes once per column and handed to a FOLDER warp
//     that adds the 16 warps' parts in warp order into Yt[8 C ..]): deterministic, no atomics, no per-warp copies of y;
//   * the rank-2*NB trailing update of a panel has no pass of its own: it is applied inside the first mat-vec pass of the
//     NEXT panel (MODE 2: DMMA on the fragment read from the ring, live elements stored straight to global memory, the
//     updated fragment then feeds the mat-vec); NB = 8 / 4 / 2 by what the panels leave for the ring;
//   * the producer runs ahead into the next pass while the O(m) sections of a column execute.
// HBM traffic per column: 4 m^2 B read + 4 m^2 / NB B written => (4/3) (1 + 1 / NB) n^3 B per matrix.
#define DTC_CWMAX 16                                // consumer warps: CW = 16 (one CTA of 18 warps per SM) or CW = 8 (10 warps, two
                                                    // CTAs per SM when two matrices fit the shared memory: the mid sizes)
#define DTC_NTHREADS(CW) (32 * ((CW) + 2))          // + the producer warp + the folder warp
#define DTC_PIECE 64                                // tiles per bulk copy (at most; a piece never spans two tile columns); 32 measured
                                                    // 0 / 2 / 3 / 8 % slower at n = 600 / 1000 / 1500 / 2000
#define DTC_NBAR 16                                 // pieces in flight (barrier pairs)

struct DtcGeom {
    static size_t fixed(int n, int nb, int ys) {
        const size_t PS = DtgGeom::ps(n);
        return (size_t)2 * DTC_CWMAX * 64 + ((size_t)2 * nb * PS + (ys + 2) * PS + 2 * 8 + 8) * 8;
    }
};

__device__ __forceinline__ unsigned dtc_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dtc_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dtc_saddr(bar)), "r"(count));
}
__device__ __forceinline__ void dtc_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dtc_bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void dtc_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DTC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DTC_DONE_%=;\n"
        "bra DTC_WAIT_%=;\n"
        "DTC_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680) : "memory");      // suspend-time hint: the warp sleeps in the barrier unit instead of polling
}

struct DtcSh {
    double *ring, *tpart, *Vs, *Ws, *Yd4, *Yt, *Vn, *ts;
    int *szs;                                     // producer: bytes held by the pieces in flight (by barrier index)
    unsigned full, empty, tfull, tempty;          // shared-memory addresses of the barrier arrays (8 B per barrier)
    int ringb;                                    // ring size in bytes
};

// Per-thread pipeline position, carried over the passes of the whole kernel.  The ring is a BYTE ring: a piece takes
// exactly its size (the short pieces of the last tile columns of a pass do not occupy a 16 KB stage each, so that many of
// them are in flight), placed at the running offset, or at 0 when it would not fit before the end of the ring -- the same
// rule on both sides.  g = index of the next piece (barrier g mod 16, phase (g / 16) & 1).
struct DtcPos {
    unsigned g;                                    // consumers: next piece to take; producer: next piece to issue
    int off;                                       // ring offset (bytes) of that piece
    unsigned tc;                                   // tile columns handed over so far (slot = tc & 1, phase from tc >> 1)
    // producer only
    unsigned tail;                                 // oldest piece not yet released
    int infl;                                      // bytes held by pieces tail .. g - 1 (skipped ring ends included)
    int pre, pC, pR0;                              // the current pass was issued ahead up to (pC, pR0)
};

template <int YS>
__device__ __forceinline__ void dtc_row_part(const DtcSh &S, const double2 x, const double2 vc, int R, int lane) {
    if (YS == 4) {
        double *yd = S.Yd4 + 32 * R + lane;
        *yd += fma(x.x, vc.x, x.y * vc.y);
    } else {
        double d = fma(x.x, vc.x, x.y * vc.y);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if ((lane & 3) == 0) S.Yd4[8 * R + (lane >> 2)] += d;
    }
}

// One pass over the tile columns >= T0, pieces = runs of at most DTC_PIECE tiles of one tile column.  MODE 0: y = A v
// (v = vj) and, at the tail, the panel dot products ts[k] = W_k.v, ts[8 + k] = V_k.v (k < jj).
// MODE 2 (first column of a panel): the pending rank-2*NB update of the previous panel, A -= V W^T + W V^T on the elements
// with row, column >= b0 (only those are stored back: the reflectors next to them may be newer in global memory than in a
// piece that was copied ahead), FUSED with y = A v on the updated tile -- the update costs no pass of its own.
// YS = 4: row-part partial sums per lane of the quad (no shuffles); YS = 1: reduced over the quad before the accumulation
// (a quarter of the shared memory: the large sizes).
template <int NB, int YS, int MODE, int CW>
__device__ __forceinline__ void dtc_stream(double *tiles, const DtcSh &S, const double *vj, int n, int NT, int PS, int T0, int j,
                                           int jj, int b0, int nextT0, DtcPos &P, int wp, int lane) {
    if (wp == CW) {
        // ---- producer: a piece is issued as soon as the ring has room for it and a barrier pair is free (pieces are
        // released in order).  After the last piece of this pass it runs AHEAD into the next pass (`nextT0` >= 0: the tiles
        // are not modified in between) as far as that is possible without waiting for a piece of the next pass itself, so
        // that the O(m) sections between two passes overlap the latency of the first copies. ----
        if (lane == 0) {
            const unsigned ring = dtc_saddr(S.ring);
            // returns false when (ahead) the piece would have to wait for a piece of the pass being issued ahead
            auto issue = [&](int &C, int &R0, bool ahead, unsigned first) -> bool {
                while (C < NT) {
                    const int cnt = min(DTC_PIECE, NT - R0), sz = cnt * 512;
                    int off = P.off, skip = 0;
                    if (off + sz > S.ringb) { skip = S.ringb - off; off = 0; }
                    const int need = sz + skip;
                    while (P.infl + need > S.ringb || P.g - P.tail >= (unsigned)DTC_NBAR) {
                        if (ahead && P.tail >= first) return false;
                        dtc_mbar_wait(S.empty + 8 * (P.tail & (DTC_NBAR - 1)), (P.tail >> 4) & 1u);
                        P.infl -= S.szs[P.tail & (DTC_NBAR - 1)];
                        ++P.tail;
                    }
                    S.szs[P.g & (DTC_NBAR - 1)] = need;
                    dtc_bulk_load(ring + (unsigned)off, tiles + dtg_off<true>(R0, C, NT), (unsigned)sz, S.full + 8 * (P.g & (DTC_NBAR - 1)));
                    P.infl += need;
                    P.off = off + sz;
                    ++P.g;
                    R0 += DTC_PIECE;
                    if (R0 >= NT) { ++C; R0 = C; }
                }
                return true;
            };
            int C = P.pre ? P.pC : T0, R0 = P.pre ? P.pR0 : T0;
            issue(C, R0, false, 0u);
            P.pre = 0;
            if (nextT0 >= 0 && nextT0 < NT) {
                P.pC = nextT0; P.pR0 = nextT0; P.pre = 1;
                issue(P.pC, P.pR0, true, P.g);
            }
        }
        __syncwarp();
    } else if (wp == CW + 1) {
        // ---- folder: column parts of the consumer warps, added in warp order ----
        {
            for (int C = T0; C < NT; ++C) {
                const unsigned slot = P.tc & 1u, use = P.tc >> 1;
                dtc_mbar_wait(S.tfull + 8 * slot, use & 1u);
                if (lane < 8) {
                    const double *tp = S.tpart + (size_t)slot * CW * 8 + lane;
                    double a[CW];
#pragma unroll
                    for (int w = 0; w < CW; ++w) a[w] = tp[w * 8];
                    double sum = 0.0;
#pragma unroll
                    for (int w = 0; w < CW; ++w) sum += a[w];
                    S.Yt[8 * C + lane] = sum;
                }
                __syncwarp();
                if (lane == 0) dtc_mbar_arrive(S.tempty + 8 * slot);
                ++P.tc;
            }
        }
    } else {
        const int g8 = lane >> 2, t4 = lane & 3;
        const int lofs = g8 * 8 + 2 * t4;
        for (int C = T0; C < NT; ++C) {
            double vb[NB > 4 ? 2 : 1], wb[NB > 4 ? 2 : 1];
            double tc0 = 0.0, tc1 = 0.0;
            const double2 vc = *reinterpret_cast<const double2 *>(vj + 8 * C + 2 * t4);
            if (MODE == 2) {
                const bool cok = 8 * C + g8 >= b0;
#pragma unroll
                for (int h = 0; h < (NB > 4 ? 2 : 1); ++h) {
                    const bool kok = t4 + 4 * h < NB;
                    vb[h] = cok && kok ? S.Vs[(t4 + 4 * h) * PS + 8 * C + g8] : 0.0;
                    wb[h] = cok && kok ? S.Ws[(t4 + 4 * h) * PS + 8 * C + g8] : 0.0;
                }
            }
            for (int R0 = C; R0 < NT; R0 += DTC_PIECE) {
                const int cnt = min(DTC_PIECE, NT - R0), sz = cnt * 512;
                if (P.off + sz > S.ringb) P.off = 0;
                const unsigned bi = P.g & (DTC_NBAR - 1);
                dtc_mbar_wait(S.full + 8 * bi, (P.g >> 4) & 1u);
                const double *tb = S.ring + (P.off >> 3) + lofs;
#pragma unroll
                for (int u = 0; u < DTC_PIECE / CW; ++u) {
                    const int q = ((wp - R0) & (CW - 1)) + u * CW;             // rows with R mod CW == wp
                    if (q < cnt) {
                        const int R = R0 + q;
                        double2 x = *reinterpret_cast<const double2 *>(tb + q * 64);
                        if (MODE == 2) {
                            const bool rok = 8 * R + g8 >= b0;
#pragma unroll
                            for (int h = 0; h < (NB > 4 ? 2 : 1); ++h) {
                                const bool kok = t4 + 4 * h < NB;
                                const double va = rok && kok ? -S.Vs[(t4 + 4 * h) * PS + 8 * R + g8] : 0.0;
                                const double wa = rok && kok ? -S.Ws[(t4 + 4 * h) * PS + 8 * R + g8] : 0.0;
                                tc_dmma(x.x, x.y, va, wb[h]);
                                tc_dmma(x.x, x.y, wa, vb[h]);
                            }
                            double *dst = tiles + dtg_off<true>(R, C, NT) + lofs;
                            if (rok) {
                                const int c0 = 8 * C + 2 * t4;
                                if (c0 >= b0) *reinterpret_cast<double2 *>(dst) = x;
                                else if (c0 + 1 >= b0) dst[1] = x.y;
                            }
                        }
                        const double vr = R > C ? vj[8 * R + g8] : 0.0;          // the diagonal tile is complete
                        dtc_row_part<YS>(S, x, vc, R, lane);
                        tc0 = fma(x.x, vr, tc0);
                        tc1 = fma(x.y, vr, tc1);
                    }
                }
                __syncwarp();
                if (lane == 0) dtc_mbar_arrive(S.empty + 8 * bi);
                P.off += sz;
                ++P.g;
            }
            {
                // column part of this warp: sum over the rows g8, lanes 0..3 hold columns 2 t4, 2 t4 + 1 (nothing to reduce
                // when the warp had no tile below the diagonal in this column)
                const int Rf = C + ((wp - C) & (CW - 1));
                if (Rf < NT && (Rf > C || Rf + CW < NT)) {
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        tc0 += __shfl_xor_sync(0xffffffffu, tc0, o);
                        tc1 += __shfl_xor_sync(0xffffffffu, tc1, o);
                    }
                }
                const unsigned slot = P.tc & 1u, use = P.tc >> 1;
                if (use >= 1u) dtc_mbar_wait(S.tempty + 8 * slot, (use - 1u) & 1u);
                if (lane < 4) *reinterpret_cast<double2 *>(S.tpart + ((size_t)slot * CW + wp) * 8 + 2 * lane) = make_double2(tc0, tc1);
                __syncwarp();
                if (lane == 0) dtc_mbar_arrive(S.tfull + 8 * slot);
                ++P.tc;
            }
        }
        if (MODE == 2) {
            // a later pass reads these tiles through the async proxy
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        {
            // panel dot products for the corrections of this column (warp kt: column kt of the panel)
            if (wp < jj) {
                double t1 = 0.0, t2 = 0.0;
                for (int i = j + 1 + lane; i < n; i += 32) {
                    const double vi = vj[i];
                    t1 = fma(S.Ws[wp * PS + i], vi, t1);
                    t2 = fma(S.Vs[wp * PS + i], vi, t2);
                }
                t1 = warp_sum(t1); t2 = warp_sum(t2);
                if (lane == 0) { S.ts[wp] = t1; S.ts[8 + wp] = t2; }
            }
        }
    }
}

template <int NW>
__device__ __forceinline__ double dtc_block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NW; ++q) s += red[q];
    return s;
}

// One persistent CTA per matrix (column-major tiles).  Outputs d, e, tau [B][n]; reflectors stay in the tiles.
template <int NB, int YS, int CW, bool GP>
static __global__ void __maxnreg__(96)                 // 576 threads, one CTA per SM (__launch_bounds__(576) makes ptxas aim at 56 registers and spill)
dtc_tridiag_kernel(double *__restrict__ tiles_all, int n, long long B, double *__restrict__ dd_all, double *__restrict__ ee_all,
                   double *__restrict__ tau_all, int ringb, double *__restrict__ gpanels) {
    // GP: the panels V, W (2 NB PS doubles per CTA) live in GLOBAL memory (L2 resident: <= 0.5 MB per CTA) and
    // the shared memory they would take goes to the ring instead -- the large sizes, where NB = 8 panels of 2 x 8 x n doubles
    // would leave no ring at all.  The current reflector then always lives in Vn (shared: the consumers read it per tile)
    // and is copied into its panel column afterwards; head, tail and the MODE 2 operands read the panels through L2.
    constexpr int NTHR = DTC_NTHREADS(CW);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[NTHR / 32];
    __shared__ double s_scal[2];
    __shared__ __align__(8) unsigned long long bars[2 * DTC_NBAR + 4];
    __shared__ int szs[DTC_NBAR];
    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31;
    const int NT = DtgGeom::nt(n), PS = DtgGeom::ps(n);
    DtcSh S;
    S.ringb = ringb;
    S.ring = (double *)smem;
    S.tpart = S.ring + (size_t)(ringb >> 3);
    constexpr bool gp = GP;                   // compile time: with a run-time choice every panel access becomes a generic load (+10 %)
    S.Vs = gp ? gpanels + (size_t)blockIdx.x * 2 * NB * PS : S.tpart + (size_t)2 * DTC_CWMAX * 8;
    S.Ws = S.Vs + NB * PS;
    S.Yd4 = gp ? S.tpart + (size_t)2 * DTC_CWMAX * 8 : S.Ws + NB * PS;
    S.Yt = S.Yd4 + YS * PS;
    S.Vn = S.Yt + PS;
    S.ts = S.Vn + PS;
    S.szs = szs;
    S.full = dtc_saddr(bars);
    S.empty = S.full + 8 * DTC_NBAR;
    S.tfull = S.empty + 8 * DTC_NBAR;
    S.tempty = S.tfull + 16;
    double *Vs = S.Vs, *Ws = S.Ws;
    if (tid == 0) {
        for (int s = 0; s < DTC_NBAR; ++s) { dtc_mbar_init(&bars[s], 1); dtc_mbar_init(&bars[DTC_NBAR + s], CW); }
        for (int s = 0; s < 2; ++s) { dtc_mbar_init(&bars[2 * DTC_NBAR + s], CW); dtc_mbar_init(&bars[2 * DTC_NBAR + 2 + s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    DtcPos P;
    P.g = 0u; P.off = 0; P.tc = 0u; P.tail = 0u; P.infl = 0; P.pre = 0; P.pC = 0; P.pR0 = 0;
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        double *tiles = tiles_all + (size_t)b * DtgGeom::tile_doubles(n);
        double *dd = dd_all + (size_t)b * n, *ee = ee_all + (size_t)b * n, *tauo = tau_all + (size_t)b * n;
        for (int q = tid; q < 2 * NB * PS; q += NTHR) Vs[q] = 0.0;
        for (int q = tid; q < (YS + 2) * PS; q += NTHR) S.Yd4[q] = 0.0;             // Yd4, Yt, Vn
        __syncthreads();
        for (int j0 = 0; j0 < n; j0 += NB) {
            const int pn = min(NB, n - j0);
            for (int jj = 0; jj < pn; ++jj) {
                const int j = j0 + jj, Cj = j >> 3, cj = j & 7;
                // First column of a panel after the first: the panel arrays still hold the PREVIOUS panel, whose rank-2*NB
                // update is pending; the column is made current from all NB of its columns, its reflector goes to Vn, and the
                // update is applied to the rest of the trailing block inside this column's mat-vec pass (MODE 2).
                const bool fused = jj == 0 && j0 > 0;
                const int nc = fused ? NB : jj;
                const bool in_vn = fused || gp;
                double *vj = in_vn ? S.Vn : Vs + jj * PS;
                // ---- column j made current (stale column + the corrections of the panel) ----
                double xn = 0.0;
                if (in_vn) for (int r = tid; r < j; r += NTHR) vj[r] = 0.0;
                for (int r = j + tid; r < n; r += NTHR) {
                    double cv = tiles[dtg_off<true>(r >> 3, Cj, NT) + (r & 7) * 8 + cj];
                    for (int k = 0; k < nc; ++k)
                        cv = fma(-Vs[k * PS + r], Ws[k * PS + j], fma(-Ws[k * PS + r], Vs[k * PS + j], cv));
                    vj[r] = cv;
                    if (r == j) dd[j] = cv;
                    if (r == j + 1) s_scal[0] = cv;
                    if (r >= j + 2) xn = fma(cv, cv, xn);
                }
                if (j == n - 1) {
                    if (tid == 0) { ee[j] = 0.0; tauo[j] = 0.0; }
                    break;
                }
                {
                    // row-part accumulators of the live rows (the column parts are assigned, not accumulated)
                    double2 *z = reinterpret_cast<double2 *>(S.Yd4 + 8 * YS * ((j + 1) >> 3));
                    const int nz = YS * (PS - 8 * ((j + 1) >> 3)) / 2;
                    for (int q = tid; q < nz; q += NTHR) z[q] = make_double2(0.0, 0.0);
                }
                xn = dtc_block_sum<NTHR / 32>(xn, red);
                double tau, beta, sc;
                s1t_reflector_scalars(s_scal[0], xn, tau, beta, sc);
                for (int r = j + tid; r < n; r += NTHR) {
                    const double v = r == j ? 0.0 : (r == j + 1 ? 1.0 : vj[r] * sc);
                    vj[r] = v;
                    if (r > j) tiles[dtg_off<true>(r >> 3, Cj, NT) + (r & 7) * 8 + cj] = v;       // kept for the back-transformation
                }
                if (tid == 0) { ee[j] = beta; tauo[j] = tau; }
                __syncthreads();
                if (fused) {
                    // (the producer does not run ahead out of this pass: it rewrites the tiles)
                    dtc_stream<NB, YS, 2, CW>(tiles, S, vj, n, NT, PS, (j + 1) >> 3, j, 0, j + 1, -1, P, wp, lane);
                    __syncthreads();
                    for (int q = tid; q < 2 * NB * PS; q += NTHR) Vs[q] = 0.0;
                    __syncthreads();
                } else {
                    // the producer may run ahead into the next column's pass (plain or fused: same tiles, unchanged until then)
                    const int nextT0 = j + 1 < n - 1 ? (j + 2) >> 3 : -1;
                    dtc_stream<NB, YS, 0, CW>(tiles, S, vj, n, NT, PS, (j + 1) >> 3, j, jj, 0, nextT0, P, wp, lane);
                    __syncthreads();
                }
                // ---- corrections, w ----
                double dt = 0.0;
                for (int r = j + 1 + tid; r < n; r += NTHR) {
                    double y;
                    if (YS == 4) {
                        const double2 ya = *reinterpret_cast<const double2 *>(S.Yd4 + 4 * r), yb = *reinterpret_cast<const double2 *>(S.Yd4 + 4 * r + 2);
                        y = ((ya.x + ya.y) + (yb.x + yb.y)) + S.Yt[r];
                    } else y = S.Yd4[r] + S.Yt[r];
                    for (int k = 0; k < jj; ++k) y = fma(-Vs[k * PS + r], S.ts[k], fma(-Ws[k * PS + r], S.ts[8 + k], y));
                    const double pr = tau * y;
                    S.Yd4[YS * r] = pr;
                    dt = fma(pr, vj[r], dt);
                }
                dt = dtc_block_sum<NTHR / 32>(dt, red);
                const double hc = -0.5 * tau * dt;
                for (int r = tid; r < PS; r += NTHR) {
                    Ws[jj * PS + r] = (r > j && r < n) ? fma(hc, vj[r], S.Yd4[YS * r]) : 0.0;
                    if (in_vn) Vs[jj * PS + r] = vj[r];                        // the reflector becomes column jj of the panel
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}

}  // namespace scqed
