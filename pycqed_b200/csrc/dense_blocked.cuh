// dense_blocked.cuh -- S2b: blocked Householder tridiagonalisation (LAPACK zlatrd / zhetrd form)
// for batches of HBM-resident Hermitian matrices, with the trailing rank-2k update on the FP64
// tensor cores (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind on sm_100a).
//
// Replaces the LAPACK zheevr reduction behind pyscqed/util.py:153 (scipy.linalg.eigh) for a
// whole batch of sweep points.  Storage: A[b] n x n complex, column-major, both triangles kept
// consistent (the Hermitian mat-vec streams contiguous columns).
//
// Per column c of a panel (two fully parallel launches, batched over B matrices):
//   blk_reflect_hemv : finish reflector c (norm partials -> tau, v), partial p = A22 v per
//                      (row block, column segment), partial dots W^H v, V^H v, v^H A v
//                      HBM/L2 bound: 16 m^2 B per column  => (16/3) n^3 B per matrix
//   blk_w_nextcol    : w = tau (A v - V (W^H v) - W (V^H v)) - 1/2 tau (w^H v) v, then the
//                      next column a = A[:,c+1] - V W[c+1,:]^H - W V[c+1,:]^H and its norm partials
// Per panel (nb = 32 columns):
//   blk_her2k_dmma   : A22 -= V W^H + W V^H   -- the only dense contraction, (16/3) n^3 flops
//                      per matrix in total, on DMMA
// Back-transformation (compact WY, per panel):  X -= V T (V^H X)   with T from the stored
// Gram columns V^H v (no extra pass over V).
#pragma once
#include "common.cuh"

namespace scqed {

#define BLK_NB 32
#define BLK_ROWS 128

struct BlkBufs {
    cplx *A;            // [B][n*n]
    cplx *Vp, *Wp;      // [B][n*NB]
    cplx *acol;         // [B][n]
    double *normpart;   // [B][NRB]
    cplx *part;         // [B][nseg][n]   row sums   sum_{c in seg, c <= r} A[r][c] v_c
    cplx *cpart;        // [B][NRB][n]    column sums sum_{r in block, r > c} conj(A[r][c]) v_r
    cplx *s1part, *s2part;   // [B][NRB][NB]
    double *vAvpart;    // [B][NRB*nseg]
    cplx *G;            // [B][n][NB]   G[c][t] = V[:,t]^H v_c  (t < i)
    cplx *Tbuf;         // [B][npanels][NB*NB]
    double *dd, *ee;    // [B][n]
    cplx *tau;          // [B][n]
    int n, nseg, nrb, npanels;
};

__device__ __forceinline__ double blk_block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = l < nw ? red[l] : 0.0;
    return warp_sum(s);
}

// Householder scalars of column c from alpha = a[c+1] and xn = ||a[c+2:]||^2 (zlarfg)
__device__ __forceinline__ void blk_reflector(cplx alpha, double xn, cplx &tj, double &beta, cplx &sc, bool &zero) {
    if (xn == 0.0 && alpha.y == 0.0) { tj = cmake(0.0, 0.0); beta = alpha.x; }
    else {
        beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn), alpha.x);
        tj = cmake((beta - alpha.x) / beta, -alpha.y / beta);
    }
    cplx den = cmake(alpha.x - beta, alpha.y);
    double dn = cabs2(den);
    zero = (tj.x == 0.0 && tj.y == 0.0);
    sc = zero ? cmake(0.0, 0.0) : cmake(den.x / dn, -den.y / dn);
}

// grid (NRB, nseg, B), block BLK_ROWS.  Only the LOWER triangle of the trailing matrix is read:
// an element A[r][c] (r > c) serves y_r += A[r][c] v_c in the owning thread and
// y_c += conj(A[r][c]) v_r through a warp shuffle reduction, so the kernel moves 8 m^2 bytes per
// column instead of 16 m^2.
static __global__ void __launch_bounds__(BLK_ROWS)
blk_reflect_hemv(BlkBufs K, int c, int i) {
    extern __shared__ cplx s_dyn[];               // s_v[seg] | colacc[4][seg]
    __shared__ double red[32];
    const int n = K.n, b = blockIdx.z, rb = blockIdx.x, sg = blockIdx.y, tid = threadIdx.x;
    const int m = n - c - 1;
    if ((rb + 1) * BLK_ROWS <= c + 1) return;     // block entirely above the trailing matrix
    const cplx *acol = K.acol + (size_t)b * n;
    double xn = 0.0;
    for (int q = 0; q < K.nrb; ++q) xn += K.normpart[(size_t)b * K.nrb + q];
    const cplx alpha = acol[c + 1];
    cplx tj, sc; double beta; bool zero;
    blk_reflector(alpha, xn, tj, beta, sc, zero);
    const int seg = (m + K.nseg - 1) / K.nseg;
    const int segmax = (n + K.nseg - 1) / K.nseg;
    const int c0 = sg * seg, c1 = min(m, c0 + seg);
    cplx *s_v = s_dyn, *colacc = s_dyn + segmax;
    for (int cc = c0 + tid; cc < c1; cc += blockDim.x)
        s_v[cc - c0] = cc == 0 ? cmake(1.0, 0.0) : cmul(acol[c + 1 + cc], sc);
    for (int q = tid; q < 4 * segmax; q += blockDim.x) colacc[q] = cmake(0.0, 0.0);
    __syncthreads();
    const int rg = rb * BLK_ROWS + tid;
    const bool active = rg >= c + 1 && rg < n;
    const int rr = rg - (c + 1);                  // relative row
    const int w = tid >> 5, l = tid & 31;
    cplx vr = cmake(0.0, 0.0), acc = cmake(0.0, 0.0);
    if (active) vr = rg == c + 1 ? cmake(1.0, 0.0) : cmul(acol[rg], sc);
    // columns this warp can touch: cc <= largest relative row of the warp
    const int rr_hi = min(m - 1, rb * BLK_ROWS + w * 32 + 31 - (c + 1));
    const int cend = min(c1, rr_hi + 1);
    const cplx *Ar = K.A + (size_t)b * n * n + (active ? rg : (c + 1)) + (size_t)(c + 1) * n;
    cplx *cw = colacc + w * segmax;
    constexpr int U = 8;                            // columns in flight per thread
    for (int cc = c0; cc < cend; cc += U) {
        cplx a[U], t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int cu = cc + u;
            a[u] = (active && cu < cend && cu <= rr) ? __ldg(Ar + (size_t)cu * n) : cmake(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int cu = cc + u;
            t[u] = cmake(0.0, 0.0);
            if (cu == rr) {                        // diagonal: real, counted once
                acc.x = fma(a[u].x, vr.x, acc.x);
                acc.y = fma(a[u].x, vr.y, acc.y);
            } else if (cu < cend) {
                cfma(acc, a[u], s_v[cu - c0]);     // a is zero above the diagonal
                cfmac(t[u], a[u], vr);             // conj(a) v_r -> y_c
            }
        }
        // U independent shuffle chains interleaved (latency hidden by ILP)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                t[u].x += __shfl_xor_sync(0xffffffffu, t[u].x, o);
                t[u].y += __shfl_xor_sync(0xffffffffu, t[u].y, o);
            }
        }
        if (l < U && cc + l < cend) {
            cplx tv = t[0];
#pragma unroll
            for (int u = 1; u < U; ++u) if (l == u) tv = t[u];
            cw[cc + l - c0] = tv;
        }
    }
    if (active) K.part[((size_t)b * K.nseg + sg) * n + rg] = acc;
    __syncthreads();
    // column sums of this row block, and the v^H (A v) partial (real for Hermitian A)
    double pv = active ? (vr.x * acc.x + vr.y * acc.y) : 0.0;
    for (int cc = c0 + tid; cc < c1; cc += blockDim.x) {
        int q = cc - c0;
        cplx t = cadd(cadd(colacc[q], colacc[segmax + q]), cadd(colacc[2 * segmax + q], colacc[3 * segmax + q]));
        K.cpart[((size_t)b * K.nrb + rb) * n + (c + 1 + cc)] = t;
        cplx vc = s_v[q];
        pv += vc.x * t.x + vc.y * t.y;
    }
    pv = blk_block_sum(pv, red);
    if (tid == 0) K.vAvpart[((size_t)b * K.nrb + rb) * K.nseg + sg] = pv;
    if (sg != 0) return;
    // first column segment also finalises the reflector and the panel dot products
    cplx *Vp = K.Vp + (size_t)b * n * BLK_NB, *Wp = K.Wp + (size_t)b * n * BLK_NB;
    if (active) {
        Vp[rg + (size_t)i * n] = vr;
        K.A[(size_t)b * n * n + rg + (size_t)c * n] = vr;      // reflector kept for the back-transform
        if (rg == c + 1) {
            K.dd[(size_t)b * n + c] = acol[c].x;
            K.ee[(size_t)b * n + c] = beta;
            K.tau[(size_t)b * n + c] = tj;
        }
    }
    __shared__ double sred[4][2][BLK_NB][2];
    for (int t = 0; t < i; ++t) {
        cplx d1 = cmake(0.0, 0.0), d2 = cmake(0.0, 0.0);
        if (active) {
            cfmac(d1, Wp[rg + (size_t)t * n], vr);
            cfmac(d2, Vp[rg + (size_t)t * n], vr);
        }
        d1 = warp_sum(d1); d2 = warp_sum(d2);
        if (l == 0) { sred[w][0][t][0] = d1.x; sred[w][0][t][1] = d1.y; sred[w][1][t][0] = d2.x; sred[w][1][t][1] = d2.y; }
    }
    __syncthreads();
    if (tid < i) {
        cplx d1 = cmake(0.0, 0.0), d2 = cmake(0.0, 0.0);
        for (int q = 0; q < 4; ++q) {
            d1.x += sred[q][0][tid][0]; d1.y += sred[q][0][tid][1];
            d2.x += sred[q][1][tid][0]; d2.y += sred[q][1][tid][1];
        }
        K.s1part[((size_t)b * K.nrb + rb) * BLK_NB + tid] = d1;
        K.s2part[((size_t)b * K.nrb + rb) * BLK_NB + tid] = d2;
    }
}

// grid (NRB, B), block BLK_ROWS.  i = -1: only load column 0 (start of the reduction).
static __global__ void __launch_bounds__(BLK_ROWS)
blk_w_nextcol(BlkBufs K, int c, int i) {
    __shared__ cplx s1[BLK_NB], s2[BLK_NB], vrow[BLK_NB + 1], wrow[BLK_NB + 1];
    __shared__ double red[32];
    __shared__ double s_a2;
    const int n = K.n, b = blockIdx.y, rb = blockIdx.x, tid = threadIdx.x;
    const int rg = rb * BLK_ROWS + tid;
    cplx *Vp = K.Vp + (size_t)b * n * BLK_NB, *Wp = K.Wp + (size_t)b * n * BLK_NB;
    const int cn = c + 1;                          // next column
    const int rb0 = (c + 1) / BLK_ROWS;            // first row block that holds trailing rows
    cplx wr = cmake(0.0, 0.0);
    cplx tj = cmake(0.0, 0.0);
    if (i >= 0) {
        tj = K.tau[(size_t)b * n + c];
        if (tid < i) {
            cplx a = cmake(0.0, 0.0), d = cmake(0.0, 0.0);
            for (int q = rb0; q < K.nrb; ++q) {
                a = cadd(a, K.s1part[((size_t)b * K.nrb + q) * BLK_NB + tid]);
                d = cadd(d, K.s2part[((size_t)b * K.nrb + q) * BLK_NB + tid]);
            }
            s1[tid] = a; s2[tid] = d;
            if (rb == rb0) K.G[((size_t)b * n + c) * BLK_NB + tid] = d;
            vrow[tid] = Vp[cn + (size_t)tid * n];
            wrow[tid] = Wp[cn + (size_t)tid * n];
        }
        double vav = 0.0;
        for (int q = rb0 * K.nseg + tid; q < K.nrb * K.nseg; q += blockDim.x) vav += K.vAvpart[(size_t)b * K.nrb * K.nseg + q];
        vav = blk_block_sum(vav, red);             // (syncs: s1/s2 visible afterwards)
        if (tid == 0) {
            double cross = 0.0;
            for (int t = 0; t < i; ++t) cross += s1[t].x * s2[t].x + s1[t].y * s2[t].y;   // Re(conj(s1) s2)
            double X = vav - 2.0 * cross;
            s_a2 = -0.5 * (tj.x * tj.x + tj.y * tj.y) * X;
        }
        __syncthreads();
        const double a2 = s_a2;
        // w for own row, and (every block, redundantly) for row cn
        auto w_finish = [&](int r, cplx av, cplx corr) -> cplx {
            cplx v = r == cn ? cmake(1.0, 0.0) : Vp[r + (size_t)i * n];
            cplx p = cmul(tj, csub(av, corr));
            return cmake(p.x + a2 * v.x, p.y + a2 * v.y);
        };
        if (rg >= cn && rg < n) {
            cplx av = cmake(0.0, 0.0);
            for (int sgm = 0; sgm < K.nseg; ++sgm) av = cadd(av, K.part[((size_t)b * K.nseg + sgm) * n + rg]);
            for (int q = rb0; q < K.nrb; ++q) av = cadd(av, K.cpart[((size_t)b * K.nrb + q) * n + rg]);
            cplx corr = cmake(0.0, 0.0);
            for (int t = 0; t < i; ++t) {
                cfma(corr, Vp[rg + (size_t)t * n], s1[t]);
                cfma(corr, Wp[rg + (size_t)t * n], s2[t]);
            }
            wr = w_finish(rg, av, corr);
            Wp[rg + (size_t)i * n] = wr;
        }
        {   // row cn, cooperatively: threads stride over the partial-sum slots and the panel columns
            cplx av = cmake(0.0, 0.0), corr = cmake(0.0, 0.0);
            for (int q = tid; q < K.nseg; q += blockDim.x) av = cadd(av, K.part[((size_t)b * K.nseg + q) * n + cn]);
            for (int q = rb0 + tid; q < K.nrb; q += blockDim.x) av = cadd(av, K.cpart[((size_t)b * K.nrb + q) * n + cn]);
            if (tid < i) { cfma(corr, vrow[tid], s1[tid]); cfma(corr, wrow[tid], s2[tid]); }
            double a_x = blk_block_sum(av.x, red), a_y = blk_block_sum(av.y, red);
            double c_x = blk_block_sum(corr.x, red), c_y = blk_block_sum(corr.y, red);
            if (tid == 0) {
                vrow[i] = cmake(1.0, 0.0);
                wrow[i] = w_finish(cn, cmake(a_x, a_y), cmake(c_x, c_y));
            }
        }
        __syncthreads();
    }
    // ---- next column: a = A[:, cn] - V W[cn,:]^H - W V[cn,:]^H  (t <= i) -------------------------
    double nrm = 0.0;
    if (cn <= n - 1 && rg >= cn && rg < n) {
        cplx a = K.A[(size_t)b * n * n + rg + (size_t)cn * n];
        for (int t = 0; t <= i; ++t) {
            cplx vt = Vp[rg + (size_t)t * n];
            cplx wt = t == i ? wr : Wp[rg + (size_t)t * n];
            cfnmac(a, vt, wrow[t]);
            cfnmac(a, wt, vrow[t]);
        }
        K.acol[(size_t)b * n + rg] = a;
        if (rg >= cn + 2) nrm = cabs2(a);
        if (cn == n - 1) { K.dd[(size_t)b * n + n - 1] = a.x; K.ee[(size_t)b * n + n - 1] = 0.0; }
    }
    nrm = blk_block_sum(nrm, red);
    if (tid == 0) K.normpart[(size_t)b * K.nrb + rb] = nrm;
}

// ---- trailing update on the FP64 tensor cores ---------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#define H2K_TILE 64
#define H2K_TK 8
#define H2K_LD 12          // row stride (doubles): 2*12 = 24 words -> rows 0..3 hit distinct 8-word groups
// grid (tiles, tiles, B), 256 threads.  C[j1+.., j1+..] -= V W^H + W V^H over panel columns [0, nbp)
static __global__ void __launch_bounds__(256)
blk_her2k_dmma(BlkBufs K, int j1, int nbp) {
    __shared__ double S[8][H2K_TILE][H2K_LD];   // 0..3: row side vr, vi, wr, wi ; 4..7: col side
    if (blockIdx.y > blockIdx.x) return;          // only the lower triangle is ever read again
    const int n = K.n, b = blockIdx.z, tid = threadIdx.x;
    const int row0 = j1 + blockIdx.x * H2K_TILE, col0 = j1 + blockIdx.y * H2K_TILE;
    const cplx *Vp = K.Vp + (size_t)b * n * BLK_NB, *Wp = K.Wp + (size_t)b * n * BLK_NB;
    const int warp = tid >> 5, l = tid & 31;
    const int wm = warp & 1, wn = warp >> 1;
    const int ar = l >> 2, ak = l & 3;
    double cre[4][2][2], cim[4][2][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) { cre[mt][nt][0] = cre[mt][nt][1] = 0.0; cim[mt][nt][0] = cim[mt][nt][1] = 0.0; }
    const int lr = tid & 63, lq = tid >> 6;
    for (int t0 = 0; t0 < nbp; t0 += H2K_TK) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int tt = lq + 4 * h, t = t0 + tt;
            int gr = row0 + lr, gc = col0 + lr;
            cplx v = cmake(0.0, 0.0), w = v, v2 = v, w2 = v;
            if (t < nbp) {
                if (gr < n) { v = Vp[gr + (size_t)t * n]; w = Wp[gr + (size_t)t * n]; }
                if (gc < n) { v2 = Vp[gc + (size_t)t * n]; w2 = Wp[gc + (size_t)t * n]; }
            }
            S[0][lr][tt] = v.x; S[1][lr][tt] = v.y; S[2][lr][tt] = w.x; S[3][lr][tt] = w.y;
            S[4][lr][tt] = v2.x; S[5][lr][tt] = v2.y; S[6][lr][tt] = w2.x; S[7][lr][tt] = w2.y;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            double bvr[2], bvi[2], bwr[2], bwi[2];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                int rr = wn * 16 + nt * 8 + ar, kk = ks * 4 + ak;
                bvr[nt] = S[4][rr][kk]; bvi[nt] = S[5][rr][kk]; bwr[nt] = S[6][rr][kk]; bwi[nt] = S[7][rr][kk];
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                int rr = wm * 32 + mt * 8 + ar, kk = ks * 4 + ak;
                double avr = S[0][rr][kk], avi = S[1][rr][kk], awr = S[2][rr][kk], awi = S[3][rr][kk];
                double navr = -avr, nawr = -awr;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    // Re: vr wr' + vi wi' + wr vr' + wi vi'
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], avr, bwr[nt]);
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], avi, bwi[nt]);
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], awr, bvr[nt]);
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], awi, bvi[nt]);
                    // Im: vi wr' - vr wi' + wi vr' - wr vi'
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], avi, bwr[nt]);
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], navr, bwi[nt]);
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], awi, bvr[nt]);
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], nawr, bvi[nt]);
                }
            }
        }
        __syncthreads();
    }
    cplx *Ab = K.A + (size_t)b * n * n;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        int gr = row0 + wm * 32 + mt * 8 + ar;
        if (gr >= n) continue;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int gc = col0 + wn * 16 + nt * 8 + 2 * ak + h;
                if (gc >= n) continue;
                cplx z = Ab[gr + (size_t)gc * n];
                z.x -= cre[mt][nt][h];
                z.y -= cim[mt][nt][h];
                Ab[gr + (size_t)gc * n] = z;
            }
    }
}

// ---- compact-WY back-transformation ---------------------------------------------------------------
// T of every panel: T[0:i, i] = -tau_i T[0:i,0:i] G[0:i, i], T[i,i] = tau_i (zlarft, forward columnwise)
// grid (npanels, B), block 32 (one warp).
static __global__ void __launch_bounds__(32)
blk_build_T(BlkBufs K) {
    __shared__ cplx T[BLK_NB][BLK_NB + 1];
    const int n = K.n, p = blockIdx.x, b = blockIdx.y, l = threadIdx.x;
    const int j0 = p * BLK_NB, nbp = min(BLK_NB, (n - 1) - j0);
    for (int q = l; q < BLK_NB * (BLK_NB + 1); q += 32) (&T[0][0])[q] = cmake(0.0, 0.0);
    __syncwarp();
    for (int i = 0; i < nbp; ++i) {
        const cplx ti = K.tau[(size_t)b * n + j0 + i];
        const cplx *g = K.G + ((size_t)b * n + j0 + i) * BLK_NB;
        cplx acc = cmake(0.0, 0.0);
        if (l < i) {
            for (int t = l; t < i; ++t) cfma(acc, T[l][t], g[t]);      // row l of upper-triangular T
            acc = cmul(cmake(-ti.x, -ti.y), acc);
        }
        __syncwarp();
        if (l < i) T[l][i] = acc;
        if (l == i) T[i][i] = ti;
        __syncwarp();
    }
    cplx *out = K.Tbuf + ((size_t)b * K.npanels + p) * BLK_NB * BLK_NB;
    for (int q = l; q < BLK_NB * BLK_NB; q += 32) out[q] = T[q / BLK_NB][q % BLK_NB];
}

// Y partial = V_p^H X over this block's rows.  grid (NRB, B), block 128. Ypart [B][NRB][NB][k]
#define BLK_KMAX 64
static __global__ void __launch_bounds__(BLK_ROWS)
blk_bt_dots(BlkBufs K, int p, int k, const cplx *__restrict__ X, cplx *__restrict__ Ypart) {
    __shared__ double sred[4][BLK_KMAX][2];
    const int n = K.n, b = blockIdx.y, rb = blockIdx.x, tid = threadIdx.x;
    const int j0 = p * BLK_NB, nbp = min(BLK_NB, (n - 1) - j0);
    const int rg = rb * BLK_ROWS + tid;
    const int w = tid >> 5, l = tid & 31;
    cplx *yp = Ypart + ((size_t)b * K.nrb + rb) * BLK_NB * k;
    if ((rb + 1) * BLK_ROWS <= j0 + 1) return;
    const cplx *Ab = K.A + (size_t)b * n * n;
    const cplx *Xb = X + (size_t)b * k * n;
    for (int t = 0; t < nbp; ++t) {
        const int c = j0 + t;
        const bool act = rg > c && rg < n;
        cplx v = act ? Ab[rg + (size_t)c * n] : cmake(0.0, 0.0);
        for (int jv = 0; jv < k; ++jv) {
            cplx d = cmake(0.0, 0.0);
            if (act) cfmac(d, v, Xb[(size_t)jv * n + rg]);
            d = warp_sum(d);
            if (l == 0) { sred[w][jv][0] = d.x; sred[w][jv][1] = d.y; }
        }
        __syncthreads();
        if (tid < k)
            yp[t * k + tid] = cmake(sred[0][tid][0] + sred[1][tid][0] + sred[2][tid][0] + sred[3][tid][0],
                                    sred[0][tid][1] + sred[1][tid][1] + sred[2][tid][1] + sred[3][tid][1]);
        __syncthreads();
    }
}

// X_rows -= V_p (T_p Y).  grid (NRB, B), block 128
static __global__ void __launch_bounds__(BLK_ROWS)
blk_bt_apply(BlkBufs K, int p, int k, cplx *__restrict__ X, const cplx *__restrict__ Ypart) {
    extern __shared__ cplx sm[];          // Y [NB][k], TY [NB][k]
    const int n = K.n, b = blockIdx.y, rb = blockIdx.x, tid = threadIdx.x;
    const int j0 = p * BLK_NB, nbp = min(BLK_NB, (n - 1) - j0);
    if ((rb + 1) * BLK_ROWS <= j0 + 1) return;
    cplx *Y = sm, *TY = sm + BLK_NB * k;
    const int rb0 = (j0 + 1) / BLK_ROWS;
    for (int q = tid; q < nbp * k; q += blockDim.x) {
        cplx a = cmake(0.0, 0.0);
        for (int r = rb0; r < K.nrb; ++r) a = cadd(a, Ypart[((size_t)b * K.nrb + r) * BLK_NB * k + q]);
        Y[q] = a;
    }
    __syncthreads();
    const cplx *T = K.Tbuf + ((size_t)b * K.npanels + p) * BLK_NB * BLK_NB;
    for (int q = tid; q < nbp * k; q += blockDim.x) {
        int t = q / k, jv = q - t * k;
        cplx a = cmake(0.0, 0.0);
        for (int u = t; u < nbp; ++u) cfma(a, T[t * BLK_NB + u], Y[u * k + jv]);   // upper triangular
        TY[q] = a;
    }
    __syncthreads();
    const int rg = rb * BLK_ROWS + tid;
    if (rg <= j0 || rg >= n) return;
    const cplx *Ab = K.A + (size_t)b * n * n;
    for (int jv = 0; jv < k; ++jv) {
        cplx x = X[((size_t)b * k + jv) * n + rg];
        for (int t = 0; t < nbp; ++t) {
            int c = j0 + t;
            if (rg > c) {
                cplx v = Ab[rg + (size_t)c * n], f = TY[t * k + jv];
                x.x -= v.x * f.x - v.y * f.y;
                x.y -= v.x * f.y + v.y * f.x;
            }
        }
        X[((size_t)b * k + jv) * n + rg] = x;
    }
}

// X[b][jv][:] = Z (real) ; and the final optional conjugation
static __global__ void blk_init_x(const double *__restrict__ Z, cplx *__restrict__ X, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) X[i] = cmake(Z[i], 0.0);
}
static __global__ void blk_conj_x(cplx *__restrict__ X, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) X[i].y = -X[i].y;
}

}  // namespace scqed
