// scqed.cu -- C ABI of the sweep engine (see include/scqed.h for the contract and the reference
// functions each entry point replaces).  Host-side orchestration only; kernels live in the
// .cuh files next to this one.
#include "../../include/scqed.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"
#include "coef.cuh"
#include "assemble.cuh"
#include "tridiag_eig.cuh"
#include "resonator.cuh"
#include "small_eigh.cuh"
#include "small_split.cuh"
#include "small_packed.cuh"
#include "dense_eigh.cuh"
#include "dense_blocked.cuh"
#include "matfree.cuh"
#include "matfree_d2.cuh"
#include "probe.cuh"
#include "pauli.cuh"

using namespace scqed;

// ------------------------------------------------------------------------------------------------
// error handling / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SCQED_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// Optional per-kernel timing (bench.py's roofline): CUDA events recorded on the launching stream
// around the dominant kernel of each path; read back (and reset) with scqed_timing_read().
static std::atomic<int> g_timing{0};
struct TimedSpan { cudaEvent_t e0, e1; int tag; };
static std::vector<TimedSpan> g_spans;
static void timing_begin(cudaStream_t st, int tag) {
    if (!g_timing.load()) return;
    TimedSpan sp; sp.tag = tag;
    if (cudaEventCreate(&sp.e0) != cudaSuccess || cudaEventCreate(&sp.e1) != cudaSuccess) return;
    cudaEventRecord(sp.e0, st);
    g_spans.push_back(sp);
}
static void timing_end(cudaStream_t st) {
    if (!g_timing.load() || g_spans.empty()) return;
    cudaEventRecord(g_spans.back().e1, st);
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
            want = bytes;
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct scqed_plan {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    PlanDev P;
    int n_coef_scalar = 0;
    long long mf_work_per_row = 0;     // complex MACs per output element of the matrix-free mat-vec
    int s1_real_mode = -1;             // fused small path: -1 unknown, 0 complex Hermitian, 1 real symmetric
    DevBuf s1work, s1vec, auxc, resvec, dcoef;
    std::vector<int> xform_node;
    std::vector<size_t> xform_off;     // byte offsets of W_i^T inside `tables`
    int pair_max_index = -1;
    DevBuf mfwarm;
    bool d2_ok = false;                // two-node plan with dense single-node operators: combined mode products
    DevBuf d2h;
    std::vector<int> factor_node;
    // device copies of the tables
    DevBuf tables;
    // per-call workspaces (grown on demand, reused across calls)
    DevBuf regs, coef, scal, work, in_params, out_evals, out_evecs, out_post, out_info, out_scal, out_pack;
    cudaStream_t own_stream = nullptr, copy_stream = nullptr;
};

static int set_device(int device) {
    CK(cudaSetDevice(device));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
extern "C" int scqed_plan_create(const scqed_plan_desc *d, int device, scqed_plan **out) {
    if (!d || !out) return fail(SCQED_EINVAL, "null argument");
    if (d->n_nodes < 1 || d->n_nodes > SCQED_MAX_NODES) return fail(SCQED_EINVAL, "n_nodes=%d outside 1..%d", d->n_nodes, SCQED_MAX_NODES);
    if (d->n_terms < 1) return fail(SCQED_EINVAL, "plan has no terms");
    long long n = 1;
    for (int k = 0; k < d->n_nodes; ++k) {
        if (d->dims[k] < 1) return fail(SCQED_EINVAL, "dims[%d]=%d", k, d->dims[k]);
        n *= d->dims[k];
        if (n > (1LL << 30)) return fail(SCQED_EINVAL, "Hilbert dimension overflows");
    }
    for (int f = 0; f < d->n_factors; ++f)
        if (d->factor_node[f] < 0 || d->factor_node[f] >= d->n_nodes) return fail(SCQED_EINVAL, "factor %d: bad node", f);
    for (int t = 0; t < d->n_terms; ++t)
        for (int k = 0; k < d->n_nodes; ++k) {
            int f = d->term_factors[t * d->n_nodes + k];
            if (f >= d->n_factors) return fail(SCQED_EINVAL, "term %d: factor id %d out of range", t, f);
            if (f >= 0 && d->factor_node[f] != k) return fail(SCQED_EINVAL, "term %d: factor %d is not on node %d", t, f, k);
        }
    for (int i = 0; i < d->n_instr; ++i) {
        int op = d->instr[3 * i], a = d->instr[3 * i + 1], b = d->instr[3 * i + 2];
        if (op < 0 || op > 19) return fail(SCQED_EINVAL, "instr %d: unknown opcode %d", i, op);
        if (op == OP_CONST && (a < 0 || a >= d->n_consts)) return fail(SCQED_EINVAL, "instr %d: const index", i);
        if (op == OP_LOAD && (a < 0 || a >= d->n_swept)) return fail(SCQED_EINVAL, "instr %d: param index", i);
        if (op >= OP_ADD && (a < 0 || a >= i)) return fail(SCQED_EINVAL, "instr %d: operand a not SSA", i);
        if (((op >= OP_ADD && op <= OP_DIV) || op == OP_POW) && (b < 0 || b >= i)) return fail(SCQED_EINVAL, "instr %d: operand b not SSA", i);
    }
    for (int t = 0; t < d->n_terms; ++t)
        if (d->coef_re[t] >= d->n_instr || d->coef_im[t] >= d->n_instr) return fail(SCQED_EINVAL, "coef register out of range");
    for (int s = 0; s < d->n_scalar; ++s)
        if (d->scalar_regs[s] < 0 || d->scalar_regs[s] >= d->n_instr) return fail(SCQED_EINVAL, "scalar register out of range");

    if (set_device(device)) return SCQED_ECUDA;
    scqed_plan *pl = new scqed_plan();
    pl->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete pl; return fail(SCQED_ECUDA, "cudaGetDeviceProperties failed"); }
    pl->sm_count = prop.multiProcessorCount;
    pl->smem_optin = prop.sharedMemPerBlockOptin;
    pl->factor_node.assign(d->factor_node, d->factor_node + d->n_factors);
    for (int q = 0; q < d->n_pairs; ++q) { int m = d->pairs[4 * q + 2] > d->pairs[4 * q + 3] ? d->pairs[4 * q + 2] : d->pairs[4 * q + 3]; if (m > pl->pair_max_index) pl->pair_max_index = m; }

    // pack all tables into one allocation
    size_t fac_elems = 0;
    for (int f = 0; f < d->n_factors; ++f) {
        size_t dk = d->dims[d->factor_node[f]];
        size_t end = (size_t)d->factor_offset[f] + dk * dk;
        if (end > fac_elems) fac_elems = end;
    }
    std::vector<unsigned char> host;
    auto put = [&](const void *src, size_t bytes) {
        size_t off = (host.size() + 255) & ~(size_t)255;
        host.resize(off + bytes);
        if (bytes) memcpy(host.data() + off, src, bytes);
        return off;
    };
    std::vector<long long> foff(d->factor_offset, d->factor_offset + d->n_factors);
    size_t o_fdata = put(d->factor_data, fac_elems * 16);
    size_t o_foff = put(foff.data(), foff.size() * 8);
    size_t o_terms = put(d->term_factors, (size_t)d->n_terms * d->n_nodes * 4);
    size_t o_instr = put(d->instr, (size_t)d->n_instr * 12);
    size_t o_consts = put(d->consts, (size_t)d->n_consts * 8);
    size_t o_cre = put(d->coef_re, (size_t)d->n_terms * 4);
    size_t o_cim = put(d->coef_im, (size_t)d->n_terms * 4);
    size_t o_sreg = put(d->scalar_regs, (size_t)d->n_scalar * 4);
    // ELL view of the factors + per-term metadata for the matrix-free mat-vec
    std::vector<int> ell_w(d->n_factors > 0 ? d->n_factors : 1, 0);
    std::vector<long long> ell_off(d->n_factors > 0 ? d->n_factors : 1, 0);
    std::vector<int> ell_cols;
    std::vector<double> ell_vals;      // interleaved complex
    std::vector<double> rowsum(d->n_factors > 0 ? d->n_factors : 1, 0.0);
    for (int f = 0; f < d->n_factors; ++f) {
        const int dk = d->dims[d->factor_node[f]];
        const double *F = d->factor_data + 2 * (size_t)d->factor_offset[f];
        int w = 0;
        double rs = 0.0;
        for (int i = 0; i < dk; ++i) {
            int cnt = 0;
            double sum = 0.0;
            for (int j = 0; j < dk; ++j) {
                double re = F[2 * ((size_t)i * dk + j)], im = F[2 * ((size_t)i * dk + j) + 1];
                if (re != 0.0 || im != 0.0) { ++cnt; sum += std::sqrt(re * re + im * im); }
            }
            if (cnt > w) w = cnt;
            if (sum > rs) rs = sum;
        }
        if (w < 1) w = 1;
        ell_w[f] = w; ell_off[f] = (long long)ell_cols.size(); rowsum[f] = rs;
        for (int i = 0; i < dk; ++i) {
            int cnt = 0;
            for (int j = 0; j < dk; ++j) {
                double re = F[2 * ((size_t)i * dk + j)], im = F[2 * ((size_t)i * dk + j) + 1];
                if (re != 0.0 || im != 0.0) { ell_cols.push_back(j); ell_vals.push_back(re); ell_vals.push_back(im); ++cnt; }
            }
            for (; cnt < w; ++cnt) { ell_cols.push_back(i); ell_vals.push_back(0.0); ell_vals.push_back(0.0); }
        }
    }
    if (d->n_aux < 0 || d->n_pairs < 0) return fail(SCQED_EINVAL, "negative aux counts");
    std::vector<int> aux_meta((size_t)(d->n_aux > 0 ? d->n_aux : 1) * 7, 0);
    for (int t = 0; t < d->n_aux; ++t) {
        int nn = 0;
        for (int k = 0; k < d->n_nodes; ++k) {
            int f = d->aux_term_factors[t * d->n_nodes + k];
            if (f >= d->n_factors || (f >= 0 && d->factor_node[f] != k)) return fail(SCQED_EINVAL, "aux term %d: bad factor", t);
            if (f >= 0) {
                if (nn >= 3) return fail(SCQED_EUNSUPPORTED, "aux term %d has more than 3 non-identity factors", t);
                aux_meta[(size_t)t * 7 + 1 + nn] = k; aux_meta[(size_t)t * 7 + 4 + nn] = f; ++nn;
            }
        }
        aux_meta[(size_t)t * 7] = nn;
        if (d->aux_re[t] >= d->n_instr || d->aux_im[t] >= d->n_instr) return fail(SCQED_EINVAL, "aux coefficient register out of range");
    }
    for (int q = 0; q < d->n_pairs; ++q) {
        const int32_t *pr = d->pairs + 4 * q;
        if (pr[0] < 0 || pr[1] > d->n_aux || pr[0] > pr[1] || pr[2] < 0 || pr[3] < 0) return fail(SCQED_EINVAL, "pair %d malformed", q);
    }
    std::vector<int> term_meta((size_t)d->n_terms * 7, 0);
    int max_nn = 0;
    for (int t = 0; t < d->n_terms; ++t) {
        int nn = 0;
        for (int k = 0; k < d->n_nodes; ++k) {
            int f = d->term_factors[t * d->n_nodes + k];
            if (f >= 0) {
                if (nn < 3) { term_meta[(size_t)t * 7 + 1 + nn] = k; term_meta[(size_t)t * 7 + 4 + nn] = f; }
                ++nn;
            }
        }
        term_meta[(size_t)t * 7] = nn;
        if (nn > max_nn) max_nn = nn;
    }
    if (ell_cols.empty()) { ell_cols.push_back(0); ell_vals.push_back(0.0); ell_vals.push_back(0.0); }
    // ---- stencil view: factors with exactly one non-zero diagonal ------------------------------------
    std::vector<int> band_off(d->n_factors > 0 ? d->n_factors : 1, 0), tab_off(d->n_factors > 0 ? d->n_factors : 1, 0);
    std::vector<double> st_val;     // interleaved complex, per factor dk entries: value in row i (column i + off)
    bool stencil_ok = d->n_nodes <= 3;
    for (int f = 0; f < d->n_factors && stencil_ok; ++f) {
        const int dk = d->dims[d->factor_node[f]];
        const double *F = d->factor_data + 2 * (size_t)d->factor_offset[f];
        int off = 0; bool have = false;
        for (int i = 0; i < dk && stencil_ok; ++i)
            for (int j = 0; j < dk; ++j) {
                double re = F[2 * ((size_t)i * dk + j)], im = F[2 * ((size_t)i * dk + j) + 1];
                if (re == 0.0 && im == 0.0) continue;
                if (!have) { off = j - i; have = true; }
                else if (j - i != off) { stencil_ok = false; break; }
            }
        band_off[f] = off;
        tab_off[f] = (int)(st_val.size() / 2);
        for (int i = 0; i < dk; ++i) {
            int j = i + off;
            double re = 0.0, im = 0.0;
            if (j >= 0 && j < dk) { re = F[2 * ((size_t)i * dk + j)]; im = F[2 * ((size_t)i * dk + j) + 1]; }
            st_val.push_back(re); st_val.push_back(im);
        }
    }
    if (st_val.size() / 2 > 2048) stencil_ok = false;       // tables must fit a corner of shared memory
    std::vector<int> st_terms;
    int n_sdiag = 0, n_soff = 0;
    if (stencil_ok) {
        std::vector<long long> strides(d->n_nodes, 1);
        for (int k = d->n_nodes - 2; k >= 0; --k) strides[k] = strides[k + 1] * d->dims[k + 1];
        for (int pass = 0; pass < 2 && stencil_ok; ++pass)
            for (int t = 0; t < d->n_terms; ++t) {
                int nn = 0, ks[3] = {0, 0, 0}, tabs[3] = {0, 0, 0};
                long long delta = 0;
                for (int k = 0; k < d->n_nodes; ++k) {
                    int f = d->term_factors[t * d->n_nodes + k];
                    if (f < 0) continue;
                    if (nn >= 3) { stencil_ok = false; break; }
                    ks[nn] = k; tabs[nn] = tab_off[f]; ++nn;
                    delta += (long long)band_off[f] * strides[k];
                }
                if (!stencil_ok) break;
                if ((delta == 0) != (pass == 0)) continue;
                int rec[9] = {nn, ks[0], ks[1], ks[2], tabs[0], tabs[1], tabs[2], (int)delta, t};
                st_terms.insert(st_terms.end(), rec, rec + 9);
                if (pass == 0) ++n_sdiag; else ++n_soff;
            }
    }
    // pure shift: every factor used by an off-diagonal term has only 0/1 entries
    bool pure_shift = stencil_ok;
    if (stencil_ok) {
        for (size_t q = (size_t)n_sdiag; q < (size_t)(n_sdiag + n_soff) && pure_shift; ++q) {
            const int *r = &st_terms[q * 9];
            for (int u = 0; u < r[0] && pure_shift; ++u) {
                int dk = d->dims[r[1 + u]];
                for (int i = 0; i < dk; ++i) {
                    double re = st_val[2 * (size_t)(r[4 + u] + i)], im = st_val[2 * (size_t)(r[4 + u] + i) + 1];
                    if (!((re == 0.0 && im == 0.0) || (re == 1.0 && im == 0.0))) { pure_shift = false; break; }
                }
            }
        }
    }
    if (st_val.empty()) { st_val.push_back(0.0); st_val.push_back(0.0); }
    if (st_terms.empty()) st_terms.assign(9, 0);
    size_t o_stt = put(st_terms.data(), st_terms.size() * 4);
    size_t o_stv = put(st_val.data(), st_val.size() * 8);
    size_t o_ellw = put(ell_w.data(), ell_w.size() * 4);
    size_t o_elloff = put(ell_off.data(), ell_off.size() * 8);
    size_t o_ellc = put(ell_cols.data(), ell_cols.size() * 4);
    size_t o_ellv = put(ell_vals.data(), ell_vals.size() * 8);
    size_t o_rowsum = put(rowsum.data(), rowsum.size() * 8);
    size_t o_tmeta = put(term_meta.data(), term_meta.size() * 4);
    // two-node plans with a dense single-node operator: coupling terms listed separately (matfree_d2.cuh)
    std::vector<int> d2_meta, d2_cidx, d2_phase;
    int d2_nmulti = 0;
    auto factor_phase = [&](int f) {          // 0: all entries real, 1: all imaginary, 2: mixed
        const int dk = d->dims[d->factor_node[f]];
        const double *fd = d->factor_data + 2 * (size_t)d->factor_offset[f];
        bool any_re = false, any_im = false;
        for (size_t e = 0; e < (size_t)dk * dk; ++e) { any_re |= fd[2 * e] != 0.0; any_im |= fd[2 * e + 1] != 0.0; }
        return any_re && any_im ? 2 : (any_im ? 1 : 0);
    };
    {
        bool dense_single = false, ok = d->n_nodes == 2 && max_nn <= 2 && d->dims[0] <= 256 && d->dims[1] <= 256;
        for (int t = 0; t < d->n_terms && ok; ++t) {
            const int nn = term_meta[(size_t)t * 7];
            if (nn == 1 && ell_w[term_meta[(size_t)t * 7 + 4]] >= 16) dense_single = true;
            if (nn == 2) {
                if ((long long)ell_w[term_meta[(size_t)t * 7 + 4]] * ell_w[term_meta[(size_t)t * 7 + 5]] > 64) ok = false;   // dense coupling: generic path
                for (int q = 0; q < 7; ++q) d2_meta.push_back(term_meta[(size_t)t * 7 + q]);
                d2_cidx.push_back(t);
                {
                    const int pa = factor_phase(term_meta[(size_t)t * 7 + 4]), pb = factor_phase(term_meta[(size_t)t * 7 + 5]);
                    d2_phase.push_back((pa == 2 || pb == 2) ? 2 : ((pa + pb) & 1));      // i * i = -1 is real
                }
                ++d2_nmulti;
            }
        }
        pl->d2_ok = ok && dense_single;
        if (d2_meta.empty()) { d2_meta.assign(7, 0); d2_cidx.assign(1, 0); }
        if (d2_phase.empty()) d2_phase.assign(1, 0);
    }
    size_t o_d2m = put(d2_meta.data(), d2_meta.size() * 4), o_d2c = put(d2_cidx.data(), d2_cidx.size() * 4);
    size_t o_d2p = put(d2_phase.data(), d2_phase.size() * 4);
    size_t o_ameta = put(aux_meta.data(), aux_meta.size() * 4);
    size_t o_are = put(d->aux_re, (size_t)d->n_aux * 4), o_aim = put(d->aux_im, (size_t)d->n_aux * 4);
    size_t o_pairs = put(d->pairs, (size_t)d->n_pairs * 16);
    int n_dense = d->n_dense;
    if (n_dense < 0 || (n_dense > 0 && (!d->dense_tab || !d->dense_re || !d->dense_im || !d->dense_gmax)))
        { delete pl; return fail(SCQED_EINVAL, "dense term table malformed"); }
    for (int g = 0; g < n_dense; ++g)
        if (d->dense_re[g] >= d->n_instr || d->dense_im[g] >= d->n_instr) { delete pl; return fail(SCQED_EINVAL, "dense table register out of range"); }
    if (d->n_dense_diag < 0 || d->n_dense_diag > n_dense) { delete pl; return fail(SCQED_EINVAL, "n_dense_diag"); }
    size_t o_dtab = put(d->dense_tab, ((size_t)(n_dense - d->n_dense_diag) * n * n + (size_t)d->n_dense_diag * n) * 8);
    size_t o_dre = put(d->dense_re, (size_t)n_dense * 4), o_dim = put(d->dense_im, (size_t)n_dense * 4);
    size_t o_dgm = put(d->dense_gmax, (size_t)n_dense * 8);
    size_t o_ekey = 0, o_eptr = 0, o_eterm = 0, o_eval = 0, o_emap = 0;
    if (d->n_entries < 0 || (d->n_entries > 0 && (!d->ent_key || !d->ent_ptr || !d->ent_term || !d->ent_val))) { delete pl; return fail(SCQED_EINVAL, "entry list malformed"); }
    if (d->n_entries > 0) {
        const int ne = d->n_entries;
        const long long np_tot = d->ent_ptr[ne];
        for (int e = 0; e < ne; ++e) {
            if (d->ent_key[e] < 0 || d->ent_key[e] >= n * n || (e > 0 && d->ent_key[e] <= d->ent_key[e - 1]) || d->ent_ptr[e + 1] < d->ent_ptr[e])
                { delete pl; return fail(SCQED_EINVAL, "entry %d malformed", e); }
        }
        for (long long q = 0; q < np_tot; ++q)
            if (d->ent_term[q] < 0 || d->ent_term[q] >= d->n_terms) { delete pl; return fail(SCQED_EINVAL, "entry pair %lld: term out of range", q); }
        o_ekey = put(d->ent_key, (size_t)ne * 8);
        o_eptr = put(d->ent_ptr, (size_t)(ne + 1) * 4);
        o_eterm = put(d->ent_term, (size_t)np_tot * 4);
        o_eval = put(d->ent_val, (size_t)np_tot * 16);
        if (d->ent_dense) {
            std::vector<int> emap((size_t)(n * n), -1);
            for (int e = 0; e < ne; ++e) emap[(size_t)d->ent_key[e]] = e;
            o_emap = put(emap.data(), emap.size() * 4);
        }
    }
    if (d->n_xform < 0 || (d->n_xform > 0 && (!d->xform_node || !d->xform_data))) { delete pl; return fail(SCQED_EINVAL, "basis transforms malformed"); }
    {
        const double *src = d->xform_data;
        for (int i = 0; i < d->n_xform; ++i) {
            const int k = d->xform_node[i];
            if (k < 0 || k >= d->n_nodes) { delete pl; return fail(SCQED_EINVAL, "xform_node[%d]=%d", i, k); }
            const int dk = d->dims[k];
            std::vector<double> wt((size_t)dk * dk);
            for (int a = 0; a < dk; ++a)
                for (int b = 0; b < dk; ++b) wt[(size_t)b * dk + a] = src[(size_t)a * dk + b];     // W^T: thread a reads wt[b*d + a]
            pl->xform_node.push_back(k);
            pl->xform_off.push_back(put(wt.data(), wt.size() * 8));
            src += (size_t)dk * dk;
        }
    }
    if (pl->tables.ensure(host.size() + 256)) { delete pl; return fail(SCQED_ENOMEM, "plan tables"); }
    if (cudaMemcpy(pl->tables.p, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        pl->tables.release(); delete pl; return fail(SCQED_ECUDA, "uploading plan tables failed");
    }
    unsigned char *base = (unsigned char *)pl->tables.p;
    PlanDev &P = pl->P;
    memset(&P, 0, sizeof(P));
    P.n_nodes = d->n_nodes;
    P.n = (int)n;
    int stride = 1;
    for (int k = d->n_nodes - 1; k >= 0; --k) { P.dims[k] = d->dims[k]; P.strides[k] = stride; stride *= d->dims[k]; }
    P.n_factors = d->n_factors;
    P.factor_data = (const cplx *)(base + o_fdata);
    P.factor_offset = (const long long *)(base + o_foff);
    P.n_terms = d->n_terms;
    P.term_factors = (const int *)(base + o_terms);
    P.ell_w = (const int *)(base + o_ellw);
    P.ell_off = (const long long *)(base + o_elloff);
    P.ell_cols = (const int *)(base + o_ellc);
    P.ell_vals = (const cplx *)(base + o_ellv);
    P.fac_rowsum = (const double *)(base + o_rowsum);
    P.term_meta = (const int *)(base + o_tmeta);
    P.n_aux = d->n_aux; P.n_pairs = d->n_pairs;
    P.aux_meta = (const int *)(base + o_ameta);
    P.aux_re = (const int *)(base + o_are); P.aux_im = (const int *)(base + o_aim);
    P.pairs = (const int *)(base + o_pairs);
    P.d2_nmulti = d2_nmulti; P.d2_meta = (const int *)(base + o_d2m); P.d2_cidx = (const int *)(base + o_d2c); P.d2_phase = (const int *)(base + o_d2p);
    P.n_entries = d->n_entries; P.ent_dense = d->n_entries > 0 && d->ent_dense;
    P.ent_key = (const long long *)(base + o_ekey); P.ent_ptr = (const int *)(base + o_eptr);
    P.ent_term = (const int *)(base + o_eterm); P.ent_val = (const cplx *)(base + o_eval);
    P.ent_map = (const int *)(base + o_emap);
    P.n_dense = n_dense; P.n_dense_diag = d->n_dense_diag; P.dense_tridiag = n_dense > 0 && d->dense_tridiag;
    P.dense_tab = (const double *)(base + o_dtab);
    P.dense_re = (const int *)(base + o_dre); P.dense_im = (const int *)(base + o_dim);
    P.dense_gmax = (const double *)(base + o_dgm);
    P.max_term_nn = max_nn;
    P.stencil_ok = stencil_ok ? 1 : 0;
    P.st_pure_shift = pure_shift ? 1 : 0;
    P.n_sdiag = n_sdiag; P.n_soff = n_soff; P.st_val_total = (int)(st_val.size() / 2);
    P.st_terms = (const int *)(base + o_stt);
    P.st_val = (const cplx *)(base + o_stv);
    pl->mf_work_per_row = 0;
    for (int t = 0; t < d->n_terms; ++t) {
        long long wprod = 1;
        for (int k = 0; k < d->n_nodes; ++k) {
            int f = d->term_factors[t * d->n_nodes + k];
            if (f >= 0) wprod *= ell_w[f];
        }
        pl->mf_work_per_row += wprod;
    }
    P.n_swept = d->n_swept; P.n_instr = d->n_instr; P.n_scalar = d->n_scalar;
    P.instr = (const int *)(base + o_instr);
    P.consts = (const double *)(base + o_consts);
    P.coef_re = (const int *)(base + o_cre);
    P.coef_im = (const int *)(base + o_cim);
    P.scalar_regs = (const int *)(base + o_sreg);
    *out = pl;
    return SCQED_OK;
}

extern "C" void scqed_plan_destroy(scqed_plan *pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    DevBuf *bufs[] = {&pl->tables, &pl->regs, &pl->coef, &pl->scal, &pl->work, &pl->s1work, &pl->s1vec, &pl->auxc, &pl->resvec, &pl->dcoef, &pl->d2h, &pl->mfwarm, &pl->in_params, &pl->out_evals,
                      &pl->out_evecs, &pl->out_post, &pl->out_info, &pl->out_scal, &pl->out_pack};
    for (DevBuf *b : bufs) b->release();
    if (pl->own_stream) cudaStreamDestroy(pl->own_stream);
    if (pl->copy_stream) cudaStreamDestroy(pl->copy_stream);
    delete pl;
}

extern "C" int scqed_plan_hilbert_dim(const scqed_plan *pl) { return pl ? pl->P.n : SCQED_EINVAL; }
extern "C" int64_t scqed_launch_count(void) { return g_launches.load(); }
extern "C" void scqed_timing_enable(int on) { g_timing.store(on ? 1 : 0); }
extern "C" int scqed_timing_read(int tag, double *total_ms, int64_t *count) {
    if (!total_ms || !count) return fail(SCQED_EINVAL, "null argument");
    double tot = 0.0; long long cnt = 0;
    std::vector<TimedSpan> keep;
    for (auto &sp : g_spans) {
        if (sp.tag != tag) { keep.push_back(sp); continue; }
        float ms = 0.f;
        if (cudaEventSynchronize(sp.e1) == cudaSuccess && cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { tot += ms; ++cnt; }
        cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1);
    }
    g_spans.swap(keep);
    *total_ms = tot; *count = cnt;
    return 0;
}
extern "C" const char *scqed_last_error(void) { return g_err.c_str(); }
extern "C" const char *scqed_version(void) { return "scqed 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// K1 / K2
// ------------------------------------------------------------------------------------------------
static int run_coef(scqed_plan *pl, const double *params_dev, long long n_pts, cplx *coef_dev, double *scal_dev,
                    cudaStream_t st, cplx *aux_dev = nullptr, bool dense = false) {
    const PlanDev &P = pl->P;
    if (pl->regs.ensure((size_t)(P.n_instr > 0 ? P.n_instr : 1) * n_pts * 8)) return fail(SCQED_ENOMEM, "coefficient registers");
    int threads = 128;
    long long blocks = (n_pts + threads - 1) / threads;
    double *dcoef = nullptr;
    if (dense && P.n_dense > 0) {
        if (pl->dcoef.ensure((size_t)n_pts * 2 * P.n_dense * 8)) return fail(SCQED_ENOMEM, "dense-table coefficients");
        dcoef = (double *)pl->dcoef.p;
    }
    coef_eval_kernel<<<(unsigned)blocks, threads, 0, st>>>(P, params_dev, n_pts, (double *)pl->regs.p, coef_dev, scal_dev, aux_dev, dcoef);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int scqed_matrix_elements(scqed_plan *pl, const double *params_dev, int64_t n_pts, int32_t k,
                                     const double *evecs_dev, double *out_dev, void *stream) {
    if (!pl || !evecs_dev || !out_dev || n_pts < 1 || k < 1) return fail(SCQED_EINVAL, "bad argument");
    const PlanDev &P = pl->P;
    if (P.n_pairs < 1) return fail(SCQED_EINVAL, "the plan has no matrix-element requests");
    if (P.n_swept > 0 && !params_dev) return fail(SCQED_EINVAL, "params_dev is NULL");
    if (pl->pair_max_index >= k) return fail(SCQED_EINVAL, "a requested eigenvector index exceeds k=%d", k);
    CK(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->coef.ensure((size_t)n_pts * P.n_terms * 16)) return fail(SCQED_ENOMEM, "coefficients");
    if (pl->scal.ensure((size_t)n_pts * (P.n_scalar + 1) * 8)) return fail(SCQED_ENOMEM, "scalars");
    if (pl->auxc.ensure((size_t)n_pts * (P.n_aux + 1) * 16)) return fail(SCQED_ENOMEM, "aux coefficients");
    int rc = run_coef(pl, params_dev, n_pts, (cplx *)pl->coef.p, (double *)pl->scal.p, st, (cplx *)pl->auxc.p);
    if (rc) return rc;
    const long long warps = n_pts * P.n_pairs;
    aux_matel_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(P, (const cplx *)pl->auxc.p, (const cplx *)evecs_dev, n_pts, k, (cplx *)out_dev);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// v <- (I x W x I) v on tensor mode `node`, in place: one CTA per (vector, fiber); the fiber of d
// amplitudes is staged in shared memory, thread a forms sum_b W[a][b] x[b] (W^T stored: coalesced).
__global__ void __launch_bounds__(128)
mode_rotate_kernel(cplx *__restrict__ V, long long n_fibers_total, int n, int d, int stride, const double *__restrict__ Wt) {
    extern __shared__ cplx s_x[];
    const int fibers_per_vec = n / d;
    for (long long fb = blockIdx.x; fb < n_fibers_total; fb += gridDim.x) {
        const long long vec = fb / fibers_per_vec;
        const int f = (int)(fb - vec * fibers_per_vec);
        const int left = f / stride, right = f - left * stride;
        cplx *base = V + (size_t)vec * n + (size_t)left * d * stride + right;
        for (int b = threadIdx.x; b < d; b += blockDim.x) s_x[b] = base[(size_t)b * stride];
        __syncthreads();
        for (int a = threadIdx.x; a < d; a += blockDim.x) {
            double re = 0.0, im = 0.0;
            for (int b = 0; b < d; ++b) {
                const double w = __ldg(Wt + (size_t)b * d + a);
                re = fma(w, s_x[b].x, re);
                im = fma(w, s_x[b].y, im);
            }
            base[(size_t)a * stride] = cmake(re, im);
        }
        __syncthreads();
    }
}

extern "C" int scqed_back_transform(scqed_plan *pl, double *evecs_dev, int64_t n_vectors, void *stream) {
    if (!pl || !evecs_dev || n_vectors < 1) return fail(SCQED_EINVAL, "bad argument");
    if (set_device(pl->device)) return SCQED_ECUDA;
    const PlanDev &P = pl->P;
    for (size_t i = 0; i < pl->xform_node.size(); ++i) {
        const int k = pl->xform_node[i], dk = P.dims[k];
        const long long fibers = (long long)n_vectors * (P.n / dk);
        const long long grid = fibers < (long long)pl->sm_count * 16 ? fibers : (long long)pl->sm_count * 16;
        mode_rotate_kernel<<<(unsigned)grid, 128, (size_t)dk * 16, (cudaStream_t)stream>>>(
            (cplx *)evecs_dev, fibers, P.n, dk, P.strides[k], (const double *)((const unsigned char *)pl->tables.p + pl->xform_off[i]));
        LAUNCHED();
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int scqed_pauli_coefficients(int device, const double *e01_dev, const double *op_dev, int64_t n_pts,
                                        double *h_dev, void *stream) {
    if (!e01_dev || !op_dev || !h_dev || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    if (set_device(device)) return SCQED_ECUDA;
    pauli_coef_kernel<<<(unsigned)((n_pts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e01_dev, (const cplx *)op_dev, n_pts, h_dev);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int scqed_eval_coefficients(scqed_plan *pl, const double *params_dev, int64_t n_pts, double *coef_dev,
                                       double *scalars_dev, void *stream) {
    if (!pl || !coef_dev || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    if (pl->P.n_swept > 0 && !params_dev) return fail(SCQED_EINVAL, "params_dev is NULL");
    if (pl->P.n_scalar > 0 && !scalars_dev) return fail(SCQED_EINVAL, "scalars_dev is NULL but the plan has scalar outputs");
    if (set_device(pl->device)) return SCQED_ECUDA;
    return run_coef(pl, params_dev, n_pts, (cplx *)coef_dev, scalars_dev, (cudaStream_t)stream);
}

static int run_assemble(scqed_plan *pl, const cplx *coef_dev, long long n_pts, int tidyup, cplx *H, cudaStream_t st) {
    const PlanDev &P = pl->P;
    if (P.n_entries > 0 && !getenv("SCQED_NO_ENTRY_LIST")) {
        const size_t nn = (size_t)P.n * P.n;
        const long long eb = P.ent_dense ? (long long)((nn + 255) / 256) : (P.n_entries + 255) / 256;
        if (eb <= 65535 && n_pts <= 2147483647LL) {
            timing_begin(st, 3);
            if (P.ent_dense) {
                assemble_mapped_kernel<<<dim3((unsigned)((n_pts + ASM_PG - 1) / ASM_PG), (unsigned)eb), 256, (size_t)ASM_PG * P.n_terms * 16, st>>>(P, coef_dev, n_pts, tidyup, H);
            } else {
                CK(cudaMemsetAsync(H, 0, (size_t)n_pts * nn * 16, st));
                assemble_entries_kernel<<<dim3((unsigned)((n_pts + ASM_PG - 1) / ASM_PG), (unsigned)eb), 256, (size_t)ASM_PG * P.n_terms * 16, st>>>(P, coef_dev, n_pts, tidyup, H);
            }
            timing_end(st);
            LAUNCHED();
            CK(cudaGetLastError());
            return 0;
        }
    }
    long long col_tiles = (P.n + ASM_TC - 1) / ASM_TC, row_tiles = (P.n + ASM_TR - 1) / ASM_TR;
    long long blocks = n_pts * col_tiles * row_tiles;
    if (blocks > 2147483647LL) return fail(SCQED_EINVAL, "assembly grid too large; split the batch");
    assemble_dense_kernel<<<(unsigned)blocks, ASM_TC, (size_t)P.n_terms * 16, st>>>(P, coef_dev, n_pts, tidyup, H);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int scqed_assemble_dense(scqed_plan *pl, const double *params_dev, int64_t n_pts, int32_t tidyup,
                                    double *H_dev, void *stream) {
    if (!pl || !H_dev || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    if (pl->P.n_swept > 0 && !params_dev) return fail(SCQED_EINVAL, "params_dev is NULL");
    if (set_device(pl->device)) return SCQED_ECUDA;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->coef.ensure((size_t)n_pts * pl->P.n_terms * 16)) return fail(SCQED_ENOMEM, "coefficients");
    if (pl->scal.ensure((size_t)n_pts * (pl->P.n_scalar + 1) * 8)) return fail(SCQED_ENOMEM, "scalars");
    int rc = run_coef(pl, params_dev, n_pts, (cplx *)pl->coef.p, (double *)pl->scal.p, st);
    if (rc) return rc;
    // row-major H[i][j]: the kernel's fast index is the column
    return run_assemble(pl, (const cplx *)pl->coef.p, n_pts, tidyup, (cplx *)H_dev, st);
}

// ------------------------------------------------------------------------------------------------
// solvers
// ------------------------------------------------------------------------------------------------
static bool small_fits(const scqed_plan *pl, int n, int k, int want_vec, int n_terms, SmallLayout *Lout) {
    SmallLayout L = small_smem_layout(n, k, want_vec, n_terms, pl->smem_optin);
    if ((size_t)L.total + 1024 > pl->smem_optin) return false;
    if (want_vec && L.vb < 1) return false;
    if (Lout) *Lout = L;
    return true;
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
    static bool attr_set = false;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(small_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_optin - 1024));
        attr_set = true;
    }
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
static int s1_packed_mode() {
    const char *e = getenv("SCQED_S1_PACKED");
    return e ? atoi(e) : 2;
}

#define S1_NOT_REAL 77001       // internal: 128 < n <= 224 only fits the shared-memory path when H is real
// real_hint: -1 unknown, 1 known real, 0 known complex
static bool split_fits(const scqed_plan *pl, int n, int k, int n_terms, int real_hint = -1) {
    if (k > 32 || n < 2) return false;
    if (n <= 32 * SMALL_RQ)
        return S1Smem<cplx>::bytes(n, n_terms, 512) <= pl->smem_optin ||
               (s1_packed_mode() && S1PSmem<cplx>::bytes(n, n_terms, 256) <= pl->smem_optin);
    if (n > 224 || real_hint == 0 || !s1_packed_mode()) return false;
    return S1PSmem<double>::bytes(n, n_terms, 512) <= pl->smem_optin;     // packed lower triangle, real symmetric only
}

template <typename T, int NQ>
static int s1p_launch(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st, bool *launched) {
    const int threads = NQ <= 4 ? 256 : 512;
    const size_t smem = S1PSmem<T>::bytes(a.n, n_terms, threads);
    *launched = false;
    if (smem > pl->smem_optin) return 0;
    CK(cudaFuncSetAttribute(s1p_tridiag_kernel<T, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1p_tridiag_kernel<T, NQ>, threads, smem));
    if (occ < 1) return 0;
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    timing_begin(st, sizeof(T) == sizeof(double) ? 1 : 2);
    s1p_tridiag_kernel<T, NQ><<<(unsigned)grid, threads, smem, st>>>(a);
    timing_end(st);
    LAUNCHED();
    CK(cudaGetLastError());
    *launched = true;
    return 0;
}

template <typename T>
static int s1_launch_tridiag(scqed_plan *pl, const S1Args &a, int n_terms, cudaStream_t st) {
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
    CK(cudaFuncSetAttribute(s1_tridiag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s1_tridiag_kernel<T>, threads, smem));
    if (occ < 1) return fail(SCQED_EUNSUPPORTED, "small tridiagonalisation kernel does not fit (n=%d)", a.n);
    long long grid = (long long)occ * pl->sm_count;
    if (grid > a.n_pts) grid = a.n_pts;
    timing_begin(st, sizeof(T) == sizeof(double) ? 1 : 2);
    s1_tridiag_kernel<T><<<(unsigned)grid, threads, smem, st>>>(a);
    timing_end(st);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
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
    (void)o_cnt;
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
            timing_begin(st, 4);
            s1_direct_tridiag_kernel<<<(unsigned)((np * 32 + 127) / 128), 128, 0, st>>>(a);
            timing_end(st);
            LAUNCHED();
            real_mode = true;
        } else if (pl->s1_real_mode == 1 && !Hin) {
            // known real-symmetric plan: no host round trip; a point that turns out complex is
            // flagged info = -2 by the kernel and the host-level driver re-runs in complex mode
            int rc = s1_launch_tridiag<double>(pl, a, P.n_terms > P.n_dense ? P.n_terms : P.n_dense, st);
            if (rc) return rc;
            real_mode = true;
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
        const bool thread_bisect = bis_mode == 2 || (bis_mode == 0 && np * k >= 4096 && k <= 32);
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
    return 0;
}

static int run_small_any(scqed_plan *pl, const PlanDev &P, const cplx *coef, const cplx *Hin, const double *scal,
                         long long n_pts, int n, int k, int want_vec, int tidyup, const ResonatorArgs &res,
                         double *evals, cplx *evecs, double *post, int *info, cudaStream_t st, bool monolithic) {
    if (!monolithic && split_fits(pl, n, k, P.n_terms > P.n_dense ? P.n_terms : P.n_dense, Hin ? -1 : pl->s1_real_mode))
        return run_small_split(pl, P, coef, Hin, scal, n_pts, n, k, want_vec, tidyup, res, evals, evecs, post, info, st);
    return run_small(pl, P, coef, Hin, scal, n_pts, n, k, want_vec, tidyup, res, evals, evecs, post, info, st);
}

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
        static bool attr_set = false;
        if (!attr_set) {
            CK(cudaFuncSetAttribute(d2_backtransform, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
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

static int pick_dense_batch(int n, long long n_pts, int k, int sm_count) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
    size_t budget = free_b / 2;
    size_t per = (size_t)n * n * 16 + blk_ws_bytes(n, k, 1, 64) + dense_ws_bytes(n, k, 1, 64);
    long long B = (long long)(budget / per);
    // enough matrices in flight to fill the SMs during the skinny tail of the reduction
    // (the per-column kernels are launch / latency bound for small n: n = 256 runs 2.4 k matrices/s with 32 matrices
    // per launch and 21.8 k/s with 2048)
    long long want = n <= 512 ? 4096 : (n <= 1536 ? 512 : (n <= 4096 ? 64 : 8));
    if (const char *e = getenv("SCQED_DENSE_BATCH")) { long long v = atoll(e); if (v >= 1) want = v; }
    if (B > want) B = want;
    if (B > n_pts) B = n_pts;
    if (B > 65535) B = 65535;
    if (B < 1) B = 0;
    (void)sm_count;
    return (int)B;
}

// ---- S3: matrix-free Chebyshev-filtered subspace iteration -------------------------------------------
static int run_small(scqed_plan *pl, const PlanDev &P, const cplx *coef, const cplx *Hin, const double *scal,
                     long long n_pts, int n, int k, int want_vec, int tidyup, const ResonatorArgs &res,
                     double *evals, cplx *evecs, double *post, int *info, cudaStream_t st);

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
static int run_matfree(scqed_plan *pl, const cplx *coef_all, long long n_pts, int k, int want_vec, double *evals,
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
            static bool st_attr = false;
            if (!st_attr) {
                CK(cudaFuncSetAttribute(mf_cheb_stencil<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
                CK(cudaFuncSetAttribute(mf_cheb_stencil<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
                CK(cudaFuncSetAttribute(mf_cheb_stencil<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
                CK(cudaFuncSetAttribute(mf_cheb_stencil<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
                st_attr = true;
            }
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(mf_cheb_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        attr_set = true;
    }
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
        CK(cudaFuncSetAttribute(mf_d2_kernel<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_real));
        CK(cudaFuncSetAttribute(mf_d2_kernel<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_real));
        CK(cudaFuncSetAttribute(mf_d2_kernel<0, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_real));
        CK(cudaFuncSetAttribute(mf_d2_kernel<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_real));
        CK(cudaFuncSetAttribute(mf_d2_kernel<0, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_cplx));
        CK(cudaFuncSetAttribute(mf_d2_kernel<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_cplx));
    }
    bool d2_real = false, d2_cpl_real = false, d2_realx = false;
    // TMA-pipelined real kernel: 1-D bulk copies need 16-byte aligned sources and sizes (rows of X and of H1)
    size_t d2_smem_tma = 0, d2_smem_dmma = 0;
    bool d2_tma = false, d2_dmma = false;
    if (use_d2) {      // FP64 tensor-core kernel: the default for real H and real vectors
        d2_smem_dmma = (size_t)D2_RB * (P.dims[0] + 16) * 8 + (size_t)D2_RB * (P.dims[1] + 16) * 8 + (size_t)(P.d2_nmulti + 2) * 16 + 64;
        d2_dmma = P.dims[1] <= 8 * 8 * D2_NT && d2_smem_dmma <= 200 * 1024 && !getenv("SCQED_NO_D2_DMMA");
        if (d2_dmma) {
            CK(cudaFuncSetAttribute(mf_d2_dmma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_dmma));
            CK(cudaFuncSetAttribute(mf_d2_dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_dmma));
        }
    }
    if (use_d2) {
        const int d0_ = P.dims[0], d1_ = P.dims[1];
        const size_t stage = (((size_t)D2_BK * d1_ * 16) + 127) & ~(size_t)127;
        d2_smem_tma = D2_ST * stage + (size_t)D2_RB * d0_ * 8 + (size_t)D2_RB * ((d1_ + 1) & ~1) * 8 + (size_t)(P.d2_nmulti + 2) * 16 + 128;
        d2_tma = (d1_ % 2 == 0) && (d0_ % 2 == 0) && (n % 1 == 0) && d2_smem_tma <= 200 * 1024 &&
                 getenv("SCQED_D2_TMA") && atoi(getenv("SCQED_D2_TMA")) != 0;      // opt-in: measured 7 % slower than the plain kernel (35.6 vs 38.4 pts/s)
        if (d2_tma) {
            CK(cudaFuncSetAttribute(mf_d2_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_tma));
            CK(cudaFuncSetAttribute(mf_d2_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d2_smem_tma));
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
        CK(cudaMemsetAsync(info_c, 1, (size_t)np * 4, st));
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
                    CK(cudaFuncSetAttribute(mf_cheb_smem_d2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_rx));
                    mf_cheb_smem_d2<true><<<dim3(b, (unsigned)np), 512, sm_rx, st>>>(a, X, Y, deg, d2re);
                } else if (use_d2 && d2_real && sm_cx <= 220 * 1024) {
                    CK(cudaFuncSetAttribute(mf_cheb_smem_d2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cx));
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

extern "C" int scqed_eigh_batched(int device, double *H_dev, int32_t n, int64_t batch, int32_t k, double *evals_dev,
                                  double *evecs_dev, int32_t *info_dev, int32_t solver, void *stream) {
    if (!H_dev || !evals_dev || !info_dev || n < 1 || batch < 1 || k < 1 || k > n) return fail(SCQED_EINVAL, "bad argument");
    if (set_device(device)) return SCQED_ECUDA;
    cudaStream_t st = (cudaStream_t)stream;
    scqed_plan tmp;   // only carries device properties + workspaces
    tmp.device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    tmp.sm_count = prop.multiProcessorCount;
    tmp.smem_optin = prop.sharedMemPerBlockOptin;
    memset(&tmp.P, 0, sizeof(tmp.P));
    int want_vec = evecs_dev != nullptr;
    ResonatorArgs res;
    memset(&res, 0, sizeof(res));
    bool fits = (solver != 5 && split_fits(&tmp, n, k, 1)) || small_fits(&tmp, n, k, want_vec, 1, nullptr);
    if (solver == 1 && !fits) return fail(SCQED_EUNSUPPORTED, "n=%d k=%d does not fit the fused path", n, k);
    if (n < 2) solver = 1;
    int rc;
    if (solver == 5 && !fits) return fail(SCQED_EUNSUPPORTED, "n=%d k=%d does not fit the fused path", n, k);
    bool small = (solver == 0 && fits) || solver == 1 || solver == 5;
    if (small) {
        rc = run_small_any(&tmp, tmp.P, nullptr, (const cplx *)H_dev, nullptr, batch, n, k, want_vec, 0, res, evals_dev,
                           (cplx *)evecs_dev, nullptr, info_dev, st, solver == 5);
        if (rc == S1_NOT_REAL) {            // 128 < n <= 224 and the matrices are not real: HBM solver instead
            if (solver != 0) rc = fail(SCQED_EUNSUPPORTED, "n=%d fits the shared-memory path only for real symmetric matrices", n);
            else small = false;
        }
        if (small && !rc) { cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) rc = fail(SCQED_ECUDA, "small eigh: %s", cudaGetErrorString(e)); }
        tmp.s1work.release(); tmp.s1vec.release();
    }
    if (!small) {
        long long B = pick_dense_batch(n, batch, k, tmp.sm_count);
        if (B < 1) return fail(SCQED_ENOMEM, "not enough device memory for one %d x %d matrix", n, n);
        int nseg = dense_nseg(n, tmp.sm_count, (int)B);
        const bool blocked = solver != 3;
        size_t wsb = blocked ? blk_ws_bytes(n, k, (int)B, nseg) : dense_ws_bytes(n, k, (int)B, nseg);
        if (tmp.work.ensure(wsb)) return fail(SCQED_ENOMEM, "dense workspace");
        DenseWs w;
        BlkWs bw;
        if (blocked) blk_ws_carve((unsigned char *)tmp.work.p, nullptr, n, k, (int)B, nseg, &bw);
        else dense_ws_carve((unsigned char *)tmp.work.p, n, k, (int)B, nseg, &w);
        rc = 0;
        for (long long b0 = 0; b0 < batch && !rc; b0 += B) {
            int nb = (int)(batch - b0 < B ? batch - b0 : B);
            // caller's matrices are row-major = column-major storage of conj(H): same spectrum,
            // conjugated eigenvectors
            cplx *Ab = (cplx *)H_dev + (size_t)b0 * n * n;
            double *ev = evals_dev + (size_t)b0 * k;
            cplx *vec = evecs_dev ? (cplx *)evecs_dev + (size_t)b0 * k * n : nullptr;
            rc = blocked ? run_dense_blocked_batch(Ab, n, nb, k, want_vec, 1, bw, ev, vec, info_dev + b0, st)
                         : run_dense_batch(&tmp, tmp.sm_count, Ab, n, nb, k, want_vec, 1, w, ev, vec, info_dev + b0, st);
        }
        if (!rc) { cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) rc = fail(SCQED_ECUDA, "dense eigh: %s", cudaGetErrorString(e)); }
        tmp.work.release();
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// the sweep
// ------------------------------------------------------------------------------------------------
static int check_opts(const scqed_plan *pl, const scqed_sweep_opts *o) {
    if (!o) return fail(SCQED_EINVAL, "opts is NULL");
    if (o->k < 1 || o->k > pl->P.n) return fail(SCQED_EINVAL, "k=%d outside 1..n=%d", o->k, pl->P.n);
    if (o->solver < 0 || o->solver > 5) return fail(SCQED_EINVAL, "solver=%d", o->solver);
    if (o->resonator) {
        if (!o->want_vectors) return fail(SCQED_EINVAL, "the Resonator evaluable needs eigenvectors (get_vectors=True)");
        if (o->k > SCQED_RES_KMAX || o->k < 2) return fail(SCQED_EUNSUPPORTED, "resonator strips support 2 <= k <= %d", SCQED_RES_KMAX);
        if (o->res_op_factor < 0 || o->res_op_factor >= pl->P.n_factors) return fail(SCQED_EINVAL, "res_op_factor");
        int ns = pl->P.n_scalar;
        if (o->res_gc < 0 || o->res_gc >= ns || o->res_frl < 0 || o->res_frl >= ns || o->res_qb < 0 || o->res_qb >= ns)
            return fail(SCQED_EINVAL, "resonator scalar slots out of range");
        if (o->res_nmax < 0) return fail(SCQED_EINVAL, "res_nmax");
    }
    return 0;
}

static int sweep_run_impl(scqed_plan *pl, const scqed_sweep_opts *o, const double *params_dev, int64_t n_pts,
                          double *evals_dev, double *evecs_dev, double *scalars_dev, double *post_dev,
                          int32_t *info_dev, void *stream);

extern "C" int scqed_sweep_run(scqed_plan *pl, const scqed_sweep_opts *o, const double *params_dev, int64_t n_pts,
                               double *evals_dev, double *evecs_dev, double *scalars_dev, double *post_dev,
                               int32_t *info_dev, void *stream) {
    int rc = sweep_run_impl(pl, o, params_dev, n_pts, evals_dev, evecs_dev, scalars_dev, post_dev, info_dev, stream);
    if (rc) return rc;
    // nodes solved in a rotated basis: eigenvectors go back to the reference's basis (post-ops ran in the rotated one)
    if (evecs_dev && o->want_vectors && !o->raw_basis && !pl->xform_node.empty())
        return scqed_back_transform(pl, evecs_dev, n_pts * o->k, stream);
    return 0;
}

static int sweep_run_impl(scqed_plan *pl, const scqed_sweep_opts *o, const double *params_dev, int64_t n_pts,
                          double *evals_dev, double *evecs_dev, double *scalars_dev, double *post_dev,
                          int32_t *info_dev, void *stream) {
    if (!pl || !evals_dev || !info_dev || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    if (pl->P.n_swept > 0 && !params_dev) return fail(SCQED_EINVAL, "params_dev is NULL");
    int rc = check_opts(pl, o);
    if (rc) return rc;
    if (o->want_vectors && !evecs_dev && !o->resonator) return fail(SCQED_EINVAL, "want_vectors set but evecs_dev is NULL");
    if (o->resonator && !post_dev) return fail(SCQED_EINVAL, "resonator set but post_dev is NULL");
    if (set_device(pl->device)) return SCQED_ECUDA;
    cudaStream_t st = (cudaStream_t)stream;
    const PlanDev &P = pl->P;
    const int n = P.n, k = o->k;

    if (pl->coef.ensure((size_t)n_pts * P.n_terms * 16)) return fail(SCQED_ENOMEM, "coefficients");
    double *scal = scalars_dev;
    if (!scal) {
        if (pl->scal.ensure((size_t)n_pts * (P.n_scalar + 1) * 8)) return fail(SCQED_ENOMEM, "scalars");
        scal = (double *)pl->scal.p;
    }
    rc = run_coef(pl, params_dev, n_pts, (cplx *)pl->coef.p, scal, st, nullptr, true);
    if (rc) return rc;

    ResonatorArgs res;
    memset(&res, 0, sizeof(res));
    if (o->resonator) {
        res.enabled = 1;
        res.nstrips = o->res_nmax + 1 + k;
        res.op_factor = o->res_op_factor;
        res.op_node = pl->factor_node[o->res_op_factor];
        res.s_gc = o->res_gc; res.s_frl = o->res_frl; res.s_qb = o->res_qb;
    }
    bool fits = (o->solver != 5 && !(o->resonator && o->k > SCQED_RES_KMAX) && split_fits(pl, n, k, P.n_terms > P.n_dense ? P.n_terms : P.n_dense, pl->s1_real_mode)) ||
                small_fits(pl, n, k, o->want_vectors, P.n_terms, nullptr);
    int solver = o->solver;
    if ((solver == 1 || solver == 5) && !fits) return fail(SCQED_EUNSUPPORTED, "n=%d k=%d does not fit the fused shared-memory path", n, k);
    if (solver == 0) {
        if (fits) solver = 1;
        else if (P.max_term_nn <= 3 && o->k + 6 <= MF_BMAX && (n > 4096 || pl->mf_work_per_row <= n / 2)) solver = 4;
        else solver = 2;
    }
    if (solver == 1 || solver == 5) {
        rc = run_small_any(pl, P, (const cplx *)pl->coef.p, nullptr, scal, n_pts, n, k, o->want_vectors, o->tidyup, res,
                           evals_dev, (cplx *)evecs_dev, post_dev, info_dev, st, solver == 5);
        if (rc != S1_NOT_REAL) return rc;
        // 128 < n <= 224 and H is not real symmetric: the packed shared-memory path does not apply
        if (o->solver == 1) return fail(SCQED_EUNSUPPORTED, "n=%d fits the shared-memory path only for real symmetric H", n);
        solver = (P.max_term_nn <= 3 && o->k + 6 <= MF_BMAX && pl->mf_work_per_row <= n / 2) ? 4 : 2;
    }
    // S2 / S3 with the Resonator evaluable: eigenvectors are needed on the device even when the
    // caller does not want them back
    cplx *vecs = (cplx *)evecs_dev;
    if (o->resonator && !vecs) {
        if (pl->resvec.ensure((size_t)n_pts * k * n * 16)) return fail(SCQED_ENOMEM, "eigenvector buffer for the Resonator evaluable");
        vecs = (cplx *)pl->resvec.p;
    }
    auto finish = [&]() -> int {
        if (!o->resonator) return 0;
        resonator_generic_kernel<<<(unsigned)n_pts, 128, 0, st>>>(P, res, scal, evals_dev, vecs, n, k, post_dev);
        LAUNCHED();
        CK(cudaGetLastError());
        return 0;
    };
    if (solver == 4) {
        const int mf_deg = getenv("SCQED_MF_DEG") ? atoi(getenv("SCQED_MF_DEG")) : 0;      // 0 = auto
        rc = run_matfree(pl, (const cplx *)pl->coef.p, n_pts, k, o->want_vectors, evals_dev, vecs, info_dev, st, mf_deg, 60, 1e-9);
        return rc ? rc : finish();
    }
    // ---- S2: batches of B matrices through HBM ------------------------------------------------
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
        rc = run_assemble(pl, (const cplx *)pl->coef.p + (size_t)p0 * P.n_terms, nb, o->tidyup, A, st);
        if (rc) return rc;
        double *ev = evals_dev + (size_t)p0 * k;
        cplx *vec = vecs ? vecs + (size_t)p0 * k * n : nullptr;
        rc = blocked ? run_dense_blocked_batch(A, n, nb, k, o->want_vectors, 1, bw, ev, vec, info_dev + p0, st)
                     : run_dense_batch(pl, pl->sm_count, A, n, nb, k, o->want_vectors, 1, w, ev, vec, info_dev + p0, st);
        if (rc) return rc;
    }
    return finish();
}

// real parts of a chunk of eigenvectors -> packed float64; flag[0] |= any non-zero imaginary part
__global__ void __launch_bounds__(256)
pack_real_kernel(const cplx *__restrict__ in, double *__restrict__ out, size_t count, int *__restrict__ flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < count) { const cplx v = in[i]; out[i] = v.x; bad = v.y != 0.0; }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

extern "C" int scqed_sweep_run_host(scqed_plan *pl, const scqed_sweep_opts *o, const double *params, int64_t n_pts,
                                    double *evals, double *evecs, double *scalars, double *post, int32_t *info) {
    if (!pl || !evals || !info || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    int rc = check_opts(pl, o);
    if (rc) return rc;
    if (pl->P.n_swept > 0 && !params) return fail(SCQED_EINVAL, "params is NULL");
    if (o->want_vectors && !evecs && !o->resonator) return fail(SCQED_EINVAL, "want_vectors set but evecs is NULL");
    if (o->resonator && !post) return fail(SCQED_EINVAL, "resonator set but post is NULL");
    if (set_device(pl->device)) return SCQED_ECUDA;
    if (!pl->own_stream) CK(cudaStreamCreateWithFlags(&pl->own_stream, cudaStreamNonBlocking));
    cudaStream_t st = pl->own_stream;
    const PlanDev &P = pl->P;
    const int n = P.n, k = o->k;
    const int nstrips = o->resonator ? o->res_nmax + 1 + k : 0;
    size_t b_par = (size_t)n_pts * (P.n_swept > 0 ? P.n_swept : 1) * 8;
    size_t b_ev = (size_t)n_pts * k * 8;
    size_t b_vec = (o->want_vectors && (evecs || o->resonator)) ? (size_t)n_pts * k * n * 16 : 0;
    size_t b_sc = (size_t)n_pts * (P.n_scalar > 0 ? P.n_scalar : 1) * 8;
    size_t b_post = (size_t)n_pts * nstrips * k * 8;
    if (pl->in_params.ensure(b_par) || pl->out_evals.ensure(b_ev) || pl->out_info.ensure((size_t)n_pts * 4) ||
        pl->out_scal.ensure(b_sc) || (b_vec && evecs && pl->out_evecs.ensure(b_vec)) || (b_post && pl->out_post.ensure(b_post)))
        return fail(SCQED_ENOMEM, "device staging buffers");
    if (!pl->copy_stream) CK(cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = pl->copy_stream;
    if (P.n_swept > 0) CK(cudaMemcpyAsync(pl->in_params.p, params, (size_t)n_pts * P.n_swept * 8, cudaMemcpyHostToDevice, st));
    // Chunked pipeline: chunk c+1 is solved on the compute stream while the results of chunk c
    // travel to the host on the copy stream (host buffers should be pinned for real overlap).
    long long nchunks = n_pts / 5000;       // chunks stay >= ~17 waves of the small-path kernels
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 8) nchunks = 8;
    double geom = 0.0;
    // measured on the bench workload (10 000 points, 251 MB of results, 171 MB with packed real eigenvectors):
    // complex vectors 3 chunks x 0.6 (12.9 -> 11.9 ms end to end), packed real vectors 2 chunks 77 % / 23 % (11.2 ms)
    if (nchunks == 2) { if (o->real_vectors) geom = 0.3; else { nchunks = 3; geom = 0.6; } }
    if (const char *envc = getenv("SCQED_HOST_CHUNKS")) { long long v = atoll(envc); if (v >= 1 && v <= 64) nchunks = v < n_pts ? v : n_pts; }
    const long long per = (n_pts + nchunks - 1) / nchunks;
    // Chunk sizes.  Default: equal.  SCQED_HOST_GEOM=r (0 < r < 1): geometric, chunk i+1 = r * chunk i, so that the
    // copy of chunk i hides behind the (shorter) solve of chunk i+1 and only a small last copy is exposed.
    std::vector<long long> sizes;
    {
        double r = geom;
        if (const char *eg = getenv("SCQED_HOST_GEOM")) r = atof(eg);
        else if (getenv("SCQED_HOST_CHUNKS")) r = 0.0;
        if (r > 0.05 && r < 0.999 && nchunks > 1) {
            double tot = 0.0, w = 1.0;
            for (long long c = 0; c < nchunks; ++c) { tot += w; w *= r; }
            long long left = n_pts;
            w = 1.0;
            for (long long c = 0; c < nchunks && left > 0; ++c) {
                long long sz = c == nchunks - 1 ? left : (long long)(n_pts * (w / tot) + 0.5);
                if (sz < 1) sz = 1;
                if (sz > left) sz = left;
                sizes.push_back(sz); left -= sz; w *= r;
            }
            if (left > 0) sizes.back() += left;
        } else {
            for (long long c0 = 0; c0 < n_pts; c0 += per) sizes.push_back(n_pts - c0 < per ? n_pts - c0 : per);
        }
    }
    double *d_ev = (double *)pl->out_evals.p, *d_sc = (double *)pl->out_scal.p;
    cplx *d_vec = (evecs && b_vec) ? (cplx *)pl->out_evecs.p : nullptr;
    const bool pack = o->real_vectors && d_vec;
    double *d_pack = nullptr;
    int *d_flag = nullptr;
    if (pack) {
        if (pl->out_pack.ensure(b_vec / 2 + 256)) return fail(SCQED_ENOMEM, "packed eigenvector buffer");
        d_flag = (int *)pl->out_pack.p;
        d_pack = (double *)((unsigned char *)pl->out_pack.p + 256);
        CK(cudaMemsetAsync(d_flag, 0, 4, st));
    }
    double *d_post = b_post ? (double *)pl->out_post.p : nullptr;
    int32_t *d_info = (int32_t *)pl->out_info.p;
    const int ns = P.n_scalar;
    auto copy_out = [&](long long c0, long long np, cudaEvent_t evt) -> int {
        CK(cudaStreamWaitEvent(cs, evt, 0));
        CK(cudaEventDestroy(evt));
        CK(cudaMemcpyAsync(evals + (size_t)c0 * k, d_ev + (size_t)c0 * k, (size_t)np * k * 8, cudaMemcpyDeviceToHost, cs));
        if (pack) CK(cudaMemcpyAsync(evecs + (size_t)c0 * k * n, d_pack + (size_t)c0 * k * n, (size_t)np * k * n * 8, cudaMemcpyDeviceToHost, cs));
        else if (d_vec) CK(cudaMemcpyAsync(evecs + (size_t)2 * c0 * k * n, d_vec + (size_t)c0 * k * n, (size_t)np * k * n * 16, cudaMemcpyDeviceToHost, cs));
        if (scalars && ns > 0) CK(cudaMemcpyAsync(scalars + (size_t)c0 * ns, d_sc + (size_t)c0 * ns, (size_t)np * ns * 8, cudaMemcpyDeviceToHost, cs));
        if (post && d_post) CK(cudaMemcpyAsync(post + (size_t)c0 * nstrips * k, d_post + (size_t)c0 * nstrips * k, (size_t)np * nstrips * k * 8, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(info + c0, d_info + c0, (size_t)np * 4, cudaMemcpyDeviceToHost, cs));
        return 0;
    };
    // Enqueue order: compute(c), then the copies of chunk c-1 -- a copy into pageable host memory
    // blocks the host, and must not delay the next chunk's kernels.
    long long prev_c0 = -1, prev_np = 0;
    cudaEvent_t prev_evt = nullptr;
    long long c0 = 0;
    for (size_t ci = 0; ci < sizes.size(); c0 += sizes[ci], ++ci) {
        const long long np = sizes[ci];
        rc = scqed_sweep_run(pl, o, (const double *)pl->in_params.p + (size_t)c0 * P.n_swept, np, d_ev + (size_t)c0 * k,
                             d_vec ? (double *)(d_vec + (size_t)c0 * k * n) : nullptr, d_sc + (size_t)c0 * (ns > 0 ? ns : 1),
                             d_post ? d_post + (size_t)c0 * nstrips * k : nullptr, d_info + c0, st);
        if (rc) { cudaStreamSynchronize(cs); return rc; }
        if (pack) {
            const size_t cnt = (size_t)np * k * n;
            pack_real_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_vec + (size_t)c0 * k * n, d_pack + (size_t)c0 * k * n, cnt, d_flag);
            LAUNCHED();
        }
        cudaEvent_t evt;
        CK(cudaEventCreateWithFlags(&evt, cudaEventDisableTiming));
        CK(cudaEventRecord(evt, st));
        if (prev_c0 >= 0) { rc = copy_out(prev_c0, prev_np, prev_evt); if (rc) return rc; }
        prev_c0 = c0; prev_np = np; prev_evt = evt;
    }
    if (prev_c0 >= 0) { rc = copy_out(prev_c0, prev_np, prev_evt); if (rc) return rc; }
    int h_flag = 0;
    if (pack) CK(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(cs));
    if (pack && h_flag) { g_err = "real_vectors: some eigenvector has a non-zero imaginary part"; return SCQED_ECOMPLEX; }
    // a plan assumed real-symmetric produced a complex point: redo the sweep in complex mode (once)
    bool redo = false;
    for (long long q = 0; q < n_pts; ++q) if (info[q] == -2) { redo = true; break; }
    if (redo && pl->s1_real_mode != 0) {
        pl->s1_real_mode = 0;
        return scqed_sweep_run_host(pl, o, params, n_pts, evals, evecs, scalars, post, info);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// probes
// ------------------------------------------------------------------------------------------------
template <typename K>
static int run_probe(int device, K kernel, double flops_per_thread_iter, double *out) {
    if (!out) return fail(SCQED_EINVAL, "null argument");
    if (set_device(device)) return SCQED_ECUDA;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double *dummy;
    CK(cudaMalloc(&dummy, 8));
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    kernel<<<blocks, threads>>>(iters / 10, dummy);
    CK(cudaDeviceSynchronize());
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kernel<<<blocks, threads>>>(iters, dummy);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double tf = flops_per_thread_iter * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    g_launches.fetch_add(4, std::memory_order_relaxed);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dummy);
    *out = best;
    return 0;
}

extern "C" int scqed_probe_fp64_tensor(int device, double *tflops_out) {
    // 8 DMMA.8x8x4 per iteration per warp: 8 * 2*8*8*4 flops per warp = 4096/32 per thread
    return run_probe(device, probe_dmma_kernel, 8.0 * 512.0 / 32.0, tflops_out);
}
extern "C" int scqed_probe_fp64_fma(int device, double *tflops_out) {
    return run_probe(device, probe_dfma_kernel, 16.0 * 2.0, tflops_out);
}
