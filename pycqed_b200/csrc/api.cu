// api.cu -- C ABI of the sweep engine (see include/scqed.h for the contract and the reference
// functions each entry point replaces).  Host-side orchestration only; the solver families live in
// small.cu / dense.cu / matfree.cu, kernels in the .cuh files next to this one.
#include "internal.h"
#include <mutex>
#include <immintrin.h>
#include "coef.cuh"
#include "assemble.cuh"
#include "probe.cuh"
#include "pauli.cuh"

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<long long> g_matvecs{0};
std::atomic<long long> g_mf_flops{0};
std::atomic<int> g_timing{0};
std::vector<TimedSpan> g_spans;
std::mutex g_spans_mu;

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
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

    ON_DEVICE(device);
    // owned until the very end: every early `return fail(...)` below frees the plan and its device tables
    struct PlanDeleter { void operator()(scqed_plan *q) const { if (q) { q->tables.release(); delete q; } } };
    std::unique_ptr<scqed_plan, PlanDeleter> owner(new scqed_plan());
    scqed_plan *pl = owner.get();
    pl->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(SCQED_ECUDA, "cudaGetDeviceProperties failed");
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
        { return fail(SCQED_EINVAL, "dense term table malformed"); }
    for (int g = 0; g < n_dense; ++g)
        if (d->dense_re[g] >= d->n_instr || d->dense_im[g] >= d->n_instr) { return fail(SCQED_EINVAL, "dense table register out of range"); }
    if (d->n_dense_diag < 0 || d->n_dense_diag > n_dense) { return fail(SCQED_EINVAL, "n_dense_diag"); }
    size_t o_dtab = put(d->dense_tab, ((size_t)(n_dense - d->n_dense_diag) * n * n + (size_t)d->n_dense_diag * n) * 8);
    size_t o_dre = put(d->dense_re, (size_t)n_dense * 4), o_dim = put(d->dense_im, (size_t)n_dense * 4);
    size_t o_dgm = put(d->dense_gmax, (size_t)n_dense * 8);
    size_t o_ekey = 0, o_eptr = 0, o_eterm = 0, o_eval = 0, o_emap = 0;
    if (d->n_entries < 0 || (d->n_entries > 0 && (!d->ent_key || !d->ent_ptr || !d->ent_term || !d->ent_val))) { return fail(SCQED_EINVAL, "entry list malformed"); }
    if (d->n_entries > 0) {
        const int ne = d->n_entries;
        const long long np_tot = d->ent_ptr[ne];
        for (int e = 0; e < ne; ++e) {
            if (d->ent_key[e] < 0 || d->ent_key[e] >= n * n || (e > 0 && d->ent_key[e] <= d->ent_key[e - 1]) || d->ent_ptr[e + 1] < d->ent_ptr[e])
                { return fail(SCQED_EINVAL, "entry %d malformed", e); }
        }
        for (long long q = 0; q < np_tot; ++q)
            if (d->ent_term[q] < 0 || d->ent_term[q] >= d->n_terms) { return fail(SCQED_EINVAL, "entry pair %lld: term out of range", q); }
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
    if (d->n_xform < 0 || (d->n_xform > 0 && (!d->xform_node || !d->xform_data))) { return fail(SCQED_EINVAL, "basis transforms malformed"); }
    {
        const double *src = d->xform_data;
        for (int i = 0; i < d->n_xform; ++i) {
            const int k = d->xform_node[i];
            if (k < 0 || k >= d->n_nodes) { return fail(SCQED_EINVAL, "xform_node[%d]=%d", i, k); }
            const int dk = d->dims[k];
            std::vector<double> wt((size_t)dk * dk);
            for (int a = 0; a < dk; ++a)
                for (int b = 0; b < dk; ++b) wt[(size_t)b * dk + a] = src[(size_t)a * dk + b];     // W^T: thread a reads wt[b*d + a]
            pl->xform_node.push_back(k);
            pl->xform_off.push_back(put(wt.data(), wt.size() * 8));
            src += (size_t)dk * dk;
        }
    }
    if (pl->tables.ensure(host.size() + 256)) { return fail(SCQED_ENOMEM, "plan tables"); }
    if (cudaMemcpy(pl->tables.p, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        return fail(SCQED_ECUDA, "uploading plan tables failed");
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
    *out = owner.release();
    return SCQED_OK;
}

extern "C" void scqed_plan_destroy(scqed_plan *pl) {
    if (!pl) return;
    DeviceGuard dev_guard_(pl->device);
    DevBuf *bufs[] = {&pl->tables, &pl->regs, &pl->coef, &pl->scal, &pl->work, &pl->s1work, &pl->s1vec, &pl->auxc, &pl->resvec, &pl->dcoef, &pl->s1ladder, &pl->s1slots, &pl->dtg, &pl->d2h, &pl->mfwarm, &pl->in_params, &pl->out_evals,
                      &pl->out_evecs, &pl->out_post, &pl->out_info, &pl->out_scal, &pl->out_pack};
    for (DevBuf *b : bufs) b->release();
    pl->hstage.release();
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
    std::vector<TimedSpan> keep, mine;
    { std::lock_guard<std::mutex> lk(g_spans_mu); mine.swap(g_spans); }
    for (auto &sp : mine) {
        if (sp.tag != tag) { keep.push_back(sp); continue; }
        float ms = 0.f;
        if (cudaEventSynchronize(sp.e1) == cudaSuccess && cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { tot += ms; ++cnt; }
        cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1);
    }
    { std::lock_guard<std::mutex> lk(g_spans_mu); g_spans.insert(g_spans.end(), keep.begin(), keep.end()); }
    *total_ms = tot; *count = cnt;
    return 0;
}
extern "C" int scqed_counter_read(int which, int64_t *value) {
    if (!value) return fail(SCQED_EINVAL, "null argument");
    if (which == 1) { *value = g_matvecs.exchange(0); return 0; }
    if (which == 2) { *value = g_mf_flops.exchange(0); return 0; }
    return fail(SCQED_EINVAL, "unknown counter %d", which);
}
extern "C" const char *scqed_last_error(void) { return g_err.c_str(); }
extern "C" const char *scqed_version(void) { return "scqed 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// K1 / K2
// ------------------------------------------------------------------------------------------------
int run_coef(scqed_plan *pl, const double *params_dev, long long n_pts, cplx *coef_dev, double *scal_dev,
             cudaStream_t st, cplx *aux_dev, bool dense) {
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
    ON_DEVICE(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->coef.ensure((size_t)n_pts * P.n_terms * 16)) return fail(SCQED_ENOMEM, "coefficients");
    if (pl->scal.ensure((size_t)n_pts * (P.n_scalar + 1) * 8)) return fail(SCQED_ENOMEM, "scalars");
    if (pl->auxc.ensure((size_t)n_pts * (P.n_aux + 1) * 16)) return fail(SCQED_ENOMEM, "aux coefficients");
    int rc = run_coef(pl, params_dev, n_pts, (cplx *)pl->coef.p, (double *)pl->scal.p, st, (cplx *)pl->auxc.p);
    if (rc) return rc;
    const long long warps = n_pts * P.n_pairs;
    (void)warps;
    return run_aux_matel(P, (const cplx *)pl->auxc.p, (const cplx *)evecs_dev, n_pts, k, (cplx *)out_dev, st);
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
    ON_DEVICE(pl->device);
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
    ON_DEVICE(device);
    pauli_coef_kernel<<<(unsigned)((n_pts + 127) / 128), 128, 0, (cudaStream_t)stream>>>(e01_dev, (const cplx *)op_dev, n_pts, h_dev);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int scqed_dense_matrix_elements(int device, const double *op_dev, const double *evecs_dev, int32_t n, int32_t m,
                                           int64_t n_pts, double *out_dev, void *stream) {
    if (!op_dev || !evecs_dev || !out_dev || n < 1 || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    if (m < 1 || m > DMEL_MMAX) return fail(SCQED_EUNSUPPORTED, "1 <= m <= %d vectors per point", DMEL_MMAX);
    ON_DEVICE(device);
    const size_t smem = ((size_t)m * n + (size_t)8 * m * m) * 16;
    if (smem > 200 * 1024) return fail(SCQED_EUNSUPPORTED, "n=%d too large for the dense matrix-element kernel", n);
    CK(cudaFuncSetAttribute(dense_matel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // constant: safe under concurrent callers
    dense_matel_kernel<<<(unsigned)n_pts, 256, smem, (cudaStream_t)stream>>>((const cplx *)op_dev, (const cplx *)evecs_dev, n, m, n_pts, (cplx *)out_dev);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int scqed_eval_coefficients(scqed_plan *pl, const double *params_dev, int64_t n_pts, double *coef_dev,
                                       double *scalars_dev, void *stream) {
    if (!pl || !coef_dev || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    if (pl->P.n_swept > 0 && !params_dev) return fail(SCQED_EINVAL, "params_dev is NULL");
    if (pl->P.n_scalar > 0 && !scalars_dev) return fail(SCQED_EINVAL, "scalars_dev is NULL but the plan has scalar outputs");
    ON_DEVICE(pl->device);
    return run_coef(pl, params_dev, n_pts, (cplx *)coef_dev, scalars_dev, (cudaStream_t)stream);
}

int run_assemble(scqed_plan *pl, const cplx *coef_dev, long long n_pts, int tidyup, cplx *H, cudaStream_t st) {
    const PlanDev &P = pl->P;
    if (P.n_entries > 0 && !getenv("SCQED_NO_ENTRY_LIST")) {
        const size_t nn = (size_t)P.n * P.n;
        const long long eb = P.ent_dense ? (long long)((nn + 255) / 256) : (P.n_entries + 255) / 256;
        if (eb <= 65535 && n_pts <= 2147483647LL) {
            TimingScope tscope_(st, 3);
            if (P.ent_dense) {
                assemble_mapped_kernel<<<dim3((unsigned)((n_pts + ASM_PG - 1) / ASM_PG), (unsigned)eb), 256, (size_t)ASM_PG * P.n_terms * 16, st>>>(P, coef_dev, n_pts, tidyup, H);
            } else {
                CK(cudaMemsetAsync(H, 0, (size_t)n_pts * nn * 16, st));
                assemble_entries_kernel<<<dim3((unsigned)((n_pts + ASM_PG - 1) / ASM_PG), (unsigned)eb), 256, (size_t)ASM_PG * P.n_terms * 16, st>>>(P, coef_dev, n_pts, tidyup, H);
            }
            tscope_.end();
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
    ON_DEVICE(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->coef.ensure((size_t)n_pts * pl->P.n_terms * 16)) return fail(SCQED_ENOMEM, "coefficients");
    if (pl->scal.ensure((size_t)n_pts * (pl->P.n_scalar + 1) * 8)) return fail(SCQED_ENOMEM, "scalars");
    int rc = run_coef(pl, params_dev, n_pts, (cplx *)pl->coef.p, (double *)pl->scal.p, st);
    if (rc) return rc;
    // row-major H[i][j]: the kernel's fast index is the column
    return run_assemble(pl, (const cplx *)pl->coef.p, n_pts, tidyup, (cplx *)H_dev, st);
}

extern "C" int scqed_eigh_batched(int device, double *H_dev, int32_t n, int64_t batch, int32_t k, double *evals_dev,
                                  double *evecs_dev, int32_t *info_dev, int32_t solver, void *stream) {
    if (!H_dev || !evals_dev || !info_dev || n < 1 || batch < 1 || k < 1 || k > n) return fail(SCQED_EINVAL, "bad argument");
    ON_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    // A per-device context carries the device properties and the workspaces from call to call (cudaMalloc / cudaFree of the
    // GB-sized dense workspaces cost 10-500 ms per call, several times the solve itself for mid-size batches).  Calls on
    // one device are serialised by its mutex; the buffers are kept while they hold <= 16 GB, else released at the end.
    if (device < 0 || device >= 64) return fail(SCQED_EINVAL, "device %d", device);
    static std::mutex eigh_mu[64];
    static scqed_plan *eigh_ctx[64];
    std::lock_guard<std::mutex> eigh_lock(eigh_mu[device]);
    if (!eigh_ctx[device]) {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        scqed_plan *c = new scqed_plan();
        c->device = device;
        c->sm_count = prop.multiProcessorCount;
        c->smem_optin = prop.sharedMemPerBlockOptin;
        memset(&c->P, 0, sizeof(c->P));
        eigh_ctx[device] = c;
    }
    scqed_plan &tmp = *eigh_ctx[device];
    tmp.s1_real_mode = -1;                         // per-call state: the realness of THIS batch is probed afresh
    struct Trim {
        scqed_plan &t;
        ~Trim() {
            DevBuf *b[] = {&t.s1work, &t.s1vec, &t.s1ladder, &t.s1slots, &t.dcoef, &t.work, &t.dtg};
            size_t tot = 0;
            for (DevBuf *x : b) tot += x->cap;
            if (tot > ((size_t)16 << 30)) for (DevBuf *x : b) x->release();
        }
    } trim_{tmp};
    int want_vec = evecs_dev != nullptr;
    ResonatorArgs res;
    memset(&res, 0, sizeof(res));
    bool fits = (solver != 5 && split_fits(&tmp, n, k, 1)) || small_fits(&tmp, n, k, want_vec, 1);
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
    }
    if (!small) {
        rc = dense_eigh_resident(&tmp, (cplx *)H_dev, n, batch, k, want_vec, evals_dev, (cplx *)evecs_dev, info_dev, solver, st);
        if (!rc) { cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) rc = fail(SCQED_ECUDA, "dense eigh: %s", cudaGetErrorString(e)); }
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
    ON_DEVICE(pl->device);
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
                small_fits(pl, n, k, o->want_vectors, P.n_terms);
    int solver = o->solver;
    if ((solver == 1 || solver == 5) && !fits) return fail(SCQED_EUNSUPPORTED, "n=%d k=%d does not fit the fused shared-memory path", n, k);
    if (solver == 0) {
        if (fits) solver = 1;
        else if (P.max_term_nn <= 3 && o->k + 6 <= MF_BMAX_HOST && (n > 4096 || pl->mf_work_per_row <= n / 2)) solver = 4;
        else solver = 2;
    }
    if (solver == 1 || solver == 5) {
        rc = run_small_any(pl, P, (const cplx *)pl->coef.p, nullptr, scal, n_pts, n, k, o->want_vectors, o->tidyup, res,
                           evals_dev, (cplx *)evecs_dev, post_dev, info_dev, st, solver == 5);
        if (rc != S1_NOT_REAL) return rc;
        // 128 < n <= 224 and H is not real symmetric: the packed shared-memory path does not apply
        if (o->solver == 1) return fail(SCQED_EUNSUPPORTED, "n=%d fits the shared-memory path only for real symmetric H", n);
        solver = (P.max_term_nn <= 3 && o->k + 6 <= MF_BMAX_HOST && pl->mf_work_per_row <= n / 2) ? 4 : 2;
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
        return run_resonator_generic(P, res, scal, evals_dev, vecs, n_pts, n, k, post_dev, info_dev, st);
    };
    if (solver == 4) {
        const int mf_deg = getenv("SCQED_MF_DEG") ? atoi(getenv("SCQED_MF_DEG")) : 0;      // 0 = auto
        rc = run_matfree(pl, (const cplx *)pl->coef.p, n_pts, k, o->want_vectors, evals_dev, vecs, info_dev, st, mf_deg, 60, 1e-9);
        return rc ? rc : finish();
    }
    // ---- S2: batches of B matrices through HBM (dense.cu) ----------------------------------------
    // (o->solver == 0: the tile-packed real symmetric path S2t may take the batch; an explicit solver = 2 / 3 keeps the
    // complex blocked / unblocked kernels, e.g. for the cross-checks of the test suite)
    rc = dense_sweep(pl, (const cplx *)pl->coef.p, n_pts, k, o->want_vectors, o->tidyup, solver, o->solver == 0, evals_dev, vecs, info_dev, st);
    if (rc) return rc;
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

// ---- host-buffer sweep ------------------------------------------------------------------------------------
// true when `ptr` is page-locked host memory known to CUDA (cudaHostAlloc / cudaHostRegister / torch pin_memory)
static bool host_ptr_is_pinned(const void *ptr) {
    if (!ptr) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// Copy out of the page-locked staging ring into the caller's pageable array.  The destination is written once and not
// read again by this library: non-temporal 32-byte stores skip the read-for-ownership of every destination line (a plain
// memcpy moves 3 bytes of DRAM traffic per byte copied, this moves 2).
__attribute__((target("avx"))) static void stream_copy_avx(char *dst, const char *src, size_t bytes) {
    while (bytes && ((uintptr_t)dst & 31)) { *dst++ = *src++; --bytes; }
    const size_t nv = bytes / 32;
    for (size_t i = 0; i < nv; ++i)
        _mm256_stream_si256((__m256i *)dst + i, _mm256_loadu_si256((const __m256i *)src + i));
    _mm_sfence();
    if (bytes - nv * 32) memcpy(dst + nv * 32, src + nv * 32, bytes - nv * 32);
}
static void stream_copy(char *dst, const char *src, size_t bytes) {
    static const bool avx = __builtin_cpu_supports("avx");
    if (avx && bytes >= 4096) stream_copy_avx(dst, src, bytes);
    else memcpy(dst, src, bytes);
}

// a large block on up to 8 threads (one core moves ~5-10 GB/s; results are 100s of MB per sweep)
static void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    const size_t min_part = (size_t)2 << 20;
    unsigned hw = std::thread::hardware_concurrency();
    int parts = (int)(bytes / min_part);
    if (parts > 8) parts = 8;
    if (hw > 1 && parts > (int)hw / 2) parts = (int)hw / 2;
    if (parts <= 1) { stream_copy((char *)dst, (const char *)src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes / parts) + 4095) & ~(size_t)4095;
    for (int i = 1; i < parts; ++i) {
        const size_t off = (size_t)i * per;
        if (off >= bytes) break;
        const size_t len = bytes - off < per ? bytes - off : per;
        th.emplace_back([=] { stream_copy((char *)dst + off, (const char *)src + off, len); });
    }
    stream_copy((char *)dst, (const char *)src, per < bytes ? per : bytes);
    for (auto &t : th) t.join();
}

struct EventList {                                  // events of one call, destroyed on every exit path
    std::vector<cudaEvent_t> ev;
    int add(cudaEvent_t *out) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return -1; }
        ev.push_back(e); *out = e; return 0;
    }
    ~EventList() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
};

extern "C" int scqed_sweep_run_host(scqed_plan *pl, const scqed_sweep_opts *o, const double *params, int64_t n_pts,
                                    double *evals, double *evecs, double *scalars, double *post, int32_t *info) {
    if (!pl || !evals || !info || n_pts < 1) return fail(SCQED_EINVAL, "bad argument");
    int rc = check_opts(pl, o);
    if (rc) return rc;
    if (pl->P.n_swept > 0 && !params) return fail(SCQED_EINVAL, "params is NULL");
    if (o->want_vectors && !evecs && !o->resonator) return fail(SCQED_EINVAL, "want_vectors set but evecs is NULL");
    if (o->resonator && !post) return fail(SCQED_EINVAL, "resonator set but post is NULL");
    ON_DEVICE(pl->device);
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
    const bool vec_dev = o->evecs_on_device && evecs && b_vec;      // eigenvectors stay in the caller's device buffer
    if (vec_dev && o->real_vectors) return fail(SCQED_EINVAL, "real_vectors and evecs_on_device are mutually exclusive");
    if (pl->in_params.ensure(b_par) || pl->out_evals.ensure(b_ev) || pl->out_info.ensure((size_t)n_pts * 4) ||
        pl->out_scal.ensure(b_sc) || (b_vec && evecs && !vec_dev && pl->out_evecs.ensure(b_vec)) || (b_post && pl->out_post.ensure(b_post)))
        return fail(SCQED_ENOMEM, "device staging buffers");
    if (!pl->copy_stream) CK(cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = pl->copy_stream;
    const int ns = P.n_scalar;
    const bool pack = o->real_vectors && evecs && b_vec;

    // ---- which caller buffers are pageable?  Those go through the library's pinned staging ring: the GPU copies
    // into a page-locked slot (a true asynchronous DMA) and a host thread moves the slot into the caller's array
    // while the next chunk is solved.  Page-locked caller buffers (cudaHostAlloc, torch pin_memory) are written directly.
    struct Seg { char *user; const char *dev; size_t per_pt; bool staged; };
    Seg segs[5] = {
        {(char *)evals, nullptr, (size_t)k * 8, false},
        {(char *)evecs, nullptr, (evecs && b_vec && !vec_dev) ? (size_t)k * n * (pack ? 8 : 16) : 0, false},
        {(char *)scalars, nullptr, (scalars && ns > 0) ? (size_t)ns * 8 : 0, false},
        {(char *)post, nullptr, (post && b_post) ? (size_t)nstrips * k * 8 : 0, false},
        {(char *)info, nullptr, 4, false},
    };
    const bool force_stage = getenv("SCQED_FORCE_STAGING") != nullptr;
    size_t staged_per_pt = 0, bytes_per_pt = 0;
    for (Seg &sg : segs) {
        if (!sg.per_pt) continue;
        sg.staged = force_stage || !host_ptr_is_pinned(sg.user);
        if (sg.staged) staged_per_pt += sg.per_pt;
        bytes_per_pt += sg.per_pt;
    }
    const bool params_pinned = P.n_swept == 0 || host_ptr_is_pinned(params);
    if (P.n_swept > 0) {
        // a pageable source makes this call synchronous, which is fine here: nothing has been enqueued yet
        CK(cudaMemcpyAsync(pl->in_params.p, params, (size_t)n_pts * P.n_swept * 8, cudaMemcpyHostToDevice, st));
        (void)params_pinned;
    }
    // Chunked pipeline: chunk c+1 is solved on the compute stream while the results of chunk c
    // travel to the host on the copy stream.
    long long nchunks = n_pts / 5000;       // chunks stay >= ~17 waves of the small-path kernels
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 8) nchunks = 8;
    double geom = 0.0;
    // measured on the bench workload (10 000 points, 251 MB of results, 171 MB with packed real eigenvectors):
    // complex vectors 3 chunks x 0.6 (12.9 -> 11.9 ms end to end), packed real vectors 2 chunks 77 % / 23 % (11.2 ms)
    if (nchunks == 2) { if (o->real_vectors) geom = 0.3; else { nchunks = 3; geom = 0.6; } }
    // Adaptive: when the copies are the long pole (measured on the previous call of this plan, see the end of this
    // function) the first copy must start as early as possible and every copy needs a solve to hide behind: four equal
    // chunks.  Measured with 2 / 4 / 8 chunks of a 10 000-point sweep: solve 8.7 / 10.2 / 12.7 ms -- chunks are not free,
    // so the switch needs copies that outlast the solve by 45 % of its time (4 or more GPUs sharing this host).  8 ranks of one box sharing the host: 170 MB per rank at 11.5 GB/s = 14.9 ms against 9.1 ms of solving.
    if (pl->host_copy_pts != n_pts) { pl->host_copy_pts = n_pts; pl->host_copy_bound = 0; }     // measured per sweep size
    const bool reprobe = pl->host_copy_bound && (++pl->host_calls % 16) == 0;      // now and then try the default again
    if (pl->host_copy_bound && !reprobe && n_pts >= 5000) {
        nchunks = 4;            // every extra chunk costs ~0.5 ms of kernel tails (8 launches per chunk): 8 chunks were not better
        geom = 0.0;
    }
    if (const char *envc = getenv("SCQED_HOST_CHUNKS")) { long long v = atoll(envc); if (v >= 1 && v <= 64) nchunks = v < n_pts ? v : n_pts; }
    // Chunk sizes.  Default: equal.  SCQED_HOST_GEOM=r (0 < r < 1): geometric, chunk i+1 = r * chunk i, so that the
    // copy of chunk i hides behind the (shorter) solve of chunk i+1 and only a small last copy is exposed.
    std::vector<long long> sizes;
    {
        double r = geom;
        if (const char *eg = getenv("SCQED_HOST_GEOM")) r = atof(eg);
        else if (getenv("SCQED_HOST_CHUNKS")) r = 0.0;
        // the staging ring holds two slots of the largest chunk: keep a slot under 256 MB
        const size_t slot_cap = (size_t)256 << 20;
        long long max_pts = staged_per_pt ? (long long)(slot_cap / staged_per_pt) : n_pts;
        if (max_pts < 1) max_pts = 1;
        const long long per0 = (n_pts + nchunks - 1) / nchunks;
        if (staged_per_pt && per0 > max_pts) { nchunks = (n_pts + max_pts - 1) / max_pts; r = 0.0; }
        const long long per = (n_pts + nchunks - 1) / nchunks;
        if (r > 0.05 && r < 0.999 && nchunks > 1) {
            double tot = 0.0, w = 1.0;
            for (long long c = 0; c < nchunks; ++c) { tot += w; w *= r; }
            long long left = n_pts;
            w = 1.0;
            for (long long c = 0; c < nchunks && left > 0; ++c) {
                long long sz = c == nchunks - 1 ? left : (long long)(n_pts * (w / tot) + 0.5);
                if (sz < 1) sz = 1;
                if (sz > left) sz = left;
                if (staged_per_pt && sz > max_pts) sz = max_pts;
                sizes.push_back(sz); left -= sz; w *= r;
            }
            while (left > 0) { long long sz = left < max_pts ? left : max_pts; sizes.push_back(sz); left -= sz; }
        } else {
            for (long long c0 = 0; c0 < n_pts; c0 += per) sizes.push_back(n_pts - c0 < per ? n_pts - c0 : per);
        }
    }
    long long max_chunk = 0;
    for (long long sz : sizes) if (sz > max_chunk) max_chunk = sz;
    const size_t slot_bytes = ((size_t)max_chunk * staged_per_pt + 4095) & ~(size_t)4095;
    if (staged_per_pt && pl->hstage.ensure(2 * slot_bytes)) return fail(SCQED_ENOMEM, "pinned staging ring (%zu MB)", (2 * slot_bytes) >> 20);

    double *d_ev = (double *)pl->out_evals.p, *d_sc = (double *)pl->out_scal.p;
    cplx *d_vec = (evecs && b_vec) ? (vec_dev ? (cplx *)evecs : (cplx *)pl->out_evecs.p) : nullptr;
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
    segs[0].dev = (const char *)d_ev;
    segs[1].dev = pack ? (const char *)d_pack : (const char *)d_vec;
    segs[2].dev = (const char *)d_sc;
    segs[3].dev = (const char *)d_post;
    segs[4].dev = (const char *)d_info;

    // ---- the host-side mover of the staging ring ------------------------------------------------------------
    // A chunk's staged bytes travel in pieces of <= 32 MB: the DMA of piece i+1 overlaps the host copy of piece i.
    struct Job { char *dst; const char *src; size_t bytes; cudaEvent_t copied; bool last_of_chunk; };
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Job> jobs;
    size_t chunks_moved = 0;
    bool stop = false, mover_failed = false;
    std::thread mover;
    if (staged_per_pt) {
        mover = std::thread([&] {
            cudaSetDevice(pl->device);
            for (size_t j = 0;; ++j) {
                Job job;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || jobs.size() > j; });
                    if (jobs.size() <= j) return;
                    job = jobs[j];
                }
                if (cudaEventSynchronize(job.copied) != cudaSuccess) { cudaGetLastError(); mover_failed = true; }
                else parallel_memcpy(job.dst, job.src, job.bytes);
                if (job.last_of_chunk) {
                    { std::lock_guard<std::mutex> lk(mu); ++chunks_moved; }
                    cv.notify_all();
                }
            }
        });
    }
    auto finish_mover = [&]() {
        if (!mover.joinable()) return;
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        mover.join();
    };
    EventList events;
    // every exit below goes through here: both streams drained (the host buffers may still be written), mover joined
    auto bail = [&](int code) { cudaStreamSynchronize(st); cudaStreamSynchronize(cs); finish_mover(); return code; };
#define CKH(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return bail(fail(SCQED_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)

    size_t n_staged_chunks = 0;
    const size_t piece = (size_t)32 << 20;
    auto copy_out = [&](long long c0, long long np, cudaEvent_t solved) -> int {
        CKH(cudaStreamWaitEvent(cs, solved, 0));
        char *dst_stage = nullptr;
        if (staged_per_pt) {
            const int slot = (int)(n_staged_chunks & 1);
            if (n_staged_chunks >= 2) {              // the slot's previous content must have reached the caller
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return chunks_moved + 2 > n_staged_chunks; });
            }
            dst_stage = (char *)pl->hstage.p + (size_t)slot * slot_bytes;
        }
        int last_staged = -1;
        for (int q = 0; q < 5; ++q) if (segs[q].per_pt && segs[q].staged) last_staged = q;
        for (int q = 0; q < 5; ++q) {
            const Seg &sg = segs[q];
            if (!sg.per_pt) continue;
            const size_t bytes = (size_t)np * sg.per_pt;
            const char *src = sg.dev + (size_t)c0 * sg.per_pt;
            if (!sg.staged) {
                CKH(cudaMemcpyAsync(sg.user + (size_t)c0 * sg.per_pt, src, bytes, cudaMemcpyDeviceToHost, cs));
                continue;
            }
            for (size_t off = 0; off < bytes; off += piece) {
                const size_t len = bytes - off < piece ? bytes - off : piece;
                CKH(cudaMemcpyAsync(dst_stage + off, src + off, len, cudaMemcpyDeviceToHost, cs));
                cudaEvent_t copied;
                if (events.add(&copied)) return bail(fail(SCQED_ECUDA, "cudaEventCreate failed"));
                CKH(cudaEventRecord(copied, cs));
                { std::lock_guard<std::mutex> lk(mu);
                  jobs.push_back({sg.user + (size_t)c0 * sg.per_pt + off, dst_stage + off, len, copied, q == last_staged && off + len >= bytes}); }
                cv.notify_all();
            }
            dst_stage += bytes;
        }
        if (staged_per_pt) ++n_staged_chunks;
        return 0;
    };
    // The per-chunk solves stay asynchronous: the "a real plan turned complex" check of the shared-memory solver
    // (one stream synchronisation per call on the device path) is done once, below, on the host copy of info.
    struct DeferGuard { scqed_plan *p; bool old; DeferGuard(scqed_plan *q) : p(q), old(q->defer_real_check) { q->defer_real_check = true; }
                        ~DeferGuard() { p->defer_real_check = old; } } defer_guard(pl);
    // three timed events decide the chunking of the NEXT call: start, last chunk solved, last copy done
    cudaEvent_t t_start = nullptr, t_solved = nullptr, t_copied = nullptr;
    struct TimedEvents { cudaEvent_t *e[3]; ~TimedEvents() { for (auto p_ : e) if (*p_) cudaEventDestroy(*p_); } } timed_guard{{&t_start, &t_solved, &t_copied}};
    if (cudaEventCreate(&t_start) != cudaSuccess || cudaEventCreate(&t_solved) != cudaSuccess || cudaEventCreate(&t_copied) != cudaSuccess)
        return bail(fail(SCQED_ECUDA, "cudaEventCreate failed"));
    CKH(cudaEventRecord(t_start, st));
    long long c0 = 0;
    for (size_t ci = 0; ci < sizes.size(); c0 += sizes[ci], ++ci) {
        const long long np = sizes[ci];
        rc = scqed_sweep_run(pl, o, (const double *)pl->in_params.p + (size_t)c0 * P.n_swept, np, d_ev + (size_t)c0 * k,
                             d_vec ? (double *)(d_vec + (size_t)c0 * k * n) : nullptr, d_sc + (size_t)c0 * (ns > 0 ? ns : 1),
                             d_post ? d_post + (size_t)c0 * nstrips * k : nullptr, d_info + c0, st);
        if (rc) return bail(rc);
        if (pack) {
            const size_t cnt = (size_t)np * k * n;
            pack_real_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_vec + (size_t)c0 * k * n, d_pack + (size_t)c0 * k * n, cnt, d_flag);
            LAUNCHED();
        }
        cudaEvent_t evt;
        if (events.add(&evt)) return bail(fail(SCQED_ECUDA, "cudaEventCreate failed"));
        CKH(cudaEventRecord(evt, st));
        // the copies of this chunk go to the copy stream at once (asynchronous: pinned destination or staging slot);
        // they run while the next chunk is solved
        rc = copy_out(c0, np, evt);
        if (rc) return rc;
    }
    int h_flag = 0;
    CKH(cudaEventRecord(t_solved, st));
    CKH(cudaEventRecord(t_copied, cs));
    if (pack) CKH(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
    CKH(cudaStreamSynchronize(st));
    CKH(cudaStreamSynchronize(cs));
    {
        float ms_solve = 0.f, ms_all = 0.f;
        if (cudaEventElapsedTime(&ms_solve, t_start, t_solved) == cudaSuccess && cudaEventElapsedTime(&ms_all, t_start, t_copied) == cudaSuccess) {
            const float tail = ms_all - ms_solve;              // copies still running after the last solve
            if (getenv("SCQED_HOST_DEBUG"))
                fprintf(stderr, "[scqed host] dev %d pts %lld chunks %zu level %d%s solve %.2f ms all %.2f ms\n", pl->device, (long long)n_pts,
                        sizes.size(), pl->host_copy_bound, reprobe ? " (reprobe)" : "", ms_solve, ms_all);
            if (!pl->host_copy_bound && tail > 0.45f * ms_solve && tail > 1.0f) pl->host_copy_bound = 1;
            else if (reprobe && tail < 0.30f * ms_solve) pl->host_copy_bound = 0;
        } else cudaGetLastError();
    }
    if (staged_per_pt) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return chunks_moved >= n_staged_chunks; });
    }
    finish_mover();
#undef CKH
    if (mover_failed) return fail(SCQED_ECUDA, "staging copy failed");
    // a plan assumed real symmetric produced a complex point (info = -2): redo the sweep in complex mode, once
    if (pl->s1_real_mode == 1) {
        bool redo = false;
        for (long long q = 0; q < n_pts; ++q) if (info[q] == -2) { redo = true; break; }
        if (redo) {
            pl->s1_real_mode = 0;
            defer_guard.p->defer_real_check = defer_guard.old;
            return scqed_sweep_run_host(pl, o, params, n_pts, evals, evecs, scalars, post, info);
        }
    }
    if (pack && h_flag) { g_err = "real_vectors: some eigenvector has a non-zero imaginary part"; return SCQED_ECOMPLEX; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// probes
// ------------------------------------------------------------------------------------------------
template <typename K>
static int run_probe(int device, K kernel, double flops_per_thread_iter, double *out) {
    if (!out) return fail(SCQED_EINVAL, "null argument");
    ON_DEVICE(device);
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
