/*
 * scqed.h -- C ABI of the B200 sweep engine (libscqed.so).
 *
 * The reference (pyscqed 0.13.0) has no FFI/plugin interface; its hot path sits behind the
 * Python class API of pyscqed/numerical_system.py.  The entry points below are what a binding
 * of that path needs; each one names the reference function it replaces.  Conventions:
 *   - plain pointers and sizes only, no C++ or torch types;
 *   - every function returns 0 on success or a negative SCQED_E* code (one positive code, SCQED_ECOMPLEX,
 *     asks the caller to repeat a real_vectors request with a complex buffer), and never throws;
 *     scqed_last_error() returns a thread-local message for the last failure;
 *   - all *_dev pointers are device pointers owned by the caller (e.g. torch tensors),
 *     complex128 data is interleaved (re, im) doubles;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - a plan is immutable after creation; one in-flight call per (plan, stream).
 */
#ifndef SCQED_H
#define SCQED_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCQED_OK            0
#define SCQED_EINVAL       -1   /* bad argument / inconsistent plan description            */
#define SCQED_ECUDA        -2   /* CUDA runtime error (message in scqed_last_error)          */
#define SCQED_ENOMEM       -3   /* workspace allocation failed                               */
#define SCQED_EUNSUPPORTED -4   /* configuration outside what this build implements         */

typedef struct scqed_plan scqed_plan;

/*
 * Plan description = the three tables that replace the reference's full-space QuTiP operators
 * (numerical_system.py:185-253 getExpandedOperatorsMap) and its per-point sympy substitution
 * (numerical_system.py:395-437 _postsub):
 *   factor table : n_factors dense d x d complex matrices (row-major), factor f belongs to node
 *                  factor_node[f] and starts at factor_data + 2*factor_offset[f] doubles
 *   term table   : H(p) = sum_t c_t(p) * kron_k F[term_factors[t*n_nodes + k]]   (-1 = identity)
 *   program      : SSA bytecode, instruction i = (op, a, b) writes register i; c_t = reg[coef_re[t]]
 *                  + i*reg[coef_im[t]] (-1 = 0); scalar outputs reg[scalar_regs[s]].
 *                  Opcodes: see pycqed_b200/coefprog.py.
 */
typedef struct {
    int32_t n_nodes;
    const int32_t *dims;            /* [n_nodes] first node = slowest Hilbert index          */
    int32_t n_factors;
    const int32_t *factor_node;     /* [n_factors]                                           */
    const int64_t *factor_offset;   /* [n_factors] in complex elements                       */
    const double  *factor_data;     /* interleaved complex128                                */
    int32_t n_terms;
    const int32_t *term_factors;    /* [n_terms * n_nodes]                                   */
    int32_t n_swept;
    int32_t n_instr;
    const int32_t *instr;           /* [n_instr * 3]                                         */
    int32_t n_consts;
    const double  *consts;
    const int32_t *coef_re;         /* [n_terms]                                             */
    const int32_t *coef_im;         /* [n_terms]                                             */
    int32_t n_scalar;
    const int32_t *scalar_regs;     /* [n_scalar]                                            */
    /* Auxiliary operators for the Current / Voltage evaluables (numerical_system.py:596-694):
     * n_aux extra terms (same factor table), coefficient registers aux_re/aux_im, and n_pairs
     * requested matrix elements, pair q = (term_begin, term_end, i, j): <i| sum_t c_t kron F |j>. */
    int32_t n_aux;
    const int32_t *aux_term_factors; /* [n_aux * n_nodes]                                    */
    const int32_t *aux_re;          /* [n_aux]                                               */
    const int32_t *aux_im;          /* [n_aux]                                               */
    int32_t n_pairs;
    const int32_t *pairs;           /* [n_pairs * 4]                                         */
    /* Optional dense real term table for the fused small-n solver (n <= 128): Re H = sum_g r[dense_re[g]] G_g,
     * Im H = sum_g r[dense_im[g]] G_g with G_g real n x n, element (i, j) at dense_tab[g*n*n + j*n + i];
     * register -1 = identically zero.  n_dense = 0 disables it (the term table is walked per element). */
    int32_t n_dense;
    int32_t n_dense_diag;           /* the last n_dense_diag matrices are diagonal, stored as n-vectors */
    int32_t dense_tridiag;          /* 1: every G_g is tridiagonal, i.e. H(p) is Hermitian tridiagonal (single-node
                                       charge basis): the Householder stage is skipped                        */
    const double *dense_tab;        /* [(n_dense - n_dense_diag) * n * n + n_dense_diag * n] */
    const int32_t *dense_re;        /* [n_dense]                                             */
    const int32_t *dense_im;        /* [n_dense]                                             */
    const double *dense_gmax;       /* [n_dense] max |G_g|                                   */
    /* Optional entry list for scqed_assemble_dense (pattern of H expanded once): entry e is position
     * ent_key[e] = i*n + j (sorted, distinct) and H[i][j] = sum over pairs q in [ent_ptr[e], ent_ptr[e+1]) of
     * c[ent_term[q]] * ent_val[q].  ent_dense != 0: the pattern covers >= 25 % of the matrix (one thread per
     * element, no zero fill pass).  n_entries = 0: walk the term table per element. */
    int32_t n_entries;
    int32_t ent_dense;
    const int64_t *ent_key;         /* [n_entries]                                           */
    const int32_t *ent_ptr;         /* [n_entries + 1]                                       */
    const int32_t *ent_term;        /* [ent_ptr[n_entries]]                                  */
    const double *ent_val;          /* complex [ent_ptr[n_entries]]                          */
    /* Nodes solved in a rotated basis (oscillator nodes whose impedance is swept, numerical_system.py:378-393:
     * the plan's factors of node xform_node[i] are expressed in the eigenbasis of a + a^dagger).  Eigenvectors
     * are rotated back on that tensor mode with the real orthogonal d x d matrix W_i (row-major,
     * concatenated in xform_data) before they are returned. */
    int32_t n_xform;
    const int32_t *xform_node;      /* [n_xform]                                             */
    const double *xform_data;       /* concatenated W_i, d_i * d_i doubles each              */
} scqed_plan_desc;

/* Options of one sweep (numerical_system.py:790-804 setDiagConfig + :868-875 addEvaluation). */
typedef struct {
    int32_t k;                /* eigvalues: number of lowest eigenpairs                      */
    int32_t want_vectors;     /* get_vectors                                                 */
    int32_t solver;           /* 0 auto, 1 fused small (H in smem), 2 dense blocked (HBM, DMMA),
                                 3 dense unblocked (HBM, reference implementation),
                                 4 matrix-free Chebyshev-filtered subspace iteration,
                                 5 small path as ONE fused kernel (1 = phase pipeline)       */
    int32_t tidyup;           /* 1: threshold H entries at 1e-12 like Qobj.tidyup (:594)     */
    /* Resonator evaluable (numerical_system.py:736-784 getResonatorResponse); needs vectors */
    int32_t resonator;        /* 0/1                                                         */
    int32_t res_nmax;         /* nmax (strips = nmax + 1 + k)                                */
    int32_t res_op_factor;    /* factor id of the coupled node's charge operator             */
    int32_t res_gc;           /* scalar slot of g_C                                          */
    int32_t res_frl;          /* scalar slot of the loaded resonator frequency               */
    int32_t res_qb;           /* scalar slot of the node's charge bias                       */
    int32_t raw_basis;        /* 1: leave eigenvectors in the plan's (rotated) basis; the caller runs
                                 scqed_matrix_elements on them and then scqed_back_transform      */
    int32_t real_vectors;     /* scqed_sweep_run_host only: the caller's `evecs` buffer is float64 [n_pts, k, n]; the
                                 real parts are packed on the device and only they cross PCIe.  Valid when every
                                 eigenvector is exactly real (real symmetric H: fluxonium, RF-SQUID ...); otherwise the
                                 call returns SCQED_ECOMPLEX (1) and must be repeated with a complex buffer.        */
    int32_t evecs_on_device;  /* scqed_sweep_run_host only: `evecs` is a DEVICE pointer (complex [n_pts, k, n]); the eigenvectors
                                 are written there and stay on the GPU (fetched lazily by the caller, or consumed by
                                 scqed_matrix_elements); everything else is copied to the host buffers as usual         */
    int32_t reserved[3];
} scqed_sweep_opts;

#define SCQED_ECOMPLEX 1   /* real_vectors requested but some eigenvector has a non-zero imaginary part */

/* Create / destroy.  Uploads the tables to `device`.  No entry point of this library changes the caller's
 * current CUDA device (it is saved and restored). */
int  scqed_plan_create(const scqed_plan_desc *desc, int device, scqed_plan **plan_out);
void scqed_plan_destroy(scqed_plan *plan);
int  scqed_plan_hilbert_dim(const scqed_plan *plan);

/* _postsub replacement (numerical_system.py:395-437): coefficients of all points.
 * params_dev [n_pts, n_swept] row-major; coef_dev complex [n_pts, n_terms]; scalars_dev
 * [n_pts, n_scalar] (may be NULL when n_scalar == 0). */
int scqed_eval_coefficients(scqed_plan *plan, const double *params_dev, int64_t n_pts,
                            double *coef_dev, double *scalars_dev, void *stream);

/* getHamiltonian replacement (numerical_system.py:509-594): dense H of every point,
 * H_dev complex [n_pts, n, n] row-major. */
int scqed_assemble_dense(scqed_plan *plan, const double *params_dev, int64_t n_pts, int32_t tidyup,
                         double *H_dev, void *stream);

/* diagonalize / util.diagDenseH replacement (numerical_system.py:809-810, util.py:127-176):
 * lowest k eigenpairs of `batch` Hermitian matrices (row-major, only consistent Hermitian input
 * is supported; H_dev is overwritten).  evals_dev [batch, k] ascending; evecs_dev complex
 * [batch, k, n] unit-norm (NULL = values only); info_dev [batch] (0 = converged).
 * Solver by size and content (solver = 0): H in shared memory (n <= 128, real symmetric <= 224); real symmetric batches of
 * >= 48 matrices up to n = 3400 on the tile-packed lower triangle, one persistent CTA per matrix (dense_tiled.cuh); the
 * complex blocked solver otherwise.  The call is synchronous on `stream`; calls on one device are serialised, and a
 * per-device context keeps the workspaces between calls (released when they exceed 16 GB). */
int scqed_eigh_batched(int device, double *H_dev, int32_t n, int64_t batch, int32_t k,
                       double *evals_dev, double *evecs_dev, int32_t *info_dev, int32_t solver,
                       void *stream);

/* paramSweep replacement (numerical_system.py:885-1049): assemble + diagonalise (+ post-ops) for
 * all points.  Device-buffer form.  post_dev: resonator strips [n_pts, nmax+1+k, k] or NULL. */
int scqed_sweep_run(scqed_plan *plan, const scqed_sweep_opts *opts, const double *params_dev,
                    int64_t n_pts, double *evals_dev, double *evecs_dev, double *scalars_dev,
                    double *post_dev, int32_t *info_dev, void *stream);

/* Same with HOST buffers: copies params in, results out.  This is the call the Python drop-in makes for one
 * GPU shard.  The sweep is cut into 2-3 chunks; the results of chunk c travel while chunk c+1 is solved.
 * Page-locked caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory; detected per buffer with
 * cudaPointerGetAttributes) are written directly by the DMA engine.  PAGEABLE caller buffers go through a
 * library-owned page-locked staging ring (two slots of <= 256 MB, allocated once per plan and kept): the GPU
 * copies into a slot asynchronously and a host thread moves the slot into the caller's array (multi-threaded
 * memcpy) while the next chunk is solved.  The caller's current CUDA device is left unchanged.  Thread-safe
 * for different plans (one host thread per GPU is how a single process drives several GPUs). */
int scqed_sweep_run_host(scqed_plan *plan, const scqed_sweep_opts *opts, const double *params,
                         int64_t n_pts, double *evals, double *evecs, double *scalars,
                         double *post, int32_t *info);

/* getCurrentMatrixElement / getVoltageMatrixElement replacement (numerical_system.py:656-694):
 * matrix elements of the plan's auxiliary operators between the eigenvectors of each point.
 * evecs_dev complex [n_pts, k, n] (as returned by scqed_sweep_run); out_dev complex [n_pts, n_pairs]
 * (interleaved re, im): <i|Op|j>.  The Current / Voltage evaluables keep the real part (:665,:692);
 * util.pauliCoefficients (util.py:322-328) uses the complex value. */
int scqed_matrix_elements(scqed_plan *plan, const double *params_dev, int64_t n_pts, int32_t k,
                          const double *evecs_dev, double *out_dev, void *stream);

/* Rotate eigenvectors [n_vectors, n] (complex, in place) from the plan's basis back to the reference's
 * Fock basis on every node listed in xform_node.  A no-op for plans without rotated nodes. */
int scqed_back_transform(scqed_plan *plan, double *evecs_dev, int64_t n_vectors, void *stream);

/* util.pauliCoefficients replacement (util.py:300-391): local-basis reduction of the two lowest
 * levels.  e01_dev [n_pts, 2] = (E0, E1); op_dev complex [n_pts, 4] = the computational-basis
 * operator in the {|0>, |1>} eigenbasis, row-major (m00, m01, m10, m11); h_dev [n_pts, 3] =
 * (hx, hy, hz).  float64 throughout (the reference rounds to complex64). */
int scqed_pauli_coefficients(int device, const double *e01_dev, const double *op_dev, int64_t n_pts,
                             double *h_dev, void *stream);

/* Matrix elements of ONE fixed dense operator between the first m (<= 4) eigenvectors of every point: the
 * fixed-operator branch of util.pauliCoefficients (util.py:320-342, `basis_op.matrix_element(V[a], V[b])` in a
 * Python loop over sweep points).  op_dev complex [n, n] row-major; evecs_dev complex [n_pts, m, n];
 * out_dev complex [n_pts, m, m] = <V_a| Op |V_b>. */
int scqed_dense_matrix_elements(int device, const double *op_dev, const double *evecs_dev, int32_t n, int32_t m,
                                int64_t n_pts, double *out_dev, void *stream);

/* Kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t scqed_launch_count(void);

/* Per-kernel device timing for roofline reporting: when enabled, CUDA events are recorded on the
 * launching stream around the dominant kernel of a path (tag 1: small-path tridiagonalisation, real
 * symmetric; tag 2: same, complex Hermitian).  scqed_timing_read sums and clears the spans. */
void scqed_timing_enable(int on);
int scqed_timing_read(int tag, double *total_ms, int64_t *count);   /* tags: 1 / 2 real / complex tridiagonalisation,
                                                                        3 dense assembly, 4 direct tridiagonal,
                                                                        5 / 6 / 7 matrix-free filter (stencil /
                                                                        shared-memory / global-memory form),
                                                                        8 dense (HBM) tridiagonal reduction */
/* Work counters for roofline reporting, cleared by the read.  which = 1: vector mat-vecs issued by the matrix-free
 * filters (active points x block size x filter degree); which = 2: algorithmic flops of those mat-vecs (8 n x complex
 * MACs per row of the term walk, or n (d0 + d1) real MACs of the combined two-node form). */
int scqed_counter_read(int which, int64_t *value);

/* FP64 peak probe: runs a register-resident DMMA (mma.sync m8n8k4 f64) loop on every SM and
 * returns TFLOP/s; the roofline denominator for the dense tridiagonalisation. */
int scqed_probe_fp64_tensor(int device, double *tflops_out);
int scqed_probe_fp64_fma(int device, double *tflops_out);

const char *scqed_last_error(void);
const char *scqed_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SCQED_H */
