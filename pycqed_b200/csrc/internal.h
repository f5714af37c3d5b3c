// internal.h -- shared by the translation units of libscqed.so (api.cu, small.cu, dense.cu, matfree.cu):
// error reporting, launch accounting, per-kernel timing, device buffers, the plan object and the host-side
// entry points of each solver family.  Not part of the public ABI (that is include/scqed.h).
#pragma once
#include "../../include/scqed.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "resonator.cuh"

using namespace scqed;


using namespace scqed;

// ------------------------------------------------------------------------------------------------
// error handling / bookkeeping
// ------------------------------------------------------------------------------------------------
extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;
extern std::atomic<long long> g_mf_flops;     // algorithmic flops of those mat-vecs (scqed_counter_read(2))
extern std::atomic<long long> g_matvecs;      // vector mat-vecs issued by the matrix-free filters (scqed_counter_read(1))

int fail(int code, const char *fmt, ...);

// Opt a kernel into the device's maximum dynamic shared memory (the opt-in limit minus the kernel's static shared
// memory).  Always the same value for a given kernel and device: the attribute is per function, and host threads driving
// different engines launch the same kernel with different footprints, so a per-launch value could undercut another thread.
template <typename Kern>
static inline cudaError_t set_max_dynamic_smem(Kern kern, size_t smem_optin) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin - (int)fa.sharedSizeBytes);
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
extern std::atomic<int> g_timing;
struct TimedSpan { cudaEvent_t e0, e1; int tag; };
extern std::vector<TimedSpan> g_spans;         // guarded by g_spans_mu (several host threads may drive several GPUs)
extern std::mutex g_spans_mu;
struct TimingScope {                           // records e0 now and e1 when it goes out of scope
    cudaStream_t st; cudaEvent_t e1 = nullptr;
    TimingScope(cudaStream_t st_, int tag) : st(st_) {
        if (!g_timing.load()) return;
        TimedSpan sp; sp.tag = tag;
        if (cudaEventCreate(&sp.e0) != cudaSuccess) return;
        if (cudaEventCreate(&sp.e1) != cudaSuccess) { cudaEventDestroy(sp.e0); return; }
        cudaEventRecord(sp.e0, st);
        e1 = sp.e1;
        std::lock_guard<std::mutex> lk(g_spans_mu);
        g_spans.push_back(sp);
    }
    void end() { if (e1) { cudaEventRecord(e1, st); e1 = nullptr; } }
    ~TimingScope() { end(); }
};

// Every entry point works on the plan's device and leaves the caller's current device untouched
// (the caller is typically PyTorch, whose notion of the current device must not change behind its back).
struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); ok = false; }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev) DeviceGuard dev_guard_(dev); if (!dev_guard_.ok) return fail(SCQED_ECUDA, "cudaSetDevice(%d) failed", (int)(dev))

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
    long long host_calls = 0, host_copy_pts = -1;
    int host_copy_bound = 0;           // scqed_sweep_run_host: the previous call spent longer draining its device -> host copies than
                                       // solving (several GPUs sharing the host's PCIe / memory bandwidth): cut into more chunks
    bool defer_real_check = false;     // scqed_sweep_run_host checks the "turned complex" flags itself, once, after its last chunk
    DevBuf s1work, s1vec, auxc, resvec, dcoef, s1ladder, s1slots, dtg;
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
    // library-owned page-locked staging ring of scqed_sweep_run_host (two slots), kept across calls
    struct HostStage {
        void *p = nullptr; size_t cap = 0;
        int ensure(size_t bytes) {
            if (bytes <= cap) return 0;
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; }
            cap = bytes;
            return 0;
        }
        void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    } hstage;
};



// ---- host-side entry points of the solver families (one translation unit each) --------------------------
#define S1_NOT_REAL 77001       // internal: 128 < n <= 224 only fits the shared-memory path when H is real
#define MF_BMAX_HOST 32         // == MF_BMAX (matfree.cuh): largest block of the subspace iteration

// api.cu
int run_coef(scqed_plan *pl, const double *params_dev, long long n_pts, cplx *coef_dev, double *scal_dev,
             cudaStream_t st, cplx *aux_dev = nullptr, bool dense = false);
int run_assemble(scqed_plan *pl, const cplx *coef_dev, long long n_pts, int tidyup, cplx *H, cudaStream_t st);
// small.cu (S1: H in shared memory)
bool small_fits(const scqed_plan *pl, int n, int k, int want_vec, int n_terms);
bool split_fits(const scqed_plan *pl, int n, int k, int n_terms, int real_hint = -1);
int run_small_any(scqed_plan *pl, const PlanDev &P, const cplx *coef, const cplx *Hin, const double *scal,
                  long long n_pts, int n, int k, int want_vec, int tidyup, const ResonatorArgs &res,
                  double *evals, cplx *evecs, double *post, int *info, cudaStream_t st, bool monolithic);
int run_resonator_generic(const PlanDev &P, const ResonatorArgs &res, const double *scal, const double *evals,
                          const cplx *vecs, long long n_pts, int n, int k, double *post, int *info, cudaStream_t st);
// dense.cu (S2: batches of dense matrices through HBM)
int s1_tridiag_eig_batch(scqed_plan *pl, double *dd, double *ee, int n, int k, long long B, int want_vec, double *evals,
                         const double **zw_out, long long *G_out, int *info, cudaStream_t st);
int dense_eigh_resident(scqed_plan *tmp, cplx *H_dev, int n, long long batch, int k, int want_vec, double *evals,
                        cplx *evecs, int *info, int solver, cudaStream_t st);
int dense_sweep(scqed_plan *pl, const cplx *coef, long long n_pts, int k, int want_vec, int tidyup, int solver, bool allow_tiled,
                double *evals, cplx *vecs, int *info, cudaStream_t st);
// matfree.cu (S3: Chebyshev-filtered subspace iteration on the Kronecker mat-vec)
int run_matfree(scqed_plan *pl, const cplx *coef_all, long long n_pts, int k, int want_vec, double *evals,
                cplx *evecs, int *info, cudaStream_t st, int deg, int maxit, double tol);
int run_aux_matel(const PlanDev &P, const cplx *auxcoef, const cplx *evecs, long long n_pts, int k, cplx *out, cudaStream_t st);
