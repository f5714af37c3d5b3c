// coef.cuh -- K1: the flat coefficient program, one thread per sweep point.
//
// Replaces the per-point sympy substitution of the reference
// (pyscqed/numerical_system.py:395-437 _postsub; pyscqed/parameters.py:1207-1240):
// every sweep-dependent scalar of H(p) is a straight-line SSA program over the swept values.
// Register i of point p lives at regs[i * n_pts + p], so all accesses of a warp are coalesced.
#pragma once
#include "common.cuh"

namespace scqed {

enum ScqedOp {
    OP_CONST = 0, OP_LOAD = 1, OP_ADD = 2, OP_SUB = 3, OP_MUL = 4, OP_DIV = 5, OP_NEG = 6,
    OP_SQRT = 7, OP_SIN = 8, OP_COS = 9, OP_EXP = 10, OP_LOG = 11, OP_POW = 12, OP_POWI = 13,
    OP_ABS = 14, OP_TAN = 15, OP_ATAN = 16, OP_TANH = 17, OP_SINH = 18, OP_COSH = 19
};

__device__ __forceinline__ double powi(double x, int n) {
    bool neg = n < 0;
    unsigned m = neg ? (unsigned)(-n) : (unsigned)n;
    double r = 1.0;
    while (m) {
        if (m & 1u) r *= x;
        x *= x;
        m >>= 1;
    }
    return neg ? 1.0 / r : r;
}

static __global__ void __launch_bounds__(128)
coef_eval_kernel(PlanDev P, const double *__restrict__ params, long long n_pts,
                 double *__restrict__ regs, cplx *__restrict__ coef, double *__restrict__ scalars,
                 cplx *__restrict__ auxcoef, double *__restrict__ dcoef) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const double *prm = params + p * P.n_swept;
    for (int i = 0; i < P.n_instr; ++i) {
        int op = __ldg(P.instr + 3 * i), a = __ldg(P.instr + 3 * i + 1), b = __ldg(P.instr + 3 * i + 2);
        double x = 0.0, y = 0.0, r;
        if (op >= OP_ADD) {
            x = regs[(long long)a * n_pts + p];
            if (op <= OP_DIV || op == OP_POW) y = regs[(long long)b * n_pts + p];
        }
        switch (op) {
            case OP_CONST: r = __ldg(P.consts + a); break;
            case OP_LOAD:  r = prm[a]; break;
            case OP_ADD:   r = x + y; break;
            case OP_SUB:   r = x - y; break;
            case OP_MUL:   r = x * y; break;
            case OP_DIV:   r = x / y; break;
            case OP_NEG:   r = -x; break;
            case OP_SQRT:  r = sqrt(x); break;
            case OP_SIN:   r = sin(x); break;
            case OP_COS:   r = cos(x); break;
            case OP_EXP:   r = exp(x); break;
            case OP_LOG:   r = log(x); break;
            case OP_POW:   r = pow(x, y); break;
            case OP_POWI:  r = powi(x, b); break;
            case OP_ABS:   r = fabs(x); break;
            case OP_TAN:   r = tan(x); break;
            case OP_ATAN:  r = atan(x); break;
            case OP_TANH:  r = tanh(x); break;
            case OP_SINH:  r = sinh(x); break;
            case OP_COSH:  r = cosh(x); break;
            default:       r = nan(""); break;
        }
        regs[(long long)i * n_pts + p] = r;
    }
    for (int t = 0; t < P.n_terms; ++t) {
        int ir = __ldg(P.coef_re + t), ii = __ldg(P.coef_im + t);
        double re = ir >= 0 ? regs[(long long)ir * n_pts + p] : 0.0;
        double im = ii >= 0 ? regs[(long long)ii * n_pts + p] : 0.0;
        coef[p * P.n_terms + t] = cmake(re, im);
    }
    for (int s = 0; s < P.n_scalar; ++s)
        scalars[p * P.n_scalar + s] = regs[(long long)__ldg(P.scalar_regs + s) * n_pts + p];
    if (dcoef) {
        const int ng = P.n_dense;
        for (int g = 0; g < ng; ++g) {
            int ir = __ldg(P.dense_re + g), ii = __ldg(P.dense_im + g);
            dcoef[p * 2 * ng + g] = ir >= 0 ? regs[(long long)ir * n_pts + p] : 0.0;
            dcoef[p * 2 * ng + ng + g] = ii >= 0 ? regs[(long long)ii * n_pts + p] : 0.0;
        }
    }
    if (auxcoef) {
        for (int t = 0; t < P.n_aux; ++t) {
            int ir = __ldg(P.aux_re + t), ii = __ldg(P.aux_im + t);
            double re = ir >= 0 ? regs[(long long)ir * n_pts + p] : 0.0;
            double im = ii >= 0 ? regs[(long long)ii * n_pts + p] : 0.0;
            auxcoef[p * P.n_aux + t] = cmake(re, im);
        }
    }
}

}  // namespace scqed
