// probe.cuh -- FP64 peak probes (roofline denominators that MEASURED_PEAKS.json does not carry).
// Register-resident loops, no memory traffic: the DMMA probe issues mma.sync.m8n8k4.f64 (SASS
// DMMA.8x8x4, the only FP64 tensor instruction on sm_100a -- tcgen05 has no f64 kind), the FMA
// probe issues dependent-free DFMA chains.
#pragma once
#include "common.cuh"

namespace scqed {

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

static __global__ void __launch_bounds__(256)
probe_dmma_kernel(int iters, double *out) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma_8x8x4(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

static __global__ void __launch_bounds__(256)
probe_dfma_kernel(int iters, double *out) {
    double x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], y, x);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

}  // namespace scqed
