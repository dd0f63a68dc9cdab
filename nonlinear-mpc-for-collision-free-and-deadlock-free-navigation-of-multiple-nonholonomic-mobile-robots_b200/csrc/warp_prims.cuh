// Warp-collective primitives used by solver_body.cuh (CUDA implementation).
// tests/emul/ provides a host implementation of the same interface so that the very same solver
// source can be stepped through on a CPU (32 fibers per warp) in the `-m "not gpu"` tests.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define NMPC_DEV __device__ __forceinline__
#define NMPC_HD __host__ __device__
#define NMPC_UNROLL _Pragma("unroll")

namespace wp {
NMPC_DEV int lane() { return threadIdx.x & 31; }
NMPC_DEV void sync() { __syncwarp(); }
NMPC_DEV double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
NMPC_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
NMPC_DEV int shfl_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
NMPC_DEV bool any(bool p) { return __any_sync(0xffffffffu, p); }
NMPC_DEV bool all(bool p) { return __all_sync(0xffffffffu, p); }
NMPC_DEV int atomic_next(int *counter) { return atomicAdd(counter, 1); }
NMPC_DEV void sincos_(double x, double *s, double *c) { sincos(x, s, c); }

NMPC_DEV double red_sum(double v)
{
    NMPC_UNROLL
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
    return v;
}
NMPC_DEV double red_max(double v)
{
    NMPC_UNROLL
    for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor(v, m));
    return v;
}
NMPC_DEV double red_min(double v)
{
    NMPC_UNROLL
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor(v, m));
    return v;
}
}  // namespace wp
