// Warp-collective primitives used by solver_body.cuh (CUDA implementation).
// tests/emul/ provides a host implementation of the same interface so that the very same solver
// source can be stepped through on a CPU (32 fibers per warp) in the `-m "not gpu"` tests.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define NMPC_DEV __device__ __forceinline__
#define NMPC_HD __host__ __device__
#define NMPC_PASS __device__ __noinline__   // one copy of each solver pass: keeps the hot code inside the instruction cache
#define NMPC_UNROLL _Pragma("unroll")
#define NMPC_NOUNROLL _Pragma("unroll 1")

namespace wp {
NMPC_DEV int lane() { return threadIdx.x & 31; }
NMPC_DEV int team_lane(int lw) { return threadIdx.x & (lw - 1); }   // lane inside a 32- or 64-thread team
NMPC_DEV void sync_cta() { __syncthreads(); }
// barrier of one 64-thread team (named barrier 1 + team index; barrier 0 stays free for the whole CTA)
NMPC_DEV void sync_team64() { asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory"); }
NMPC_DEV int cta_count(bool pred) { return __syncthreads_count(pred); }   // CTA barrier + number of threads with pred
NMPC_DEV void sync() { __syncwarp(); }
NMPC_DEV double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
NMPC_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
NMPC_DEV int shfl_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
NMPC_DEV bool any(bool p) { return __any_sync(0xffffffffu, p); }
NMPC_DEV bool all(bool p) { return __all_sync(0xffffffffu, p); }
NMPC_DEV int atomic_next(int *counter) { return atomicAdd(counter, 1); }
// address-space hints for pointers that travel through the solver object
extern __shared__ double nmpc_dyn_smem[];
NMPC_DEV double *shared_ptr(double *p) { return nmpc_dyn_smem + (p - (double *)nmpc_dyn_smem); }
template <class Tp> NMPC_DEV Tp *global_ptr(Tp *p) { __builtin_assume(__isGlobal(p)); return p; }
// reciprocal of a positive, finite, normal double: MUFU seed + two Newton steps (<= 1 ulp), no range checks
NMPC_DEV double rcp_pos(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}
// 1 / sqrt(d) for a positive, finite, normal d: MUFU seed (2^-22) and two Newton steps, no special-case branches
NMPC_DEV double rsqrt_pos(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-d * y, y, 1.0);
    return fma(0.5 * y, e, y);
}
// L2 prefetch of a line that a later stage of the same pass will read (the per-warp scratch does not fit L1)
NMPC_DEV void prefetch(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// asynchronous 16-byte global -> shared copy (LDGSTS, L2-only caching): stages the scratch rows of the NEXT Riccati
// stage while the current one is being processed, without holding registers
NMPC_DEV void cp_async16(double *smem_dst, const double *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
NMPC_DEV void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// TMA bulk copies global -> shared (cp.async.bulk, SASS UBLKCP.S.G), completion on an mbarrier in shared memory.  One elected lane
// arms the barrier with the byte count of a stage (expect_tx) and issues one copy per contiguous range of rows; every lane of the
// team waits on the barrier's phase parity.  The staged rows were written with ordinary stores: the issuing lane orders them
// before the bulk engine's reads with a proxy fence (after the team-wide sync that already separates writers and issuer).
NMPC_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
NMPC_DEV void mbar_init(double *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the initialised barrier becomes visible to the bulk-copy engine
}
// generic-proxy global stores (the scratch rows) before async-proxy reads of them (the bulk copies).  The .global form: the
// unqualified fence.proxy.async also covers shared memory and cost 5 % of the kernel's throughput (47.9 k against 50.6 k solves/s).
NMPC_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
NMPC_DEV void bulk_expect(double *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NMPC_DEV void bulk_g2s(double *dst, const double *src, unsigned bytes, double *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
NMPC_DEV void mbar_wait(double *bar, unsigned parity)
{
    asm volatile("{\n .reg .pred p;\n NMPC_MBAR_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra NMPC_MBAR_WAIT;\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// compiler-level fence: memory accesses are not moved across it (keeps the batched loads of the sweep apart, which bounds
// their register footprint)
NMPC_DEV void sched_fence() { asm volatile("" ::: "memory"); }
// positive, finite, normal double (one integer compare on the high word; nvcc turns the two floating-point compares into ~15 integer instructions)
NMPC_DEV bool pos_normal(double v) { return (unsigned)(__double2hiint(v) - 0x00100000) < 0x7fe00000u; }
NMPC_DEV unsigned nth_set_bit(unsigned mask, int n) { return __fns(mask, 0, n + 1); }
// one shared copy of the long math routines: keeps the passes small enough for the instruction cache
__device__ __noinline__ void sincos_(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __noinline__ double log_(double x) { return log(x); }
NMPC_DEV double frexp_(double x, int *e) { return frexp(x, e); }

__device__ __noinline__ double red_sum(double v)
{
    NMPC_UNROLL
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
    return v;
}
__device__ __noinline__ double red_max(double v)
{
    NMPC_UNROLL
    for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor(v, m));
    return v;
}
__device__ __noinline__ double red_min(double v)
{
    NMPC_UNROLL
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor(v, m));
    return v;
}
}  // namespace wp
