// Host implementation of the wp:: interface of csrc/warp_prims.cuh -- TEST INFRASTRUCTURE ONLY.
// A warp is emulated by 32 ucontext fibres that run one after the other between warp-wide
// synchronisation points, so csrc/solver_body.cuh (the product's device source) can be executed
// and debugged on a CPU.  Running lanes strictly one after another also makes most missing
// __syncwarp() bugs show up as wrong results (and `reverse` flips the lane order).
#pragma once
#include <math.h>
#include <string.h>
#include <ucontext.h>

#define NMPC_DEV inline
#define NMPC_HD
#define NMPC_PASS inline
#define NMPC_UNROLL
#define NMPC_NOUNROLL

namespace wp {
struct Emu {
    ucontext_t main, ctx[64];
    int cur;
    bool done[64];
    double xd[64];
    int xi[64];
    bool xb[64];
};
extern thread_local Emu *emu;
inline int lane() { return emu->cur & 31; }
inline int team_lane(int lw) { return emu->cur & (lw - 1); }
inline void sync() { swapcontext(&emu->ctx[emu->cur], &emu->main); }
inline void sync_cta() { sync(); }
inline void sync_team64() { sync(); }
inline int cta_count(bool pred) { return pred ? 1 : 0; }
inline double shfl(double v, int src) { emu->xd[emu->cur] = v; sync(); double r = emu->xd[(emu->cur & 32) | (src & 31)]; sync(); return r; }
inline double shfl_xor(double v, int m) { return shfl(v, (emu->cur & 31) ^ m); }
inline int shfl_i(int v, int src) { emu->xi[emu->cur] = v; sync(); int r = emu->xi[(emu->cur & 32) | (src & 31)]; sync(); return r; }
inline bool any(bool p) { emu->xb[emu->cur] = p; sync(); bool r = false; for (int i = 0; i < 32; i++) r = r || emu->xb[i]; sync(); return r; }
inline bool all(bool p) { emu->xb[emu->cur] = p; sync(); bool r = true; for (int i = 0; i < 32; i++) r = r && emu->xb[i]; sync(); return r; }
inline int atomic_next(int *c) { return (*c)++; }
inline double *shared_ptr(double *p) { return p; }
template <class Tp> inline Tp *global_ptr(Tp *p) { return p; }
inline double rcp_pos(double d) { return 1.0 / d; }
inline void sched_fence() {}
inline bool pos_normal(double v) { return v >= 2.2250738585072014e-308 && v < INFINITY; }
inline void prefetch(const void *) {}
inline void cp_async16(double *dst, const double *src) { dst[0] = src[0]; dst[1] = src[1]; }
inline void cp_async_wait() {}
inline void mbar_init(double *) {}
inline void fence_proxy_async() {}
inline void bulk_expect(double *, unsigned) {}
inline void bulk_g2s(double *dst, const double *src, unsigned bytes, double *) { memcpy(dst, src, bytes); }
inline void mbar_wait(double *, unsigned) {}
inline unsigned nth_set_bit(unsigned mask, int n) { for (unsigned b = 0; b < 32; b++) if (mask >> b & 1u) { if (n == 0) return b; n--; } return 0xffffffffu; }
inline double log_(double x) { return ::log(x); }
inline double frexp_(double x, int *e) { return ::frexp(x, e); }
inline void sincos_(double x, double *s, double *c) { ::sincos(x, s, c); }
inline double red_sum(double v) { for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m); return v; }
inline double red_max(double v) { for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor(v, m)); return v; }
inline double red_min(double v) { for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor(v, m)); return v; }
}  // namespace wp
