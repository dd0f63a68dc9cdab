// Stage-layout bound rows from the reference's flat lbx/ubx/lbg/ubg arrays
// (centralized_six_robots_implementation.py:349-352), relaxed by IPOPT's bound_relax_factor.
// One call per (bound set, stage k, lane); shared by prep_bounds_kernel and the CPU emulation.
#pragma once
#include "nmpc_internal.h"

#ifndef NMPC_HD
#ifdef __CUDACC__
#define NMPC_HD __host__ __device__
#else
#define NMPC_HD
#endif
#endif
#ifndef NMPC_INF
#define NMPC_INF ((double)INFINITY)
#endif

NMPC_HD inline double nmpc_relax_lo(double v, double f) { return v > -NMPC_INF ? v - f * fmax(1.0, fabs(v)) : -NMPC_INF; }
NMPC_HD inline double nmpc_relax_hi(double v, double f) { return v < NMPC_INF ? v + f * fmax(1.0, fabs(v)) : NMPC_INF; }

// rows: [NMPC_BR_COUNT][S][lw] (or, stage_major, [S][NMPC_BR_COUNT][lw]) of this bound set (lw = 32 or 64 lanes per instance).  Returns 0 or a negative NMPC_E* code.
// nobs / family: rows per block = pair rows + Nr * nobs obstacle rows; family 1 (static obstacles,
// first_scenario_mpc_obstacle_avoidance.py:150) has no inequality rows in block 0, so block k starts at ns + (k-1)(ns+M).
NMPC_HD inline int nmpc_prep_bounds_elem(int Nr, int N, double relax, const double *lbx, const double *ubx,
                                         const double *lbg, const double *ubg, int k, int lane, int lw, double *rows,
                                         int nobs = 0, int family = 0, int stage_major = 0)
{
    const int ns = 3 * Nr, nc = 2 * Nr, nz = 5 * Nr, M = Nr * (Nr - 1) / 2 + Nr * nobs, S = N + 1;
    const long long goff = family ? (k == 0 ? 0 : ns + (long long)(k - 1) * (ns + M)) : (long long)k * (ns + M);
    // row x of stage k: one array per row ([x][S][lw], the dense-block path) or one record of NMPC_BR_COUNT rows per stage
    // ([S][x][lw], the warp-per-instance path, which stages [BL, BU] and [CE, DL, DU] as contiguous ranges)
    const long long rs = stage_major ? (long long)lw : (long long)S * lw;
    const long long e = stage_major ? (long long)k * NMPC_BR_COUNT * lw + lane : (long long)k * lw + lane;
    int err = 0;
    double lo = -NMPC_INF, hi = NMPC_INF;
    if (lane < ns) { lo = lbx[k * ns + lane]; hi = ubx[k * ns + lane]; }
    else if (lane < nz && k < N) { lo = lbx[ns * S + k * nc + (lane - ns)]; hi = ubx[ns * S + k * nc + (lane - ns)]; }
    if (!(lo <= hi)) err = NMPC_EBOUNDS;
    else if (lo == hi) err = NMPC_ENOTSUP;  // fixed variables are not part of this path
    rows[NMPC_BR_BL * rs + e] = nmpc_relax_lo(lo, relax);
    rows[NMPC_BR_BU * rs + e] = nmpc_relax_hi(hi, relax);
    double ce = 0.0;
    if (lane < ns) {
        double l = lbg[goff + lane], u = ubg[goff + lane];
        if (!(l == u) || !(l > -NMPC_INF && l < NMPC_INF)) err = err ? err : NMPC_ENOTSUP;  // dynamics rows are equalities
        ce = l;
    }
    rows[NMPC_BR_CE * rs + e] = ce;
    double dl = -NMPC_INF, du = NMPC_INF;
    if (lane < M && !(family && k == 0)) {
        dl = lbg[goff + ns + lane]; du = ubg[goff + ns + lane];
        if (!(dl <= du)) err = err ? err : NMPC_EBOUNDS;
        else if (dl == du) err = err ? err : NMPC_ENOTSUP;  // equality on a distance row
    }
    rows[NMPC_BR_DL * rs + e] = nmpc_relax_lo(dl, relax);
    rows[NMPC_BR_DU * rs + e] = nmpc_relax_hi(du, relax);
    return err;
}
