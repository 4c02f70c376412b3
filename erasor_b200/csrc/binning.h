// binning.h -- exact polar bin index of one point (host + device).
//
// Replaces the per-point arithmetic of ERASOR::voi2r_pod / xy2theta / xy2radius
// (reference src/offline_map_updater/src/erasor.cpp:11-21, 100-144):
//
//     if (pt.z < max_h && pt.z > min_h) {
//         double r = sqrt(pow(x,2) + pow(y,2));               // double
//         if (r <= max_r) {
//             double theta = (y >= 0) ? atan2(y,x) : 2*PI + atan2(y,x);   // PI = 3.1415926535
//             sector = min(int(theta / sector_size), S-1);  ring = min(int(r / ring_size), R-1);
//
// The result must be BIT-EXACT, but a double-precision sqrt + atan2 + two divisions per point
// would make the kernel issue-bound far below the HBM roofline.  Instead every comparison the
// reference makes is turned into an equivalent comparison against a threshold computed ONCE on
// the host (binning_tables.cpp):
//
//   * z window: float thresholds z_lo / z_hi with  (double)z < max_h  <=>  z < z_hi  etc.
//   * range and ring: s = x*x + y*y is exact in double (24-bit mantissas squared; one fma);
//     sqrt and the division are monotone, so  r <= max_r  <=>  s <= s_max  and
//     ring >= k  <=>  s >= ring_thr[k], with s_max / ring_thr[k] found by bisection over doubles
//     on the reference's own expression.  A float guess picks k, two double compares verify it.
//   * sector: a float polynomial atan2 gives q ~ theta/sector_size with a proven error < eps_q;
//     if q is further than eps_q from an integer the sector is certain.  Otherwise (a few points
//     in 10^4) the decision  atan2(y,x) >= T_j  -- T_j again from bisection on the reference's
//     expression, separately for the y >= 0 and the y < 0 branch -- is taken exactly as the sign of
//     the cross product of (x,y) with the direction of the rounding midpoint below T_j, first in
//     double, then in double-double (direction tabulated in quad precision).  This equals a
//     correctly rounded atan2; glibc's is within 0.55 ulp of that, a difference no float input
//     can expose short of lying within 2^-58 rad of a midpoint direction.
//
// Bin id = sector * R + ring (the reference's flatten order, erasor.cpp:309-320); -1 = not binned.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define ERASOR_HD __host__ __device__ __forceinline__
#define ERASOR_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define ERASOR_HD inline
#define ERASOR_HD_NOINLINE inline
#endif

namespace erasor {

struct SectorBoundary {
    double  c_hi, c_lo, s_hi, s_lo;   // cos / sin of the rounding midpoint below T_j, as double-double
    int32_t kind;                     // 0 regular, 1 always true, 2 never true (for this y-branch)
    int32_t pad_;
};

struct BinTablesView {
    float  z_lo, z_hi;       // binned iff z > z_lo && z < z_hi
    float  inv_ring;         // float(1 / ring_size)   (guess only)
    float  inv_ss;           // float(1 / sector_size) (guess only)
    float  eps_q;            // guard band on the float sector coordinate
    int    R, S;
    int    sec_of_pi;        // sector of theta = atan2(+0, -1)
    double s_max;            // r <= max_r  <=>  s <= s_max
    float  smax_lo, smax_hi; // float guard band around s_max: sf <= smax_lo => s <= s_max for sure; sf > smax_hi => s > s_max for sure
    const double*         ring_thr;   // [R+1]: [0] = -inf, [k] = min s with ring >= k, [R] = +inf
    const float*          ring_guard; // [2(R+1)]: {up_k, dn_k}: sf >= up_k => s >= ring_thr[k] for sure; sf < dn_k => s < ring_thr[k] for sure
    const SectorBoundary* sec_pos;    // [S+1], entries 1..S-1 used, y > 0 branch (A in (0, pi])
    const SectorBoundary* sec_neg;    // [S+1], y < 0 branch (A in [-pi, 0))
};

struct BinFenceCounters {
    unsigned negzero;      // SURVEY App. B-1 points
    unsigned ambiguous;    // sector decisions still ambiguous after double-double
    unsigned slow;         // points that took the exact sector path (instrumentation)
};

ERASOR_HD bool sign_bit(float f) {
#if defined(__CUDA_ARCH__)
    return (__float_as_uint(f) >> 31) != 0u;
#else
    uint32_t u; __builtin_memcpy(&u, &f, 4); return (u >> 31) != 0u;
#endif
}

// atan(t) for t in [0,1], |error| < 1.1e-7 in float arithmetic (tests/test_host_logic.py)
ERASOR_HD float atan_unit(float t) {
    const float u = t * t;
    float p = 0.0028340641874819994f;
    p = fmaf(p, u, -0.016005029901862144f);
    p = fmaf(p, u, 0.042587608098983765f);
    p = fmaf(p, u, -0.07495445758104324f);
    p = fmaf(p, u, 0.10636754333972931f);
    p = fmaf(p, u, -0.14202570915222168f);
    p = fmaf(p, u, 0.19992484152317047f);
    p = fmaf(p, u, -0.3333306610584259f);
    p = fmaf(p, u, 1.0f);
    return p * t;
}

// sign of  cos(m)*y - sin(m)*x  >  0, exactly.
ERASOR_HD_NOINLINE bool angle_above(const SectorBoundary& b, float x, float y, BinFenceCounters* fc) {
    const double xd = (double)x, yd = (double)y;
    const double p1 = yd * b.c_hi;
    const double p2 = xd * b.s_hi;
    const double d   = fma(yd, b.c_hi, -p2);
    const double mag = fabs(p1) + fabs(p2);
    if (fabs(d) > mag * 0x1p-48) return d > 0.0;
    // double-double
    const double e1 = fma(yd, b.c_hi, -p1);
    const double e2 = fma(xd, b.s_hi, -p2);
    const double sh = p1 - p2;
    const double bb = sh - p1;
    const double sl = (p1 - (sh - bb)) + (-p2 - bb);
    const double rest = sl + ((e1 - e2) + (yd * b.c_lo - xd * b.s_lo));
    const double tot  = sh + rest;
    if (!(fabs(tot) > mag * 0x1p-96)) fc->ambiguous++;
    return tot > 0.0;
}

ERASOR_HD_NOINLINE int sector_exact(const BinTablesView& T, float x, float y, int guess, BinFenceCounters* fc) {
    const SectorBoundary* tab = (y > 0.0f) ? T.sec_pos : T.sec_neg;
    fc->slow++;
    int g = guess;
    if (g < 0) g = 0;
    if (g > T.S - 1) g = T.S - 1;
    // P(j) = "atan2(y,x) >= T_j" is monotone decreasing in j; P(0) = true.  Find the largest j <= S-1 with P(j).
    while (g >= 1) {
        const SectorBoundary& b = tab[g];
        const bool pj = (b.kind == 1) ? true : (b.kind == 2) ? false : angle_above(b, x, y, fc);
        if (pj) break;
        --g;
    }
    while (g + 1 <= T.S - 1) {
        const SectorBoundary& b = tab[g + 1];
        const bool pj = (b.kind == 1) ? true : (b.kind == 2) ? false : angle_above(b, x, y, fc);
        if (!pj) break;
        ++g;
    }
    return g;
}

// ring_thr may point to a shared-memory copy of T.ring_thr.
ERASOR_HD int bin_of_point(const BinTablesView& T, const double* ring_thr, float x, float y, float z, BinFenceCounters* fc) {
    if (!(z < T.z_hi && z > T.z_lo)) return -1;
    const double xd = (double)x, yd = (double)y;
    const double s  = fma(yd, yd, xd * xd);          // exact x^2 + y^2 rounded once == the reference's pow+pow+add
    if (!(s <= T.s_max)) return -1;
    // ---- ring ----
    const float sf = fmaxf((float)s, 1e-30f);
#if defined(__CUDA_ARCH__)
    const float rf = sf * rsqrtf(sf);
#else
    const float rf = sqrtf(sf);
#endif
    int g = (int)(rf * T.inv_ring);
    g = g < 0 ? 0 : (g > T.R - 1 ? T.R - 1 : g);
    while (s >= ring_thr[g + 1]) ++g;                // ring_thr[R] = +inf
    while (s < ring_thr[g]) --g;                     // ring_thr[0] = -inf
    // ---- sector ----
    const float ax = fabsf(x), ay = fabsf(y);
    int sector;
    if (ay == 0.0f) {
        const bool xneg = sign_bit(x);
        if (xneg && sign_bit(y)) fc->negzero++;       // reference throws here (App. B-1); fenced to y = +0
        sector = xneg ? T.sec_of_pi : 0;
    } else {
        const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
#if defined(__CUDA_ARCH__)
        const float t = __fdiv_rn(mn, mx);
#else
        const float t = mn / mx;
#endif
        float a = atan_unit(t);
        if (ay > ax)  a = 1.57079637f - a;
        if (x < 0.0f) a = 3.14159274f - a;
        if (y < 0.0f) a = 6.28318548f - a;           // the 1.8e-10 between 2*3.1415926535 and 2*pi is far inside eps_q
        const float q  = a * T.inv_ss;
        const int   k  = (int)q;
        const float fr = q - (float)k;
        if (fr < T.eps_q || fr > 1.0f - T.eps_q) {
            sector = sector_exact(T, x, y, (int)(q + 0.5f) - 1, fc);
        } else {
            sector = k > T.S - 1 ? T.S - 1 : k;
        }
    }
    return sector * T.R + g;
}

// order-preserving float <-> uint32 (for atomicMin/Max on z); -0.0 < +0.0 in this order, which is
// harmless: min/max are only used through their values and -0.0 == +0.0 in every later comparison.
ERASOR_HD uint32_t float_to_ordered(float f) {
#if defined(__CUDA_ARCH__)
    const uint32_t u = __float_as_uint(f);
#else
    uint32_t u; __builtin_memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
ERASOR_HD float ordered_to_float(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}

}  // namespace erasor
