// binning_tables.cpp -- host-side construction of the exact thresholds used by binning.h.
// Runs once per erasor_create().  Every threshold is found by bisection over the ordered set of
// doubles on the reference's own double-precision expression (erasor.cpp:11-21,104-110 with the
// constants of erasor.h:3-4,63-64), so it is exact by construction on this host's libm/compiler.
// Compiled with g++ (needs __float128 / libquadmath for the midpoint directions).
#include "binning_tables.h"

#include <quadmath.h>

#include <cmath>
#include <cstring>
#include <limits>

namespace erasor {

namespace {
constexpr double kPI_TRUNC = 3.1415926535;   // reference erasor.h:4

// monotone map double <-> int64
inline int64_t to_ord(double d) {
    int64_t i;
    std::memcpy(&i, &d, 8);
    return i < 0 ? (int64_t)0x8000000000000000ull - i : i;
}
inline double from_ord(int64_t o) {
    int64_t i = o < 0 ? (int64_t)0x8000000000000000ull - o : o;
    double d;
    std::memcpy(&d, &i, 8);
    return d;
}
// smallest double v in [lo, hi] with pred(v) true; requires pred monotone (false..true), pred(hi) true.
template <class F>
double first_true(double lo, double hi, F pred) {
    int64_t a = to_ord(lo), b = to_ord(hi);
    if (pred(lo)) return lo;
    while (b - a > 1) {
        int64_t m = a + (b - a) / 2;
        if (pred(from_ord(m))) b = m; else a = m;
    }
    return from_ord(b);
}

// keep the reference's expressions out of the optimiser's reach (no reassociation / constant folding surprises)
__attribute__((noinline)) int ring_of_s(volatile double s, volatile double ring_size) {
    volatile double r = std::sqrt(s);
    volatile double q = r / ring_size;
    return static_cast<int>(q);
}
__attribute__((noinline)) bool range_ok(volatile double s, volatile double max_r) {
    volatile double r = std::sqrt(s);
    return r <= max_r;
}
__attribute__((noinline)) int sector_of_theta(volatile double theta, volatile double sector_size) {
    volatile double q = theta / sector_size;
    return static_cast<int>(q);
}
__attribute__((noinline)) double theta_neg_branch(volatile double a) {   // y < 0: 2*PI + atan2(y,x)
    volatile double two_pi = 2 * kPI_TRUNC;
    volatile double t = two_pi + a;
    return t;
}

void fill_direction(double T, SectorBoundary& b) {
    // rounding midpoint below T: a correctly rounded atan2 returns >= T  <=>  true angle > (pred(T) + T)/2
    const double     pred = std::nextafter(T, -std::numeric_limits<double>::infinity());
    const __float128 m    = ((__float128)pred + (__float128)T) / 2;
    const __float128 c = cosq(m), s = sinq(m);
    b.c_hi = (double)c; b.c_lo = (double)(c - (__float128)b.c_hi);
    b.s_hi = (double)s; b.s_lo = (double)(s - (__float128)b.s_hi);
    b.kind = 0; b.pad_ = 0;
}
}  // namespace

int build_bin_tables(const erasor_params_t& p, HostBinTables& out, std::string& err) {
    const int R = p.num_rings, S = p.num_sectors;
    if (R < 1 || S < 1 || !(p.max_range > 0) || !(p.max_h > p.min_h)) { err = "invalid geometry parameters"; return -1; }
    const double max_r       = p.max_range;
    const double ring_size   = max_r / R;              // erasor.h:63
    const double sector_size = 2 * kPI_TRUNC / S;      // erasor.h:64
    out.R = R; out.S = S;
    out.ring_size = ring_size; out.sector_size = sector_size;

    // z window (erasor.cpp:104: pt.z < max_h && pt.z > min_h, float promoted to double)
    float zh = (float)p.max_h;
    if ((double)zh < p.max_h) zh = std::nextafterf(zh, std::numeric_limits<float>::infinity());
    float zl = (float)p.min_h;
    if ((double)zl > p.min_h) zl = std::nextafterf(zl, -std::numeric_limits<float>::infinity());
    out.z_hi = zh; out.z_lo = zl;

    // range: r <= max_r
    const double s_hi_bound = 4.0 * max_r * max_r + 1.0;
    const double first_bad  = first_true(0.0, s_hi_bound, [&](double s) { return !range_ok(s, max_r); });
    out.s_max = std::nextafter(first_bad, -std::numeric_limits<double>::infinity());

    // rings
    out.ring_thr.assign(R + 1, 0.0);
    out.ring_thr[0] = -std::numeric_limits<double>::infinity();
    out.ring_thr[R] = std::numeric_limits<double>::infinity();
    for (int k = 1; k <= R - 1; ++k)
        out.ring_thr[k] = first_true(0.0, s_hi_bound, [&](double s) { return ring_of_s(s, ring_size) >= k; });

    // float guard bands for the device's fast path: sf = fmaf(y, y, x * x) in float differs from the exact s by at most
    // 2^-23 relative (x * x rounded, then the fused add rounded; all terms non-negative).  With delta = 2^-22:
    //   sf >= up(T) = T (1 + 2 delta), rounded up    =>  s >= T        sf < dn(T) = T (1 - 2 delta), rounded down  =>  s < T
    // Points inside a band (a few in 10^6) take the exact double path.
    {
        const double delta2 = 2.0 * 0x1p-22;
        const float finf = std::numeric_limits<float>::infinity();
        auto up_of = [&](double T) { float f = (float)(T * (1.0 + delta2)); if ((double)f < T * (1.0 + delta2)) f = std::nextafterf(f, finf); return std::nextafterf(f, finf); };
        auto dn_of = [&](double T) { float f = (float)(T * (1.0 - delta2)); if ((double)f > T * (1.0 - delta2)) f = std::nextafterf(f, -finf); return std::nextafterf(f, -finf); };
        out.ring_guard.assign(2 * (size_t)(R + 1), 0.0f);
        out.ring_guard[0] = -finf; out.ring_guard[1] = -finf;                       // T_0 = -inf: every s is >= it
        for (int k = 1; k <= R - 1; ++k) { out.ring_guard[2 * k] = up_of(out.ring_thr[k]); out.ring_guard[2 * k + 1] = dn_of(out.ring_thr[k]); }
        out.ring_guard[2 * R] = finf; out.ring_guard[2 * R + 1] = finf;             // T_R = +inf: every finite s is below it
        out.smax_lo = dn_of(out.s_max);        // sf <= smax_lo  =>  s <= s_max   (dn is strictly below s_max (1 - 2 delta))
        out.smax_hi = up_of(out.s_max);        // sf >  smax_hi  =>  s >  s_max
    }

    // sectors
    const double A_MAX = std::atan2(0.0, -1.0);        // largest value atan2 can return
    out.sec_of_pi = std::min(sector_of_theta(A_MAX, sector_size), S - 1);
    out.sec_pos.assign(S + 1, SectorBoundary{0, 0, 0, 0, 2, 0});
    out.sec_neg.assign(S + 1, SectorBoundary{0, 0, 0, 0, 2, 0});
    out.sec_pos[0].kind = 1; out.sec_neg[0].kind = 1;
    for (int j = 1; j <= S - 1; ++j) {
        // y >= 0 branch: theta = A, A in [0, A_MAX]
        if (sector_of_theta(A_MAX, sector_size) < j) {
            out.sec_pos[j].kind = 2;                     // never reached on this branch
        } else {
            const double T = first_true(0.0, A_MAX, [&](double t) { return sector_of_theta(t, sector_size) >= j; });
            if (T <= 0.0) out.sec_pos[j].kind = 1; else fill_direction(T, out.sec_pos[j]);
        }
        // y < 0 branch: theta = 2*PI + A, A in [-A_MAX, -0)
        const double lo = -A_MAX, hi = -std::numeric_limits<double>::denorm_min();
        if (sector_of_theta(theta_neg_branch(lo), sector_size) >= j) {
            out.sec_neg[j].kind = 1;                     // always
        } else if (sector_of_theta(theta_neg_branch(hi), sector_size) < j) {
            out.sec_neg[j].kind = 2;
        } else {
            const double T = first_true(lo, hi, [&](double a) { return sector_of_theta(theta_neg_branch(a), sector_size) >= j; });
            fill_direction(T, out.sec_neg[j]);
        }
    }
    out.inv_ring = (float)(1.0 / ring_size);
    out.inv_ss   = (float)(1.0 / sector_size);
    // guard band of the float sector coordinate q = theta_f * inv_ss (binning.h):
    //   |theta_f - theta| <= 1e-6 rad (polynomial 1.1e-7, division 3e-8, three float "C - a" steps with
    //   rounded constants 8.4e-7 worst case) and q carries two more float roundings (<= 1.2e-7 * S).
    //   Doubled for margin; verified empirically by tests/test_host_logic.py.
    out.eps_q = (float)(2.0 * (1.0e-6 / sector_size + 1.5e-7 * S));
    if (!(out.eps_q < 0.25f)) { err = "num_sectors too large for the float sector guess"; return -1; }
    return 0;
}

}  // namespace erasor
