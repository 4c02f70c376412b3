// binning_tables.h -- host container for the exact binning thresholds (see binning.h).
#pragma once
#include <string>
#include <vector>

#include "../../include/erasor_b200.h"
#include "binning.h"

namespace erasor {

struct HostBinTables {
    int    R = 0, S = 0;
    double ring_size = 0, sector_size = 0;
    float  z_lo = 0, z_hi = 0, inv_ring = 0, inv_ss = 0, eps_q = 0;
    int    sec_of_pi = 0;
    double s_max = 0;
    float  smax_lo = 0, smax_hi = 0;
    std::vector<double>         ring_thr;   // R+1
    std::vector<float>          ring_guard; // 2(R+1): {up, dn} per threshold (float guard band, see binning.h)
    std::vector<SectorBoundary> sec_pos;    // S+1
    std::vector<SectorBoundary> sec_neg;    // S+1

    BinTablesView view(const double* ring, const SectorBoundary* pos, const SectorBoundary* neg, const float* guard) const {
        BinTablesView v;
        v.z_lo = z_lo; v.z_hi = z_hi; v.inv_ring = inv_ring; v.inv_ss = inv_ss; v.eps_q = eps_q;
        v.R = R; v.S = S; v.sec_of_pi = sec_of_pi; v.s_max = s_max; v.smax_lo = smax_lo; v.smax_hi = smax_hi;
        v.ring_thr = ring; v.sec_pos = pos; v.sec_neg = neg; v.ring_guard = guard;
        return v;
    }
    BinTablesView host_view() const { return view(ring_thr.data(), sec_pos.data(), sec_neg.data(), ring_guard.data()); }
};

// returns 0 on success
int build_bin_tables(const erasor_params_t& p, HostBinTables& out, std::string& err);

}  // namespace erasor
