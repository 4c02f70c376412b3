// include/erasor/erasor.hpp -- ROS-free C++ host mirror of the reference's `class ERASOR`
// (reference include/erasor/erasor.h:43-147) on top of the C ABI in include/erasor_b200.h.
//
// Same public method names, argument meaning, call order and error behaviour (exceptions) as the
// reference, so that OfflineMapUpdater::callback_node (OfflineMapUpdater.cpp:266-284) compiles against it
// with two substitutions only:
//     pcl::PointCloud<pcl::PointXYZI>   ->  erasor_b200::PointCloud   (x, y, z, intensity; 16 bytes)
//     ERASOR(ros::NodeHandle*)          ->  ERASOR(const erasor_params_t&)   (the same /erasor/* keys)
// Everything the methods compute runs in the sm_100a kernels behind the C ABI; this header holds no
// arithmetic.  The rviz / debug publishers of the reference (erasor.h:67-77) are not mirrored.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../erasor_b200.h"

namespace erasor_b200 {

struct PointXYZI {          // the four floats of pcl::PointXYZI that the path uses
    float x, y, z, intensity;
};
static_assert(sizeof(PointXYZI) == 16, "PointXYZI must be one float4");
using PointCloud = std::vector<PointXYZI>;

// defaults of erasor.h:47-61 and OfflineMapUpdater.cpp:81
inline erasor_params_t default_params() {
    erasor_params_t p{};
    p.max_range = 10.0; p.num_rings = 20; p.num_sectors = 60; p.max_h = 3.0; p.min_h = 0.0; p.th_bin_max_h = 0.39;
    p.scan_ratio_threshold = 0.22; p.num_lowest_pts = 5; p.minimum_num_pts = 4; p.rejection_ratio = 0.33;
    p.gf_dist_thr = 0.05; p.gf_iter = 3; p.gf_num_lpr = 10; p.gf_th_seeds_height = 0.5; p.map_voxel_size = 0.2;
    p.version = 3; p.cov_mode = 0; p.sort_mode = 1; p.skip_voxelize = 0;
    return p;
}

class ERASOR {
public:
    // replaces ERASOR(ros::NodeHandle*): parameters are read once, R-PODs are allocated once (erasor.h:46-103)
    explicit ERASOR(const erasor_params_t& params, int device = 0) : params_(params) {
        const int rc = erasor_create(&params_, device, &h_);
        if (rc != ERASOR_OK) throw std::runtime_error(std::string("ERASOR: ") + erasor_last_error(nullptr));
    }
    ~ERASOR() { erasor_destroy(h_); }
    ERASOR(const ERASOR&) = delete;
    ERASOR& operator=(const ERASOR&) = delete;

    // Inputs: transformed & cut pcs, both in the egocentric body frame (erasor.cpp:54-59)
    void set_inputs(const PointCloud& map_voi, const PointCloud& query_voi) {
        check(erasor_set_inputs(h_, reinterpret_cast<const float*>(map_voi.data()), map_voi.size(),
                                reinterpret_cast<const float*>(query_voi.data()), query_voi.size(), ERASOR_PTR_HOST));
    }
    // Version 2 algorithm (erasor.cpp:332-434)
    void compare_vois_and_revert_ground(int frame) { check(erasor_compare(h_, 2, frame)); }
    // Version 3 algorithm (erasor.cpp:438-571)
    void compare_vois_and_revert_ground_w_block(int frame) { check(erasor_compare(h_, 3, frame)); }

    void get_static_estimate(PointCloud& arranged, PointCloud& complement) {   // erasor.cpp:612-626
        size_t na = 0, nc = 0;
        check(erasor_get_output_sizes(h_, &na, &nc, nullptr, nullptr));
        arranged.resize(na); complement.resize(nc);
        check(erasor_get_static_estimate(h_, reinterpret_cast<float*>(arranged.data()), na, &na,
                                         reinterpret_cast<float*>(complement.data()), nc, &nc, ERASOR_PTR_HOST));
    }
    void get_outliers(PointCloud& map_rejected, PointCloud& curr_rejected) {   // erasor.cpp:322-327
        size_t nm = 0, nq = 0;
        check(erasor_get_output_sizes(h_, nullptr, nullptr, &nm, &nq));
        map_rejected.resize(nm); curr_rejected.resize(nq);
        check(erasor_get_outliers(h_, reinterpret_cast<float*>(map_rejected.data()), nm, &nm,
                                  reinterpret_cast<float*>(curr_rejected.data()), nq, &nq, ERASOR_PTR_HOST));
    }
    double get_max_range() { return erasor_get_max_range(h_); }               // erasor.cpp:628

    // Frame-independent batch mode (no counterpart in the reference; BASELINE.json north_star): F independent
    // (map VoI, query VoI) pairs in one submission.  keep[f][i] == 0 where frame f rejects the i-th point of its map VoI.
    std::vector<std::vector<uint8_t>> process_frames(const std::vector<PointCloud>& map_vois, const std::vector<PointCloud>& query_vois) {
        if (map_vois.size() != query_vois.size() || map_vois.empty()) throw std::invalid_argument("ERASOR: one query VoI per map VoI");
        const size_t F = map_vois.size();
        std::vector<uint64_t> mo(F + 1, 0), qo(F + 1, 0);
        for (size_t f = 0; f < F; ++f) { mo[f + 1] = mo[f] + map_vois[f].size(); qo[f + 1] = qo[f] + query_vois[f].size(); }
        PointCloud m(mo[F]), q(qo[F]);
        for (size_t f = 0; f < F; ++f) {
            std::copy(map_vois[f].begin(), map_vois[f].end(), m.begin() + static_cast<std::ptrdiff_t>(mo[f]));
            std::copy(query_vois[f].begin(), query_vois[f].end(), q.begin() + static_cast<std::ptrdiff_t>(qo[f]));
        }
        std::vector<uint8_t> keep(mo[F] ? mo[F] : 1);
        check(erasor_process_frames(h_, reinterpret_cast<const float*>(m.data()), mo.data(), reinterpret_cast<const float*>(q.data()), qo.data(),
                                    static_cast<int>(F), keep.data(), ERASOR_PTR_HOST));
        std::vector<std::vector<uint8_t>> out(F);
        for (size_t f = 0; f < F; ++f) out[f].assign(keep.begin() + static_cast<std::ptrdiff_t>(mo[f]), keep.begin() + static_cast<std::ptrdiff_t>(mo[f + 1]));
        return out;
    }

    // what the reference exposes as /SCDR/debug/polygons_marker likelihoods (erasor.cpp:439-441,570);
    // index = sector * num_rings + ring
    std::vector<float> get_status() {
        std::vector<float> st(static_cast<size_t>(params_.num_rings) * params_.num_sectors);
        check(erasor_get_status(h_, st.data()));
        return st;
    }
    erasor_handle_t handle() const { return h_; }
    const erasor_params_t& params() const { return params_; }

private:
    void check(int rc) const {
        if (rc == ERASOR_OK) return;
        const std::string msg = std::string("ERASOR: ") + erasor_last_error(h_);
        if (rc == ERASOR_E_INVALID) throw std::invalid_argument(msg);   // e.g. "Other version is not implemented!" (OfflineMapUpdater.cpp:274)
        throw std::runtime_error(msg);
    }
    erasor_params_t params_;
    erasor_handle_t h_ = nullptr;
};

}  // namespace erasor_b200
