// examples/erasor_cpp_demo.cpp -- the reference's call sequence (OfflineMapUpdater.cpp:266-284) against the
// ROS-free C++ class.  Reads two clouds as raw float32 [n][4] files, prints the output sizes.
//   usage: erasor_cpp_demo map_voi.f32 query_voi.f32 [version]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#define ERASOR_B200_GLOBAL_NAMES          // `ERASOR` in the global namespace, as the reference spells it (erasor.h:43)
#include "../include/erasor/erasor.hpp"

static erasor_b200::PointCloud load(const char* path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::invalid_argument(std::string("cannot open ") + path);
    const size_t bytes = static_cast<size_t>(f.tellg());
    erasor_b200::PointCloud c(bytes / sizeof(erasor_b200::PointXYZI));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(c.data()), c.size() * sizeof(erasor_b200::PointXYZI));
    return c;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s map_voi.f32 query_voi.f32 [version]\n", argv[0]); return 2; }
    try {
        erasor_params_t p = erasor_b200::default_params();   // config/seq_05.yaml
        p.max_range = 60.0; p.num_rings = 15; p.num_sectors = 60; p.min_h = -1.3; p.max_h = 3.2; p.th_bin_max_h = 0.05;
        p.scan_ratio_threshold = 0.3; p.minimum_num_pts = 10; p.rejection_ratio = 0; p.gf_dist_thr = 0.15; p.gf_iter = 3;
        p.gf_num_lpr = 10; p.gf_th_seeds_height = 0.5;
        p.version = argc > 3 ? std::atoi(argv[3]) : 3;
        ERASOR erasor(p);
        const auto map_voi = load(argv[1]), query_voi = load(argv[2]);
        erasor_b200::PointCloud map_static_estimate, map_egocentric_complement, map_rejected, query_rejected;
        erasor.set_inputs(map_voi, query_voi);
        if (p.version == 2) erasor.compare_vois_and_revert_ground(0);
        else                erasor.compare_vois_and_revert_ground_w_block(0);
        erasor.get_static_estimate(map_static_estimate, map_egocentric_complement);
        erasor.get_outliers(map_rejected, query_rejected);
        std::cout << "ERASOR Input: " << map_voi.size() << " = " << map_static_estimate.size() << " + "
                  << map_egocentric_complement.size() << " - " << map_rejected.size() << std::endl;
        // the frame-independent batch mode on the same pair, twice in one submission: one keep byte per map point per frame
        const auto keep = erasor.process_frames({map_voi, map_voi}, {query_voi, query_voi});
        size_t rejected0 = 0, rejected1 = 0;
        for (uint8_t k : keep[0]) rejected0 += (k == 0);
        for (uint8_t k : keep[1]) rejected1 += (k == 0);
        std::cout << "batch mode: " << rejected0 << " / " << rejected1 << " map points rejected (cloud mode: " << map_rejected.size() << ")" << std::endl;
        if (rejected0 != rejected1 || (p.gf_iter > 0 && rejected0 != map_rejected.size())) { std::cerr << "batch / cloud mode disagree" << std::endl; return 1; }
        // the reference's public members (erasor.h:127,139-141) and is_dynamic_obj_close (erasor.h:132)
        erasor.set_inputs(map_voi, query_voi);
        if (p.version == 2) erasor.compare_vois_and_revert_ground(0);
        else                erasor.compare_vois_and_revert_ground_w_block(0);
        erasor.refresh_debug_members();
        bool tail_ok = erasor.ground_viz.size() <= map_static_estimate.size();
        for (size_t i = 0; tail_ok && i < erasor.ground_viz.size(); ++i) {
            const auto& a = erasor.ground_viz[i];
            const auto& b = map_static_estimate[map_static_estimate.size() - erasor.ground_viz.size() + i];
            tail_ok = a.x == b.x && a.y == b.y && a.z == b.z && a.intensity == b.intensity;
        }
        int n_close = 0;
        for (int t = 0; t < p.num_sectors; ++t) for (int r = 0; r < p.num_rings; ++r) n_close += erasor.is_dynamic_obj_close(r, t) ? 1 : 0;
        std::cout << "members: ground_viz " << erasor.ground_viz.size() << ", debug_map_rejected " << erasor.debug_map_rejected.size()
                  << ", map_complement " << erasor.map_complement.size() << ", bins next to a CURR_IS_HIGHER bin " << n_close << std::endl;
        if (!tail_ok || erasor.debug_map_rejected.size() != map_rejected.size() || erasor.map_complement.size() != map_egocentric_complement.size()) {
            std::cerr << "public members disagree with the getters" << std::endl; return 1;
        }
        // map-resident mode: the same cloud as the "global map", identity pose -> the same rejected set, on global indices
        erasor.load_global_map(map_voi);
        const std::array<double, 7> identity{0, 0, 0, 0, 0, 0, 1};
        const auto keep_nodes = erasor.process_nodes({identity}, {query_voi});
        size_t rejected_nodes = 0;
        for (uint8_t k : keep_nodes) rejected_nodes += (k == 0);
        std::cout << "node mode: " << rejected_nodes << " map points rejected" << std::endl;
        if (p.gf_iter > 0 && rejected_nodes != map_rejected.size()) { std::cerr << "node / cloud mode disagree" << std::endl; return 1; }
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
