// examples/offline_map_updater_main.cpp -- ROS-free counterpart of the reference's file-based driver
// src/offline_map_updater/main_in_your_env.cpp:60-127: reads <data_dir>/dense_global_map.pcd,
// <data_dir>/poses_lidar2body.csv (header line, then  idx,time,x,y,z,qx,qy,qz,qw ; columns 2..8 are used, :46-49)
// and <data_dir>/pcds/%06d.pcd, feeds every node to the device-resident OfflineMapUpdater and writes
// <save_path>/<data_name>_result.pcd with save_static_map(0.2) (:123).
//   usage: offline_map_updater_main <config.yaml> <data_dir> [init_idx]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../include/erasor/offline_map_updater.hpp"

using namespace erasor_b200;

static std::vector<std::array<double, 7>> load_all_poses(const std::string& txt) {
    std::vector<std::array<double, 7>> poses;
    std::ifstream in(txt);
    if (!in) throw std::invalid_argument("cannot open " + txt);
    std::string line;
    bool first = true;
    while (std::getline(in, line)) {
        if (first) { first = false; continue; }
        std::vector<float> v; std::stringstream ss(line); std::string t;
        while (std::getline(ss, t, ',')) v.push_back(std::stof(t));
        if (v.size() < 9) continue;
        // the reference builds Eigen::Quaternionf(w=pose[8], x=pose[5], y=pose[6], z=pose[7]) and a float translation (:46-47)
        poses.push_back({v[2], v[3], v[4], v[5], v[6], v[7], v[8]});
    }
    std::cout << "Total " << poses.size() << " poses are loaded" << std::endl;
    return poses;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <config.yaml> <data_dir> [init_idx]\n", argv[0]); return 2; }
    try {
        Config cfg = load_config(argv[1]);
        const std::string data_dir = argv[2];
        const int init_idx = argc > 3 ? std::atoi(argv[3]) : 0;
        cfg.initial_map_path = data_dir + "/dense_global_map.pcd";
        OfflineMapUpdater updater(cfg);
        const auto poses = load_all_poses(data_dir + "/poses_lidar2body.csv");
        // Scans come from files, so the next one is known: it is loaded before the current node is processed and, when that next
        // call will process its node (removal_interval), handed to the updater's look-ahead -- its upload and voxelisation
        // then run under the current node's path.
        auto scan_path = [&](int i) { char name[64]; std::snprintf(name, sizeof(name), "/pcds/%06d.pcd", i); return data_dir + name; };
        const int n_nodes = (int)poses.size();
        PointCloud cur, nxt;
        if (init_idx < n_nodes) cur = load_pcd(scan_path(init_idx));
        for (int i = init_idx; i < n_nodes; ++i) {
            if (i + 1 < n_nodes) {
                nxt = load_pcd(scan_path(i + 1));
                if (updater.processes_call(2)) updater.prefetch(nxt);
            }
            updater.callback_node(i, poses[i].data(), cur);
            std::swap(cur, nxt);
        }
        updater.save_static_map(0.2f);
        std::cout << "Static map building complete!" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
