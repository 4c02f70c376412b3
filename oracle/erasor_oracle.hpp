// ============================================================================
//  oracle/erasor_oracle.hpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE
//
//  CPU restatement (ROS/PCL/Eigen-free, C++17, zero dependencies) of the
//  reference's per-frame  R-POD -> Scan Ratio Test -> R-GPF  hot path:
//      /root/reference/include/erasor/erasor.h
//      /root/reference/src/offline_map_updater/src/erasor.cpp
//  plus the pieces of its one caller that sit either side of the path:
//      /root/reference/src/offline_map_updater/src/OfflineMapUpdater.cpp:203-449
//      /root/reference/src/offline_map_updater/src/erasor_utils.cpp:57-114
//
//  PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
//   and cannot be compiled in this image (needs ROS, PCL,
//  Eigen, Boost, FLANN).  Third-party arithmetic that the path calls
//  (pcl::computeMeanAndCovarianceMatrix, Eigen::JacobiSVD, pcl::VoxelGrid,
//  FLANN 1-NN, pcl::transformPointCloud) is restated below from the published
//  algorithms of the versions the reference's stated environment implies
//  (Ubuntu 18.04 / ROS Melodic => PCL 1.8.1, Eigen 3.3.4, FLANN 1.9.1); each
//  restatement is marked [3P].  The oracle is hardened by the source-derived
//  invariants in tests/test_oracle_invariants.py and by an independent numpy
//  restatement of binning + SRT (tests/np_restatement.py).
//
//  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
//  reference legs may use anything in this directory.
//
//  Data structures deliberately mirror the reference (32-byte points, one
//  std::vector per bin, whole-bin deep copies) so that the same code is also
//  the honest "restated reference path" for CPU timing.
// ============================================================================
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <string>

namespace oracle {

// Same constants as erasor.h:3-18
constexpr double kINF            = 10000000000000.0;   // erasor.h:3
constexpr double kPI             = 3.1415926535;       // erasor.h:4 (truncated on purpose)
constexpr int    kENOUGH_NUM     = 8000;               // erasor.h:5
constexpr double MAP_IS_HIGHER   = 0.5;                // erasor.h:12
constexpr double CURR_IS_HIGHER  = 1.0;                // erasor.h:13
constexpr double LITTLE_NUM      = 0.0;                // erasor.h:14
constexpr double BLOCKED         = 0.8;                // erasor.h:15
constexpr double MERGE_BINS      = 0.25;               // erasor.h:17
constexpr double NOT_ASSIGNED    = 0.0;                // erasor.h:18

// pcl::PointXYZI is a 32-byte, 16-aligned struct: {x,y,z,pad}{intensity,pad*3}.
// The first pad word carries the SOURCE INDEX of the point in the cloud that
// was handed to set_inputs (map: idx, query: idx | 0x80000000, synthesised
// voxel centroids: 0xFFFFFFFF).  It is never read by the algorithm; it only
// lets the tests follow a point through the reference's copies.
struct alignas(16) PointXYZI {
    float    x, y, z;
    uint32_t src;
    float    intensity;
    uint32_t pad_[3];
};
static_assert(sizeof(PointXYZI) == 32, "PointXYZI must be 32 bytes like pcl::PointXYZI");
using Cloud = std::vector<PointXYZI>;

constexpr uint32_t SRC_QUERY_BIT = 0x80000000u;
constexpr uint32_t SRC_NONE      = 0xFFFFFFFFu;

// Parameters: exactly the 15 keys ERASOR's ctor pulls from /erasor/* (erasor.h:47-61)
// + /erasor/version (OfflineMapUpdater.cpp:81) + oracle-only mode switches.
struct Params {
    double max_range            = 10.0;   // erasor.h:47
    double min_h                = 0.0;    // erasor.h:51
    double max_h                = 3.0;    // erasor.h:50
    double th_bin_max_h         = 0.39;   // erasor.h:52
    double scan_ratio_threshold = 0.22;   // erasor.h:53
    double rejection_ratio      = 0.33;   // erasor.h:56 (unused by the path)
    double gf_dist_thr          = 0.05;   // erasor.h:57
    double gf_th_seeds_height   = 0.5;    // erasor.h:60
    double map_voxel_size       = 0.2;    // erasor.h:61
    int    num_rings            = 20;     // erasor.h:48
    int    num_sectors          = 60;     // erasor.h:49
    int    num_lowest_pts       = 5;      // erasor.h:54
    int    minimum_num_pts      = 4;      // erasor.h:55
    int    gf_iter              = 3;      // erasor.h:58
    int    gf_num_lpr           = 10;     // erasor.h:59
    int    version              = 3;      // OfflineMapUpdater.cpp:81
    // ---- mode switches (not in the reference) ----
    // cov_mode 0: pcl::computeMeanAndCovarianceMatrix of PCL <= 1.10 (unshifted single pass, Melodic default)
    // cov_mode 1: PCL >= 1.11 (first point subtracted before accumulation)
    int    cov_mode             = 0;
    // sort_mode 0: std::sort (as written, erasor.cpp:240; order of equal z is libstdc++'s)
    // sort_mode 1: stable order (z, then source position) -- a legal outcome of the
    //              reference's unstable sort and the one the CUDA path implements
    int    sort_mode            = 1;
    // skip_voxelize 1: leave out the in-bin voxelize_preserving_labels of v3
    //              (erasor.cpp:526-528); used to isolate R-GPF in tests
    int    skip_voxelize        = 0;
};

struct Bin {                      // erasor.h:24-33
    double max_h;
    double min_h;
    double x;
    double y;
    double status;
    bool   is_occupied;
    Cloud  points;
};
using Ring  = std::vector<Bin>;
using R_POD = std::vector<Ring>;  // r_pod[ring][sector], erasor.h:40

// One record per bin that ran extract_ground (in processing order: theta outer, r inner).
struct PlaneTap {
    int   ring, sector;
    int   n_points;                       // |bin_map.points|
    int   n_seeds;                        // |initial seeds|
    double lpr_height;
    std::vector<float>  normal;           // 3 per iteration
    std::vector<double> d;                // 1 per iteration
    std::vector<int>    n_ground;         // ground count after each iteration
    int   n_empty_fits;                   // how many estimate_plane_ calls saw an empty set (App. B-3 fence)
};

class ERASOR {
public:
    explicit ERASOR(const Params& p);

    // erasor.cpp:57-85
    void set_inputs(const Cloud& map_voi, const Cloud& query_voi);
    // erasor.cpp:332-434 (version 2)
    void compare_vois_and_revert_ground(int frame);
    // erasor.cpp:438-571 (version 3)
    void compare_vois_and_revert_ground_w_block(int frame);
    // erasor.cpp:573-595
    bool is_dynamic_obj_close(R_POD& r_pod_selected, int r_target, int theta_target, int r_range, int theta_range);
    // erasor.cpp:612-626
    void get_static_estimate(Cloud& arranged, Cloud& complement);
    // erasor.cpp:322-327
    void get_outliers(Cloud& map_rejected, Cloud& curr_rejected);
    // erasor.cpp:628
    double get_max_range() const { return max_r; }

    Cloud ground_viz;             // erasor.h:127
    Cloud debug_curr_rejected;    // erasor.h:139
    Cloud debug_map_rejected;     // erasor.h:140
    Cloud map_complement;         // erasor.h:141
    R_POD r_pod_map;              // erasor.h:143
    R_POD r_pod_curr;             // erasor.h:144
    R_POD r_pod_selected;         // erasor.h:145

    // ---- parity taps (oracle only) ----
    std::vector<int32_t>  tap_bin_of_map;     // per map_voi point: sector*R + ring, or -1 (complement)
    std::vector<int32_t>  tap_bin_of_query;   // per query_voi point, or -1 (dropped)
    std::vector<double>   tap_status;         // final status per bin, index sector*R + ring
    std::vector<double>   tap_status_pass1;   // v3: status after pass 1
    std::vector<PlaneTap> tap_planes;
    long                  tap_negzero_fenced = 0;  // App. B-1 points (y == -0.0f, x < 0)

    Params p;
    int    num_rings, num_sectors;
    double max_r, ring_size, sector_size, min_h, max_h;

    // exposed for unit tests
    double xy2theta(const double& x, const double& y);        // erasor.cpp:11-17
    double xy2radius(const double& x, const double& y);       // erasor.cpp:19-21
    void   extract_ground(const Cloud& src, Cloud& dst, Cloud& outliers);   // erasor.cpp:233-294

private:
    void init(R_POD& r_pod);                                   // erasor.cpp:29-42
    void clear_bin(Bin& bin);                                  // erasor.cpp:44-52
    void pt2r_pod(const PointXYZI& pt, Bin& bin);              // erasor.cpp:87-98
    void voi2r_pod(const Cloud& src, R_POD& r_pod, std::vector<int32_t>& tap);                      // erasor.cpp:100-122
    void voi2r_pod(const Cloud& src, R_POD& r_pod, Cloud& complement, std::vector<int32_t>& tap);   // erasor.cpp:124-144
    void estimate_plane_(const Cloud& ground);                 // erasor.cpp:183-198
    void extract_initial_seeds_(const Cloud& p_sorted, Cloud& init_seeds);   // erasor.cpp:204-231
    void merge_bins(const Bin& src1, const Bin& src2, Bin& dst);             // erasor.cpp:296-307
    void r_pod2pc(const R_POD& sc, Cloud& pc);                 // erasor.cpp:309-320
    bool bin_index(const PointXYZI& pt, int& ring_idx, int& sector_idx);

    Cloud  piecewise_ground_, non_ground_, ground_pc_, non_ground_pc_;
    float  normal_[3] = {0.f, 0.f, 1.f};
    double th_dist_d_ = 0, d_ = 0;
    PlaneTap* cur_tap_ = nullptr;
};

// ---------------------------------------------------------------------------
// [3P] restatements, usable on their own in unit tests
// ---------------------------------------------------------------------------
// pcl::computeMeanAndCovarianceMatrix<PointXYZI,float> ; returns point count.
// mode 0 = PCL<=1.10, mode 1 = PCL>=1.11.  cov row-major 3x3, mean[4].
unsigned compute_mean_and_covariance(const Cloud& cloud, float cov[9], float mean[4], int mode);
// Eigen 3.3 JacobiSVD<MatrixXf>(A, ComputeFullU) for a 3x3 input; U row-major; sv[3].
void jacobi_svd_3x3_full_u(const float A[9], float U[9], float sv[3]);
// erasor_utils::voxelize_preserving_labels (erasor_utils.cpp:80-114) = pcl::VoxelGrid + FLANN 1-NN.
void voxelize_preserving_labels(const Cloud& src, Cloud& dst, double leaf_size);
// pcl::transformPointCloud(cloud_in, cloud_out, Eigen::Matrix4f) (PCL 1.8 scalar path); T row-major 4x4.
extern int g_study[3];   // blast-radius study switches (erasor_oracle.cpp); all 0 outside scripts/blast_radius.py
void transform_point_cloud(const Cloud& in, Cloud& out, const float T[16]);
// Eigen::Matrix4f::inverse() stand-in (general 4x4, float cofactors).
void invert_4x4(const float T[16], float Tinv[16]);
// erasor_utils::geoPose2eigen (erasor_utils.cpp:35-55): pose = {x,y,z,qx,qy,qz,qw} -> row-major float 4x4
void geo_pose_to_matrix(const double pose[7], float T[16]);
// erasor_utils::parse_dynamic_obj label test (erasor_utils.cpp:63-72)
bool is_dynamic_label(float intensity);

// ---------------------------------------------------------------------------
// The caller, OfflineMapUpdater::callback_node, ROS stripped
// (OfflineMapUpdater.cpp:203-330, 332-449).  Only what feeds or drains the path.
// ---------------------------------------------------------------------------
struct UpdaterParams {
    double query_voxel_size = 0.05;   // OfflineMapUpdater.cpp:66
    double map_voxel_size   = 0.05;   // :67 (unused by the path)
    int    removal_interval = 2;      // :69
    bool   is_large_scale   = false;  // :75
    double submap_size      = 200.0;  // :76
    double max_range        = 60.0;   // :78
    int    version          = 3;      // :81
    double lidar2body[7]    = {0, 0, 0, 0, 0, 0, 1};   // /tf/lidar2body
};

class OfflineMapUpdater {
public:
    OfflineMapUpdater(const UpdaterParams& up, const Params& ep, const Cloud& initial_map);
    // One erasor/node message: pose body->origin {x,y,z,qx,qy,qz,qw}, raw lidar scan (lidar frame).
    // Returns true when the node was processed (every removal_interval-th call).
    bool callback_node(int seq, const double odom[7], const Cloud& lidar);
    // OfflineMapUpdater.cpp:174-196 minus the file write
    void save_static_map(float voxel_size, Cloud& map_to_be_saved);

    Cloud map_arranged_, map_arranged_global_, map_arranged_complement_;
    Cloud query_voi_, map_voi_, map_voi_wrt_origin_, map_outskirts_;
    Cloud map_static_estimate_, map_egocentric_complement_, map_filtered_;
    Cloud query_rejected_, map_rejected_, total_query_rejected_, total_map_rejected_;
    ERASOR erasor_;
    double last_erasor_seconds = 0, last_voi_seconds = 0;
    int    stack_count = 0;
    float  tf_lidar2body_[16], tf_body2origin_[16];

private:
    void reassign_submap(double pose_x, double pose_y);                          // :332-358
    void set_submap(const Cloud& map_global, Cloud& submap, Cloud& submap_complement,
                    double x, double y, double submap_size);                     // :360-379
    void fetch_VoI(double x_criterion, double y_criterion, Cloud& dst, Cloud& outskirts);   // :381-438 ("naive" mode)
    void body2origin(const Cloud src, Cloud& dst);                               // :441-449
    UpdaterParams up_;
    bool   is_submap_not_initialized_ = true;
    double submap_center_x_ = 0, submap_center_y_ = 0;
    size_t num_pcs_init_ = 0;
};


// ---------------------------------------------------------------------------
// mapgen, ROS stripped (src/mapgen/mapgen.hpp:198-309; driver src/mapgen/main.cpp:40-49): the naive map builder that
// produces the path's input map.  Per node: drop the points within CAR_BODY_SIZE of the sensor, lift by 1.73 m,
// move to the map frame with the node's pose, voxelize at 0.2 m, append; at the end voxelize the whole map at `leafsize`.
// ---------------------------------------------------------------------------
class NaiveMapGen {
public:
    NaiveMapGen(float leafsize, bool is_large_scale) : leafsize_(leafsize), is_large_scale_(is_large_scale) {}
    void accum_point_cloud(const double odom[7], const Cloud& lidar);     // accumPointCloud, mapgen.hpp:198-263
    void save_naive_map(Cloud& original, Cloud& voxelized) const;          // saveNaiveMap, mapgen.hpp:271-307
    Cloud cloud_map, cloud_curr;                                           // :38-39 (getPointClouds :265-269)
    std::vector<Cloud> cloud_maps;                                         // :37 (large-scale submaps)
private:
    float leafsize_;
    bool  is_large_scale_;
    bool  is_initial_ = true;                                              // :33
    int   cnt_voxel_  = 0;                                                 // function-static in the reference (:248)
};

}  // namespace oracle
