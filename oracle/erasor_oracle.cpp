// ============================================================================
//  oracle/erasor_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE
//  See erasor_oracle.hpp for scope, provenance and the "parity unpinned" note.
//  Build with -ffp-contract=off (the reference's CMakeLists.txt:3-4 passes no
//  -march, so its x86-64 build has no FMA contraction either).
// ============================================================================
#include "erasor_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <unordered_map>

namespace oracle {

using std::max;
using std::min;

// ----------------------------------------------------------------------------
// ERASOR
// ----------------------------------------------------------------------------
ERASOR::ERASOR(const Params& prm) : p(prm) {
    // erasor.h:47-64
    max_r       = p.max_range;
    num_rings   = p.num_rings;
    num_sectors = p.num_sectors;
    max_h       = p.max_h;
    min_h       = p.min_h;
    ring_size   = max_r / num_rings;              // erasor.h:63
    sector_size = 2 * kPI / num_sectors;          // erasor.h:64
    init(r_pod_map);                              // erasor.h:95-97
    init(r_pod_curr);
    init(r_pod_selected);
    piecewise_ground_.reserve(130000);            // erasor.h:99-102
    non_ground_.reserve(130000);
    ground_pc_.reserve(130000);
    non_ground_pc_.reserve(130000);
}

double ERASOR::xy2theta(const double& x, const double& y) {   // erasor.cpp:11-17
    if (y >= 0) {
        return atan2(y, x);
    } else {
        return 2 * kPI + atan2(y, x);
    }
}

double ERASOR::xy2radius(const double& x, const double& y) {  // erasor.cpp:19-21
    return sqrt(pow(x, 2) + pow(y, 2));
}

void ERASOR::init(R_POD& r_pod) {                              // erasor.cpp:29-42
    if (!r_pod.empty()) r_pod.clear();
    Ring ring;
    Bin  bin = {-kINF, kINF, 0, 0, static_cast<double>(false), static_cast<bool>(NOT_ASSIGNED), {}};
    bin.points.reserve(kENOUGH_NUM);
    for (int i = 0; i < num_sectors; i++) ring.emplace_back(bin);
    for (int j = 0; j < num_rings; j++) r_pod.emplace_back(ring);
}

void ERASOR::clear_bin(Bin& bin) {                             // erasor.cpp:44-52
    bin.max_h       = -kINF;
    bin.min_h       = kINF;
    bin.x           = 0;
    bin.y           = 0;
    bin.is_occupied = false;
    bin.status      = NOT_ASSIGNED;
    if (!bin.points.empty()) bin.points.clear();
}

void ERASOR::set_inputs(const Cloud& map_voi, const Cloud& query_voi) {   // erasor.cpp:57-85
    debug_curr_rejected.clear();
    debug_map_rejected.clear();
    map_complement.clear();
    for (int theta = 0; theta < num_sectors; ++theta) {
        for (int r = 0; r < num_rings; ++r) {
            clear_bin(r_pod_map[r][theta]);
            clear_bin(r_pod_curr[r][theta]);
            clear_bin(r_pod_selected[r][theta]);
        }
    }
    tap_negzero_fenced = 0;
    voi2r_pod(query_voi, r_pod_curr, tap_bin_of_query);
    voi2r_pod(map_voi, r_pod_map, map_complement, tap_bin_of_map);
}

void ERASOR::pt2r_pod(const PointXYZI& pt, Bin& bin) {         // erasor.cpp:87-98
    bin.is_occupied = true;
    bin.points.push_back(pt);
    if (pt.z >= bin.max_h) {
        bin.max_h = pt.z;
        bin.x     = pt.x;
        bin.y     = pt.y;
    }
    if (pt.z <= bin.min_h) {
        bin.min_h = pt.z;
    }
}

// The index arithmetic shared by both voi2r_pod overloads (erasor.cpp:104-110 / 128-135).
// Returns false when the point fails the z window or the range test.
bool ERASOR::bin_index(const PointXYZI& pt, int& ring_idx, int& sector_idx) {
    if (pt.z < max_h && pt.z > min_h) {
        double r = xy2radius(pt.x, pt.y);
        if (r <= max_r) {
            double theta = xy2theta(pt.x, pt.y);
            sector_idx = min(static_cast<int>((theta / sector_size)), num_sectors - 1);
            ring_idx   = min(static_cast<int>((r / ring_size)), num_rings - 1);
            if (sector_idx < 0) {
                // y == -0.0f with x <= -0 gives theta = -pi, a negative
                // sector, and r_pod.at() throws std::out_of_range in the reference
                // (erasor.cpp:112,136).  FENCE (oracle and CUDA path alike): treat y as
                // +0.0f and count the event.
                tap_negzero_fenced++;
                theta      = xy2theta(pt.x, 0.0);
                sector_idx = min(static_cast<int>((theta / sector_size)), num_sectors - 1);
            }
            return true;
        }
    }
    return false;
}

void ERASOR::voi2r_pod(const Cloud& src, R_POD& r_pod, std::vector<int32_t>& tap) {   // erasor.cpp:100-122
    tap.assign(src.size(), -1);
    size_t i = 0;
    for (auto const& pt : src) {
        int ring_idx, sector_idx;
        if (bin_index(pt, ring_idx, sector_idx)) {
            pt2r_pod(pt, r_pod.at(ring_idx).at(sector_idx));
            tap[i] = sector_idx * num_rings + ring_idx;
        }
        ++i;
    }
    // erasor.cpp:117-121 (debug r_pod2pc + publish) omitted: no observable effect on the path.
}

void ERASOR::voi2r_pod(const Cloud& src, R_POD& r_pod, Cloud& complement, std::vector<int32_t>& tap) {   // erasor.cpp:124-144
    tap.assign(src.size(), -1);
    size_t i = 0;
    for (auto const& pt : src) {
        int ring_idx, sector_idx;
        if (bin_index(pt, ring_idx, sector_idx)) {
            pt2r_pod(pt, r_pod.at(ring_idx).at(sector_idx));
            tap[i] = sector_idx * num_rings + ring_idx;
        } else {
            complement.push_back(pt);
        }
        ++i;
    }
}

// ----------------------------------------------------------------------------
// [3P] pcl::computeMeanAndCovarianceMatrix<PointXYZI, float>  
// ----------------------------------------------------------------------------
unsigned compute_mean_and_covariance(const Cloud& cloud, float cov[9], float mean[4], int mode) {
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float K[3]    = {0, 0, 0};
    if (mode == 1 && !cloud.empty()) {   // PCL >= 1.11: shift by the first (finite) point
        K[0] = cloud[0].x; K[1] = cloud[0].y; K[2] = cloud[0].z;
    }
    const size_t point_count = cloud.size();
    for (size_t i = 0; i < point_count; ++i) {
        const float x = (mode == 1) ? cloud[i].x - K[0] : cloud[i].x;
        const float y = (mode == 1) ? cloud[i].y - K[1] : cloud[i].y;
        const float z = (mode == 1) ? cloud[i].z - K[2] : cloud[i].z;
        accu[0] += x * x;
        accu[1] += x * y;
        accu[2] += x * z;
        accu[3] += y * y;
        accu[4] += y * z;
        accu[5] += z * z;
        accu[6] += x;
        accu[7] += y;
        accu[8] += z;
    }
    if (point_count != 0) {
        const float n = static_cast<float>(point_count);
        for (int k = 0; k < 9; ++k) accu[k] /= n;
        mean[0] = (mode == 1) ? accu[6] + K[0] : accu[6];
        mean[1] = (mode == 1) ? accu[7] + K[1] : accu[7];
        mean[2] = (mode == 1) ? accu[8] + K[2] : accu[8];
        mean[3] = 1;
        cov[0] = accu[0] - accu[6] * accu[6];
        cov[1] = accu[1] - accu[6] * accu[7];
        cov[2] = accu[2] - accu[6] * accu[8];
        cov[4] = accu[3] - accu[7] * accu[7];
        cov[5] = accu[4] - accu[7] * accu[8];
        cov[8] = accu[5] - accu[8] * accu[8];
        cov[3] = cov[1];
        cov[6] = cov[2];
        cov[7] = cov[5];
    }
    return static_cast<unsigned>(point_count);
}

// ----------------------------------------------------------------------------
// [3P] Eigen 3.3 JacobiSVD, 3x3 float, ComputeFullU, square => no QR preconditioner
// (Eigen/src/SVD/JacobiSVD.h, Eigen/src/Jacobi/Jacobi.h)
// ----------------------------------------------------------------------------
namespace {
struct Rot { float c, s; };

inline Rot rot_transpose(Rot j) { return Rot{j.c, -j.s}; }
inline Rot rot_mul(Rot a, Rot b) {   // JacobiRotation::operator*
    return Rot{a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c};
}
// internal::apply_rotation_in_the_plane on two strided 'vectors' of length n
inline void apply_rot(float* x, int incx, float* y, int incy, int n, Rot j) {
    if (j.c == 1.0f && j.s == 0.0f) return;
    for (int i = 0; i < n; ++i) {
        const float xi = *x, yi = *y;
        *x = j.c * xi + j.s * yi;
        *y = -j.s * xi + j.c * yi;
        x += incx; y += incy;
    }
}
// JacobiRotation::makeJacobi(x, y, z)
inline Rot make_jacobi(float x, float y, float z) {
    Rot r;
    const float deno = 2.0f * std::fabs(y);
    if (deno < FLT_MIN) {
        r.c = 1.0f; r.s = 0.0f;
    } else {
        const float tau = (x - z) / deno;
        const float w   = std::sqrt(tau * tau + 1.0f);
        float t;
        if (tau > 0.0f) t = 1.0f / (tau + w);
        else            t = 1.0f / (tau - w);
        const float sign_t = t > 0.0f ? 1.0f : -1.0f;
        const float n      = 1.0f / std::sqrt(t * t + 1.0f);
        r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
        r.c = n;
    }
    return r;
}
// internal::real_2x2_jacobi_svd on W (row-major 3x3) at (p,q)
inline void real_2x2_jacobi_svd(const float* W, int p, int q, Rot* j_left, Rot* j_right) {
    float m[4] = {W[p * 3 + p], W[p * 3 + q], W[q * 3 + p], W[q * 3 + q]};   // m00 m01 m10 m11
    Rot rot1;
    const float t = m[0] + m[3];
    const float d = m[2] - m[1];
    if (std::fabs(d) < FLT_MIN) {
        rot1.s = 0.0f; rot1.c = 1.0f;
    } else {
        const float u   = t / d;
        const float tmp = std::sqrt(1.0f + u * u);
        rot1.s = 1.0f / tmp;
        rot1.c = u / tmp;
    }
    apply_rot(&m[0], 1, &m[2], 1, 2, rot1);            // m.applyOnTheLeft(0,1,rot1)
    *j_right = make_jacobi(m[0], m[1], m[3]);          // j_right->makeJacobi(m,0,1)
    *j_left  = rot_mul(rot1, rot_transpose(*j_right));
}
}  // namespace

int g_last_sweeps = 0;
void jacobi_svd_3x3_full_u(const float A[9], float U[9], float sv[3]) {
    const float precision      = 2.0f * FLT_EPSILON;
    const float considerAsZero = FLT_MIN;
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) scale = (std::fabs(A[i]) > scale) ? std::fabs(A[i]) : scale;   // cwiseAbs().maxCoeff()
    if (scale == 0.0f) scale = 1.0f;
    float W[9];
    for (int i = 0; i < 9; ++i) W[i] = A[i] / scale;
    for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    float maxDiagEntry = 0.0f;
    for (int i = 0; i < 3; ++i) maxDiagEntry = (std::fabs(W[i * 4]) > maxDiagEntry) ? std::fabs(W[i * 4]) : maxDiagEntry;

    bool finished = false;
    int  sweeps   = 0;
    while (!finished && sweeps < 1000) {   // the 1000 cap is an oracle safety net only
        finished = true;
        ++sweeps;
        for (int p = 1; p < 3; ++p) {
            for (int q = 0; q < p; ++q) {
                const float threshold = std::max(considerAsZero, precision * maxDiagEntry);
                if (std::fabs(W[p * 3 + q]) > threshold || std::fabs(W[q * 3 + p]) > threshold) {
                    finished = false;
                    Rot j_left, j_right;
                    real_2x2_jacobi_svd(W, p, q, &j_left, &j_right);
                    apply_rot(&W[p * 3], 1, &W[q * 3], 1, 3, j_left);                   // W.applyOnTheLeft(p,q,j_left)
                    apply_rot(&U[p], 3, &U[q], 3, 3, rot_transpose(rot_transpose(j_left)));   // U.applyOnTheRight(p,q,j_left.transpose())
                    apply_rot(&W[p], 3, &W[q], 3, 3, rot_transpose(j_right));           // W.applyOnTheRight(p,q,j_right)
                    maxDiagEntry = std::max(maxDiagEntry, std::max(std::fabs(W[p * 4]), std::fabs(W[q * 4])));
                }
            }
        }
    }
    g_last_sweeps = sweeps;
    for (int i = 0; i < 3; ++i) {
        const float a = W[i * 4];
        sv[i] = std::fabs(a);
        if (a < 0.0f) for (int r = 0; r < 3; ++r) U[r * 3 + i] = -U[r * 3 + i];
    }
    for (int i = 0; i < 3; ++i) sv[i] *= scale;
    for (int i = 0; i < 3; ++i) {   // selection sort, descending
        int   pos = 0;
        float mx  = sv[i];
        for (int k = i + 1; k < 3; ++k) if (sv[k] > mx) { mx = sv[k]; pos = k - i; }
        if (mx == 0.0f) break;
        if (pos) {
            pos += i;
            std::swap(sv[i], sv[pos]);
            for (int r = 0; r < 3; ++r) std::swap(U[r * 3 + pos], U[r * 3 + i]);
        }
    }
}

// ----------------------------------------------------------------------------
// R-GPF
// ----------------------------------------------------------------------------
void ERASOR::estimate_plane_(const Cloud& ground) {            // erasor.cpp:183-198
    float cov[9], pc_mean[4];
    // the reference leaves cov / pc_mean uninitialised when the set is
    // empty (UB).  FENCE: an empty set fits cov = 0, mean = 0 => U = I, normal = (0,0,1), d = 0.
    for (int i = 0; i < 9; ++i) cov[i] = 0.0f;
    for (int i = 0; i < 4; ++i) pc_mean[i] = 0.0f;
    const unsigned n = compute_mean_and_covariance(ground, cov, pc_mean, p.cov_mode);
    if (n == 0 && cur_tap_) cur_tap_->n_empty_fits++;
    float U[9], sv[3];
    jacobi_svd_3x3_full_u(cov, U, sv);
    normal_[0] = U[2]; normal_[1] = U[5]; normal_[2] = U[8];   // svd.matrixU().col(2)
    const float seeds_mean[3] = {pc_mean[0], pc_mean[1], pc_mean[2]};
    const float dot = (normal_[0] * seeds_mean[0] + normal_[1] * seeds_mean[1]) + normal_[2] * seeds_mean[2];
    d_         = -dot;                                         // erasor.cpp:195
    th_dist_d_ = p.gf_dist_thr - d_;                           // erasor.cpp:197
}

static bool point_cmp(PointXYZI a, PointXYZI b) { return a.z < b.z; }   // erasor.cpp:200-202

// "Blast radius" study switches (scripts/blast_radius.py; never set by tests or the bench): alternative legal outcomes of the
// choices the reference leaves to its libraries / compiler, to measure how far the results can move if a real PCL build
// differs from the pinned ones.  [0] ties of the z-sort in REVERSE source order (the opposite extreme from the stable
// order), [1] the classification dot product (erasor.cpp:271) FMA-contracted, [2] 1-NN distance ties to the HIGHEST index.
int g_study[3] = {0, 0, 0};

void ERASOR::extract_initial_seeds_(const Cloud& p_sorted, Cloud& init_seeds) {   // erasor.cpp:204-231
    init_seeds.clear();
    Cloud  g_seeds_pc;
    double sum = 0;
    int    cnt = 0;
    for (int i = p.num_lowest_pts; static_cast<size_t>(i) < p_sorted.size() && cnt < p.gf_num_lpr; i++) {
        sum += p_sorted[i].z;
        cnt++;
    }
    double lpr_height = cnt != 0 ? sum / cnt : 0;
    g_seeds_pc.clear();
    for (size_t i = 0; i < p_sorted.size(); i++) {
        if (p_sorted[i].z < lpr_height + p.gf_th_seeds_height) {
            g_seeds_pc.push_back(p_sorted[i]);
        }
    }
    init_seeds = g_seeds_pc;
    if (cur_tap_) { cur_tap_->lpr_height = lpr_height; cur_tap_->n_seeds = static_cast<int>(init_seeds.size()); }
}

void ERASOR::extract_ground(const Cloud& src, Cloud& dst, Cloud& outliers) {       // erasor.cpp:233-294
    if (!dst.empty()) dst.clear();
    if (!outliers.empty()) outliers.clear();

    auto src_copy = src;
    if (g_study[0]) { std::reverse(src_copy.begin(), src_copy.end()); std::stable_sort(src_copy.begin(), src_copy.end(), point_cmp); }
    else if (p.sort_mode == 0) std::sort(src_copy.begin(), src_copy.end(), point_cmp);
    else                  std::stable_sort(src_copy.begin(), src_copy.end(), point_cmp);
    // 1. remove_outliers (erasor.cpp:242-251)
    auto it = src_copy.begin();
    for (size_t i = 0; i < src_copy.size(); i++) {
        if (src_copy[i].z < min_h) it++;
        else break;
    }
    src_copy.erase(src_copy.begin(), it);

    // 2. set seeds
    if (!ground_pc_.empty()) ground_pc_.clear();
    if (!non_ground_pc_.empty()) non_ground_pc_.clear();
    extract_initial_seeds_(src_copy, ground_pc_);

    // 3. Extract ground (erasor.cpp:260-283)
    for (int i = 0; i < p.gf_iter; i++) {
        estimate_plane_(ground_pc_);
        ground_pc_.clear();
        if (cur_tap_) {
            cur_tap_->normal.push_back(normal_[0]); cur_tap_->normal.push_back(normal_[1]); cur_tap_->normal.push_back(normal_[2]);
            cur_tap_->d.push_back(d_);
        }
        // points(n,3) * normal_ : depth-3 product accumulated in order, no FMA  [3P Eigen GEBP]
        for (size_t r = 0; r < src.size(); r++) {
            const float result = g_study[1] ? std::fmaf(src[r].z, normal_[2], std::fmaf(src[r].y, normal_[1], src[r].x * normal_[0]))
                                            : (src[r].x * normal_[0] + src[r].y * normal_[1]) + src[r].z * normal_[2];
            if (result < th_dist_d_) {
                ground_pc_.push_back(src[r]);
            } else {
                if (i == (p.gf_iter - 1)) non_ground_pc_.push_back(src[r]);
            }
        }
        if (cur_tap_) cur_tap_->n_ground.push_back(static_cast<int>(ground_pc_.size()));
    }
    dst      = ground_pc_;
    outliers = non_ground_pc_;
}

void ERASOR::merge_bins(const Bin& src1, const Bin& src2, Bin& dst) {   // erasor.cpp:296-307
    dst.max_h       = max(src1.max_h, src2.max_h);
    dst.min_h       = min(src1.min_h, src2.min_h);
    dst.is_occupied = true;
    dst.points.clear();
    for (auto const& pt : src1.points) dst.points.push_back(pt);
    for (auto const& pt : src2.points) dst.points.push_back(pt);
}

void ERASOR::r_pod2pc(const R_POD& sc, Cloud& pc) {            // erasor.cpp:309-320
    pc.clear();
    for (int theta = 0; theta < num_sectors; theta++) {
        for (int r = 0; r < num_rings; r++) {
            if (sc.at(r).at(theta).is_occupied) {
                for (auto const& pt : sc.at(r).at(theta).points) pc.push_back(pt);
            }
        }
    }
}

void ERASOR::get_outliers(Cloud& map_rejected, Cloud& curr_rejected) {   // erasor.cpp:322-327
    map_rejected  = debug_map_rejected;
    curr_rejected = debug_curr_rejected;
}

static PlaneTap new_tap(int r, int theta, size_t n) {
    PlaneTap t;
    t.ring = r; t.sector = theta; t.n_points = static_cast<int>(n);
    t.n_seeds = 0; t.lpr_height = 0; t.n_empty_fits = 0;
    return t;
}

// Version 2 (erasor.cpp:332-434)
void ERASOR::compare_vois_and_revert_ground(int /*frame*/) {
    ground_viz.clear();
    tap_planes.clear();
    tap_status.assign(static_cast<size_t>(num_rings) * num_sectors, NOT_ASSIGNED);
    tap_status_pass1.clear();
    for (int theta = 0; theta < num_sectors; theta++) {
        for (int r = 0; r < num_rings; r++) {
            Bin& bin_curr = r_pod_curr[r][theta];
            Bin& bin_map  = r_pod_map[r][theta];
            double& st    = tap_status[static_cast<size_t>(theta) * num_rings + r];

            if (bin_curr.points.size() < static_cast<size_t>(p.minimum_num_pts)) {   // :354 (size_t vs int, App. B-9)
                r_pod_selected[r][theta] = bin_map;
                st = LITTLE_NUM;
                continue;
            }
            if (bin_curr.is_occupied && bin_map.is_occupied) {
                double map_h_diff  = bin_map.max_h - bin_map.min_h;
                double curr_h_diff = bin_curr.max_h - bin_curr.min_h;
                double scan_ratio  = min(map_h_diff / curr_h_diff, curr_h_diff / map_h_diff);
                if (scan_ratio < p.scan_ratio_threshold) {
                    if (map_h_diff >= curr_h_diff) {
                        st = MAP_IS_HIGHER;
                        if (bin_map.max_h > p.th_bin_max_h) {
                            r_pod_selected[r][theta] = bin_curr;
                            if (!piecewise_ground_.empty()) piecewise_ground_.clear();
                            if (!non_ground_.empty()) non_ground_.clear();
                            tap_planes.push_back(new_tap(r, theta, bin_map.points.size()));
                            cur_tap_ = &tap_planes.back();
                            extract_ground(bin_map.points, piecewise_ground_, non_ground_);
                            cur_tap_ = nullptr;
                            r_pod_selected[r][theta].points.insert(r_pod_selected[r][theta].points.end(),
                                                                   piecewise_ground_.begin(), piecewise_ground_.end());
                            ground_viz.insert(ground_viz.end(), piecewise_ground_.begin(), piecewise_ground_.end());
                            debug_map_rejected.insert(debug_map_rejected.end(), non_ground_.begin(), non_ground_.end());
                        } else {
                            r_pod_selected[r][theta] = bin_map;
                        }
                    } else if (map_h_diff <= curr_h_diff) {
                        st = CURR_IS_HIGHER;
                        r_pod_selected[r][theta] = bin_map;
                        if (bin_curr.max_h > p.th_bin_max_h) {
                            debug_curr_rejected.insert(debug_curr_rejected.end(), bin_curr.points.begin(), bin_curr.points.end());
                        }
                    }
                } else {
                    st = MERGE_BINS;
                    Bin bin_merged;
                    merge_bins(bin_curr, bin_map, bin_merged);
                    r_pod_selected[r][theta] = bin_merged;
                }
            } else if (bin_curr.is_occupied) {
                r_pod_selected[r][theta] = bin_curr;
            } else if (bin_map.is_occupied) {
                r_pod_selected[r][theta] = bin_map;
            }
        }
    }
}

// Version 3 (erasor.cpp:438-571)
void ERASOR::compare_vois_and_revert_ground_w_block(int /*frame*/) {
    ground_viz.clear();
    tap_planes.clear();

    // 1. Update status (erasor.cpp:448-486)
    for (int theta = 0; theta < num_sectors; theta++) {
        for (int r = 0; r < num_rings; r++) {
            Bin& bin_curr = r_pod_curr[r][theta];
            Bin& bin_map  = r_pod_map[r][theta];
            if (bin_map.points.empty()) {
                r_pod_selected[r][theta].status = LITTLE_NUM;
                continue;
            }
            if (bin_curr.points.size() < static_cast<size_t>(p.minimum_num_pts)) {
                r_pod_selected[r][theta].status = LITTLE_NUM;
            } else {
                double map_h_diff  = bin_map.max_h - bin_map.min_h;
                double curr_h_diff = bin_curr.max_h - bin_curr.min_h;
                double scan_ratio  = min(map_h_diff / curr_h_diff, curr_h_diff / map_h_diff);
                if (bin_curr.is_occupied && bin_map.is_occupied) {
                    if (scan_ratio < p.scan_ratio_threshold) {
                        if (map_h_diff >= curr_h_diff) {
                            r_pod_selected[r][theta].status = MAP_IS_HIGHER;
                        } else if (map_h_diff <= curr_h_diff) {
                            r_pod_selected[r][theta].status = CURR_IS_HIGHER;
                        }
                    } else {
                        r_pod_selected[r][theta].status = MERGE_BINS;
                    }
                } else if (bin_map.is_occupied) {
                    r_pod_selected[r][theta].status = LITTLE_NUM;
                }
            }
        }
    }
    tap_status_pass1.resize(static_cast<size_t>(num_rings) * num_sectors);
    for (int theta = 0; theta < num_sectors; theta++)
        for (int r = 0; r < num_rings; r++)
            tap_status_pass1[static_cast<size_t>(theta) * num_rings + r] = r_pod_selected[r][theta].status;

    // 2. set bins (erasor.cpp:493-563)
    for (int theta = 0; theta < num_sectors; theta++) {
        for (int r = 0; r < num_rings; r++) {
            Bin& bin_curr = r_pod_curr[r][theta];
            Bin& bin_map  = r_pod_map[r][theta];

            double OCCUPANCY_STATUS = r_pod_selected[r][theta].status;
            if (OCCUPANCY_STATUS == LITTLE_NUM) {
                r_pod_selected[r][theta]        = bin_map;
                r_pod_selected[r][theta].status = LITTLE_NUM;
            } else if (OCCUPANCY_STATUS == MAP_IS_HIGHER) {
                if ((bin_map.max_h - bin_map.min_h) > 0.5) {       // hard-coded, erasor.cpp:511
                    r_pod_selected[r][theta]        = bin_curr;
                    r_pod_selected[r][theta].status = MAP_IS_HIGHER;
                    if (!piecewise_ground_.empty()) piecewise_ground_.clear();
                    if (!non_ground_.empty()) non_ground_.clear();
                    tap_planes.push_back(new_tap(r, theta, bin_map.points.size()));
                    cur_tap_ = &tap_planes.back();
                    extract_ground(bin_map.points, piecewise_ground_, non_ground_);
                    cur_tap_ = nullptr;
                    r_pod_selected[r][theta].points.insert(r_pod_selected[r][theta].points.end(),
                                                           piecewise_ground_.begin(), piecewise_ground_.end());
                    if (!p.skip_voxelize) {                          // erasor.cpp:526-528
                        Cloud tmp = r_pod_selected[r][theta].points;
                        voxelize_preserving_labels(tmp, r_pod_selected[r][theta].points, p.map_voxel_size);
                    }
                    ground_viz.insert(ground_viz.end(), piecewise_ground_.begin(), piecewise_ground_.end());
                    debug_map_rejected.insert(debug_map_rejected.end(), non_ground_.begin(), non_ground_.end());
                } else {
                    r_pod_selected[r][theta]        = bin_map;
                    r_pod_selected[r][theta].status = NOT_ASSIGNED;
                }
            } else if (OCCUPANCY_STATUS == CURR_IS_HIGHER) {
                r_pod_selected[r][theta]        = bin_map;
                r_pod_selected[r][theta].status = CURR_IS_HIGHER;
            } else if (OCCUPANCY_STATUS == MERGE_BINS) {
                if (is_dynamic_obj_close(r_pod_selected, r, theta, 1, 1)) {
                    r_pod_selected[r][theta]        = bin_map;
                    r_pod_selected[r][theta].status = BLOCKED;
                } else {
                    r_pod_selected[r][theta]        = bin_map;
                    r_pod_selected[r][theta].status = MERGE_BINS;
                }
            }
        }
    }
    tap_status.resize(static_cast<size_t>(num_rings) * num_sectors);
    for (int theta = 0; theta < num_sectors; theta++)
        for (int r = 0; r < num_rings; r++)
            tap_status[static_cast<size_t>(theta) * num_rings + r] = r_pod_selected[r][theta].status;
}

bool ERASOR::is_dynamic_obj_close(R_POD& r_pod_sel, int r_target, int theta_target, int r_range, int theta_range) {   // erasor.cpp:573-595
    std::vector<int> theta_candidates;
    for (int j = theta_target - theta_range; j <= theta_target + theta_range; j++) {
        if (j < 0) {
            theta_candidates.push_back(j + num_rings);       // sic: num_rings
        } else if (j >= num_sectors) {
            theta_candidates.push_back(j - num_rings);       // sic
        } else {
            theta_candidates.push_back(j);
        }
    }
    for (int r = std::max(0, r_target - r_range); r <= std::min(r_target + r_range, num_rings - 1); r++) {
        for (const auto& theta : theta_candidates) {
            if ((r == r_target) && (theta == theta_target)) continue;
            // FENCE: with num_rings > num_sectors the reference indexes out of range (UB);
            // skip such candidates (none of the shipped yamls can reach this).
            if (theta < 0 || theta >= num_sectors) continue;
            if (r_pod_sel[r][theta].status == CURR_IS_HIGHER) return true;
        }
    }
    return false;
}

void ERASOR::get_static_estimate(Cloud& arranged, Cloud& complement) {   // erasor.cpp:612-626
    r_pod2pc(r_pod_selected, arranged);
    arranged.insert(arranged.end(), ground_viz.begin(), ground_viz.end());
    complement = map_complement;
}

// ----------------------------------------------------------------------------
// [3P] pcl::VoxelGrid<PointXYZI> (PCL 1.8, downsample_all_data = true,
// min_points_per_voxel = 0) followed by exact 1-NN label restore
// (erasor_utils.cpp:80-114).
// Unpinned choices, fixed here and in the CUDA path alike:
//   * points of one voxel are summed in cloud order (std::sort there is unstable);
//   * 1-NN ties go to the lowest cloud index (FLANN's tie order is tree-dependent).
// ----------------------------------------------------------------------------
namespace {
struct CellKey { int i, j, k; };
struct IdxPair { unsigned idx; unsigned cloud_point_index; };
}

void voxelize_preserving_labels(const Cloud& src, Cloud& dst, double leaf_size) {
    Cloud voxelized;
    if (src.empty()) { dst.clear(); return; }
    const float leaf = static_cast<float>(leaf_size);          // setLeafSize(float,float,float)
    const float inv  = 1.0f / leaf;                            // inverse_leaf_size_ = 1 / leaf_size_
    float min_p[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, max_p[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const auto& pt : src) {                               // getMinMax3D
        min_p[0] = std::min(min_p[0], pt.x); max_p[0] = std::max(max_p[0], pt.x);
        min_p[1] = std::min(min_p[1], pt.y); max_p[1] = std::max(max_p[1], pt.y);
        min_p[2] = std::min(min_p[2], pt.z); max_p[2] = std::max(max_p[2], pt.z);
    }
    const int64_t dx = static_cast<int64_t>((max_p[0] - min_p[0]) * inv) + 1;
    const int64_t dy = static_cast<int64_t>((max_p[1] - min_p[1]) * inv) + 1;
    const int64_t dz = static_cast<int64_t>((max_p[2] - min_p[2]) * inv) + 1;
    bool overflow = (dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max());
    if (overflow) {
        voxelized = src;                                       // "Leaf size is too small": output = input
    } else {
        int min_b[3], max_b[3], div_b[3], mul[3];
        for (int a = 0; a < 3; ++a) {
            min_b[a] = static_cast<int>(std::floor(min_p[a] * inv));
            max_b[a] = static_cast<int>(std::floor(max_p[a] * inv));
            div_b[a] = max_b[a] - min_b[a] + 1;
        }
        mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
        std::vector<IdxPair> index_vector;
        index_vector.reserve(src.size());
        for (unsigned i = 0; i < src.size(); ++i) {
            const int ijk0 = static_cast<int>(std::floor(src[i].x * inv) - static_cast<float>(min_b[0]));
            const int ijk1 = static_cast<int>(std::floor(src[i].y * inv) - static_cast<float>(min_b[1]));
            const int ijk2 = static_cast<int>(std::floor(src[i].z * inv) - static_cast<float>(min_b[2]));
            const int idx  = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
            index_vector.push_back(IdxPair{static_cast<unsigned>(idx), i});
        }
        std::stable_sort(index_vector.begin(), index_vector.end(),
                         [](const IdxPair& a, const IdxPair& b) { return a.idx < b.idx; });
        size_t index = 0;
        while (index < index_vector.size()) {
            size_t i = index + 1;
            while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx) ++i;
            float sx = 0, sy = 0, sz = 0, si = 0;               // CentroidPoint accumulators (float)
            for (size_t li = index; li < i; ++li) {
                const PointXYZI& q = src[index_vector[li].cloud_point_index];
                sx += q.x; sy += q.y; sz += q.z; si += q.intensity;
            }
            const float n = static_cast<float>(i - index);
            PointXYZI c{};
            c.x = sx / n; c.y = sy / n; c.z = sz / n; c.intensity = si / n;
            c.src = SRC_NONE;
            voxelized.push_back(c);
            index = i;
        }
    }
    // 2. exact nearest source point per centroid; uniform grid of pitch `leaf` as the search index
    struct KeyHash {
        size_t operator()(const int64_t& k) const { return std::hash<int64_t>()(k); }
    };
    auto cell_of = [&](float v) { return static_cast<int64_t>(std::floor(v * inv)); };
    auto pack = [](int64_t a, int64_t b, int64_t c) {
        return ((a & 0x1FFFFF) << 42) | ((b & 0x1FFFFF) << 21) | (c & 0x1FFFFF);
    };
    std::unordered_map<int64_t, std::vector<unsigned>, KeyHash> grid;
    grid.reserve(src.size());
    for (unsigned i = 0; i < src.size(); ++i)
        grid[pack(cell_of(src[i].x), cell_of(src[i].y), cell_of(src[i].z))].push_back(i);

    Cloud reassigned;
    reassigned.reserve(voxelized.size());
    for (const auto& pt : voxelized) {
        const int64_t ci = cell_of(pt.x), cj = cell_of(pt.y), ck = cell_of(pt.z);
        float    best_d = std::numeric_limits<float>::infinity();
        unsigned best_i = 0xFFFFFFFFu;
        auto scan = [&](int rad_lo, int rad_hi) {
            for (int a = -rad_hi; a <= rad_hi; ++a)
                for (int b = -rad_hi; b <= rad_hi; ++b)
                    for (int c = -rad_hi; c <= rad_hi; ++c) {
                        const int cheb = std::max(std::abs(a), std::max(std::abs(b), std::abs(c)));
                        if (cheb < rad_lo) continue;
                        auto itc = grid.find(pack(ci + a, cj + b, ck + c));
                        if (itc == grid.end()) continue;
                        for (unsigned i : itc->second) {
                            const float ddx = pt.x - src[i].x, ddy = pt.y - src[i].y, ddz = pt.z - src[i].z;
                            const float d = (ddx * ddx + ddy * ddy) + ddz * ddz;    // flann::L2_Simple<float>
                            if (d < best_d || (d == best_d && (g_study[2] ? (best_i == 0xFFFFFFFFu || i > best_i) : i < best_i))) { best_d = d; best_i = i; }
                        }
                    }
        };
        // grow the shell until no unvisited cell can hold a closer point
        int rad = 0;
        scan(0, 0);
        while (true) {
            // every point closer than `reach` has been seen; the 1e-4 margin keeps the early exit identical to a
            // brute-force scan even when float rounding puts an unseen point's distance exactly on best_d
            const float reach = static_cast<float>(rad) * leaf * 0.9999f;
            if (best_i != 0xFFFFFFFFu && best_d < reach * reach) break;
            ++rad;
            scan(rad, rad);
            if (rad > 64 && best_i != 0xFFFFFFFFu) break;
            if (rad > 4096) break;
        }
        if (best_i != 0xFFFFFFFFu) {
            PointXYZI updated = pt;
            updated.intensity = src[best_i].intensity;         // erasor_utils.cpp:109
            reassigned.push_back(updated);
        }
    }
    dst = reassigned;
}

// ----------------------------------------------------------------------------
// [3P] pcl::transformPointCloud (PCL 1.8 scalar path), Eigen inverse, tf quaternion
// ----------------------------------------------------------------------------
void transform_point_cloud(const Cloud& in, Cloud& out, const float T[16]) {
    if (&in != &out) out = in;
    for (size_t i = 0; i < out.size(); ++i) {
        const float x = in[i].x, y = in[i].y, z = in[i].z;
        out[i].x = static_cast<float>(T[0] * x + T[1] * y + T[2] * z + T[3]);
        out[i].y = static_cast<float>(T[4] * x + T[5] * y + T[6] * z + T[7]);
        out[i].z = static_cast<float>(T[8] * x + T[9] * y + T[10] * z + T[11]);
    }
}

void invert_4x4(const float m[16], float inv_out[16]) {
    // general cofactor inverse in float (Eigen uses an SSE cofactor kernel for Matrix4f; bits unpinned)
    float inv[16];
    inv[0]  =  m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4]  = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8]  =  m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1]  = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5]  =  m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9]  = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] =  m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2]  =  m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6]  = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] =  m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3]  = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7]  =  m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] =  m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    det = 1.0f / det;
    for (int i = 0; i < 16; i++) inv_out[i] = inv[i] * det;
}

void geo_pose_to_matrix(const double pose[7], float T[16]) {   // erasor_utils.cpp:35-55 via tf::Matrix3x3(q)
    const double qx = pose[3], qy = pose[4], qz = pose[5], qw = pose[6];
    // tf::Matrix3x3::setRotation
    const double d  = qx * qx + qy * qy + qz * qz + qw * qw;
    const double s  = 2.0 / d;
    const double xs = qx * s, ys = qy * s, zs = qz * s;
    const double wx = qw * xs, wy = qw * ys, wz = qw * zs;
    const double xx = qx * xs, xy = qx * ys, xz = qx * zs;
    const double yy = qy * ys, yz = qy * zs, zz = qz * zs;
    const double m[9] = {1.0 - (yy + zz), xy - wz, xz + wy,
                         xy + wz, 1.0 - (xx + zz), yz - wx,
                         xz - wy, yz + wx, 1.0 - (xx + yy)};
    for (int i = 0; i < 16; ++i) T[i] = 0.0f;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T[r * 4 + c] = static_cast<float>(m[r * 3 + c]);
    T[3] = static_cast<float>(pose[0]); T[7] = static_cast<float>(pose[1]); T[11] = static_cast<float>(pose[2]);
    T[15] = 1.0f;
}

bool is_dynamic_label(float intensity) {                      // erasor_utils.cpp:63-72
    const uint32_t float2int      = static_cast<uint32_t>(intensity);
    const uint32_t semantic_label = float2int & 0xFFFF;
    return semantic_label >= 252 && semantic_label <= 259;
}

// ----------------------------------------------------------------------------
// OfflineMapUpdater (the caller), ROS stripped
// ----------------------------------------------------------------------------
static void mat4_mul(const float A[16], const float B[16], float C[16]) {
    float t[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) acc += A[r * 4 + k] * B[k * 4 + c];
            t[r * 4 + c] = acc;
        }
    std::memcpy(C, t, sizeof(t));
}

OfflineMapUpdater::OfflineMapUpdater(const UpdaterParams& up, const Params& ep, const Cloud& initial_map)
    : erasor_(ep), up_(up) {
    // set_params (:63-105): tf_lidar2body_ = geoPose2eigen(pose) * Identity
    float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, G[16];
    geo_pose_to_matrix(up.lidar2body, G);
    mat4_mul(G, I, tf_lidar2body_);
    std::memcpy(tf_body2origin_, I, sizeof(I));
    // load_global_map (:107-167), outdoor
    num_pcs_init_ = initial_map.size();
    map_arranged_ = initial_map;
    if (up.is_large_scale) map_arranged_global_ = map_arranged_;
}

void OfflineMapUpdater::set_submap(const Cloud& map_global, Cloud& submap, Cloud& submap_complement,
                                   double x, double y, double submap_size) {   // :360-379
    submap.clear();
    submap_complement.clear();
    for (const auto pt : map_global) {
        double diff_x = fabs(x - pt.x);
        double diff_y = fabs(y - pt.y);
        if ((diff_x < submap_size) && (diff_y < submap_size)) submap.emplace_back(pt);
        else submap_complement.emplace_back(pt);
    }
}

void OfflineMapUpdater::reassign_submap(double pose_x, double pose_y) {   // :332-358
    if (is_submap_not_initialized_) {
        set_submap(map_arranged_global_, map_arranged_, map_arranged_complement_, pose_x, pose_y, up_.submap_size);
        submap_center_x_ = pose_x;
        submap_center_y_ = pose_y;
        is_submap_not_initialized_ = false;
    } else {
        double diff_x = std::abs(submap_center_x_ - pose_x);
        double diff_y = std::abs(submap_center_y_ - pose_y);
        const double half_size = up_.submap_size / 2.0;
        if ((diff_x > half_size) || (diff_y > half_size)) {
            map_arranged_global_.clear();
            map_arranged_global_ = map_arranged_;
            map_arranged_global_.insert(map_arranged_global_.end(), map_arranged_complement_.begin(), map_arranged_complement_.end());
            set_submap(map_arranged_global_, map_arranged_, map_arranged_complement_, pose_x, pose_y, up_.submap_size);
            submap_center_x_ = pose_x;
            submap_center_y_ = pose_y;
        }
    }
}

void OfflineMapUpdater::fetch_VoI(double x_criterion, double y_criterion, Cloud& dst, Cloud& outskirts) {   // :381-438
    if (!dst.empty()) dst.clear();
    if (!outskirts.empty()) outskirts.clear();
    if (!map_voi_wrt_origin_.empty()) map_voi_wrt_origin_.clear();
    double max_dist_square = pow(up_.max_range + 0.0, 2);
    for (auto const& pt : map_arranged_) {
        double dist_square = pow(pt.x - x_criterion, 2) + pow(pt.y - y_criterion, 2);
        if (dist_square < max_dist_square) map_voi_wrt_origin_.emplace_back(pt);
        else outskirts.emplace_back(pt);
    }
    float Tinv[16];
    invert_4x4(tf_body2origin_, Tinv);
    Cloud transformed;
    transform_point_cloud(map_voi_wrt_origin_, transformed, Tinv);
    dst = transformed;
}

void OfflineMapUpdater::body2origin(const Cloud src, Cloud& dst) {   // :441-449
    Cloud transformed;
    transform_point_cloud(src, transformed, tf_body2origin_);
    dst = transformed;
}

bool OfflineMapUpdater::callback_node(int seq, const double odom[7], const Cloud& lidar) {   // :203-330
    stack_count++;
    if (stack_count % up_.removal_interval != 0) return false;
    using clk = std::chrono::steady_clock;

    geo_pose_to_matrix(odom, tf_body2origin_);                                   // :219
    Cloud query_voxel, query_body;
    voxelize_preserving_labels(lidar, query_voxel, up_.query_voxel_size);        // :238
    transform_point_cloud(query_voxel, query_body, tf_lidar2body_);              // :240
    query_voi_ = query_body;                                                     // :241
    // tag query points with their position in query_voi_
    for (size_t i = 0; i < query_voi_.size(); ++i) query_voi_[i].src = static_cast<uint32_t>(i) | SRC_QUERY_BIT;

    double x_curr = tf_body2origin_[3];                                          // :246
    double y_curr = tf_body2origin_[7];
    if (up_.is_large_scale) reassign_submap(x_curr, y_curr);                     // :249-251

    auto t0 = clk::now();
    fetch_VoI(x_curr, y_curr, map_voi_, map_outskirts_);                         // :254
    auto t1 = clk::now();
    for (size_t i = 0; i < map_voi_.size(); ++i) map_voi_[i].src = static_cast<uint32_t>(i);
    last_voi_seconds = std::chrono::duration<double>(t1 - t0).count();

    auto s0 = clk::now();
    erasor_.set_inputs(map_voi_, query_voi_);                                    // :266
    if (up_.version == 2) {
        erasor_.compare_vois_and_revert_ground(seq);
        erasor_.get_static_estimate(map_static_estimate_, map_egocentric_complement_);
    } else if (up_.version == 3) {
        erasor_.compare_vois_and_revert_ground_w_block(seq);
        erasor_.get_static_estimate(map_static_estimate_, map_egocentric_complement_);
    } else {
        throw std::invalid_argument("Other version is not implemented!");
    }
    auto s1 = clk::now();
    last_erasor_seconds = std::chrono::duration<double>(s1 - s0).count();        // "ERASOR takes" :264-279

    map_filtered_ = map_static_estimate_;                                        // :281
    map_filtered_.insert(map_filtered_.end(), map_egocentric_complement_.begin(), map_egocentric_complement_.end());
    erasor_.get_outliers(map_rejected_, query_rejected_);                        // :284
    body2origin(map_filtered_, map_filtered_);                                   // :286-288
    body2origin(map_rejected_, map_rejected_);
    body2origin(query_rejected_, query_rejected_);
    map_arranged_ = map_filtered_;                                               // :290
    map_arranged_.insert(map_arranged_.end(), map_outskirts_.begin(), map_outskirts_.end());
    total_map_rejected_.insert(total_map_rejected_.end(), map_rejected_.begin(), map_rejected_.end());       // :297-298
    total_query_rejected_.insert(total_query_rejected_.end(), query_rejected_.begin(), query_rejected_.end());
    return true;
}

void OfflineMapUpdater::save_static_map(float voxel_size, Cloud& map_to_be_saved) {   // :174-196
    Cloud src;
    if (up_.is_large_scale) {
        src = map_arranged_;
        src.insert(src.end(), map_arranged_complement_.begin(), map_arranged_complement_.end());
    } else {
        src = map_arranged_;
    }
    voxelize_preserving_labels(src, map_to_be_saved, voxel_size);
}

// ----------------------------------------------------------------------------
// mapgen (src/mapgen/mapgen.hpp), ROS stripped
// ----------------------------------------------------------------------------
void NaiveMapGen::accum_point_cloud(const double odom[7], const Cloud& lidar) {
    // tf_lidar2origin (:209-214): identity with z + 1.73
    const float tf_lidar2origin[16] = {1, 0, 0, 0,  0, 1, 0, 0,  0, 0, 1, 1.73f,  0, 0, 0, 1};
    // "To remove some noisy points in the vicinity of the vehicles" (:218-229): float threshold, double distance
    const float max_dist_square = static_cast<float>(std::pow(2.7, 2));            // CAR_BODY_SIZE 2.7 (:8)
    Cloud outliers;
    outliers.reserve(lidar.size());
    for (const auto& pt : lidar) {
        const double dist_square = std::pow(static_cast<double>(pt.x), 2) + std::pow(static_cast<double>(pt.y), 2);
        if (!(dist_square < max_dist_square)) outliers.push_back(pt);               // inliers are dropped
    }
    Cloud lifted, world;
    transform_point_cloud(outliers, lifted, tf_lidar2origin);                       // :231-232
    float pose[16];
    geo_pose_to_matrix(odom, pose);                                                 // :234
    transform_point_cloud(lifted, world, pose);                                     // :236-237
    voxelize_preserving_labels(world, cloud_curr, 0.2);                             // :239 (fixed 0.2, not leafsize)
    if (is_initial_) {
        cloud_map   = cloud_curr;                                                   // :241-243
        is_initial_ = false;
    } else {
        cloud_map.insert(cloud_map.end(), cloud_curr.begin(), cloud_curr.end());    // :245
        if (is_large_scale_) {
            if (cnt_voxel_++ % 500 == 0) {                                          // :247-258
                Cloud vox;
                voxelize_preserving_labels(cloud_map, vox, leafsize_);
                cloud_maps.push_back(vox);
                cloud_map.clear();
            }
        }
    }
}

void NaiveMapGen::save_naive_map(Cloud& original, Cloud& voxelized) const {
    original.clear();
    if (is_large_scale_) {                                                          // :275-281
        for (const auto& submap : cloud_maps) original.insert(original.end(), submap.begin(), submap.end());
        original.insert(original.end(), cloud_map.begin(), cloud_map.end());
    } else {
        original = cloud_map;                                                       // :283
    }
    voxelize_preserving_labels(original, voxelized, leafsize_);                     // :296
}

}  // namespace oracle
