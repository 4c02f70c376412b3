// ============================================================================
//  oracle/oracle_capi.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE
//  extern "C" shim so that tests/ and bench.py's cpu_baseline leg can drive the
//  oracle through ctypes.  Clouds cross the boundary as float[n][4] = x,y,z,intensity.
// ============================================================================
#include "erasor_oracle.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>

using namespace oracle;

namespace {
struct Session {
    std::unique_ptr<ERASOR> e;
    Cloud map_voi, query_voi;
    Cloud arranged, complement, map_rejected, curr_rejected;
};
struct USession {
    std::unique_ptr<OfflineMapUpdater> u;
};

void to_cloud(const float* xyzi, size_t n, uint32_t tag_bit, Cloud& c) {
    c.resize(n);
    for (size_t i = 0; i < n; ++i) {
        PointXYZI p{};
        p.x = xyzi[4 * i + 0]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3];
        p.src = static_cast<uint32_t>(i) | tag_bit;
        c[i] = p;
    }
}
size_t from_cloud(const Cloud& c, float* xyzi, uint32_t* src, size_t cap) {
    const size_t n = c.size() < cap ? c.size() : cap;
    for (size_t i = 0; i < n; ++i) {
        if (xyzi) { xyzi[4 * i + 0] = c[i].x; xyzi[4 * i + 1] = c[i].y; xyzi[4 * i + 2] = c[i].z; xyzi[4 * i + 3] = c[i].intensity; }
        if (src) src[i] = c[i].src;
    }
    return c.size();
}
const Cloud* pick(Session* s, int which) {
    switch (which) {
        case 0: return &s->arranged;
        case 1: return &s->complement;
        case 2: return &s->map_rejected;
        case 3: return &s->curr_rejected;
        case 4: return &s->e->ground_viz;
        default: return nullptr;
    }
}
}  // namespace

extern "C" {

void* oracle_create(const Params* p) {
    auto* s = new Session();
    s->e.reset(new ERASOR(*p));
    return s;
}
void oracle_destroy(void* h) { delete static_cast<Session*>(h); }

// One pass of the path: set_inputs + compare (version from params) + get_static_estimate + get_outliers.
// Returns seconds spent in set_inputs..get_static_estimate (the reference's "ERASOR takes" span).
double oracle_run(void* h, const float* map_xyzi, size_t n_map, const float* query_xyzi, size_t n_query, int frame) {
    auto* s = static_cast<Session*>(h);
    to_cloud(map_xyzi, n_map, 0u, s->map_voi);
    to_cloud(query_xyzi, n_query, SRC_QUERY_BIT, s->query_voi);
    auto t0 = std::chrono::steady_clock::now();
    s->e->set_inputs(s->map_voi, s->query_voi);
    if (s->e->p.version == 2) s->e->compare_vois_and_revert_ground(frame);
    else                      s->e->compare_vois_and_revert_ground_w_block(frame);
    s->e->get_static_estimate(s->arranged, s->complement);
    auto t1 = std::chrono::steady_clock::now();
    s->e->get_outliers(s->map_rejected, s->curr_rejected);
    return std::chrono::duration<double>(t1 - t0).count();
}

// bin id per point (sector*R + ring, -1 = not binned)
void oracle_get_bin_of_point(void* h, int which /*0 map, 1 query*/, int32_t* out) {
    auto* s = static_cast<Session*>(h);
    const auto& v = which == 0 ? s->e->tap_bin_of_map : s->e->tap_bin_of_query;
    std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
}
// per-bin tables, index sector*R + ring; min/max are the doubles the reference holds (+-1e13 when empty)
void oracle_get_bins(void* h, int which /*0 map, 1 query*/, double* min_h, double* max_h, uint32_t* count, uint8_t* occupied) {
    auto* s = static_cast<Session*>(h);
    const R_POD& rp = which == 0 ? s->e->r_pod_map : s->e->r_pod_curr;
    const int R = s->e->num_rings, S = s->e->num_sectors;
    for (int t = 0; t < S; ++t)
        for (int r = 0; r < R; ++r) {
            const Bin& b = rp[r][t];
            const size_t i = static_cast<size_t>(t) * R + r;
            if (min_h) min_h[i] = b.min_h;
            if (max_h) max_h[i] = b.max_h;
            if (count) count[i] = static_cast<uint32_t>(b.points.size());
            if (occupied) occupied[i] = b.is_occupied ? 1 : 0;
        }
}
void oracle_get_status(void* h, double* status, double* status_pass1) {
    auto* s = static_cast<Session*>(h);
    if (status) std::memcpy(status, s->e->tap_status.data(), s->e->tap_status.size() * sizeof(double));
    if (status_pass1 && !s->e->tap_status_pass1.empty())
        std::memcpy(status_pass1, s->e->tap_status_pass1.data(), s->e->tap_status_pass1.size() * sizeof(double));
}
long oracle_get_negzero_fenced(void* h) { return static_cast<Session*>(h)->e->tap_negzero_fenced; }

int oracle_num_planes(void* h) { return static_cast<int>(static_cast<Session*>(h)->e->tap_planes.size()); }
// per flagged bin i: bin id, n_points, n_seeds, lpr, n_empty_fits; normal_d[iter][4]; n_ground[iter]
void oracle_get_plane(void* h, int i, int32_t* bin, int32_t* n_points, int32_t* n_seeds, double* lpr, int32_t* n_empty,
                      double* normal_d /*gf_iter*4*/, int32_t* n_ground /*gf_iter*/) {
    auto* s = static_cast<Session*>(h);
    const PlaneTap& t = s->e->tap_planes[i];
    *bin = t.sector * s->e->num_rings + t.ring;
    *n_points = t.n_points; *n_seeds = t.n_seeds; *lpr = t.lpr_height; *n_empty = t.n_empty_fits;
    for (size_t k = 0; k < t.d.size(); ++k) {
        normal_d[4 * k + 0] = t.normal[3 * k + 0];
        normal_d[4 * k + 1] = t.normal[3 * k + 1];
        normal_d[4 * k + 2] = t.normal[3 * k + 2];
        normal_d[4 * k + 3] = t.d[k];
        n_ground[k] = t.n_ground[k];
    }
}
// which: 0 arranged, 1 complement, 2 map_rejected, 3 curr_rejected, 4 ground_viz
size_t oracle_cloud_size(void* h, int which) { return pick(static_cast<Session*>(h), which)->size(); }
size_t oracle_get_cloud(void* h, int which, float* xyzi, uint32_t* src, size_t cap) {
    return from_cloud(*pick(static_cast<Session*>(h), which), xyzi, src, cap);
}

// blast-radius study switches (scripts/blast_radius.py only)
void oracle_set_study(int reverse_sort_ties, int fma_classification, int nn_tie_highest) {
    g_study[0] = reverse_sort_ties; g_study[1] = fma_classification; g_study[2] = nn_tie_highest;
}

// ---- unit entry points for the [3P] restatements ----
unsigned oracle_mean_cov(const float* xyzi, size_t n, int mode, float* cov9, float* mean4) {
    Cloud c; to_cloud(xyzi, n, 0u, c);
    for (int i = 0; i < 9; ++i) cov9[i] = 0;
    for (int i = 0; i < 4; ++i) mean4[i] = 0;
    return compute_mean_and_covariance(c, cov9, mean4, mode);
}
void oracle_jacobi_svd(const float* A9, float* U9, float* sv3) { jacobi_svd_3x3_full_u(A9, U9, sv3); }
size_t oracle_voxelize(const float* xyzi, size_t n, double leaf, float* out_xyzi, size_t cap) {
    Cloud c, d; to_cloud(xyzi, n, 0u, c);
    voxelize_preserving_labels(c, d, leaf);
    return from_cloud(d, out_xyzi, nullptr, cap);
}
void oracle_transform(const float* xyzi, size_t n, const float* T16, float* out_xyzi) {
    Cloud c, d; to_cloud(xyzi, n, 0u, c);
    transform_point_cloud(c, d, T16);
    from_cloud(d, out_xyzi, nullptr, n);
}
void oracle_pose_to_matrix(const double* pose7, float* T16) { geo_pose_to_matrix(pose7, T16); }
// OfflineMapUpdater::fetch_VoI as a free function (OfflineMapUpdater.cpp:381-438 with the pose handling of :219,246-247):
// 2-D radius cut around the body position in double on float differences, then origin -> body with the float inverse.
// Writes the VoI (body frame) and the map index of every VoI point; returns the VoI size (outputs truncated at cap).
size_t oracle_fetch_voi(const float* map_xyzi, size_t n_map, const double* pose7, double max_range, float* voi_xyzi, uint32_t* voi_index, size_t cap) {
    float T[16], Tinv[16];
    geo_pose_to_matrix(pose7, T);
    const double x_criterion = T[3], y_criterion = T[7];
    const double max_dist_square = std::pow(max_range + 0.0, 2);
    Cloud sel;
    for (size_t i = 0; i < n_map; ++i) {
        PointXYZI p{};
        p.x = map_xyzi[4 * i + 0]; p.y = map_xyzi[4 * i + 1]; p.z = map_xyzi[4 * i + 2]; p.intensity = map_xyzi[4 * i + 3];
        p.src = static_cast<uint32_t>(i);
        const double dist_square = std::pow(p.x - x_criterion, 2) + std::pow(p.y - y_criterion, 2);
        if (dist_square < max_dist_square) sel.push_back(p);
    }
    invert_4x4(T, Tinv);
    Cloud out;
    transform_point_cloud(sel, out, Tinv);
    return from_cloud(out, voi_xyzi, voi_index, cap);
}
void oracle_invert4(const float* T16, float* out16) { invert_4x4(T16, out16); }
// extract_ground on a free-standing cloud (R-GPF unit test): returns ground flags per src point
void oracle_extract_ground(void* h, const float* xyzi, size_t n, uint8_t* is_ground) {
    auto* s = static_cast<Session*>(h);
    Cloud c, g, o; to_cloud(xyzi, n, 0u, c);
    s->e->extract_ground(c, g, o);
    std::memset(is_ground, 0, n);
    for (const auto& p : g) is_ground[p.src] = 1;
}

// ---- the caller loop ----
void* oracle_updater_create(const UpdaterParams* up, const Params* ep, const float* map_xyzi, size_t n_map) {
    auto* s = new USession();
    Cloud m; to_cloud(map_xyzi, n_map, 0u, m);
    s->u.reset(new OfflineMapUpdater(*up, *ep, m));
    return s;
}
void oracle_updater_destroy(void* h) { delete static_cast<USession*>(h); }
int oracle_updater_callback_node(void* h, int seq, const double* odom7, const float* lidar_xyzi, size_t n) {
    auto* s = static_cast<USession*>(h);
    Cloud l; to_cloud(lidar_xyzi, n, 0u, l);
    return s->u->callback_node(seq, odom7, l) ? 1 : 0;
}
double oracle_updater_last_erasor_seconds(void* h) { return static_cast<USession*>(h)->u->last_erasor_seconds; }
double oracle_updater_last_voi_seconds(void* h) { return static_cast<USession*>(h)->u->last_voi_seconds; }
// which: 0 map_arranged_, 1 map_voi_, 2 query_voi_, 3 map_static_estimate_, 4 map_egocentric_complement_,
//        5 map_rejected_ (origin frame), 6 total_map_rejected_, 7 map_outskirts_, 8 map_arranged_complement_
static const Cloud* upick(USession* s, int which) {
    switch (which) {
        case 0: return &s->u->map_arranged_;
        case 1: return &s->u->map_voi_;
        case 2: return &s->u->query_voi_;
        case 3: return &s->u->map_static_estimate_;
        case 4: return &s->u->map_egocentric_complement_;
        case 5: return &s->u->map_rejected_;
        case 6: return &s->u->total_map_rejected_;
        case 7: return &s->u->map_outskirts_;
        case 8: return &s->u->map_arranged_complement_;
        default: return nullptr;
    }
}
size_t oracle_updater_cloud_size(void* h, int which) { return upick(static_cast<USession*>(h), which)->size(); }
size_t oracle_updater_get_cloud(void* h, int which, float* xyzi, uint32_t* src, size_t cap) {
    return from_cloud(*upick(static_cast<USession*>(h), which), xyzi, src, cap);
}
size_t oracle_updater_save_static_map(void* h, float voxel_size, float* out_xyzi, size_t cap) {
    auto* s = static_cast<USession*>(h);
    Cloud out;
    s->u->save_static_map(voxel_size, out);
    return from_cloud(out, out_xyzi, nullptr, cap);
}

// ---- mapgen ----
void* oracle_mapgen_create(float leafsize, int is_large_scale) { return new NaiveMapGen(leafsize, is_large_scale != 0); }
void  oracle_mapgen_destroy(void* g) { delete static_cast<NaiveMapGen*>(g); }
void  oracle_mapgen_accum(void* g, const double* odom7, const float* xyzi, size_t n) {
    Cloud c; to_cloud(xyzi, n, 0u, c);
    static_cast<NaiveMapGen*>(g)->accum_point_cloud(odom7, c);
}
// which: 0 cloud_map, 1 cloud_curr, 2 saved original, 3 saved voxelized
size_t oracle_mapgen_get(void* g, int which, float* out_xyzi, size_t cap) {
    auto* m = static_cast<NaiveMapGen*>(g);
    if (which == 0) return from_cloud(m->cloud_map, out_xyzi, nullptr, cap);
    if (which == 1) return from_cloud(m->cloud_curr, out_xyzi, nullptr, cap);
    Cloud orig, vox;
    m->save_naive_map(orig, vox);
    return from_cloud(which == 2 ? orig : vox, out_xyzi, nullptr, cap);
}

}  // extern "C"
