// pose_math.h -- host-side pose arithmetic of the caller (reference erasor_utils.cpp:35-55, OfflineMapUpdater.cpp:219,
// 246-247, 434-436), shared by the sequential updater (updater_capi.cu) and the map-resident batch mode
// (erasor_capi.cu, erasor_process_nodes) so that both hand the device bit-identical transforms.
// Compile the including translation unit with -ffp-contract=off: the oracle restates the same expressions without FMA.
#pragma once
#include <cmath>

#include "device_types.h"
#include "updater_kernels.h"

namespace erasor {

// erasor_utils::geoPose2eigen via tf::Matrix3x3(q) (erasor_utils.cpp:35-55): double quaternion math, cast to float
inline void pose_to_mat(const double pose[7], Mat4& T) {
    const double qx = pose[3], qy = pose[4], qz = pose[5], qw = pose[6];
    const double d = qx * qx + qy * qy + qz * qz + qw * qw, s = 2.0 / d;
    const double xs = qx * s, ys = qy * s, zs = qz * s;
    const double wx = qw * xs, wy = qw * ys, wz = qw * zs, xx = qx * xs, xy = qx * ys, xz = qx * zs, yy = qy * ys, yz = qy * zs, zz = qz * zs;
    const double m[9] = {1.0 - (yy + zz), xy - wz, xz + wy, xy + wz, 1.0 - (xx + zz), yz - wx, xz - wy, yz + wx, 1.0 - (xx + yy)};
    for (int i = 0; i < 16; ++i) T.m[i] = 0.0f;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T.m[r * 4 + c] = (float)m[r * 3 + c];
    T.m[3] = (float)pose[0]; T.m[7] = (float)pose[1]; T.m[11] = (float)pose[2]; T.m[15] = 1.0f;
}
inline void mat_mul(const Mat4& A, const Mat4& B, Mat4& C) {
    Mat4 t;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { float acc = 0.0f; for (int k = 0; k < 4; ++k) acc += A.m[r * 4 + k] * B.m[k * 4 + c]; t.m[r * 4 + c] = acc; }
    C = t;
}
// Eigen::Matrix4f::inverse() stand-in: general cofactor inverse in float (bits of Eigen's SSE kernel are unpinned)
inline void mat_inv(const Mat4& M, Mat4& O) {
    const float* m = M.m; float inv[16];
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
    for (int i = 0; i < 16; ++i) O.m[i] = inv[i] * det;
}

// what fetch_VoI needs for one node (OfflineMapUpdater.cpp:219,246-247,381-438): criterion point, squared radius, origin -> body rows
inline void node_pose_of(const double odom7[7], double voi_max_range, NodePose& out) {
    Mat4 T, Tinv;
    pose_to_mat(odom7, T);
    mat_inv(T, Tinv);
    out.px = T.m[3]; out.py = T.m[7];                       // double x_curr = tf_body2origin_(0, 3) (float -> double)
    out.limit = std::pow(voi_max_range + 0.0, 2);
    for (int i = 0; i < 12; ++i) out.T[i] = Tinv.m[i];
    // float pre-test of the radius cut: d2f = fmaf(dyf, dyf, dxf * dxf) on float differences is within 4 * 2^-24 relative of
    // the reference's double expression; a 1e-6 band on either side of the limit decides everything else exactly in double
    out.pxf = T.m[3]; out.pyf = T.m[7];
    const float finf = INFINITY;
    out.lim_lo = std::nextafterf((float)(out.limit * (1.0 - 1.0e-6)), -finf);
    out.lim_hi = std::nextafterf((float)(out.limit * (1.0 + 1.0e-6)), finf);
}

}  // namespace erasor
