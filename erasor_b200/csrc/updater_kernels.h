// updater_kernels.h -- launch wrappers of the caller-side device steps (definitions in updater_kernels.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace erasor {

struct Mat4 { float m[16]; };            // row-major 4x4, like Eigen::Matrix4f(r, c) = m[4 r + c]

enum { PART_RADIUS = 0, PART_SUBMAP = 1, PART_NOT_NEAR = 2 };
struct PartPred {
    int    kind;
    int    pad_;
    double x, y;                         // criterion point
    double limit;                        // PART_RADIUS: max_dist_square; PART_SUBMAP: submap_size; PART_NOT_NEAR: squared vehicle-body radius
};

struct VoxGrid {                         // pcl::VoxelGrid state of one filter call, kept on the device
    uint32_t mn[3], mx[3];               // order-preserving encodings of the cloud's min / max
    float    leaf, inv;
    int      min_b[3], div[3];
    uint32_t n_vox;
    int      overflow;
};

size_t partition_tmp_words(uint32_t n);
// stable partition of in[0..n): pred-true points (optionally through the affine T_sel) to out_sel, the rest to out_rest,
// both in source order; *d_total_sel (device) receives the number selected.
cudaError_t launch_partition(cudaStream_t st, const PartPred& P, const Mat4& T_sel, bool transform_sel, const float4* in, uint32_t n,
                             uint32_t* chunk_tmp, uint32_t* d_total_sel, float4* out_sel, float4* out_rest);
cudaError_t launch_affine_copy(cudaStream_t st, const Mat4& T, bool do_transform, const float4* in, float4* out, uint32_t n);

size_t voxelize_tmp_bytes(uint32_t n);
int    voxelize_num_launches();
// erasor_utils::voxelize_preserving_labels: out[0..*d_n_out) in ascending voxel key, labels restored by exact 1-NN
cudaError_t launch_voxelize(cudaStream_t st, const float4* in, uint32_t n, float leaf, VoxGrid* grid, void* tmp, float4* out, uint32_t* d_n_out);

}  // namespace erasor
