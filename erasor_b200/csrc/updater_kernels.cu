// updater_kernels.cu -- device versions of the caller-side steps either side of the R-POD -> SRT -> R-GPF path
// (SURVEY.md section 8f, rows 1-3), so that the map never leaves HBM between frames:
//
//   U1  stable partition of the map by a per-point predicate, with the float affine of pcl::transformPointCloud
//       fused into the copy of the selected side:
//         * OfflineMapUpdater::fetch_VoI  ("naive" mode)          reference OfflineMapUpdater.cpp:381-438
//         * OfflineMapUpdater::set_submap (large-scale window)    reference OfflineMapUpdater.cpp:360-379
//   U2  affine copy (body2origin, lidar2body: pcl::transformPointCloud)             :240, :286-288, :441-449
//   U3  erasor_utils::voxelize_preserving_labels for clouds of any size             erasor_utils.cpp:80-114
//       (pcl::VoxelGrid keys -> stable global radix sort -> centroids -> exact 1-NN label restore);
//       used on the query scan (:238) and by save_static_map (:186)
//
// Arithmetic is spelled with round-to-nearest intrinsics (no FMA contraction), exactly as restated in oracle/.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "updater_kernels.h"

namespace erasor {

#define FULL_MASK 0xFFFFFFFFu
#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FD(a, b) __fdiv_rn((a), (b))

// pcl::transformPointCloud, PCL 1.8 scalar path: ((m0*x + m1*y) + m2*z) + m3, float, no contraction
__device__ __forceinline__ float4 affine(const Mat4& T, float4 p) {
    float4 o;
    o.x = FA(FA(FA(FM(T.m[0], p.x), FM(T.m[1], p.y)), FM(T.m[2], p.z)), T.m[3]);
    o.y = FA(FA(FA(FM(T.m[4], p.x), FM(T.m[5], p.y)), FM(T.m[6], p.z)), T.m[7]);
    o.z = FA(FA(FA(FM(T.m[8], p.x), FM(T.m[9], p.y)), FM(T.m[10], p.z)), T.m[11]);
    o.w = p.w;
    return o;
}

__device__ __forceinline__ bool part_pred(const PartPred& P, float4 p) {
    if (P.kind == PART_RADIUS) {
        // double dist_square = pow(pt.x - x_criterion, 2) + pow(pt.y - y_criterion, 2); dist_square < max_dist_square
        const double dx = __dsub_rn((double)p.x, P.x), dy = __dsub_rn((double)p.y, P.y);
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        return d2 < P.limit;
    }
    if (P.kind == PART_NOT_NEAR) {
        // mapgen's vehicle-body cut (src/mapgen/mapgen.hpp:218-229): dist_square = pow(pt.x, 2) + pow(pt.y, 2) in double;
        // points with dist_square < max_dist_square are dropped, everything else (NaN included) is kept
        const double xd = (double)p.x, yd = (double)p.y;
        const double d2 = __dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd));
        return !(d2 < P.limit);
    }
    // set_submap: fabs(x - pt.x) < submap_size && fabs(y - pt.y) < submap_size
    const double dx = fabs(__dsub_rn(P.x, (double)p.x)), dy = fabs(__dsub_rn(P.y, (double)p.y));
    return (dx < P.limit) && (dy < P.limit);
}

constexpr int PART_THREADS = 256;
constexpr int PART_CHUNK   = 4096;

__global__ void __launch_bounds__(PART_THREADS)
k_part_count(PartPred P, const float4* __restrict__ in, uint32_t n, uint32_t* __restrict__ chunk_cnt) {
    __shared__ uint32_t s_w[PART_THREADS / 32];
    const uint32_t b0 = blockIdx.x * PART_CHUNK;
    uint32_t c = 0;
    for (uint32_t i = b0 + threadIdx.x; i < min(n, b0 + PART_CHUNK); i += PART_THREADS) c += part_pred(P, in[i]) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < PART_THREADS / 32; ++w) t += s_w[w];
        chunk_cnt[blockIdx.x] = t;
    }
}

// exclusive scan of up to a few hundred thousand counters by one CTA; writes total to *total
__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t s_part[34];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg = (n + 1023u) / 1024u;
    const uint32_t b0 = min(n, tid * seg), b1 = min(n, b0 + seg);
    uint32_t sum = 0;
    for (uint32_t i = b0; i < b1; ++i) sum += v[i];
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_part[lane];
        uint32_t wi = w;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, wi, o); if (lane >= o) wi += t; }
        s_part[lane] = wi - w;
        if (lane == 31) s_part[33] = wi;
    }
    __syncthreads();
    uint32_t run = s_part[warp] + incl - sum;
    for (uint32_t i = b0; i < b1; ++i) { const uint32_t t = v[i]; v[i] = run; run += t; }
    if (tid == 0 && total) *total = s_part[33];
}

// stable scatter: selected points (optionally transformed by T_sel) to out_sel in order, the rest to out_rest in order
__global__ void __launch_bounds__(PART_THREADS)
k_part_scatter(PartPred P, Mat4 T_sel, int transform_sel, const float4* __restrict__ in, uint32_t n,
               const uint32_t* __restrict__ chunk_off, float4* __restrict__ out_sel, float4* __restrict__ out_rest) {
    __shared__ uint32_t s_w[PART_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b0 = blockIdx.x * PART_CHUNK, b1 = min(n, b0 + PART_CHUNK);
    uint32_t sel_base = chunk_off[blockIdx.x];
    uint32_t rest_base = b0 - sel_base;
    for (uint32_t r0 = b0; r0 < b1; r0 += PART_THREADS) {
        const uint32_t i = r0 + tid;
        const bool ok = i < b1;
        float4 p = ok ? in[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool s = ok && part_pred(P, p);
        const unsigned bal = __ballot_sync(FULL_MASK, s);
        if (lane == 0) s_w[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, round = 0;
#pragma unroll
        for (int w = 0; w < PART_THREADS / 32; ++w) { const uint32_t c = s_w[w]; before += (w < warp) ? c : 0u; round += c; }
        const uint32_t rank_sel = before + __popc(bal & ((1u << lane) - 1u));
        if (ok) {
            if (s) out_sel[sel_base + rank_sel] = transform_sel ? affine(T_sel, p) : p;
            else   out_rest[rest_base + (i - r0) - rank_sel] = p;
        }
        const uint32_t valid = min((uint32_t)PART_THREADS, b1 - r0);
        sel_base += round;
        rest_base += valid - round;
        __syncthreads();
    }
}

cudaError_t launch_partition(cudaStream_t st, const PartPred& P, const Mat4& T_sel, bool transform_sel, const float4* in, uint32_t n,
                             uint32_t* chunk_tmp, uint32_t* d_total_sel, float4* out_sel, float4* out_rest) {
    if (n == 0) return cudaMemsetAsync(d_total_sel, 0, sizeof(uint32_t), st);
    const uint32_t chunks = (n + PART_CHUNK - 1) / PART_CHUNK;
    k_part_count<<<chunks, PART_THREADS, 0, st>>>(P, in, n, chunk_tmp);
    k_scan_u32<<<1, 1024, 0, st>>>(chunk_tmp, chunks, d_total_sel);
    k_part_scatter<<<chunks, PART_THREADS, 0, st>>>(P, T_sel, transform_sel ? 1 : 0, in, n, chunk_tmp, out_sel, out_rest);
    return cudaGetLastError();
}
size_t partition_tmp_words(uint32_t n) { return (size_t)(n + PART_CHUNK - 1) / PART_CHUNK + 1; }

__global__ void k_affine_copy(Mat4 T, int do_transform, const float4* in, float4* out, uint32_t n) {   // in may alias out
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = do_transform ? affine(T, in[i]) : in[i];
}
cudaError_t launch_affine_copy(cudaStream_t st, const Mat4& T, bool do_transform, const float4* in, float4* out, uint32_t n) {
    if (n == 0) return cudaSuccess;
    k_affine_copy<<<(n + 255) / 256, 256, 0, st>>>(T, do_transform ? 1 : 0, in, out, n);
    return cudaGetLastError();
}

// ============================================================================================
// U3  voxelize_preserving_labels
// ============================================================================================
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

__global__ void k_vox_init(VoxGrid* g) {
    if (threadIdx.x < 3) { g->mn[threadIdx.x] = 0xFFFFFFFFu; g->mx[threadIdx.x] = 0u; }
    if (threadIdx.x == 0) { g->n_vox = 0; g->overflow = 0; }
}
// getMinMax3D
__global__ void __launch_bounds__(256) k_vox_minmax(const float4* __restrict__ in, uint32_t n, VoxGrid* __restrict__ g) {
    uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float4 p = in[i];
        // fold -0.0 onto +0.0 so that the ordered encoding agrees with std::min / std::max on floats
        const uint32_t a = f2ord(p.x + 0.0f), b = f2ord(p.y + 0.0f), c = f2ord(p.z + 0.0f);
        mn[0] = min(mn[0], a); mx[0] = max(mx[0], a);
        mn[1] = min(mn[1], b); mx[1] = max(mx[1], b);
        mn[2] = min(mn[2], c); mx[2] = max(mx[2], c);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t a = __reduce_min_sync(FULL_MASK, mn[k]), b = __reduce_max_sync(FULL_MASK, mx[k]);
        if ((threadIdx.x & 31) == 0) { atomicMin(&g->mn[k], a); atomicMax(&g->mx[k], b); }
    }
}
// pcl::VoxelGrid::applyFilter set-up: inverse leaf, min_b, div_b, multipliers, overflow check
__global__ void k_vox_setup(float leaf, uint32_t n, VoxGrid* g) {
    if (threadIdx.x != 0) return;
    const float inv = FD(1.0f, leaf);
    g->leaf = leaf; g->inv = inv;
    if (n == 0) { g->div[0] = g->div[1] = g->div[2] = 1; g->min_b[0] = g->min_b[1] = g->min_b[2] = 0; return; }
    float mnf[3], mxf[3];
    for (int k = 0; k < 3; ++k) { mnf[k] = ord2f(g->mn[k]); mxf[k] = ord2f(g->mx[k]); }
    const long long dx = (long long)FM(FS(mxf[0], mnf[0]), inv) + 1;
    const long long dy = (long long)FM(FS(mxf[1], mnf[1]), inv) + 1;
    const long long dz = (long long)FM(FS(mxf[2], mnf[2]), inv) + 1;
    g->overflow = (dx * dy * dz) > 2147483647LL ? 1 : 0;
    for (int k = 0; k < 3; ++k) {
        g->min_b[k] = (int)floorf(FM(mnf[k], inv));
        const int max_b = (int)floorf(FM(mxf[k], inv));
        g->div[k] = max_b - g->min_b[k] + 1;
    }
}
__global__ void k_vox_keys(const float4* __restrict__ in, uint32_t n, const VoxGrid* __restrict__ g, uint32_t* __restrict__ key, uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    uint32_t k;
    if (g->overflow) {
        k = i;      // "Leaf size is too small": output = input, one point per voxel in cloud order
    } else {
        const float inv = g->inv;
        const int ijk0 = (int)FS(floorf(FM(p.x, inv)), (float)g->min_b[0]);
        const int ijk1 = (int)FS(floorf(FM(p.y, inv)), (float)g->min_b[1]);
        const int ijk2 = (int)FS(floorf(FM(p.z, inv)), (float)g->min_b[2]);
        k = (uint32_t)(ijk0 + ijk1 * g->div[0] + ijk2 * g->div[0] * g->div[1]);
    }
    key[i] = k; idx[i] = i;
}

// ---- global stable LSD radix sort of (key, idx) pairs: 8-bit digits, one warp per 4096-element segment -------------
constexpr int RS_SEG = 1024;   // one warp per segment: small segments = many warps in flight (the walk is latency-bound)
__global__ void __launch_bounds__(32) k_rs_hist(const uint32_t* __restrict__ key, uint32_t n, int shift, uint32_t nseg, uint32_t* __restrict__ cnt /*[256][nseg]*/) {
    __shared__ uint32_t s_c[256];
    const int lane = threadIdx.x;
    const uint32_t seg = blockIdx.x, b0 = seg * RS_SEG, b1 = min(n, b0 + RS_SEG);
    for (int i = lane; i < 256; i += 32) s_c[i] = 0u;
    __syncwarp();
    for (uint32_t e0 = b0; e0 < b1; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool valid = e < b1;
        const unsigned vm = __ballot_sync(FULL_MASK, valid);
        if (valid) {
            const uint32_t d = (key[e] >> shift) & 255u;
            const unsigned peers = __match_any_sync(vm, d);
            if (lane == __ffs(peers) - 1) s_c[d] += __popc(peers);
        }
        __syncwarp();
    }
    for (int i = lane; i < 256; i += 32) cnt[(size_t)i * nseg + seg] = s_c[i];
}
__global__ void __launch_bounds__(32) k_rs_scatter(const uint32_t* __restrict__ key, const uint32_t* __restrict__ idx, uint32_t n, int shift,
                                                   uint32_t nseg, const uint32_t* __restrict__ off /*[256][nseg] scanned*/,
                                                   uint32_t* __restrict__ key_out, uint32_t* __restrict__ idx_out) {
    __shared__ uint32_t s_o[256];
    const int lane = threadIdx.x;
    const uint32_t seg = blockIdx.x, b0 = seg * RS_SEG, b1 = min(n, b0 + RS_SEG);
    for (int i = lane; i < 256; i += 32) s_o[i] = off[(size_t)i * nseg + seg];
    __syncwarp();
    for (uint32_t e0 = b0; e0 < b1; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool valid = e < b1;
        const unsigned vm = __ballot_sync(FULL_MASK, valid);
        if (valid) {
            const uint32_t k = key[e], v = idx[e];
            const uint32_t d = (k >> shift) & 255u;
            const unsigned peers = __match_any_sync(vm, d);
            const uint32_t base = s_o[d];
            __syncwarp(vm);
            if (lane == __ffs(peers) - 1) s_o[d] = base + __popc(peers);
            const uint32_t o = base + __popc(peers & ((1u << lane) - 1u));
            key_out[o] = k; idx_out[o] = v;
        }
        __syncwarp();
    }
}

// heads of equal-key runs in the sorted order -> voxel_start[] (ordered), n_vox
__global__ void __launch_bounds__(PART_THREADS) k_vox_head_count(const uint32_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ chunk_cnt) {
    __shared__ uint32_t s_w[PART_THREADS / 32];
    const uint32_t b0 = blockIdx.x * PART_CHUNK;
    uint32_t c = 0;
    for (uint32_t i = b0 + threadIdx.x; i < min(n, b0 + PART_CHUNK); i += PART_THREADS) c += (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < PART_THREADS / 32; ++w) t += s_w[w]; chunk_cnt[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(PART_THREADS) k_vox_head_scatter(const uint32_t* __restrict__ key, uint32_t n, const uint32_t* __restrict__ chunk_off,
                                                                   uint32_t* __restrict__ vox_start, uint32_t* __restrict__ vox_key) {
    __shared__ uint32_t s_w[PART_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b0 = blockIdx.x * PART_CHUNK, b1 = min(n, b0 + PART_CHUNK);
    uint32_t base = chunk_off[blockIdx.x];
    for (uint32_t r0 = b0; r0 < b1; r0 += PART_THREADS) {
        const uint32_t i = r0 + tid;
        const bool hd = (i < b1) && (i == 0 || key[i] != key[i - 1]);
        const unsigned bal = __ballot_sync(FULL_MASK, hd);
        if (lane == 0) s_w[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, round = 0;
#pragma unroll
        for (int w = 0; w < PART_THREADS / 32; ++w) { const uint32_t c = s_w[w]; before += (w < warp) ? c : 0u; round += c; }
        if (hd) { const uint32_t v = base + before + __popc(bal & ((1u << lane) - 1u)); vox_start[v] = i; vox_key[v] = key[i]; }
        base += round;
        __syncthreads();
    }
}
__global__ void k_vox_finish_heads(VoxGrid* g, const uint32_t* __restrict__ total, uint32_t n, uint32_t* __restrict__ vox_start) {
    if (threadIdx.x == 0) { g->n_vox = *total; vox_start[*total] = n; }
}

// pcl::CentroidPoint: float sums over the voxel's members in cloud order, divided by float(count)
__global__ void k_vox_centroid(const float4* __restrict__ in, const uint32_t* __restrict__ sidx, const uint32_t* __restrict__ vox_start,
                               const VoxGrid* __restrict__ g, float4* __restrict__ out) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g->n_vox) return;
    const uint32_t a = vox_start[v], e = vox_start[v + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (uint32_t k = a; k < e; ++k) {
        const float4 p = in[sidx[k]];
        sx = FA(sx, p.x); sy = FA(sy, p.y); sz = FA(sz, p.z); si = FA(si, p.w);
    }
    const float cn = (float)(e - a);
    out[v] = make_float4(FD(sx, cn), FD(sy, cn), FD(sz, cn), FD(si, cn));
}

// exact 1-NN of every centroid into the source cloud through the voxel grid itself (cells = voxels): grow the
// Chebyshev shell until nothing unseen can be closer; ties go to the lowest cloud index.  Mirrors oracle/ line by line.
__global__ void k_vox_label(const float4* __restrict__ in, const uint32_t* __restrict__ sidx, const uint32_t* __restrict__ vox_start,
                            const uint32_t* __restrict__ vox_key, const VoxGrid* __restrict__ g, float4* __restrict__ out) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nv = g->n_vox;
    if (v >= nv) return;
    float4 c = out[v];
    if (g->overflow) { return; }       // every point is its own voxel: the nearest point is itself (or an identical earlier one)
    const float inv = g->inv, leaf = g->leaf;
    const int d0 = g->div[0], d1 = g->div[1], d2 = g->div[2];
    const int ci = (int)FS(floorf(FM(c.x, inv)), (float)g->min_b[0]);
    const int cj = (int)FS(floorf(FM(c.y, inv)), (float)g->min_b[1]);
    const int ck = (int)FS(floorf(FM(c.z, inv)), (float)g->min_b[2]);
    float best_d = __int_as_float(0x7f800000);
    uint32_t best_i = 0xFFFFFFFFu;
    int rad = 0;
    while (true) {
        // scan the shell of Chebyshev radius `rad`
        for (int a = -rad; a <= rad; ++a) {
            const int ii = ci + a;
            if (ii < 0 || ii >= d0) continue;
            for (int b = -rad; b <= rad; ++b) {
                const int jj = cj + b;
                if (jj < 0 || jj >= d1) continue;
                for (int cc = -rad; cc <= rad; ++cc) {
                    if (max(abs(a), max(abs(b), abs(cc))) != rad) continue;
                    const int kk = ck + cc;
                    if (kk < 0 || kk >= d2) continue;
                    const uint32_t key = (uint32_t)(ii + jj * d0 + kk * d0 * d1);
                    uint32_t lo = 0, hi = nv;                 // first voxel with vox_key >= key
                    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (vox_key[mid] < key) lo = mid + 1; else hi = mid; }
                    if (lo >= nv || vox_key[lo] != key) continue;
                    for (uint32_t k = vox_start[lo]; k < vox_start[lo + 1]; ++k) {
                        const uint32_t i = sidx[k];
                        const float4 p = in[i];
                        const float dx = FS(c.x, p.x), dy = FS(c.y, p.y), dz = FS(c.z, p.z);
                        const float d = FA(FA(FM(dx, dx), FM(dy, dy)), FM(dz, dz));
                        if (d < best_d || (d == best_d && i < best_i)) { best_d = d; best_i = i; }
                    }
                }
            }
        }
        const float reach = FM(FM((float)rad, leaf), 0.9999f);
        if (best_i != 0xFFFFFFFFu && best_d < FM(reach, reach)) break;
        ++rad;
        if (rad > 64 && best_i != 0xFFFFFFFFu) break;
        if (rad > 4096) break;
    }
    if (best_i != 0xFFFFFFFFu) c.w = in[best_i].w;
    out[v] = c;
}

size_t voxelize_tmp_bytes(uint32_t n) {
    const size_t nseg = (n + RS_SEG - 1) / RS_SEG + 1;
    // key/idx ping-pong (4 arrays), radix counters, partition chunk counters, voxel starts/keys
    return sizeof(uint32_t) * ((size_t)4 * n + 256 * nseg + partition_tmp_words(n) + 2 * ((size_t)n + 2) + 64);
}

cudaError_t launch_voxelize(cudaStream_t st, const float4* in, uint32_t n, float leaf, VoxGrid* grid, void* tmp, float4* out, uint32_t* d_n_out) {
    k_vox_init<<<1, 32, 0, st>>>(grid);
    if (n == 0) return cudaMemsetAsync(d_n_out, 0, sizeof(uint32_t), st);
    uint32_t* w = reinterpret_cast<uint32_t*>(tmp);
    const uint32_t nseg = (n + RS_SEG - 1) / RS_SEG;
    uint32_t* key_a = w;            uint32_t* idx_a = key_a + n;
    uint32_t* key_b = idx_a + n;    uint32_t* idx_b = key_b + n;
    uint32_t* cnt   = idx_b + n;                              // 256 * nseg
    uint32_t* chunk = cnt + (size_t)256 * (nseg + 1);         // partition_tmp_words(n)
    uint32_t* vstart = chunk + partition_tmp_words(n);        // n + 2
    uint32_t* vkey   = vstart + (n + 2);                      // n + 2
    const uint32_t mm_want = (n + 255) / 256;
    const int mm_blocks = (int)(mm_want < 1184u ? mm_want : 1184u);
    k_vox_minmax<<<mm_blocks, 256, 0, st>>>(in, n, grid);
    k_vox_setup<<<1, 32, 0, st>>>(leaf, n, grid);
    k_vox_keys<<<(n + 255) / 256, 256, 0, st>>>(in, n, grid, key_a, idx_a);
    for (int pass = 0; pass < 4; ++pass) {
        const uint32_t* ki = (pass & 1) ? key_b : key_a;  const uint32_t* ii = (pass & 1) ? idx_b : idx_a;
        uint32_t* ko = (pass & 1) ? key_a : key_b;        uint32_t* io = (pass & 1) ? idx_a : idx_b;
        k_rs_hist<<<nseg, 32, 0, st>>>(ki, n, pass * 8, nseg, cnt);
        k_scan_u32<<<1, 1024, 0, st>>>(cnt, 256u * nseg, nullptr);
        k_rs_scatter<<<nseg, 32, 0, st>>>(ki, ii, n, pass * 8, nseg, cnt, ko, io);
    }
    // after 4 passes the sorted pairs are back in (key_a, idx_a)
    const uint32_t chunks = (n + PART_CHUNK - 1) / PART_CHUNK;
    k_vox_head_count<<<chunks, PART_THREADS, 0, st>>>(key_a, n, chunk);
    k_scan_u32<<<1, 1024, 0, st>>>(chunk, chunks, d_n_out);
    k_vox_head_scatter<<<chunks, PART_THREADS, 0, st>>>(key_a, n, chunk, vstart, vkey);
    k_vox_finish_heads<<<1, 32, 0, st>>>(grid, d_n_out, n, vstart);
    // n_vox is only known on the device: launch for the upper bound n and let threads beyond n_vox exit
    k_vox_centroid<<<(n + 127) / 128, 128, 0, st>>>(in, idx_a, vstart, grid, out);
    k_vox_label<<<(n + 127) / 128, 128, 0, st>>>(in, idx_a, vstart, vkey, grid, out);
    return cudaGetLastError();
}
int voxelize_num_launches() { return 3 + 12 + 4 + 2 + 1; }

}  // namespace erasor
