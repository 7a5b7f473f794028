// prep.cu -- F4: the reference's PCpreprocessing (src/CommonFunc.cpp:423-452) on the device:
// pcl::VoxelGrid with a cubic leaf and the first pass of pcl::StatisticalOutlierRemoval (mean distance of every
// point to its k nearest other points).  [PCL 1.8.1 filters/voxel_grid.hpp applyFilter,
// statistical_outlier_removal.hpp applyFilterIndices; tests/test_gpu_parity.py compares both bit for bit with the
// CPU restatement of the same two filters.]
//
//   voxel_key_kernel -> cub radix sort (stable: points of a voxel stay in input order, which fixes the float
//   summation order) -> voxel_mark_kernel -> cub inclusive scan -> voxel_centroid_kernel (one thread per voxel run,
//   sequential float sums, IEEE division) ; output in ascending voxel index, like PCL.
//   knn_mean_dist_kernel: exact k-NN on the cell-sorted grid of nn_search.cuh (level 0 only): the k smallest float
//   distances are kept sorted in registers; a cube of cells around the query is scanned and the search ends when the
//   k-th distance fits inside the scanned cube, else the cube grows to that distance.
//
// Both are HBM/L2-latency bound gather kernels (algorithmic bytes: voxel grid 12 B read + 12 B/voxel written + 12 B
// key/index traffic per sort pass; k-NN 16 B per candidate visited, ~60 candidates per point on a sampled surface).
#include <cub/cub.cuh>

#include "common.cuh"
#include "nn_search.cuh"

namespace pwicp {

__global__ void voxel_key_kernel(const float* __restrict__ xyz, int n, float inv, long long mbx, long long mby, long long mbz,
                                 long long div0, long long div01, unsigned long long* __restrict__ keys,
                                 uint32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // ijk = floor(p * inverse_leaf) - min_b (voxel_grid.hpp), products rounded to float first
    const long long ix = (long long)floorf(__fmul_rn(xyz[3 * (size_t)i], inv)) - mbx;
    const long long iy = (long long)floorf(__fmul_rn(xyz[3 * (size_t)i + 1], inv)) - mby;
    const long long iz = (long long)floorf(__fmul_rn(xyz[3 * (size_t)i + 2], inv)) - mbz;
    keys[i] = (unsigned long long)(ix + iy * div0 + iz * div01);
    idx[i] = (uint32_t)i;
}

__global__ void voxel_mark_kernel(const unsigned long long* __restrict__ keys, int n, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per sorted position that starts a voxel run
__global__ void voxel_centroid_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx,
                                      const int* __restrict__ flags, const int* __restrict__ incl, const float* __restrict__ xyz,
                                      int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const unsigned long long k = keys[i];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int j = i;
    for (; j < n && keys[j] == k; ++j) {
        const float* p = xyz + 3 * (size_t)idx[j];
        sx = __fadd_rn(sx, p[0]); sy = __fadd_rn(sy, p[1]); sz = __fadd_rn(sz, p[2]);
    }
    const float cnt = (float)(j - i);
    float* o = out + 3 * (size_t)(incl[i] - 1);
    o[0] = __fdiv_rn(sx, cnt); o[1] = __fdiv_rn(sy, cnt); o[2] = __fdiv_rn(sz, cnt);
}

// result: number of voxels; the voxel centroids (packed xyz) in `out_dev` (capacity n points)
int voxel_grid_dev(Ctx* ctx, const float* xyz_dev, int n, float leaf, float* out_dev, int* n_out) {
    float mn[3], mx[3];
    PW_TRY(bbox_packed_dev(ctx, xyz_dev, (size_t)n, mn, mx));
    const float inv = 1.0f / leaf;
    long long minb[3], div[3];
    for (int c = 0; c < 3; ++c) {
        minb[c] = (long long)floorf(mn[c] * inv);
        div[c] = (long long)floorf(mx[c] * inv) - minb[c] + 1;
    }
    // PCL refuses grids whose index overflows int32 ("Leaf size is too small"); 64-bit keys carry them, up to 2^62
    const long double cells = (long double)div[0] * (long double)div[1] * (long double)div[2];
    if (cells > 4.0e18L) { set_error(ctx, "voxel_grid: leaf size too small for the extent of the cloud"); return PWICP_ERR_ARG; }
    int bits = 1;
    while (bits < 64 && (long double)(1ULL << bits) < cells) ++bits;
    PW_TRY(ctx->keys.reserve(ctx, (size_t)n * 8));
    PW_TRY(ctx->keys2.reserve(ctx, (size_t)n * 8));
    PW_TRY(ctx->vals.reserve(ctx, (size_t)n * 4));
    PW_TRY(ctx->vals2.reserve(ctx, (size_t)n * 4));
    PW_TRY(ctx->flags.reserve(ctx, (size_t)n * 4));
    PW_TRY(ctx->pos.reserve(ctx, (size_t)n * 4));
    const int blocks = (n + 255) / 256;
    auto* k1 = ctx->keys.as<unsigned long long>();
    auto* k2 = ctx->keys2.as<unsigned long long>();
    voxel_key_kernel<<<blocks, 256, 0, ctx->stream>>>(xyz_dev, n, inv, minb[0], minb[1], minb[2], div[0], div[0] * div[1], k1,
                                                     ctx->vals.as<uint32_t>());
    size_t tmp = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, k1, k2, ctx->vals.as<uint32_t>(), ctx->vals2.as<uint32_t>(), n, 0, bits, ctx->stream);
    cub::DeviceScan::InclusiveSum(nullptr, tmp2, ctx->flags.as<int>(), ctx->pos.as<int>(), n, ctx->stream);
    PW_TRY(ctx->cub_tmp.reserve(ctx, std::max(tmp, tmp2)));
    size_t cap = ctx->cub_tmp.cap;
    PW_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, cap, k1, k2, ctx->vals.as<uint32_t>(), ctx->vals2.as<uint32_t>(), n, 0, bits,
                                            ctx->stream));
    voxel_mark_kernel<<<blocks, 256, 0, ctx->stream>>>(k2, n, ctx->flags.as<int>());
    cap = ctx->cub_tmp.cap;
    PW_CUDA(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, cap, ctx->flags.as<int>(), ctx->pos.as<int>(), n, ctx->stream));
    voxel_centroid_kernel<<<blocks, 256, 0, ctx->stream>>>(k2, ctx->vals2.as<uint32_t>(), ctx->flags.as<int>(), ctx->pos.as<int>(), xyz_dev, n,
                                                          out_dev);
    ctx->launches += 5;
    PW_CUDA(cudaGetLastError());
    int m = 0;
    PW_CUDA(cudaMemcpyAsync(&m, ctx->pos.as<int>() + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_out = m;
    return PWICP_OK;
}

// ---- k nearest other points: mean of the k distances -----------------------------------------------------------------
// One thread per cell-sorted position.  best[0..K) ascending float squared distances (registers: K is a template
// parameter, the insertion is fully unrolled).  The cube of cells covering [p - R, p + R] is scanned (rows along x are
// contiguous ranges of the sorted array); when K candidates are known and sqrt(best[K-1]) <= R the K nearest all lie in
// the cube (closed ball inside the scanned box, with the cell-assignment margin of nn_search.cuh) and the search ends,
// else R grows (to the K-th distance when known, else by doubling) and the cube is rescanned.
template <int K>
__global__ void __launch_bounds__(128)
knn_mean_dist_kernel(GridDev g, int k_runtime, float r0, float* __restrict__ out) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const GridLevel& L = g.lv[0];
    const float4 p = __ldg(L.pts + pos);
    const int self = __float_as_int(p.w);
    const float fx = (p.x - g.ox) * L.inv_h, fy = (p.y - g.oy) * L.inv_h, fz = (p.z - g.oz) * L.inv_h;
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f, mz = 0.01f + fabsf(fz) * 4e-6f;
    const float inf = __int_as_float(0x7f800000);
    float best[K];
    float R = r0;                                          // in units of cells
    for (int pass = 0; pass < 64; ++pass) {
#pragma unroll
        for (int j = 0; j < K; ++j) best[j] = inf;
        const int lx = min(max((int)floorf(fx - R - mx), 0), L.dx - 1), hx = min(max((int)floorf(fx + R + mx), 0), L.dx - 1);
        const int ly = min(max((int)floorf(fy - R - my), 0), L.dy - 1), hy = min(max((int)floorf(fy + R + my), 0), L.dy - 1);
        const int lz = min(max((int)floorf(fz - R - mz), 0), L.dz - 1), hz = min(max((int)floorf(fz + R + mz), 0), L.dz - 1);
        for (int kz = lz; kz <= hz; ++kz)
            for (int ky = ly; ky <= hy; ++ky) {
                const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
                const uint32_t b = __ldg(L.cell_start + row + lx), e = __ldg(L.cell_start + row + hx + 1);
                for (uint32_t i = b; i < e; ++i) {
                    const float4 q = __ldg(L.pts + i);
                    if (__float_as_int(q.w) == self) continue;
                    float d = l2_simple(p.x, p.y, p.z, q.x, q.y, q.z);
                    if (d < best[K - 1]) {
                        // sorted insertion: the new value replaces the largest and sinks to its place (one bubble pass,
                        // static indices -> registers; equal values do not swap)
                        best[K - 1] = d;
#pragma unroll
                        for (int j = K - 1; j > 0; --j)
                            if (best[j] < best[j - 1]) { const float t = best[j]; best[j] = best[j - 1]; best[j - 1] = t; }
                    }
                }
            }
        const bool whole = lx == 0 && ly == 0 && lz == 0 && hx == L.dx - 1 && hy == L.dy - 1 && hz == L.dz - 1;
        float kth = inf;                                   // the k-th smallest (k <= K), selected without dynamic indexing
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j == k_runtime - 1) kth = best[j];
        if (whole || (kth < inf && sqrtf(kth) * L.inv_h * 1.00001f <= R)) break;
        R = (kth < inf) ? sqrtf(kth) * L.inv_h * 1.0001f : R * 2.0f;
    }
    // dist_sum (double) of sqrt(d2) in ascending order, float(dist_sum / k)   (statistical_outlier_removal.hpp)
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if (j < k_runtime) s += (double)__fsqrt_rn(best[j]);
    out[self] = (float)(s / (double)k_runtime);
}

int knn_mean_dist_dev(Ctx* ctx, const GridDev& g, int k, float* out_dev) {
    if (k < 1 || k > 32 || g.n <= k) { set_error(ctx, "knn_mean_dist: need 1 <= k <= 32 and more than k points"); return PWICP_ERR_ARG; }
    // first cube: one cell around the point in every direction (3 x 3 x 3 cells; on a sampled surface a cell holds ~10 points,
    // so the 14 nearest usually lie inside; otherwise the second pass scans the cube of the k-th distance found)
    const float r0 = 1.0f;
    const int blocks = (g.n + 127) / 128;
    // K is a compile-time bound of the register-resident list; k <= K entries are summed
    if (k <= 8) knn_mean_dist_kernel<8><<<blocks, 128, 0, ctx->stream>>>(g, k, r0, out_dev);
    else if (k <= 16) knn_mean_dist_kernel<16><<<blocks, 128, 0, ctx->stream>>>(g, k, r0, out_dev);
    else knn_mean_dist_kernel<32><<<blocks, 128, 0, ctx->stream>>>(g, k, r0, out_dev);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

// ---- k nearest neighbours WITH indices + PCA normal (segmentation front end, F4) ---------------------------------------
// Replaces the per-point loop of the reference's PatchGenerationAndRefinement (src/Segmentation.cpp:28-46:
// kdtree.FindKNearestNeighbors(points[i], k, &neighbors) and PCAEstimateNormal over them; 0.9 s of its 1.6 s per 170k-point
// cloud).  One thread per cell-sorted point; the k best (squared distance in the codelibrary's double metric,
// t += double(a - b) * (a - b) over x, y, z; ties by index; the point itself first) live in a sorted list in local memory.
// The cube of cells covering [p - R, p + R] is scanned; when the k-th distance fits inside the cube the list is complete,
// else R grows to it and the cube is rescanned (as knn_mean_dist_kernel).  The normal is the eigenvector of the smallest
// eigenvalue of the neighbours' covariance in closed form (codelibrary/geometry/point_cloud/pca_estimate_normals.h:47-117,
// unit weights, sums in neighbour order); its orientation is not defined (the reference says so).
constexpr int kKnnMax = 64;
__global__ void __launch_bounds__(128)
knn_normals_kernel(GridDev g, int k, float r0, int* __restrict__ neighbors, double* __restrict__ normals) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const GridLevel& L = g.lv[0];
    const float4 p = __ldg(L.pts + pos);
    const int self = __float_as_int(p.w);
    const float fx = (p.x - g.ox) * L.inv_h, fy = (p.y - g.oy) * L.inv_h, fz = (p.z - g.oz) * L.inv_h;
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f, mz = 0.01f + fabsf(fz) * 4e-6f;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double bd[kKnnMax];
    int bi[kKnnMax], bp[kKnnMax];                          // original index (the order key), sorted position (the address)
    float R = r0;                                          // in units of cells
    for (int pass = 0; pass < 64; ++pass) {
        for (int j = 0; j < k; ++j) { bd[j] = inf; bi[j] = 0x7fffffff; bp[j] = pos; }
        const int lx = min(max((int)floorf(fx - R - mx), 0), L.dx - 1), hx = min(max((int)floorf(fx + R + mx), 0), L.dx - 1);
        const int ly = min(max((int)floorf(fy - R - my), 0), L.dy - 1), hy = min(max((int)floorf(fy + R + my), 0), L.dy - 1);
        const int lz = min(max((int)floorf(fz - R - mz), 0), L.dz - 1), hz = min(max((int)floorf(fz + R + mz), 0), L.dz - 1);
        for (int kz = lz; kz <= hz; ++kz)
            for (int ky = ly; ky <= hy; ++ky) {
                const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
                const uint32_t b = __ldg(L.cell_start + row + lx), e = __ldg(L.cell_start + row + hx + 1);
                for (uint32_t i = b; i < e; ++i) {
                    const float4 q = __ldg(L.pts + i);
                    const double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
                    double d = 0.0;
                    d += dx * dx; d += dy * dy; d += dz * dz;
                    const int id = __float_as_int(q.w);
                    if (d < bd[k - 1] || (d == bd[k - 1] && id < bi[k - 1])) {
                        int j = k - 1;
                        while (j > 0 && (bd[j - 1] > d || (bd[j - 1] == d && bi[j - 1] > id))) {
                            bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; bp[j] = bp[j - 1]; --j;
                        }
                        bd[j] = d; bi[j] = id; bp[j] = (int)i;
                    }
                }
            }
        const bool whole = lx == 0 && ly == 0 && lz == 0 && hx == L.dx - 1 && hy == L.dy - 1 && hz == L.dz - 1;
        const double kth = bd[k - 1];
        if (whole || (kth < inf && (float)sqrt(kth) * L.inv_h * 1.00001f <= R)) break;
        R = (kth < inf) ? (float)sqrt(kth) * L.inv_h * 1.0001f : R * 2.0f;
    }
    if (neighbors) for (int j = 0; j < k; ++j) neighbors[(size_t)self * k + j] = bi[j];
    if (!normals) return;
    double cx = 0, cy = 0, cz = 0, sum = 0;                                   // Centroid3D, center_3d.h:82-108
    for (int j = 0; j < k; ++j) {
        const float4 q = __ldg(L.pts + bp[j]);
        const double w = 1.0;
        cx += w * q.x; cy += w * q.y; cz += w * q.z; sum += w;
    }
    sum = 1.0 / sum; cx *= sum; cy *= sum; cz *= sum;
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, wsum = 0;
    for (int j = 0; j < k; ++j) {
        const float4 q = __ldg(L.pts + bp[j]);
        const double x = q.x - cx, y = q.y - cy, z = q.z - cz, w = 1.0;
        a00 += w * x * x; a01 += w * x * y; a02 += w * x * z; a11 += w * y * y; a12 += w * y * z; a22 += w * z * z;
        wsum += w;
    }
    const double t = 1.0 / wsum;
    a00 *= t; a01 *= t; a02 *= t; a11 *= t; a12 *= t; a22 *= t;
    // smallest eigenvalue of the symmetric 3x3: trigonometric solution of the characteristic cubic
    const double q3 = (a00 + a11 + a22) / 3.0;
    double pq = (a00 - q3) * (a00 - q3) + (a11 - q3) * (a11 - q3) + (a22 - q3) * (a22 - q3) + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12);
    pq = sqrt(pq / 6.0);
    const double mpq = pow(1.0 / pq, 3.0);
    const double det_b = mpq * ((a00 - q3) * ((a11 - q3) * (a22 - q3) - a12 * a12) - a01 * (a01 * (a22 - q3) - a12 * a02) +
                                a02 * (a01 * a12 - (a11 - q3) * a02));
    const double r = 0.5 * det_b;
    const double kPi = 3.14159265358979323846;
    double phi = 0.0;
    if (r <= -1.0) phi = kPi / 3.0;
    else if (r >= 1.0) phi = 0.0;
    else phi = acos(r) / 3.0;
    const double eig = q3 + 2.0 * pq * cos(phi + kPi * (2.0 / 3.0));
    double nx = a01 * a12 - a02 * (a11 - eig);
    double ny = a01 * a02 - a12 * (a00 - eig);
    double nz = (a00 - eig) * (a11 - eig) - a01 * a01;
    const double norm = sqrt(nx * nx + ny * ny + nz * nz);
    if (norm == 0.0) { nx = 0.0; ny = 0.0; nz = 1.0; }
    else { const double sc = 1.0 / norm; nx *= sc; ny *= sc; nz *= sc; }
    normals[3 * (size_t)self] = nx; normals[3 * (size_t)self + 1] = ny; normals[3 * (size_t)self + 2] = nz;
}

int knn_normals_dev(Ctx* ctx, const GridDev& g, int k, int* neighbors_dev, double* normals_dev) {
    if (k < 1 || k > kKnnMax || g.n < k) { set_error(ctx, "knn_normals: need 1 <= k <= 64 <= n"); return PWICP_ERR_ARG; }
    // first cube: on a sampled surface the k nearest lie within sqrt(k / pi) point spacings; a cell holds a few points
    const float r0 = k <= 16 ? 1.0f : 2.0f;
    knn_normals_kernel<<<(g.n + 127) / 128, 128, 0, ctx->stream>>>(g, k, r0, neighbors_dev, normals_dev);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

}  // namespace pwicp
