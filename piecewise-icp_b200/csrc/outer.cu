// outer.cu -- one outer iteration of Piecewise-ICP on the device-resident pair.
//
// Follows PwICP_singleIteration (reference src/Registration.cpp:704-972) statement by statement
// (SURVEY.md appendix A): (1) NN CT2->CT1 and BP2->CT1, (2) LoD per patch, (3) point-to-plane
// distances, (4) stable/unstable classification + order-preserving compaction, (5) inner ICP,
// (6) bounding-cube corner change, (7) DT schedule incl. the stage-1 percentile, (8) transform of
// cloud2 / CT2 / BP2 / patches, (9) VCM on the pre-update stable centroids.
#include <cub/cub.cuh>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>

#include "common.cuh"
#include "nn_search.cuh"
#include "small_algebra.cuh"

namespace pwicp {

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return (i >= 0) ? i : i ^ 0x7fffffff; }
static inline float ord2f(int i) { int j = (i >= 0) ? i : i ^ 0x7fffffff; float f; memcpy(&f, &j, 4); return f; }

// ---- (1)-(3): distances of the n2 centroids and 6*n2 boundary points --------------------------
// t < n2: centroid t; t >= n2: boundary point t - n2 (6 per patch).  src/Registration.cpp:737-812.
__global__ void __launch_bounds__(256)
classify_dist_kernel(GridDev g, const float4* __restrict__ aux, const unsigned char* __restrict__ ok,
                     const float4* __restrict__ ct2, const float4* __restrict__ bp2,
                     const float* __restrict__ bpstd2, int n2, float minLoD, float maxLoD,
                     float* __restrict__ pl, float* __restrict__ pt2pt, float* __restrict__ lod,
                     int* __restrict__ lod_minmax, const uint32_t* __restrict__ order,
                     int* __restrict__ ct_seed, int* __restrict__ bp_seed) {
    // thread u handles the u-th query of the Morton-ordered patch list (spatially compact groups);
    // t is the query's slot in the caller's order: centroid t < n2, boundary point t - n2
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = u < 7 * n2;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    int t = 0, seed = -1;
    if (active) {
        if (u < n2) { t = (int)order[u]; p = __ldg(ct2 + t); seed = ct_seed[t]; }
        else {
            const int v = u - n2, bpi = 6 * (int)order[v / 6] + v % 6;
            t = n2 + bpi; p = __ldg(bp2 + bpi); seed = bp_seed[bpi];
        }
    }
    if (!active) return;
    const Best b = nn_search_seeded(g, p.x, p.y, p.z, seed);
    if (t < n2) ct_seed[t] = b.pos; else bp_seed[t - n2] = b.pos;
    const float4 a = __ldg(aux + b.pos);
    float resDis;
    if (__ldg(ok + b.pos)) {
        const float DisDx = b.qx - p.x, DisDy = b.qy - p.y, DisDz = b.qz - p.z;
        resDis = fabsf(DisDx * a.x + DisDy * a.y + DisDz * a.z);
    } else {
        resDis = sqrtf(b.d2);
    }
    pl[t] = resDis;
    if (t < n2) {
        pt2pt[t] = sqrtf(b.d2);
        const float sigm1 = a.w, sigm2 = __ldg(bpstd2 + t);
        float LoD = (float)(1.96 * (double)sqrtf(sigm1 * sigm1 + sigm2 * sigm2));
        if (LoD > maxLoD) LoD = maxLoD;
        else if (LoD < minLoD) LoD = minLoD;
        lod[t] = LoD;
        atomicMin(lod_minmax, f2ord(LoD));
        atomicMax(lod_minmax + 1, f2ord(LoD));
    }
}

// ---- (4) classification, src/Registration.cpp:815-862 ---------------------------------------
__global__ void classify_flag_kernel(const float* __restrict__ pl, const float* __restrict__ pt2pt,
                                     const float* __restrict__ lod, const int* __restrict__ patch_off,
                                     int n2, float currDT, float DTctct, int* __restrict__ flags,
                                     unsigned long long* __restrict__ n_pts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;      // n2 tail: the remaining lanes of the warp stay converged below
    const float L = lod[i];
    const float thr = (currDT <= L) ? L : currDT;
    bool pass = true;
    const float* bp = pl + n2 + 6 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 6; ++k) if (thr < bp[k]) pass = false;
    if (thr < pl[i]) pass = false;
    const bool stable = pass && (pt2pt[i] < DTctct);
    flags[i] = stable ? 1 : 0;
    // points of the stable patches: warp-aggregated (one atomic per warp, integer -> deterministic)
    unsigned int np = stable ? (unsigned int)(patch_off[i + 1] - patch_off[i]) : 0u;
    np = __reduce_add_sync(__activemask(), np);
    if ((threadIdx.x & 31) == 0 && np) atomicAdd(n_pts, (unsigned long long)np);
}

__global__ void compact_kernel(const float4* __restrict__ ct2, const int* __restrict__ flags,
                               const int* __restrict__ pos, int n2, float4* __restrict__ out,
                               unsigned char* __restrict__ flags_u8, const int* __restrict__ ct_seed,
                               int* __restrict__ seed_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const int f = flags[i];
    if (f) { out[pos[i]] = ct2[i]; seed_out[pos[i]] = ct_seed[i]; }
    if (flags_u8) flags_u8[i] = (unsigned char)f;
}

// ---- (7) stage-1 percentile: NN distances of (flagged) points against a full-cloud grid ------
// calPercentileDistBetween2PC, src/CommonFunc.cpp:266-281.  Unflagged points get +inf so that
// they sort behind every valid distance.
__global__ void __launch_bounds__(256)
percentile_d2_kernel(GridDev g, const float* __restrict__ q, int nq, const int* __restrict__ patch_id,
                     const int* __restrict__ flags, float* __restrict__ d2, int* __restrict__ seeds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inrange = i < nq;
    const bool active = inrange && !(flags && !flags[patch_id[i]]);
    float px = 0.f, py = 0.f, pz = 0.f;
    if (active) { px = q[3 * (size_t)i]; py = q[3 * (size_t)i + 1]; pz = q[3 * (size_t)i + 2]; }
    const int seed = (active && seeds) ? seeds[i] : -1;
    float dist2 = __int_as_float(0x7f800000);
    if (active) {
        const Best b = nn_search_seeded(g, px, py, pz, seed);
        if (seeds) seeds[i] = b.pos;
        dist2 = b.d2;
    }
    if (inrange) d2[i] = dist2;
}

int percentile_dev(Ctx* ctx, const GridDev& g, const float* q, int nq, const int* patch_id,
                   const int* flags, long long n_valid, float pct, double* out, int* seeds) {
    if (nq < 1 || n_valid < 1) { set_error(ctx, "percentile: empty query set"); return PWICP_ERR_ARG; }
    PW_TRY(ctx->scratch_a.reserve(ctx, (size_t)nq * 4));
    PW_TRY(ctx->scratch_b.reserve(ctx, (size_t)nq * 4));
    float* d2 = ctx->scratch_a.as<float>();
    float* d2s = ctx->scratch_b.as<float>();
    const size_t smem = 0;
    percentile_d2_kernel<<<(nq + 255) / 256, 256, smem, ctx->stream>>>(g, q, nq, patch_id, flags, d2, seeds);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, d2, d2s, nq, 0, 32, ctx->stream);
    PW_TRY(ctx->cub_tmp.reserve(ctx, tmp));
    size_t cap = ctx->cub_tmp.cap;
    PW_CUDA(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, cap, d2, d2s, nq, 0, 32, ctx->stream));
    ctx->launches += 4;
    int leftnum = (int)((float)n_valid * pct);           // int leftnum = n * percentile (:177)
    if (leftnum >= n_valid) leftnum = (int)n_valid - 1;
    float v = 0.f;
    PW_CUDA(cudaMemcpyAsync(&v, d2s + leftnum, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = (double)sqrtf(v);                             // distArray[i] = sqrt(float) (:277)
    return PWICP_OK;
}

// ---- (8) transforms, src/Registration.cpp:942-954 -------------------------------------------
struct Mat34 { float m[12]; };

// packed xyz: each thread moves 4 points = three 16-byte vectors (coalesced 128-bit accesses)
__global__ void __launch_bounds__(256)
transform_packed_kernel(float* __restrict__ xyz, size_t n, Mat34 T) {
    const size_t nquad = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float4* v = reinterpret_cast<float4*>(xyz);
    for (size_t qd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; qd < nquad; qd += stride) {
        float4 a = v[3 * qd], b = v[3 * qd + 1], c = v[3 * qd + 2];
        float o[12];
        xform_point(T.m, a.x, a.y, a.z, o[0], o[1], o[2]);
        xform_point(T.m, a.w, b.x, b.y, o[3], o[4], o[5]);
        xform_point(T.m, b.z, b.w, c.x, o[6], o[7], o[8]);
        xform_point(T.m, c.y, c.z, c.w, o[9], o[10], o[11]);
        v[3 * qd] = make_float4(o[0], o[1], o[2], o[3]);
        v[3 * qd + 1] = make_float4(o[4], o[5], o[6], o[7]);
        v[3 * qd + 2] = make_float4(o[8], o[9], o[10], o[11]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = nquad * 4 + threadIdx.x;
        float x, y, z;
        xform_point(T.m, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], x, y, z);
        xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
    }
}

__global__ void transform_f4_kernel(float4* __restrict__ p, int n, Mat34 T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = p[i];
    float x, y, z;
    xform_point(T.m, v.x, v.y, v.z, x, y, z);
    p[i] = make_float4(x, y, z, v.w);
}

int transform_packed_dev(Ctx* ctx, float* xyz, size_t n, const float* T16) {
    if (!n) return PWICP_OK;
    Mat34 T; for (int k = 0; k < 12; ++k) T.m[k] = T16[k];
    size_t nquad = n / 4;
    int blocks = (int)std::min<size_t>(std::max<size_t>((nquad + 255) / 256, 1), (size_t)ctx->num_sms * 16);
    transform_packed_kernel<<<blocks, 256, 0, ctx->stream>>>(xyz, n, T);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

static int transform_f4_dev(Ctx* ctx, float4* p, int n, const float* T16) {
    if (!n) return PWICP_OK;
    Mat34 T; for (int k = 0; k < 12; ++k) T.m[k] = T16[k];
    transform_f4_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(p, n, T);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

// ---- (6) octree bounding cube, src/Registration.cpp:881-886 ---------------------------------
// pcl::octree::OctreePointCloud::defineBoundingBox() + getKeyBitSize() (PCL 1.8.1)
void octree_cube(const float* mn, const float* mx, double res, double* bb) {
    const float minValue512 = std::numeric_limits<float>::epsilon() * 512.0f;
    const float minValue = std::numeric_limits<float>::epsilon();
    double lo[3], hi[3];
    for (int c = 0; c < 3; ++c) { lo[c] = mn[c]; hi[c] = (float)(mx[c] + minValue512); }
    unsigned int key[3];
    for (int c = 0; c < 3; ++c) key[c] = (unsigned int)std::ceil((hi[c] - lo[c] - minValue) / res);
    unsigned int max_voxels = std::max(std::max(std::max(key[0], key[1]), key[2]), 2u);
    unsigned int depth = (unsigned int)std::ceil(std::log((double)max_voxels) / std::log(2.0) - minValue);
    depth = std::min(depth, 32u);
    const double side = (double)(1u << depth) * res;
    for (int c = 0; c < 3; ++c) {
        const double over = (side - (hi[c] - lo[c])) / 2.0;
        if (over > minValue) { lo[c] -= over; hi[c] += over; }
    }
    bb[0] = lo[0]; bb[1] = lo[1]; bb[2] = lo[2]; bb[3] = hi[0]; bb[4] = hi[1]; bb[5] = hi[2];
}

// ---- (9) VCM, src/Registration.cpp:1273-1343 ------------------------------------------------
constexpr int kVcmBlocks = 296;

__device__ __forceinline__ void vcm_row(const float4* aux, const Best& b, const float4 q, double* a, double& L) {
    const float4 nq = __ldg(aux + b.pos);
    const double Qx = q.x, Qy = q.y, Qz = q.z, Px = b.qx, Py = b.qy, Pz = b.qz;
    const double Nx = nq.x, Ny = nq.y, Nz = nq.z;
    a[0] = Nz * Qy - Ny * Qz; a[1] = Nx * Qz - Nz * Qx; a[2] = Ny * Qx - Nx * Qy;
    a[3] = Nx; a[4] = Ny; a[5] = Nz;
    L = Nx * (Px - Qx) + Ny * (Py - Qy) + Nz * (Pz - Qz);
}

// pass 0: per-block partial sums of the 21 upper-triangle A^T A entries and 6 A^T L entries
// pass 1: per-block partial sums of v^T v with v = A x - L
__global__ void __launch_bounds__(256)
vcm_kernel(GridDev g, const float4* __restrict__ aux, const float4* __restrict__ src, int n, int pass,
           const double* __restrict__ X, double* __restrict__ partials, int* __restrict__ seeds) {
    __shared__ double sm[8][27];
    double acc[27];
#pragma unroll
    for (int v = 0; v < 27; ++v) acc[v] = 0.0;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool active = i < n;
        const float4 q = active ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int seed = (active && seeds) ? seeds[i] : -1;
        if (!active) continue;
        const Best b = nn_search_seeded(g, q.x, q.y, q.z, seed);
        if (seeds) seeds[i] = b.pos;
        double a[6], L;
        vcm_row(aux, b, q, a, L);
        if (pass == 0) {
            int v = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = r; c < 6; ++c) acc[v++] += a[r] * a[c];
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[21 + r] += a[r] * L;
        } else {
            double vv = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) vv += a[c] * X[c];
            vv -= L;
            acc[0] += vv * vv;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nv = pass == 0 ? 27 : 1;
    for (int v = 0; v < nv; ++v) {
        double s = acc[v];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sm[warp][v] = s;
    }
    __syncthreads();
    if (threadIdx.x < nv) {
        double s = sm[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) s += sm[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 27 + threadIdx.x] = s;
    }
}

int vcm_dev(Ctx* ctx, const float4* src, int n, double* vcm36, int* singular, int* seeds) {
    if (n < 7) { set_error(ctx, "vcm: needs more than 6 stable patches"); return PWICP_ERR_TOO_FEW_STABLE; }
    const int blocks = std::min(kVcmBlocks, (n + 255) / 256);
    PW_TRY(ctx->scratch_c.reserve(ctx, (size_t)kVcmBlocks * 27 * 8 + 64));
    double* part = ctx->scratch_c.as<double>();
    double* Xd = part + (size_t)kVcmBlocks * 27;
    std::vector<double> hp((size_t)blocks * 27);
    const size_t smem = 0;
    vcm_kernel<<<blocks, 256, smem, ctx->stream>>>(ctx->tgt.dev, ctx->tgt_aux.as<float4>(), src, n, 0, nullptr, part, seeds);
    ctx->launches++;
    PW_CUDA(cudaMemcpyAsync(hp.data(), part, hp.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    double s27[27];
    for (int v = 0; v < 27; ++v) { double s = 0; for (int b = 0; b < blocks; ++b) s += hp[(size_t)b * 27 + v]; s27[v] = s; }
    double ATA[36], ATL[6], Q[36], X[6];
    int v = 0;
    for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { ATA[r * 6 + c] = s27[v]; ATA[c * 6 + r] = s27[v]; ++v; }
    for (int r = 0; r < 6; ++r) ATL[r] = s27[21 + r];
    const double det = inverse6(ATA, Q);
    if (singular) *singular = (std::fabs(det) < 1e-9) ? 1 : 0;      // :1324-1325 (reported only)
    for (int r = 0; r < 6; ++r) { double s = 0; for (int c = 0; c < 6; ++c) s += Q[r * 6 + c] * ATL[c]; X[r] = s; }
    PW_CUDA(cudaMemcpyAsync(Xd, X, sizeof(X), cudaMemcpyHostToDevice, ctx->stream));
    vcm_kernel<<<blocks, 256, smem, ctx->stream>>>(ctx->tgt.dev, ctx->tgt_aux.as<float4>(), src, n, 1, Xd, part, seeds);
    ctx->launches++;
    PW_CUDA(cudaMemcpyAsync(hp.data(), part, hp.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    double vtpv = 0;
    for (int b = 0; b < blocks; ++b) vtpv += hp[(size_t)b * 27];
    const double STD0 = 1 * std::sqrt(vtpv / double(n - 6));
    for (int k = 0; k < 36; ++k) vcm36[k] = STD0 * STD0 * Q[k];
    return PWICP_OK;
}

// ---- Morton order of the source patches (once per pair) ------------------------------------
__device__ __forceinline__ unsigned long long spread21_o(unsigned int v) {
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void patch_key_kernel(const float4* __restrict__ ct2, int n, float ox, float oy, float oz, float inv_h,
                                 int dx, int dy, int dz, unsigned long long* keys, uint32_t* vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = ct2[i];
    int cx = min(max((int)floorf((p.x - ox) * inv_h), 0), dx - 1);
    int cy = min(max((int)floorf((p.y - oy) * inv_h), 0), dy - 1);
    int cz = min(max((int)floorf((p.z - oz) * inv_h), 0), dz - 1);
    keys[i] = spread21_o(cx) | (spread21_o(cy) << 1) | (spread21_o(cz) << 2);
    vals[i] = (uint32_t)i;
}

// Processing order of the classification queries: patches sorted by the Morton code of the
// target-grid cell of their centroid.  Results are written back in the caller's order, so the
// order only shapes the query groups of the tile search.
static int ensure_patch_order(Ctx* ctx) {
    if (ctx->ct_order_valid) return PWICP_OK;
    const int n = ctx->n2;
    const GridLevel& L = ctx->tgt.dev.lv[0];
    PW_TRY(ctx->keys.reserve(ctx, (size_t)n * 8));
    PW_TRY(ctx->vals.reserve(ctx, (size_t)n * 4));
    PW_TRY(ctx->keys2.reserve(ctx, (size_t)n * 8));
    PW_TRY(ctx->ct_order.reserve(ctx, (size_t)n * 4));
    patch_key_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->ct2.as<float4>(), n, ctx->tgt.dev.ox, ctx->tgt.dev.oy,
                                                             ctx->tgt.dev.oz, L.inv_h, L.dx, L.dy, L.dz,
                                                             ctx->keys.as<unsigned long long>(), ctx->vals.as<uint32_t>());
    int maxd = std::max(L.dx, std::max(L.dy, L.dz));
    int b1 = 1; while ((1 << b1) < maxd && b1 < 21) ++b1;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->keys.as<unsigned long long>(), ctx->keys2.as<unsigned long long>(),
                                    ctx->vals.as<uint32_t>(), ctx->ct_order.as<uint32_t>(), n, 0, 3 * b1, ctx->stream);
    PW_TRY(ctx->cub_tmp.reserve(ctx, tmp));
    size_t cap = ctx->cub_tmp.cap;
    PW_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, cap, ctx->keys.as<unsigned long long>(), ctx->keys2.as<unsigned long long>(),
                                            ctx->vals.as<uint32_t>(), ctx->ct_order.as<uint32_t>(), n, 0, 3 * b1, ctx->stream));
    ctx->launches += 4;
    ctx->ct_order_valid = true;
    return PWICP_OK;
}

// ---- the outer iteration -------------------------------------------------------------------
float bbox_corner_change_host(const double* bb, const float* T) {
    float best = 0.0f;
    for (int k = 0; k < 2; ++k) {
        const float c[4] = {(float)bb[3 * k], (float)bb[3 * k + 1], (float)bb[3 * k + 2], 1.0f};
        float d[3];
        for (int r = 0; r < 3; ++r) {
            float s = T[r * 4] * c[0];
            s += T[r * 4 + 1] * c[1];
            s += T[r * 4 + 2] * c[2];
            s += T[r * 4 + 3] * c[3];
            d[r] = s - c[r];
        }
        const float nrm = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        best = (k == 0) ? nrm : std::max(best, nrm);
    }
    return best;
}

int outer_single_iteration(Ctx* ctx, const pwicp_pair_params& pp, pwicp_state* st,
                           const pwicp_icp_params& icp, float* T16, double* vcm36,
                           unsigned char* stable_flags, pwicp_iter_stats* stats) {
    const int n2 = ctx->n2;
    if (!ctx->tgt.dev.nlevels || n2 < 1 || !ctx->c1.dev.nlevels || ctx->m2 < 1) {
        set_error(ctx, "single_iteration: target, source and clouds must be uploaded first");
        return PWICP_ERR_ARG;
    }
    float& currDT = st->currDT;
    const float DTmin = pp.DTmin;
    if (currDT <= DTmin) currDT = DTmin;                                   // :724-725
    if (4 > n2) { set_error(ctx, "No enough stable points left (<4)"); return PWICP_ERR_TOO_FEW_PATCHES; }

    cudaEvent_t e0, e1;
    PW_CUDA(cudaEventCreate(&e0)); PW_CUDA(cudaEventCreate(&e1));
    PW_CUDA(cudaEventRecord(e0, ctx->stream));

    const float max2minLoD = 2.0f;
    const float maxLoD = DTmin * max2minLoD, minLoD = DTmin;               // :751-753
    // scratch layout: pl[7*n2], pt2pt[n2], lod[n2] | flags[n2], pos[n2] | minmax + npts
    PW_TRY(ctx->scratch_a.reserve(ctx, (size_t)9 * n2 * 4));
    PW_TRY(ctx->flags.reserve(ctx, (size_t)n2 * 4 + (size_t)n2));
    PW_TRY(ctx->pos.reserve(ctx, (size_t)n2 * 4));
    PW_TRY(ctx->scratch_d.reserve(ctx, 256));
    float* pl = ctx->scratch_a.as<float>();
    float* pt2pt = pl + (size_t)7 * n2;
    float* lod = pt2pt + n2;
    int* flags = ctx->flags.as<int>();
    unsigned char* flags_u8 = reinterpret_cast<unsigned char*>(flags + n2);
    int* pos = ctx->pos.as<int>();
    int* minmax = ctx->scratch_d.as<int>();
    unsigned long long* npts = reinterpret_cast<unsigned long long*>(minmax + 2);
    struct { int mn, mx; unsigned long long np; } init = {0x7fffffff, (int)0x80000000, 0ull};
    PW_CUDA(cudaMemcpyAsync(minmax, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));

    const size_t tsmem = 0;
    PW_TRY(ensure_patch_order(ctx));
    classify_dist_kernel<<<(7 * n2 + 255) / 256, 256, tsmem, ctx->stream>>>(
        ctx->tgt.dev, ctx->tgt_aux.as<float4>(), ctx->tgt_ok.as<unsigned char>(), ctx->ct2.as<float4>(),
        ctx->bp2.as<float4>(), ctx->bpstd2.as<float>(), n2, minLoD, maxLoD, pl, pt2pt, lod, minmax,
        ctx->ct_order.as<uint32_t>(), ctx->ct_seed.as<int>(), ctx->bp_seed.as<int>());
    const float DTctct = currDT + 1 * (pp.SVRes1 + pp.SVRes2);             // :817
    classify_flag_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(
        pl, pt2pt, lod, ctx->patch_off.as<int>(), n2, currDT, DTctct, flags, npts);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, flags, pos, n2, ctx->stream);
    PW_TRY(ctx->cub_tmp.reserve(ctx, tmp));
    size_t cap = ctx->cub_tmp.cap;
    PW_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, cap, flags, pos, n2, ctx->stream));
    PW_TRY(ctx->icp_src.reserve(ctx, (size_t)n2 * sizeof(float4)));
    PW_TRY(ctx->icp_seed.reserve(ctx, (size_t)n2 * sizeof(int)));
    compact_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->ct2.as<float4>(), flags, pos, n2,
                                                              ctx->icp_src.as<float4>(), flags_u8,
                                                              ctx->ct_seed.as<int>(), ctx->icp_seed.as<int>());
    ctx->launches += 5;
    struct { int mn, mx; unsigned long long np; int lastpos, lastflag; } h;
    PW_CUDA(cudaMemcpyAsync(&h, minmax, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaMemcpyAsync(&h.lastpos, pos + (n2 - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaMemcpyAsync(&h.lastflag, flags + (n2 - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (stable_flags) PW_CUDA(cudaMemcpyAsync(stable_flags, flags_u8, n2, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    const float LoDet_min = ord2f(h.mn), LoDet_max = ord2f(h.mx);          // :768-769
    const int nStable = h.lastpos + h.lastflag;
    const long long nStablePts = (long long)h.np;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_stable = nStable; stats->n_stable_pts = (int)nStablePts;
        stats->LoDet_min = LoDet_min; stats->LoDet_max = LoDet_max;
        stats->P75 = std::numeric_limits<double>::quiet_NaN();
    }
    if (4 > nStable) {                                                     // :864-867
        set_error(ctx, "No enough stable points left, no enough overlapping areas");
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return PWICP_ERR_TOO_FEW_STABLE;
    }

    // (5) inner ICP on the stable centroids against ALL target centroids, :877
    ctx->n_icp = nStable;
    ctx->icp_seed_valid = true;          // the classification matches are the exact NN of ICP iteration 0
    float transMatICP[16];
    pwicp_icp_result ir;
    PW_TRY(icp_run_device(ctx, icp, transMatICP, &ir, nullptr, nullptr, nullptr));

    // (6) bounding cube of the CURRENT cloud2, :880-888
    float mn[3], mx[3];
    PW_TRY(bbox_packed_dev(ctx, ctx->cloud2.as<float>(), (size_t)ctx->m2, mn, mx));
    double BoundingBox[6];
    octree_cube(mn, mx, (double)(float)(pp.Res2 * 2), BoundingBox);
    const float maxBBchange = bbox_corner_change_host(BoundingBox, transMatICP);
    if (stats) {
        stats->icp_iters = ir.n_iter; stats->icp_state = ir.conv_state; stats->maxBBchange = maxBBchange;
        memcpy(stats->bb6, BoundingBox, sizeof(BoundingBox));
    }

    // (7) DT update, :891-935 (stage-1 block may fall through into the stage-2 block)
    if (!st->toStage2 && maxBBchange < minLoD) st->toStage2 = 1;
    else if (currDT == LoDet_min) st->toStage3 = 1;

    if (!st->toStage2) {
        double Dist75 = 0;
        PW_TRY(percentile_dev(ctx, ctx->c1.dev, ctx->patch_xyz.as<float>(), ctx->mp2, ctx->patch_id.as<int>(),
                              flags, nStablePts, 0.75f, &Dist75, ctx->pp_seed.as<int>()));   // :905
        if (stats) stats->P75 = Dist75;
        if (currDT > Dist75) currDT = Dist75;
        else st->toStage2 = 1;
        if (currDT <= LoDet_min) currDT = LoDet_min;
        st->BBchange_2 = st->BBchange_1;
        st->BBchange_1 = maxBBchange;
    }
    if (st->toStage2 && !st->toStage3) {
        const float upperBound = 0.8f, lowerBound = 0.5f;
        const float alpha = std::abs(st->BBchange_1 / st->BBchange_2);
        if (std::isnan(alpha) || std::isinf(alpha)) currDT = currDT * upperBound;
        else if (alpha < lowerBound) currDT = currDT * lowerBound;
        else if (alpha > upperBound) currDT = currDT * upperBound;
        else currDT = currDT * alpha;
        if (currDT <= LoDet_min) currDT = LoDet_min;
        st->BBchange_2 = st->BBchange_1;
        st->BBchange_1 = maxBBchange;
    }

    // (8) apply the transform to cloud2, CT2, BP2 and every patch, :942-954
    PW_TRY(transform_packed_dev(ctx, ctx->cloud2.as<float>(), (size_t)ctx->m2, transMatICP));
    PW_TRY(transform_f4_dev(ctx, ctx->ct2.as<float4>(), n2, transMatICP));
    PW_TRY(transform_f4_dev(ctx, ctx->bp2.as<float4>(), 6 * n2, transMatICP));
    PW_TRY(transform_packed_dev(ctx, ctx->patch_xyz.as<float>(), (size_t)ctx->mp2, transMatICP));

    // (9) VCM from the pre-update stable centroids (icp_src is never modified by the loop), :957-961
    if (st->toStage3 && vcm36) {
        int sing = 0;
        PW_TRY(vcm_dev(ctx, ctx->icp_src.as<float4>(), nStable, vcm36, &sing, ctx->icp_seed.as<int>()));
        if (stats) { stats->vcm_written = 1; stats->vcm_singular = sing; }
    }
    PW_CUDA(cudaEventRecord(e1, ctx->stream));
    PW_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->last_ms = ms;
    if (stats) stats->device_ms = ms;
    memcpy(T16, transMatICP, sizeof(transMatICP));
    return PWICP_OK;
}

}  // namespace pwicp
