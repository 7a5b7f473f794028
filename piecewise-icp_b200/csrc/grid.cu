// grid.cu -- device-side build of the cell-sorted grid pyramid + the batched exact 1-NN kernel.
//
// One build per target replaces the reference's repeated KD-tree builds over the same cloud
// (five per outer iteration over CTcloud1: src/Registration.cpp:738, :744, :1294 and two inside
// pcl::IterativeClosestPoint; one per stage-1 iteration over cloud1: src/CommonFunc.cpp:269-273).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "nn_search.cuh"

namespace pwicp {

// ---- small utilities ----------------------------------------------------------------------
void GridOwner::release() {
    for (int l = 0; l < kMaxLevels; ++l) { pts[l].release(); cells[l].release(); }
    inv_perm.release(); perm0_buf.release();
    perm0 = nullptr; n = 0;
    dev = GridDev{};
}

int DevBuf::reserve(Ctx* ctx, size_t bytes) {
    if (bytes <= cap && p) return PWICP_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error(ctx, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return PWICP_ERR_NOMEM;
    }
    cap = want;
    return PWICP_OK;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

static int ensure_pinned(Ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return PWICP_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr; ctx->pinned_cap = 0;
    PW_CUDA(cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_cap = bytes;
    return PWICP_OK;
}

int upload_packed(Ctx* ctx, DevBuf& buf, const float* host_xyz, size_t n_floats) {
    PW_TRY(buf.reserve(ctx, n_floats * sizeof(float) + 64));
    if (n_floats)
        PW_CUDA(cudaMemcpyAsync(buf.p, host_xyz, n_floats * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return PWICP_OK;
}

__global__ void finite_kernel(const float* __restrict__ v, size_t n, int* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    int b = 0;
    for (; i < n; i += stride) {
        float x = v[i];
        if (!(fabsf(x) <= FLT_MAX)) b = 1;
    }
    if (b) atomicOr(bad, 1);
}

int check_finite_dev(Ctx* ctx, const float* dev, size_t n_floats, bool* ok) {
    PW_TRY(ensure_pinned(ctx, 4096));
    PW_TRY(ctx->scratch_d.reserve(ctx, 256));
    int* flag = ctx->scratch_d.as<int>();
    PW_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    if (n_floats) {
        int blocks = (int)std::min<size_t>((n_floats + 255) / 256, (size_t)ctx->num_sms * 8);
        finite_kernel<<<blocks, 256, 0, ctx->stream>>>(dev, n_floats, flag);
        ctx->launches++;
    }
    PW_CUDA(cudaMemcpyAsync(ctx->pinned, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    *ok = (*(int*)ctx->pinned == 0);
    return PWICP_OK;
}

int finite_accumulate_dev(Ctx* ctx, const float* dev, size_t n_floats, int* flag_dev) {
    if (!n_floats) return PWICP_OK;
    int blocks = (int)std::min<size_t>((n_floats + 255) / 256, (size_t)ctx->num_sms * 8);
    finite_kernel<<<blocks, 256, 0, ctx->stream>>>(dev, n_floats, flag_dev);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

// ---- bounding box -------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return (i >= 0) ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_to_float(int i) {
    int j = (i >= 0) ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(j);
#else
    float f; memcpy(&f, &j, 4); return f;
#endif
}

// min/max over packed xyz.  out[0..2] = ordered-int min, out[3..5] = ordered-int max.
// Four points = three 16-byte loads per thread and trip (the arrays come from cudaMalloc: 16-byte aligned), the warps of
// a block folded in shared memory before the six atomics (round 1 / 2: scalar loads, six atomics per warp -- 43 us per
// 2.4M points in the ncu launch list of the outer loop, 7 % of it).
__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ xyz, size_t n, int* out) {
    __shared__ float s_mn[8][3], s_mx[8][3];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    auto take = [&](float x, float y, float z) {
        mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
        mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
        mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
    };
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    const size_t nq = (reinterpret_cast<uintptr_t>(xyz) & 15) ? 0 : n / 4;   // groups of four points (an unaligned array: scalar loads)
    const float4* __restrict__ v = reinterpret_cast<const float4*>(xyz);
    for (size_t g = tid; g < nq; g += stride) {
        const float4 a = v[3 * g], b = v[3 * g + 1], c = v[3 * g + 2];
        take(a.x, a.y, a.z); take(a.w, b.x, b.y); take(b.z, b.w, c.x); take(c.y, c.z, c.w);
    }
    for (size_t i = 4 * nq + tid; i < n; i += stride) take(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    for (int c = 0; c < 3; ++c)
        for (int o = 16; o; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int c = 0; c < 3; ++c) { s_mn[warp][c] = mn[c]; s_mx[warp][c] = mx[c]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float a = s_mn[0][c], b = s_mx[0][c];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fminf(a, s_mn[w][c]); b = fmaxf(b, s_mx[w][c]); }
        atomicMin(out + c, float_to_ordered(a));
        atomicMax(out + 3 + c, float_to_ordered(b));
    }
}

__global__ void bbox_init_kernel(int* out) {
    if (threadIdx.x < 3) out[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) out[threadIdx.x] = (int)0x80000000;
}

// min / max accumulated into six ordered ints that the caller has initialised on the device (no host round trip)
int bbox_accumulate_dev(Ctx* ctx, const float* xyz_dev, size_t n, int* out6_dev) {
    int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 8);
    bbox_kernel<<<std::max(blocks, 1), 256, 0, ctx->stream>>>(xyz_dev, n, out6_dev);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

int bbox_packed_dev(Ctx* ctx, const float* xyz_dev, size_t n, float* mn3, float* mx3, const int* flag_dev, int* flag_out) {
    PW_TRY(ensure_pinned(ctx, 4096));
    PW_TRY(ctx->scratch_d.reserve(ctx, 256));
    int* out = ctx->scratch_d.as<int>();
    bbox_init_kernel<<<1, 32, 0, ctx->stream>>>(out);
    int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 8);
    if (blocks < 1) blocks = 1;
    bbox_kernel<<<blocks, 256, 0, ctx->stream>>>(xyz_dev, n, out);
    ctx->launches += 2;
    PW_CUDA(cudaMemcpyAsync(ctx->pinned, out, 6 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    // a device flag of the caller (the verdict of a finite check) rides on the same round trip
    if (flag_dev) PW_CUDA(cudaMemcpyAsync((int*)ctx->pinned + 8, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    const int* h = (const int*)ctx->pinned;
    for (int c = 0; c < 3; ++c) { mn3[c] = ordered_to_float(h[c]); mx3[c] = ordered_to_float(h[3 + c]); }
    if (flag_dev && flag_out) *flag_out = h[8];
    return PWICP_OK;
}

// ---- build ---------------------------------------------------------------------------------
// Counting sort by cell, hand-written (round 1 went through cub::DeviceRadixSort + DeviceScan: a third of the build was
// the 32-bit key sort, profiles/r01j_launch_shares.txt).  Cell keys are bounded by the number of cells, so per level:
//   count   : one atomic per (warp, distinct cell) into A[2 + cell]
//   scan    : inclusive prefix sum of A in place (three launches)  ->  A[1 + cell] = first slot of the cell
//   scatter : slot = A[1 + cell]++, again one atomic per (warp, distinct cell); the point goes straight into the
//             sorted float4 array.  Afterwards A[1 + cell] is the END of the cell, i.e. A[cell] its start: A is the
//             cell_start array the searches read (A[0] = 0, A[ncells] = n).
// The order of the points INSIDE a cell is whatever the atomics make it.  Nothing reads it: every search compares
// (distance, original index) explicitly, so indices and distances do not depend on it.
__device__ __forceinline__ uint32_t cell_key_of(const float* __restrict__ xyz, int i, float ox, float oy, float oz,
                                                float inv_h, int dx, int dy, int dz, float4& p) {
    p.x = xyz[3 * (size_t)i]; p.y = xyz[3 * (size_t)i + 1]; p.z = xyz[3 * (size_t)i + 2];
    // identical expression to the searches: (p - origin) * inv_h, floor, clamp
    const float fx = (p.x - ox) * inv_h, fy = (p.y - oy) * inv_h, fz = (p.z - oz) * inv_h;
    const int cx = min(max((int)floorf(fx), 0), dx - 1);
    const int cy = min(max((int)floorf(fy), 0), dy - 1);
    const int cz = min(max((int)floorf(fz), 0), dz - 1);
    return ((uint32_t)cz * (uint32_t)dy + (uint32_t)cy) * (uint32_t)dx + (uint32_t)cx;
}

// one level per launch (large clouds, see grid_build)
__global__ void __launch_bounds__(256)
cell_count_kernel(const float* __restrict__ xyz, int n, float ox, float oy, float oz, float inv_h, int dx, int dy, int dz,
                  uint32_t* __restrict__ A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p;
    const uint32_t key = cell_key_of(xyz, i, ox, oy, oz, inv_h, dx, dy, dz, p);
    // neighbours in the caller's order usually share a cell (always on the coarse levels): one atomic per group
    const unsigned m = __match_any_sync(__activemask(), key);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(A + 2 + key, (uint32_t)__popc(m));
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(const float* __restrict__ xyz, int n, float ox, float oy, float oz, float inv_h, int dx, int dy, int dz,
                    uint32_t* __restrict__ A, float4* __restrict__ out, uint32_t* __restrict__ inv_perm,
                    uint32_t* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p;
    const uint32_t key = cell_key_of(xyz, i, ox, oy, oz, inv_h, dx, dy, dz, p);
    const unsigned m = __match_any_sync(__activemask(), key);
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(A + 1 + key, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1));
    p.w = __int_as_float(i);
    out[slot] = p;
    if (inv_perm) { inv_perm[i] = slot; perm[slot] = (uint32_t)i; }
}

// all levels of a grid in one pass over the points: one count launch, one segmented scan (three launches), one scatter
// launch -- 8 launches for three levels where a loop over the levels took 18 (0.245 -> see DESIGN.md 2)
constexpr int kFuseLevelsMaxPoints = 2000000;   // grid_build: up to this size all levels share the count / scan / scatter launches
struct LevelBuild { float inv_h; int dx, dy, dz; uint32_t* A; float4* pts; };
struct LevelsBuild { LevelBuild lv[kMaxLevels]; int nlev; };

__global__ void __launch_bounds__(256)
cell_count_levels_kernel(const float* __restrict__ xyz, int n, float ox, float oy, float oz, const LevelsBuild B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int l = 0; l < B.nlev; ++l) {
        float4 p;
        const LevelBuild& L = B.lv[l];
        const uint32_t key = cell_key_of(xyz, i, ox, oy, oz, L.inv_h, L.dx, L.dy, L.dz, p);
        const unsigned m = __match_any_sync(__activemask(), key);
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(L.A + 2 + key, (uint32_t)__popc(m));
    }
}

__global__ void __launch_bounds__(256)
cell_scatter_levels_kernel(const float* __restrict__ xyz, int n, float ox, float oy, float oz, const LevelsBuild B,
                           uint32_t* __restrict__ inv_perm, uint32_t* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    for (int l = 0; l < B.nlev; ++l) {
        float4 p;
        const LevelBuild& L = B.lv[l];
        const uint32_t key = cell_key_of(xyz, i, ox, oy, oz, L.inv_h, L.dx, L.dy, L.dz, p);
        const unsigned m = __match_any_sync(__activemask(), key);
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(L.A + 1 + key, (uint32_t)__popc(m));
        base = __shfl_sync(m, base, leader);
        const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1));
        p.w = __int_as_float(i);
        L.pts[slot] = p;
        if (l == 0) { inv_perm[i] = slot; perm[slot] = (uint32_t)i; }
    }
}

// inclusive prefix sum of n 32-bit counters in place: tile sums, scan of the tile sums (one block), tiles
constexpr int kScanThreads = 512, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ void scan_tile_sum_body(const uint32_t* __restrict__ a, size_t n, uint32_t* __restrict__ tsum, int tile) {
    __shared__ uint32_t s_w[kScanThreads / 32];
    const size_t base = (size_t)tile * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) s += a[base + k];
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = (threadIdx.x < kScanThreads / 32) ? s_w[threadIdx.x] : 0u;
        v = __reduce_add_sync(0xffffffffu, v);
        if (threadIdx.x == 0) tsum[tile] = v;
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sum_kernel(const uint32_t* __restrict__ a, size_t n, uint32_t* __restrict__ tsum) {
    scan_tile_sum_body(a, n, tsum, blockIdx.x);
}

// several arrays at once (the cell counters of all levels of a grid): block b works on tile b - tile0[s] of segment s
struct ScanSegs {
    uint32_t* a[kMaxLevels];
    size_t n[kMaxLevels];
    uint32_t* tsum[kMaxLevels];
    int tile0[kMaxLevels + 1];
    int nseg;
};
__device__ __forceinline__ int scan_seg_of(const ScanSegs& S, int b) {
    int s = 0;
    while (s + 1 < S.nseg && b >= S.tile0[s + 1]) ++s;
    return s;
}
__global__ void __launch_bounds__(kScanThreads)
scan_tile_sum_segs_kernel(const ScanSegs S) {
    const int s = scan_seg_of(S, blockIdx.x);
    scan_tile_sum_body(S.a[s], S.n[s], S.tsum[s], blockIdx.x - S.tile0[s]);
}

// exclusive scan of the tile sums in place (one block, any count)
__device__ __forceinline__ void scan_tsum_body(uint32_t* __restrict__ tsum, int nt) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nt; b0 += 1024) {
        const int i = b0 + tid;
        const uint32_t v = (i < nt) ? tsum[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (warp ? s_warp[warp - 1] : 0u) + s_carry;
        if (i < nt) tsum[i] = incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
scan_tsum_kernel(uint32_t* __restrict__ tsum, int nt) { scan_tsum_body(tsum, nt); }
__global__ void __launch_bounds__(1024)
scan_tsum_segs_kernel(const ScanSegs S) { scan_tsum_body(S.tsum[blockIdx.x], S.tile0[blockIdx.x + 1] - S.tile0[blockIdx.x]); }

__device__ __forceinline__ void scan_tile_apply_body(uint32_t* __restrict__ a, size_t n, const uint32_t* __restrict__ toff, int tile) {
    __shared__ uint32_t s_w[kScanThreads / 32];
    const size_t base = (size_t)tile * kScanTile + (size_t)threadIdx.x * kScanItems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < n) ? a[base + k] : 0u; s += v[k]; v[k] = s; }
    uint32_t x = s;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < kScanThreads / 32) ? s_w[lane] : 0u;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        if (lane < kScanThreads / 32) s_w[lane] = w;
    }
    __syncthreads();
    const uint32_t before = toff[tile] + (warp ? s_w[warp - 1] : 0u) + (x - s);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) a[base + k] = before + v[k];
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_apply_kernel(uint32_t* __restrict__ a, size_t n, const uint32_t* __restrict__ toff) {
    scan_tile_apply_body(a, n, toff, blockIdx.x);
}
__global__ void __launch_bounds__(kScanThreads)
scan_tile_apply_segs_kernel(const ScanSegs S) {
    const int s = scan_seg_of(S, blockIdx.x);
    scan_tile_apply_body(S.a[s], S.n[s], S.tsum[s], blockIdx.x - S.tile0[s]);
}

int scan_inclusive_inplace(Ctx* ctx, uint32_t* a, size_t n) {
    const int nt = (int)((n + kScanTile - 1) / kScanTile);
    PW_TRY(ctx->vals.reserve(ctx, (size_t)nt * 4));
    uint32_t* tsum = ctx->vals.as<uint32_t>();
    scan_tile_sum_kernel<<<nt, kScanThreads, 0, ctx->stream>>>(a, n, tsum);
    scan_tsum_kernel<<<1, 1024, 0, ctx->stream>>>(tsum, nt);
    scan_tile_apply_kernel<<<nt, kScanThreads, 0, ctx->stream>>>(a, n, tsum);
    ctx->launches += 3;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

// ... of up to kMaxLevels arrays in three launches
static int scan_inclusive_inplace_segs(Ctx* ctx, uint32_t* const* a, const size_t* n, int nseg) {
    ScanSegs S{};
    S.nseg = nseg;
    int tiles = 0;
    for (int k = 0; k < nseg; ++k) { S.tile0[k] = tiles; tiles += (int)((n[k] + kScanTile - 1) / kScanTile); }
    S.tile0[nseg] = tiles;
    PW_TRY(ctx->vals.reserve(ctx, (size_t)tiles * 4));
    for (int k = 0; k < nseg; ++k) { S.a[k] = a[k]; S.n[k] = n[k]; S.tsum[k] = ctx->vals.as<uint32_t>() + S.tile0[k]; }
    scan_tile_sum_segs_kernel<<<tiles, kScanThreads, 0, ctx->stream>>>(S);
    scan_tsum_segs_kernel<<<nseg, 1024, 0, ctx->stream>>>(S);
    scan_tile_apply_segs_kernel<<<tiles, kScanThreads, 0, ctx->stream>>>(S);
    ctx->launches += 3;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

// ---- processing order of a query set -----------------------------------------------------------
// The inner loop and the classification work through their queries in an order in which consecutive queries are
// spatial neighbours (the 32 lanes of a warp then walk the same few cell rows).  Round 1 and the first half of round
// 2 sorted Morton codes of the target-grid cell with cub::DeviceRadixSort (four onesweep passes, 3 % of a bench
// step); the keys are bounded, so this is a counting sort like the grid build's: bins = cells of twice the finest
// cell size, enumerated block by block (8 x 8 x 8 bins, Morton order inside a block, blocks row-major), one
// warp-aggregated atomic per bin for the count and for the slot, and a last pass that ranks the points of a bin by
// their index in the caller's order -- the order is a function of the input alone, whatever the atomics do (the
// summation order of the inner loop depends on it); only bins of more than 512 points, i.e. degenerate inputs, keep
// the order of the atomics.
constexpr uint32_t kRankMaxBin = 512;   // order_rank_kernel: bins up to this size are ranked by index in the caller's order
__device__ __forceinline__ uint32_t order_key_of(float x, float y, float z, float ox, float oy, float oz, float inv_h,
                                                 int dx, int dy, int dz, int nbx, int nby) {
    const int cx = min(max((int)floorf((x - ox) * inv_h), 0), dx - 1) >> 1;
    const int cy = min(max((int)floorf((y - oy) * inv_h), 0), dy - 1) >> 1;
    const int cz = min(max((int)floorf((z - oz) * inv_h), 0), dz - 1) >> 1;
    const uint32_t lx = cx & 7, ly = cy & 7, lz = cz & 7;
    auto spread3 = [](uint32_t v) { return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4); };
    const uint32_t m = spread3(lx) | (spread3(ly) << 1) | (spread3(lz) << 2);
    return (((uint32_t)(cz >> 3) * (uint32_t)nby + (uint32_t)(cy >> 3)) * (uint32_t)nbx + (uint32_t)(cx >> 3)) * 512u + m;
}

__global__ void __launch_bounds__(256)
order_count_kernel(const float4* __restrict__ pts, int n, float ox, float oy, float oz, float inv_h, int dx, int dy, int dz,
                   int nbx, int nby, uint32_t* __restrict__ A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const uint32_t key = order_key_of(p.x, p.y, p.z, ox, oy, oz, inv_h, dx, dy, dz, nbx, nby);
    const unsigned m = __match_any_sync(__activemask(), key);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(A + 2 + key, (uint32_t)__popc(m));
}

__global__ void __launch_bounds__(256)
order_scatter_kernel(const float4* __restrict__ pts, int n, float ox, float oy, float oz, float inv_h, int dx, int dy, int dz,
                     int nbx, int nby, uint32_t* __restrict__ A, uint32_t* __restrict__ slots, uint32_t* __restrict__ myslot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const uint32_t key = order_key_of(p.x, p.y, p.z, ox, oy, oz, inv_h, dx, dy, dz, nbx, nby);
    const unsigned m = __match_any_sync(__activemask(), key);
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(A + 1 + key, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1));
    slots[slot] = (uint32_t)i;
    myslot[i] = slot;
}

// after the scatter A[key] .. A[key + 1] is the bin of `key`; the point takes the place of its rank among the bin's
// points by index in the caller's order
__global__ void __launch_bounds__(256)
order_rank_kernel(const float4* __restrict__ pts, int n, float ox, float oy, float oz, float inv_h, int dx, int dy, int dz,
                  int nbx, int nby, const uint32_t* __restrict__ A, const uint32_t* __restrict__ slots,
                  const uint32_t* __restrict__ myslot, uint32_t* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const uint32_t key = order_key_of(p.x, p.y, p.z, ox, oy, oz, inv_h, dx, dy, dz, nbx, nby);
    const uint32_t s = A[key], e = A[key + 1];
    // a bin that holds more than kRankMaxBin points (a degenerate input: a source far outside the target's box clamps
    // into one boundary cell) keeps the order the atomics made -- the quadratic ranking would take minutes at 1M
    if (e - s > kRankMaxBin) { order[myslot[i]] = (uint32_t)i; return; }
    uint32_t rank = 0;
    for (uint32_t j = s; j < e; ++j) rank += (slots[j] < (uint32_t)i) ? 1u : 0u;
    order[s + rank] = (uint32_t)i;
}

int spatial_order_dev(Ctx* ctx, const GridDev& g, const float4* pts, int n, uint32_t* order) {
    const GridLevel& L = g.lv[0];
    const int nbx = (((L.dx - 1) >> 1) >> 3) + 1, nby = (((L.dy - 1) >> 1) >> 3) + 1, nbz = (((L.dz - 1) >> 1) >> 3) + 1;
    const uint64_t nbins = (uint64_t)nbx * nby * nbz * 512ull;
    if (nbins > 0x7ffffff0ull) { set_error(ctx, "spatial_order: too many bins"); return PWICP_ERR_ARG; }
    PW_TRY(ctx->keys.reserve(ctx, (nbins + 2) * sizeof(uint32_t)));
    PW_TRY(ctx->keys2.reserve(ctx, (size_t)2 * n * sizeof(uint32_t)));
    uint32_t* A = ctx->keys.as<uint32_t>();
    uint32_t* slots = ctx->keys2.as<uint32_t>();
    uint32_t* myslot = slots + n;
    const int blocks = (n + 255) / 256;
    PW_CUDA(cudaMemsetAsync(A, 0, (nbins + 2) * sizeof(uint32_t), ctx->stream));
    order_count_kernel<<<blocks, 256, 0, ctx->stream>>>(pts, n, g.ox, g.oy, g.oz, L.inv_h, L.dx, L.dy, L.dz, nbx, nby, A);
    PW_TRY(scan_inclusive_inplace(ctx, A, (size_t)nbins + 2));
    order_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(pts, n, g.ox, g.oy, g.oz, L.inv_h, L.dx, L.dy, L.dz, nbx, nby, A, slots, myslot);
    order_rank_kernel<<<blocks, 256, 0, ctx->stream>>>(pts, n, g.ox, g.oy, g.oz, L.inv_h, L.dx, L.dy, L.dz, nbx, nby, A, slots, myslot, order);
    ctx->launches += 3;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

// bad_flag_dev: a device flag that a finite check of xyz accumulates into (enqueued before this call); it is read with
// the bounding box, and a set flag ends the build with PWICP_ERR_NONFINITE before the box is used.  sync_at_end = false:
// the caller keeps enqueuing on the library stream (host-buffer call of the inner loop).
int grid_build(Ctx* ctx, GridOwner& g, const float* xyz, int n, const int* bad_flag_dev, bool sync_at_end) {
    g.dev = GridDev{};                 // invalid until the build completes; buffers are reused
    g.n = 0;
    if (n < 1) { set_error(ctx, "grid_build: empty target"); return PWICP_ERR_ARG; }
    float mn[3], mx[3];
    int bad = 0;
    PW_TRY(bbox_packed_dev(ctx, xyz, (size_t)n, mn, mx, bad_flag_dev, &bad));
    if (bad) { set_error(ctx, "non-finite value in input"); return PWICP_ERR_NONFINITE; }
    double ext[3];
    for (int c = 0; c < 3; ++c) ext[c] = std::max((double)mx[c] - (double)mn[c], 0.0);
    double emax = std::max(ext[0], std::max(ext[1], ext[2]));
    if (!(emax > 0)) emax = 1.0;
    // finest cell size: about cells_per_point cells per point over the bounding box, where
    // degenerate (flat) axes count as one cell
    const double budget = std::max(64.0, (double)ctx->cells_per_point * (double)n);
    double h = emax / 1024.0;
    {
        // solve prod(max(1, ext/h)) ~= budget by bisection on log h
        double lo = emax * 1e-7, hi = emax * 2.0;
        for (int it = 0; it < 80; ++it) {
            double mid = std::sqrt(lo * hi);
            double cells = 1.0;
            for (int c = 0; c < 3; ++c) cells *= std::max(1.0, std::ceil(ext[c] / mid + 1e-9));
            if (cells > budget) lo = mid; else hi = mid;
        }
        h = hi;
    }
    // keep the coordinates in cell units small enough for float (|f| * 4e-6 margin stays tiny)
    h = std::max(h, emax / 30000.0);

    g.h0 = (float)h;
    g.n = n;
    g.dev.ox = mn[0]; g.dev.oy = mn[1]; g.dev.oz = mn[2];
    g.dev.n = n;
    int nlev = 0;
    double hl = h;
    const int blocks = (n + 255) / 256;
    LevelsBuild B{};
    uint32_t* Aseg[kMaxLevels];
    size_t nseg[kMaxLevels];
    PW_TRY(g.inv_perm.reserve(ctx, (size_t)n * sizeof(uint32_t)));
    PW_TRY(g.perm0_buf.reserve(ctx, (size_t)n * sizeof(uint32_t)));
    g.perm0 = g.perm0_buf.as<uint32_t>();
    for (int l = 0; l < kMaxLevels; ++l) {
        int d[3];
        for (int c = 0; c < 3; ++c) d[c] = (int)std::max(1.0, std::floor(ext[c] / hl) + 1.0);
        uint64_t ncells = (uint64_t)d[0] * d[1] * d[2];
        if (ncells > 0x7fffffffull) { set_error(ctx, "grid_build: too many cells"); return PWICP_ERR_ARG; }
        float inv_h = (float)(1.0 / hl);

        PW_TRY(g.cells[l].reserve(ctx, (ncells + 2) * sizeof(uint32_t)));
        PW_TRY(g.pts[l].reserve(ctx, (size_t)n * sizeof(float4)));
        uint32_t* A = g.cells[l].as<uint32_t>();
        float4* pts = g.pts[l].as<float4>();
        PW_CUDA(cudaMemsetAsync(A, 0, (ncells + 2) * sizeof(uint32_t), ctx->stream));
        B.lv[l] = LevelBuild{inv_h, d[0], d[1], d[2], A, pts};
        Aseg[l] = A; nseg[l] = (size_t)ncells + 2;

        GridLevel& L = g.dev.lv[l];
        L.pts = pts; L.cell_start = A;
        L.dx = d[0]; L.dy = d[1]; L.dz = d[2];
        L.inv_h = inv_h; L.inv_h2 = inv_h * inv_h;
        nlev = l + 1;
        if (d[0] <= 4 && d[1] <= 4 && d[2] <= 4) break;
        hl *= kLevelFactor;
    }
    // count, scan, scatter: every level in the same three steps (8 launches for three levels instead of 18)
    B.nlev = nlev;
    if (n <= kFuseLevelsMaxPoints) {
        // count, scan, scatter: every level in the same three steps (8 launches for three levels instead of 18)
        cell_count_levels_kernel<<<blocks, 256, 0, ctx->stream>>>(xyz, n, mn[0], mn[1], mn[2], B);
        PW_TRY(scan_inclusive_inplace_segs(ctx, Aseg, nseg, nlev));
        cell_scatter_levels_kernel<<<blocks, 256, 0, ctx->stream>>>(xyz, n, mn[0], mn[1], mn[2], B, g.inv_perm.as<uint32_t>(), g.perm0);
        ctx->launches += 2;
    } else {
        // large clouds: level after level (at 10M the shared launches are 45 % slower, profiles/r02as_build_ab.txt)
        for (int l = 0; l < nlev; ++l) {
            const LevelBuild& V = B.lv[l];
            cell_count_kernel<<<blocks, 256, 0, ctx->stream>>>(xyz, n, mn[0], mn[1], mn[2], V.inv_h, V.dx, V.dy, V.dz, V.A);
            PW_TRY(scan_inclusive_inplace(ctx, V.A, nseg[l]));
            cell_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(xyz, n, mn[0], mn[1], mn[2], V.inv_h, V.dx, V.dy, V.dz, V.A, V.pts,
                                                                l == 0 ? g.inv_perm.as<uint32_t>() : nullptr, g.perm0);
            ctx->launches += 2;
        }
    }
    g.dev.nlevels = nlev;
    g.dev.inv_perm = g.inv_perm.as<uint32_t>();
    PW_CUDA(cudaGetLastError());
    if (sync_at_end) PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

// ---- batched query -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nn_kernel(GridDev g, const float* __restrict__ q, int nq, int* __restrict__ idx, float* __restrict__ d2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < nq;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (active) { px = q[3 * (size_t)i]; py = q[3 * (size_t)i + 1]; pz = q[3 * (size_t)i + 2]; }
    const Best b = nn_search_warp(g, px, py, pz, -1, active);     // warp-collective (nn_search.cuh, team search)
    if (active) {
        if (idx) idx[i] = b.idx;
        if (d2) d2[i] = b.d2;
    }
}

// ---- self NN: distance of every target point to its nearest OTHER point ---------------------
// Serves calPCresolution (reference src/CommonFunc.cpp:239-263: KdTreeFLANN::nearestKSearch(i, 2),
// second neighbour).  One thread per sorted position; the sorted-array neighbour is the seed, the
// ball through it is scanned with the thread's own index excluded.  d2 is written in the
// caller's (original) order.
__global__ void __launch_bounds__(256)
self_nn_kernel(GridDev g, float* __restrict__ d2_out) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const GridLevel& L = g.lv[0];
    const float4 p = __ldg(L.pts + pos);
    const int self = __float_as_int(p.w);
    const float4 s = __ldg(L.pts + (pos + 1 < g.n ? pos + 1 : pos - 1));
    float bd = l2_simple(p.x, p.y, p.z, s.x, s.y, s.z);
    const float fx = (p.x - g.ox) * L.inv_h, fy = (p.y - g.oy) * L.inv_h, fz = (p.z - g.oz) * L.inv_h;
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f, mz = 0.01f + fabsf(fz) * 4e-6f;
    const float r = sqrtf(bd) * L.inv_h * 1.00001f;
    const int lx = min(max((int)floorf(fx - r - mx), 0), L.dx - 1), hx = min(max((int)floorf(fx + r + mx), 0), L.dx - 1);
    const int ly = min(max((int)floorf(fy - r - my), 0), L.dy - 1), hy = min(max((int)floorf(fy + r + my), 0), L.dy - 1);
    const int lz = min(max((int)floorf(fz - r - mz), 0), L.dz - 1), hz = min(max((int)floorf(fz + r + mz), 0), L.dz - 1);
    for (int kz = lz; kz <= hz; ++kz)
        for (int ky = ly; ky <= hy; ++ky) {
            const float gy = axis_gap(fy, ky, my), gz = axis_gap(fz, kz, mz);
            if (gy * gy + gz * gz > bd * L.inv_h2) continue;
            const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
            const uint32_t b = __ldg(L.cell_start + row + lx), e = __ldg(L.cell_start + row + hx + 1);
            for (uint32_t i = b; i < e; ++i) {
                const float4 q = __ldg(L.pts + i);
                if (__float_as_int(q.w) == self) continue;
                const float d = l2_simple(p.x, p.y, p.z, q.x, q.y, q.z);
                if (d < bd) bd = d;
            }
        }
    d2_out[self] = bd;
}

int self_nn_dev(Ctx* ctx, const GridDev& g, float* d2_dev) {
    if (g.n < 2) { set_error(ctx, "self_nn: needs at least 2 points"); return PWICP_ERR_ARG; }
    self_nn_kernel<<<(g.n + 255) / 256, 256, 0, ctx->stream>>>(g, d2_dev);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

int nn_query_packed(Ctx* ctx, const GridDev& g, const float* q_dev, int nq, int* idx_dev, float* d2_dev) {
    if (nq <= 0) return PWICP_OK;
    const size_t smem = 0;
    int blocks = (nq + 255) / 256;
    nn_kernel<<<blocks, 256, smem, ctx->stream>>>(g, q_dev, nq, idx_dev, d2_dev);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

}  // namespace pwicp
