// outer.cu -- one outer iteration of Piecewise-ICP on the device-resident pair.
//
// Follows PwICP_singleIteration (reference src/Registration.cpp:704-972) statement by statement
// (SURVEY.md appendix A): (1) NN CT2->CT1 and BP2->CT1, (2) LoD per patch, (3) point-to-plane
// distances, (4) stable/unstable classification + order-preserving compaction, (5) inner ICP,
// (6) bounding-cube corner change, (7) DT schedule incl. the stage-1 percentile, (8) transform of
// cloud2 / CT2 / BP2 / patches, (9) VCM on the pre-update stable centroids.
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>

#include "common.cuh"
#include "nn_search.cuh"
#include "small_algebra.cuh"

namespace pwicp {

// Device-resident state of one outer iteration.  The host writes the first block (state carried between iterations
// + the reset values of the reductions) at the start, every kernel of the iteration reads / writes it in place, and
// the host reads it back ONCE, after the last kernel: one synchronisation per outer iteration (two in the last one,
// for the VCM).  Round 1 needed six.
struct SelectDev {                       // k-th smallest of non-negative floats: radix select, 11 + 11 + 10 bits
    unsigned prefix;                     // bit pattern found so far
    unsigned ticket;                     // blocks that have added their histogram (last one picks the bin)
    unsigned long long rank;             // rank of the wanted element among those sharing the prefix
    unsigned hist[2048];
};
struct OuterDev {
    // -- written by the host at the start of the iteration
    float currDT, BBchange_1, BBchange_2;
    int toStage2, toStage3;
    float minLoD;
    double res2;                         // octree resolution (float(Res2 * 2) widened, src/Registration.cpp:883)
    int lod_min, lod_max;                // ordered-int min / max of the per-patch LoDetection
    unsigned long long n_stable_pts;     // points of the stable patches
    int bbox[6];                         // ordered-int min xyz / max xyz of cloud2
    int err;                             // pwicp_status raised on the device (0 = ok)
    int ran_p75;                         // the stage-1 percentile was evaluated in this iteration
    // -- results
    int n_stable;
    int need_p75;                        // stage 1 goes on: the percentile of this iteration is wanted
    int icp_iters, icp_state;
    float T[16];
    float maxBBchange, LoDet_min, LoDet_max;
    int pad0;
    double bb6[6];
    double P75;
    SelectDev sel;
};
constexpr size_t kOuterInitBytes = offsetof(OuterDev, n_stable);
constexpr size_t kOuterReadBytes = offsetof(OuterDev, sel);

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
    // thread u handles the u-th query of the spatially ordered patch list (spatial_order_dev) (spatially compact groups);
    // t is the query's slot in the caller's order: centroid t < n2, boundary point t - n2.  (Seeding the boundary points
    // of the first iteration with their centroid's match, in a second launch, was measured: no gain, r02r.)
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
    const Best b = nn_search_warp(g, p.x, p.y, p.z, seed, active);   // warp-collective (nn_search.cuh, team search)
    if (!active) return;
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
// Thread u handles the u-th patch of the processing order (t = order[u]); the flags are stored in the caller's order, the
// number of stable patches per 256-thread block in processing order (first level of the compaction scan).
constexpr int kCompactBlock = 256;
__global__ void __launch_bounds__(kCompactBlock)
classify_flag_kernel(const float* __restrict__ pl, const float* __restrict__ pt2pt, const float* __restrict__ lod,
                     const int* __restrict__ patch_off, const uint32_t* __restrict__ order, int n2, float currDT,
                     float DTctct, int* __restrict__ flags, unsigned char* __restrict__ flags_u8,
                     int* __restrict__ block_cnt, OuterDev* st) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    bool stable = false;
    unsigned int np = 0;
    if (u < n2) {
        const int i = (int)order[u];
        const float L = lod[i];
        const float thr = (currDT <= L) ? L : currDT;
        bool pass = true;
        const float* bp = pl + n2 + 6 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 6; ++k) if (thr < bp[k]) pass = false;
        if (thr < pl[i]) pass = false;
        stable = pass && (pt2pt[i] < DTctct);
        flags[i] = stable ? 1 : 0;
        if (flags_u8) flags_u8[i] = stable ? 1 : 0;
        np = stable ? (unsigned int)(patch_off[i + 1] - patch_off[i]) : 0u;
    }
    // points of the stable patches: one atomic per warp (integer: deterministic); all 32 lanes are here
    np = __reduce_add_sync(0xffffffffu, np);
    if ((threadIdx.x & 31) == 0 && np) atomicAdd(&st->n_stable_pts, (unsigned long long)np);
    const int cnt = __syncthreads_count(stable);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = cnt;
}

// a padding point of the inner loop's arrays (icp.cu): zero normal
__device__ __forceinline__ void write_icp_pad(int i, float4* sorted, float4* cn, float4* cq) {
    sorted[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    cq[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    cn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// second level of the scan: exclusive prefix of the block counts (one block), the total, and the pads of the inner
// loop's arrays up to the next whole batch
__global__ void __launch_bounds__(1024)
scan_blocks_kernel(int* __restrict__ block_cnt, int nblocks, OuterDev* st, float4* sorted, float4* cn, float4* cq) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + tid;
        const int v = (i < nblocks) ? block_cnt[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int incl = x + (warp ? s_warp[warp - 1] : 0) + s_carry;
        if (i < nblocks) block_cnt[i] = incl - v;              // exclusive
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
    const int total = s_carry;
    if (tid == 0) st->n_stable = total;
    const int padded = (total + 31) / 32 * 32;
    if (total + tid < padded) write_icp_pad(total + tid, sorted, cn, cq);
}

// third level: position of every stable patch, and the stable set written in the layout of the inner loop
// (icp.cu): point (w = its rank), its classification match and that match's normal inline.  The stable set is in
// processing order of the patches, so the inner loop needs no sort of its own.
__global__ void __launch_bounds__(kCompactBlock)
compact_kernel(const float4* __restrict__ ct2, const int* __restrict__ flags, const uint32_t* __restrict__ order,
               const int* __restrict__ block_off, int n2, const int* __restrict__ ct_seed,
               const float4* __restrict__ tgt_pts, const float4* __restrict__ tgt_aux,
               float4* __restrict__ sorted, float4* __restrict__ cn, float4* __restrict__ cq) {
    __shared__ int s_warp[kCompactBlock / 32];
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = (u < n2) ? (int)order[u] : 0;
    const bool f = (u < n2) && flags[t] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = block_off[blockIdx.x] + __popc(m & ((1u << lane) - 1));
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (!f) return;
    const float4 p = ct2[t];
    const int sd = ct_seed[t];                                  // the classification match: exact NN of iteration 0
    const float4 q = __ldg(tgt_pts + sd), nq = __ldg(tgt_aux + sd);
    sorted[before] = make_float4(p.x, p.y, p.z, __int_as_float(before));
    cq[before] = make_float4(q.x, q.y, q.z, __int_as_float(sd));
    cn[before] = make_float4(nq.x, nq.y, nq.z, nq.x * q.x + nq.y * q.y + nq.z * q.z);
}

// ---- k-th smallest of n non-negative floats: three histogram passes over the values ------------------------------
// kPass 0: bits 31..21, 1: bits 20..10, 2: bits 9..0 of the float pattern (non-negative floats order like unsigned
// integers).  The last block to add its histogram picks the bin that holds the wanted rank and resets the state for
// the next pass.  gate: when not null and zero, the pass is skipped (stage 1 of the DT schedule is over).
template <int kPass>
__global__ void __launch_bounds__(256)
select_pass_kernel(const float* __restrict__ v, int n, SelectDev* s, const int* gate) {
    if (gate && *gate == 0) return;
    constexpr int kShift = kPass == 0 ? 21 : kPass == 1 ? 10 : 0;
    constexpr int kBins = kPass == 2 ? 1024 : 2048;
    constexpr unsigned kHi = kPass == 0 ? 0u : kPass == 1 ? 0xffe00000u : 0xfffffc00u;
    __shared__ unsigned sh[2048];
    __shared__ int s_last;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned prefix = s->prefix;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned u = __float_as_uint(v[i]);
        if ((u & kHi) == prefix) atomicAdd(&sh[(u >> kShift) & (kBins - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) if (sh[i]) atomicAdd(&s->hist[i], sh[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&s->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // inclusive scan of the bins (8 per thread), then the first bin whose running count exceeds the rank
    constexpr int kPer = 2048 / 256;
    unsigned loc[kPer], run = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) { const int b = threadIdx.x * kPer + k; loc[k] = (b < kBins) ? __ldcg(&s->hist[b]) : 0u; run += loc[k]; }
    sh[threadIdx.x] = run;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const unsigned y = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += y;
        __syncthreads();
    }
    unsigned long long before = (unsigned long long)(sh[threadIdx.x] - run);
    const unsigned long long rank = s->rank;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if (rank >= before && rank < before + loc[k]) {          // exactly one thread, one bin
            s->prefix = prefix | ((unsigned)(threadIdx.x * kPer + k) << kShift);
            s->rank = rank - before;
        }
        before += loc[k];
    }
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s->hist[i] = 0;
    if (threadIdx.x == 0) s->ticket = 0;
}

static void select_enqueue(Ctx* ctx, const float* v, int n, SelectDev* s, const int* gate) {
    const int blocks = std::max(1, std::min((n + 255) / 256, ctx->num_sms * 8));
    select_pass_kernel<0><<<blocks, 256, 0, ctx->stream>>>(v, n, s, gate);
    select_pass_kernel<1><<<blocks, 256, 0, ctx->stream>>>(v, n, s, gate);
    select_pass_kernel<2><<<blocks, 256, 0, ctx->stream>>>(v, n, s, gate);
    ctx->launches += 3;
}

// ---- (7) stage-1 percentile: NN distances of (flagged) points against a full-cloud grid ------
// calPercentileDistBetween2PC, src/CommonFunc.cpp:266-281.  Unflagged points get +inf so that
// they rank behind every valid distance.  gate: see select_pass_kernel.
__global__ void __launch_bounds__(256)
percentile_d2_kernel(GridDev g, const float* __restrict__ q, int nq, const int* __restrict__ patch_id,
                     const int* __restrict__ flags, float* __restrict__ d2, int* __restrict__ seeds,
                     const int* gate) {
    if (gate && *gate == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inrange = i < nq;
    const bool active = inrange && !(flags && !flags[patch_id[i]]);
    float px = 0.f, py = 0.f, pz = 0.f;
    if (active) { px = q[3 * (size_t)i]; py = q[3 * (size_t)i + 1]; pz = q[3 * (size_t)i + 2]; }
    const int seed = (active && seeds) ? seeds[i] : -1;
    float dist2 = __int_as_float(0x7f800000);
    const Best b = nn_search_warp(g, px, py, pz, seed, active);      // warp-collective (nn_search.cuh, team search)
    if (active) {
        if (seeds) seeds[i] = b.pos;
        dist2 = b.d2;
    }
    if (inrange) d2[i] = dist2;
}

// stand-alone percentile (DTinit, pwicp_percentile_nn): every query counts, the rank is known on the host
int percentile_dev(Ctx* ctx, const GridDev& g, const float* q, int nq, const int* patch_id,
                   const int* flags, long long n_valid, float pct, double* out, int* seeds) {
    if (nq < 1 || n_valid < 1) { set_error(ctx, "percentile: empty query set"); return PWICP_ERR_ARG; }
    PW_TRY(ctx->scratch_a.reserve(ctx, (size_t)nq * 4));
    PW_TRY(ctx->outer_state.reserve(ctx, sizeof(OuterDev)));
    float* d2 = ctx->scratch_a.as<float>();
    SelectDev* sel = &ctx->outer_state.as<OuterDev>()->sel;
    percentile_d2_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(g, q, nq, patch_id, flags, d2, seeds, nullptr);
    ctx->launches++;
    int leftnum = (int)((float)n_valid * pct);           // int leftnum = n * percentile (:177)
    if (leftnum >= n_valid) leftnum = (int)n_valid - 1;
    PW_CUDA(cudaMemsetAsync(sel, 0, sizeof(SelectDev), ctx->stream));
    const unsigned long long rank = (unsigned long long)leftnum;
    PW_CUDA(cudaMemcpyAsync(&sel->rank, &rank, sizeof(rank), cudaMemcpyHostToDevice, ctx->stream));
    select_enqueue(ctx, d2, nq, sel, nullptr);
    unsigned bits = 0;
    PW_CUDA(cudaMemcpyAsync(&bits, &sel->prefix, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    float v; memcpy(&v, &bits, 4);
    *out = (double)sqrtf(v);                             // distArray[i] = sqrt(float) (:277)
    return PWICP_OK;
}

// ---- (8) transforms, src/Registration.cpp:942-954 -------------------------------------------
struct Mat34 { float m[12]; };

// packed xyz: each thread moves 4 points = three 16-byte vectors (coalesced 128-bit accesses)
__device__ __forceinline__ void transform_packed_body(float* __restrict__ xyz, size_t n, const float* T, size_t first, size_t stride) {
    const size_t nquad = n / 4;
    float4* v = reinterpret_cast<float4*>(xyz);
    for (size_t qd = first; qd < nquad; qd += stride) {
        float4 a = v[3 * qd], b = v[3 * qd + 1], c = v[3 * qd + 2];
        float o[12];
        xform_point(T, a.x, a.y, a.z, o[0], o[1], o[2]);
        xform_point(T, a.w, b.x, b.y, o[3], o[4], o[5]);
        xform_point(T, b.z, b.w, c.x, o[6], o[7], o[8]);
        xform_point(T, c.y, c.z, c.w, o[9], o[10], o[11]);
        v[3 * qd] = make_float4(o[0], o[1], o[2], o[3]);
        v[3 * qd + 1] = make_float4(o[4], o[5], o[6], o[7]);
        v[3 * qd + 2] = make_float4(o[8], o[9], o[10], o[11]);
    }
    if (first < (n & 3)) {
        const size_t i = nquad * 4 + first;
        float x, y, z;
        xform_point(T, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], x, y, z);
        xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
    }
}

__global__ void __launch_bounds__(256)
transform_packed_kernel(float* __restrict__ xyz, size_t n, Mat34 T) {
    transform_packed_body(xyz, n, T.m, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

// the four transforms of an outer iteration in one launch, the matrix read from the device state: blocks
// [0, b0) cloud2, [b0, b1) patch points, then CT2 and BP2 (float4)
__global__ void __launch_bounds__(256)
transform_all_kernel(const OuterDev* st, float* __restrict__ cloud2, size_t m2, float* __restrict__ patch_xyz, size_t mp2,
                     float4* __restrict__ ct2, int n2, float4* __restrict__ bp2, int b0, int b1) {
    __shared__ float T[12];
    if (threadIdx.x < 12) T[threadIdx.x] = st->T[threadIdx.x];
    __syncthreads();
    const int b = blockIdx.x;
    if (b < b0) { transform_packed_body(cloud2, m2, T, (size_t)b * blockDim.x + threadIdx.x, (size_t)b0 * blockDim.x); return; }
    if (b < b1) { transform_packed_body(patch_xyz, mp2, T, (size_t)(b - b0) * blockDim.x + threadIdx.x, (size_t)(b1 - b0) * blockDim.x); return; }
    const int nb2 = (int)gridDim.x - b1;
    for (int i = (b - b1) * blockDim.x + threadIdx.x; i < 7 * n2; i += nb2 * blockDim.x) {
        float4* p = (i < n2) ? ct2 + i : bp2 + (i - n2);
        float4 v = *p;
        float x, y, z;
        xform_point(T, v.x, v.y, v.z, x, y, z);
        *p = make_float4(x, y, z, v.w);
    }
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

// ---- (6) octree bounding cube, src/Registration.cpp:881-886 ---------------------------------
// pcl::octree::OctreePointCloud::defineBoundingBox() + getKeyBitSize() (PCL 1.8.1)
PW_HD void octree_cube_hd(const float* mn, const float* mx, double res, double* bb) {
    const float minValue512 = FLT_EPSILON * 512.0f;
    const float minValue = FLT_EPSILON;
    double lo[3], hi[3];
    for (int c = 0; c < 3; ++c) { lo[c] = mn[c]; hi[c] = (float)(mx[c] + minValue512); }
    unsigned int key[3];
    for (int c = 0; c < 3; ++c) key[c] = (unsigned int)ceil((hi[c] - lo[c] - minValue) / res);
    unsigned int max_voxels = key[0] > key[1] ? key[0] : key[1];
    if (key[2] > max_voxels) max_voxels = key[2];
    if (max_voxels < 2u) max_voxels = 2u;
    unsigned int depth = (unsigned int)ceil(log((double)max_voxels) / log(2.0) - minValue);
    if (depth > 32u) depth = 32u;
    const double side = (double)(1ull << depth) * res;
    for (int c = 0; c < 3; ++c) {
        const double over = (side - (hi[c] - lo[c])) / 2.0;
        if (over > minValue) { lo[c] -= over; hi[c] += over; }
    }
    bb[0] = lo[0]; bb[1] = lo[1]; bb[2] = lo[2]; bb[3] = hi[0]; bb[4] = hi[1]; bb[5] = hi[2];
}

void octree_cube(const float* mn, const float* mx, double res, double* bb) { octree_cube_hd(mn, mx, res, bb); }

// ---- the DT schedule on the device, src/Registration.cpp:880-935 ---------------------------------------------
PW_HD float bbox_corner_change_hd(const double* bb, const float* T) {
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
        const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        best = (k == 0) ? nrm : (nrm > best ? nrm : best);
    }
    return best;
}
float bbox_corner_change_host(const double* bb, const float* T) { return bbox_corner_change_hd(bb, T); }

__device__ __forceinline__ float ord2f_dev(int i) { return __int_as_float((i >= 0) ? i : i ^ 0x7fffffff); }

// after the inner loop: LoDetection range, bounding cube of the CURRENT cloud2, corner change, first half of the
// stage logic (:891-893) and the rank the stage-1 percentile has to deliver
__global__ void outer_state1_kernel(OuterDev* st, const float* __restrict__ icp_T, const int* __restrict__ icp_state) {
    if (threadIdx.x) return;
    st->LoDet_min = ord2f_dev(st->lod_min); st->LoDet_max = ord2f_dev(st->lod_max);          // :768-769
    st->icp_iters = icp_state[0]; st->icp_state = icp_state[1];
    for (int k = 0; k < 16; ++k) st->T[k] = icp_T[k];
    st->need_p75 = 0;
    if (st->n_stable < 4) {                                                                 // :864-867
        st->err = PWICP_ERR_TOO_FEW_STABLE;
        for (int k = 0; k < 16; ++k) st->T[k] = (k % 5 == 0) ? 1.0f : 0.0f;
        return;
    }
    float mn[3], mx[3];
    for (int c = 0; c < 3; ++c) { mn[c] = ord2f_dev(st->bbox[c]); mx[c] = ord2f_dev(st->bbox[3 + c]); }
    octree_cube_hd(mn, mx, st->res2, st->bb6);                                              // :881-886
    const float maxBB = bbox_corner_change_hd(st->bb6, st->T);
    st->maxBBchange = maxBB;
    if (!st->toStage2 && maxBB < st->minLoD) st->toStage2 = 1;                              // :891-893
    else if (st->currDT == st->LoDet_min) st->toStage3 = 1;
    if (!st->toStage2) {
        const long long n_valid = (long long)st->n_stable_pts;
        if (n_valid < 1) { st->err = PWICP_ERR_ARG; return; }
        int leftnum = (int)((float)n_valid * 0.75f);                                        // int leftnum = n * percentile (:177)
        if (leftnum >= n_valid) leftnum = (int)n_valid - 1;
        st->sel.prefix = 0; st->sel.ticket = 0; st->sel.rank = (unsigned long long)leftnum;
        st->need_p75 = 1;
    }
}

// second half of the stage logic, :896-935 (the stage-1 block may fall through into the stage-2 block)
__global__ void outer_state2_kernel(OuterDev* st) {
    if (threadIdx.x || st->err) return;
    float currDT = st->currDT;
    const float LoDet_min = st->LoDet_min, maxBBchange = st->maxBBchange;
    if (!st->toStage2) {
        const double Dist75 = (double)sqrtf(__uint_as_float(st->sel.prefix));               // distArray[i] = sqrt(float) (:277)
        st->P75 = Dist75; st->ran_p75 = 1;
        if (currDT > Dist75) currDT = Dist75;
        else st->toStage2 = 1;
        if (currDT <= LoDet_min) currDT = LoDet_min;
        st->BBchange_2 = st->BBchange_1;
        st->BBchange_1 = maxBBchange;
    }
    if (st->toStage2 && !st->toStage3) {
        const float upperBound = 0.8f, lowerBound = 0.5f;
        const float alpha = fabsf(st->BBchange_1 / st->BBchange_2);
        if (isnan(alpha) || isinf(alpha)) currDT = currDT * upperBound;
        else if (alpha < lowerBound) currDT = currDT * lowerBound;
        else if (alpha > upperBound) currDT = currDT * upperBound;
        else currDT = currDT * alpha;
        if (currDT <= LoDet_min) currDT = LoDet_min;
        st->BBchange_2 = st->BBchange_1;
        st->BBchange_1 = maxBBchange;
    }
    st->currDT = currDT;
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
           const double* __restrict__ X, double* __restrict__ partials, int* __restrict__ seeds,
           const float4* __restrict__ cq_seed) {
    __shared__ double sm[8][27];
    double acc[27];
#pragma unroll
    for (int v = 0; v < 27; ++v) acc[v] = 0.0;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool active = i < n;
        const float4 q = active ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (!active) continue;
        // seed: the caller's list, or the last match the inner loop left for this point (level-0 position in cq[].w)
        const int seed = seeds ? seeds[i] : (cq_seed ? (__float_as_int(cq_seed[i].w) & 0x3fffffff) : -1);
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

int vcm_dev(Ctx* ctx, const float4* src, int n, double* vcm36, int* singular, int* seeds, const float4* cq_seed) {
    // n <= 6: the reference divides by (n - 6) all the same (:1331) and hands back an infinite / NaN matrix; so does this
    if (n < 1) { set_error(ctx, "vcm: empty source"); return PWICP_ERR_ARG; }
    const int blocks = std::min(kVcmBlocks, (n + 255) / 256);
    PW_TRY(ctx->scratch_c.reserve(ctx, (size_t)kVcmBlocks * 27 * 8 + 64));
    double* part = ctx->scratch_c.as<double>();
    double* Xd = part + (size_t)kVcmBlocks * 27;
    std::vector<double> hp((size_t)blocks * 27);
    const size_t smem = 0;
    vcm_kernel<<<blocks, 256, smem, ctx->stream>>>(ctx->tgt.dev, ctx->tgt_aux.as<float4>(), src, n, 0, nullptr, part, seeds, cq_seed);
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
    vcm_kernel<<<blocks, 256, smem, ctx->stream>>>(ctx->tgt.dev, ctx->tgt_aux.as<float4>(), src, n, 1, Xd, part, seeds, cq_seed);
    ctx->launches++;
    PW_CUDA(cudaMemcpyAsync(hp.data(), part, hp.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    double vtpv = 0;
    for (int b = 0; b < blocks; ++b) vtpv += hp[(size_t)b * 27];
    const double STD0 = 1 * std::sqrt(vtpv / double(n - 6));
    for (int k = 0; k < 36; ++k) vcm36[k] = STD0 * STD0 * Q[k];
    return PWICP_OK;
}

// ---- processing order of the source patches (once per pair) --------------------------------

// Processing order of the classification queries: patches in the spatial order of their centroid (spatial_order_dev,
// grid.cu).  Results are written back in the caller's order, so the order only shapes the warps of the search.
static int ensure_patch_order(Ctx* ctx) {
    if (ctx->ct_order_valid) return PWICP_OK;
    const int n = ctx->n2;
    PW_TRY(ctx->ct_order.reserve(ctx, (size_t)n * 4));
    PW_TRY(spatial_order_dev(ctx, ctx->tgt.dev, ctx->ct2.as<float4>(), n, ctx->ct_order.as<uint32_t>()));
    ctx->ct_order_valid = true;
    return PWICP_OK;
}

// ---- the outer iteration -------------------------------------------------------------------
int outer_single_iteration(Ctx* ctx, const pwicp_pair_params& pp, pwicp_state* st,
                           const pwicp_icp_params& icp, float* T16, double* vcm36,
                           unsigned char* stable_flags, pwicp_iter_stats* stats) {
    const int n2 = ctx->n2;
    if (!ctx->tgt.dev.nlevels || n2 < 1 || !ctx->c1.dev.nlevels || ctx->m2 < 1) {
        set_error(ctx, "single_iteration: target, source and clouds must be uploaded first");
        return PWICP_ERR_ARG;
    }
    const float DTmin = pp.DTmin;
    if (st->currDT <= DTmin) st->currDT = DTmin;                           // :724-725
    if (4 > n2) { set_error(ctx, "No enough stable points left (<4)"); return PWICP_ERR_TOO_FEW_PATCHES; }

    const float max2minLoD = 2.0f;
    const float maxLoD = DTmin * max2minLoD, minLoD = DTmin;               // :751-753
    // scratch: pl[7*n2], pt2pt[n2], lod[n2] | flags[n2] + bytes[n2] | block counts | state | pinned mirror
    const int nblocks = (n2 + kCompactBlock - 1) / kCompactBlock;
    const int n2_pad = (n2 + 31) / 32 * 32;
    PW_TRY(ctx->scratch_a.reserve(ctx, (size_t)9 * n2 * 4));
    PW_TRY(ctx->flags.reserve(ctx, (size_t)n2 * 4 + (size_t)n2));
    PW_TRY(ctx->pos.reserve(ctx, (size_t)nblocks * 4));
    PW_TRY(ctx->outer_state.reserve(ctx, sizeof(OuterDev)));
    PW_TRY(ctx->icp_sorted.reserve(ctx, (size_t)n2_pad * sizeof(float4)));
    PW_TRY(ctx->icp_match.reserve(ctx, (size_t)n2_pad * 3 * sizeof(float4)));
    if (ctx->pinned_cap < 16384) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr; ctx->pinned_cap = 0;
        PW_CUDA(cudaMallocHost(&ctx->pinned, 16384));
        ctx->pinned_cap = 16384;
    }
    if (!ctx->outer_state_zeroed) {                    // the select histogram is kept zero by its last pass
        PW_CUDA(cudaMemsetAsync(ctx->outer_state.p, 0, sizeof(OuterDev), ctx->stream));
        ctx->outer_state_zeroed = true;
    }
    float* pl = ctx->scratch_a.as<float>();
    float* pt2pt = pl + (size_t)7 * n2;
    float* lod = pt2pt + n2;
    int* flags = ctx->flags.as<int>();
    unsigned char* flags_u8 = reinterpret_cast<unsigned char*>(flags + n2);
    int* block_cnt = ctx->pos.as<int>();
    OuterDev* sd = ctx->outer_state.as<OuterDev>();
    OuterDev* h_in = reinterpret_cast<OuterDev*>(ctx->pinned);                                   // host -> device
    OuterDev* h_out = reinterpret_cast<OuterDev*>(reinterpret_cast<char*>(ctx->pinned) + 8192);  // device -> host
    float4* sorted = ctx->icp_sorted.as<float4>();
    float4* cn = ctx->icp_match.as<float4>();
    float4* cq = cn + n2_pad;

    memset(h_in, 0, kOuterInitBytes);
    h_in->currDT = st->currDT; h_in->BBchange_1 = st->BBchange_1; h_in->BBchange_2 = st->BBchange_2;
    h_in->toStage2 = st->toStage2; h_in->toStage3 = st->toStage3;
    h_in->minLoD = minLoD;
    h_in->res2 = (double)(float)(pp.Res2 * 2);
    h_in->lod_min = 0x7fffffff; h_in->lod_max = (int)0x80000000;
    for (int c = 0; c < 3; ++c) { h_in->bbox[c] = 0x7fffffff; h_in->bbox[3 + c] = (int)0x80000000; }
    PW_CUDA(cudaEventRecord(ctx->ev_o0, ctx->stream));
    PW_CUDA(cudaMemcpyAsync(sd, h_in, kOuterInitBytes, cudaMemcpyHostToDevice, ctx->stream));

    PW_TRY(ensure_patch_order(ctx));
    // (1)-(3) distances of the 7 queries of every patch, LoDetection
    classify_dist_kernel<<<(7 * n2 + 255) / 256, 256, 0, ctx->stream>>>(
        ctx->tgt.dev, ctx->tgt_aux.as<float4>(), ctx->tgt_ok.as<unsigned char>(), ctx->ct2.as<float4>(),
        ctx->bp2.as<float4>(), ctx->bpstd2.as<float>(), n2, minLoD, maxLoD, pl, pt2pt, lod, &sd->lod_min,
        ctx->ct_order.as<uint32_t>(), ctx->ct_seed.as<int>(), ctx->bp_seed.as<int>());
    // (4) classification + compaction of the stable set straight into the inner loop's layout
    const float DTctct = st->currDT + 1 * (pp.SVRes1 + pp.SVRes2);          // :817
    classify_flag_kernel<<<nblocks, kCompactBlock, 0, ctx->stream>>>(
        pl, pt2pt, lod, ctx->patch_off.as<int>(), ctx->ct_order.as<uint32_t>(), n2, st->currDT, DTctct, flags, flags_u8,
        block_cnt, sd);
    scan_blocks_kernel<<<1, 1024, 0, ctx->stream>>>(block_cnt, nblocks, sd, sorted, cn, cq);
    compact_kernel<<<nblocks, kCompactBlock, 0, ctx->stream>>>(
        ctx->ct2.as<float4>(), flags, ctx->ct_order.as<uint32_t>(), block_cnt, n2, ctx->ct_seed.as<int>(),
        ctx->tgt.dev.lv[0].pts, ctx->tgt_aux.as<float4>(), sorted, cn, cq);
    ctx->launches += 4;
    // (5) inner ICP on the stable centroids against ALL target centroids, :877 -- the count stays on the device
    IcpLaunch L;
    PW_TRY(icp_enqueue(ctx, icp, n2, &sd->n_stable, true, false, false, false, &L));
    // (6) bounding cube of the CURRENT cloud2, :880-888
    PW_TRY(bbox_accumulate_dev(ctx, ctx->cloud2.as<float>(), (size_t)ctx->m2, sd->bbox));
    outer_state1_kernel<<<1, 32, 0, ctx->stream>>>(sd, reinterpret_cast<const float*>(L.out), reinterpret_cast<const int*>(L.out + 64));
    ctx->launches++;
    // (7) stage 1: P75 of the stable patches' points against cloud1, :905.  toStage2 never goes back to 0, so once the
    // host has seen it the kernels are not even launched; in the iteration that leaves stage 1 they return at once
    if (!st->toStage2) {
        PW_TRY(ctx->scratch_b.reserve(ctx, (size_t)std::max(ctx->mp2, 1) * 4));
        float* d2 = ctx->scratch_b.as<float>();
        percentile_d2_kernel<<<(ctx->mp2 + 255) / 256, 256, 0, ctx->stream>>>(
            ctx->c1.dev, ctx->patch_xyz.as<float>(), ctx->mp2, ctx->patch_id.as<int>(), flags, d2, ctx->pp_seed.as<int>(),
            &sd->need_p75);
        ctx->launches++;
        select_enqueue(ctx, d2, ctx->mp2, &sd->sel, &sd->need_p75);
    }
    outer_state2_kernel<<<1, 32, 0, ctx->stream>>>(sd);
    // (8) apply the transform to cloud2, CT2, BP2 and every patch, :942-954
    {
        const int cap = ctx->num_sms * 8;
        const int b0 = (int)std::min<size_t>(std::max<size_t>(((size_t)ctx->m2 / 4 + 255) / 256, 1), (size_t)cap);
        const int bp = (int)std::min<size_t>(std::max<size_t>(((size_t)ctx->mp2 / 4 + 255) / 256, 1), (size_t)cap);
        const int bc = std::min(std::max((7 * n2 + 255) / 256, 1), cap);
        transform_all_kernel<<<b0 + bp + bc, 256, 0, ctx->stream>>>(sd, ctx->cloud2.as<float>(), (size_t)ctx->m2,
                                                                   ctx->patch_xyz.as<float>(), (size_t)ctx->mp2,
                                                                   ctx->ct2.as<float4>(), n2, ctx->bp2.as<float4>(), b0, b0 + bp);
    }
    ctx->launches += 2;
    PW_CUDA(cudaMemcpyAsync(h_out, sd, kOuterReadBytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (stable_flags) PW_CUDA(cudaMemcpyAsync(stable_flags, flags_u8, n2, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaEventRecord(ctx->ev_o1, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));                            // the one synchronisation of the iteration

    const int nStable = h_out->n_stable;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_stable = nStable; stats->n_stable_pts = (int)h_out->n_stable_pts;
        stats->LoDet_min = h_out->LoDet_min; stats->LoDet_max = h_out->LoDet_max;
        stats->P75 = h_out->ran_p75 ? h_out->P75 : std::numeric_limits<double>::quiet_NaN();
        stats->icp_iters = h_out->icp_iters; stats->icp_state = h_out->icp_state; stats->maxBBchange = h_out->maxBBchange;
        memcpy(stats->bb6, h_out->bb6, sizeof(h_out->bb6));
    }
    if (h_out->err == PWICP_ERR_TOO_FEW_STABLE) {                           // :864-867
        set_error(ctx, "No enough stable points left, no enough overlapping areas");
        return PWICP_ERR_TOO_FEW_STABLE;
    }
    if (h_out->err) { set_error(ctx, "single_iteration: stage-1 percentile over an empty point set"); return h_out->err; }
    ctx->n_icp = nStable;
    ctx->icp_prof_iters = h_out->icp_iters;
    st->currDT = h_out->currDT; st->BBchange_1 = h_out->BBchange_1; st->BBchange_2 = h_out->BBchange_2;
    st->toStage2 = h_out->toStage2; st->toStage3 = h_out->toStage3;
    memcpy(T16, h_out->T, 16 * sizeof(float));

    // (9) VCM from the pre-update stable centroids (the inner loop never modifies its source array), :957-961
    float ms = 0.f;
    if (st->toStage3 && vcm36) {
        int sing = 0;
        PW_TRY(vcm_dev(ctx, sorted, nStable, vcm36, &sing, nullptr, cq));
        if (stats) { stats->vcm_written = 1; stats->vcm_singular = sing; }
        PW_CUDA(cudaEventRecord(ctx->ev_o1, ctx->stream));
        PW_CUDA(cudaEventSynchronize(ctx->ev_o1));
    }
    cudaEventElapsedTime(&ms, ctx->ev_o0, ctx->ev_o1);
    ctx->last_ms = ms;
    if (stats) stats->device_ms = ms;
    return PWICP_OK;
}

}  // namespace pwicp
