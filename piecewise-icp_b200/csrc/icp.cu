// icp.cu -- the whole inner point-to-plane ICP loop as ONE persistent cooperative kernel.
//
// Replaces P2PICPwithPatchNormal (reference src/Registration.cpp:1255-1269), i.e.
// pcl::IterativeClosestPointWithNormals::align with TransformationEstimationPointToPlaneLLS and
// DefaultConvergenceCriteria.
//
// Per inner iteration every thread: (a) applies the previous incremental transform to its
// source point in place (float, pcl::transformPointCloudWithNormals order), (b) finds the exact
// nearest target centroid in the grid, (c) forms the 7 float row terms of the LLS system; each
// warp accumulates the 27 (+1: sum of squared NN distances) double sums of its 32-point batch
// sequentially through shared memory, CTAs publish partials, ONE grid-wide barrier, then every
// CTA redundantly reduces the partials in a fixed order, solves the 6x6 system, builds the float
// transform and evaluates the convergence criteria -- identical instructions on identical
// inputs, so all CTAs take the same decision without a second barrier or a host round trip.
//
// The summation order ("reduction geometry") is deterministic and is reproduced by the
// oracle's reduce_mode=1 for the bit-exact whole-loop parity test.
#include <cooperative_groups.h>

#include "common.cuh"
#include <cub/cub.cuh>

#include "nn_search.cuh"
#include "small_algebra.cuh"

namespace cg = cooperative_groups;

namespace pwicp {

struct IcpArgs {
    GridDev g;
    const float4* aux;        // level-0 order: nx, ny, nz, ctstd
    const float4* src;        // source set (read only)
    float4* work;             // transformed copy, updated in place every iteration
    int* match;               // per source point: level-0 position of its last match (seed of the next search)
    int use_seed0;            // match[] already holds seeds for the first iteration
    int n;
    int max_iter;
    int force_iters;
    double rot_thr, transl_thr, mse_rel, mse_abs;
    double* partials;         // [2][gridDim.x][28]
    float* out_T;             // 16: final transformation
    int* out_state;           // [0] n_iter, [1] conv_state
    double* mse_trace;        // nullable
    float* T_trace;           // nullable
    int* idx_trace;           // nullable, [iter][n]
};

// Thread 0 of every CTA: totals -> 6x6 solve -> float transform -> convergence decision.
// Kept out of line so its local arrays do not inflate the register budget of the search loop.
__device__ __noinline__ int icp_finish(const IcpArgs& a, int it, const double* s_tot, float* s_T,
                                       float* s_Tfinal, double& prev_mse) {
    double tot[kNumVals], x[6];
    for (int v = 0; v < kNumVals; ++v) tot[v] = s_tot[v];
    float Tn[16];
    solve_from28(tot, x, Tn);
    for (int k = 0; k < 16; ++k) s_T[k] = Tn[k];
    float Tf[16];
    mat4_mul(Tn, s_Tfinal, Tf);                       // final = T * final
    for (int k = 0; k < 16; ++k) s_Tfinal[k] = Tf[k];
    const double mse = tot[27] / (double)a.n;
    const int iters = it + 1;
    int state = 0;
    // DefaultConvergenceCriteria<float>::hasConverged(), in PCL's order
    if (iters >= a.max_iter) state = PWICP_CONV_ITERATIONS;
    else if (!a.force_iters) {
        const double cos_angle = 0.5 * (double)(Tn[0] + Tn[5] + Tn[10] - 1.0f);
        const double transl_sq = (double)(Tn[3] * Tn[3] + Tn[7] * Tn[7] + Tn[11] * Tn[11]);
        if (cos_angle >= a.rot_thr && transl_sq <= a.transl_thr) state = PWICP_CONV_TRANSFORM;
        else if (fabs(mse - prev_mse) < a.mse_abs) state = PWICP_CONV_ABS_MSE;
        else if (fabs(mse - prev_mse) / prev_mse < a.mse_rel) state = PWICP_CONV_REL_MSE;
    }
    prev_mse = mse;
    if (blockIdx.x == 0) {
        if (a.mse_trace) a.mse_trace[it] = mse;
        if (a.T_trace) for (int k = 0; k < 16; ++k) a.T_trace[(size_t)it * 16 + k] = Tn[k];
        if (state) {
            for (int k = 0; k < 16; ++k) a.out_T[k] = Tf[k];
            a.out_state[0] = iters;
            a.out_state[1] = state;
        }
    }
    return state;
}

__global__ void __launch_bounds__(kIcpThreads, 3) icp_persistent_kernel(const IcpArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) float s_rows[kIcpWarps][32][8];
    __shared__ double s_wacc[kIcpWarps][kNumVals];
    __shared__ double s_tot[kNumVals];
    __shared__ float s_T[16];
    __shared__ float s_Tfinal[16];
    __shared__ int s_stop;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const long NW = (long)G * kIcpWarps;
    const long gw = (long)blockIdx.x * kIcpWarps + warp;
    const long nb = ((long)a.n + 31) / 32;

    // which pair of row terms this lane accumulates: 21 upper-triangle ATA entries (row-major),
    // 6 ATb entries (u_r * u_6), lane 27 = sum of squared NN distances
    int va = 0, vb = 0;
    {
        int v = 0;
        for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c) { if (v == lane) { va = r; vb = c; } ++v; }
        for (int r = 0; r < 6; ++r) { if (v == lane) { va = r; vb = 6; } ++v; }
        if (lane == 27) { va = 7; vb = 7; }
    }
    if (tid < 16) { s_T[tid] = (tid % 5 == 0) ? 1.0f : 0.0f; s_Tfinal[tid] = s_T[tid]; }
    if (tid == 0) s_stop = 0;
    double prev_mse = 1.7976931348623157e308;   // DBL_MAX
    __syncthreads();

    for (int it = 0;; ++it) {
        double acc = 0.0;
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = s_T[k];

        for (long b = gw; b < nb; b += NW) {
            const long i = b * 32 + lane;
            const bool active = i < a.n;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
            float4 p = lo;
            if (active) {
                p = (it == 0) ? __ldg(a.src + i) : a.work[i];
                if (it > 0) {
                    float x, y, z;
                    xform_point(T, p.x, p.y, p.z, x, y, z);
                    p.x = x; p.y = y; p.z = z;
                }
                a.work[i] = p;
            }
            int seed = -1;
            if (active && (it > 0 || a.use_seed0)) seed = a.match[i];
            if (active) {
                const Best bb = nn_search_seeded<true>(a.g, p.x, p.y, p.z, seed);
                a.match[i] = bb.pos;
                const float4 nq = __ldg(a.aux + bb.pos);
                const float sx = p.x, sy = p.y, sz = p.z;
                const float dx = bb.qx, dy = bb.qy, dz = bb.qz;
                const float nx = nq.x, ny = nq.y, nz = nq.z;
                // float expressions of TransformationEstimationPointToPlaneLLS (no FMA)
                lo.x = nz * sy - ny * sz;
                lo.y = nx * sz - nz * sx;
                lo.z = ny * sx - nx * sy;
                lo.w = nx;
                hi.x = ny;
                hi.y = nz;
                hi.z = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz;
                hi.w = bb.d2;
                // traces are reported in the caller's order (p.w = original source index)
                if (a.idx_trace) a.idx_trace[(size_t)it * a.n + __float_as_int(p.w)] = bb.idx;
            }
            float4* row = reinterpret_cast<float4*>(&s_rows[warp][lane][0]);
            row[0] = lo; row[1] = hi;
            __syncwarp();
            if (lane < kNumVals) {
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const double x = (double)s_rows[warp][r][va];
                    const double y = (lane == 27) ? 1.0 : (double)s_rows[warp][r][vb];
                    acc += x * y;      // exact product of two float values, then one rounding
                }
            }
            __syncwarp();
        }
        if (lane < kNumVals) s_wacc[warp][lane] = acc;
        __syncthreads();
        if (tid < kNumVals) {
            double s = s_wacc[0][tid];
#pragma unroll
            for (int w = 1; w < kIcpWarps; ++w) s += s_wacc[w][tid];
            a.partials[((size_t)(it & 1) * G + blockIdx.x) * kNumVals + tid] = s;
        }
        grid.sync();

        // fixed-order reduction of the CTA partials (every CTA computes the same totals)
        const double* P = a.partials + (size_t)(it & 1) * G * kNumVals;
        for (int v = warp; v < kNumVals; v += kIcpWarps) {
            double s = 0.0;
            for (int b = lane; b < G; b += 32) s += __ldcg(P + (size_t)b * kNumVals + v);
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) s_tot[v] = s;
        }
        __syncthreads();

        if (tid == 0) s_stop = icp_finish(a, it, s_tot, s_T, s_Tfinal, prev_mse);
        __syncthreads();
        if (s_stop) break;
    }
}

__global__ void expand_xyz_kernel(const float* __restrict__ xyz, int n, float4* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.f);
}

int icp_expand_source(Ctx* ctx, const float* packed_dev, int n) {
    PW_TRY(ctx->icp_src.reserve(ctx, (size_t)n * sizeof(float4)));
    expand_xyz_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(packed_dev, n, ctx->icp_src.as<float4>());
    ctx->launches++;
    ctx->n_icp = n;
    return PWICP_OK;
}

// key = Morton code of the finest-level cell of the source point (same cell expression as the
// grid build), so that consecutive points of the sorted order sit in a compact 3-D neighbourhood
__device__ __forceinline__ unsigned long long spread21(unsigned int v) {
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void src_key_kernel(const float4* __restrict__ src, int n, float ox, float oy, float oz, float inv_h,
                               int dx, int dy, int dz, unsigned long long* keys, uint32_t* vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = src[i];
    float fx = (p.x - ox) * inv_h, fy = (p.y - oy) * inv_h, fz = (p.z - oz) * inv_h;
    int cx = min(max((int)floorf(fx), 0), dx - 1);
    int cy = min(max((int)floorf(fy), 0), dy - 1);
    int cz = min(max((int)floorf(fz), 0), dz - 1);
    keys[i] = spread21(cx) | (spread21(cy) << 1) | (spread21(cz) << 2);
    vals[i] = (uint32_t)i;
}

__global__ void src_gather_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ order, int n, float4* out,
                                  const int* __restrict__ seed_in, int* __restrict__ seed_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = order[i];
    float4 p = src[o];
    p.w = __int_as_float((int)o);
    out[i] = p;
    if (seed_in) seed_out[i] = seed_in[o];
}

// Sorts the source set by the target-grid cell it starts in (stable: ties keep the caller's
// order), so that the 8 queries of a tile group share a small candidate block.  The processing
// order only affects the order of the double sums (DESIGN.md "reduction geometry").
static int icp_sort_source(Ctx* ctx, int n, bool have_seed) {
    const GridLevel& L = ctx->tgt.dev.lv[0];
    PW_TRY(ctx->keys.reserve(ctx, (size_t)n * 8));
    PW_TRY(ctx->vals.reserve(ctx, (size_t)n * 4));
    PW_TRY(ctx->keys2.reserve(ctx, (size_t)n * 8));
    PW_TRY(ctx->icp_perm.reserve(ctx, (size_t)n * 4));
    PW_TRY(ctx->icp_sorted.reserve(ctx, (size_t)n * sizeof(float4)));
    PW_TRY(ctx->icp_match.reserve(ctx, (size_t)n * sizeof(int)));
    const int blocks = (n + 255) / 256;
    src_key_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->icp_src.as<float4>(), n, ctx->tgt.dev.ox, ctx->tgt.dev.oy,
                                                    ctx->tgt.dev.oz, L.inv_h, L.dx, L.dy, L.dz,
                                                    ctx->keys.as<unsigned long long>(), ctx->vals.as<uint32_t>());
    int maxd = std::max(L.dx, std::max(L.dy, L.dz));
    int b1 = 1; while ((1 << b1) < maxd && b1 < 21) ++b1;
    const int bits = 3 * b1;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->keys.as<unsigned long long>(), ctx->keys2.as<unsigned long long>(),
                                    ctx->vals.as<uint32_t>(), ctx->icp_perm.as<uint32_t>(), n, 0, bits, ctx->stream);
    PW_TRY(ctx->cub_tmp.reserve(ctx, tmp));
    size_t cap = ctx->cub_tmp.cap;
    PW_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, cap, ctx->keys.as<unsigned long long>(), ctx->keys2.as<unsigned long long>(),
                                            ctx->vals.as<uint32_t>(), ctx->icp_perm.as<uint32_t>(), n, 0, bits, ctx->stream));
    src_gather_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->icp_src.as<float4>(), ctx->icp_perm.as<uint32_t>(), n,
                                                       ctx->icp_sorted.as<float4>(),
                                                       have_seed ? ctx->icp_seed.as<int>() : nullptr, ctx->icp_match.as<int>());
    ctx->launches += 5;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

int icp_run_device(Ctx* ctx, const pwicp_icp_params& prm, float* T16, pwicp_icp_result* res,
                   double* mse_trace, float* T_trace, int* idx_trace) {
    const int n = ctx->n_icp;
    if (ctx->n1 < 1 || !ctx->tgt.dev.nlevels) { set_error(ctx, "icp: no target uploaded"); return PWICP_ERR_ARG; }
    if (n < 3) { set_error(ctx, "icp: fewer than 3 correspondences"); return PWICP_ERR_TOO_FEW_CORR; }
    if (prm.max_iter < 1 || prm.max_iter > kMaxIcpIter) { set_error(ctx, "icp: max_iter out of range"); return PWICP_ERR_ARG; }

    const size_t smem = 0;
    int occ = 0;
    PW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_persistent_kernel, kIcpThreads, smem));
    if (occ < 1) { set_error(ctx, "icp: kernel does not fit on an SM"); return PWICP_ERR_CUDA; }
    const long nb = ((long)n + 31) / 32;
    long want = (nb + kIcpWarps - 1) / kIcpWarps;
    int grid = (int)std::min<long>((long)occ * ctx->num_sms, std::max<long>(1, want));

    PW_TRY(ctx->icp_work.reserve(ctx, (size_t)n * sizeof(float4)));
    PW_TRY(ctx->icp_partials.reserve(ctx, (size_t)2 * grid * kNumVals * sizeof(double)));
    const size_t out_bytes = 64 + 16 + (size_t)prm.max_iter * (8 + 64);
    PW_TRY(ctx->icp_out.reserve(ctx, out_bytes));
    char* ob = ctx->icp_out.as<char>();
    PW_CUDA(cudaMemsetAsync(ob, 0, out_bytes, ctx->stream));
    if (idx_trace) PW_TRY(ctx->icp_idx.reserve(ctx, (size_t)prm.max_iter * n * sizeof(int)));

    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    const bool have_seed = ctx->icp_seed_valid && ctx->icp_seed.p != nullptr;
    PW_TRY(icp_sort_source(ctx, n, have_seed));
    ctx->icp_seed_valid = false;                       // seeds belong to one source set

    IcpArgs a;
    a.g = ctx->tgt.dev;
    a.aux = ctx->tgt_aux.as<float4>();
    a.src = ctx->icp_sorted.as<float4>();
    a.work = ctx->icp_work.as<float4>();
    a.match = ctx->icp_match.as<int>();
    a.use_seed0 = have_seed ? 1 : 0;
    a.n = n;
    a.max_iter = prm.max_iter;
    a.force_iters = prm.force_iters;
    a.rot_thr = prm.rot_thr_default ? 0.99999 : (1.0 - prm.tf_eps);
    a.transl_thr = prm.tf_eps;
    a.mse_rel = prm.fit_eps;
    a.mse_abs = 1e-12;
    a.partials = ctx->icp_partials.as<double>();
    a.out_T = reinterpret_cast<float*>(ob);
    a.out_state = reinterpret_cast<int*>(ob + 64);
    a.mse_trace = mse_trace ? reinterpret_cast<double*>(ob + 80) : nullptr;
    a.T_trace = T_trace ? reinterpret_cast<float*>(ob + 80 + (size_t)prm.max_iter * 8) : nullptr;
    a.idx_trace = idx_trace ? ctx->icp_idx.as<int>() : nullptr;

    void* kargs[] = {(void*)&a};
    PW_CUDA(cudaLaunchCooperativeKernel((void*)icp_persistent_kernel, dim3(grid), dim3(kIcpThreads), kargs, smem, ctx->stream));
    ctx->launches++;
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));

    struct { float T[16]; int st[4]; } host;
    PW_CUDA(cudaMemcpyAsync(&host, ob, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    PW_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    const int n_iter = host.st[0];
    if (T16) for (int k = 0; k < 16; ++k) T16[k] = host.T[k];
    if (res) {
        res->n_iter = n_iter; res->conv_state = host.st[1];
        res->grid_blocks = grid; res->warps_per_block = kIcpWarps;
        res->device_ms = ms; res->correspondences = (long long)n_iter * n;
    }
    if (mse_trace) PW_CUDA(cudaMemcpy(mse_trace, ob + 80, (size_t)n_iter * 8, cudaMemcpyDeviceToHost));
    if (T_trace) PW_CUDA(cudaMemcpy(T_trace, ob + 80 + (size_t)prm.max_iter * 8, (size_t)n_iter * 64, cudaMemcpyDeviceToHost));
    if (idx_trace) PW_CUDA(cudaMemcpy(idx_trace, ctx->icp_idx.p, (size_t)n_iter * n * sizeof(int), cudaMemcpyDeviceToHost));
    return PWICP_OK;
}

}  // namespace pwicp
