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
// Batches of 32 points are handed out dynamically (their cost is data dependent; a static split
// left half of the SM time waiting at the grid barrier, profiles/r01c_*), yet the summation order
// ("reduction geometry") is fixed: rows of a batch in order, batches of a 64-batch group
// in order (a second, cheap grid-wide pass), groups by a lane-strided sum + butterfly.
// The oracle's reduce_mode=1 reproduces it for the bit-exact whole-loop parity test.
#include <cooperative_groups.h>

#include "common.cuh"
#include <cub/cub.cuh>

#include "nn_search.cuh"
#include "small_algebra.cuh"

#ifndef PWICP_ICP_MINBLOCKS
#define PWICP_ICP_MINBLOCKS 3
#endif
namespace cg = cooperative_groups;

namespace pwicp {

struct IcpArgs {
    GridDev g;
    const float4* aux;        // level-0 order: nx, ny, nz, ctstd
    const float4* src;        // source set (read only)
    float4* work;             // transformed copy, updated in place every iteration
    int4* cand;               // per source point: candidate cache (nn_search.cuh), .x = last match = seed of the next search
    float4* anchor;           // per source point: position the cache was built at, w = validity radius (0: none)
    float slack;              // cache radius beyond the NN distance
    float build_step2;        // a cache is built only when the point moved less than sqrt(this) in the last step
    int n;
    int max_iter;
    int force_iters;
    double rot_thr, transl_thr, mse_rel, mse_abs;
    double* batch_part;       // [nb][28]: sums of one 32-point batch
    double* group_part;       // [ng][28]: sums of kGroupBatches consecutive batches
    int* batch_counter;       // [max_iter], zeroed before the launch: dynamic batch hand-out
    float* out_T;             // 16: final transformation
    int* out_state;           // [0] n_iter, [1] conv_state
    double* mse_trace;        // nullable
    float* T_trace;           // nullable
    int* idx_trace;           // nullable, [iter][n]
    long long* timing;        // debug build only
};

// Shared scratch of the per-iteration solve.
struct FinishSmem {
    double A[6][6];      // ATA, then its LU factors in place
    double inv[6][6];
    double b[6], x[6];
    double sc[3][2];     // sin / cos of alpha, beta, gamma
    int piv[6];
    float Tn[16];
};

// Warp 0 of every CTA: totals -> 6x6 solve -> float transform -> convergence decision.
// Every scalar operation is the one small_algebra.cuh's sequential inverse6()/solve_from28()
// performs (same operands, same order per element), so the result is bit-identical to the
// single-thread version and to the oracle; the lanes only shorten the critical path (the
// single-thread solve cost ~25 us per inner iteration, profiles/r01d_*).
__device__ __forceinline__ int icp_finish_warp(const IcpArgs& a, int it, const double* s_tot, float* s_T,
                                               float* s_Tfinal, double& prev_mse, FinishSmem& F, int lane) {
    // ATA (mirrored) and ATb from the 28 totals
    for (int idx = lane; idx < 36; idx += 32) {
        const int r = idx / 6, c = idx % 6, lo = min(r, c), hi = max(r, c);
        F.A[r][c] = s_tot[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
    }
    if (lane < 6) { F.b[lane] = s_tot[21 + lane]; F.piv[lane] = lane; }
    __syncwarp();
    // LU with partial pivoting (inverse6)
    for (int k = 0; k < 6; ++k) {
        int p = k;
        if (lane == 0) {
            double big = fabs(F.A[k][k]);
            for (int r = k + 1; r < 6; ++r) { const double v = fabs(F.A[r][k]); if (v > big) { big = v; p = r; } }
        }
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p != k) {
            if (lane < 6) { const double t = F.A[k][lane]; F.A[k][lane] = F.A[p][lane]; F.A[p][lane] = t; }
            if (lane == 0) { const int t = F.piv[k]; F.piv[k] = F.piv[p]; F.piv[p] = t; }
        }
        __syncwarp();
        const double d = F.A[k][k];
        if (d == 0.0) continue;
        const int m = 5 - k;                                    // trailing block is m x m
        double f = 0.0, upd = 0.0;
        int r = 0, c = 0;
        const bool on = lane < m * m;
        if (on) {
            r = k + 1 + lane / m; c = k + 1 + lane % m;
            f = F.A[r][k] / d;
            upd = F.A[r][c] - f * F.A[k][c];
        }
        __syncwarp();
        if (on) { F.A[r][c] = upd; if (c == k + 1) F.A[r][k] = f; }
        __syncwarp();
    }
    // inverse: lane j solves L U x = P e_j
    if (lane < 6) {
        double y[6], xs[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double sacc = (F.piv[r] == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < r; ++c) sacc -= F.A[r][c] * y[c];
            y[r] = sacc;
        }
#pragma unroll
        for (int r = 5; r >= 0; --r) {
            double sacc = y[r];
#pragma unroll
            for (int c = r + 1; c < 6; ++c) sacc -= F.A[r][c] * xs[c];
            xs[r] = sacc / F.A[r][r];
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) F.inv[r][lane] = xs[r];
    }
    __syncwarp();
    if (lane < 6) {
        double sacc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) sacc += F.inv[lane][c] * F.b[c];
        F.x[lane] = sacc;
    }
    __syncwarp();
    if (lane < 3) { F.sc[lane][0] = sin(F.x[lane]); F.sc[lane][1] = cos(F.x[lane]); }
    __syncwarp();
    if (lane == 0) {
        float Tn[16];
        construct_T_sc(F.sc[0][0], F.sc[0][1], F.sc[1][0], F.sc[1][1], F.sc[2][0], F.sc[2][1], F.x, Tn);
        for (int k = 0; k < 16; ++k) F.Tn[k] = Tn[k];
    }
    __syncwarp();
    // final = T * final, one entry per lane (mat4_mul's order: sum over k = 0..3)
    float tf = 0.f;
    if (lane < 16) {
        const int i = lane / 4, j = lane % 4;
        tf = F.Tn[i * 4 + 0] * s_Tfinal[0 * 4 + j];
        tf += F.Tn[i * 4 + 1] * s_Tfinal[1 * 4 + j];
        tf += F.Tn[i * 4 + 2] * s_Tfinal[2 * 4 + j];
        tf += F.Tn[i * 4 + 3] * s_Tfinal[3 * 4 + j];
    }
    __syncwarp();
    if (lane < 16) { s_Tfinal[lane] = tf; s_T[lane] = F.Tn[lane]; }
    __syncwarp();
    int state = 0;
    if (lane == 0) {
        const float* Tn = F.Tn;
        const double mse = s_tot[27] / (double)a.n;
        const int iters = it + 1;
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
                for (int k = 0; k < 16; ++k) a.out_T[k] = s_Tfinal[k];
                a.out_state[0] = iters;
                a.out_state[1] = state;
            }
        }
    }
    return __shfl_sync(0xffffffffu, state, 0);
}

__global__ void __launch_bounds__(kIcpThreads, PWICP_ICP_MINBLOCKS) icp_persistent_kernel(const IcpArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) float s_rows[kIcpWarps][32][8];
    __shared__ double s_tot[kNumVals];
    __shared__ float s_T[16];
    __shared__ float s_Tfinal[16];
    __shared__ int s_stop;
    __shared__ FinishSmem s_fin;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = (a.n + 31) / 32;                              // 32-point batches
    const int ng = (nb + kGroupBatches - 1) / kGroupBatches;     // groups of kGroupBatches batches

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

#ifdef PWICP_TIMING
#define PW_TS(k) do { if (tid == 0 && it == 10 && a.timing) a.timing[blockIdx.x * 8 + (k)] = clock64(); } while (0)
#else
#define PW_TS(k)
#endif
    for (int it = 0;; ++it) {
        PW_TS(0);
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = s_T[k];

        // ---- phase A: batches are handed out dynamically, kGrab at a time (the cost of a batch
        // depends on the data), but every sum below is formed in an order that does not depend on
        // which warp does it
        for (;;) {
            int b0 = 0;
            if (lane == 0) b0 = atomicAdd(a.batch_counter + it, kGrab);
            b0 = __shfl_sync(0xffffffffu, b0, 0);
            if (b0 >= nb) break;
            const int b1 = min(b0 + kGrab, nb);
            for (int b = b0; b < b1; ++b) {
                const int i = b * 32 + lane;
                const bool active = i < a.n;
                float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
                if (active) {
                    float4 p = (it == 0) ? __ldg(a.src + i) : a.work[i];
                    const int4 c = a.cand[i];
                    const float4 an = a.anchor[i];
                    float step2 = __int_as_float(0x7f800000);
                    if (it > 0) {
                        float x, y, z;
                        xform_point(T, p.x, p.y, p.z, x, y, z);
                        step2 = l2_simple(x, y, z, p.x, p.y, p.z);
                        p.x = x; p.y = y; p.z = z;
                    }
                    a.work[i] = p;
                    // (b) exact NN: best of the cached candidates when the cache still covers the
                    // query (nn_search.cuh, "candidate cache"), else the seeded ball search
                    Best bb;
                    bool ok = false;
                    int seed = c.x;
                    if (seed >= 0) {
                        const float4* __restrict__ pts = a.g.lv[0].pts;
                        const float4 q0 = __ldg(pts + c.x), q1 = __ldg(pts + c.y), q2 = __ldg(pts + c.z), q3 = __ldg(pts + c.w);
                        bb.d2 = l2_simple(p.x, p.y, p.z, q0.x, q0.y, q0.z);
                        bb.idx = __float_as_int(q0.w); bb.pos = c.x; bb.qx = q0.x; bb.qy = q0.y; bb.qz = q0.z;
#define PW_CAND(q, cp)                                                                         \
                        {                                                                      \
                            const float d = l2_simple(p.x, p.y, p.z, q.x, q.y, q.z);           \
                            const int id = __float_as_int(q.w);                                \
                            if (d < bb.d2 || (d == bb.d2 && id < bb.idx)) {                    \
                                bb.d2 = d; bb.idx = id; bb.pos = cp; bb.qx = q.x; bb.qy = q.y; bb.qz = q.z; \
                            }                                                                  \
                        }
                        PW_CAND(q1, c.y) PW_CAND(q2, c.z) PW_CAND(q3, c.w)
#undef PW_CAND
                        const float da = l2_simple(p.x, p.y, p.z, an.x, an.y, an.z);
                        ok = sqrtf(bb.d2) + sqrtf(da) < an.w;
                        seed = bb.pos;
                    }
                    if (!ok) {
                        bb = nn_search_seeded<true>(a.g, p.x, p.y, p.z, seed);
                        CandCache cc;
                        cc.rho = 0.f;
                        if (step2 < a.build_step2) {
                            const float rm = sqrtf(bb.d2) + a.slack;
                            cc = ball_collect(a.g.lv[0], a.g.ox, a.g.oy, a.g.oz, p.x, p.y, p.z, rm * rm);
                        }
                        if (cc.rho > 0.f) {
                            a.cand[i] = make_int4(cc.pos[0], cc.pos[1], cc.pos[2], cc.pos[3]);
                            a.anchor[i] = make_float4(p.x, p.y, p.z, cc.rho);
                        } else {
                            a.cand[i] = make_int4(bb.pos, bb.pos, bb.pos, bb.pos);
                            if (an.w != 0.f) a.anchor[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
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
                // batch sums: lane v adds its product over rows 0..31 in order, starting from 0
                if (lane < kNumVals) {
                    double acc = 0.0;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        const double x = (double)s_rows[warp][r][va];
                        const double y = (lane == 27) ? 1.0 : (double)s_rows[warp][r][vb];
                        acc = __fma_rn(x, y, acc);   // the product of two float values is exact in double, so
                                                     // this is acc + x*y with one rounding, fused or not
                    }
                    __stcg(a.batch_part + (size_t)b * kNumVals + lane, acc);
                }
                __syncwarp();
            }
        }
        PW_TS(1);
        grid.sync();
        PW_TS(2);

        // ---- phase B1: one warp per (group, value): lane l adds the group's batch sums l, l+32, ...
        // in ascending order, then an xor butterfly 16,8,4,2,1 -- a fixed order, 16 independent loads
        {
            const int gwarp = blockIdx.x * kIcpWarps + warp, nwarps = gridDim.x * kIcpWarps;
            for (int w = gwarp; w < ng * kNumVals; w += nwarps) {
                const int g = w / kNumVals, v = w % kNumVals;
                const int gsize = min(kGroupBatches, nb - g * kGroupBatches);
                const double* bp = a.batch_part + (size_t)g * kGroupBatches * kNumVals + v;
                double sg = 0.0;
                for (int k = lane; k < gsize; k += 32) sg += __ldcg(bp + (size_t)k * kNumVals);
#pragma unroll
                for (int o = 16; o; o >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);
                if (lane == 0) __stcg(a.group_part + (size_t)g * kNumVals + v, sg);
            }
        }
        PW_TS(3);
        grid.sync();
        PW_TS(4);

        // ---- phase B2: the same pattern over the group sums (every CTA computes the same totals)
        const double* P = a.group_part;
        for (int v = warp; v < kNumVals; v += kIcpWarps) {
            double sv = 0.0;
            for (int g = lane; g < ng; g += 32) sv += __ldcg(P + (size_t)g * kNumVals + v);
#pragma unroll
            for (int o = 16; o; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
            if (lane == 0) s_tot[v] = sv;
        }
        __syncthreads();

        if (warp == 0) {
            const int st = icp_finish_warp(a, it, s_tot, s_T, s_Tfinal, prev_mse, s_fin, lane);
            if (lane == 0) s_stop = st;
        }
        __syncthreads();
        PW_TS(5);
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
                                  const int* __restrict__ seed_in, int4* __restrict__ cand, float4* __restrict__ anchor) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = order[i];
    float4 p = src[o];
    p.w = __int_as_float((int)o);
    out[i] = p;
    const int sd = seed_in ? seed_in[o] : -1;
    cand[i] = make_int4(sd, sd, sd, sd);
    anchor[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    PW_TRY(ctx->icp_match.reserve(ctx, (size_t)n * (sizeof(int4) + sizeof(float4))));
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
                                                       have_seed ? ctx->icp_seed.as<int>() : nullptr, ctx->icp_match.as<int4>(),
                                                       reinterpret_cast<float4*>(ctx->icp_match.as<int4>() + n));
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
    // enough CTAs that every warp can own a batch, never more than can be co-resident
    long want = (nb + kIcpWarps - 1) / kIcpWarps;
    int grid = (int)std::min<long>((long)occ * ctx->num_sms, std::max<long>(1, want));

    PW_TRY(ctx->icp_work.reserve(ctx, (size_t)n * sizeof(float4)));
    const int ngroups = (int)((nb + kGroupBatches - 1) / kGroupBatches);
    const size_t bytes_batch = (size_t)nb * kNumVals * sizeof(double);
    const size_t bytes_group = (size_t)2 * ngroups * kNumVals * sizeof(double);
    const size_t bytes_cnt = ((size_t)prm.max_iter + ngroups) * sizeof(int);
    PW_TRY(ctx->icp_partials.reserve(ctx, bytes_batch + bytes_group + bytes_cnt + 64));
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
    a.cand = ctx->icp_match.as<int4>();
    a.anchor = reinterpret_cast<float4*>(a.cand + n);
    a.slack = 0.03f / ctx->tgt.dev.lv[0].inv_h;
    a.build_step2 = (0.25f * a.slack) * (0.25f * a.slack);

    a.n = n;
    a.max_iter = prm.max_iter;
    a.force_iters = prm.force_iters;
    a.rot_thr = prm.rot_thr_default ? 0.99999 : (1.0 - prm.tf_eps);
    a.transl_thr = prm.tf_eps;
    a.mse_rel = prm.fit_eps;
    a.mse_abs = 1e-12;
    a.batch_part = ctx->icp_partials.as<double>();
    a.group_part = a.batch_part + (size_t)nb * kNumVals;
    a.batch_counter = reinterpret_cast<int*>(ctx->icp_partials.as<char>() + bytes_batch + bytes_group);
    PW_CUDA(cudaMemsetAsync(a.batch_counter, 0, bytes_cnt, ctx->stream));
    a.out_T = reinterpret_cast<float*>(ob);
    a.out_state = reinterpret_cast<int*>(ob + 64);
    a.mse_trace = mse_trace ? reinterpret_cast<double*>(ob + 80) : nullptr;
    a.T_trace = T_trace ? reinterpret_cast<float*>(ob + 80 + (size_t)prm.max_iter * 8) : nullptr;
    a.idx_trace = idx_trace ? ctx->icp_idx.as<int>() : nullptr;
    a.timing = nullptr;
#ifdef PWICP_TIMING
    static long long* d_timing = nullptr;
    if (!d_timing) cudaMalloc(&d_timing, 1024 * 8 * sizeof(long long));
    cudaMemsetAsync(d_timing, 0, 1024 * 8 * sizeof(long long), ctx->stream);
    a.timing = d_timing;
#endif

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
#ifdef PWICP_TIMING
    {
        std::vector<long long> ht(1024 * 8);
        cudaMemcpy(ht.data(), d_timing, ht.size() * 8, cudaMemcpyDeviceToHost);
        long long t0min = -1;
        for (int b = 0; b < grid; ++b) if (ht[b * 8]) t0min = (t0min < 0 || ht[b * 8] < t0min) ? ht[b * 8] : t0min;
        double sum[6] = {0}, mx[6] = {0}, mn[6] = {1e30, 1e30, 1e30, 1e30, 1e30, 1e30};
        for (int b = 0; b < grid; ++b) for (int k = 0; k < 6; ++k) { double v = (double)(ht[b * 8 + k] - t0min); sum[k] += v; mx[k] = std::max(mx[k], v); mn[k] = std::min(mn[k], v); }
        if (t0min > 0) { printf("TIMING it=10 cycles (min/avg/max over %d CTAs):", grid); for (int k = 0; k < 6; ++k) printf(" [%d] %.0f/%.0f/%.0f", k, mn[k], sum[k] / grid, mx[k]); printf("\n"); }
    }
#endif
    const int n_iter = host.st[0];
    if (T16) for (int k = 0; k < 16; ++k) T16[k] = host.T[k];
    if (res) {
        res->n_iter = n_iter; res->conv_state = host.st[1];
        res->grid_blocks = grid; res->warps_per_block = kIcpWarps; res->group_batches = kGroupBatches;
        res->device_ms = ms; res->correspondences = (long long)n_iter * n;
    }
    if (mse_trace) PW_CUDA(cudaMemcpy(mse_trace, ob + 80, (size_t)n_iter * 8, cudaMemcpyDeviceToHost));
    if (T_trace) PW_CUDA(cudaMemcpy(T_trace, ob + 80 + (size_t)prm.max_iter * 8, (size_t)n_iter * 64, cudaMemcpyDeviceToHost));
    if (idx_trace) PW_CUDA(cudaMemcpy(idx_trace, ctx->icp_idx.p, (size_t)n_iter * n * sizeof(int), cudaMemcpyDeviceToHost));
    return PWICP_OK;
}

}  // namespace pwicp
