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

#ifndef PWICP_STATIC_EIGHTHS
// candidate cache: radius beyond the NN distance (in level-0 cells) and the largest last step (as a fraction of it) at
// which a cache is built
#ifndef PWICP_SLACK_CELLS
#define PWICP_SLACK_CELLS 0.03f
#endif
#ifndef PWICP_BUILD_FRAC
#define PWICP_BUILD_FRAC 0.25f
#endif
#define PWICP_STATIC_EIGHTHS 6     // share of a warp's batches that is assigned statically once the loop is calm
#endif
#ifndef PWICP_ICP_MINBLOCKS
#define PWICP_ICP_MINBLOCKS 3
#endif
namespace cg = cooperative_groups;

namespace pwicp {

struct IcpArgs {
    GridDev g;
    const float4* aux;        // level-0 order: nx, ny, nz, ctstd
    const float4* src;        // source set (read only)
    float4* work;             // transformed copy, updated in place every iteration; w = path length since the
                              // candidate cache of the point was built
    // per source point: the candidate cache (nn_search.cuh)
    float4* cq0;              // primary candidate INLINE: x, y, z of the last match, w = its level-0 position (int bits, -1: none)
    float4* cn0;              // normal of the primary candidate INLINE, w = validity radius of the cache (0: none); the
                              // path length the query has travelled since the cache was built lives in work[].w
    int4* cmore;              // x, y, z: positions of up to three further candidates (unused = primary), w = original index of the primary
    int seed_exact;           // cand[].x is the exact NN of the untransformed source (iteration 0 needs no search)
    float slack;              // cache radius beyond the NN distance
    float build_step2;        // a cache is built only when the point moved less than sqrt(this) in the last step
    int n;
    int max_iter;
    int force_iters;
    double rot_thr, transl_thr, mse_rel, mse_abs;
    double* part[kMaxRedLevels];  // [count[l]][28]: level 0 = sums of one 32-point batch, level l+1 = sums of
                                  // kFanIn consecutive level-l entries
    int* done;                    // [count[2]]: groups finished per supergroup, monotonic over iterations
    int count[kMaxRedLevels];     // entries per level
    int nlevels;
    const double* top_part;       // = part[nlevels - 1]
    int top_count;                // = count[nlevels - 1]
    int* batch_counter;       // [max_iter][kHandoutLanes * 32], zeroed before the launch: dynamic batch hand-out
                              // (kHandoutLanes counters per iteration, 128 bytes apart)
    int* fallbacks;           // [max_iter], zeroed before the launch: queries that ran the ball search
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
    double b[6], x[6], rcp[6];
    double sc[3][2];     // sin / cos of alpha, beta, gamma
    double prev_mse;
    int piv[6];
    float Tn[16];
};

#ifdef PWICP_TIMING
#define PW_TSF(k) do { if (lane == 0 && it == 30 && a.timing) a.timing[blockIdx.x * 16 + (k)] = clock64(); } while (0)
#else
#define PW_TSF(k)
#endif

// Warp 0 of every CTA: totals -> 6x6 solve -> float transform -> convergence decision.
// Every scalar operation is the one small_algebra.cuh's sequential inverse6()/solve_from28()
// performs (same operands, same order per element), so the result is bit-identical to the
// single-thread version and to the oracle; the lanes only shorten the critical path.
// This code runs once per inner iteration on one warp, i.e. always from a cold instruction cache
// (profiles/r01d_*: a straight-line version spent ~10 us mostly fetching instructions), hence the
// rolled loops and the out-of-line placement: few instruction lines, re-used from L0.
static __device__ __noinline__ int icp_finish_warp(const IcpArgs& a, int it, const double* s_tot, float* s_T,
                                                   float* s_Tfinal, FinishSmem& F, int lane) {
    // ATA (mirrored) and ATb from the 28 totals
    for (int idx = lane; idx < 36; idx += 32) {
        const int r = idx / 6, c = idx % 6, lo = min(r, c), hi = max(r, c);
        F.A[r][c] = s_tot[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
    }
    if (lane < 6) { F.b[lane] = s_tot[21 + lane]; F.piv[lane] = lane; }
    __syncwarp();
    // LU with partial pivoting (inverse6); lanes 0..24 own the entries of the trailing 5x5 block
    const int rr = lane / 5 + 1, cc = lane % 5 + 1;
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
        // first row r >= k with the largest |A[r][k]|: the bit pattern of a non-negative double
        // orders like an unsigned integer
        const bool part = lane >= k && lane < 6;
        const double v = part ? fabs(F.A[lane][k]) : 0.0;
        const unsigned vh = (unsigned)__double2hiint(v), vl = (unsigned)__double2loint(v);
        const unsigned mh = __reduce_max_sync(0xffffffffu, vh);
        const unsigned ml = __reduce_max_sync(0xffffffffu, vh == mh ? vl : 0u);
        const int p = __ffs(__ballot_sync(0xffffffffu, part && vh == mh && vl == ml)) - 1;
        if (p != k) {
            if (lane < 6) { const double t = F.A[k][lane]; F.A[k][lane] = F.A[p][lane]; F.A[p][lane] = t; }
            if (lane == 0) { const int t = F.piv[k]; F.piv[k] = F.piv[p]; F.piv[p] = t; }
        }
        __syncwarp();
        const double d = F.A[k][k];
        if (d == 0.0) continue;
        const bool on = lane < 25 && rr > k && cc > k;
        double f = 0.0, upd = 0.0;
        if (on) {
            f = F.A[rr][k] / d;
            upd = F.A[rr][cc] - f * F.A[k][cc];
        }
        __syncwarp();
        if (on) { F.A[rr][cc] = upd; if (cc == k + 1) F.A[rr][k] = f; }
        __syncwarp();
    }
    PW_TSF(8);
    // inverse: lane j solves L U x = P e_j (one reciprocal per pivot, formed by six lanes at once)
    if (lane < 6) F.rcp[lane] = 1.0 / F.A[lane][lane];
    __syncwarp();
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
            xs[r] = sacc * F.rcp[r];
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
    PW_TSF(6);
    if (lane < 3) sincos(F.x[lane], &F.sc[lane][0], &F.sc[lane][1]);
    __syncwarp();
    PW_TSF(7);
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
        const double prev_mse = F.prev_mse;
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
        F.prev_mse = mse;
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

// The search path of a query whose candidate cache does not cover it: seeded ball search, and --
// once the point has (nearly) stopped moving -- a new cache around its position.  Out of line on
// purpose: it runs for every query during the first few iterations and (almost) never afterwards,
// and the steady-state loop has to stay small enough to live in the instruction cache together
// with the per-iteration solve (profiles/r01e_*: 6 us per iteration otherwise).
struct Fallback {
    Best bb;
    float nx, ny, nz;     // normal of the match
    int built;            // a new cache was written: the caller restarts the path length
};
static __device__ __noinline__ Fallback icp_search_fallback(const IcpArgs& a, int it, int i, float px, float py, float pz,
                                                            int seed, float step2) {
    {   // queries that needed the search this iteration (decides the scheduling of the next one)
        const unsigned m = __activemask();
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(a.fallbacks + it, __popc(m));
    }
#ifdef PWICP_TIMING
    if (a.timing && it < 64) {
        atomicAdd((unsigned long long*)a.timing + 16 * 1024 + it * 2, 1ull);
        if (step2 < a.build_step2) atomicAdd((unsigned long long*)a.timing + 16 * 1024 + it * 2 + 1, 1ull);
    }
#endif
    Fallback f;
    f.bb = nn_search_seeded<true>(a.g, px, py, pz, seed);
    const Best& bb = f.bb;
    const float4 nq = __ldg(a.aux + bb.pos);
    f.nx = nq.x; f.ny = nq.y; f.nz = nq.z;
    CandCache cc;
    cc.rho = 0.f;
    if (step2 < a.build_step2) {
        const float rm = sqrtf(bb.d2) + a.slack;
        cc = ball_collect(a.g.lv[0], a.g.ox, a.g.oy, a.g.oz, px, py, pz, rm * rm);
    }
    // the match is the primary candidate; the other cached targets follow in any order
    int o1 = bb.pos, o2 = bb.pos, o3 = bb.pos;
    if (cc.rho > 0.f) {
        int k = 0;
#pragma unroll
        for (int j = 0; j < kCacheCands; ++j)
            if (cc.pos[j] != bb.pos) { if (k == 0) o1 = cc.pos[j]; else if (k == 1) o2 = cc.pos[j]; else if (k == 2) o3 = cc.pos[j]; ++k; }
        // all four slots taken by targets other than the match cannot happen (the match is the nearest
        // of the collected set); guard anyway: no cache rather than a wrong one
        if (k > 3) { cc.rho = 0.f; o1 = o2 = o3 = bb.pos; }
    }
    f.built = cc.rho > 0.f;
    a.cq0[i] = make_float4(bb.qx, bb.qy, bb.qz, __int_as_float(bb.pos));
    a.cn0[i] = make_float4(nq.x, nq.y, nq.z, cc.rho);
    a.cmore[i] = make_int4(o1, o2, o3, bb.idx);
    return f;
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only: the sources are rewritten by other
// SMs every iteration), per-thread completion groups.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Sum of up to kFanIn entries (stride kNumVals doubles) in order, starting from 0.  All loads are
// issued before the first add; absent entries contribute +0.0, which leaves the sum unchanged.
__device__ __forceinline__ double sum_entries(const double* __restrict__ src, int size) {
    double v[kFanIn];
#pragma unroll
    for (int k = 0; k < kFanIn; ++k) v[k] = (k < size) ? __ldcg(src + (size_t)k * kNumVals) : 0.0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kFanIn; ++k) acc += v[k];
    return acc;
}

__global__ void __launch_bounds__(kIcpThreads, PWICP_ICP_MINBLOCKS) icp_persistent_kernel(const IcpArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) float s_rows[kIcpWarps][32][8];
    // staging of the streamed per-point data, two batches deep per warp, filled by cp.async:
    // [slot][field: point + path length, normal of the primary + radius, primary candidate, further candidates][lane]
    extern __shared__ __align__(16) unsigned char s_dyn[];
    float4 (*s_stage)[2][4][32] = reinterpret_cast<float4 (*)[2][4][32]>(s_dyn);
    __shared__ double s_tot[kNumVals];
    __shared__ float s_T[16];
    __shared__ float s_Tfinal[16];
    __shared__ int s_stop;
    __shared__ int s_fb;      // queries that ran the search in the previous iteration
    __shared__ FinishSmem s_fin;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = (a.n + 31) / 32;                              // 32-point batches

    // DMMA fragment roles of this lane (row m = lane / 4 of A and column m of B; D[m][2k], D[m][2k+1]
    // with k = lane % 4).  Row layout in shared memory: a b c nx ny nz e d2.  The 28 values are the
    // 21 upper-triangle ATA entries (row-major), the 6 ATb entries, and the sum of squared NN distances.
    const int offA = (lane >> 2) < 6 ? (lane >> 2) : 7;      // A row 6 (and the unused row 7): d2
    const int offB = (lane >> 2) < 6 ? (lane >> 2) : 6;      // B column 6: e; column 7 is the constant 1
    int v0 = -1, v1 = -1;
    {
        const int rD = lane >> 2;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int c = 2 * (lane & 3) + jj;
            int v = -1;
            if (rD < 6) {
                if (c < 6 && c >= rD) v = rD * 6 - rD * (rD - 1) / 2 + (c - rD);
                else if (c == 6) v = 21 + rD;
            } else if (rD == 6 && c == 7) v = 27;
            if (jj == 0) v0 = v; else v1 = v;
        }
    }
    if (tid < 16) { s_T[tid] = (tid % 5 == 0) ? 1.0f : 0.0f; s_Tfinal[tid] = s_T[tid]; }
    if (tid == 0) { s_stop = 0; s_fin.prev_mse = 1.7976931348623157e308; }   // DBL_MAX
    __syncthreads();

#ifdef PWICP_TIMING
#define PW_TS(k) do { if (tid == 0 && it == 30 && a.timing) a.timing[blockIdx.x * 16 + (k)] = clock64(); } while (0)
#else
#define PW_TS(k)
#endif
    bool staged = false;      // the first batch of this iteration is already being copied
    for (int it = 0;; ++it) {
        PW_TS(0);
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = s_T[k];
        // copies of a batch: point + path length, normal + radius, primary candidate, further candidates
        auto stage_batch = [&](const float4* __restrict__ psrc, int bb_, int slot) {
            const int i_ = bb_ * 32 + lane;
            if (bb_ < nb && i_ < a.n) {
                cp_async16(&s_stage[warp][slot][0][lane], psrc + i_);
                cp_async16(&s_stage[warp][slot][1][lane], a.cn0 + i_);
                cp_async16(&s_stage[warp][slot][2][lane], a.cq0 + i_);
                cp_async16(&s_stage[warp][slot][3][lane], a.cmore + i_);
            }
            cp_async_commit();
        };
        // ---- phase A.  Every warp first works through a static share of the batches (b = j * NW + W,
        // three quarters of its fair share once the loop is calm), then takes single batches from a counter: the cost of a batch is
        // data dependent while the ball search runs, and one atomic per batch on one address would
        // serialise in L2 (~0.85 cycles each).  Every sum below is formed in an order that does not
        // depend on which warp does it.  The loop is software pipelined: the loads of the next
        // batch and the hand-out of the one after are in flight while the current one is processed.
        {
            const int NW = gridDim.x * kIcpWarps, W = blockIdx.x * kIcpWarps + warp;
            // static batches per warp: three quarters of the fair share once (nearly) every query is answered from its
            // cache (uniform cost per batch), else only the two that cover the pipeline depth of the hand-out
            const int fb_prev = (it > 1) ? s_fb : a.n;
            const bool calm = (long long)fb_prev * 64 < (long long)a.n;
            const int J = (calm ? (nb / NW) * PWICP_STATIC_EIGHTHS / 8 : 0) + 2;
            const bool has_dyn = (long long)J * NW < (long long)nb;   // else no hand-out tickets at all
            const float4* __restrict__ psrc = (it == 0) ? a.src : a.work;
            // kHandoutLanes counters, counter c hands out the dynamic batches D0 + c, D0 + c + kHandoutLanes, ...:
            // a single address makes every ticket queue behind thousands of others in the L2 atomic unit
            // (16 % of all stall samples, profiles/r01h_*); interleaving keeps the lanes equally loaded
            const int nl = min(kHandoutLanes, NW), hl = W % nl;
            int* counter = a.batch_counter + ((size_t)it * kHandoutLanes + hl) * 32;
            int seq = 0;                                       // position in this warp's batch sequence
            int tkt = 0;                                       // lane 0: hand-out ticket in flight
            int b = W, bn = NW + W;                            // positions 0 and 1
            if (!staged) {                                     // else: issued during the previous iteration's solve
                stage_batch(psrc, b, 0);
            }
            while (b < nb) {
                const int slot = seq & 1;
                // copies of the next batch and the ticket for the one after are in flight during this one
                stage_batch(psrc, bn, slot ^ 1);
                if (has_dyn && seq + 2 >= J && lane == 0) tkt = atomicAdd(counter, 1);   // position seq+2 is dynamic
                cp_async_wait_group1();                        // everything but the copies just issued has landed

                const int i = b * 32 + lane;
                const bool active = i < a.n;
                float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
                if (active) {
                    float4 p = s_stage[warp][slot][0][lane];
                    const float4 cn = s_stage[warp][slot][1][lane];
                    const float4 q0 = s_stage[warp][slot][2][lane];
                    const int4 cm = *reinterpret_cast<const int4*>(&s_stage[warp][slot][3][lane]);
                    const int pos0 = __float_as_int(q0.w);
                    float step2 = __int_as_float(0x7f800000);
                    float path = 0.f;                          // travelled since the cache was built (upper bound)
                    if (it > 0) {
                        float x, y, z;
                        xform_point(T, p.x, p.y, p.z, x, y, z);
                        step2 = l2_simple(x, y, z, p.x, p.y, p.z);
                        path = p.w + sqrtf(step2) * 1.000001f;
                        p.x = x; p.y = y; p.z = z;
                    }
                    // (b) exact NN: best of the cached candidates when the cache still covers the
                    // query (nn_search.cuh, "candidate cache"), else the seeded ball search
                    Best bb;
                    bool ok = false;
                    int seed = pos0;
                    if (pos0 >= 0) {
                        const float4* __restrict__ pts = a.g.lv[0].pts;
                        bb.d2 = l2_simple(p.x, p.y, p.z, q0.x, q0.y, q0.z);
                        bb.idx = cm.w; bb.pos = pos0; bb.qx = q0.x; bb.qy = q0.y; bb.qz = q0.z;
                        // unused slots repeat the primary: their loads are predicated off (no L1 traffic); the
                        // three loads are issued together, one round trip instead of up to three
                        float4 q1 = q0, q2 = q0, q3 = q0;
                        if (cm.x != pos0) q1 = __ldg(pts + cm.x);
                        if (cm.y != pos0) q2 = __ldg(pts + cm.y);
                        if (cm.z != pos0) q3 = __ldg(pts + cm.z);
#define PW_CAND(q, cp)                                                                         \
                        if ((cp) != pos0) {                                                    \
                            const float d = l2_simple(p.x, p.y, p.z, q.x, q.y, q.z);           \
                            const int id = __float_as_int(q.w);                                \
                            if (d < bb.d2 || (d == bb.d2 && id < bb.idx)) {                    \
                                bb.d2 = d; bb.idx = id; bb.pos = (cp); bb.qx = q.x; bb.qy = q.y; bb.qz = q.z; \
                            }                                                                  \
                        }
                        PW_CAND(q1, cm.x) PW_CAND(q2, cm.y) PW_CAND(q3, cm.z)
#undef PW_CAND
                        // |p - anchor| <= path (triangle inequality over the steps actually taken)
                        ok = (it == 0 && a.seed_exact) || sqrtf(bb.d2) + path * 1.00001f < cn.w;
                        seed = bb.pos;
                    }
                    float nx = cn.x, ny = cn.y, nz = cn.z;
                    if (!ok) {
                        const Fallback f = icp_search_fallback(a, it, i, p.x, p.y, p.z, seed, step2);
                        bb = f.bb; nx = f.nx; ny = f.ny; nz = f.nz;
                        if (f.built) path = 0.f;
                    } else if (bb.pos != pos0) {
                        // another cached target has become the nearest: make it the primary
                        const float4 nq = __ldg(a.aux + bb.pos);
                        nx = nq.x; ny = nq.y; nz = nq.z;
                        a.cq0[i] = make_float4(bb.qx, bb.qy, bb.qz, __int_as_float(bb.pos));
                        a.cn0[i] = make_float4(nx, ny, nz, cn.w);
                        a.cmore[i] = make_int4(cm.x == bb.pos ? pos0 : cm.x, cm.y == bb.pos ? pos0 : cm.y,
                                               cm.z == bb.pos ? pos0 : cm.z, bb.idx);
                    }
                    a.work[i] = make_float4(p.x, p.y, p.z, path);
                    const float sx = p.x, sy = p.y, sz = p.z;
                    const float dx = bb.qx, dy = bb.qy, dz = bb.qz;
                    // float expressions of TransformationEstimationPointToPlaneLLS (no FMA)
                    lo.x = nz * sy - ny * sz;
                    lo.y = nx * sz - nz * sx;
                    lo.z = ny * sx - nx * sy;
                    lo.w = nx;
                    hi.x = ny;
                    hi.y = nz;
                    hi.z = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz;
                    hi.w = bb.d2;
                    // traces are reported in the caller's order (src[].w = original source index)
                    if (a.idx_trace) a.idx_trace[(size_t)it * a.n + __float_as_int(__ldg(a.src + i).w)] = bb.idx;
                }
                // rows through shared memory as floats: a broadcast LDS.32 is one wavefront, an LDS.64 two,
                // and the L1/shared pipe is the busiest unit of this kernel (profiles/r01e_*)
                float4* row = reinterpret_cast<float4*>(&s_rows[warp][lane][0]);
                row[0] = lo; row[1] = hi;
                __syncwarp();
                // batch sums on the FP64 tensor cores: the 28 sums are entries of D = A * B with
                // A = [a b c nx ny nz d2 -]^T (8 x 32) and B = [a b c nx ny nz e 1] (32 x 8), formed by eight
                // chained DMMA.8x8x4.  On B200 a DMMA adds its four products to the accumulator one after
                // the other, each with one rounding, in k order (scripts/micro/dmma_order.cu: 0 mismatches in
                // 128 000 sums against a DFMA chain), and a product of two float values is exact in
                // double: every sum is over rows 0..31 in order, starting from 0 -- the order the oracle uses.
                {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float* r = &s_rows[warp][4 * j + (lane & 3)][0];
                        const double av = (double)r[offA];
                        const double bv = (lane >= 28) ? 1.0 : (double)r[offB];
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
                    }
                    double* dst = a.part[0] + (size_t)b * kNumVals;
                    if (v0 >= 0) __stcg(dst + v0, c0);
                    if (v1 >= 0) __stcg(dst + v1, c1);
                }
                __syncwarp();
                // rotate the pipeline
                ++seq;
                b = bn;
                bn = (seq + 1 < J) ? (seq + 1) * NW + W
                                   : (has_dyn ? J * NW + __shfl_sync(0xffffffffu, tkt, 0) * nl + hl : nb);
            }
            cp_async_wait_all();
        }
        PW_TS(1);
        grid.sync();
        PW_TS(2);
        // the next iteration's first batch (always position 0 = batch W): its copies do not depend on the
        // transform being solved for, so they fly during the reduction and the solve
        stage_batch(a.work, blockIdx.x * kIcpWarps + warp, 0);
        staged = true;

        // ---- phase B1: one warp per group of kFanIn batches: lane v sums the group's entries of
        // value v in order, starting from 0 (loads are independent, the adds sequential).  The warp
        // that completes the last group of a kFanIn-group supergroup sums that one the same way.
        if (a.nlevels > 1) {
            const int gwarp = blockIdx.x * kIcpWarps + warp, nwarps = gridDim.x * kIcpWarps;
            for (int g = gwarp; g < a.count[1]; g += nwarps) {
                if (lane < kNumVals)
                    __stcg(a.part[1] + (size_t)g * kNumVals + lane,
                           sum_entries(a.part[0] + (size_t)g * kFanIn * kNumVals + lane, min(kFanIn, a.count[0] - g * kFanIn)));
                if (a.nlevels > 2) {
                    const int sg = g / kFanIn, size = min(kFanIn, a.count[1] - sg * kFanIn);
                    __threadfence();
                    __syncwarp();
                    int last = 0;
                    if (lane == 0) last = (atomicAdd(a.done + sg, 1) + 1 == (it + 1) * size);
                    last = __shfl_sync(0xffffffffu, last, 0);
                    if (last) {
                        __threadfence();
                        if (lane < kNumVals)
                            __stcg(a.part[2] + (size_t)sg * kNumVals + lane,
                                   sum_entries(a.part[1] + (size_t)sg * kFanIn * kNumVals + lane, size));
                    }
                }
            }
            PW_TS(3);
            grid.sync();
        }

        // ---- phase B2: every CTA forms the same totals from the top-level entries, in order
        if (warp == 0) {
            if (lane < kNumVals) {
                double acc = 0.0;
                for (int k0 = 0; k0 < a.top_count; k0 += kFanIn)
                    acc += sum_entries(a.top_part + (size_t)k0 * kNumVals + lane, min(kFanIn, a.top_count - k0));
                s_tot[lane] = acc;
            }
            if (lane == 31) s_fb = __ldcg(a.fallbacks + it);   // complete since the first grid barrier
            __syncwarp();
        }
        PW_TS(4);

        if (warp == 0) {
            const int st = icp_finish_warp(a, it, s_tot, s_T, s_Tfinal, s_fin, lane);
            if (lane == 0) s_stop = st;
        }
        __syncthreads();
        PW_TS(5);
        if (s_stop) break;
    }
    cp_async_wait_all();
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
                                  const int* __restrict__ seed_in, const float4* __restrict__ tgt_pts,
                                  const float4* __restrict__ tgt_aux,
                                  float4* __restrict__ cn0, float4* __restrict__ cq0, int4* __restrict__ cmore) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = order[i];
    float4 p = src[o];
    p.w = __int_as_float((int)o);
    out[i] = p;
    const int sd = seed_in ? seed_in[o] : -1;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    int idx = 0;
    float4 nq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sd >= 0) { q = __ldg(tgt_pts + sd); idx = __float_as_int(q.w); nq = __ldg(tgt_aux + sd); }
    cq0[i] = make_float4(q.x, q.y, q.z, __int_as_float(sd));
    cmore[i] = make_int4(sd, sd, sd, idx);
    cn0[i] = make_float4(nq.x, nq.y, nq.z, 0.f);
}

// Iteration 0 of a source set without seeds: the plain search at full occupancy (the persistent
// kernel is register-capped and runs it about twice as slowly, profiles/r01f_*).  The matches go
// into the candidate slots; the persistent kernel takes them as the exact answer of iteration 0.
__global__ void __launch_bounds__(256)
icp_seed_kernel(GridDev g, const float4* __restrict__ tgt_aux, const float4* __restrict__ src, int n,
                float4* __restrict__ cn0, float4* __restrict__ cq0, int4* __restrict__ cmore) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(src + i);
    const Best b = nn_search_seeded(g, p.x, p.y, p.z, -1);
    const float4 nq = __ldg(tgt_aux + b.pos);
    cq0[i] = make_float4(b.qx, b.qy, b.qz, __int_as_float(b.pos));
    cn0[i] = make_float4(nq.x, nq.y, nq.z, 0.f);
    cmore[i] = make_int4(b.pos, b.pos, b.pos, b.idx);
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
    PW_TRY(ctx->icp_match.reserve(ctx, (size_t)n * 3 * sizeof(float4)));      // cn0, cq0, cmore
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
                                                       have_seed ? ctx->icp_seed.as<int>() : nullptr, ctx->tgt.dev.lv[0].pts,
                                                       ctx->tgt_aux.as<float4>(), ctx->icp_match.as<float4>(), ctx->icp_match.as<float4>() + n,
                                                       reinterpret_cast<int4*>(ctx->icp_match.as<float4>() + 2 * (size_t)n));
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

    const size_t smem = (size_t)kIcpWarps * 2 * 4 * 32 * sizeof(float4);      // s_stage
    PW_CUDA(cudaFuncSetAttribute(icp_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_persistent_kernel, kIcpThreads, smem));
    if (occ < 1) { set_error(ctx, "icp: kernel does not fit on an SM"); return PWICP_ERR_CUDA; }
    const long nb = ((long)n + 31) / 32;
    // enough CTAs that every warp can own a batch, never more than can be co-resident
    long want = (nb + kIcpWarps - 1) / kIcpWarps;
    int grid = (int)std::min<long>((long)occ * ctx->num_sms, std::max<long>(1, want));

    PW_TRY(ctx->icp_work.reserve(ctx, (size_t)n * sizeof(float4)));
    // reduction levels: level 0 = one entry per batch, then up to two levels that sum kFanIn
    // consecutive entries each; the top level (any length) is summed in order by every CTA
    int lcount[kMaxRedLevels], nlevels = 1;
    lcount[0] = (int)nb;
    while (nlevels < kMaxRedLevels && lcount[nlevels - 1] > kFanIn) {
        lcount[nlevels] = (lcount[nlevels - 1] + kFanIn - 1) / kFanIn;
        ++nlevels;
    }
    size_t part_entries = 0, done_entries = 0;
    for (int l = 0; l < nlevels; ++l) part_entries += (size_t)lcount[l];
    if (nlevels > 2) done_entries = (size_t)lcount[2];
    const size_t bytes_part = part_entries * kNumVals * sizeof(double);
    const size_t bytes_cnt = ((size_t)prm.max_iter * (kHandoutLanes * 32 + 1) + done_entries) * sizeof(int);
    PW_TRY(ctx->icp_partials.reserve(ctx, bytes_part + bytes_cnt + 64));
    const size_t out_bytes = 64 + 16 + (size_t)prm.max_iter * (8 + 64);
    PW_TRY(ctx->icp_out.reserve(ctx, out_bytes));
    char* ob = ctx->icp_out.as<char>();
    PW_CUDA(cudaMemsetAsync(ob, 0, out_bytes, ctx->stream));
    if (idx_trace) PW_TRY(ctx->icp_idx.reserve(ctx, (size_t)prm.max_iter * n * sizeof(int)));

    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    const bool have_seed = ctx->icp_seed_valid && ctx->icp_seed.p != nullptr;
    PW_TRY(icp_sort_source(ctx, n, have_seed));
    ctx->icp_seed_valid = false;                       // seeds belong to one source set
    if (!have_seed) {
        icp_seed_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->tgt.dev, ctx->tgt_aux.as<float4>(),
                                                                  ctx->icp_sorted.as<float4>(), n,
                                                                  ctx->icp_match.as<float4>(), ctx->icp_match.as<float4>() + n,
                                                                  reinterpret_cast<int4*>(ctx->icp_match.as<float4>() + 2 * (size_t)n));
        ctx->launches++;
    }

    IcpArgs a;
    a.g = ctx->tgt.dev;
    a.aux = ctx->tgt_aux.as<float4>();
    a.src = ctx->icp_sorted.as<float4>();
    a.work = ctx->icp_work.as<float4>();
    a.cn0 = ctx->icp_match.as<float4>();
    a.cq0 = a.cn0 + n;
    a.cmore = reinterpret_cast<int4*>(a.cn0 + 2 * (size_t)n);
    a.seed_exact = 1;         // classification matches (outer.cu) or icp_seed_kernel
    a.slack = PWICP_SLACK_CELLS / ctx->tgt.dev.lv[0].inv_h;
    a.build_step2 = (PWICP_BUILD_FRAC * a.slack) * (PWICP_BUILD_FRAC * a.slack);

    a.n = n;
    a.max_iter = prm.max_iter;
    a.force_iters = prm.force_iters;
    a.rot_thr = prm.rot_thr_default ? 0.99999 : (1.0 - prm.tf_eps);
    a.transl_thr = prm.tf_eps;
    a.mse_rel = prm.fit_eps;
    a.mse_abs = 1e-12;
    {
        double* pp = ctx->icp_partials.as<double>();
        int* cc = reinterpret_cast<int*>(ctx->icp_partials.as<char>() + bytes_part);
        PW_CUDA(cudaMemsetAsync(cc, 0, bytes_cnt, ctx->stream));
        a.batch_counter = cc;
        cc += (size_t)prm.max_iter * kHandoutLanes * 32;
        a.fallbacks = cc;
        cc += prm.max_iter;
        a.done = cc;
        for (int l = 0; l < kMaxRedLevels; ++l) { a.part[l] = nullptr; a.count[l] = 0; }
        for (int l = 0; l < nlevels; ++l) {
            a.part[l] = pp; pp += (size_t)lcount[l] * kNumVals;
            a.count[l] = lcount[l];
        }
        a.nlevels = nlevels;
        a.top_part = a.part[nlevels - 1];
        a.top_count = lcount[nlevels - 1];
    }
    a.out_T = reinterpret_cast<float*>(ob);
    a.out_state = reinterpret_cast<int*>(ob + 64);
    a.mse_trace = mse_trace ? reinterpret_cast<double*>(ob + 80) : nullptr;
    a.T_trace = T_trace ? reinterpret_cast<float*>(ob + 80 + (size_t)prm.max_iter * 8) : nullptr;
    a.idx_trace = idx_trace ? ctx->icp_idx.as<int>() : nullptr;
    a.timing = nullptr;
#ifdef PWICP_TIMING
    static long long* d_timing = nullptr;
    if (!d_timing) cudaMalloc(&d_timing, (1024 * 16 + 128) * sizeof(long long));
    cudaMemsetAsync(d_timing, 0, (1024 * 16 + 128) * sizeof(long long), ctx->stream);
    a.timing = d_timing;
#endif

    void* kargs[] = {(void*)&a};
    PW_CUDA(cudaEventRecord(ctx->ev2, ctx->stream));
    PW_CUDA(cudaLaunchCooperativeKernel((void*)icp_persistent_kernel, dim3(grid), dim3(kIcpThreads), kargs, smem, ctx->stream));
    ctx->launches++;
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));

    struct { float T[16]; int st[4]; } host;
    PW_CUDA(cudaMemcpyAsync(&host, ob, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f, kms = 0.f;
    PW_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    PW_CUDA(cudaEventElapsedTime(&kms, ctx->ev2, ctx->ev1));
    ctx->last_ms = ms;
#ifdef PWICP_TIMING
    {
        std::vector<long long> ht(1024 * 16 + 128);
        cudaMemcpy(ht.data(), d_timing, ht.size() * 8, cudaMemcpyDeviceToHost);
        // per-CTA deltas against its own first stamp (SM clocks are not synchronised)
        double sum[16] = {0}, mx[16] = {0}, mn[16]; for (double& m : mn) m = 1e30;
        int cnt = 0;
        for (int b = 0; b < grid; ++b) {
            if (!ht[b * 16]) continue;
            ++cnt;
            for (int k = 0; k < 9; ++k) { double v = (double)(ht[b * 16 + k] - ht[b * 16]); sum[k] += v; mx[k] = std::max(mx[k], v); mn[k] = std::min(mn[k], v); }
        }
        printf("FALLBACK lanes (cache builds) per iteration:");
        for (int k = 0; k < std::min(64, host.st[0]); ++k) printf(" %lld(%lld)", ht[16 * 1024 + 2 * k], ht[16 * 1024 + 2 * k + 1]);
        printf("\n");
        if (cnt) { printf("TIMING it=30 n=%d cycles since iteration start (min/avg/max over %d CTAs):", n, cnt); for (int k : {1, 2, 3, 4, 8, 6, 7, 5}) printf(" [%d] %.0f/%.0f/%.0f", k, mn[k], sum[k] / cnt, mx[k]); printf("\n"); }
    }
#endif
    const int n_iter = host.st[0];
    if (T16) for (int k = 0; k < 16; ++k) T16[k] = host.T[k];
    if (res) {
        res->n_iter = n_iter; res->conv_state = host.st[1];
        res->grid_blocks = grid; res->warps_per_block = kIcpWarps; res->group_batches = kFanIn;
        res->device_ms = ms; res->correspondences = (long long)n_iter * n;
        res->kernel_ms = kms; res->reserved0 = 0.f;
    }
    if (mse_trace) PW_CUDA(cudaMemcpy(mse_trace, ob + 80, (size_t)n_iter * 8, cudaMemcpyDeviceToHost));
    if (T_trace) PW_CUDA(cudaMemcpy(T_trace, ob + 80 + (size_t)prm.max_iter * 8, (size_t)n_iter * 64, cudaMemcpyDeviceToHost));
    if (idx_trace) PW_CUDA(cudaMemcpy(idx_trace, ctx->icp_idx.p, (size_t)n_iter * n * sizeof(int), cudaMemcpyDeviceToHost));
    return PWICP_OK;
}

}  // namespace pwicp
