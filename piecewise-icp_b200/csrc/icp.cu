// icp.cu -- the whole inner point-to-plane ICP loop as ONE persistent cooperative kernel.
//
// Replaces P2PICPwithPatchNormal (reference src/Registration.cpp:1255-1269), i.e.
// pcl::IterativeClosestPointWithNormals::align with TransformationEstimationPointToPlaneLLS and
// DefaultConvergenceCriteria.
//
// Layout (round 2).  One CTA of 16 warps per SM.  The 32-point batches of the spatially ordered source
// are dealt out round robin to the warps of the whole grid (batch b belongs to warp b mod NWT) and stay with that
// warp for the whole loop: every piece of per-point state -- the transformed point, its matched target and normal,
// the candidate cache -- is read and written by one thread only, so no iteration needs a grid-wide fence for it, and
// every warp samples the whole cloud, so the data-dependent cost of the search iterations is balanced without a
// hand-out counter.  Per inner iteration every lane (a) applies the previous incremental transform to its point in
// place (float, pcl::transformPointCloudWithNormals order), (b) confirms the cached match as the exact nearest
// target (nn_search.cuh, "candidate cache") or runs the seeded ball search, (c) forms the 7 float row terms of the
// LLS system and adds its 27 products + d2 to 28 double accumulators of its own (DFMA; a product of two float
// values is exact in double, so fma(a, b, acc) is the separately rounded acc + a*b of the reference).  At the end
// of the iteration a warp folds its 32 lanes with a reduce-scatter butterfly (partners 16, 8, 4, 2, 1), the CTA
// adds its warps in order and posts the sum as tagged packets (no grid barrier: the readers poll for the tag), every CTA
// adds the CTA sums in the same fixed order, solves the 6x6
// system, builds the float transform and evaluates the convergence criteria -- identical instructions on identical
// inputs, so all CTAs take the same decision without a second barrier or a host round trip.
//
// Why not the FP64 tensor cores (round 1 formed the batch sums by DMMA.8x8x4): profiles/r02a_hw_rates.txt -- on
// B200 a DMMA.8x8x4 issues every 3.9 cycles per SM and a warp-wide DFMA every 0.55, i.e. the same FLOP rate, but
// the 8x8 tile spends 64 outputs on 28 sums and needs 16 F2F.F64.F32 per batch (2 cycles per SM each, the slowest
// instruction of the loop) against 8 for the per-lane sums: 64 against 31 FP64-pipe cycles per batch.
// The summation order ("reduction geometry") is fixed by (n, gridDim.x, warps per CTA) alone and the oracle's
// reduce_mode = 2 reproduces it for the bit-exact whole-loop parity tests.
#include "common.cuh"

#include "nn_search.cuh"
#include "small_algebra.cuh"

// candidate cache (nn_search.cuh), in level-0 cells: how far beyond the match distance targets are collected when a
// cache is built, how close to the match distance a target must be to be cached with the match, and the largest last
// step (L1) after which a cache is still built (a point that has just moved several cells will move again)
#ifndef PWICP_COLLECT_CELLS
#define PWICP_COLLECT_CELLS 0.1f
#endif
#ifndef PWICP_TIE_CELLS
#define PWICP_TIE_CELLS 0.0005f
#endif
#ifndef PWICP_RESEARCH_ALWAYS_BLOCK
#define PWICP_RESEARCH_ALWAYS_BLOCK 1     // icp_research_kernel: one code path for fresh and stale seeds (A/B switch)
#endif
#ifndef PWICP_BUILD_STEP_CELLS
#define PWICP_BUILD_STEP_CELLS 1e30f
#endif

namespace pwicp {

constexpr int kSplitMinPoints = 500000;   // icp_enqueue: source sets from this size on run iteration 1's search as its own kernel
constexpr int kMoreBit = 0x40000000;      // in cq[].w: further cached candidates in cmore[]
constexpr float kPadMargin = 1e18f;       // margin of a padding point: its square is finite, no step ever uses it up

// target-only prefix of the point-to-plane residual: ((nx*dx + ny*dy) + nz*dz), float, no FMA (-fmad=false)
__device__ __forceinline__ float nq_dot(float nx, float ny, float nz, float qx, float qy, float qz) {
    return nx * qx + ny * qy + nz * qz;
}

struct IcpArgs {
    GridDev g;
    const float4* aux;        // level-0 order: nx, ny, nz, ctstd
    const float4* src;        // source set in processing order (read only): x, y, z, w = index in the caller's order
    float4* work;             // x, y, z: transformed copy, updated in place every iteration; w = what is left of the
                              // validity radius of the point's candidate cache after the path it has travelled since
                              // the cache was built (<= 0: no cache)
    // per source point: the candidate cache (nn_search.cuh); rewritten only when the match changes
    float4* cq;               // matched target INLINE: x, y, z, w = its level-0 position (int bits) | kMoreBit when
                              // cmore[] holds further cached candidates
    float4* cn;               // its normal INLINE: x, y, z, w = (nx*qx + ny*qy) + nz*qz, the target-only prefix of the
                              // residual expression of TransformationEstimationPointToPlaneLLS (float, unfused, in order)
    int4* cmore;              // x, y, z: positions of up to three further candidates (unused = primary), w = original index of the primary
    // all per-point arrays are padded to a multiple of 32 points: a pad has a zero normal (every row term and product is
    // exactly zero), a huge margin (never searches) and is excluded from the sum of squared distances
    int seed_exact;           // cq[] is the exact NN of the untransformed source (iteration 0 needs no search)
    float collect;            // targets within this of the match distance are looked at when a cache is built
    float tie;                // ... and those within this of the match distance are cached with it
    float tie_step;           // ... or within this fraction of the step the query has just taken, if that is more
    float build_step;         // a cache is built only when the point moved less than this in the last step (L1 length)
    int n;                    // number of source points ...
    const int* n_dev;         // ... or, when not null, where the device holds it (outer loop: the stable set is counted
                              // on the device; the launch geometry then comes from the capacity n)
    int max_iter;
    int force_iters;
    // the loop runs as two launches with a stand-alone search in between (icp_research_kernel): this launch does the
    // iterations [it_begin, it_end); it_begin > 0: the points in work[] are already where this iteration wants them and
    // carry their fresh caches, the state of the loop so far comes from *carry
    int it_begin, it_end;
    struct IcpCarry* carry;
    double rot_thr, transl_thr, mse_rel, mse_abs;
    unsigned long long* part; // [2][gridDim.x][28][2]: CTA sums of an iteration as tagged 8-byte packets, double buffered
    int* searched;            // [max_iter], zeroed before the launch: queries that ran the ball search (diagnostic)
    float* out_T;             // 16: final transformation
    int* out_state;           // [0] n_iter, [1] conv_state, [2] first iteration at which the criteria were met, [3] its state
    double* mse_trace;        // nullable
    float* T_trace;           // nullable
    int* idx_trace;           // nullable, [iter][n]
    unsigned long long* iter_ns;   // [max_iter + 1]: %globaltimer at the start of the loop and after every iteration (CTA 0)
    unsigned long long* phase_ns;  // [max_iter][4]: CTA 0, warp 0: end of its batches, CTA sum posted, totals formed (all packets in), solved
};

// What a launch that ends before the loop does hands to the next one (written by CTA 0).
struct IcpCarry {
    float T[16];         // the incremental transform of the last iteration done
    float Tfinal[16];    // the accumulated one
    double prev_mse;
    int met;             // the convergence criteria have been met before
    int stop;            // conv_state != 0: the loop is over, the launches that follow return at once
};

// Shared scratch of the per-iteration solve.
struct FinishSmem {
    double A[6][6];      // ATA, then its LU factors in place
    double inv[6][6];
    double b[6], x[6], rcp[6];
    double sc[3][2];     // sin / cos of alpha, beta, gamma
    double prev_mse;
    int piv[6];
    int met;             // the convergence criteria have been met before (force_iters keeps looping)
    float Tn[16];
};

// Warp 0 of every CTA: totals -> 6x6 solve -> float transform -> convergence decision.
// Every scalar operation is the one small_algebra.cuh's sequential inverse6()/solve_from28()
// performs (same operands, same order per element), so the result is bit-identical to the
// single-thread version and to the oracle; the lanes only shorten the critical path.
// This code runs once per inner iteration on one warp, i.e. always from a cold instruction cache
// (profiles/r01d_*: a straight-line version spent ~10 us mostly fetching instructions), hence the
// rolled loops and the out-of-line placement: few instruction lines, re-used from L0.
static __device__ __noinline__ int icp_finish_warp(const IcpArgs& a, int n, int it, const double* s_tot, float* s_T,
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
    if (lane < 3) sincos(F.x[lane], &F.sc[lane][0], &F.sc[lane][1]);
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
        const double mse = s_tot[27] / (double)n;
        const double prev_mse = F.prev_mse;
        const int iters = it + 1;
        // DefaultConvergenceCriteria<float>::hasConverged(), in PCL's order
        // force_iters (benchmark mode): the criteria are evaluated and recorded but may not stop the loop
        int crit = 0;
        if (iters >= a.max_iter) crit = PWICP_CONV_ITERATIONS;
        else {
            const double cos_angle = 0.5 * (double)(Tn[0] + Tn[5] + Tn[10] - 1.0f);
            const double transl_sq = (double)(Tn[3] * Tn[3] + Tn[7] * Tn[7] + Tn[11] * Tn[11]);
            if (cos_angle >= a.rot_thr && transl_sq <= a.transl_thr) crit = PWICP_CONV_TRANSFORM;
            else if (fabs(mse - prev_mse) < a.mse_abs) crit = PWICP_CONV_ABS_MSE;
            else if (fabs(mse - prev_mse) / prev_mse < a.mse_rel) crit = PWICP_CONV_REL_MSE;
        }
        state = (crit == PWICP_CONV_ITERATIONS || !a.force_iters) ? crit : 0;
        if (crit && !F.met) {
            F.met = 1;
            if (blockIdx.x == 0) { a.out_state[2] = iters; a.out_state[3] = crit; }
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

// Everything a query can need beyond its inline match, out of line on purpose: the steady-state loop has to stay small
// (instruction cache, together with the per-iteration solve: profiles/r01e_*) and, above all, keep its 28 double
// accumulators in registers -- inlined, the live values of this path spilled them in every trip (profiles/r02j).
//   (1) further cached candidates (cq[].w has kMoreBit): best of the cached set under the tie rule, promoted to the
//       inline slot when it overtakes the match;
//   (2) the certificate fails: seeded ball search and a new cache around the position (nn_search.cuh).
struct SlowOut {
    float d2;             // squared distance of the match
    int pos;              // its level-0 position
    float qx, qy, qz;     // the match
    float nx, ny, nz, nq; // its normal, nq_dot(normal, match)
    float margin;         // what is left of the cache radius
};
static __device__ __noinline__ SlowOut icp_slow_path(const IcpArgs& a, int it, int i, float px, float py, float pz,
                                                     float4 q0, float4 cn, float margin, float mchk, float step) {
    const float4* __restrict__ pts = a.g.lv[0].pts;
    const int pos0 = __float_as_int(q0.w) & ~kMoreBit;
    SlowOut o;
    o.d2 = l2_simple(px, py, pz, q0.x, q0.y, q0.z);
    o.pos = pos0; o.qx = q0.x; o.qy = q0.y; o.qz = q0.z;
    o.nx = cn.x; o.ny = cn.y; o.nz = cn.z; o.nq = cn.w;
    o.margin = margin;
    // how close to the match a target must be to be cached with it: a fraction of the step the query has just taken
    // (the next one is smaller), at least `tie`, at most the collection radius
    const float tie = fminf(fmaxf(a.tie_step * step, a.tie), a.collect);
    if (__float_as_int(q0.w) & kMoreBit) {
        // positions from the side array, the three loads are issued together, unused slots repeat the primary
        const int4 cm = __ldcg(a.cmore + i);
        int bidx = cm.w;
        float second = __int_as_float(0x7f800000);   // squared distance of the second-nearest cached target
        float4 q1 = q0, q2 = q0, q3 = q0;
        if (cm.x != pos0) q1 = __ldg(pts + cm.x);
        if (cm.y != pos0) q2 = __ldg(pts + cm.y);
        if (cm.z != pos0) q3 = __ldg(pts + cm.z);
#define PW_CAND(q, cp)                                                                         \
        if ((cp) != pos0) {                                                                    \
            const float d = l2_simple(px, py, pz, q.x, q.y, q.z);                              \
            const int id = __float_as_int(q.w);                                                \
            if (d < o.d2 || (d == o.d2 && id < bidx)) {                                        \
                second = o.d2;                                                                 \
                o.d2 = d; bidx = id; o.pos = (cp); o.qx = q.x; o.qy = q.y; o.qz = q.z;         \
            } else second = fminf(second, d);                                                  \
        }
        PW_CAND(q1, cm.x) PW_CAND(q2, cm.y) PW_CAND(q3, cm.z)
#undef PW_CAND
        if (mchk > 0.f && o.d2 * 1.00003f < mchk * mchk) {
            // the cache answers.  Every target within `margin` of this position is cached (the ball lies inside the
            // cache ball around the anchor), so a NEW cache can be cut out of the old one without a search: this
            // position as the anchor, the match alone, radius just below the second-nearest cached target.  Done as
            // soon as that target is clear of the match: the query drops its side list and leaves this path.
            const float d1 = sqrtf(o.d2);
            const bool alone = second > (d1 + tie) * (d1 + tie);
            if (o.pos != pos0 || alone) {
                if (o.pos != pos0) {
                    const float4 nv = __ldg(a.aux + o.pos);
                    o.nx = nv.x; o.ny = nv.y; o.nz = nv.z;
                    o.nq = nq_dot(o.nx, o.ny, o.nz, o.qx, o.qy, o.qz);
                    a.cn[i] = make_float4(o.nx, o.ny, o.nz, o.nq);
                }
                a.cq[i] = make_float4(o.qx, o.qy, o.qz, __int_as_float(o.pos | (alone ? 0 : kMoreBit)));
                if (!alone)      // another cached target has become the nearest: it is the primary now
                    a.cmore[i] = make_int4(cm.x == o.pos ? pos0 : cm.x, cm.y == o.pos ? pos0 : cm.y,
                                           cm.z == o.pos ? pos0 : cm.z, bidx);
            }
            if (alone) o.margin = fminf(mchk, sqrtf(second) * 0.9999f);
            return o;
        }
    }
    // |p - anchor| <= path (triangle inequality over the steps actually taken): every target at least as close to p
    // as the best cached one lies within the cache radius of the anchor; compared as squares
    if (mchk > 0.f && o.d2 * 1.00003f < mchk * mchk) return o;

    {   // queries that needed the search this iteration (diagnostic: pwicp_icp_profile)
        const unsigned m = __activemask();
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(a.searched + it, __popc(m));
    }
    const Best bb = nn_search_seeded<true>(a.g, px, py, pz, o.pos);
    const float4 nq = __ldg(a.aux + bb.pos);
    // a new cache around this position (nn_search.cuh, "candidate cache"): the match, the targets within `tie` of it,
    // and the radius just below the first target left out
    int o1 = bb.pos, o2 = bb.pos, o3 = bb.pos, k = 0;
    float rho = 0.f;
    const float d1 = sqrtf(bb.d2);
    if (step < a.build_step) {
        // first with the wide radius (a large gap to the second-nearest target = a long-lived cache); a query far from
        // its match would need more than the 3x3-row scan for that: then only the ties are looked for
        for (int attempt = 0; attempt < 2; ++attempt) {
            const float R = d1 + (attempt == 0 ? a.collect : fminf(4.0f * tie, a.collect));
            if (R * a.g.lv[0].inv_h >= 0.95f) continue;
            const Near5 nb = ball_collect(a.g.lv[0], a.g.ox, a.g.oy, a.g.oz, px, py, pz, R * R);
            if (!nb.complete) continue;
            const float lim = (d1 + tie) * (d1 + tie);
            float rho2 = R * R;                      // complete up to the scanned radius unless a target is left out
            bool open = true;                        // still taking targets into the cache
#pragma unroll
            for (int j = 0; j <= kCacheCands; ++j) {
                if (!open || nb.pos[j] < 0) continue;
                if (nb.pos[j] == bb.pos) continue;   // the match itself: always cached (the primary)
                if (nb.d2[j] < lim && k < 3) {
                    if (k == 0) o1 = nb.pos[j]; else if (k == 1) o2 = nb.pos[j]; else o3 = nb.pos[j];
                    ++k;
                } else { rho2 = nb.d2[j]; open = false; }
            }
            rho = sqrtf(rho2) * 0.9999f;
            break;
        }
    }
    o.d2 = bb.d2; o.pos = bb.pos; o.qx = bb.qx; o.qy = bb.qy; o.qz = bb.qz;
    o.nx = nq.x; o.ny = nq.y; o.nz = nq.z;
    o.nq = nq_dot(nq.x, nq.y, nq.z, bb.qx, bb.qy, bb.qz);
    o.margin = rho;
    a.cq[i] = make_float4(bb.qx, bb.qy, bb.qz, __int_as_float(bb.pos | (k > 0 ? kMoreBit : 0)));
    a.cn[i] = make_float4(nq.x, nq.y, nq.z, o.nq);
    if (k > 0) a.cmore[i] = make_int4(o1, o2, o3, bb.idx);
    return o;
}

// 16-byte asynchronous copy global -> shared (LDGSTS through L2), per-thread completion groups.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group2() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }
// ---- bulk-copy (TMA) staging of the per-batch streams, PWICP_STAGE_TMA = 1 ------------------------------------------
// One elected lane issues three 512-byte cp.async.bulk copies per batch (SASS UBLKCP) that complete on the slot's
// mbarrier; the warp waits on the barrier's phase instead of a per-thread cp.async group.  A/B against the per-lane
// LDGSTS copies: profiles/r02t_tma_ab.txt.
#ifndef PWICP_STAGE_TMA
#define PWICP_STAGE_TMA 0
#endif
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
}

// 16-byte load from the shared window (volatile: never hoisted out of the loop, never cached in registers)
__device__ __forceinline__ void lds128(float4& v, unsigned addr) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Reduce-scatter butterfly over the lanes of a warp.  One step with partner distance kS: the lanes whose bit kS is
// clear keep entries [0, kS) of their array, the others [kS, 2 kS); every lane adds what its partner (lane ^ kS)
// held of the kept half.  After the steps 16, 8, 4, 2, 1 lane v holds the sum over all lanes of entry v, each
// formed as the balanced tree ((x[l] + x[l + 16]) + ...) -- addition commutes bit for bit, so it does not matter
// which partner does the add.  The first step takes the 28 sums (entries 28..31 are implicit zeros).
__device__ __forceinline__ void fold_lanes16(const double (&x)[kNumVals], double (&y)[16], int lane) {
    const bool upper = (lane & 16) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const double hi = (j + 16 < kNumVals) ? x[(j + 16 < kNumVals) ? j + 16 : 0] : 0.0;
        const double keep = upper ? hi : x[j];
        const double send = upper ? x[j] : hi;
        y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
}
template <int kS>
__device__ __forceinline__ void fold_lanes(double (&x)[16], int lane) {
    const bool upper = (lane & kS) != 0;
#pragma unroll
    for (int j = 0; j < kS; ++j) {
        const double keep = upper ? x[j + kS] : x[j];
        const double send = upper ? x[j] : x[j + kS];
        x[j] = keep + __shfl_xor_sync(0xffffffffu, send, kS);
    }
}

// The grid-wide exchange of the CTA sums, without a grid barrier.  Every CTA sum travels as two 8-byte packets {half of
// the double, iteration tag}: an aligned 8-byte store is single-copy atomic, so a reader that sees the tag sees the data
// with it and no fence or counter is needed (the LL protocol of collective libraries).  The buffers alternate with the
// iteration parity; a CTA can run at most one iteration ahead of the slowest one (it needs everybody's packets of
// iteration k + 1 before it can post k + 2 into the buffer of k), so a packet is never overwritten before it has been
// read.  The buffers are zeroed before the launch (tag 0 = nothing posted).  All CTAs are resident (cooperative launch).
// In: s_wsum, the warp sums of this CTA (after a CTA barrier).  Out: s_csum[w] = sum of the CTA sums of chunk w, in order.
static __device__ __noinline__ void icp_exchange(const IcpArgs& a, int it, int warp, int lane, const double* s_wsum, double* s_csum) {
    const int G = gridDim.x;
    const int per = (G + kIcpWarps - 1) / kIcpWarps;             // CTA sums per chunk of the final sum
    const int nchunks = (G + per - 1) / per;
    const unsigned tag = (unsigned)it + 1u;
    if (lane >= kNumVals) return;
    if (warp == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kIcpWarps; ++w) s += s_wsum[w * kNumVals + lane];
        unsigned long long* dst = a.part + (((size_t)(it & 1) * G + blockIdx.x) * kNumVals + lane) * 2;
        const unsigned long long t = (unsigned long long)tag << 32;
        st_relaxed_u64(dst, t | (unsigned)__double2loint(s));
        st_relaxed_u64(dst + 1, t | (unsigned)__double2hiint(s));
        if (blockIdx.x == 0 && lane == 0) a.phase_ns[it * 4 + 1] = globaltimer_ns();
    }
    // every CTA forms the same totals: warp w adds the CTA sums w*per .. in order (warp 0 adds the chunks afterwards)
    if (warp < nchunks) {
        const unsigned long long* __restrict__ src = a.part + (((size_t)(it & 1) * G + (size_t)warp * per) * kNumVals + lane) * 2;
        const int cnt = min(per, G - warp * per);
        double s = 0.0;
        for (int g0 = 0; g0 < cnt; g0 += 10) {                   // loads of a round are issued before its first add
            unsigned long long v[20];
            bool ok;
            do {
                ok = true;
#pragma unroll
                for (int g = 0; g < 10; ++g) {
                    const unsigned long long* q = src + (size_t)min(g0 + g, cnt - 1) * (kNumVals * 2);
                    v[2 * g] = ld_relaxed_u64(q); v[2 * g + 1] = ld_relaxed_u64(q + 1);
                }
#pragma unroll
                for (int g = 0; g < 20; ++g) ok = ok && ((unsigned)(v[g] >> 32) == tag);
            } while (!ok);
#pragma unroll
            for (int g = 0; g < 10; ++g)
                if (g0 + g < cnt) s += __hiloint2double((int)(unsigned)v[2 * g + 1], (int)(unsigned)v[2 * g]);
        }
        s_csum[warp * kNumVals + lane] = s;
    }
}

// kTrace: the variant that also records the correspondence indices of every iteration (parity tests).
template <bool kTrace>
__global__ void __launch_bounds__(kIcpThreads, 1) icp_persistent_kernel(const IcpArgs a) {
    // landing zone of the streamed per-point data, kStageSlots batches deep per warp, filled by cp.async; every lane
    // reads back only the 16-byte cells it copied itself: [warp][slot][point + margin, matched target, normal][lane]
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ double s_wsum[kIcpWarps][kNumVals];   // warp sums of the iteration
    __shared__ double s_csum[kIcpWarps][kNumVals];   // chunk sums of the CTA sums of the grid (a second array: the warps
                                                     // start polling for the packets while warp 0 still adds the warp sums)
    __shared__ double s_tot[kNumVals];
    __shared__ float s_T[16];
    __shared__ float s_Tfinal[16];
    __shared__ int s_stop;
    __shared__ FinishSmem s_fin;
#if PWICP_STAGE_TMA
    __shared__ __align__(8) unsigned long long s_bar[kIcpWarps][kStageSlots];
#endif

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n_dev ? __ldg(a.n_dev) : a.n;                // same value in every thread of the grid
    if (n < 3) {                                                 // pcl: min_number_correspondences_ (device-side count only)
        if (a.it_begin > 0) return;
        if (blockIdx.x == 0 && tid == 0 && a.carry) a.carry->stop = PWICP_CONV_NO_CORR;
        if (blockIdx.x == 0 && tid < 16) a.out_T[tid] = (tid % 5 == 0) ? 1.0f : 0.0f;
        if (blockIdx.x == 0 && tid == 0) { a.out_state[0] = 0; a.out_state[1] = PWICP_CONV_NO_CORR; }
        return;
    }
    const int nb = (n + 31) / 32;                                // 32-point batches
    const int G = gridDim.x, NWT = G * kIcpWarps, ws = blockIdx.x * kIcpWarps + warp;
    const int K = (ws < nb) ? (nb - 1 - ws) / NWT + 1 : 0;       // batches ws, ws + NWT, ... of this warp
    const int stride = NWT * 32;                                 // distance of this lane's consecutive points
    const int i0 = ws * 32 + lane;
    constexpr int kSlotBytes = 3 * 32 * (int)sizeof(float4);
    // shared-window address of this lane's landing cells, pinned in a register (an opaque move: otherwise the compiler
    // re-derives it from the thread index in every trip of the loop, 15 instructions)
    unsigned s_lane_sh;
    {
        const unsigned v = (unsigned)__cvta_generic_to_shared(s_dyn + (size_t)warp * (kStageSlots * kSlotBytes)) + lane * 16;
        asm volatile("mov.u32 %0, %1;" : "=r"(s_lane_sh) : "r"(v));
    }

    if (a.it_begin == 0) {
        if (tid < 16) { s_T[tid] = (tid % 5 == 0) ? 1.0f : 0.0f; s_Tfinal[tid] = s_T[tid]; }
        if (tid == 0) { s_stop = 0; s_fin.prev_mse = 1.7976931348623157e308; s_fin.met = 0; }   // DBL_MAX
        if (blockIdx.x == 0 && tid == 0) a.iter_ns[0] = globaltimer_ns();
    } else {
        if (__ldcg(&a.carry->stop)) return;                      // the loop ended in the previous launch (every thread)
        if (tid < 16) { s_T[tid] = __ldcg(&a.carry->T[tid]); s_Tfinal[tid] = __ldcg(&a.carry->Tfinal[tid]); }
        if (tid == 0) { s_stop = 0; s_fin.prev_mse = __ldcg(&a.carry->prev_mse); s_fin.met = __ldcg(&a.carry->met); }
    }
    __syncthreads();

    // copies of point i_ of this lane (one of the batches of this warp; past the end: an empty group): point +
    // margin, matched target, its normal
    const int n_pad = nb * 32;
#if PWICP_STAGE_TMA
    const unsigned s_bar_sh = (unsigned)__cvta_generic_to_shared(&s_bar[warp][0]);
    const unsigned s_warp_sh = s_lane_sh - lane * 16;            // this warp's landing zone
    if (lane == 0) {
        for (int k = 0; k < kStageSlots; ++k) mbar_init(s_bar_sh + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phases = 0;                                         // bit s: parity the next wait on slot s expects
    // one lane copies the three 512-byte rows of the batch that starts at point i_ - lane (past the end: nothing)
    auto stage_point = [&](const float4* __restrict__ psrc, int i_, int slot_off) {
        __syncwarp();                                            // every lane has read what the slot held before
        if (lane == 0 && i_ < n_pad) {
            const unsigned bar = s_bar_sh + 8 * (slot_off / kSlotBytes), d = s_warp_sh + slot_off;
            mbar_expect_tx(bar, 3 * 512);
            bulk_g2s(d, psrc + i_, 512, bar);
            bulk_g2s(d + 512, a.cq + i_, 512, bar);
            bulk_g2s(d + 1024, a.cn + i_, 512, bar);
        }
    };
#else
    auto stage_point = [&](const float4* __restrict__ psrc, int i_, int slot_off) {
        if (i_ < n_pad) {
            const unsigned d = s_lane_sh + slot_off;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(psrc + i_) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512), "l"(a.cq + i_) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 1024), "l"(a.cn + i_) : "memory");
        }
        cp_async_commit();
    };
#endif
    stage_point(a.it_begin == 0 ? a.src : a.work, i0, 0);
    stage_point(a.it_begin == 0 ? a.src : a.work, i0 + stride, kSlotBytes);
    const unsigned s_T_sh = (unsigned)__cvta_generic_to_shared(s_T);

    for (int it = a.it_begin;; ++it) {
        double acc[kNumVals];
#pragma unroll
        for (int v = 0; v < kNumVals; ++v) acc[v] = 0.0;
        const float4* __restrict__ psrc = (it == 0) ? a.src : a.work;
        const float4* __restrict__ pts = a.g.lv[0].pts;
        const bool first = it == 0;
        const bool resumed = it > 0 && it == a.it_begin;         // work[] holds this iteration's positions and caches

        // ---- phase A: this warp's batches, two batches of copies in flight ahead of the one being processed
        int slot = 0, pslot = 2 * kSlotBytes, i = i0;            // byte offsets of the ring slots
#pragma unroll 1
        for (int k = 0; k < K; ++k, i += stride) {
            stage_point(psrc, i + 2 * stride, pslot);
#if PWICP_STAGE_TMA
            {
                const int sidx = slot / kSlotBytes;
                mbar_wait(s_bar_sh + 8 * sidx, (phases >> sidx) & 1u);
                phases ^= 1u << sidx;
            }
#else
            cp_async_wait_group2();                              // everything but the two youngest groups has landed
#endif
            float4 p, q0, cn;
            lds128(p, s_lane_sh + slot);
            lds128(q0, s_lane_sh + slot + 512);
            lds128(cn, s_lane_sh + slot + 1024);
            float step = __int_as_float(0x7f800000);             // L1 length of the last step (>= its Euclidean length)
            float margin;                                        // cache radius left after the path travelled (lower bound)
            float mchk;                                          // ... as the certificate below uses it
            if (resumed) {
                margin = p.w;
                mchk = margin;
            } else if (!first) {
                // the transform is read from shared memory where it is used (three broadcast loads): twelve more
                // live registers across the loop would spill the accumulators
                float4 r0, r1, r2;
                lds128(r0, s_T_sh); lds128(r1, s_T_sh + 16); lds128(r2, s_T_sh + 32);
                const float x = r0.x * p.x + r0.y * p.y + r0.z * p.z + r0.w;      // xform_point (small_algebra.cuh)
                const float y = r1.x * p.x + r1.y * p.y + r1.z * p.z + r1.w;
                const float z = r2.x * p.x + r2.y * p.y + r2.z * p.z + r2.w;
                step = (fabsf(x - p.x) + fabsf(y - p.y)) + fabsf(z - p.z);
                margin = __fmaf_rn(step, -1.00001f, p.w);
                mchk = margin;
                p.x = x; p.y = y; p.z = z;
            } else {
                margin = (i < n) ? 0.f : kPadMargin;              // no cache yet; pads never search
                mchk = (a.seed_exact || i >= n) ? kPadMargin : 0.f;
            }
            // (b) exact NN: the inline match when the candidate cache still covers the query and holds nothing else
            // (nn_search.cuh, "candidate cache"); further cached candidates and the seeded ball search are out of line
            float bd2 = l2_simple(p.x, p.y, p.z, q0.x, q0.y, q0.z);
            float qx = q0.x, qy = q0.y, qz = q0.z;
            float nx = cn.x, ny = cn.y, nz = cn.z, nq = cn.w;
            int bpos = __float_as_int(q0.w);
            if ((bpos & kMoreBit) || !(mchk > 0.f && bd2 * 1.00003f < mchk * mchk)) {
                const SlowOut o = icp_slow_path(a, it, i, p.x, p.y, p.z, q0, cn, margin, mchk, step);
                bd2 = o.d2; bpos = o.pos; qx = o.qx; qy = o.qy; qz = o.qz;
                nx = o.nx; ny = o.ny; nz = o.nz; nq = o.nq;
                margin = o.margin;
            }
            a.work[i] = make_float4(p.x, p.y, p.z, margin);
            // float expressions of TransformationEstimationPointToPlaneLLS (no FMA); nq = (nx*dx + ny*dy) + nz*dz
            const float u0 = nz * p.y - ny * p.z;
            const float u1 = nx * p.z - nz * p.x;
            const float u2 = ny * p.x - nx * p.y;
            const float u6 = nq - nx * p.x - ny * p.y - nz * p.z;
            const float u7 = (i < n) ? bd2 : 0.f;                // a pad's rows are zero through its zero normal
            // traces are reported in the caller's order (src[].w = original source index)
            if (kTrace && i < n)
                a.idx_trace[(size_t)it * n + __float_as_int(__ldg(a.src + i).w)] = __float_as_int(__ldg(pts + bpos).w);
            // 27 products + d2 into this lane's sums
            {
                const double d0 = (double)u0, d1 = (double)u1, d2 = (double)u2, d3 = (double)nx, d4 = (double)ny,
                             d5 = (double)nz, d6 = (double)u6;
                acc[0] = fma(d0, d0, acc[0]);   acc[1] = fma(d0, d1, acc[1]);   acc[2] = fma(d0, d2, acc[2]);
                acc[3] = fma(d0, d3, acc[3]);   acc[4] = fma(d0, d4, acc[4]);   acc[5] = fma(d0, d5, acc[5]);
                acc[6] = fma(d1, d1, acc[6]);   acc[7] = fma(d1, d2, acc[7]);   acc[8] = fma(d1, d3, acc[8]);
                acc[9] = fma(d1, d4, acc[9]);   acc[10] = fma(d1, d5, acc[10]); acc[11] = fma(d2, d2, acc[11]);
                acc[12] = fma(d2, d3, acc[12]); acc[13] = fma(d2, d4, acc[13]); acc[14] = fma(d2, d5, acc[14]);
                acc[15] = fma(d3, d3, acc[15]); acc[16] = fma(d3, d4, acc[16]); acc[17] = fma(d3, d5, acc[17]);
                acc[18] = fma(d4, d4, acc[18]); acc[19] = fma(d4, d5, acc[19]); acc[20] = fma(d5, d5, acc[20]);
                acc[21] = fma(d0, d6, acc[21]); acc[22] = fma(d1, d6, acc[22]); acc[23] = fma(d2, d6, acc[23]);
                acc[24] = fma(d3, d6, acc[24]); acc[25] = fma(d4, d6, acc[25]); acc[26] = fma(d5, d6, acc[26]);
                acc[27] += (double)u7;
            }
            slot = (slot == (kStageSlots - 1) * kSlotBytes) ? 0 : slot + kSlotBytes;
            pslot = (pslot == (kStageSlots - 1) * kSlotBytes) ? 0 : pslot + kSlotBytes;
        }

        if (blockIdx.x == 0 && tid == 0) a.phase_ns[it * 4] = globaltimer_ns();
        // ---- warp sums (lane v ends up with value v), CTA sum over the warps in order, published for the grid
        {
            double y[16];
            fold_lanes16(acc, y, lane);
            fold_lanes<8>(y, lane);
            fold_lanes<4>(y, lane);
            fold_lanes<2>(y, lane);
            fold_lanes<1>(y, lane);
            if (lane < kNumVals) s_wsum[warp][lane] = y[0];
        }
#if PWICP_STAGE_TMA
        asm volatile("fence.proxy.async;" ::: "memory");         // work[] was written through the generic proxy
#endif
        // the next iteration's first two batches: their copies do not depend on the transform being solved
        // for, so they fly during the reduction and the solve (every slot of the ring has been consumed)
        stage_point(a.work, i0, 0);
        stage_point(a.work, i0 + stride, kSlotBytes);
        __syncthreads();
        // ---- the grid-wide exchange of the CTA sums (icp_exchange, out of line: its address arithmetic and the twenty
        // packets in flight per lane must not compete with the accumulators of the batch loop for registers)
        icp_exchange(a, it, warp, lane, &s_wsum[0][0], &s_csum[0][0]);
        __syncthreads();
        if (warp == 0) {
            if (lane < kNumVals) {
                const int per = (G + kIcpWarps - 1) / kIcpWarps, nchunks = (G + per - 1) / per;   // as icp_exchange
                double v[kIcpWarps];                            // all loads before the first add (nchunks <= kIcpWarps)
#pragma unroll
                for (int c = 0; c < kIcpWarps; ++c) v[c] = s_csum[c][lane];
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < kIcpWarps; ++c) if (c < nchunks) s += v[c];
                s_tot[lane] = s;
            }
            __syncwarp();
            if (blockIdx.x == 0 && lane == 0) a.phase_ns[it * 4 + 2] = globaltimer_ns();
            const int st = icp_finish_warp(a, n, it, s_tot, s_T, s_Tfinal, s_fin, lane);
            if (blockIdx.x == 0 && a.carry && it + 1 == a.it_end) {   // hand-over to the next launch
                if (lane < 16) { a.carry->T[lane] = s_T[lane]; a.carry->Tfinal[lane] = s_Tfinal[lane]; }
                if (lane == 0) { a.carry->prev_mse = s_fin.prev_mse; a.carry->met = s_fin.met; a.carry->stop = st; }
            }
            if (lane == 0) {
                s_stop = st;
                if (blockIdx.x == 0) { a.iter_ns[it + 1] = globaltimer_ns(); a.phase_ns[it * 4 + 3] = a.iter_ns[it + 1]; }
            }
        }
        __syncthreads();
        if (s_stop || it + 1 >= a.it_end) break;
    }
#if PWICP_STAGE_TMA
    // the copies staged for an iteration that does not run must land before the CTA gives up its shared memory
    if (K > 0) mbar_wait(s_bar_sh, phases & 1u);
    if (K > 1) mbar_wait(s_bar_sh + 8, (phases >> 1) & 1u);
#else
    cp_async_wait_all();
#endif
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

// a padding point (i >= n, up to the next multiple of 32): zero normal, see IcpArgs
__device__ __forceinline__ void write_pad(int i, float4* out, float4* cn0, float4* cq0) {
    out[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    cq0[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    cn0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void src_gather_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ order, int n, int n_pad,
                                  float4* out, const int* __restrict__ seed_in, const float4* __restrict__ tgt_pts,
                                  const float4* __restrict__ tgt_aux,
                                  float4* __restrict__ cn0, float4* __restrict__ cq0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i >= n) { write_pad(i, out, cn0, cq0); return; }
    const uint32_t o = order[i];
    float4 p = src[o];
    p.w = __int_as_float((int)o);
    out[i] = p;
    if (!seed_in) return;                 // icp_seed_kernel fills the match
    const int sd = seed_in[o];            // classification match (outer.cu): always a valid position
    const float4 q = __ldg(tgt_pts + sd), nq = __ldg(tgt_aux + sd);
    cq0[i] = make_float4(q.x, q.y, q.z, __int_as_float(sd));
    cn0[i] = make_float4(nq.x, nq.y, nq.z, nq_dot(nq.x, nq.y, nq.z, q.x, q.y, q.z));
}

// Iteration 0 of a source set without seeds: the plain search at full occupancy (the persistent
// kernel is register-capped and runs it about twice as slowly, profiles/r01f_*).  The matches go
// into the candidate slots; the persistent kernel takes them as the exact answer of iteration 0.
__global__ void __launch_bounds__(256)
icp_seed_kernel(GridDev g, const float4* __restrict__ tgt_aux, const float4* __restrict__ src, int n,
                float4* __restrict__ cn0, float4* __restrict__ cq0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float4 p = (i < n) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const Best b = nn_search_warp(g, p.x, p.y, p.z, -1, i < n);   // warp-collective (nn_search.cuh, team search)
    if (i >= n) return;
    cq0[i] = make_float4(b.qx, b.qy, b.qz, __int_as_float(b.pos));
    if (!tgt_aux) return;                  // the normals are not on the device yet: icp_fill_cn_kernel
    const float4 nq = __ldg(tgt_aux + b.pos);
    cn0[i] = make_float4(nq.x, nq.y, nq.z, nq_dot(nq.x, nq.y, nq.z, b.qx, b.qy, b.qz));
}

// ... the normals of the matches icp_seed_kernel found, once the normals have arrived (pwicp_icp_p2plane)
__global__ void __launch_bounds__(256)
icp_fill_cn_kernel(const float4* __restrict__ tgt_aux, const float4* __restrict__ cq0, int n, float4* __restrict__ cn0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = cq0[i];
    const float4 nq = __ldg(tgt_aux + __float_as_int(q.w));
    cn0[i] = make_float4(nq.x, nq.y, nq.z, nq_dot(nq.x, nq.y, nq.z, q.x, q.y, q.z));
}

// Iteration 1 of the loop searches for EVERY query (no cache exists yet; after the first, large ICP step the matches of
// iteration 0 are stale) -- 0.43 of the 1.39 ms the persistent kernel took at 1M, at 16 warps per SM and 128 registers.
// This kernel does that iteration's per-point work up to the match at full occupancy, between two launches of the
// persistent kernel: the transform of iteration 0 applied to the point (the same float expression), the seeded search,
// the candidate cache around the new position (what icp_slow_path does for a query without a cache).  The second
// launch finds work[] = the transformed point + the radius of its fresh cache and cq / cn / cmore = the match, and
// starts with the row terms of iteration 1.
__global__ void __launch_bounds__(256)
icp_research_kernel(const IcpArgs a) {
    const int n = a.n_dev ? __ldg(a.n_dev) : a.n;
    if (n < 3 || __ldcg(&a.carry->stop)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_pad = (n + 31) / 32 * 32;
    if (i >= n_pad) return;
    const float* T = a.carry->T;
    const float4 p = a.work[i];
    const float x = __ldcg(T + 0) * p.x + __ldcg(T + 1) * p.y + __ldcg(T + 2) * p.z + __ldcg(T + 3);     // xform_point (small_algebra.cuh)
    const float y = __ldcg(T + 4) * p.x + __ldcg(T + 5) * p.y + __ldcg(T + 6) * p.z + __ldcg(T + 7);
    const float z = __ldcg(T + 8) * p.x + __ldcg(T + 9) * p.y + __ldcg(T + 10) * p.z + __ldcg(T + 11);
    if (i >= n) { a.work[i] = make_float4(x, y, z, kPadMargin); return; }
    const float step = (fabsf(x - p.x) + fabsf(y - p.y)) + fabsf(z - p.z);
    const float tie = fminf(fmaxf(a.tie_step * step, a.tie), a.collect);
    // a query whose cache still covers it (second pass, before iteration 2: most of them) only moves; one with further
    // cached candidates is left to the loop (icp_slow_path, (1))
    const float margin = __fmaf_rn(step, -1.00001f, p.w);
    const float4 q0 = a.cq[i];
    if ((__float_as_int(q0.w) & kMoreBit) ||
        (margin > 0.f && l2_simple(x, y, z, q0.x, q0.y, q0.z) * 1.00003f < margin * margin)) {
        a.work[i] = make_float4(x, y, z, margin);
        return;
    }
    {
        const unsigned m = __activemask();
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(a.searched + a.it_begin, __popc(m));
    }
    const Best bb = nn_search_seeded<false>(a.g, x, y, z, __float_as_int(q0.w), PWICP_RESEARCH_ALWAYS_BLOCK != 0);
    const float4 nq = __ldg(a.aux + bb.pos);
    int o1 = bb.pos, o2 = bb.pos, o3 = bb.pos, k = 0;
    float rho = 0.f;
    const float d1 = sqrtf(bb.d2);
    if (step < a.build_step) {
        for (int attempt = 0; attempt < 2; ++attempt) {
            const float R = d1 + (attempt == 0 ? a.collect : fminf(4.0f * tie, a.collect));
            if (R * a.g.lv[0].inv_h >= 0.95f) continue;
            const Near5 nb = ball_collect(a.g.lv[0], a.g.ox, a.g.oy, a.g.oz, x, y, z, R * R);
            if (!nb.complete) continue;
            const float lim = (d1 + tie) * (d1 + tie);
            float rho2 = R * R;
            bool open = true;
#pragma unroll
            for (int j = 0; j <= kCacheCands; ++j) {
                if (!open || nb.pos[j] < 0) continue;
                if (nb.pos[j] == bb.pos) continue;
                if (nb.d2[j] < lim && k < 3) {
                    if (k == 0) o1 = nb.pos[j]; else if (k == 1) o2 = nb.pos[j]; else o3 = nb.pos[j];
                    ++k;
                } else { rho2 = nb.d2[j]; open = false; }
            }
            rho = sqrtf(rho2) * 0.9999f;
            break;
        }
    }
    a.work[i] = make_float4(x, y, z, rho);
    a.cq[i] = make_float4(bb.qx, bb.qy, bb.qz, __int_as_float(bb.pos | (k > 0 ? kMoreBit : 0)));
    a.cn[i] = make_float4(nq.x, nq.y, nq.z, nq_dot(nq.x, nq.y, nq.z, bb.qx, bb.qy, bb.qz));
    if (k > 0) a.cmore[i] = make_int4(o1, o2, o3, bb.idx);
}

// Puts the source set into a spatially compact processing order (spatial_order_dev, grid.cu: bins of the target grid,
// ties in the caller's order) so that the lanes of a warp walk the same few cell rows.  The processing order only
// affects the order of the double sums (DESIGN.md "reduction geometry").
static int icp_sort_source(Ctx* ctx, int n, bool have_seed) {
    PW_TRY(ctx->icp_perm.reserve(ctx, (size_t)n * 4));
    const int n_pad = (n + 31) / 32 * 32;            // the per-point arrays of the loop are padded to whole batches
    PW_TRY(ctx->icp_sorted.reserve(ctx, (size_t)n_pad * sizeof(float4)));
    PW_TRY(ctx->icp_match.reserve(ctx, (size_t)n_pad * 3 * sizeof(float4)));      // cn, cq, cmore
    PW_TRY(spatial_order_dev(ctx, ctx->tgt.dev, ctx->icp_src.as<float4>(), n, ctx->icp_perm.as<uint32_t>()));
    src_gather_kernel<<<(n_pad + 255) / 256, 256, 0, ctx->stream>>>(ctx->icp_src.as<float4>(), ctx->icp_perm.as<uint32_t>(), n, n_pad,
                                                       ctx->icp_sorted.as<float4>(),
                                                       have_seed ? ctx->icp_seed.as<int>() : nullptr, ctx->tgt.dev.lv[0].pts,
                                                       ctx->tgt_aux.as<float4>(), ctx->icp_match.as<float4>(), ctx->icp_match.as<float4>() + n_pad);
    ctx->launches += 1;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

// Enqueues the inner loop on the library stream without waiting for it.  n = number of source points -- or, with
// n_dev, the capacity (the device holds the count: outer loop).  presorted: the caller has filled icp_sorted / icp_match
// (points in processing order + their matches, padded to whole batches) itself.
int icp_enqueue(Ctx* ctx, const pwicp_icp_params& prm, int n, const int* n_dev, bool presorted,
                bool want_mse, bool want_T, bool want_idx, IcpLaunch* L) {
    if (ctx->n1 < 1 || !ctx->tgt.dev.nlevels) { set_error(ctx, "icp: no target uploaded"); return PWICP_ERR_ARG; }
    if (n < 3) { set_error(ctx, "icp: fewer than 3 correspondences"); return PWICP_ERR_TOO_FEW_CORR; }
    if (prm.max_iter < 1 || prm.max_iter > kMaxIcpIter) { set_error(ctx, "icp: max_iter out of range"); return PWICP_ERR_ARG; }

    const size_t smem = (size_t)kIcpWarps * kStageSlots * 3 * 32 * sizeof(float4);      // s_stage
    void* kern = want_idx ? (void*)icp_persistent_kernel<true> : (void*)icp_persistent_kernel<false>;
    if (!ctx->icp_attr_set) {
        PW_CUDA(cudaFuncSetAttribute((void*)icp_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PW_CUDA(cudaFuncSetAttribute((void*)icp_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        PW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_persistent_kernel<true>, kIcpThreads, smem));
        if (occ < 1) { set_error(ctx, "icp: kernel does not fit on an SM"); return PWICP_ERR_CUDA; }
        ctx->icp_attr_set = true;
    }
    const long nb = ((long)n + 31) / 32;
    // one CTA per SM; fewer when there are not enough batches to give every warp one
    long want = (nb + kIcpWarps - 1) / kIcpWarps;
    int grid = (int)std::min<long>((long)ctx->num_sms, std::max<long>(1, want));

    const int n_pad = (int)(nb * 32);
    PW_TRY(ctx->icp_work.reserve(ctx, (size_t)n_pad * sizeof(float4)));
    // device scratch: CTA sums (double buffered) | searched[max_iter] | iter_ns[max_iter + 1] | phase_ns[max_iter][4]
    const size_t bytes_part = (size_t)2 * grid * kNumVals * 2 * sizeof(unsigned long long);   // tagged packets (icp_persistent_kernel)
    const size_t bytes_cnt = (size_t)prm.max_iter * sizeof(int);
    const size_t off_ns = (bytes_part + bytes_cnt + 7) & ~(size_t)7;
    const size_t off_carry = (off_ns + (size_t)(prm.max_iter + 1) * 8 + (size_t)prm.max_iter * 32 + 15) & ~(size_t)15;
    PW_TRY(ctx->icp_partials.reserve(ctx, off_carry + sizeof(IcpCarry) + 64));
    const size_t out_bytes = 64 + 16 + (size_t)prm.max_iter * (8 + 64);
    PW_TRY(ctx->icp_out.reserve(ctx, out_bytes));
    char* ob = ctx->icp_out.as<char>();
    PW_CUDA(cudaMemsetAsync(ob, 0, out_bytes, ctx->stream));
    if (want_idx) PW_TRY(ctx->icp_idx.reserve(ctx, (size_t)prm.max_iter * n * sizeof(int)));

    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (!presorted) {
        const bool have_seed = ctx->icp_seed_valid && ctx->icp_seed.p != nullptr;
        if (have_seed) PW_TRY(finish_deferred_aux(ctx));   // the gather of the seeded matches reads the normals
        PW_TRY(icp_sort_source(ctx, n, have_seed));
        ctx->icp_seed_valid = false;                       // seeds belong to one source set
        PW_CUDA(cudaEventRecord(ctx->ev3, ctx->stream));
        if (!have_seed) {
            const bool late = ctx->aux_deferred;           // normals still on their way up: search first, then wait
            icp_seed_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->tgt.dev, late ? nullptr : ctx->tgt_aux.as<float4>(),
                                                                      ctx->icp_sorted.as<float4>(), n,
                                                                      ctx->icp_match.as<float4>(), ctx->icp_match.as<float4>() + n_pad);
            ctx->launches++;
            if (late) {
                PW_TRY(finish_deferred_aux(ctx));
                icp_fill_cn_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->tgt_aux.as<float4>(), ctx->icp_match.as<float4>() + n_pad,
                                                                            n, ctx->icp_match.as<float4>());
                ctx->launches++;
            }
        }
    } else {
        PW_CUDA(cudaEventRecord(ctx->ev3, ctx->stream));
    }
    PW_TRY(finish_deferred_aux(ctx));

    IcpArgs a;
    a.g = ctx->tgt.dev;
    a.aux = ctx->tgt_aux.as<float4>();
    a.src = ctx->icp_sorted.as<float4>();
    a.work = ctx->icp_work.as<float4>();
    a.cn = ctx->icp_match.as<float4>();
    a.cq = a.cn + n_pad;
    a.cmore = reinterpret_cast<int4*>(a.cn + 2 * (size_t)n_pad);
    a.seed_exact = 1;         // classification matches (outer.cu) or icp_seed_kernel
    {
        // tuning knobs (cells of the finest grid level); the environment overrides are for A/B runs only
        auto knob = [](const char* name, float dflt) { const char* e = getenv(name); return e ? (float)atof(e) : dflt; };
        const float h = 1.0f / ctx->tgt.dev.lv[0].inv_h;
        a.collect = knob("PWICP_COLLECT_CELLS", PWICP_COLLECT_CELLS) * h;
        a.tie = knob("PWICP_TIE_CELLS", PWICP_TIE_CELLS) * h;
        a.tie_step = knob("PWICP_TIE_STEP_FRAC", 0.25f);
        a.build_step = knob("PWICP_BUILD_STEP_CELLS", PWICP_BUILD_STEP_CELLS) * h;
    }

    a.n = n;
    a.n_dev = n_dev;
    a.max_iter = prm.max_iter;
    a.force_iters = prm.force_iters;
    a.rot_thr = prm.rot_thr_default ? 0.99999 : (1.0 - prm.tf_eps);
    a.transl_thr = prm.tf_eps;
    a.mse_rel = prm.fit_eps;
    a.mse_abs = 1e-12;
    a.part = ctx->icp_partials.as<unsigned long long>();
    a.searched = reinterpret_cast<int*>(ctx->icp_partials.as<char>() + bytes_part);
    a.iter_ns = reinterpret_cast<unsigned long long*>(ctx->icp_partials.as<char>() + off_ns);
    a.phase_ns = a.iter_ns + prm.max_iter + 1;
    // packets (tag 0 = nothing posted), search counters, timers
    PW_CUDA(cudaMemsetAsync(a.part, 0, off_carry + sizeof(IcpCarry), ctx->stream));
    a.carry = reinterpret_cast<IcpCarry*>(ctx->icp_partials.as<char>() + off_carry);
    a.out_T = reinterpret_cast<float*>(ob);
    a.out_state = reinterpret_cast<int*>(ob + 64);
    a.mse_trace = want_mse ? reinterpret_cast<double*>(ob + 80) : nullptr;
    a.T_trace = want_T ? reinterpret_cast<float*>(ob + 80 + (size_t)prm.max_iter * 8) : nullptr;
    a.idx_trace = want_idx ? ctx->icp_idx.as<int>() : nullptr;

    // Two launches with the search of iteration 1 between them (icp_research_kernel) -- or one, when there is no
    // iteration 1 or the A/B switch says so.  (Cooperative launches: every CTA must be resident, the CTA sums are
    // exchanged by polling.)
    // PWICP_SPLIT_ITER1 = number of leading iterations whose search runs as its own kernel (A/B: 0, 1, 2).  One: a second
    // pass before iteration 2 -- 1.5 % of the queries search there -- loses (1M: iteration 2 105 -> 209 us and weaker
    // caches afterwards, profiles/r02ak_split2_ab.txt)
    // (read at every call: the tests switch it)  PWICP_SPLIT_MIN_POINTS overrides the size from which the split is used.
    const int split_env = [] { const char* e = getenv("PWICP_SPLIT_ITER1"); return e ? atoi(e) : 1; }();
    const int split_min = [] { const char* e = getenv("PWICP_SPLIT_MIN_POINTS"); return e ? atoi(e) : kSplitMinPoints; }();
    // worth it for large source sets only: below ~half a million points the extra launches cost more than the
    // occupancy gains (outer loop at 300k patches: 4.86 -> 5.06 ms with the split, profiles/r02af_split_launch_ab.txt)
    const int nsplit = (n >= split_min) ? std::max(0, std::min(std::min(split_env, 2), prm.max_iter - 1)) : 0;     // searches taken out: before iterations 1 .. nsplit
    void* kargs[] = {(void*)&a};
    cudaEvent_t ev_k[4] = {ctx->ev2, ctx->ev5, ctx->ev7, nullptr}, ev_r[3] = {ctx->ev4, ctx->ev6, nullptr};
    for (int part = 0; part <= nsplit; ++part) {
        a.it_begin = part; a.it_end = (part < nsplit) ? part + 1 : prm.max_iter;
        if (part > 0) {
            icp_research_kernel<<<(n_pad + 255) / 256, 256, 0, ctx->stream>>>(a);           // the search of iteration `part`
            ctx->launches++;
        }
        PW_CUDA(cudaEventRecord(ev_k[part], ctx->stream));
        PW_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kIcpThreads), kargs, smem, ctx->stream));
        ctx->launches++;
        if (part < nsplit) PW_CUDA(cudaEventRecord(ev_r[part], ctx->stream));
    }
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->icp_nsplit = nsplit;
    ctx->icp_prof_max_iter = prm.max_iter;
    ctx->icp_prof_off_searched = bytes_part; ctx->icp_prof_off_ns = off_ns;
    if (L) { L->grid = grid; L->out = ob; L->max_iter = prm.max_iter; }
    return PWICP_OK;
}

int icp_run_device(Ctx* ctx, const pwicp_icp_params& prm, float* T16, pwicp_icp_result* res,
                   double* mse_trace, float* T_trace, int* idx_trace) {
    const int n = ctx->n_icp;
    IcpLaunch L;
    PW_TRY(icp_enqueue(ctx, prm, n, nullptr, false, mse_trace != nullptr, T_trace != nullptr, idx_trace != nullptr, &L));
    const char* ob = L.out;
    const int grid = L.grid;
    struct { float T[16]; int st[4]; } host;
    PW_CUDA(cudaMemcpyAsync(&host, ob, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->tail_flag_dev)                                  // host-buffer call: the finite verdict rides on the same round trip
        PW_CUDA(cudaMemcpyAsync(&ctx->tail_flag_host, ctx->tail_flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f, kms = 0.f, rms = 0.f, sms = 0.f, pms = 0.f;
    PW_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    {   // the persistent kernel (its launches summed) and the stand-alone searches between them
        cudaEvent_t ev_k[3] = {ctx->ev2, ctx->ev5, ctx->ev7}, ev_r[3] = {ctx->ev4, ctx->ev6, ctx->ev1};
        for (int part = 0; part <= ctx->icp_nsplit; ++part) {
            float t = 0.f;
            PW_CUDA(cudaEventElapsedTime(&t, ev_k[part], part < ctx->icp_nsplit ? ev_r[part] : ctx->ev1));
            kms += t;
            if (part < ctx->icp_nsplit) { PW_CUDA(cudaEventElapsedTime(&t, ev_r[part], ev_k[part + 1])); rms += t; }
        }
    }
    PW_CUDA(cudaEventElapsedTime(&sms, ctx->ev0, ctx->ev3));
    PW_CUDA(cudaEventElapsedTime(&pms, ctx->ev3, ctx->ev2));
    ctx->last_ms = ms;
    const int n_iter = host.st[0];
    ctx->icp_prof_iters = n_iter;
    if (T16) for (int k = 0; k < 16; ++k) T16[k] = host.T[k];
    if (res) {
        res->n_iter = n_iter; res->conv_state = host.st[1];
        res->grid_blocks = grid; res->warps_per_block = kIcpWarps;
        res->group_batches = (grid * kIcpWarps) | (kIcpWarps << 16);   // reduction geometry for the oracle's reduce_mode 2
        res->device_ms = ms; res->correspondences = (long long)n_iter * n;
        res->kernel_ms = kms; res->natural_iters = host.st[2]; res->natural_state = host.st[3];
        res->sort_ms = sms; res->prepass_ms = pms; res->research_ms = rms;
    }
    if (mse_trace) PW_CUDA(cudaMemcpy(mse_trace, ob + 80, (size_t)n_iter * 8, cudaMemcpyDeviceToHost));
    if (T_trace) PW_CUDA(cudaMemcpy(T_trace, ob + 80 + (size_t)prm.max_iter * 8, (size_t)n_iter * 64, cudaMemcpyDeviceToHost));
    if (idx_trace) PW_CUDA(cudaMemcpy(idx_trace, ctx->icp_idx.p, (size_t)n_iter * n * sizeof(int), cudaMemcpyDeviceToHost));
    return PWICP_OK;
}

}  // namespace pwicp
