// nn_search.cuh -- exact 1-NN over the cell-sorted grid pyramid: "seed + ball" search (device code).
//
// Replaces pcl::KdTreeFLANN<PointXYZ, flann::L2_Simple<float>>::nearestKSearch(p, 1) as reached
// from pcl::registration::CorrespondenceEstimation::determineCorrespondences
// (reference call sites: src/Registration.cpp:737-747, :1293-1297, :597-601,
// src/CommonFunc.cpp:269-273 and the inner loop of src/Registration.cpp:1266).
//
// Idea: any target point s gives an upper bound d(p, s) of the NN distance, so the exact answer
// lies in the cells the closed ball B(p, d) touches.  ICP is temporally coherent -- the match of
// the previous iteration is almost always still (nearly) the nearest -- so with that match as the
// seed the ball covers one to three cells and every lane of a warp does the same small amount of
// work.  Without a seed the best point of the query's home cell (or of the first non-empty home
// cell on a coarser level) is taken.  When the ball spans many fine cells the scan runs on the
// pyramid level whose cells are about as large as the ball.
//
// Parity rules:
//   * distance = ((dx*dx) + dy*dy) + dz*dz in float32 with separately rounded operations
//     (__fmul_rn/__fadd_rn never contract into FMA);
//   * exact: the scanned box covers the ball with a conservative margin for the float rounding of
//     the cell assignment; a row is skipped only when its lower bound exceeds the current best;
//   * exact float ties resolve to the lowest original target index.
//
// History (profiles/r01a-c): a ring search with per-lane pruning was exact but ran 8 of 32 lanes;
// warp-cooperative group tiles (TMA-staged or read through L1) lost to load imbalance at the grid
// barrier.  Round 2 (profiles/r02j_union_search_ab.txt): one uniform scan per warp over the bounding
// box of its 32 queries' cells (every candidate broadcast to all lanes, 32 of 32 lanes busy, no
// divergent branch) -- the halo makes the union 3-5 times the cells a single query needs and the
// gain in lane efficiency is eaten exactly: classification 1.03x, pre-pass 1.03x, inner-loop search
// iterations 0.85x of this walk.  The seeded ball stays.
#pragma once
#include "common.cuh"

namespace pwicp {

// Squared distance (in cells) beyond which a seed counts as stale and the home cell is tried as well.  Stand-alone
// searches: half a cell.  Inner loop: one cell -- after the large first ICP steps the previous match IS stale (at 10M
// centroids the first transform moves the cloud edge by several cells: second iteration 5.7 -> 3.2 ms with the
// re-seed), but at half a cell the extra home-cell scan costs more than the larger ball on the following
// iterations (profiles/r01k_ab_reseed.txt: 1M loop 2.85 ms without, 3.18 ms at 0.25, 2.77 ms at 1.0).
#ifndef PWICP_RESEED_CELLS2
#define PWICP_RESEED_CELLS2 0.25f
#endif
#ifndef PWICP_RESEED_CELLS2_LOOP
#define PWICP_RESEED_CELLS2_LOOP 1.0f
#endif
#ifndef PWICP_BLOCK_WIDE
#define PWICP_BLOCK_WIDE 0       // stand-alone kernels: candidates of the block scan four at a time (A/B switch)
#endif
constexpr float kBallMaxCells = 6.0f;   // scan on the finest level whose ball radius is <= this many cells
                                        // (finer is cheaper unless the ball is mostly empty space: rows ~ (2r+1)^2)

struct Best {
    float d2;      // squared distance of the match (float, reference arithmetic)
    int idx;       // original target index
    int pos;       // position in the level-0 sorted array
    float qx, qy, qz;  // the matched target point (bit-identical to the caller's array)
};

__device__ __forceinline__ float l2_simple(float px, float py, float pz, float qx, float qy, float qz) {
    float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// lower bound (in cell units) of |p - q| along one axis for any q stored in cell k.
// f = (p - origin) * inv_h as computed for p, m = safety margin covering float rounding of the
// cell assignment (SURVEY "hard parts": bit-exact indices need a conservative bound).
__device__ __forceinline__ float axis_gap(float f, int k, float m) {
    float lo = (float)k - f, hi = f - (float)(k + 1);
    float g = fmaxf(lo, hi) - m;
    return fmaxf(g, 0.0f);
}

// Candidates [s, e) of the sorted array against the current best.  kWide: four loads in flight at
// a time -- in the register-capped persistent ICP kernel (24 warps per SM) the plain loop is a chain
// of exposed L1/L2 round trips (its trip count is unknown to the compiler); the stand-alone search
// kernels run at higher occupancy and are faster with the plain loop (profiles/r01f_*).
template <bool kWide>
__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, uint32_t s, uint32_t e,
                                           float px, float py, float pz, bool level0,
                                           float& bd, int& bi, int& bpos) {
    if (kWide) {
        for (uint32_t i = s; i < e; i += 4) {
            float4 q[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) q[k] = __ldg(pts + min(i + k, e - 1));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // slots past the end repeat the last candidate: same distance, same index, no effect
                const float d = l2_simple(px, py, pz, q[k].x, q[k].y, q[k].z);
                const int id = __float_as_int(q[k].w);
                if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bpos = level0 ? (int)min(i + k, e - 1) : -1; }
            }
        }
    } else {
        for (uint32_t i = s; i < e; ++i) {
            const float4 q = __ldg(pts + i);
            const float d = l2_simple(px, py, pz, q.x, q.y, q.z);
            const int id = __float_as_int(q.w);
            if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bpos = level0 ? (int)i : -1; }
        }
    }
}

// One cell row (ky, kz) of the ball: chord test against the CURRENT best, then a range scan.
template <bool kWide>
__device__ __forceinline__ void ball_row(const GridLevel& L, int ky, int kz, float fx, float fy, float fz,
                                         float mx, float my, float mz, int lx, int hx,
                                         float px, float py, float pz, bool level0,
                                         float& bd, int& bi, int& bpos) {
    const float gy = axis_gap(fy, ky, my), gz = axis_gap(fz, kz, mz);
    const float gyz = gy * gy + gz * gz, bc = bd * L.inv_h2;
    if (gyz > bc) return;                                        // row entirely outside the ball
    const float w = sqrtf(bc - gyz) * 1.00001f;                  // half chord of the ball along x
    const int lxr = max(lx, (int)floorf(fx - w - mx)), hxr = min(hx, (int)floorf(fx + w + mx));
    if (lxr > hxr) return;
    const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
    const uint32_t s = __ldg(L.cell_start + row + lxr), e = __ldg(L.cell_start + row + hxr + 1);
    scan_range<kWide>(L.pts, s, e, px, py, pz, level0, bd, bi, bpos);
}

// Up to 3 x 3 cell rows in one flattened pass (stand-alone search kernels; PWICP_FLAT_ROWS).  The nested walk -- row
// after row, every lane with its own trip counts -- keeps 11 of 32 lanes busy in the candidate loop and is issue-bound
// (ncu: profiles/r02ac_ncu_nn_kernel_bw.txt).  Here every lane (1) tests all its rows against the bound it has --
// gap, chord, and the cell_start loads of all surviving rows in flight together, one round trip instead of up to nine --
// (2) queues the non-empty ranges in a small local array and (3) runs ONE candidate loop over the queue: the warp's
// trip count is the largest per-lane candidate total instead of the sum of the per-row maxima.  The chords are cut
// against the incoming bound only (it is not tightened between rows), which is still exact: a row or cell is left
// out only when its lower bound exceeds an upper bound of the NN distance.  (cy, cz): a row to leave out (already
// scanned), or outside the box for none.
// Measured (profiles/r02ag_flat_rows_ab.txt): loses 5-8 % -- without the bound tightening from row to row the lanes look
// at more candidates than the flattening saves.  Off.
#ifndef PWICP_FLAT_ROWS
#define PWICP_FLAT_ROWS 0
#endif
__device__ __forceinline__ void rows_flat(const GridLevel& L, float fx, float fy, float fz, float mx, float my, float mz,
                                          int lx, int hx, int ly, int hy, int lz, int hz, int cy, int cz,
                                          float px, float py, float pz, float& bd, int& bi, int& bpos) {
    uint32_t rs[9], re[9];
    const float bc = bd * L.inv_h2;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        rs[r] = 0; re[r] = 0;
        const int ky = ly + r % 3, kz = lz + r / 3;
        if (ky <= hy && kz <= hz && !(ky == cy && kz == cz)) {
            const float gy = axis_gap(fy, ky, my), gz = axis_gap(fz, kz, mz);
            const float gyz = gy * gy + gz * gz;
            if (gyz <= bc) {
                const float w = sqrtf(bc - gyz) * 1.00001f;
                const int lxr = max(lx, (int)floorf(fx - w - mx)), hxr = min(hx, (int)floorf(fx + w + mx));
                if (lxr <= hxr) {
                    const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
                    rs[r] = __ldg(L.cell_start + row + lxr); re[r] = __ldg(L.cell_start + row + hxr + 1);
                }
            }
        }
    }
    uint2 queue[9];
    int nr = 0;
#pragma unroll
    for (int r = 0; r < 9; ++r)
        if (re[r] > rs[r]) queue[nr++] = make_uint2(rs[r], re[r]);
    int r = 0;
    uint32_t i = 0, e = 0;
    for (;;) {
        if (i >= e) {
            if (r >= nr) break;
            const uint2 t = queue[r++];
            i = t.x; e = t.y;
        }
        const float4 q = __ldg(L.pts + i);
        const float d = l2_simple(px, py, pz, q.x, q.y, q.z);
        const int id = __float_as_int(q.w);
        if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bpos = (int)i; }
        ++i;
    }
}

// Large balls (more than 3x3 rows): rows nearest-first, as square rings in y/z around the home
// row.  As soon as a closer target is met the ball shrinks and the remaining rings fall outside
// it, so a query far from the surface does not pay for the whole initial ball.  Out of line: this
// is the rare path and must not cost the common one registers.
template <bool kWide>
__device__ __forceinline__ void ball_scan_rings(const GridLevel& L, float fx, float fy, float fz,
                                                    float mx, float my, float mz,
                                                    int lx, int hx, int ly, int hy, int lz, int hz,
                                                    float px, float py, float pz, bool level0,
                                                    float& bd, int& bi, int& bpos) {
    const int cy = min(max((int)floorf(fy), ly), hy), cz = min(max((int)floorf(fz), lz), hz);
    const int R = max(max(cy - ly, hy - cy), max(cz - lz, hz - cz));
    ball_row<kWide>(L, cy, cz, fx, fy, fz, mx, my, mz, lx, hx, px, py, pz, level0, bd, bi, bpos);
    for (int t = 1; t <= R; ++t) {
        // every row of ring t (and of all later rings) is at least this far away in y or z
        const float g = fminf(fminf(axis_gap(fy, cy - t, my), axis_gap(fy, cy + t, my)),
                              fminf(axis_gap(fz, cz - t, mz), axis_gap(fz, cz + t, mz)));
        if (g * g > bd * L.inv_h2) break;
        const int z0 = max(cz - t, lz), z1 = min(cz + t, hz), y0 = max(cy - t, ly), y1 = min(cy + t, hy);
        for (int kz = z0; kz <= z1; ++kz) {
            const bool full = (kz == cz - t || kz == cz + t);
            for (int ky = y0; ky <= y1; ++ky) {
                if (!full && ky != cy - t && ky != cy + t) continue;
                ball_row<kWide>(L, ky, kz, fx, fy, fz, mx, my, mz, lx, hx, px, py, pz, level0, bd, bi, bpos);
            }
        }
    }
}

static __device__ __noinline__ void ball_scan_rings_ool(const GridLevel& L, float fx, float fy, float fz,
                                                        float mx, float my, float mz,
                                                        int lx, int hx, int ly, int hy, int lz, int hz,
                                                        float px, float py, float pz, bool level0,
                                                        float& bd, int& bi, int& bpos) {
    ball_scan_rings<true>(L, fx, fy, fz, mx, my, mz, lx, hx, ly, hy, lz, hz, px, py, pz, level0, bd, bi, bpos);
}

// Scans every cell of level L that the closed ball of radius sqrt(bd) around p touches.
// bd / bi / bpos are updated in place (bpos only on level 0, else -1 on improvement).
// kLean: the caller is register-bound (persistent ICP kernel); the rare large-ball path goes
// through an out-of-line copy.
template <bool kLean>
__device__ __forceinline__ void ball_scan(const GridLevel& L, float ox, float oy, float oz,
                                          float px, float py, float pz, bool level0,
                                          float& bd, int& bi, int& bpos) {
    const float fx = (px - ox) * L.inv_h, fy = (py - oy) * L.inv_h, fz = (pz - oz) * L.inv_h;
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f, mz = 0.01f + fabsf(fz) * 4e-6f;
    // radius in cell units, rounded up generously (float sqrt/mul errors are ~1e-7 relative)
    const float r = sqrtf(bd) * L.inv_h * 1.00001f;
    const int lx = min(max((int)floorf(fx - r - mx), 0), L.dx - 1), hx = min(max((int)floorf(fx + r + mx), 0), L.dx - 1);
    const int ly = min(max((int)floorf(fy - r - my), 0), L.dy - 1), hy = min(max((int)floorf(fy + r + my), 0), L.dy - 1);
    const int lz = min(max((int)floorf(fz - r - mz), 0), L.dz - 1), hz = min(max((int)floorf(fz + r + mz), 0), L.dz - 1);
    if (hy - ly <= 2 && hz - lz <= 2) {
        // one to nine rows (the common case of a seeded query): plain nested loops.  Two flattened
        // variants (one state-machine loop; row list in shared memory + one candidate loop) were
        // measured and were not faster: the batch time is a chain of dependent L2 round trips.
        if (!kLean && PWICP_FLAT_ROWS && level0) {
            rows_flat(L, fx, fy, fz, mx, my, mz, lx, hx, ly, hy, lz, hz, -1, -1, px, py, pz, bd, bi, bpos);
        } else {
            for (int kz = lz; kz <= hz; ++kz)
                for (int ky = ly; ky <= hy; ++ky)
                    ball_row<kLean>(L, ky, kz, fx, fy, fz, mx, my, mz, lx, hx, px, py, pz, level0, bd, bi, bpos);
        }
    } else if (kLean) {
        ball_scan_rings_ool(L, fx, fy, fz, mx, my, mz, lx, hx, ly, hy, lz, hz, px, py, pz, level0, bd, bi, bpos);
    } else {
        ball_scan_rings<false>(L, fx, fy, fz, mx, my, mz, lx, hx, ly, hy, lz, hz, px, py, pz, level0, bd, bi, bpos);
    }
}

// Best of (a sample of) the points in [s, e): candidate seed.
__device__ __forceinline__ void seed_range(const float4* __restrict__ pts, uint32_t s, uint32_t e, uint32_t step,
                                           float px, float py, float pz, bool level0, float& bd, int& bi, int& bpos) {
    for (uint32_t i = s; i < e; i += step) {
        const float4 q = __ldg(pts + i);
        const float d = l2_simple(px, py, pz, q.x, q.y, q.z);
        const int id = __float_as_int(q.w);
        if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bpos = level0 ? (int)i : -1; }
    }
}

// A candidate for a query without (usable) history.  Any target point is a valid seed; a close
// one keeps the ball small.  Per level, finest first: the home cell, then its 3x3x3 block; the
// first level that yields a candidate wins.  Coarse cells are sampled, not scanned.
__device__ __forceinline__ void find_seed(const GridDev& g, float px, float py, float pz,
                                          float& bd, int& bi, int& bpos) {
    for (int l = 0; l < g.nlevels; ++l) {
        const GridLevel& L = g.lv[l];
        const int cx = min(max((int)floorf((px - g.ox) * L.inv_h), 0), L.dx - 1);
        const int cy = min(max((int)floorf((py - g.oy) * L.inv_h), 0), L.dy - 1);
        const int cz = min(max((int)floorf((pz - g.oz) * L.inv_h), 0), L.dz - 1);
        const uint32_t c = ((uint32_t)cz * (uint32_t)L.dy + (uint32_t)cy) * (uint32_t)L.dx + (uint32_t)cx;
        const uint32_t s = __ldg(L.cell_start + c), e = __ldg(L.cell_start + c + 1);
        if (e > s) {
            const uint32_t step = (l == 0 || e - s <= 16u) ? 1u : (e - s) / 16u;
            seed_range(L.pts, s, e, step, px, py, pz, l == 0, bd, bi, bpos);
            return;
        }
        bool found = false;
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, L.dx - 1);
        for (int kz = max(cz - 1, 0); kz <= min(cz + 1, L.dz - 1); ++kz)
            for (int ky = max(cy - 1, 0); ky <= min(cy + 1, L.dy - 1); ++ky) {
                const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
                const uint32_t rs = __ldg(L.cell_start + row + x0), re = __ldg(L.cell_start + row + x1 + 1);
                if (re > rs) {
                    const uint32_t step = (l == 0 || re - rs <= 8u) ? 1u : (re - rs) / 8u;
                    seed_range(L.pts, rs, re, step, px, py, pz, l == 0, bd, bi, bpos);
                    found = true;
                }
            }
        if (found) return;
    }
    // nothing near on any level (query far outside the target): take the first sorted target
    const float4 q = __ldg(g.lv[0].pts);
    const float d = l2_simple(px, py, pz, q.x, q.y, q.z);
    if (d < bd) { bd = d; bi = __float_as_int(q.w); bpos = 0; }
}

static __device__ __noinline__ void find_seed_ool(const GridDev& g, float px, float py, float pz,
                                                  float& bd, int& bi, int& bpos) {
    find_seed(g, px, py, pz, bd, bi, bpos);
}

// A query without a usable seed: the 3x3x3 block of finest-level cells around its home cell in ONE pass -- the home row
// first (for overlapping clouds it holds the nearest target or one nearly as close), then the eight rows around it,
// each pruned against the running best by the same gap / chord tests as a ball scan.  (Round 1 first took the best
// of the home cell as a seed and then scanned the ball through it, home cell included a second time.)  Returns true
// when the closed ball of the best distance lies inside the block (faces on the grid boundary are open: there is no
// target beyond them), i.e. the answer is exact; else bd / bi / bpos hold a valid upper bound (or nothing).
template <bool kWide, bool kFlat = false>
__device__ __forceinline__ bool block_scan(const GridLevel& L, float ox, float oy, float oz, float px, float py, float pz,
                                           float& bd, int& bi, int& bpos) {
    const float fx = (px - ox) * L.inv_h, fy = (py - oy) * L.inv_h, fz = (pz - oz) * L.inv_h;
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f, mz = 0.01f + fabsf(fz) * 4e-6f;
    const int cx = min(max((int)floorf(fx), 0), L.dx - 1), cy = min(max((int)floorf(fy), 0), L.dy - 1),
              cz = min(max((int)floorf(fz), 0), L.dz - 1);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, L.dx - 1);
    const int y0 = max(cy - 1, 0), y1 = min(cy + 1, L.dy - 1), z0 = max(cz - 1, 0), z1 = min(cz + 1, L.dz - 1);
    ball_row<kWide>(L, cy, cz, fx, fy, fz, mx, my, mz, x0, x1, px, py, pz, true, bd, bi, bpos);
    if (kFlat) {
        rows_flat(L, fx, fy, fz, mx, my, mz, x0, x1, y0, y1, z0, z1, cy, cz, px, py, pz, bd, bi, bpos);
    } else {
        for (int kz = z0; kz <= z1; ++kz)
            for (int ky = y0; ky <= y1; ++ky) {
                if (ky == cy && kz == cz) continue;
                ball_row<kWide>(L, ky, kz, fx, fy, fz, mx, my, mz, x0, x1, px, py, pz, true, bd, bi, bpos);
            }
    }
    if (bpos < 0) return false;
    const float r = sqrtf(bd) * L.inv_h * 1.00001f;
    return (x0 == 0 || fx - r - mx >= (float)x0) && (x1 == L.dx - 1 || fx + r + mx < (float)(x1 + 1)) &&
           (y0 == 0 || fy - r - my >= (float)y0) && (y1 == L.dy - 1 || fy + r + my < (float)(y1 + 1)) &&
           (z0 == 0 || fz - r - mz >= (float)z0) && (z1 == L.dz - 1 || fz + r + mz < (float)(z1 + 1));
}

static __device__ __noinline__ bool block_scan_ool(const GridLevel& L, float ox, float oy, float oz, float px, float py,
                                                   float pz, float& bd, int& bi, int& bpos) {
    return block_scan<true>(L, ox, oy, oz, px, py, pz, bd, bi, bpos);
}

// Exact nearest neighbour.  seed_pos >= 0: level-0 position of a target known to be close
// (normally the previous match); -1: none.
// kLean = true (inner ICP loop): the rare paths are kept out of line.
template <bool kLean = false>
__device__ __forceinline__ Best nn_search_seeded(const GridDev& g, float px, float py, float pz, int seed_pos,
                                                 bool always_block = false) {
    float bd = __int_as_float(0x7f800000);
    int bi = 0x7fffffff, bpos = -1;
    bool done = false;
    if (seed_pos >= 0) {
        const float4 q = __ldg(g.lv[0].pts + seed_pos);
        bd = l2_simple(px, py, pz, q.x, q.y, q.z);
        bi = __float_as_int(q.w);
        bpos = seed_pos;
    }
    // no seed, or a stale one (the cloud moved by a good part of a cell since it was recorded: the ball through it
    // would be large): the block around the home cell, with the seed as the first bound
    const float stale = kLean ? PWICP_RESEED_CELLS2_LOOP : PWICP_RESEED_CELLS2;
    // always_block: every lane takes the block scan, the seed only bounds it (callers whose seeds are a mix of fresh
    // and stale ones: two code paths in one warp run one after the other)
    if (seed_pos < 0 || always_block || bd * g.lv[0].inv_h2 > stale) {
        done = kLean ? block_scan_ool(g.lv[0], g.ox, g.oy, g.oz, px, py, pz, bd, bi, bpos)
                     : block_scan<PWICP_BLOCK_WIDE != 0, PWICP_FLAT_ROWS != 0>(g.lv[0], g.ox, g.oy, g.oz, px, py, pz, bd, bi, bpos);
        if (!done && bpos < 0) {             // empty block: a sampled point of the first non-empty coarser home cell
            if (kLean) find_seed_ool(g, px, py, pz, bd, bi, bpos); else find_seed(g, px, py, pz, bd, bi, bpos);
        }
    }
    if (!done) {
        // finest level on which the ball spans only a few cells
        int l = 0;
        float r = sqrtf(bd) * g.lv[0].inv_h;
        while (l < g.nlevels - 1 && r > kBallMaxCells) { ++l; r *= (1.0f / kLevelFactor); }
        ball_scan<kLean>(g.lv[l], g.ox, g.oy, g.oz, px, py, pz, l == 0, bd, bi, bpos);
    }

    Best b;
    b.d2 = bd; b.idx = bi;
    if (bpos < 0) bpos = (int)__ldg(g.inv_perm + bi);
    b.pos = bpos;
    const float4 q = __ldg(g.lv[0].pts + bpos);
    b.qx = q.x; b.qy = q.y; b.qz = q.z;
    return b;
}

// ---------------------------------------------------------------------------------------------
// Team search (round 2): kTeam adjacent lanes work on ONE query at a time.
//
// The thread-per-query walk above spends its time in divergence: the 32 lanes of a warp scan rows of different
// lengths and prune at different points, so a warp issues ~8000 instructions for 32 queries of ~70 candidates each
// (ncu, round 1: 11 of 32 lanes active) and every query is a chain of up to 18 dependent L2 round trips.  Two
// warp-wide formulations lost (one tile / one union box per warp: the halo of 32 queries is 3-5x what one needs).
// Here the unit of cooperation is a team of kTeam lanes and the unit of work is still ONE query's own 3x3x3 block:
//   step 1  the home row (three cells, one contiguous range): candidate j of the range goes to lane j mod kTeam,
//           arg-min over the team by xor shuffles -> an upper bound bd;
//   step 2  the eight rows around it, one (or two) per lane: gap test and chord cut against bd -- most rows die
//           here without a memory access; the surviving ranges are scanned by the whole team, one after the other;
//   step 3  arg-min again; the answer is exact when the ball of the result lies inside the block.
// Four dependent memory round trips per query instead of eighteen, every candidate load a coalesced row of kTeam
// 16-byte cells, and the rows are pruned against a bound exactly as before, so the candidate count does not grow.
// The rounds of the 32 / kTeam teams of a warp run independently (team masks on every shuffle).
// Pruning never changes the result: a row or chord is skipped only when its lower bound exceeds an upper bound of
// the NN distance, and the arg-min is over (distance, original index) pairs, which is order-independent.
#ifndef PWICP_TEAM_SEARCH
#define PWICP_TEAM_SEARCH 0      // measured (profiles/r02ab_team_search_ab.txt): loses to the thread-per-query walk
#endif
#ifndef PWICP_TEAM
#define PWICP_TEAM 8
#endif
constexpr int kTeam = PWICP_TEAM;

template <int kT>
__device__ __forceinline__ void team_argmin(unsigned tmask, float& bd, int& bi, int& bpos) {
#pragma unroll
    for (int o = kT / 2; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(tmask, bd, o);
        const int oi = __shfl_xor_sync(tmask, bi, o), op = __shfl_xor_sync(tmask, bpos, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bpos = op; }
    }
}

// candidates [s, e) against the lane's own best, candidate s + sub + k * kT for lane `sub` of the team
template <int kT>
__device__ __forceinline__ void team_scan_range(const float4* __restrict__ pts, uint32_t s, uint32_t e, int sub,
                                                float px, float py, float pz, float& bd, int& bi, int& bpos) {
    for (uint32_t i = s + (uint32_t)sub; i < e; i += kT) {
        const float4 q = __ldg(pts + i);
        const float d = l2_simple(px, py, pz, q.x, q.y, q.z);
        const int id = __float_as_int(q.w);
        if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bpos = (int)i; }
    }
}

// One query (the same px, py, pz and the same incoming bound in all lanes of the team), its 3x3x3 block on level 0.
// All lanes return the same bd / bi / bpos; true = the result is exact (see block_scan).
template <int kT>
__device__ __forceinline__ bool team_block_scan(const GridLevel& L, float ox, float oy, float oz, float px, float py,
                                                float pz, unsigned tmask, int tbase, int sub,
                                                float& bd, int& bi, int& bpos) {
    const float fx = (px - ox) * L.inv_h, fy = (py - oy) * L.inv_h, fz = (pz - oz) * L.inv_h;
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f, mz = 0.01f + fabsf(fz) * 4e-6f;
    const int cx = min(max((int)floorf(fx), 0), L.dx - 1), cy = min(max((int)floorf(fy), 0), L.dy - 1),
              cz = min(max((int)floorf(fz), 0), L.dz - 1);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, L.dx - 1);
    const int y0 = max(cy - 1, 0), y1 = min(cy + 1, L.dy - 1), z0 = max(cz - 1, 0), z1 = min(cz + 1, L.dz - 1);
    // step 1: the home row, cut by the incoming bound if there is one (a stale seed)
    {
        uint32_t s = 0, e = 0;
        const float gy = axis_gap(fy, cy, my), gz = axis_gap(fz, cz, mz);
        const float gyz = gy * gy + gz * gz, bc = bd * L.inv_h2;
        if (gyz <= bc) {
            const float w = sqrtf(bc - gyz) * 1.00001f;
            const int lxr = max(x0, (int)floorf(fx - w - mx)), hxr = min(x1, (int)floorf(fx + w + mx));
            if (lxr <= hxr) {
                const uint32_t row = ((uint32_t)cz * (uint32_t)L.dy + (uint32_t)cy) * (uint32_t)L.dx;
                s = __ldg(L.cell_start + row + lxr); e = __ldg(L.cell_start + row + hxr + 1);
            }
        }
        team_scan_range<kT>(L.pts, s, e, sub, px, py, pz, bd, bi, bpos);
        team_argmin<kT>(tmask, bd, bi, bpos);
    }
    // step 2: the eight rows around it; lane `sub` tests rows sub, sub + kT, ... against the bound of step 1
    constexpr int kSlots = (8 + kT - 1) / kT;
    uint32_t rs[kSlots], re[kSlots];
    const float bc = bd * L.inv_h2;
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
        rs[k] = 0; re[k] = 0;
        const int r = sub + k * kT;                      // 0..7 -> the eight (dy, dz) != (0, 0)
        const int r9 = r + (r >= 4 ? 1 : 0);
        const int ky = cy + r9 % 3 - 1, kz = cz + r9 / 3 - 1;
        if (r < 8 && ky >= y0 && ky <= y1 && kz >= z0 && kz <= z1) {
            const float gy = axis_gap(fy, ky, my), gz = axis_gap(fz, kz, mz);
            const float gyz = gy * gy + gz * gz;
            if (gyz <= bc) {
                const float w = sqrtf(bc - gyz) * 1.00001f;
                const int lxr = max(x0, (int)floorf(fx - w - mx)), hxr = min(x1, (int)floorf(fx + w + mx));
                if (lxr <= hxr) {
                    const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
                    rs[k] = __ldg(L.cell_start + row + lxr); re[k] = __ldg(L.cell_start + row + hxr + 1);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
        unsigned live = (__ballot_sync(tmask, re[k] > rs[k]) >> tbase) & ((kT == 32) ? 0xffffffffu : ((1u << kT) - 1u));
        while (live) {
            const int j = __ffs(live) - 1;
            live &= live - 1;
            const uint32_t s = __shfl_sync(tmask, rs[k], tbase + j), e = __shfl_sync(tmask, re[k], tbase + j);
            team_scan_range<kT>(L.pts, s, e, sub, px, py, pz, bd, bi, bpos);
        }
    }
    team_argmin<kT>(tmask, bd, bi, bpos);
    if (bpos < 0) return false;
    const float r = sqrtf(bd) * L.inv_h * 1.00001f;
    return (x0 == 0 || fx - r - mx >= (float)x0) && (x1 == L.dx - 1 || fx + r + mx < (float)(x1 + 1)) &&
           (y0 == 0 || fy - r - my >= (float)y0) && (y1 == L.dy - 1 || fy + r + my < (float)(y1 + 1)) &&
           (z0 == 0 || fz - r - mz >= (float)z0) && (z1 == L.dz - 1 || fz + r + mz < (float)(z1 + 1));
}

// Exact nearest neighbour for every lane of a warp (a warp-collective call: ALL 32 lanes must arrive; lanes without a
// query pass active = false).  Queries with a fresh seed run the seeded ball as before (one to nine short rows, the
// same trip counts across the warp); queries without a usable seed go through the team block scan, the teams taking
// the queries of their own kT lanes one after the other; what the block does not settle (empty block, ball beyond
// the block) finishes on the serial path from the bound reached.
template <int kT = kTeam>
__device__ __forceinline__ Best nn_search_warp(const GridDev& g, float px, float py, float pz, int seed_pos, bool active) {
#if !PWICP_TEAM_SEARCH
    if (!active) { Best none; none.d2 = 0.f; none.idx = 0; none.pos = 0; none.qx = none.qy = none.qz = 0.f; return none; }
    return nn_search_seeded(g, px, py, pz, seed_pos);
#endif
    const int lane = (int)(threadIdx.x & 31), sub = lane % kT, tbase = lane - sub;
    const unsigned tmask = ((kT == 32) ? 0xffffffffu : ((1u << kT) - 1u)) << tbase;
    float bd = __int_as_float(0x7f800000);
    int bi = 0x7fffffff, bpos = -1;
    if (active && seed_pos >= 0) {
        const float4 q = __ldg(g.lv[0].pts + seed_pos);
        bd = l2_simple(px, py, pz, q.x, q.y, q.z);
        bi = __float_as_int(q.w);
        bpos = seed_pos;
    }
    const bool want_block = active && (seed_pos < 0 || bd * g.lv[0].inv_h2 > PWICP_RESEED_CELLS2);
    bool done = false;
    unsigned todo = (__ballot_sync(0xffffffffu, want_block) >> tbase) & ((kT == 32) ? 0xffffffffu : ((1u << kT) - 1u));
    while (todo) {                                       // team-uniform
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const int owner = tbase + j;
        const float qx = __shfl_sync(tmask, px, owner), qy = __shfl_sync(tmask, py, owner), qz = __shfl_sync(tmask, pz, owner);
        float tbd = __shfl_sync(tmask, bd, owner);
        int tbi = __shfl_sync(tmask, bi, owner), tbp = __shfl_sync(tmask, bpos, owner);
        const bool ok = team_block_scan<kT>(g.lv[0], g.ox, g.oy, g.oz, qx, qy, qz, tmask, tbase, sub, tbd, tbi, tbp);
        if (sub == j) { bd = tbd; bi = tbi; bpos = tbp; done = ok; }
    }
    __syncwarp();
    if (active && !done) {
        if (bpos < 0) find_seed(g, px, py, pz, bd, bi, bpos);
        int l = 0;
        float r = sqrtf(bd) * g.lv[0].inv_h;
        while (l < g.nlevels - 1 && r > kBallMaxCells) { ++l; r *= (1.0f / kLevelFactor); }
        ball_scan<false>(g.lv[l], g.ox, g.oy, g.oz, px, py, pz, l == 0, bd, bi, bpos);
    }
    Best b;
    b.d2 = bd; b.idx = bi; b.pos = 0; b.qx = b.qy = b.qz = 0.f;
    if (active) {
        if (bpos < 0) bpos = (int)__ldg(g.inv_perm + bi);
        b.pos = bpos;
        const float4 q = __ldg(g.lv[0].pts + bpos);
        b.qx = q.x; b.qy = q.y; b.qz = q.z;
    }
    return b;
}

// ---------------------------------------------------------------------------------------------
// Candidate cache (inner ICP loop).  Around the position a at which a query was last searched (its anchor) the
// targets are known in order of distance: d_1 <= d_2 <= ...  The cache holds the first m of them (the match and up to
// three more) and the radius rho just below d_(m+1), so that S = { q : d(a, q) < rho } is exactly the cached set.
// Later, at position p with |p - a| <= path:  if  |p - best(S)| + path < rho  then every target at least as close to
// p as best(S) lies within rho of a, i.e. in S, and best(S) under the tie rule is the exact answer.  m is chosen per
// query: targets within `tie` of the match (they can overtake it after a tiny move) are cached with it, the first
// one beyond that sets rho -- for most queries m = 1 and rho - d_1 is the gap to the second-nearest target (a good
// part of the point spacing), so a cache built right after the first, large ICP step survives the later small ones.
constexpr int kCacheCands = 4;

struct Near5 {
    float d2[kCacheCands + 1];   // squared distances, ascending; +inf = no such target within the scanned ball
    int pos[kCacheCands + 1];    // level-0 positions, -1 = none
    int complete;                // the scan covered the whole ball (else: ball too large for the 3x3-row scan)
};

// The kCacheCands + 1 nearest targets within sqrt(rho2max) of p on level 0 (ordered by distance only; the tie rule
// is the caller's business).  Out of line: runs once per query when its cache is (re)built.
static __device__ __noinline__ Near5 ball_collect(const GridLevel& L, float ox, float oy, float oz,
                                                  float px, float py, float pz, float rho2max) {
    Near5 out;
#pragma unroll
    for (int k = 0; k <= kCacheCands; ++k) { out.d2[k] = __int_as_float(0x7f800000); out.pos[k] = -1; }
    out.complete = 0;
    const float fx = (px - ox) * L.inv_h, fy = (py - oy) * L.inv_h, fz = (pz - oz) * L.inv_h;
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f, mz = 0.01f + fabsf(fz) * 4e-6f;
    const float r = sqrtf(rho2max) * L.inv_h * 1.00001f;
    const int lx = min(max((int)floorf(fx - r - mx), 0), L.dx - 1), hx = min(max((int)floorf(fx + r + mx), 0), L.dx - 1);
    const int ly = min(max((int)floorf(fy - r - my), 0), L.dy - 1), hy = min(max((int)floorf(fy + r + my), 0), L.dy - 1);
    const int lz = min(max((int)floorf(fz - r - mz), 0), L.dz - 1), hz = min(max((int)floorf(fz + r + mz), 0), L.dz - 1);
    if (hy - ly > 2 || hz - lz > 2) return out;
    float D[kCacheCands + 1];
    int P[kCacheCands + 1];
#pragma unroll
    for (int k = 0; k <= kCacheCands; ++k) { D[k] = __int_as_float(0x7f800000); P[k] = -1; }
    const float bc = rho2max * L.inv_h2;
    // Targets inside the radius are rare (one to three per query).  Sorting each one into the list where it is met ran
    // the 25-instruction insertion with two lanes of the warp active (16 % of the instructions of icp_research_kernel,
    // profiles/r02x): hits are parked in a small local list in scan order and sorted in after the walk, when the lanes
    // are back together.  Same list as before, same order among equal distances.
    constexpr int kPark = 8;
    float LD[kPark];
    int LP[kPark], cnt = 0;
    auto flush = [&]() {
        for (int j = 0; j < cnt; ++j) {
            float cd = LD[j];
            int cp = LP[j];
            if (cd < D[kCacheCands]) {          // sorted insertion, ascending distance
#pragma unroll
                for (int k = 0; k <= kCacheCands; ++k)
                    if (cd < D[k]) { const float td = D[k]; const int tp = P[k]; D[k] = cd; P[k] = cp; cd = td; cp = tp; }
            }
        }
        cnt = 0;
    };
    for (int kz = lz; kz <= hz; ++kz)
        for (int ky = ly; ky <= hy; ++ky) {
            const float gy = axis_gap(fy, ky, my), gz = axis_gap(fz, kz, mz);
            const float gyz = gy * gy + gz * gz;
            if (gyz > bc) continue;
            const float w = sqrtf(bc - gyz) * 1.00001f;
            const int lxr = max(lx, (int)floorf(fx - w - mx)), hxr = min(hx, (int)floorf(fx + w + mx));
            if (lxr > hxr) continue;
            const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
            const uint32_t s = __ldg(L.cell_start + row + lxr), e = __ldg(L.cell_start + row + hxr + 1);
            for (uint32_t i = s; i < e; ++i) {
                const float4 q = __ldg(L.pts + i);
                const float cd = l2_simple(px, py, pz, q.x, q.y, q.z);
                if (cd <= rho2max) {
                    if (cnt == kPark) flush();
                    LD[cnt] = cd; LP[cnt] = (int)i; ++cnt;
                }
            }
        }
    flush();
#pragma unroll
    for (int k = 0; k <= kCacheCands; ++k) { out.d2[k] = D[k]; out.pos[k] = P[k]; }
    out.complete = 1;
    return out;
}

// (A whole-warp version of the seeded search + collect for batches in which only a few queries miss their cache was
// measured and rejected: profiles/r02ae_warp_search_ab.txt.)
__device__ __forceinline__ Best nn_search(const GridDev& g, float px, float py, float pz) {
    return nn_search_seeded(g, px, py, pz, -1);
}

}  // namespace pwicp
