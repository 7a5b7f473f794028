// nn_search.cuh -- exact 1-NN over the cell-sorted grid pyramid (device code).
//
// Replaces pcl::KdTreeFLANN<PointXYZ, flann::L2_Simple<float>>::nearestKSearch(p, 1) as reached
// from pcl::registration::CorrespondenceEstimation::determineCorrespondences
// (reference call sites: src/Registration.cpp:737-747, :1293-1297, :597-601,
// src/CommonFunc.cpp:269-273 and the inner loop of src/Registration.cpp:1266).
//
// Parity rules:
//   * distance = ((dx*dx) + dy*dy) + dz*dz in float32 with separately rounded operations
//     (__fmul_rn/__fadd_rn never contract into FMA);
//   * the search is exact: a cell is skipped only when a conservative lower bound of the
//     distance to anything inside it exceeds the current best;
//   * exact float ties resolve to the lowest original target index.
#pragma once
#include "common.cuh"

namespace pwicp {

struct Best {
    float d2;      // best squared distance so far (float, reference arithmetic)
    int idx;       // original target index
    int pos;       // position in the level-0 sorted array, -1 when found on a coarser level
    float qx, qy, qz;  // the matched target point (bit-identical to the caller's array)
};

__device__ __forceinline__ float l2_simple(float px, float py, float pz, float qx, float qy, float qz) {
    float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, uint32_t s, uint32_t e,
                                           float px, float py, float pz, bool level0, Best& b) {
    for (uint32_t i = s; i < e; ++i) {
        float4 q = __ldg(pts + i);
        float d = l2_simple(px, py, pz, q.x, q.y, q.z);
        int id = __float_as_int(q.w);
        if (d < b.d2 || (d == b.d2 && id < b.idx)) {
            b.d2 = d; b.idx = id; b.pos = level0 ? (int)i : -1;
            b.qx = q.x; b.qy = q.y; b.qz = q.z;
        }
    }
}

// lower bound (in cell units) of |p - q| along one axis for any q stored in cell k.
// f = (p - origin) * inv_h as computed for p, m = safety margin covering float rounding of the
// cell assignment (SURVEY "hard parts": bit-exact indices need a conservative bound).
__device__ __forceinline__ float axis_gap(float f, int k, float m) {
    float lo = (float)k - f, hi = f - (float)(k + 1);
    float g = fmaxf(lo, hi) - m;
    return fmaxf(g, 0.0f);
}

// Searches rings 0..rmax (Chebyshev distance in cells around the home cell) of one level.
// Returns true when the result is proven exact.
__device__ __forceinline__ bool search_level(const GridLevel& L, float ox, float oy, float oz,
                                             float px, float py, float pz, int rmax, bool level0,
                                             Best& b) {
    const float fx = (px - ox) * L.inv_h, fy = (py - oy) * L.inv_h, fz = (pz - oz) * L.inv_h;
    const int cx = min(max((int)floorf(fx), 0), L.dx - 1);
    const int cy = min(max((int)floorf(fy), 0), L.dy - 1);
    const int cz = min(max((int)floorf(fz), 0), L.dz - 1);
    const float mx = 0.01f + fabsf(fx) * 4e-6f, my = 0.01f + fabsf(fy) * 4e-6f,
                mz = 0.01f + fabsf(fz) * 4e-6f;
    const uint32_t* __restrict__ cs = L.cell_start;

    for (int r = 0; r <= rmax; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, L.dz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, L.dy - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, L.dx - 1);
        for (int kz = z0; kz <= z1; ++kz) {
            const float gz = axis_gap(fz, kz, mz);
            const bool zshell = (kz - cz == r) || (cz - kz == r);
            for (int ky = y0; ky <= y1; ++ky) {
                const float gy = axis_gap(fy, ky, my);
                const float gyz = gy * gy + gz * gz;
                float bc = b.d2 * L.inv_h2;          // best in cell units^2
                if (gyz > bc) continue;
                const uint32_t row = ((uint32_t)kz * (uint32_t)L.dy + (uint32_t)ky) * (uint32_t)L.dx;
                const bool shell_row = zshell || (ky - cy == r) || (cy - ky == r);
                if (shell_row) {
                    int xl = x0, xh = x1;
                    while (xl <= xh) { float g = axis_gap(fx, xl, mx); if (g * g + gyz > bc) ++xl; else break; }
                    while (xh >= xl) { float g = axis_gap(fx, xh, mx); if (g * g + gyz > bc) --xh; else break; }
                    if (xl <= xh) scan_range(L.pts, __ldg(cs + row + xl), __ldg(cs + row + xh + 1), px, py, pz, level0, b);
                } else {
                    // interior row of the shell: only the two end cells are new
                    if (cx - r >= 0) {
                        float g = axis_gap(fx, cx - r, mx);
                        if (g * g + gyz <= bc)
                            scan_range(L.pts, __ldg(cs + row + cx - r), __ldg(cs + row + cx - r + 1), px, py, pz, level0, b);
                    }
                    if (cx + r <= L.dx - 1) {
                        bc = b.d2 * L.inv_h2;
                        float g = axis_gap(fx, cx + r, mx);
                        if (g * g + gyz <= bc)
                            scan_range(L.pts, __ldg(cs + row + cx + r), __ldg(cs + row + cx + r + 1), px, py, pz, level0, b);
                    }
                }
            }
        }
        // exactness test: everything not yet visited lies outside the block of radius r
        float bound = 3.0e38f;
        bool any = false;
        if (cx - r - 1 >= 0)    { bound = fminf(bound, axis_gap(fx, cx - r - 1, mx)); any = true; }
        if (cx + r + 1 < L.dx)  { bound = fminf(bound, axis_gap(fx, cx + r + 1, mx)); any = true; }
        if (cy - r - 1 >= 0)    { bound = fminf(bound, axis_gap(fy, cy - r - 1, my)); any = true; }
        if (cy + r + 1 < L.dy)  { bound = fminf(bound, axis_gap(fy, cy + r + 1, my)); any = true; }
        if (cz - r - 1 >= 0)    { bound = fminf(bound, axis_gap(fz, cz - r - 1, mz)); any = true; }
        if (cz + r + 1 < L.dz)  { bound = fminf(bound, axis_gap(fz, cz + r + 1, mz)); any = true; }
        if (!any) return true;                        // the whole level has been covered
        if (b.d2 * L.inv_h2 < bound * bound) return true;
    }
    return false;
}

__device__ __forceinline__ Best nn_search(const GridDev& g, float px, float py, float pz) {
    Best b;
    b.d2 = __int_as_float(0x7f800000); b.idx = 0x7fffffff; b.pos = -1;
    b.qx = b.qy = b.qz = 0.f;
    const int last = g.nlevels - 1;
    for (int l = 0; l <= last; ++l) {
        const int rmax = (l == last) ? 0x3fffffff : kRingsPerLevel;
        if (search_level(g.lv[l], g.ox, g.oy, g.oz, px, py, pz, rmax, l == 0, b)) break;
    }
    if (b.pos < 0 && b.idx != 0x7fffffff) b.pos = (int)__ldg(g.inv_perm + b.idx);
    return b;
}

}  // namespace pwicp
