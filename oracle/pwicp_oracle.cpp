/*
 * pwicp_oracle.cpp -- CPU ORACLE (test infrastructure; see pwicp_oracle.h for the rules).
 * PARITY PINNED by the reference: the reference's recorded results/4DPCReg matrices on its shipped
 * scans (tests/test_oracle.py, tests/golden/refpair_e2.npz), its own KD-tree compiled into oracle/_ref, and its recorded
 * aggregate files (tests/golden/recorded_4d).
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: the reference was built with MSVC /fp:precise for x64/SSE2, which
 * never fuses a*b+c, and the float expressions below feed comparisons and index decisions.
 *
 * All "[PCL]" comments restate PCL 1.8.1 / FLANN / Eigen behaviour from the published sources
 * (not present in this container); each is tied to the reference call site that triggers it.
 */
#include "../piecewise-icp_b200/host/msvc_sort.h"   /* the Microsoft STL's std::sort order: one copy, shared with the host mirror */
#include "pwicp_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

/* ------------------------------------------------------------------------------------------
 * A1. Exact nearest neighbour.
 * [PCL] pcl::KdTreeFLANN<PointXYZ, flann::L2_Simple<float>>::nearestKSearch(p, 1), exact search
 * (checks = -1, eps = 0).  L2_Simple: result = 0; for c in x,y,z { diff = a[c]-b[c];
 * result += diff*diff; } in float.  Call sites: src/Registration.cpp:737-747, :1293-1297,
 * :597-601, src/CommonFunc.cpp:269-273 and inside pcl::IterativeClosestPoint (:1266).
 * Any exact search returns the same (index, distance) except on exact float ties, where FLANN
 * returns the first visited; the oracle and the CUDA path both return the lowest index.
 * ------------------------------------------------------------------------------------------ */
inline float l2_simple(const float* a, const float* b) {
    float r = 0.0f;
    float d0 = a[0] - b[0];
    r += d0 * d0;
    float d1 = a[1] - b[1];
    r += d1 * d1;
    float d2 = a[2] - b[2];
    r += d2 * d2;
    return r;
}

struct KdTree {
    struct Node {
        int lo, hi;        /* point range [lo,hi) in the reordered arrays (leaf) */
        int dim;           /* -1 = leaf */
        float divlow, divhigh;
        int left, right;
    };
    std::vector<float> pts;  /* reordered xyz */
    std::vector<int> ids;    /* original index of reordered point */
    std::vector<Node> nodes;
    float bbmin[3], bbmax[3];
    int n = 0;
    static constexpr int LEAF = 15; /* [PCL] KDTreeSingleIndexParams(15) */

    void build(const float* p, int n_) {
        n = n_;
        ids.resize(n);
        std::iota(ids.begin(), ids.end(), 0);
        for (int c = 0; c < 3; ++c) { bbmin[c] = FLT_MAX; bbmax[c] = -FLT_MAX; }
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) {
                bbmin[c] = std::min(bbmin[c], p[3 * i + c]);
                bbmax[c] = std::max(bbmax[c], p[3 * i + c]);
            }
        nodes.clear();
        nodes.reserve(2 * (n / 8 + 2));
        if (n > 0) {
            float mn[3] = {bbmin[0], bbmin[1], bbmin[2]}, mx[3] = {bbmax[0], bbmax[1], bbmax[2]};
            divide(p, 0, n, mn, mx);
        }
        pts.resize(3 * (size_t)n);
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) pts[3 * (size_t)i + c] = p[3 * (size_t)ids[i] + c];
    }

    int divide(const float* p, int lo, int hi, float* mn, float* mx) {
        int me = (int)nodes.size();
        nodes.push_back(Node());
        if (hi - lo <= LEAF) {
            nodes[me].lo = lo; nodes[me].hi = hi; nodes[me].dim = -1;
            nodes[me].left = nodes[me].right = -1;
            /* lowest original index first inside a leaf, so strict '<' keeps the lowest on ties */
            std::sort(ids.begin() + lo, ids.begin() + hi);
            return me;
        }
        /* split the widest dimension of the actual point spread at the median */
        float smn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, smx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int i = lo; i < hi; ++i)
            for (int c = 0; c < 3; ++c) {
                float v = p[3 * (size_t)ids[i] + c];
                smn[c] = std::min(smn[c], v); smx[c] = std::max(smx[c], v);
            }
        int dim = 0;
        for (int c = 1; c < 3; ++c) if (smx[c] - smn[c] > smx[dim] - smn[dim]) dim = c;
        int mid = lo + (hi - lo) / 2;
        std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi,
                         [&](int a, int b) {
                             float va = p[3 * (size_t)a + dim], vb = p[3 * (size_t)b + dim];
                             return va < vb || (va == vb && a < b);
                         });
        float divlow = -FLT_MAX, divhigh = FLT_MAX;
        for (int i = lo; i < mid; ++i) divlow = std::max(divlow, p[3 * (size_t)ids[i] + dim]);
        divhigh = FLT_MAX;
        for (int i = mid; i < hi; ++i) divhigh = std::min(divhigh, p[3 * (size_t)ids[i] + dim]);
        nodes[me].dim = dim; nodes[me].divlow = divlow; nodes[me].divhigh = divhigh;
        nodes[me].lo = lo; nodes[me].hi = hi;
        float save = mx[dim];
        mx[dim] = divlow;
        int l = divide(p, lo, mid, mn, mx);
        mx[dim] = save;
        save = mn[dim];
        mn[dim] = divhigh;
        int r = divide(p, mid, hi, mn, mx);
        mn[dim] = save;
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }

    struct Best { float d2; int idx; };

    void search(int ni, const float* q, double mindist, double* dists, Best& best) const {
        const Node& nd = nodes[ni];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) {
                float d = l2_simple(q, &pts[3 * (size_t)i]);
                if (d < best.d2 || (d == best.d2 && ids[i] < best.idx)) { best.d2 = d; best.idx = ids[i]; }
            }
            return;
        }
        int dim = nd.dim;
        double val = q[dim];
        double diff1 = val - (double)nd.divlow, diff2 = val - (double)nd.divhigh;
        int nearc, farc; double cut;
        if (diff1 + diff2 < 0) { nearc = nd.left; farc = nd.right; cut = diff2 * diff2; }
        else                   { nearc = nd.right; farc = nd.left; cut = diff1 * diff1; }
        search(nearc, q, mindist, dists, best);
        double dsave = dists[dim];
        double md = mindist + cut - dsave;
        /* conservative pruning: float-evaluated distances may undershoot the real one by ~2e-7
         * relative, and equal distances must still be visited for the lowest-index tie rule */
        if (md * 0.999999 <= (double)best.d2) {
            dists[dim] = cut;
            search(farc, q, md, dists, best);
            dists[dim] = dsave;
        }
    }

    void query(const float* q, int& idx, float& d2) const {
        Best b{std::numeric_limits<float>::infinity(), INT32_MAX};
        if (n > 0) {
            double dists[3] = {0, 0, 0}, md = 0;
            for (int c = 0; c < 3; ++c) {
                if (q[c] < bbmin[c]) { double d = (double)q[c] - bbmin[c]; dists[c] = d * d; }
                if (q[c] > bbmax[c]) { double d = (double)q[c] - bbmax[c]; dists[c] = d * d; }
                md += dists[c];
            }
            search(0, q, md, dists, b);
        }
        idx = (b.idx == INT32_MAX) ? -1 : b.idx;
        d2 = b.d2;
    }

    /* k nearest neighbours (F4: StatisticalOutlierRemoval's nearestKSearch): best[] holds the k
     * smallest float distances in ascending order; `skip` (an original index, or -1) is left out. */
    void search_k(int ni, const float* q, double mindist, double* dists, int skip, int k, float* best) const {
        const Node& nd = nodes[ni];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) {
                if (ids[i] == skip) continue;
                float d = l2_simple(q, &pts[3 * (size_t)i]);
                if (d < best[k - 1]) {
                    int j = k - 1;
                    while (j > 0 && best[j - 1] > d) { best[j] = best[j - 1]; --j; }
                    best[j] = d;
                }
            }
            return;
        }
        int dim = nd.dim;
        double val = q[dim];
        double diff1 = val - (double)nd.divlow, diff2 = val - (double)nd.divhigh;
        int nearc, farc; double cut;
        if (diff1 + diff2 < 0) { nearc = nd.left; farc = nd.right; cut = diff2 * diff2; }
        else                   { nearc = nd.right; farc = nd.left; cut = diff1 * diff1; }
        search_k(nearc, q, mindist, dists, skip, k, best);
        double dsave = dists[dim];
        double md = mindist + cut - dsave;
        if (md * 0.999999 <= (double)best[k - 1]) {
            dists[dim] = cut;
            search_k(farc, q, md, dists, skip, k, best);
            dists[dim] = dsave;
        }
    }

    /* k nearest neighbours WITH indices in the metric of the reference's codelibrary (F4 segmentation front end,
     * codelibrary/util/metric/squared_euclidean.h:33-36): t += double(a - b) * (a - b) over x, y, z on double coordinates
     * (float inputs widened).  bd / bi ascending; equal distances keep the lower original index first. */
    static double l2_double(const float* a, const float* b) {
        double t = 0.0;
        for (int c = 0; c < 3; ++c) { const double d = (double)a[c] - (double)b[c]; t += d * d; }
        return t;
    }
    void search_kd(int ni, const float* q, double mindist, double* dists, int k, double* bd, int* bi) const {
        const Node& nd = nodes[ni];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) {
                const double d = l2_double(q, &pts[3 * (size_t)i]);
                const int id = ids[i];
                if (d < bd[k - 1] || (d == bd[k - 1] && id < bi[k - 1])) {
                    int j = k - 1;
                    while (j > 0 && (bd[j - 1] > d || (bd[j - 1] == d && bi[j - 1] > id))) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
                    bd[j] = d; bi[j] = id;
                }
            }
            return;
        }
        int dim = nd.dim;
        double val = q[dim];
        double diff1 = val - (double)nd.divlow, diff2 = val - (double)nd.divhigh;
        int nearc, farc; double cut;
        if (diff1 + diff2 < 0) { nearc = nd.left; farc = nd.right; cut = diff2 * diff2; }
        else                   { nearc = nd.right; farc = nd.left; cut = diff1 * diff1; }
        search_kd(nearc, q, mindist, dists, k, bd, bi);
        double dsave = dists[dim];
        double md = mindist + cut - dsave;
        if (md * 0.999999 <= bd[k - 1]) {
            dists[dim] = cut;
            search_kd(farc, q, md, dists, k, bd, bi);
            dists[dim] = dsave;
        }
    }
    void query_kd(const float* q, int k, double* bd, int* bi) const {
        for (int j = 0; j < k; ++j) { bd[j] = std::numeric_limits<double>::infinity(); bi[j] = INT32_MAX; }
        if (n <= 0) return;
        double dists[3] = {0, 0, 0}, md = 0;
        for (int c = 0; c < 3; ++c) {
            if (q[c] < bbmin[c]) { double d = (double)q[c] - bbmin[c]; dists[c] = d * d; }
            if (q[c] > bbmax[c]) { double d = (double)q[c] - bbmax[c]; dists[c] = d * d; }
            md += dists[c];
        }
        search_kd(0, q, md, dists, k, bd, bi);
    }

    void query_k(const float* q, int skip, int k, float* best) const {
        for (int j = 0; j < k; ++j) best[j] = std::numeric_limits<float>::infinity();
        if (n <= 0) return;
        double dists[3] = {0, 0, 0}, md = 0;
        for (int c = 0; c < 3; ++c) {
            if (q[c] < bbmin[c]) { double d = (double)q[c] - bbmin[c]; dists[c] = d * d; }
            if (q[c] > bbmax[c]) { double d = (double)q[c] - bbmax[c]; dists[c] = d * d; }
            md += dists[c];
        }
        search_k(0, q, md, dists, skip, k, best);
    }
};

/* ------------------------------------------------------------------------------------------
 * Small dense algebra in double: 6x6 inverse through partial-pivot LU.
 * [PCL] x = ATA.inverse() * ATb with Eigen fixed-size 6x6 -> PartialPivLU-based inverse, then a
 * matrix-vector product.  Eigen's internal operation order is not reproduced (bit
 * parity with Eigen is not attainable); values agree to ~1e-15 relative.
 * ------------------------------------------------------------------------------------------ */
bool inverse6(const double* A, double* Ainv, double* det_out) {
    double lu[36];
    int piv[6];
    std::memcpy(lu, A, sizeof(lu));
    double det = 1.0;
    for (int i = 0; i < 6; ++i) piv[i] = i;
    bool ok = true;
    for (int k = 0; k < 6; ++k) {
        int p = k; double big = std::fabs(lu[k * 6 + k]);
        for (int r = k + 1; r < 6; ++r) {
            double v = std::fabs(lu[r * 6 + k]);
            if (v > big) { big = v; p = r; }
        }
        if (p != k) {
            for (int c = 0; c < 6; ++c) std::swap(lu[k * 6 + c], lu[p * 6 + c]);
            std::swap(piv[k], piv[p]);
            det = -det;
        }
        double d = lu[k * 6 + k];
        det *= d;
        if (d == 0.0) { ok = false; continue; }
        for (int r = k + 1; r < 6; ++r) {
            double f = lu[r * 6 + k] / d;
            lu[r * 6 + k] = f;
            for (int c = k + 1; c < 6; ++c) lu[r * 6 + c] -= f * lu[k * 6 + c];
        }
    }
    if (det_out) *det_out = det;
    /* solve LU * X = P * I column by column (one reciprocal per pivot, as the device does) */
    double rcp[6];
    for (int r = 0; r < 6; ++r) rcp[r] = 1.0 / lu[r * 6 + r];
    for (int col = 0; col < 6; ++col) {
        double y[6];
        for (int r = 0; r < 6; ++r) {
            double s = (piv[r] == col) ? 1.0 : 0.0;
            for (int c = 0; c < r; ++c) s -= lu[r * 6 + c] * y[c];
            y[r] = s;
        }
        for (int r = 5; r >= 0; --r) {
            double s = y[r];
            for (int c = r + 1; c < 6; ++c) s -= lu[r * 6 + c] * Ainv[c * 6 + col];
            Ainv[r * 6 + col] = s * rcp[r];
        }
    }
    return ok;
}

/* [PCL] TransformationEstimationPointToPlaneLLS::constructTransformationMatrix: double trig,
 * every entry cast to float.  R = Rz(gamma) * Ry(beta) * Rx(alpha). */
void construct_T(const double* x, float* T) {
    double alpha = x[0], beta = x[1], gamma = x[2];
    for (int i = 0; i < 16; ++i) T[i] = 0.0f;
    T[0]  = (float)( std::cos(gamma) * std::cos(beta));
    T[1]  = (float)(-std::sin(gamma) * std::cos(alpha) + std::cos(gamma) * std::sin(beta) * std::sin(alpha));
    T[2]  = (float)( std::sin(gamma) * std::sin(alpha) + std::cos(gamma) * std::sin(beta) * std::cos(alpha));
    T[4]  = (float)( std::sin(gamma) * std::cos(beta));
    T[5]  = (float)( std::cos(gamma) * std::cos(alpha) + std::sin(gamma) * std::sin(beta) * std::sin(alpha));
    T[6]  = (float)(-std::cos(gamma) * std::sin(alpha) + std::sin(gamma) * std::sin(beta) * std::cos(alpha));
    T[8]  = (float)(-std::sin(beta));
    T[9]  = (float)( std::cos(beta) * std::sin(alpha));
    T[10] = (float)( std::cos(beta) * std::cos(alpha));
    T[3]  = (float)x[3];
    T[7]  = (float)x[4];
    T[11] = (float)x[5];
    T[15] = 1.0f;
}

/* Row terms of one correspondence.  [PCL] the right-hand sides are expressions of
 * `const float&`, hence evaluated in float and only then widened to double. */
inline void lls_row(const float* s, const float* d, const float* n, float* u /*7*/) {
    float sx = s[0], sy = s[1], sz = s[2];
    float dx = d[0], dy = d[1], dz = d[2];
    float nx = n[0], ny = n[1], nz = n[2];
    u[0] = nz * sy - ny * sz;
    u[1] = nx * sz - nz * sx;
    u[2] = ny * sx - nx * sy;
    u[3] = nx; u[4] = ny; u[5] = nz;
    u[6] = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz;
}

/* index tables of the 28 accumulated values: 21 upper-triangle ATA entries (row-major),
 * 6 ATb entries, and the sum of squared NN distances (MSE numerator, A6) */
struct ValTab {
    int a[28], b[28];
    ValTab() {
        int v = 0;
        for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { a[v] = r; b[v] = c; ++v; }
        for (int r = 0; r < 6; ++r) { a[v] = r; b[v] = 6; ++v; }
        a[27] = b[27] = -1;
    }
};
const ValTab g_vt;

/* Accumulate the 28 values over all rows.  u7: n x 7 floats, d2: n floats, valid: n flags. */
void accumulate28(const float* u7, const float* d2, const unsigned char* valid, int n,
                  int reduce_mode, int group_batches, int /*unused*/, double* out28) {
    if (reduce_mode == 0) {
        /* [PCL] one sequential loop in correspondence order */
        for (int v = 0; v < 28; ++v) out28[v] = 0.0;
        for (int i = 0; i < n; ++i) {
            const float* u = u7 + 7 * (size_t)i;
            if (valid[i])
                for (int v = 0; v < 27; ++v) out28[v] += (double)u[g_vt.a[v]] * (double)u[g_vt.b[v]];
            out28[27] += (double)d2[i];
        }
        return;
    }
    if (reduce_mode == 2) {
        /* Summation order of the round-2 CUDA kernel (DESIGN.md 3.2 "reduction geometry").  group_batches = total
         * warps of the grid NWT | warps per CTA WPC << 16.  Batch b (rows 32 b .. 32 b + 31) belongs to warp
         * b mod NWT; lane l of that warp adds row l of its batches b = w, w + NWT, ... in that order to 28 sums of
         * its own (fused multiply-add: the product of two float values is exact in double); the 32 lanes are
         * folded as a balanced tree with partner distances 16, 8, 4, 2, 1; a CTA adds its WPC warps in order; the
         * CTA sums are added in chunks of ceil(G / WPC) consecutive CTAs, each from 0, and the chunks in order. */
        const long nb = ((long)n + 31) / 32;
        const int NWT = group_batches & 0xffff, WPC = (group_batches >> 16) & 0x7fff;
        for (int v = 0; v < 28; ++v) out28[v] = 0.0;
        if (NWT < 1 || WPC < 1 || NWT % WPC) return;
        const int G = NWT / WPC;
        std::vector<double> cta((size_t)G * 28, 0.0);
        std::vector<double> lanes(32 * 28);
        for (int w = 0; w < NWT; ++w) {
            std::fill(lanes.begin(), lanes.end(), 0.0);
            for (long b = w; b < nb; b += NWT)
                for (int l = 0; l < 32; ++l) {
                    const long i = b * 32 + l;
                    if (i >= n) break;
                    double* acc = &lanes[(size_t)l * 28];
                    const float* u = u7 + 7 * (size_t)i;
                    if (valid[i])
                        for (int v = 0; v < 27; ++v) acc[v] = std::fma((double)u[g_vt.a[v]], (double)u[g_vt.b[v]], acc[v]);
                    acc[27] += (double)d2[i];
                }
            for (int stride = 16; stride >= 1; stride >>= 1)
                for (int l = 0; l < stride; ++l)
                    for (int v = 0; v < 28; ++v) lanes[(size_t)l * 28 + v] += lanes[(size_t)(l + stride) * 28 + v];
            /* CTA sum: warps in order, starting from 0 */
            double* c = &cta[(size_t)(w / WPC) * 28];
            for (int v = 0; v < 28; ++v) c[v] += lanes[v];
        }
        const int per = (G + WPC - 1) / WPC;
        for (int v = 0; v < 28; ++v) {
            double s = 0.0;
            for (int g0 = 0; g0 < G; g0 += per) {
                double ch = 0.0;
                for (int g = g0; g < std::min(G, g0 + per); ++g) ch += cta[(size_t)g * 28 + v];
                s += ch;
            }
            out28[v] = s;
        }
        return;
    }
    /* Summation order of the round-1 CUDA kernel: per 32-point batch the
     * rows in order starting from 0; then a hierarchy with fan-in `group_batches` (default 32):
     * every parent is the sum of its consecutive children in order, starting from 0; after at most
     * two such levels (or once at most fan-in entries remain) the entries are summed in order into
     * the grand total. */
    const long nb = ((long)n + 31) / 32;
    /* group_batches: low 16 bits = fan-in of every level; bits 16..30 (optional) = a different fan-in for the first level
     * (batches -> groups), for kernels that sum a short run of consecutive batches in registers before the hierarchy */
    const int fan_hi = group_batches & 0xffff, fan_lo = (group_batches >> 16) & 0x7fff;
    const int fan = fan_hi > 1 ? fan_hi : 32;                 /* a fan-in below 2 is meaningless */
    const int fan0 = fan_lo > 1 ? fan_lo : fan;
    std::vector<double> cur((size_t)nb * 28, 0.0);
    for (long b = 0; b < nb; ++b) {
        double* acc = &cur[(size_t)b * 28];
        for (int r = 0; r < 32; ++r) {
            long i = b * 32 + r;
            if (i >= n) break;
            const float* u = u7 + 7 * (size_t)i;
            if (valid[i])
                for (int v = 0; v < 27; ++v) acc[v] += (double)u[g_vt.a[v]] * (double)u[g_vt.b[v]];
            acc[27] += (double)d2[i];
        }
    }
    long count = nb;
    for (int level = 1; level < 3 && count > (level == 1 ? fan0 : fan); ++level) {
        const long f = (level == 1) ? fan0 : fan;
        const long np = (count + f - 1) / f;
        std::vector<double> nxt((size_t)np * 28, 0.0);
        for (long g = 0; g < np; ++g) {
            const long size = std::min(f, count - g * f);
            for (int v = 0; v < 28; ++v) {
                double s = 0.0;
                for (long k = 0; k < size; ++k) s += cur[(size_t)(g * f + k) * 28 + v];
                nxt[(size_t)g * 28 + v] = s;
            }
        }
        cur.swap(nxt);
        count = np;
    }
    for (int v = 0; v < 28; ++v) {
        double s = 0.0;
        for (long k0 = 0; k0 < count; k0 += fan) {       /* chunks of fan-in entries, each from 0 */
            double c = 0.0;
            for (long k = k0; k < std::min(count, k0 + fan); ++k) c += cur[(size_t)k * 28 + v];
            s += c;
        }
        out28[v] = s;
    }
}

void solve_from28(const double* s28, double* ATA, double* ATb, double* x, float* T) {
    int v = 0;
    for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { ATA[r * 6 + c] = s28[v]; ATA[c * 6 + r] = s28[v]; ++v; }
    for (int r = 0; r < 6; ++r) ATb[r] = s28[v++];
    double inv[36];
    inverse6(ATA, inv, nullptr);
    for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int c = 0; c < 6; ++c) s += inv[r * 6 + c] * ATb[c];
        x[r] = s;
    }
    construct_T(x, T);
}

inline bool finite3(const float* p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); }

/* Eigen 4x4 float product, coefficient = sum over k in order (no FMA). */
void mat4_mul(const float* A, const float* B, float* C) {
    float R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = A[i * 4 + 0] * B[0 * 4 + j];
            s += A[i * 4 + 1] * B[1 * 4 + j];
            s += A[i * 4 + 2] * B[2 * 4 + j];
            s += A[i * 4 + 3] * B[3 * 4 + j];
            R[i * 4 + j] = s;
        }
    std::memcpy(C, R, sizeof(R));
}

void set_identity(float* T) { for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0f : 0.0f; }

orc_icp_params default_icp() {
    orc_icp_params p;
    p.max_iter = 100; p.tf_eps = 1e-8; p.fit_eps = 1e-6; p.force_iters = 0;
    p.reduce_mode = 0; p.group_batches = 0; p.reserved = 0; p.rot_thr_default = 0;
    return p;
}

/* Threads for the INDEPENDENT nearest-neighbour queries of the outer iteration, the percentile and the VCM (1 = a plain
 * loop, what the reference does).  Only which thread answers a query changes; every query, every sum and every
 * selection is the same, so the results are bit-identical for any thread count (tests/test_oracle.py). */
static int g_threads = 1;
extern "C" void orc_set_threads(int threads) { g_threads = threads > 1 ? threads : 1; }

/* i = 0..n-1 on `threads` threads in contiguous chunks (1: a plain loop, what the reference does) */
template <typename F>
void parallel_for(int n, int threads, F&& f) {
    if (threads <= 1 || n < 1024) { for (int i = 0; i < n; ++i) f(i); return; }
    std::vector<std::thread> pool;
    const int chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        pool.emplace_back([lo, hi, &f]() { for (int i = lo; i < hi; ++i) f(i); });
    }
    for (auto& th : pool) th.join();
}

/* A3/A6 on a prebuilt target tree */
int icp_run(const KdTree& tree, const float* tgt, const float* nrm, const float* src, int n2s,
            const orc_icp_params& prm, float* Tfinal, int* n_iter, int* conv_state,
            double* mse_trace, float* T_trace, int* idx_trace) {
    std::vector<float> cur(src, src + 3 * (size_t)n2s);   /* input_transformed = *source */
    std::vector<int> idx(n2s);
    std::vector<float> d2(n2s), u7(7 * (size_t)n2s);
    std::vector<unsigned char> valid(n2s);
    set_identity(Tfinal);
    int iters = 0, state = 0;
    /* [PCL] DefaultConvergenceCriteria wiring in IterativeClosestPoint::computeTransformation */
    const double rot_thr = prm.rot_thr_default ? 0.99999 : (1.0 - prm.tf_eps);
    const double transl_thr = prm.tf_eps;
    const double mse_rel = prm.fit_eps, mse_abs = 1e-12;
    double prev_mse = std::numeric_limits<double>::max();
    if (n2s < 3) { *n_iter = 0; *conv_state = 5; return -1; }   /* min_number_correspondences_ = 3 */
    for (;;) {
        /* [reference: single thread]  prm.reserved > 1: the independent queries and rows on that many
         * threads (bench.py's multi-thread CPU figure); every sum stays in its sequential order. */
        const int threads = prm.reserved > 1 ? prm.reserved : 1;
        parallel_for(n2s, threads, [&](int i) { tree.query(&cur[3 * (size_t)i], idx[i], d2[i]); });
        if (idx_trace) std::memcpy(idx_trace + (size_t)iters * n2s, idx.data(), sizeof(int) * (size_t)n2s);
        parallel_for(n2s, threads, [&](int i) {
            const float* s = &cur[3 * (size_t)i];
            const float* d = tgt + 3 * (size_t)idx[i];
            const float* n = nrm + 3 * (size_t)idx[i];
            valid[i] = finite3(s) && finite3(d) && finite3(n);
            lls_row(s, d, n, &u7[7 * (size_t)i]);
        });
        double s28[28], ATA[36], ATb[6], x[6];
        float T[16];
        accumulate28(u7.data(), d2.data(), valid.data(), n2s, prm.reduce_mode, prm.group_batches, 0, s28);
        solve_from28(s28, ATA, ATb, x, T);
        orc_transform(cur.data(), n2s, T);            /* transformCloud(input_transformed, ..) */
        mat4_mul(T, Tfinal, Tfinal);                  /* final = T * final */
        if (T_trace) std::memcpy(T_trace + 16 * (size_t)iters, T, sizeof(T));
        double mse = s28[27] / (double)n2s;
        if (mse_trace) mse_trace[iters] = mse;
        ++iters;
        /* [PCL] DefaultConvergenceCriteria<float>::hasConverged(), in this order */
        if (iters >= prm.max_iter) { state = 1; break; }
        if (!prm.force_iters) {
            double cos_angle = 0.5 * (double)(T[0] + T[5] + T[10] - 1);
            double transl_sq = (double)(T[3] * T[3] + T[7] * T[7] + T[11] * T[11]);
            if (cos_angle >= rot_thr && transl_sq <= transl_thr) { state = 2; break; }
            if (std::fabs(mse - prev_mse) < mse_abs) { state = 3; break; }
            if (std::fabs(mse - prev_mse) / prev_mse < mse_rel) { state = 4; break; }
        }
        prev_mse = mse;
    }
    *n_iter = iters; *conv_state = state;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A9. [PCL] pcl::computePointNormal -> computeMeanAndCovarianceMatrix (single pass, float) ->
 * solvePlaneParameters -> pcl::eigen33 (closed-form roots + best cross product).
 * ------------------------------------------------------------------------------------------ */
void compute_roots2(float b, float c, float* roots) {
    roots[0] = 0.0f;
    float d = (float)(b * b - 4.0 * c);
    if (d < 0.0) d = 0.0f;
    float sd = std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

void compute_roots(const float m[3][3], float* roots) {
    float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2]
             - m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2]
             + m[1][1] * m[2][2] - m[1][2] * m[1][2];
    float c2 = m[0][0] + m[1][1] + m[2][2];
    if (std::fabs(c0) < FLT_EPSILON) { compute_roots2(c2, c1, roots); return; }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
    float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
        std::swap(roots[1], roots[2]);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0) compute_roots2(c2, c1, roots);
}

inline void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

void eigen33_smallest(const float C[3][3], float* evec) {
    float scale = 0.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(C[i][j]));
    if (scale <= FLT_MIN) scale = 1.0f;
    float m[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = C[i][j] / scale;
    float roots[3];
    compute_roots(m, roots);
    for (int i = 0; i < 3; ++i) m[i][i] -= roots[0];
    float v1[3], v2[3], v3[3];
    cross3(m[0], m[1], v1); cross3(m[0], m[2], v2); cross3(m[1], m[2], v3);
    float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
    const float* v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    float s = std::sqrt(l);
    evec[0] = v[0] / s; evec[1] = v[1] / s; evec[2] = v[2] / s;
}

/* fallback of src/CommonFunc.cpp:303-326: smallest singular vector of the centred covariance.
 * The reference uses Eigen::JacobiSVD<Matrix3f>; restated as cyclic Jacobi on the symmetric 3x3
 * (same subspace; not bit-identical -- this branch needs |len-1| >= 1e-5 and is not reached by
 * any data set in the repository). */
void jacobi_smallest(float A[3][3], float* evec) {
    float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        float off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        if (off < 1e-30f) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0f) continue;
                float theta = (A[q][q] - A[p][p]) / (2.0f * A[p][q]);
                float t = (theta >= 0 ? 1.0f : -1.0f) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0f));
                float c = 1.0f / std::sqrt(t * t + 1.0f), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    float akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    float apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    float vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    for (int i = 1; i < 3; ++i) if (A[i][i] < A[m][m]) m = i;
    evec[0] = V[0][m]; evec[1] = V[1][m]; evec[2] = V[2][m];
}

}  // namespace

extern "C" {

int orc_nn_brute(const float* tgt, int n1, const float* qry, int nq, int* idx, float* d2) {
    for (int i = 0; i < nq; ++i) {
        float best = std::numeric_limits<float>::infinity(); int bi = -1;
        for (int j = 0; j < n1; ++j) {
            float d = l2_simple(qry + 3 * (size_t)i, tgt + 3 * (size_t)j);
            if (d < best) { best = d; bi = j; }   /* ascending j + strict '<' = lowest index on ties */
        }
        idx[i] = bi; d2[i] = best;
    }
    return 0;
}

void* orc_tree_build(const float* tgt, int n1) { KdTree* t = new KdTree(); t->build(tgt, n1); return t; }
void orc_tree_query(const void* tree, const float* qry, int nq, int* idx, float* d2) {
    const KdTree* t = (const KdTree*)tree;
    for (int i = 0; i < nq; ++i) t->query(qry + 3 * (size_t)i, idx[i], d2[i]);
}
void orc_tree_free(void* tree) { delete (KdTree*)tree; }

int orc_nn(const float* tgt, int n1, const float* qry, int nq, int* idx, float* d2) {
    KdTree t; t.build(tgt, n1);
    for (int i = 0; i < nq; ++i) t.query(qry + 3 * (size_t)i, idx[i], d2[i]);
    return 0;
}

/* A5 [PCL] pcl::transformPointCloud / transformPointCloudWithNormals point formula:
 * x' = m00*x + m01*y + m02*z + m03, float, left to right. */
void orc_transform(float* pts, int n, const float* T) {
    for (int i = 0; i < n; ++i) {
        float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        pts[3 * (size_t)i]     = T[0] * x + T[1] * y + T[2] * z + T[3];
        pts[3 * (size_t)i + 1] = T[4] * x + T[5] * y + T[6] * z + T[7];
        pts[3 * (size_t)i + 2] = T[8] * x + T[9] * y + T[10] * z + T[11];
    }
}

void orc_mat4_mul(const float* A, const float* B, float* C) { mat4_mul(A, B, C); }

int orc_lls_step(const float* src, const int* match, int n, const float* tgt, const float* nrm,
                 int reduce_mode, int group_batches, int reserved,
                 double* ATA36, double* ATb6, double* x6, float* T16) {
    std::vector<float> u7(7 * (size_t)n), d2(n);
    std::vector<unsigned char> valid(n);
    for (int i = 0; i < n; ++i) {
        const float* s = src + 3 * (size_t)i;
        const float* d = tgt + 3 * (size_t)match[i];
        const float* nn = nrm + 3 * (size_t)match[i];
        valid[i] = finite3(s) && finite3(d) && finite3(nn);
        lls_row(s, d, nn, &u7[7 * (size_t)i]);
        d2[i] = l2_simple(s, d);
    }
    double s28[28];
    accumulate28(u7.data(), d2.data(), valid.data(), n, reduce_mode, group_batches, reserved, s28);
    solve_from28(s28, ATA36, ATb6, x6, T16);
    return 0;
}

int orc_icp_p2plane(const float* tgt, const float* nrm, int n1, const float* src, int n2s,
                    const orc_icp_params* prm, float* T_final16, int* n_iter, int* conv_state,
                    double* mse_trace, float* T_trace, int* idx_trace) {
    KdTree tree; tree.build(tgt, n1);
    orc_icp_params p = prm ? *prm : default_icp();
    int it = 0, st = 0;
    int rc = icp_run(tree, tgt, nrm, src, n2s, p, T_final16, &it, &st, mse_trace, T_trace, idx_trace);
    if (n_iter) *n_iter = it;
    if (conv_state) *conv_state = st;
    return rc;
}

/* A7 [PCL] OctreePointCloud::defineBoundingBox() + getKeyBitSize() + getBoundingBox()
 * (src/Registration.cpp:881-886) */
void orc_octree_bbox(const float* pts, int n, double res, double* bb) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            float v = pts[3 * (size_t)i + c];
            mn[c] = std::min(mn[c], v); mx[c] = std::max(mx[c], v);
        }
    const float minValue512 = std::numeric_limits<float>::epsilon() * 512.0f;
    double lo[3], hi[3];
    for (int c = 0; c < 3; ++c) { lo[c] = mn[c]; hi[c] = (float)(mx[c] + minValue512); }
    const float minValue = std::numeric_limits<float>::epsilon();
    unsigned int key[3];
    for (int c = 0; c < 3; ++c) key[c] = (unsigned int)std::ceil((hi[c] - lo[c] - minValue) / res);
    unsigned int max_voxels = std::max(std::max(std::max(key[0], key[1]), key[2]), 2u);
    unsigned int depth = (unsigned int)std::ceil(std::log((double)max_voxels) / std::log(2.0) - minValue);
    depth = std::min(depth, 32u);
    double side = (double)(1u << depth) * res;
    for (int c = 0; c < 3; ++c) {
        double over = (side - (hi[c] - lo[c])) / 2.0;
        if (over > minValue) { lo[c] -= over; hi[c] += over; }
    }
    bb[0] = lo[0]; bb[1] = lo[1]; bb[2] = lo[2]; bb[3] = hi[0]; bb[4] = hi[1]; bb[5] = hi[2];
}

/* calBoundingBoxCornerChange (src/CommonFunc.cpp:410-419), Eigen float 4x4 * 4-vector */
float orc_bbox_corner_change(const double* bb, const float* T) {
    float best = 0.0f;
    for (int k = 0; k < 2; ++k) {
        float c[4] = {(float)bb[3 * k], (float)bb[3 * k + 1], (float)bb[3 * k + 2], 1.0f};
        float d[3];
        for (int r = 0; r < 3; ++r) {
            float s = T[r * 4] * c[0];
            s += T[r * 4 + 1] * c[1];
            s += T[r * 4 + 2] * c[2];
            s += T[r * 4 + 3] * c[3];
            d[r] = s - c[r];
        }
        float nrm = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        best = (k == 0) ? nrm : std::max(best, nrm);
    }
    return best;
}

/* calPercentileDistBetween2PC (src/CommonFunc.cpp:266-281): sqrt of the float squared distance
 * (std::sqrt(float) under `using namespace std`), stored as double, sorted ascending
 * (QuickSort :145-170; only the value at [int(n*percentile)] matters, :177-178). */
double orc_percentile_nn(const float* cloud1, int m1, const float* cloud2, int m2, float pct) {
    KdTree t; t.build(cloud1, m1);
    std::vector<double> dist(m2);
    parallel_for(m2, g_threads, [&](int i) {
        int j; float d2;
        t.query(cloud2 + 3 * (size_t)i, j, d2);
        dist[i] = std::sqrt(d2);
    });
    int leftnum = m2 * pct;                          /* int = int * float, :177 */
    if (leftnum >= m2) leftnum = m2 - 1;             /* the reference would read out of bounds */
    std::nth_element(dist.begin(), dist.begin() + leftnum, dist.end());
    return dist[leftnum];
}

/* A8 calTransParaVCM (src/Registration.cpp:1273-1343), straightforward double evaluation */
int orc_vcm(const float* tgt, const float* nrm, int n1, const float* src, int n2s,
            double* vcm36, int* singular) {
    KdTree t; t.build(tgt, n1);
    std::vector<double> A(6 * (size_t)n2s), L(n2s);
    for (int i = 0; i < n2s; ++i) {
        int j; float d2;
        t.query(src + 3 * (size_t)i, j, d2);
        double Qx = src[3 * (size_t)i], Qy = src[3 * (size_t)i + 1], Qz = src[3 * (size_t)i + 2];
        double Px = tgt[3 * (size_t)j], Py = tgt[3 * (size_t)j + 1], Pz = tgt[3 * (size_t)j + 2];
        double Nx = nrm[3 * (size_t)j], Ny = nrm[3 * (size_t)j + 1], Nz = nrm[3 * (size_t)j + 2];
        double* a = &A[6 * (size_t)i];
        a[0] = Nz * Qy - Ny * Qz; a[1] = Nx * Qz - Nz * Qx; a[2] = Ny * Qx - Nx * Qy;
        a[3] = Nx; a[4] = Ny; a[5] = Nz;
        L[i] = Nx * (Px - Qx) + Ny * (Py - Qy) + Nz * (Pz - Qz);
    }
    double ATA[36] = {0}, ATL[6] = {0};
    for (int i = 0; i < n2s; ++i) {
        const double* a = &A[6 * (size_t)i];
        for (int r = 0; r < 6; ++r) {
            for (int c = 0; c < 6; ++c) ATA[r * 6 + c] += a[r] * a[c];
            ATL[r] += a[r] * L[i];
        }
    }
    double Q[36], det = 0;
    inverse6(ATA, Q, &det);
    if (singular) *singular = (std::fabs(det) < 1e-9) ? 1 : 0;
    double X[6];
    for (int r = 0; r < 6; ++r) { double s = 0; for (int c = 0; c < 6; ++c) s += Q[r * 6 + c] * ATL[c]; X[r] = s; }
    double vtpv = 0;
    for (int i = 0; i < n2s; ++i) {
        const double* a = &A[6 * (size_t)i];
        double v = 0; for (int c = 0; c < 6; ++c) v += a[c] * X[c];
        v -= L[i];
        vtpv += v * v;
    }
    double STD0 = 1 * std::sqrt(vtpv / double(n2s - 6));
    for (int k = 0; k < 36; ++k) vcm36[k] = STD0 * STD0 * Q[k];
    return 0;
}

/* A9 calPatchNormal (src/CommonFunc.cpp:284-333) */
int orc_patch_normal(const float* pts, int n, float* n3) {
    if (!(n > 4)) { n3[0] = 0; n3[1] = 0; n3[2] = 1; return 0; }
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
        accu[0] += x * x; accu[1] += x * y; accu[2] += x * z;
        accu[3] += y * y; accu[4] += y * z; accu[5] += z * z;
        accu[6] += x; accu[7] += y; accu[8] += z;
    }
    for (int k = 0; k < 9; ++k) accu[k] /= (float)n;
    float C[3][3];
    C[0][0] = accu[0] - accu[6] * accu[6];
    C[0][1] = accu[1] - accu[6] * accu[7];
    C[0][2] = accu[2] - accu[6] * accu[8];
    C[1][1] = accu[3] - accu[7] * accu[7];
    C[1][2] = accu[4] - accu[7] * accu[8];
    C[2][2] = accu[5] - accu[8] * accu[8];
    C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
    float v[3];
    eigen33_smallest(C, v);
    float nLen = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (std::fabs(nLen - 1.0) < 1e-5) { n3[0] = v[0]; n3[1] = v[1]; n3[2] = v[2]; return 1; }
    /* recalculation branch (:303-326): centred covariance / Pn, smallest singular vector */
    float mean[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) mean[c] += pts[3 * (size_t)i + c];
    for (int c = 0; c < 3; ++c) mean[c] /= (float)n;
    float M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < n; ++i) {
        float d[3];
        for (int c = 0; c < 3; ++c) d[c] = pts[3 * (size_t)i + c] - mean[c];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] += d[r] * d[c];
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] /= (float)n;
    jacobi_smallest(M, n3);
    float nLen2 = std::sqrt(n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]);
    return (std::fabs(nLen2 - 1.0) < 1e-5) ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * F3. Patch constants next to the normal: calPatchCTandBP (src/Segmentation.cpp:260-303),
 * calPatchSTD (src/CommonFunc.cpp:336-354) and calBPandCTSTD (src/Segmentation.cpp:306-321).
 * [PCL] compute3DCentroid<PointXYZ, float>: float sums in point order, divided by n.
 * [PCL] pcl::PCA (1.8.1): float mean, demeaned cloud, covariance = D D^T / (n - 1) in float,
 * SelfAdjointEigenSolver<Matrix3f>, eigenvectors by descending eigenvalue -> col(2) = smallest.
 * Eigen's float solver is not reproduced; the eigenvector comes from a double Jacobi sweep of the
 * float covariance (agrees to ~1e-7, compared with a tolerance).
 * [PCL] pointToPlaneDistance(p, a, b, c, d) = |a x + b y + c z + d| / sqrt(a^2 + b^2 + c^2) in double.
 * ------------------------------------------------------------------------------------------ */
void orc_patch_ct_bp(const float* pts, int n, float* ct3, float* bp18) {
    float s[3] = {0, 0, 0};
    float e[6][3] = {{-FLT_MAX, 0, 0}, {FLT_MAX, 0, 0}, {0, -FLT_MAX, 0}, {0, FLT_MAX, 0}, {0, 0, -FLT_MAX}, {0, 0, FLT_MAX}};
    for (int i = 0; i < n; ++i) {
        const float* p = pts + 3 * (size_t)i;
        for (int c = 0; c < 3; ++c) s[c] += p[c];
        for (int c = 0; c < 3; ++c) {
            if (p[c] > e[2 * c][c]) std::memcpy(e[2 * c], p, 12);           /* :282-293, strict comparisons */
            if (p[c] < e[2 * c + 1][c]) std::memcpy(e[2 * c + 1], p, 12);
        }
    }
    for (int c = 0; c < 3; ++c) ct3[c] = s[c] / (float)n;
    std::memcpy(bp18, e, sizeof(e));                                         /* Xmax Xmin Ymax Ymin Zmax Zmin, :295-300 */
}

float orc_patch_std(const float* pts, int n) {
    float mean[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) mean[c] += pts[3 * (size_t)i + c];
    for (int c = 0; c < 3; ++c) mean[c] /= (float)n;
    float M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < n; ++i) {
        float d[3];
        for (int c = 0; c < 3; ++c) d[c] = pts[3 * (size_t)i + c] - mean[c];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] += d[r] * d[c];
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] /= (float)(n - 1);
    float nv[3];
    jacobi_smallest(M, nv);
    const float A = nv[0], B = nv[1], C = nv[2];
    const float D = -(A * mean[0] + B * mean[1] + C * mean[2]);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const float* p = pts + 3 * (size_t)i;
        const double dist = std::fabs((double)A * p[0] + (double)B * p[1] + (double)C * p[2] + (double)D) /
                            std::sqrt((double)A * A + (double)B * B + (double)C * C);
        acc += dist * dist;
    }
    return (float)std::sqrt(acc / double(n - 1));
}

/* all patch constants of a cloud: points packed patch by patch, off[np + 1] */
void orc_patch_stats(const float* pts, const int* off, int np, float* ct, float* bp, float* nrm,
                     unsigned char* ok, float* bpstd, float* ctstd) {
    for (int i = 0; i < np; ++i) {
        const float* p = pts + 3 * (size_t)off[i];
        const int n = off[i + 1] - off[i];
        orc_patch_ct_bp(p, n, ct + 3 * (size_t)i, bp + 18 * (size_t)i);
        ok[i] = (unsigned char)orc_patch_normal(p, n, nrm + 3 * (size_t)i);
        bpstd[i] = orc_patch_std(p, n);                  /* BPstd = sigma, :316 */
        ctstd[i] = bpstd[i] / (float)n;                  /* CTstd = sigma / n, :319 */
    }
}

/* A10 matrix2angle (src/CommonFunc.cpp:385-407) */
void orc_matrix2angle(const float* T, float* ang) {
    double ax, ay, az;
    if (T[8] == 1 || T[8] == -1) {
        az = 0;
        double dlta = std::atan2(T[1], T[2]);
        if (T[8] == -1) { ay = M_PI / 2; ax = az + dlta; }
        else { ay = -M_PI / 2; ax = -az + dlta; }
    } else {
        ay = -std::asin(T[8]);
        ax = std::atan2(T[9] / std::cos(ay), T[10] / std::cos(ay));
        az = std::atan2(T[4] / std::cos(ay), T[0] / std::cos(ay));
    }
    ang[0] = (float)ax; ang[1] = (float)ay; ang[2] = (float)az;
}

/* PwICP_singleIteration (src/Registration.cpp:704-972), statement order preserved. */
int orc_single_iteration(orc_pair* pr, orc_state* st, const orc_icp_params* icp_in,
                         float* T16, double* vcm36, int* vcm_written,
                         unsigned char* stable_flags, orc_iter_stats* stats) {
    const int SVnumPC2 = pr->n2;                                       /* :722 */
    float& currDT = st->currDT;
    const float DTmin = pr->DTmin;
    if (vcm_written) *vcm_written = 0;
    if (currDT <= DTmin) currDT = DTmin;                               /* :724-725 */
    if (4 > SVnumPC2) return -1;                                       /* :728-731 */

    /* (1) correspondences CT2->CT1 and BP2->CT1, :737-747 */
    KdTree treeCT; treeCT.build(pr->ct1, pr->n1);
    std::vector<int> ctIdx(SVnumPC2), bpIdx(6 * (size_t)SVnumPC2);
    std::vector<float> ctD2(SVnumPC2), bpD2(6 * (size_t)SVnumPC2);
    parallel_for(SVnumPC2, g_threads, [&](int i) { treeCT.query(pr->ct2 + 3 * (size_t)i, ctIdx[i], ctD2[i]); });
    parallel_for(6 * SVnumPC2, g_threads, [&](int i) { treeCT.query(pr->bp2 + 3 * (size_t)i, bpIdx[i], bpD2[i]); });

    /* (2) LoDetection per patch, :750-769 */
    float max2minLoD = 2.0;
    float maxLoD = DTmin * max2minLoD;
    float minLoD = DTmin;
    std::vector<float> LoDet(SVnumPC2);
    for (int i = 0; i < SVnumPC2; ++i) {
        float sigm1 = pr->ctstd1[ctIdx[i]];
        float sigm2 = pr->bpstd2[i];
        float LoD = 1.96 * std::sqrt(sigm1 * sigm1 + sigm2 * sigm2);   /* double product -> float */
        if (LoD > maxLoD) LoDet[i] = maxLoD;
        else if (LoD < minLoD) LoDet[i] = minLoD;
        else LoDet[i] = LoD;
    }
    float LoDet_min = *std::min_element(LoDet.begin(), LoDet.end());
    float LoDet_max = *std::max_element(LoDet.begin(), LoDet.end());

    /* (3) correspondence distances, :774-812 */
    auto nrm_ok = [&](int j) { return pr->nrm1_ok ? pr->nrm1_ok[j] != 0 : true; };
    std::vector<float> Pt2Pl_CT(SVnumPC2), Pt2Pt_CT(SVnumPC2), Pt2Pl_BP(6 * (size_t)SVnumPC2);
    for (int i = 0; i < SVnumPC2; ++i) {
        int j = ctIdx[i];
        float resDis;
        if (nrm_ok(j)) {
            float DisDx = pr->ct1[3 * (size_t)j] - pr->ct2[3 * (size_t)i];
            float DisDy = pr->ct1[3 * (size_t)j + 1] - pr->ct2[3 * (size_t)i + 1];
            float DisDz = pr->ct1[3 * (size_t)j + 2] - pr->ct2[3 * (size_t)i + 2];
            const float* nm = pr->nrm1 + 3 * (size_t)j;
            resDis = std::fabs(DisDx * nm[0] + DisDy * nm[1] + DisDz * nm[2]);
        } else resDis = std::sqrt(ctD2[i]);
        Pt2Pl_CT[i] = resDis;
        Pt2Pt_CT[i] = std::sqrt(ctD2[i]);
    }
    for (int i = 0; i < 6 * SVnumPC2; ++i) {
        int j = bpIdx[i];
        float resDis;
        if (nrm_ok(j)) {
            float DisDx = pr->ct1[3 * (size_t)j] - pr->bp2[3 * (size_t)i];
            float DisDy = pr->ct1[3 * (size_t)j + 1] - pr->bp2[3 * (size_t)i + 1];
            float DisDz = pr->ct1[3 * (size_t)j + 2] - pr->bp2[3 * (size_t)i + 2];
            const float* nm = pr->nrm1 + 3 * (size_t)j;
            resDis = std::fabs(DisDx * nm[0] + DisDy * nm[1] + DisDz * nm[2]);
        } else resDis = std::sqrt(bpD2[i]);
        Pt2Pl_BP[i] = resDis;
    }

    /* (4) classification, :815-871 */
    float DTctct = currDT + 1 * (pr->SVRes1 + pr->SVRes2);
    std::vector<float> stableCT2;             /* pre-update stable centroids (copied at :868) */
    std::vector<float> stablePC2;
    std::vector<unsigned char> flags(SVnumPC2);
    int BPidx = 0;
    for (int i = 0; i < SVnumPC2; ++i) {
        bool BPdisPass = true;
        for (int k = 0; k < 6; ++k) {
            if (currDT <= LoDet[i]) { if (LoDet[i] < Pt2Pl_BP[BPidx + k]) BPdisPass = false; }
            else { if (currDT < Pt2Pl_BP[BPidx + k]) BPdisPass = false; }
        }
        BPidx += 6;
        bool CTdisPass = true;
        if (currDT <= LoDet[i]) { if (LoDet[i] < Pt2Pl_CT[i]) CTdisPass = false; }
        else { if (currDT < Pt2Pl_CT[i]) CTdisPass = false; }
        bool stable = CTdisPass && BPdisPass && (Pt2Pt_CT[i] < DTctct);
        flags[i] = stable ? 1 : 0;
        if (stable) {
            stableCT2.insert(stableCT2.end(), pr->ct2 + 3 * (size_t)i, pr->ct2 + 3 * (size_t)i + 3);
            stablePC2.insert(stablePC2.end(), pr->patch_pts2 + 3 * (size_t)pr->patch_off2[i],
                             pr->patch_pts2 + 3 * (size_t)pr->patch_off2[i + 1]);
        }
    }
    if (stable_flags) std::memcpy(stable_flags, flags.data(), SVnumPC2);
    const int nStable = (int)(stableCT2.size() / 3);
    if (stats) {
        stats->n_stable = nStable; stats->n_stable_pts = (int)(stablePC2.size() / 3);
        stats->LoDet_min = LoDet_min; stats->LoDet_max = LoDet_max;
        stats->P75 = std::numeric_limits<double>::quiet_NaN();
        stats->icp_iters = 0; stats->icp_state = 0; stats->maxBBchange = 0;
    }
    if (4 > nStable) return -2;                                        /* :864-867 */

    /* (5) point-to-plane ICP of the stable centroids against ALL target centroids, :877 */
    orc_icp_params icp = icp_in ? *icp_in : default_icp();
    float transMatICP[16];
    int it = 0, cs = 0;
    icp_run(treeCT, pr->ct1, pr->nrm1, stableCT2.data(), nStable, icp, transMatICP, &it, &cs,
            nullptr, nullptr, nullptr);

    /* (6) bounding-box corner change on the CURRENT cloud2, :880-888 */
    double BoundingBox[6];
    orc_octree_bbox(pr->cloud2, pr->m2, (double)(float)(pr->Res2 * 2), BoundingBox);
    float maxBBchange = orc_bbox_corner_change(BoundingBox, transMatICP);
    if (stats) {
        stats->icp_iters = it; stats->icp_state = cs; stats->maxBBchange = maxBBchange;
        std::memcpy(stats->bb6, BoundingBox, sizeof(BoundingBox));
    }

    /* (7) DT update, :891-935 (the fall-through from the stage-1 block into the stage-2 block in
     * the same call is the reference's behaviour) */
    if (!st->toStage2 && maxBBchange < minLoD) st->toStage2 = 1;
    else if (currDT == LoDet_min) st->toStage3 = 1;

    if (!st->toStage2) {
        double Dist75 = orc_percentile_nn(pr->cloud1, pr->m1, stablePC2.data(), (int)(stablePC2.size() / 3), 0.75f);
        if (stats) stats->P75 = Dist75;
        if (currDT > Dist75) currDT = Dist75;
        else st->toStage2 = 1;
        if (currDT <= LoDet_min) currDT = LoDet_min;
        st->BBchange_2 = st->BBchange_1;
        st->BBchange_1 = maxBBchange;
    }
    if (st->toStage2 && !st->toStage3) {
        float upperBound = 0.8;
        float lowerBound = 0.5;
        float alpha = std::abs(st->BBchange_1 / st->BBchange_2);
        if (std::isnan(alpha) || std::isinf(alpha)) currDT = currDT * upperBound;
        else if (alpha < lowerBound) currDT = currDT * lowerBound;
        else if (alpha > upperBound) currDT = currDT * upperBound;
        else currDT = currDT * alpha;
        if (currDT <= LoDet_min) currDT = LoDet_min;
        st->BBchange_2 = st->BBchange_1;
        st->BBchange_1 = maxBBchange;
    }

    /* (8) apply the transform to cloud2, CT2, BP2 and every patch, :942-954 */
    orc_transform(pr->cloud2, pr->m2, transMatICP);
    orc_transform(pr->ct2, pr->n2, transMatICP);
    orc_transform(pr->bp2, 6 * pr->n2, transMatICP);
    orc_transform(pr->patch_pts2, pr->patch_off2[pr->n2], transMatICP);

    /* (9) VCM from the pre-update stable centroids, :957-961 */
    if (st->toStage3 && vcm36) {
        int sing = 0;
        orc_vcm(pr->ct1, pr->nrm1, pr->n1, stableCT2.data(), nStable, vcm36, &sing);
        if (vcm_written) *vcm_written = 1;
    }
    std::memcpy(T16, transMatICP, sizeof(transMatICP));
    return 0;
}

/* Piecewise_ICP (src/Registration.cpp:618-700), after patch generation */
int orc_piecewise_icp(orc_pair* pr, int isManualDTinit, float DTinit, const orc_icp_params* icp,
                      int max_outer, float* DTseries, int* n_series, float* T16, double* vcm36,
                      orc_iter_stats* stats_per_iter) {
    orc_state st;
    st.toStage2 = 0; st.toStage3 = 0;                                  /* :623-624 */
    if (!isManualDTinit) {
        double Dist75 = orc_percentile_nn(pr->cloud1, pr->m1, pr->cloud2, pr->m2, 0.75);
        DTinit = Dist75 * 3.0;                                         /* :627-630 */
    }
    st.currDT = DTinit;
    st.BBchange_1 = 0.0f; st.BBchange_2 = 0.0f;
    float transMat[16]; set_identity(transMat);
    int ns = 0, count = 0;
    DTseries[ns++] = st.currDT;
    while (!st.toStage3 && count < max_outer) {
        float cur[16]; int vw = 0;
        int rc = orc_single_iteration(pr, &st, icp, cur, vcm36, &vw, nullptr,
                                      stats_per_iter ? stats_per_iter + count : nullptr);
        if (rc != 0) { *n_series = ns; std::memcpy(T16, transMat, sizeof(transMat)); return rc; }
        mat4_mul(cur, transMat, transMat);                             /* :687 */
        DTseries[ns++] = st.currDT;
        ++count;
    }
    *n_series = ns;
    std::memcpy(T16, transMat, sizeof(transMat));
    return count;
}

/* ------------------------------------------------------------------------------------------
 * F4: PCpreprocessing (src/CommonFunc.cpp:423-452) = pcl::VoxelGrid + pcl::StatisticalOutlierRemoval
 * [PCL-recalled, PCL 1.8.1 filters/voxel_grid.hpp applyFilter, statistical_outlier_removal.hpp
 * applyFilterIndices].
 * ------------------------------------------------------------------------------------------ */

/* VoxelGrid with a cubic leaf: inverse_leaf = 1/leaf (float); min_b/max_b = floor(min/max * inverse_leaf);
 * voxel index = ijk . (1, div_x, div_x*div_y) with ijk = floor(p * inverse_leaf) - min_b; points sorted by
 * voxel index; one output point per occupied voxel in ascending index order, the float centroid of its
 * points.  The order of the points INSIDE a voxel follows std::sort in PCL (unspecified); here it is the
 * input order (what a stable sort gives), which fixes the float summation order.  Returns the number of
 * output points (out has room for n). */
static int voxel_grid_impl(const float* xyz, int n, float leaf, float* out, int msvc_order);
int orc_voxel_grid(const float* xyz, int n, float leaf, float* out) { return voxel_grid_impl(xyz, n, leaf, out, 0); }
/* the same with the points of a voxel in the order the Microsoft STL's std::sort leaves them (msvc_sort.h): what the
 * reference's Windows build, which recorded results/4DPCReg, computed */
int orc_voxel_grid_msvc(const float* xyz, int n, float leaf, float* out) { return voxel_grid_impl(xyz, n, leaf, out, 1); }
static int voxel_grid_impl(const float* xyz, int n, float leaf, float* out, int msvc_order) {
    if (n <= 0) return 0;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) { mn[c] = std::min(mn[c], xyz[3 * (size_t)i + c]); mx[c] = std::max(mx[c], xyz[3 * (size_t)i + c]); }
    const float inv = 1.0f / leaf;
    long long minb[3], div[3];
    for (int c = 0; c < 3; ++c) {
        minb[c] = (long long)std::floor(mn[c] * inv);
        div[c] = (long long)std::floor(mx[c] * inv) - minb[c] + 1;
    }
    std::vector<std::pair<long long, int>> keyed((size_t)n);
    for (int i = 0; i < n; ++i) {
        const float* p = xyz + 3 * (size_t)i;
        const long long ix = (long long)std::floor(p[0] * inv) - minb[0];
        const long long iy = (long long)std::floor(p[1] * inv) - minb[1];
        const long long iz = (long long)std::floor(p[2] * inv) - minb[2];
        keyed[i] = {ix + iy * div[0] + iz * div[0] * div[1], i};
    }
    auto by_key = [](const std::pair<long long, int>& a, const std::pair<long long, int>& b) { return a.first < b.first; };
    if (msvc_order) msvc::sort(keyed.begin(), keyed.end(), by_key);
    else std::stable_sort(keyed.begin(), keyed.end(), by_key);
    int m = 0;
    size_t i = 0;
    while (i < keyed.size()) {
        size_t j = i;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        while (j < keyed.size() && keyed[j].first == keyed[i].first) {
            const float* p = xyz + 3 * (size_t)keyed[j].second;
            sx += p[0]; sy += p[1]; sz += p[2]; ++j;
        }
        const float cnt = (float)(j - i);
        out[3 * (size_t)m] = sx / cnt; out[3 * (size_t)m + 1] = sy / cnt; out[3 * (size_t)m + 2] = sz / cnt;
        ++m;
        i = j;
    }
    return m;
}

/* StatisticalOutlierRemoval, first pass: for every point the mean distance to its k nearest OTHER points
 * (nearestKSearch(point, k + 1), first hit = the point itself, skipped): dist_sum (double) of sqrt(d2) (float d2,
 * float sqrt) in ascending order, distances[i] = float(dist_sum / k).  n must exceed k. */
int orc_knn_mean_dist(const float* xyz, int n, int k, float* mean_dist) {
    if (k < 1 || k > 256 || n <= k) return -1;
    KdTree tree;
    tree.build(xyz, n);
    std::vector<float> best((size_t)k);
    for (int i = 0; i < n; ++i) {
        tree.query_k(xyz + 3 * (size_t)i, i, k, best.data());
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += std::sqrt(best[j]);
        mean_dist[i] = (float)(s / k);
    }
    return 0;
}

/* second pass [PCL 1.8.1 StatisticalOutlierRemoval::applyFilterIndices]: mean and standard deviation of the mean distances
 * (double sums over a vector<float>: `sq_sum += distances[i] * distances[i]` is a FLOAT product, rounded before it is
 * widened; the variance is not clamped), threshold = mean + mult * stddev, a point is an outlier iff distance > threshold;
 * the others are kept in input order (negative = false).  Returns the number kept. */
int orc_sor_select(const float* xyz, int n, const float* mean_dist, double std_mult, float* out, double* threshold) {
    double sum = 0, sq = 0;
    for (int i = 0; i < n; ++i) { const float d_sq = mean_dist[i] * mean_dist[i]; sum += mean_dist[i]; sq += d_sq; }
    const double mean = sum / n;
    const double var = (sq - sum * sum / n) / (n - 1);
    const double thr = mean + std_mult * std::sqrt(var);
    if (threshold) *threshold = thr;
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (!(mean_dist[i] > thr)) { std::memcpy(out + 3 * (size_t)m, xyz + 3 * (size_t)i, 12); ++m; }
    return m;
}

/* F4, segmentation front end (src/Segmentation.cpp:28-46): for every point its k nearest neighbours (itself included, ascending
 * distance in the codelibrary's double metric, ties by index) and the normal of
 * cl::geometry::point_cloud::PCAEstimateNormal over them (codelibrary/geometry/point_cloud/pca_estimate_normals.h:47-117,
 * unit weights): centroid and covariance sums in neighbour order, smallest eigenvalue in closed form
 * (trigonometric solution of the characteristic cubic), eigenvector from the cross product of two rows, normalised;
 * (0,0,1) when degenerate.  The orientation of the normal is not defined (the reference says so).
 * neighbors: n x k, normals: n x 3 doubles. */
int orc_knn_normals(const float* xyz, int n, int k, int* neighbors, double* normals) {
    if (k < 1 || k > 256 || n < k) return -1;
    KdTree tree;
    tree.build(xyz, n);
    std::vector<double> bd((size_t)k);
    std::vector<int> bi((size_t)k);
    for (int i = 0; i < n; ++i) {
        tree.query_kd(xyz + 3 * (size_t)i, k, bd.data(), bi.data());
        for (int j = 0; j < k; ++j) neighbors[(size_t)i * k + j] = bi[j];
        double cx = 0, cy = 0, cz = 0, sum = 0;                               /* Centroid3D, center_3d.h:82-108 */
        for (int j = 0; j < k; ++j) {
            const float* p = xyz + 3 * (size_t)bi[j];
            const double w = 1.0;
            cx += w * p[0]; cy += w * p[1]; cz += w * p[2]; sum += w;
        }
        sum = 1.0 / sum; cx *= sum; cy *= sum; cz *= sum;
        double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, wsum = 0;
        for (int j = 0; j < k; ++j) {
            const float* p = xyz + 3 * (size_t)bi[j];
            const double x = p[0] - cx, y = p[1] - cy, z = p[2] - cz, w = 1.0;
            a00 += w * x * x; a01 += w * x * y; a02 += w * x * z; a11 += w * y * y; a12 += w * y * z; a22 += w * z * z;
            wsum += w;
        }
        const double t = 1.0 / wsum;
        a00 *= t; a01 *= t; a02 *= t; a11 *= t; a12 *= t; a22 *= t;
        const double q = (a00 + a11 + a22) / 3.0;
        double pq = (a00 - q) * (a00 - q) + (a11 - q) * (a11 - q) + (a22 - q) * (a22 - q) + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12);
        pq = std::sqrt(pq / 6.0);
        const double mpq = std::pow(1.0 / pq, 3.0);
        const double det_b = mpq * ((a00 - q) * ((a11 - q) * (a22 - q) - a12 * a12) - a01 * (a01 * (a22 - q) - a12 * a02) +
                                    a02 * (a01 * a12 - (a11 - q) * a02));
        const double r = 0.5 * det_b;
        double phi = 0.0;
        if (r <= -1.0) phi = M_PI / 3.0;
        else if (r >= 1.0) phi = 0.0;
        else phi = std::acos(r) / 3.0;
        const double eig = q + 2.0 * pq * std::cos(phi + M_PI * (2.0 / 3.0));
        double nx = a01 * a12 - a02 * (a11 - eig);
        double ny = a01 * a02 - a12 * (a00 - eig);
        double nz = (a00 - eig) * (a11 - eig) - a01 * a01;
        const double norm = std::sqrt(nx * nx + ny * ny + nz * nz);
        if (norm == 0.0) { nx = 0.0; ny = 0.0; nz = 1.0; }
        else { const double s = 1.0 / norm; nx *= s; ny *= s; nz *= s; }
        normals[3 * (size_t)i] = nx; normals[3 * (size_t)i + 1] = ny; normals[3 * (size_t)i + 2] = nz;
    }
    return 0;
}

}  // extern "C"
