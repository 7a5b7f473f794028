// Supervoxel.cpp -- a built-in segmenter for the plug-in point of Segmentation.cpp (include/pwicp_host.h:
// pwicp_segmenter_fn): boundary-preserving supervoxels after Lin et al. 2018 (ISPRS J. 143, 39-47), the segmentation the
// reference runs in front of the hot path (src/Segmentation.cpp:17-66 on top of its codelibrary:
// geometry/point_cloud/supervoxel_segmentation.h:65-265, pca_estimate_normals.h:47-117, grid_sample.h, VCCS metric
// include/Segmentation.h:362-375).  Written from that description, host C++ (the merge is a sequential union-find):
//
//   1. k nearest neighbours of every point (k = kNN = 45, the point itself first) in double precision;
//   2. normal of every point = eigenvector of the smallest eigenvalue of the neighbours' covariance (closed form);
//   3. fusion: every point starts as its own supervoxel; for lambda = median nearest-neighbour metric, doubling each
//      round, every representative absorbs the adjacent supervoxels j with  size(j) * D(i, j) < lambda  (region growing
//      over the adjacency, union-find) until as many supervoxels are left as the cloud has occupied cubes of side
//      svResolution;  D(p, q) = 1 - |n_p . n_q| + 0.4 |p - q| / svResolution;
//   4. boundary refinement: points adjacent to another supervoxel move to it while that lowers D(point, representative);
//   5. labels 0 .. numSV-1 in representative order.
//
// Statement order and arithmetic follow the reference so that the labels are the same (tests/test_host_cpu.py compares
// them with the reference's own code when oracle/_ref is built); exact distance ties between neighbours are the one
// place where the visiting order of the reference's KD-tree could differ.
//
// Not the default yet: PatchGenerationAndRefinement uses it when PWICP_SEGMENTER=supervoxel is set or when it is
// registered (pwicp_host_set_segmenter(&pwicp_host_builtin_supervoxels)); otherwise the cubic-cell stand-in stays.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <queue>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/pwicp_host.h"

namespace {

using std::vector;

struct P3 { double x, y, z; };

// ---- 1. k nearest neighbours on a dense cell grid (double metric: t = dx*dx + dy*dy + dz*dz, ties by index) ----------
struct CellGrid {
    double ox, oy, oz, h;
    int dx, dy, dz;
    vector<int> start, order;
    int cellOf(double v, double o, int d) const { int c = (int)std::floor((v - o) / h); return c < 0 ? 0 : (c >= d ? d - 1 : c); }
};

void buildGrid(const vector<P3>& p, CellGrid& g) {
    const int n = (int)p.size();
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (const P3& q : p) {
        mn[0] = std::min(mn[0], q.x); mn[1] = std::min(mn[1], q.y); mn[2] = std::min(mn[2], q.z);
        mx[0] = std::max(mx[0], q.x); mx[1] = std::max(mx[1], q.y); mx[2] = std::max(mx[2], q.z);
    }
    const double ex = mx[0] - mn[0] + 1e-9, ey = mx[1] - mn[1] + 1e-9, ez = mx[2] - mn[2] + 1e-9;
    // scans are sampled surfaces: size the cells for ~12 points per occupied cell, judged by the largest face of the box
    const double area = std::max(ex * ey, std::max(ex * ez, ey * ez));
    double h = std::sqrt(12.0 * area / std::max(n, 1));
    const double cap = 64e6;                                        // dense cell array stays below ~256 MB
    while ((ex / h + 1) * (ey / h + 1) * (ez / h + 1) > cap) h *= 1.26;
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.h = h;
    g.dx = (int)(ex / h) + 1; g.dy = (int)(ey / h) + 1; g.dz = (int)(ez / h) + 1;
    const size_t nc = (size_t)g.dx * g.dy * g.dz;
    g.start.assign(nc + 1, 0);
    vector<int> cell(n);
    for (int i = 0; i < n; ++i) {
        cell[i] = (g.cellOf(p[i].z, g.oz, g.dz) * g.dy + g.cellOf(p[i].y, g.oy, g.dy)) * g.dx + g.cellOf(p[i].x, g.ox, g.dx);
        ++g.start[cell[i] + 1];
    }
    for (size_t c = 0; c < nc; ++c) g.start[c + 1] += g.start[c];
    g.order.resize(n);
    vector<int> fill(g.start.begin(), g.start.end() - 1);
    for (int i = 0; i < n; ++i) g.order[fill[cell[i]]++] = i;      // ascending index inside a cell
}

// the k nearest of point i (itself included), ascending (distance, index)
void knnOf(const vector<P3>& p, const CellGrid& g, int i, int k, vector<double>& bd, vector<int>& bi) {
    const P3& q = p[i];
    const int cx = g.cellOf(q.x, g.ox, g.dx), cy = g.cellOf(q.y, g.oy, g.dy), cz = g.cellOf(q.z, g.oz, g.dz);
    int cnt = 0;
    auto offer = [&](int j) {
        const double tx = q.x - p[j].x, ty = q.y - p[j].y, tz = q.z - p[j].z;
        const double d = tx * tx + ty * ty + tz * tz;
        if (cnt == k && !(d < bd[k - 1] || (d == bd[k - 1] && j < bi[k - 1]))) return;
        int pos = (cnt < k) ? cnt++ : k - 1;
        while (pos > 0 && (bd[pos - 1] > d || (bd[pos - 1] == d && bi[pos - 1] > j))) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bd[pos] = d; bi[pos] = j;
    };
    const int rmax = std::max(g.dx, std::max(g.dy, g.dz));
    for (int r = 0; r <= rmax; ++r) {
        const int x0 = std::max(cx - r, 0), x1 = std::min(cx + r, g.dx - 1);
        const int y0 = std::max(cy - r, 0), y1 = std::min(cy + r, g.dy - 1);
        const int z0 = std::max(cz - r, 0), z1 = std::min(cz + r, g.dz - 1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const bool shellRow = (std::abs(z - cz) == r || std::abs(y - cy) == r);
                const size_t row = ((size_t)z * g.dy + y) * g.dx;
                if (shellRow) {
                    for (int t = g.start[row + x0]; t < g.start[row + x1 + 1]; ++t) offer(g.order[t]);
                } else {                                             // only the two end cells of the row are new
                    if (cx - r >= 0) for (int t = g.start[row + cx - r]; t < g.start[row + cx - r + 1]; ++t) offer(g.order[t]);
                    if (r > 0 && cx + r < g.dx) for (int t = g.start[row + cx + r]; t < g.start[row + cx + r + 1]; ++t) offer(g.order[t]);
                }
            }
        if (cnt == k) {
            // everything outside the scanned cube is at least this far away (faces beyond the grid do not count)
            double face = DBL_MAX;
            if (cx - r > 0) face = std::min(face, q.x - (g.ox + (cx - r) * g.h));
            if (cx + r < g.dx - 1) face = std::min(face, g.ox + (cx + r + 1) * g.h - q.x);
            if (cy - r > 0) face = std::min(face, q.y - (g.oy + (cy - r) * g.h));
            if (cy + r < g.dy - 1) face = std::min(face, g.oy + (cy + r + 1) * g.h - q.y);
            if (cz - r > 0) face = std::min(face, q.z - (g.oz + (cz - r) * g.h));
            if (cz + r < g.dz - 1) face = std::min(face, g.oz + (cz + r + 1) * g.h - q.z);
            if (face == DBL_MAX) break;
            if (face > 0 && bd[k - 1] < face * face * (1.0 - 1e-9)) break;
        }
    }
}

// ---- 2. normal over a neighbourhood (pca_estimate_normals.h:47-117 with unit weights) ---------------------------------
P3 pcaNormal(const vector<P3>& p, const int* nb, int k) {
    double cx = 0, cy = 0, cz = 0, sum = 0;
    for (int j = 0; j < k; ++j) { const double w = 1.0; cx += w * p[nb[j]].x; cy += w * p[nb[j]].y; cz += w * p[nb[j]].z; sum += w; }
    sum = 1.0 / sum; cx *= sum; cy *= sum; cz *= sum;
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, ws = 0;
    for (int j = 0; j < k; ++j) {
        const double x = p[nb[j]].x - cx, y = p[nb[j]].y - cy, z = p[nb[j]].z - cz, w = 1.0;
        a00 += w * x * x; a01 += w * x * y; a02 += w * x * z; a11 += w * y * y; a12 += w * y * z; a22 += w * z * z;
        ws += w;
    }
    const double t = 1.0 / ws;
    a00 *= t; a01 *= t; a02 *= t; a11 *= t; a12 *= t; a22 *= t;
    // least eigenvalue of the symmetric 3x3 in closed form (trigonometric solution of the characteristic cubic)
    const double q = (a00 + a11 + a22) / 3.0;
    double pq = (a00 - q) * (a00 - q) + (a11 - q) * (a11 - q) + (a22 - q) * (a22 - q) + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12);
    pq = std::sqrt(pq / 6.0);
    const double mpq = std::pow(1.0 / pq, 3.0);
    const double detB = mpq * ((a00 - q) * ((a11 - q) * (a22 - q) - a12 * a12) - a01 * (a01 * (a22 - q) - a12 * a02) +
                               a02 * (a01 * a12 - (a11 - q) * a02));
    const double r = 0.5 * detB;
    double phi = 0.0;
    if (r <= -1.0) phi = M_PI / 3.0;
    else if (r >= 1.0) phi = 0.0;
    else phi = std::acos(r) / 3.0;
    const double eig = q + 2.0 * pq * std::cos(phi + M_PI * (2.0 / 3.0));
    P3 nrm;
    nrm.x = a01 * a12 - a02 * (a11 - eig);
    nrm.y = a01 * a02 - a12 * (a00 - eig);
    nrm.z = (a00 - eig) * (a11 - eig) - a01 * a01;
    const double len = std::sqrt(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
    if (len == 0.0) { nrm.x = 0.0; nrm.y = 0.0; nrm.z = 1.0; }
    else { const double s = 1.0 / len; nrm.x *= s; nrm.y *= s; nrm.z *= s; }
    return nrm;
}

// ---- union-find with path halving; link(i, j) hangs root i under root j ------------------------------------------------
struct Forest {
    mutable vector<int> up;
    explicit Forest(int n) : up(n) { for (int i = 0; i < n; ++i) up[i] = i; }
    int find(int i) const { while (i != up[i]) { up[i] = up[up[i]]; i = up[i]; } return i; }
    void link(int i, int j) { up[i] = j; }
};

}  // namespace

extern "C" int pwicp_host_builtin_supervoxels(const float* xyz, int n, float svResolution, int knn, int* labels) {
    if (!xyz || !labels || n <= knn || knn < 1 || !(svResolution > 0.f)) return -1;
    vector<P3> p(n);
    for (int i = 0; i < n; ++i) { p[i].x = xyz[3 * i]; p[i].y = xyz[3 * i + 1]; p[i].z = xyz[3 * i + 2]; }
    const double resolution = svResolution;

    // 1 + 2: neighbours and normals
    CellGrid g;
    buildGrid(p, g);
    vector<int> nbr((size_t)n * knn);
    vector<P3> nrm(n);
    {
        // every point is independent: host threads (the fusion below is sequential by nature)
        const int nt = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        auto work = [&](int t) {
            vector<double> bd(knn); vector<int> bi(knn);
            const int lo = (int)((long long)n * t / nt), hi = (int)((long long)n * (t + 1) / nt);     // contiguous: scan order is local
            for (int i = lo; i < hi; ++i) {
                knnOf(p, g, i, knn, bd, bi);
                std::copy(bi.begin(), bi.end(), nbr.begin() + (size_t)i * knn);
                nrm[i] = pcaNormal(p, &nbr[(size_t)i * knn], knn);
            }
        };
        vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
    }
    auto metric = [&](int a, int b) {
        const double dot = nrm[a].x * nrm[b].x + nrm[a].y * nrm[b].y + nrm[a].z * nrm[b].z;
        const double t1 = p[a].x - p[b].x, t2 = p[a].y - p[b].y, t3 = p[a].z - p[b].z;
        return 1.0 - std::fabs(dot) + std::sqrt(t1 * t1 + t2 * t2 + t3 * t3) / resolution * 0.4;
    };

    // target number of supervoxels: occupied cubes of side `resolution` over the bounding box (grid_sample.h)
    int nTarget;
    {
        double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (const P3& q : p) {
            mn[0] = std::min(mn[0], q.x); mn[1] = std::min(mn[1], q.y); mn[2] = std::min(mn[2], q.z);
            mx[0] = std::max(mx[0], q.x); mx[1] = std::max(mx[1], q.y); mx[2] = std::max(mx[2], q.z);
        }
        const long long s1 = (long long)((mx[0] - mn[0]) / resolution) + 1, s2 = (long long)((mx[1] - mn[1]) / resolution) + 1,
                        s3 = (long long)((mx[2] - mn[2]) / resolution) + 1;
        std::unordered_set<long long> seen;
        seen.reserve((size_t)n / 8 + 16);
        for (const P3& q : p) {
            long long x = (long long)((q.x - mn[0]) / resolution), y = (long long)((q.y - mn[1]) / resolution), z = (long long)((q.z - mn[2]) / resolution);
            x = std::min(std::max(x, 0LL), s1 - 1); y = std::min(std::max(y, 0LL), s2 - 1); z = std::min(std::max(z, 0LL), s3 - 1);
            seen.insert((z * s2 + y) * s1 + x);
        }
        nTarget = (int)seen.size();
    }

    // 3. fusion
    Forest set(n);
    vector<int> reps(n);
    for (int i = 0; i < n; ++i) reps[i] = i;
    vector<int> sizes(n, 1), queue(n);
    vector<vector<int>> adj(n);
    for (int i = 0; i < n; ++i) adj[i].assign(nbr.begin() + (size_t)i * knn, nbr.begin() + (size_t)(i + 1) * knn);
    int count = n;
    vector<char> visited(n, 0);
    vector<double> dis(n, DBL_MAX);
    for (int i = 0; i < n; ++i)
        for (int j : adj[i])
            if (i != j) dis[i] = std::min(dis[i], metric(i, j));
    double lambda;
    {
        vector<double> v(dis);
        std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
        lambda = std::max(DBL_EPSILON, v[v.size() / 2]);
    }
    for (;; lambda *= 2.0) {
        if (reps.size() <= 1) break;
        for (int i : reps) {
            if (adj[i].empty()) continue;
            visited[i] = 1;
            int front = 0, back = 1;
            queue[front++] = i;
            for (int j : adj[i]) {
                j = set.find(j);
                if (!visited[j]) { visited[j] = 1; queue[back++] = j; }
            }
            vector<int> keep;
            while (front < back) {
                const int j = queue[front++];
                const double loss = sizes[j] * metric(i, j);
                if (lambda - loss > 0.0) {
                    set.link(j, i);
                    sizes[i] += sizes[j];
                    for (int k : adj[j]) {
                        k = set.find(k);
                        if (!visited[k]) { visited[k] = 1; queue[back++] = k; }
                    }
                    adj[j].clear();
                    if (--count == nTarget) break;
                } else {
                    keep.push_back(j);
                }
            }
            adj[i].swap(keep);
            for (int j = 0; j < back; ++j) visited[queue[j]] = 0;
            if (count == nTarget) break;
        }
        count = 0;
        for (int i : reps)
            if (set.find(i) == i) reps[count++] = i;
        reps.resize(count);
        if (count == nTarget) break;
    }
    for (int i = 0; i < n; ++i) labels[i] = set.find(i);

    // 4. boundary refinement over the original neighbour lists
    for (int i = 0; i < n; ++i) dis[i] = metric(i, labels[i]);
    std::queue<int> q;
    vector<char> inq(n, 0);
    for (int i = 0; i < n; ++i)
        for (int t = 0; t < knn; ++t) {
            const int j = nbr[(size_t)i * knn + t];
            if (labels[i] != labels[j]) {
                if (!inq[i]) { q.push(i); inq[i] = 1; }
                if (!inq[j]) { q.push(j); inq[j] = 1; }
            }
        }
    while (!q.empty()) {
        const int i = q.front();
        q.pop();
        inq[i] = 0;
        bool change = false;
        for (int t = 0; t < knn; ++t) {
            const int j = nbr[(size_t)i * knn + t];
            const int a = labels[i], b = labels[j];
            if (a == b) continue;
            const double d = metric(i, b);
            if (d < dis[i]) { labels[i] = b; dis[i] = d; change = true; }
        }
        if (change)
            for (int t = 0; t < knn; ++t) {
                const int j = nbr[(size_t)i * knn + t];
                if (labels[i] != labels[j] && !inq[j]) { q.push(j); inq[j] = 1; }
            }
    }

    // 5. labels in representative order
    vector<int> map(n, -1);
    for (size_t s = 0; s < reps.size(); ++s) map[reps[s]] = (int)s;
    for (int i = 0; i < n; ++i) labels[i] = map[labels[i]];
    return (int)reps.size();
}
