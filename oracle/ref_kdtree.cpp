// oracle/ref_kdtree.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The one piece of the reference that compiles in this container is its header-only, STL-only
// codelibrary.  Its KD-tree (codelibrary/util/tree/kd_tree.h: build :378-392,
// divide :481-511, middle split :536-623, search :401-473) has the structure of the FLANN
// KDTreeSingleIndex PCL calls on the hot path (FLANN itself is absent), with two differences:
// leaf size 10 instead of 15 and the metric accumulated in double
// (codelibrary/util/metric/squared_euclidean.h:33-36) instead of float (flann::L2_Simple<float>).
// This file compiles that header WHERE IT LIES under /root/reference (include path given by
// oracle/Makefile; no reference source is copied) into oracle/_ref/libref_kdtree.so, so that the
// oracle's exact-1-NN restatement can be pinned against reference-authored code: the two must
// agree on the index of every query whose two best float distances differ (the rounding-level
// ties are the only place the double metric can pick another point).
#include <cstdint>
#include <memory>      // codelibrary/base/array.h uses std::uninitialized_* without including it (MSVC pulls it in)
#include <cstring>
#include <numeric>     // std::iota, same reason (codelibrary/base/algorithm.h:53)

#include "codelibrary/geometry/kernel/point_3d.h"
#include "codelibrary/util/tree/kd_tree.h"

typedef cl::Point3D<float> Pt;

extern "C" {

// exact 1-NN of every query; d2 = the tree's own metric value (double) for the returned index
int ref_kdtree_nn(const float* tgt, int n1, const float* qry, int nq, int32_t* idx, double* d2) {
    if (n1 <= 0) return -1;
    cl::Array<Pt> pts(n1);
    for (int i = 0; i < n1; ++i) pts[i] = Pt(tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2]);
    cl::KDTree<Pt> tree(pts.begin(), pts.end());
    cl::metric::SquaredEuclidean m;
    for (int i = 0; i < nq; ++i) {
        const Pt q(qry[3 * i], qry[3 * i + 1], qry[3 * i + 2]);
        int j = -1;
        tree.FindNearestPoint(q, &j);
        idx[i] = j;
        d2[i] = m(q, tree.points()[j]);
    }
    return 0;
}

// k nearest neighbours (ascending distance) of every query: idx is nq x k
int ref_kdtree_knn(const float* tgt, int n1, const float* qry, int nq, int k, int32_t* idx) {
    if (n1 <= 0 || k <= 0 || k > n1) return -1;
    cl::Array<Pt> pts(n1);
    for (int i = 0; i < n1; ++i) pts[i] = Pt(tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2]);
    cl::KDTree<Pt> tree(pts.begin(), pts.end());
    cl::Array<int> nb;
    for (int i = 0; i < nq; ++i) {
        const Pt q(qry[3 * i], qry[3 * i + 1], qry[3 * i + 2]);
        tree.FindKNearestNeighbors(q, k, &nb);
        for (int t = 0; t < k; ++t) idx[(size_t)i * k + t] = nb[t];
    }
    return 0;
}

}  // extern "C"
