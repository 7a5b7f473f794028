// oracle/ref_supervoxel.cpp -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.
//
// The reference's patch generator in front of the hot path (src/Segmentation.cpp:11-72) is built
// entirely on its header-only, STL-only codelibrary: k-NN (k = kNN = 45, include/CommonFunc.h:41)
// on cl::KDTree, cl::geometry::point_cloud::PCAEstimateNormal, and Lin's boundary-preserving
// cl::geometry::point_cloud::SupervoxelSegmentation with the VCCS metric.  Those headers compile
// here, so this file compiles them WHERE THEY LIE under /root/reference (include path given by
// oracle/Makefile; no reference source is copied) into oracle/_ref/libref_supervoxel.so.  It
// returns the supervoxel label of every point, i.e. exactly the `lin_labels` the reference's
// PatchGenerationAndRefinement groups points by (src/Segmentation.cpp:95-100).  Segmentation is
// out of scope for the product; the file exists so that the hot path can be run on
// the reference's shipped scans behind the reference's OWN segmentation and compared with the
// results the reference recorded (results/4DPCReg), through the product's segmenter plug-in
// (pwicp_host_set_segmenter, include/pwicp_host.h).
//
// The two 5-line types below live in include/Segmentation.h (:24-28 PointWithNormal, :362-375
// VCCSMetric), which cannot be included because it pulls in PCL through CommonFunc.h.
#include <cstdint>
#include <cstring>
#include <memory>       // codelibrary/base/array.h uses std::uninitialized_* without including it (MSVC pulls it in)
#include <numeric>      // std::iota, same reason (codelibrary/base/algorithm.h:53)

#include "codelibrary/geometry/kernel/point_3d.h"
#include "codelibrary/geometry/util/distance_3d.h"
#include "codelibrary/geometry/point_cloud/pca_estimate_normals.h"
#include "codelibrary/geometry/point_cloud/supervoxel_segmentation.h"
#include "codelibrary/util/tree/kd_tree.h"

namespace {

struct PointWithNormal : cl::RPoint3D {       // include/Segmentation.h:24-28
    PointWithNormal() {}
    cl::RVector3D normal;
};

class VCCSMetric {                            // include/Segmentation.h:362-375
public:
    explicit VCCSMetric(double resolution) : resolution_(resolution) {}
    double operator()(const PointWithNormal& p1, const PointWithNormal& p2) const {
        return 1.0 - std::fabs(p1.normal * p2.normal) + cl::geometry::Distance(p1, p2) / resolution_ * 0.4;
    }
private:
    double resolution_;
};

}  // namespace

// Statement order of src/Segmentation.cpp:17-66.  Returns the number of supervoxels (labels are in
// [0, return)), or -1 when the cloud has no more than knn points (the reference asserts).
extern "C" int ref_supervoxel_labels(const float* xyz, int n, float sv_resolution, int knn, int32_t* labels) {
    if (n <= knn || knn <= 0) return -1;
    cl::Array<cl::RPoint3D> points;
    for (int i = 0; i < n; ++i) points.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    const int numPoints = points.size();

    cl::KDTree<cl::RPoint3D> kdtree;
    kdtree.SwapPoints(&points);
    cl::Array<cl::RVector3D> normals(numPoints);
    cl::Array<cl::Array<int>> neighbors(numPoints);
    cl::Array<cl::RPoint3D> neighbor_points(knn);
    for (int i = 0; i < numPoints; ++i) {
        kdtree.FindKNearestNeighbors(kdtree.points()[i], knn, &neighbors[i]);
        for (int k = 0; k < knn; ++k) neighbor_points[k] = kdtree.points()[neighbors[i][k]];
        cl::geometry::point_cloud::PCAEstimateNormal(neighbor_points.begin(), neighbor_points.end(), &normals[i]);
    }
    kdtree.SwapPoints(&points);

    VCCSMetric metric(sv_resolution);
    cl::Array<int> lin_supervoxels, lin_labels;
    cl::Array<PointWithNormal> oriented_points(numPoints);
    for (int i = 0; i < numPoints; ++i) {
        oriented_points[i].x = points[i].x;
        oriented_points[i].y = points[i].y;
        oriented_points[i].z = points[i].z;
        oriented_points[i].normal = normals[i];
    }
    cl::geometry::point_cloud::SupervoxelSegmentation(oriented_points, neighbors, sv_resolution, metric,
                                                      &lin_supervoxels, &lin_labels);
    for (int i = 0; i < numPoints; ++i) labels[i] = lin_labels[i];
    return lin_supervoxels.size();
}

// The first half of src/Segmentation.cpp:17-66 alone (statements :28-46): the k nearest neighbours of every point (the point
// itself first) and the PCA normal over them, as the reference computes them.  neighbors: n x knn indices; normals: n x 3 doubles.
extern "C" int ref_knn_normals(const float* xyz, int n, int knn, int32_t* neighbors_out, double* normals_out) {
    if (n <= knn || knn <= 0) return -1;
    cl::Array<cl::RPoint3D> points;
    for (int i = 0; i < n; ++i) points.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    cl::KDTree<cl::RPoint3D> kdtree;
    kdtree.SwapPoints(&points);
    cl::Array<int> nb;
    cl::Array<cl::RPoint3D> neighbor_points(knn);
    for (int i = 0; i < n; ++i) {
        kdtree.FindKNearestNeighbors(kdtree.points()[i], knn, &nb);
        for (int k = 0; k < knn; ++k) { neighbor_points[k] = kdtree.points()[nb[k]]; neighbors_out[(size_t)i * knn + k] = nb[k]; }
        cl::RVector3D nrm;
        cl::geometry::point_cloud::PCAEstimateNormal(neighbor_points.begin(), neighbor_points.end(), &nrm);
        normals_out[3 * (size_t)i] = nrm.x; normals_out[3 * (size_t)i + 1] = nrm.y; normals_out[3 * (size_t)i + 2] = nrm.z;
    }
    return 0;
}

