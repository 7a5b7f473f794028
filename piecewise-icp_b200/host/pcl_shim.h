// pcl_shim.h -- the few PCL / Eigen names the reference's public headers use, PCL-free.
//
// The reference's API (include/Registration.h, include/CommonFunc.h) is written in terms of
// pcl::PointCloud<pcl::PointXYZ>::Ptr, pcl::PointNormal, Eigen::Matrix4f and Eigen::MatrixXd.
// PCL, Eigen and Boost are not available in this build environment, so this header
// supplies minimal same-named types with the same memory layout (PointXYZ = 4 x f32, PointNormal =
// 12 x f32) and the handful of members the drivers touch.  A build that has the real libraries
// can define PWICP_USE_REAL_PCL and include them instead; every function of the mirror only uses
// the subset below.
#pragma once
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace Eigen {

struct Matrix4f {
    float m[16];                                     // row-major
    Matrix4f() { std::memset(m, 0, sizeof(m)); }
    static Matrix4f Identity() { Matrix4f I; I.m[0] = I.m[5] = I.m[10] = I.m[15] = 1.0f; return I; }
    float& operator()(int r, int c) { return m[r * 4 + c]; }
    float operator()(int r, int c) const { return m[r * 4 + c]; }
    const float* data() const { return m; }
    float* data() { return m; }
};
// product in Eigen's coefficient order (sum over k = 0..3, float, unfused)
Matrix4f operator*(const Matrix4f& a, const Matrix4f& b);

struct Vector3f {
    float v[3] = {0, 0, 0};
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
};
struct Vector4f {
    float v[4] = {0, 0, 0, 0};
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
};

struct MatrixXd {
    int r = 0, c = 0;
    std::vector<double> d;                           // row-major
    MatrixXd() {}
    MatrixXd(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
    static MatrixXd Zero(int rows, int cols) { return MatrixXd(rows, cols); }
    void resize(int rows, int cols) { r = rows; c = cols; d.assign((size_t)rows * cols, 0.0); }
    int rows() const { return r; }
    int cols() const { return c; }
    double& operator()(int i, int j) { return d[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
};

}  // namespace Eigen

namespace pcl {

struct alignas(16) PointXYZ {
    float x = 0, y = 0, z = 0, pad = 1.0f;
    PointXYZ() {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct alignas(16) PointNormal {
    float x = 0, y = 0, z = 0, pad0 = 1.0f;
    float normal_x = 0, normal_y = 0, normal_z = 0, pad1 = 0;
    float curvature = 0, pad2[3] = {0, 0, 0};
};

template <typename PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    unsigned width = 0, height = 1;
    bool is_dense = true;

    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = 0; height = 1; }
    void resize(size_t n) { points.resize(n); width = (unsigned)n; height = 1; }
    void push_back(const PointT& p) { points.push_back(p); width = (unsigned)points.size(); height = 1; }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
    PointT& at(size_t i) { return points.at(i); }
    PointCloud& operator+=(const PointCloud& o) {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (unsigned)points.size(); height = 1;
        return *this;
    }
    Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};

template <typename A, typename B>
inline void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out);
template <>
inline void copyPointCloud(const PointCloud<PointXYZ>& in, PointCloud<PointXYZ>& out) { out = in; }
template <>
inline void copyPointCloud(const PointCloud<PointNormal>& in, PointCloud<PointXYZ>& out) {
    out.resize(in.size());
    for (size_t i = 0; i < in.size(); ++i) { out.points[i].x = in.points[i].x; out.points[i].y = in.points[i].y; out.points[i].z = in.points[i].z; }
}

// x' = m00*x + m01*y + m02*z + m03 ..., float, left to right (pcl::transformPointCloud)
void transformPointCloud(const PointCloud<PointXYZ>& in, PointCloud<PointXYZ>& out, const Eigen::Matrix4f& T);
// float accumulation like pcl::compute3DCentroid
unsigned compute3DCentroid(const PointCloud<PointXYZ>& cloud, Eigen::Vector4f& centroid);

namespace io {
// ascii and binary PCD files with float x y z fields (extra fields are skipped)
int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud);
int savePCDFileBinary(const std::string& path, const PointCloud<PointXYZ>& cloud);
}  // namespace io

namespace console {
struct TicToc {
    std::chrono::steady_clock::time_point t0;
    void tic() { t0 = std::chrono::steady_clock::now(); }
    double toc() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
}  // namespace console

}  // namespace pcl
