// patch_algebra.cuh -- per-patch statistics shared by the device kernel (patch.cu) and host code.
//
// Restates, for one planar patch given as n packed xyz points:
//   * calPatchCTandBP (reference src/Segmentation.cpp:260-303): pcl::compute3DCentroid (float sums in
//     point order) and the six axis-extremal points Xmax,Xmin,Ymax,Ymin,Zmax,Zmin (strict comparisons,
//     first extremal point wins);
//   * calPatchNormal (src/CommonFunc.cpp:284-333): pcl::computePointNormal = single-pass float mean +
//     covariance, pcl::eigen33 smallest eigenvector (closed-form roots + best cross product), accepted
//     if | |n| - 1 | < 1e-5, else the centred-covariance recomputation (:303-326);
//   * calPatchSTD (src/CommonFunc.cpp:336-354): plane through the centroid with the smallest-variance
//     direction, std of the point-to-plane distances with (n - 1); CTstd = std / n
//     (calBPandCTSTD, src/Segmentation.cpp:306-321).
// Compiled without FMA contraction (-fmad=false / -ffp-contract=off), so the float sums are the
// sequential sums of the reference on host and device alike; only the libm calls of eigen33
// (atan2f, cosf, sinf) may differ in the last ulp between CUDA and the host C library.
#pragma once
#include <cfloat>
#include <cmath>

#ifndef PW_HD
#ifdef __CUDACC__
#define PW_HD __host__ __device__ __forceinline__
#else
#define PW_HD inline
#endif
#endif

namespace pwicp {

PW_HD void pa_roots2(float b, float c, float* roots) {
    roots[0] = 0.0f;
    float d = (float)(b * b - 4.0 * c);
    if (d < 0.0) d = 0.0f;
    const float sd = sqrtf(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

PW_HD void pa_swap(float& a, float& b) { const float t = a; a = b; b = t; }

PW_HD void pa_roots3(const float m[3][3], float* roots) {
    const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2]
                   - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
    const float c2 = m[0][0] + m[1][1] + m[2][2];
    if (fabsf(c0) < FLT_EPSILON) { pa_roots2(c2, c1, roots); return; }
    const float inv3 = (float)(1.0 / 3.0), sqrt3 = sqrtf(3.0f);
    const float c2_3 = c2 * inv3;
    float a_3 = (c1 - c2 * c2_3) * inv3;
    if (a_3 > 0.0f) a_3 = 0.0f;
    const float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
    float q = half_b * half_b + a_3 * a_3 * a_3;
    if (q > 0.0f) q = 0.0f;
    const float rho = sqrtf(-a_3);
    const float theta = atan2f(sqrtf(-q), half_b) * inv3;
    const float ct = cosf(theta), st = sinf(theta);
    roots[0] = c2_3 + 2.0f * rho * ct;
    roots[1] = c2_3 - rho * (ct + sqrt3 * st);
    roots[2] = c2_3 - rho * (ct - sqrt3 * st);
    if (roots[0] >= roots[1]) pa_swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) { pa_swap(roots[1], roots[2]); if (roots[0] >= roots[1]) pa_swap(roots[0], roots[1]); }
    if (roots[0] <= 0) pa_roots2(c2, c1, roots);
}

PW_HD void pa_cross(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// pcl::eigen33(mat, eigenvalue, eigenvector): eigenvector of the smallest eigenvalue
PW_HD void pa_smallest_eigenvector(const float C[3][3], float* v) {
    float scale = 0.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = fmaxf(scale, fabsf(C[i][j]));
    if (scale <= FLT_MIN) scale = 1.0f;
    float m[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = C[i][j] / scale;
    float roots[3];
    pa_roots3(m, roots);
    for (int i = 0; i < 3; ++i) m[i][i] -= roots[0];
    float v1[3], v2[3], v3[3];
    pa_cross(m[0], m[1], v1); pa_cross(m[0], m[2], v2); pa_cross(m[1], m[2], v3);
    const float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    const float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    const float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
    const float* b; float l;
    if (l1 >= l2 && l1 >= l3) { b = v1; l = l1; } else if (l2 >= l1 && l2 >= l3) { b = v2; l = l2; } else { b = v3; l = l3; }
    const float s = sqrtf(l);
    v[0] = b[0] / s; v[1] = b[1] / s; v[2] = b[2] / s;
}

// symmetric 3x3 eigen decomposition (cyclic Jacobi, double): column `vmin` of the smallest eigenvalue
PW_HD void pa_jacobi3_smallest(double A[3][3], double vmin[3]) {
    double V[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
                for (int k = 0; k < 3; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
                for (int k = 0; k < 3; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
            }
    }
    int j = 0;                                          // first smallest diagonal entry (stable order)
    if (A[1][1] < A[j][j]) j = 1;
    if (A[2][2] < A[j][j]) j = 2;
    vmin[0] = V[0][j]; vmin[1] = V[1][j]; vmin[2] = V[2][j];
}

// centred second moments in double (sum, not divided): mean[3], M[3][3]
PW_HD void pa_centred_moments(const float* pts, int n, double mean[3], double M[3][3]) {
    mean[0] = mean[1] = mean[2] = 0.0;
    for (int i = 0; i < n; ++i) { mean[0] += pts[3 * i]; mean[1] += pts[3 * i + 1]; mean[2] += pts[3 * i + 2]; }
    for (int c = 0; c < 3; ++c) mean[c] /= (double)n;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] = 0.0;
    for (int i = 0; i < n; ++i) {
        const double d[3] = {pts[3 * i] - mean[0], pts[3 * i + 1] - mean[1], pts[3 * i + 2] - mean[2]};
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] += d[r] * d[c];
    }
}

// calPatchNormal: returns 1 on success; (0,0,1) and 0 for patches of 4 points or fewer
PW_HD int pa_patch_normal(const float* pts, int n, float* n3) {
    if (!(n > 4)) { n3[0] = 0.f; n3[1] = 0.f; n3[2] = 1.f; return 0; }
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        acc[0] += x * x; acc[1] += x * y; acc[2] += x * z;
        acc[3] += y * y; acc[4] += y * z; acc[5] += z * z;
        acc[6] += x; acc[7] += y; acc[8] += z;
    }
    for (int k = 0; k < 9; ++k) acc[k] /= (float)n;
    float C[3][3];
    C[0][0] = acc[0] - acc[6] * acc[6]; C[0][1] = acc[1] - acc[6] * acc[7]; C[0][2] = acc[2] - acc[6] * acc[8];
    C[1][1] = acc[3] - acc[7] * acc[7]; C[1][2] = acc[4] - acc[7] * acc[8]; C[2][2] = acc[5] - acc[8] * acc[8];
    C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
    float v[3];
    pa_smallest_eigenvector(C, v);
    const float nLen = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (fabs(nLen - 1.0) < 1e-5) { n3[0] = v[0]; n3[1] = v[1]; n3[2] = v[2]; return 1; }
    double mean[3], M[3][3], w[3];
    pa_centred_moments(pts, n, mean, M);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r][c] /= (double)n;
    pa_jacobi3_smallest(M, w);
    n3[0] = (float)w[0]; n3[1] = (float)w[1]; n3[2] = (float)w[2];
    const float nLen2 = sqrtf(n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]);
    return (fabs(nLen2 - 1.0) < 1e-5) ? 1 : 0;
}

// calPatchSTD
PW_HD float pa_patch_std(const float* pts, int n) {
    double mean[3], M[3][3], w[3];
    pa_centred_moments(pts, n, mean, M);
    pa_jacobi3_smallest(M, w);
    const float A = (float)w[0], B = (float)w[1], C = (float)w[2];
    const float D = -(A * (float)mean[0] + B * (float)mean[1] + C * (float)mean[2]);
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        // pcl::pointToPlaneDistance(p, double a, double b, double c, double d): float coefficients, double arithmetic
        const double dist = fabs((double)A * pts[3 * i] + (double)B * pts[3 * i + 1] + (double)C * pts[3 * i + 2] + (double)D) /
                            sqrt((double)A * A + (double)B * B + (double)C * C);
        s += dist * dist;
    }
    return (float)sqrt(s / (double)(n - 1));
}

// calPatchCTandBP: ct[3], bp[18] = Xmax, Xmin, Ymax, Ymin, Zmax, Zmin
PW_HD void pa_patch_ct_bp(const float* pts, int n, float* ct, float* bp) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    float e[6][3] = {{-FLT_MAX, 0, 0}, {FLT_MAX, 0, 0}, {0, -FLT_MAX, 0}, {0, FLT_MAX, 0}, {0, 0, -FLT_MAX}, {0, 0, FLT_MAX}};
    for (int i = 0; i < n; ++i) {
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        sx += x; sy += y; sz += z;
        if (x > e[0][0]) { e[0][0] = x; e[0][1] = y; e[0][2] = z; }
        if (x < e[1][0]) { e[1][0] = x; e[1][1] = y; e[1][2] = z; }
        if (y > e[2][1]) { e[2][0] = x; e[2][1] = y; e[2][2] = z; }
        if (y < e[3][1]) { e[3][0] = x; e[3][1] = y; e[3][2] = z; }
        if (z > e[4][2]) { e[4][0] = x; e[4][1] = y; e[4][2] = z; }
        if (z < e[5][2]) { e[5][0] = x; e[5][1] = y; e[5][2] = z; }
    }
    if (ct) { ct[0] = sx / (float)n; ct[1] = sy / (float)n; ct[2] = sz / (float)n; }
    if (bp) for (int k = 0; k < 6; ++k) for (int c = 0; c < 3; ++c) bp[3 * k + c] = e[k][c];
}

}  // namespace pwicp
