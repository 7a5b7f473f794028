// small_algebra.cuh -- 6x6 / 4x4 helpers shared by host and device code of libpwicp.so.
//
// Follows the arithmetic PCL 1.8.1 performs behind the reference's call at
// src/Registration.cpp:1266 (TransformationEstimationPointToPlaneLLS): x = ATA.inverse() * ATb
// through a partial-pivot LU, then constructTransformationMatrix (double trig, float cast), and
// behind src/Registration.cpp:1328 (Q_xx = var_ATA.inverse()).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define PW_HD __host__ __device__ __forceinline__
#else
#define PW_HD inline
#endif

namespace pwicp {

// Inverse of a 6x6 (row-major) via LU with partial pivoting.  Returns the determinant.
PW_HD double inverse6(const double* A, double* Ainv) {
    double lu[36];
    int piv[6];
    for (int i = 0; i < 36; ++i) lu[i] = A[i];
    for (int i = 0; i < 6; ++i) piv[i] = i;
    double det = 1.0;
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double big = fabs(lu[k * 6 + k]);
        for (int r = k + 1; r < 6; ++r) {
            double v = fabs(lu[r * 6 + k]);
            if (v > big) { big = v; p = r; }
        }
        if (p != k) {
            for (int c = 0; c < 6; ++c) { double t = lu[k * 6 + c]; lu[k * 6 + c] = lu[p * 6 + c]; lu[p * 6 + c] = t; }
            int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
            det = -det;
        }
        double d = lu[k * 6 + k];
        det *= d;
        if (d == 0.0) continue;
        for (int r = k + 1; r < 6; ++r) {
            double f = lu[r * 6 + k] / d;
            lu[r * 6 + k] = f;
            for (int c = k + 1; c < 6; ++c) lu[r * 6 + c] -= f * lu[k * 6 + c];
        }
    }
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
    return det;
}

// constructTransformationMatrix(alpha, beta, gamma, tx, ty, tz): R = Rz(gamma) Ry(beta) Rx(alpha)
// (sines / cosines passed in so that a warp can evaluate them on three lanes)
PW_HD void construct_T_sc(double sa, double ca, double sb, double cb, double sg, double cg, const double* x, float* T) {
    for (int i = 0; i < 16; ++i) T[i] = 0.0f;
    T[0] = (float)(cg * cb);
    T[1] = (float)(-sg * ca + cg * sb * sa);
    T[2] = (float)(sg * sa + cg * sb * ca);
    T[4] = (float)(sg * cb);
    T[5] = (float)(cg * ca + sg * sb * sa);
    T[6] = (float)(-cg * sa + sg * sb * ca);
    T[8] = (float)(-sb);
    T[9] = (float)(cb * sa);
    T[10] = (float)(cb * ca);
    T[3] = (float)x[3];
    T[7] = (float)x[4];
    T[11] = (float)x[5];
    T[15] = 1.0f;
}

PW_HD void construct_T(const double* x, float* T) {
    construct_T_sc(sin(x[0]), cos(x[0]), sin(x[1]), cos(x[1]), sin(x[2]), cos(x[2]), x, T);
}

// 28 accumulated values -> ATA (mirrored), ATb, x, T
PW_HD void solve_from28(const double* s28, double* x, float* T) {
    double ATA[36], ATb[6], inv[36];
    int v = 0;
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) { ATA[r * 6 + c] = s28[v]; ATA[c * 6 + r] = s28[v]; ++v; }
    for (int r = 0; r < 6; ++r) ATb[r] = s28[v++];
    inverse6(ATA, inv);
    for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int c = 0; c < 6; ++c) s += inv[r * 6 + c] * ATb[c];
        x[r] = s;
    }
    construct_T(x, T);
}

// C = A * B for row-major 4x4 float (coefficient = products summed over k in order, no FMA:
// this file is compiled with -fmad=false / -ffp-contract=off)
PW_HD void mat4_mul(const float* A, const float* B, float* C) {
    float R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = A[i * 4 + 0] * B[0 * 4 + j];
            s += A[i * 4 + 1] * B[1 * 4 + j];
            s += A[i * 4 + 2] * B[2 * 4 + j];
            s += A[i * 4 + 3] * B[3 * 4 + j];
            R[i * 4 + j] = s;
        }
    for (int i = 0; i < 16; ++i) C[i] = R[i];
}

// x' = m00*x + m01*y + m02*z + m03 (pcl::transformPointCloud point formula, float, in order)
PW_HD void xform_point(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = T[0] * x + T[1] * y + T[2] * z + T[3];
    oy = T[4] * x + T[5] * y + T[6] * z + T[7];
    oz = T[8] * x + T[9] * y + T[10] * z + T[11];
}

}  // namespace pwicp
