// patch.cu -- batched per-patch statistics on the device.
//
// One thread per planar patch: centroid + six boundary points (calPatchCTandBP, reference
// src/Segmentation.cpp:260-303), patch normal (calPatchNormal, src/CommonFunc.cpp:284-333), plane
// sigma and its centroid form (calPatchSTD :336-354, calBPandCTSTD src/Segmentation.cpp:306-321).
// The reference recomputes the normal of every patch 7*N2 + N1 + N2 times per outer iteration
// (SURVEY.md 8a A9); here all patch constants of a cloud come from one launch.  The arithmetic is
// patch_algebra.cuh: sequential sums in point order, one thread per patch, so the float results do
// not depend on the launch geometry.
#include "common.cuh"
#include "patch_algebra.cuh"

namespace pwicp {

__global__ void __launch_bounds__(128)
patch_stats_kernel(const float* __restrict__ xyz, const int* __restrict__ off, int np, float* __restrict__ ct,
                   float* __restrict__ bp, float* __restrict__ nrm, unsigned char* __restrict__ ok,
                   float* __restrict__ bpstd, float* __restrict__ ctstd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const int s = off[i], n = off[i + 1] - s;
    const float* pts = xyz + 3 * (size_t)s;
    if (ct || bp) {
        float c3[3], b18[18];
        pa_patch_ct_bp(pts, n, c3, b18);
        if (ct) for (int k = 0; k < 3; ++k) ct[3 * (size_t)i + k] = c3[k];
        if (bp) for (int k = 0; k < 18; ++k) bp[18 * (size_t)i + k] = b18[k];
    }
    if (nrm || ok) {
        float n3[3];
        const int good = pa_patch_normal(pts, n, n3);
        if (nrm) for (int k = 0; k < 3; ++k) nrm[3 * (size_t)i + k] = n3[k];
        if (ok) ok[i] = (unsigned char)good;
    }
    if (bpstd || ctstd) {
        const float sd = pa_patch_std(pts, n);
        if (bpstd) bpstd[i] = sd;
        if (ctstd) ctstd[i] = sd / (float)n;
    }
}

int patch_stats_dev(Ctx* ctx, const float* xyz_dev, const int* off_dev, int np, float* ct, float* bp, float* nrm,
                    unsigned char* ok, float* bpstd, float* ctstd) {
    if (np < 1) return PWICP_OK;
    patch_stats_kernel<<<(np + 127) / 128, 128, 0, ctx->stream>>>(xyz_dev, off_dev, np, ct, bp, nrm, ok, bpstd, ctstd);
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

}  // namespace pwicp
