// capi.cu -- the extern "C" boundary of libpwicp.so (declared in include/pwicp.h).
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>

#include "common.cuh"
#include "nn_search.cuh"
#include "small_algebra.cuh"

namespace pwicp {

static std::string g_last_error;
static std::mutex g_err_mu;

void set_error(Ctx* c, const std::string& msg) {
    if (c) c->err = msg;
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_error = msg;
}

// nx,ny,nz,ctstd and ok flags gathered into level-0 order
__global__ void gather_aux_kernel(const float* __restrict__ nrm, const float* __restrict__ ctstd,
                                  const unsigned char* __restrict__ ok, const uint32_t* __restrict__ perm,
                                  int n, float4* aux, unsigned char* ok_sorted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o = perm[i];
    aux[i] = make_float4(nrm[3 * (size_t)o], nrm[3 * (size_t)o + 1], nrm[3 * (size_t)o + 2], ctstd ? ctstd[o] : 0.f);
    ok_sorted[i] = ok ? ok[o] : (unsigned char)1;
}

__global__ void expand_f4_kernel(const float* __restrict__ xyz, int n, float4* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.f);
}

__global__ void pack_f4_kernel(const float4* __restrict__ in, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 v = in[i]; out[3 * (size_t)i] = v.x; out[3 * (size_t)i + 1] = v.y; out[3 * (size_t)i + 2] = v.z; }
}

__global__ void patch_id_kernel(const int* __restrict__ off, int n2, int m, int* pid) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    int lo = 0, hi = n2;               // largest i with off[i] <= j
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= j) lo = mid; else hi = mid; }
    pid[j] = lo;
}

__global__ void count_below_kernel(const float* __restrict__ d2, int n, float thr, unsigned long long* cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool below = (i < n) && (sqrtf(d2[i]) < thr);
    unsigned m = __ballot_sync(0xffffffffu, below);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(cnt, (unsigned long long)__popc(m));
}

__global__ void fill_kernel(uint4* p, size_t n, unsigned v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_uint4(v, v, v, v);
}

static int upload_checked(Ctx* ctx, DevBuf& buf, const float* host, size_t n_floats, const char* what) {
    if (!host && n_floats) { set_error(ctx, std::string(what) + ": null pointer"); return PWICP_ERR_ARG; }
    PW_TRY(upload_packed(ctx, buf, host, n_floats));
    bool ok = true;
    PW_TRY(check_finite_dev(ctx, buf.as<float>(), n_floats, &ok));
    if (!ok) { set_error(ctx, std::string(what) + ": non-finite value in input"); return PWICP_ERR_NONFINITE; }
    return PWICP_OK;
}

// The normals of a host-buffer call have landed (or are waited for here, on the stream): finite check, level-0 order.
int finish_deferred_aux(Ctx* ctx) {
    if (!ctx->aux_deferred) return PWICP_OK;
    ctx->aux_deferred = false;
    const int n1 = ctx->aux_deferred_n1;
    PW_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[1], 0));
    PW_TRY(finite_accumulate_dev(ctx, ctx->tgt_nrm_raw.as<float>(), (size_t)3 * n1, ctx->aux_deferred_flag));
    gather_aux_kernel<<<(n1 + 255) / 256, 256, 0, ctx->stream>>>(ctx->tgt_nrm_raw.as<float>(), nullptr, nullptr,
                                                                 ctx->tgt.perm0, n1, ctx->tgt_aux.as<float4>(),
                                                                 ctx->tgt_ok.as<unsigned char>());
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

}  // namespace pwicp

using namespace pwicp;

extern "C" {

int pwicp_version(void) { return 100; }

int pwicp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* pwicp_last_error(const pwicp_ctx* c) {
    if (c) return reinterpret_cast<const Ctx*>(c)->err.c_str();
    return g_last_error.c_str();
}

int pwicp_ctx_create(int device, pwicp_ctx** out) {
    Ctx* ctx = nullptr;
    if (!out) return PWICP_ERR_ARG;
    *out = nullptr;
    int ndev = pwicp_device_count();
    if (ndev < 1) { set_error(nullptr, "no CUDA device available (libpwicp has no CPU fallback)"); return PWICP_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error(nullptr, "device index out of range"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(device));
    Ctx* c = new Ctx();
    c->device = device;
    ctx = c;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; set_error(nullptr, "cudaGetDeviceProperties failed"); return PWICP_ERR_CUDA; }
    if (prop.major < 10) {
        delete c;
        set_error(nullptr, "libpwicp is built for sm_100a (B200) only");
        return PWICP_ERR_CUDA;
    }
    c->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaEventCreate(&c->ev2) != cudaSuccess || cudaEventCreate(&c->ev3) != cudaSuccess ||
        cudaEventCreate(&c->ev4) != cudaSuccess || cudaEventCreate(&c->ev5) != cudaSuccess ||
        cudaEventCreate(&c->ev6) != cudaSuccess || cudaEventCreate(&c->ev7) != cudaSuccess ||
        cudaEventCreate(&c->ev_o0) != cudaSuccess || cudaEventCreate(&c->ev_o1) != cudaSuccess) {
        delete c; set_error(nullptr, "stream/event creation failed"); return PWICP_ERR_CUDA;
    }
    *out = reinterpret_cast<pwicp_ctx*>(c);
    return PWICP_OK;
}

void pwicp_ctx_destroy(pwicp_ctx* p) {
    if (!p) return;
    Ctx* c = reinterpret_cast<Ctx*>(p);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->tgt.release(); c->c1.release(); c->prep.release();
    DevBuf* bufs[] = {&c->tgt_aux, &c->tgt_ok, &c->ct2, &c->bp2, &c->bpstd2, &c->patch_xyz, &c->patch_id,
                      &c->patch_off, &c->cloud2, &c->icp_src, &c->icp_work, &c->icp_partials, &c->icp_out,
                      &c->icp_idx, &c->icp_sorted, &c->icp_perm, &c->icp_seed, &c->icp_match, &c->ct_seed, &c->bp_seed, &c->pp_seed, &c->ct_order, &c->tgt_xyz, &c->tgt_nrm_raw, &c->tgt_std_raw, &c->tgt_ok_raw, &c->keys, &c->vals, &c->keys2, &c->vals2, &c->cub_tmp, &c->scratch_a,
                      &c->scratch_b, &c->scratch_c, &c->scratch_d, &c->flags, &c->pos, &c->l2flush, &c->outer_state};
    for (DevBuf* b : bufs) b->release();
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev2); cudaEventDestroy(c->ev3); cudaEventDestroy(c->ev4); cudaEventDestroy(c->ev5); cudaEventDestroy(c->ev6); cudaEventDestroy(c->ev7);
    cudaEventDestroy(c->ev_o0); cudaEventDestroy(c->ev_o1);
    cudaStreamDestroy(c->stream);
    if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); for (auto& e : c->copy_ev) if (e) cudaEventDestroy(e); }
    delete c;
}

float pwicp_last_device_ms(const pwicp_ctx* p) { return p ? reinterpret_cast<const Ctx*>(p)->last_ms : 0.f; }
float pwicp_last_knn_kernel_ms(const pwicp_ctx* p) { return p ? reinterpret_cast<const Ctx*>(p)->prep_kernel_ms : 0.f; }
long long pwicp_launch_count(const pwicp_ctx* p) { return p ? reinterpret_cast<const Ctx*>(p)->launches : 0; }

int pwicp_sync(pwicp_ctx* p) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx) return PWICP_ERR_ARG;
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

int pwicp_flush_l2(pwicp_ctx* p) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx) return PWICP_ERR_ARG;
    PW_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)384 << 20;           // 3x the 126 MB L2
    PW_TRY(ctx->l2flush.reserve(ctx, bytes));
    fill_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->l2flush.as<uint4>(), bytes / 16, (unsigned)ctx->launches);
    ctx->launches++;
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

int pwicp_set_cells_per_point(pwicp_ctx* p, float cpp) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !(cpp > 0.01f) || cpp > 1024.f) return PWICP_ERR_ARG;
    ctx->cells_per_point = cpp;
    return PWICP_OK;
}

void pwicp_icp_default_params(pwicp_icp_params* p) {
    if (!p) return;
    p->max_iter = 100; p->tf_eps = 1e-8; p->fit_eps = 1e-6; p->force_iters = 0; p->rot_thr_default = 0;
}

// ---- uploads -------------------------------------------------------------------------------
static int reset_seeds(Ctx* ctx) {
    ctx->icp_seed_valid = false;
    ctx->ct_order_valid = false;
    if (ctx->n2 > 0) {
        PW_TRY(ctx->ct_seed.reserve(ctx, (size_t)ctx->n2 * 4));
        PW_TRY(ctx->bp_seed.reserve(ctx, (size_t)ctx->n2 * 24));
        PW_CUDA(cudaMemsetAsync(ctx->ct_seed.p, 0xff, (size_t)ctx->n2 * 4, ctx->stream));
        PW_CUDA(cudaMemsetAsync(ctx->bp_seed.p, 0xff, (size_t)ctx->n2 * 24, ctx->stream));
    }
    if (ctx->mp2 > 0) {
        PW_TRY(ctx->pp_seed.reserve(ctx, (size_t)ctx->mp2 * 4));
        PW_CUDA(cudaMemsetAsync(ctx->pp_seed.p, 0xff, (size_t)ctx->mp2 * 4, ctx->stream));
    }
    return PWICP_OK;
}

static int target_build_resident(Ctx* ctx, int n1) {
    PW_TRY(grid_build(ctx, ctx->tgt, ctx->tgt_xyz.as<float>(), n1));
    PW_TRY(reset_seeds(ctx));
    // normals / sigma / ok flags into level-0 order
    PW_TRY(ctx->tgt_aux.reserve(ctx, (size_t)n1 * sizeof(float4)));
    PW_TRY(ctx->tgt_ok.reserve(ctx, (size_t)n1));
    gather_aux_kernel<<<(n1 + 255) / 256, 256, 0, ctx->stream>>>(
        ctx->tgt_nrm_raw.as<float>(), ctx->tgt_has_std ? ctx->tgt_std_raw.as<float>() : nullptr,
        ctx->tgt_has_ok ? ctx->tgt_ok_raw.as<unsigned char>() : nullptr, ctx->tgt.perm0, n1,
        ctx->tgt_aux.as<float4>(), ctx->tgt_ok.as<unsigned char>());
    ctx->launches++;
    PW_CUDA(cudaGetLastError());
    return PWICP_OK;
}

int pwicp_target_upload(pwicp_ctx* p, const float* ct_xyz, const float* nrm, const unsigned char* nrm_ok,
                        const float* ct_std, int n1) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || n1 < 1 || !ct_xyz) { set_error(ctx, "target_upload: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    ctx->n1 = 0;
    PW_TRY(upload_checked(ctx, ctx->tgt_xyz, ct_xyz, (size_t)3 * n1, "target centroids"));
    if (nrm) {
        PW_TRY(upload_checked(ctx, ctx->tgt_nrm_raw, nrm, (size_t)3 * n1, "target normals"));
    } else {
        PW_TRY(ctx->tgt_nrm_raw.reserve(ctx, (size_t)3 * n1 * 4));
        PW_CUDA(cudaMemsetAsync(ctx->tgt_nrm_raw.p, 0, (size_t)3 * n1 * 4, ctx->stream));
    }
    ctx->tgt_has_std = ct_std != nullptr;
    if (ct_std) PW_TRY(upload_checked(ctx, ctx->tgt_std_raw, ct_std, (size_t)n1, "CTstd1"));
    ctx->tgt_has_ok = nrm_ok != nullptr;
    if (nrm_ok) {
        PW_TRY(ctx->tgt_ok_raw.reserve(ctx, (size_t)n1));
        PW_CUDA(cudaMemcpyAsync(ctx->tgt_ok_raw.p, nrm_ok, (size_t)n1, cudaMemcpyHostToDevice, ctx->stream));
    }
    PW_TRY(target_build_resident(ctx, n1));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->n1 = n1;
    return PWICP_OK;
}

int pwicp_target_rebuild(pwicp_ctx* p) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || ctx->n1 < 1) { set_error(ctx, "target_rebuild: no target uploaded"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    const int n1 = ctx->n1;
    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    PW_TRY(target_build_resident(ctx, n1));
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    PW_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return PWICP_OK;
}

int pwicp_source_upload(pwicp_ctx* p, const float* ct_xyz, const float* bp_xyz, const float* bp_std,
                        const int* patch_off, const float* patch_xyz, int n2) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || n2 < 1 || !ct_xyz) { set_error(ctx, "source_upload: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    ctx->n2 = 0;
    PW_TRY(upload_checked(ctx, ctx->scratch_a, ct_xyz, (size_t)3 * n2, "source centroids"));
    PW_TRY(ctx->ct2.reserve(ctx, (size_t)n2 * sizeof(float4)));
    expand_f4_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch_a.as<float>(), n2, ctx->ct2.as<float4>());
    ctx->launches++;
    PW_TRY(ctx->bp2.reserve(ctx, (size_t)6 * n2 * sizeof(float4)));
    if (bp_xyz) {
        PW_TRY(upload_checked(ctx, ctx->scratch_b, bp_xyz, (size_t)18 * n2, "source boundary points"));
        expand_f4_kernel<<<(6 * n2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch_b.as<float>(), 6 * n2, ctx->bp2.as<float4>());
        ctx->launches++;
    }
    PW_TRY(ctx->bpstd2.reserve(ctx, (size_t)n2 * 4));
    if (bp_std) { PW_TRY(upload_checked(ctx, ctx->bpstd2, bp_std, (size_t)n2, "BPstd2")); }
    else PW_CUDA(cudaMemsetAsync(ctx->bpstd2.p, 0, (size_t)n2 * 4, ctx->stream));
    PW_TRY(ctx->patch_off.reserve(ctx, (size_t)(n2 + 1) * 4));
    int mp = 0;
    if (patch_off) {
        mp = patch_off[n2];
        if (mp < 0 || patch_off[0] != 0) { set_error(ctx, "source_upload: bad patch offsets"); return PWICP_ERR_ARG; }
        PW_CUDA(cudaMemcpyAsync(ctx->patch_off.p, patch_off, (size_t)(n2 + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        PW_CUDA(cudaMemsetAsync(ctx->patch_off.p, 0, (size_t)(n2 + 1) * 4, ctx->stream));
    }
    ctx->mp2 = mp;
    if (mp > 0) {
        PW_TRY(upload_checked(ctx, ctx->patch_xyz, patch_xyz, (size_t)3 * mp, "source patch points"));
        PW_TRY(ctx->patch_id.reserve(ctx, (size_t)mp * 4));
        patch_id_kernel<<<(mp + 255) / 256, 256, 0, ctx->stream>>>(ctx->patch_off.as<int>(), n2, mp, ctx->patch_id.as<int>());
        ctx->launches++;
    }
    ctx->n2 = n2;
    PW_TRY(reset_seeds(ctx));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

int pwicp_clouds_upload(pwicp_ctx* p, const float* cloud1, int m1, const float* cloud2, int m2) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || m2 < 1 || !cloud2 || (cloud1 && m1 < 1)) { set_error(ctx, "clouds_upload: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    if (!cloud1) {
        // the reference epoch of a series stays resident: cloud1 and its grid are kept, only cloud2 is replaced
        if (ctx->m1 < 1 || !ctx->c1.dev.nlevels) { set_error(ctx, "clouds_upload: no resident cloud1 to keep"); return PWICP_ERR_ARG; }
        m1 = ctx->m1;
        ctx->m2 = 0;
    } else {
        ctx->m1 = ctx->m2 = 0;
        PW_TRY(upload_checked(ctx, ctx->scratch_a, cloud1, (size_t)3 * m1, "cloud1"));
        PW_TRY(grid_build(ctx, ctx->c1, ctx->scratch_a.as<float>(), m1));
    }
    PW_TRY(reset_seeds(ctx));
    PW_TRY(upload_checked(ctx, ctx->cloud2, cloud2, (size_t)3 * m2, "cloud2"));
    ctx->m1 = m1; ctx->m2 = m2;
    return PWICP_OK;
}

int pwicp_source_download(pwicp_ctx* p, float* cloud2, float* ct_xyz, float* bp_xyz, float* patch_xyz) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx) return PWICP_ERR_ARG;
    PW_CUDA(cudaSetDevice(ctx->device));
    if (cloud2 && ctx->m2) PW_CUDA(cudaMemcpyAsync(cloud2, ctx->cloud2.p, (size_t)3 * ctx->m2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (patch_xyz && ctx->mp2) PW_CUDA(cudaMemcpyAsync(patch_xyz, ctx->patch_xyz.p, (size_t)3 * ctx->mp2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    const int n2 = ctx->n2;
    if ((ct_xyz || bp_xyz) && n2) {
        PW_TRY(ctx->scratch_b.reserve(ctx, (size_t)21 * n2 * 4));
        float* tmp = ctx->scratch_b.as<float>();
        if (ct_xyz) {
            pack_f4_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->ct2.as<float4>(), n2, tmp);
            PW_CUDA(cudaMemcpyAsync(ct_xyz, tmp, (size_t)3 * n2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (bp_xyz) {
            pack_f4_kernel<<<(6 * n2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->bp2.as<float4>(), 6 * n2, tmp + (size_t)3 * n2);
            PW_CUDA(cudaMemcpyAsync(bp_xyz, tmp + (size_t)3 * n2, (size_t)18 * n2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        ctx->launches += 2;
    }
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

// ---- A1 --------------------------------------------------------------------------------------
int pwicp_nn(pwicp_ctx* p, int which, const float* qry, int nq, int* idx, float* d2) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || nq < 0 || (nq && !qry)) { set_error(ctx, "nn: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    const GridOwner& g = (which == PWICP_TGT_CLOUD1) ? ctx->c1 : ctx->tgt;
    if (!g.dev.nlevels) { set_error(ctx, "nn: target not uploaded"); return PWICP_ERR_ARG; }
    if (nq == 0) return PWICP_OK;
    PW_TRY(upload_checked(ctx, ctx->scratch_a, qry, (size_t)3 * nq, "nn queries"));
    PW_TRY(ctx->scratch_b.reserve(ctx, (size_t)nq * 8));
    int* idx_d = ctx->scratch_b.as<int>();
    float* d2_d = reinterpret_cast<float*>(idx_d + nq);
    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    PW_TRY(nn_query_packed(ctx, g.dev, ctx->scratch_a.as<float>(), nq, idx_d, d2_d));
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (idx) PW_CUDA(cudaMemcpyAsync(idx, idx_d, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (d2) PW_CUDA(cudaMemcpyAsync(d2, d2_d, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    PW_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return PWICP_OK;
}

// ---- A3-A6 -----------------------------------------------------------------------------------
int pwicp_icp_source_upload(pwicp_ctx* p, const float* src, int n) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || n < 1 || !src) { set_error(ctx, "icp_source_upload: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(upload_checked(ctx, ctx->scratch_a, src, (size_t)3 * n, "icp source"));
    ctx->icp_seed_valid = false;
    PW_TRY(icp_expand_source(ctx, ctx->scratch_a.as<float>(), n));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

int pwicp_icp_source_all(pwicp_ctx* p) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || ctx->n2 < 1) { set_error(ctx, "icp_source_all: no source uploaded"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(ctx->icp_src.reserve(ctx, (size_t)ctx->n2 * sizeof(float4)));
    PW_CUDA(cudaMemcpyAsync(ctx->icp_src.p, ctx->ct2.p, (size_t)ctx->n2 * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->n_icp = ctx->n2;
    ctx->icp_seed_valid = false;
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

int pwicp_icp_run(pwicp_ctx* p, const pwicp_icp_params* prm, float* T16, pwicp_icp_result* res,
                  double* mse_trace, float* T_trace, int* idx_trace) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx) return PWICP_ERR_ARG;
    PW_CUDA(cudaSetDevice(ctx->device));
    pwicp_icp_params d; pwicp_icp_default_params(&d);
    return icp_run_device(ctx, prm ? *prm : d, T16, res, mse_trace, T_trace, idx_trace);
}

int pwicp_icp_phase_profile(pwicp_ctx* p, double* phase_us, int cap) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || cap < 0 || !phase_us || ctx->icp_prof_iters < 1 || !ctx->icp_partials.p) { set_error(ctx, "icp_phase_profile: no inner loop has run"); return 0; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return 0;
    const int m = std::min(cap, ctx->icp_prof_iters), M = ctx->icp_prof_max_iter;
    std::vector<unsigned long long> ns((size_t)M + 1 + (size_t)M * 4);
    cudaMemcpy(ns.data(), ctx->icp_partials.as<char>() + ctx->icp_prof_off_ns, ns.size() * 8, cudaMemcpyDeviceToHost);
    for (int k = 0; k < m; ++k) {
        const unsigned long long t0 = ns[k];
        for (int j = 0; j < 4; ++j) phase_us[k * 4 + j] = (double)(ns[(size_t)M + 1 + (size_t)k * 4 + j] - t0) * 1e-3;
    }
    return m;
}

int pwicp_icp_profile(pwicp_ctx* p, double* iter_us, int* searched, int cap) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || cap < 0 || ctx->icp_prof_iters < 1 || !ctx->icp_partials.p) { set_error(ctx, "icp_profile: no inner loop has run"); return 0; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return 0;
    const int m = std::min(cap, ctx->icp_prof_iters);
    if (searched && m) cudaMemcpy(searched, ctx->icp_partials.as<char>() + ctx->icp_prof_off_searched, (size_t)m * 4, cudaMemcpyDeviceToHost);
    if (iter_us && m) {
        std::vector<unsigned long long> ns((size_t)m + 1);
        cudaMemcpy(ns.data(), ctx->icp_partials.as<char>() + ctx->icp_prof_off_ns, ((size_t)m + 1) * 8, cudaMemcpyDeviceToHost);
        for (int k = 0; k < m; ++k) iter_us[k] = (double)(ns[k + 1] - ns[k]) * 1e-3;
    }
    return m;
}

int pwicp_icp_order(pwicp_ctx* p, int* perm) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !perm || ctx->n_icp < 1 || !ctx->icp_perm.p) { set_error(ctx, "icp_order: no inner loop has run"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_CUDA(cudaMemcpy(perm, ctx->icp_perm.p, (size_t)ctx->n_icp * 4, cudaMemcpyDeviceToHost));
    return PWICP_OK;
}

// Host-buffer call in the shape of P2PICPwithPatchNormal (reference src/Registration.cpp:1255-1269).
// The three uploads run on a copy stream; the grid build over the target overlaps the transfer of
// the normals and of the source, the finite checks accumulate into one device flag that is read
// once, right before the inner loop starts.
int pwicp_icp_p2plane(pwicp_ctx* p, const float* tgt_xyz, const float* tgt_nrm, int n1, const float* src_xyz,
                      int n2, const pwicp_icp_params* prm, float* T16, pwicp_icp_result* res) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || n1 < 1 || n2 < 1 || !tgt_xyz || !tgt_nrm || !src_xyz) {
        set_error(ctx, "icp_p2plane: bad arguments"); return PWICP_ERR_ARG;
    }
    PW_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) {
        PW_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->copy_ev) PW_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    ctx->n1 = 0;
    PW_TRY(ctx->tgt_xyz.reserve(ctx, (size_t)3 * n1 * 4 + 64));
    PW_TRY(ctx->tgt_nrm_raw.reserve(ctx, (size_t)3 * n1 * 4 + 64));
    PW_TRY(ctx->scratch_a.reserve(ctx, (size_t)3 * n2 * 4 + 64));
    PW_TRY(ctx->scratch_d.reserve(ctx, 256));
    // earlier work on the library stream may still read these buffers
    PW_CUDA(cudaEventRecord(ctx->copy_ev[3], ctx->stream));
    PW_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[3], 0));
    // order of the uploads = order of first use: target points (grid build), source (sort + iteration-0 search), target
    // normals (row terms of the loop) -- the last 12 bytes per target travel while the source is sorted and searched
    PW_CUDA(cudaMemcpyAsync(ctx->tgt_xyz.p, tgt_xyz, (size_t)3 * n1 * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    PW_CUDA(cudaEventRecord(ctx->copy_ev[0], ctx->copy_stream));
    PW_CUDA(cudaMemcpyAsync(ctx->scratch_a.p, src_xyz, (size_t)3 * n2 * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    PW_CUDA(cudaEventRecord(ctx->copy_ev[2], ctx->copy_stream));
    PW_CUDA(cudaMemcpyAsync(ctx->tgt_nrm_raw.p, tgt_nrm, (size_t)3 * n1 * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    PW_CUDA(cudaEventRecord(ctx->copy_ev[1], ctx->copy_stream));

    int* flag = ctx->scratch_d.as<int>() + 16;            // [0..5] belong to the bounding-box reduction
    PW_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    PW_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[0], 0));
    PW_TRY(finite_accumulate_dev(ctx, ctx->tgt_xyz.as<float>(), (size_t)3 * n1, flag));
    ctx->tgt_has_std = false;
    ctx->tgt_has_ok = false;
    // a non-finite coordinate would poison the bounding box: the verdict on the target comes back with the box (one
    // round trip), and the build does not wait for its own kernels at the end
    int bad = 0;
    {
        const int rc = grid_build(ctx, ctx->tgt, ctx->tgt_xyz.as<float>(), n1, flag, false);
        if (rc != PWICP_OK) {
            cudaStreamSynchronize(ctx->copy_stream);
            if (rc == PWICP_ERR_NONFINITE) set_error(ctx, "target centroids: non-finite value in input");
            return rc;
        }
    }
    PW_TRY(reset_seeds(ctx));
    PW_TRY(ctx->tgt_aux.reserve(ctx, (size_t)n1 * sizeof(float4)));
    PW_TRY(ctx->tgt_ok.reserve(ctx, (size_t)n1));
    PW_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[2], 0));
    PW_TRY(finite_accumulate_dev(ctx, ctx->scratch_a.as<float>(), (size_t)3 * n2, flag));
    ctx->icp_seed_valid = false;
    PW_TRY(icp_expand_source(ctx, ctx->scratch_a.as<float>(), n2));
    // a non-finite source coordinate would send the search to an invalid address: the verdict on the source is read
    // before the loop is enqueued (the normals are still travelling; theirs only poison arithmetic and is read last)
    PW_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) { cudaStreamSynchronize(ctx->copy_stream); set_error(ctx, "icp_p2plane: non-finite value in input"); return PWICP_ERR_NONFINITE; }
    ctx->n1 = n1;
    ctx->aux_deferred = true; ctx->aux_deferred_n1 = n1; ctx->aux_deferred_flag = flag;
    ctx->tail_flag_dev = flag; ctx->tail_flag_host = 0;      // the verdict on the normals comes back with the result
    int st = pwicp_icp_run(p, prm, T16, res, nullptr, nullptr, nullptr);
    ctx->tail_flag_dev = nullptr;
    if (ctx->aux_deferred) { const int s2 = finish_deferred_aux(ctx); if (st == PWICP_OK) st = s2; }   // the loop failed before it got there
    if (st != PWICP_OK) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return st; }
    if (ctx->tail_flag_host) { set_error(ctx, "icp_p2plane: non-finite value in input"); return PWICP_ERR_NONFINITE; }
    return PWICP_OK;
}


// ---- F3: per-patch statistics -----------------------------------------------------------------
int pwicp_patch_stats(pwicp_ctx* p, const float* patch_xyz, const int* patch_off, int n_patches,
                      float* ct3, float* bp18, float* nrm3, unsigned char* nrm_ok, float* bp_std, float* ct_std) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || n_patches < 1 || !patch_xyz || !patch_off) { set_error(ctx, "patch_stats: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    const long long m = patch_off[n_patches];
    if (patch_off[0] != 0 || m < 1) { set_error(ctx, "patch_stats: offsets must start at 0 and cover at least one point"); return PWICP_ERR_ARG; }
    for (int i = 0; i < n_patches; ++i)
        if (patch_off[i + 1] < patch_off[i]) { set_error(ctx, "patch_stats: offsets must be non-decreasing"); return PWICP_ERR_ARG; }
    const size_t np = (size_t)n_patches;
    PW_TRY(upload_checked(ctx, ctx->scratch_a, patch_xyz, (size_t)3 * m, "patch points"));
    PW_TRY(ctx->scratch_b.reserve(ctx, (np + 1) * sizeof(int)));
    PW_CUDA(cudaMemcpyAsync(ctx->scratch_b.p, patch_off, (np + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    // outputs: ct 3, bp 18, nrm 3, bpstd 1, ctstd 1 floats per patch, then the ok bytes
    PW_TRY(ctx->scratch_c.reserve(ctx, np * 26 * sizeof(float) + np));
    float* d = ctx->scratch_c.as<float>();
    float *d_ct = d, *d_bp = d + 3 * np, *d_nrm = d + 21 * np, *d_bs = d + 24 * np, *d_cs = d + 25 * np;
    unsigned char* d_ok = reinterpret_cast<unsigned char*>(d + 26 * np);
    PW_TRY(patch_stats_dev(ctx, ctx->scratch_a.as<float>(), ctx->scratch_b.as<int>(), n_patches,
                           (ct3 || bp18) ? d_ct : nullptr, bp18 ? d_bp : nullptr, (nrm3 || nrm_ok) ? d_nrm : nullptr,
                           (nrm3 || nrm_ok) ? d_ok : nullptr, (bp_std || ct_std) ? d_bs : nullptr,
                           (bp_std || ct_std) ? d_cs : nullptr));
    if (ct3) PW_CUDA(cudaMemcpyAsync(ct3, d_ct, np * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (bp18) PW_CUDA(cudaMemcpyAsync(bp18, d_bp, np * 18 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (nrm3) PW_CUDA(cudaMemcpyAsync(nrm3, d_nrm, np * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (nrm_ok) PW_CUDA(cudaMemcpyAsync(nrm_ok, d_ok, np, cudaMemcpyDeviceToHost, ctx->stream));
    if (bp_std) PW_CUDA(cudaMemcpyAsync(bp_std, d_bs, np * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (ct_std) PW_CUDA(cudaMemcpyAsync(ct_std, d_cs, np * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    return PWICP_OK;
}

// ---- outer iteration / loop -------------------------------------------------------------------
int pwicp_single_iteration(pwicp_ctx* p, const pwicp_pair_params* pp, pwicp_state* st,
                           const pwicp_icp_params* icp, float* T16, double* vcm36,
                           unsigned char* stable_flags, pwicp_iter_stats* stats) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !pp || !st || !T16) { set_error(ctx, "single_iteration: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    pwicp_icp_params d; pwicp_icp_default_params(&d);
    return outer_single_iteration(ctx, *pp, st, icp ? *icp : d, T16, vcm36, stable_flags, stats);
}

int pwicp_piecewise_icp(pwicp_ctx* p, const pwicp_pair_params* pp, int is_manual, float DTinit,
                        const pwicp_icp_params* icp, int max_outer, float* DTseries, int* n_series,
                        float* T16, double* vcm36, int* n_outer, pwicp_iter_stats* stats_per_iter) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !pp || !DTseries || !n_series || !T16 || max_outer < 1) { set_error(ctx, "piecewise_icp: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    pwicp_state st;
    st.toStage2 = 0; st.toStage3 = 0;                                       // :623-624
    if (!is_manual) {
        if (!ctx->c1.dev.nlevels || ctx->m2 < 1) { set_error(ctx, "piecewise_icp: clouds not uploaded"); return PWICP_ERR_ARG; }
        double Dist75 = 0;
        PW_TRY(percentile_dev(ctx, ctx->c1.dev, ctx->cloud2.as<float>(), ctx->m2, nullptr, nullptr, ctx->m2, 0.75f, &Dist75, nullptr));
        DTinit = Dist75 * 3.0;                                              // :627-630
    }
    st.currDT = DTinit; st.BBchange_1 = 0.f; st.BBchange_2 = 0.f;
    float transMat[16];
    for (int k = 0; k < 16; ++k) transMat[k] = (k % 5 == 0) ? 1.f : 0.f;
    int ns = 0, count = 0;
    DTseries[ns++] = st.currDT;
    float total_ms = 0.f;
    pwicp_icp_params ip;
    pwicp_icp_default_params(&ip);
    if (icp) ip = *icp;
    while (!st.toStage3 && count < max_outer) {
        float cur[16];
        int rc = outer_single_iteration(ctx, *pp, &st, ip, cur, vcm36, nullptr,
                                        stats_per_iter ? stats_per_iter + count : nullptr);
        if (rc != PWICP_OK) { *n_series = ns; if (n_outer) *n_outer = count; memcpy(T16, transMat, sizeof(transMat)); return rc; }
        total_ms += ctx->last_ms;
        mat4_mul(cur, transMat, transMat);                                  // transMat = cur * transMat, :687
        DTseries[ns++] = st.currDT;
        ++count;
    }
    ctx->last_ms = total_ms;
    *n_series = ns;
    if (n_outer) *n_outer = count;
    memcpy(T16, transMat, sizeof(transMat));
    if (!st.toStage3) {          // the reference loops until stage 3 (:682); max_outer is this library's guard
        set_error(ctx, "piecewise_icp: max_outer iterations without reaching stage 3 (no VCM; T16 = transformation so far)");
        return PWICP_ERR_MAX_OUTER;
    }
    return PWICP_OK;
}

// ---- stand-alone pieces -------------------------------------------------------------------------
int pwicp_percentile_nn(pwicp_ctx* p, const float* cloud1, int m1, const float* cloud2, int m2,
                        float pct, double* out) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || m1 < 1 || m2 < 1 || !out) { set_error(ctx, "percentile_nn: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    GridOwner& g = ctx->prep;            // persistent buffers (cudaMalloc/cudaFree cost ms), rebuilt per call
    PW_TRY(upload_checked(ctx, ctx->scratch_a, cloud1, (size_t)3 * m1, "cloud1"));
    int rc = grid_build(ctx, g, ctx->scratch_a.as<float>(), m1);
    if (rc == PWICP_OK) rc = upload_checked(ctx, ctx->scratch_c, cloud2, (size_t)3 * m2, "cloud2");
    if (rc == PWICP_OK) rc = percentile_dev(ctx, g.dev, ctx->scratch_c.as<float>(), m2, nullptr, nullptr, m2, pct, out, nullptr);
    return rc;
}

int pwicp_overlap_ratio(pwicp_ctx* p, const float* cloud1, int m1, const float* cloud2, int m2,
                        float DTinit, float* out) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || m1 < 1 || m2 < 1 || !out) { set_error(ctx, "overlap_ratio: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    GridOwner& g = ctx->prep;            // persistent buffers (cudaMalloc/cudaFree cost ms), rebuilt per call
    PW_TRY(upload_checked(ctx, ctx->scratch_a, cloud1, (size_t)3 * m1, "cloud1"));
    int rc = grid_build(ctx, g, ctx->scratch_a.as<float>(), m1);
    if (rc == PWICP_OK) rc = upload_checked(ctx, ctx->scratch_c, cloud2, (size_t)3 * m2, "cloud2");
    unsigned long long cnt = 0;
    if (rc == PWICP_OK) rc = ctx->scratch_b.reserve(ctx, (size_t)m2 * 4 + 64);
    if (rc == PWICP_OK) {
        float* d2 = ctx->scratch_b.as<float>();
        unsigned long long* cd = reinterpret_cast<unsigned long long*>(ctx->scratch_d.p);
        cudaMemsetAsync(cd, 0, 8, ctx->stream);
        rc = nn_query_packed(ctx, g.dev, ctx->scratch_c.as<float>(), m2, nullptr, d2);
        if (rc == PWICP_OK) {
            count_below_kernel<<<(m2 + 255) / 256, 256, 0, ctx->stream>>>(d2, m2, DTinit, cd);
            ctx->launches++;
            if (cudaMemcpyAsync(&cnt, cd, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error(ctx, "overlap_ratio: copy failed"); rc = PWICP_ERR_CUDA; }
        }
    }
    if (rc == PWICP_OK) *out = float((int)cnt) / float(m2);        // :613
    return rc;
}

int pwicp_self_nn(pwicp_ctx* p, const float* xyz, int n, float* d2) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !xyz || !d2 || n < 2) { set_error(ctx, "self_nn: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    GridOwner& g = ctx->prep;            // persistent buffers (cudaMalloc/cudaFree cost ms), rebuilt per call
    PW_TRY(upload_checked(ctx, ctx->scratch_a, xyz, (size_t)3 * n, "cloud"));
    int rc = grid_build(ctx, g, ctx->scratch_a.as<float>(), n);
    if (rc == PWICP_OK) rc = ctx->scratch_b.reserve(ctx, (size_t)n * 4);
    if (rc == PWICP_OK) rc = self_nn_dev(ctx, g.dev, ctx->scratch_b.as<float>());
    if (rc == PWICP_OK) {
        if (cudaMemcpyAsync(d2, ctx->scratch_b.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error(ctx, "self_nn: copy failed"); rc = PWICP_ERR_CUDA; }
    }
    return rc;
}

// ---- F4: PCpreprocessing (src/CommonFunc.cpp:423-452) ----------------------------------------------------------------
int pwicp_voxel_grid(pwicp_ctx* p, const float* xyz, int n, float leaf, float* out_xyz, int* n_out) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !xyz || !out_xyz || !n_out || n < 1 || !(leaf > 0.f)) { set_error(ctx, "voxel_grid: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(upload_checked(ctx, ctx->scratch_a, xyz, (size_t)3 * n, "cloud"));
    PW_TRY(ctx->scratch_c.reserve(ctx, (size_t)3 * n * 4));
    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    PW_TRY(voxel_grid_dev(ctx, ctx->scratch_a.as<float>(), n, leaf, ctx->scratch_c.as<float>(), n_out));
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PW_CUDA(cudaMemcpyAsync(out_xyz, ctx->scratch_c.p, (size_t)3 * *n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    PW_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return PWICP_OK;
}

static int knn_mean_dist_resident(Ctx* ctx, const float* xyz_dev, int n, int k, float* mean_dist_host) {
    GridOwner& g = ctx->prep;                      // persistent buffers, rebuilt for this cloud
    PW_TRY(grid_build(ctx, g, xyz_dev, n));
    PW_TRY(ctx->scratch_b.reserve(ctx, (size_t)n * 4));
    PW_CUDA(cudaEventRecord(ctx->ev2, ctx->stream));
    PW_TRY(knn_mean_dist_dev(ctx, g.dev, k, ctx->scratch_b.as<float>()));
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PW_CUDA(cudaMemcpyAsync(mean_dist_host, ctx->scratch_b.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    PW_CUDA(cudaEventElapsedTime(&ctx->prep_kernel_ms, ctx->ev2, ctx->ev1));
    return PWICP_OK;
}

int pwicp_knn_mean_dist(pwicp_ctx* p, const float* xyz, int n, int k, float* mean_dist) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !xyz || !mean_dist || k < 1 || k > 32 || n <= k) { set_error(ctx, "knn_mean_dist: bad arguments (1 <= k <= 32 < n)"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(upload_checked(ctx, ctx->scratch_a, xyz, (size_t)3 * n, "cloud"));
    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    PW_TRY(knn_mean_dist_resident(ctx, ctx->scratch_a.as<float>(), n, k, mean_dist));
    PW_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));      // grid build + k-NN kernel
    return PWICP_OK;
}

// k nearest neighbours (indices, the point itself first) and PCA normals of every point: the segmentation front end
int pwicp_knn_normals(pwicp_ctx* p, const float* xyz, int n, int k, int* neighbors, double* normals) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !xyz || n < 1 || k < 1 || k > 64 || n < k || (!neighbors && !normals)) { set_error(ctx, "knn_normals: bad arguments (1 <= k <= 64 <= n)"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(upload_checked(ctx, ctx->scratch_a, xyz, (size_t)3 * n, "cloud"));
    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    GridOwner& g = ctx->prep;                      // persistent buffers, rebuilt for this cloud
    PW_TRY(grid_build(ctx, g, ctx->scratch_a.as<float>(), n));
    PW_TRY(ctx->scratch_b.reserve(ctx, (size_t)n * k * sizeof(int)));
    PW_TRY(ctx->scratch_c.reserve(ctx, (size_t)n * 3 * sizeof(double)));
    PW_CUDA(cudaEventRecord(ctx->ev2, ctx->stream));
    PW_TRY(knn_normals_dev(ctx, g.dev, k, neighbors ? ctx->scratch_b.as<int>() : nullptr, normals ? ctx->scratch_c.as<double>() : nullptr));
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (neighbors) PW_CUDA(cudaMemcpyAsync(neighbors, ctx->scratch_b.p, (size_t)n * k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (normals) PW_CUDA(cudaMemcpyAsync(normals, ctx->scratch_c.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    PW_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));           // grid build + kernel
    PW_CUDA(cudaEventElapsedTime(&ctx->prep_kernel_ms, ctx->ev2, ctx->ev1));    // the kernel alone
    return PWICP_OK;
}

// VoxelGrid (optional) + StatisticalOutlierRemoval in one call: the cloud stays on the device between the two steps; the
// selection (mean + mult * stddev of the mean distances, sequential double sums like PCL) runs on the host over the
// downloaded 4 bytes per point.  out_xyz must hold n points.
int pwicp_preprocess(pwicp_ctx* p, const float* xyz, int n, int downsample, float leaf, int k, double std_mult,
                     float* out_xyz, int* n_out) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !xyz || !out_xyz || !n_out || n < 1 || (downsample && !(leaf > 0.f)) || k < 1 || k > 32) {
        set_error(ctx, "preprocess: bad arguments"); return PWICP_ERR_ARG;
    }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(upload_checked(ctx, ctx->scratch_a, xyz, (size_t)3 * n, "cloud"));
    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    const float* cur = ctx->scratch_a.as<float>();
    int m = n;
    if (downsample) {
        PW_TRY(ctx->scratch_c.reserve(ctx, (size_t)3 * n * 4));
        PW_TRY(voxel_grid_dev(ctx, ctx->scratch_a.as<float>(), n, leaf, ctx->scratch_c.as<float>(), &m));
        cur = ctx->scratch_c.as<float>();
    }
    std::vector<float> pts((size_t)3 * m), md((size_t)m);
    PW_CUDA(cudaMemcpyAsync(pts.data(), cur, (size_t)3 * m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (m <= k) {                                  // StatisticalOutlierRemoval needs more than k points: pass through
        PW_CUDA(cudaStreamSynchronize(ctx->stream));
        std::memcpy(out_xyz, pts.data(), (size_t)3 * m * 4);
        *n_out = m;
        return PWICP_OK;
    }
    PW_TRY(knn_mean_dist_resident(ctx, cur, m, k, md.data()));
    PW_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));      // voxel grid + cloud download + grid build + k-NN kernel
    double sum = 0, sq = 0;
    // [PCL 1.8.1 StatisticalOutlierRemoval::applyFilterIndices] distances is a vector<float>: the square is a float
    // product, rounded before it is widened; no clamp of the variance; a point is an outlier iff distance > threshold
    for (int i = 0; i < m; ++i) { const float d_sq = md[i] * md[i]; sum += md[i]; sq += d_sq; }
    const double mean = sum / m;
    const double var = (sq - sum * sum / m) / (m - 1);
    const double thr = mean + std_mult * std::sqrt(var);
    int kept = 0;
    for (int i = 0; i < m; ++i)
        if (!(md[i] > thr)) { std::memcpy(out_xyz + 3 * (size_t)kept, pts.data() + 3 * (size_t)i, 12); ++kept; }
    *n_out = kept;
    return PWICP_OK;
}

int pwicp_vcm(pwicp_ctx* p, const float* src, int n, double* vcm36, int* singular) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !src || !vcm36 || n < 1) { set_error(ctx, "vcm: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->tgt.dev.nlevels) { set_error(ctx, "vcm: target not uploaded"); return PWICP_ERR_ARG; }
    PW_TRY(upload_checked(ctx, ctx->scratch_a, src, (size_t)3 * n, "vcm source"));
    PW_TRY(ctx->scratch_b.reserve(ctx, (size_t)n * sizeof(float4)));
    expand_f4_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch_a.as<float>(), n, ctx->scratch_b.as<float4>());
    ctx->launches++;
    return vcm_dev(ctx, ctx->scratch_b.as<float4>(), n, vcm36, singular, nullptr, nullptr);
}

int pwicp_transform(pwicp_ctx* p, float* xyz, int n, const float* T16) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !xyz || n < 0 || !T16) { set_error(ctx, "transform: bad arguments"); return PWICP_ERR_ARG; }
    if (n == 0) return PWICP_OK;
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(upload_packed(ctx, ctx->scratch_a, xyz, (size_t)3 * n));
    PW_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    PW_TRY(transform_packed_dev(ctx, ctx->scratch_a.as<float>(), (size_t)n, T16));
    PW_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    PW_CUDA(cudaMemcpyAsync(xyz, ctx->scratch_a.p, (size_t)3 * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PW_CUDA(cudaStreamSynchronize(ctx->stream));
    PW_CUDA(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return PWICP_OK;
}

int pwicp_octree_bbox(pwicp_ctx* p, const float* xyz, int n, double res, double* bb6) {
    Ctx* ctx = reinterpret_cast<Ctx*>(p);
    if (!ctx || !xyz || n < 1 || !bb6) { set_error(ctx, "octree_bbox: bad arguments"); return PWICP_ERR_ARG; }
    PW_CUDA(cudaSetDevice(ctx->device));
    PW_TRY(upload_packed(ctx, ctx->scratch_a, xyz, (size_t)3 * n));
    float mn[3], mx[3];
    PW_TRY(bbox_packed_dev(ctx, ctx->scratch_a.as<float>(), (size_t)n, mn, mx));
    octree_cube(mn, mx, res, bb6);
    return PWICP_OK;
}

float pwicp_bbox_corner_change(const double* bb6, const float* T16) { return bbox_corner_change_host(bb6, T16); }

void pwicp_matrix2angle(const float* T, float* ang) {
    double ax, ay, az;
    if (T[8] == 1 || T[8] == -1) {
        az = 0;
        const double dlta = std::atan2(T[1], T[2]);
        if (T[8] == -1) { ay = M_PI / 2; ax = az + dlta; }
        else { ay = -M_PI / 2; ax = -az + dlta; }
    } else {
        ay = -std::asin(T[8]);
        ax = std::atan2(T[9] / std::cos(ay), T[10] / std::cos(ay));
        az = std::atan2(T[4] / std::cos(ay), T[0] / std::cos(ay));
    }
    ang[0] = (float)ax; ang[1] = (float)ay; ang[2] = (float)az;
}

void pwicp_mat4_mul(const float* A, const float* B, float* C) { mat4_mul(A, B, C); }

}  // extern "C"
