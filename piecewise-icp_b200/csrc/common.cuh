// common.cuh -- shared declarations of libpwicp.so (sm_100a only; compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pwicp.h"

namespace pwicp {

constexpr int kMaxLevels = 3;      // grid pyramid: cell sizes h, 8h, 64h
constexpr int kLevelFactor = 8;
constexpr int kRingsPerLevel = 2;  // rings searched on a level before moving to the coarser one
#ifndef PWICP_ICP_THREADS
#define PWICP_ICP_THREADS 512
#endif
constexpr int kIcpThreads = PWICP_ICP_THREADS;   // one CTA of 16 warps per SM; part of the reduction geometry
constexpr int kIcpWarps = kIcpThreads / 32;
constexpr int kNumVals = 28;       // 21 ATA + 6 ATb + sum d2
constexpr int kStageSlots = 3;     // batches of streamed per-point data in flight per warp (cp.async ring)
constexpr int kMaxIcpIter = 1024;

struct GridLevel {
    const float4* pts;           // cell-sorted targets: x, y, z, original index (int bits)
    const uint32_t* cell_start;  // ncells + 1
    int dx, dy, dz;
    float inv_h, inv_h2;         // 1/h, 1/h^2
};

struct GridDev {
    float ox, oy, oz;
    int nlevels;
    int n;
    GridLevel lv[kMaxLevels];
    const uint32_t* inv_perm;    // original index -> position in level 0
};

// simple device buffer that grows on demand (never shrinks: rebuilding a grid of the same size
// does not touch the allocator -- cudaMalloc/cudaFree serialise across processes and cost ms)
struct Ctx;
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(Ctx* ctx, size_t bytes);
    void release();
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Owns the device memory of one grid pyramid.
struct GridOwner {
    GridDev dev{};
    DevBuf pts[kMaxLevels], cells[kMaxLevels];
    DevBuf inv_perm;
    DevBuf perm0_buf;            // level-0 order: position -> original index
    uint32_t* perm0 = nullptr;
    float h0 = 0.f;
    int n = 0;
    void release();
};

struct Ctx;

// error helpers ---------------------------------------------------------------------------
void set_error(Ctx* c, const std::string& msg);
#define PW_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e__));            \
            return PWICP_ERR_CUDA;                                                           \
        }                                                                                    \
    } while (0)
#define PW_TRY(call)                                                                         \
    do {                                                                                     \
        int s__ = (call);                                                                    \
        if (s__ != PWICP_OK) return s__;                                                     \
    } while (0)

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // host-buffer calls: uploads overlap the grid build
    cudaEvent_t copy_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr, ev5 = nullptr, ev6 = nullptr, ev7 = nullptr;
    cudaEvent_t ev_o0 = nullptr, ev_o1 = nullptr;   // outer iteration
    std::string err;
    float last_ms = 0.f;
    long long launches = 0;
    float cells_per_point = 4.0f;
    int num_sms = 0;

    // target (centroids)
    GridOwner tgt;
    DevBuf tgt_aux;      // float4 per target in level-0 order: nx, ny, nz, ctstd
    DevBuf tgt_ok;       // uint8 per target in level-0 order (calPatchNormal success)
    DevBuf tgt_xyz, tgt_nrm_raw, tgt_std_raw, tgt_ok_raw;   // original-order copies (rebuild)
    bool tgt_has_std = false, tgt_has_ok = false;
    int n1 = 0;
    // full cloud1
    GridOwner c1;
    int m1 = 0;
    // F4 pre-processing: grid over the cloud being filtered (buffers persist across calls: cudaMalloc/cudaFree cost ms)
    GridOwner prep;
    float prep_kernel_ms = 0.f;                  // duration of the last k-NN kernel alone
    // source
    DevBuf ct2, bp2, bpstd2, patch_xyz, patch_id, patch_off, cloud2;
    int n2 = 0, m2 = 0, mp2 = 0;
    // inner loop
    DevBuf icp_src, icp_sorted, icp_perm, icp_work, icp_partials, icp_out, icp_idx;
    DevBuf icp_seed, icp_match;   // seeds in icp_src order / matches in processing order
    bool icp_seed_valid = false;
    // host-buffer call (pwicp_icp_p2plane): the target normals are still on their way up while the source is sorted and
    // searched; the inner loop picks them up right before it needs them (icp_enqueue -> finish_deferred_aux)
    bool aux_deferred = false;
    int aux_deferred_n1 = 0;
    int* aux_deferred_flag = nullptr;            // device flag the finite check of the normals accumulates into
    const int* tail_flag_dev = nullptr;          // ... read back with the result of the loop (icp_run_device) ...
    int tail_flag_host = 0;                      // ... into this
    // temporal-coherence seeds of the outer iteration (level-0 positions, -1 = none)
    DevBuf ct_seed, bp_seed, pp_seed, ct_order;
    bool ct_order_valid = false;
    int n_icp = 0;
    int icp_nsplit = 0;                          // stand-alone search passes of the last inner loop (icp_enqueue)
    bool icp_attr_set = false;                   // kernel attributes of the persistent kernel set for this device
    int icp_prof_max_iter = 0;
    int icp_prof_iters = 0;                      // pwicp_icp_profile: iterations of the last run, offsets into icp_partials
    size_t icp_prof_off_searched = 0, icp_prof_off_ns = 0;
    // scratch
    DevBuf keys, vals, keys2, vals2, cub_tmp, scratch_a, scratch_b, scratch_c, scratch_d, flags, pos;
    DevBuf l2flush;
    DevBuf outer_state;                          // OuterDev (outer.cu): device-resident state of an outer iteration
    bool outer_state_zeroed = false;
    void* pinned = nullptr;   // small pinned staging area
    size_t pinned_cap = 0;
};

// grid.cu
int grid_build(Ctx* ctx, GridOwner& g, const float* xyz_dev_packed, int n, const int* bad_flag_dev = nullptr,
               bool sync_at_end = true);
int nn_query_packed(Ctx* ctx, const GridDev& g, const float* q_dev_packed, int nq, int* idx_dev,
                    float* d2_dev);
int self_nn_dev(Ctx* ctx, const GridDev& g, float* d2_dev);
int upload_packed(Ctx* ctx, DevBuf& buf, const float* host_xyz, size_t n_floats);
int check_finite_dev(Ctx* ctx, const float* dev, size_t n_floats, bool* ok);
int scan_inclusive_inplace(Ctx* ctx, uint32_t* a, size_t n);               // a[i] = a[0] + .. + a[i], any n
int spatial_order_dev(Ctx* ctx, const GridDev& g, const float4* pts, int n, uint32_t* order);   // processing order of a query set
int finite_accumulate_dev(Ctx* ctx, const float* dev, size_t n_floats, int* flag_dev);   // no host sync

// capi.cu
int finish_deferred_aux(Ctx* ctx);

// icp.cu
struct IcpLaunch { int grid; const char* out; int max_iter; };   // out: device: T[16] | n_iter, conv_state, natural iter, state
int icp_enqueue(Ctx* ctx, const pwicp_icp_params& prm, int n, const int* n_dev, bool presorted,
                bool want_mse, bool want_T, bool want_idx, IcpLaunch* L);
int icp_run_device(Ctx* ctx, const pwicp_icp_params& prm, float* T16, pwicp_icp_result* res,
                   double* mse_trace, float* T_trace, int* idx_trace);

// outer.cu
int outer_single_iteration(Ctx* ctx, const pwicp_pair_params& pp, pwicp_state* st,
                           const pwicp_icp_params& icp, float* T16, double* vcm36,
                           unsigned char* stable_flags, pwicp_iter_stats* stats);
int percentile_dev(Ctx* ctx, const GridDev& g, const float* q_packed_dev, int nq,
                   const int* patch_id_dev, const int* flags_dev, long long n_valid,
                   float pct, double* out, int* seeds_dev);
float bbox_corner_change_host(const double* bb6, const float* T16);
int icp_expand_source(Ctx* ctx, const float* packed_dev, int n);
int vcm_dev(Ctx* ctx, const float4* src_dev, int n, double* vcm36, int* singular, int* seeds_dev,
            const float4* cq_seed_dev);
int transform_packed_dev(Ctx* ctx, float* xyz_dev, size_t n, const float* T16);
int bbox_packed_dev(Ctx* ctx, const float* xyz_dev, size_t n, float* mn3, float* mx3, const int* flag_dev = nullptr,
                    int* flag_out = nullptr);
int bbox_accumulate_dev(Ctx* ctx, const float* xyz_dev, size_t n, int* out6_dev);
void octree_cube(const float* mn, const float* mx, double res, double* bb6);

// prep.cu (F4)
int voxel_grid_dev(Ctx* ctx, const float* xyz_dev, int n, float leaf, float* out_dev, int* n_out);
int knn_mean_dist_dev(Ctx* ctx, const GridDev& g, int k, float* out_dev);
int knn_normals_dev(Ctx* ctx, const GridDev& g, int k, int* neighbors_dev, double* normals_dev);

// patch.cu
int patch_stats_dev(Ctx* ctx, const float* xyz_dev, const int* off_dev, int np, float* ct, float* bp, float* nrm,
                    unsigned char* ok, float* bpstd, float* ctstd);

}  // namespace pwicp
