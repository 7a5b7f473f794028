/*
 * pwicp.h -- C ABI of the B200-native Piecewise-ICP inner registration loop (libpwicp.so).
 *
 * This is the drop-in boundary.  The reference has no plugin registry; its FFI
 * surface is the two undecorated DLL exports bound by python/main.py:12-18
 *     bool PiecewiseICP_pair_call(const char* confile, const char* outfile)       include/Registration.h:49
 *     bool PiecewiseICP_4D_call(const char*, int, int, int, float)                include/Registration.h:36
 * (exported by libpwicp_host.so, see include/pwicp_host.h) plus the C++ free functions of
 * include/Registration.h:149-229 and include/CommonFunc.h:127-183 that reach PCL.  Each entry
 * point below names the reference interface whose arithmetic it replaces.
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers unless a name says `_dev`,
 * packed float32 xyz (n x 3), row-major matrices, the context owns all device memory and one
 * CUDA stream.  Every function returns a pwicp_status (0 = ok, negative = error) and never calls
 * exit(); pwicp_last_error() gives the text.  Inputs must be finite (PWICP_ERR_NONFINITE).
 * There is no CPU fallback: without a CUDA device every compute entry fails with PWICP_ERR_CUDA.
 */
#ifndef PWICP_H
#define PWICP_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pwicp_ctx pwicp_ctx;

typedef enum {
    PWICP_OK = 0,
    PWICP_ERR_CUDA = -1,          /* no device / CUDA runtime error */
    PWICP_ERR_ARG = -2,           /* bad argument or call order */
    PWICP_ERR_NONFINITE = -3,     /* NaN/Inf in an input cloud */
    PWICP_ERR_TOO_FEW_PATCHES = -4,   /* < 4 source patches   (reference: exit, src/Registration.cpp:728-731) */
    PWICP_ERR_TOO_FEW_STABLE = -5,    /* < 4 stable patches   (reference: exit, src/Registration.cpp:864-867) */
    PWICP_ERR_TOO_FEW_CORR = -6,      /* < 3 correspondences  (PCL min_number_correspondences_) */
    PWICP_ERR_NOMEM = -7,
    PWICP_ERR_MAX_OUTER = -8      /* pwicp_piecewise_icp stopped by max_outer before stage 3: T16 valid, no VCM */
} pwicp_status;

/* convergence states of the inner loop (pcl::registration::DefaultConvergenceCriteria) */
enum { PWICP_CONV_NONE = 0, PWICP_CONV_ITERATIONS = 1, PWICP_CONV_TRANSFORM = 2,
       PWICP_CONV_ABS_MSE = 3, PWICP_CONV_REL_MSE = 4, PWICP_CONV_NO_CORR = 5 };

/* which device-resident target a query runs against */
enum { PWICP_TGT_CENTROIDS = 0,   /* CTcloud1 (pwicp_target_upload)      */
       PWICP_TGT_CLOUD1 = 1 };    /* full cloud1 (pwicp_clouds_upload)   */

/* ---- context ------------------------------------------------------------------------------ */
int  pwicp_version(void);
int  pwicp_device_count(void);                       /* 0 when no CUDA device is usable */
int  pwicp_ctx_create(int device, pwicp_ctx** out);
void pwicp_ctx_destroy(pwicp_ctx* ctx);
const char* pwicp_last_error(const pwicp_ctx* ctx);  /* ctx may be NULL: last global error */
/* device time (ms, CUDA events on the context's stream) of the last timed entry point */
float pwicp_last_device_ms(const pwicp_ctx* ctx);
/* duration of the k-NN kernel alone in the last pwicp_knn_mean_dist / pwicp_preprocess call (CUDA events) */
float pwicp_last_knn_kernel_ms(const pwicp_ctx* ctx);
/* number of kernels this library launched on the context so far (bench "gpu_launches") */
long long pwicp_launch_count(const pwicp_ctx* ctx);
/* writes a buffer larger than L2 (bench hygiene between timed steps) */
int  pwicp_flush_l2(pwicp_ctx* ctx);
int  pwicp_sync(pwicp_ctx* ctx);
/* tuning knob: average grid cells per target point of the finest level (default 4) */
int  pwicp_set_cells_per_point(pwicp_ctx* ctx, float cpp);

/* ---- uploads -------------------------------------------------------------------------------
 * pwicp_target_upload replaces the five KD-tree builds over CTcloud1 per outer iteration
 * (CorrespondenceEstimation::setInputTarget at src/Registration.cpp:738, :744, :1294 and the two
 * inside IterativeClosestPoint, :1260/:1266): one device grid per pair.  nrm = patch normals of
 * generateCentroidCloudWithPatchNormals (src/CommonFunc.cpp:357-382), nrm_ok = calPatchNormal
 * success per patch (NULL = all ok), ct_std = CTstd1 (src/Segmentation.cpp:319). */
int pwicp_target_upload(pwicp_ctx* ctx, const float* ct_xyz, const float* nrm,
                        const unsigned char* nrm_ok, const float* ct_std, int n1);
/* Rebuilds the grid from the device-resident copy of the target centroids (same result as
 * pwicp_target_upload without the host copy; what the device-resident bench step times). */
int pwicp_target_rebuild(pwicp_ctx* ctx);
/* CTcloud2, BPcloud2 (6 per patch), BPstd2 and the patch point lists SVcloud2[] as one
 * concatenated array with n2+1 offsets (src/Registration.cpp:646-664). */
int pwicp_source_upload(pwicp_ctx* ctx, const float* ct_xyz, const float* bp_xyz,
                        const float* bp_std, const int* patch_off, const float* patch_xyz, int n2);
/* the pre-processed full clouds cloud1 / cloud2 of Piecewise_ICP (src/Registration.cpp:618);
 * cloud1 gets its own grid (replaces the tree build of src/CommonFunc.cpp:269-273).
 * cloud1 = NULL keeps the resident cloud1 and its grid (a 4D series registers every epoch against the same reference
 * epoch, src/Registration.cpp:95-97: the reference side is uploaded once, pwicp_target_upload likewise). */
int pwicp_clouds_upload(pwicp_ctx* ctx, const float* cloud1, int m1, const float* cloud2, int m2);
/* current (transformed) source-side data back to the host; any pointer may be NULL */
int pwicp_source_download(pwicp_ctx* ctx, float* cloud2, float* ct_xyz, float* bp_xyz, float* patch_xyz);

/* ---- A1: batched exact 1-NN ---------------------------------------------------------------
 * Replaces pcl::registration::CorrespondenceEstimation::determineCorrespondences(corrs, DBL_MAX)
 * (src/Registration.cpp:737-747, :1293-1297, :597-601; src/CommonFunc.cpp:269-273): for every
 * query the index of the nearest target point and the float squared distance
 * ((dx*dx)+dy*dy)+dz*dz, ties -> lowest index.  idx / d2 may be NULL. */
int pwicp_nn(pwicp_ctx* ctx, int which_target, const float* qry_xyz, int nq, int* idx, float* d2);

/* ---- A3-A6: inner point-to-plane ICP ------------------------------------------------------ */
typedef struct {
    int    max_iter;        /* setMaximumIterations(100)        src/Registration.cpp:1264 */
    double tf_eps;          /* setTransformationEpsilon(1e-8)   src/Registration.cpp:1262 */
    double fit_eps;         /* setEuclideanFitnessEpsilon(1e-6) src/Registration.cpp:877, :1263 */
    int    force_iters;     /* benchmark mode: run exactly max_iter iterations */
    int    rot_thr_default; /* leave PCL's rotation threshold at 0.99999 instead of 1-tf_eps */
} pwicp_icp_params;
void pwicp_icp_default_params(pwicp_icp_params* p);

typedef struct {
    int   n_iter;           /* inner iterations done */
    int   conv_state;       /* PWICP_CONV_* */
    int   grid_blocks;      /* launch geometry (informational): CTAs x warps_per_block */
    int   warps_per_block;
    int   group_batches;    /* reduction geometry: total warps of the grid | warps per CTA << 16;
                               the summation order is a function of (n_source, this) alone */
    float device_ms;        /* whole inner loop (source sort, iteration-0 pre-pass, persistent kernel), CUDA events */
    long long correspondences;  /* n_iter * n_source */
    float kernel_ms;        /* the persistent kernel alone (CUDA events around its launches, summed) */
    int   natural_iters;    /* first iteration at which DefaultConvergenceCriteria was met (= n_iter unless force_iters) */
    int   natural_state;    /* ... and the PWICP_CONV_* state it reported */
    float sort_ms;          /* Morton sort + gather of the source set (CUDA events) */
    float prepass_ms;       /* iteration-0 search pre-pass of an unseeded source set (0 when seeded) */
    float research_ms;      /* the stand-alone search of iteration 1 between the two launches of the persistent kernel */
} pwicp_icp_result;

/* source set of the inner loop: host upload (stand-alone use) ... */
int pwicp_icp_source_upload(pwicp_ctx* ctx, const float* src_xyz, int n);
/* ... or "all source centroids uploaded by pwicp_source_upload" */
int pwicp_icp_source_all(pwicp_ctx* ctx);
/* Runs the loop on the device-resident source set against the resident centroid target.
 * Replaces P2PICPwithPatchNormal (src/Registration.cpp:1255-1269) =
 * pcl::IterativeClosestPointWithNormals::align.  T16 = final transformation (row-major f32).
 * Optional traces, each may be NULL: mse_trace[max_iter] (double), T_trace[max_iter*16] (float),
 * idx_trace[max_iter*n] (int, correspondence indices of every inner iteration). */
int pwicp_icp_run(pwicp_ctx* ctx, const pwicp_icp_params* prm, float* T16, pwicp_icp_result* res,
                  double* mse_trace, float* T_trace, int* idx_trace);
/* Per-iteration profile of the last pwicp_icp_run (diagnostic): iter_us[k] = device time of inner iteration k
 * (%globaltimer of CTA 0, microseconds), searched[k] = queries of iteration k that were not answered from their
 * candidate cache and ran the ball search.  cap = entries available in each array (either may be NULL);
 * returns the number of iterations written. */
int pwicp_icp_profile(pwicp_ctx* ctx, double* iter_us, int* searched, int cap);
/* ... and where CTA 0 spent each iteration: phase_us[4 k + j] = microseconds from the start of iteration k to the end of
 * its own batches (j = 0), its CTA sum posted (1), the packets of all CTAs in and the totals formed (2), the 6x6 system solved (3). */
int pwicp_icp_phase_profile(pwicp_ctx* ctx, double* phase_us, int cap);
/* Processing order of the last pwicp_icp_run: perm[k] = index (in the uploaded source set) of the
 * k-th point in the order the device accumulated the normal equations (source points are sorted
 * by the target-grid cell they start in).  Only the order of the double sums depends on it; the
 * bit-exact parity test feeds the oracle the same order.  perm holds n_source ints. */
int pwicp_icp_order(pwicp_ctx* ctx, int* perm);
/* Host-buffer convenience with the call shape of P2PICPwithPatchNormal(target, source, eps):
 * uploads both clouds, builds the grid, runs the loop (this is the path `e2e` times). */
int pwicp_icp_p2plane(pwicp_ctx* ctx, const float* tgt_xyz, const float* tgt_nrm, int n1,
                      const float* src_xyz, int n2, const pwicp_icp_params* prm,
                      float* T16, pwicp_icp_result* res);

/* ---- F3: constants of every planar patch of a cloud in one launch --------------
 * Replaces, per patch: calPatchCTandBP (src/Segmentation.cpp:260-303: centroid + the six boundary
 * points Xmax,Xmin,Ymax,Ymin,Zmax,Zmin), calPatchNormal (src/CommonFunc.cpp:284-333; the reference
 * re-runs it 7*N2+N1+N2 times per outer iteration), calPatchSTD (src/CommonFunc.cpp:336-354) and the
 * CTstd = std / n of calBPandCTSTD (src/Segmentation.cpp:306-321).
 * patch_xyz: points packed patch by patch; patch_off[n_patches+1]: first point of every patch
 * (patch_off[0] = 0).  Outputs (each may be NULL): ct3[n*3], bp18[n*18], nrm3[n*3], nrm_ok[n]
 * (calPatchNormal's verdict: 0 and (0,0,1) for patches of 4 points or fewer), bp_std[n], ct_std[n]. */
int pwicp_patch_stats(pwicp_ctx* ctx, const float* patch_xyz, const int* patch_off, int n_patches,
                      float* ct3, float* bp18, float* nrm3, unsigned char* nrm_ok,
                      float* bp_std, float* ct_std);

/* ---- A2 + A7 + A8: one outer iteration / the outer loop ----------------------------------- */
typedef struct {
    float Res1, Res2, SVRes1, SVRes2, DTmin;   /* src/Registration.cpp:706, :710 */
} pwicp_pair_params;

typedef struct {                 /* replaces the reference's globals and in/out refs */
    float currDT, BBchange_1, BBchange_2;      /* src/Registration.cpp:711 */
    int   toStage2, toStage3;                  /* g_toStage2 / g_toStage3, src/Registration.cpp:11-12 */
} pwicp_state;

typedef struct {
    int   n_stable, n_stable_pts;
    int   icp_iters, icp_state;
    float LoDet_min, LoDet_max, maxBBchange;
    double P75;                  /* NaN when stage 1 did not run */
    double bb6[6];
    int   vcm_written, vcm_singular;
    float device_ms;
} pwicp_iter_stats;

/* PwICP_singleIteration (src/Registration.cpp:704-972) on the resident pair.  T16 = transMatICP,
 * vcm36 written when stage 3 is reached in this call (calTransParaVCM, :1273-1343).
 * stable_flags (n2 bytes) may be NULL. */
int pwicp_single_iteration(pwicp_ctx* ctx, const pwicp_pair_params* pp, pwicp_state* st,
                           const pwicp_icp_params* icp, float* T16, double* vcm36,
                           unsigned char* stable_flags, pwicp_iter_stats* stats);

/* Piecewise_ICP (src/Registration.cpp:618-700) from the centroid-level boundary on: DTinit
 * (manual, or 3*P75(cloud1, cloud2) when is_manual_dtinit == 0), the while(!stage3) loop,
 * transMat = cur * transMat, DTseries (max_outer+1 floats).  Returns the iteration count in
 * *n_outer.  stats_per_iter (max_outer entries) may be NULL. */
int pwicp_piecewise_icp(pwicp_ctx* ctx, const pwicp_pair_params* pp, int is_manual_dtinit,
                        float DTinit, const pwicp_icp_params* icp, int max_outer,
                        float* DTseries, int* n_series, float* T16, double* vcm36,
                        int* n_outer, pwicp_iter_stats* stats_per_iter);

/* ---- stand-alone pieces (parity tests, F1 consumers) --------------------------------------- */
/* calPercentileDistBetween2PC (src/CommonFunc.cpp:266-281): host clouds in, value out. */
int pwicp_percentile_nn(pwicp_ctx* ctx, const float* cloud1, int m1, const float* cloud2, int m2,
                        float percentile, double* out);
/* calOverlapRatioByC2Cdist (src/Registration.cpp:593-614) */
int pwicp_overlap_ratio(pwicp_ctx* ctx, const float* cloud1, int m1, const float* cloud2, int m2,
                        float DTinit, float* out);
/* squared distance of every point to its nearest OTHER point of the same cloud: the second
 * neighbour of KdTreeFLANN::nearestKSearch(i, 2) in calPCresolution (src/CommonFunc.cpp:239-263) */
int pwicp_self_nn(pwicp_ctx* ctx, const float* xyz, int n, float* d2);
/* ---- F4: PCpreprocessing (src/CommonFunc.cpp:423-452), [PCL 1.8.1 VoxelGrid / StatisticalOutlierRemoval] ---- */
/* pcl::VoxelGrid with a cubic leaf (:430-433): one centroid per occupied voxel, ascending voxel index; points of a
 * voxel are summed in input order (float).  out_xyz must hold n points. */
int pwicp_voxel_grid(pwicp_ctx* ctx, const float* xyz, int n, float leaf, float* out_xyz, int* n_out);
/* first pass of pcl::StatisticalOutlierRemoval (:446-451): mean distance of every point to its k nearest other
 * points (nearestKSearch(point, k + 1) without the point itself), float(dist_sum / k).  1 <= k <= 32 < n. */
int pwicp_knn_mean_dist(pwicp_ctx* ctx, const float* xyz, int n, int k, float* mean_dist);
/* Front end of the supervoxel segmentation (src/Segmentation.cpp:28-46: kdtree.FindKNearestNeighbors(points[i], kNN = 45) and
 * PCAEstimateNormal over the neighbours, per point): neighbors[n * k] = indices of the k nearest points of every point (the
 * point itself first; squared distance in double, ties by index), normals[n * 3] = unit eigenvector of the smallest eigenvalue
 * of the neighbours' covariance (double; orientation undefined).  Either output may be NULL.  1 <= k <= 64 <= n.  These are
 * the inputs the reference's SupervoxelSegmentation takes; the merge itself stays with the caller (segmenter plug-in). */
int pwicp_knn_normals(pwicp_ctx* ctx, const float* xyz, int n, int k, int* neighbors, double* normals);
/* PCpreprocessing(cloud_in, cloud_out, isDownSamp, voxelSize, SOR_NeighborNum, SOR_StdMult) (:423-439): VoxelGrid
 * (when downsample != 0) then StatisticalOutlierRemoval; points with mean distance <= mean + std_mult * stddev are
 * kept in order.  out_xyz must hold n points. */
int pwicp_preprocess(pwicp_ctx* ctx, const float* xyz, int n, int downsample, float leaf, int k, double std_mult,
                     float* out_xyz, int* n_out);
/* calTransParaVCM (src/Registration.cpp:1273-1343) on the resident centroid target */
int pwicp_vcm(pwicp_ctx* ctx, const float* src_stable_xyz, int n, double* vcm36, int* singular);
/* pcl::transformPointCloud (src/Registration.cpp:943-954), in place on a host array */
int pwicp_transform(pwicp_ctx* ctx, float* xyz, int n, const float* T16);
/* octree bounding cube (src/Registration.cpp:881-886) of a host cloud */
int pwicp_octree_bbox(pwicp_ctx* ctx, const float* xyz, int n, double res, double* bb6);

/* pure host helpers of the path (no device needed) */
float pwicp_bbox_corner_change(const double* bb6, const float* T16);  /* src/CommonFunc.cpp:410-419 */
void  pwicp_matrix2angle(const float* T16, float* ang3);              /* src/CommonFunc.cpp:385-407 */
void  pwicp_mat4_mul(const float* A, const float* B, float* C);       /* Eigen Matrix4f product, :687 */

#ifdef __cplusplus
}
#endif
#endif
