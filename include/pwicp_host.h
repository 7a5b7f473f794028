/*
 * pwicp_host.h -- C ABI of libpwicp_host.so: the reference's two public entry points and the
 * epoch-sharded form of the 4D loop.
 *
 * The reference exports exactly two undecorated symbols from its DLL, bound with ctypes by
 * python/main.py:12-18:
 *     bool PiecewiseICP_pair_call(const char* confile, const char* outfile)     include/Registration.h:49
 *     bool PiecewiseICP_4D_call(const char* confile, int startEpoch, int epochNum,
 *                               int pairMode, float overlapThd)                 include/Registration.h:36
 * Both are kept with identical types.  The epoch loop of the 4D entry (src/Registration.cpp:89-187)
 * is embarrassingly parallel: PiecewiseICP_4D_shard runs the pairs of one rank and
 * fills fixed-size records, the caller gathers the records of all ranks (torch.distributed /
 * NCCL all-gather in bench.py and tests), PiecewiseICP_4D_finalize writes TransMatrices.txt /
 * TransParameters.txt and chains to the reference epoch (src/Registration.cpp:977-1153).
 */
#ifndef PWICP_HOST_H
#define PWICP_HOST_H
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

bool PiecewiseICP_pair_call(const char* confile, const char* outfile);
#ifdef __cplusplus
bool PiecewiseICP_4D_call(const char* confile, int startEpoch, int epochNum, int pairMode, float overlapThd = 0.75f);
#else
bool PiecewiseICP_4D_call(const char* confile, int startEpoch, int epochNum, int pairMode, float overlapThd);
#endif

/* one registered epoch pair: what src/Registration.cpp:151-180 writes per epoch */
typedef struct {
    int    step;          /* 1-based index of the pair in the epoch loop (i - startEpoch + 1) */
    int    status;        /* 1 = registered, 0 = not this rank's / failed */
    long long time_stamp; /* scanTimeList[i + 1] */
    float  T[16];         /* final transformation, row-major */
    float  para[6];       /* Rx Ry Rz [gon], tx ty tz [m] */
    double VCM[36];
    float  seconds;       /* wall time of the pair */
    int    pad;
} pwicp_epoch_record;     /* 400 bytes */

/* Runs the pairs with (step - 1) % world == rank on `device`.  records must hold epochNum entries;
 * entry (step - 1) is filled for this rank's pairs, the others get status 0.  Returns the number of
 * pairs this rank registered, or -1 on a configuration / file error. */
int  PiecewiseICP_4D_shard(const char* confile, int startEpoch, int epochNum, int pairMode, float overlapThd,
                           int rank, int world, int device, pwicp_epoch_record* records);
/* Writes the result files from the gathered records (all ranks' entries merged) and chains to
 * the reference epoch; call on rank 0 only. */
bool PiecewiseICP_4D_finalize(const char* confile, int startEpoch, int epochNum, int pairMode,
                              const pwicp_epoch_record* records);
/* Segmenter plug-in.  The supervoxel segmentation in front of the hot path (src/Segmentation.cpp:17-66: kNN-45 PCA
 * normals + Lin's supervoxels, all in the reference's header-only codelibrary) is out of scope of this library, which
 * ships a documented stand-in (cubic cells).  A caller that owns a segmentation registers it here: it receives the packed
 * xyz of a pre-processed cloud, the supervoxel resolution and kNN (include/CommonFunc.h:41), writes one label per point
 * (the reference's lin_labels) and returns the number of supervoxels (< 0 = failure).  PatchGenerationAndRefinement then
 * groups the points as src/Segmentation.cpp:95-100 does and applies the reference's refinement / planarity gates.
 * NULL restores the stand-in.  Not thread-safe (like the reference's globals). */
typedef int (*pwicp_segmenter_fn)(const float* xyz, int n, float svResolution, int knn, int* labels);
void pwicp_host_set_segmenter(pwicp_segmenter_fn fn);
/* A segmentation can also be named by the environment, for the file-level drivers: PWICP_SEGMENTER_PLUGIN=<shared object>:<symbol>
 * is dlopen()ed on first use and the symbol (of type pwicp_segmenter_fn) registered.  The library itself ships no
 * supervoxel implementation (out of scope); without a plug-in the drivers use the stand-in and say so. */

/* Result-file tools (no device needed).  pwicp_host_write_transmatrix: the per-pair file of src/Registration.cpp:340-388
 * (TransMatrix.txt / <time>_<mode>_TransMatrix.txt) from a row-major 4x4 and a 6x6; returns 1 on success.
 * pwicp_host_chain_to_reference: calTransToReferenceEpoch (:977-1153) file to file.  pwicp_host_abs_error:
 * calAbsErrorOfTransPara (:1157-1251) file to file. */
int  pwicp_host_write_transmatrix(const char* path, const float* T16, const double* vcm36);
void pwicp_host_chain_to_reference(const char* transMatFile, int pairMode, const char* pairFile, int epochNum,
                                   const char* outTM, const char* outTP);
void pwicp_host_abs_error(const char* transMatFile, const char* gtFile, int allEpochNum, int startEpoch, const char* outFile);

/* device used by the reference-shaped entry points (default: PWICP_DEVICE, LOCAL_RANK or 0) */
void pwicp_host_set_device(int device);

#ifdef __cplusplus
}
#endif
#endif
