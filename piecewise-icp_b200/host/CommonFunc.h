// CommonFunc.h -- mirror of the reference's include/CommonFunc.h (same names, argument meaning and
// error behaviour) on top of libpwicp.so.  Functions that reach PCL in the reference are served by
// the CUDA library; the rest is plain host C++.  Citations: /root/reference paths.
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "pcl_shim.h"

const double ARC_TO_DEG = 57.29577951308238;   ///< include/CommonFunc.h:37
const double DEG_TO_ARC = 0.0174532925199433;
const double GON_TO_ARC = 0.0157079632679;
const double ARC_TO_GON = 63.6619772368;       ///< include/CommonFunc.h:40
const int kNN = 45;                             ///< include/CommonFunc.h:41 (passed to a registered segmenter; unused by the stand-in segmentation)
const int minPtNum = 20;                        ///< include/CommonFunc.h:42

/// include/CommonFunc.h:48-61
struct ConfigPara {
    std::string FolderFilePath1;
    std::string FolderFilePath2;
    bool isSetResSVsize = false;
    float PCres1 = 0, PCres2 = 0;
    float SVsize1 = 0, SVsize2 = 0;
    bool isSetDTinit = false;
    float DTinit = 0;
    float DTmin = 0;
    bool isVisual = false;
};

/// 11-line positional config parser, src/CommonFunc.cpp:11-136 (CR of CRLF files is stripped).
bool readConfigFile(std::string conFile, ConfigPara& confPara);
/// src/CommonFunc.cpp:182-208 (POSIX directory walk instead of _findfirst)
int extractAllFilesFromFolder(std::string folderPath, std::vector<std::string>& fileNameList, std::vector<long>& fileTimeList);
void getFiles(std::string folderpath, std::vector<std::string>& files);
long extractTimeFromFileName(std::string fileName, const std::string substring, int timeLength);
/// src/CommonFunc.cpp:239-263: mean distance to the nearest other point (device: pwicp_mean_nn_spacing)
float calPCresolution(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud);
/// src/CommonFunc.cpp:266-281 (device: pwicp_percentile_nn)
double calPercentileDistBetween2PC(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2, float percentile);
/// src/CommonFunc.cpp:284-333 (pcl::computePointNormal + eigen33 restated on the host, float)
bool calPatchNormal(pcl::PointCloud<pcl::PointXYZ> cloud, float& nx, float& ny, float& nz);
/// src/CommonFunc.cpp:336-354
float calPatchSTD(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud);
/// src/CommonFunc.cpp:357-382
void generateCentroidCloudWithPatchNormals(pcl::PointCloud<pcl::PointXYZ>::Ptr cloudCentroids,
                                           pcl::PointCloud<pcl::PointXYZ>* cloudPatch,
                                           pcl::PointCloud<pcl::PointNormal>::Ptr cloudCentroids_normals);
/// src/CommonFunc.cpp:385-407
void matrix2angle(Eigen::Matrix4f transMat, Eigen::Vector3f& rotAngle);
/// src/CommonFunc.cpp:410-419
float calBoundingBoxCornerChange(const double* boundingBox, const Eigen::Matrix4f transMat);
/// src/CommonFunc.cpp:423-452.  OUT OF SCOPE stand-ins: plain host voxel-grid
/// centroiding + statistical outlier removal with PCL's documented semantics, not parity-checked.
/// host-only statements of the same two filters (no device; used by c_hooks.cpp tools)
void PCpreprocessingHost(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
                         bool isDownSamp, float voxelSize, int SOR_NeighborNum, double SOR_StdMult);
void SORfilterHost(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
                   int SOR_NeighborNum, double SOR_StdMult);
void PCpreprocessing(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
                     bool isDownSamp, float voxelSize, int SOR_NeighborNum, double SOR_StdMult);
void SORfilter(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
               int SOR_NeighborNum, double SOR_StdMult);
/// GUI of the reference (src/CommonFunc.cpp:456-493): no-ops here, the isVisual key is accepted.
void visualizeTwoPC(pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int,
                    pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int, double, double, double);
void visualizeThreePC(pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int,
                      pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int,
                      pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int, double, double, double);

// ---- device context of the mirror (the reference API has no device argument) -----------------
struct pwicp_ctx;
/// process-wide context, created on first use on device PWICP_DEVICE / LOCAL_RANK / 0
pwicp_ctx* pwicpHostContext();
void pwicpHostSetDevice(int device);
