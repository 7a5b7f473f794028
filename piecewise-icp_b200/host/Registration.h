// Registration.h -- mirror of the reference's include/Registration.h: same names, argument
// meaning and error behaviour; the arithmetic of the inner registration loop runs in libpwicp.so.
#pragma once
#include "CommonFunc.h"
#include "Segmentation.h"

// The two public entry points (include/Registration.h:36, :49; src/Registration.cpp:17-215,
// :219-398) and the epoch-sharded form of the 4D loop are declared, with C linkage,
// in include/pwicp_host.h:
//   bool PiecewiseICP_4D_call(const char* confile, int startEpoch, int epochNum, int pairMode, float overlapThd = 0.75f);
//   bool PiecewiseICP_pair_call(const char* confile, const char* outfile);
#include "../../include/pwicp_host.h"

/// include/Registration.h:74-78 (src/Registration.cpp:402-548)
bool Piecewise_ICP_4D(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2,
                      bool isSetResSVsize, float Res1, float Res2, float SVsize1, float SVsize2,
                      bool isManualDTinit, float DTinit, float DTmin, std::string outfileIdx,
                      Eigen::Matrix4f& transMat, std::vector<float>& transPara, Eigen::MatrixXd& VCM);
/// include/Registration.h:93-94 (src/Registration.cpp:552-589)
bool calAdaptivePairSequence(std::vector<std::string> fileNameList, int startEpoch, float DTinit, float ratioThd,
                             std::map<int, int>& RegPairs, std::string adaptivePairFile);
/// include/Registration.h:106-107 (src/Registration.cpp:593-614; device: pwicp_overlap_ratio)
float calOverlapRatioByC2Cdist(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2, float DTinit);
/// include/Registration.h:127-129 (src/Registration.cpp:977-1153)
void calTransToReferenceEpoch(std::string transMatFile, int pairMode, std::string adaptivePairFile, int epochNum,
                              std::string transMat2RefFile, std::string transPara2RefFile, std::vector<int>& timeStamp,
                              std::vector<Eigen::Matrix4f>& allTransMat2Ref, std::vector<Eigen::MatrixXd>& allVCM2Ref);
/// include/Registration.h:149-153 (src/Registration.cpp:618-700; device: pwicp_piecewise_icp)
void Piecewise_ICP(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2,
                   bool isSetResSVsize, float Res1, float Res2, float SVsize1, float SVsize2,
                   bool isManualDTinit, float DTinit, float DTmin,
                   std::vector<float>& DTseries, Eigen::Matrix4f& transMat, Eigen::MatrixXd& VCM);
/// include/Registration.h:181-188 (src/Registration.cpp:704-972; device: pwicp_single_iteration)
Eigen::Matrix4f PwICP_singleIteration(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2,
                                      float Res1, float Res2, float SVRes1, float SVRes2,
                                      pcl::PointCloud<pcl::PointXYZ>*& SVcloud1, pcl::PointCloud<pcl::PointXYZ>*& SVcloud2,
                                      pcl::PointCloud<pcl::PointXYZ>::Ptr CTcloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr CTcloud2,
                                      pcl::PointCloud<pcl::PointXYZ>::Ptr BPcloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr BPcloud2,
                                      std::vector<float> CTstd1, std::vector<float> BPstd2, float DTmin,
                                      float& currDT, float& BBchange_1, float& BBchange_2, Eigen::MatrixXd& VCM);
/// include/Registration.h:200-201 (src/Registration.cpp:1157-1251)
void calAbsErrorOfTransPara(std::string transMatFile, std::string GTtransMatFile, int allEpochNum, int startEpoch,
                            std::string transParaErrorFile);
/// include/Registration.h:213-214 (src/Registration.cpp:1255-1269; device: pwicp_icp_p2plane)
Eigen::Matrix4f P2PICPwithPatchNormal(pcl::PointCloud<pcl::PointNormal>::Ptr cloudTarget,
                                      pcl::PointCloud<pcl::PointNormal>::Ptr cloudSource, double EucldEpsilon);
/// include/Registration.h:227-229 (src/Registration.cpp:1273-1343; device: pwicp_vcm)
Eigen::MatrixXd calTransParaVCM(pcl::PointCloud<pcl::PointXYZ>::Ptr cloudTarget,
                                pcl::PointCloud<pcl::PointNormal>::Ptr cloudTargetwithNormals,
                                pcl::PointCloud<pcl::PointXYZ>::Ptr cloudSourceStable);
