// Segmentation.h -- input side of the hot path.  OUT OF SCOPE: the
// reference segments with Lin-2018 supervoxels over kNN-45 PCA normals (src/Segmentation.cpp:11-192,
// codelibrary/).  This mirror keeps the interface and the per-patch post-processing of the
// reference (2-sigma refinement :195-228, planarity gates :231-257 / :127, centroid + 6 boundary
// points :260-303, sigma per patch :306-321) but generates the initial patches with a documented
// stand-in: one patch per occupied cubic cell of side svResolution (SURVEY.md 8(d), C1 fixture).
#pragma once
#include "CommonFunc.h"

int PatchGenerationAndRefinement(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, float svResolution,
                                 pcl::PointCloud<pcl::PointXYZ>::Ptr cloudCentroid,
                                 pcl::PointCloud<pcl::PointXYZ>::Ptr cloudBoundary,
                                 pcl::PointCloud<pcl::PointXYZ>*& cloudPatches, bool isVis);
int PatchRefinement(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, pcl::PointCloud<pcl::PointXYZ>::Ptr refinedPatch, double sigmaMul);
void calPatchFeature(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, float& variation, float& planarity, float& linearity);
int calPatchCTandBP(pcl::PointCloud<pcl::PointXYZ> cloud, pcl::PointXYZ& centroid, pcl::PointCloud<pcl::PointXYZ>::Ptr cloudBP);
void calBPandCTSTD(pcl::PointCloud<pcl::PointXYZ>* cloudPatches, int patchNum, std::vector<float>& stdBP, std::vector<float>& stdCT);
