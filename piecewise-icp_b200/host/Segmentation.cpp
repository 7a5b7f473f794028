// Segmentation.cpp -- input side of the hot path (OUT OF SCOPE; see Segmentation.h).
// Initial patches: one per occupied cubic cell of side svResolution (documented stand-in for the
// reference's supervoxel segmentation).  Everything after that follows the reference's
// per-patch post-processing: src/Segmentation.cpp:107-150, :195-321.
#include "Segmentation.h"
#include <dlfcn.h>
#include "../../include/pwicp_host.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

using namespace std;

void pwicpJacobi3(double A[3][3], double w[3], double V[3][3]);   // CommonFunc.cpp

namespace {

// centroid (double) and centred scatter matrix / n
void scatter(const pcl::PointCloud<pcl::PointXYZ>& c, double mean[3], double M[3][3]) {
    mean[0] = mean[1] = mean[2] = 0;
    for (const auto& p : c.points) { mean[0] += p.x; mean[1] += p.y; mean[2] += p.z; }
    for (int k = 0; k < 3; ++k) mean[k] /= (double)c.size();
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) M[r][q] = 0;
    for (const auto& p : c.points) {
        const double d[3] = {p.x - mean[0], p.y - mean[1], p.z - mean[2]};
        for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) M[r][q] += d[r] * d[q];
    }
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) M[r][q] /= (double)c.size();
}

}  // namespace

// src/Segmentation.cpp:195-228: drop points farther than sigmaMul * RMS from the PCA plane
int PatchRefinement(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, pcl::PointCloud<pcl::PointXYZ>::Ptr refinedPatch, double sigmaMul) {
    const int n = (int)cloud->size();
    refinedPatch->clear();
    double mean[3], M[3][3], w[3], V[3][3];
    scatter(*cloud, mean, M);
    pwicpJacobi3(M, w, V);
    const float A = (float)V[0][0], B = (float)V[1][0], C = (float)V[2][0];
    const float D = -(A * (float)mean[0] + B * (float)mean[1] + C * (float)mean[2]);
    const double nrm = std::sqrt((double)(A * A + B * B + C * C));
    vector<double> dist(n);
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const auto& p = cloud->points[i];
        dist[i] = std::fabs(A * p.x + B * p.y + C * p.z + D) / nrm;
        s += dist[i] * dist[i];
    }
    const double stdDist = std::sqrt(s / double(n));
    for (int j = 0; j < n; ++j)
        if (std::fabs(dist[j]) < std::fabs(sigmaMul * stdDist)) refinedPatch->push_back(cloud->points[j]);
    return (int)refinedPatch->size();
}

// src/Segmentation.cpp:231-257: singular values E1 >= E2 >= E3 of the covariance
void calPatchFeature(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, float& variation, float& planarity, float& linearity) {
    double mean[3], M[3][3], w[3], V[3][3];
    scatter(*cloud, mean, M);
    pwicpJacobi3(M, w, V);
    const float E1 = (float)std::max(w[2], 0.0), E2 = (float)std::max(w[1], 0.0), E3 = (float)std::max(w[0], 0.0);
    variation = E3 / (E1 + E2 + E3);
    planarity = (E2 - E3) / E1;
    linearity = (E1 - E2) / E1;
}

// src/Segmentation.cpp:260-303: centroid + the six axis-extremal points (Xmax,Xmin,Ymax,Ymin,Zmax,Zmin)
int calPatchCTandBP(pcl::PointCloud<pcl::PointXYZ> cloud, pcl::PointXYZ& centroid, pcl::PointCloud<pcl::PointXYZ>::Ptr cloudBP) {
    cloudBP->clear();
    Eigen::Vector4f ct;
    pcl::compute3DCentroid(cloud, ct);
    centroid.x = ct[0]; centroid.y = ct[1]; centroid.z = ct[2];
    pcl::PointXYZ ext[6];
    ext[0] = pcl::PointXYZ(-FLT_MAX, 0, 0); ext[1] = pcl::PointXYZ(FLT_MAX, 0, 0);
    ext[2] = pcl::PointXYZ(0, -FLT_MAX, 0); ext[3] = pcl::PointXYZ(0, FLT_MAX, 0);
    ext[4] = pcl::PointXYZ(0, 0, -FLT_MAX); ext[5] = pcl::PointXYZ(0, 0, FLT_MAX);
    for (const auto& p : cloud.points) {
        if (p.x > ext[0].x) ext[0] = p;
        if (p.x < ext[1].x) ext[1] = p;
        if (p.y > ext[2].y) ext[2] = p;
        if (p.y < ext[3].y) ext[3] = p;
        if (p.z > ext[4].z) ext[4] = p;
        if (p.z < ext[5].z) ext[5] = p;
    }
    for (int k = 0; k < 6; ++k) cloudBP->push_back(ext[k]);
    return (int)cloudBP->size();
}

// src/Segmentation.cpp:306-321
void calBPandCTSTD(pcl::PointCloud<pcl::PointXYZ>* cloudPatches, int patchNum, std::vector<float>& stdBP, std::vector<float>& stdCT) {
    stdBP.clear(); stdCT.clear();
    for (int i = 0; i < patchNum; ++i) {
        const int pointNum = (int)cloudPatches[i].size();
        const float patchStd = calPatchSTD(cloudPatches[i].makeShared());
        stdBP.push_back(patchStd);
        const float Neffe = float(pointNum);
        stdCT.push_back(patchStd / Neffe);
    }
}

// ---- segmenter plug-in (include/pwicp_host.h: pwicp_host_set_segmenter) ----------------------------------------
// Segmentation is out of scope here; a caller that owns one (the reference does: Lin's supervoxels,
// src/Segmentation.cpp:17-66) registers it and gets the reference's grouping (:95-100: points appended to their
// supervoxel in cloud order, supervoxels visited in label order) followed by the same per-patch post-processing.
static pwicp_segmenter_fn g_segmenter = nullptr;
extern "C" void pwicp_host_set_segmenter(pwicp_segmenter_fn fn) { g_segmenter = fn; }

// PWICP_SEGMENTER_PLUGIN=<shared object>:<symbol>: a segmentation named by the environment (include/pwicp_host.h)
static void loadSegmenterPlugin() {
    static bool tried = false;
    if (tried) return;
    tried = true;
    const char* e = getenv("PWICP_SEGMENTER_PLUGIN");
    if (!e || !*e) {
        cout << "--->>> no segmenter registered: cubic-cell stand-in instead of the reference's supervoxels "
                "(pwicp_host_set_segmenter / PWICP_SEGMENTER_PLUGIN)" << endl;
        return;
    }
    const std::string spec(e);
    const size_t colon = spec.rfind(':');
    if (colon == std::string::npos) { cerr << "PWICP_SEGMENTER_PLUGIN: expected <shared object>:<symbol>" << endl; return; }
    void* h = dlopen(spec.substr(0, colon).c_str(), RTLD_NOW | RTLD_LOCAL);
    void* f = h ? dlsym(h, spec.substr(colon + 1).c_str()) : nullptr;
    if (!f) { cerr << "PWICP_SEGMENTER_PLUGIN: cannot load " << spec << ": " << dlerror() << endl; return; }
    g_segmenter = reinterpret_cast<pwicp_segmenter_fn>(f);
}

namespace {

// src/Segmentation.cpp:107-150 for one initial patch; returns true when the patch is kept
bool acceptPatch(pcl::PointCloud<pcl::PointXYZ>::Ptr raw, pcl::PointCloud<pcl::PointXYZ>::Ptr refined,
                 pcl::PointCloud<pcl::PointXYZ>::Ptr cloudCentroid, pcl::PointCloud<pcl::PointXYZ>::Ptr cloudBoundary,
                 pcl::PointCloud<pcl::PointXYZ>& slot) {
    if ((int)raw->size() < minPtNum) return false;                                        // :109-112
    const int kept = PatchRefinement(raw, refined, 2.0);                                  // :116
    if (kept < minPtNum) return false;                                                    // :119-122
    float variation, planarity, linearity;
    calPatchFeature(refined, variation, planarity, linearity);
    if (variation > 0.02f || planarity < 0.25f) return false;                             // :127
    slot = *refined;
    pcl::PointXYZ centroid;
    pcl::PointCloud<pcl::PointXYZ>::Ptr bp(new pcl::PointCloud<pcl::PointXYZ>);
    if (calPatchCTandBP(*refined, centroid, bp) != 6) {
        std::cerr << "Error: Incorrect number of boundary points calculated! Aborting.\n";
        std::exit(EXIT_FAILURE);
    }
    cloudCentroid->push_back(centroid);
    *cloudBoundary += *bp;
    return true;
}

int patchesFromSegmenter(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, float svResolution,
                         pcl::PointCloud<pcl::PointXYZ>::Ptr cloudCentroid, pcl::PointCloud<pcl::PointXYZ>::Ptr cloudBoundary,
                         pcl::PointCloud<pcl::PointXYZ>*& cloudPatches) {
    const int n = (int)cloud->size();
    vector<float> xyz(3 * (size_t)n);
    for (int i = 0; i < n; ++i) { xyz[3 * i] = cloud->points[i].x; xyz[3 * i + 1] = cloud->points[i].y; xyz[3 * i + 2] = cloud->points[i].z; }
    vector<int> labels(n, -1);
    const int numSV = g_segmenter(xyz.data(), n, svResolution, kNN, labels.data());
    if (numSV < 0) { std::cerr << "Error: the registered segmenter failed! Aborting.\n"; std::exit(EXIT_FAILURE); }
    cout << "--->>> " << numSV << " supervoxels are generated." << endl;
    vector<pcl::PointCloud<pcl::PointXYZ>> all(std::max(numSV, 1));
    for (int i = 0; i < n; ++i)
        if (labels[i] >= 0 && labels[i] < numSV) all[labels[i]].push_back(cloud->points[i]);                       // :95-100
    cloudPatches = new pcl::PointCloud<pcl::PointXYZ>[std::max(numSV, 1)];
    int validSV = 0, validSVPtNum = 0;
    pcl::PointCloud<pcl::PointXYZ>::Ptr refined(new pcl::PointCloud<pcl::PointXYZ>);
    for (int i = 0; i < numSV; ++i) {
        if (!acceptPatch(all[i].makeShared(), refined, cloudCentroid, cloudBoundary, cloudPatches[validSV])) continue;
        validSVPtNum += (int)cloudPatches[validSV].size();
        ++validSV;
    }
    cout << "--->>> Number of selected patches = " << validSV << "   Ratio of selected patches = "
         << 100.0 * validSV / std::max(numSV, 1) << "% \n"
         << "--->>> Number of points in selected patches = " << validSVPtNum << "   Ratio of selected points = "
         << 100.0 * validSVPtNum / std::max(n, 1) << "% \n\n";
    return validSV;
}

}  // namespace

int PatchGenerationAndRefinement(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud, float svResolution,
                                 pcl::PointCloud<pcl::PointXYZ>::Ptr cloudCentroid,
                                 pcl::PointCloud<pcl::PointXYZ>::Ptr cloudBoundary,
                                 pcl::PointCloud<pcl::PointXYZ>*& cloudPatches, bool /*isVis*/) {
    cloudCentroid->clear(); cloudBoundary->clear();
    const int n = (int)cloud->size();
    if (!g_segmenter) loadSegmenterPlugin();
    if (g_segmenter) return patchesFromSegmenter(cloud, svResolution, cloudCentroid, cloudBoundary, cloudPatches);
    // stand-in segmentation: sort points by cubic cell of side svResolution
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (const auto& p : cloud->points) { mn[0] = min(mn[0], p.x); mn[1] = min(mn[1], p.y); mn[2] = min(mn[2], p.z); }
    vector<pair<unsigned long long, int>> keyed(n);
    for (int i = 0; i < n; ++i) {
        const auto& p = cloud->points[i];
        const unsigned long long ix = (unsigned long long)floor((p.x - mn[0]) / svResolution);
        const unsigned long long iy = (unsigned long long)floor((p.y - mn[1]) / svResolution);
        const unsigned long long iz = (unsigned long long)floor((p.z - mn[2]) / svResolution);
        keyed[i] = {(iz << 42) | (iy << 21) | ix, i};
    }
    sort(keyed.begin(), keyed.end());
    int numSV = 0;
    for (int i = 0; i < n; ++i) if (i == 0 || keyed[i].first != keyed[i - 1].first) ++numSV;
    cout << "--->>> " << numSV << " supervoxels are generated." << endl;

    cloudPatches = new pcl::PointCloud<pcl::PointXYZ>[std::max(numSV, 1)];
    int validSV = 0, invalidSV = 0, validSVPtNum = 0;
    pcl::PointCloud<pcl::PointXYZ>::Ptr raw(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::PointCloud<pcl::PointXYZ>::Ptr refined(new pcl::PointCloud<pcl::PointXYZ>);
    int i = 0;
    while (i < n) {
        int j = i;
        raw->clear();
        while (j < n && keyed[j].first == keyed[i].first) { raw->push_back(cloud->points[keyed[j].second]); ++j; }
        i = j;
        if (!acceptPatch(raw, refined, cloudCentroid, cloudBoundary, cloudPatches[validSV])) { ++invalidSV; continue; }
        validSVPtNum += (int)cloudPatches[validSV].size();
        ++validSV;
    }
    cout << "--->>> Number of selected patches = " << validSV << "   Ratio of selected patches = "
         << 100.0 * validSV / std::max(numSV, 1) << "% \n"
         << "--->>> Number of points in selected patches = " << validSVPtNum << "   Ratio of selected points = "
         << 100.0 * validSVPtNum / std::max(n, 1) << "% \n\n";
    return validSV;
}
