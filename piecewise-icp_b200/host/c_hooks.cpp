// c_hooks.cpp -- small C-linkage views of host-side mirror functions (used by the Python binding
// and by the CPU tests; none of them needs a CUDA device).
#include <cstring>

#include "Registration.h"

extern "C" {

struct pwicp_config_c {
    char path1[1024], path2[1024];
    int isSetResSVsize; float PCres1, PCres2, SVsize1, SVsize2;
    int isSetDTinit; float DTinit, DTmin; int isVisual;
};

int pwicp_host_read_config(const char* path, pwicp_config_c* out) {
    ConfigPara c;
    if (!readConfigFile(path, c)) return 0;
    std::memset(out, 0, sizeof(*out));
    std::strncpy(out->path1, c.FolderFilePath1.c_str(), sizeof(out->path1) - 1);
    std::strncpy(out->path2, c.FolderFilePath2.c_str(), sizeof(out->path2) - 1);
    out->isSetResSVsize = c.isSetResSVsize; out->PCres1 = c.PCres1; out->PCres2 = c.PCres2;
    out->SVsize1 = c.SVsize1; out->SVsize2 = c.SVsize2; out->isSetDTinit = c.isSetDTinit;
    out->DTinit = c.DTinit; out->DTmin = c.DTmin; out->isVisual = c.isVisual;
    return 1;
}

int pwicp_host_patch_normal(const float* xyz, int n, float* n3) {
    pcl::PointCloud<pcl::PointXYZ> c;
    c.resize(n);
    for (int i = 0; i < n; ++i) { c.points[i].x = xyz[3 * i]; c.points[i].y = xyz[3 * i + 1]; c.points[i].z = xyz[3 * i + 2]; }
    return calPatchNormal(c, n3[0], n3[1], n3[2]) ? 1 : 0;
}

// number of points, or -1; xyz may be NULL to query the size only
int pwicp_host_load_pcd(const char* path, float* xyz, int cap) {
    pcl::PointCloud<pcl::PointXYZ> c;
    if (pcl::io::loadPCDFile(path, c) != 0) return -1;
    if (xyz) for (int i = 0; i < (int)c.size() && i < cap; ++i) { xyz[3 * i] = c.points[i].x; xyz[3 * i + 1] = c.points[i].y; xyz[3 * i + 2] = c.points[i].z; }
    return (int)c.size();
}

int pwicp_host_save_pcd(const char* path, const float* xyz, int n) {
    pcl::PointCloud<pcl::PointXYZ> c;
    c.resize(n);
    for (int i = 0; i < n; ++i) { c.points[i].x = xyz[3 * i]; c.points[i].y = xyz[3 * i + 1]; c.points[i].z = xyz[3 * i + 2]; }
    return pcl::io::savePCDFileBinary(path, c);
}

// time stamps of the listed scan files in processing order; returns the count
int pwicp_host_list_epochs(const char* folder, long* times, int cap) {
    std::vector<std::string> files; std::vector<long> t;
    const int n = extractAllFilesFromFolder(folder, files, t);
    for (int i = 0; i < n && i < cap; ++i) times[i] = t[i];
    return n;
}

// patch generation stand-in + per-patch statistics on a host cloud: returns the number of patches,
// fills centroids (3 floats each), boundary points (18), sigmas (BPstd, CTstd), up to cap patches
int pwicp_host_patches(const float* xyz, int n, float svRes, float* ct, float* bp, float* bpstd, float* ctstd, int cap) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr c(new pcl::PointCloud<pcl::PointXYZ>), CT(new pcl::PointCloud<pcl::PointXYZ>), BP(new pcl::PointCloud<pcl::PointXYZ>);
    c->resize(n);
    for (int i = 0; i < n; ++i) { c->points[i].x = xyz[3 * i]; c->points[i].y = xyz[3 * i + 1]; c->points[i].z = xyz[3 * i + 2]; }
    pcl::PointCloud<pcl::PointXYZ>* patches = nullptr;
    const int np = PatchGenerationAndRefinement(c, svRes, CT, BP, patches, false);
    std::vector<float> sb, sc;
    calBPandCTSTD(patches, np, sb, sc);
    for (int i = 0; i < np && i < cap; ++i) {
        ct[3 * i] = CT->points[i].x; ct[3 * i + 1] = CT->points[i].y; ct[3 * i + 2] = CT->points[i].z;
        for (int k = 0; k < 6; ++k) { bp[18 * i + 3 * k] = BP->points[6 * i + k].x; bp[18 * i + 3 * k + 1] = BP->points[6 * i + k].y; bp[18 * i + 3 * k + 2] = BP->points[6 * i + k].z; }
        bpstd[i] = sb[i]; ctstd[i] = sc[i];
    }
    delete[] patches;
    return np;
}

// Input side of one 4D pair on the host, no device needed: PCpreprocessing of both clouds (src/Registration.cpp:415-416),
// the shift by -centroid(cloud1_prep) (:420-436) and PatchGenerationAndRefinement of both (:653-654).  Outputs are packed
// xyz; patch points are concatenated by patch with CSR offsets.  Every out pointer may be NULL (sizes only); sizes[8] =
// {m1, m2, np1, np2, patch points 1, patch points 2, 0, 0}.  Returns 0, or -1 when a capacity is too small.
int pwicp_host_prepare_pair(const float* xyz1, int n1, const float* xyz2, int n2, float Res1, float Res2, float SV1, float SV2,
                            float* cloud1, float* cloud2, int capCloud, float* patch1, int* off1, float* patch2, int* off2,
                            int capPatchPts, int capPatches, float* shift3, int* sizes) {
    typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
    Cloud::Ptr in1(new Cloud), in2(new Cloud), p1(new Cloud), p2(new Cloud), r1(new Cloud), r2(new Cloud);
    in1->resize(n1); in2->resize(n2);
    for (int i = 0; i < n1; ++i) { in1->points[i].x = xyz1[3 * i]; in1->points[i].y = xyz1[3 * i + 1]; in1->points[i].z = xyz1[3 * i + 2]; }
    for (int i = 0; i < n2; ++i) { in2->points[i].x = xyz2[3 * i]; in2->points[i].y = xyz2[3 * i + 1]; in2->points[i].z = xyz2[3 * i + 2]; }
    PCpreprocessingHost(in1, p1, true, Res1, 14, 5.0);
    PCpreprocessingHost(in2, p2, true, Res2, 14, 5.0);
    Eigen::Vector4f c;
    pcl::compute3DCentroid(*p1, c);
    Eigen::Matrix4f S = Eigen::Matrix4f::Identity();
    S(0, 3) = -1 * c[0]; S(1, 3) = -1 * c[1]; S(2, 3) = -1 * c[2];
    pcl::transformPointCloud(*p1, *r1, S);
    pcl::transformPointCloud(*p2, *r2, S);
    if (shift3) { shift3[0] = S(0, 3); shift3[1] = S(1, 3); shift3[2] = S(2, 3); }
    Cloud::Ptr CT1(new Cloud), CT2(new Cloud), BP1(new Cloud), BP2(new Cloud);
    Cloud* SV1c = nullptr; Cloud* SV2c = nullptr;
    const int np1 = PatchGenerationAndRefinement(r1, SV1, CT1, BP1, SV1c, false);
    const int np2 = PatchGenerationAndRefinement(r2, SV2, CT2, BP2, SV2c, false);
    int tot1 = 0, tot2 = 0;
    for (int i = 0; i < np1; ++i) tot1 += (int)SV1c[i].size();
    for (int i = 0; i < np2; ++i) tot2 += (int)SV2c[i].size();
    const int sz[8] = {(int)r1->size(), (int)r2->size(), np1, np2, tot1, tot2, 0, 0};
    if (sizes) std::memcpy(sizes, sz, sizeof(sz));
    int rc = 0;
    if ((cloud1 || cloud2) && (sz[0] > capCloud || sz[1] > capCloud)) rc = -1;
    if ((patch1 || patch2) && (tot1 > capPatchPts || tot2 > capPatchPts || np1 > capPatches || np2 > capPatches)) rc = -1;
    if (rc == 0) {
        auto packCloud = [](const Cloud& cl, float* o) { if (o) for (size_t i = 0; i < cl.size(); ++i) { o[3 * i] = cl.points[i].x; o[3 * i + 1] = cl.points[i].y; o[3 * i + 2] = cl.points[i].z; } };
        packCloud(*r1, cloud1); packCloud(*r2, cloud2);
        auto packPatches = [](Cloud* sv, int np, float* o, int* off) {
            if (!o || !off) return;
            int k = 0; off[0] = 0;
            for (int i = 0; i < np; ++i) {
                for (const auto& p : sv[i].points) { o[3 * k] = p.x; o[3 * k + 1] = p.y; o[3 * k + 2] = p.z; ++k; }
                off[i + 1] = k;
            }
        };
        packPatches(SV1c, np1, patch1, off1); packPatches(SV2c, np2, patch2, off2);
    }
    delete[] SV1c; delete[] SV2c;
    return rc;
}

// PCpreprocessing on a host array: device != 0 -> the driver's path (pwicp_preprocess), else the host statements
int pwicp_host_preprocess(const float* xyz, int n, int downsample, float leaf, int k, double mult, int device, float* out, int cap) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr in(new pcl::PointCloud<pcl::PointXYZ>), res(new pcl::PointCloud<pcl::PointXYZ>);
    in->resize(n);
    for (int i = 0; i < n; ++i) { in->points[i].x = xyz[3 * i]; in->points[i].y = xyz[3 * i + 1]; in->points[i].z = xyz[3 * i + 2]; }
    if (device) PCpreprocessing(in, res, downsample != 0, leaf, k, mult);
    else PCpreprocessingHost(in, res, downsample != 0, leaf, k, mult);
    const int m = (int)res->size();
    for (int i = 0; i < m && i < cap; ++i) { out[3 * i] = res->points[i].x; out[3 * i + 1] = res->points[i].y; out[3 * i + 2] = res->points[i].z; }
    return m;
}

// calTransToReferenceEpoch as a file-to-file operation (F2)
void pwicp_host_chain_to_reference(const char* transMatFile, int pairMode, const char* pairFile, int epochNum,
                                   const char* outTM, const char* outTP) {
    std::vector<int> ts; std::vector<Eigen::Matrix4f> T; std::vector<Eigen::MatrixXd> V;
    calTransToReferenceEpoch(transMatFile, pairMode, pairFile, epochNum, outTM, outTP, ts, T, V);
}

// calAbsErrorOfTransPara as a file-to-file operation (src/Registration.cpp:1157-1251)
void pwicp_host_abs_error(const char* transMatFile, const char* gtFile, int allEpochNum, int startEpoch, const char* outFile) {
    calAbsErrorOfTransPara(transMatFile, gtFile, allEpochNum, startEpoch, outFile);
}

}  // extern "C"

// ---- the reference's outer loop written with the mirror's per-iteration function ---------------
// mode 0: Piecewise_ICP (one upload, device loop); mode 1: the reference's own loop shape
// (src/Registration.cpp:680-694): while(!g_toStage3) PwICP_singleIteration(...), composing
// transMat = cur * transMat on the host.  Both must give the same answer.
extern bool g_toStage2, g_toStage3;

extern "C" int pwicp_host_register_clouds(const float* xyz1, int n1, const float* xyz2, int n2, float Res, float SVsize,
                                          float DTinit, float DTmin, int mode, float* T16, double* vcm36, float* dtseries, int cap) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr c1(new pcl::PointCloud<pcl::PointXYZ>), c2(new pcl::PointCloud<pcl::PointXYZ>);
    c1->resize(n1); c2->resize(n2);
    for (int i = 0; i < n1; ++i) { c1->points[i].x = xyz1[3 * i]; c1->points[i].y = xyz1[3 * i + 1]; c1->points[i].z = xyz1[3 * i + 2]; }
    for (int i = 0; i < n2; ++i) { c2->points[i].x = xyz2[3 * i]; c2->points[i].y = xyz2[3 * i + 1]; c2->points[i].z = xyz2[3 * i + 2]; }
    std::vector<float> series;
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    Eigen::MatrixXd VCM;
    if (mode == 0) {
        Piecewise_ICP(c1, c2, true, Res, Res, SVsize, SVsize, true, DTinit, DTmin, series, T, VCM);
    } else {
        g_toStage2 = false; g_toStage3 = false;
        pcl::PointCloud<pcl::PointXYZ>* SV1 = nullptr; pcl::PointCloud<pcl::PointXYZ>* SV2 = nullptr;
        pcl::PointCloud<pcl::PointXYZ>::Ptr CT1(new pcl::PointCloud<pcl::PointXYZ>), CT2(new pcl::PointCloud<pcl::PointXYZ>);
        pcl::PointCloud<pcl::PointXYZ>::Ptr BP1(new pcl::PointCloud<pcl::PointXYZ>), BP2(new pcl::PointCloud<pcl::PointXYZ>);
        const int m1 = PatchGenerationAndRefinement(c1, SVsize, CT1, BP1, SV1, false);
        const int m2 = PatchGenerationAndRefinement(c2, SVsize, CT2, BP2, SV2, false);
        std::vector<float> BPstd1, BPstd2, CTstd1, CTstd2;
        calBPandCTSTD(SV1, m1, BPstd1, CTstd1);
        calBPandCTSTD(SV2, m2, BPstd2, CTstd2);
        float currDT = DTinit, BB1 = 0.0f, BB2 = 0.0f;
        series.push_back(currDT);
        while (!g_toStage3) {
            Eigen::Matrix4f cur = PwICP_singleIteration(c1, c2, Res, Res, SVsize, SVsize, SV1, SV2, CT1, CT2, BP1, BP2,
                                                        CTstd1, BPstd2, DTmin, currDT, BB1, BB2, VCM);
            T = cur * T;
            series.push_back(currDT);
        }
        delete[] SV1; delete[] SV2;
    }
    std::memcpy(T16, T.m, sizeof(T.m));
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) vcm36[r * 6 + c] = VCM(r, c);
    for (int k = 0; k < (int)series.size() && k < cap; ++k) dtseries[k] = series[k];
    return (int)series.size();
}
