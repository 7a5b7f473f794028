// Registration.cpp -- host mirror of the reference's src/Registration.cpp on top of libpwicp.so.
// Drivers, file formats and error behaviour follow the reference (citations below); the
// arithmetic of the registration loop runs on the device through the C ABI of include/pwicp.h.
#include "Registration.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>

#include "../../include/pwicp.h"

using namespace std;

// The reference keeps the DT-stage state in file-scope globals (src/Registration.cpp:11-14), which
// makes it non re-entrant.  They are kept for PwICP_singleIteration's call shape; the device outer
// loop (Piecewise_ICP) carries the state in a pwicp_state per pair instead.
bool g_toStage2 = false;
bool g_toStage3 = false;
bool g_isVis = false;

namespace {

vector<float> packXYZ(const pcl::PointCloud<pcl::PointXYZ>& c) {
    vector<float> v(3 * c.size());
    for (size_t i = 0; i < c.size(); ++i) { v[3 * i] = c.points[i].x; v[3 * i + 1] = c.points[i].y; v[3 * i + 2] = c.points[i].z; }
    return v;
}

void unpackXYZ(const vector<float>& v, pcl::PointCloud<pcl::PointXYZ>& c) {
    for (size_t i = 0; i < c.size(); ++i) { c.points[i].x = v[3 * i]; c.points[i].y = v[3 * i + 1]; c.points[i].z = v[3 * i + 2]; }
}

[[noreturn]] void fatal(const string& what) {
    std::cerr << "Error: " << what << " Aborting.\n";
    std::exit(EXIT_FAILURE);
}

void check(int status, const char* where) {
    if (status == PWICP_OK) return;
    // the reference exits the process on these conditions (src/Registration.cpp:728-731, :864-867)
    if (status == PWICP_ERR_TOO_FEW_PATCHES) fatal("No enough stable points left (<4)!");
    if (status == PWICP_ERR_TOO_FEW_STABLE) fatal("No enough stable points left, no enough overlapping areas!!!");
    fatal(string(where) + ": " + pwicp_last_error(pwicpHostContext()));
}

// calBPandCTSTD (src/Segmentation.cpp:306-321) for all patches of a cloud in one launch of the batched
// device kernel (pwicp_patch_stats; the arithmetic of calPatchSTD, csrc/patch_algebra.cuh).
// The per-patch host function of the same name stays available for callers without a device.
void packPatches(pcl::PointCloud<pcl::PointXYZ>* patches, int n, vector<float>& xyz, vector<int>& off) {
    off.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + (int)patches[i].size();
    xyz.resize(3 * (size_t)off[n]);
    for (int i = 0; i < n; ++i)
        for (size_t k = 0; k < patches[i].size(); ++k) {
            const auto& p = patches[i].points[k];
            float* o = &xyz[3 * ((size_t)off[i] + k)];
            o[0] = p.x; o[1] = p.y; o[2] = p.z;
        }
}

void calBPandCTSTDDevice(pcl::PointCloud<pcl::PointXYZ>* patches, int n, vector<float>& stdBP, vector<float>& stdCT) {
    stdBP.assign(std::max(n, 0), 0.f); stdCT.assign(std::max(n, 0), 0.f);
    if (n < 1) return;
    vector<float> xyz; vector<int> off;
    packPatches(patches, n, xyz, off);
    check(pwicp_patch_stats(pwicpHostContext(), xyz.data(), off.data(), n, nullptr, nullptr, nullptr, nullptr,
                            stdBP.data(), stdCT.data()), "patch sigmas");
}

// Everything a pair needs on the device, derived from the patch arrays the way
// PwICP_singleIteration derives it every iteration (normals :821-824, concatenated patches).
struct PairHost {
    vector<float> ct1, nrm1, ctstd1, ct2, bp2, bpstd2, patch2, cloud1, cloud2;
    vector<unsigned char> ok1;
    vector<int> off2;
};

void buildPairHost(PairHost& h, const pcl::PointCloud<pcl::PointXYZ>& cloud1, const pcl::PointCloud<pcl::PointXYZ>& cloud2,
                   pcl::PointCloud<pcl::PointXYZ>* SVcloud1, pcl::PointCloud<pcl::PointXYZ>* SVcloud2,
                   const pcl::PointCloud<pcl::PointXYZ>& CT1, const pcl::PointCloud<pcl::PointXYZ>& CT2,
                   const pcl::PointCloud<pcl::PointXYZ>& BP2, const vector<float>& CTstd1, const vector<float>& BPstd2) {
    const int n1 = (int)CT1.size(), n2 = (int)CT2.size();
    h.ct1 = packXYZ(CT1); h.ct2 = packXYZ(CT2); h.bp2 = packXYZ(BP2);
    h.cloud1 = packXYZ(cloud1); h.cloud2 = packXYZ(cloud2);
    h.ctstd1 = CTstd1; h.bpstd2 = BPstd2;
    h.nrm1.assign(3 * (size_t)n1, 0.f); h.ok1.assign(n1, 1);
    {
        // the constant target patch normals: computed once per pair, all patches in
        // one launch (pwicp_patch_stats = calPatchNormal per patch).  The classification uses
        // calPatchNormal's verdict (:783), the ICP target cloud the ">6 points, else (0,0,1)" rule
        // of generateCentroidCloudWithPatchNormals (:367).
        vector<int> off1; vector<float> xyz1;
        packPatches(SVcloud1, n1, xyz1, off1);
        check(pwicp_patch_stats(pwicpHostContext(), xyz1.data(), off1.data(), n1, nullptr, nullptr, h.nrm1.data(),
                                h.ok1.data(), nullptr, nullptr), "patch normals");
        for (int i = 0; i < n1; ++i)
            if (!(SVcloud1[i].size() > 6 && h.ok1[i])) { h.nrm1[3 * i] = 0; h.nrm1[3 * i + 1] = 0; h.nrm1[3 * i + 2] = 1; }
    }
    h.off2.assign(n2 + 1, 0);
    for (int i = 0; i < n2; ++i) h.off2[i + 1] = h.off2[i] + (int)SVcloud2[i].size();
    h.patch2.resize(3 * (size_t)h.off2[n2]);
    for (int i = 0; i < n2; ++i)
        for (size_t k = 0; k < SVcloud2[i].size(); ++k) {
            const auto& p = SVcloud2[i].points[k];
            float* o = &h.patch2[3 * ((size_t)h.off2[i] + k)];
            o[0] = p.x; o[1] = p.y; o[2] = p.z;
        }
}

void uploadPair(pwicp_ctx* ctx, const PairHost& h) {
    check(pwicp_target_upload(ctx, h.ct1.data(), h.nrm1.data(), h.ok1.data(), h.ctstd1.data(), (int)h.ctstd1.size()), "target upload");
    check(pwicp_source_upload(ctx, h.ct2.data(), h.bp2.data(), h.bpstd2.data(), h.off2.data(), h.patch2.data(), (int)h.bpstd2.size()), "source upload");
    check(pwicp_clouds_upload(ctx, h.cloud1.data(), (int)(h.cloud1.size() / 3), h.cloud2.data(), (int)(h.cloud2.size() / 3)), "cloud upload");
}

// mutated source-side data back into the caller's clouds (src/Registration.cpp:942-954)
void downloadSource(pwicp_ctx* ctx, PairHost& h, pcl::PointCloud<pcl::PointXYZ>& cloud2, pcl::PointCloud<pcl::PointXYZ>* SVcloud2,
                    pcl::PointCloud<pcl::PointXYZ>& CT2, pcl::PointCloud<pcl::PointXYZ>& BP2) {
    check(pwicp_source_download(ctx, h.cloud2.data(), h.ct2.data(), h.bp2.data(), h.patch2.data()), "download");
    unpackXYZ(h.cloud2, cloud2); unpackXYZ(h.ct2, CT2); unpackXYZ(h.bp2, BP2);
    for (size_t i = 0; i + 1 < h.off2.size(); ++i)
        for (size_t k = 0; k < SVcloud2[i].size(); ++k) {
            const float* o = &h.patch2[3 * ((size_t)h.off2[i] + k)];
            SVcloud2[i].points[k].x = o[0]; SVcloud2[i].points[k].y = o[1]; SVcloud2[i].points[k].z = o[2];
        }
}

void writeTransMatrixFile(ofstream& out, const Eigen::Matrix4f& T, const Eigen::Vector3f& ang, const Eigen::Vector3f& tr,
                          const Eigen::MatrixXd& VCM) {
    // layout of src/Registration.cpp:349-386 / :500-538
    out << "4x4 Transformation Matrix:\n";
    out << fixed << setprecision(12);
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) out << T(i, j) << " "; out << "\n"; }
    out << endl;
    out << "Rotation Angles (unit: gon):\n" << fixed << setprecision(10)
        << "Rx = " << ang[0] * ARC_TO_GON << "\n" << "Ry = " << ang[1] * ARC_TO_GON << "\n" << "Rz = " << ang[2] * ARC_TO_GON << "\n";
    out << "Translation (unit: m):\n" << "tx = " << tr[0] << "\n" << "ty = " << tr[1] << "\n" << "tz = " << tr[2] << "\n";
    out << endl;
    out << "6x6 Variance-Covariance Matrix of transformation parameters:\n";
    out << fixed << setprecision(12);
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) out << VCM(i, j) << " "; out << "\n"; }
    out << endl;
    out << "Standard Deviations of estimated transformation parameters:\n";
    out << fixed << setprecision(10)
        << "Std_Rx = " << 1000 * ARC_TO_GON * sqrt(VCM(0, 0)) << " mgon\n"
        << "Std_Ry = " << 1000 * ARC_TO_GON * sqrt(VCM(1, 1)) << " mgon\n"
        << "Std_Rz = " << 1000 * ARC_TO_GON * sqrt(VCM(2, 2)) << " mgon\n"
        << "Std_tx = " << 1000 * sqrt(VCM(3, 3)) << " mm\n"
        << "Std_ty = " << 1000 * sqrt(VCM(4, 4)) << " mm\n"
        << "Std_tz = " << 1000 * sqrt(VCM(5, 5)) << " mm\n";
}

}  // namespace

// the per-pair result file (TransMatrix.txt / <time>_<mode>_TransMatrix.txt) from a 4x4 and a 6x6, for tools and tests
extern "C" int pwicp_host_write_transmatrix(const char* path, const float* T16, const double* vcm36) {
    Eigen::Matrix4f T; std::memcpy(T.m, T16, sizeof(T.m));
    Eigen::MatrixXd V; V.resize(6, 6);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) V(r, c) = vcm36[r * 6 + c];
    Eigen::Vector3f ang, tr;
    matrix2angle(T, ang);
    tr[0] = T(0, 3); tr[1] = T(1, 3); tr[2] = T(2, 3);
    ofstream out(path);
    if (!out) return 0;
    writeTransMatrixFile(out, T, ang, tr, V);
    return 1;
}

namespace {

// shift by -centroid(cloud1_prep), run the core, conjugate back (src/Registration.cpp:276-319, :419-461)
struct CoreResult { Eigen::Matrix4f T_final; Eigen::Vector3f ang, tr; Eigen::MatrixXd VCM; vector<float> DTseries; };

CoreResult runShiftedCore(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1_prep, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2_prep,
                          bool isSetResSVsize, float Res1, float Res2, float SVsize1, float SVsize2,
                          bool isManualDTinit, float DTinit, float DTmin) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr red1(new pcl::PointCloud<pcl::PointXYZ>), red2(new pcl::PointCloud<pcl::PointXYZ>);
    Eigen::Vector4f PC1CT;
    pcl::compute3DCentroid(*cloud1_prep, PC1CT);
    const float sx = -1 * PC1CT[0], sy = -1 * PC1CT[1], sz = -1 * PC1CT[2];
    Eigen::Matrix4f S = Eigen::Matrix4f::Identity(), Sinv = Eigen::Matrix4f::Identity();
    S(0, 3) = sx; S(1, 3) = sy; S(2, 3) = sz;
    Sinv(0, 3) = -1 * sx; Sinv(1, 3) = -1 * sy; Sinv(2, 3) = -1 * sz;
    pcl::transformPointCloud(*cloud1_prep, *red1, S);
    pcl::transformPointCloud(*cloud2_prep, *red2, S);
    cout << "\nPreprocessed PC-1 point number: " << red1->size() << "\tPreprocessed PC-2 point number: " << red2->size() << endl << endl;

    pcl::console::TicToc time;
    cout << "\n--->>> Compute Core TransMat... "; time.tic();
    CoreResult r;
    Eigen::Matrix4f transMat = Eigen::Matrix4f::Identity();
    Piecewise_ICP(red1, red2, isSetResSVsize, Res1, Res2, SVsize1, SVsize2, isManualDTinit, DTinit, DTmin, r.DTseries, transMat, r.VCM);
    cout << "--->>> Computing time of core TransMat: " << int(0.001 * time.toc()) << " s \n\n";
    r.T_final = Sinv * transMat * S;
    cout << "Final Registration TransMatrix: \n";
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) cout << r.T_final(i, j) << " "; cout << "\n"; }
    matrix2angle(r.T_final, r.ang);
    r.tr[0] = r.T_final(0, 3); r.tr[1] = r.T_final(1, 3); r.tr[2] = r.T_final(2, 3);
    cout << fixed << setprecision(6);
    cout << "Rotation (degree): " << r.ang[0] * 180.0 / M_PI << " " << r.ang[1] * 180.0 / M_PI << " " << r.ang[2] * 180.0 / M_PI << "\n";
    cout << "Translation (m): " << r.tr[0] << " " << r.tr[1] << " " << r.tr[2] << "\n\n";
    return r;
}

}  // namespace

// ================================================================================================
// The hot path: Piecewise_ICP / PwICP_singleIteration / P2PICPwithPatchNormal / calTransParaVCM
// ================================================================================================

void Piecewise_ICP(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2,
                   bool isSetResSVsize, float Res1, float Res2, float SVsize1, float SVsize2,
                   bool isManualDTinit, float DTinit, float DTmin,
                   std::vector<float>& DTseries, Eigen::Matrix4f& transMat, Eigen::MatrixXd& VCM) {
    g_toStage2 = false; g_toStage3 = false;                                  // :623-624
    float SVRes1 = Res1 * 10, SVRes2 = Res2 * 10;                            // :635-640
    if (isSetResSVsize) { SVRes1 = SVsize1; SVRes2 = SVsize2; }

    pcl::PointCloud<pcl::PointXYZ>* SVcloud1 = nullptr;
    pcl::PointCloud<pcl::PointXYZ>* SVcloud2 = nullptr;
    pcl::PointCloud<pcl::PointXYZ>::Ptr CT1(new pcl::PointCloud<pcl::PointXYZ>), CT2(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::PointCloud<pcl::PointXYZ>::Ptr BP1(new pcl::PointCloud<pcl::PointXYZ>), BP2(new pcl::PointCloud<pcl::PointXYZ>);
    const int num1 = PatchGenerationAndRefinement(cloud1, SVRes1, CT1, BP1, SVcloud1, g_isVis);    // :653-654
    const int num2 = PatchGenerationAndRefinement(cloud2, SVRes2, CT2, BP2, SVcloud2, g_isVis);
    cout << "PC-1 selected patch number: " << num1 << "\tPC-2 selected patch number: " << num2 << endl;
    cout << "---------------------------------------------------------------------------- \n\n";
    if (num1 < 1) fatal("no patch in the target cloud!");
    std::vector<float> BPstd1, BPstd2, CTstd1, CTstd2;
    calBPandCTSTDDevice(SVcloud1, num1, BPstd1, CTstd1);                                            // :663-664
    calBPandCTSTDDevice(SVcloud2, num2, BPstd2, CTstd2);
    if (4 > num2) fatal("No enough stable points left (<4)!");                                      // :728-731

    // one upload, the whole while(!g_toStage3) loop on the device (:680-694)
    pwicp_ctx* ctx = pwicpHostContext();
    PairHost h;
    buildPairHost(h, *cloud1, *cloud2, SVcloud1, SVcloud2, *CT1, *CT2, *BP2, CTstd1, BPstd2);
    uploadPair(ctx, h);
    pwicp_pair_params pp = {Res1, Res2, SVRes1, SVRes2, DTmin};
    const int max_outer = 1000;
    vector<float> series(max_outer + 1);
    vector<pwicp_iter_stats> stats(max_outer);
    int ns = 0, n_outer = 0;
    float T16[16];
    double vcm36[36] = {0};
    cout << "Start Piecewise-ICP iteration... \n\n";
    check(pwicp_piecewise_icp(ctx, &pp, isManualDTinit ? 1 : 0, DTinit, nullptr, max_outer, series.data(), &ns, T16, vcm36,
                              &n_outer, stats.data()), "Piecewise_ICP");
    cout << endl << "DT initial value = " << series[0] << " m \n\n";
    for (int k = 0; k < n_outer; ++k)
        cout << "--->>> Iteration No." << k + 1 << " | Current DT = " << series[k + 1] * 100 << " cm. \t stable patches: "
             << stats[k].n_stable << " | inner iterations: " << stats[k].icp_iters << " | bbox change: " << stats[k].maxBBchange * 100
             << " cm | device " << stats[k].device_ms << " ms\n";
    g_toStage2 = g_toStage3 = true;
    if (const char* tr = getenv("PWICP_TRACE_JSON")) {
        // machine-readable trace (one JSON line per registered pair, appended): what the reference only prints
        // (:691-693, :870-871, :888) plus the device time of every outer iteration
        ofstream js(tr, std::ios::app);
        if (js) {
            js << std::setprecision(9) << "{\"n1\": " << num1 << ", \"n2\": " << num2 << ", \"outer_iterations\": " << n_outer << ", \"iterations\": [";
            for (int k = 0; k < n_outer; ++k) {
                const pwicp_iter_stats& st = stats[k];
                js << (k ? ", " : "") << "{\"DT\": " << series[k] << ", \"DT_next\": " << series[k + 1] << ", \"n_stable\": " << st.n_stable
                   << ", \"n_stable_points\": " << st.n_stable_pts << ", \"inner_iterations\": " << st.icp_iters
                   << ", \"bbox_change_m\": " << st.maxBBchange << ", \"vcm_written\": " << st.vcm_written
                   << ", \"device_ms\": " << st.device_ms << "}";
            }
            js << "], \"T\": [";
            for (int k = 0; k < 16; ++k) js << (k ? ", " : "") << T16[k];
            js << "]}\n";
        }
    }
    DTseries.assign(series.begin(), series.begin() + ns);                                           // :675-688
    memcpy(transMat.m, T16, sizeof(T16));
    VCM.resize(6, 6);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) VCM(r, c) = vcm36[r * 6 + c];
    downloadSource(ctx, h, *cloud2, SVcloud2, *CT2, *BP2);          // cloud2 is left transformed, like the reference
    delete[] SVcloud1; delete[] SVcloud2;                                                           // :696-697
    cout << "Computed TransMat after " << n_outer << " iterations.\n*******************************************\n\n";
}

Eigen::Matrix4f PwICP_singleIteration(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2,
                                      float Res1, float Res2, float SVRes1, float SVRes2,
                                      pcl::PointCloud<pcl::PointXYZ>*& SVcloud1, pcl::PointCloud<pcl::PointXYZ>*& SVcloud2,
                                      pcl::PointCloud<pcl::PointXYZ>::Ptr CTcloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr CTcloud2,
                                      pcl::PointCloud<pcl::PointXYZ>::Ptr /*BPcloud1*/, pcl::PointCloud<pcl::PointXYZ>::Ptr BPcloud2,
                                      std::vector<float> CTstd1, std::vector<float> BPstd2, float DTmin,
                                      float& currDT, float& BBchange_1, float& BBchange_2, Eigen::MatrixXd& VCM) {
    pwicp_ctx* ctx = pwicpHostContext();
    PairHost h;
    buildPairHost(h, *cloud1, *cloud2, SVcloud1, SVcloud2, *CTcloud1, *CTcloud2, *BPcloud2, CTstd1, BPstd2);
    uploadPair(ctx, h);
    pwicp_pair_params pp = {Res1, Res2, SVRes1, SVRes2, DTmin};
    pwicp_state st = {currDT, BBchange_1, BBchange_2, g_toStage2 ? 1 : 0, g_toStage3 ? 1 : 0};
    pwicp_iter_stats stats;
    float T16[16];
    double vcm36[36];
    check(pwicp_single_iteration(ctx, &pp, &st, nullptr, T16, vcm36, nullptr, &stats), "PwICP_singleIteration");
    cout << "Ratio of stable points: " << 100.0 * stats.n_stable_pts / float(h.off2.back()) << " %\n";
    cout << "Change of bounding box = " << stats.maxBBchange * 100 << " cm \n";
    currDT = st.currDT; BBchange_1 = st.BBchange_1; BBchange_2 = st.BBchange_2;
    g_toStage2 = st.toStage2 != 0; g_toStage3 = st.toStage3 != 0;
    if (stats.vcm_written) {                                                                        // :958-961
        VCM.resize(6, 6);
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) VCM(r, c) = vcm36[r * 6 + c];
    }
    downloadSource(ctx, h, *cloud2, SVcloud2, *CTcloud2, *BPcloud2);                                // :942-954
    Eigen::Matrix4f T;
    memcpy(T.m, T16, sizeof(T16));
    return T;
}

Eigen::Matrix4f P2PICPwithPatchNormal(pcl::PointCloud<pcl::PointNormal>::Ptr cloudTarget,
                                      pcl::PointCloud<pcl::PointNormal>::Ptr cloudSource, double EucldEpsilon) {
    const int n1 = (int)cloudTarget->size(), n2 = (int)cloudSource->size();
    vector<float> t(3 * (size_t)n1), nr(3 * (size_t)n1), s(3 * (size_t)n2);
    for (int i = 0; i < n1; ++i) {
        const auto& p = cloudTarget->points[i];
        t[3 * i] = p.x; t[3 * i + 1] = p.y; t[3 * i + 2] = p.z;
        nr[3 * i] = p.normal_x; nr[3 * i + 1] = p.normal_y; nr[3 * i + 2] = p.normal_z;
    }
    for (int i = 0; i < n2; ++i) { const auto& p = cloudSource->points[i]; s[3 * i] = p.x; s[3 * i + 1] = p.y; s[3 * i + 2] = p.z; }
    pwicp_icp_params prm;
    pwicp_icp_default_params(&prm);                   // 1e-8 / 100 iterations (:1262-1264)
    prm.fit_eps = EucldEpsilon;                       // :1263
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    pwicp_icp_result res;
    check(pwicp_icp_p2plane(pwicpHostContext(), t.data(), nr.data(), n1, s.data(), n2, &prm, T.m, &res), "P2PICPwithPatchNormal");
    return T;
}

Eigen::MatrixXd calTransParaVCM(pcl::PointCloud<pcl::PointXYZ>::Ptr cloudTarget,
                                pcl::PointCloud<pcl::PointNormal>::Ptr cloudTargetwithNormals,
                                pcl::PointCloud<pcl::PointXYZ>::Ptr cloudSourceStable) {
    const int n1 = (int)cloudTarget->size();
    vector<float> t = packXYZ(*cloudTarget), nr(3 * (size_t)n1), s = packXYZ(*cloudSourceStable);
    for (int i = 0; i < n1; ++i) {
        const auto& p = cloudTargetwithNormals->points[i];
        nr[3 * i] = p.normal_x; nr[3 * i + 1] = p.normal_y; nr[3 * i + 2] = p.normal_z;
    }
    pwicp_ctx* ctx = pwicpHostContext();
    check(pwicp_target_upload(ctx, t.data(), nr.data(), nullptr, nullptr, n1), "calTransParaVCM");
    double v[36]; int singular = 0;
    check(pwicp_vcm(ctx, s.data(), (int)cloudSourceStable->size(), v, &singular), "calTransParaVCM");
    if (singular) cout << "\n This is a singular matrix! \n" << endl;                              // :1324-1325
    Eigen::MatrixXd D(6, 6);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) D(r, c) = v[r * 6 + c];
    return D;
}

float calOverlapRatioByC2Cdist(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2, float DTinit) {
    vector<float> a = packXYZ(*cloud1), b = packXYZ(*cloud2);
    float out = 0;
    check(pwicp_overlap_ratio(pwicpHostContext(), a.data(), (int)cloud1->size(), b.data(), (int)cloud2->size(), DTinit, &out),
          "calOverlapRatioByC2Cdist");
    return out;
}

// ================================================================================================
// Drivers (API surface + output formats)
// ================================================================================================

bool Piecewise_ICP_4D(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2,
                      bool isSetResSVsize, float Res1, float Res2, float SVsize1, float SVsize2,
                      bool isManualDTinit, float DTinit, float DTmin, std::string outfileIdx,
                      Eigen::Matrix4f& transMat, std::vector<float>& transPara, Eigen::MatrixXd& VCM) {
    cout << "Original PC-1 point number: " << cloud1->size() << "\t Original PC-2 point number: " << cloud2->size() << endl;
    cout << "PC-1 avg. point spacing: " << Res1 << "\t PC-2 avg. point spacing: " << Res2 << endl << endl;
    pcl::PointCloud<pcl::PointXYZ>::Ptr prep1(new pcl::PointCloud<pcl::PointXYZ>), prep2(new pcl::PointCloud<pcl::PointXYZ>);
    PCpreprocessing(cloud1, prep1, true, Res1, 14, 5.0);                                            // :415-416
    PCpreprocessing(cloud2, prep2, true, Res2, 14, 5.0);
    CoreResult r = runShiftedCore(prep1, prep2, isSetResSVsize, Res1, Res2, SVsize1, SVsize2, isManualDTinit, DTinit, DTmin);
    transMat = r.T_final;
    transPara.resize(6);                                                                            // :474-480
    transPara[0] = r.ang[0] * ARC_TO_GON; transPara[1] = r.ang[1] * ARC_TO_GON; transPara[2] = r.ang[2] * ARC_TO_GON;
    transPara[3] = r.tr[0]; transPara[4] = r.tr[1]; transPara[5] = r.tr[2];
    VCM = r.VCM;
    string name(outfileIdx);
    name.append("TransMatrix.txt");
    ofstream out(name.c_str());
    if (!out) { std::cerr << "Cannot open TransMatrix.txt for writing!\n\n"; return false; }
    writeTransMatrixFile(out, r.T_final, r.ang, r.tr, r.VCM);
    out.close();
    cout << "--->>> Transformation results saved.\n\n";
    return true;
}

extern "C" bool PiecewiseICP_pair_call(const char* confile, const char* outfile) {
    ConfigPara cfg;
    string confilename = confile;
    std::cout << "Loading parameter configuration file: " << confilename << "\n\n";
    if (!readConfigFile(confilename, cfg)) { std::cerr << "Error: Cannot open configuration file! Aborting.\n\n"; return false; }
    float Res1 = cfg.PCres1, Res2 = cfg.PCres2;
    g_isVis = cfg.isVisual;
    pcl::PointCloud<pcl::PointXYZ>::Ptr ori1(new pcl::PointCloud<pcl::PointXYZ>), ori2(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::io::loadPCDFile(cfg.FolderFilePath1, *ori1);
    pcl::io::loadPCDFile(cfg.FolderFilePath2, *ori2);
    if (ori1->size() < 1 || ori2->size() < 1) return false;                                         // :254-256
    if (!cfg.isSetResSVsize) { Res1 = calPCresolution(ori1); Res2 = calPCresolution(ori2); }         // :259-262
    cout << "Original PC-1 point number: " << ori1->size() << "\t Original PC-2 point number: " << ori2->size() << endl;
    cout << "PC-1 avg. point spacing: " << Res1 << "\t PC-2 avg. point spacing: " << Res2 << endl << endl;
    pcl::PointCloud<pcl::PointXYZ>::Ptr prep1(new pcl::PointCloud<pcl::PointXYZ>), prep2(new pcl::PointCloud<pcl::PointXYZ>);
    PCpreprocessing(ori1, prep1, true, Res1, 14, 2.7);                                              // :272-273
    PCpreprocessing(ori2, prep2, true, Res2, 14, 2.7);
    CoreResult r = runShiftedCore(prep1, prep2, cfg.isSetResSVsize, Res1, Res2, cfg.SVsize1, cfg.SVsize2, cfg.isSetDTinit, cfg.DTinit, cfg.DTmin);

    pcl::PointCloud<pcl::PointXYZ>::Ptr moved(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::transformPointCloud(*ori2, *moved, r.T_final);                                             // :333
    string nameTM(outfile);
    nameTM.append("TransMatrix.txt");
    ofstream out(nameTM.c_str());
    if (!out) { std::cerr << "Cannot open TransMatrix.txt for writing!\n\n"; return false; }
    writeTransMatrixFile(out, r.T_final, r.ang, r.tr, r.VCM);
    out.close();
    cout << "--->>> Transformation results saved.\n";
    string namePC(outfile);
    namePC.append("RegisteredSourceCloud.pcd");
    pcl::io::savePCDFileBinary(namePC, *moved);                                                     // :392-394
    cout << "--->>> Registered source cloud saved.\n\n";
    return true;
}

bool calAdaptivePairSequence(std::vector<std::string> fileNameList, int startEpoch, float DTinit, float ratioThd,
                             std::map<int, int>& RegPairs, std::string adaptivePairFile) {
    int IdxTarget = startEpoch;
    for (int j = startEpoch + 1; j < (int)fileNameList.size(); ++j) {
        cout << "--> Computing for source cloud - " << j << " ... ";
        float OverlapRatio = 0;
        pcl::PointCloud<pcl::PointXYZ>::Ptr c2(new pcl::PointCloud<pcl::PointXYZ>);
        pcl::io::loadPCDFile(fileNameList[j], *c2);
        for (int i = IdxTarget; i < j; ++i) {
            pcl::PointCloud<pcl::PointXYZ>::Ptr c1(new pcl::PointCloud<pcl::PointXYZ>);
            pcl::io::loadPCDFile(fileNameList[i], *c1);
            OverlapRatio = calOverlapRatioByC2Cdist(c1, c2, DTinit);
            IdxTarget = i;
            if (OverlapRatio > ratioThd) break;
        }
        RegPairs.insert(std::make_pair(j - startEpoch, IdxTarget - startEpoch));
        cout << "Pair: " << IdxTarget - startEpoch << " - " << j - startEpoch << ";  Overlap ratio = " << 100 * OverlapRatio << "% \n";
    }
    if (RegPairs.size() != fileNameList.size() - 1) return false;
    ofstream out(adaptivePairFile);
    if (!out) fatal("Cannot open adaptivePairFile!");
    for (auto it = RegPairs.begin(); it != RegPairs.end(); ++it) out << it->first << " " << it->second << endl;
    out.close();
    return true;
}

namespace {

// per-epoch block of TransMatrices.txt and row of TransParameters.txt (src/Registration.cpp:151-180)
void appendEpochRecord(ofstream& outTM, ofstream& outTP, const pwicp_epoch_record& r) {
    outTM << fixed << setprecision(12);
    outTM << r.time_stamp << "\n";
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) outTM << r.T[i * 4 + j] << " "; outTM << "\n"; }
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) outTM << r.VCM[i * 6 + j] << " "; outTM << "\n"; }
    outTP << fixed << setprecision(10);
    outTP << r.time_stamp << " ";
    for (int p = 0; p < 6; ++p) outTP << r.para[p] << " ";
    outTP << 1000 * sqrt(r.VCM[0]) * ARC_TO_GON << " " << 1000 * sqrt(r.VCM[7]) * ARC_TO_GON << " "
          << 1000 * sqrt(r.VCM[14]) * ARC_TO_GON << " " << 1000 * sqrt(r.VCM[21]) << " " << 1000 * sqrt(r.VCM[28]) << " "
          << 1000 * sqrt(r.VCM[35]) << "\n";
}

struct Plan4D {
    ConfigPara cfg;
    vector<string> files;
    vector<long> times;
    std::map<int, int> regPairs;
};

bool plan4D(const char* confile, int startEpoch, int epochNum, int pairMode, float overlapThd, Plan4D& pl, bool writePairFile) {
    string confilename = confile;
    std::cout << "Loading parameter configuration file: " << confilename << "\n\n";
    if (!readConfigFile(confilename, pl.cfg)) { std::cerr << "Error: Cannot open configuration file! Aborting.\n\n"; return false; }
    g_isVis = pl.cfg.isVisual;
    const int count = extractAllFilesFromFolder(pl.cfg.FolderFilePath1, pl.files, pl.times);
    cout << "--->>> " << count << " scan files are successfully extracted. \n\n";
    if (count < 2 || epochNum > count || startEpoch < 0 || startEpoch >= epochNum) {
        std::cerr << "Error: not enough scan files for the requested epochs.\n";
        return false;
    }
    if (pairMode < 0) {                                                                             // :57-61
        cout << "--->>> Adaptive pair sequence determination... \n";
        calAdaptivePairSequence(pl.files, startEpoch, pl.cfg.DTinit, overlapThd, pl.regPairs,
                                writePairFile ? string("RegPairFile.txt") : string("/dev/null"));
    }
    return true;
}

// one iteration of the epoch loop, src/Registration.cpp:89-187
bool registerEpoch(const Plan4D& pl, int startEpoch, int pairMode, int i, pcl::PointCloud<pcl::PointXYZ>::Ptr refCloud,
                   pwicp_epoch_record& rec) {
    const int step = i - startEpoch + 1;
    pcl::console::TicToc time; time.tic();
    int refIdx = startEpoch;
    if (pairMode > 0) refIdx = (pairMode >= step) ? startEpoch : (i + 1 - pairMode);                // :95-97
    else if (pairMode < 0) {                                                                        // :98-100
        // the reference reads regPairs[i + 1] although calAdaptivePairSequence stores keys and targets relative to
        // startEpoch (:563): the two agree for startEpoch == 0 only; a missing key is std::map::operator[]'s 0 there.
        // Mirrored as is (a drop-in must pick the same pairs), without operator[]'s insertion and without a throw
        // across the extern "C" boundary.
        const auto it = pl.regPairs.find(i + 1);
        refIdx = (it == pl.regPairs.end()) ? 0 : it->second;
    }
    cout << "\n//////////////////////  Process Pair_" << step << ":  Epoch-" << pl.times[refIdx] << " and Epoch-"
         << pl.times[i + 1] << "   //////////////////////////////////////////// \n\n";
    std::string prefix = pl.cfg.FolderFilePath2 + std::to_string(pl.times[i + 1]);
    pcl::PointCloud<pcl::PointXYZ>::Ptr ori1(new pcl::PointCloud<pcl::PointXYZ>), ori2(new pcl::PointCloud<pcl::PointXYZ>);
    if (pairMode == 0) { prefix.append("_Direct2Ref_"); pcl::copyPointCloud(*refCloud, *ori1); }
    else if (pairMode > 0) { prefix.append("_Fixed_"); pcl::io::loadPCDFile(pl.files[refIdx], *ori1); }
    else { prefix.append("_Adaptive_"); pcl::io::loadPCDFile(pl.files[refIdx], *ori1); }
    pcl::io::loadPCDFile(pl.files[i + 1], *ori2);
    float Res1 = pl.cfg.PCres1, Res2 = pl.cfg.PCres2;
    if (!pl.cfg.isSetResSVsize) { Res1 = calPCresolution(ori1); Res2 = calPCresolution(ori2); }
    Eigen::Matrix4f T;
    std::vector<float> para;
    Eigen::MatrixXd VCM;
    const bool ok = Piecewise_ICP_4D(ori1, ori2, pl.cfg.isSetResSVsize, Res1, Res2, pl.cfg.SVsize1, pl.cfg.SVsize2,
                                     pl.cfg.isSetDTinit, pl.cfg.DTinit, pl.cfg.DTmin, prefix, T, para, VCM);
    memset(&rec, 0, sizeof(rec));
    rec.step = step;
    rec.time_stamp = pl.times[i + 1];
    rec.seconds = (float)(0.001 * time.toc());
    if (!ok) { std::cerr << "Step " << step << " failed. Skipping to next.\n\n"; return false; }
    rec.status = 1;
    memcpy(rec.T, T.m, sizeof(rec.T));
    for (int p = 0; p < 6; ++p) rec.para[p] = para[p];
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) rec.VCM[r * 6 + c] = VCM(r, c);
    cout << "--->>> Step-" << step << "  Total computing time: " << int(rec.seconds) << " s \n\n";
    return true;
}

bool finalize4D(const Plan4D& pl, int startEpoch, int epochNum, int pairMode, const pwicp_epoch_record* records) {
    const string outTMname = pl.cfg.FolderFilePath2 + "TransMatrices.txt", outTPname = pl.cfg.FolderFilePath2 + "TransParameters.txt";
    ofstream outTM(outTMname.c_str()), outTP(outTPname.c_str());
    if (!outTM || !outTP) { std::cerr << "Error: Unable to open output file(s).\n"; return false; }
    outTP << "Epoch  Rx[gon]  Ry[gon]  Rz[gon]  tx[m]  ty[m]  tz[m]  Std_Rx[mgon]  Std_Ry[mgon]  Std_Rz[mgon]  "
          << "Std_tx[mm]  Std_ty[mm]  Std_tz[mm]" << endl;
    int written = 0;
    for (int i = startEpoch; i < epochNum - 1; ++i) {
        const pwicp_epoch_record& r = records[i - startEpoch];
        if (r.status != 1) continue;                       // a failed epoch is skipped, like :145-147
        appendEpochRecord(outTM, outTP, r);
        ++written;
    }
    outTM.close(); outTP.close();
    if (written != epochNum - startEpoch - 1) {
        std::cerr << "Error: " << (epochNum - startEpoch - 1 - written) << " epoch pair(s) failed; the chain to the reference epoch "
                  << "needs every pair (the reference would read a misaligned file here).\n";
        return false;
    }
    std::vector<int> timeStamp;
    std::vector<Eigen::Matrix4f> T2Ref;
    std::vector<Eigen::MatrixXd> VCM2Ref;
    calTransToReferenceEpoch(outTMname, pairMode, "RegPairFile.txt", epochNum - startEpoch - 1,
                             pl.cfg.FolderFilePath2 + "TransMatrices_toRef.txt", pl.cfg.FolderFilePath2 + "TransParameters_toRef.txt",
                             timeStamp, T2Ref, VCM2Ref);
    // accuracy analysis against the ground truth, when it is there (the reference hard-codes this
    // CWD-relative path and exits if it is missing, :207-211; here it is optional)
    const char* gt = getenv("PWICP_GROUND_TRUTH");
    string gtFile = gt ? gt : "data/data_synthetic/defined_transformations.txt";
    if (ifstream(gtFile).good())
        calAbsErrorOfTransPara(pl.cfg.FolderFilePath2 + "TransMatrices_toRef.txt", gtFile, epochNum, startEpoch,
                               pl.cfg.FolderFilePath2 + "TransPara_AbsError.txt");
    return true;
}

}  // namespace

extern "C" void pwicp_host_set_device(int device) { pwicpHostSetDevice(device); }

extern "C" int PiecewiseICP_4D_shard(const char* confile, int startEpoch, int epochNum, int pairMode, float overlapThd,
                                     int rank, int world, int device, pwicp_epoch_record* records) {
    if (!records || world < 1 || rank < 0 || rank >= world) return -1;
    if (device >= 0) pwicpHostSetDevice(device);
    Plan4D pl;
    if (!plan4D(confile, startEpoch, epochNum, pairMode, overlapThd, pl, rank == 0)) return -1;
    pcl::PointCloud<pcl::PointXYZ>::Ptr refCloud(new pcl::PointCloud<pcl::PointXYZ>);
    pcl::io::loadPCDFile(pl.files[startEpoch], *refCloud);                                          // :87
    int done = 0;
    for (int i = startEpoch; i < epochNum - 1; ++i) {
        pwicp_epoch_record& rec = records[i - startEpoch];
        memset(&rec, 0, sizeof(rec));
        rec.step = i - startEpoch + 1;
        if ((i - startEpoch) % world != rank) continue;     // epoch sharding, SURVEY.md 8(e)
        if (registerEpoch(pl, startEpoch, pairMode, i, refCloud, rec)) ++done;
    }
    return done;
}

extern "C" bool PiecewiseICP_4D_finalize(const char* confile, int startEpoch, int epochNum, int pairMode,
                                         const pwicp_epoch_record* records) {
    Plan4D pl;
    string confilename = confile;
    if (!readConfigFile(confilename, pl.cfg)) return false;
    extractAllFilesFromFolder(pl.cfg.FolderFilePath1, pl.files, pl.times);
    return finalize4D(pl, startEpoch, epochNum, pairMode, records);
}

extern "C" bool PiecewiseICP_4D_call(const char* confile, int startEpoch, int epochNum, int pairMode, float overlapThd) {
    // single-process form: one shard that holds every pair, then the same finalisation
    std::vector<pwicp_epoch_record> records((size_t)std::max(epochNum, 1));
    if (PiecewiseICP_4D_shard(confile, startEpoch, epochNum, pairMode, overlapThd, 0, 1, -1, records.data()) < 0) return false;
    Plan4D pl;
    string confilename = confile;
    if (!readConfigFile(confilename, pl.cfg)) return false;
    extractAllFilesFromFolder(pl.cfg.FolderFilePath1, pl.files, pl.times);
    return finalize4D(pl, startEpoch, epochNum, pairMode, records.data());
}

// ---- chaining to the reference epoch (src/Registration.cpp:977-1153) -----------------------------
void calTransToReferenceEpoch(std::string transMatFile, int pairMode, std::string adaptivePairFile, int epochNum,
                              std::string transMat2RefFile, std::string transPara2RefFile, std::vector<int>& timeStamp,
                              std::vector<Eigen::Matrix4f>& allTransMat2Ref, std::vector<Eigen::MatrixXd>& allVCM2Ref) {
    cout << "\n--->>> Calculate the transformation of each epoch to the reference epoch...\n";
    ifstream in(transMatFile);
    if (!in) fatal("Cannot open transMatFile!");
    std::vector<Eigen::Matrix4f> allT;
    std::vector<Eigen::MatrixXd> allV;
    for (int i = 0; i < epochNum; ++i) {
        int t; in >> t;
        Eigen::Matrix4f M = Eigen::Matrix4f::Identity();
        Eigen::MatrixXd V = Eigen::MatrixXd::Zero(6, 6);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) in >> M(r, c);
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) in >> V(r, c);
        timeStamp.push_back(t); allT.push_back(M); allV.push_back(V);
    }
    in.close();
    std::map<int, int> RegPair;
    if (pairMode < 0) {
        ifstream pin(adaptivePairFile);
        if (!pin) fatal("Cannot open adaptivePairFile!");
        for (int i = 0; i < epochNum; i++) { int s, t; pin >> s >> t; RegPair.insert(std::make_pair(s, t)); }
        pin.close();
        for (auto it = RegPair.begin(); it != RegPair.end(); ++it) cout << "--> Optimal pair: " << it->first << "-->" << it->second << endl;
    }
    ofstream outTM(transMat2RefFile.c_str());
    if (!outTM) fatal("Cannot open transMat2RefFile!");
    ofstream outTP(transPara2RefFile.c_str());
    if (!outTP) fatal("Cannot open transPara2RefFile!");
    outTP << "Epoch  Rx[gon]  Ry[gon]  Rz[gon]  tx[m]  ty[m]  tz[m]  Std_Rx[mgon]  Std_Ry[mgon]  Std_Rz[mgon]  "
          << "Std_tx[mm]  Std_ty[mm]  Std_tz[mm]" << endl;

    auto addVCM = [](const Eigen::MatrixXd& a, const Eigen::MatrixXd& b) {
        Eigen::MatrixXd c(6, 6);
        for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) c(r, k) = a(r, k) + b(r, k);
        return c;
    };
    for (int i = 0; i < epochNum; i++) {
        Eigen::Matrix4f accT = Eigen::Matrix4f::Identity();
        Eigen::MatrixXd accV = Eigen::MatrixXd::Zero(6, 6);
        if (pairMode < 0) {
            accT = allT[i]; accV = allV[i];
            int target = i + 1, times = 1;
            for (int j = 0; j < i + 1; j++) {
                target = RegPair[target];
                if (target == 0) break;                            // reached the first epoch
                const Eigen::Matrix4f Mnew = allT[target - 1];
                accT = Mnew * accT;
                // rigorous propagation with the adjoint Ad = [[R, 0], [t^ R, R]] (:1072-1083)
                double R[3][3], Sx[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, tR[3][3];
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r][c] = (double)Mnew(r, c);
                const double tx = Mnew(0, 3), ty = Mnew(1, 3), tz = Mnew(2, 3);
                Sx[0][1] = -tz; Sx[0][2] = ty; Sx[1][0] = tz; Sx[1][2] = -tx; Sx[2][0] = -ty; Sx[2][1] = tx;
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += Sx[r][k] * R[k][c]; tR[r][c] = s; }
                double Ad[6][6] = {{0}};
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { Ad[r][c] = R[r][c]; Ad[r + 3][c + 3] = R[r][c]; Ad[r + 3][c] = tR[r][c]; }
                Eigen::MatrixXd tmp(6, 6), prop(6, 6);
                for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double s = 0; for (int k = 0; k < 6; ++k) s += Ad[r][k] * accV(k, c); tmp(r, c) = s; }
                for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double s = 0; for (int k = 0; k < 6; ++k) s += tmp(r, k) * Ad[c][k]; prop(r, c) = s; }
                accV = addVCM(allV[target - 1], prop);
                times++;
            }
            cout << "Source cloud- " << i + 1 << " is aligned " << times << " time(s)\n";
        } else if ((pairMode == 0) || (i < pairMode)) {
            accT = allT[i]; accV = allV[i];
        } else {                                                   // fixed interval (:1099-1106)
            for (int j = 0; j < epochNum; j++) {
                accT = allT[i - pairMode * j] * accT;
                accV = addVCM(allV[i - pairMode * j], accV);
                if (i - pairMode * j < pairMode) break;
            }
        }
        allTransMat2Ref.push_back(accT);
        allVCM2Ref.push_back(accV);
        outTM << fixed << setprecision(12);
        outTM << timeStamp[i] << "\n";
        for (int r = 0; r < 4; r++) { for (int c = 0; c < 4; c++) outTM << accT(r, c) << " "; outTM << endl; }
        for (int r = 0; r < 6; r++) { for (int c = 0; c < 6; c++) outTM << accV(r, c) << " "; outTM << endl; }
        outTP << fixed << setprecision(10);
        outTP << timeStamp[i] << " ";
        Eigen::Vector3f ang;
        matrix2angle(accT, ang);
        const float Rx = ang[0] * ARC_TO_GON, Ry = ang[1] * ARC_TO_GON, Rz = ang[2] * ARC_TO_GON;
        const float tx = accT(0, 3), ty = accT(1, 3), tz = accT(2, 3);
        outTP << Rx << " " << Ry << " " << Rz << " " << tx << " " << ty << " " << tz << " "
              << 1000 * sqrt(accV(0, 0)) * ARC_TO_GON << " " << 1000 * sqrt(accV(1, 1)) * ARC_TO_GON << " "
              << 1000 * sqrt(accV(2, 2)) * ARC_TO_GON << " " << 1000 * sqrt(accV(3, 3)) << " " << 1000 * sqrt(accV(4, 4)) << " "
              << 1000 * sqrt(accV(5, 5)) << endl;
    }
    outTM.close(); outTP.close();
}

// ---- absolute error against the ground truth (src/Registration.cpp:1157-1251) --------------------
void calAbsErrorOfTransPara(std::string transMatFile, std::string GTtransMatFile, int allEpochNum, int startEpoch,
                            std::string transParaErrorFile) {
    const int EpoNum = allEpochNum - startEpoch - 1;
    ifstream in1(transMatFile);
    if (!in1) fatal("Cannot open transMatFile!");
    std::vector<Eigen::Matrix4f> est;
    for (int i = 0; i < EpoNum; i++) {
        int t; in1 >> t;
        Eigen::Matrix4f M = Eigen::Matrix4f::Identity();
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) in1 >> M(r, c);
        double skip; for (int k = 0; k < 36; ++k) in1 >> skip;
        est.push_back(M);
    }
    in1.close();
    ifstream in2(GTtransMatFile);
    if (!in2) fatal("Cannot open GTtransMatFile!");
    std::vector<Eigen::Matrix4f> ref;
    for (int i = 0; i < allEpochNum; i++) {
        int t; in2 >> t;
        Eigen::Matrix4f M = Eigen::Matrix4f::Identity();
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) in2 >> M(r, c);
        ref.push_back(M);
    }
    in2.close();
    ofstream out(transParaErrorFile);
    if (!out) fatal("Cannot open transParaErrorFile!");
    out << "Err_Rx[mgon]  Err_Ry[mgon]  Err_Rz[mgon]  Err_tx[mm]  Err_ty[mm]  Err_tz[mm]" << endl;
    for (int i = 0; i < EpoNum; i++) {
        Eigen::Vector3f a, b;
        matrix2angle(est[i], a);
        matrix2angle(ref[startEpoch + 1 + i], b);
        const float Rx = a[0] * ARC_TO_GON, Ry = a[1] * ARC_TO_GON, Rz = a[2] * ARC_TO_GON;
        const float Rxr = b[0] * ARC_TO_GON, Ryr = b[1] * ARC_TO_GON, Rzr = b[2] * ARC_TO_GON;
        const float errRx = 1000 * fabs(Rxr - Rx), errRy = 1000 * fabs(Ryr - Ry), errRz = 1000 * fabs(Rzr - Rz);
        const float errtx = 1000 * fabs(ref[startEpoch + 1 + i](0, 3) - est[i](0, 3));
        const float errty = 1000 * fabs(ref[startEpoch + 1 + i](1, 3) - est[i](1, 3));
        const float errtz = 1000 * fabs(ref[startEpoch + 1 + i](2, 3) - est[i](2, 3));
        out << errRx << " " << errRy << " " << errRz << " " << errtx << " " << errty << " " << errtz << " " << endl;
    }
    out.close();
}
