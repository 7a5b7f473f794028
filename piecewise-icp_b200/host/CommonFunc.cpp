// CommonFunc.cpp -- host mirror of the reference's src/CommonFunc.cpp on top of libpwicp.so.
// Same function names, argument meaning and error behaviour; see CommonFunc.h.
#include "CommonFunc.h"
#include "msvc_sort.h"

#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <limits>
#include <sstream>
#include <unordered_map>

#include "../../include/pwicp.h"
#include "../csrc/patch_algebra.cuh"

using namespace std;

// ---- device context of the mirror --------------------------------------------------------------
static pwicp_ctx* g_ctx = nullptr;
static int g_device = -1;

void pwicpHostSetDevice(int device) {
    if (g_ctx && device != g_device) { pwicp_ctx_destroy(g_ctx); g_ctx = nullptr; }
    g_device = device;
}

pwicp_ctx* pwicpHostContext() {
    if (g_ctx) return g_ctx;
    int dev = g_device;
    if (dev < 0) {
        const char* e = getenv("PWICP_DEVICE");
        if (!e) e = getenv("LOCAL_RANK");
        dev = e ? atoi(e) : 0;
    }
    if (pwicp_ctx_create(dev, &g_ctx) != PWICP_OK) {
        // no CPU fallback: the reference-shaped functions cannot work without the device library
        cerr << "Error: cannot create the CUDA context: " << pwicp_last_error(nullptr) << " Aborting.\n";
        exit(EXIT_FAILURE);
    }
    g_device = dev;
    return g_ctx;
}

static vector<float> packXYZ(const pcl::PointCloud<pcl::PointXYZ>& c) {
    vector<float> v(3 * c.size());
    for (size_t i = 0; i < c.size(); ++i) { v[3 * i] = c.points[i].x; v[3 * i + 1] = c.points[i].y; v[3 * i + 2] = c.points[i].z; }
    return v;
}

// ---- Eigen / pcl shim bodies -------------------------------------------------------------------
namespace Eigen {
Matrix4f operator*(const Matrix4f& a, const Matrix4f& b) {
    Matrix4f c;
    pwicp_mat4_mul(a.m, b.m, c.m);
    return c;
}
}  // namespace Eigen

namespace pcl {

void transformPointCloud(const PointCloud<PointXYZ>& in, PointCloud<PointXYZ>& out, const Eigen::Matrix4f& T) {
    PointCloud<PointXYZ> res;
    res.resize(in.size());
    const float* m = T.m;
    for (size_t i = 0; i < in.size(); ++i) {
        const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
        res.points[i].x = m[0] * x + m[1] * y + m[2] * z + m[3];
        res.points[i].y = m[4] * x + m[5] * y + m[6] * z + m[7];
        res.points[i].z = m[8] * x + m[9] * y + m[10] * z + m[11];
    }
    out = res;
}

unsigned compute3DCentroid(const PointCloud<PointXYZ>& cloud, Eigen::Vector4f& centroid) {
    float sx = 0, sy = 0, sz = 0;
    for (const auto& p : cloud.points) { sx += p.x; sy += p.y; sz += p.z; }
    const float n = (float)cloud.size();
    centroid[0] = sx / n; centroid[1] = sy / n; centroid[2] = sz / n; centroid[3] = 1;
    return (unsigned)cloud.size();
}

namespace io {

int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud) {
    cloud.clear();
    ifstream f(path, ios::binary);
    if (!f) { cerr << "[pcl::io::loadPCDFile] cannot open " << path << "\n"; return -1; }
    vector<string> fields; vector<int> sizes, counts; vector<char> types;
    size_t npts = 0; string data;
    string line;
    while (getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        istringstream ls(line);
        string key; ls >> key;
        if (key == "FIELDS") { string s; while (ls >> s) fields.push_back(s); }
        else if (key == "SIZE") { int v; while (ls >> v) sizes.push_back(v); }
        else if (key == "TYPE") { char c; while (ls >> c) types.push_back(c); }
        else if (key == "COUNT") { int v; while (ls >> v) counts.push_back(v); }
        else if (key == "POINTS") { ls >> npts; }
        else if (key == "WIDTH") { size_t w; ls >> w; if (!npts) npts = w; }
        else if (key == "DATA") { ls >> data; break; }
    }
    if (counts.empty()) counts.assign(fields.size(), 1);
    if (fields.size() != sizes.size() || fields.size() != types.size()) { cerr << "[pcl::io::loadPCDFile] bad header\n"; return -1; }
    int off[3] = {-1, -1, -1}, col[3] = {-1, -1, -1}, stride = 0, ncol = 0;
    for (size_t k = 0; k < fields.size(); ++k) {
        for (int a = 0; a < 3; ++a)
            if (fields[k] == (a == 0 ? "x" : a == 1 ? "y" : "z")) {
                if (types[k] != 'F' || sizes[k] != 4) { cerr << "[pcl::io::loadPCDFile] x/y/z must be float32\n"; return -1; }
                off[a] = stride; col[a] = ncol;
            }
        stride += sizes[k] * counts[k];
        ncol += counts[k];
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0) { cerr << "[pcl::io::loadPCDFile] no x y z fields\n"; return -1; }
    cloud.resize(npts);
    if (data == "binary") {
        vector<char> buf((size_t)stride * npts);
        f.read(buf.data(), (streamsize)buf.size());
        if ((size_t)f.gcount() != buf.size()) { cerr << "[pcl::io::loadPCDFile] truncated file\n"; cloud.clear(); return -1; }
        for (size_t i = 0; i < npts; ++i) {
            const char* rec = buf.data() + i * stride;
            memcpy(&cloud.points[i].x, rec + off[0], 4);
            memcpy(&cloud.points[i].y, rec + off[1], 4);
            memcpy(&cloud.points[i].z, rec + off[2], 4);
        }
    } else if (data == "ascii") {
        for (size_t i = 0; i < npts; ++i) {
            if (!getline(f, line)) { cloud.resize(i); break; }
            istringstream ls(line);
            string tok; int c = 0;
            while (ls >> tok) {
                for (int a = 0; a < 3; ++a)
                    if (c == col[a]) (a == 0 ? cloud.points[i].x : a == 1 ? cloud.points[i].y : cloud.points[i].z) = strtof(tok.c_str(), nullptr);
                ++c;
            }
        }
    } else {
        cerr << "[pcl::io::loadPCDFile] unsupported DATA " << data << "\n";
        cloud.clear();
        return -1;
    }
    return 0;
}

int savePCDFileBinary(const std::string& path, const PointCloud<PointXYZ>& cloud) {
    ofstream f(path, ios::binary);
    if (!f) return -1;
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
      << "WIDTH " << cloud.size() << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << cloud.size() << "\nDATA binary\n";
    for (const auto& p : cloud.points) f.write(reinterpret_cast<const char*>(&p.x), 12);
    return f ? 0 : -1;
}

}  // namespace io
}  // namespace pcl

// ---- configuration file (src/CommonFunc.cpp:11-136) --------------------------------------------
// Eleven getline()s in fixed order; the value is the text after the first ':' (+2 for the two path
// lines, +1 for the numeric ones); labels are ignored; range checks return false.
static string valueAfterColon(string line, int skip) {
    if (!line.empty() && line.back() == '\r') line.pop_back();     // the shipped files use CRLF
    size_t p = line.find(":");
    if (p == string::npos || p + skip > line.size()) return string();
    return line.substr(p + skip);
}

bool readConfigFile(std::string conFile, ConfigPara& confPara) {
    ifstream in;
    in.open(conFile);
    if (!in || conFile.empty()) {
        std::cerr << "Cannot open configuration file! Aborting.\n";
        return false;
    }
    string s;
    auto next = [&](int skip, string& out) -> bool {
        s.clear();
        getline(in, s);
        if (!s.empty() && s.back() == '\r') s.pop_back();
        if (s.empty()) return false;
        out = valueAfterColon(s, skip);
        return true;
    };
    string v;
    if (next(2, v)) { confPara.FolderFilePath1 = v; cout << "1 FolderFilePath1: *" << confPara.FolderFilePath1 << "*" << endl; }
    if (next(2, v)) { confPara.FolderFilePath2 = v; cout << "2 FolderFilePath2: *" << confPara.FolderFilePath2 << "*" << endl; }
    if (next(1, v)) { confPara.isSetResSVsize = std::stoi(v); cout << "3 isSetResSVsize: *" << confPara.isSetResSVsize << "*" << endl; }
    if (next(1, v)) { confPara.PCres1 = std::stof(v); cout << "4 PCres1: *" << confPara.PCres1 << "*" << endl; }
    if (confPara.PCres1 <= 0) { std::cerr << "PCres1 out of limits! \n" << endl; return false; }
    if (next(1, v)) { confPara.PCres2 = std::stof(v); cout << "5 PCres2: *" << confPara.PCres2 << "*" << endl; }
    if (confPara.PCres2 <= 0) { std::cerr << "PCres2 out of limits! \n" << endl; return false; }
    if (next(1, v)) { confPara.SVsize1 = std::stof(v); cout << "6 SVsize1: *" << confPara.SVsize1 << "*" << endl; }
    if (confPara.SVsize1 < confPara.PCres1 || confPara.SVsize1 > 40 * confPara.PCres1) { std::cerr << "SVsize1 out of limits! \n" << endl; return false; }
    if (next(1, v)) { confPara.SVsize2 = std::stof(v); cout << "7 SVsize2: *" << confPara.SVsize2 << "*" << endl; }
    if (confPara.SVsize2 < confPara.PCres2 || confPara.SVsize2 > 40 * confPara.PCres2) { std::cerr << "SVsize2 out of limits! \n" << endl; return false; }
    if (next(1, v)) { confPara.isSetDTinit = std::stoi(v); cout << "8 isSetDTinit: *" << confPara.isSetDTinit << "*" << endl; }
    if (next(1, v)) { confPara.DTinit = std::stof(v); cout << "9 DTinit: *" << confPara.DTinit << "*" << endl; }
    if (confPara.DTinit <= 0) { std::cerr << "DTinit out of limits! \n" << endl; return false; }
    if (next(1, v)) { confPara.DTmin = std::stof(v); cout << "10 DisThrhdmin: *" << confPara.DTmin << "*" << endl; }
    if (confPara.DTinit < confPara.DTmin) { std::cerr << "DTmin out of limits! \n" << endl; return false; }
    if (next(1, v)) { confPara.isVisual = std::stoi(v); cout << "11 isVisual: *" << confPara.isVisual << "*" << endl; }
    cout << endl << endl;
    in.close();
    return true;
}

// ---- file listing (src/CommonFunc.cpp:182-236) --------------------------------------------------
void getFiles(std::string folderpath, std::vector<std::string>& files) {
    DIR* d = opendir(folderpath.c_str());
    if (!d) return;
    vector<string> names;
    while (dirent* e = readdir(d)) names.push_back(e->d_name);
    closedir(d);
    sort(names.begin(), names.end());
    for (const string& n : names) {
        if (n == "." || n == "..") continue;
        string full = folderpath + "/" + n;
        struct stat st;
        if (stat(full.c_str(), &st) != 0) continue;
        if (S_ISDIR(st.st_mode)) getFiles(full, files);
        else files.push_back(full);
    }
}

long extractTimeFromFileName(std::string fileName, const std::string substring, int timeLength) {
    int startpos = (int)fileName.find(substring) + (int)substring.size();
    string alltime = fileName.substr(startpos, timeLength);
    return stol(alltime);
}

int extractAllFilesFromFolder(std::string folderPath, std::vector<std::string>& fileNameList, std::vector<long>& fileTimeList) {
    fileNameList.clear(); fileTimeList.clear();
    while (folderPath.size() > 1 && folderPath.back() == '/') folderPath.pop_back();
    vector<string> all;
    getFiles(folderPath, all);
    cout << "--->>> " << all.size() << " files found in folder: " << folderPath << "\n";
    vector<pair<string, long>> byTime;
    for (const string& f : all) byTime.emplace_back(f, extractTimeFromFileName(f, "Epoch_", 3));
    // the reference goes through a std::map keyed by name, then sorts by time
    sort(byTime.begin(), byTime.end(), [](const pair<string, long>& a, const pair<string, long>& b) { return a.first < b.first; });
    stable_sort(byTime.begin(), byTime.end(), [](const pair<string, long>& a, const pair<string, long>& b) { return a.second < b.second; });
    for (auto& p : byTime) { fileNameList.push_back(p.first); fileTimeList.push_back(p.second); }
    return (int)fileNameList.size();
}

// ---- NN-based helpers served by the device ------------------------------------------------------
float calPCresolution(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud) {
    const int n = (int)cloud->size();
    if (n < 2) { std::cerr << "Error: Number of neighbor is 0! \n\n"; return 0.0; }
    vector<float> xyz = packXYZ(*cloud), d2(n);
    if (pwicp_self_nn(pwicpHostContext(), xyz.data(), n, d2.data()) != PWICP_OK) {
        std::cerr << "Error: " << pwicp_last_error(pwicpHostContext()) << "\n\n";
        return 0.0;
    }
    float res = 0.0;                                  // sequential float sum, like the reference
    for (int i = 0; i < n; ++i) res += sqrt(d2[i]);
    res /= n;
    return res;
}

double calPercentileDistBetween2PC(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud1, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud2, float percentile) {
    vector<float> a = packXYZ(*cloud1), b = packXYZ(*cloud2);
    double out = 0;
    if (pwicp_percentile_nn(pwicpHostContext(), a.data(), (int)cloud1->size(), b.data(), (int)cloud2->size(), percentile, &out) != PWICP_OK) {
        std::cerr << "Error: " << pwicp_last_error(pwicpHostContext()) << " Aborting.\n";
        std::exit(EXIT_FAILURE);
    }
    return out;
}

// ---- patch normals (src/CommonFunc.cpp:284-333; pcl::computePointNormal + pcl::eigen33) --------
// symmetric 3x3 eigen decomposition (cyclic Jacobi, double): eigenvalues ascending, columns of V
void pwicpJacobi3(double A[3][3], double w[3], double V[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
                for (int k = 0; k < 3; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
                for (int k = 0; k < 3; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
            }
    }
    int idx[3] = {0, 1, 2};
    sort(idx, idx + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
    double Vs[3][3];
    for (int k = 0; k < 3; ++k) { w[k] = A[idx[k]][idx[k]]; for (int r = 0; r < 3; ++r) Vs[r][k] = V[r][idx[k]]; }
    memcpy(V, Vs, sizeof(Vs));
}

// The arithmetic of calPatchNormal / calPatchSTD is csrc/patch_algebra.cuh, shared with the batched
// device kernel (pwicp_patch_stats); these are the reference-shaped single-patch entry points.
bool calPatchNormal(pcl::PointCloud<pcl::PointXYZ> cloud, float& nx, float& ny, float& nz) {
    if (!(cloud.size() > 4)) {
        std::cerr << "Patch normal calculation fails !!! Assigned with (0, 0, 1) \n\n";
        nx = 0; ny = 0; nz = 1;
        return false;
    }
    const vector<float> xyz = packXYZ(cloud);
    float n3[3];
    const bool ok = pwicp::pa_patch_normal(xyz.data(), (int)cloud.size(), n3) != 0;
    nx = n3[0]; ny = n3[1]; nz = n3[2];
    if (!ok) std::cerr << "Incalculable normals: " << nx << ", " << ny << ", " << nz << " !!! \n\n";
    return ok;
}

// plane through the centroid with the smallest-eigenvalue direction, std of the point-to-plane
// distances with (n - 1) (src/CommonFunc.cpp:336-354; pcl::PCA replaced by a Jacobi solve)
float calPatchSTD(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud) {
    const vector<float> xyz = packXYZ(*cloud);
    return pwicp::pa_patch_std(xyz.data(), (int)cloud->size());
}

void generateCentroidCloudWithPatchNormals(pcl::PointCloud<pcl::PointXYZ>::Ptr cloudCentroids,
                                           pcl::PointCloud<pcl::PointXYZ>* cloudPatch,
                                           pcl::PointCloud<pcl::PointNormal>::Ptr cloudCentroids_normals) {
    cloudCentroids_normals->clear();
    cloudCentroids_normals->resize(cloudCentroids->size());
    for (size_t i = 0; i < cloudCentroids->size(); ++i) {
        float nx = 0.0f, ny = 0.0f, nz = 1.0f;
        pcl::PointNormal& o = cloudCentroids_normals->points[i];
        o.x = cloudCentroids->points[i].x; o.y = cloudCentroids->points[i].y; o.z = cloudCentroids->points[i].z;
        if (cloudPatch[i].size() > 6 && calPatchNormal(cloudPatch[i], nx, ny, nz)) {
            o.normal_x = nx; o.normal_y = ny; o.normal_z = nz;
        } else {
            std::cerr << "[Warning] Failed to estimate normal for patch " << i << ". Assigning default normal (0, 0, 1).\n";
            o.normal_x = 0.0f; o.normal_y = 0.0f; o.normal_z = 1.0f;
        }
    }
}

void matrix2angle(Eigen::Matrix4f transMat, Eigen::Vector3f& rotAngle) {
    float a[3];
    pwicp_matrix2angle(transMat.m, a);
    rotAngle[0] = a[0]; rotAngle[1] = a[1]; rotAngle[2] = a[2];
}

float calBoundingBoxCornerChange(const double* boundingBox, const Eigen::Matrix4f transMat) {
    return pwicp_bbox_corner_change(boundingBox, transMat.m);
}

// ---- pre-processing -------------------------------------------------
// PCpreprocessing / SORfilter run on the device (pwicp_preprocess: csrc/prep.cu).  The plain host statements of the
// same two PCL filters below (PCpreprocessingHost / SORfilterHost) are kept for tools that must work without a
// device (c_hooks.cpp: pwicp_host_prepare_pair, the CPU pinning harness); both give identical clouds
// (tests/test_host_gpu.py).
static void voxelGrid(const pcl::PointCloud<pcl::PointXYZ>& in, float leaf, pcl::PointCloud<pcl::PointXYZ>& out) {
    out.clear();
    if (in.empty()) return;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const auto& p : in.points) {
        mn[0] = min(mn[0], p.x); mn[1] = min(mn[1], p.y); mn[2] = min(mn[2], p.z);
        mx[0] = max(mx[0], p.x); mx[1] = max(mx[1], p.y); mx[2] = max(mx[2], p.z);
    }
    const float inv = 1.0f / leaf;
    long long minb[3], div[3];
    for (int c = 0; c < 3; ++c) { minb[c] = (long long)floor(mn[c] * inv); div[c] = (long long)floor(mx[c] * inv) - minb[c] + 1; }
    vector<pair<long long, int>> keyed(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
        const auto& p = in.points[i];
        const long long ix = (long long)floor(p.x * inv) - minb[0], iy = (long long)floor(p.y * inv) - minb[1], iz = (long long)floor(p.z * inv) - minb[2];
        keyed[i] = {ix + iy * div[0] + iz * div[0] * div[1], (int)i};
    }
    // order of the points inside a voxel = order of the float sums: input order (what a stable sort gives, and what the
    // device path does), or -- PWICP_VOXEL_ORDER=msvc -- the order the reference's Windows build produced (msvc_sort.h)
    const char* vo = getenv("PWICP_VOXEL_ORDER");
    if (vo && string(vo) == "msvc")
        msvc::sort(keyed.begin(), keyed.end(), [](const pair<long long, int>& a, const pair<long long, int>& b) { return a.first < b.first; });
    else sort(keyed.begin(), keyed.end());
    size_t i = 0;
    while (i < keyed.size()) {
        size_t j = i;
        float sx = 0, sy = 0, sz = 0;
        while (j < keyed.size() && keyed[j].first == keyed[i].first) {
            const auto& p = in.points[keyed[j].second];
            sx += p.x; sy += p.y; sz += p.z; ++j;
        }
        const float n = (float)(j - i);
        out.push_back(pcl::PointXYZ(sx / n, sy / n, sz / n));
        i = j;
    }
}

void SORfilterHost(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
                   int SOR_NeighborNum, double SOR_StdMult) {
    const auto& pts = cloud_in->points;
    const int n = (int)pts.size(), k = SOR_NeighborNum;
    cloud_out->clear();
    if (n <= k) { *cloud_out = *cloud_in; return; }
    // host uniform grid sized for ~k points per 3x3x3 block
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const auto& p : pts) {
        mn[0] = min(mn[0], p.x); mn[1] = min(mn[1], p.y); mn[2] = min(mn[2], p.z);
        mx[0] = max(mx[0], p.x); mx[1] = max(mx[1], p.y); mx[2] = max(mx[2], p.z);
    }
    const double ext[3] = {mx[0] - mn[0] + 1e-6, mx[1] - mn[1] + 1e-6, mx[2] - mn[2] + 1e-6};
    double area = max(ext[0] * ext[1], max(ext[0] * ext[2], ext[1] * ext[2]));
    const double h = max(sqrt(area / n) * 2.0, 1e-9);
    auto cellOf = [&](float v, int c) { return (long long)floor((v - mn[c]) / h); };
    unordered_map<long long, vector<int>> grid;
    grid.reserve(n);
    auto key = [](long long x, long long y, long long z) { return (x * 73856093LL) ^ (y * 19349663LL) ^ (z * 83492791LL); };
    for (int i = 0; i < n; ++i) grid[key(cellOf(pts[i].x, 0), cellOf(pts[i].y, 1), cellOf(pts[i].z, 2))].push_back(i);
    vector<float> meanDist(n);
    vector<float> cand;
    for (int i = 0; i < n; ++i) {
        const long long cx = cellOf(pts[i].x, 0), cy = cellOf(pts[i].y, 1), cz = cellOf(pts[i].z, 2);
        for (int r = 1;; ++r) {
            cand.clear();
            for (long long z = cz - r; z <= cz + r; ++z)
                for (long long y = cy - r; y <= cy + r; ++y)
                    for (long long x = cx - r; x <= cx + r; ++x) {
                        auto it = grid.find(key(x, y, z));
                        if (it == grid.end()) continue;
                        for (int j : it->second) {
                            if (cellOf(pts[j].x, 0) != x || cellOf(pts[j].y, 1) != y || cellOf(pts[j].z, 2) != z) continue;  // hash collision
                            if (j == i) continue;
                            const float dx = pts[j].x - pts[i].x, dy = pts[j].y - pts[i].y, dz = pts[j].z - pts[i].z;
                            cand.push_back(dx * dx + dy * dy + dz * dz);
                        }
                    }
            if ((int)cand.size() >= k) {
                nth_element(cand.begin(), cand.begin() + (k - 1), cand.end());
                if (std::sqrt(cand[k - 1]) <= r * h || r > 64) break;
            } else if (r > 64) break;
        }
        sort(cand.begin(), cand.end());
        double s = 0;
        const int kk = min(k, (int)cand.size());
        for (int j = 0; j < kk; ++j) s += std::sqrt(cand[j]);
        meanDist[i] = kk ? (float)(s / kk) : 0.f;
    }
    double sum = 0, sq = 0;
    // [PCL 1.8.1 StatisticalOutlierRemoval::applyFilterIndices] distances is a vector<float>: the square is a float
    // product, rounded before it is widened; no clamp of the variance; a point is an outlier iff distance > threshold
    for (float d : meanDist) { const float d_sq = d * d; sum += d; sq += d_sq; }
    const double mean = sum / n;
    const double var = (sq - sum * sum / n) / (n - 1);
    const double thr = mean + SOR_StdMult * sqrt(var);
    for (int i = 0; i < n; ++i) if (!(meanDist[i] > thr)) cloud_out->push_back(pts[i]);
}

void PCpreprocessingHost(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
                         bool isDownSamp, float voxelSize, int SOR_NeighborNum, double SOR_StdMult) {
    pcl::PointCloud<pcl::PointXYZ>::Ptr filtered(new pcl::PointCloud<pcl::PointXYZ>);
    if (isDownSamp) voxelGrid(*cloud_in, voxelSize, *filtered);
    else pcl::copyPointCloud(*cloud_in, *filtered);
    SORfilterHost(filtered, cloud_out, SOR_NeighborNum, SOR_StdMult);
}

// src/CommonFunc.cpp:423-439 and :442-452 on the device
static void preprocessDevice(const pcl::PointCloud<pcl::PointXYZ>& in, pcl::PointCloud<pcl::PointXYZ>& out,
                             bool isDownSamp, float voxelSize, int k, double mult) {
    const int n = (int)in.size();
    out.clear();
    if (n < 1) return;
    vector<float> xyz(3 * (size_t)n), res(3 * (size_t)n);
    for (int i = 0; i < n; ++i) { xyz[3 * i] = in.points[i].x; xyz[3 * i + 1] = in.points[i].y; xyz[3 * i + 2] = in.points[i].z; }
    int m = 0;
    if (pwicp_preprocess(pwicpHostContext(), xyz.data(), n, isDownSamp ? 1 : 0, voxelSize, k, mult, res.data(), &m) != PWICP_OK) {
        std::cerr << "Error: pre-processing failed on the device: " << pwicp_last_error(pwicpHostContext()) << "\n";
        std::exit(EXIT_FAILURE);
    }
    out.resize(m);
    for (int i = 0; i < m; ++i) { out.points[i].x = res[3 * i]; out.points[i].y = res[3 * i + 1]; out.points[i].z = res[3 * i + 2]; }
}

void SORfilter(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
               int SOR_NeighborNum, double SOR_StdMult) {
    pcl::PointCloud<pcl::PointXYZ> res;
    preprocessDevice(*cloud_in, res, false, 0.f, SOR_NeighborNum, SOR_StdMult);
    *cloud_out = res;
}

void PCpreprocessing(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out,
                     bool isDownSamp, float voxelSize, int SOR_NeighborNum, double SOR_StdMult) {
    pcl::PointCloud<pcl::PointXYZ> res;
    const char* vo = getenv("PWICP_VOXEL_ORDER");
    if (isDownSamp && vo && string(vo) == "msvc") {
        // the reference's Windows build: the voxel centroids in the Microsoft std::sort order (host, msvc_sort.h), then
        // the outlier removal on the device as usual
        pcl::PointCloud<pcl::PointXYZ> vox;
        voxelGrid(*cloud_in, voxelSize, vox);
        preprocessDevice(vox, res, false, 0.f, SOR_NeighborNum, SOR_StdMult);
    } else {
        preprocessDevice(*cloud_in, res, isDownSamp, voxelSize, SOR_NeighborNum, SOR_StdMult);
    }
    *cloud_out = res;
}

void visualizeTwoPC(pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int,
                    pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int, double, double, double) {}
void visualizeThreePC(pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int,
                      pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int,
                      pcl::PointCloud<pcl::PointXYZ>::Ptr, std::string, double, double, double, int, double, double, double) {}
