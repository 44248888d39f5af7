// extern "C" facade over the host classes, for the ctypes-driven tests and tools (tests/test_host_classes.py).
// Not part of the drop-in boundary (that is include/hkf.h); it only lets Python drive the C++ classes.
#include <cstring>
#include <string>
#include "AppParamsProducer.h"
#include "CameraPoseFinderFromFile.h"
#include "CudaDeviceDataMan.h"
#include "DataSourceProducerRGBDDataset.h"
#include "DataSourceProducerSynthetic.h"
#include "PngReader.h"
#include "MainController.h"
#include "TrajectoryRecorder.h"
#include "../csrc/hkf_solver.h"

static std::string g_err;

extern "C" {

const char* hkfhost_last_error() { return g_err.c_str(); }

// ---- parameters
int hkfhost_read_params(const char* path) { return AppParamsProducer::readParams(path) ? 0 : 1; }
const char* hkfhost_missing_keys() { static std::string s; s = AppParamsProducer::lastMissingKeys(); return s.c_str(); }
void hkfhost_get_frame_params(HkfFrameParams* fp, HkfInitParams* ip)
{
    if (fp) *fp = AppParams::instance()->frameParams();
    if (ip) *ip = AppParams::instance()->initParams();
}
int hkfhost_switches()
{
    const SwitchParams& s = AppParams::instance()->_switch_params;
    return (s.useRGBData ? 1 : 0) | (s.useTrajFromFile ? 2 : 0) | (s.useSdfTracker ? 4 : 0) | (s.recordTrajectory ? 8 : 0) |
           (s.useSyntheticData ? 16 : 0) | (s.stepwiseTracking ? 32 : 0);
}

// ---- synthetic producer (pure host)
void hkfhost_synth_pose(unsigned i, unsigned n, double phase, double volume_size, double trunc_min, double* out16)
{
    DataSourceProducerSynthetic::trajectoryPose(i, n, phase, volume_size, trunc_min, out16);
}
void hkfhost_synth_render(const double* pose16, unsigned cols, unsigned rows, float cx, float cy, float fx, float fy, uint16_t* out)
{
    CameraParams cam; cam.cols = cols; cam.rows = rows; cam.cx = cx; cam.cy = cy; cam.fx = fx; cam.fy = fy;
    DataSourceProducerSynthetic::renderDepthMm(pose16, cam, out);
}
void hkfhost_synth_render_color(const double* pose16, unsigned cols, unsigned rows, float cx, float cy, float fx, float fy, uint8_t* out)
{
    CameraParams cam; cam.cols = cols; cam.rows = rows; cam.cx = cx; cam.cy = cy; cam.fx = fx; cam.fy = fy;
    DataSourceProducerSynthetic::renderColor(pose16, cam, out);
}
int hkfhost_write_tum(const char* path, unsigned n, double phase, double volume_size, double trunc_min)
{
    return DataSourceProducerSynthetic::writeTumTrajectory(path, n, phase, volume_size, trunc_min) ? 0 : 1;
}

// ---- TUM dataset reader (pure host): read up to n frames of the dataset at `dir` (with trailing slash) at the
// configured depth camera size; returns the number of frames read, timestamps in ts
namespace { struct OpenDataset : DataSourceProducerRGBDDataset { using DataSourceProducer::init; }; }
int hkfhost_read_dataset(const char* dir, unsigned cols, unsigned rows, unsigned n, uint16_t* out, double* ts)
{
    AppParams* P = AppParams::instance();
    P->_io_params.rgbdReadFilename = dir;
    P->_switch_params.useRGBData = false; P->_switch_params.recordRGBD = false;
    P->_depth_camera_params.cols = cols; P->_depth_camera_params.rows = rows;
    OpenDataset src;
    if (!src.init()) { g_err = src.lastError(); return -1; }
    unsigned k = 0;
    for (; k < n; ++k) {
        DepthFrameData d; ColorFrameData c;
        if (!src.readNewFrame(d, c)) break;
        if ((unsigned)d.cols() != cols || (unsigned)d.rows() != rows) { g_err = "size"; return -2; }
        std::memcpy(out + (size_t)k * cols * rows, d.mm(), (size_t)cols * rows * 2);
        ts[k] = d.timeStamp();
    }
    if (k < n) g_err = src.lastError();
    return (int)k;
}
int hkfhost_read_png(const char* path, unsigned* w, unsigned* h, unsigned* ch, unsigned* depth, uint16_t* out, size_t cap)
{
    PngImage img; std::string e;
    if (!readPng(path, img, &e)) { g_err = e; return 1; }
    *w = img.width; *h = img.height; *ch = img.channels; *depth = img.bit_depth;
    if (out && cap >= img.samples.size()) std::memcpy(out, img.samples.data(), img.samples.size() * 2);
    return 0;
}

// ---- trajectory file round trip (pure host): record poses, then replay them through CameraPoseFinderFromFile
int hkfhost_record_trajectory(const char* path, const float* poses16, const double* timestamps, unsigned n)
{
    TrajectoryRecorder rec(path);
    for (unsigned i = 0; i < n; ++i) {
        Mat44 m; std::memcpy(m.m, poses16 + 16 * i, 64);
        if (!rec.recordCameraPose(m, timestamps[i])) return 1;
    }
    return 0;
}
int hkfhost_replay_trajectory(const char* path, const float* pose0_16, const double* timestamps, unsigned n, float* poses_out16)
{
    AppParams::instance()->_io_params.trajReadFilename = path;
    CameraPoseFinderFromFile f;
    Mat44 p0; std::memcpy(p0.m, pose0_16, 64);
    if (!f.init(p0)) { g_err = "cannot open trajectory"; return 1; }
    ColorFrameData rgb;
    for (unsigned i = 0; i < n; ++i) {
        DepthFrameData d(i);
        uint16_t px = 0;
        d.setFrameData(&px, 1, 1, timestamps[i]);
        if (!f.findCameraPose(d, rgb)) { g_err = "no record for timestamp"; return 2; }
        std::memcpy(poses_out16 + 16 * i, f.getCameraPose().m, 64);
    }
    return 0;
}

// ---- the host-side solves of the two trackers as the stepwise host classes run them (csrc/hkf_solver.h compiled for
// the host): exposed so that the CPU tests can pin them against the oracle's independent restatement (oracle/host_math.h)
int hkfhost_icp_update(const float* sums27, float angle_shake, float dist_shake, float* cur16) { return hkf_icp_update(sums27, angle_shake, dist_shake, cur16); }
int hkfhost_sdf_update(const float* sums27, float angle_shake, float dist_shake, float* cur16, int* converged) { return hkf_sdf_update(sums27, angle_shake, dist_shake, cur16, converged); }
void hkfhost_llt_solve6(const float* sums27, float* x6) { float A[36], b[6]; hkf_unpack27(sums27, A, b); hkf_llt_solve6(A, b, x6); }
float hkfhost_det6(const float* sums27) { float A[36], b[6]; hkf_unpack27(sums27, A, b); return hkf_det6(A); }

// ---- mesh post-processing (pure host)
int hkfhost_weld(const HkfTriangle* soup, unsigned n, const char* obj_path, unsigned* n_vertices, unsigned* n_faces)
{
    std::vector<HkfTriangle> s(soup, soup + n);
    std::vector<std::array<float, 3>> v;
    std::vector<std::array<unsigned, 3>> f;
    MeshGeneratorMarchingcube::weld(s, v, f);
    if (n_vertices) *n_vertices = (unsigned)v.size();
    if (n_faces) *n_faces = (unsigned)f.size();
    if (obj_path && *obj_path) {
        std::vector<std::array<float, 3>> nrm;
        MeshGeneratorMarchingcube::vertexNormals(v, f, nrm);
        return MeshGeneratorMarchingcube::writeMesh(obj_path, v, f, nrm) ? 0 : 1;
    }
    return 0;
}

// the indexed mesh itself (for the comparison with the reference's ml::MeshData, tests/test_mesh_reference.py): arrays sized
// by the caller for 3 n vertices / n faces
int hkfhost_weld_arrays(const HkfTriangle* soup, unsigned n, int use_color, unsigned* n_vertices, unsigned* n_faces,
                        float* vertices3, float* normals3, float* colors4, unsigned* faces3)
{
    std::vector<HkfTriangle> s(soup, soup + n);
    std::vector<std::array<float, 3>> v, nrm;
    std::vector<std::array<unsigned, 3>> f;
    std::vector<std::array<float, 4>> col;
    MeshGeneratorMarchingcube::weld(s, v, f, use_color ? &col : nullptr);
    MeshGeneratorMarchingcube::vertexNormals(v, f, nrm);
    *n_vertices = (unsigned)v.size(); *n_faces = (unsigned)f.size();
    for (size_t i = 0; i < v.size(); ++i) {
        std::memcpy(vertices3 + 3 * i, v[i].data(), 12);
        std::memcpy(normals3 + 3 * i, nrm[i].data(), 12);
        if (use_color) std::memcpy(colors4 + 4 * i, col[i].data(), 16);
    }
    for (size_t i = 0; i < f.size(); ++i) std::memcpy(faces3 + 3 * i, f[i].data(), 12);
    return 0;
}

// ---- the whole pipeline through MainController / HybKinectfu (needs a B200)
int hkfhost_init(const char* config_path, int device)
{
    if (!MainController::instance()->init(config_path, device)) { g_err = MainController::instance()->lastError(); return 1; }
    return 0;
}
// process one externally supplied frame through HybKinectfu::processNewFrame
static int process_frame(const uint16_t* depth_mm, const uint8_t* bgr, unsigned cols, unsigned rows, unsigned frame_id, double timestamp, float* pose_out16, int* status);
int hkfhost_process(const uint16_t* depth_mm, unsigned cols, unsigned rows, unsigned frame_id, double timestamp, float* pose_out16, int* status)
{
    return process_frame(depth_mm, nullptr, cols, rows, frame_id, timestamp, pose_out16, status);
}
// the same with a colour frame of the rgb camera's size (use_color=1)
int hkfhost_process_color(const uint16_t* depth_mm, const uint8_t* bgr, unsigned cols, unsigned rows, unsigned frame_id, double timestamp, float* pose_out16, int* status)
{
    return process_frame(depth_mm, bgr, cols, rows, frame_id, timestamp, pose_out16, status);
}
static int process_frame(const uint16_t* depth_mm, const uint8_t* bgr, unsigned cols, unsigned rows, unsigned frame_id, double timestamp, float* pose_out16, int* status)
{
    HybKinectfu* k = MainController::instance()->fusioner();
    if (!k) { g_err = "not initialised"; return 1; }
    DepthFrameData d(frame_id);
    ColorFrameData rgb(frame_id);
    d.setFrameData(depth_mm, (int)cols, (int)rows, timestamp);
    const CameraParams& rc = AppParams::instance()->_rgb_camera_params;
    if (bgr) rgb.setFrameData(bgr, (int)rc.cols, (int)rc.rows, timestamp); else rgb.setFrameData(nullptr, 0, 0, timestamp);
    if (!k->processNewFrame(d, rgb)) { g_err = std::string("processNewFrame: ") + hkf_last_error(CudaDeviceDataMan::instance()->ctx()); return 2; }
    if (pose_out16) std::memcpy(pose_out16, k->cameraPose().m, 64);
    if (status) *status = k->lastTrackingStatus();
    return 0;
}
// pull `n` frames from the configured data source (MainController::mainLoop)
int hkfhost_run(unsigned n, int save_mesh) { return MainController::instance()->mainLoop(n, save_mesh != 0, false); }
int hkfhost_mesh(const char* obj_path, unsigned* n_tri, unsigned* n_vertices, unsigned* n_faces)
{
    MeshGeneratorMarchingcube* m = MainController::instance()->meshGenerator();
    if (!m) return 1;
    m->generateMesh();
    if (!m->saveMesh(obj_path)) return 2;
    if (n_tri) *n_tri = m->triangleCount();
    if (n_vertices) *n_vertices = (unsigned)m->vertexCount();
    if (n_faces) *n_faces = (unsigned)m->faceCount();
    return 0;
}
void* hkfhost_ctx() { return CudaDeviceDataMan::instance()->ctx(); }
void hkfhost_release() { MainController::instance()->release(); }

}  // extern "C"
