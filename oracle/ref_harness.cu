// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// ref_harness.cu: a thin C-ABI driver around the UNMODIFIED reference device code
// (/root/reference/src/cuda/*.cu), compiled from where it lies by oracle/Makefile
// into oracle/_ref/libhkf_ref.so.  Nothing from the reference is copied here; the
// reference headers are #included from $(REF)/src/cuda at build time.
//
// It fills the reference's AppParams singleton directly (its fields are public,
// src/AppParams.h:156-166), calls CudaDeviceDataMan::instance()->init()
// (src/cuda/CudaDeviceDataMan.h:24-51), and then calls the reference's own
// extern "C" wrappers (src/cuda/CudaWrappers.h:22-36) exactly as
// HybKinectfu::processNewFrame (src/HybKinectfu.cpp:98-160) and the pose finders do.
//
// Fences around reference UB:
//  (i)  pyramids: the four cudaDownSample* wrappers launch <<<block,grid>>> swapped
//       (src/cuda/sample.cu:72,84,96,109).  We include sample.cu in this TU and launch the
//       unmodified pryDown*Kernel ourselves with a correct config into padded buffers.
//  (ii) colour: off unless ref_set_color() switches it on.  The reference's SOURCE reads the colour pixel before its bounds
//       test (src/cuda/integrateVolume.cu:58-62); the code nvcc 12.9 generates for sm_100 sinks the three byte loads below
//       every test (checked in the SASS of _ref/integrateVolume.o), so the compiled oracle never reads out of bounds.
//  (iii) the MC triangle counter is cleared before each cudaMarchingcube call.
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <vector>

#include "CudaWrappers.h"
#include "CudaDeviceDataMan.h"
#include "DepthCamera.h"
// the unmodified pyramid kernels (and their broken wrappers, which we never call)
#include "sample.cu"

#include "host_math.h"

struct RefParams {
    uint32_t cols, rows;
    float cx, cy, fx, fy;
    float trunc_min, trunc_max, sigma_pixel, sigma_depth;
    uint32_t resolution;
    float volume_size, weight_max;
    float sdf_trunc, max_integrate_dist;
    float ray_increment;
    uint32_t pyramid_levels;
    float icp_dist_thres, icp_norm_sin_thres, icp_dist_shake, icp_angle_shake;
    uint32_t sdf_max_iter;
    float sdf_dist_shake, sdf_angle_shake;
    uint32_t max_triangles;
};

enum RefBuffer {
    REF_RAW_DEPTH = 0, REF_TRUNC_DEPTH = 1, REF_FILTERED_DEPTH = 2,
    REF_NEW_VERTICES = 3, REF_NEW_NORMALS = 4, REF_MODEL_VERTICES = 5, REF_MODEL_NORMALS = 6,
    REF_VOLUME = 7 /* AoS Voxel 12 B */, REF_ICP_REDUCED = 8, REF_TRIANGLES = 9,
    REF_VOLUME_COLOR = 10 /* uchar4 per voxel (.w = 0) */, REF_RAYCAST_RGB = 11, REF_RAW_RGB = 12 /* uchar3 maps */
};

static bool g_inited = false;
static Mat44 g_pose;
static std::vector<int> g_iter_nums;
static std::vector<CameraParams> g_cam_pyr;

static CudaDeviceDataMan* DM() { return CudaDeviceDataMan::instance(); }
static AppParams* AP() { return AppParams::instance(); }

extern "C" int ref_init(const RefParams* p)
{
    AppParams* a = AP();
    a->_switch_params.recordRGBD = false;
    a->_switch_params.recordTrajectory = false;
    a->_switch_params.useRGBData = false;
    a->_switch_params.colorAngleWeight = false;
    a->_switch_params.useDatasetRGBD = false;
    a->_switch_params.useTrajFromFile = false;
    a->_switch_params.useSdfTracker = false;
    CameraParams c; c.cols = p->cols; c.rows = p->rows; c.cx = p->cx; c.cy = p->cy; c.fx = p->fx; c.fy = p->fy;
    a->_depth_camera_params = c;
    a->_rgb_camera_params = c;
    a->_depth_prepocess_params.fMaxTrunc = p->trunc_max;
    a->_depth_prepocess_params.fMinTrunc = p->trunc_min;
    a->_depth_prepocess_params.fSigmaDepth = p->sigma_depth;
    a->_depth_prepocess_params.fSigmaPixel = p->sigma_pixel;
    a->_icp_params.nPyramidLevels = p->pyramid_levels;
    a->_icp_params.fNormSinThres = p->icp_norm_sin_thres;
    a->_icp_params.fDistThres = p->icp_dist_thres;
    a->_icp_params.fDistShake = p->icp_dist_shake;
    a->_icp_params.fAngleShake = p->icp_angle_shake;
    a->_sdf_tracker_params.maxIterNums = p->sdf_max_iter;
    a->_sdf_tracker_params.fDistShake = p->sdf_dist_shake;
    a->_sdf_tracker_params.fAngleShake = p->sdf_angle_shake;
    a->_volume_params.nResolution = p->resolution;
    a->_volume_params.fVolumeMeterSize = p->volume_size;
    a->_volume_params.fWeightMax = p->weight_max;
    a->_integrate_params.fSdfTruncation = p->sdf_trunc;
    a->_integrate_params.fMaxIntegrateDist = p->max_integrate_dist;
    a->_raycast_params.fRayIncrement = p->ray_increment;
    a->_marchingcube_params.uMaxTriangles = p->max_triangles;

    CudaDeviceDataMan* d = DM();
    // init() push_back()s pyramid levels, so clear them to allow re-init within one process
    d->corrs_vertices_pyramid.clear();
    d->new_vertices_pyramid.clear();
    d->new_normals_pyramid.clear();
    d->model_vertices_pyramid.clear();
    d->model_normals_pyramid.clear();
    d->rigid_align_buf_temp_pyramid.clear();
    d->init();
    // model maps start undefined in the reference (cudaMalloc); zero them so downloads are deterministic
    for (size_t l = 0; l < d->model_vertices_pyramid.size(); ++l) {
        d->model_vertices_pyramid[l].clearData();
        d->model_normals_pyramid[l].clearData();
        d->new_vertices_pyramid[l].clearData();
        d->new_normals_pyramid[l].clearData();
    }
    d->raw_depth.clearData(); d->trunced_depth.clearData(); d->filtered_depth.clearData();

    // CameraPoseFinderICP::initPoseFinder (src/CameraPoseFinderICP.cpp:12-48)
    unsigned levels = p->pyramid_levels;
    g_iter_nums.assign(levels, 0);
    if (levels == 1) { g_iter_nums[0] = 3; }
    else if (levels == 2) { g_iter_nums[0] = 10; g_iter_nums[1] = 5; }
    else if (levels == 3) { g_iter_nums[0] = 10; g_iter_nums[1] = 5; g_iter_nums[2] = 4; }
    else return -1;
    g_cam_pyr.clear();
    g_cam_pyr.push_back(c);
    for (unsigned l = 1; l < levels; ++l) {
        CameraParams q;
        q.cols = g_cam_pyr[l - 1].cols / 2; q.rows = g_cam_pyr[l - 1].rows / 2;
        q.cx = g_cam_pyr[l - 1].cx / 2; q.cy = g_cam_pyr[l - 1].cy / 2;
        q.fx = g_cam_pyr[l - 1].fx / 2; q.fy = g_cam_pyr[l - 1].fy / 2;
        g_cam_pyr.push_back(q);
    }
    // HybKinectfu::init pose0 (src/HybKinectfu.cpp:51-54)
    g_pose = Mat44::getIdentity();
    g_pose.setTranslation(make_float3(p->volume_size / 2.0, p->volume_size / 2.0, -p->trunc_min));
    g_inited = true;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "ref_init: %s\n", cudaGetErrorString(e)); return -2; }
    return 0;
}

// colour mode (src/HybKinectfu.cpp:134-135,149; src/MeshGeneratorMarchingcube.cpp:26)
extern "C" void ref_set_color(int use_color, int angle_weight)
{
    AP()->_switch_params.useRGBData = use_color != 0;
    AP()->_switch_params.colorAngleWeight = angle_weight != 0;
}
// the colour half of HybKinectfu::copyFrameToGPU (src/HybKinectfu.cpp:80-95): 3 bytes per pixel, rgb camera size
extern "C" void ref_upload_rgb(const uint8_t* rgb3)
{
    unsigned cols = AP()->_rgb_camera_params.cols, rows = AP()->_rgb_camera_params.rows;
    Color3uMap2D cpu(CPU, cols, rows);
    for (unsigned r = 0; r < rows; ++r)
        for (unsigned c = 0; c < cols; ++c) {
            const uint8_t* q = rgb3 + 3 * ((size_t)r * cols + c);
            cpu.at(c, r) = make_uchar3(q[0], q[1], q[2]);
        }
    DM()->raw_rgb.copyDataFrom(cpu);
}

extern "C" void ref_reset_volume()
{
    DM()->volume.init(AP()->_volume_params);
    g_pose = Mat44::getIdentity();
    g_pose.setTranslation(make_float3(AP()->_volume_params.fVolumeMeterSize / 2.0,
                                      AP()->_volume_params.fVolumeMeterSize / 2.0,
                                      -AP()->_depth_prepocess_params.fMinTrunc));
}

extern "C" void ref_get_pose(float out[16]) { memcpy(out, g_pose.entries, 64); }
extern "C" void ref_set_pose(const float in[16]) { g_pose = Mat44(in); }

// HybKinectfu::copyFrameToGPU (src/HybKinectfu.cpp:63-78): double multiply, then narrow
extern "C" void ref_upload_depth_mm(const uint16_t* mm)
{
    unsigned cols = AP()->_depth_camera_params.cols, rows = AP()->_depth_camera_params.rows;
    DepthfMap2D cpu(CPU, cols, rows);
    for (unsigned r = 0; r < rows; ++r)
        for (unsigned c = 0; c < cols; ++c) {
            float v = mm[r * cols + c] * 0.001;
            cpu.at(c, r) = v;
        }
    DM()->raw_depth.copyDataFrom(cpu);
}

extern "C" void ref_trunc() { cudaTruncDepth(AP()->_depth_prepocess_params.fMinTrunc, AP()->_depth_prepocess_params.fMaxTrunc); }
extern "C" void ref_bilateral() { cudaBiliearFilterDepth(AP()->_depth_prepocess_params.fSigmaPixel, AP()->_depth_prepocess_params.fSigmaDepth); }
extern "C" void ref_vertices() { cudaCalculateNewVertices(AP()->_depth_camera_params); }
extern "C" void ref_normals() { cudaCalculateNewNormals(); }

// fenced pyramid: unmodified kernels, corrected launch, padded buffers (see header comment).  The padded buffers are
// allocated once per image size (the stock path has no allocation in a frame either); the padding rows are zeroed at
// allocation and never written (only cols x rows texels are copied in), everything runs in order on the default stream.
struct PadBufs { std::vector<Point4fMap2D> in, out; unsigned c0 = 0, r0 = 0; };
static PadBufs g_pad[2];
template <bool NORMALS>
static void fenced_down(std::vector<CudaMap2D<float4>>& pyr)
{
    const int levels = (int)pyr.size();
    if (levels < 2) return;
    PadBufs& P = g_pad[NORMALS ? 1 : 0];
    if ((int)P.in.size() != levels - 1 || P.c0 != pyr[0].cols() || P.r0 != pyr[0].rows()) {
        P.in.clear(); P.out.clear();
        for (int i = 0; i < levels - 1; ++i) {
            P.in.push_back(Point4fMap2D(GPU, pyr[i].cols(), pyr[i].rows() + 64));
            P.out.push_back(Point4fMap2D(GPU, pyr[i + 1].cols(), pyr[i + 1].rows() + 32));
            P.in.back().clearData(); P.out.back().clearData();
        }
        P.c0 = pyr[0].cols(); P.r0 = pyr[0].rows();
    }
    for (int i = 0; i < levels - 1; ++i) {
        unsigned ci = pyr[i].cols(), ri = pyr[i].rows();
        unsigned co = pyr[i + 1].cols(), ro = pyr[i + 1].rows();
        cudaMemcpyAsync(P.in[i].ptr(), pyr[i].ptr(), sizeof(float4) * (size_t)ci * ri, cudaMemcpyDeviceToDevice, 0);
        const dim3 blockSize(BLOCK_SIZE_2D_X, BLOCK_SIZE_2D_Y);
        const dim3 gridSize(divUp(co, BLOCK_SIZE_2D_X), divUp(ro, BLOCK_SIZE_2D_Y));
        if (NORMALS) pryDownNormalsKernel<<<gridSize, blockSize>>>(P.in[i], P.out[i]);
        else pryDownVerticesKernel<<<gridSize, blockSize>>>(P.in[i], P.out[i]);
        cudaMemcpyAsync(pyr[i + 1].ptr(), P.out[i].ptr(), sizeof(float4) * (size_t)co * ro, cudaMemcpyDeviceToDevice, 0);
    }
}
extern "C" void ref_downsample_new()
{
    fenced_down<false>(DM()->new_vertices_pyramid);
    fenced_down<true>(DM()->new_normals_pyramid);
}
extern "C" void ref_downsample_model()
{
    fenced_down<false>(DM()->model_vertices_pyramid);
    fenced_down<true>(DM()->model_normals_pyramid);
}

extern "C" void ref_icp_params(unsigned level, const float cur[16], const float last_inv[16], float out27[27])
{
    Mat44 c(cur), li(last_inv);
    cudaCalPointToPlaneErrSolverParams(level, c, li, g_cam_pyr[level], AP()->_icp_params.fDistThres, AP()->_icp_params.fNormSinThres);
    CudaMap1D<float> buf = DM()->rigid_align_buf_reduced.clone(CPU);
    for (int i = 0; i < 27; ++i) out27[i] = buf.at(i);
}
extern "C" void ref_sdf_params(const float cur[16], float out27[27])
{
    Mat44 c(cur);
    cudaCalSDFSolverParams(AP()->_depth_camera_params, c);
    CudaMap1D<float> buf = DM()->rigid_align_buf_reduced.clone(CPU);
    for (int i = 0; i < 27; ++i) out27[i] = buf.at(i);
}
extern "C" void ref_integrate(const float pose[16])
{
    Mat44 p(pose);
    // silence the wrapper's per-call cout
    std::streambuf* old = std::cout.rdbuf(nullptr);
    cudaIntegrateVolume(AP()->_switch_params.useRGBData, AP()->_switch_params.colorAngleWeight, p, AP()->_integrate_params,
                        AP()->_depth_camera_params, AP()->_rgb_camera_params);
    std::cout.rdbuf(old);
}
extern "C" void ref_raycast(const float pose[16])
{
    Mat44 p(pose);
    cudaRaycastingVolume(AP()->_switch_params.useRGBData, p, AP()->_raycast_params, AP()->_depth_camera_params,
                         AP()->_depth_prepocess_params.fMinTrunc, AP()->_depth_prepocess_params.fMaxTrunc);
}
extern "C" unsigned ref_marchingcube()
{
    DM()->marchingcube_data.clearData();
    // MeshGeneratorMarchingcube::generateMesh threshold (src/MeshGeneratorMarchingcube.cpp:26-27)
    cudaMarchingcube(AP()->_switch_params.useRGBData, 300 * AP()->_volume_params.fVolumeMeterSize / AP()->_volume_params.nResolution);
    return DM()->marchingcube_data.triangleNums();
}

static size_t map_bytes(int buffer, unsigned level, void** dptr)
{
    CudaDeviceDataMan* d = DM();
    switch (buffer) {
    case REF_RAW_DEPTH: *dptr = d->raw_depth.ptr(); return sizeof(float) * d->raw_depth.cols() * d->raw_depth.rows();
    case REF_TRUNC_DEPTH: *dptr = d->trunced_depth.ptr(); return sizeof(float) * d->trunced_depth.cols() * d->trunced_depth.rows();
    case REF_FILTERED_DEPTH: *dptr = d->filtered_depth.ptr(); return sizeof(float) * d->filtered_depth.cols() * d->filtered_depth.rows();
    case REF_NEW_VERTICES: *dptr = d->new_vertices_pyramid[level].ptr(); return sizeof(float4) * d->new_vertices_pyramid[level].cols() * d->new_vertices_pyramid[level].rows();
    case REF_NEW_NORMALS: *dptr = d->new_normals_pyramid[level].ptr(); return sizeof(float4) * d->new_normals_pyramid[level].cols() * d->new_normals_pyramid[level].rows();
    case REF_MODEL_VERTICES: *dptr = d->model_vertices_pyramid[level].ptr(); return sizeof(float4) * d->model_vertices_pyramid[level].cols() * d->model_vertices_pyramid[level].rows();
    case REF_MODEL_NORMALS: *dptr = d->model_normals_pyramid[level].ptr(); return sizeof(float4) * d->model_normals_pyramid[level].cols() * d->model_normals_pyramid[level].rows();
    case REF_ICP_REDUCED: *dptr = d->rigid_align_buf_reduced.ptr(); return sizeof(float) * 27;
    case REF_RAYCAST_RGB: *dptr = d->raycast_rgb.ptr(); return sizeof(uchar3) * d->raycast_rgb.cols() * d->raycast_rgb.rows();
    case REF_RAW_RGB: *dptr = d->raw_rgb.ptr(); return sizeof(uchar3) * d->raw_rgb.cols() * d->raw_rgb.rows();
    default: *dptr = nullptr; return 0;
    }
}

// Volume view: the reference stores AoS Voxel{tsdf,weight,uchar3} (12 B) in a protected member.
// A derived type reaches it without touching the reference source.
struct VolAccess : public tsdfvolume {
    VolAccess(const tsdfvolume& v) : tsdfvolume(v) {}
    Voxel* raw() { return _data.ptr(); }
};
__global__ void ref_volume_to_float2(const Voxel* vox, float2* out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_float2(vox[i].tsdf, vox[i].weight);
}
__global__ void ref_volume_to_color(const Voxel* vox, uchar4* out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_uchar4(vox[i].color.x, vox[i].color.y, vox[i].color.z, 0);
}
__global__ void ref_volume_from_float2(Voxel* vox, const float2* in, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        vox[i].tsdf = in[i].x; vox[i].weight = in[i].y;
    }
}

extern "C" long ref_download(int buffer, unsigned level, void* dst, size_t cap)
{
    if (buffer == REF_VOLUME) {
        int R = AP()->_volume_params.nResolution;
        size_t n = (size_t)R * R * R;
        if (cap < n * sizeof(float2)) return -1;
        float2* tmp; cudaMalloc(&tmp, n * sizeof(float2));
        VolAccess va(DM()->volume);
        ref_volume_to_float2<<<1184, 256>>>(va.raw(), tmp, n);
        cudaMemcpy(dst, tmp, n * sizeof(float2), cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        return (long)(n * sizeof(float2));
    }
    if (buffer == REF_VOLUME_COLOR) {
        int R = AP()->_volume_params.nResolution;
        size_t n = (size_t)R * R * R;
        if (cap < n * sizeof(uchar4)) return -1;
        uchar4* tmp; cudaMalloc(&tmp, n * sizeof(uchar4));
        VolAccess va(DM()->volume);
        ref_volume_to_color<<<1184, 256>>>(va.raw(), tmp, n);
        cudaMemcpy(dst, tmp, n * sizeof(uchar4), cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        return (long)(n * sizeof(uchar4));
    }
    if (buffer == REF_TRIANGLES) {
        MarchingcubeData cpu = DM()->marchingcube_data.clone(CPU);
        unsigned n = cpu.triangleNums();
        if (cap < (size_t)n * sizeof(Triangle)) return -1;
        for (unsigned i = 0; i < n; ++i) memcpy((char*)dst + (size_t)i * sizeof(Triangle), &cpu.at(i), sizeof(Triangle));
        return (long)((size_t)n * sizeof(Triangle));
    }
    void* dptr; size_t nb = map_bytes(buffer, level, &dptr);
    if (!dptr || cap < nb) return -1;
    cudaMemcpy(dst, dptr, nb, cudaMemcpyDeviceToHost);
    return (long)nb;
}

// planes [z0, z1) of the volume as packed {tsdf, weight} (a 1024^3 volume does not fit a host array comfortably)
extern "C" long ref_download_volume_planes(unsigned z0, unsigned z1, void* dst, size_t cap)
{
    int R = AP()->_volume_params.nResolution;
    if (z0 >= z1 || z1 > (unsigned)R) return -1;
    size_t n = (size_t)(z1 - z0) * R * R;
    if (cap < n * sizeof(float2)) return -1;
    float2* tmp; if (cudaMalloc(&tmp, n * sizeof(float2)) != cudaSuccess) return -2;
    VolAccess va(DM()->volume);
    ref_volume_to_float2<<<1184, 256>>>(va.raw() + (size_t)z0 * R * R, tmp, n);
    cudaMemcpy(dst, tmp, n * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    return (long)(n * sizeof(float2));
}

extern "C" long ref_upload(int buffer, unsigned level, const void* src, size_t nbytes)
{
    if (buffer == REF_VOLUME) {
        int R = AP()->_volume_params.nResolution;
        size_t n = (size_t)R * R * R;
        if (nbytes != n * sizeof(float2)) return -1;
        float2* tmp; cudaMalloc(&tmp, n * sizeof(float2));
        cudaMemcpy(tmp, src, n * sizeof(float2), cudaMemcpyHostToDevice);
        VolAccess va(DM()->volume);
        ref_volume_from_float2<<<1184, 256>>>(va.raw(), tmp, n);
        cudaDeviceSynchronize();
        cudaFree(tmp);
        return (long)nbytes;
    }
    void* dptr; size_t nb = map_bytes(buffer, level, &dptr);
    if (!dptr || nbytes != nb) return -1;
    cudaMemcpy(dptr, src, nb, cudaMemcpyHostToDevice);
    return (long)nb;
}

// ---- full per-frame pipeline, as HybKinectfu::processNewFrame + the pose finders drive it ----
// status: 0 ok, 1 singular (det<1e-10), 2 shake, 3 not inited
static int ref_track_icp()
{
    // CameraPoseFinderICP::estimateCameraPose (src/CameraPoseFinderICP.cpp:50-86)
    ref_downsample_new();
    ref_downsample_model();
    Mat44 cur = g_pose;
    Mat44 last_inv = g_pose.getInverse();
    for (int l = (int)g_iter_nums.size() - 1; l >= 0; --l) {
        int it = g_iter_nums[l];
        while (it--) {
            float s27[27];
            ref_icp_params(l, cur.entries, last_inv.entries, s27);
            float x[6];
            int st = hm_solve_icp(s27, x);     // det guard + LLT  (src/CameraPoseFinderICP.cpp:118-144)
            if (st != 0) return 1;
            float T[16];
            if (!hm_vector6_to_transform(x, AP()->_icp_params.fAngleShake, AP()->_icp_params.fDistShake, T)) return 2;
            Mat44 inc(T);
            cur = inc * cur;
        }
    }
    g_pose = cur;
    return 0;
}
static int ref_track_sdf()
{
    // CameraPoseFinderSDF::estimateCameraPose (src/CameraPoseFinderSDF.cpp:44-106)
    Mat44 cur = g_pose;
    unsigned iter = 0;
    while (iter < AP()->_sdf_tracker_params.maxIterNums) {
        float s27[27];
        ref_sdf_params(cur.entries, s27);
        float x[6];
        hm_solve_llt(s27, x);                  // no det guard (src/CameraPoseFinderSDF.cpp:79)
        float T[16];
        if (!hm_vector6_to_transform(x, AP()->_sdf_tracker_params.fAngleShake, AP()->_sdf_tracker_params.fDistShake, T)) return 2;
        if (hm_norm6(x) < 0.001f) break;
        float next[16];
        hm_sdf_compose(cur.entries, x, next);
        cur = Mat44(next);
        iter++;
    }
    g_pose = cur;
    return 0;
}

// mode: 0 ICP, 1 SDF, 2 pose given (FromFile)
extern "C" int ref_process_frame(const uint16_t* depth_mm, unsigned frame_id, int mode, const float* given_pose, float pose_out[16])
{
    if (!g_inited) return 3;
    ref_upload_depth_mm(depth_mm);
    ref_trunc(); ref_bilateral(); ref_vertices(); ref_normals();
    int st = 0;
    if (mode == 2) { g_pose = Mat44(given_pose); }
    else if (frame_id != 0) { st = (mode == 1) ? ref_track_sdf() : ref_track_icp(); }
    if (st == 0) ref_integrate(g_pose.entries);
    ref_raycast(g_pose.entries);
    memcpy(pose_out, g_pose.entries, 64);
    return st;
}

extern "C" int ref_last_cuda_error()
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) fprintf(stderr, "ref: CUDA error %s\n", cudaGetErrorString(e));
    return (int)e;
}
