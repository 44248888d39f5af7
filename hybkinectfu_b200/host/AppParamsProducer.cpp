#include "AppParamsProducer.h"
#include <iostream>
#include "AppParams.h"
#include "INIParser.h"

static std::string g_missing;
std::string AppParamsProducer::lastMissingKeys() { return g_missing; }

bool AppParamsProducer::readParams(const std::string& filename)
{
    INIParser ini;
    if (!ini.readINI(filename)) return false;
    AppParams* P = AppParams::instance();
    auto num = [&](const char* s, const char* k, double d) { return ini.getNumber(s, k, d); };
    auto flag = [&](const char* k, bool d) { return ini.has("switch", k) ? ini.getNumber("switch", k, d) != 0 : d; };
    SwitchParams& sw = P->_switch_params;
    sw.useRGBData = flag("use_color", sw.useRGBData);
    sw.recordRGBD = flag("record_rgbd", sw.recordRGBD);
    sw.useDatasetRGBD = flag("read_rgbd_from_rgbddataset", sw.useDatasetRGBD);
    sw.recordTrajectory = flag("record_trajectory", sw.recordTrajectory);
    sw.useTrajFromFile = flag("read_trajactory_from_file", sw.useTrajFromFile);   // the reference's spelling
    sw.colorAngleWeight = flag("color_angle_weight", sw.colorAngleWeight);
    sw.useSdfTracker = flag("use_sdf_tracker", sw.useSdfTracker);
    sw.useSyntheticData = flag("use_synthetic_data", sw.useSyntheticData);
    sw.stepwiseTracking = flag("stepwise_tracking", sw.stepwiseTracking);

    CameraParams& rc = P->_rgb_camera_params;
    rc.cols = (unsigned)num("camera", "rgb_cols", rc.cols); rc.rows = (unsigned)num("camera", "rgb_rows", rc.rows);
    rc.cx = (float)num("camera", "rgb_cx", rc.cx); rc.cy = (float)num("camera", "rgb_cy", rc.cy);
    rc.fx = (float)num("camera", "rgb_fx", rc.fx); rc.fy = (float)num("camera", "rgb_fy", rc.fy);
    CameraParams& dc = P->_depth_camera_params;
    dc.cols = (unsigned)num("camera", "depth_cols", dc.cols); dc.rows = (unsigned)num("camera", "depth_rows", dc.rows);
    dc.cx = (float)num("camera", "depth_cx", dc.cx); dc.cy = (float)num("camera", "depth_cy", dc.cy);
    dc.fx = (float)num("camera", "depth_fx", dc.fx); dc.fy = (float)num("camera", "depth_fy", dc.fy);

    DepthPrepocessParams& dp = P->_depth_prepocess_params;
    dp.fMaxTrunc = (float)num("kinectfusion", "depth_trunc_max", dp.fMaxTrunc);
    dp.fMinTrunc = (float)num("kinectfusion", "depth_trunc_min", dp.fMinTrunc);
    dp.fSigmaPixel = (float)num("kinectfusion", "filter_sigma_pixel", dp.fSigmaPixel);
    dp.fSigmaDepth = (float)num("kinectfusion", "filter_sigma_depth", dp.fSigmaDepth);
    tsdfVolumeParams& vp = P->_volume_params;
    vp.nResolution = (unsigned)num("kinectfusion", "volume_resolution", vp.nResolution);
    vp.fVolumeMeterSize = (float)num("kinectfusion", "volume_size_meter", vp.fVolumeMeterSize);
    vp.fWeightMax = (float)num("kinectfusion", "volume_max_weight", vp.fWeightMax);
    IntegrateParams& ip = P->_integrate_params;
    ip.fSdfTruncation = (float)num("kinectfusion", "integrate_sdf_trunc", ip.fSdfTruncation);
    ip.fMaxIntegrateDist = (float)num("kinectfusion", "integrate_depth_trunc", ip.fMaxIntegrateDist);
    // fRayIncrement = raycast_increment_factor * integrate_sdf_trunc, a float product (src/AppParamsProducer.cpp:113-117)
    float factor = (float)num("kinectfusion", "raycast_increment_factor", 0.8);
    P->_raycast_params.fRayIncrement = factor * ip.fSdfTruncation;

    IcpParams& icp = P->_icp_params;
    icp.nPyramidLevels = (unsigned)num("icp", "icp_pyramid_level", icp.nPyramidLevels);
    icp.fDistThres = (float)num("icp", "icp_thre_dist", icp.fDistThres);
    icp.fNormSinThres = (float)num("icp", "icp_thre_sin_angle", icp.fNormSinThres);
    icp.fDistShake = (float)num("icp", "camera_shake_dist", icp.fDistShake);
    icp.fAngleShake = (float)num("icp", "camera_shake_angle", icp.fAngleShake);
    SDFTrackerParams& sp = P->_sdf_tracker_params;
    sp.maxIterNums = (unsigned)num("sdfTracker", "max_iter_nums", sp.maxIterNums);
    sp.fDistShake = (float)num("sdfTracker", "camera_shake_dist", sp.fDistShake);
    sp.fAngleShake = (float)num("sdfTracker", "camera_shake_angle", sp.fAngleShake);
    P->_marchingcube_params.uMaxTriangles = (unsigned)num("mesh", "max_triangle_num", P->_marchingcube_params.uMaxTriangles);

    IOParams& io = P->_io_params;
    io.rgbdWriteFilename = ini.getValue("IO", "rgbd_record_filename");
    io.rgbdReadFilename = ini.getValue("IO", "rgbd_source_filename");
    io.trajWriteFilename = ini.getValue("IO", "trajectory_record_filename");
    io.trajReadFilename = ini.getValue("IO", "trajectory_source_filename");
    io.meshFilename = ini.getValue("IO", "mesh_filename");

    SyntheticParams& sy = P->_synthetic_params;
    if (ini.has("synthetic", "frames")) sy.nFrames = (unsigned)ini.getNumber("synthetic", "frames", sy.nFrames);
    if (ini.has("synthetic", "phase")) sy.phase = ini.getNumber("synthetic", "phase", sy.phase);
    if (ini.has("synthetic", "noise_sigma_mm")) sy.noiseSigmaMm = ini.getNumber("synthetic", "noise_sigma_mm", sy.noiseSigmaMm);
    g_missing = ini.missing();
    return true;
}

void AppParams::print() const
{
    std::cout << "depth camera " << _depth_camera_params.cols << "x" << _depth_camera_params.rows << " c=(" << _depth_camera_params.cx << ","
              << _depth_camera_params.cy << ") f=(" << _depth_camera_params.fx << "," << _depth_camera_params.fy << ")\n"
              << "depth truncation (" << _depth_prepocess_params.fMinTrunc << "," << _depth_prepocess_params.fMaxTrunc << ") bilateral sigma px "
              << _depth_prepocess_params.fSigmaPixel << " depth " << _depth_prepocess_params.fSigmaDepth << "\n"
              << "volume " << _volume_params.nResolution << "^3 over " << _volume_params.fVolumeMeterSize << " m, max weight " << _volume_params.fWeightMax
              << ", sdf truncation " << _integrate_params.fSdfTruncation << ", integrate up to " << _integrate_params.fMaxIntegrateDist << " m, ray step "
              << _raycast_params.fRayIncrement << "\n"
              << "tracker " << (_switch_params.useTrajFromFile ? "poses from file" : _switch_params.useSdfTracker ? "SDF" : "ICP") << ", icp levels "
              << _icp_params.nPyramidLevels << ", sdf max iterations " << _sdf_tracker_params.maxIterNums << "\n"
              << "max triangles " << _marchingcube_params.uMaxTriangles << ", mesh file " << _io_params.meshFilename << std::endl;
}

HkfFrameParams AppParams::frameParams() const
{
    HkfFrameParams f{};
    f.trunc_min = _depth_prepocess_params.fMinTrunc; f.trunc_max = _depth_prepocess_params.fMaxTrunc;
    f.sigma_pixel = _depth_prepocess_params.fSigmaPixel; f.sigma_depth = _depth_prepocess_params.fSigmaDepth;
    f.integrate.fSdfTruncation = _integrate_params.fSdfTruncation; f.integrate.fMaxIntegrateDist = _integrate_params.fMaxIntegrateDist;
    f.raycast.fRayIncrement = _raycast_params.fRayIncrement;
    f.icp_dist_thres = _icp_params.fDistThres; f.icp_norm_sin_thres = _icp_params.fNormSinThres;
    f.icp_dist_shake = _icp_params.fDistShake; f.icp_angle_shake = _icp_params.fAngleShake;
    f.sdf_max_iter = _sdf_tracker_params.maxIterNums; f.sdf_dist_shake = _sdf_tracker_params.fDistShake; f.sdf_angle_shake = _sdf_tracker_params.fAngleShake;
    f.use_color = _switch_params.useRGBData ? 1u : 0u; f.color_angle_weight = _switch_params.colorAngleWeight ? 1u : 0u;
    return f;
}

HkfInitParams AppParams::initParams() const
{
    HkfInitParams p{};
    p.depth_camera.cols = _depth_camera_params.cols; p.depth_camera.rows = _depth_camera_params.rows;
    p.depth_camera.cx = _depth_camera_params.cx; p.depth_camera.cy = _depth_camera_params.cy;
    p.depth_camera.fx = _depth_camera_params.fx; p.depth_camera.fy = _depth_camera_params.fy;
    p.pyramid_levels = _icp_params.nPyramidLevels;
    p.volume_resolution = _volume_params.nResolution; p.volume_size = _volume_params.fVolumeMeterSize; p.weight_max = _volume_params.fWeightMax;
    p.max_triangles = _marchingcube_params.uMaxTriangles;
    p.use_color = _switch_params.useRGBData ? 1u : 0u;
    p.rgb_camera.cols = _rgb_camera_params.cols; p.rgb_camera.rows = _rgb_camera_params.rows;
    p.rgb_camera.cx = _rgb_camera_params.cx; p.rgb_camera.cy = _rgb_camera_params.cy;
    p.rgb_camera.fx = _rgb_camera_params.fx; p.rgb_camera.fy = _rgb_camera_params.fy;
    return p;
}
