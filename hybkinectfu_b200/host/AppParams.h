// AppParams.h — the parameter singleton the reference's host classes read (src/AppParams.h:12-173): same struct
// and field names so that code written against the reference compiles against this header, with defaults taken
// from the reference's root config.ini where it has one.  The PODs that cross the C ABI (CameraParams,
// IntegrateParams, RayCasterParams) are layout-identical to HkfCameraParams etc. in include/hkf.h.
#pragma once
#include <string>
#include "../../include/hkf.h"

struct SwitchParams {
    bool recordRGBD = false, recordTrajectory = false, useRGBData = false, colorAngleWeight = false;
    bool useDatasetRGBD = false, useTrajFromFile = false, useSdfTracker = false;
    bool useSyntheticData = true;     // new: DataSourceProducerSynthetic (the reference's OpenNI2 source needs hardware)
    bool stepwiseTracking = false;    // new: run the pose finders iteration by iteration on the host (reference control flow)
};
struct CameraParams {
    unsigned cols = 640, rows = 480;
    float cx = 319.5f, cy = 239.5f, fx = 525.0f, fy = 525.0f;
    CameraParams operator*(int rate) const
    {
        CameraParams r;
        r.cols = cols * rate; r.rows = rows * rate;
        r.cx = cx * rate; r.cy = cy * rate; r.fx = fx * rate; r.fy = fy * rate;
        return r;
    }
};
struct DepthPrepocessParams { float fMaxTrunc = 4.0f, fMinTrunc = 0.3f, fSigmaDepth = 0.02f, fSigmaPixel = 2.0f; };
struct IcpParams { unsigned nPyramidLevels = 3; float fNormSinThres = 0.2f, fDistThres = 0.1f, fDistShake = 0.4f, fAngleShake = 0.4f; };
struct SDFTrackerParams { unsigned maxIterNums = 6; float fDistShake = 0.4f, fAngleShake = 0.4f; };
struct tsdfVolumeParams { unsigned nResolution = 256; float fVolumeMeterSize = 2.0f, fWeightMax = 128.0f; };
struct IntegrateParams { float fSdfTruncation = 0.04f, fMaxIntegrateDist = 2.0f; };
struct RayCasterParams { float fRayIncrement = 0.8f * 0.04f; };
struct MarchingcubeParams { unsigned uMaxTriangles = 1500000; };
struct IOParams { std::string meshFilename, rgbdReadFilename, rgbdWriteFilename, trajReadFilename, trajWriteFilename; };
// new: synthetic sequence and z-slab sharding
struct SyntheticParams { unsigned nFrames = 500; double phase = 0.0; double noiseSigmaMm = 0.0; };

static_assert(sizeof(CameraParams) == sizeof(HkfCameraParams), "CameraParams crosses the C ABI");
static_assert(sizeof(IntegrateParams) == sizeof(HkfIntegrateParams), "IntegrateParams crosses the C ABI");
static_assert(sizeof(RayCasterParams) == sizeof(HkfRayCasterParams), "RayCasterParams crosses the C ABI");

class AppParams {
public:
    static AppParams* instance() { static AppParams v; return &v; }
    void print() const;
    // the per-frame PODs the fused C-ABI paths take
    HkfFrameParams frameParams() const;
    HkfInitParams initParams() const;
    SwitchParams _switch_params;
    CameraParams _rgb_camera_params, _depth_camera_params;
    DepthPrepocessParams _depth_prepocess_params;
    IcpParams _icp_params;
    SDFTrackerParams _sdf_tracker_params;
    tsdfVolumeParams _volume_params;
    IntegrateParams _integrate_params;
    RayCasterParams _raycast_params;
    MarchingcubeParams _marchingcube_params;
    IOParams _io_params;
    SyntheticParams _synthetic_params;
protected:
    AppParams() {}
};
