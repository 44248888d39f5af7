#include "HybKinectfu.h"
#include <iostream>
#include "CameraPoseFinderFromFile.h"
#include "CameraPoseFinderICP.h"
#include "CameraPoseFinderSDF.h"
#include "CudaDeviceDataMan.h"

HybKinectfu::HybKinectfu() : _camera_pose_finder(nullptr), _camera_pose_recorder(nullptr), _inited(false), _last_success(false), _last_status(HKF_TRACK_OK) {}

HybKinectfu::~HybKinectfu()
{
    delete _camera_pose_recorder;
    delete _camera_pose_finder;
}

bool HybKinectfu::init()
{
    if (_inited) return false;
    const AppParams* P = AppParams::instance();
    if (!CudaDeviceDataMan::instance()->ctx()) return false;
    if (P->_switch_params.useTrajFromFile) _camera_pose_finder = new CameraPoseFinderFromFile();
    else if (P->_switch_params.useSdfTracker) _camera_pose_finder = new CameraPoseFinderSDF();
    else _camera_pose_finder = new CameraPoseFinderICP();
    if (P->_switch_params.recordTrajectory) _camera_pose_recorder = new TrajectoryRecorder(P->_io_params.trajWriteFilename);
    // identity rotation, t = (S/2, S/2, -depth_trunc_min); S/2.0 is a double division narrowed to float (:51-54)
    Mat44 pose0 = Mat44::getIdentity();
    pose0.setTranslation((float)(P->_volume_params.fVolumeMeterSize / 2.0), (float)(P->_volume_params.fVolumeMeterSize / 2.0), -P->_depth_prepocess_params.fMinTrunc);
    if (!_camera_pose_finder->init(pose0)) return false;
    if (hkf_set_pose(CudaDeviceDataMan::instance()->ctx(), pose0.abi()) != HKF_OK) return false;
    _inited = true;
    return true;
}

bool HybKinectfu::processNewFrame(const DepthFrameData& depth_frame, const ColorFrameData& rgb_frame)
{
    if (!_inited) return false;
    hkf_ctx* c = CudaDeviceDataMan::instance()->ctx();
    const AppParams* P = AppParams::instance();
    const HkfFrameParams fp = P->frameParams();
    if ((unsigned)depth_frame.cols() != P->_depth_camera_params.cols || (unsigned)depth_frame.rows() != P->_depth_camera_params.rows) return false;
    // copyFrameToGPU + cudaTruncDepth + cudaBiliearFilterDepth + cudaCalculateNewVertices/Normals (+ new-frame pyramids)
    if (hkf_preprocess(c, depth_frame.mm(), 0, &fp) != HKF_OK) return false;
    const bool color = P->_switch_params.useRGBData;
    if (color) {
        // the colour half of copyFrameToGPU (src/HybKinectfu.cpp:80-95): the frame's 3-byte pixels as they are
        if (rgb_frame.empty() || (unsigned)rgb_frame.cols() != P->_rgb_camera_params.cols || (unsigned)rgb_frame.rows() != P->_rgb_camera_params.rows) return false;
        if (hkf_upload_color(c, reinterpret_cast<const uint8_t*>(rgb_frame.data()), 0) != HKF_OK) return false;
    }

    bool ok = _camera_pose_finder->findCameraPose(depth_frame, rgb_frame);
    const bool resident = _camera_pose_finder->poseIsDeviceResident();
    if (resident) {
        // integration is skipped on the device when tracking failed; raycast uses the last good pose
        if (hkf_integrate_tracked(c, &fp) != HKF_OK || hkf_raycast_tracked(c, &fp) != HKF_OK) return false;
        Mat44 pose; int st = HKF_TRACK_OK;
        if (hkf_frame_result(c, pose.abi(), &st) != HKF_OK) return false;       // the frame's only synchronisation
        _last_status = st;
        ok = ok && st == HKF_TRACK_OK;
        if (ok) _camera_pose_finder->setCameraPose(pose);
    } else {
        _last_status = ok ? HKF_TRACK_OK : _camera_pose_finder->lastStatus();
        const Mat44 pose = _camera_pose_finder->getCameraPose();
        const CameraParams& dc = P->_depth_camera_params;
        const CameraParams& rc = P->_rgb_camera_params;
        if (ok && hkf_integrate_volume(c, color ? 1 : 0, (color && P->_switch_params.colorAngleWeight) ? 1 : 0, pose.abi(), reinterpret_cast<const HkfIntegrateParams*>(&P->_integrate_params),
                                       reinterpret_cast<const HkfCameraParams*>(&dc), reinterpret_cast<const HkfCameraParams*>(&rc)) != HKF_OK) return false;
        if (hkf_raycasting_volume(c, color ? 1 : 0, pose.abi(), reinterpret_cast<const HkfRayCasterParams*>(&P->_raycast_params),
                                  reinterpret_cast<const HkfCameraParams*>(&dc), P->_depth_prepocess_params.fMinTrunc, P->_depth_prepocess_params.fMaxTrunc) != HKF_OK) return false;
        if (ok && hkf_set_pose(c, pose.abi()) != HKF_OK) return false;          // keep the device copy in step
    }
    _last_success = ok;
    if (ok && _camera_pose_recorder) _camera_pose_recorder->recordCameraPose(_camera_pose_finder->getCameraPose(), depth_frame.timeStamp());
    if (!ok) std::cout << "camera lost" << std::endl;
    return true;
}
