#include "CameraPoseFinderSDF.h"
#include "../csrc/hkf_solver.h"
#include "CudaDeviceDataMan.h"

bool CameraPoseFinderSDF::poseIsDeviceResident() const { return !AppParams::instance()->_switch_params.stepwiseTracking; }

bool CameraPoseFinderSDF::estimateCameraPose(const DepthFrameData& depth_frame, const ColorFrameData&)
{
    _status = HKF_TRACK_OK;
    if (depth_frame.frameId() == 0) return true;                  // src/CameraPoseFinderSDF.cpp:46-49
    hkf_ctx* c = CudaDeviceDataMan::instance()->ctx();
    const AppParams* P = AppParams::instance();
    if (!P->_switch_params.stepwiseTracking) {
        HkfFrameParams fp = P->frameParams();
        if (hkf_sdf_track(c, &fp) != HKF_OK) { _status = HKF_TRACK_SHAKE; return false; }
        return true;
    }
    Mat44 cur = _pose;
    const CameraParams& cam = P->_depth_camera_params;
    for (unsigned it = 0; it < P->_sdf_tracker_params.maxIterNums; ++it) {
        if (hkf_cal_sdf_solver_params(c, reinterpret_cast<const HkfCameraParams*>(&cam), cur.abi()) != HKF_OK) return false;
        float sums[27];
        if (hkf_download(c, HKF_BUF_SOLVER_SUMS, 0, sums, sizeof(sums)) != (long)sizeof(sums)) return false;
        int converged = 0;
        // llt solve without a determinant guard, shake gate, |x| < 1e-3 stop BEFORE composing, se(3) exp, R <- R_d^T R (:79-100)
        if (hkf_sdf_update(sums, P->_sdf_tracker_params.fAngleShake, P->_sdf_tracker_params.fDistShake, cur.m, &converged) != HKF_SOLVE_OK) {
            _status = HKF_TRACK_SHAKE;
            return false;
        }
        if (converged) break;
    }
    _pose = cur;
    return true;
}
