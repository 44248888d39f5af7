#include "CameraPoseFinderICP.h"
#include "../csrc/hkf_solver.h"
#include "CudaDeviceDataMan.h"

bool CameraPoseFinderICP::poseIsDeviceResident() const { return !AppParams::instance()->_switch_params.stepwiseTracking; }

bool CameraPoseFinderICP::initPoseFinder()
{
    // iteration counts indexed by level, level 0 = finest (src/CameraPoseFinderICP.cpp:16-30)
    const unsigned levels = AppParams::instance()->_icp_params.nPyramidLevels;
    if (levels == 1) _iter_nums = { 3 };
    else if (levels == 2) _iter_nums = { 10, 5 };
    else if (levels == 3) _iter_nums = { 10, 5, 4 };
    else return false;
    // every intrinsic halves per level (:36-46)
    CameraParams cam = AppParams::instance()->_depth_camera_params;
    _camera_params_pyramid.clear();
    for (unsigned l = 0; l < levels; ++l) {
        _camera_params_pyramid.push_back(cam);
        cam.cols /= 2; cam.rows /= 2; cam.cx = cam.cx / 2; cam.cy = cam.cy / 2; cam.fx = cam.fx / 2; cam.fy = cam.fy / 2;
    }
    return true;
}

bool CameraPoseFinderICP::minimizePointToPlaneErrFunc(unsigned level, float six_dof[6], const Mat44& cur_transform, const Mat44& last_transform_inv)
{
    hkf_ctx* c = CudaDeviceDataMan::instance()->ctx();
    const IcpParams& ip = AppParams::instance()->_icp_params;
    const CameraParams& cam = _camera_params_pyramid[level];
    if (hkf_cal_point_to_plane_err_solver_params(c, level, cur_transform.abi(), last_transform_inv.abi(),
                                                 reinterpret_cast<const HkfCameraParams*>(&cam), ip.fDistThres, ip.fNormSinThres) != HKF_OK) return false;
    float sums[27];
    if (hkf_download(c, HKF_BUF_SOLVER_SUMS, 0, sums, sizeof(sums)) != (long)sizeof(sums)) return false;   // rigid_align_buf_reduced.clone(CPU)
    if (hkf_icp_singular(sums)) { _status = HKF_TRACK_SINGULAR; return false; }                              // det(ATA) < 1e-10
    float A[36], b[6];
    hkf_unpack27(sums, A, b);
    hkf_llt_solve6(A, b, six_dof);
    return true;
}

bool CameraPoseFinderICP::estimateCameraPose(const DepthFrameData& depth_frame, const ColorFrameData&)
{
    _status = HKF_TRACK_OK;
    if (depth_frame.frameId() == 0) return true;                  // nothing to align against yet (:52-55)
    hkf_ctx* c = CudaDeviceDataMan::instance()->ctx();
    const AppParams* P = AppParams::instance();
    if (!P->_switch_params.stepwiseTracking) {
        HkfFrameParams fp = P->frameParams();
        if (hkf_icp_track(c, &fp) != HKF_OK) { _status = HKF_TRACK_SINGULAR; return false; }
        return true;      // status and pose stay on the device; HybKinectfu reads them back once per frame
    }
    // ---- the reference's host loop
    hkf_downsample_new_vertices(c); hkf_downsample_new_normals(c);
    hkf_downsample_model_vertices(c); hkf_downsample_model_normals(c);
    Mat44 cur_transform = _pose;
    const Mat44 last_transform_inv = _pose.getInverse();
    for (int level = (int)_iter_nums.size() - 1; level >= 0; --level)
        for (int it = 0; it < _iter_nums[level]; ++it) {
            float x[6], T[16];
            if (!minimizePointToPlaneErrFunc((unsigned)level, x, cur_transform, last_transform_inv)) return false;
            if (!hkf_vector6_to_transform(x, P->_icp_params.fAngleShake, P->_icp_params.fDistShake, T)) { _status = HKF_TRACK_SHAKE; return false; }
            Mat44 inc; std::memcpy(inc.m, T, sizeof(T));
            cur_transform = inc * cur_transform;
        }
    _pose = cur_transform;
    return true;
}
