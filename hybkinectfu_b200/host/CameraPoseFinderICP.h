// CameraPoseFinderICP — projective point-to-plane ICP (src/CameraPoseFinderICP.{h,cpp}).
// Two execution modes behind the same interface:
//   fused (default)  hkf_icp_track(): model pyramids, the coarse-to-fine schedule {3} / {10,5} / {10,5,4}, the 27
//                    sums, the 6x6 solve, the determinant and shake gates and the pose update all run on the device;
//                    one status + pose read-back per frame instead of 19 blocking 27-float copies;
//   stepwise         the reference's control flow verbatim: per iteration hkf_cal_point_to_plane_err_solver_params
//                    (= cudaCalPointToPlaneErrSolverParams), download the 27 sums, solve on the host
//                    (csrc/hkf_solver.h: Eigen's LLT / determinant / AngleAxis restated), left-multiply.
#pragma once
#include <vector>
#include "CameraPoseFinder.h"

class CameraPoseFinderICP : public CameraPoseFinder {
public:
    bool poseIsDeviceResident() const override;
protected:
    bool initPoseFinder() override;
    bool estimateCameraPose(const DepthFrameData& depth_frame, const ColorFrameData& color_frame) override;
private:
    bool minimizePointToPlaneErrFunc(unsigned level, float six_dof[6], const Mat44& cur_transform, const Mat44& last_transform_inv);
    std::vector<int> _iter_nums;
    std::vector<CameraParams> _camera_params_pyramid;
};
