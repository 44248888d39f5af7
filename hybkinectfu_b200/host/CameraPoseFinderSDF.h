// CameraPoseFinderSDF — direct SDF tracking (src/CameraPoseFinderSDF.{h,cpp}): up to max_iter_nums Gauss-Newton steps
// of the full-resolution depth against the TSDF.  fused: hkf_sdf_track() (device loop, se(3) exponential map on the
// device); stepwise: hkf_cal_sdf_solver_params (= cudaCalSDFSolverParams) + host solve per iteration.
#pragma once
#include "CameraPoseFinder.h"

class CameraPoseFinderSDF : public CameraPoseFinder {
public:
    bool poseIsDeviceResident() const override;
protected:
    bool initPoseFinder() override { return true; }
    bool estimateCameraPose(const DepthFrameData& depth_frame, const ColorFrameData& color_frame) override;
};
