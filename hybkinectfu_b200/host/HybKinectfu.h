// HybKinectfu — the per-frame orchestrator (src/HybKinectfu.{h,cpp}): init() picks the pose finder from the
// switches and sets the initial pose; processNewFrame() = upload + preprocess, track, integrate on success, raycast.
// Same public interface; behind it the frame travels as the raw 16-bit buffer (the mm -> m double multiply of
// copyFrameToGPU, :73, is done on the device), the preprocessing stages are one fused launch, and with a
// device-resident tracker the host synchronises once per frame instead of ~45 times.
// On tracking failure the reference blocks in cv::waitKey (:142-147); here processNewFrame returns true with
// lastTrackingSuccess() == false, integration skipped and the raycast taken from the previous pose - the same
// state transition, headless.
#pragma once
#include "CameraPoseFinder.h"
#include "FrameData.h"
#include "TrajectoryRecorder.h"

class HybKinectfu {
public:
    HybKinectfu();
    virtual ~HybKinectfu();
    bool init();
    bool processNewFrame(const DepthFrameData& depth_frame, const ColorFrameData& rgb_frame);
    bool lastTrackingSuccess() const { return _last_success; }
    int lastTrackingStatus() const { return _last_status; }
    Mat44 cameraPose() const { return _camera_pose_finder ? _camera_pose_finder->getCameraPose() : Mat44::getIdentity(); }
private:
    CameraPoseFinder* _camera_pose_finder;
    TrajectoryRecorder* _camera_pose_recorder;
    bool _inited, _last_success;
    int _last_status;
};
