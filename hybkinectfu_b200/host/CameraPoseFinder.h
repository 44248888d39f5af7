// CameraPoseFinder.h — pose-finder interface (src/CameraPoseFinder.h:15-43), unchanged: init(reference pose),
// findCameraPose(depth, colour) -> bool, getCameraPose()/setCameraPose(); subclasses implement initPoseFinder() and
// estimateCameraPose().  _pose is only updated on success.
#pragma once
#include "AppParams.h"
#include "FrameData.h"
#include "Mat.h"

class CameraPoseFinder {
public:
    CameraPoseFinder() {}
    virtual ~CameraPoseFinder() {}
    virtual bool init(const Mat44& reference_transform)
    {
        if (_inited) return false;
        _pose = reference_transform;
        if (!initPoseFinder()) return false;
        _inited = true;
        return true;
    }
    virtual bool findCameraPose(const DepthFrameData& depth_frame, const ColorFrameData& color_frame)
    {
        if (!_inited) return false;
        return estimateCameraPose(depth_frame, color_frame);
    }
    Mat44 getCameraPose() const { return _pose; }
    void setCameraPose(const Mat44& transform) { _pose = transform; }
    // HKF_TRACK_* of the last findCameraPose (the reference only has the bool)
    int lastStatus() const { return _status; }
    // true when the finder leaves the pose on the device and the caller should use the *_tracked C-ABI paths
    virtual bool poseIsDeviceResident() const { return false; }
protected:
    Mat44 _pose;
    int _status = HKF_TRACK_OK;
    virtual bool initPoseFinder() = 0;
    virtual bool estimateCameraPose(const DepthFrameData& depth_frame, const ColorFrameData& color_frame) = 0;
private:
    bool _inited = false;
};
