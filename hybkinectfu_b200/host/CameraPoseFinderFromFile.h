// CameraPoseFinderFromFile — poses replayed from a TUM-format trajectory "timestamp tx ty tz qx qy qz qw"
// (src/CameraPoseFinderFromFile.{h,cpp}): three header lines skipped, nearest-timestamp match with a one-line rewind,
// pose = (pose0 * file0^-1) * file_i.  Pure host code; the quaternion -> matrix conversion is Eigen's, restated.
#pragma once
#include <fstream>
#include "CameraPoseFinder.h"

class CameraPoseFinderFromFile : public CameraPoseFinder {
public:
    ~CameraPoseFinderFromFile() override { if (_trajfile_src.is_open()) _trajfile_src.close(); }
    // exposed for tests: the record the reference would pick for `timestamp`
    bool parseFrameFromFile(double timestamp, float t[3], float q[4], double& res_timestamp);
protected:
    bool initPoseFinder() override;
    bool estimateCameraPose(const DepthFrameData& depth_frame, const ColorFrameData& color_frame) override;
private:
    std::ifstream _trajfile_src;
    Mat44 _refer_transform;
};
