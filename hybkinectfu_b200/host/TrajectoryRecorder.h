// TrajectoryRecorder — TUM-format trajectory writer (src/TrajectoryRecorder.{h,cpp}): three header lines, then
// "timestamp tx ty tz qx qy qz qw" with the stream precisions of the reference (14 significant digits for the
// timestamp, 6 for the rest).  Eigen's Matrix3f -> Quaternionf conversion restated in float.
#pragma once
#include <fstream>
#include <string>
#include "Mat.h"

class TrajectoryRecorder {
public:
    explicit TrajectoryRecorder(const std::string& record_filename);
    ~TrajectoryRecorder();
    bool recordCameraPose(const Mat44& mat, double timestamp);
    static void rotationToQuaternion(const float* r9, float q[4]);   // -> (x, y, z, w)
private:
    std::ofstream _record_file;
};
