#include "TrajectoryRecorder.h"
#include <cmath>
#include <iomanip>

TrajectoryRecorder::TrajectoryRecorder(const std::string& record_filename)
{
    _record_file.open(record_filename.c_str());
    if (!_record_file.is_open()) return;
    _record_file << "# trajectory\n# file: " << record_filename << "\n# timestamp tx ty tz qx qy qz qw" << std::endl;
}

TrajectoryRecorder::~TrajectoryRecorder() { if (_record_file.is_open()) _record_file.close(); }

// Eigen::Quaternionf(Matrix3f): the branch on the trace, then on the largest diagonal entry
void TrajectoryRecorder::rotationToQuaternion(const float* m, float q[4])
{
    auto R = [&](int i, int j) { return m[i * 3 + j]; };
    float t = R(0, 0) + R(1, 1) + R(2, 2);
    if (t > 0.0f) {
        t = std::sqrt(t + 1.0f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (R(2, 1) - R(1, 2)) * t; q[1] = (R(0, 2) - R(2, 0)) * t; q[2] = (R(1, 0) - R(0, 1)) * t;
        return;
    }
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0f);
    q[i] = 0.5f * t;
    t = 0.5f / t;
    q[3] = (R(k, j) - R(j, k)) * t;
    q[j] = (R(j, i) + R(i, j)) * t;
    q[k] = (R(k, i) + R(i, k)) * t;
}

bool TrajectoryRecorder::recordCameraPose(const Mat44& mat, double timestamp)
{
    if (!_record_file.is_open()) return false;
    float r[9], q[4];
    mat.getRotation(r);
    rotationToQuaternion(r, q);
    const float3_ t = mat.getTranslation();
    _record_file << std::setprecision(14) << timestamp << " ";
    _record_file << std::setprecision(6) << t.x << " " << t.y << " " << t.z << " ";
    _record_file << std::setprecision(6) << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
    return true;
}
