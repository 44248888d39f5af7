#include "CameraPoseFinderFromFile.h"
#include <sstream>
#include <string>
#include "../csrc/hkf_solver.h"

bool CameraPoseFinderFromFile::initPoseFinder()
{
    _trajfile_src.open(AppParams::instance()->_io_params.trajReadFilename.c_str());
    if (!_trajfile_src.is_open()) return false;
    std::string line;
    for (int i = 0; i < 3; ++i) std::getline(_trajfile_src, line);
    return true;
}

bool CameraPoseFinderFromFile::parseFrameFromFile(double target, float t[3], float q[4], double& res)
{
    double last_ts = 0;
    float last_t[3] = { 0, 0, 0 }, last_q[4] = { 0, 0, 0, 0 };
    std::string line;
    while (!_trajfile_src.eof()) {
        const std::streampos pos = _trajfile_src.tellg();
        std::getline(_trajfile_src, line);
        std::stringstream ss(line);
        ss >> res >> t[0] >> t[1] >> t[2] >> q[0] >> q[1] >> q[2] >> q[3];     // q = (x, y, z, w)
        if (res >= target) {
            if (res - target > target - last_ts) {                                 // the previous record is nearer
                for (int i = 0; i < 3; ++i) t[i] = last_t[i];
                for (int i = 0; i < 4; ++i) q[i] = last_q[i];
                res = last_ts;
                _trajfile_src.seekg(pos);                                          // re-read this line for the next frame
            }
            return true;
        }
        last_ts = res;
        for (int i = 0; i < 3; ++i) last_t[i] = t[i];
        for (int i = 0; i < 4; ++i) last_q[i] = q[i];
    }
    return false;
}

bool CameraPoseFinderFromFile::estimateCameraPose(const DepthFrameData& depth_frame, const ColorFrameData&)
{
    float t[3], q[4];
    double ts;
    _status = HKF_TRACK_OK;
    if (!parseFrameFromFile(depth_frame.timeStamp(), t, q, ts)) { _status = HKF_TRACK_NO_POSE; return false; }
    // Eigen: Quaternion(w,x,y,z) -> Matrix3f r (column-major storage); the reference then builds Mat33 from
    // r.transpose().data(), i.e. row-major r: the file rotation is used as is
    float r[9];
    hkf_quat_to_mat3(q[3], q[0], q[1], q[2], r);
    Mat44 file_transform = Mat44::getIdentity();
    file_transform.setRotation(r);
    file_transform.setTranslation(t[0], t[1], t[2]);
    if (depth_frame.frameId() == 0) {
        _refer_transform = _pose * file_transform.getInverse();
        return true;
    }
    _pose = _refer_transform * file_transform;
    return true;
}
