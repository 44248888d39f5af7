#include "DataSourceProducerSynthetic.h"
#include <cmath>
#include <cstdio>
#include <limits>

namespace {
const double kPlanes[3][2] = { { 1, 1.7 }, { 2, 1.6 }, { 0, 0.2 } };           // (axis, offset)
const double kSpheres[2][4] = { { 1.0, 1.1, 1.2, 0.35 }, { 1.45, 1.45, 0.9, 0.2 } };
const double kCentre[3] = { 1.0, 1.0, 1.2 };

void mul44(const double* a, const double* b, double* c)
{
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    for (int i = 0; i < 16; ++i) c[i] = r[i];
}
void identity(double* m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
// rigid inverse is exact enough here only for phase == 0 (identity); general 4x4 via Gauss-Jordan
void inverse44(const double* m, double* out)
{
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = m[i * 4 + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
        for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[p][j]);
        double d = a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] /= d;
        for (int r = 0; r < 4; ++r) if (r != c) { double f = a[r][c]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[i * 4 + j] = a[i][4 + j];
}
void motion(double a, double* m)
{
    const double th = 0.25 * std::sin(a);
    const double d[3] = { 0.10 * std::sin(a), 0.05 * std::sin(2 * a), 0.05 * (1.0 - std::cos(a)) };
    double ry[16], tc[16], tmc[16], td[16], t1[16], t2[16];
    identity(ry); identity(tc); identity(tmc); identity(td);
    const double c = std::cos(th), s = std::sin(th);
    ry[0] = c; ry[2] = s; ry[8] = -s; ry[10] = c;
    for (int i = 0; i < 3; ++i) { tc[i * 4 + 3] = kCentre[i]; tmc[i * 4 + 3] = -kCentre[i]; td[i * 4 + 3] = d[i]; }
    mul44(td, tc, t1); mul44(t1, ry, t2); mul44(t2, tmc, m);
}
}  // namespace

void DataSourceProducerSynthetic::trajectoryPose(unsigned i, unsigned n, double phase, double volume_size, double trunc_min, double out[16])
{
    const double kTwoPi = 2.0 * 3.14159265358979323846;
    double p0[16]; identity(p0);
    p0[3] = volume_size / 2.0; p0[7] = volume_size / 2.0; p0[11] = -trunc_min;
    double m0[16], base[16], m[16], t[16];
    motion(kTwoPi * 0 / n + phase, m0);
    inverse44(m0, base);                      // with a phase offset, re-anchor so that frame 0 == pose0
    motion(kTwoPi * i / n + phase, m);
    mul44(m, base, t); mul44(t, p0, out);
}

void DataSourceProducerSynthetic::renderDepthMm(const double pose[16], const CameraParams& cam, uint16_t* out)
{
    const double inf = std::numeric_limits<double>::infinity();
    const double o[3] = { pose[3], pose[7], pose[11] };
    for (unsigned v = 0; v < cam.rows; ++v)
        for (unsigned u = 0; u < cam.cols; ++u) {
            const double dc[3] = { ((double)u - (double)cam.cx) / (double)cam.fx, ((double)v - (double)cam.cy) / (double)cam.fy, 1.0 };
            double d[3];
            for (int i = 0; i < 3; ++i) d[i] = dc[0] * pose[i * 4 + 0] + dc[1] * pose[i * 4 + 1] + dc[2] * pose[i * 4 + 2];
            double best = inf;
            for (const auto& pl : kPlanes) {
                const int ax = (int)pl[0];
                const double s = (pl[1] - o[ax]) / d[ax];
                if (s > 1e-9 && std::isfinite(s) && s < best) best = s;
            }
            for (const auto& sp : kSpheres) {
                const double oc[3] = { o[0] - sp[0], o[1] - sp[1], o[2] - sp[2] };
                const double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                const double b = 2.0 * (d[0] * oc[0] + d[1] * oc[1] + d[2] * oc[2]);
                const double c = (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - sp[3] * sp[3];
                const double disc = b * b - 4 * a * c;
                if (disc < 0) continue;
                const double sq = std::sqrt(disc);
                const double s1 = (-b - sq) / (2 * a), s2 = (-b + sq) / (2 * a);
                const double s = s1 > 1e-9 ? s1 : s2;
                if (s > 1e-9 && s < best) best = s;
            }
            const double mm = best * 1000.0;
            out[(size_t)v * cam.cols + u] = (std::isfinite(mm) && mm < 65535.0 && mm > 0) ? (uint16_t)std::nearbyint(mm) : 0;
        }
}

void DataSourceProducerSynthetic::renderColor(const double pose[16], const CameraParams& cam, uint8_t* out)
{
    static const double kBgr[5][3] = { { 90, 180, 200 }, { 220, 160, 90 }, { 120, 220, 120 }, { 80, 80, 230 }, { 60, 240, 240 } };
    const double inf = std::numeric_limits<double>::infinity();
    const double o[3] = { pose[3], pose[7], pose[11] };
    for (unsigned v = 0; v < cam.rows; ++v)
        for (unsigned u = 0; u < cam.cols; ++u) {
            const double dc[3] = { ((double)u - (double)cam.cx) / (double)cam.fx, ((double)v - (double)cam.cy) / (double)cam.fy, 1.0 };
            double d[3];
            for (int i = 0; i < 3; ++i) d[i] = dc[0] * pose[i * 4 + 0] + dc[1] * pose[i * 4 + 1] + dc[2] * pose[i * 4 + 2];
            double best = inf;
            int surf = -1, k = 0;
            for (const auto& pl : kPlanes) {
                const int ax = (int)pl[0];
                const double s = (pl[1] - o[ax]) / d[ax];
                if (s > 1e-9 && std::isfinite(s) && s < best) { best = s; surf = k; }
                ++k;
            }
            for (const auto& sp : kSpheres) {
                const double oc[3] = { o[0] - sp[0], o[1] - sp[1], o[2] - sp[2] };
                const double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                const double b = 2.0 * (d[0] * oc[0] + d[1] * oc[1] + d[2] * oc[2]);
                const double c = (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - sp[3] * sp[3];
                const double disc = b * b - 4 * a * c;
                if (disc >= 0) {
                    const double sq = std::sqrt(disc);
                    const double s1 = (-b - sq) / (2 * a), s2 = (-b + sq) / (2 * a);
                    const double s = s1 > 1e-9 ? s1 : s2;
                    if (s > 1e-9 && s < best) { best = s; surf = k; }
                }
                ++k;
            }
            uint8_t* px = out + 3 * ((size_t)v * cam.cols + u);
            if (!std::isfinite(best)) { px[0] = px[1] = px[2] = 0; continue; }
            double p[3];
            long long cell = 0;
            for (int i = 0; i < 3; ++i) { p[i] = o[i] + d[i] * best; cell += (long long)std::floor(p[i] / 0.1 + 1e-6); }
            const double check = (double)(cell & 1);
            const double ramp = p[0] / 2.0 < 0.0 ? 0.0 : p[0] / 2.0 > 1.0 ? 1.0 : p[0] / 2.0;
            const double shade = (1.0 - 0.4 * check) * (0.75 + 0.25 * ramp);
            for (int i = 0; i < 3; ++i) {
                const double c = std::nearbyint(kBgr[surf][i] * shade);
                px[i] = (uint8_t)(c < 0 ? 0 : c > 255 ? 255 : c);
            }
        }
}

static void mat_to_quat(const double* p, double q[4])   // (qx,qy,qz,qw) of the rotation block of a row-major 4x4
{
    auto R = [&](int i, int j) { return p[i * 4 + j]; };
    const double t = R(0, 0) + R(1, 1) + R(2, 2);
    if (t > 0) {
        const double s = std::sqrt(t + 1.0) * 2;
        q[0] = (R(2, 1) - R(1, 2)) / s; q[1] = (R(0, 2) - R(2, 0)) / s; q[2] = (R(1, 0) - R(0, 1)) / s; q[3] = 0.25 * s;
        return;
    }
    int i = 0;
    if (R(1, 1) > R(i, i)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    const double s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0) * 2;
    q[i] = 0.25 * s; q[j] = (R(j, i) + R(i, j)) / s; q[k] = (R(k, i) + R(i, k)) / s; q[3] = (R(k, j) - R(j, k)) / s;
}

bool DataSourceProducerSynthetic::writeTumTrajectory(const std::string& path, unsigned n, double phase, double volume_size, double trunc_min, double fps)
{
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "# ground truth trajectory\n# file: synthetic\n# timestamp tx ty tz qx qy qz qw\n");
    for (unsigned i = 0; i < n; ++i) {
        double p[16], q[4];
        trajectoryPose(i, n, phase, volume_size, trunc_min, p);
        mat_to_quat(p, q);
        std::fprintf(f, "%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n", i / fps, p[3], p[7], p[11], q[0], q[1], q[2], q[3]);
    }
    std::fclose(f);
    return true;
}

bool DataSourceProducerSynthetic::initDataSource()
{
    const AppParams* P = AppParams::instance();
    _n = P->_synthetic_params.nFrames;
    _phase = P->_synthetic_params.phase;
    _buf.resize((size_t)P->_depth_camera_params.cols * P->_depth_camera_params.rows);
    if (P->_switch_params.useRGBData) _rgb.resize((size_t)P->_rgb_camera_params.cols * P->_rgb_camera_params.rows * 3);
    _next = 0;
    return _n > 0 && !_buf.empty();
}

bool DataSourceProducerSynthetic::readDataFromSource(DepthFrameData& depth_data, ColorFrameData& rgb_data)
{
    const AppParams* P = AppParams::instance();
    double pose[16];
    trajectoryPose(_next % _n, _n, _phase, P->_volume_params.fVolumeMeterSize, P->_depth_prepocess_params.fMinTrunc, pose);
    renderDepthMm(pose, P->_depth_camera_params, _buf.data());
    const double ts = _next / 30.0;
    depth_data.setFrameId(_next); rgb_data.setFrameId(_next);
    depth_data.setFrameData(_buf.data(), (int)P->_depth_camera_params.cols, (int)P->_depth_camera_params.rows, ts);
    if (P->_switch_params.useRGBData) {
        renderColor(pose, P->_rgb_camera_params, _rgb.data());
        rgb_data.setFrameData(_rgb.data(), (int)P->_rgb_camera_params.cols, (int)P->_rgb_camera_params.rows, ts);
    } else {
        rgb_data.setFrameData(nullptr, 0, 0, ts);
    }
    ++_next;
    return true;
}
