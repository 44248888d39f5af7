// DataSourceProducerSynthetic — analytic scene rendered along a known trajectory: floor y = 1.7,
// back wall z = 1.6, side wall x = 0.2, spheres c = (1.0,1.1,1.2) r = 0.35 and c = (1.45,1.45,0.9) r = 0.2, inside the
// 2 m volume cube; 16-bit millimetre depth; trajectory starting exactly at HybKinectfu's initial pose.
// The C++ twin of hybkinectfu_b200/synth.py (same double-precision arithmetic, same frames).  It takes the place of
// the reference's OpenNI2 / TUM-dataset producers, whose hardware, libraries and data are not available here.
#pragma once
#include <vector>
#include "DataSourceProducer.h"

class DataSourceProducerSynthetic : public DataSourceProducer {
public:
    // camera -> world pose of frame i of n (row-major 4x4, double)
    static void trajectoryPose(unsigned i, unsigned n, double phase, double volume_size, double trunc_min, double out[16]);
    static void renderDepthMm(const double pose[16], const CameraParams& cam, uint16_t* out);
    // 3 bytes per pixel (B, G, R like the reference's OpenCV frames) of the same scene: synth.render_color
    static void renderColor(const double pose[16], const CameraParams& cam, uint8_t* out);
    // "timestamp tx ty tz qx qy qz qw" with the three header lines CameraPoseFinderFromFile skips
    static bool writeTumTrajectory(const std::string& path, unsigned n, double phase, double volume_size, double trunc_min, double fps = 30.0);
    unsigned framesRead() const { return _next; }
protected:
    bool initDataSource() override;
    bool readDataFromSource(DepthFrameData& depth_data, ColorFrameData& rgb_data) override;
    bool recordFrameToFile() override { return true; }
    bool initRecorder(std::string) override { return true; }
private:
    unsigned _next = 0, _n = 0;
    double _phase = 0.0;
    std::vector<uint16_t> _buf;
    std::vector<uint8_t> _rgb;
};
