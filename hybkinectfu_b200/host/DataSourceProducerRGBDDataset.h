// DataSourceProducerRGBDDataset — TUM RGB-D dataset reader (src/DataSourceProducerRGBDDataset.{h,cpp}): depth.txt
// lists "timestamp path" after three header lines; each depth image is a 16-bit PNG in 1/5000 m, divided by
// depth_factor = 5 to millimetres (cv::Mat /= 5 rounds to nearest); when the configured depth camera is half the
// image size the image goes through cv::pyrDown (5x5 Gaussian, BORDER_REFLECT_101, (sum + 128) >> 8), restated here.
// Colour frames are not read (use_color = 0 on this path).
#pragma once
#include <fstream>
#include <vector>
#include "DataSourceProducer.h"

class DataSourceProducerRGBDDataset : public DataSourceProducer {
public:
    ~DataSourceProducerRGBDDataset() override { if (_dptfile.is_open()) _dptfile.close(); }
    // exposed for tests
    static void toMillimetres(std::vector<uint16_t>& img, unsigned depth_factor);
    static void pyrDown(const std::vector<uint16_t>& src, unsigned w, unsigned h, std::vector<uint16_t>& dst);
    const std::string& lastError() const { return _err; }
protected:
    bool initDataSource() override;
    bool readDataFromSource(DepthFrameData& depth_data, ColorFrameData& rgb_data) override;
    bool recordFrameToFile() override { return false; }      // not implemented in the reference either
    bool initRecorder(std::string) override { return false; }
private:
    std::ifstream _dptfile;
    unsigned _depth_factor = 5, _next = 0;
    std::string _err;
};
