#include "DataSourceProducerRGBDDataset.h"
#include <sstream>
#include "PngReader.h"

bool DataSourceProducerRGBDDataset::initDataSource()
{
    _dptfile.open((_sourcefilename + "depth.txt").c_str());
    if (!_dptfile.is_open()) { _err = "cannot open " + _sourcefilename + "depth.txt"; return false; }
    std::string line;
    for (int i = 0; i < 3; ++i) std::getline(_dptfile, line);
    _next = 0;
    return true;
}

void DataSourceProducerRGBDDataset::toMillimetres(std::vector<uint16_t>& img, unsigned f)
{
    for (auto& v : img) v = (uint16_t)((v + f / 2) / f);          // cv::Mat /= f on CV_16U: round to nearest (no exact halves for odd f)
}

static inline int reflect101(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return i; }

void DataSourceProducerRGBDDataset::pyrDown(const std::vector<uint16_t>& src, unsigned w, unsigned h, std::vector<uint16_t>& dst)
{
    const unsigned dw = (w + 1) / 2, dh = (h + 1) / 2;
    dst.assign((size_t)dw * dh, 0);
    static const int k[5] = { 1, 4, 6, 4, 1 };
    std::vector<int> rowsum((size_t)dw * h);
    for (unsigned y = 0; y < h; ++y)                                // horizontal pass at the decimated columns
        for (unsigned x = 0; x < dw; ++x) {
            int s = 0;
            for (int i = -2; i <= 2; ++i) s += k[i + 2] * src[(size_t)y * w + reflect101(2 * (int)x + i, (int)w)];
            rowsum[(size_t)y * dw + x] = s;
        }
    for (unsigned y = 0; y < dh; ++y)
        for (unsigned x = 0; x < dw; ++x) {
            int s = 0;
            for (int i = -2; i <= 2; ++i) s += k[i + 2] * rowsum[(size_t)reflect101(2 * (int)y + i, (int)h) * dw + x];
            dst[(size_t)y * dw + x] = (uint16_t)((s + 128) >> 8);
        }
}

bool DataSourceProducerRGBDDataset::readDataFromSource(DepthFrameData& depth_data, ColorFrameData& rgb_data)
{
    if (_dptfile.eof()) return false;
    std::string line, rel;
    std::getline(_dptfile, line);
    std::stringstream ss(line);
    double ts = 0;
    if (!(ss >> ts >> rel)) return false;
    PngImage img;
    if (!readPng(_sourcefilename + rel, img, &_err)) return false;
    if (img.channels != 1) { _err = "depth image is not greyscale"; return false; }
    toMillimetres(img.samples, _depth_factor);
    const CameraParams& cam = AppParams::instance()->_depth_camera_params;
    depth_data.setFrameId(_next); rgb_data.setFrameId(_next);
    if (cam.cols != img.width || cam.rows != img.height) {
        std::vector<uint16_t> half;
        pyrDown(img.samples, img.width, img.height, half);
        const unsigned dw = (img.width + 1) / 2, dh = (img.height + 1) / 2;
        if (cam.cols != dw || cam.rows != dh) { _err = "depth image size does not match the configured depth camera (nor half of it)"; return false; }
        depth_data.setFrameData(half.data(), (int)dw, (int)dh, ts);
    } else {
        depth_data.setFrameData(img.samples.data(), (int)img.width, (int)img.height, ts);
    }
    rgb_data.setFrameData(nullptr, 0, 0, ts);
    ++_next;
    return true;
}
