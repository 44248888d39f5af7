// PngReader — minimal PNG decoder for the TUM RGB-D dataset's depth images (16-bit greyscale, also 8-bit grey /
// RGB / RGBA), non-interlaced, zlib inflate + the five scanline filters.  Stands in for cv::imread
// (src/DataSourceProducerRGBDDataset.cpp:108), OpenCV's C++ API not being a dependency of this path.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct PngImage {
    unsigned width = 0, height = 0, channels = 0, bit_depth = 0;
    std::vector<uint16_t> samples;      // [height][width][channels], 8-bit samples widened
};

bool readPng(const std::string& path, PngImage& out, std::string* err = nullptr);
