// FrameData.h — frame containers handed to HybKinectfu::processNewFrame (src/FrameData.h:16-119), re-backed by a
// plain buffer because OpenCV's C++ API is not a dependency of this path.  Same accessors: frameId(), timeStamp(),
// setFrameData(data, cols, rows, time_stamp); mat() is replaced by data()/cols()/rows().
// DepthFrameData is 16-bit unsigned MILLIMETRES, 0 = invalid (src/FrameData.h:57,70).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

class FrameData {
public:
    FrameData() {}
    explicit FrameData(unsigned frameId) : _frame_id(frameId) {}
    virtual ~FrameData() {}
    virtual void setFrameData(const void* data, int cols, int rows, double time_stamp) = 0;
    double timeStamp() const { return _time_stamp; }
    unsigned frameId() const { return _frame_id; }
    void setFrameId(unsigned id) { _frame_id = id; }
    int cols() const { return _cols; }
    int rows() const { return _rows; }
    bool empty() const { return _bytes.empty(); }
    const void* data() const { return _bytes.data(); }
    void copyTo(FrameData& o) const { o._bytes = _bytes; o._cols = _cols; o._rows = _rows; o._time_stamp = _time_stamp; o._frame_id = _frame_id; o._is_setted = _is_setted; }
protected:
    void assign(const void* data, size_t nbytes, int cols, int rows, double ts)
    {
        _bytes.resize(nbytes);
        if (nbytes) std::memcpy(_bytes.data(), data, nbytes);
        _cols = cols; _rows = rows; _time_stamp = ts; _is_setted = true;
    }
    std::vector<uint8_t> _bytes;
    int _cols = 0, _rows = 0;
    double _time_stamp = 0.0;
    bool _is_setted = false;
    unsigned _frame_id = 0;
};

class DepthFrameData : public FrameData {
public:
    DepthFrameData() {}
    explicit DepthFrameData(unsigned frameId) : FrameData(frameId) {}
    void setFrameData(const void* data, int cols, int rows, double time_stamp) override { assign(data, (size_t)cols * rows * 2, cols, rows, time_stamp); }
    const uint16_t* mm() const { return reinterpret_cast<const uint16_t*>(_bytes.data()); }
};

class ColorFrameData : public FrameData {
public:
    ColorFrameData() {}
    explicit ColorFrameData(unsigned frameId) : FrameData(frameId) {}
    void setFrameData(const void* data, int cols, int rows, double time_stamp) override { assign(data, data ? (size_t)cols * rows * 3 : 0, cols, rows, time_stamp); }
};
