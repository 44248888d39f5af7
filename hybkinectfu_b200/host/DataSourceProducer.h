// DataSourceProducer.h — frame-source interface (src/DataSourceProducer.h:13-50), unchanged in shape: init(),
// readNewFrame(depth, rgb) over the protected virtuals.  readNewFrame returns true when not recording (the
// reference falls off the end of the function there: undefined behaviour).
#pragma once
#include <string>
#include "AppParams.h"
#include "FrameData.h"

class DataSourceProducer {
public:
    DataSourceProducer() {}
    virtual ~DataSourceProducer() {}
    bool init()
    {
        if (_inited) return false;
        _sourcefilename = AppParams::instance()->_io_params.rgbdReadFilename;
        _capture_color = AppParams::instance()->_switch_params.useRGBData;
        _recording = AppParams::instance()->_switch_params.recordRGBD;
        if (!initDataSource()) return false;
        if (_recording && !initRecorder(AppParams::instance()->_io_params.rgbdWriteFilename)) return false;
        _inited = true;
        return true;
    }
    bool readNewFrame(DepthFrameData& depth_data, ColorFrameData& rgb_data)
    {
        if (!_inited) return false;
        if (!readDataFromSource(depth_data, rgb_data)) return false;
        return _recording ? recordFrameToFile() : true;
    }
protected:
    virtual bool initDataSource() = 0;
    virtual bool readDataFromSource(DepthFrameData& depth_data, ColorFrameData& rgb_data) = 0;
    virtual bool recordFrameToFile() = 0;
    virtual bool initRecorder(std::string record_filename) = 0;
    std::string _sourcefilename;
    bool _capture_color = false;
private:
    bool _inited = false, _recording = false;
};
