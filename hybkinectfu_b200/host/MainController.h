// MainController — process-level driver (src/MainController.{h,cpp}) without the GUI: init(config) reads the
// parameters, allocates the device state, builds the data source, the fusion pipeline and the mesh generator;
// mainLoop(max_frames) pulls frames until the source ends, and generates + saves the mesh at the end (the reference
// does so on key 's' or at frame 1990).
#pragma once
#include <string>
#include "DataSourceProducer.h"
#include "HybKinectfu.h"
#include "MeshGeneratorMarchingcube.h"

class MainController {
public:
    static MainController* instance() { static MainController v; return &v; }
    bool init(const std::string& config_filename, int device = 0);
    // returns the number of frames processed; prints per-frame tracking state when verbose
    int mainLoop(unsigned max_frames, bool save_mesh, bool verbose);
    void release();
    HybKinectfu* fusioner() { return _kinect_fusioner; }
    DataSourceProducer* dataSource() { return _data_source; }
    MeshGeneratorMarchingcube* meshGenerator() { return _mesh_generator; }
    const std::string& lastError() const { return _err; }
    unsigned framesLost() const { return _lost; }
    ~MainController() { release(); }
private:
    MainController() {}
    DataSourceProducer* _data_source = nullptr;
    HybKinectfu* _kinect_fusioner = nullptr;
    MeshGeneratorMarchingcube* _mesh_generator = nullptr;
    std::string _err;
    unsigned _lost = 0;
};
