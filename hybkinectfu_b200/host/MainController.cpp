#include "MainController.h"
#include <chrono>
#include <iostream>
#include "AppParamsProducer.h"
#include "CudaDeviceDataMan.h"
#include "DataSourceProducerRGBDDataset.h"
#include "DataSourceProducerSynthetic.h"

bool MainController::init(const std::string& config_filename, int device)
{
    release();
    if (!AppParamsProducer::readParams(config_filename)) { _err = "cannot read " + config_filename; return false; }
    if (!CudaDeviceDataMan::instance()->init(device)) { _err = "device init failed: " + CudaDeviceDataMan::instance()->lastError(); return false; }
    const SwitchParams& sw = AppParams::instance()->_switch_params;
    if (sw.useSyntheticData) _data_source = new DataSourceProducerSynthetic();
    else if (sw.useDatasetRGBD) _data_source = new DataSourceProducerRGBDDataset();
    else { _err = "the OpenNI2 live source needs hardware and a library that are not available: use the synthetic or the TUM dataset source"; return false; }
    if (!_data_source->init()) { _err = "data source init failed"; return false; }
    _kinect_fusioner = new HybKinectfu();
    if (!_kinect_fusioner->init()) { _err = "HybKinectfu::init failed"; return false; }
    _mesh_generator = new MeshGeneratorMarchingcube();
    return true;
}

int MainController::mainLoop(unsigned max_frames, bool save_mesh, bool verbose)
{
    if (!_kinect_fusioner) return -1;
    DepthFrameData depth;
    ColorFrameData rgb;
    unsigned n = 0;
    _lost = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (; n < max_frames; ++n) {
        if (!_data_source->readNewFrame(depth, rgb)) break;
        if (!_kinect_fusioner->processNewFrame(depth, rgb)) { _err = std::string("processNewFrame failed: ") + hkf_last_error(CudaDeviceDataMan::instance()->ctx()); return -1; }
        if (!_kinect_fusioner->lastTrackingSuccess()) ++_lost;
        if (verbose) std::cout << "frame " << n << " status " << _kinect_fusioner->lastTrackingStatus() << "\n" << _kinect_fusioner->cameraPose();
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << n << " frames in " << s << " s (" << (s > 0 ? n / s : 0.0) << " frames/s incl. synthetic rendering), lost " << _lost << std::endl;
    if (save_mesh) {
        _mesh_generator->generateMesh();
        const std::string& fn = AppParams::instance()->_io_params.meshFilename;
        if (!fn.empty() && !_mesh_generator->saveMesh(fn)) { _err = "saveMesh failed"; return -1; }
    }
    return (int)n;
}

void MainController::release()
{
    delete _mesh_generator; _mesh_generator = nullptr;
    delete _kinect_fusioner; _kinect_fusioner = nullptr;
    delete _data_source; _data_source = nullptr;
    CudaDeviceDataMan::instance()->release();
}
