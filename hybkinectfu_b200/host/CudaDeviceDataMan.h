// CudaDeviceDataMan.h — the process-wide owner of all device buffers (src/cuda/CudaDeviceDataMan.h:14-77).  In the
// reference this singleton IS the hidden half of the wrapper interface; here it is a thin holder of the hkf_ctx
// handle of libhkf_b200.so, so the retained host classes keep their "reach the singleton" shape while the library
// itself stays re-entrant (several contexts per process for replicas / z-slab ranks).
#pragma once
#include <string>
#include "AppParams.h"

class CudaDeviceDataMan {
public:
    static CudaDeviceDataMan* instance() { static CudaDeviceDataMan v; return &v; }
    // allocates every buffer from AppParams (reference: init() reads the same singleton, :24-51)
    bool init(int device = 0)
    {
        if (_ctx) return false;
        HkfInitParams ip = AppParams::instance()->initParams();
        int rc = hkf_create(&ip, device, &_ctx);
        if (rc != HKF_OK) { _err = hkf_last_error(nullptr); _ctx = nullptr; return false; }
        return true;
    }
    void release() { if (_ctx) hkf_destroy(_ctx); _ctx = nullptr; }
    hkf_ctx* ctx() const { return _ctx; }
    const std::string& lastError() const { return _err; }
    ~CudaDeviceDataMan() { release(); }
private:
    CudaDeviceDataMan() {}
    hkf_ctx* _ctx = nullptr;
    std::string _err;
};
