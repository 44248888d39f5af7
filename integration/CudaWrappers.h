// CudaWrappers.h — drop-in replacement of the reference's src/cuda/CudaWrappers.h (the 13 live `extern "C" void cuda*()`
// wrappers, src/cuda/CudaWrappers.h:22-36): same names, same argument lists, forwarding to libhkf_b200.so through one
// process-wide context handle.  A maintainer of the reference copies THIS file over src/cuda/CudaWrappers.h, adds the
// three call-site changes listed in INTEGRATION.md §2 and links with -lhkf_b200.
//
// tests/test_abi.py::test_shim_compiles_against_the_reference_headers compiles it against the reference's own
// src/AppParams.h and src/cuda/Mat.h (where /root/reference is present) and checks the layout claims below.
#ifndef CUDAWRAPPERS_H_
#define CUDAWRAPPERS_H_
#include "hkf.h"
#include "../AppParams.h"
#include "Mat.h"

extern hkf_ctx* g_hkf;                                    // created once in MainController::init()

// the PODs cross the ABI by reinterpretation: the layouts are identical
static_assert(sizeof(Mat44) == sizeof(HkfMat44), "Mat44 is 16 row-major floats");
static_assert(sizeof(CameraParams) == sizeof(HkfCameraParams), "CameraParams: cols, rows, cx, cy, fx, fy");
static_assert(sizeof(IntegrateParams) == sizeof(HkfIntegrateParams), "IntegrateParams: fSdfTruncation, fMaxIntegrateDist");
static_assert(sizeof(RayCasterParams) == sizeof(HkfRayCasterParams), "RayCasterParams: fRayIncrement");

static inline const HkfMat44* hkf_M(const Mat44& m) { return reinterpret_cast<const HkfMat44*>(&m); }
static inline const HkfCameraParams* hkf_C(const CameraParams& c) { return reinterpret_cast<const HkfCameraParams*>(&c); }

inline void cudaTruncDepth(float mn, float mx)                      { hkf_trunc_depth(g_hkf, mn, mx); }
inline void cudaBiliearFilterDepth(float sp, float sd)              { hkf_bilateral_filter_depth(g_hkf, sp, sd); }
inline void cudaCalculateNewVertices(const CameraParams& c)         { hkf_calculate_new_vertices(g_hkf, hkf_C(c)); }
inline void cudaCalculateNewNormals()                               { hkf_calculate_new_normals(g_hkf); }
inline void cudaDownSampleNewVertices()                             { hkf_downsample_new_vertices(g_hkf); }
inline void cudaDownSampleNewNormals()                              { hkf_downsample_new_normals(g_hkf); }
inline void cudaDownSampleModelVertices()                           { hkf_downsample_model_vertices(g_hkf); }
inline void cudaDownSampleModelNormals()                            { hkf_downsample_model_normals(g_hkf); }
inline void cudaCalPointToPlaneErrSolverParams(unsigned l, const Mat44& cur, const Mat44& li, const CameraParams& c, float d, float s)
                                                                    { hkf_cal_point_to_plane_err_solver_params(g_hkf, l, hkf_M(cur), hkf_M(li), hkf_C(c), d, s); }
inline void cudaCalSDFSolverParams(const CameraParams& c, const Mat44& cur) { hkf_cal_sdf_solver_params(g_hkf, hkf_C(c), hkf_M(cur)); }
inline void cudaIntegrateVolume(bool col, bool ang, const Mat44& p, const IntegrateParams& ip, const CameraParams& dc, const CameraParams& rc)
    { hkf_integrate_volume(g_hkf, col, ang, hkf_M(p), reinterpret_cast<const HkfIntegrateParams*>(&ip), hkf_C(dc), hkf_C(rc)); }
inline void cudaRaycastingVolume(bool col, const Mat44& p, const RayCasterParams& rp, const CameraParams& c, float n, float f)
    { hkf_raycasting_volume(g_hkf, col, hkf_M(p), reinterpret_cast<const HkfRayCasterParams*>(&rp), hkf_C(c), n, f); }
inline void cudaMarchingcube(bool col, float thr)                   { hkf_marchingcube(g_hkf, col, thr); }
#endif
