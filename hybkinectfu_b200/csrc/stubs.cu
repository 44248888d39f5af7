// temporary: entry points not implemented yet return HKF_ERR_STATE (never a silent fallback)
#include "hkf_internal.cuh"
#define NOT_YET(c) do { if (c) (c)->err = std::string(__func__) + ": not implemented yet"; return HKF_ERR_STATE; } while (0)
extern "C" int hkf_cal_point_to_plane_err_solver_params(hkf_ctx* c, unsigned, const HkfMat44*, const HkfMat44*, const HkfCameraParams*, float, float) { NOT_YET(c); }
extern "C" int hkf_cal_sdf_solver_params(hkf_ctx* c, const HkfCameraParams*, const HkfMat44*) { NOT_YET(c); }
extern "C" int hkf_raycasting_volume(hkf_ctx* c, int, const HkfMat44*, const HkfRayCasterParams*, const HkfCameraParams*, float, float) { NOT_YET(c); }
extern "C" int hkf_marchingcube(hkf_ctx* c, int, float) { NOT_YET(c); }
extern "C" int hkf_triangle_count(hkf_ctx* c, unsigned* n) { if (n) *n = 0; NOT_YET(c); }
extern "C" int hkf_icp_track(hkf_ctx* c, const HkfFrameParams*) { NOT_YET(c); }
extern "C" int hkf_sdf_track(hkf_ctx* c, const HkfFrameParams*) { NOT_YET(c); }
extern "C" int hkf_raycast_tracked(hkf_ctx* c, const HkfFrameParams*) { NOT_YET(c); }
extern "C" int hkf_shard_unique_id(void*) { return HKF_ERR_STATE; }
extern "C" int hkf_shard_init(hkf_ctx* c, int, int, const void*) { NOT_YET(c); }
extern "C" int hkf_shard_composite_raycast(hkf_ctx* c) { NOT_YET(c); }
void hkf_shard_release(hkf_ctx*) {}
