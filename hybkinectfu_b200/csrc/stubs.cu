// temporary: entry points not implemented yet return HKF_ERR_STATE (never a silent fallback)
#include "hkf_internal.cuh"
#define NOT_YET(c) do { if (c) (c)->err = std::string(__func__) + ": not implemented yet"; return HKF_ERR_STATE; } while (0)
extern "C" int hkf_shard_unique_id(void*) { return HKF_ERR_STATE; }
extern "C" int hkf_shard_init(hkf_ctx* c, int, int, const void*) { NOT_YET(c); }
extern "C" int hkf_shard_composite_raycast(hkf_ctx* c) { NOT_YET(c); }
void hkf_shard_release(hkf_ctx*) {}
