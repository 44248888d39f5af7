// shard.cu — z-slab sharding of one volume across the GPUs of a box (BASELINE.json config 4; new, the reference
// is single-GPU).  One process per GPU; rank r stores planes [slab_z0 - halo, slab_z1 + halo) and owns
// [slab_z0, slab_z1).
//
//  * integration needs no exchange: a voxel is a pure function of (depth map, pose), so each rank integrates its
//    stored planes, halo included, and no volume byte ever crosses NVLink;
//  * ICP: every rank evaluates a band of pixel rows of each level, the 27 partial sums are summed with one
//    ncclAllReduce per iteration enqueued on the context's stream, and every rank runs the same 6x6 solve on the
//    same bits, so the device-resident poses stay identical on all ranks without ever being compared;
//  * raycast: every rank marches the global sample sequence over its stored planes and reports the first (+ -> -)
//    crossing whose later sample lies in its OWNED planes as a 64-bit key (t | ok | rank).  Compositing = the
//    minimum key over ranks (ncclAllReduce MIN on u64; a nearer failed refinement beats a farther success, as in
//    the reference where the ray ends at its first crossing), then the winner's vertex/normal: losers zero their
//    payload and an ncclAllReduce SUM delivers x + 0 + ... + 0 = x bit for bit to every rank.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the library torch already loaded in this process), so
// libhkf_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include "hkf_internal.cuh"

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint32 = 3, ncclUint64 = 5, ncclFloat32 = 7 };
enum { ncclSum = 0, ncclMin = 3 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

static NcclApi* nccl_api()
{
    static NcclApi api;
    if (api.lib || !api.err.empty()) return &api;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return &api; }
#define HKF_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name); if (!api.field) { api.err = std::string("missing NCCL symbol ") + name; api.lib = nullptr; return &api; }
    HKF_SYM(GetUniqueId, "ncclGetUniqueId")
    HKF_SYM(CommInitRank, "ncclCommInitRank")
    HKF_SYM(CommDestroy, "ncclCommDestroy")
    HKF_SYM(AllReduce, "ncclAllReduce")
    HKF_SYM(GetErrorString, "ncclGetErrorString")
#undef HKF_SYM
    return &api;
}

#define HKF_NCCL(ctx, call)                                                                         \
    do {                                                                                            \
        ncclResult_t _r = (call);                                                                   \
        if (_r != 0) { (ctx)->err = std::string(#call) + ": " + nccl_api()->GetErrorString(_r); return HKF_ERR_NCCL; } \
    } while (0)

extern "C" int hkf_shard_unique_id(void* unique_id_128)
{
    if (!unique_id_128) return HKF_ERR_ARG;
    NcclApi* a = nccl_api();
    if (!a->lib) return HKF_ERR_NCCL;
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != 0) return HKF_ERR_NCCL;
    memcpy(unique_id_128, &id, sizeof(id));
    return HKF_OK;
}

extern "C" int hkf_shard_init(hkf_ctx* c, int rank, int nranks, const void* unique_id_128)
{
    if (!c || !unique_id_128 || nranks < 1 || rank < 0 || rank >= nranks) return HKF_ERR_ARG;
    if (nranks > 255) { c->err = "at most 255 ranks (8 bits of the compositing key)"; return HKF_ERR_ARG; }
    if (c->nccl_comm) { c->err = "hkf_shard_init: already initialised"; return HKF_ERR_STATE; }
    NcclApi* a = nccl_api();
    if (!a->lib) { c->err = a->err; return HKF_ERR_NCCL; }
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof(id));
    HKF_CUDA(c, cudaSetDevice(c->device));
    ncclComm_t comm = nullptr;
    HKF_NCCL(c, a->CommInitRank(&comm, nranks, id, rank));
    c->nccl_comm = comm; c->rank = rank; c->nranks = nranks;
    return HKF_OK;
}

void hkf_shard_release(hkf_ctx* c)
{
    if (c && c->nccl_comm) { nccl_api()->CommDestroy((ncclComm_t)c->nccl_comm); c->nccl_comm = nullptr; }
}

// sum the 27 per-rank partial sums in place (st->sums), enqueued on the context's stream
int hkf_shard_allreduce_sums(hkf_ctx* c)
{
    if (!c->nccl_comm) return HKF_OK;
    HKF_NCCL(c, nccl_api()->AllReduce(c->st->sums, c->st->sums, 28, ncclFloat32, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    return HKF_OK;
}

// keep this rank's vertex/normal only where its key is the global minimum; payload = [v | n] as 8 floats per pixel
__global__ void k_composite_select(const unsigned long long* __restrict__ mine, const unsigned long long* __restrict__ best,
                                   const float4* __restrict__ v, const float4* __restrict__ n, float4* __restrict__ payload, unsigned npix)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const bool win = mine[i] != ~0ull && mine[i] == best[i];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    payload[i] = win ? v[i] : z;
    payload[npix + i] = win ? n[i] : z;
}

extern "C" int hkf_shard_composite_raycast(hkf_ctx* c)
{
    if (!c) return HKF_ERR_ARG;
    if (!c->nccl_comm) { c->err = "hkf_shard_composite_raycast: hkf_shard_init has not been called"; return HKF_ERR_STATE; }
    NcclApi* a = nccl_api();
    const unsigned npix = c->cols[0] * c->rows[0];
    unsigned long long* best = c->ray_keys + npix;            // second half of the key buffer
    HKF_NCCL(c, a->AllReduce(c->ray_keys, best, npix, ncclUint64, ncclMin, (ncclComm_t)c->nccl_comm, c->stream));
    k_composite_select<<<div_up(npix, 256), 256, 0, c->stream>>>(c->ray_keys, best, c->model_v[0], c->model_n[0], (float4*)c->ray_payload, npix);
    HKF_LAUNCH_CHECK(c);
    HKF_NCCL(c, a->AllReduce(c->ray_payload, c->ray_payload, (size_t)npix * 8, ncclFloat32, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    HKF_CUDA(c, cudaMemcpyAsync(c->model_v[0], c->ray_payload, (size_t)npix * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    HKF_CUDA(c, cudaMemcpyAsync(c->model_n[0], c->ray_payload + (size_t)npix * 4, (size_t)npix * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    return HKF_OK;
}
