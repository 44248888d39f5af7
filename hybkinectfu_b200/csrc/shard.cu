// shard.cu — z-slab sharding of one volume across the GPUs of a box (BASELINE.json config 4; new, the reference
// is single-GPU).  One process per GPU; rank r owns planes [slab_z0, slab_z1).
//
//  * integration needs no exchange: a voxel is a pure function of (depth map, pose).  The slab raycast needs `halo`
//    planes on either side of the owned range: with peer memory it reads them in the neighbours' volumes
//    (hkf_shard_peer_halo), otherwise every rank stores and integrates them itself;
//  * ICP: replicated on every rank, or banded - every rank evaluates a band of pixel rows and the 27 partial sums are
//    exchanged (inside k_icp_frame over peer memory, ncclAllReduce otherwise) and added in rank order, so that every rank
//    runs the same 6x6 solve on the same bits and the device-resident poses stay identical without ever being compared;
//  * raycast: every rank marches the global sample sequence over the planes it can read and reports the first
//    (+ -> -) crossing whose later sample lies in its OWNED planes as a 64-bit key (t | ok | rank).  Compositing = the
//    minimum key over ranks (a nearer failed refinement beats a farther success, as in the reference where the ray
//    ends at its first crossing), then the winner's vertex / normal.  Peer memory: the raycast kernel pushes its hits
//    into every rank's inbox and k_composite_p2p reads local memory only.  NCCL fallback: ncclAllReduce MIN on the keys,
//    losers zero their payload and an ncclAllReduce SUM delivers x + 0 + ... + 0 = x bit for bit to every rank.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the library torch already loaded in this process), so
// libhkf_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include "hkf_internal.cuh"

// ---------------------------------------------------------------- device kernels (tests/emu compiles these for the host)
// One kernel composites the per-slab raycasts from this rank's OWN inboxes (the peers' raycast kernels pushed their hits
// there).  Block 0 is the only one that talks to the peers' flags: it announces that this rank's raycast of the epoch -
// and with it every store it pushed - is complete (stream order + one system-scope fence), waits until every rank has
// announced the same, and releases the other blocks through a device-local word.  Then, per pixel, the minimum of the
// nranks local keys and the winner's payload; the key slots are reset to "no hit" for their next use two epochs later.
// The inboxes are double-buffered by epoch parity: a peer can only push epoch e + 2 after this rank has announced
// e + 1, i.e. after this kernel of epoch e has completed - no trailing barrier.
__global__ void __launch_bounds__(256)
k_composite_p2p(ShardPeers P, float4* __restrict__ out_v, float4* __restrict__ out_n, unsigned npix, unsigned epoch, HkfDeviceState* st)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        bool ok = true;
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (int r = 0; r < P.nranks; ++r) *(volatile unsigned*)&P.mail[r]->ready[P.rank] = epoch;
            const long long t0 = clock64();
            for (int r = 0; r < P.nranks && ok; ++r)
                while ((int)(*(volatile const unsigned*)&P.mail[P.rank]->ready[r] - epoch) < 0)
                    if (clock64() - t0 > 4000000000ll) { ok = false; break; }                          // ~2 s: a peer died
            __threadfence_system();
            if (!ok) st->track_status = HKF_TRACK_PEER_TIMEOUT;
            *(volatile unsigned*)&st->comp_go = ok ? epoch : (epoch | 0x80000000u);
        } else {
            unsigned g;
            const long long t0 = clock64();
            while ((int)(((g = *(volatile const unsigned*)&st->comp_go) & 0x7fffffffu) - epoch) < 0)
                if (clock64() - t0 > 6000000000ll) { g = 0x80000000u; break; }
            ok = !(g & 0x80000000u);
            __threadfence();
        }
        s_ok = ok;
    }
    __syncthreads();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (s_ok && i < npix) {
        // the inboxes were written by peers: read through L2 only (never a stale L1 line); all loads first
        unsigned long long* kin = P.keys_in[P.rank] + (size_t)(epoch & 1u) * HKF_MAX_RANKS * npix + i;
        unsigned long long k[HKF_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < HKF_MAX_RANKS; ++r) k[r] = r < P.nranks ? __ldcg(kin + (size_t)r * npix) : ~0ull;
        unsigned long long best = ~0ull;
#pragma unroll
        for (int r = 0; r < HKF_MAX_RANKS; ++r) best = k[r] < best ? k[r] : best;
        const int w = (int)(best & 0xffu);
        if (best == ~0ull || !((best >> 8) & 1ull)) {
            out_v[i] = make_float4(0.f, 0.f, 0.f, 0.f); out_n[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (w != P.rank) {
            const float4* pay = P.pay_in[P.rank] + 2 * (((size_t)(epoch & 1u) * HKF_MAX_RANKS + (unsigned)w) * npix + i);
            out_v[i] = __ldcg(pay); out_n[i] = __ldcg(pay + 1);
        }
#pragma unroll
        for (int r = 0; r < HKF_MAX_RANKS; ++r) if (r < P.nranks && k[r] != ~0ull) kin[(size_t)r * npix] = ~0ull;
    }
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

#ifndef HKF_EMULATE_DEVICE      /* host code below (tests/emu stops here) */
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
    c->icp_replicated = getenv("HKF_SHARD_ICP") && !strcmp(getenv("HKF_SHARD_ICP"), "replicated");
    return HKF_OK;
}

extern "C" int hkf_shard_set_icp_mode(hkf_ctx* c, int replicated)
{
    if (!c) return HKF_ERR_ARG;
    c->icp_replicated = replicated != 0;
    return HKF_OK;
}

void hkf_shard_release(hkf_ctx* c)
{
    if (!c) return;
    for (int r = 0; r < HKF_MAX_RANKS; ++r)
        for (int k = 0; k < 4; ++k)
            if (c->ipc_opened[r][k]) { cudaIpcCloseMemHandle(c->ipc_opened[r][k]); c->ipc_opened[r][k] = nullptr; }
    c->p2p = false;
    if (c->mailbox) { cudaFree(c->mailbox); c->mailbox = nullptr; }
    if (c->keys_in) { cudaFree(c->keys_in); c->keys_in = nullptr; }
    if (c->pay_in) { cudaFree(c->pay_in); c->pay_in = nullptr; }
    if (c->nccl_comm) { nccl_api()->CommDestroy((ncclComm_t)c->nccl_comm); c->nccl_comm = nullptr; }
}

// ---------------------------------------------------------------- peer-memory fast path
extern "C" int hkf_shard_ipc_handles(hkf_ctx* c, void* out)
{
    if (!c || !out) return HKF_ERR_ARG;
    static_assert(4 * sizeof(cudaIpcMemHandle_t) == HKF_SHARD_IPC_BYTES, "handle block size");
    HKF_CUDA(c, cudaSetDevice(c->device));
    if (!c->mailbox) {
        HKF_CUDA(c, cudaMalloc(&c->mailbox, sizeof(ShardMailbox)));
        HKF_CUDA(c, cudaMemset(c->mailbox, 0, sizeof(ShardMailbox)));
    }
    const size_t npix = (size_t)c->cols[0] * c->rows[0];
    if (!c->keys_in) {
        const size_t nk = 2 * (size_t)HKF_MAX_RANKS * npix;
        HKF_CUDA(c, cudaMalloc(&c->keys_in, nk * sizeof(unsigned long long)));
        HKF_CUDA(c, cudaMemset(c->keys_in, 0xff, nk * sizeof(unsigned long long)));       // "no hit" everywhere
        HKF_CUDA(c, cudaMalloc(&c->pay_in, nk * 2 * sizeof(float4)));
        HKF_CUDA(c, cudaMemset(c->pay_in, 0, nk * 2 * sizeof(float4)));
        HKF_CUDA(c, cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h[4];
    memset(h, 0, sizeof(h));
    HKF_CUDA(c, cudaIpcGetMemHandle(&h[0], c->mailbox));
    HKF_CUDA(c, cudaIpcGetMemHandle(&h[1], c->keys_in));
    HKF_CUDA(c, cudaIpcGetMemHandle(&h[2], c->pay_in));
    HKF_CUDA(c, cudaIpcGetMemHandle(&h[3], c->vol));
    memcpy(out, h, sizeof(h));
    return HKF_OK;
}

extern "C" int hkf_shard_connect(hkf_ctx* c, const void* all_handles, int nranks)
{
    if (!c || !all_handles) return HKF_ERR_ARG;
    if (nranks != c->nranks || nranks > HKF_MAX_RANKS) { c->err = "hkf_shard_connect: rank count mismatch (or > 8 ranks)"; return HKF_ERR_ARG; }
    if (!c->mailbox) { c->err = "hkf_shard_connect: call hkf_shard_ipc_handles first"; return HKF_ERR_STATE; }
    HKF_CUDA(c, cudaSetDevice(c->device));
    const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
    ShardPeers P{};
    P.rank = c->rank; P.nranks = nranks;
    for (int r = 0; r < nranks; ++r) {
        void* ptr[4];
        if (r == c->rank) {
            ptr[0] = c->mailbox; ptr[1] = c->keys_in; ptr[2] = c->pay_in; ptr[3] = c->vol;
        } else {
            for (int k = 0; k < 4; ++k) {
                HKF_CUDA(c, cudaIpcOpenMemHandle(&ptr[k], h[r * 4 + k], cudaIpcMemLazyEnablePeerAccess));
                c->ipc_opened[r][k] = ptr[k];
            }
        }
        P.mail[r] = (ShardMailbox*)ptr[0]; P.keys_in[r] = (unsigned long long*)ptr[1]; P.pay_in[r] = (float4*)ptr[2];
        c->peer_vol[r] = (const float2*)ptr[3];
    }
    c->peers = P;
    c->p2p = true;
    return HKF_OK;
}

// Peer-memory halo: instead of storing and integrating `halo` planes of its neighbours a second time (at 1024^3 over 8
// ranks that is 40 planes on top of 60-80 owned ones), a rank reads them in the neighbours' memory over NVLink: only the
// slab raycast needs them - about one 4-byte sample per ray at the slab's entry, and the refinement of crossings next to a
// boundary.  own_ranges = nranks pairs (z0, z1) of the OWNED planes of every rank (= what each rank stores: contexts are
// created with slab_halo = 0); every slab must be at least `halo` planes thick.  Call after hkf_shard_connect.
extern "C" int hkf_shard_peer_halo(hkf_ctx* c, unsigned halo, const unsigned* own_ranges, int nranks)
{
    if (!c || !own_ranges) return HKF_ERR_ARG;
    if (!c->p2p || nranks != c->nranks) { c->err = "hkf_shard_peer_halo: call hkf_shard_connect first (same rank count)"; return HKF_ERR_STATE; }
    if (c->z_lo != c->own_z0 || c->z_hi != c->own_z1) { c->err = "hkf_shard_peer_halo: the context stores a halo of its own (create it with slab_halo = 0)"; return HKF_ERR_STATE; }
    const unsigned R = c->ip.volume_resolution;
    for (int r = 0; r < nranks; ++r) {
        const unsigned z0 = own_ranges[2 * r], z1 = own_ranges[2 * r + 1];
        if (z0 >= z1 || z1 > R || (r > 0 && z0 != own_ranges[2 * r - 1]) || z1 - z0 < halo) { c->err = "hkf_shard_peer_halo: slabs must tile [0, R) and be at least `halo` planes thick"; return HKF_ERR_ARG; }
    }
    if (own_ranges[2 * c->rank] != c->own_z0 || own_ranges[2 * c->rank + 1] != c->own_z1) { c->err = "hkf_shard_peer_halo: own range mismatch"; return HKF_ERR_ARG; }
    const size_t pl = (size_t)R * R;
    c->halo_lo = c->halo_hi = nullptr;
    if (c->rank > 0 && halo) {            // planes [own_z0 - halo, own_z0) of the lower neighbour
        c->halo_z_lo = c->own_z0 - halo;
        c->halo_lo = c->peer_vol[c->rank - 1] + (size_t)(c->halo_z_lo - own_ranges[2 * (c->rank - 1)]) * pl;
    } else c->halo_z_lo = c->own_z0;
    if (c->rank + 1 < nranks && halo) {   // planes [own_z1, own_z1 + halo) of the upper neighbour: its first planes
        c->halo_z_hi = c->own_z1 + halo;
        c->halo_hi = c->peer_vol[c->rank + 1];
    } else c->halo_z_hi = c->own_z1;
    return HKF_OK;
}

extern "C" int hkf_shard_composite_raycast(hkf_ctx* c);
int hkf_shard_flush(hkf_ctx* c)
{
    if (!c || !c->composite_pending) return HKF_OK;
    return hkf_shard_composite_raycast(c);
}

// sum the 27 per-rank partial sums in place (st->sums), enqueued on the context's stream
int hkf_shard_allreduce_sums(hkf_ctx* c)
{
    if (!c->nccl_comm) return HKF_OK;
    HKF_NCCL(c, nccl_api()->AllReduce(c->st->sums, c->st->sums, 28, ncclFloat32, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    return HKF_OK;
}

extern "C" int hkf_shard_composite_raycast(hkf_ctx* c)
{
    if (!c) return HKF_ERR_ARG;
    if (!c->nccl_comm) { c->err = "hkf_shard_composite_raycast: hkf_shard_init has not been called"; return HKF_ERR_STATE; }
    c->composite_pending = false;
    NcclApi* a = nccl_api();
    const unsigned npix = c->cols[0] * c->rows[0];
    if (c->p2p) {
        k_composite_p2p<<<div_up(npix, 256), 256, 0, c->stream>>>(c->peers, c->model_v[0], c->model_n[0], npix, c->shard_epoch, c->st);
        HKF_LAUNCH_CHECK(c);
        return HKF_OK;
    }
    unsigned long long* best = c->ray_keys + npix;            // second half of the key buffer
    HKF_NCCL(c, a->AllReduce(c->ray_keys, best, npix, ncclUint64, ncclMin, (ncclComm_t)c->nccl_comm, c->stream));
    k_composite_select<<<div_up(npix, 256), 256, 0, c->stream>>>(c->ray_keys, best, c->model_v[0], c->model_n[0], (float4*)c->ray_payload, npix);
    HKF_LAUNCH_CHECK(c);
    HKF_NCCL(c, a->AllReduce(c->ray_payload, c->ray_payload, (size_t)npix * 8, ncclFloat32, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    HKF_CUDA(c, cudaMemcpyAsync(c->model_v[0], c->ray_payload, (size_t)npix * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    HKF_CUDA(c, cudaMemcpyAsync(c->model_n[0], c->ray_payload + (size_t)npix * 4, (size_t)npix * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    return HKF_OK;
}
#endif // HKF_EMULATE_DEVICE
