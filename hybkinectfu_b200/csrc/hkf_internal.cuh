// hkf_internal.cuh — context, error plumbing and the device-side arithmetic contract.
//
// The reference's float arithmetic is whatever nvcc 12.9 made of its sources with default
// flags (-fmad=true, IEEE div/sqrt).  Which a*b+c pairs were fused into FMAs decides which
// pixel a voxel projects to and which side of zero a near-zero SDF falls on, so every
// expression that feeds an index, a threshold or a parity-checked value is written here with
// explicit round-to-nearest intrinsics in the association order read off the reference's
// PTX/SASS (compiled for sm_100 by oracle/Makefile).  Nothing in this file is left to the
// compiler's contraction heuristics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/hkf.h"

#define HKF_MAX_LEVELS 3
#define HKF_SUMS 27

struct DMat44 { float m[16]; };

// Depth extrema of the current truncated depth map for range queries (integration chunk classification).  Base tiles are
// 8 x 8 pixels; tab holds, for window sizes of 2^kx x 2^ky base tiles (kx < HKF_BKX, ky < HKF_BKY), the extrema of the
// window whose top-left tile is (x, y) - a sparse table, so that the extrema over ANY rectangle of up to 32 x 8 tiles
// (256 x 64 pixels) are the combination of four overlapping power-of-two windows: four 8-byte loads, no loop.
// An entry is {float bits of the largest valid depth (0 = no valid pixel), INVERTED bits of the smallest depth with
// invalid pixels counted as 0}.  Level (0,0) is written by the producers of the truncated depth (plain stores: every
// base tile belongs to one producer block), the other levels by k_bounds_rmq; base[0] = float bits of the frame's
// largest valid depth (atomicMax, cleared before each frame).
// (Round 1 / early round 2 kept extrema of 32 x 8 pixel tiles and three coarser levels and looped over up to 4 x 4 of
// them per chunk: looser bounds and most of the classification pass's instructions, profiles/r2_integrate_*.)
#define HKF_TILE_W 32
#define HKF_TILE_H 8
#define HKF_BT 8
#define HKF_BKX 6
#define HKF_BKY 4
struct DepthBounds {
    unsigned* base;
    uint2* tab;          // [(kx * HKF_BKY + ky)][ty][tx]
    int tx, ty;          // base tiles per row / per column
};

// device-resident per-frame state shared by the fused kernels
struct HkfDeviceState {
    float pose[16];          // current camera pose (camera -> world)
    float pose_inv[16];      // its inverse (float cofactors, Mat44::getInverse arithmetic)
    float last_inv[16];      // snapshot of pose_inv at tracking start = inverse of the pose the model maps were raycast from
    float cur[16];           // ICP/SDF working transform
    int   track_status;      // HKF_TRACK_*
    int   frame_valid;       // 0 until the first frame has been integrated
    unsigned dmax_bits;      // max valid truncated depth of the current frame (float bits)
    unsigned ticket;         // last-block-done counter for the solver kernels
    unsigned gen;            // arrival counter of k_icp_frame (monotonic within a frame, cleared when tracking begins)
    unsigned comp_go;        // raycast compositing: epoch that block 0 has released to the other blocks
    unsigned mc_count;       // marching-cubes triangle counter
    unsigned mc_ncand;       // marching-cubes candidate-cell counter
    unsigned pad;
    unsigned long long upd_count;
    float sums[HKF_SUMS + 5]; // reduced 27 sums (+ padding)
    int   sdf_done;          // SDF tracker: converged flag (|x| < 1e-3)
#ifdef HKF_ICP_STAMPS          /* profiling build only (make stamps): %globaltimer stamps of the tracking launches */
    unsigned long long stamps[32][8];
    unsigned stamp_it;
#endif
    unsigned int_items[4];   // integration work-item counters [parity][free, general] (double-buffered: one pair is appended to while the other is cleared)
};

// ---- z-slab sharding over peer memory (csrc/shard.cu).  Every rank owns one mailbox that its peers write into
// through NVLink (cudaIpc-mapped pointers); flags are monotonically increasing sequence numbers, so nothing is
// ever reset.
#define HKF_MAX_RANKS 8
struct ShardMailbox {
    // [iteration parity][source rank][27 partial sums + padding]: every 8-byte word carries (float bits, iteration
    // number), stored with ONE 64-bit store, so the reader needs no separate flag and the writer no fence (the
    // low-latency protocol NCCL calls LL)
    unsigned long long sums[2][HKF_MAX_RANKS][32];
    unsigned ready[HKF_MAX_RANKS];         // raycast epoch whose keys/payload the source rank has finished writing
    unsigned integ[HKF_MAX_RANKS];         // raycast epoch for which the source rank's integration is complete (peer-memory halo)
};
struct ShardPeers {
    ShardMailbox* mail[HKF_MAX_RANKS];                 // mail[r] = rank r's mailbox as mapped in this process
    // rank r's compositing inboxes: keys_in[parity][source rank][pixel] (u64) and pay_in[parity][source rank][pixel][v|n]
    // (2 x float4).  The raycast kernel of a rank PUSHES the key / payload of every pixel it hits into the inbox of every
    // rank (posted stores over NVLink, overlapped with the march); compositing then reads local memory only.
    unsigned long long* keys_in[HKF_MAX_RANKS];
    float4* pay_in[HKF_MAX_RANKS];
    int rank, nranks;
};

// one set of per-frame input buffers (the context keeps two when frames are pipelined, preprocess.cu)
struct HkfFrameSet {
    uint16_t* depth_mm = nullptr;
    float *raw_depth = nullptr, *trunc_depth = nullptr, *filtered_depth = nullptr;
    float4* new_v[HKF_MAX_LEVELS]{};
    float4* new_n[HKF_MAX_LEVELS]{};
    DepthBounds bounds{};
};

struct hkf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    HkfInitParams ip{};
    std::string err;
    uint64_t launches = 0;

    int levels = 1;
    unsigned cols[HKF_MAX_LEVELS]{}, rows[HKF_MAX_LEVELS]{};
    // image buffers
    uint16_t* depth_mm = nullptr;     // device copy of the raw 16-bit frame
    uint16_t* depth_mm_pinned = nullptr; // pinned staging for pageable callers
    float* raw_depth = nullptr;
    float* trunc_depth = nullptr;
    float* filtered_depth = nullptr;
    DepthBounds bounds{};             // per-tile depth extrema of the current frame (integration chunk classification)
    size_t bounds_zero_bytes = 0;     // prefix of bounds.base that is cleared before each frame (the frame maximum)
    size_t bounds_tab_bytes = 0;      // size of bounds.tab
    float4* new_v[HKF_MAX_LEVELS]{};
    float4* new_n[HKF_MAX_LEVELS]{};
    float4* model_v[HKF_MAX_LEVELS]{};
    float4* model_n[HKF_MAX_LEVELS]{};
    unsigned long long* ray_keys = nullptr;
    // frame pipelining (hkf_preprocess_next / hkf_swap_frame)
    HkfFrameSet back;
    bool back_allocated = false, back_valid = false;
    const uint16_t* back_src = nullptr;
    HkfFrameParams back_fp{};
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_pre_done = nullptr, ev_swap = nullptr;
    // solver scratch: [27][nblocks] partials for the largest level
    float* partials = nullptr;
    unsigned max_blocks = 0;
    bool icp_coop = true;              // k_icp_frame: cooperative launch attribute accepted (cleared on the first refusal)
    // volume slab
    float2* vol = nullptr;
    // colour path (use_color): voxel colours as a plane of their own, the frame's colour image, the raycast colour map
    uchar4* color = nullptr;
    unsigned char* raw_rgb = nullptr;      // uchar3 [rgb rows][rgb cols]
    unsigned char* raycast_rgb = nullptr;  // uchar3 [rows][cols]
    uint4* int_items = nullptr;        // integration work items (integrate.cu: make_free_item / make_general_item)
    unsigned int_capacity = 0;
    int int_parity = 0;
    unsigned* int_hist = nullptr;      // per-plane work profile (hkf_integrate_profile)
    bool int_hist_on = false;
    bool int_attr_set[3][2]{};         // cudaFuncSetAttribute done for k_integrate<FROM_STATE, {plain, COUNT_ONLY, COLOR}> on this context's device
    int sm_count = 148;
    unsigned z_lo = 0, z_hi = 0;       // stored planes
    unsigned own_z0 = 0, own_z1 = 0;   // owned planes
    // marching cubes
    HkfTriangle* tris = nullptr;
    unsigned* mc_cand = nullptr;       // candidate cells of the two-pass extraction
    unsigned mc_cand_cap = 0;
    // device state
    HkfDeviceState* st = nullptr;
    HkfDeviceState* st_host = nullptr; // pinned mirror for result read-back
    // sharding
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    float* ray_payload = nullptr;
    ShardMailbox* mailbox = nullptr;   // this rank's mailbox (device memory, IPC-exported)
    unsigned long long* keys_in = nullptr;   // this rank's compositing inboxes (IPC-exported), see ShardPeers
    float4* pay_in = nullptr;
    ShardPeers peers{};
    bool p2p = false;                  // peers mapped: exchanges go over peer memory instead of NCCL
    // peer-memory halo (hkf_shard_peer_halo): the neighbours' planes this rank reads instead of storing them
    // sharded mode: the compositing of the last hkf_raycast_tracked is DEFERRED to the next call that needs the model maps
    // (hkf_icp_track, hkf_frame_result, hkf_download, hkf_sync, the next raycast), so that the next frame's preprocess -
    // which does not - runs while the peers' hits are still travelling and the slowest rank is still marching
    bool composite_pending = false;
    const float2* peer_vol[HKF_MAX_RANKS]{};      // rank r's stored planes as mapped in this process (host side only:
                                                  // the kernels get the two halo pointers below through VolView)
    const float2* halo_lo = nullptr; const float2* halo_hi = nullptr;
    unsigned halo_z_lo = 0, halo_z_hi = 0;
    bool icp_replicated = false;       // sharded mode: every rank tracks on the whole image, no exchange of the 27 sums
    unsigned shard_seq = 0, shard_epoch = 0;
    void* ipc_opened[HKF_MAX_RANKS][4]{};
};

#define HKF_CUDA(ctx, call)                                                        \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);       \
            return HKF_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

#define HKF_LAUNCH_CHECK(ctx)                                                      \
    do {                                                                           \
        (ctx)->launches++;                                                         \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) {                                                   \
            (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(_e);  \
            return HKF_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

static inline unsigned div_up(unsigned a, unsigned b) { return (a + b - 1) / b; }

// dynamic shared memory of a kernel: `HKF_DYN_SMEM(float, smem);` (tests/emu maps it onto a host buffer)
#ifdef HKF_EMULATE_DEVICE
#define HKF_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu_dyn_smem())
#else
#define HKF_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

#if defined(__CUDACC__) || defined(HKF_EMULATE_DEVICE)   /* the latter: tests/emu compiles kernels for the host */
// Launch with programmatic stream serialisation: the grid may become resident while its predecessor in the stream is
// still draining (every kernel launched this way starts with hkf_pdl_enter(), which releases ITS successor early and
// then waits for the predecessor's completion and memory flush), which hides the launch latency between the
// dependent kernels of a frame.
template <typename... KArgs, typename... Args>
static inline cudaError_t hkf_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
__device__ __forceinline__ void hkf_pdl_enter()
{
#ifndef HKF_EMULATE_DEVICE
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// The same with the successor released only AFTER the predecessor has completed: for a kernel whose successor reads,
// before its own wait, data written two launches back (k_icp loads the new frame's pyramid level ahead of its wait;
// with the early release above it could become resident while the kernel that builds that pyramid is still running).
__device__ __forceinline__ void hkf_pdl_enter_ordered()
{
#ifndef HKF_EMULATE_DEVICE
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif

// internal cross-file helpers (C++ linkage)
void hkf_shard_release(hkf_ctx* c);
int hkf_shard_allreduce_sums(hkf_ctx* c);
int hkf_shard_flush(hkf_ctx* c);          // run the deferred compositing, if any
int hkf_pyramids_fused(hkf_ctx* c, float4** pv, float4** pn, cudaStream_t stream = nullptr, bool track_begin = false);
int hkf_depth_bounds(hkf_ctx* c);   // depth extrema of trunc_depth (after any write to it outside the fused preprocess)
int hkf_bounds_rmq(hkf_ctx* c, const DepthBounds& B, cudaStream_t stream);   // the coarser levels of the range-query table

// ---------------------------------------------------------------- device arithmetic contract
#if defined(__CUDACC__) || defined(HKF_EMULATE_DEVICE)

// (int)(p + 0.5) as the reference compiles it: the literal 0.5 is a double, so the add is in
// fp64 and the cast truncates toward zero (src/cuda/DepthCamera.h:39-43;
// PTX: cvt.f64.f32 ; add.f64 0.5 ; cvt.rzi.s32.f64)
__device__ __forceinline__ int hkf_round_px(float p)
{
    return __double2int_rz(__dadd_rn((double)p, 0.5));
}

// projectSkeletonToScreenfloat: (v*f)/z + c   (mul, div.rn, add)
__device__ __forceinline__ float hkf_project1(float v, float f, float z, float c)
{
    return __fadd_rn(__fdiv_rn(__fmul_rn(v, f), z), c);
}

// dot(a,a) / dot(a,b) as nvcc contracts x*x' + y*y' + z*z':  fma(z,z', fma(x,x', y*y'))
__device__ __forceinline__ float hkf_dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by)));
}

// a*b - c*d as ptxas fuses it everywhere in the reference: fma(a, b, -(c*d))
__device__ __forceinline__ float hkf_det2(float a, float b, float c, float d)
{
    return __fmaf_rn(a, b, -__fmul_rn(c, d));
}

// normalize(float3) (src/cuda/cuda_declar.h:89-94): len = sqrt.rn(dot); if ((double)len < 1e-8) -> 0;
// v * (float)(1.0/len) which nvcc lowers to an IEEE float reciprocal (rcp.rn)
__device__ __forceinline__ float3 hkf_normalize(float x, float y, float z)
{
    float len = __fsqrt_rn(hkf_dot3(x, y, z, x, y, z));
    if ((double)len < 1e-8) return make_float3(0.f, 0.f, 0.f);
    float r = __frcp_rn(len);
    return make_float3(__fmul_rn(x, r), __fmul_rn(y, r), __fmul_rn(z, r));
}

__device__ __forceinline__ bool hkf_is_zero4(const float4& v)
{
    return v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f;
}

// Mat44 * (x,y,z,w): per row ((m1*x + m2*y) + m3*z) + m4*w.  The reference's PTX shows two
// different contractions of the leading pair depending on the call site:
//   LEAD_FIRST : mul(m1*x) then fma(y,m2,.), fma(z,m3,.), fma(w,m4,.)
//   LEAD_SECOND: mul(m2*y) then fma(x,m1,.), fma(z,m3,.), fma(w,m4,.)
template <bool LEAD_FIRST>
__device__ __forceinline__ float hkf_row4(float m1, float m2, float m3, float m4, float x, float y, float z, float w)
{
    float t = LEAD_FIRST ? __fmaf_rn(y, m2, __fmul_rn(x, m1)) : __fmaf_rn(x, m1, __fmul_rn(y, m2));
    return __fmaf_rn(w, m4, __fmaf_rn(z, m3, t));
}

__device__ __forceinline__ float4 ld_f4(const float4* p) { return __ldg(p); }

// flag plumbing of the peer-memory exchanges: release-store / acquire-spin at system scope, bounded
__device__ __forceinline__ void shard_signal(unsigned* p, unsigned v)
{
    __threadfence_system();
    *(volatile unsigned*)p = v;
}
__device__ __forceinline__ bool shard_wait_ge(const unsigned* p, unsigned v)
{
    const long long t0 = clock64();
    while ((int)(*(volatile const unsigned*)p - v) < 0)
        if (clock64() - t0 > 4000000000ll) return false;            // ~2 s: a peer died; do not hang the GPU
    __threadfence_system();
    return true;
}

// raw MUFU.RCP (1 ulp): only ever used to PRE-classify a value whose exact IEEE result is then either
// provably implied or recomputed with div.rn
__device__ __forceinline__ float hkf_rcp_approx(float x)
{
#ifdef HKF_EMULATE_DEVICE
    return 1.0f / x;                    // any value within the guarded error must lead to the same result
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// raw MUFU.EX2 (no scaling path for results below 2^-126, see bilateral_px_tile in preprocess.cu)
__device__ __forceinline__ float hkf_ex2_ftz(float x)
{
#ifdef HKF_EMULATE_DEVICE
    return exp2f(x);                    // the CPU oracle's stand-in for ex2.approx as well
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

#endif // __CUDACC__
