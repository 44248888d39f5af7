// context.cu — context lifetime and buffer access.
//
// Replaces the CudaDeviceDataMan singleton (src/cuda/CudaDeviceDataMan.h:14-77) and the
// CudaMap1D/2D copyDataFrom/clone(CPU) traffic (src/cuda/DataMap.h:81-121,242-268) the host
// classes use to move data (src/HybKinectfu.cpp:78, src/CameraPoseFinderICP.cpp:117,
// src/MeshGeneratorMarchingcube.cpp:32).
//
// HBM layout (per context):
//   depth_mm        u16  [rows][cols]                       raw frame as uploaded
//   raw/trunc/filt  f32  [rows][cols]
//   new_v/new_n     f32x4 [levels][rows_l][cols_l]          camera frame
//   model_v/model_n f32x4 [levels][rows_l][cols_l]          world frame (raycast output)
//   vol             {f32 tsdf, f32 weight} [z_hi-z_lo][R][R] packed 8 B voxels, x fastest
//   partials        f32  [27][max_blocks]                   solver block partials
//   tris            HkfTriangle [max_triangles]
//   st              HkfDeviceState                          pose / status / counters
// A 512^3 volume is 1.07 GB, a 1024^3 slab of 128+40 planes 1.4 GB: far below 180 GB, so nothing
// is ever paged or re-allocated per frame (the reference cudaMalloc/cudaFree's inside
// cudaIntegrateVolume, src/cuda/integrateVolume.cu:87-95).
#include "hkf_internal.cuh"
#include "hkf_hostdev_math.h"

static thread_local std::string g_create_err;

extern "C" const char* hkf_last_error(const hkf_ctx* c)
{
    if (!c) return g_create_err.c_str();
    return c->err.c_str();
}

static void free_all(hkf_ctx* c)
{
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    if (c->ev_pre_done) cudaEventDestroy(c->ev_pre_done);
    if (c->ev_swap) cudaEventDestroy(c->ev_swap);
    cudaFree(c->back.depth_mm); cudaFree(c->back.raw_depth); cudaFree(c->back.trunc_depth); cudaFree(c->back.filtered_depth); cudaFree(c->back.bounds.base); cudaFree(c->back.bounds.tab);
    for (int l = 0; l < HKF_MAX_LEVELS; ++l) { cudaFree(c->back.new_v[l]); cudaFree(c->back.new_n[l]); }
    cudaFree(c->depth_mm); cudaFree(c->raw_depth); cudaFree(c->trunc_depth); cudaFree(c->filtered_depth); cudaFree(c->bounds.base); cudaFree(c->bounds.tab);
    for (int l = 0; l < HKF_MAX_LEVELS; ++l) { cudaFree(c->new_v[l]); cudaFree(c->new_n[l]); cudaFree(c->model_v[l]); cudaFree(c->model_n[l]); }
    cudaFree(c->color); cudaFree(c->raw_rgb); cudaFree(c->raycast_rgb);
    cudaFree(c->ray_keys); cudaFree(c->ray_payload); cudaFree(c->partials); cudaFree(c->vol); cudaFree(c->int_items); cudaFree(c->int_hist); cudaFree(c->tris); cudaFree(c->mc_cand); cudaFree(c->st);
    if (c->st_host) cudaFreeHost(c->st_host);
    if (c->depth_mm_pinned) cudaFreeHost(c->depth_mm_pinned);
    if (c->stream) cudaStreamDestroy(c->stream);
}

#define CREATE_CUDA(call)                                                               \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) {                                                        \
            g_create_err = std::string(#call) + ": " + cudaGetErrorString(_e);          \
            free_all(c); delete c; return (_e == cudaErrorMemoryAllocation) ? HKF_ERR_NOMEM : HKF_ERR_CUDA; \
        }                                                                               \
    } while (0)

extern "C" int hkf_create(const HkfInitParams* p, int device, hkf_ctx** out)
{
    if (!p || !out) return HKF_ERR_ARG;
    *out = nullptr;
    if (p->pyramid_levels < 1 || p->pyramid_levels > HKF_MAX_LEVELS) { g_create_err = "pyramid_levels must be 1..3 (CameraPoseFinderICP::initPoseFinder)"; return HKF_ERR_ARG; }
    if (p->volume_resolution < 4 || (p->volume_resolution & 3) || p->volume_resolution > 2048) { g_create_err = "volume_resolution must be a multiple of 4 in 4..2048"; return HKF_ERR_ARG; }
    if (p->depth_camera.cols == 0 || p->depth_camera.rows == 0) { g_create_err = "empty depth camera"; return HKF_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)"; return HKF_ERR_CUDA; }
    if (device < 0 || device >= ndev) { g_create_err = "bad device ordinal"; return HKF_ERR_ARG; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) { g_create_err = "libhkf_b200 is built for sm_100a only"; return HKF_ERR_CUDA; }

    hkf_ctx* c = new hkf_ctx();
    c->device = device;
    c->ip = *p;
    c->levels = (int)p->pyramid_levels;
    CREATE_CUDA(cudaSetDevice(device));
    {   // the per-frame chain runs at the highest priority; the side stream that prepares the next frame at the lowest
        int lo = 0, hi = 0;
        CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CREATE_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
    }
    unsigned cols = p->depth_camera.cols, rows = p->depth_camera.rows;
    for (int l = 0; l < c->levels; ++l) { c->cols[l] = cols; c->rows[l] = rows; cols >>= 1; rows >>= 1; }
    size_t n0 = (size_t)c->cols[0] * c->rows[0];
    CREATE_CUDA(cudaMalloc(&c->depth_mm, n0 * sizeof(uint16_t)));
    CREATE_CUDA(cudaMalloc(&c->raw_depth, n0 * sizeof(float)));
    CREATE_CUDA(cudaMalloc(&c->trunc_depth, n0 * sizeof(float)));
    CREATE_CUDA(cudaMalloc(&c->filtered_depth, n0 * sizeof(float)));
    CREATE_CUDA(cudaMemset(c->raw_depth, 0, n0 * sizeof(float)));
    CREATE_CUDA(cudaMemset(c->trunc_depth, 0, n0 * sizeof(float)));
    CREATE_CUDA(cudaMemset(c->filtered_depth, 0, n0 * sizeof(float)));
    {
        DepthBounds& B = c->bounds;
        B.tx = (int)div_up(c->cols[0], HKF_BT); B.ty = (int)div_up(c->rows[0], HKF_BT);
        c->bounds_zero_bytes = 4 * sizeof(unsigned);
        c->bounds_tab_bytes = (size_t)HKF_BKX * HKF_BKY * B.tx * B.ty * sizeof(uint2);
        CREATE_CUDA(cudaMalloc(&B.base, c->bounds_zero_bytes));
        CREATE_CUDA(cudaMemset(B.base, 0, c->bounds_zero_bytes));
        CREATE_CUDA(cudaMalloc(&B.tab, c->bounds_tab_bytes));
        CREATE_CUDA(cudaMemset(B.tab, 0, c->bounds_tab_bytes));
    }
    CREATE_CUDA(cudaMalloc(&c->ray_keys, 2 * n0 * sizeof(unsigned long long)));   // own keys | composited keys
    CREATE_CUDA(cudaMalloc(&c->ray_payload, n0 * 8 * sizeof(float)));
    for (int l = 0; l < c->levels; ++l) {
        size_t nb = (size_t)c->cols[l] * c->rows[l] * sizeof(float4);
        CREATE_CUDA(cudaMalloc(&c->new_v[l], nb)); CREATE_CUDA(cudaMemset(c->new_v[l], 0, nb));
        CREATE_CUDA(cudaMalloc(&c->new_n[l], nb)); CREATE_CUDA(cudaMemset(c->new_n[l], 0, nb));
        CREATE_CUDA(cudaMalloc(&c->model_v[l], nb)); CREATE_CUDA(cudaMemset(c->model_v[l], 0, nb));
        CREATE_CUDA(cudaMalloc(&c->model_n[l], nb)); CREATE_CUDA(cudaMemset(c->model_n[l], 0, nb));
    }
    // solver partials: one slot per 16x16 pixel block of level 0 (the reference's gbuf geometry,
    // src/cuda/CudaDeviceDataMan.h:43)
    c->max_blocks = div_up(c->cols[0], 16) * div_up(c->rows[0], 16);
    // (the ICP kernels store [27][256] leaf-pair sums - k_icp_frame two such buffers -, the SDF kernel [27][tiles] tile sums)
    CREATE_CUDA(cudaMalloc(&c->partials, (size_t)HKF_SUMS * (c->max_blocks > 512 ? c->max_blocks : 512) * sizeof(float)));

    unsigned R = p->volume_resolution;
    unsigned z0 = p->slab_z0, z1 = p->slab_z1;
    if (z0 == 0 && z1 == 0) z1 = R;
    if (z0 >= z1 || z1 > R) { g_create_err = "bad z-slab"; free_all(c); delete c; return HKF_ERR_ARG; }
    c->own_z0 = z0; c->own_z1 = z1;
    c->z_lo = (z0 > p->slab_halo) ? z0 - p->slab_halo : 0;
    c->z_hi = (z1 + p->slab_halo < R) ? z1 + p->slab_halo : R;
    size_t nvox = (size_t)(c->z_hi - c->z_lo) * R * R;
    CREATE_CUDA(cudaMalloc(&c->vol, nvox * sizeof(float2)));
    CREATE_CUDA(cudaMemset(c->vol, 0, nvox * sizeof(float2)));
    if (p->use_color) {
        if (!(c->z_lo == 0 && c->z_hi == R)) { g_create_err = "use_color is not available with z-slabs"; free_all(c); delete c; return HKF_ERR_ARG; }
        if (p->rgb_camera.cols == 0 || p->rgb_camera.rows == 0) { g_create_err = "use_color needs an rgb camera size"; free_all(c); delete c; return HKF_ERR_ARG; }
        const size_t nrgb = (size_t)p->rgb_camera.cols * p->rgb_camera.rows * 3;
        CREATE_CUDA(cudaMalloc(&c->color, nvox * sizeof(uchar4)));
        CREATE_CUDA(cudaMemset(c->color, 0, nvox * sizeof(uchar4)));
        CREATE_CUDA(cudaMalloc(&c->raw_rgb, nrgb));
        CREATE_CUDA(cudaMemset(c->raw_rgb, 0, nrgb));
        CREATE_CUDA(cudaMalloc(&c->raycast_rgb, n0 * 3));
        CREATE_CUDA(cudaMemset(c->raycast_rgb, 0, n0 * 3));
    }
    // at most four work items per 4 x 128 voxel brick: one per general chunk, or one for its free chunks (integrate.cu)
    c->int_capacity = 4u * (c->z_hi - c->z_lo) * (R / 4) * div_up(R, 128);
    CREATE_CUDA(cudaMalloc(&c->int_items, (size_t)c->int_capacity * sizeof(uint4)));
    c->sm_count = prop.multiProcessorCount;
    if (p->max_triangles) CREATE_CUDA(cudaMalloc(&c->tris, (size_t)p->max_triangles * sizeof(HkfTriangle)));
    CREATE_CUDA(cudaMalloc(&c->st, sizeof(HkfDeviceState)));
    CREATE_CUDA(cudaMemset(c->st, 0, sizeof(HkfDeviceState)));
    CREATE_CUDA(cudaMallocHost(&c->st_host, sizeof(HkfDeviceState)));
    CREATE_CUDA(cudaMallocHost(&c->depth_mm_pinned, n0 * sizeof(uint16_t)));
    memset(c->st_host, 0, sizeof(HkfDeviceState));
    CREATE_CUDA(cudaDeviceSynchronize());
    *out = c;
    return HKF_OK;
}

extern "C" void hkf_destroy(hkf_ctx* c)
{
    if (!c) return;
    if (c->composite_pending) { hkf_shard_flush(c); cudaStreamSynchronize(c->stream); }      // the peers wait for this rank's word
    hkf_shard_release(c);
    free_all(c);
    delete c;
}

extern "C" int hkf_sync(hkf_ctx* c)
{
    if (!c) return HKF_ERR_ARG;
    { const int rc = hkf_shard_flush(c); if (rc != HKF_OK) return rc; }
    if (c->stream2) HKF_CUDA(c, cudaStreamSynchronize(c->stream2));
    HKF_CUDA(c, cudaStreamSynchronize(c->stream));
    return HKF_OK;
}

extern "C" void* hkf_stream(hkf_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" uint64_t hkf_launch_count(const hkf_ctx* c) { return c ? c->launches : 0; }

extern "C" int hkf_reset_volume(hkf_ctx* c)
{
    if (!c) return HKF_ERR_ARG;
    { const int rc = hkf_shard_flush(c); if (rc != HKF_OK) return rc; }      // the neighbours may still be reading these planes
    size_t nvox = (size_t)(c->z_hi - c->z_lo) * c->ip.volume_resolution * c->ip.volume_resolution;
    HKF_CUDA(c, cudaMemsetAsync(c->vol, 0, nvox * sizeof(float2), c->stream));
    if (c->color) HKF_CUDA(c, cudaMemsetAsync(c->color, 0, nvox * sizeof(uchar4), c->stream));
    HKF_CUDA(c, cudaMemsetAsync(c->st, 0, sizeof(HkfDeviceState), c->stream));
    for (int l = 0; l < c->levels; ++l) {
        size_t nb = (size_t)c->cols[l] * c->rows[l] * sizeof(float4);
        HKF_CUDA(c, cudaMemsetAsync(c->model_v[l], 0, nb, c->stream));
        HKF_CUDA(c, cudaMemsetAsync(c->model_n[l], 0, nb, c->stream));
    }
    return HKF_OK;
}

static void* buf_ptr(hkf_ctx* c, int buffer, unsigned level, size_t* nbytes)
{
    *nbytes = 0;
    size_t n0 = (size_t)c->cols[0] * c->rows[0];
    switch (buffer) {
    case HKF_BUF_RAW_DEPTH: *nbytes = n0 * 4; return c->raw_depth;
    case HKF_BUF_TRUNC_DEPTH: *nbytes = n0 * 4; return c->trunc_depth;
    case HKF_BUF_FILTERED_DEPTH: *nbytes = n0 * 4; return c->filtered_depth;
    case HKF_BUF_NEW_VERTICES: case HKF_BUF_NEW_NORMALS: case HKF_BUF_MODEL_VERTICES: case HKF_BUF_MODEL_NORMALS: {
        if ((int)level >= c->levels) return nullptr;
        *nbytes = (size_t)c->cols[level] * c->rows[level] * sizeof(float4);
        float4** pyr = buffer == HKF_BUF_NEW_VERTICES ? c->new_v : buffer == HKF_BUF_NEW_NORMALS ? c->new_n
                     : buffer == HKF_BUF_MODEL_VERTICES ? c->model_v : c->model_n;
        return pyr[level];
    }
    case HKF_BUF_VOLUME: *nbytes = (size_t)(c->z_hi - c->z_lo) * c->ip.volume_resolution * c->ip.volume_resolution * sizeof(float2); return c->vol;
    case HKF_BUF_SOLVER_SUMS: *nbytes = HKF_SUMS * sizeof(float); return c->st->sums;
    case HKF_BUF_POSE: *nbytes = 16 * sizeof(float); return c->st->pose;
    case HKF_BUF_RAYCAST_KEYS: *nbytes = n0 * sizeof(unsigned long long); return c->ray_keys;
    case HKF_BUF_RAW_RGB: if (!c->raw_rgb) return nullptr; *nbytes = (size_t)c->ip.rgb_camera.cols * c->ip.rgb_camera.rows * 3; return c->raw_rgb;
    case HKF_BUF_RAYCAST_RGB: if (!c->raycast_rgb) return nullptr; *nbytes = n0 * 3; return c->raycast_rgb;
    case HKF_BUF_VOLUME_COLOR: if (!c->color) return nullptr; *nbytes = (size_t)(c->z_hi - c->z_lo) * c->ip.volume_resolution * c->ip.volume_resolution * sizeof(uchar4); return c->color;
#ifdef HKF_ICP_STAMPS
    case 100: *nbytes = sizeof(c->st->stamps); return c->st->stamps;      // profiling build only
#endif
    default: return nullptr;
    }
}

extern "C" size_t hkf_buffer_bytes(hkf_ctx* c, int buffer, unsigned level)
{
    if (!c) return 0;
    if (buffer == HKF_BUF_TRIANGLES) return (size_t)c->ip.max_triangles * sizeof(HkfTriangle);
    size_t nb; buf_ptr(c, buffer, level, &nb); return nb;
}

extern "C" long hkf_download(hkf_ctx* c, int buffer, unsigned level, void* dst, size_t cap)
{
    if (!c || !dst) return -HKF_ERR_ARG;
    if (hkf_shard_flush(c) != HKF_OK) return -HKF_ERR_CUDA;
    if (buffer == HKF_BUF_TRIANGLES) {
        unsigned n = 0;
        if (hkf_triangle_count(c, &n) != HKF_OK) return -HKF_ERR_CUDA;
        if (n > c->ip.max_triangles) n = c->ip.max_triangles;
        size_t nb = (size_t)n * sizeof(HkfTriangle);
        if (cap < nb) { c->err = "hkf_download: destination too small"; return -HKF_ERR_ARG; }
        if (nb && cudaMemcpyAsync(dst, c->tris, nb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -HKF_ERR_CUDA;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -HKF_ERR_CUDA;
        return (long)nb;
    }
    size_t nb; void* p = buf_ptr(c, buffer, level, &nb);
    if (!p) { c->err = "hkf_download: unknown buffer/level"; return -HKF_ERR_ARG; }
    if (cap < nb) { c->err = "hkf_download: destination too small"; return -HKF_ERR_ARG; }
    cudaError_t e = cudaMemcpyAsync(dst, p, nb, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->err = std::string("hkf_download: ") + cudaGetErrorString(e); return -HKF_ERR_CUDA; }
    return (long)nb;
}

extern "C" long hkf_upload(hkf_ctx* c, int buffer, unsigned level, const void* src, size_t nbytes)
{
    if (!c || !src) return -HKF_ERR_ARG;
    if (hkf_shard_flush(c) != HKF_OK) return -HKF_ERR_CUDA;
    if (buffer == HKF_BUF_POSE) {
        if (nbytes != 64) return -HKF_ERR_ARG;
        HkfMat44 m; memcpy(m.m, src, 64);
        return hkf_set_pose(c, &m) == HKF_OK ? 64 : -HKF_ERR_CUDA;
    }
    size_t nb; void* p = buf_ptr(c, buffer, level, &nb);
    if (!p || buffer == HKF_BUF_SOLVER_SUMS) { c->err = "hkf_upload: buffer not writable"; return -HKF_ERR_ARG; }
    if (nbytes != nb) { c->err = "hkf_upload: size mismatch"; return -HKF_ERR_ARG; }
    cudaError_t e = cudaMemcpyAsync(p, src, nb, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && buffer == HKF_BUF_TRUNC_DEPTH) {
        // keep the depth bounds used by integration culling consistent with the uploaded map
        if (hkf_depth_bounds(c) != HKF_OK) return -HKF_ERR_CUDA;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->err = std::string("hkf_upload: ") + cudaGetErrorString(e); return -HKF_ERR_CUDA; }
    return (long)nb;
}

// pose and its inverse travel together so that device code never inverts on its own.  They ride in the kernel's
// parameter space: no staging buffer, no host synchronisation, so a caller that supplies poses (CameraPoseFinderFromFile,
// src/CameraPoseFinderFromFile.cpp:21-89) pipelines like the tracked paths do.
struct PosePair { float pose[16], inv[16]; };
__global__ void k_set_pose(HkfDeviceState* st, PosePair p)
{
    if (threadIdx.x < 16) { st->pose[threadIdx.x] = p.pose[threadIdx.x]; st->pose_inv[threadIdx.x] = p.inv[threadIdx.x]; }
    if (threadIdx.x == 0) st->track_status = HKF_TRACK_OK;
}

extern "C" int hkf_set_pose(hkf_ctx* c, const HkfMat44* pose)
{
    if (!c || !pose) return HKF_ERR_ARG;
    PosePair p;
    memcpy(p.pose, pose->m, 64);
    hkf_mat44_inverse(pose->m, p.inv);      // host, float cofactors: Mat44::getInverse (src/cuda/Mat.h:319-444)
    k_set_pose<<<1, 32, 0, c->stream>>>(c->st, p);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_upload_color(hkf_ctx* c, const uint8_t* rgb3, int is_device)
{
    if (!c || !rgb3) return HKF_ERR_ARG;
    if (!c->raw_rgb) { c->err = "context was created without use_color"; return HKF_ERR_STATE; }
    const size_t nb = (size_t)c->ip.rgb_camera.cols * c->ip.rgb_camera.rows * 3;
    HKF_CUDA(c, cudaMemcpyAsync(c->raw_rgb, rgb3, nb, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    return HKF_OK;
}

extern "C" int hkf_frame_result(hkf_ctx* c, HkfMat44* pose_out, int* track_status)
{
    if (!c) return HKF_ERR_ARG;
    { const int rc = hkf_shard_flush(c); if (rc != HKF_OK) return rc; }
    HKF_CUDA(c, cudaMemcpyAsync(c->st_host, c->st, offsetof(HkfDeviceState, dmax_bits), cudaMemcpyDeviceToHost, c->stream));
    HKF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (pose_out) memcpy(pose_out->m, c->st_host->pose, 64);
    if (track_status) *track_status = c->st_host->track_status;
    return HKF_OK;
}
