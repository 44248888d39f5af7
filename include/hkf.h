/*
 * hkf.h — C ABI of the B200-native hybKinectfu hot path (libhkf_b200.so).
 *
 * This is the drop-in boundary.  It replaces, one for one, the reference's
 * `extern "C" void cuda*()` wrapper layer (src/cuda/CudaWrappers.h:22-36) and the
 * second, hidden half of that interface: the process-wide CudaDeviceDataMan singleton
 * (src/cuda/CudaDeviceDataMan.h:14-77) through which every wrapper and every host class
 * reaches device buffers by name.  Here the singleton becomes an explicit context handle
 * (needed for 8 replicas / z-slab ranks in one box), every function returns an int status
 * (the reference returns void and never checks a CUDA error), PODs travel by pointer, and
 * no torch / C++ types appear in any signature.
 *
 * Layers:
 *   1. lifetime + buffer access      (replaces CudaDeviceDataMan::init, CudaMap*::copyDataFrom/clone)
 *   2. wrapper-parity entry points   (one per live reference wrapper, same arguments, same order)
 *   3. fused fast paths              (what the retained host classes call per frame: device-resident
 *                                     pose, no per-iteration host round trip)
 *   4. z-slab sharding               (new; BASELINE.json config 4)
 *
 * All work is enqueued on the context's own CUDA stream.  Layer-2 calls do NOT synchronise
 * (the reference's cudaDeviceSynchronize() after each wrapper is a reference inefficiency, not
 * a contract); hkf_sync(), hkf_download() and hkf_frame_result() do.
 *
 * There is no CPU fallback: every entry point fails with HKF_ERR_CUDA if no sm_100 device
 * is usable.
 */
#ifndef HKF_H_
#define HKF_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ POD types */

/* == CameraParams, src/AppParams.h:35-55 (24 B) */
typedef struct HkfCameraParams {
    uint32_t cols, rows;
    float cx, cy, fx, fy;
} HkfCameraParams;

/* == Mat44, src/cuda/Mat.h:160-171: 16 row-major floats; a pose maps camera -> world,
 * translation in m[3], m[7], m[11] (Mat.h:293-295) */
typedef struct HkfMat44 { float m[16]; } HkfMat44;

/* == IntegrateParams, src/AppParams.h:80-84 */
typedef struct HkfIntegrateParams { float fSdfTruncation, fMaxIntegrateDist; } HkfIntegrateParams;
/* == RayCasterParams, src/AppParams.h:56-59 */
typedef struct HkfRayCasterParams { float fRayIncrement; } HkfRayCasterParams;

/* == Vertex / Triangle, src/cuda/MarchingcubeData.h:15-25 (24 B / 72 B) */
typedef struct HkfVertex { float pos[3]; float color[3]; } HkfVertex;
typedef struct HkfTriangle { HkfVertex v0, v1, v2; } HkfTriangle;

/* Everything CudaDeviceDataMan::init() reads from the AppParams singleton
 * (src/cuda/CudaDeviceDataMan.h:24-51), plus the z-slab owned by this context. */
typedef struct HkfInitParams {
    HkfCameraParams depth_camera;   /* _depth_camera_params */
    uint32_t pyramid_levels;        /* _icp_params.nPyramidLevels (1..3) */
    uint32_t volume_resolution;     /* _volume_params.nResolution  (R, cube R^3) */
    float    volume_size;           /* _volume_params.fVolumeMeterSize (S metres) */
    float    weight_max;            /* _volume_params.fWeightMax */
    uint32_t max_triangles;         /* _marchingcube_params.uMaxTriangles */
    /* z-slab: this context stores planes [slab_z0 - halo, slab_z1 + halo) ∩ [0,R) and owns
     * [slab_z0, slab_z1).  slab_z0 = slab_z1 = 0 means "the whole volume". */
    uint32_t slab_z0, slab_z1, slab_halo;
    /* colour path (the reference's default mode, config.ini:2 use_color=1): a non-zero use_color allocates the voxel
     * colour plane (uchar4 [z][y][x], kept apart from the packed {tsdf,weight} stream so that depth-only runs never
     * touch it; the reference stores a uchar3 inside its 12-byte Voxel, src/cuda/tsdfVolume.h:14-19), raw_rgb
     * (rgb_camera.cols x rgb_camera.rows, CudaDeviceDataMan.h:30) and raycast_rgb (depth camera size, :31).
     * Not available together with z-slabs. */
    uint32_t use_color;
    HkfCameraParams rgb_camera;     /* _rgb_camera_params */
} HkfInitParams;

/* Per-frame parameters for the fused paths: the remaining AppParams PODs the per-frame
 * path reads (src/AppParams.h:19-34,60-67,80-84). */
typedef struct HkfFrameParams {
    float trunc_min, trunc_max;             /* _depth_prepocess_params.fMinTrunc/fMaxTrunc */
    float sigma_pixel, sigma_depth;         /* fSigmaPixel / fSigmaDepth */
    HkfIntegrateParams integrate;
    HkfRayCasterParams raycast;
    float icp_dist_thres, icp_norm_sin_thres, icp_dist_shake, icp_angle_shake;   /* IcpParams */
    uint32_t sdf_max_iter; float sdf_dist_shake, sdf_angle_shake;                /* SDFTrackerParams */
    uint32_t use_color, color_angle_weight; /* _switch_params.useRGBData / colorAngleWeight (src/HybKinectfu.cpp:134-135,149) */
} HkfFrameParams;

typedef struct hkf_ctx hkf_ctx;

/* ------------------------------------------------------------------ status codes */
enum {
    HKF_OK = 0,
    HKF_ERR_ARG = 1,        /* bad argument / size mismatch */
    HKF_ERR_CUDA = 2,       /* CUDA runtime error (see hkf_last_error) */
    HKF_ERR_NOMEM = 3,
    HKF_ERR_STATE = 4,      /* call not valid in this state */
    HKF_ERR_NCCL = 5
};

/* tracking status, as the reference signals it one level up through bool findCameraPose()
 * (src/CameraPoseFinder.h:29-33) */
enum {
    HKF_TRACK_OK = 0,
    HKF_TRACK_SINGULAR = 1, /* det(AtA) < 1e-10, src/CameraPoseFinderICP.cpp:138 */
    HKF_TRACK_SHAKE = 2,    /* "camera shaking detected", src/CameraPoseFinderICP.cpp:104 */
    HKF_TRACK_NO_POSE = 3,  /* CameraPoseFinderFromFile found no record for the frame, src/CameraPoseFinderFromFile.cpp:32-65 */
    HKF_TRACK_PEER_TIMEOUT = 4 /* sharded mode only: a peer rank did not deliver its share of an exchange within ~2 s */
};

/* buffers named in src/cuda/CudaDeviceDataMan.h:54-67, in the REFERENCE layouts */
enum HkfBuffer {
    HKF_BUF_RAW_DEPTH = 0,        /* f32 [rows][cols] metres */
    HKF_BUF_TRUNC_DEPTH = 1,      /* f32 */
    HKF_BUF_FILTERED_DEPTH = 2,   /* f32 */
    HKF_BUF_NEW_VERTICES = 3,     /* float4 [rows_l][cols_l], level = 0..levels-1, camera frame */
    HKF_BUF_NEW_NORMALS = 4,
    HKF_BUF_MODEL_VERTICES = 5,   /* world frame (raycast output) */
    HKF_BUF_MODEL_NORMALS = 6,
    HKF_BUF_VOLUME = 7,           /* packed {f32 tsdf, f32 weight} [z][y][x] over the stored slab */
    HKF_BUF_SOLVER_SUMS = 8,      /* 27 f32: rigid_align_buf_reduced */
    HKF_BUF_TRIANGLES = 9,        /* HkfTriangle[n], n = hkf_triangle_count() */
    HKF_BUF_POSE = 10,            /* 16 f32: device-resident current pose */
    HKF_BUF_RAYCAST_KEYS = 11,    /* u64 [rows][cols]: nearest-hit compositing keys (sharded mode) */
    HKF_BUF_RAW_RGB = 12,         /* uchar3 [rgb rows][rgb cols]: raw_rgb, channel order as uploaded (OpenCV BGR in the reference) */
    HKF_BUF_RAYCAST_RGB = 13,     /* uchar3 [rows][cols]: raycast_rgb */
    HKF_BUF_VOLUME_COLOR = 14     /* uchar4 [z][y][x]: Voxel.color of every stored voxel (.w unused) */
};

/* ------------------------------------------------------------------ 1. lifetime, buffers */
/* replaces CudaDeviceDataMan::instance()->init() (src/cuda/CudaDeviceDataMan.h:24-51) */
int  hkf_create(const HkfInitParams* params, int device, hkf_ctx** out);
void hkf_destroy(hkf_ctx* ctx);
const char* hkf_last_error(const hkf_ctx* ctx);
int  hkf_sync(hkf_ctx* ctx);
/* raw cudaStream_t of the context, for callers that order their own work against it */
void* hkf_stream(hkf_ctx* ctx);
/* clear the TSDF (tsdfvolume::init -> clearData, src/cuda/tsdfVolume.h:29-37) and reset pose/status */
int  hkf_reset_volume(hkf_ctx* ctx);
/* CudaMap*::clone(CPU) / copyDataFrom equivalents.  `level` is ignored for non-pyramid buffers.
 * hkf_download returns the number of bytes written, or a negative HKF_ERR_*. */
long hkf_download(hkf_ctx* ctx, int buffer, unsigned level, void* dst, size_t cap);
long hkf_upload(hkf_ctx* ctx, int buffer, unsigned level, const void* src, size_t nbytes);
size_t hkf_buffer_bytes(hkf_ctx* ctx, int buffer, unsigned level);
/* number of kernels this library launched on this context so far */
uint64_t hkf_launch_count(const hkf_ctx* ctx);

/* replaces HybKinectfu::copyFrameToGPU (src/HybKinectfu.cpp:63-78): uploads the 16-bit millimetre
 * frame as is (host pointer; pinned memory makes it asynchronous) and decodes on the device with
 * the reference's exact (float)((double)mm*0.001) arithmetic.  `depth_mm` may also be a device
 * pointer (is_device != 0), in which case no copy is made. */
int  hkf_upload_depth_mm(hkf_ctx* ctx, const uint16_t* depth_mm, int is_device);
/* the colour half of HybKinectfu::copyFrameToGPU (src/HybKinectfu.cpp:80-95): 3 interleaved bytes per pixel of the
 * rgb camera's image, host or device pointer (contexts created with use_color) */
int  hkf_upload_color(hkf_ctx* ctx, const uint8_t* rgb3, int is_device);

/* ------------------------------------------------------------------ 2. wrapper parity layer */
/* cudaTruncDepth, src/cuda/DataPreprocesser.cu:80-88 */
int hkf_trunc_depth(hkf_ctx* ctx, float trunc_min, float trunc_max);
/* cudaBiliearFilterDepth, src/cuda/DataPreprocesser.cu:89-100 */
int hkf_bilateral_filter_depth(hkf_ctx* ctx, float sigma_pixel, float sigma_depth);
/* cudaCalculateNewVertices, src/cuda/VerticesNormalsCalculater.cu:67-75 */
int hkf_calculate_new_vertices(hkf_ctx* ctx, const HkfCameraParams* depth_camera_params);
/* cudaCalculateNewNormals, src/cuda/VerticesNormalsCalculater.cu:77-85 */
int hkf_calculate_new_normals(hkf_ctx* ctx);
/* cudaDownSampleNewVertices / NewNormals / ModelVertices / ModelNormals, src/cuda/sample.cu:63-112
 * (with the launch configuration the reference intended) */
int hkf_downsample_new_vertices(hkf_ctx* ctx);
int hkf_downsample_new_normals(hkf_ctx* ctx);
int hkf_downsample_model_vertices(hkf_ctx* ctx);
int hkf_downsample_model_normals(hkf_ctx* ctx);
/* cudaCalPointToPlaneErrSolverParams, src/cuda/CalPointToPlaneErrSolverParams.cu:110-129.
 * Result: 27 sums in HKF_BUF_SOLVER_SUMS. */
int hkf_cal_point_to_plane_err_solver_params(hkf_ctx* ctx, unsigned pyramid_level,
        const HkfMat44* cur_transform, const HkfMat44* last_transform_inv,
        const HkfCameraParams* depth_camera_params, float dist_thres, float norm_sin_thres);
/* cudaCalSDFSolverParams, src/cuda/CalSDFErrSolverParams.cu:110-138 */
int hkf_cal_sdf_solver_params(hkf_ctx* ctx, const HkfCameraParams* depth_camera_params,
        const HkfMat44* cur_transform);
/* cudaIntegrateVolume, src/cuda/integrateVolume.cu:78-97.  has_color: the voxel colours are blended as well
 * (tsdfVolume.h:57-74) from raw_rgb at the pixel the reference computes with its hard-coded 525 / 320 / 240
 * intrinsics (integrateVolume.cu:56-62; kept, because parity is defined by the reference), and a
 * voxel whose colour pixel lies outside the rgb image is not updated at all.  The colour is read AFTER that bounds test
 * (the reference's source reads first; its compiled code does not). */
int hkf_integrate_volume(hkf_ctx* ctx, int has_color, int use_angle_weight_color,
        const HkfMat44* transform, const HkfIntegrateParams* integrate_params,
        const HkfCameraParams* depth_camera_params, const HkfCameraParams* rgb_camera_params);
/* cudaRaycastingVolume, src/cuda/raycastingVolume.cu:158-176 */
int hkf_raycasting_volume(hkf_ctx* ctx, int has_color, const HkfMat44* transform,
        const HkfRayCasterParams* raycast_params, const HkfCameraParams* depth_camera_params,
        float near_plane, float far_plane);
/* cudaMarchingcube, src/cuda/marchingcube.cu:154-164.  Unlike the reference the triangle counter
 * is cleared first (the reference only zeroes it at allocation). */
int hkf_marchingcube(hkf_ctx* ctx, int has_color, float threshold_marchingcube);
/* MarchingcubeData::triangleNums(), src/cuda/MarchingcubeData.h:86-98 (synchronises) */
int hkf_triangle_count(hkf_ctx* ctx, unsigned* n);

/* ------------------------------------------------------------------ 3. fused fast paths */
/* decode + trunc + bilateral + vertices + normals + new-frame pyramids in two launches.
 * depth_mm: host (pinned or pageable) or device pointer, see hkf_upload_depth_mm. */
int hkf_preprocess(hkf_ctx* ctx, const uint16_t* depth_mm, int is_device, const HkfFrameParams* fp);
/* Frame pipelining: upload + preprocess the NEXT frame on a second stream into a second set of per-frame buffers
 * while the current frame is still being tracked / integrated / raycast, then make it the current frame.  Same
 * results as hkf_preprocess; hkf_swap_frame orders the main stream after the prepared frame.  A host buffer must
 * stay valid until the swap. */
int hkf_preprocess_next(hkf_ctx* ctx, const uint16_t* depth_mm, int is_device, const HkfFrameParams* fp);
int hkf_swap_frame(hkf_ctx* ctx);
/* set / get the device-resident pose (HKF_TRACK_OK is restored by set) */
int hkf_set_pose(hkf_ctx* ctx, const HkfMat44* pose);
/* whole CameraPoseFinderICP::estimateCameraPose (src/CameraPoseFinderICP.cpp:50-86) on the device:
 * model pyramids, the coarse-to-fine schedule, per iteration one launch that does correspondence,
 * the 27 sums, the 6x6 solve, the gates and the pose update.  Nothing is read back. */
int hkf_icp_track(hkf_ctx* ctx, const HkfFrameParams* fp);
/* whole CameraPoseFinderSDF::estimateCameraPose (src/CameraPoseFinderSDF.cpp:44-106) on the device */
int hkf_sdf_track(hkf_ctx* ctx, const HkfFrameParams* fp);
/* integrate / raycast with the device-resident pose; integration is skipped on the device when the
 * tracking status is not HKF_TRACK_OK (HybKinectfu::processNewFrame, src/HybKinectfu.cpp:123-154) */
int hkf_integrate_tracked(hkf_ctx* ctx, const HkfFrameParams* fp);
int hkf_raycast_tracked(hkf_ctx* ctx, const HkfFrameParams* fp);
/* blocks until everything enqueued so far is done; returns pose and tracking status */
int hkf_frame_result(hkf_ctx* ctx, HkfMat44* pose_out, int* track_status);
/* count voxels the next integration with `transform` would rewrite (N_upd of the roofline), no writes */
int hkf_count_integrate_updates(hkf_ctx* ctx, const HkfMat44* transform,
        const HkfIntegrateParams* integrate_params, const HkfCameraParams* depth_camera_params,
        unsigned long long* n_upd);

/* Work profile of the integration, plane by plane (for balancing z-slabs on MEASURED work): while enabled, every
 * integration adds, per stored plane, the number of general 4x32-voxel chunks and of free 4x128-voxel bricks it
 * processed.  hkf_integrate_profile(ctx, 1) clears and enables, (ctx, 0) disables; _read copies 2 x stored-planes
 * counters {general, free} and returns the number of entries or a negative HKF_ERR_*. */
int  hkf_integrate_profile(hkf_ctx* ctx, int enable);
long hkf_integrate_profile_read(hkf_ctx* ctx, uint32_t* out, size_t cap_entries);

/* ------------------------------------------------------------------ 4. z-slab sharding */
/* Join an NCCL communicator: `unique_id` is the 128-byte ncclUniqueId produced by
 * hkf_shard_unique_id() on rank 0 and distributed by the caller (torch.distributed). */
int hkf_shard_unique_id(void* unique_id_128);
int hkf_shard_init(hkf_ctx* ctx, int rank, int nranks, const void* unique_id_128);
/* ICP in sharded mode.  0 (default, or env HKF_SHARD_ICP unset): every rank evaluates its band of pixel rows and the 27
 * partial sums are all-reduced (peer-memory exchange inside the tracking kernel, ncclAllReduce without peer memory),
 * summed in rank order so that every rank holds the same bits.  1 (or HKF_SHARD_ICP=replicated): every rank tracks on
 * the whole image, no exchange - ICP is latency-bound, so on NVLink-connected B200s this is the faster mode.  Must be
 * set identically on all ranks. */
int hkf_shard_set_icp_mode(hkf_ctx* ctx, int replicated);
/* nearest-hit compositing of per-slab partial raycasts (min over ranks of (t, ok, rank) keys,
 * then payload from the winner) -> full model maps on every rank.  hkf_raycast_tracked schedules it by itself, DEFERRED:
 * it is issued by the next call that needs the model maps or changes the volume (hkf_icp_track, hkf_integrate_*,
 * hkf_frame_result, hkf_download, hkf_upload, hkf_reset_volume, hkf_sync, the next hkf_raycast_tracked), i.e. after the
 * next frame's hkf_preprocess, which does neither.  Its handshake is also what orders a rank's next integration after
 * the neighbours' slab raycasts that read its planes (peer-memory halo).  Call it directly only with env
 * HKF_SHARD_MANUAL_COMPOSITE=1 (every rank, once per hkf_raycast_tracked, before the next hkf_icp_track /
 * hkf_integrate_tracked). */
int hkf_shard_composite_raycast(hkf_ctx* ctx);
/* Peer-memory fast path (optional, after hkf_shard_init): each rank exports HKF_SHARD_IPC_BYTES bytes of cudaIpc
 * handles (its mailbox, its compositing inboxes and its volume); the caller gathers them from all ranks (rank order) and hands the
 * concatenation to hkf_shard_connect.  Afterwards the ICP sums are exchanged by the tracking kernel itself (P2P stores
 * of (value, iteration) words over NVLink, summed in rank order so every rank holds the same bits), the raycast kernel
 * pushes the key and payload of every pixel it hits into the inbox of every rank while it marches, and compositing is
 * one kernel over local memory; NCCL is no longer on the per-frame path.  Without it the same exchanges use
 * ncclAllReduce. */
#define HKF_SHARD_IPC_BYTES 256
int hkf_shard_ipc_handles(hkf_ctx* ctx, void* out_handles);
int hkf_shard_connect(hkf_ctx* ctx, const void* all_handles, int nranks);
/* Peer-memory halo (after hkf_shard_connect; contexts created with slab_halo = 0): the `halo` planes on either side of
 * the owned range that the slab raycast needs are read in the neighbours' memory over NVLink instead of being stored and
 * integrated a second time.  own_ranges = nranks pairs (z0, z1) of the owned planes of every rank, tiling [0, R), each at
 * least `halo` planes thick.  The raycast then waits for the neighbours' integration of the frame (one more flag). */
int hkf_shard_peer_halo(hkf_ctx* ctx, unsigned halo, const unsigned* own_ranges, int nranks);

#ifdef __cplusplus
}
#endif
#endif /* HKF_H_ */
