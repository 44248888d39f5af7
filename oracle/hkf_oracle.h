/* TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-frame hot path (plain C).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path never does.
 *
 * Parity status: PINNED against outputs of the reference's own unmodified CUDA kernels
 * (oracle/_ref/libhkf_ref.so run on a B200 by tests/golden/make_golden.py; fixtures committed under
 * tests/golden/).  The reference itself ships no golden vectors, KATs or tests.
 * The Eigen-backed host solves (host_math.h) are restated from Eigen's published algorithms and are
 * UNPINNED (Eigen is not in the reference tree nor in this image).
 */
#ifndef HKF_ORACLE_H
#define HKF_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcCamera { uint32_t cols, rows; float cx, cy, fx, fy; } OrcCamera;

/* src/HybKinectfu.cpp:69-75 */
void orc_decode_mm(const uint16_t* mm, float* raw, size_t n);
/* src/cuda/DataPreprocesser.cu:17-36 */
void orc_trunc(const float* raw, float* out, size_t n, float tmin, float tmax);
/* src/cuda/DataPreprocesser.cu:37-79,89-100.  __expf is a hardware approximation (MUFU.EX2): the CPU
 * value agrees to ~2 ulp of the weight, not bit for bit. */
void orc_bilateral(const float* in, float* out, int cols, int rows, float sigma_pixel, float sigma_depth);
/* src/cuda/VerticesNormalsCalculater.cu:15-32 ; float4 maps are float[rows][cols][4] */
void orc_vertices(const float* depth, float* v4, const OrcCamera* cam);
/* src/cuda/VerticesNormalsCalculater.cu:35-66 */
void orc_normals(const float* v4, float* n4, int cols, int rows);
/* src/cuda/sample.cu:16-60 with the launch the reference intended; out is (in_cols>>1) x (in_rows>>1) */
void orc_down_vertices(const float* in, float* out, int in_cols, int in_rows);
void orc_down_normals(const float* in, float* out, int in_cols, int in_rows);

/* src/cuda/integrateVolume.cu:29-76 + tsdfVolume.h:38-74; vol = {tsdf,weight}[R][R][R] (z,y,x).
 * Integrates planes [z0,z1).  Returns the number of voxels rewritten. */
uint64_t orc_integrate(float* vol, int R, float S, float wmax, const float* trunc_depth, const OrcCamera* cam,
                       const float pose[16], float sdf_trunc, float max_dist, int z0, int z1);

/* the same with has_color (src/cuda/integrateVolume.cu:52-72, tsdfVolume.h:57-74): color = uint8 [R][R][R][4] (3 used),
 * normals4 = the new frame's level-0 normal map (angle weighting), rgb = uint8 [rgb_rows][rgb_cols][3] */
uint64_t orc_integrate_color(float* vol, uint8_t* color, int R, float S, float wmax, const float* trunc_depth, const float* normals4,
                             const uint8_t* rgb, int rgb_cols, int rgb_rows, int angled, const OrcCamera* cam,
                             const float pose[16], float sdf_trunc, float max_dist, int z0, int z1);

/* src/cuda/CalPointToPlaneErrSolverParams.cu:7-129 + device_functions.h:11-55: the 27 sums with the
 * reference's exact two-stage float reduction trees (16x16 blocks, then 512-strided). */
void orc_icp_sums(const float* new_v, const float* new_n, const float* model_v, const float* model_n,
                  const OrcCamera* cam_level, const float cur[16], const float last_inv[16],
                  float dist_thres, float norm_sin_thres, float out27[27]);

/* src/cuda/raycastingVolume.cu:15-176; outputs float4 maps (world frame) */
void orc_raycast(const float* vol, int R, float S, const OrcCamera* cam, const float pose[16], float ray_inc,
                 float near_plane, float far_plane, float* out_v4, float* out_n4);

/* has_color: also raycast_rgb (uint8 [rows][cols][3]), src/cuda/raycastingVolume.cu:101-106 + tsdfVolume.h:122-148 */
void orc_raycast_color(const float* vol, const uint8_t* color, int R, float S, const OrcCamera* cam, const float pose[16], float ray_inc,
                       float near_plane, float far_plane, float* out_v4, float* out_n4, uint8_t* out_rgb3);

/* src/cuda/CalSDFErrSolverParams.cu:7-138 */
void orc_sdf_sums(const float* vol, int R, float S, const float* trunc_depth, const OrcCamera* cam,
                  const float cur[16], float out27[27]);

/* src/cuda/marchingcube.cu:5-164 (colour off). tris: float[max][3][6] (pos, colour=0). Returns the count. */
uint32_t orc_marching_cubes(const float* vol, int R, float S, float threshold, float* tris, uint32_t max_tris);

/* has_color: vertex colours as vertexInterp computes them (marchingcube.cu:5-26, 44-75) */
uint32_t orc_marching_cubes_color(const float* vol, const uint8_t* color, int R, float S, float threshold, float* tris, uint32_t max_tris);

/* tsdfvolume::interpolateSDF (src/cuda/tsdfVolume.h:98-120,151-174). returns 0/1 */
int orc_interpolate_sdf(const float* vol, int R, float S, float px, float py, float pz, float* dist);

/* host-side solves, see host_math.h */
int  orc_solve_icp(const float s27[27], float x[6]);
void orc_solve_llt(const float s27[27], float x[6]);
int  orc_vector6_to_transform(const float x[6], float angle_shake, float dist_shake, float T[16]);
void orc_sdf_compose(const float cur[16], const float x[6], float next[16]);
void orc_mat44_inverse(const float in[16], float out[16]);
void orc_mat44_mul(const float a[16], const float b[16], float c[16]);
void orc_quat_to_mat3(float w, float x, float y, float z, float R[9]);

/* whole-frame drivers mirroring HybKinectfu::processNewFrame + the pose finders.  A tracker instance
 * owns all maps and the volume. */
typedef struct OrcParams {
    OrcCamera cam;
    float trunc_min, trunc_max, sigma_pixel, sigma_depth;
    int resolution; float volume_size, weight_max;
    float sdf_trunc, max_integrate_dist, ray_increment;
    int pyramid_levels;
    float icp_dist_thres, icp_norm_sin_thres, icp_dist_shake, icp_angle_shake;
    int sdf_max_iter; float sdf_dist_shake, sdf_angle_shake;
} OrcParams;
typedef struct OrcTracker OrcTracker;
OrcTracker* orc_tracker_create(const OrcParams* p);
void orc_tracker_destroy(OrcTracker* t);
/* mode 0 ICP, 1 SDF, 2 pose given.  returns tracking status (0 ok, 1 singular, 2 shake) */
int orc_tracker_process(OrcTracker* t, const uint16_t* depth_mm, unsigned frame_id, int mode, const float* given_pose, float pose_out[16]);
float* orc_tracker_volume(OrcTracker* t);
float* orc_tracker_map(OrcTracker* t, int which /*0 new_v 1 new_n 2 model_v 3 model_n*/, int level);
float* orc_tracker_depth(OrcTracker* t, int which /*0 raw 1 trunc 2 filtered*/);
/* per-stage wall-clock seconds accumulated by orc_tracker_process: pre, track, integrate, raycast */
void orc_tracker_times(OrcTracker* t, double out4[4]);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
