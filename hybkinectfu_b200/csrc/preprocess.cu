// preprocess.cu — depth decode, truncation, bilateral filter, vertex/normal maps, pyramids.
//
// Replaces reference kernels K1–K6:
//   trucDepthKernel / biateralFilterKernel      src/cuda/DataPreprocesser.cu:17-79
//   depthToVerticesKernel / verticesToNormals   src/cuda/VerticesNormalsCalculater.cu:15-66
//   pryDownVerticesKernel / pryDownNormals      src/cuda/sample.cu:16-60
// and the host mm->m loop of HybKinectfu::copyFrameToGPU (src/HybKinectfu.cpp:69-75).
//
// Design: the reference runs six launches with a device-wide sync after each and 81 uncached
// global loads per pixel in the bilateral filter.  Here the whole per-frame chain is two
// launches: (A) one tile kernel that decodes the 16-bit frame, truncates, stages the truncated
// tile (+apron) in shared memory and filters exactly its own 256 pixels; (B) one kernel that
// back-projects, differentiates and builds every coarser pyramid level of both maps with warp
// shuffles.  HBM traffic is the algorithmic minimum: 2 B/px in, 3 x 4 B + 2 x 16 B out at level 0.  The per-wrapper entry points reuse the same
// device functions on global memory so that each stage can be parity-checked in isolation.
#include "hkf_internal.cuh"

// ---------------------------------------------------------------- shared device functions

// (float)(ushort * 0.001): the multiply is in double (src/HybKinectfu.cpp:73)
__device__ __forceinline__ float decode_mm(uint16_t mm)
{
    return (float)__dmul_rn((double)mm, 0.001);
}

// trucDepthKernel: keep min < d < max (strict), else 0
__device__ __forceinline__ float trunc_depth(float d, float tmin, float tmax)
{
    return (d < tmax && d > tmin) ? d : 0.f;
}

// biateralFilterKernel body.  `at(cx,cy)` reads the truncated depth at image coordinates.
// Arithmetic as compiled (SASS of the reference):
//   arg  = fma(space2, ss, data2*sd);   w = __expf(-arg);   sum1 = fma(tmp,w,sum1);   sum2 += w
// Window rows outer, columns inner, clamped to the image.  Any valid neighbour further than
// 5*sigma_d from the centre aborts to pass-through (DataPreprocesser.cu:66-69).
template <typename At>
__device__ __forceinline__ float bilateral_px(At at, int x, int y, int cols, int rows, int radius,
                                              float ss, float sd, float thresh)
{
    float value = at(x, y);
    if (value == 0.f) return value;
    int x0 = max(x - radius, 0), x1 = min(x + radius, cols - 1);
    int y0 = max(y - radius, 0), y1 = min(y + radius, rows - 1);
    float sum1 = 0.f, sum2 = 0.f;
    for (int cy = y0; cy <= y1; ++cy) {
        for (int cx = x0; cx <= x1; ++cx) {
            float tmp = at(cx, cy);
            if (tmp == 0.f) continue;
            if (fabsf(__fsub_rn(tmp, value)) > thresh) return value;
            float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
            float d = __fsub_rn(value, tmp);
            float data2 = __fmul_rn(d, d);
            float arg = __fmaf_rn(space2, ss, __fmul_rn(data2, sd));
            float w = __expf(-arg);
            sum1 = __fmaf_rn(tmp, w, sum1);
            sum2 = __fadd_rn(sum2, w);
        }
    }
    if (sum2 > 0.f) return __fdiv_rn(sum1, sum2);
    return value;
}

// The same filter over a ZERO-PADDED shared-memory tile (`c` points at the centre, `stride` floats per tile row), for the
// configured window radius R.  Bit-identical to bilateral_px, with half the instructions per tap:
//  * pixels outside the image are zeros in the tile and zero neighbours are skipped, which is what clamping the window
//    does - so there is no border case, and the visiting order of the remaining taps is unchanged;
//  * the early return on the first outlier neighbour is a pure function of the window ("any valid neighbour further than
//    thresh"), so it becomes a flag tested once at the end and the tap body has no control flow;
//  * a skipped neighbour contributes nothing: tmp == 0 makes fma(tmp, w, sum1) == sum1 exactly, and w is replaced by 0 for sum2;
//  * |tmp - value| == |value - tmp| exactly, and dx*dx + dy*dy <= 2R^2 is exact in float whichever way it is summed;
//  * __expf(-a) is ex2.approx(a * -log2(e)) with a scaling path for results below 2^-126.  Every tap that reaches the
//    result has data2*sd <= 12.5 (|d| <= 5 sigma_d) and space2*ss <= 2R^2 / (2 (R-1)^2 / 4) < 15 (R = ceil(2 sigma_px)),
//    so a < 28 and the .ftz form - the same MUFU.EX2 without the scaling path - returns the same bits.
// Rows stay a loop (dx unrolled): the whole 81-tap body unrolled is 26 KB of straight-line code and 48 desynchronised warps
// per SM thrash the 6 KB L0 / 32 KB L1.5 instruction caches - measured slower than the rolled loop.
template <int R>
__device__ __forceinline__ float bilateral_px_tile(const float* __restrict__ c, int stride, float ss, float sd, float thresh)
{
    const float value = c[0];
    if (value == 0.f) return value;
    float sum1 = 0.f, sum2 = 0.f;
    bool outlier = false;
    const float* row = c - R * stride;
#pragma unroll 1
    for (int dy = -R; dy <= R; ++dy, row += stride) {
        const float dy2 = (float)(dy * dy);
#pragma unroll
        for (int dx = -R; dx <= R; ++dx) {
            const float tmp = row[dx];
            const bool valid = tmp != 0.f;
            const float d = __fsub_rn(value, tmp);
            outlier |= valid & (fabsf(d) > thresh);
            const float space2 = __fadd_rn(dy2, (float)(dx * dx));
            const float arg = __fmaf_rn(space2, ss, __fmul_rn(__fmul_rn(d, d), sd));
            const float w = hkf_ex2_ftz(__fmul_rn(arg, -1.4426950216293334961f));
            sum1 = __fmaf_rn(tmp, w, sum1);
            sum2 = __fadd_rn(sum2, valid ? w : 0.f);
        }
    }
    if (outlier) return value;
    if (sum2 > 0.f) return __fdiv_rn(sum1, sum2);
    return value;
}

// DepthCamera::depthToSkeleton (src/cuda/DepthCamera.h:19-29): (z*(u-cx))/fx
__device__ __forceinline__ float4 back_project(unsigned ux, unsigned uy, float depth, float cx, float cy, float fx, float fy)
{
    if (depth == 0.f) return make_float4(0.f, 0.f, 0.f, 0.f);
    float vx = __fdiv_rn(__fmul_rn(depth, __fsub_rn((float)ux, cx)), fx);
    float vy = __fdiv_rn(__fmul_rn(depth, __fsub_rn((float)uy, cy)), fy);
    return make_float4(vx, vy, depth, 1.0f);
}

// verticesToNormalsKernel body: normalize(cross(up - down, right - left)); any invalid -> 0
__device__ __forceinline__ float4 normal_from(const float4& v0, const float4& vr, const float4& vu,
                                              const float4& vl, const float4& vd)
{
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v0.z == 0.f || vr.z == 0.f || vu.z == 0.f || vl.z == 0.f || vd.z == 0.f) return zero;
    float ax = __fsub_rn(vu.x, vd.x), ay = __fsub_rn(vu.y, vd.y), az = __fsub_rn(vu.z, vd.z);
    float bx = __fsub_rn(vr.x, vl.x), by = __fsub_rn(vr.y, vl.y), bz = __fsub_rn(vr.z, vl.z);
    float cxp = hkf_det2(ay, bz, az, by);
    float cyp = hkf_det2(az, bx, ax, bz);
    float czp = hkf_det2(ax, by, ay, bx);
    float3 n = hkf_normalize(cxp, cyp, czp);
    return make_float4(n.x, n.y, n.z, 0.f);
}

// pryDownVerticesKernel: ((p00+p01)+p10)+p11 then *0.25, all four components; 0 if any z==0
__device__ __forceinline__ float4 down_vertex(const float4& p00, const float4& p01, const float4& p10, const float4& p11)
{
    if (p00.z == 0.f || p01.z == 0.f || p10.z == 0.f || p11.z == 0.f) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r;
    r.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00.x, p01.x), p10.x), p11.x), 0.25f);
    r.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00.y, p01.y), p10.y), p11.y), 0.25f);
    r.z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00.z, p01.z), p10.z), p11.z), 0.25f);
    r.w = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00.w, p01.w), p10.w), p11.w), 0.25f);
    return r;
}
// pryDownNormalsKernel: 0 if any of the four is all-zero; mean then renormalise
__device__ __forceinline__ float4 down_normal(const float4& p00, const float4& p01, const float4& p10, const float4& p11)
{
    if (hkf_is_zero4(p01) || hkf_is_zero4(p10) || hkf_is_zero4(p00) || hkf_is_zero4(p11)) return make_float4(0.f, 0.f, 0.f, 0.f);
    float x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00.x, p01.x), p10.x), p11.x), 0.25f);
    float y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00.y, p01.y), p10.y), p11.y), 0.25f);
    float z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p00.z, p01.z), p10.z), p11.z), 0.25f);
    float3 n = hkf_normalize(x, y, z);
    return make_float4(n.x, n.y, n.z, 0.f);
}

// ---------------------------------------------------------------- per-wrapper kernels (layer 2)

__global__ void k_decode_mm(const uint16_t* __restrict__ mm, float* __restrict__ raw, unsigned n)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) raw[i] = decode_mm(mm[i]);
}

__global__ void k_trunc(const float* __restrict__ raw, float* __restrict__ out, unsigned n, float tmin, float tmax)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = trunc_depth(raw[i], tmin, tmax);
}

__global__ void k_bilateral(const float* __restrict__ in, float* __restrict__ out, int cols, int rows, int radius,
                            float ss, float sd, float thresh)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    auto at = [&](int cx, int cy) { return __ldg(in + (size_t)cy * cols + cx); };
    out[(size_t)y * cols + x] = bilateral_px(at, x, y, cols, rows, radius, ss, sd, thresh);
}

__global__ void k_vertices(const float* __restrict__ depth, float4* __restrict__ out, unsigned cols, unsigned rows,
                           float cx, float cy, float fx, float fy)
{
    unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    out[(size_t)y * cols + x] = back_project(x, y, depth[(size_t)y * cols + x], cx, cy, fx, fy);
}

__global__ void k_normals(const float4* __restrict__ v, float4* __restrict__ out, int cols, int rows)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(x == cols - 1 || y == rows - 1 || x == 0 || y == 0)) {
        size_t i = (size_t)y * cols + x;
        r = normal_from(ld_f4(v + i), ld_f4(v + i + 1), ld_f4(v + i + cols), ld_f4(v + i - 1), ld_f4(v + i - cols));
    }
    out[(size_t)y * cols + x] = r;
}

template <bool NORMALS>
__global__ void k_down(const float4* __restrict__ in, float4* __restrict__ out, unsigned in_cols, unsigned out_cols, unsigned out_rows)
{
    unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out_cols || y >= out_rows) return;
    const float4* p = in + (size_t)(2 * y) * in_cols + 2 * x;
    float4 p00 = ld_f4(p), p01 = ld_f4(p + 1), p10 = ld_f4(p + in_cols), p11 = ld_f4(p + in_cols + 1);
    out[(size_t)y * out_cols + x] = NORMALS ? down_normal(p00, p01, p10, p11) : down_vertex(p00, p01, p10, p11);
}

#define TILE_W 32
#define TILE_H 8

// ---------------------------------------------------------------- depth range-query table (integration chunk classification)
// A producer block of 32 x 8 threads (one pixel each: warp = one row of 32 pixels) publishes the extrema of its four
// 8 x 8 base tiles.  t = the pixel's truncated depth (0 = invalid), in_image = the pixel exists.  Call from all threads;
// contains two block barriers.  s_ext: 8 words of shared memory.
__device__ __forceinline__ void hkf_bounds_publish_block(const DepthBounds& B, int block_x, int block_y, float t, bool in_image, unsigned* s_ext)
{
    const int tid = threadIdx.y * HKF_TILE_W + threadIdx.x;
    if (tid < 8) s_ext[tid] = (tid < 4) ? 0u : 0xffffffffu;
    __syncthreads();
    unsigned mx = in_image ? __float_as_uint(fmaxf(t, 0.f)) : 0u, mn = in_image ? __float_as_uint(fmaxf(t, 0.f)) : 0xffffffffu;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 7) == 0) { atomicMax(&s_ext[threadIdx.x >> 3], mx); atomicMin(&s_ext[4 + (threadIdx.x >> 3)], mn); }
    __syncthreads();
    if (tid < 4) {
        const int bx = block_x * 4 + tid;
        if (bx < B.tx && block_y < B.ty) B.tab[(size_t)block_y * B.tx + bx] = make_uint2(s_ext[tid], ~s_ext[4 + tid]);
        if (s_ext[tid]) atomicMax(&B.base[0], s_ext[tid]);
    }
}

// one block per 32x8-pixel tile publishes the depth extrema of its four 8x8 base tiles (hkf_internal.cuh: DepthBounds)
__global__ void __launch_bounds__(HKF_TILE_W * HKF_TILE_H)
k_depth_bounds(const float* __restrict__ d, int cols, int rows, DepthBounds B)
{
    __shared__ unsigned s_ext[8];
    const int x = blockIdx.x * HKF_TILE_W + threadIdx.x, y = blockIdx.y * HKF_TILE_H + threadIdx.y;
    const bool in = (x < cols && y < rows);
    hkf_bounds_publish_block(B, blockIdx.x, blockIdx.y, in ? d[(size_t)y * cols + x] : 0.f, in, s_ext);
}

// The coarser levels of the range-query table from the base tiles: block (y, ky) reduces the 2^ky tile rows below y
// column by column into shared memory, then every thread forms the windows of 2^kx columns that start at its x.
__global__ void __launch_bounds__(128)
k_bounds_rmq(DepthBounds B)
{
    __shared__ uint2 s_col[512];
    hkf_pdl_enter();
    const int y = blockIdx.x, ky = blockIdx.y;
    const int h = min(1 << ky, B.ty - y);
    for (int x = threadIdx.x; x < B.tx; x += blockDim.x) {
        uint2 a = make_uint2(0u, 0u);
        for (int r = 0; r < h; ++r) {
            const uint2 e = B.tab[(size_t)(y + r) * B.tx + x];
            a.x = max(a.x, e.x); a.y = max(a.y, e.y);
        }
        s_col[x] = a;
    }
    __syncthreads();
    const size_t plane = (size_t)B.tx * B.ty;
    for (int x = threadIdx.x; x < B.tx; x += blockDim.x) {
        uint2 a = make_uint2(0u, 0u);
        int done = 0;
#pragma unroll
        for (int kx = 0; kx < HKF_BKX; ++kx) {
            const int w = min(1 << kx, B.tx - x);
            for (; done < w; ++done) { const uint2 e = s_col[x + done]; a.x = max(a.x, e.x); a.y = max(a.y, e.y); }
            if (kx | ky) B.tab[(size_t)(kx * HKF_BKY + ky) * plane + (size_t)y * B.tx + x] = a;
        }
    }
}

// ---------------------------------------------------------------- fused kernel B: all coarser levels
// Thread per level-2 pixel: builds its four level-1 pixels (vertices and normals) from the 4x4
// level-0 block, then the level-2 pixel from those, entirely in registers.  With two levels only
// the level-1 part runs.
__global__ void k_pyramids_fused(const float4* __restrict__ v0, const float4* __restrict__ n0,
                                 float4* __restrict__ v1, float4* __restrict__ n1,
                                 float4* __restrict__ v2, float4* __restrict__ n2,
                                 unsigned c0, unsigned c1, unsigned r1, unsigned c2, unsigned r2, HkfDeviceState* track_begin)
{
    // the launch that starts a frame's tracking is followed by k_icp, which preloads the NEW frame's pyramid before its
    // grid-dependency wait: release it only once the launch that built that pyramid (our predecessor) has completed
    if (track_begin != nullptr) hkf_pdl_enter_ordered(); else hkf_pdl_enter();
    // start of a frame's tracking, folded into the model-pyramid launch (saves a launch): cur = pose,
    // last_inv = pose^-1 (CameraPoseFinderICP.cpp:62-63), status cleared
    if (track_begin != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.y == 0) {
        HkfDeviceState* st = track_begin;
        if (threadIdx.x < 16) { st->cur[threadIdx.x] = st->pose[threadIdx.x]; st->last_inv[threadIdx.x] = st->pose_inv[threadIdx.x]; }
        if (threadIdx.x == 0) { st->track_status = HKF_TRACK_OK; st->ticket = 0; st->gen = 0; st->sdf_done = 0; }
    }
    unsigned X = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned Y = blockIdx.y * blockDim.y + threadIdx.y;
    float4 lv[2][2], ln[2][2];
    bool have[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            unsigned x1 = 2 * X + i, y1 = 2 * Y + j;
            have[j][i] = (x1 < c1 && y1 < r1);
            if (have[j][i]) {
                const float4* pv = v0 + (size_t)(2 * y1) * c0 + 2 * x1;
                const float4* pn = n0 + (size_t)(2 * y1) * c0 + 2 * x1;
                lv[j][i] = down_vertex(ld_f4(pv), ld_f4(pv + 1), ld_f4(pv + c0), ld_f4(pv + c0 + 1));
                ln[j][i] = down_normal(ld_f4(pn), ld_f4(pn + 1), ld_f4(pn + c0), ld_f4(pn + c0 + 1));
                v1[(size_t)y1 * c1 + x1] = lv[j][i];
                n1[(size_t)y1 * c1 + x1] = ln[j][i];
            }
        }
    if (v2 != nullptr && X < c2 && Y < r2) {
        v2[(size_t)Y * c2 + X] = down_vertex(lv[0][0], lv[0][1], lv[1][0], lv[1][1]);
        n2[(size_t)Y * c2 + X] = down_normal(ln[0][0], ln[0][1], ln[1][0], ln[1][1]);
    }
}

// ---------------------------------------------------------------- fused kernels A + A'
// A: one 32x8 pixel tile per block filters exactly its own 256 pixels from a zero-padded shared-memory tile (apron =
// window radius) and stops at the filtered depth.  A': vertices, normals and BOTH coarser pyramid levels of the new frame
// come from one launch that reads the filtered map (L2-resident) with 8x4-pixel warp patches, so that the 2x2 and 4x4
// blocks of the pyramid are warp shuffles (lane ^ 1, ^ 8, ^ 9, then ^ 2, ^ 16, ^ 18).  The first design (round 1) formed
// the normals inside the tile kernel, which made every tile filter a 1-pixel halo as well (340 instead of 256 bilateral
// evaluations, and the filter is all of the kernel's time): 35.2 -> 28.9 us per 640x480 frame (profiles/r2_experimental.log).
__global__ void __launch_bounds__(TILE_W * TILE_H)
k_preprocess_tile(const uint16_t* __restrict__ mm, float* __restrict__ raw_out, float* __restrict__ trunc_out,
                     float* __restrict__ filt_out, int cols, int rows, int radius, float tmin, float tmax, float ss, float sd,
                     float thresh, DepthBounds B)
{
    HKF_DYN_SMEM(float, smem);
    __shared__ unsigned s_ext[8];
    const int ap = radius;
    const int tw = TILE_W + 2 * ap, th = TILE_H + 2 * ap;
    float* s_trunc = smem;              // tw*th, 0 outside the image
    const int tid = threadIdx.y * TILE_W + threadIdx.x;
    const int bx0 = blockIdx.x * TILE_W, by0 = blockIdx.y * TILE_H;
    for (int i = tid; i < tw * th; i += TILE_W * TILE_H) {
        int lx = i % tw, ly = i / tw;
        int gx = bx0 - ap + lx, gy = by0 - ap + ly;
        float t = 0.f;
        if (gx >= 0 && gx < cols && gy >= 0 && gy < rows) {
            float d = decode_mm(__ldg(mm + (size_t)gy * cols + gx));
            t = trunc_depth(d, tmin, tmax);
            if (lx >= ap && lx < ap + TILE_W && ly >= ap && ly < ap + TILE_H) {
                raw_out[(size_t)gy * cols + gx] = d;
                trunc_out[(size_t)gy * cols + gx] = t;
            }
        }
        s_trunc[i] = t;
    }
    __syncthreads();
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    // depth extrema of the block's four 8x8 base tiles (thread = its own pixel of the staged tile)
    hkf_bounds_publish_block(B, blockIdx.x, blockIdx.y, s_trunc[(threadIdx.y + ap) * tw + (threadIdx.x + ap)], x < cols && y < rows, s_ext);
    if (x >= cols || y >= rows) return;
    auto at = [&](int gx, int gy) { return s_trunc[(gy - by0 + ap) * tw + (gx - bx0 + ap)]; };
    float f;
    if (radius == 4) f = bilateral_px_tile<4>(s_trunc + (threadIdx.y + ap) * tw + (threadIdx.x + ap), tw, ss, sd, thresh);
    else f = bilateral_px(at, x, y, cols, rows, radius, ss, sd, thresh);
    filt_out[(size_t)y * cols + x] = f;
}

__global__ void __launch_bounds__(256)
k_maps_pyramids(const float* __restrict__ filt, float4* __restrict__ v0o, float4* __restrict__ n0o,
                   float4* __restrict__ v1, float4* __restrict__ n1, float4* __restrict__ v2, float4* __restrict__ n2,
                   int cols, int rows, unsigned c1, unsigned r1, unsigned c2, unsigned r2, float cx, float cy, float fx, float fy)
{
    hkf_pdl_enter();
    // 16x16 pixel block, warps as 8x4 patches (patch origins are multiples of 8 and 4: 2x2 and 4x4 blocks never straddle warps)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), n = v;
    if (x < cols && y < rows) {
        const size_t gi = (size_t)y * cols + x;
        v = back_project(x, y, filt[gi], cx, cy, fx, fy);
        if (!(x == cols - 1 || y == rows - 1 || x == 0 || y == 0)) {
            const float4 vr = back_project(x + 1, y, filt[gi + 1], cx, cy, fx, fy);
            const float4 vu = back_project(x, y + 1, filt[gi + cols], cx, cy, fx, fy);
            const float4 vl = back_project(x - 1, y, filt[gi - 1], cx, cy, fx, fy);
            const float4 vd = back_project(x, y - 1, filt[gi - cols], cx, cy, fx, fy);
            n = normal_from(v, vr, vu, vl, vd);
        }
        v0o[gi] = v;
        n0o[gi] = n;
    }
    if (v1 == nullptr) return;                                   // one level only (uniform)
    auto sh = [&](const float4& a, int m) {
        return make_float4(__shfl_xor_sync(0xffffffffu, a.x, m), __shfl_xor_sync(0xffffffffu, a.y, m),
                           __shfl_xor_sync(0xffffffffu, a.z, m), __shfl_xor_sync(0xffffffffu, a.w, m));
    };
    // level 1: the lane with even x and even y holds p00; p01 = x+1 (lane^1), p10 = y+1 (lane^8), p11 (lane^9)
    const float4 v01 = sh(v, 1), v10 = sh(v, 8), v11 = sh(v, 9);
    const float4 n01 = sh(n, 1), n10 = sh(n, 8), n11 = sh(n, 9);
    float4 lv = make_float4(0.f, 0.f, 0.f, 0.f), ln = lv;
    const unsigned x1 = (unsigned)x >> 1, y1 = (unsigned)y >> 1;
    if (!(lane & 1) && !(lane & 8) && x1 < c1 && y1 < r1) {
        lv = down_vertex(v, v01, v10, v11);
        ln = down_normal(n, n01, n10, n11);
        v1[(size_t)y1 * c1 + x1] = lv;
        n1[(size_t)y1 * c1 + x1] = ln;
    }
    if (v2 == nullptr) return;
    // level 2: level-1 pixels sit on the lanes with even x, y; p01 = x+2 (lane^2), p10 = y+2 (lane^16), p11 (lane^18)
    const float4 w01 = sh(lv, 2), w10 = sh(lv, 16), w11 = sh(lv, 18);
    const float4 m01 = sh(ln, 2), m10 = sh(ln, 16), m11 = sh(ln, 18);
    const unsigned x2 = (unsigned)x >> 2, y2 = (unsigned)y >> 2;
    if (!(lane & 3) && !(lane & 24) && x2 < c2 && y2 < r2) {
        v2[(size_t)y2 * c2 + x2] = down_vertex(lv, w01, w10, w11);
        n2[(size_t)y2 * c2 + x2] = down_normal(ln, m01, m10, m11);
    }
}

// ---------------------------------------------------------------- host entry points

static inline void bilateral_consts(float sigma_pixel, float sigma_depth, int* radius, float* ss, float* sd, float* thresh)
{
    // host wrapper arithmetic: 0.5 / (sigma*sigma) in double (float*float promoted), narrowed
    // (src/cuda/DataPreprocesser.cu:95-96); kernel: (int)ceil(2.0*sigma_pixel), 5*sigma_data (:45,66)
    float s2d = sigma_depth * sigma_depth;
    float s2p = sigma_pixel * sigma_pixel;
    *sd = (float)(0.5 / (double)s2d);
    *ss = (float)(0.5 / (double)s2p);
    *radius = (int)ceil(2.0 * (double)sigma_pixel);
    *thresh = sigma_depth * 5.0f;
}

#ifndef HKF_EMULATE_DEVICE      /* host launch code below (tests/emu stops here) */

extern "C" int hkf_upload_depth_mm(hkf_ctx* c, const uint16_t* depth_mm, int is_device)
{
    if (!c || !depth_mm) return HKF_ERR_ARG;
    unsigned n = c->cols[0] * c->rows[0];
    const uint16_t* src = depth_mm;
    if (!is_device) {
        HKF_CUDA(c, cudaMemcpyAsync(c->depth_mm, depth_mm, n * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
        src = c->depth_mm;
    }
    k_decode_mm<<<div_up(n, 256), 256, 0, c->stream>>>(src, c->raw_depth, n);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_trunc_depth(hkf_ctx* c, float tmin, float tmax)
{
    if (!c) return HKF_ERR_ARG;
    unsigned n = c->cols[0] * c->rows[0];
    k_trunc<<<div_up(n, 256), 256, 0, c->stream>>>(c->raw_depth, c->trunc_depth, n, tmin, tmax);
    HKF_LAUNCH_CHECK(c);
    return hkf_depth_bounds(c);
}

extern "C" int hkf_bilateral_filter_depth(hkf_ctx* c, float sigma_pixel, float sigma_depth)
{
    if (!c) return HKF_ERR_ARG;
    int radius; float ss, sd, thresh;
    bilateral_consts(sigma_pixel, sigma_depth, &radius, &ss, &sd, &thresh);
    dim3 b(32, 8), g(div_up(c->cols[0], 32), div_up(c->rows[0], 8));
    k_bilateral<<<g, b, 0, c->stream>>>(c->trunc_depth, c->filtered_depth, c->cols[0], c->rows[0], radius, ss, sd, thresh);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_calculate_new_vertices(hkf_ctx* c, const HkfCameraParams* cam)
{
    if (!c || !cam) return HKF_ERR_ARG;
    dim3 b(32, 8), g(div_up(c->cols[0], 32), div_up(c->rows[0], 8));
    k_vertices<<<g, b, 0, c->stream>>>(c->filtered_depth, c->new_v[0], c->cols[0], c->rows[0], cam->cx, cam->cy, cam->fx, cam->fy);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

extern "C" int hkf_calculate_new_normals(hkf_ctx* c)
{
    if (!c) return HKF_ERR_ARG;
    dim3 b(32, 8), g(div_up(c->cols[0], 32), div_up(c->rows[0], 8));
    k_normals<<<g, b, 0, c->stream>>>(c->new_v[0], c->new_n[0], c->cols[0], c->rows[0]);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

template <bool NORMALS>
static int downsample(hkf_ctx* c, float4** pyr)
{
    for (int l = 0; l < c->levels - 1; ++l) {
        dim3 b(32, 8), g(div_up(c->cols[l + 1], 32), div_up(c->rows[l + 1], 8));
        k_down<NORMALS><<<g, b, 0, c->stream>>>(pyr[l], pyr[l + 1], c->cols[l], c->cols[l + 1], c->rows[l + 1]);
        HKF_LAUNCH_CHECK(c);
    }
    return HKF_OK;
}
extern "C" int hkf_downsample_new_vertices(hkf_ctx* c) { return c ? downsample<false>(c, c->new_v) : HKF_ERR_ARG; }
extern "C" int hkf_downsample_new_normals(hkf_ctx* c) { return c ? downsample<true>(c, c->new_n) : HKF_ERR_ARG; }
extern "C" int hkf_downsample_model_vertices(hkf_ctx* c) { return c ? downsample<false>(c, c->model_v) : HKF_ERR_ARG; }
extern "C" int hkf_downsample_model_normals(hkf_ctx* c) { return c ? downsample<true>(c, c->model_n) : HKF_ERR_ARG; }

// all coarser levels of one (vertices, normals) pyramid pair in a single launch
int hkf_pyramids_fused(hkf_ctx* c, float4** pv, float4** pn, cudaStream_t stream, bool track_begin)
{
    if (stream == nullptr) stream = c->stream;
    if (c->levels < 2) return HKF_OK;
    bool three = c->levels >= 3;
    unsigned gx = div_up(c->cols[1], 2), gy = div_up(c->rows[1], 2);
    dim3 b(32, 4), g(div_up(gx, 32), div_up(gy, 4));
    hkf_launch_pdl(k_pyramids_fused, g, b, 0, stream, (const float4*)pv[0], (const float4*)pn[0], pv[1], pn[1], three ? pv[2] : (float4*)nullptr, three ? pn[2] : (float4*)nullptr,
                   c->cols[0], c->cols[1], c->rows[1], three ? c->cols[2] : 0u, three ? c->rows[2] : 0u, track_begin ? c->st : (HkfDeviceState*)nullptr);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

int hkf_bounds_rmq(hkf_ctx* c, const DepthBounds& B, cudaStream_t stream)
{
    if (B.tx > 512) { c->err = "image wider than 4096 pixels"; return HKF_ERR_ARG; }
    hkf_launch_pdl(k_bounds_rmq, dim3(B.ty, HKF_BKY), dim3(128), 0, stream, B);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

int hkf_depth_bounds(hkf_ctx* c)
{
    HKF_CUDA(c, cudaMemsetAsync(c->bounds.base, 0, c->bounds_zero_bytes, c->stream));
    dim3 b(HKF_TILE_W, HKF_TILE_H), g(div_up(c->cols[0], HKF_TILE_W), div_up(c->rows[0], HKF_TILE_H));
    k_depth_bounds<<<g, b, 0, c->stream>>>(c->trunc_depth, (int)c->cols[0], (int)c->rows[0], c->bounds);
    HKF_LAUNCH_CHECK(c);
    return hkf_bounds_rmq(c, c->bounds, c->stream);
}

// the fused preprocess of one frame into a given set of per-frame buffers, on a given stream
static int run_preprocess(hkf_ctx* c, cudaStream_t stream, HkfFrameSet& F, const uint16_t* depth_mm, int is_device, const HkfFrameParams* fp)
{
    unsigned n = c->cols[0] * c->rows[0];
    const uint16_t* src = depth_mm;
    if (!is_device) {
        HKF_CUDA(c, cudaMemcpyAsync(F.depth_mm, depth_mm, n * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));
        src = F.depth_mm;
    }
    int radius; float ss, sd, thresh;
    bilateral_consts(fp->sigma_pixel, fp->sigma_depth, &radius, &ss, &sd, &thresh);
    if (radius < 0 || radius > 12) { c->err = "hkf_preprocess: bilateral radius out of range for the fused kernel"; return HKF_ERR_ARG; }
    HKF_CUDA(c, cudaMemsetAsync(F.bounds.base, 0, c->bounds_zero_bytes, stream));
    static_assert(TILE_W == HKF_TILE_W && TILE_H == HKF_TILE_H, "preprocess tile = integration culling tile");
    dim3 b(TILE_W, TILE_H), g(div_up(c->cols[0], TILE_W), div_up(c->rows[0], TILE_H));
    const HkfCameraParams& cam = c->ip.depth_camera;
    const size_t smem2 = (size_t)(TILE_W + 2 * radius) * (TILE_H + 2 * radius) * sizeof(float);
    k_preprocess_tile<<<g, b, smem2, stream>>>(src, F.raw_depth, F.trunc_depth, F.filtered_depth, (int)c->cols[0], (int)c->rows[0], radius,
                                               fp->trunc_min, fp->trunc_max, ss, sd, thresh, F.bounds);
    HKF_LAUNCH_CHECK(c);
    // coarser levels of the depth range-query table (integration): a tiny launch, in front of the maps so that their blocks
    // become resident behind it
    { const int rc = hkf_bounds_rmq(c, F.bounds, stream); if (rc != HKF_OK) return rc; }
    const bool two = c->levels >= 2, three = c->levels >= 3;
    hkf_launch_pdl(k_maps_pyramids, dim3(div_up(c->cols[0], 16), div_up(c->rows[0], 16)), dim3(256), 0, stream, (const float*)F.filtered_depth,
                   F.new_v[0], F.new_n[0], two ? F.new_v[1] : (float4*)nullptr, two ? F.new_n[1] : (float4*)nullptr,
                   three ? F.new_v[2] : (float4*)nullptr, three ? F.new_n[2] : (float4*)nullptr, (int)c->cols[0], (int)c->rows[0],
                   two ? c->cols[1] : 0u, two ? c->rows[1] : 0u, three ? c->cols[2] : 0u, three ? c->rows[2] : 0u, cam.cx, cam.cy, cam.fx, cam.fy);
    HKF_LAUNCH_CHECK(c);
    return HKF_OK;
}

static void current_set(hkf_ctx* c, HkfFrameSet* F)
{
    F->depth_mm = c->depth_mm; F->raw_depth = c->raw_depth; F->trunc_depth = c->trunc_depth; F->filtered_depth = c->filtered_depth;
    for (int l = 0; l < HKF_MAX_LEVELS; ++l) { F->new_v[l] = c->new_v[l]; F->new_n[l] = c->new_n[l]; }
    F->bounds = c->bounds;
}
static void make_current(hkf_ctx* c, const HkfFrameSet& F)
{
    c->depth_mm = F.depth_mm; c->raw_depth = F.raw_depth; c->trunc_depth = F.trunc_depth; c->filtered_depth = F.filtered_depth;
    for (int l = 0; l < HKF_MAX_LEVELS; ++l) { c->new_v[l] = F.new_v[l]; c->new_n[l] = F.new_n[l]; }
    c->bounds = F.bounds;
}

extern "C" int hkf_preprocess(hkf_ctx* c, const uint16_t* depth_mm, int is_device, const HkfFrameParams* fp)
{
    if (!c || !depth_mm || !fp) return HKF_ERR_ARG;
    HkfFrameSet F; current_set(c, &F);
    return run_preprocess(c, c->stream, F, depth_mm, is_device, fp);
}

// ---- frame pipelining: the NEXT frame is uploaded and preprocessed on a second stream into a second set of
// per-frame buffers while the current frame is still being tracked / integrated / raycast; hkf_swap_frame runs its
// preprocess kernels on the main stream and makes it current.
static int alloc_back(hkf_ctx* c)
{
    if (c->back_allocated) return HKF_OK;
    HkfFrameSet& B = c->back;
    size_t n0 = (size_t)c->cols[0] * c->rows[0];
    HKF_CUDA(c, cudaMalloc(&B.depth_mm, n0 * sizeof(uint16_t)));
    HKF_CUDA(c, cudaMalloc(&B.raw_depth, n0 * sizeof(float)));
    HKF_CUDA(c, cudaMalloc(&B.trunc_depth, n0 * sizeof(float)));
    HKF_CUDA(c, cudaMalloc(&B.filtered_depth, n0 * sizeof(float)));
    for (int l = 0; l < c->levels; ++l) {
        size_t nb = (size_t)c->cols[l] * c->rows[l] * sizeof(float4);
        HKF_CUDA(c, cudaMalloc(&B.new_v[l], nb)); HKF_CUDA(c, cudaMalloc(&B.new_n[l], nb));
    }
    B.bounds = c->bounds;
    HKF_CUDA(c, cudaMalloc(&B.bounds.base, c->bounds_zero_bytes));
    HKF_CUDA(c, cudaMemset(B.bounds.base, 0, c->bounds_zero_bytes));
    HKF_CUDA(c, cudaMalloc(&B.bounds.tab, c->bounds_tab_bytes));
    HKF_CUDA(c, cudaMemset(B.bounds.tab, 0, c->bounds_tab_bytes));
    {
        int lo = 0, hi = 0;
        HKF_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        HKF_CUDA(c, cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, hi));
    }
    HKF_CUDA(c, cudaEventCreateWithFlags(&c->ev_pre_done, cudaEventDisableTiming));
    HKF_CUDA(c, cudaEventCreateWithFlags(&c->ev_swap, cudaEventDisableTiming));
    HKF_CUDA(c, cudaEventRecord(c->ev_swap, c->stream));
    // Only the upload runs on the side stream.  Running the next frame's kernels there as well was measured slower on
    // B200 (round 1: 800-1360 instead of 2040 frames/s): every kernel of the chain fills the register file, so the
    // bilateral blocks only delay the tracking blocks.
    c->back_allocated = true;
    return HKF_OK;
}

extern "C" int hkf_preprocess_next(hkf_ctx* c, const uint16_t* depth_mm, int is_device, const HkfFrameParams* fp)
{
    if (!c || !depth_mm || !fp) return HKF_ERR_ARG;
    int rc = alloc_back(c);
    if (rc != HKF_OK) return rc;
    // the back set was the current one until the last swap: everything enqueued before that swap may still read it
    HKF_CUDA(c, cudaStreamWaitEvent(c->stream2, c->ev_swap, 0));
    c->back_src = depth_mm;
    if (!is_device) {
        // copy engine only: the upload overlaps the current frame, the kernels run on the main stream at the swap
        unsigned n = c->cols[0] * c->rows[0];
        HKF_CUDA(c, cudaMemcpyAsync(c->back.depth_mm, depth_mm, n * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream2));
        c->back_src = c->back.depth_mm;
    }
    c->back_fp = *fp;
    HKF_CUDA(c, cudaEventRecord(c->ev_pre_done, c->stream2));
    c->back_valid = true;
    return HKF_OK;
}

extern "C" int hkf_swap_frame(hkf_ctx* c)
{
    if (!c) return HKF_ERR_ARG;
    if (!c->back_valid) { c->err = "hkf_swap_frame: no frame was prepared with hkf_preprocess_next"; return HKF_ERR_STATE; }
    HKF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_pre_done, 0));
    HkfFrameSet cur; current_set(c, &cur);
    make_current(c, c->back);
    c->back = cur;
    c->back_valid = false;
    {
        HkfFrameSet F; current_set(c, &F);
        int rc = run_preprocess(c, c->stream, F, c->back_src, 1, &c->back_fp);
        if (rc != HKF_OK) return rc;
    }
    HKF_CUDA(c, cudaEventRecord(c->ev_swap, c->stream));
    return HKF_OK;
}
#endif // HKF_EMULATE_DEVICE
