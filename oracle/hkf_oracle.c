/* TEST INFRASTRUCTURE ONLY — see hkf_oracle.h.
 *
 * CPU restatement of the reference's per-frame hot path.  Every float expression follows the
 * association and FMA contraction that nvcc 12.9 (default flags) produced for the reference kernels
 * built for sm_100 (read off their PTX/SASS): fmaf() where the device code has an FFMA, plain
 * operators (compiled with -ffp-contract=off) where it has separate FMUL/FADD, IEEE division and
 * square root throughout, double arithmetic where the source's literals promote to double.
 */
#define _GNU_SOURCE
#include "hkf_oracle.h"
#include "host_math.h"
#include "mc_tables_packed.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ small contracts */
/* x*x' + y*y' + z*z'  ->  fma(z,z', fma(x,x', y*y')) */
static inline float dot3c(float ax, float ay, float az, float bx, float by, float bz)
{
    return fmaf(az, bz, fmaf(ax, bx, ay * by));
}
/* a*b - c*d  ->  fma(a, b, -(c*d)) */
static inline float det2c(float a, float b, float c, float d) { return fmaf(a, b, -(c * d)); }
/* normalize (src/cuda/cuda_declar.h:89-94): (double)len < 1e-8 -> 0 ; v * rcp.rn(len) */
static inline void normalize3c(float x, float y, float z, float out[3])
{
    float len = sqrtf(dot3c(x, y, z, x, y, z));
    if ((double)len < 1e-8) { out[0] = out[1] = out[2] = 0.f; return; }
    float r = 1.0f / len;
    out[0] = x * r; out[1] = y * r; out[2] = z * r;
}
/* (int)(p + 0.5) with a double add, truncation toward zero (src/cuda/DepthCamera.h:39-43).
 * CUDA's cvt.rzi saturates; mimic that so that out-of-range values stay out of range. */
static inline int round_px(float p)
{
    double d = (double)p + 0.5;
    if (!(d == d)) return 0;
    if (d >= 2147483647.0) return 2147483647;
    if (d <= -2147483648.0) return (-2147483647 - 1);
    return (int)d;
}
static inline int f2i_rz(float f)
{
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
/* cvt.rzi.u32.f32 (F2I.U32.TRUNC): NaN -> 0, negative -> 0, saturating */
static inline unsigned f2u_rz(float f)
{
    if (f != f || f <= 0.f) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (unsigned)f;
}
static inline float project1(float v, float f, float z, float c) { return (v * f) / z + c; }
/* Mat44 * float4 row, two contractions (see CalPointToPlaneErrSolverParams PTX) */
static inline float row4_lead_first(const float* m, float x, float y, float z, float w)
{
    return fmaf(w, m[3], fmaf(z, m[2], fmaf(y, m[1], x * m[0])));
}
static inline float row4_lead_second(const float* m, float x, float y, float z, float w)
{
    return fmaf(w, m[3], fmaf(z, m[2], fmaf(x, m[0], y * m[1])));
}

/* ------------------------------------------------------------------ preprocess */
void orc_decode_mm(const uint16_t* mm, float* raw, size_t n)
{
    for (size_t i = 0; i < n; ++i) raw[i] = (float)(mm[i] * 0.001);
}
void orc_trunc(const float* raw, float* out, size_t n, float tmin, float tmax)
{
    for (size_t i = 0; i < n; ++i) { float d = raw[i]; out[i] = (d < tmax && d > tmin) ? d : 0.f; }
}

void orc_bilateral(const float* in, float* out, int cols, int rows, float sigma_pixel, float sigma_depth)
{
    float sd = (float)(0.5 / (double)(sigma_depth * sigma_depth));
    float ss = (float)(0.5 / (double)(sigma_pixel * sigma_pixel));
    int radius = (int)ceil(2.0 * (double)sigma_pixel);
    float thresh = sigma_depth * 5.0f;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float value = in[(size_t)y * cols + x];
            float res = value;
            if (value != 0.f) {
                int x0 = x - radius < 0 ? 0 : x - radius, x1 = x + radius > cols - 1 ? cols - 1 : x + radius;
                int y0 = y - radius < 0 ? 0 : y - radius, y1 = y + radius > rows - 1 ? rows - 1 : y + radius;
                float sum1 = 0.f, sum2 = 0.f;
                int aborted = 0;
                for (int cy = y0; cy <= y1 && !aborted; ++cy)
                    for (int cx = x0; cx <= x1; ++cx) {
                        float tmp = in[(size_t)cy * cols + cx];
                        if (tmp == 0.f) continue;
                        if (fabsf(tmp - value) > thresh) { aborted = 1; break; }
                        float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
                        float d = value - tmp;
                        float data2 = d * d;
                        float arg = fmaf(space2, ss, data2 * sd);
                        /* __expf(-arg) = ex2.approx(-arg * log2(e)); CPU: exp2f of the same product */
                        float w = exp2f(arg * -1.4426950216293334961f);
                        sum1 = fmaf(tmp, w, sum1);
                        sum2 = sum2 + w;
                    }
                if (!aborted && sum2 > 0.f) res = sum1 / sum2;
            }
            out[(size_t)y * cols + x] = res;
        }
}

void orc_vertices(const float* depth, float* v4, const OrcCamera* cam)
{
    int cols = (int)cam->cols, rows = (int)cam->rows;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float d = depth[(size_t)y * cols + x];
            float* o = v4 + 4 * ((size_t)y * cols + x);
            if (d == 0.f) { o[0] = o[1] = o[2] = o[3] = 0.f; continue; }
            o[0] = (d * ((float)(unsigned)x - cam->cx)) / cam->fx;
            o[1] = (d * ((float)(unsigned)y - cam->cy)) / cam->fy;
            o[2] = d; o[3] = 1.0f;
        }
}

void orc_normals(const float* v4, float* n4, int cols, int rows)
{
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float* o = n4 + 4 * ((size_t)y * cols + x);
            o[0] = o[1] = o[2] = o[3] = 0.f;
            if (x == cols - 1 || y == rows - 1 || x == 0 || y == 0) continue;
            const float* v0 = v4 + 4 * ((size_t)y * cols + x);
            const float* vr = v0 + 4; const float* vl = v0 - 4;
            const float* vu = v0 + 4 * (size_t)cols; const float* vd = v0 - 4 * (size_t)cols;
            if (v0[2] == 0.f || vr[2] == 0.f || vu[2] == 0.f || vl[2] == 0.f || vd[2] == 0.f) continue;
            float ax = vu[0] - vd[0], ay = vu[1] - vd[1], az = vu[2] - vd[2];
            float bx = vr[0] - vl[0], by = vr[1] - vl[1], bz = vr[2] - vl[2];
            float c[3];
            normalize3c(det2c(ay, bz, az, by), det2c(az, bx, ax, bz), det2c(ax, by, ay, bx), c);
            o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
        }
}

static inline int is_zero4(const float* p) { return p[0] == 0.f && p[1] == 0.f && p[2] == 0.f && p[3] == 0.f; }

void orc_down_vertices(const float* in, float* out, int in_cols, int in_rows)
{
    int oc = in_cols >> 1, orow = in_rows >> 1;
    for (int y = 0; y < orow; ++y)
        for (int x = 0; x < oc; ++x) {
            const float* p00 = in + 4 * ((size_t)(2 * y) * in_cols + 2 * x);
            const float* p01 = p00 + 4; const float* p10 = p00 + 4 * (size_t)in_cols; const float* p11 = p10 + 4;
            float* o = out + 4 * ((size_t)y * oc + x);
            if (p00[2] == 0.f || p01[2] == 0.f || p10[2] == 0.f || p11[2] == 0.f) { o[0] = o[1] = o[2] = o[3] = 0.f; continue; }
            for (int k = 0; k < 4; ++k) o[k] = (((p00[k] + p01[k]) + p10[k]) + p11[k]) * 0.25f;
        }
}
void orc_down_normals(const float* in, float* out, int in_cols, int in_rows)
{
    int oc = in_cols >> 1, orow = in_rows >> 1;
    for (int y = 0; y < orow; ++y)
        for (int x = 0; x < oc; ++x) {
            const float* p00 = in + 4 * ((size_t)(2 * y) * in_cols + 2 * x);
            const float* p01 = p00 + 4; const float* p10 = p00 + 4 * (size_t)in_cols; const float* p11 = p10 + 4;
            float* o = out + 4 * ((size_t)y * oc + x);
            o[0] = o[1] = o[2] = o[3] = 0.f;
            if (is_zero4(p01) || is_zero4(p10) || is_zero4(p00) || is_zero4(p11)) continue;
            float s[3];
            for (int k = 0; k < 3; ++k) s[k] = (((p00[k] + p01[k]) + p10[k]) + p11[k]) * 0.25f;
            float n[3];
            normalize3c(s[0], s[1], s[2], n);
            o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
        }
}

/* ------------------------------------------------------------------ integration */
/* has_color (color != NULL): src/cuda/integrateVolume.cu:52-72 + tsdfVolume.h:57-74 as compiled for sm_100 (SASS of
 * oracle/_ref/integrateVolume.o): colour pixel (int)(X*525/Z + 320), (int)(Y*525/Z + 240) in float with a truncating
 * convert, voxel skipped when it falls outside [1, rgb_cols-2] x [1, rgb_rows-2]; weight 2 or 2*fminf(|n.z|/0.75, 1);
 * c' = trunc(fminf(fma(wc, c_pixel, w_old * c_old) / (wc + w_old), 255)). */
uint64_t orc_integrate_color(float* vol, uint8_t* color, int R, float S, float wmax, const float* depth, const float* normals4,
                             const uint8_t* rgb, int rgb_cols, int rgb_rows, int angled, const OrcCamera* cam,
                             const float pose[16], float sdf_trunc, float max_dist, int z0, int z1)
{
    float inv[16];
    hm_mat44_inverse(pose, inv);                 /* cudaIntegrateVolume: transform.getInverse() on the host */
    const float cell = S / (float)R;             /* _size.x / _resolution.x */
    const int cols = (int)cam->cols, rows = (int)cam->rows;
    uint64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int z = z0; z < z1; ++z) {
        const float wz = ((float)z + 0.5f) * cell;
        for (int y = 0; y < R; ++y) {
            const float wy = ((float)y + 0.5f) * cell;
            const float ty1 = wy * inv[1], ty2 = wy * inv[5], ty3 = wy * inv[9];
            float* rowp = vol + 2 * (((size_t)z * R + y) * R);
            uint8_t* crow = color ? color + 4 * (((size_t)z * R + y) * R) : NULL;
            for (int x = 0; x < R; ++x) {
                const float wx = ((float)x + 0.5f) * cell;
                const float Z = fmaf(wz, inv[10], fmaf(wx, inv[8], ty3)) + inv[11];
                if (Z <= 0.f) continue;
                const float X = fmaf(wz, inv[2], fmaf(wx, inv[0], ty1)) + inv[3];
                const float Y = fmaf(wz, inv[6], fmaf(wx, inv[4], ty2)) + inv[7];
                const int u = round_px(project1(X, cam->fx, Z, cam->cx));
                const int v = round_px(project1(Y, cam->fy, Z, cam->cy));
                if (u >= cols - 1 || v >= rows - 1 || u < 1 || v < 1) continue;
                const float d = depth[(size_t)v * cols + u];
                if (d == 0.f) continue;
                int cu = 0, cv = 0;
                if (color) {
                    cu = f2i_rz((X * 525.0f) / Z + 320.0f);
                    cv = f2i_rz((Y * 525.0f) / Z + 240.0f);
                    if (cu >= rgb_cols - 1 || cv >= rgb_rows - 1 || cu < 1 || cv < 1) continue;
                }
                if (!(d < max_dist)) continue;
                const float sdf = d - Z;
                if (!(sdf > -sdf_trunc)) continue;
                const float tsdf = fminf(sdf / sdf_trunc, 1.0f);
                float t = rowp[2 * x], w = rowp[2 * x + 1];
                float w1 = w + 1.0f;
                rowp[2 * x] = fmaf(t, w, tsdf) / w1;
                rowp[2 * x + 1] = fminf(w1, wmax);
                if (color) {
                    float wc = 2.0f;
                    if (angled) { const float q = fminf(fabsf(normals4[4 * ((size_t)v * cols + u) + 2]) / 0.75f, 1.0f); wc = q + q; }
                    const uint8_t* px = rgb + 3 * ((size_t)cv * rgb_cols + cu);
                    const float den = wc + w;
                    for (int k = 0; k < 3; ++k) {
                        const float nc = fminf(fmaf(wc, (float)px[k], w * (float)crow[4 * x + k]) / den, 255.0f);
                        crow[4 * x + k] = (uint8_t)f2u_rz(nc);
                    }
                    crow[4 * x + 3] = 0;
                }
                total++;
            }
        }
    }
    return total;
}

uint64_t orc_integrate(float* vol, int R, float S, float wmax, const float* depth, const OrcCamera* cam,
                       const float pose[16], float sdf_trunc, float max_dist, int z0, int z1)
{
    return orc_integrate_color(vol, NULL, R, S, wmax, depth, NULL, NULL, 0, 0, 0, cam, pose, sdf_trunc, max_dist, z0, z1);
}

/* ------------------------------------------------------------------ ICP sums */
/* reduce<T> of src/cuda/device_functions.h:16-35 for n = 256 values */
static float tree256(float* b)
{
    for (int i = 0; i < 128; ++i) b[i] = b[i] + b[i + 128];
    for (int i = 0; i < 64; ++i) b[i] = b[i] + b[i + 64];
    for (int s = 32; s >= 1; s >>= 1)
        for (int i = 0; i < s; ++i) b[i] = b[i] + b[i + s];
    return b[0];
}
/* reduceGbufKernel with 512 threads (device_functions.h:38-55) */
static float reduce_row512(const float* row, unsigned n)
{
    float b[512];
    for (unsigned t = 0; t < 512; ++t) {
        float sum = 0.f;
        for (unsigned i = t; i < n; i += 512) sum += row[i];
        b[t] = sum;
    }
    for (int i = 0; i < 256; ++i) b[i] = b[i] + b[i + 256];
    for (int i = 0; i < 128; ++i) b[i] = b[i] + b[i + 128];
    for (int i = 0; i < 64; ++i) b[i] = b[i] + b[i + 64];
    for (int s = 32; s >= 1; s >>= 1)
        for (int i = 0; i < s; ++i) b[i] = b[i] + b[i + s];
    return b[0];
}

static void reduce_blocks(const float* rows7, int cols, int rows, float out27[27])
{
    /* rows7: per pixel 7 floats (zero when no correspondence); 16x16 blocks */
    int gx = (cols + 15) / 16, gy = (rows + 15) / 16;
    unsigned nblocks = (unsigned)(gx * gy);
    float* gbuf = (float*)malloc(sizeof(float) * 27 * (size_t)nblocks);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < (int)nblocks; ++b) {
        int bx = b % gx, by = b / gx;
        float r7[256][7];
        for (int ty = 0; ty < 16; ++ty)
            for (int tx = 0; tx < 16; ++tx) {
                int x = bx * 16 + tx, y = by * 16 + ty;
                float* d = r7[ty * 16 + tx];
                if (x < cols && y < rows) memcpy(d, rows7 + 7 * ((size_t)y * cols + x), 7 * sizeof(float));
                else memset(d, 0, 7 * sizeof(float));
            }
        int shift = 0;
        float buf[256];
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 7; ++j) {
                for (int t = 0; t < 256; ++t) buf[t] = r7[t][i] * r7[t][j];
                gbuf[(size_t)shift * nblocks + b] = tree256(buf);
                shift++;
            }
    }
    for (int k = 0; k < 27; ++k) out27[k] = reduce_row512(gbuf + (size_t)k * nblocks, nblocks);
    free(gbuf);
}

void orc_icp_sums(const float* new_v, const float* new_n, const float* model_v, const float* model_n,
                  const OrcCamera* cam, const float cur[16], const float li[16],
                  float dist_thres, float norm_sin_thres, float out27[27])
{
    int cols = (int)cam->cols, rows = (int)cam->rows;
    float* rows7 = (float*)calloc((size_t)cols * rows * 7, sizeof(float));
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const float* v = new_v + 4 * ((size_t)y * cols + x);
            const float* n = new_n + 4 * ((size_t)y * cols + x);
            if (is_zero4(n)) continue;
            /* v_input_g = cur * v  (lead-first contraction); n_input_g = cur * n (lead-second) */
            float qx = row4_lead_first(cur + 0, v[0], v[1], v[2], v[3]);
            float qy = row4_lead_first(cur + 4, v[0], v[1], v[2], v[3]);
            float qz = row4_lead_first(cur + 8, v[0], v[1], v[2], v[3]);
            float qw = row4_lead_first(cur + 12, v[0], v[1], v[2], v[3]);
            float nix = row4_lead_second(cur + 0, n[0], n[1], n[2], n[3]);
            float niy = row4_lead_second(cur + 4, n[0], n[1], n[2], n[3]);
            float niz = row4_lead_second(cur + 8, n[0], n[1], n[2], n[3]);
            /* v_cp = last_inv * v_input_g (lead-second) */
            float cxp = row4_lead_second(li + 0, qx, qy, qz, qw);
            float cyp = row4_lead_second(li + 4, qx, qy, qz, qw);
            float czp = row4_lead_second(li + 8, qx, qy, qz, qw);
            int u = round_px(project1(cxp, cam->fx, czp, cam->cx));
            int w = round_px(project1(cyp, cam->fy, czp, cam->cy));
            if (u < 0 || u >= cols || w < 0 || w >= rows) continue;
            const float* nt = model_n + 4 * ((size_t)w * cols + u);
            if (is_zero4(nt)) continue;
            const float* p = model_v + 4 * ((size_t)w * cols + u);
            float dx = p[0] - qx, dy = p[1] - qy, dz = p[2] - qz;
            float d = sqrtf(dot3c(dx, dy, dz, dx, dy, dz));
            float sx = det2c(nt[1], niz, nt[2], niy);
            float sy = det2c(nt[2], nix, nt[0], niz);
            float sz = det2c(nt[0], niy, nt[1], nix);
            float s = sqrtf(dot3c(sx, sy, sz, sx, sy, sz));
            if (d > dist_thres || s > norm_sin_thres) continue;
            float* r = rows7 + 7 * ((size_t)y * cols + x);
            r[0] = det2c(qy, nt[2], qz, nt[1]);
            r[1] = det2c(qz, nt[0], qx, nt[2]);
            r[2] = det2c(qx, nt[1], qy, nt[0]);
            r[3] = nt[0]; r[4] = nt[1]; r[5] = nt[2];
            r[6] = dot3c(nt[0], nt[1], nt[2], dx, dy, dz);
        }
    reduce_blocks(rows7, cols, rows, out27);
    free(rows7);
}

/* ------------------------------------------------------------------ TSDF sampling */
typedef struct { const float* vol; int R; float S; float cell; float Rf; } VolView;
static inline VolView volview(const float* vol, int R, float S)
{
    VolView v; v.vol = vol; v.R = R; v.S = S; v.cell = S / (float)R; v.Rf = (float)R; return v;
}
/* worldPosToVoxel: (int)(p * R / S) */
static inline int w2v(const VolView* V, float p) { return f2i_rz((p * V->Rf) / V->S); }
static inline const float* voxel(const VolView* V, int x, int y, int z) { return V->vol + 2 * (((size_t)z * V->R + y) * V->R + x); }

static int interp_sdf(const VolView* V, float px, float py, float pz, float* dist)
{
    int gx = w2v(V, px), gy = w2v(V, py), gz = w2v(V, pz);
    int R = V->R;
    if (gx <= 0 || gx >= R - 1) return 0;
    if (gy <= 0 || gy >= R - 1) return 0;
    if (gz <= 0 || gz >= R - 1) return 0;
    float cell = V->cell;
    float vx = ((float)gx + 0.5f) * cell, vy = ((float)gy + 0.5f) * cell, vz = ((float)gz + 0.5f) * cell;
    gx = (px < vx) ? gx - 1 : gx;
    gy = (py < vy) ? gy - 1 : gy;
    gz = (pz < vz) ? gz - 1 : gz;
    /* (p - (g+0.5)*cell)/cell, the numerator fused by ptxas: fma(g+0.5, -cell, p) */
    float a = fmaf((float)gx + 0.5f, -cell, px) / cell;
    float b = fmaf((float)gy + 0.5f, -cell, py) / cell;
    float c = fmaf((float)gz + 0.5f, -cell, pz) / cell;
    const float* v000 = voxel(V, gx, gy, gz);         if (v000[1] == 0.f) return 0;
    const float* v001 = voxel(V, gx, gy, gz + 1);     if (v001[1] == 0.f) return 0;
    const float* v010 = voxel(V, gx, gy + 1, gz);     if (v010[1] == 0.f) return 0;
    const float* v011 = voxel(V, gx, gy + 1, gz + 1); if (v011[1] == 0.f) return 0;
    const float* v100 = voxel(V, gx + 1, gy, gz);     if (v100[1] == 0.f) return 0;
    const float* v101 = voxel(V, gx + 1, gy, gz + 1); if (v101[1] == 0.f) return 0;
    const float* v110 = voxel(V, gx + 1, gy + 1, gz); if (v110[1] == 0.f) return 0;
    const float* v111 = voxel(V, gx + 1, gy + 1, gz + 1); if (v111[1] == 0.f) return 0;
    float ia = 1.0f - a, ib = 1.0f - b, ic = 1.0f - c;
    float acc = fmaf(ic, ib * (ia * v000[0]), c * (ib * (ia * v001[0])));
    acc = fmaf(ic, b * (ia * v010[0]), acc);
    acc = fmaf(c, b * (ia * v011[0]), acc);
    acc = fmaf(ic, ib * (a * v100[0]), acc);
    acc = fmaf(c, ib * (a * v101[0]), acc);
    acc = fmaf(ic, b * (a * v110[0]), acc);
    acc = fmaf(c, b * (a * v111[0]), acc);
    *dist = acc;
    return 1;
}
/* tsdfvolume::interpolateColor (src/cuda/tsdfVolume.h:122-148): same cell and term order, per colour byte, truncated */
static int interp_color(const VolView* V, const uint8_t* color, float px, float py, float pz, uint8_t out[3])
{
    int gx = w2v(V, px), gy = w2v(V, py), gz = w2v(V, pz);
    int R = V->R;
    if (gx <= 0 || gx >= R - 1) return 0;
    if (gy <= 0 || gy >= R - 1) return 0;
    if (gz <= 0 || gz >= R - 1) return 0;
    float cell = V->cell;
    float vx = ((float)gx + 0.5f) * cell, vy = ((float)gy + 0.5f) * cell, vz = ((float)gz + 0.5f) * cell;
    gx = (px < vx) ? gx - 1 : gx;
    gy = (py < vy) ? gy - 1 : gy;
    gz = (pz < vz) ? gz - 1 : gz;
    float a = fmaf((float)gx + 0.5f, -cell, px) / cell;
    float b = fmaf((float)gy + 0.5f, -cell, py) / cell;
    float c = fmaf((float)gz + 0.5f, -cell, pz) / cell;
    static const int off[8][3] = { {0,0,0},{0,0,1},{0,1,0},{0,1,1},{1,0,0},{1,0,1},{1,1,0},{1,1,1} };
    const uint8_t* cp[8];
    for (int k = 0; k < 8; ++k) {
        if (voxel(V, gx + off[k][0], gy + off[k][1], gz + off[k][2])[1] == 0.f) return 0;
        cp[k] = color + 4 * (((size_t)(gz + off[k][2]) * R + (gy + off[k][1])) * R + (gx + off[k][0]));
    }
    float ia = 1.0f - a, ib = 1.0f - b, ic = 1.0f - c;
    for (int ch = 0; ch < 3; ++ch) {
        float acc = fmaf(ic, ib * (ia * (float)cp[0][ch]), c * (ib * (ia * (float)cp[1][ch])));
        acc = fmaf(ic, b * (ia * (float)cp[2][ch]), acc);
        acc = fmaf(c, b * (ia * (float)cp[3][ch]), acc);
        acc = fmaf(ic, ib * (a * (float)cp[4][ch]), acc);
        acc = fmaf(c, ib * (a * (float)cp[5][ch]), acc);
        acc = fmaf(ic, b * (a * (float)cp[6][ch]), acc);
        acc = fmaf(c, b * (a * (float)cp[7][ch]), acc);
        out[ch] = (uint8_t)f2u_rz(acc);
    }
    return 1;
}
int orc_interpolate_sdf(const float* vol, int R, float S, float px, float py, float pz, float* dist)
{
    VolView V = volview(vol, R, S);
    return interp_sdf(&V, px, py, pz, dist);
}

/* ------------------------------------------------------------------ raycast */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct { float ox, oy, oz, dx, dy, dz, tmin, tmax; int live; } Ray;

/* ray set-up of raycastKernel (src/cuda/raycastingVolume.cu:137-153) */
static Ray ray_setup(const OrcCamera* cam, const float pose[16], float S, float near_plane, float far_plane, int x, int y)
{
    Ray r;
    r.ox = pose[3]; r.oy = pose[7]; r.oz = pose[11];
    /* cam_dir = normalize(depthToSkeleton(x,y,1)) ; z*z == 1 folds into "+ 1" */
    float vx = ((float)(unsigned)x - cam->cx) / cam->fx;
    float vy = ((float)(unsigned)y - cam->cy) / cam->fy;
    float len = sqrtf(fmaf(vx, vx, vy * vy) + 1.0f);
    float cdx, cdy, cdz;
    if ((double)len < 1e-8) { cdx = cdy = cdz = 0.f; }
    else { float q = 1.0f / len; cdx = vx * q; cdy = vy * q; cdz = q; }
    /* world_dir = transform * (cam_dir, 0): lead-second rows, w = 0 */
    float dx = fmaf(pose[3], 0.f, fmaf(cdz, pose[2], fmaf(cdx, pose[0], cdy * pose[1])));
    float dy = fmaf(pose[7], 0.f, fmaf(cdz, pose[6], fmaf(cdx, pose[4], cdy * pose[5])));
    float dz = fmaf(pose[11], 0.f, fmaf(cdz, pose[10], fmaf(cdx, pose[8], cdy * pose[9])));
    if (dx == 0.f) dx = 1e-15f;
    if (dy == 0.f) dy = 1e-15f;
    if (dz == 0.f) dz = 1e-15f;
    float txmin = ((dx > 0 ? 0.f : S) - r.ox) / dx, tymin = ((dy > 0 ? 0.f : S) - r.oy) / dy, tzmin = ((dz > 0 ? 0.f : S) - r.oz) / dz;
    float tmin = fmaxf(fmaxf(txmin, tymin), tzmin);
    float txmax = ((dx > 0 ? S : 0.f) - r.ox) / dx, tymax = ((dy > 0 ? S : 0.f) - r.oy) / dy, tzmax = ((dz > 0 ? S : 0.f) - r.oz) / dz;
    float tmax = fminf(fminf(txmax, tymax), tzmax);
    r.tmin = fmaxf(tmin, near_plane / cdz);
    r.tmax = fminf(tmax, far_plane / cdz);
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.live = r.tmin < r.tmax;
    return r;
}

/* the (+ -> -) pair (last sample at l*, this sample at p*, parameter t): refinement, vertex, normal
   (src/cuda/raycastingVolume.cu:65-119); leaves the pixel zero on any failure */
static void ray_hit(const VolView* V, const Ray* r, float inc, float t, float px, float py, float pz, float lx, float ly, float lz,
                    float* ov, float* on, const uint8_t* color, uint8_t* oc)
{
    const int R = V->R;
    float ftdt, ft;
    if (!interp_sdf(V, px, py, pz, &ftdt)) return;
    if (!interp_sdf(V, lx, ly, lz, &ft)) return;
    float alpha = t - (inc * ftdt) / (ftdt - ft);
    float wx = fmaf(r->dx, alpha, r->ox), wy = fmaf(r->dy, alpha, r->oy), wz = fmaf(r->dz, alpha, r->oz);
    /* has_color (raycastingVolume.cu:101-106): written when the interpolation succeeds, whatever the gradient does next */
    if (color) { uint8_t c3[3]; if (interp_color(V, color, wx, wy, wz, c3)) { oc[0] = c3[0]; oc[1] = c3[1]; oc[2] = c3[2]; } }
    /* gradientForPoint(volume, last_world_pos, vertex) */
    int hx = w2v(V, lx), hy = w2v(V, ly), hz = w2v(V, lz);
    if (hx <= 1 || hx >= R - 2) return;
    if (hy <= 1 || hy >= R - 2) return;
    if (hz <= 1 || hz >= R - 2) return;
    float cell = V->cell;
    float f1, f2, nx, ny, nz;
    if (!interp_sdf(V, wx + cell, wy, wz, &f1)) return;
    if (!interp_sdf(V, wx - cell, wy, wz, &f2)) return;
    nx = f1 - f2;
    if (!interp_sdf(V, wx, wy + cell, wz, &f1)) return;
    if (!interp_sdf(V, wx, wy - cell, wz, &f2)) return;
    ny = f1 - f2;
    if (!interp_sdf(V, wx, wy, wz + cell, &f1)) return;
    if (!interp_sdf(V, wx, wy, wz - cell, &f2)) return;
    nz = f1 - f2;
    float l = sqrtf(dot3c(nx, ny, nz, nx, ny, nz));
    if ((double)l < 1e-8) return;
    float q = 1.0f / l;
    ov[0] = wx; ov[1] = wy; ov[2] = wz; ov[3] = 1.0f;
    on[0] = nx * q; on[1] = ny * q; on[2] = nz * q; on[3] = 0.f;
}

void orc_raycast(const float* vol, int R, float S, const OrcCamera* cam, const float pose[16], float inc,
                 float near_plane, float far_plane, float* out_v4, float* out_n4)
{
    orc_raycast_color(vol, NULL, R, S, cam, pose, inc, near_plane, far_plane, out_v4, out_n4, NULL);
}

void orc_raycast_color(const float* vol, const uint8_t* color, int R, float S, const OrcCamera* cam, const float pose[16], float inc,
                       float near_plane, float far_plane, float* out_v4, float* out_n4, uint8_t* out_rgb3)
{
    VolView V = volview(vol, R, S);
    const int cols = (int)cam->cols, rows = (int)cam->rows;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float* ov = out_v4 + 4 * ((size_t)y * cols + x);
            float* on = out_n4 + 4 * ((size_t)y * cols + x);
            ov[0] = ov[1] = ov[2] = ov[3] = 0.f;
            on[0] = on[1] = on[2] = on[3] = 0.f;
            uint8_t* oc = color ? out_rgb3 + 3 * ((size_t)y * cols + x) : NULL;
            if (oc) { oc[0] = oc[1] = oc[2] = 0; }
            const Ray r = ray_setup(cam, pose, S, near_plane, far_plane, x, y);
            if (!r.live) continue;
            float t = r.tmin, last_sdf = 0.f, lx = 0.f, ly = 0.f, lz = 0.f;
            while (t < r.tmax) {
                float px = fmaf(r.dx, t, r.ox), py = fmaf(r.dy, t, r.oy), pz = fmaf(r.dz, t, r.oz);
                int gx = clampi(w2v(&V, px), 0, R - 1), gy = clampi(w2v(&V, py), 0, R - 1), gz = clampi(w2v(&V, pz), 0, R - 1);
                float sdf = voxel(&V, gx, gy, gz)[0];
                if (last_sdf > 0.0f && sdf < 0.0f) {
                    ray_hit(&V, &r, inc, t, px, py, pz, lx, ly, lz, ov, on, color, oc);
                    break;
                }
                last_sdf = sdf; lx = px; ly = py; lz = pz;
                t = inc + t;
            }
        }
}

/* ------------------------------------------------------------------ host-math exports */
int orc_solve_icp(const float s27[27], float x[6]) { return hm_solve_icp(s27, x); }
void orc_solve_llt(const float s27[27], float x[6]) { hm_solve_llt(s27, x); }
int orc_vector6_to_transform(const float x[6], float a, float d, float T[16]) { return hm_vector6_to_transform(x, a, d, T); }
void orc_sdf_compose(const float cur[16], const float x[6], float next[16]) { hm_sdf_compose(cur, x, next); }
void orc_mat44_inverse(const float in[16], float out[16]) { hm_mat44_inverse(in, out); }
void orc_mat44_mul(const float a[16], const float b[16], float c[16]) { hm_mat44_mul(a, b, c); }
void orc_quat_to_mat3(float w, float x, float y, float z, float R[9]) { hm_quat_to_mat3(w, x, y, z, R); }

#include "hkf_oracle_track.inc"
