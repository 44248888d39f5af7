// frame_emu.cpp — block-level host emulation (emu_block.h) of the preprocess and integration kernels
// (k_preprocess_tile + k_maps_pyramids, k_integrate_items + k_integrate) at sizes and views the GPU tests do not
// use.  Test infrastructure; the product library never links this.
#include "emu_shim.h"
#include "emu_block.h"
#include "../../hybkinectfu_b200/csrc/hkf_internal.cuh"
#include "emu_indices.h"
#include "../../hybkinectfu_b200/csrc/preprocess.cu"
#include "../../hybkinectfu_b200/csrc/integrate.cu"

namespace {
struct Bounds { DepthBounds B; std::vector<unsigned> store; std::vector<uint2> tab; size_t zero_bytes; };
// the allocation of hkf_create (context.cu): the frame maximum + the range-query table
Bounds make_bounds(unsigned cols, unsigned rows)
{
    Bounds b;
    DepthBounds& B = b.B;
    B.tx = (int)div_up(cols, HKF_BT); B.ty = (int)div_up(rows, HKF_BT);
    b.zero_bytes = 4 * sizeof(unsigned);
    b.store.assign(4, 0u);
    b.tab.assign((size_t)HKF_BKX * HKF_BKY * B.tx * B.ty, make_uint2(0u, 0u));
    B.base = b.store.data();
    B.tab = b.tab.data();
    return b;
}
}

struct EmuFrame {
    unsigned cols, rows;
    Bounds bounds;
    std::vector<float> raw, trunc, filt;
    std::vector<float4> v[3], n[3];
    unsigned lc[3], lr[3];
};

extern "C" void* emu_frame_create(unsigned cols, unsigned rows)
{
    EmuFrame* f = new EmuFrame();
    f->cols = cols; f->rows = rows;
    f->bounds = make_bounds(cols, rows);
    f->bounds.B.base = f->bounds.store.data();
    f->bounds.B.tab = f->bounds.tab.data();
    const size_t n0 = (size_t)cols * rows;
    f->raw.assign(n0, 0.f); f->trunc.assign(n0, 0.f); f->filt.assign(n0, 0.f);
    unsigned c = cols, r = rows;
    for (int l = 0; l < 3; ++l) { f->lc[l] = c; f->lr[l] = r; f->v[l].assign((size_t)c * r, make_float4(0, 0, 0, 0)); f->n[l] = f->v[l]; c >>= 1; r >>= 1; }
    return f;
}
extern "C" void emu_frame_destroy(void* h) { delete static_cast<EmuFrame*>(h); }

extern "C" int emu_preprocess(void* h, const uint16_t* mm, float tmin, float tmax, float sigma_pixel, float sigma_depth,
                              float cx, float cy, float fx, float fy)
{
    EmuFrame* f = static_cast<EmuFrame*>(h);
    int radius; float ss, sd, thresh;
    bilateral_consts(sigma_pixel, sigma_depth, &radius, &ss, &sd, &thresh);
    memset(f->bounds.store.data(), 0, f->bounds.zero_bytes);
    const int cols = (int)f->cols, rows = (int)f->rows;
    const dim3 b(TILE_W, TILE_H), g(div_up(f->cols, TILE_W), div_up(f->rows, TILE_H));
    const DepthBounds B = f->bounds.B;
    emu_launch(g, b, [&]() {
        k_preprocess_tile(mm, f->raw.data(), f->trunc.data(), f->filt.data(), cols, rows, radius, tmin, tmax, ss, sd, thresh, B);
    });
    emu_launch(dim3(B.ty, HKF_BKY), dim3(128), [&]() { k_bounds_rmq(B); });
    emu_launch(dim3(div_up(f->cols, 16), div_up(f->rows, 16)), dim3(256), [&]() {
        k_maps_pyramids(f->filt.data(), f->v[0].data(), f->n[0].data(), f->v[1].data(), f->n[1].data(), f->v[2].data(), f->n[2].data(),
                        cols, rows, f->lc[1], f->lr[1], f->lc[2], f->lr[2], cx, cy, fx, fy);
    });
    return 0;
}

// buffer: 0 raw, 1 trunc, 2 filt (floats); 3 + 2*level vertices, 4 + 2*level normals (float4)
extern "C" const void* emu_frame_buffer(void* h, int which)
{
    EmuFrame* f = static_cast<EmuFrame*>(h);
    if (which == 0) return f->raw.data();
    if (which == 1) return f->trunc.data();
    if (which == 2) return f->filt.data();
    const int l = (which - 3) / 2;
    return ((which - 3) & 1) ? (const void*)f->n[l].data() : (const void*)f->v[l].data();
}

// One integration of the frame's truncated depth (k_integrate_items + k_integrate) into vol [R][R][R][2].  Returns the
// number of work items written as free | general << 32.
extern "C" unsigned long long emu_integrate(void* h, float* vol, int R, float S, float wmax, const float* pose_inv16, float trunc, float maxdist,
                                            float cx, float cy, float fx, float fy)
{
    EmuFrame* f = static_cast<EmuFrame*>(h);
    IntegrateArgs A;
    for (int i = 0; i < 12; ++i) A.inv[i] = pose_inv16[i];
    A.cell = S / (float)R;
    A.fx = fx; A.fy = fy; A.cx = cx; A.cy = cy;
    A.cols = (int)f->cols; A.rows = (int)f->rows;
    A.trunc = trunc; A.maxdist = maxdist; A.wmax = wmax;
    A.R = R; A.z_lo = 0; A.z_hi = R; A.size = S;
    A.B = f->bounds.B;
    HkfDeviceState st;
    memset(&st, 0, sizeof(st));
    const unsigned capacity = 4u * (unsigned)R * (unsigned)(R / 4) * div_up((unsigned)R, 128);
    std::vector<uint4> items(capacity);
    const int ngroups = R * (R / INT_G);
    const dim3 g1(div_up((unsigned)ngroups, INT_GB)), b1(256);
    emu_launch(g1, b1, [&]() { k_integrate_items<false, false, false>(A, &st, items.data(), capacity, 0, nullptr); });
    const unsigned long long counts = (unsigned long long)st.int_items[0] | ((unsigned long long)st.int_items[1] << 32);
    emu_launch(dim3(2), dim3(INT_THREADS), [&]() {          // persistent warps: any grid size walks the whole list
        k_integrate<false, false, false>(A, f->trunc.data(), reinterpret_cast<float2*>(vol), &st, items.data(), capacity, 0, IntegrateColor{});
    });
    return counts;
}

extern "C" int emu_can_spawn(int n) { return emu_probe_threads(n) ? 1 : 0; }
