// marching_cubes.cu — iso-surface extraction of the TSDF into a triangle soup.
//
// Replaces extractIsoSurfaceKernel / extractIsoSurfaceAtPosition / vertexInterp / appendTriangle and
// cudaMarchingcube (src/cuda/marchingcube.cu:5-164) with the case table of marchingcube_table.h.
//
// Reference shape: a thread per (x,y) column looping over z (uncoalesced across the warp in z, 288 B of
// local stack per thread), one global atomicAdd per triangle, check-then-add overflow race.
//
// This design: a thread per cell with x fastest so the warp's voxel reads are contiguous, a cheap rejection of
// cells whose raw 3x3x3 neighbourhood cannot produce a triangle (free space, interior, unobserved), early exit at
// the first unobserved corner exactly in the reference's corner order, the packed case table in constant
// memory (one 64-bit word per case, the 12-bit edge mask derived from it), and ONE atomicAdd per warp:
// lanes prefix-sum their triangle counts with shuffles and write into a contiguous range.  Triangle
// ORDER is nondeterministic in the reference as well; the triangle multiset is identical.
//
// Semantics kept: the 8 "corner" SDFs are trilinear interpolations at the lattice corners, all 8 voxels
// of each must be observed; cube index bit order 010,110,100,000,011,111,101,001; cases whose edge mask
// is 0 or 0xff are skipped (the author's addition, :110); vertexInterp snaps within 1e-5 and otherwise
// returns fma(p2-p1, (0-d1)/(d2-d1), p1); |d| > threshold rejects the cell (never binds for 300*S/R).
#include "tsdf_sample.cuh"
#include "mc_tables.h"

__constant__ unsigned long long c_tri[256];
static bool g_tables_uploaded[64] = { false };

struct McArgs {
    VolView V;
    int z0, z1;          // cells (planes) to extract
    float threshold;
    HkfTriangle* tris;
    unsigned max_tris;
    const uchar4* color;     // has_color: the voxel colour plane (corner colours = interpolateColor at the lattice corners)
};

__device__ __forceinline__ void mc_vertex(const float* p1, const float* p2, float d1, float d2, float* out)
{
    const float m1 = __fsub_rn(0.0f, d1);
    if (fabsf(m1) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
    if (fabsf(__fsub_rn(0.0f, d2)) < 0.00001f) { out[0] = p2[0]; out[1] = p2[1]; out[2] = p2[2]; return; }
    if (fabsf(__fsub_rn(d1, d2)) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
    const float mu = __fdiv_rn(m1, __fsub_rn(d2, d1));
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = __fmaf_rn(__fsub_rn(p2[k], p1[k]), mu, p1[k]);
}

// the colour half of vertexInterp (marchingcube.cu:5-26 as compiled): c / 255 as a multiplication by (float)(1.0/255) at a
// snapped end point, otherwise fma((float)(c2 - c1), mu, (float)c1) / 255 with an IEEE division
__device__ __forceinline__ void mc_vertex_color(unsigned c1, unsigned c2, float d1, float d2, float* out)
{
    const float m1 = __fsub_rn(0.0f, d1);
    const float k = 0.0039215688593685626984f;
    const bool snap1 = fabsf(m1) < 0.00001f;
    const bool snap2 = !snap1 && fabsf(__fsub_rn(0.0f, d2)) < 0.00001f;
    const bool flat = !snap1 && !snap2 && fabsf(__fsub_rn(d1, d2)) < 0.00001f;
    const float mu = __fdiv_rn(m1, __fsub_rn(d2, d1));
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int a = (int)((c1 >> (8 * ch)) & 0xffu), b = (int)((c2 >> (8 * ch)) & 0xffu);
        if (snap1 || flat) out[ch] = __fmul_rn((float)a, k);
        else if (snap2) out[ch] = __fmul_rn((float)b, k);
        else out[ch] = __fdiv_rn(__fmaf_rn((float)(b - a), mu, (float)a), 255.0f);
    }
}

// Cheap rejection of cell (x, y, z) on its raw 3x3x3 neighbourhood (the union of the eight corners' 2x2x2 voxel sets),
// before any interpolation.  If one of the 27 voxels is unobserved some corner's interpolation fails and the reference
// drops the cell; if all 27 tsdf values have the same strict sign every corner value - a combination of eight of them
// with weights ~1/8 > 0 - has that sign too, the case is 0 or 255 and no triangle comes out.  Returns true when the
// cell still has to be evaluated.  (Border cells and slab edges are always evaluated.)
template <bool SLAB>
__device__ __forceinline__ bool mc_candidate(const VolView& V, int x, int y, int z, float2 ctr)
{
    const int R = V.R;
    if (!(x >= 1 && y >= 1 && z >= 1 && x <= R - 2 && y <= R - 2 && z <= R - 2 && (!SLAB || (z - 1 >= V.kz_lo && z + 1 < V.kz_hi)))) return true;
    if (ctr.y == 0.f) return false;
    const float2* p0 = V.vol + vol_index(V, x, y, z);
    bool any_pos = false, any_neg = false, unobserved = false;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        // SLAB: the plane above / below may be a neighbour's (peer-memory halo)
        const float2* pz = SLAB ? vol_plane(V, z + dz) + ((size_t)y * R + x) : p0 + (ptrdiff_t)dz * R * R;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const float2 q = __ldg(pz + (ptrdiff_t)dy * R + dx);
                unobserved |= (q.y == 0.f);
                any_pos |= (q.x > 0.f);
                any_neg |= (q.x < 0.f);
                if (!(q.x > 0.f) && !(q.x < 0.f)) { any_pos = true; any_neg = true; }   // zero (or NaN): undecided
            }
    }
    return !unobserved && any_pos && any_neg;
}

// Pass A: a lean streaming sweep (4 cells per thread per step, all centre loads in flight) that appends the cells
// that survive mc_candidate to a compact list with one atomic per warp.  ~75 % of a volume is unobserved and costs
// one 8-byte read per cell here; free space costs 27 cached reads; only the thin shell around the surface goes on.
template <bool SLAB>
__global__ void __launch_bounds__(256)
k_mc_candidates(McArgs A, unsigned* __restrict__ cand, unsigned cap, unsigned* __restrict__ ncand)
{
    const VolView& V = A.V;
    const int R = V.R;
    const int lane = threadIdx.x & 31;
    const int nrows = (A.z1 - A.z0) * R;
    const int xsteps = (R + 1023) / 1024;
    for (int job = blockIdx.x; job < nrows * xsteps; job += gridDim.x) {
        const int row = job / xsteps, xs = (job - row * xsteps) * 1024 + (int)threadIdx.x;
        const int zr = row / R, y = row - zr * R, z = zr + A.z0;
        const bool zin = !SLAB || (z >= V.z_lo && z < V.z_hi);
        float2 ctr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = xs + 256 * k;
            ctr[k] = (x < R && zin) ? __ldg(V.vol + vol_index(V, x, y, z)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = xs + 256 * k;
            const bool keep = (x < R) && mc_candidate<SLAB>(V, x, y, z, ctr[k]);
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            if (!b) continue;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(ncand, (unsigned)__popc(b));
            base = __shfl_sync(0xffffffffu, base, 0) + (unsigned)__popc(b & ((1u << lane) - 1));
            if (keep && base < cap) cand[base] = ((unsigned)(z - A.z0) * (unsigned)R + (unsigned)y) * (unsigned)R + (unsigned)x;
        }
    }
}

// Pass B (or the whole extraction when CANDIDATES is false): the reference's cell evaluation.
template <bool SLAB, bool CANDIDATES, bool COLOR = false>
__global__ void __launch_bounds__(256)
k_marching_cubes(McArgs A, unsigned* __restrict__ counter, const unsigned* __restrict__ cand, const unsigned* __restrict__ ncand)
{
    const VolView& V = A.V;
    const int R = V.R;
    const int lane = threadIdx.x & 31;
    // without a candidate list a block walks rows (y, z) of cells, its threads the x positions of the row in steps of 256
    const int nrows = (A.z1 - A.z0) * R;
    const int xsteps = (R + (int)blockDim.x - 1) / (int)blockDim.x;
    const int njobs = CANDIDATES ? (int)((*ncand + blockDim.x - 1) / blockDim.x) : nrows * xsteps;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        int x, y, z;
        bool active;
        if (CANDIDATES) {
            const unsigned i = (unsigned)job * blockDim.x + threadIdx.x;
            active = i < *ncand;
            const unsigned id = active ? cand[i] : 0u;
            x = (int)(id % (unsigned)R); y = (int)((id / (unsigned)R) % (unsigned)R); z = (int)(id / ((unsigned)R * (unsigned)R)) + A.z0;
        } else {
            const int row = job / xsteps;
            x = (job - row * xsteps) * (int)blockDim.x + (int)threadIdx.x;
            const int zr = row / R;
            y = row - zr * R; z = zr + A.z0;
            active = x < R;
        }
        int ntri = 0;
        unsigned long long packed = ~0ull;
        float pos[8][3], dist[8];          // corner c = i + 2j + 4k (i along x, j along y, k along z)
        if (active) {
            const float cell = V.cell;
            const float wx = __fmul_rn(__fadd_rn((float)x, 0.5f), cell);
            const float wy = __fmul_rn(__fadd_rn((float)y, 0.5f), cell);
            const float wz = __fmul_rn(__fadd_rn((float)z, 0.5f), cell);
            const float P = __fmul_rn(cell, 0.5f), M = __fmul_rn(cell, -0.5f);
            // reference evaluation order with early exit: 000,100,010,001,110,011,101,111
            const int order[8] = { 0, 1, 2, 4, 3, 6, 5, 7 };
            bool ok = true;
            if (!CANDIDATES) {
                const bool zin = !SLAB || (z >= V.z_lo && z < V.z_hi);
                ok = mc_candidate<SLAB>(V, x, y, z, zin ? __ldg(V.vol + vol_index(V, x, y, z)) : make_float2(0.f, 0.f));
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int c = order[o];
                if (ok) {
                    pos[c][0] = __fadd_rn(wx, (c & 1) ? P : M);
                    pos[c][1] = __fadd_rn(wy, (c & 2) ? P : M);
                    pos[c][2] = __fadd_rn(wz, (c & 4) ? P : M);
                    ok = (vol_interp_sdf<SLAB>(V, pos[c][0], pos[c][1], pos[c][2], dist[c]) == 1);
                }
            }
            if (ok) {
                unsigned ci = 0;
                if (dist[2] < 0.f) ci += 1;      // 010
                if (dist[3] < 0.f) ci += 2;      // 110
                if (dist[1] < 0.f) ci += 4;      // 100
                if (dist[0] < 0.f) ci += 8;      // 000
                if (dist[6] < 0.f) ci += 16;     // 011
                if (dist[7] < 0.f) ci += 32;     // 111
                if (dist[5] < 0.f) ci += 64;     // 101
                if (dist[4] < 0.f) ci += 128;    // 001
                bool over = false;
#pragma unroll
                for (int c = 0; c < 8; ++c) over |= (fabsf(dist[c]) > A.threshold);
                if (!over) {
                    packed = c_tri[ci];
                    unsigned em = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) { unsigned e = (unsigned)((packed >> (4 * k)) & 0xF); if (e != 0xF) em |= 1u << e; }
                    if (em == 0 || em == 255) packed = ~0ull;
                }
            }
            if (packed != ~0ull) {
#pragma unroll
                for (int k = 0; k < 16; k += 3) if (((packed >> (4 * k)) & 0xF) != 0xF) ++ntri;
            }
        }
        if (!__any_sync(0xffffffffu, ntri > 0)) continue;       // nothing to emit in this warp (the common case)
        // one atomic per warp: exclusive prefix of ntri over the lanes
        int incl = ntri;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned wbase = 0;
        if (total > 0) {
            if (lane == 31) wbase = atomicAdd(counter, (unsigned)total);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
        }
        if (ntri > 0) {
            // edge e joins corners ea[e], eb[e] (corner id = i + 2j + 4k)
            const int ea[12] = { 2, 3, 1, 0, 6, 7, 5, 4, 2, 3, 1, 0 };
            const int eb[12] = { 3, 1, 0, 2, 7, 5, 4, 6, 6, 7, 5, 4 };
            unsigned slot = wbase + (unsigned)(incl - ntri);
            unsigned col[8];
            if (COLOR) {
                // the corner interpolations succeeded for the SDF, so they succeed for the colour (same eight weight tests)
#pragma unroll
                for (int c = 0; c < 8; ++c) col[c] = vol_interp_color(V, A.color, pos[c][0], pos[c][1], pos[c][2]);
            }
#pragma unroll 1
            for (int k = 0; k < 16; k += 3) {
                const unsigned e0 = (unsigned)((packed >> (4 * k)) & 0xF);
                if (e0 == 0xF) break;
                const unsigned e1 = (unsigned)((packed >> (4 * (k + 1))) & 0xF), e2 = (unsigned)((packed >> (4 * (k + 2))) & 0xF);
                if (slot < A.max_tris) {
                    HkfTriangle t;
                    mc_vertex(pos[ea[e0]], pos[eb[e0]], dist[ea[e0]], dist[eb[e0]], t.v0.pos);
                    mc_vertex(pos[ea[e1]], pos[eb[e1]], dist[ea[e1]], dist[eb[e1]], t.v1.pos);
                    mc_vertex(pos[ea[e2]], pos[eb[e2]], dist[ea[e2]], dist[eb[e2]], t.v2.pos);
                    if (COLOR) {
                        mc_vertex_color(col[ea[e0]], col[eb[e0]], dist[ea[e0]], dist[eb[e0]], t.v0.color);
                        mc_vertex_color(col[ea[e1]], col[eb[e1]], dist[ea[e1]], dist[eb[e1]], t.v1.color);
                        mc_vertex_color(col[ea[e2]], col[eb[e2]], dist[ea[e2]], dist[eb[e2]], t.v2.color);
                    } else {
                        t.v0.color[0] = t.v0.color[1] = t.v0.color[2] = 0.f;
                        t.v1.color[0] = t.v1.color[1] = t.v1.color[2] = 0.f;
                        t.v2.color[0] = t.v2.color[1] = t.v2.color[2] = 0.f;
                    }
                    A.tris[slot] = t;
                }
                ++slot;
            }
        }
    }
}

extern "C" int hkf_marchingcube(hkf_ctx* c, int has_color, float threshold)
{
    if (!c) return HKF_ERR_ARG;
    if (has_color && !c->color) { c->err = "context was created without use_color"; return HKF_ERR_STATE; }
    if (!c->tris) { c->err = "max_triangles is 0"; return HKF_ERR_STATE; }
    HKF_CUDA(c, cudaSetDevice(c->device));
    if (!g_tables_uploaded[c->device & 63]) {
        HKF_CUDA(c, cudaMemcpyToSymbol(c_tri, HKF_MC_TRI_PACKED, sizeof(HKF_MC_TRI_PACKED)));
        g_tables_uploaded[c->device & 63] = true;
    }
    HKF_CUDA(c, cudaMemsetAsync(&c->st->mc_count, 0, sizeof(unsigned), c->stream));
    McArgs A;
    A.V = make_volview(c);
    A.z0 = (int)c->own_z0; A.z1 = (int)c->own_z1;
    A.threshold = threshold; A.tris = c->tris; A.max_tris = c->ip.max_triangles;
    A.color = has_color ? c->color : nullptr;
    const bool slab = !(c->z_lo == 0 && c->z_hi == c->ip.volume_resolution);
    const int grid = c->sm_count * 8;
    // candidate list: one entry per 16 cells is far more than the shell around a surface needs; if a volume ever
    // overflows it the single-pass kernel (same results) is used instead
    const size_t ncells = (size_t)(A.z1 - A.z0) * A.V.R * A.V.R;
    if (ncells > 0x100000000ull) {
        // candidate ids are 32-bit cell numbers: a slab of more than 2^32 owned cells (R = 2048 on one context) takes the
        // single-pass kernel, which indexes by (row, x)
        if (slab) k_marching_cubes<true, false><<<grid, 256, 0, c->stream>>>(A, &c->st->mc_count, nullptr, nullptr);
        else if (has_color) k_marching_cubes<false, false, true><<<grid, 256, 0, c->stream>>>(A, &c->st->mc_count, nullptr, nullptr);
        else k_marching_cubes<false, false><<<grid, 256, 0, c->stream>>>(A, &c->st->mc_count, nullptr, nullptr);
        HKF_LAUNCH_CHECK(c);
        return HKF_OK;
    }
    const unsigned cap = (unsigned)(ncells / 16 < 65536 ? 65536 : ncells / 16);
    if (!c->mc_cand || c->mc_cand_cap < cap) {
        if (c->mc_cand) cudaFree(c->mc_cand);
        c->mc_cand = nullptr;
        HKF_CUDA(c, cudaMalloc(&c->mc_cand, (size_t)cap * sizeof(unsigned)));
        c->mc_cand_cap = cap;
    }
    HKF_CUDA(c, cudaMemsetAsync(&c->st->mc_ncand, 0, sizeof(unsigned), c->stream));
    if (slab) k_mc_candidates<true><<<grid, 256, 0, c->stream>>>(A, c->mc_cand, cap, &c->st->mc_ncand);
    else k_mc_candidates<false><<<grid, 256, 0, c->stream>>>(A, c->mc_cand, cap, &c->st->mc_ncand);
    HKF_LAUNCH_CHECK(c);
    HKF_CUDA(c, cudaMemcpyAsync(&c->st_host->mc_ncand, &c->st->mc_ncand, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    HKF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->st_host->mc_ncand <= cap) {
        const int g2 = (int)div_up(c->st_host->mc_ncand, 256);
        if (g2 > 0) {
            if (slab) k_marching_cubes<true, true><<<g2, 256, 0, c->stream>>>(A, &c->st->mc_count, c->mc_cand, &c->st->mc_ncand);
            else if (has_color) k_marching_cubes<false, true, true><<<g2, 256, 0, c->stream>>>(A, &c->st->mc_count, c->mc_cand, &c->st->mc_ncand);
            else k_marching_cubes<false, true><<<g2, 256, 0, c->stream>>>(A, &c->st->mc_count, c->mc_cand, &c->st->mc_ncand);
            HKF_LAUNCH_CHECK(c);
        }
    } else {
        if (slab) k_marching_cubes<true, false><<<grid, 256, 0, c->stream>>>(A, &c->st->mc_count, nullptr, nullptr);
        else if (has_color) k_marching_cubes<false, false, true><<<grid, 256, 0, c->stream>>>(A, &c->st->mc_count, nullptr, nullptr);
        else k_marching_cubes<false, false><<<grid, 256, 0, c->stream>>>(A, &c->st->mc_count, nullptr, nullptr);
        HKF_LAUNCH_CHECK(c);
    }
    return HKF_OK;
}

extern "C" int hkf_triangle_count(hkf_ctx* c, unsigned* n)
{
    if (!c || !n) return HKF_ERR_ARG;
    HKF_CUDA(c, cudaMemcpyAsync(&c->st_host->mc_count, &c->st->mc_count, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    HKF_CUDA(c, cudaStreamSynchronize(c->stream));
    *n = c->st_host->mc_count;
    return HKF_OK;
}
