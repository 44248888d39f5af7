#include "MeshGeneratorMarchingcube.h"
#include <charconv>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include "AppParams.h"
#include "CudaDeviceDataMan.h"

void MeshGeneratorMarchingcube::generateMesh()
{
    const AppParams* P = AppParams::instance();
    hkf_marchingcube(CudaDeviceDataMan::instance()->ctx(), P->_switch_params.useRGBData ? 1 : 0, 300 * P->_volume_params.fVolumeMeterSize / P->_volume_params.nResolution);
}

bool MeshGeneratorMarchingcube::copyTrianglesToCPU()
{
    hkf_ctx* c = CudaDeviceDataMan::instance()->ctx();
    unsigned n = 0;
    if (hkf_triangle_count(c, &n) != HKF_OK) return false;
    n = std::min(n, AppParams::instance()->_marchingcube_params.uMaxTriangles);
    std::cout << "Marching Cubes triangles = " << n << std::endl;
    if (n == 0) return false;
    _soup.resize(n);
    return hkf_download(c, HKF_BUF_TRIANGLES, 0, _soup.data(), _soup.size() * sizeof(HkfTriangle)) == (long)(n * sizeof(HkfTriangle));
}

namespace {
inline int sign_of(float v) { return (0 < v) - (v < 0); }                      // sign(), src/cuda/cuda_declar.h:95-97
// toVirtualVoxelPos (src/utils/mesh/meshData.h:750-753): v * (float)(1.0 / voxelSize) + sign(v) * 0.5, truncated
inline int32_t virtual_cell(float v, float inv_size) { const float t = v * inv_size; return (int32_t)(t + (float)(sign_of(v) * 0.5)); }
}

// The reference's post-processing of the triangle soup, restated (src/MeshGeneratorMarchingcube.cpp:61-97 ->
// ml::MeshData::mergeCloseVertices(1e-4, approx = true), removeDegeneratedFaces, removeDuplicateFaces;
// src/utils/mesh/meshData.cpp:41-86,166-314).  The approximate merge is a hash grid of 1e-4 m cells: a vertex joins the
// FIRST registered vertex found in the 27 cells around its own (x outermost, no distance test), otherwise it
// registers itself in its cell - so vertices up to two cells apart may merge and the result depends on the order in
// which the soup is walked.  Marching cubes appends triangles in a nondeterministic order (in the reference too), so
// the soup is first put in canonical order (sorted by its nine position floats): the mesh - hence the file - is then a
// deterministic function of the volume, and equals what the reference's code makes of the same ordered soup
// (tests/test_mesh_reference.py drives the reference's own ml::MeshData next to this).
namespace {
// Open-addressing tables for the two passes of the weld (linear probing, power-of-two capacity, load <= 1/4): a lookup
// that misses - 26 of the 27 neighbour cells of almost every vertex - costs one cache line, and the 27 lines of a
// vertex are requested together before the first one is looked at.
struct CellTable {
    struct Entry { int32_t x, y, z; uint32_t id; };
    std::vector<Entry> e;
    size_t mask, count = 0;
    explicit CellTable(size_t expected)
    {
        size_t cap = 64;
        while (cap < 2 * expected) cap <<= 1;
        e.assign(cap, Entry{ 0, 0, 0, ~0u });
        mask = cap - 1;
    }
    // hash = finish(hx(x) ^ hy(y) ^ hz(z)): the three per-axis parts are shared by the 27 cells around a vertex
    static uint64_t hx(int32_t x) { return (uint32_t)x * 0x9E3779B97F4A7C15ull; }
    static uint64_t hy(int32_t y) { const uint64_t h = (uint32_t)y * 0xC2B2AE3D27D4EB4Full; return h ^ (h >> 31); }
    static uint64_t hz(int32_t z) { const uint64_t h = (uint32_t)z * 0x165667B19E3779F9ull; return (h << 17) | (h >> 47); }
    static uint64_t finish(uint64_t h) { h ^= h >> 32; return (h * 0xff51afd7ed558ccdull) >> 20; }
    static uint64_t hash(int32_t x, int32_t y, int32_t z) { return finish(hx(x) ^ hy(y) ^ hz(z)); }
    size_t slot(int32_t x, int32_t y, int32_t z) const { return (size_t)hash(x, y, z) & mask; }
    uint32_t find(size_t s, int32_t x, int32_t y, int32_t z) const
    {
        for (;; s = (s + 1) & mask) {
            const Entry& q = e[s];
            if (q.id == ~0u) return ~0u;
            if (q.x == x && q.y == y && q.z == z) return q.id;
        }
    }
    void place(int32_t x, int32_t y, int32_t z, uint32_t id)
    {
        size_t s = slot(x, y, z);
        while (e[s].id != ~0u) s = (s + 1) & mask;
        e[s] = Entry{ x, y, z, id };
    }
    // the table stays small enough for the caches as long as the soup welds well (a vertex is shared by ~6 triangles)
    // and doubles when it does not: load factor <= 1/2
    void add(int32_t x, int32_t y, int32_t z, uint32_t id)
    {
        if (2 * (count + 1) > e.size()) {
            std::vector<Entry> old;
            old.swap(e);
            e.assign(2 * old.size(), Entry{ 0, 0, 0, ~0u });
            mask = e.size() - 1;
            for (const Entry& q : old) if (q.id != ~0u) place(q.x, q.y, q.z, q.id);
        }
        place(x, y, z, id);
        ++count;
    }
};
struct SortedTriangle { float key[9]; uint32_t index; };
inline bool soup_less(const SortedTriangle& a, const SortedTriangle& b)
{
    for (int i = 0; i < 9; ++i) { if (a.key[i] < b.key[i]) return true; if (b.key[i] < a.key[i]) return false; }
    return false;
}
}

void MeshGeneratorMarchingcube::weld(std::vector<HkfTriangle> soup, std::vector<std::array<float, 3>>& vertices,
                                     std::vector<std::array<unsigned, 3>>& faces, std::vector<std::array<float, 4>>* colors)
{
    const size_t n = soup.size();
    // canonical order: the nine position floats, lexicographically (keys sorted in place: the first float decides almost
    // every comparison, and it sits in the element that is being moved, not behind an index)
    std::vector<SortedTriangle> order(n);
    for (size_t t = 0; t < n; ++t) {
        const HkfVertex* v = &soup[t].v0;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) order[t].key[i * 3 + j] = v[i].pos[j];
        order[t].index = (uint32_t)t;
    }
    std::sort(order.begin(), order.end(), soup_less);
    const float thresh = 0.0001f;
    const float inv_size = (float)(1.0 / thresh);
    CellTable grid(n / 2 + 16);               // the usual outcome: about half as many vertices as triangles
    vertices.clear(); faces.clear();
    vertices.reserve(n); faces.reserve(n);
    if (colors) { colors->clear(); colors->reserve(n); }
    for (const SortedTriangle& st : order) {
        const HkfVertex* tv = &soup[st.index].v0;
        std::array<unsigned, 3> f;
        for (int i = 0; i < 3; ++i) {
            const float* p = tv[i].pos;
            const int32_t cx = virtual_cell(p[0], inv_size), cy = virtual_cell(p[1], inv_size), cz = virtual_cell(p[2], inv_size);
            size_t slots[27];
            uint64_t ax[3], ay[3], az[3];
            for (int d = 0; d < 3; ++d) { ax[d] = CellTable::hx(cx + d - 1); ay[d] = CellTable::hy(cy + d - 1); az[d] = CellTable::hz(cz + d - 1); }
            int k = 0;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    for (int c = 0; c < 3; ++c, ++k) { slots[k] = (size_t)CellTable::finish(ax[a] ^ ay[b] ^ az[c]) & grid.mask; __builtin_prefetch(&grid.e[slots[k]]); }
            unsigned id = ~0u;
            for (int a = -1; a <= 1 && id == ~0u; ++a)           // x outermost, the first registered vertex found wins
                for (int b = -1; b <= 1 && id == ~0u; ++b)
                    for (int c = -1; c <= 1; ++c) {
                        id = grid.find(slots[(a + 1) * 9 + (b + 1) * 3 + (c + 1)], cx + a, cy + b, cz + c);
                        if (id != ~0u) break;
                    }
            if (id == ~0u) {
                id = (unsigned)vertices.size();
                grid.add(cx, cy, cz, id);
                vertices.push_back({ p[0], p[1], p[2] });
                if (colors) colors->push_back({ tv[i].color[2], tv[i].color[1], tv[i].color[0], 1.0f });   // (z, y, x, 1): BGR -> RGB, :53-55
            }
            f[i] = id;
        }
        if (f[0] == f[1] || f[1] == f[2] || f[0] == f[2]) continue;                 // removeDegeneratedFaces
        faces.push_back(f);
    }
    // removeDuplicateFaces: same vertex set, the first one stays
    CellTable seen(faces.size());
    std::vector<std::array<unsigned, 3>> uniq;
    uniq.reserve(faces.size());
    for (const auto& f : faces) {
        std::array<uint32_t, 3> s3 = { f[0], f[1], f[2] };
        std::sort(s3.begin(), s3.end());
        const size_t s = seen.slot((int32_t)s3[0], (int32_t)s3[1], (int32_t)s3[2]);
        if (seen.find(s, (int32_t)s3[0], (int32_t)s3[1], (int32_t)s3[2]) != ~0u) continue;
        seen.add((int32_t)s3[0], (int32_t)s3[1], (int32_t)s3[2], 0u);
        uniq.push_back(f);
    }
    faces.swap(uniq);
}

bool MeshGeneratorMarchingcube::writeObj(const std::string& filename, const std::vector<std::array<float, 3>>& vertices,
                                         const std::vector<std::array<unsigned, 3>>& faces)
{
    FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "# hybKinectfu marching-cubes mesh\n# vertices %zu\n# faces %zu\n", vertices.size(), faces.size());
    for (const auto& v : vertices) std::fprintf(f, "v %.9g %.9g %.9g\n", v[0], v[1], v[2]);
    for (const auto& t : faces) std::fprintf(f, "f %u %u %u\n", t[0] + 1, t[1] + 1, t[2] + 1);
    return std::fclose(f) == 0;
}

void MeshGeneratorMarchingcube::vertexNormals(const std::vector<std::array<float, 3>>& v, const std::vector<std::array<unsigned, 3>>& faces,
                                              std::vector<std::array<float, 3>>& normals)
{
    auto unit = [](std::array<float, 3> n) {
        const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);       // dot() order x, y, z (cuda_declar.h)
        if (len < 1e-8) return std::array<float, 3>{ 0.f, 0.f, 0.f };               // normalize() of src/cuda/cuda_declar.h:89-94
        const float r = (float)(1.0 / len);                                         // v * (1.0 / len): the double quotient narrowed
        return std::array<float, 3>{ n[0] * r, n[1] * r, n[2] * r };
    };
    normals.assign(v.size(), { 0.f, 0.f, 0.f });
    for (const auto& f : faces) {
        const auto &a = v[f[0]], &b = v[f[1]], &c = v[f[2]];
        const float e1[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, e2[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
        const auto n = unit({ e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] });
        for (unsigned idx : f) for (int k = 0; k < 3; ++k) normals[idx][k] += n[k];
    }
    for (auto& n : normals) n = unit(n);
}

namespace {
// Text output through std::to_chars (same characters as printf's %.9g / %.6g / %u in the C locale, several times
// faster than fprintf: a 553 k-triangle OBJ is 42 MB of numbers) into one buffer that is flushed in large pieces.
class TextOut {
public:
    explicit TextOut(FILE* f) : _f(f) { _buf.reserve(1 << 22); }
    void str(const char* s) { _buf.append(s); }
    void num(float v, int precision)
    {
        char tmp[48];
        auto r = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::general, precision);
        _buf.append(tmp, r.ptr);
    }
    void num(unsigned v)
    {
        char tmp[16];
        auto r = std::to_chars(tmp, tmp + sizeof(tmp), v);
        _buf.append(tmp, r.ptr);
    }
    void maybe_flush() { if (_buf.size() > (1u << 22) - 4096) flush(); }
    bool flush()
    {
        if (!_buf.empty() && std::fwrite(_buf.data(), 1, _buf.size(), _f) != _buf.size()) _ok = false;
        _buf.clear();
        return _ok;
    }
private:
    FILE* _f;
    std::string _buf;
    bool _ok = true;
};
}

bool MeshGeneratorMarchingcube::writeMesh(const std::string& filename, const std::vector<std::array<float, 3>>& v,
                                          const std::vector<std::array<unsigned, 3>>& faces, const std::vector<std::array<float, 3>>& normals,
                                          const std::vector<std::array<float, 4>>* colors)
{
    const bool with_c = colors && colors->size() == v.size();
    auto byte_of = [](float c) { const float x = c * 255.0f + 0.5f; return (unsigned)(x < 0.f ? 0.f : x > 255.f ? 255.f : x); };
    const size_t dot = filename.find_last_of('.');
    std::string ext = dot == std::string::npos ? "" : filename.substr(dot + 1);
    for (auto& ch : ext) ch = (char)std::tolower((unsigned char)ch);
    const bool with_n = normals.size() == v.size();
    FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) return false;
    TextOut o(f);
    auto vec3 = [&](const char* lead, const std::array<float, 3>& p, int precision, const char* tail) {
        o.str(lead); o.num(p[0], precision); o.str(" "); o.num(p[1], precision); o.str(" "); o.num(p[2], precision); o.str(tail);
        o.maybe_flush();
    };
    if (ext == "ply") {
        std::fprintf(f, "ply\nformat ascii 1.0\ncomment hybKinectfu marching-cubes mesh\nelement vertex %zu\nproperty float x\nproperty float y\nproperty float z\n", v.size());
        if (with_n) std::fprintf(f, "property float nx\nproperty float ny\nproperty float nz\n");
        if (with_c) std::fprintf(f, "property uchar red\nproperty uchar green\nproperty uchar blue\n");
        std::fprintf(f, "element face %zu\nproperty list uchar int vertex_indices\nend_header\n", faces.size());
        for (size_t i = 0; i < v.size(); ++i) {
            vec3("", v[i], 9, (with_n || with_c) ? "" : "\n");
            if (with_n) vec3(" ", normals[i], 6, with_c ? "" : "\n");
            if (with_c) { const auto& c = (*colors)[i]; o.str(" "); o.num(byte_of(c[0])); o.str(" "); o.num(byte_of(c[1])); o.str(" "); o.num(byte_of(c[2])); o.str("\n"); }
        }
        for (const auto& t : faces) { o.str("3 "); o.num(t[0]); o.str(" "); o.num(t[1]); o.str(" "); o.num(t[2]); o.str("\n"); o.maybe_flush(); }
    } else if (ext == "off") {
        std::fprintf(f, "OFF\n%zu %zu 0\n", v.size(), faces.size());
        for (const auto& p : v) vec3("", p, 9, "\n");
        for (const auto& t : faces) { o.str("3 "); o.num(t[0]); o.str(" "); o.num(t[1]); o.str(" "); o.num(t[2]); o.str("\n"); o.maybe_flush(); }
    } else {
        std::fprintf(f, "# hybKinectfu marching-cubes mesh\n# vertices %zu\n# faces %zu\n", v.size(), faces.size());
        for (size_t i = 0; i < v.size(); ++i) {
            vec3("v ", v[i], 9, with_c ? "" : "\n");
            if (with_c) { const auto& c = (*colors)[i]; vec3(" ", { c[0], c[1], c[2] }, 6, "\n"); }     // the common "v x y z r g b" extension
        }
        if (with_n) for (const auto& n : normals) vec3("vn ", n, 6, "\n");
        for (const auto& t : faces) {
            o.str("f ");
            for (int i = 0; i < 3; ++i) {
                o.num(t[i] + 1);
                if (with_n) { o.str("//"); o.num(t[i] + 1); }
                o.str(i < 2 ? " " : "\n");
            }
            o.maybe_flush();
        }
    }
    const bool ok = o.flush();
    return (std::fclose(f) == 0) && ok;
}

bool MeshGeneratorMarchingcube::saveMesh(const std::string& filename)
{
    if (!copyTrianglesToCPU()) { std::cout << "copy triangles to cpu failed" << std::endl; return false; }
    const bool color = AppParams::instance()->_switch_params.useRGBData;
    weld(_soup, _vertices, _faces, color ? &_colors : nullptr);
    vertexNormals(_vertices, _faces, _normals);
    std::cout << "vertices :\t" << _vertices.size() << "\nfaces :\t" << _faces.size() << std::endl;
    return writeMesh(filename, _vertices, _faces, _normals, color ? &_colors : nullptr);
}
