#include "MeshGeneratorMarchingcube.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include "AppParams.h"
#include "CudaDeviceDataMan.h"

void MeshGeneratorMarchingcube::generateMesh()
{
    const AppParams* P = AppParams::instance();
    hkf_marchingcube(CudaDeviceDataMan::instance()->ctx(), 0, 300 * P->_volume_params.fVolumeMeterSize / P->_volume_params.nResolution);
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

void MeshGeneratorMarchingcube::weld(std::vector<HkfTriangle> soup, std::vector<std::array<float, 3>>& vertices,
                                     std::vector<std::array<unsigned, 3>>& faces)
{
    auto key = [](const HkfTriangle& t) { std::array<float, 9> k; const HkfVertex* v = &t.v0; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) k[i * 3 + j] = v[i].pos[j]; return k; };
    std::sort(soup.begin(), soup.end(), [&](const HkfTriangle& a, const HkfTriangle& b) { return key(a) < key(b); });
    std::map<std::array<float, 3>, unsigned> index;
    vertices.clear(); faces.clear();
    for (const HkfTriangle& t : soup) {
        const HkfVertex* v = &t.v0;
        std::array<unsigned, 3> f;
        for (int i = 0; i < 3; ++i) {
            std::array<float, 3> p = { v[i].pos[0], v[i].pos[1], v[i].pos[2] };
            auto it = index.find(p);
            if (it == index.end()) { it = index.emplace(p, (unsigned)vertices.size()).first; vertices.push_back(p); }
            f[i] = it->second;
        }
        if (f[0] == f[1] || f[1] == f[2] || f[0] == f[2]) continue;                 // degenerate after welding
        faces.push_back(f);
    }
    // duplicate faces (same vertex set) are dropped, as ml::MeshData::removeDuplicateFaces does
    std::vector<std::array<unsigned, 3>> uniq;
    std::map<std::array<unsigned, 3>, bool> seen;
    for (const auto& f : faces) {
        std::array<unsigned, 3> s = f;
        std::sort(s.begin(), s.end());
        if (seen.emplace(s, true).second) uniq.push_back(f);
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

bool MeshGeneratorMarchingcube::saveMesh(const std::string& filename)
{
    if (!copyTrianglesToCPU()) { std::cout << "copy triangles to cpu failed" << std::endl; return false; }
    weld(_soup, _vertices, _faces);
    std::cout << "vertices :\t" << _vertices.size() << "\nfaces :\t" << _faces.size() << std::endl;
    return writeObj(filename, _vertices, _faces);
}
