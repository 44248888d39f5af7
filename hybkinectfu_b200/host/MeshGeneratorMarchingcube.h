// MeshGeneratorMarchingcube — marching-cubes driver (src/MeshGeneratorMarchingcube.{h,cpp}): generateMesh() launches
// the extraction with threshold 300*S/R (never binds since |tsdf| <= 1), copyTrianglesToCPU() downloads the soup,
// saveMesh() turns it into an indexed mesh with vertex normals (and vertex colours with use_color) and writes OBJ / PLY /
// OFF (by extension).  The post-processing is the reference's (mLib's approximate 1e-4 hash-grid merge, degenerate and
// duplicate face removal, vertex normals), applied to the soup in canonical order so that the file is a deterministic
// function of the volume - see weld().  saveMesh returns true on success (the reference always returns 0).
#pragma once
#include <array>
#include <vector>
#include "MeshGenerator.h"
#include "../../include/hkf.h"

class MeshGeneratorMarchingcube : public MeshGenerator {
public:
    void generateMesh() override;
    bool saveMesh(const std::string& filename) override;
    unsigned triangleCount() const { return (unsigned)_soup.size(); }
    size_t vertexCount() const { return _vertices.size(); }
    size_t faceCount() const { return _faces.size(); }
    // soup -> canonical indexed mesh (exposed for tests)
    static void weld(std::vector<HkfTriangle> soup, std::vector<std::array<float, 3>>& vertices, std::vector<std::array<unsigned, 3>>& faces,
                     std::vector<std::array<float, 4>>* colors = nullptr);
    // per-vertex normals as ml::MeshData::computeVertexNormals does (src/utils/mesh/meshData.h:713-736): the unit face
    // normal cross(v1 - v0, v2 - v0) added to its three vertices, then normalised
    static void vertexNormals(const std::vector<std::array<float, 3>>& vertices, const std::vector<std::array<unsigned, 3>>& faces,
                              std::vector<std::array<float, 3>>& normals);
    // format by extension: .obj (v / vn / f v//vn), .ply (ascii, with normals), .off
    static bool writeMesh(const std::string& filename, const std::vector<std::array<float, 3>>& vertices, const std::vector<std::array<unsigned, 3>>& faces,
                          const std::vector<std::array<float, 3>>& normals, const std::vector<std::array<float, 4>>* colors = nullptr);
    static bool writeObj(const std::string& filename, const std::vector<std::array<float, 3>>& vertices, const std::vector<std::array<unsigned, 3>>& faces);
protected:
    bool copyTrianglesToCPU();
    std::vector<HkfTriangle> _soup;
    std::vector<std::array<float, 3>> _vertices;
    std::vector<std::array<unsigned, 3>> _faces;
    std::vector<std::array<float, 3>> _normals;
    std::vector<std::array<float, 4>> _colors;      // (r, g, b, 1) per vertex with use_color (m_Colors of the reference)
};
