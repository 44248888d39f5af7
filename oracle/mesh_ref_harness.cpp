// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// mesh_ref_harness.cpp: the reference's own mesh post-processing (mLib-derived ml::MeshData,
// /root/reference/src/utils/mesh/meshData.{h,cpp}) driven exactly as MeshGeneratorMarchingcube::copyTrianglesToCPU +
// saveMesh drive it (src/MeshGeneratorMarchingcube.cpp:30-97), minus the file write.  Compiled by oracle/Makefile
// from where the sources lie (g++ -fpermissive: the class templates hold a float2 expression that only a lenient
// front end lets through uninstantiated) into oracle/_ref/libmesh_ref.so.  Nothing of the reference is copied here.
#include <iostream>
#include <string>
#include <cstring>
using namespace std;
#include "cuda/cuda_declar.h"     // the include order of the reference's own translation units (float3 operators first)
#include "cuda/Mat.h"
#include "utils/mesh/meshData.h"

// tris: n x 3 vertices x {pos[3], colour[3]} (HkfTriangle / the reference's Triangle).  Outputs are sized by the caller
// for 3 n vertices / n faces.  Returns 0.
extern "C" int ref_mesh_process(const float* tris, unsigned n, int use_color, unsigned* out_nv, unsigned* out_nf,
                                float* out_vertices3, float* out_normals3, float* out_colors4, unsigned* out_faces3)
{
    ml::MeshDataf m;
    m.m_Vertices.resize(3 * (size_t)n);
    if (use_color) m.m_Colors.resize(3 * (size_t)n);
    for (unsigned i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            const float* v = tris + ((size_t)i * 3 + k) * 6;
            m.m_Vertices[3 * i + k] = make_float3(v[0], v[1], v[2]);
            if (use_color) m.m_Colors[3 * i + k] = make_float4(v[5], v[4], v[3], 1.0);      // (color.z, color.y, color.x, 1), :53-55
        }
    m.m_FaceIndicesVertices.resize(m.m_Vertices.size());                                    // as the reference: 3 n "faces", n used
    for (unsigned i = 0; i < (unsigned)m.m_Vertices.size() / 3; i++) {
        m.m_FaceIndicesVertices[i][0] = 3 * i + 0;
        m.m_FaceIndicesVertices[i][1] = 3 * i + 1;
        m.m_FaceIndicesVertices[i][2] = 3 * i + 2;
    }
    std::streambuf* old = std::cout.rdbuf(nullptr);
    m.mergeCloseVertices(0.0001f, true);
    m.removeDuplicateFaces();
    m.computeVertexNormals();
    std::cout.rdbuf(old);
    *out_nv = (unsigned)m.m_Vertices.size();
    *out_nf = (unsigned)m.m_FaceIndicesVertices.size();
    for (size_t i = 0; i < m.m_Vertices.size(); ++i) {
        out_vertices3[3 * i] = m.m_Vertices[i].x; out_vertices3[3 * i + 1] = m.m_Vertices[i].y; out_vertices3[3 * i + 2] = m.m_Vertices[i].z;
        out_normals3[3 * i] = m.m_Normals[i].x; out_normals3[3 * i + 1] = m.m_Normals[i].y; out_normals3[3 * i + 2] = m.m_Normals[i].z;
        if (use_color) { out_colors4[4 * i] = m.m_Colors[i].x; out_colors4[4 * i + 1] = m.m_Colors[i].y; out_colors4[4 * i + 2] = m.m_Colors[i].z; out_colors4[4 * i + 3] = m.m_Colors[i].w; }
    }
    for (size_t i = 0; i < m.m_FaceIndicesVertices.size(); ++i)
        for (int k = 0; k < 3; ++k) out_faces3[3 * i + k] = m.m_FaceIndicesVertices[i][k];
    return 0;
}
