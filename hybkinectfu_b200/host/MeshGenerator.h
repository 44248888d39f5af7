// MeshGenerator.h — mesh-generator interface (src/MeshGenerator.h:11-17), unchanged.
#pragma once
#include <string>
class MeshGenerator {
public:
    virtual ~MeshGenerator() {}
    virtual void generateMesh() = 0;
    virtual bool saveMesh(const std::string& filename) = 0;
};
