// device_scene.h — host-side owner of the device-resident scene (geometry records, BVH nodes, materials,
// textures, lights).  Built once per initScene (RenderEngine/renderer/OptixRenderer.cpp:436-485).
#pragma once
#include "scene.cuh"
#include <vector>
#include <stdexcept>
#include <string>

namespace opr {

class DeviceScene {
public:
    SceneView view{};
    unsigned numPrims = 0, numTriangles = 0;
    int numNodes = 0, bvhLevels = 0, plocRounds = 0;   // 4-wide nodes, collapse levels, PLOC rounds
    float sahCost = 0.f;                                // SAH cost of the binary tree / root area
    float3 aabbMin{}, aabbMax{};
    std::vector<opr_light> hostLights;
    const void* traversalBlock = nullptr;               // quantised nodes | fp32 nodes | leaf triangles, one allocation
    size_t traversalNodeBytes = 0, traversalBytes = 0;  // bytes of the two node arrays; of the whole block

    DeviceScene() {}
    ~DeviceScene();
    DeviceScene(const DeviceScene&) = delete;
    DeviceScene& operator=(const DeviceScene&) = delete;
    void load(const opr_scene_desc* d, cudaStream_t st);
    void release();

private:
    void buildBvh(const std::vector<float4>& primLo, const std::vector<float4>& primHi, cudaStream_t st);
    std::vector<void*> owned;
    std::vector<cudaTextureObject_t> texObjects;
    std::vector<cudaArray_t> texArrays;
};

} // namespace opr
