// scene.cu — scene upload: geometry records, per-vertex attributes, materials, textures, lights
// (the values the reference hands to OptiX in scene/Scene.cpp:361-467, scene/Cornell.cpp:46-197, material/*.cpp).
// The acceleration structure is built in bvh.cu.
#include "device_scene.h"
#include "../../include/opr.h"
#include <vector>
#include <cstring>
#include <cmath>
#include <stdexcept>
#include <string>

namespace opr {

// ---- host side ------------------------------------------------------------------------------------------
template <class T> static T* upload(const std::vector<T>& v, std::vector<void*>& owned)
{
    T* d = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    devAlloc((void**)&d, bytes);
    if (!v.empty()) OPR_CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    owned.push_back(d);
    return d;
}

static inline float3 h3(const float* p) { return make_float3(p[0], p[1], p[2]); }

DeviceScene::~DeviceScene() { release(); }

void DeviceScene::release()
{
    for (void* p : owned) devFree(p);
    owned.clear();
    for (auto t : texObjects) cudaDestroyTextureObject(t);
    texObjects.clear();
    for (auto a : texArrays) cudaFreeArray(a);
    texArrays.clear();
    memset(&view, 0, sizeof view);
}

void DeviceScene::load(const opr_scene_desc* d, cudaStream_t st)
{
    release();
    { char msg[160]; if (opr_scene_desc_check(d, msg, (int)sizeof msg)) throw std::runtime_error(msg); }
    if (d->numLights <= 0) throw std::runtime_error("No lights exists in this scene.");   // OptixRenderer.cpp:443-447
    if (d->numLights > 65535) throw std::runtime_error("more than 65535 lights: the direct-light sample records hold 16-bit light indices");

    // ---- textures (material/Texture.cpp:109-130): RGBA8, wrap repeat, bilinear, normalised float reads
    std::vector<cudaTextureObject_t> tex(d->numTextures);
    for (int i = 0; i < d->numTextures; i++) {
        const opr_texture& t = d->textures[i];
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr;
        OPR_CUDA_CHECK(cudaMallocArray(&arr, &cd, t.width, t.height));
        OPR_CUDA_CHECK(cudaMemcpy2DToArray(arr, 0, 0, t.rgba, (size_t)t.width * 4, (size_t)t.width * 4, t.height, cudaMemcpyHostToDevice));
        texArrays.push_back(arr);
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
        cudaTextureDesc td; memset(&td, 0, sizeof td);
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        OPR_CUDA_CHECK(cudaCreateTextureObject(&tex[i], &rd, &td, nullptr));
        texObjects.push_back(tex[i]);
    }

    // ---- materials (values of registerGeometryInstanceValues, see include/opr_scene.h)
    std::vector<DMaterial> mats(d->numMaterials);
    for (int i = 0; i < d->numMaterials; i++) {
        const opr_material& m = d->materials[i];
        DMaterial& o = mats[i]; memset(&o, 0, sizeof o);
        o.kind = m.kind;
        o.Kd = h3(m.Kd); o.Ks = h3(m.Ks); o.exponent = m.exponent; o.Kr = h3(m.Kr); o.Kt = h3(m.Kt); o.ior = m.indexOfRefraction;
        float3 ppa = h3(m.power); ppa.x *= m.inverseArea; ppa.y *= m.inverseArea; ppa.z *= m.inverseArea;   // DiffuseEmitter.cpp:52-56
        o.powerPerArea = ppa;
        o.Lemit = make_float3(m.power[0] * m.inverseArea * INV_PI_F, m.power[1] * m.inverseArea * INV_PI_F, m.power[2] * m.inverseArea * INV_PI_F);   // :61
        o.inverseArea = m.inverseArea;
        if (m.kind == OPR_MAT_TEXTURE) {
            if (m.diffuseTexture < 0 || m.diffuseTexture >= d->numTextures) throw std::runtime_error("Texture material without a diffuse texture");
            o.diffuseTex = tex[m.diffuseTexture];
            o.hasNormalTex = (m.normalTexture >= 0);
            if (o.hasNormalTex) o.normalTex = tex[m.normalTexture];
        }
    }

    // ---- primitives.  Degenerate ones are dropped exactly as the bounds programs invalidate them
    // (TriangleMesh.cu:105-117, parallelogram.cu:89-97).
    std::vector<float4> triRec, triAttr, pgramRec, sphereRec, vN, vT, vB;
    std::vector<int4> triShade;
    std::vector<float2> vUV;
    std::vector<int> pgramMat, sphereMat;
    std::vector<float4> primLo, primHi;
    size_t totalV = 0;
    for (int m = 0; m < d->numMeshes; m++) totalV += d->meshes[m].numVertices;
    vN.assign(std::max<size_t>(totalV, 1), make_float4(0, 0, 0, 0)); vT = vN; vB = vN;
    vUV.assign(std::max<size_t>(totalV, 1), make_float2(0, 0));
    auto addPrim = [&](float3 lo, float3 hi, int ref) {
        // conservative inflation: the fp32 slab test must never cull a box whose primitive test would pass
        float3 ext = make_float3(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z);
        float scale = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
        float eps = 4e-6f * fmaxf(scale, fmaxf(ext.x, fmaxf(ext.y, ext.z))) + 1e-30f;
        primLo.push_back(make_float4(lo.x - eps, lo.y - eps, lo.z - eps, __builtin_bit_cast(float, ~ref)));
        primHi.push_back(make_float4(hi.x + eps, hi.y + eps, hi.z + eps, 0.f));
    };
    size_t vbase = 0;
    for (int m = 0; m < d->numMeshes; m++) {
        const opr_mesh& me = d->meshes[m];
        int flags = (me.normals ? TRI_HAS_NORMALS : 0) | (me.uvs ? TRI_HAS_UVS : 0) | ((me.normals && me.tangents && me.bitangents) ? TRI_HAS_TANGENTS : 0);
        for (int v = 0; v < me.numVertices; v++) {
            if (me.normals) vN[vbase + v] = make_float4(me.normals[3 * v], me.normals[3 * v + 1], me.normals[3 * v + 2], 0);
            if (me.uvs) vUV[vbase + v] = make_float2(me.uvs[2 * v], me.uvs[2 * v + 1]);
            if (flags & TRI_HAS_TANGENTS) {
                vT[vbase + v] = make_float4(me.tangents[3 * v], me.tangents[3 * v + 1], me.tangents[3 * v + 2], 0);
                vB[vbase + v] = make_float4(me.bitangents[3 * v], me.bitangents[3 * v + 1], me.bitangents[3 * v + 2], 0);
            }
        }
        for (int t = 0; t < me.numTriangles; t++) {
            const int i0 = me.indices[3 * t], i1 = me.indices[3 * t + 1], i2 = me.indices[3 * t + 2];
            const float3 p0 = h3(me.vertices + 3 * i0), p1 = h3(me.vertices + 3 * i1), p2 = h3(me.vertices + 3 * i2);
            const float3 a = p1 - p0, b = p2 - p0;
            const float area = length(cross(a, b));
            if (!(area > 0.0f) || std::isinf(area)) continue;
            const float3 e0 = p1 - p0, e1 = p0 - p2;
            const float3 n = cross(e1, e0);
            int idx = (int)triShade.size();
            triRec.push_back(make_float4(p0.x, p0.y, p0.z, e0.x));
            triRec.push_back(make_float4(e0.y, e0.z, e1.x, e1.y));
            triRec.push_back(make_float4(e1.z, n.x, n.y, n.z));
            triShade.push_back(make_int4((int)vbase + i0, (int)vbase + i1, (int)vbase + i2, me.material | (flags << 24)));
            {   // per-triangle copy of the vertex normals and texture coordinates (eval_surface)
                const float4 n0 = vN[vbase + i0], n1 = vN[vbase + i1], n2 = vN[vbase + i2];
                const float2 u0 = vUV[vbase + i0], u1 = vUV[vbase + i1], u2 = vUV[vbase + i2];
                triAttr.push_back(make_float4(n0.x, n0.y, n0.z, u0.x));
                triAttr.push_back(make_float4(n1.x, n1.y, n1.z, u0.y));
                triAttr.push_back(make_float4(n2.x, n2.y, n2.z, u1.x));
                triAttr.push_back(make_float4(u1.y, u2.x, u2.y, __builtin_bit_cast(float, me.material | (flags << 24))));
            }
            addPrim(fmin3(fmin3(p0, p1), p2), fmax3(fmax3(p0, p1), p2), (idx << 2) | PRIM_TRIANGLE);
        }
        vbase += me.numVertices;
    }
    for (int i = 0; i < d->numParallelograms; i++) {
        // Cornell::createParallelogram, scene/Cornell.cpp:46-58
        const opr_parallelogram& s = d->parallelograms[i];
        const float3 anchor = h3(s.anchor), o1 = h3(s.offset1), o2 = h3(s.offset2);
        const float3 normal = normalize(cross(o1, o2));
        const float dd = dot(normal, anchor);
        const float3 v1 = o1 / dot(o1, o1), v2 = o2 / dot(o2, o2);
        const float3 tv1 = v1 / dot(v1, v1), tv2 = v2 / dot(v2, v2);     // bounds program rescales back
        const float area = length(cross(tv1, tv2));
        if (!(area > 0.0f) || std::isinf(area)) continue;
        int idx = (int)pgramMat.size();
        pgramRec.push_back(make_float4(normal.x, normal.y, normal.z, dd));
        pgramRec.push_back(make_float4(v1.x, v1.y, v1.z, 0.f));
        pgramRec.push_back(make_float4(v2.x, v2.y, v2.z, 0.f));
        pgramRec.push_back(make_float4(anchor.x, anchor.y, anchor.z, 0.f));
        pgramMat.push_back(s.material);
        const float3 p00 = anchor, p01 = anchor + tv1, p10 = anchor + tv2, p11 = anchor + tv1 + tv2;
        addPrim(fmin3(fmin3(p00, p01), fmin3(p10, p11)), fmax3(fmax3(p00, p01), fmax3(p10, p11)), (idx << 2) | PRIM_PARALLELOGRAM);
    }
    for (int i = 0; i < d->numSpheres; i++) {
        const opr_sphere& s = d->spheres[i];
        int idx = (int)sphereMat.size();
        sphereRec.push_back(make_float4(s.center[0], s.center[1], s.center[2], s.radius));
        sphereMat.push_back(s.material);
        const float3 c = h3(s.center);
        addPrim(c - f3(s.radius), c + f3(s.radius), (idx << 2) | PRIM_SPHERE);
    }
    numTriangles = (unsigned)triShade.size();
    numPrims = (unsigned)primLo.size();
    if (numPrims == 0) throw std::runtime_error("scene has no intersectable geometry");

    view.triRec = upload(triRec, owned); view.triShade = upload(triShade, owned); view.triAttr = upload(triAttr, owned);
    view.pgramRec = upload(pgramRec, owned); view.pgramMat = upload(pgramMat, owned);
    view.sphereRec = upload(sphereRec, owned); view.sphereMat = upload(sphereMat, owned);
    view.vNormal = upload(vN, owned); view.vUV = upload(vUV, owned); view.vTangent = upload(vT, owned); view.vBitangent = upload(vB, owned);
    view.materials = upload(mats, owned);
    std::vector<opr_light> lights(d->lights, d->lights + d->numLights);
    view.lights = upload(lights, owned);
    view.numLights = d->numLights;
    hostLights = lights;

    // sceneBoundingSphere: AAB::getBoundingSphere (math/AAB.cpp:31-38) through Vector3::dot, which
    // multiplies a.z*b.x (math/Vector3.cpp:30-33).  Kept: it feeds point-light emission.
    {
        float3 mn = h3(d->aabbMin), mx = h3(d->aabbMax);
        float3 center = (mn + mx) * 0.5f;
        float3 dd = mx - center;
        float dotbug = dd.x * dd.x + dd.y * dd.y + dd.z * dd.x;
        view.boundingSphere.center = center;
        view.boundingSphere.radius = sqrtf(dotbug);
        aabbMin = mn; aabbMax = mx;
    }

    buildBvh(primLo, primHi, st);
}

} // namespace opr
