// ORACLE — test infrastructure only (see ora_math.h header).  CPU restatement of the reference's
// per-iteration light-transport path.  Parity status: pinned against the reference's own device
// programs host-compiled in oracle/_ref (see oracle/refshim/README.md and tests/test_oracle_vs_ref.py);
// the reference itself ships no tests or golden vectors.
#pragma once
#include "ora_math.h"
#include "../include/opr_scene.h"
#include <vector>
#include <string>
#include <atomic>

namespace ora {

// ---- config.h:33-43 ------------------------------------------------------------------------------
static const unsigned MAX_PHOTON_TRACE_DEPTH_ = 7;
static const unsigned MAX_RADIANCE_TRACE_DEPTH_ = 9;
static const unsigned PHOTON_TRACING_RR_START_DEPTH_ = 3;
static const float RAY_LEN_MIN_ = 0.0001f;
static const float EPS_COSINE_ = 1e-6f;
static const float EPS_RAY_ = 1e-3f;
static const float EPS_PHONG_ = 1e-3f;   // BxDF.h:266
static const unsigned MAX_PHOTON_DEPOSITS = 4;   // config.h:26
static const unsigned PHOTON_GRID_MAX_SIZE_ = 100 * 100 * 100;   // OptixRenderer.cpp:39

// ---- renderer/RadiancePRD.h:31-36 -----------------------------------------------------------------
static const uint PRD_HIT_EMITTER = 1u << 31;
static const uint PRD_ERROR = 1u << 30;
static const uint PRD_MISS = 1u << 29;
static const uint PRD_HIT_SPECULAR = 1u << 28;
static const uint PRD_HIT_NON_SPECULAR = 1u << 27;
static const uint PRD_PATH_TRACING = 1u << 26;

// ---- curand XORWOW (CUDA curand_kernel.h: _curand_init_inplace / curand(); identical in CUDA 5.5 and
// 12.9) — helpers/random.h:59-69 ----------------------------------------------------------------------
struct RandomState { uint d; uint v[5]; };
void rng_init(RandomState* s, unsigned seed, unsigned index);   // curand_init(seed+index,0,0)
uint rng_next_u32(RandomState* s);                              // curand()
float getRandomUniformFloat(RandomState* s);                    // max(curand_uniform - FLT_EPSILON, 0)
inline float2 getRandomUniformFloat2(RandomState* s) { float2 r; r.x = getRandomUniformFloat(s); r.y = getRandomUniformFloat(s); return r; }
inline float3 getRandomUniformFloat3(RandomState* s) { float3 r; r.x = getRandomUniformFloat(s); r.y = getRandomUniformFloat(s); r.z = getRandomUniformFloat(s); return r; }

// ---- PODs ------------------------------------------------------------------------------------------
struct Photon { float3 power, position, rayDirection; };                       // ppm/Photon.h:10-33 (36 B)
struct Hitpoint { float3 position, normal, attenuation, radiance; uint flags; float3 volumetricRadiance; }; // Hitpoint.h:10-18 (64 B)
struct Sphere { float3 center; float radius; };                                 // math/Sphere.h
typedef opr_light Light;
typedef opr_camera Camera;
inline float3 f3(const float* p) { return {p[0], p[1], p[2]}; }

// ---- scene ----------------------------------------------------------------------------------------
enum PrimType { PRIM_TRIANGLE = 0, PRIM_PARALLELOGRAM = 1, PRIM_SPHERE = 2 };

struct Mesh {
    std::vector<float3> vertices, normals, tangents, bitangents;
    std::vector<float2> uvs;
    std::vector<int3> indices;
    int material;
};
struct Parallelogram { float4 plane; float3 v1, v2, anchor; int material; };   // parallelogram.cu:35-39, Cornell.cpp:46-58
struct SpherePrim { float3 center; float radius; int material; };
struct Texture { int width, height; std::vector<uint8_t> rgba; };

struct Prim { int type; int a; int b; };   // triangle: a=mesh, b=tri; pgram: a=index; sphere: a=index

struct Aabb { float3 lo, hi; };

struct Hit {
    float t; int prim;
    float3 geometricNormal, shadingNormal, tangent, bitangent;
    float2 textureCoordinate;
    int material;
    bool hasTangents;   // TriangleMesh.cu:61 hasTangentsAndBitangents of the mesh that was hit
};

struct BvhNode { Aabb box; int left, right; int first, count; };   // leaf if count > 0

struct Scene {
    std::string name;
    std::vector<Mesh> meshes;
    std::vector<Parallelogram> pgrams;
    std::vector<SpherePrim> spheres;
    std::vector<opr_material> materials;
    std::vector<Texture> textures;
    std::vector<Light> lights;
    float3 aabbMin, aabbMax;
    Sphere boundingSphere;
    std::vector<Prim> prims;       // BVH leaf order
    std::vector<BvhNode> nodes;

    void load(const opr_scene_desc* d);
    void buildBvh();
    bool intersectPrim(int primIndex, const Ray& ray, float tmax, Hit& hit) const;
    bool closestHit(const Ray& ray, Hit& hit) const;
    bool anyHit(const Ray& ray) const;

    // Generic front-to-back candidate enumeration: leaf(primIndex, tmax) may shrink tmax and returns true to stop.
    // Used by oracle/refshim to call the reference's own intersection programs at the leaves.
    template <class Leaf> void forEachCandidate(const Ray& ray, float& tmax, Leaf&& leaf) const
    {
        if (nodes.empty()) return;
        const float3 inv = {1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z};
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const BvhNode& nd = nodes[stack[--sp]];
            if (!slabTest(nd.box, ray.origin, inv, ray.tmin, tmax)) continue;
            if (nd.count > 0) { for (int i = nd.first; i < nd.first + nd.count; i++) if (leaf(i, tmax)) return; }
            else { stack[sp++] = nd.right; stack[sp++] = nd.left; }
        }
    }
    static inline bool slabTest(const Aabb& b, const float3& o, const float3& inv, float tmin, float tmax)
    {
        float t0 = (b.lo.x - o.x) * inv.x, t1 = (b.hi.x - o.x) * inv.x;
        float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
        t0 = (b.lo.y - o.y) * inv.y; t1 = (b.hi.y - o.y) * inv.y;
        lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
        t0 = (b.lo.z - o.z) * inv.z; t1 = (b.hi.z - o.z) * inv.z;
        lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
        lo = lo - fabsf(lo) * 1e-4f - 1e-6f;   // conservative: never cull a primitive lying exactly on a box face
        hi = hi + fabsf(hi) * 1e-4f + 1e-6f;
        return lo <= hi && hi >= tmin && lo <= tmax;
    }
};

// host helpers restated from the reference's scene/camera/light classes
void camera_setup(Camera& c);                                               // Camera.cpp:333-345
Camera make_camera(float3 eye, float3 lookat, float3 up, float hfov, float vfov, float aperture, int arm);
void camera_set_aspect_ratio(Camera& c, float ratio);                       // Camera.cpp:294-318
Light make_area_light(float3 power, float3 position, float3 v1, float3 v2); // Light.cpp:14-30
Light make_point_light(float3 power, float3 position);                      // Light.cpp:32-40
Sphere aabb_bounding_sphere(float3 mn, float3 mx);                          // AAB.cpp:31-38 + Vector3.cpp:30-33 (dot bug)
float initial_ppm_radius(float3 mn, float3 mx);                             // IScene.cpp:23-31

// ---- photon grid (OptixRenderer_SpatialHash.cu:209-282) ---------------------------------------------
struct PhotonGrid {
    float cellSize; uint3 gridSize; float3 origo; uint numCells; uint numValid;
};
// Sorts `photons` in place (stable, by key), fills keys (sorted), offsets[numCells+1].
// Returns false when numCells > PHOTON_GRID_MAX_SIZE (reference throws, :244-248).
bool build_photon_grid(Photon* photons, uint n, std::vector<uint>& keys, std::vector<uint>& offsets, PhotonGrid& g);

// ---- ray statistics for the CPU baseline -----------------------------------------------------------
struct RayStats { std::atomic<unsigned long long> closest{0}, any{0}; };

} // namespace ora
