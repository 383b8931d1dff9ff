// scene.cuh — device-resident scene, BVH2 node layout and the closest-/any-hit traversal that replace
// OptiX's acceleration structure + rtTrace.
//
// HBM layout
//   nodes      4 x float4 per internal node (64 B, one 16-byte load each):
//                n0xy = (L.lo.x, L.hi.x, L.lo.y, L.hi.y)   n1xy = same for the right child
//                nz   = (L.lo.z, L.hi.z, R.lo.z, R.hi.z)   ch   = (leftRef, rightRef, -, -) as int bits
//              child ref >= 0: internal node index; < 0: leaf, ~ref = (primIndex << 2) | primType
//   triRec     3 x float4 per triangle (48 B): (p0.xyz, e0.x) (e0.yz, e1.xy) (e1.z, n.xyz) with
//              e0 = p1-p0, e1 = p0-p2, n = cross(e1,e0) — the operands of OptiX intersect_triangle_branchless
//              (called from geometry_instance/TriangleMesh.cu:46), precomputed with the same fp32 ops
//   triShade   int4 per triangle: global vertex indices + (material | flags << 24), read only when shading
//   pgramRec   4 x float4 per parallelogram (geometry_instance/parallelogram.cu:35-39)
//   sphereRec  float4 (center, radius) per sphere (geometry_instance/Sphere.cu:26-27)
#pragma once
#include "common.cuh"
#include "../../include/opr_scene.h"

namespace opr {

enum { PRIM_TRIANGLE = 0, PRIM_PARALLELOGRAM = 1, PRIM_SPHERE = 2 };
enum { TRI_HAS_NORMALS = 1, TRI_HAS_UVS = 2, TRI_HAS_TANGENTS = 4 };

struct DMaterial {
    int kind;
    float3 Kd, Ks; float exponent;
    float3 Kr, Kt; float ior;
    float3 powerPerArea;     // DiffuseEmitter.cpp:52-58
    float3 Lemit;            // DiffuseEmitter.cpp:61
    float inverseArea;
    int hasNormalTex;
    cudaTextureObject_t diffuseTex, normalTex;
};

struct DSphere { float3 center; float radius; };

struct SceneView {
    const float4* nodes; int rootRef;
    const float4* triRec; const int4* triShade;
    const float4* pgramRec; const int* pgramMat;
    const float4* sphereRec; const int* sphereMat;
    const float4* vNormal; const float2* vUV; const float4* vTangent; const float4* vBitangent;
    const DMaterial* materials;
    const opr_light* lights; int numLights;
    DSphere boundingSphere;
};

struct HitRec { float t; int prim; float u, v; };   // prim = (index << 2) | type, or -1 on miss

OPR_DEV int prim_material(const SceneView& sc, int prim)
{
    int type = prim & 3, idx = prim >> 2;
    if (type == PRIM_TRIANGLE) return __ldg(&sc.triShade[idx].w) & 0xffffff;
    if (type == PRIM_PARALLELOGRAM) return __ldg(&sc.pgramMat[idx]);
    return __ldg(&sc.sphereMat[idx]);
}

// Leaf test.  Accepts tmin < t < tcur (rtPotentialIntersection).  Returns true and updates hit on success.
OPR_DEV bool intersect_prim(const SceneView& sc, int prim, const float3& o, const float3& d, float tmin, float tcur, HitRec& hit)
{
    const int type = prim & 3, idx = prim >> 2;
    if (type == PRIM_TRIANGLE) {
        const float4 a = __ldg(&sc.triRec[3 * idx]), b = __ldg(&sc.triRec[3 * idx + 1]), c = __ldg(&sc.triRec[3 * idx + 2]);
        const float3 p0 = f3(a.x, a.y, a.z), e0 = f3(a.w, b.x, b.y), e1 = f3(b.z, b.w, c.x), n = f3(c.y, c.z, c.w);
        const float3 e2 = (1.0f / dot(n, d)) * (p0 - o);
        const float3 i = cross(d, e2);
        const float beta = dot(i, e1);
        const float gamma = dot(i, e0);
        const float t = dot(n, e2);
        if ((t < tcur) & (t > tmin) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1)) {
            hit.t = t; hit.prim = prim; hit.u = beta; hit.v = gamma;
            return true;
        }
        return false;
    }
    if (type == PRIM_PARALLELOGRAM) {
        const float4 plane = __ldg(&sc.pgramRec[4 * idx]), r1 = __ldg(&sc.pgramRec[4 * idx + 1]), r2 = __ldg(&sc.pgramRec[4 * idx + 2]), r3 = __ldg(&sc.pgramRec[4 * idx + 3]);
        const float3 n = f3(plane.x, plane.y, plane.z), v1 = f3(r1.x, r1.y, r1.z), v2 = f3(r2.x, r2.y, r2.z), anchor = f3(r3.x, r3.y, r3.z);
        float dt = dot(d, n);
        float t = (plane.w - dot(n, o)) / dt;
        if (t > tmin && t < tcur) {
            float3 p = o + d * t;
            float3 vi = p - anchor;
            float a1 = dot(v1, vi);
            if (a1 >= 0 && a1 <= 1) {
                float a2 = dot(v2, vi);
                if (a2 >= 0 && a2 <= 1) { hit.t = t; hit.prim = prim; hit.u = a1; hit.v = a2; return true; }
            }
        }
        return false;
    }
    {
        const float4 s = __ldg(&sc.sphereRec[idx]);
        float3 O = o - f3(s.x, s.y, s.z);
        float b = dot(O, d);
        float c = dot(O, O) - s.w * s.w;
        float disc = b * b - c;
        if (disc > 0.0f) {
            float sdisc = sqrtf(disc);
            float root1 = (-b - sdisc);
            if (root1 > tmin && root1 < tcur) { hit.t = root1; hit.prim = prim; hit.u = hit.v = 0.f; return true; }
            float root2 = (-b + sdisc);
            if (root2 > tmin && root2 < tcur) { hit.t = root2; hit.prim = prim; hit.u = hit.v = 0.f; return true; }
        }
        return false;
    }
}

constexpr int TRAVERSAL_STACK = 64;

// While-while BVH2 traversal.  ANY = true returns at the first accepted intersection (shadow rays:
// ppm/DirectRadianceEstimation.cu:79-83 bound to every material by material/Material.cpp:18-26).
template <bool ANY>
OPR_DEV bool traverse(const SceneView& sc, const float3& o, const float3& d, float tmin, float tmax, HitRec& hit)
{
    hit.t = tmax; hit.prim = -1; hit.u = hit.v = 0.f;
    const float3 idir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[TRAVERSAL_STACK];
    int sp = 0;
    int cur = sc.rootRef;
    for (;;) {
        while (cur >= 0) {
            const float4 n0 = __ldg(&sc.nodes[4 * cur]), n1 = __ldg(&sc.nodes[4 * cur + 1]), nz = __ldg(&sc.nodes[4 * cur + 2]);
            const float4 chf = __ldg(&sc.nodes[4 * cur + 3]);
            const int left = __float_as_int(chf.x), right = __float_as_int(chf.y);
            // slab tests; (plane - o) * idir keeps +-inf (not NaN) for axis-parallel rays
            float lx0 = (n0.x - o.x) * idir.x, lx1 = (n0.y - o.x) * idir.x;
            float ly0 = (n0.z - o.y) * idir.y, ly1 = (n0.w - o.y) * idir.y;
            float lz0 = (nz.x - o.z) * idir.z, lz1 = (nz.y - o.z) * idir.z;
            float lmin = fmaxf(fmaxf(fminf(lx0, lx1), fminf(ly0, ly1)), fmaxf(fminf(lz0, lz1), tmin));
            float lmax = fminf(fminf(fmaxf(lx0, lx1), fmaxf(ly0, ly1)), fminf(fmaxf(lz0, lz1), hit.t));
            float rx0 = (n1.x - o.x) * idir.x, rx1 = (n1.y - o.x) * idir.x;
            float ry0 = (n1.z - o.y) * idir.y, ry1 = (n1.w - o.y) * idir.y;
            float rz0 = (nz.z - o.z) * idir.z, rz1 = (nz.w - o.z) * idir.z;
            float rmin = fmaxf(fmaxf(fminf(rx0, rx1), fminf(ry0, ry1)), fmaxf(fminf(rz0, rz1), tmin));
            float rmax = fminf(fminf(fmaxf(rx0, rx1), fmaxf(ry0, ry1)), fminf(fmaxf(rz0, rz1), hit.t));
            const bool hl = lmin <= lmax, hr = rmin <= rmax;
            if (!hl && !hr) {
                if (sp == 0) return hit.prim != -1;
                cur = stack[--sp];
            } else {
                cur = hl ? left : right;
                if (hl && hr) {
                    int far = right;
                    if (rmin < lmin) { far = left; cur = right; }
                    stack[sp++] = far;
                }
            }
        }
        if (intersect_prim(sc, ~cur, o, d, tmin, hit.t, hit) && ANY) return true;
        if (sp == 0) return hit.prim != -1;
        cur = stack[--sp];
    }
}

// Attributes the reference's intersection programs hand to the material programs
struct Surface { float3 geometricNormal, shadingNormal, tangent, bitangent; float2 uv; int material; };

OPR_DEV Surface eval_surface(const SceneView& sc, const HitRec& hit, const float3& o, const float3& d)
{
    Surface s;
    s.tangent = s.bitangent = f3(0.f); s.uv = make_float2(0.f, 0.f);
    const int type = hit.prim & 3, idx = hit.prim >> 2;
    if (type == PRIM_TRIANGLE) {                       // TriangleMesh.cu:48-88
        const float4 c = __ldg(&sc.triRec[3 * idx + 2]);
        const float3 n = f3(c.y, c.z, c.w);
        const int4 sh = __ldg(&sc.triShade[idx]);
        const int flags = sh.w >> 24;
        s.material = sh.w & 0xffffff;
        const float beta = hit.u, gamma = hit.v;
        if (flags & TRI_HAS_NORMALS) {
            const float3 n0 = f3(__ldg(&sc.vNormal[sh.x])), n1 = f3(__ldg(&sc.vNormal[sh.y])), n2 = f3(__ldg(&sc.vNormal[sh.z]));
            s.shadingNormal = normalize(n1 * beta + n2 * gamma + n0 * (1.0f - beta - gamma));
            if (flags & TRI_HAS_TANGENTS) {
                const float3 t0 = f3(__ldg(&sc.vTangent[sh.x])), t1 = f3(__ldg(&sc.vTangent[sh.y])), t2 = f3(__ldg(&sc.vTangent[sh.z]));
                s.tangent = normalize(t1 * beta + t2 * gamma + t0 * (1.0f - beta - gamma));
                const float3 b0 = f3(__ldg(&sc.vBitangent[sh.x])), b1 = f3(__ldg(&sc.vBitangent[sh.y])), b2 = f3(__ldg(&sc.vBitangent[sh.z]));
                s.bitangent = normalize(b1 * beta + b2 * gamma + b0 * (1.0f - beta - gamma));
            }
        } else {
            s.shadingNormal = normalize(n);
        }
        s.geometricNormal = normalize(n);
        if (flags & TRI_HAS_UVS) {
            const float2 t0 = __ldg(&sc.vUV[sh.x]), t1 = __ldg(&sc.vUV[sh.y]), t2 = __ldg(&sc.vUV[sh.z]);
            const float w0 = 1.0f - beta - gamma;
            s.uv = make_float2(t1.x * beta + t2.x * gamma + t0.x * w0, t1.y * beta + t2.y * gamma + t0.y * w0);
        }
    } else if (type == PRIM_PARALLELOGRAM) {           // parallelogram.cu:63-64
        const float4 plane = __ldg(&sc.pgramRec[4 * idx]);
        s.shadingNormal = s.geometricNormal = f3(plane.x, plane.y, plane.z);
        s.material = __ldg(&sc.pgramMat[idx]);
    } else {                                           // Sphere.cu:44,51
        const float4 sp = __ldg(&sc.sphereRec[idx]);
        const float3 O = o - f3(sp.x, sp.y, sp.z);
        s.shadingNormal = s.geometricNormal = (O + hit.t * d) / sp.w;
        s.material = __ldg(&sc.sphereMat[idx]);
    }
    return s;
}

} // namespace opr
