// scene.cuh — device-resident scene, BVH2 node layout and the closest-/any-hit traversal that replace
// OptiX's acceleration structure + rtTrace.
//
// HBM layout
//   nodes      8 x float4 per 4-wide node (128 B = one line): centre.x[4] half.x[4] centre.y[4] half.y[4] centre.z[4] half.z[4] refs[4] spare
//              child ref >= 0: node index; < 0: leaf, ~ref = (firstSlot << 4) | ((count-1) << 1) | mixed; empty: box +inf
//   leafTri    4 x float4 per leaf slot (48-byte triangle record padded to 64 B; leaf = contiguous run of <= 8 slots)
//   leafPrim   int per leaf slot: (primIndex << 2) | primType, read once per ray to name the hit
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
    const float4* nodes; int rootRef;                 // 4-wide BVH (bvh.cu), 8 float4 per node
    const float4* qnodes; float3 qOrigin, qExtent;   // the same tree with child boxes on a 15-bit scene grid, 4 float4 (64 B) per node
    const float4* leafTri; const int* leafPrim;       // leaf-slot order: 3 float4 per slot; (primIndex << 2) | primType per slot
    const float4* triRec; const int4* triShade;       // triShade: vertex indices (tangent frames only) + material | flags << 24
    const float4* triAttr;                            // 64 B per triangle: n0, uv0.x | n1, uv0.y | n2, uv1.x | uv1.y, uv2.x, uv2.y, material|flags
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
    if (type == PRIM_TRIANGLE) return __float_as_int(__ldg(&sc.triAttr[4 * (size_t)idx + 3].w)) & 0xffffff;
    if (type == PRIM_PARALLELOGRAM) return __ldg(&sc.pgramMat[idx]);
    return __ldg(&sc.sphereMat[idx]);
}

// Triangle test on one 48-byte record (the operands of OptiX intersect_triangle_branchless, TriangleMesh.cu:46)
OPR_DEV bool intersect_tri_rec(const float4& a, const float4& b, const float4& c, const float3& o, const float3& d, float tmin, float tcur,
                               float& tOut, float& uOut, float& vOut)
{
    const float3 p0 = f3(a.x, a.y, a.z), e0 = f3(a.w, b.x, b.y), e1 = f3(b.z, b.w, c.x), n = f3(c.y, c.z, c.w);
    const float3 e2 = (1.0f / dot(n, d)) * (p0 - o);
    const float3 i = cross(d, e2);
    const float beta = dot(i, e1);
    const float gamma = dot(i, e0);
    const float t = dot(n, e2);
    if ((t < tcur) & (t > tmin) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1)) { tOut = t; uOut = beta; vOut = gamma; return true; }
    return false;
}

// Leaf test.  Accepts tmin < t < tcur (rtPotentialIntersection).  Returns true and updates hit on success.
OPR_DEV bool intersect_prim(const SceneView& sc, int prim, const float3& o, const float3& d, float tmin, float tcur, HitRec& hit)
{
    const int type = prim & 3, idx = prim >> 2;
    if (type == PRIM_TRIANGLE) {
        const float4 a = __ldg(&sc.triRec[3 * idx]), b = __ldg(&sc.triRec[3 * idx + 1]), c = __ldg(&sc.triRec[3 * idx + 2]);
        float t, u, v;
        if (intersect_tri_rec(a, b, c, o, d, tmin, tcur, t, u, v)) { hit.t = t; hit.prim = prim; hit.u = u; hit.v = v; return true; }
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

// 256-bit read-only global load (LDG.E.ENL2.256 on sm_100a): a lane's 128-byte node is 4 L1 wavefronts instead of 8, and the
// L1 data pipe — not DRAM, not issue — is what bounds incoherent traversal (profiles/r1d: l1tex data-pipe wavefronts 72 %).
struct __align__(32) Float8 { float4 lo, hi; };
OPR_DEV Float8 ldg256(const float4* p)
{
    Float8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
}

constexpr int LEAF_TRI_STRIDE = 4;                 // float4 per leaf slot: 48-byte triangle record padded to 64 B = two 256-bit loads
constexpr int TRAVERSAL_DONE = (int)0x80000000;   // negative like a leaf ref, never a real one (bvh.cu caps the slot count)
constexpr int TRAVERSAL_STACK = 96;    // entries; bvh.cu refuses trees with 3*levels + 4 > TRAVERSAL_STACK

// ---- traversal of the 4-wide BVH (bvh.cu) -----------------------------------------------------------------------------
// One node visit = one 128-byte line: six float4 of child boxes and one int4 of child refs.  Hit children are visited front
// to back (5-comparator sorting network); the far ones are pushed with their entry distance so that a pop can discard
// subtrees already behind the closest hit without fetching their node.
//
// The loops are warp-lockstep ("while-while" with votes): all lanes of `mask` run the node loop until every lane holds a
// leaf or is finished, then the lanes that hold a leaf test it together.  Without the votes the leaf code ran with 2 of 32
// lanes active (profiles/r1d): lanes reached their leaves at different iterations and each paid the leaf code alone.
//
// The steps are macros over the per-lane variables (o, d, idir, ood, tmin, hit, stack, sp, cur) because the persistent
// wavefront kernel in trace.cu runs the same steps around its own ray fetch.

// Slab test of one child: entry distance, or +inf when the ray misses the box within (tmin, hit.t).  A child box is stored as
// centre c and half-extent h per axis (bvh.cu): tc = c * idir + ood, t_near = tc - h |idir|, t_far = tc + h |idir| — three FFMAs per
// axis on the FP32 pipe, no per-axis min / max on the half-rate ALU pipe that bounds these kernels (|x|, -x are operand modifiers).
// Fused form c * idir + ood with ood = -o * idir (one FFMA instead of FADD + FMUL).  Its rounding error is
// <= 2^-24 |o| |idir| in t, i.e. 6e-8 * scale in space, against a box inflation of 4e-6 * scale (scene.cu addPrim): the
// conservatively inflated boxes still never cull a primitive the exact triangle test would accept.
#define OPR_SLAB(QN, CX, HX, CY, HY, CZ, HZ, TOUT)                                                                 \
    {                                                                                                              \
        float nx, fx, ny, fy, nz, fz;                                                                              \
        if (QN) {   /* decoded values are 1 + q / 32768: tc carries the +-|idir| of the half-extent's offset */     \
            const float tx = fmaf((CX), idir.x, ood.x), ty = fmaf((CY), idir.y, ood.y), tz = fmaf((CZ), idir.z, ood.z); \
            nx = fmaf((HX), -fabsf(idir.x), tx + fabsf(idir.x)); fx = fmaf((HX), fabsf(idir.x), tx - fabsf(idir.x)); \
            ny = fmaf((HY), -fabsf(idir.y), ty + fabsf(idir.y)); fy = fmaf((HY), fabsf(idir.y), ty - fabsf(idir.y)); \
            nz = fmaf((HZ), -fabsf(idir.z), tz + fabsf(idir.z)); fz = fmaf((HZ), fabsf(idir.z), tz - fabsf(idir.z)); \
        } else {                                                                                                   \
            const float tx = fmaf((CX), idir.x, ood.x), ty = fmaf((CY), idir.y, ood.y), tz = fmaf((CZ), idir.z, ood.z); \
            nx = fmaf((HX), -fabsf(idir.x), tx); fx = fmaf((HX), fabsf(idir.x), tx);                               \
            ny = fmaf((HY), -fabsf(idir.y), ty); fy = fmaf((HY), fabsf(idir.y), ty);                               \
            nz = fmaf((HZ), -fabsf(idir.z), tz); fz = fmaf((HZ), fabsf(idir.z), tz);                               \
        }                                                                                                          \
        const float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));                                                    \
        const float tf = fminf(fminf(fx, fy), fminf(fz, hit.t));                                                   \
        TOUT = tn <= tf ? tn : INF_;                                                                               \
    }
// The traversal stack of a lane: entries are (subtree ref, entry distance bits).  The default keeps it in local memory
// (`int2 stack[TRAVERSAL_STACK]`); the wavefront kernels of trace.cu redefine the two accessors to keep the first entries in
// shared memory (see there).
#define OPR_STACK_ST(I, V) stack[(I)] = (V);
#define OPR_STACK_LD(DST, I) DST = stack[(I)];
#define OPR_STACK_PUSH(R, TB) { OPR_STACK_ST(sp, make_int2((R), (TB))) sp++; }
#define OPR_CSWAP(TA, RA, TB, RB) { if (TB < TA) { const float tt = TA; TA = TB; TB = tt; const int rr = RA; RA = RB; RB = rr; } }
// pop the next subtree; those that start behind the closest hit so far are dropped unfetched
#define OPR_POP(ANYHIT)                                                                                            \
    for (;;) {                                                                                                     \
        if (sp == 0) { cur = TRAVERSAL_DONE; break; }                                                              \
        int2 e_; --sp; OPR_STACK_LD(e_, sp)                                                                        \
        if ((ANYHIT) || __int_as_float(e_.y) < hit.t) { cur = e_.x; break; }                                       \
    }
// The four child boxes (centres cX cY cZ, half-extents hX hY hZ) and child refs (r0 .. r3) of node `cur`, in one of two formats chosen per launch (QN):
//   QN = false  128-byte fp32 nodes (four 256-bit loads);
//   QN = true   64-byte quantised nodes (two 256-bit loads): box centres and half-extents as 15-bit values on a scene grid, each decoded by ONE byte
//               permute into a float in [1, 2), so the slab test stays one FFMA per plane with per-ray constants folded in
//               (ray_slab_setup).  Boxes are rounded outwards by a cell (bvh.cu), so every primitive the fp32 nodes reach is reached.
// Measured (1.1 M triangles, profiles/): incoherent rays (photon bounces, tails) 11 % faster on the quantised nodes — half the L1
// wavefronts per node visit — coherent rays (primary, shadow, specular chains) 11-17 % slower (the decode costs issue slots that the
// coherent kernels do not have to spare).  Hence per launch, not per build.
#define OPR_QLO(W) __uint_as_float(__byte_perm(__float_as_uint(W), 0x3F800000u, 0x7104u))
#define OPR_QHI(W) __uint_as_float(__byte_perm(__float_as_uint(W), 0x3F800000u, 0x7324u))
#define OPR_Q4(A, B) make_float4(OPR_QLO(A), OPR_QHI(A), OPR_QLO(B), OPR_QHI(B))
template <bool QN>
OPR_DEV void load_node(const SceneView& sc, int cur, float4& cX, float4& hX, float4& cY, float4& hY, float4& cZ, float4& hZ, int& r0, int& r1, int& r2, int& r3)
{
    float4 rf;
    if (QN) {
        const float4* N_ = sc.qnodes + 4 * (size_t)cur;
        const Float8 na_ = ldg256(N_), nb_ = ldg256(N_ + 2);
        cX = OPR_Q4(na_.lo.x, na_.lo.y); hX = OPR_Q4(na_.lo.z, na_.lo.w);
        cY = OPR_Q4(na_.hi.x, na_.hi.y); hY = OPR_Q4(na_.hi.z, na_.hi.w);
        cZ = OPR_Q4(nb_.lo.x, nb_.lo.y); hZ = OPR_Q4(nb_.lo.z, nb_.lo.w); rf = nb_.hi;
    } else {
        const float4* N_ = sc.nodes + 8 * (size_t)cur;
        const Float8 nx_ = ldg256(N_), ny_ = ldg256(N_ + 2), nz_ = ldg256(N_ + 4), nr_ = ldg256(N_ + 6);
        cX = nx_.lo; hX = nx_.hi; cY = ny_.lo; hY = ny_.hi; cZ = nz_.lo; hZ = nz_.hi; rf = nr_.lo;
    }
    r0 = __float_as_int(rf.x); r1 = __float_as_int(rf.y); r2 = __float_as_int(rf.z); r3 = __float_as_int(rf.w);
}
#define OPR_NODE_LOAD(QN)                                                                                          \
        float4 cX, hX, cY, hY, cZ, hZ; int r0, r1, r2, r3;                                                   \
        load_node<QN>(sc, cur, cX, hX, cY, hY, cZ, hZ, r0, r1, r2, r3);
// a quantised absent child has no box on the grid that every ray misses: recognise it by its ref (child 0 always exists)
#define OPR_NODE_EMPTY(QN)                                                                                         \
        if (QN) {                                                                                                  \
            if (r1 == 0x7fffffff) t1 = INF_;                                                                       \
            if (r2 == 0x7fffffff) t2 = INF_;                                                                       \
            if (r3 == 0x7fffffff) t3 = INF_;                                                                       \
        }
// Order of the hit children.  Full: 5-comparator network, visit and pop strictly front to back.  OPR_SORT_NEAREST: only the
// nearest child is found (3 comparators); the others are pushed as they come, each with its entry distance, so a pop still
// drops the ones behind the closest hit.
#ifndef OPR_SORT_NEAREST
#define OPR_SORT_NEAREST 0
#endif
#if OPR_SORT_NEAREST
#define OPR_SORT_CHILDREN OPR_CSWAP(t0, r0, t1, r1) OPR_CSWAP(t0, r0, t2, r2) OPR_CSWAP(t0, r0, t3, r3)
#else
#define OPR_SORT_CHILDREN OPR_CSWAP(t0, r0, t1, r1) OPR_CSWAP(t2, r2, t3, r3) OPR_CSWAP(t0, r0, t2, r2) OPR_CSWAP(t1, r1, t3, r3) OPR_CSWAP(t1, r1, t2, r2)
#endif
// one node visit of a lane with cur >= 0
#define OPR_NODE_STEP(ANYHIT, QN)                                                                                  \
    {                                                                                                              \
        const float INF_ = __int_as_float(0x7f800000);                                                             \
        OPR_NODE_LOAD(QN)                                                                                          \
        float t0, t1, t2, t3;                                                                                      \
        OPR_SLAB(QN, cX.x, hX.x, cY.x, hY.x, cZ.x, hZ.x, t0)                                                 \
        OPR_SLAB(QN, cX.y, hX.y, cY.y, hY.y, cZ.y, hZ.y, t1)                                                 \
        OPR_SLAB(QN, cX.z, hX.z, cY.z, hY.z, cZ.z, hZ.z, t2)                                                 \
        OPR_SLAB(QN, cX.w, hX.w, cY.w, hY.w, cZ.w, hZ.w, t3)                                                 \
        OPR_NODE_EMPTY(QN)                                                                                         \
        if (ANYHIT) {          /* any order will do for an occlusion query: skip the sorting network */             \
            if (t1 < INF_) { if (t0 < INF_) OPR_STACK_PUSH(r1, 0) else { t0 = t1; r0 = r1; } }                         \
            if (t2 < INF_) { if (t0 < INF_) OPR_STACK_PUSH(r2, 0) else { t0 = t2; r0 = r2; } }                         \
            if (t3 < INF_) { if (t0 < INF_) OPR_STACK_PUSH(r3, 0) else { t0 = t3; r0 = r3; } }                         \
            if (t0 < INF_) cur = r0; else OPR_POP(true)                                                            \
        } else {                                                                                                   \
            OPR_SORT_CHILDREN                                                                                      \
            if (t0 < INF_) {                                                                                       \
                if (t3 < INF_) OPR_STACK_PUSH(r3, __float_as_int(t3))                                              \
                if (t2 < INF_) OPR_STACK_PUSH(r2, __float_as_int(t2))                                              \
                if (t1 < INF_) OPR_STACK_PUSH(r1, __float_as_int(t1))                                              \
                cur = r0;                                                                                          \
            } else {                                                                                               \
                OPR_POP(false)                                                                                     \
            }                                                                                                      \
        }                                                                                                          \
    }
// one leaf visit of a lane with cur < 0, cur != TRAVERSAL_DONE; FOUND is set when a primitive was accepted
#define OPR_LEAF_STEP(FOUND)                                                                                       \
    {                                                                                                              \
        const unsigned L_ = (unsigned)~cur;                                                                        \
        const unsigned first_ = L_ >> 4, cnt_ = ((L_ >> 1) & 7u) + 1u;                                             \
        if (!(L_ & 1u)) {      /* all triangles: one contiguous run of 48-byte records */                          \
            for (unsigned k_ = 0; k_ < cnt_; k_++) {                                                               \
                const float4* R_ = sc.leafTri + LEAF_TRI_STRIDE * (size_t)(first_ + k_);                           \
                const Float8 ab_ = ldg256(R_), cd_ = ldg256(R_ + 2);                                               \
                const float4 a_ = ab_.lo, b_ = ab_.hi, c_ = cd_.lo;                                                \
                if (intersect_tri_rec(a_, b_, c_, o, d, tmin, hit.t, hit.t, hit.u, hit.v)) { hit.prim = (int)(first_ + k_); FOUND = true; } \
            }                                                                                                      \
        } else {               /* analytic primitives of the built-in scenes (parallelogram.cu, Sphere.cu) mixed in */ \
            for (unsigned k_ = 0; k_ < cnt_; k_++) {                                                               \
                const int prim_ = __ldg(&sc.leafPrim[first_ + k_]);                                                \
                if ((prim_ & 3) == PRIM_TRIANGLE) {                                                                \
                    const float4* R_ = sc.leafTri + LEAF_TRI_STRIDE * (size_t)(first_ + k_);                       \
                    const Float8 ab_ = ldg256(R_), cd_ = ldg256(R_ + 2);                                           \
                    const float4 a_ = ab_.lo, b_ = ab_.hi, c_ = cd_.lo;                                            \
                    if (intersect_tri_rec(a_, b_, c_, o, d, tmin, hit.t, hit.t, hit.u, hit.v)) { hit.prim = (int)(first_ + k_); FOUND = true; } \
                } else if (intersect_prim(sc, prim_, o, d, tmin, hit.t, hit)) { hit.prim = (int)(first_ + k_); FOUND = true; } \
            }                                                                                                      \
        }                                                                                                          \
    }

// Node phase of the warp-lockstep loops: lanes that hold an inner node step together until the lanes that wait with a leaf
// outnumber them (weights OPR_VOTE_NODE : OPR_VOTE_LEAF); then the leaf phase serves those, while the lanes that still hold a
// node keep it in `cur` and resume afterwards.  Every lane visits exactly the nodes and leaves it would visit alone, in the
// same order.  With OPR_VOTE_LEAF = 0 the phase runs until no lane holds a node (the classic while-while loop, whose node
// phase lasts as long as the slowest lane's descent).  Measured: traces 5 % faster with 1:1, 2:1, 1:2 or 3:1 alike.
#ifndef OPR_VOTE_NODE
#define OPR_VOTE_NODE 2
#endif
#ifndef OPR_VOTE_LEAF
#define OPR_VOTE_LEAF 1
#endif
#define OPR_NODE_PHASE(MASK, ANYHIT, QN)                                                                           \
    for (;;) {                                                                                                     \
        const unsigned nm_ = __ballot_sync((MASK), cur >= 0);                                                      \
        if (!nm_) break;                                                                                           \
        if (OPR_VOTE_LEAF) {                                                                                       \
            const unsigned lm_ = __ballot_sync((MASK), cur < 0 && cur != TRAVERSAL_DONE);                          \
            if (__popc(nm_) * OPR_VOTE_NODE < __popc(lm_) * OPR_VOTE_LEAF) break;                                  \
        }                                                                                                          \
        if (cur >= 0) OPR_NODE_STEP(ANYHIT, QN)                                                                    \
    }

// 1/d with |d| clamped away from zero: keeps every slab product finite-or-inf (never 0 * inf = NaN)
OPR_DEV float3 safe_inverse_direction(const float3& d)
{
    return f3(1.0f / (fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x)), 1.0f / (fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y)),
              1.0f / (fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z)));
}

// Per-ray constants of the slab test.  Full-precision nodes: t = plane * idir + ood.  Quantised nodes (QN): a plane is
// stored as q in [0, 32767] on the scene grid, decoded by ONE byte permute into the float v = 1 + q / 32768 (the stored
// half-word carries the exponent's low bit), and world = qOrigin + (v - 1) * qExtent, so t = v * (qExtent * idir) +
// (qOrigin - qExtent - o) * idir: the same single FFMA per plane with different per-ray constants.
template <bool QN>
OPR_DEV void ray_slab_setup(const SceneView& sc, const float3& o, const float3& d, float3& idir, float3& ood)
{
    const float3 inv = safe_inverse_direction(d);
    if (QN) {
        idir = f3(sc.qExtent.x * inv.x, sc.qExtent.y * inv.y, sc.qExtent.z * inv.z);
        ood = f3((sc.qOrigin.x - sc.qExtent.x - o.x) * inv.x, (sc.qOrigin.y - sc.qExtent.y - o.y) * inv.y, (sc.qOrigin.z - sc.qExtent.z - o.z) * inv.z);
    } else {
        idir = inv;
        ood = f3(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    }
}

// Closest hit (ANY = false) or first accepted intersection (ANY = true: shadow rays, ppm/DirectRadianceEstimation.cu:79-83,
// bound to every material by material/Material.cpp:18-26) for the lanes of `mask`, which must all make this call together;
// lanes with active = false ride along.  On return hit.prim = (primIndex << 2) | primType, or -1.
template <bool ANY, bool QN = false>
OPR_DEV bool traverse_warp(const SceneView& sc, unsigned mask, bool active, const float3& o, const float3& d, float tmin, float tmax, HitRec& hit)
{
    hit.t = tmax; hit.prim = -1; hit.u = hit.v = 0.f;
    float3 idir, ood;
    ray_slab_setup<QN>(sc, o, d, idir, ood);
    int2 stack[TRAVERSAL_STACK];
    int sp = 0;
    int cur = active ? sc.rootRef : TRAVERSAL_DONE;
    for (;;) {
        OPR_NODE_PHASE(mask, ANY, QN)
        if (__all_sync(mask, cur == TRAVERSAL_DONE)) break;
        if (cur < 0 && cur != TRAVERSAL_DONE) {
            bool found = false;
            OPR_LEAF_STEP(found)
            if (ANY && found) { cur = TRAVERSAL_DONE; sp = 0; }
            else OPR_POP(ANY)
        }
    }
    if (hit.prim < 0) return false;
    if (!ANY) hit.prim = __ldg(&sc.leafPrim[hit.prim]);     // leaf slot -> (primIndex << 2) | primType
    return true;
}

// No votes: every lane runs at its own pace (for latency-bound kernels with a handful of rays per warp).
template <bool ANY, bool QN = false>
OPR_DEV bool traverse_free(const SceneView& sc, const float3& o, const float3& d, float tmin, float tmax, HitRec& hit)
{
    hit.t = tmax; hit.prim = -1; hit.u = hit.v = 0.f;
    float3 idir, ood;
    ray_slab_setup<QN>(sc, o, d, idir, ood);
    int2 stack[TRAVERSAL_STACK];
    int sp = 0;
    int cur = sc.rootRef;
    for (;;) {
        while (cur >= 0) OPR_NODE_STEP(ANY, QN)
        if (cur == TRAVERSAL_DONE) break;
        bool found = false;
        OPR_LEAF_STEP(found)
        if (ANY && found) return true;
        OPR_POP(ANY)
    }
    if (hit.prim < 0) return false;
    if (!ANY) hit.prim = __ldg(&sc.leafPrim[hit.prim]);
    return true;
}

// Call from possibly divergent code: the lanes that arrive together traverse together.
template <bool ANY>
OPR_DEV bool traverse(const SceneView& sc, const float3& o, const float3& d, float tmin, float tmax, HitRec& hit)
{
    return traverse_warp<ANY>(sc, __activemask(), true, o, d, tmin, tmax, hit);
}

// Attributes the reference's intersection programs hand to the material programs
struct Surface { float3 geometricNormal, shadingNormal, tangent, bitangent; float2 uv; int material; bool hasTangents; };

OPR_DEV Surface eval_surface(const SceneView& sc, const HitRec& hit, const float3& o, const float3& d)
{
    Surface s;
    s.tangent = s.bitangent = f3(0.f); s.uv = make_float2(0.f, 0.f); s.hasTangents = false;
    const int type = hit.prim & 3, idx = hit.prim >> 2;
    if (type == PRIM_TRIANGLE) {                       // TriangleMesh.cu:48-88
        const float4 c = __ldg(&sc.triRec[3 * idx + 2]);
        const float3 n = f3(c.y, c.z, c.w);
        // the three vertex normals and texture coordinates of the triangle sit in one 64-byte record (two 256-bit loads) instead
        // of behind the vertex indices: one dependent load less per hit and two sectors instead of seven
        const float4* A = sc.triAttr + 4 * (size_t)idx;
        const Float8 a01 = ldg256(A), a23 = ldg256(A + 2);
        const int mf = __float_as_int(a23.hi.w);
        const int flags = mf >> 24;
        s.material = mf & 0xffffff;
        const float beta = hit.u, gamma = hit.v;
        if (flags & TRI_HAS_NORMALS) {
            const float3 n0 = f3(a01.lo), n1 = f3(a01.hi), n2 = f3(a23.lo);
            s.shadingNormal = normalize(n1 * beta + n2 * gamma + n0 * (1.0f - beta - gamma));
            if (flags & TRI_HAS_TANGENTS) {
                s.hasTangents = true;
                const int4 sh = __ldg(&sc.triShade[idx]);
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
            const float2 t0 = make_float2(a01.lo.w, a01.hi.w), t1 = make_float2(a23.lo.w, a23.hi.x), t2 = make_float2(a23.hi.y, a23.hi.z);
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
