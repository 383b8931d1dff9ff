// ORACLE — test infrastructure only (see ora_math.h).  RNG, camera/light host helpers, scene compile,
// CPU BVH and the primitive intersectors restated from geometry_instance/*.cu.
#include "ora_core.h"
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace ora {

// ---------------------------------------------------------------------------------------------------
// curand XORWOW.  curand_kernel.h `_curand_init_inplace` with subsequence=offset=0 (both skip-aheads
// are no-ops) and `curand()`; `_curand_uniform` (curand_uniform.h:69-72) is x*2^-32 + 2^-33, which
// nvcc contracts to one FMA in device code, so it is written as fmaf here and in the CUDA kernels.
// helpers/random.h:59-69.
void rng_init(RandomState* s, unsigned seed, unsigned index)
{
    unsigned long long sd = (unsigned)(seed + index);   // unsigned int sum, then widened (random.h:61)
    unsigned s0 = ((unsigned)sd) ^ 0xaad26b49u;
    unsigned s1 = (unsigned)(sd >> 32) ^ 0xf7dcefddu;
    unsigned t0 = 1099087573u * s0;
    unsigned t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
}

uint rng_next_u32(RandomState* s)
{
    unsigned t = (s->v[0] ^ (s->v[0] >> 2));
    s->v[0] = s->v[1];
    s->v[1] = s->v[2];
    s->v[2] = s->v[3];
    s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return s->v[4] + s->d;
}

float getRandomUniformFloat(RandomState* s)
{
    const float k = 2.3283064e-10f;   // CURAND_2POW32_INV
    float u = fmaf((float)rng_next_u32(s), k, k / 2.0f);
    float r = u - FLT_EPSILON;
    return r > 0.0f ? r : 0.0f;       // maxf(a,b) = a > b ? a : b (helpers.h:139-142)
}

// ---------------------------------------------------------------------------------------------------
// Camera (Camera.cpp:262-345)
static inline bool finite3(const float3& v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }
static inline float DtoR(float d) { return d * (static_cast<float>(M_PI) / 180.f); }
static inline float RtoD(float r) { return r * (180.f / static_cast<float>(M_PI)); }
static inline void st3(float* p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

void camera_setup(Camera& c)
{
    float3 lookdir = f3(c.lookdir), up = f3(c.up), cu = f3(c.camera_u), cv = f3(c.camera_v);
    float3 t = f3(c.lookat) - f3(c.eye);
    if (finite3(t)) lookdir = t;
    float lookdir_len = length(lookdir);
    t = normalize(up); if (finite3(t)) up = t;
    t = normalize(cross(lookdir, up)); if (finite3(t)) cu = t;
    t = normalize(cross(cu, lookdir)); if (finite3(t)) cv = t;
    float ulen = lookdir_len * tanf(DtoR(c.hfov * 0.5f));
    t = cu * ulen; if (finite3(t)) cu = t;
    float vlen = lookdir_len * tanf(DtoR(c.vfov * 0.5f));
    t = cv * vlen; if (finite3(t)) cv = t;
    st3(c.lookdir, lookdir); st3(c.up, up); st3(c.camera_u, cu); st3(c.camera_v, cv);
    c.imagePlaneSize[0] = 2.0f * ulen;
    c.imagePlaneSize[1] = 2.0f * vlen;
}

Camera make_camera(float3 eye, float3 lookat, float3 up, float hfov, float vfov, float aperture, int arm)
{
    Camera c; memset(&c, 0, sizeof c);
    st3(c.eye, eye); st3(c.lookat, lookat); st3(c.up, up);
    c.hfov = hfov; c.vfov = vfov; c.aperture = aperture; c.aspectRatioMode = arm;
    camera_setup(c);
    return c;
}

void camera_set_aspect_ratio(Camera& c, float ratio)
{
    float realRatio = ratio;
    const float* in = 0; float* out = 0;
    switch (c.aspectRatioMode) {
    case OPR_KEEP_HORIZONTAL: in = &c.hfov; out = &c.vfov; realRatio = 1.f / ratio; break;
    case OPR_KEEP_VERTICAL: in = &c.vfov; out = &c.hfov; break;
    default: return;
    }
    *out = RtoD(2.0f * atanf(realRatio * tanf(DtoR(0.5f * (*in)))));
    camera_setup(c);
}

// Light (Light.cpp:14-40)
Light make_area_light(float3 power, float3 position, float3 v1, float3 v2)
{
    Light l; memset(&l, 0, sizeof l);
    st3(l.power, power); st3(l.position, position); st3(l.v1, v1); st3(l.v2, v2);
    l.lightType = OPR_LIGHT_AREA;
    float3 cp = cross(v1, v2);
    st3(l.normal, normalize(cp));
    l.area_or_angle = length(cp);
    l.inverseArea = 1.0f / l.area_or_angle;
    st3(l.Lemit, power * l.inverseArea * M_1_PIf_);
    l.isDelta = 0; l.isFinite = 1;
    return l;
}

Light make_point_light(float3 power, float3 position)
{
    Light l; memset(&l, 0, sizeof l);
    st3(l.power, power); st3(l.position, position);
    l.lightType = OPR_LIGHT_POINT;
    st3(l.Lemit, power * 0.25f * M_1_PIf_);   // intensity
    l.isDelta = 1; l.isFinite = 1;
    return l;
}

// AAB::getBoundingSphere (AAB.cpp:31-38) through Vector3::length -> Vector3::dot, which multiplies
// a.z*b.x (Vector3.cpp:30-33).  Reproduced on purpose.
Sphere aabb_bounding_sphere(float3 mn, float3 mx)
{
    float3 center = (mn + mx) * 0.5f;
    float3 d = mx - center;
    float dotbug = d.x * d.x + d.y * d.y + d.z * d.x;
    Sphere s; s.center = center; s.radius = sqrtf(dotbug);
    return s;
}

float initial_ppm_radius(float3 mn, float3 mx)
{
    float3 e = mx - mn;
    float volume = e.x * e.y * e.z;
    float cubelength = powf(volume, 1.f / 3.f);      // pow(float,float) -> float overload
    float A = 6 * cubelength * cubelength;
    float radius = (float)(A * 3.94e-6);             // float*double then narrowed (IScene.cpp:29)
    return radius;
}

// ---------------------------------------------------------------------------------------------------
void Scene::load(const opr_scene_desc* d)
{
    name = d->name ? d->name : "";
    meshes.clear(); pgrams.clear(); spheres.clear(); materials.clear(); textures.clear(); lights.clear();
    for (int m = 0; m < d->numMeshes; m++) {
        const opr_mesh& s = d->meshes[m];
        Mesh o; o.material = s.material;
        o.vertices.resize(s.numVertices);
        memcpy(o.vertices.data(), s.vertices, sizeof(float3) * s.numVertices);
        if (s.normals) { o.normals.resize(s.numVertices); memcpy(o.normals.data(), s.normals, sizeof(float3) * s.numVertices); }
        if (s.uvs) { o.uvs.resize(s.numVertices); memcpy(o.uvs.data(), s.uvs, sizeof(float2) * s.numVertices); }
        if (s.tangents && s.bitangents) {
            o.tangents.resize(s.numVertices); memcpy(o.tangents.data(), s.tangents, sizeof(float3) * s.numVertices);
            o.bitangents.resize(s.numVertices); memcpy(o.bitangents.data(), s.bitangents, sizeof(float3) * s.numVertices);
        }
        o.indices.resize(s.numTriangles);
        memcpy(o.indices.data(), s.indices, sizeof(int3) * s.numTriangles);
        meshes.push_back(std::move(o));
    }
    for (int i = 0; i < d->numParallelograms; i++) {
        // Cornell::createParallelogram, scene/Cornell.cpp:46-58
        const opr_parallelogram& s = d->parallelograms[i];
        float3 anchor = f3(s.anchor), o1 = f3(s.offset1), o2 = f3(s.offset2);
        float3 normal = normalize(cross(o1, o2));
        float dd = dot(normal, anchor);
        Parallelogram p;
        p.plane = make_float4(normal, dd);
        p.v1 = o1 / dot(o1, o1);
        p.v2 = o2 / dot(o2, o2);
        p.anchor = anchor;
        p.material = s.material;
        pgrams.push_back(p);
    }
    for (int i = 0; i < d->numSpheres; i++) {
        SpherePrim s; s.center = f3(d->spheres[i].center); s.radius = d->spheres[i].radius; s.material = d->spheres[i].material;
        spheres.push_back(s);
    }
    materials.assign(d->materials, d->materials + d->numMaterials);
    for (int i = 0; i < d->numTextures; i++) {
        Texture t; t.width = d->textures[i].width; t.height = d->textures[i].height;
        t.rgba.assign(d->textures[i].rgba, d->textures[i].rgba + (size_t)4 * t.width * t.height);
        textures.push_back(std::move(t));
    }
    lights.assign(d->lights, d->lights + d->numLights);
    aabbMin = f3(d->aabbMin); aabbMax = f3(d->aabbMax);
    boundingSphere = aabb_bounding_sphere(aabbMin, aabbMax);
    buildBvh();
}

// bounds programs: TriangleMesh.cu:98-118, parallelogram.cu:78-98, Sphere.cu:58-70.
static bool primBounds(const Scene& s, const Prim& p, Aabb& box)
{
    if (p.type == PRIM_TRIANGLE) {
        const Mesh& m = s.meshes[p.a];
        const int3 idx = m.indices[p.b];
        const float3 v0 = m.vertices[idx.x], v1 = m.vertices[idx.y], v2 = m.vertices[idx.z];
        const float area = length(cross(v1 - v0, v2 - v0));
        if (area > 0.0f && !std::isinf(area)) {
            box.lo = fminf3(fminf3(v0, v1), v2);
            box.hi = fmaxf3(fmaxf3(v0, v1), v2);
            return true;
        }
        return false;
    }
    if (p.type == PRIM_PARALLELOGRAM) {
        const Parallelogram& g = s.pgrams[p.a];
        const float3 tv1 = g.v1 / dot(g.v1, g.v1);
        const float3 tv2 = g.v2 / dot(g.v2, g.v2);
        const float3 p00 = g.anchor, p01 = g.anchor + tv1, p10 = g.anchor + tv2, p11 = g.anchor + tv1 + tv2;
        const float area = length(cross(tv1, tv2));
        if (area > 0.0f && !std::isinf(area)) {
            box.lo = fminf3(fminf3(p00, p01), fminf3(p10, p11));
            box.hi = fmaxf3(fmaxf3(p00, p01), fmaxf3(p10, p11));
            return true;
        }
        return false;
    }
    const SpherePrim& sp = s.spheres[p.a];
    box.lo = sp.center - make_float3(sp.radius);
    box.hi = sp.center + make_float3(sp.radius);
    return true;
}

static inline float boxArea(const Aabb& b)
{
    float3 e = b.hi - b.lo;
    return 2.f * (e.x * e.y + e.y * e.z + e.z * e.x);
}

// The acceleration structure itself is OptiX-internal in the reference ("Trbvh"/"Bvh"/"NoAccel",
// scene/Scene.cpp:353,513,547; Cornell.cpp:183,188).  Any exact BVH returns the same closest hit, so the
// oracle uses a plain binned-SAH BVH2 on the host.
void Scene::buildBvh()
{
    prims.clear(); nodes.clear();
    std::vector<Prim> all;
    for (int m = 0; m < (int)meshes.size(); m++)
        for (int t = 0; t < (int)meshes[m].indices.size(); t++) all.push_back({PRIM_TRIANGLE, m, t});
    for (int i = 0; i < (int)pgrams.size(); i++) all.push_back({PRIM_PARALLELOGRAM, i, 0});
    for (int i = 0; i < (int)spheres.size(); i++) all.push_back({PRIM_SPHERE, i, 0});
    std::vector<Aabb> boxes; std::vector<float3> cent;
    for (auto& p : all) {
        Aabb b;
        if (!primBounds(*this, p, b)) continue;   // aabb->invalidate(): never intersected
        prims.push_back(p); boxes.push_back(b); cent.push_back((b.lo + b.hi) * 0.5f);
    }
    const int n = (int)prims.size();
    if (n == 0) return;
    std::vector<int> order(n); std::iota(order.begin(), order.end(), 0);
    nodes.reserve(2 * n);
    struct Task { int node, lo, hi; };
    std::vector<Task> stack;
    nodes.push_back(BvhNode());
    stack.push_back({0, 0, n});
    const int NB = 16;
    while (!stack.empty()) {
        Task tk = stack.back(); stack.pop_back();
        Aabb box{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
        Aabb cb = box;
        for (int i = tk.lo; i < tk.hi; i++) {
            box.lo = fminf3(box.lo, boxes[order[i]].lo); box.hi = fmaxf3(box.hi, boxes[order[i]].hi);
            cb.lo = fminf3(cb.lo, cent[order[i]]); cb.hi = fmaxf3(cb.hi, cent[order[i]]);
        }
        BvhNode nd; nd.box = box; nd.left = nd.right = -1; nd.first = tk.lo; nd.count = 0;
        int cnt = tk.hi - tk.lo;
        int mid = -1;
        if (cnt > 4) {
            float3 ce = cb.hi - cb.lo;
            int axis = (ce.x >= ce.y && ce.x >= ce.z) ? 0 : (ce.y >= ce.z ? 1 : 2);
            float lo = (&cb.lo.x)[axis], ext = (&ce.x)[axis];
            if (ext > 0.f) {
                Aabb bb[NB]; int bc[NB];
                for (int b = 0; b < NB; b++) { bb[b] = Aabb{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; bc[b] = 0; }
                auto binOf = [&](int id) { int b = (int)(((&cent[id].x)[axis] - lo) / ext * NB); return b < 0 ? 0 : (b >= NB ? NB - 1 : b); };
                for (int i = tk.lo; i < tk.hi; i++) {
                    int b = binOf(order[i]); bc[b]++;
                    bb[b].lo = fminf3(bb[b].lo, boxes[order[i]].lo); bb[b].hi = fmaxf3(bb[b].hi, boxes[order[i]].hi);
                }
                float best = FLT_MAX; int bestSplit = -1;
                Aabb rb[NB]; int rc[NB];
                Aabb acc{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; int ac = 0;
                for (int b = NB - 1; b > 0; b--) { acc.lo = fminf3(acc.lo, bb[b].lo); acc.hi = fmaxf3(acc.hi, bb[b].hi); ac += bc[b]; rb[b] = acc; rc[b] = ac; }
                acc = Aabb{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; ac = 0;
                for (int b = 0; b < NB - 1; b++) {
                    acc.lo = fminf3(acc.lo, bb[b].lo); acc.hi = fmaxf3(acc.hi, bb[b].hi); ac += bc[b];
                    if (ac == 0 || rc[b + 1] == 0) continue;
                    float cost = boxArea(acc) * ac + boxArea(rb[b + 1]) * rc[b + 1];
                    if (cost < best) { best = cost; bestSplit = b; }
                }
                if (bestSplit >= 0) {
                    auto it = std::partition(order.begin() + tk.lo, order.begin() + tk.hi, [&](int id) { return binOf(id) <= bestSplit; });
                    mid = (int)(it - order.begin());
                }
            }
            if (mid <= tk.lo || mid >= tk.hi) {
                mid = (tk.lo + tk.hi) / 2;
                std::nth_element(order.begin() + tk.lo, order.begin() + mid, order.begin() + tk.hi,
                                 [&](int a, int b) { return (&cent[a].x)[axis] < (&cent[b].x)[axis]; });
            }
        }
        if (mid < 0) { nd.count = cnt; nodes[tk.node] = nd; continue; }
        nd.left = (int)nodes.size(); nodes.push_back(BvhNode());
        nd.right = (int)nodes.size(); nodes.push_back(BvhNode());
        nodes[tk.node] = nd;
        stack.push_back({nd.left, tk.lo, mid});
        stack.push_back({nd.right, mid, tk.hi});
    }
    std::vector<Prim> sorted(n);
    for (int i = 0; i < n; i++) sorted[i] = prims[order[i]];
    prims.swap(sorted);
}

// Primitive intersectors.  `tmax` is the current closest distance: rtPotentialIntersection accepts
// tmin < t < tmax.
bool Scene::intersectPrim(int pi, const Ray& ray, float tmax, Hit& hit) const
{
    const Prim& p = prims[pi];
    if (p.type == PRIM_TRIANGLE) {
        // TriangleMesh.cu:35-96
        const Mesh& m = meshes[p.a];
        const int3 index = m.indices[p.b];
        const float3 p0 = m.vertices[index.x], p1 = m.vertices[index.y], p2 = m.vertices[index.z];
        float3 n; float t, beta, gamma;
        Ray r = ray; r.tmax = tmax;
        if (!intersect_triangle(r, p0, p1, p2, n, t, beta, gamma)) return false;
        hit.t = t; hit.prim = pi; hit.material = m.material;
        hit.tangent = hit.bitangent = make_float3(0.f); hit.hasTangents = false;
        if (m.normals.empty()) {
            hit.shadingNormal = normalize(n);
        } else {
            const float3 n0 = m.normals[index.x], n1 = m.normals[index.y], n2 = m.normals[index.z];
            hit.shadingNormal = normalize(n1 * beta + n2 * gamma + n0 * (1.0f - beta - gamma));
            if (!m.tangents.empty()) {
                hit.hasTangents = true;
                const float3 t0 = m.tangents[index.x], t1 = m.tangents[index.y], t2 = m.tangents[index.z];
                hit.tangent = normalize(t1 * beta + t2 * gamma + t0 * (1.0f - beta - gamma));
                const float3 b0 = m.bitangents[index.x], b1 = m.bitangents[index.y], b2 = m.bitangents[index.z];
                hit.bitangent = normalize(b1 * beta + b2 * gamma + b0 * (1.0f - beta - gamma));
            }
        }
        hit.geometricNormal = normalize(n);
        if (m.uvs.empty()) {
            hit.textureCoordinate = make_float2(0.f, 0.f);
        } else {
            const float2 t0 = m.uvs[index.x], t1 = m.uvs[index.y], t2 = m.uvs[index.z];
            hit.textureCoordinate = t1 * beta + t2 * gamma + t0 * (1.0f - beta - gamma);
        }
        return true;
    }
    if (p.type == PRIM_PARALLELOGRAM) {
        // parallelogram.cu:49-76
        const Parallelogram& g = pgrams[p.a];
        float3 n = make_float3(g.plane);
        float dt = dot(ray.direction, n);
        float t = (g.plane.w - dot(n, ray.origin)) / dt;
        if (t > ray.tmin && t < tmax) {
            float3 pp = ray.origin + ray.direction * t;
            float3 vi = pp - g.anchor;
            float a1 = dot(g.v1, vi);
            if (a1 >= 0 && a1 <= 1) {
                float a2 = dot(g.v2, vi);
                if (a2 >= 0 && a2 <= 1) {
                    hit.t = t; hit.prim = pi; hit.material = g.material;
                    hit.shadingNormal = hit.geometricNormal = n;
                    hit.tangent = hit.bitangent = make_float3(0.f); hit.hasTangents = false;
                    hit.textureCoordinate = make_float2(0.f, 0.f);   // `texcoord` attribute is not `textureCoordinate`
                    return true;
                }
            }
        }
        return false;
    }
    // Sphere.cu:32-56
    const SpherePrim& s = spheres[p.a];
    float3 O = ray.origin - s.center;
    float b = dot(O, ray.direction);
    float c = dot(O, O) - s.radius * s.radius;
    float disc = b * b - c;
    if (disc > 0.0f) {
        float sdisc = sqrtf(disc);
        float root1 = (-b - sdisc);
        if (root1 > ray.tmin && root1 < tmax) {
            hit.t = root1; hit.prim = pi; hit.material = s.material;
            hit.shadingNormal = hit.geometricNormal = (O + root1 * ray.direction) / s.radius;
            hit.tangent = hit.bitangent = make_float3(0.f); hit.hasTangents = false; hit.textureCoordinate = make_float2(0.f, 0.f);
            return true;
        }
        float root2 = (-b + sdisc);
        if (root2 > ray.tmin && root2 < tmax) {
            hit.t = root2; hit.prim = pi; hit.material = s.material;
            hit.shadingNormal = hit.geometricNormal = (O + root2 * ray.direction) / s.radius;
            hit.tangent = hit.bitangent = make_float3(0.f); hit.hasTangents = false; hit.textureCoordinate = make_float2(0.f, 0.f);
            return true;
        }
    }
    return false;
}

static inline bool slab(const Aabb& b, const float3& o, const float3& inv, float tmin, float tmax)
{
    float t0 = (b.lo.x - o.x) * inv.x, t1 = (b.hi.x - o.x) * inv.x;
    float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
    t0 = (b.lo.y - o.y) * inv.y; t1 = (b.hi.y - o.y) * inv.y;
    lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
    t0 = (b.lo.z - o.z) * inv.z; t1 = (b.hi.z - o.z) * inv.z;
    lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
    // conservative: widen the interval so that a primitive lying exactly on a box face is never culled
    lo = lo - fabsf(lo) * 1e-4f - 1e-6f;
    hi = hi + fabsf(hi) * 1e-4f + 1e-6f;
    return lo <= hi && hi >= tmin && lo <= tmax;
}

bool Scene::closestHit(const Ray& ray, Hit& best) const
{
    if (nodes.empty()) return false;
    float3 inv = {1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z};
    float tmax = ray.tmax; bool found = false;
    int stack[128]; int sp = 0; stack[sp++] = 0;
    Hit h;
    while (sp) {
        const BvhNode& nd = nodes[stack[--sp]];
        if (!slab(nd.box, ray.origin, inv, ray.tmin, tmax)) continue;
        if (nd.count > 0) {
            for (int i = nd.first; i < nd.first + nd.count; i++)
                if (intersectPrim(i, ray, tmax, h)) { tmax = h.t; best = h; found = true; }
        } else { stack[sp++] = nd.right; stack[sp++] = nd.left; }
    }
    return found;
}

bool Scene::anyHit(const Ray& ray) const
{
    if (nodes.empty()) return false;
    float3 inv = {1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z};
    int stack[128]; int sp = 0; stack[sp++] = 0;
    Hit h;
    while (sp) {
        const BvhNode& nd = nodes[stack[--sp]];
        if (!slab(nd.box, ray.origin, inv, ray.tmin, ray.tmax)) continue;
        if (nd.count > 0) {
            for (int i = nd.first; i < nd.first + nd.count; i++)
                if (intersectPrim(i, ray, ray.tmax, h)) return true;
        } else { stack[sp++] = nd.right; stack[sp++] = nd.left; }
    }
    return false;
}

} // namespace ora
