// host_api.cpp — host-only pieces of the drop-in surface: Camera::setup / setAspectRatio, the Light
// constructors, the initial-radius estimate, the PPM radius schedule and the built-in Cornell /
// CornellSmall scene tables.  No CUDA here; these run on a machine without a GPU.
#include <cstdio>
#include "../../include/opr.h"
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    void put(float* p) const { p[0] = x; p[1] = y; p[2] = z; }
};
inline V3 add(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 divs(V3 a, float s) { float inv = 1.0f / s; return mul(a, inv); }        // OptiX float3 / float
inline float dotp(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 crossp(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float len(V3 a) { return sqrtf(dotp(a, a)); }
inline V3 unit(V3 a) { float inv = 1.0f / sqrtf(dotp(a, a)); return mul(a, inv); }  // OptiX normalize
inline bool finite(V3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }
inline float deg2rad(float d) { return d * (static_cast<float>(M_PI) / 180.f); }   // Camera.cpp DtoR
inline float rad2deg(float r) { return r * (180.f / static_cast<float>(M_PI)); }   // Camera.cpp RtoD

}  // namespace

extern "C" {

// Camera::setup, renderer/Camera.cpp:333-345 (assignWithCheck keeps the old value when the new one is not finite)
void opr_camera_setup(opr_camera* c)
{
    V3 lookdir(c->lookdir), up(c->up), u(c->camera_u), v(c->camera_v);
    V3 cand = sub(V3(c->lookat), V3(c->eye));
    if (finite(cand)) lookdir = cand;
    const float lookdir_len = len(lookdir);
    cand = unit(up); if (finite(cand)) up = cand;
    cand = unit(crossp(lookdir, up)); if (finite(cand)) u = cand;
    cand = unit(crossp(u, lookdir)); if (finite(cand)) v = cand;
    const float ulen = lookdir_len * tanf(deg2rad(c->hfov * 0.5f));
    cand = mul(u, ulen); if (finite(cand)) u = cand;
    const float vlen = lookdir_len * tanf(deg2rad(c->vfov * 0.5f));
    cand = mul(v, vlen); if (finite(cand)) v = cand;
    lookdir.put(c->lookdir); up.put(c->up); u.put(c->camera_u); v.put(c->camera_v);
    c->imagePlaneSize[0] = 2.0f * ulen;
    c->imagePlaneSize[1] = 2.0f * vlen;
}

// Camera::setAspectRatio, renderer/Camera.cpp:294-318
void opr_camera_set_aspect_ratio(opr_camera* c, float ratio)
{
    float realRatio = ratio;
    if (c->aspectRatioMode == OPR_KEEP_HORIZONTAL) {
        realRatio = 1.f / ratio;
        c->vfov = rad2deg(2.0f * atanf(realRatio * tanf(deg2rad(0.5f * c->hfov))));
    } else if (c->aspectRatioMode == OPR_KEEP_VERTICAL) {
        c->hfov = rad2deg(2.0f * atanf(realRatio * tanf(deg2rad(0.5f * c->vfov))));
    } else {
        return;
    }
    opr_camera_setup(c);
}

// Light::Light(power, position, v1, v2), renderer/Light.cpp:14-30
void opr_make_area_light(opr_light* out, const float* power3, const float* position3, const float* v1, const float* v2)
{
    memset(out, 0, sizeof *out);
    memcpy(out->power, power3, 12); memcpy(out->position, position3, 12); memcpy(out->v1, v1, 12); memcpy(out->v2, v2, 12);
    out->lightType = OPR_LIGHT_AREA;
    const V3 cp = crossp(V3(v1), V3(v2));
    unit(cp).put(out->normal);
    out->area_or_angle = len(cp);
    out->inverseArea = 1.0f / out->area_or_angle;
    const float kInvPi = 0.318309886183790671538f;
    for (int i = 0; i < 3; i++) out->Lemit[i] = power3[i] * out->inverseArea * kInvPi;
    out->isDelta = 0; out->isFinite = 1;
}

// Light::Light(power, position), renderer/Light.cpp:32-40
void opr_make_point_light(opr_light* out, const float* power3, const float* position3)
{
    memset(out, 0, sizeof *out);
    memcpy(out->power, power3, 12); memcpy(out->position, position3, 12);
    out->lightType = OPR_LIGHT_POINT;
    const float kInvPi = 0.318309886183790671538f;
    for (int i = 0; i < 3; i++) out->Lemit[i] = power3[i] * 0.25f * kInvPi;   // intensity
    out->isDelta = 1; out->isFinite = 1;
}

// IScene::getSceneInitialPPMRadiusEstimate, scene/IScene.cpp:23-31
float opr_initial_ppm_radius(const float* mn, const float* mx)
{
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    const float volume = ex * ey * ez;
    const float cubelength = powf(volume, 1.f / 3.f);
    const float A = 6 * cubelength * cubelength;
    return (float)(A * 3.94e-6);
}

// r_i from r_0: i steps of r^2 <- r^2 (k+alpha)/(k+1) in double with a sqrt per step, exactly as the driver
// loops (Standalone/StandaloneRenderManager.cpp:105-107)
double opr_ppm_radius_for_iteration(double initialRadius, uint64_t iteration, double alpha)
{
    double r = initialRadius;
    for (uint64_t k = 0; k < iteration; k++) {
        const double r2 = r * r;
        const double r2new = r2 * (k + alpha) / double(k + 1);
        r = sqrt(r2new);
    }
    return r;
}

}  // extern "C"

// ---- built-in scenes --------------------------------------------------------------------------------------
namespace {

struct SceneTable {
    std::string name;
    std::vector<opr_parallelogram> quads;
    std::vector<opr_sphere> balls;
    std::vector<opr_material> mats;
    std::vector<opr_light> lights;
    opr_camera cam;
    opr_scene_desc desc;

    int material(int kind, const float* Kd = nullptr, const float* Kr = nullptr, const float* Kt = nullptr, float ior = 0.f)
    {
        opr_material m; memset(&m, 0, sizeof m);
        m.kind = kind; m.diffuseTexture = m.normalTexture = -1; m.indexOfRefraction = ior;
        if (Kd) memcpy(m.Kd, Kd, 12);
        if (Kr) memcpy(m.Kr, Kr, 12);
        if (Kt) memcpy(m.Kt, Kt, 12);
        mats.push_back(m);
        return (int)mats.size() - 1;
    }
    int diffuse(float r, float g, float b) { float k[3] = {r, g, b}; return material(OPR_MAT_DIFFUSE, k); }
    int emitter(const opr_light& l)
    {   // DiffuseEmitter(power, Vector3(1)) + setInverseArea (Cornell.cpp:150-151, DiffuseEmitter.cpp:16-21)
        float one[3] = {1.f, 1.f, 1.f};
        int id = material(OPR_MAT_DIFFUSE_EMITTER, one);
        for (int i = 0; i < 3; i++) mats[id].power[i] = l.power[i] * one[i];
        mats[id].inverseArea = l.inverseArea;
        return id;
    }
    void quad(float ax, float ay, float az, float ux, float uy, float uz, float vx, float vy, float vz, int mat, float scale = 0.f)
    {
        opr_parallelogram p;
        V3 a(ax, ay, az), u(ux, uy, uz), v(vx, vy, vz);
        if (scale != 0.f) { a = divs(a, scale); u = divs(u, scale); v = divs(v, scale); }   // make_float3(..) / 220.f
        a.put(p.anchor); u.put(p.offset1); v.put(p.offset2); p.material = mat;
        quads.push_back(p);
    }
    void ball(float cx, float cy, float cz, float radius, int mat)
    {
        opr_sphere s; s.center[0] = cx; s.center[1] = cy; s.center[2] = cz; s.radius = radius; s.material = mat;
        balls.push_back(s);
    }
    void seal(float lo, float hx, float hy, float hz, float pad)
    {
        memset(&desc, 0, sizeof desc);
        desc.name = name.c_str();
        desc.numParallelograms = (int)quads.size(); desc.parallelograms = quads.data();
        desc.numSpheres = (int)balls.size(); desc.spheres = balls.data();
        desc.numMaterials = (int)mats.size(); desc.materials = mats.data();
        desc.numLights = (int)lights.size(); desc.lights = lights.data();
        desc.aabbMin[0] = desc.aabbMin[1] = desc.aabbMin[2] = lo;
        desc.aabbMax[0] = hx + pad; desc.aabbMax[1] = hy + pad; desc.aabbMax[2] = hz + pad;
    }
    void camera(float ex, float ey, float ez, float lx, float ly, float lz, float fov)
    {
        memset(&cam, 0, sizeof cam);
        cam.eye[0] = ex; cam.eye[1] = ey; cam.eye[2] = ez; cam.lookat[0] = lx; cam.lookat[1] = ly; cam.lookat[2] = lz;
        cam.up[1] = 1.f; cam.hfov = cam.vfov = fov;
        cam.aperture = 0.f;   // getDefaultCamera passes Camera::KeepVertical (== 0) in the aperture slot (Cornell.cpp:199-207)
        cam.aspectRatioMode = OPR_KEEP_VERTICAL;
        opr_camera_setup(&cam);
    }
};

// scene/Cornell.cpp:20-192
SceneTable* cornell()
{
    SceneTable* s = new SceneTable; s->name = "Cornell";
    const float anchor[3] = {343.0f, 548.7999f, 227.0f}, v1[3] = {0.0f, 0.0f, 105.0f}, v2[3] = {-130.0f, 0.0f, 0.0f};
    const float power[3] = {0.5e6f, 0.4e6f, 0.2e6f};
    opr_light l; opr_make_area_light(&l, power, anchor, v1, v2);
    s->lights.push_back(l);
    const int white = s->diffuse(0.8f, 0.8f, 0.8f), green = s->diffuse(0.05f, 0.8f, 0.05f), red = s->diffuse(1.f, 0.05f, 0.05f);
    s->quad(0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 559.2f, 556.0f, 0.0f, 0.0f, white);       // floor
    s->quad(0.0f, 548.80f, 0.0f, 556.0f, 0.0f, 0.0f, 0.0f, 0.0f, 559.2f, white);    // ceiling
    s->quad(0.0f, 0.0f, 559.2f, 0.0f, 548.8f, 0.0f, 556.0f, 0.0f, 0.0f, white);     // back wall
    s->quad(0.0f, 0.0f, 0.0f, 0.0f, 548.8f, 0.0f, 0.0f, 0.0f, 559.2f, green);       // right wall
    s->quad(556.0f, 0.0f, 0.0f, 0.0f, 0.0f, 559.2f, 0.0f, 548.8f, 0.0f, red);       // left wall
    const int em = s->emitter(s->lights[0]);
    for (const opr_light& li : s->lights)
        s->quad(li.position[0], li.position[1], li.position[2], li.v1[0], li.v1[1], li.v1[2], li.v2[0], li.v2[1], li.v2[2], em);
    s->seal(-5.f, 556.0f, 548.85f, 559.2f, 5.f);                                      // Cornell.cpp:29-30
    s->camera(278.0f, 273.0f, -850.0f, 278.0f, 273.0f, 0.0f, 35.0f);
    return s;
}

enum {   // scene/CornellSmall.h:25-42
    kSmallVCMColors = 1 << 0, kLightArea = 1 << 1, kLightAreaUpwards = 1 << 2, kLightPoint = 1 << 3, kLightPointStrong = 1 << 4,
    kLightPointDistant = 1 << 5, kBackwallBlue = 1 << 6, kFloorMirror = 1 << 7, kFloorGlossy = 1 << 8, kBlocks = 1 << 9,
    kLargeMirrorSphere = 1 << 10, kLargeGlassSphere = 1 << 11, kSmallMirrorSphere = 1 << 12, kSmallGlassSphere = 1 << 13
};

// scene/CornellSmall.cpp:25-326
SceneTable* cornellSmall(unsigned cfg)
{
    SceneTable* s = new SceneTable; s->name = "CornellSmall";
    const bool area = (cfg & kLightArea) || (cfg & kLightAreaUpwards);
    if (area) {
        float anchor[3] = {1.f, 2.499f, 1.f}, v1[3] = {0.5f, 0.0f, 0.0f}, v2[3] = {0.0f, 0.0f, 0.5f};
        if (cfg & kLightAreaUpwards) { for (int i = 0; i < 3; i++) std::swap(v1[i], v2[i]); anchor[1] -= .1f; }
        const float p = 19.661107023935260172519494336416f;
        const float power[3] = {p, p, p};
        opr_light l; opr_make_area_light(&l, power, anchor, v1, v2);
        s->lights.push_back(l);
    } else if ((cfg & kLightPoint) || (cfg & kLightPointStrong) || (cfg & kLightPointDistant)) {
        float anchor[3] = {1.25f, 2.25f, 1.25f};
        float p = 30.f;
        if (cfg & kLightPointStrong) p = 70.f;
        if (cfg & kLightPointDistant) { p = 200.f; anchor[1] += 5.f; }
        const float power[3] = {p, p, p};
        opr_light l; opr_make_point_light(&l, power, anchor);
        s->lights.push_back(l);
    }
    int white, green, red;
    if (cfg & kSmallVCMColors) {
        white = s->diffuse(0.803922f, 0.803922f, 0.803922f); green = s->diffuse(0.156863f, 0.803922f, 0.172549f); red = s->diffuse(0.803922f, 0.152941f, 0.152941f);
    } else {
        white = s->diffuse(0.8f, 0.8f, 0.8f); green = s->diffuse(0.05f, 0.8f, 0.05f); red = s->diffuse(1.f, 0.05f, 0.05f);
    }
    const int blue = s->diffuse(0.156863f, 0.172549f, 0.803922f);
    const float one[3] = {1.f, 1.f, 1.f};
    const int mirror = s->material(OPR_MAT_MIRROR, nullptr, one);
    int glossy;
    {   // Glossy(Kd=.1, Ks=.7, 90): rescaled only if max(Kd+Ks) > 1 (material/Glossy.cpp:16-33)
        float Kd[3] = {.1f, .1f, .1f}, Ks[3] = {.7f, .7f, .7f};
        float m = Kd[0] + Ks[0]; for (int i = 1; i < 3; i++) m = (Kd[i] + Ks[i]) > m ? (Kd[i] + Ks[i]) : m;
        const float sumScale = 1.f / m;
        if (sumScale < 1.f) for (int i = 0; i < 3; i++) { Kd[i] *= sumScale; Ks[i] *= sumScale; }
        glossy = s->material(OPR_MAT_GLOSSY, Kd);
        memcpy(s->mats[glossy].Ks, Ks, 12); s->mats[glossy].exponent = 90.f;
    }
    const int glass = s->material(OPR_MAT_GLASS, nullptr, one, one, (float)1.6);

    int floorMat = white;
    if (cfg & kFloorMirror) floorMat = mirror; else if (cfg & kFloorGlossy) floorMat = glossy;
    const int back = (cfg & kBackwallBlue) ? blue : white;
    const int rightWall = (cfg & kSmallVCMColors) ? red : green;
    const int leftWall = (cfg & kSmallVCMColors) ? green : red;

    s->quad(0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 2.5f, 2.5f, 0.0f, 0.0f, floorMat);
    if (!(cfg & kLightPointDistant)) s->quad(0.0f, 2.5f, 0.0f, 2.5f, 0.0f, 0.0f, 0.0f, 0.0f, 2.5f, white);
    s->quad(0.0f, 0.0f, 2.5f, 0.0f, 2.5f, 0.0f, 2.5f, 0.0f, 0.0f, back);
    s->quad(0.0f, 0.0f, 0.0f, 0.0f, 2.5f, 0.0f, 0.0f, 0.0f, 2.5f, rightWall);
    s->quad(2.5f, 0.0f, 0.0f, 0.0f, 0.0f, 2.5f, 0.0f, 2.5f, 0.0f, leftWall);
    if (cfg & kBlocks) {
        const float k = 220.f;
        s->quad(130.0f, 165.0f, 65.0f, -48.0f, 0.0f, 160.0f, 160.0f, 0.0f, 49.0f, white, k);   // short block
        s->quad(290.0f, 0.0f, 114.0f, 0.0f, 165.0f, 0.0f, -50.0f, 0.0f, 158.0f, white, k);
        s->quad(130.0f, 0.0f, 65.0f, 0.0f, 165.0f, 0.0f, 160.0f, 0.0f, 49.0f, white, k);
        s->quad(82.0f, 0.0f, 225.0f, 0.0f, 165.0f, 0.0f, 48.0f, 0.0f, -160.0f, white, k);
        s->quad(240.0f, 0.0f, 272.0f, 0.0f, 165.0f, 0.0f, -158.0f, 0.0f, -47.0f, white, k);
        s->quad(423.0f, 340.0f, 247.0f, -158.0f, 0.0f, 49.0f, 49.0f, 0.0f, 159.0f, white, k);  // tall block
        s->quad(423.0f, 0.0f, 247.0f, 0.0f, 340.0f, 0.0f, 49.0f, 0.0f, 159.0f, white, k);
        s->quad(472.0f, 0.0f, 406.0f, 0.0f, 340.0f, 0.0f, -158.0f, 0.0f, 50.0f, white, k);
        s->quad(314.0f, 0.0f, 456.0f, 0.0f, 340.0f, 0.0f, -49.0f, 0.0f, -160.0f, white, k);
        s->quad(265.0f, 0.0f, 296.0f, 0.0f, 340.1f, 0.0f, 158.0f, 0.0f, -49.0f, white, k);
    }
    if (area) {
        const int em = s->emitter(s->lights[0]);
        for (const opr_light& li : s->lights)
            s->quad(li.position[0], li.position[1], li.position[2], li.v1[0], li.v1[1], li.v1[2], li.v2[0], li.v2[1], li.v2[2], em);
    }
    if ((cfg & kLargeMirrorSphere) || (cfg & kLargeGlassSphere)) {
        const float radius = (float)0.8;
        s->ball(1.25f, radius, 1.25f, radius, (cfg & kLargeGlassSphere) ? glass : mirror);
    }
    if (cfg & kSmallGlassSphere) s->ball(1.25f - 0.535714269f, 0.5f, 1.25f, 0.5f, glass);
    if (cfg & kSmallMirrorSphere) s->ball(1.25f + 0.535714269f, 0.5f, 1.25f, 0.5f, mirror);
    s->seal(-0.1f, 2.5f, 2.5f, 2.5f, 0.1f);                                           // CornellSmall.cpp:66-67
    s->camera(1.25f, 1.25f, -2.85f, 1.25f, 1.25f, 0.0f, 45.0f);
    return s;
}

}  // namespace

extern "C" {

void* opr_builtin_scene_create(const char* name, uint32_t config)
{
    if (!name) return nullptr;
    if (!strcmp(name, "Cornell")) return cornell();
    if (!strcmp(name, "CornellSmall")) return cornellSmall(config ? config : (kLightArea | kBlocks));
    return nullptr;
}
void opr_builtin_scene_destroy(void* scene) { delete static_cast<SceneTable*>(scene); }
const opr_scene_desc* opr_builtin_scene_desc(void* scene) { return &static_cast<SceneTable*>(scene)->desc; }
void opr_builtin_scene_camera(void* scene, opr_camera* out) { *out = static_cast<SceneTable*>(scene)->cam; }

}  // extern "C"

// ---- opr_scene_desc sanity (called by initScene before anything is uploaded) ---------------------------------------------
// The reference trusts its own Scene / Cornell classes; a C ABI cannot: every index the device code will follow is checked here.
extern "C" int opr_scene_desc_check(const opr_scene_desc* d, char* msg, int capacity)
{
    auto fail = [&](const char* fmt, int a, int b) { if (msg && capacity > 0) snprintf(msg, (size_t)capacity, fmt, a, b); return 1; };
    if (!d) return fail("scene description is NULL", 0, 0);
    if (d->numMeshes < 0 || d->numParallelograms < 0 || d->numSpheres < 0 || d->numMaterials < 0 || d->numTextures < 0 || d->numLights < 0)
        return fail("scene description with a negative count", 0, 0);
    if ((d->numMeshes && !d->meshes) || (d->numParallelograms && !d->parallelograms) || (d->numSpheres && !d->spheres) ||
        (d->numMaterials && !d->materials) || (d->numTextures && !d->textures) || (d->numLights && !d->lights))
        return fail("scene description with a count but no array", 0, 0);
    for (int i = 0; i < d->numTextures; i++) {
        const opr_texture& t = d->textures[i];
        if (t.width <= 0 || t.height <= 0 || !t.rgba) return fail("texture %d has no texels", i, 0);
    }
    for (int i = 0; i < d->numMaterials; i++) {
        const opr_material& m = d->materials[i];
        if (m.kind < 0 || m.kind >= OPR_MAT_NUM_KINDS) return fail("material %d has unknown kind %d", i, m.kind);
        if (m.kind == OPR_MAT_TEXTURE) {
            if (m.diffuseTexture < 0 || m.diffuseTexture >= d->numTextures) return fail("Texture material %d without a diffuse texture", i, 0);
            if (m.normalTexture >= d->numTextures) return fail("Texture material %d: normal map %d does not exist", i, m.normalTexture);
        }
    }
    auto badMaterial = [&](int m) { return m < 0 || m >= d->numMaterials; };
    for (int i = 0; i < d->numMeshes; i++) {
        const opr_mesh& me = d->meshes[i];
        if (me.numVertices < 0 || me.numTriangles < 0) return fail("mesh %d with a negative count", i, 0);
        if (badMaterial(me.material)) return fail("mesh %d: material %d does not exist", i, me.material);
        if (me.numTriangles && (!me.indices || !me.vertices)) return fail("mesh %d has triangles but no index / vertex array", i, 0);
        for (long long k = 0; k < 3LL * me.numTriangles; k++)
            if (me.indices[k] < 0 || me.indices[k] >= me.numVertices) return fail("mesh %d: vertex index %d out of range", i, me.indices[k]);
    }
    for (int i = 0; i < d->numParallelograms; i++)
        if (badMaterial(d->parallelograms[i].material)) return fail("parallelogram %d: material %d does not exist", i, d->parallelograms[i].material);
    for (int i = 0; i < d->numSpheres; i++)
        if (badMaterial(d->spheres[i].material)) return fail("sphere %d: material %d does not exist", i, d->spheres[i].material);
    for (int i = 0; i < d->numLights; i++)
        if (d->lights[i].lightType < OPR_LIGHT_AREA || d->lights[i].lightType > OPR_LIGHT_SPOT) return fail("light %d has unknown type %d", i, d->lights[i].lightType);
    return 0;
}
