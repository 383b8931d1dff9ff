// ORACLE — test infrastructure only (see ora_math.h).  Built-in scenes restated from
// RenderEngine/scene/Cornell.cpp:20-207 and RenderEngine/scene/CornellSmall.cpp:25-342.
#include "ora_core.h"
#include <cstring>
#include <memory>

namespace ora {

struct BuiltinScene {
    std::string name;
    std::vector<opr_parallelogram> pgrams;
    std::vector<opr_sphere> spheres;
    std::vector<opr_material> mats;
    std::vector<opr_light> lights;
    opr_scene_desc desc;
    Camera camera;

    int addMat(const opr_material& m) { mats.push_back(m); return (int)mats.size() - 1; }
    void addPgram(float3 a, float3 o1, float3 o2, int mat)
    {
        opr_parallelogram p; p.anchor[0] = a.x; p.anchor[1] = a.y; p.anchor[2] = a.z;
        p.offset1[0] = o1.x; p.offset1[1] = o1.y; p.offset1[2] = o1.z;
        p.offset2[0] = o2.x; p.offset2[1] = o2.y; p.offset2[2] = o2.z; p.material = mat;
        pgrams.push_back(p);
    }
    void finish(float3 mn, float3 mx)
    {
        memset(&desc, 0, sizeof desc);
        desc.name = name.c_str();
        desc.numParallelograms = (int)pgrams.size(); desc.parallelograms = pgrams.data();
        desc.numSpheres = (int)spheres.size(); desc.spheres = spheres.data();
        desc.numMaterials = (int)mats.size(); desc.materials = mats.data();
        desc.numLights = (int)lights.size(); desc.lights = lights.data();
        desc.aabbMin[0] = mn.x; desc.aabbMin[1] = mn.y; desc.aabbMin[2] = mn.z;
        desc.aabbMax[0] = mx.x; desc.aabbMax[1] = mx.y; desc.aabbMax[2] = mx.z;
    }
};

static opr_material blank(int kind)
{
    opr_material m; memset(&m, 0, sizeof m); m.kind = kind; m.diffuseTexture = m.normalTexture = -1; return m;
}
static opr_material diffuse(float3 kd) { opr_material m = blank(OPR_MAT_DIFFUSE); m.Kd[0] = kd.x; m.Kd[1] = kd.y; m.Kd[2] = kd.z; return m; }
static opr_material mirror(float3 kr) { opr_material m = blank(OPR_MAT_MIRROR); m.Kr[0] = kr.x; m.Kr[1] = kr.y; m.Kr[2] = kr.z; return m; }
static opr_material glass(float ior, float3 kr, float3 kt)
{ opr_material m = blank(OPR_MAT_GLASS); m.indexOfRefraction = ior; m.Kr[0] = kr.x; m.Kr[1] = kr.y; m.Kr[2] = kr.z; m.Kt[0] = kt.x; m.Kt[1] = kt.y; m.Kt[2] = kt.z; return m; }
static opr_material glossy(float3 Kd, float3 Ks, float exponent)
{   // material/Glossy.cpp:16-33
    opr_material m = blank(OPR_MAT_GLOSSY);
    float3 sumK = Kd + Ks;
    float sumScale = 1.f / fmaxf3(sumK);
    if (sumScale < 1.f) { Kd = Kd * sumScale; Ks = Ks * sumScale; }
    m.Kd[0] = Kd.x; m.Kd[1] = Kd.y; m.Kd[2] = Kd.z; m.Ks[0] = Ks.x; m.Ks[1] = Ks.y; m.Ks[2] = Ks.z; m.exponent = exponent;
    return m;
}
static opr_material emitter(float3 power, float3 Kd, float inverseArea)
{   // material/DiffuseEmitter.cpp:16-21
    opr_material m = blank(OPR_MAT_DIFFUSE_EMITTER);
    m.power[0] = power.x * Kd.x; m.power[1] = power.y * Kd.y; m.power[2] = power.z * Kd.z;
    m.Kd[0] = Kd.x; m.Kd[1] = Kd.y; m.Kd[2] = Kd.z; m.inverseArea = inverseArea;
    return m;
}

static BuiltinScene* makeCornell()
{
    BuiltinScene* s = new BuiltinScene(); s->name = "Cornell";
    float3 anchor = make_float3(343.0f, 548.7999f, 227.0f);
    float3 v1 = make_float3(0.0f, 0.0f, 105.0f), v2 = make_float3(-130.0f, 0.0f, 0.0f);
    float3 power = make_float3(0.5e6f, 0.4e6f, 0.2e6f);
    s->lights.push_back(make_area_light(power, anchor, v1, v2));
    int white = s->addMat(diffuse(make_float3(0.8f)));
    int green = s->addMat(diffuse(make_float3(0.05f, 0.8f, 0.05f)));
    int red = s->addMat(diffuse(make_float3(1.f, 0.05f, 0.05f)));
    s->addPgram(make_float3(0.0f, 0.0f, 0.0f), make_float3(0.0f, 0.0f, 559.2f), make_float3(556.0f, 0.0f, 0.0f), white);      // floor
    s->addPgram(make_float3(0.0f, 548.80f, 0.0f), make_float3(556.0f, 0.0f, 0.0f), make_float3(0.0f, 0.0f, 559.2f), white);  // ceiling
    s->addPgram(make_float3(0.0f, 0.0f, 559.2f), make_float3(0.0f, 548.8f, 0.0f), make_float3(556.0f, 0.0f, 0.0f), white);   // back
    s->addPgram(make_float3(0.0f, 0.0f, 0.0f), make_float3(0.0f, 548.8f, 0.0f), make_float3(0.0f, 0.0f, 559.2f), green);     // right
    s->addPgram(make_float3(556.0f, 0.0f, 0.0f), make_float3(0.0f, 0.0f, 559.2f), make_float3(0.0f, 548.8f, 0.0f), red);     // left
    int em = s->addMat(emitter(f3(s->lights[0].power), make_float3(1.f), s->lights[0].inverseArea));
    for (size_t i = 0; i < s->lights.size(); i++)
        s->addPgram(f3(s->lights[i].position), f3(s->lights[i].v1), f3(s->lights[i].v2), em);
    s->finish(make_float3(-5.f), make_float3(556.0f, 548.85f, 559.2f) + 5.f);
    // Cornell::getDefaultCamera passes Camera::KeepVertical (=0) in the aperture slot (Cornell.cpp:199-207)
    s->camera = make_camera(make_float3(278.0f, 273.0f, -850.0f), make_float3(278.0f, 273.0f, 0.0f), make_float3(0.0f, 1.0f, 0.0f), 35.0f, 35.0f, 0.f, OPR_KEEP_VERTICAL);
    return s;
}

enum {   // CornellSmall.h:25-42
    SmallVCMColors = 1 << 0, LightArea = 1 << 1, LightAreaUpwards = 1 << 2, LightPoint = 1 << 3, LightPointStrong = 1 << 4,
    LightPointDistant = 1 << 5, BackwallBlue = 1 << 6, FloorMirror = 1 << 7, FloorGlossy = 1 << 8, Blocks = 1 << 9,
    LargeMirrorSphere = 1 << 10, LargeGlassSphere = 1 << 11, SmallMirrorSphere = 1 << 12, SmallGlassSphere = 1 << 13
};

static BuiltinScene* makeCornellSmall(unsigned cfg)
{
    BuiltinScene* s = new BuiltinScene(); s->name = "CornellSmall";
    if ((cfg & LightArea) || (cfg & LightAreaUpwards)) {
        float3 anchor = make_float3(1.f, 2.499f, 1.f);
        float3 v1 = make_float3(0.5f, 0.0f, 0.0f), v2 = make_float3(0.0f, 0.0f, 0.5f);
        if (cfg & LightAreaUpwards) { float3 t = v1; v1 = v2; v2 = t; anchor -= make_float3(0.0f, .1f, 0.f); }
        float3 power = make_float3(19.661107023935260172519494336416f);
        s->lights.push_back(make_area_light(power, anchor, v1, v2));
    } else if ((cfg & LightPoint) || (cfg & LightPointStrong) || (cfg & LightPointDistant)) {
        float3 anchor = make_float3(1.25f, 2.25f, 1.25f);
        float power = 30.f;
        if (cfg & LightPointStrong) power = 70.f;
        if (cfg & LightPointDistant) { power = 200.f; anchor += make_float3(0.f, 5.f, 0.f); }
        s->lights.push_back(make_point_light(make_float3(power), anchor));
    }
    opr_material dWhite = diffuse(make_float3(0.8f)), dGreen = diffuse(make_float3(0.05f, 0.8f, 0.05f)), dRed = diffuse(make_float3(1.f, 0.05f, 0.05f));
    if (cfg & SmallVCMColors) {
        dWhite = diffuse(make_float3(0.803922f, 0.803922f, 0.803922f));
        dGreen = diffuse(make_float3(0.156863f, 0.803922f, 0.172549f));
        dRed = diffuse(make_float3(0.803922f, 0.152941f, 0.152941f));
    }
    int white = s->addMat(dWhite), green = s->addMat(dGreen), red = s->addMat(dRed);
    int blue = s->addMat(diffuse(make_float3(0.156863f, 0.172549f, 0.803922f)));
    int mir = s->addMat(mirror(make_float3(1.f, 1.f, 1.f)));
    int glo = s->addMat(glossy(make_float3(.1f, .1f, .1f), make_float3(.7f, .7f, .7f), 90.f));
    int gla = s->addMat(glass((float)1.6, make_float3(1.f, 1.f, 1.f), make_float3(1.f, 1.f, 1.f)));
    int matFloor = white;
    if (cfg & FloorMirror) matFloor = mir; else if (cfg & FloorGlossy) matFloor = glo;
    int matBack = (cfg & BackwallBlue) ? blue : white;
    int matRight = (cfg & SmallVCMColors) ? red : green;
    int matLeft = (cfg & SmallVCMColors) ? green : red;

    s->addPgram(make_float3(0.0f, 0.0f, 0.0f), make_float3(0.0f, 0.0f, 2.5f), make_float3(2.5f, 0.0f, 0.0f), matFloor);
    if (!(cfg & LightPointDistant))
        s->addPgram(make_float3(0.0f, 2.5f, 0.0f), make_float3(2.5f, 0.0f, 0.0f), make_float3(0.0f, 0.0f, 2.5f), white);
    s->addPgram(make_float3(0.0f, 0.0f, 2.5f), make_float3(0.0f, 2.5f, 0.0f), make_float3(2.5f, 0.0f, 0.0f), matBack);
    s->addPgram(make_float3(0.0f, 0.0f, 0.0f), make_float3(0.0f, 2.5f, 0.0f), make_float3(0.0f, 0.0f, 2.5f), matRight);
    s->addPgram(make_float3(2.5f, 0.0f, 0.0f), make_float3(0.0f, 0.0f, 2.5f), make_float3(0.0f, 2.5f, 0.0f), matLeft);
    if (cfg & Blocks) {
        const float k = 220.f;
        auto P = [&](float ax, float ay, float az, float bx, float by, float bz, float cx, float cy, float cz) {
            s->addPgram(make_float3(ax, ay, az) / k, make_float3(bx, by, bz) / k, make_float3(cx, cy, cz) / k, white);
        };
        P(130.0f, 165.0f, 65.0f, -48.0f, 0.0f, 160.0f, 160.0f, 0.0f, 49.0f);
        P(290.0f, 0.0f, 114.0f, 0.0f, 165.0f, 0.0f, -50.0f, 0.0f, 158.0f);
        P(130.0f, 0.0f, 65.0f, 0.0f, 165.0f, 0.0f, 160.0f, 0.0f, 49.0f);
        P(82.0f, 0.0f, 225.0f, 0.0f, 165.0f, 0.0f, 48.0f, 0.0f, -160.0f);
        P(240.0f, 0.0f, 272.0f, 0.0f, 165.0f, 0.0f, -158.0f, 0.0f, -47.0f);
        P(423.0f, 340.0f, 247.0f, -158.0f, 0.0f, 49.0f, 49.0f, 0.0f, 159.0f);
        P(423.0f, 0.0f, 247.0f, 0.0f, 340.0f, 0.0f, 49.0f, 0.0f, 159.0f);
        P(472.0f, 0.0f, 406.0f, 0.0f, 340.0f, 0.0f, -158.0f, 0.0f, 50.0f);
        P(314.0f, 0.0f, 456.0f, 0.0f, 340.0f, 0.0f, -49.0f, 0.0f, -160.0f);
        P(265.0f, 0.0f, 296.0f, 0.0f, 340.1f, 0.0f, 158.0f, 0.0f, -49.0f);
    }
    if ((cfg & LightArea) || (cfg & LightAreaUpwards)) {
        int em = s->addMat(emitter(f3(s->lights[0].power), make_float3(1.f), s->lights[0].inverseArea));
        for (size_t i = 0; i < s->lights.size(); i++)
            s->addPgram(f3(s->lights[i].position), f3(s->lights[i].v1), f3(s->lights[i].v2), em);
    }
    auto addSphere = [&](float3 c, float rad, int mat) { opr_sphere sp; sp.center[0] = c.x; sp.center[1] = c.y; sp.center[2] = c.z; sp.radius = rad; sp.material = mat; s->spheres.push_back(sp); };
    if ((cfg & LargeMirrorSphere) || (cfg & LargeGlassSphere)) {
        float radius = (float)0.8;
        addSphere(make_float3(1.25f, radius, 1.25f), radius, (cfg & LargeGlassSphere) ? gla : mir);
    }
    if (cfg & SmallGlassSphere) { float radius = 0.5f; addSphere(make_float3(1.25f - 0.535714269f, radius, 1.25f), radius, gla); }
    if (cfg & SmallMirrorSphere) { float radius = 0.5f; addSphere(make_float3(1.25f + 0.535714269f, radius, 1.25f), radius, mir); }
    s->finish(make_float3(-0.1f), make_float3(2.5f, 2.5f, 2.5f) + 0.1f);
    s->camera = make_camera(make_float3(1.25f, 1.25f, -2.85f), make_float3(1.25f, 1.25f, 0.0f), make_float3(0.0f, 1.0f, 0.0f), 45.0f, 45.0f, 0.f, OPR_KEEP_VERTICAL);
    return s;
}

} // namespace ora

extern "C" {
// name: "Cornell" or "CornellSmall"; config: CornellSmall::Config bits (0 => Default = LightArea|Blocks)
void* ora_builtin_scene_create(const char* name, unsigned config)
{
    if (!strcmp(name, "Cornell")) return ora::makeCornell();
    if (!strcmp(name, "CornellSmall")) return ora::makeCornellSmall(config ? config : (ora::LightArea | ora::Blocks));
    return nullptr;
}
void ora_builtin_scene_destroy(void* s) { delete (ora::BuiltinScene*)s; }
const opr_scene_desc* ora_builtin_scene_desc(void* s) { return &((ora::BuiltinScene*)s)->desc; }
void ora_builtin_scene_camera(void* s, opr_camera* out) { *out = ((ora::BuiltinScene*)s)->camera; }
}
