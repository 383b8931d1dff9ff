// ORACLE/_ref — test infrastructure only.  Runtime of the OptiX stand-in: variable binding, rtTrace over the
// oracle's CPU BVH with the REFERENCE's intersection programs at the leaves, dispatch to the REFERENCE's
// material programs per ray type (binding table: SURVEY.md Appendix D / material/*.cpp), and the ora_* program
// entry points (ppm_generateRay, ..., vcm_cameraPass) that oracle/ora_render.cpp's orchestration calls.
// Linked with oracle/ora_{scene,grid,render,builtin}.cpp instead of ora_{ppm,vcm}.cpp -> oracle/_ref/liboppref.so.
#include "shim/shim_optix.h"
#include "shim_rt.h"
#include "../ora_render.h"
#include <map>
#include <unordered_map>
#include <mutex>

namespace ora { float4 tex2D_raw(const uint8_t* rgba, int width, int height, float u, float v); }

namespace shim {

// ---- registry -------------------------------------------------------------------------------------------------
static std::map<std::string, TURuntime*>& registry() { static std::map<std::string, TURuntime*> r; return r; }
static std::vector<std::string>& internNames() { static std::vector<std::string> v; return v; }
static std::mutex g_internMutex;
static int intern(const std::string& s)
{
    std::lock_guard<std::mutex> l(g_internMutex);
    auto& v = internNames();
    for (size_t i = 0; i < v.size(); i++) if (v[i] == s) return (int)i;
    v.push_back(s);
    return (int)v.size() - 1;
}

TU::TU(const char* n) : name(n) {}

TURuntime::TURuntime(TU* t) : tu(t) { registry()[t->name] = this; }

static void finalize(TURuntime& r)
{
    if (r.finalized) return;
    for (const Var& v : r.tu->vars) {
        VarPlan p; p.var = v; p.id = -1; p.kind = VK_CONTEXT;
        std::string sem = v.semantic;
        if (sem == "rtLaunchIndex") p.kind = VK_LAUNCH_INDEX;
        else if (sem == "rtLaunchDim") p.kind = VK_LAUNCH_DIM;
        else if (sem == "rtCurrentRay") p.kind = VK_RAY;
        else if (sem == "rtIntersectionDistance") p.kind = VK_THIT;
        else if (sem == "rtPayload") p.kind = VK_PAYLOAD;
        else if (sem.rfind("attribute", 0) == 0) { p.kind = VK_ATTRIBUTE; p.id = intern("attr:" + sem.substr(sem.find(' ') + 1)); }
        else p.id = intern(v.name);
        r.plan.push_back(p);
    }
    for (auto& b : r.buffers) b.id = intern(std::string("buf:") + b.name);
    r.finalized = true;
}

static TURuntime& getTU(const char* name)
{
    auto it = registry().find(name);
    if (it == registry().end()) { fprintf(stderr, "refshim: translation unit %s is not linked\n", name); abort(); }
    finalize(*it->second);
    return *it->second;
}

// ---- per-render context ------------------------------------------------------------------------------------------
struct Slot { const void* ptr; size_t size; };
struct RefCtx {
    ora::Renderer* r = nullptr; unsigned long long uid = 0;
    std::vector<curandState> randomStates; unsigned rngEpoch = ~0u;
    std::vector<unsigned> dbgCells, dbgPhotons;
    std::vector<char> lightVertexRaw; size_t lightVertexElem = 0; size_t lightVertexCount = 0;
    // context scalars materialised with the reference's types
    unsigned maxPhotonDepositsPerEmitted = 4; int bufId[8];
    uint3 photonsGridSize; float3 photonsWorldOrigo;
    unsigned lightSubpathLengthDummy = 0;
    std::unordered_map<int, Slot> vars;           // interned name -> value
    std::unordered_map<int, BufferDesc> buffers;  // interned "buf:name" -> descriptor
    std::vector<TextureData> textures;
    int bufIntern[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    struct SceneTables* tables = nullptr;
    unsigned long long epoch = 0;                  // bumped every launch so that threads re-bind context variables
};
static std::mutex g_ctxMutex;
static std::map<ora::Renderer*, RefCtx*> g_ctx;
static RefCtx& ctxFor(ora::Renderer& r)
{
    std::lock_guard<std::mutex> l(g_ctxMutex);
    RefCtx*& c = g_ctx[&r];
    if (c && c->uid != r.uid) { delete c; c = nullptr; }   // a new Renderer was allocated at a recycled address
    if (!c) { c = new RefCtx(); c->r = &r; c->uid = r.uid; }
    return *c;
}

static thread_local RefCtx* t_ctx = nullptr;
static thread_local uint2 t_launchIndex, t_launchDim;

static void setVar(RefCtx& c, const char* name, const void* p, size_t n) { c.vars[intern(name)] = Slot{p, n}; }
static void setBuf(RefCtx& c, const char* name, void* data, size_t elem, size_t d0, size_t d1 = 1, size_t d2 = 1)
{ BufferDesc b = {data, elem, {d0, d1, d2}}; c.buffers[intern(std::string("buf:") + name)] = b; }

enum { BID_OUTPUT = 1, BID_LIGHT_VERTEX, BID_LIGHT_VERTEX_INDEX, BID_VERTEX_COUNT, BID_VERTEX_INDEX, BID_LIGHTS };
static const char* bufNameOfId(int id)
{
    switch (id) {
    case BID_OUTPUT: return "outputBuffer"; case BID_LIGHT_VERTEX: return "lightVertexBuffer";
    case BID_LIGHT_VERTEX_INDEX: return "lightVertexBufferIndexBuffer"; case BID_VERTEX_COUNT: return "lightSubpathVertexCountBuffer";
    case BID_VERTEX_INDEX: return "lightSubpathVertexIndexBuffer"; case BID_LIGHTS: return "lights";
    }
    return "";
}
BufferDesc* bufferById(int id)
{
    static thread_local BufferDesc empty = {nullptr, 0, {0, 0, 0}};
    auto it = t_ctx->buffers.find(t_ctx->bufIntern[id & 7]);
    return it == t_ctx->buffers.end() ? &empty : &it->second;
}

float4 sampleTexture(const TextureData* t, float u, float v)
{
    if (!t) { float4 z = {0, 0, 0, 0}; return z; }
    ora::float4 r = ora::tex2D_raw(t->rgba, t->width, t->height, u, v);
    float4 o = {r.x, r.y, r.z, r.w};
    return o;
}

// Refresh the context tables from the oracle Renderer (the values OptixRenderer.cpp uploads: :555-557, :583-589,
// :616-617, :694-696, :714-715 and OptixRenderer_SpatialHash.cu:278-280).
static struct SceneTables& tablesFor(RefCtx& c);
static void refreshContext(RefCtx& c)
{
    ora::Renderer& r = *c.r;
    for (auto& kv : registry()) finalize(*kv.second);   // serial: pass_begin runs before the parallel launch
    for (int i = 1; i <= 6; i++) c.bufIntern[i] = intern(std::string("buf:") + bufNameOfId(i));
    c.tables = &tablesFor(c);
    static std::atomic<unsigned long long> g_epoch{0};   // globally unique: thread-local bind caches compare (ctx, epoch)
    c.epoch = ++g_epoch;
    if (c.rngEpoch != r.rngEpoch) {   // initializeRandomStates (OptixRenderer_SpatialHash.cu:310-347) on curandState records
        c.randomStates.resize(r.randomStates.size());
        for (size_t i = 0; i < c.randomStates.size(); i++) curand_init((unsigned)(r.seed + (unsigned)i), 0, 0, &c.randomStates[i]);
        c.rngEpoch = r.rngEpoch;
    }
    const size_t px = (size_t)r.width * r.height;
    c.dbgCells.resize(px); c.dbgPhotons.resize(px);
    c.vars.clear(); c.buffers.clear();
    setVar(c, "camera", &r.camera, sizeof r.camera);
    setVar(c, "localIterationNumber", &r.localIterationNumber, 4);
    setVar(c, "iterationNumber", &r.iterationNumber, 4);
    setVar(c, "ppmAlpha", &r.ppmAlpha, 4); setVar(c, "alpha", &r.ppmAlpha, 4);
    setVar(c, "ppmRadius", &r.ppmRadius, 4); setVar(c, "ppmRadiusSquared", &r.ppmRadiusSquared, 4);
    setVar(c, "emittedPhotonsPerIterationFloat", &r.emittedPhotonsPerIterationFloat, 4);
    setVar(c, "maxPhotonDepositsPerEmitted", &c.maxPhotonDepositsPerEmitted, 4);
    setVar(c, "photonLaunchWidth", &r.photonLaunchWidth, 4);
    setVar(c, "sceneBoundingSphere", &r.scene.boundingSphere, 16);
    c.photonsGridSize.x = r.grid.gridSize.x; c.photonsGridSize.y = r.grid.gridSize.y; c.photonsGridSize.z = r.grid.gridSize.z;
    c.photonsWorldOrigo.x = r.grid.origo.x; c.photonsWorldOrigo.y = r.grid.origo.y; c.photonsWorldOrigo.z = r.grid.origo.z;
    setVar(c, "photonsGridSize", &c.photonsGridSize, 12); setVar(c, "photonsWorldOrigo", &c.photonsWorldOrigo, 12);
    setVar(c, "photonsGridCellSize", &r.grid.cellSize, 4);
    setVar(c, "lightVertexCountEstimatePass", &r.lightVertexCountEstimatePass, 4);
    setVar(c, "maxPathLen", &r.maxPathLen, 4);
    setVar(c, "lightSubpathCount", &r.lightSubpathCount, 4);
    setVar(c, "misVcWeightFactor", &r.misVcWeightFactor, 4); setVar(c, "misVmWeightFactor", &r.misVmWeightFactor, 4);
    setVar(c, "vmNormalizationFactor", &r.vmNormalizationFactor, 4);
    setVar(c, "pixelSizeFactor", &r.pixelSizeFactor, 8);
    setVar(c, "vcmUseVC", &r.vcmUseVC, 4); setVar(c, "vcmUseVM", &r.vcmUseVM, 4);
    static int ptDirectLightSampling = 1;                                   // OptixRenderer.cpp:561
    setVar(c, "ptDirectLightSampling", &ptDirectLightSampling, 4);
    for (int i = 0; i < 8; i++) c.bufId[i] = i;
    setVar(c, "outputBufferId", &c.bufId[BID_OUTPUT], 4); setVar(c, "lightVertexBufferId", &c.bufId[BID_LIGHT_VERTEX], 4);
    setVar(c, "lightVertexBufferIndexBufferId", &c.bufId[BID_LIGHT_VERTEX_INDEX], 4);
    setVar(c, "lightSubpathVertexCountBufferId", &c.bufId[BID_VERTEX_COUNT], 4);
    setVar(c, "lightSubpathVertexIndexBufferId", &c.bufId[BID_VERTEX_INDEX], 4); setVar(c, "lightsBufferId", &c.bufId[BID_LIGHTS], 4);

    setBuf(c, "randomStates", c.randomStates.data(), sizeof(curandState), r.rsW, r.rsH);
    setBuf(c, "raytracePassOutputBuffer", r.hitpoints.data(), 64, r.width, r.height);
    setBuf(c, "photons", r.photons.data(), 36, r.photons.size());
    setBuf(c, "hashmapOffsetTable", r.offsets.data(), 4, r.offsets.size());
    setBuf(c, "indirectRadianceBuffer", r.indirect.data(), 12, r.width, r.height);
    setBuf(c, "directRadianceBuffer", r.direct.data(), 12, r.width, r.height);
    setBuf(c, "outputBuffer", r.output.data(), 12, r.width, r.height);
    setBuf(c, "lights", r.scene.lights.data(), 88, r.scene.lights.size());
    setBuf(c, "debugPhotonPathLengthBuffer", r.photonPathLength.data(), 4, r.photonLaunchWidth, r.photonLaunchHeight);
    setBuf(c, "debugIndirectRadianceCellsVisisted", c.dbgCells.data(), 4, r.width, r.height);
    setBuf(c, "debugIndirectRadiancePhotonsVisisted", c.dbgPhotons.data(), 4, r.width, r.height);
    setBuf(c, "lightSubpathVertexCountBuffer", r.lightSubpathVertexCount.data(), 4, r.lightSubpathVertexCount.size());
    setBuf(c, "lightVertexBuffer", c.lightVertexRaw.data(), c.lightVertexElem, c.lightVertexCount);
    setBuf(c, "lightVertexBufferIndexBuffer", &r.lightVertexBufferIndex, 4, 1);
    setBuf(c, "lightSubpathVertexIndexBuffer", r.lightSubpathVertexIndex.data(), 4, (size_t)r.width * r.height, r.vcmMaxPathLength - 1);
    c.textures.resize(r.scene.textures.size());
    for (size_t i = 0; i < c.textures.size(); i++) { c.textures[i].width = r.scene.textures[i].width; c.textures[i].height = r.scene.textures[i].height; c.textures[i].rgba = r.scene.textures[i].rgba.data(); }
}

// ---- variable binding --------------------------------------------------------------------------------------------
struct InstVar { int id; const void* ptr; size_t size; };
struct Instance { const InstVar* vars; int n; };

struct TraceState {
    optix::Ray ray; float tmax; float pendingT; bool shadow, terminated, found;
    int hitPrim; void* payload; const std::type_info* ptype;
    struct Attr { int id; size_t size; char bytes[16]; } attrs[8]; int nattrs;
    TURuntime* curGeomTU;
};
static thread_local TraceState* t_trace = nullptr;

static void bindContext(TURuntime& tu)
{
    RefCtx& c = *t_ctx;
    ThreadState& ts = tu.threadState();
    if (ts.ctx == &c && ts.epoch == c.epoch) return;
    for (const VarPlan& p : tu.plan) {
        if (p.kind != VK_CONTEXT) continue;
        auto it = c.vars.find(p.id);
        if (it != c.vars.end()) memcpy(p.var.get(), it->second.ptr, std::min(p.var.size, it->second.size));
    }
    for (const BufReg& b : tu.buffers) {
        auto it = c.buffers.find(b.id);
        BufferDesc empty = {nullptr, 0, {0, 0, 0}};
        b.bind(it == c.buffers.end() ? empty : it->second);
    }
    ts.ctx = &c; ts.epoch = c.epoch; ts.lastInstance = nullptr;
}

static void bindInstance(TURuntime& tu, const Instance& inst)
{
    for (const VarPlan& p : tu.plan) {
        if (p.kind != VK_CONTEXT) continue;
        for (int i = 0; i < inst.n; i++)
            if (inst.vars[i].id == p.id) { memcpy(p.var.get(), inst.vars[i].ptr, std::min(p.var.size, inst.vars[i].size)); break; }
    }
}

// Runs one RT_PROGRAM of `tu` with the OptiX semantic variables bound.  Variables of the unit are saved and
// restored around the call so that recursion through rtTrace into the same unit (Mirror -> Mirror) sees what
// OptiX guarantees: the outer program's ray / tHit / attributes are intact after rtTrace returns.
static void runProgram(TURuntime& tu, void (*fn)(), TraceState* st, const Instance* inst, void* payload, const std::type_info* ptype)
{
    bindContext(tu);
    char* save = (char*)alloca(tu.saveBytes());
    size_t off = 0;
    void* payloadVar = nullptr; size_t payloadSize = 0;
    for (const VarPlan& p : tu.plan) {
        void* v = p.var.get();
        if (p.kind == VK_PAYLOAD && ptype && !strcmp(p.var.type->name(), ptype->name())) { payloadVar = v; payloadSize = p.var.size; off += p.var.size; continue; }
        memcpy(save + off, v, p.var.size); off += p.var.size;
        switch (p.kind) {
        case VK_LAUNCH_INDEX: memcpy(v, &t_launchIndex, std::min(p.var.size, sizeof(uint2))); break;
        case VK_LAUNCH_DIM: memcpy(v, &t_launchDim, std::min(p.var.size, sizeof(uint2))); break;
        case VK_RAY: if (st) memcpy(v, &st->ray, std::min(p.var.size, sizeof(optix::Ray))); break;
        case VK_THIT: if (st) memcpy(v, &st->tmax, 4); break;
        case VK_ATTRIBUTE:
            if (st) for (int i = 0; i < st->nattrs; i++) if (st->attrs[i].id == p.id) { memcpy(v, st->attrs[i].bytes, std::min(p.var.size, st->attrs[i].size)); break; }
            break;
        default: break;
        }
    }
    if (inst) { bindInstance(tu, *inst); tu.threadState().lastInstance = nullptr; }
    if (payloadVar && payloadVar != payload) memcpy(payloadVar, payload, payloadSize);
    fn();
    if (payloadVar && payloadVar != payload) memcpy(payload, payloadVar, payloadSize);
    off = 0;
    for (const VarPlan& p : tu.plan) {
        if (p.kind == VK_PAYLOAD && ptype && !strcmp(p.var.type->name(), ptype->name())) { off += p.var.size; continue; }
        if (p.kind != VK_CONTEXT || inst) memcpy(p.var.get(), save + off, p.var.size);   // context values stay bound
        off += p.var.size;
    }
    if (inst) tu.threadState().epoch = ~0ull;   // instance values were restored over context values: re-bind next time
}

// ---- material / geometry instance tables ---------------------------------------------------------------------------
struct MaterialInst { InstVar vars[12]; int n; float3 powerPerArea, Lemit; unsigned hasNormals; TextureData* diffuse; TextureData* normal; };
struct SceneTables {
    std::vector<MaterialInst> mats;
    struct PgramInst { InstVar vars[5]; float4 plane; float3 v1, v2, anchor; int lgt; };
    std::vector<PgramInst> pgrams;
    struct SphereInst { InstVar vars[2]; float3 center; float radius; };
    std::vector<SphereInst> spheres;
    struct MeshInst { BufferDesc vertex, normal, tangent, bitangent, texcoord, index; unsigned hasTB; unsigned meshId; InstVar vars[2]; };
    std::vector<MeshInst> meshes;
    const void* sceneKey = nullptr; size_t primCount = 0; unsigned long long uid = 0;
};
static std::map<RefCtx*, SceneTables*> g_tables;

static SceneTables& tablesFor(RefCtx& c)
{
    std::lock_guard<std::mutex> l(g_ctxMutex);
    SceneTables*& t = g_tables[&c];
    ora::Scene& s = c.r->scene;
    if (t && t->uid == c.uid && t->sceneKey == (const void*)s.prims.data() && t->primCount == s.prims.size()) return *t;
    delete t; t = new SceneTables(); t->sceneKey = s.prims.data(); t->primCount = s.prims.size(); t->uid = c.uid;
    const int idKd = intern("Kd"), idKs = intern("Ks"), idExp = intern("exponent"), idKr = intern("Kr"), idKt = intern("Kt"), idIor = intern("indexOfRefraction"),
              idPpa = intern("powerPerArea"), idPower = intern("power"), idLemit = intern("Lemit"), idInvA = intern("inverseArea"), idHasN = intern("hasNormals");
    t->mats.resize(s.materials.size());
    for (size_t i = 0; i < s.materials.size(); i++) {
        const opr_material& m = s.materials[i]; MaterialInst& mi = t->mats[i];
        // DiffuseEmitter::registerGeometryInstanceValues (material/DiffuseEmitter.cpp:48-64)
        mi.powerPerArea.x = m.power[0] * m.inverseArea; mi.powerPerArea.y = m.power[1] * m.inverseArea; mi.powerPerArea.z = m.power[2] * m.inverseArea;
        mi.Lemit.x = m.power[0] * m.inverseArea * M_1_PIf; mi.Lemit.y = m.power[1] * m.inverseArea * M_1_PIf; mi.Lemit.z = m.power[2] * m.inverseArea * M_1_PIf;
        mi.hasNormals = m.normalTexture >= 0;
        mi.diffuse = mi.normal = nullptr;
        InstVar v[] = {{idKd, m.Kd, 12}, {idKs, m.Ks, 12}, {idExp, &m.exponent, 4}, {idKr, m.Kr, 12}, {idKt, m.Kt, 12}, {idIor, &m.indexOfRefraction, 4},
                       {idPpa, &mi.powerPerArea, 12}, {idPower, m.power, 12}, {idLemit, &mi.Lemit, 12}, {idInvA, &m.inverseArea, 4}, {idHasN, &mi.hasNormals, 4}};
        mi.n = 11; for (int k = 0; k < 11; k++) mi.vars[k] = v[k];
    }
    t->pgrams.resize(s.pgrams.size());
    for (size_t i = 0; i < s.pgrams.size(); i++) {
        auto& p = t->pgrams[i]; const ora::Parallelogram& g = s.pgrams[i];
        memcpy(&p.plane, &g.plane, 16); memcpy(&p.v1, &g.v1, 12); memcpy(&p.v2, &g.v2, 12); memcpy(&p.anchor, &g.anchor, 12); p.lgt = 0;
        InstVar v[] = {{intern("plane"), &p.plane, 16}, {intern("v1"), &p.v1, 12}, {intern("v2"), &p.v2, 12}, {intern("anchor"), &p.anchor, 12}, {intern("lgt_instance"), &p.lgt, 4}};
        for (int k = 0; k < 5; k++) p.vars[k] = v[k];
    }
    t->spheres.resize(s.spheres.size());
    for (size_t i = 0; i < s.spheres.size(); i++) {
        auto& p = t->spheres[i]; memcpy(&p.center, &s.spheres[i].center, 12); p.radius = s.spheres[i].radius;
        InstVar v[] = {{intern("center"), &p.center, 12}, {intern("radius"), &p.radius, 4}};
        for (int k = 0; k < 2; k++) p.vars[k] = v[k];
    }
    t->meshes.resize(s.meshes.size());
    for (size_t i = 0; i < s.meshes.size(); i++) {
        auto& m = t->meshes[i]; ora::Mesh& src = s.meshes[i];
        m.vertex = BufferDesc{src.vertices.data(), 12, {src.vertices.size(), 1, 1}};
        m.normal = BufferDesc{src.normals.data(), 12, {src.normals.size(), 1, 1}};
        m.tangent = BufferDesc{src.tangents.data(), 12, {src.tangents.size(), 1, 1}};
        m.bitangent = BufferDesc{src.bitangents.data(), 12, {src.bitangents.size(), 1, 1}};
        m.texcoord = BufferDesc{src.uvs.data(), 8, {src.uvs.size(), 1, 1}};
        m.index = BufferDesc{src.indices.data(), 12, {src.indices.size(), 1, 1}};
        m.hasTB = !src.tangents.empty() && !src.bitangents.empty(); m.meshId = (unsigned)i;
        InstVar v[] = {{intern("hasTangentsAndBitangents"), &m.hasTB, 4}, {intern("meshId"), &m.meshId, 4}};
        for (int k = 0; k < 2; k++) m.vars[k] = v[k];
    }
    return *t;
}

// ---- intersection API --------------------------------------------------------------------------------------------
bool potentialIntersection(float t)
{
    TraceState* st = t_trace;
    if (t > st->ray.tmin && t < st->tmax) { st->pendingT = t; return true; }
    return false;
}

static void (*g_anyHitShadow)() = nullptr;   // gatherAnyHitOnNonEmitter: bound to every material (material/Material.cpp:18-26)
static TURuntime* g_anyHitTU = nullptr;

bool reportIntersection(unsigned)
{
    TraceState* st = t_trace;
    if (st->shadow) {
        TraceState hitState = *st; hitState.tmax = st->pendingT;
        runProgram(*g_anyHitTU, g_anyHitShadow, &hitState, nullptr, st->payload, st->ptype);
        return true;   // the program called rtTerminateRay()
    }
    st->tmax = st->pendingT; st->found = true;
    st->nattrs = 0;
    for (const VarPlan& p : st->curGeomTU->plan)
        if (p.kind == VK_ATTRIBUTE && st->nattrs < 8) {
            TraceState::Attr& a = st->attrs[st->nattrs++];
            a.id = p.id; a.size = std::min<size_t>(p.var.size, 16); memcpy(a.bytes, p.var.get(), a.size);
        }
    return true;
}
void terminateRay() { t_trace->terminated = true; }
void ignoreIntersection() {}

// ---- rtTrace --------------------------------------------------------------------------------------------------------
enum { RT_RADIANCE = 0, RT_PHOTON = 1, RT_CAMERA_VCM = 2, RT_LIGHT_VCM = 3, RT_RADIANCE_PM = 4, RT_PHOTON_PM = 5, RT_SHADOW = 7 };   // renderer/RayType.h

static const char* materialTU(int kind)
{
    switch (kind) {
    case OPR_MAT_DIFFUSE: return "Diffuse"; case OPR_MAT_DIFFUSE_EMITTER: return "DiffuseEmitter"; case OPR_MAT_GLASS: return "Glass";
    case OPR_MAT_MIRROR: return "Mirror"; case OPR_MAT_TEXTURE: return "Texture"; case OPR_MAT_GLOSSY: return "Glossy";
    }
    return "Diffuse";
}
static const char* programFor(unsigned rayType)
{
    switch (rayType) {
    case RT_RADIANCE: case RT_RADIANCE_PM: return "closestHitRadiance";
    case RT_PHOTON: case RT_PHOTON_PM: return "closestHitPhoton";
    case RT_LIGHT_VCM: return "vcmClosestHitLight";
    case RT_CAMERA_VCM: return "vcmClosestHitCamera";
    }
    return nullptr;
}

void trace(const optix::Ray& ray, void* payload, const std::type_info& ptype)
{
    RefCtx& c = *t_ctx;
    ora::Renderer& r = *c.r;
    SceneTables& tabs = *c.tables;
    TraceState st; st.ray = ray; st.tmax = ray.tmax; st.pendingT = 0.f; st.shadow = (ray.ray_type == RT_SHADOW); st.terminated = false; st.found = false;
    st.hitPrim = -1; st.payload = payload; st.ptype = &ptype; st.nattrs = 0; st.curGeomTU = nullptr;
    TraceState* prev = t_trace; t_trace = &st;
    if (st.shadow) r.stats.any.fetch_add(1, std::memory_order_relaxed); else r.stats.closest.fetch_add(1, std::memory_order_relaxed);

    static TURuntime& tuTri = getTU("TriangleMesh"); static TURuntime& tuPg = getTU("parallelogram"); static TURuntime& tuSp = getTU("Sphere");
    static void (*fnTri)(int) = tuTri.intProgram("mesh_intersect"); static void (*fnPg)(int) = tuPg.intProgram("intersect"); static void (*fnSp)(int) = tuSp.intProgram("intersect");
    if (!g_anyHitTU) { g_anyHitTU = &getTU("DirectRadianceEstimation"); g_anyHitShadow = g_anyHitTU->program("gatherAnyHitOnNonEmitter"); }
    bool rayBound[3] = {false, false, false};
    ora::Ray oray; memcpy(&oray.origin, &ray.origin, 12); memcpy(&oray.direction, &ray.direction, 12); oray.ray_type = ray.ray_type; oray.tmin = ray.tmin; oray.tmax = ray.tmax;

    r.scene.forEachCandidate(oray, st.tmax, [&](int pi, float& tmaxRef) -> bool {
        const ora::Prim& p = r.scene.prims[pi];
        TURuntime& tu = p.type == ora::PRIM_TRIANGLE ? tuTri : (p.type == ora::PRIM_PARALLELOGRAM ? tuPg : tuSp);
        bindContext(tu);
        ThreadState& ts = tu.threadState();
        if (!rayBound[p.type]) {
            for (const VarPlan& vp : tu.plan) if (vp.kind == VK_RAY) memcpy(vp.var.get(), &st.ray, sizeof(optix::Ray));
            rayBound[p.type] = true;
        }
        const void* instKey;
        if (p.type == ora::PRIM_TRIANGLE) {
            auto& m = tabs.meshes[p.a]; instKey = &m;
            if (ts.lastInstance != instKey) {
                Instance inst = {m.vars, 2}; bindInstance(tu, inst);
                for (const BufReg& b : tu.buffers) {
                    std::string n = b.name;
                    if (n == "vertexBuffer") b.bind(m.vertex); else if (n == "normalBuffer") b.bind(m.normal); else if (n == "tangentBuffer") b.bind(m.tangent);
                    else if (n == "bitangentBuffer") b.bind(m.bitangent); else if (n == "texCoordBuffer") b.bind(m.texcoord); else if (n == "indexBuffer") b.bind(m.index);
                }
                ts.lastInstance = instKey;
            }
        } else if (p.type == ora::PRIM_PARALLELOGRAM) {
            auto& g = tabs.pgrams[p.a]; instKey = &g;
            if (ts.lastInstance != instKey) { Instance inst = {g.vars, 5}; bindInstance(tu, inst); ts.lastInstance = instKey; }
        } else {
            auto& g = tabs.spheres[p.a]; instKey = &g;
            if (ts.lastInstance != instKey) { Instance inst = {g.vars, 2}; bindInstance(tu, inst); ts.lastInstance = instKey; }
        }
        st.curGeomTU = &tu;
        const bool before = st.found; const float tBefore = st.tmax;
        (p.type == ora::PRIM_TRIANGLE ? fnTri : (p.type == ora::PRIM_PARALLELOGRAM ? fnPg : fnSp))(p.type == ora::PRIM_TRIANGLE ? p.b : 0);
        if (st.found && (!before || st.tmax != tBefore)) st.hitPrim = pi;
        tmaxRef = st.tmax;
        return st.terminated;
    });

    if (!st.shadow) {
        if (st.found) {
            const ora::Prim& p = r.scene.prims[st.hitPrim];
            int mat = p.type == ora::PRIM_TRIANGLE ? r.scene.meshes[p.a].material : (p.type == ora::PRIM_PARALLELOGRAM ? r.scene.pgrams[p.a].material : r.scene.spheres[p.a].material);
            const opr_material& m = r.scene.materials[mat];
            const char* prog = programFor(ray.ray_type);
            TURuntime& tu = getTU(materialTU(m.kind));
            void (*fn)() = prog ? tu.programOrNull(prog) : nullptr;
            if (fn) {
                MaterialInst& mi = tabs.mats[mat];
                Instance inst = {mi.vars, mi.n};
                if (m.kind == OPR_MAT_TEXTURE) {
                    for (const SamplerReg& s : tu.samplers) {
                        int ti = std::string(s.name) == "diffuseSampler" ? m.diffuseTexture : m.normalTexture;
                        s.bind(ti >= 0 ? &c.textures[ti] : nullptr);
                    }
                }
                runProgram(tu, fn, &st, &inst, payload, &ptype);
            }
        } else {
            const char* tuName = nullptr;
            switch (ray.ray_type) {
            case RT_RADIANCE: case RT_RADIANCE_PM: tuName = "RayGeneratorPPM"; break;   // OptixRenderer.cpp:170-171
            case RT_PHOTON: case RT_PHOTON_PM: tuName = "PhotonGenerator"; break;       // :180
            case RT_LIGHT_VCM: tuName = "VCMLightPass"; break;                          // :356
            case RT_CAMERA_VCM: tuName = "VCMCameraPass"; break;                        // :365
            }
            if (tuName) { TURuntime& tu = getTU(tuName); runProgram(tu, tu.program("miss"), &st, nullptr, payload, &ptype); }
        }
    }
    t_trace = prev;
}

static void launchOne(ora::Renderer& r, const char* tuName, const char* prog, unsigned x, unsigned y, unsigned w, unsigned h)
{
    RefCtx& c = ctxFor(r);
    t_ctx = &c;
    t_launchIndex.x = x; t_launchIndex.y = y; t_launchDim.x = w; t_launchDim.y = h;
    TURuntime& tu = getTU(tuName);
    runProgram(tu, tu.program(prog), nullptr, nullptr, nullptr, nullptr);
}

} // namespace shim

// ---- the program entry points the shared orchestration (oracle/ora_render.cpp) launches ---------------------------------
namespace ora {
using shim::launchOne;

void pass_begin(Renderer& r) { shim::refreshContext(shim::ctxFor(r)); }
}  // namespace ora

namespace ora {
void ppm_generateRay(Renderer& r, uint x, uint y) { launchOne(r, "RayGeneratorPPM", "generateRay", x, y, r.width, r.height); }
void ppm_photonGenerator(Renderer& r, uint x, uint y) { launchOne(r, "PhotonGenerator", "generator", x, y, r.photonLaunchWidth, r.photonLaunchHeight); }
void ppm_indirect(Renderer& r, uint x, uint y) { launchOne(r, "IndirectRadianceEstimation", "kernel", x, y, r.width, r.height); }
void ppm_direct(Renderer& r, uint x, uint y) { launchOne(r, "DirectRadianceEstimation", "kernel", x, y, r.width, r.height); }
void ppm_output(Renderer& r, uint x, uint y) { launchOne(r, "Output", "kernel", x, y, r.width, r.height); }
void pt_generateRay(Renderer& r, uint x, uint y) { launchOne(r, "RayGeneratorPT", "generateRay", x, y, r.width, r.height); }
void vcm_lightPass(Renderer& r, uint x, uint y) { launchOne(r, "VCMLightPass", "lightPass", x, y, r.width, r.height); }
void vcm_cameraPass(Renderer& r, uint x, uint y) { launchOne(r, "VCMCameraPass", "cameraPass", x, y, r.width, r.height); }
void vcm_alloc_vertices(Renderer& r, uint n)
{
    shim::RefCtx& c = shim::ctxFor(r);
    c.lightVertexElem = shim::getTU("Diffuse").bufferElemSize("lightVertexBuffer");
    c.lightVertexCount = n;
    // the reference never checks the cursor against the buffer size (vcm.h:281-282); reserve the worst case so that an
    // overflowing vertex is stored instead of corrupting memory
    const size_t cap = std::max<size_t>(n, (size_t)r.width * r.height * (r.vcmMaxPathLength - 1));
    c.lightVertexRaw.assign(std::max<size_t>(cap, 1) * c.lightVertexElem, 0);
    r.lightVertexBufferSize = n;
}
void vcm_free(Renderer&) {}
}  // namespace ora
