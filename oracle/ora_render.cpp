// ORACLE — test infrastructure only (see ora_math.h).  Host orchestration restated from
// RenderEngine/renderer/OptixRenderer.cpp:113-400 (initialize), :436-485 (initScene), :507-821
// (renderNextIteration), :828-848 (resizeBuffers), plus the flat C API the tests/bench use via ctypes.
#include "ora_render.h"
#include <cstring>
#include <stdexcept>
#include <omp.h>

namespace ora {

static std::atomic<unsigned long long> g_nextUid{1};
Renderer::Renderer() { uid = g_nextUid.fetch_add(1); }
Renderer::~Renderer() { vcm_free(*this); }

void Renderer::initialize()
{
    if (initialized) throw std::runtime_error("ERROR: Multiple OptixRenderer::initialize!\n");   // :115-118
    photons.assign((size_t)photonLaunchWidth * photonLaunchHeight * MAX_PHOTON_DEPOSITS, Photon{});
    photonPathLength.assign((size_t)photonLaunchWidth * photonLaunchHeight, 0u);
    emittedPhotonsPerIterationFloat = float(photonLaunchWidth * photonLaunchHeight);
    initialized = true;
}

void Renderer::initScene(const opr_scene_desc* d)
{
    if (!initialized) throw std::runtime_error("Cannot initialize scene before OptixRenderer.");   // :438-441
    if (d->numLights == 0) throw std::runtime_error("No lights exists in this scene.");            // :443-447
    lightVertexCountEstimated = false;
    scene.load(d);
    sceneLoaded = true;
}

void Renderer::initializeRandomStates()
{   // OptixRenderer_SpatialHash.cu:310-347
    const size_t num = (size_t)rsW * rsH;
    rngEpoch++;
    randomStates.resize(num);
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)num; i++) rng_init(&randomStates[i], seed, (unsigned)i);
}

void Renderer::resizeBuffers(uint w, uint h)
{   // OptixRenderer.cpp:828-848
    output.assign((size_t)w * h, make_float3(0.f));
    hitpoints.assign((size_t)w * h, Hitpoint{});
    direct.assign((size_t)w * h, make_float3(0.f));
    indirect.assign((size_t)w * h, make_float3(0.f));
    rsW = std::max(photonLaunchWidth, w); rsH = std::max(photonLaunchHeight, h);
    initializeRandomStates();
    width = w; height = h;
    lightSubpathCount = w * h;
    lightSubpathVertexIndex.assign((size_t)w * h * (vcmMaxPathLength - 1), 0u);
    pixelSizeFactor = make_float2(1.0f / w, 1.0f / h);
    lightVertexCountEstimated = false;
}

template <class F> static void launch2D(Renderer& r, uint w, uint h, F f)
{
    const long long n = (long long)w * h;
    pass_begin(r);
    #pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < n; i++) f(r, (uint)(i % w), (uint)(i / w));
}

void Renderer::renderNextIteration(unsigned long long iterationNumber_, unsigned long long localIterationNumber_,
                                   float PPMRadius, bool /*createOutput: ignored by the reference*/, const opr_request_details& details)
{
    if (!initialized) throw std::runtime_error("Traced before OptixRenderer was initialized.");   // :513-516
    if (!sceneLoaded) throw std::runtime_error("Traced before a scene was initialized.");
    if (numThreads > 0) omp_set_num_threads(numThreads);
    if (details.width != width || details.height != height) resizeBuffers(details.width, details.height);   // :534-537
    if (localIterationNumber_ == 0) std::fill(output.begin(), output.end(), make_float3(0.f));               // :540-545

    camera = details.camera;
    iterationNumber = static_cast<float>(iterationNumber_);
    localIterationNumber = (unsigned)localIterationNumber_;

    if (details.renderMethod == OPR_PROGRESSIVE_PHOTON_MAPPING) {
        launch2D(*this, width, height, ppm_generateRay);                         // :576
        ppmAlpha = details.ppmAlpha;
        ppmRadius = PPMRadius;
        ppmRadiusSquared = PPMRadius * PPMRadius;                               // :585
        launch2D(*this, photonLaunchWidth, photonLaunchHeight, ppm_photonGenerator);   // :612
        photonsAsEmitted = photons;
        if (!build_photon_grid(photons.data(), (uint)photons.size(), photonKeys, offsets, grid))   // :628
            throw std::runtime_error("Too many cells in SpatialHash.cu, over defined PHOTON_GRID_MAX_SIZE.");
        launch2D(*this, width, height, ppm_indirect);                            // :657
        directSampleMask.assign((size_t)width * height, 0);
        launch2D(*this, width, height, ppm_direct);                              // :664
        launch2D(*this, width, height, ppm_output);                              // :670
    } else if (details.renderMethod == OPR_VCM_BIDIRECTIONAL_PATH_TRACING) {
        const unsigned lightSubPathCount = width * height;                       // :677-696
        const unsigned lightSubpathsConnected = 1;
        const unsigned lightSubpathsMerged = lightSubPathCount;
        const float ppmRadiusSquared_ = PPMRadius * PPMRadius;
        vmNormalizationFactor = 1.f / (ppmRadiusSquared_ * M_PIf_ * lightSubPathCount);
        const float etaVCM = (float(lightSubpathsMerged) / lightSubpathsConnected) * M_PIf_ * ppmRadiusSquared_;
        misVmWeightFactor = vcmUseVM ? etaVCM : 0.f;
        misVcWeightFactor = vcmUseVC ? (1.f / etaVCM) : 0.f;
        lightSubpathCount = lightSubPathCount;

        if (!lightVertexCountEstimated) {                                        // :699-763
            lightSubpathVertexCount.assign((size_t)width * height, 0u);
            maxPathLen = vcmMaxPathLength;
            lightVertexCountEstimatePass = 1u;
            launch2D(*this, width, height, vcm_lightPass);
            unsigned long long sumStoredVerts = 0;
            const unsigned subpathEstimateCount = width * height;
            for (unsigned i = 0; i < subpathEstimateCount; i++) sumStoredVerts += lightSubpathVertexCount[i];
            float avgSubpathVerts = float(sumStoredVerts) / subpathEstimateCount;
            avgSubpathVerts = std::min(avgSubpathVerts / 0.5f, float(vcmMaxPathLength - 1));
            const unsigned vertBufSize = (unsigned)(lightSubPathCount * avgSubpathVerts);
            vcm_alloc_vertices(*this, vertBufSize);
            lightVertexCountEstimated = true;
            lightVertexCountEstimatePass = 0u;
        }
        lightVertexBufferIndex.store(0);                                         // :776-778
        connectionCount.assign((size_t)width * height, 0u);
        pathSignature.assign((size_t)width * height * 2, 0u);
        launch2D(*this, width, height, vcm_lightPass);                           // :784
        launch2D(*this, width, height, vcm_cameraPass);                          // :793
    } else if (details.renderMethod == OPR_PATH_TRACING) {
        ptSignature.assign((size_t)width * height, 0);
        launch2D(*this, width, height, pt_generateRay);                          // :559-567
    } else {
        throw std::runtime_error("unknown render method");
    }
}

} // namespace ora

// ---------------------------------------------------------------------------------------------------
// Flat C API (ctypes).  Every entry returns 0 on success, non-zero + ora_last_error() otherwise.
using namespace ora;
static thread_local std::string g_err;
#define ORA_TRY(body) try { body; return 0; } catch (const std::exception& e) { g_err = e.what(); return 1; }

extern "C" {

const char* ora_last_error() { return g_err.c_str(); }

void* ora_create(const opr_config* cfg)
{
    Renderer* r = new Renderer();
    if (cfg) {
        if (cfg->photonLaunchWidth) r->photonLaunchWidth = cfg->photonLaunchWidth;
        if (cfg->photonLaunchHeight) r->photonLaunchHeight = cfg->photonLaunchHeight;
        if (cfg->vcmMaxPathLength) r->vcmMaxPathLength = cfg->vcmMaxPathLength;
        if (cfg->seedIsSet) r->seed = cfg->randomSeed;
    }
    return r;
}
void ora_destroy(void* h) { delete (Renderer*)h; }
int ora_initialize(void* h) { ORA_TRY(((Renderer*)h)->initialize()) }
int ora_init_scene(void* h, const opr_scene_desc* d) { ORA_TRY(((Renderer*)h)->initScene(d)) }
int ora_set_threads(void* h, int n) { ((Renderer*)h)->numThreads = n; return 0; }
int ora_render_next_iteration(void* h, unsigned long long it, unsigned long long lit, float radius, int createOutput, const opr_request_details* d)
{ ORA_TRY(((Renderer*)h)->renderNextIteration(it, lit, radius, createOutput != 0, *d)) }
int ora_get_output_buffer(void* h, void* dst)
{ Renderer* r = (Renderer*)h; memcpy(dst, r->output.data(), sizeof(float3) * r->output.size()); return 0; }
unsigned ora_get_width(void* h) { return ((Renderer*)h)->width; }
unsigned ora_get_height(void* h) { return ((Renderer*)h)->height; }

// intermediate buffers (reference layouts) for stage-by-stage parity
int ora_get_hitpoints(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->hitpoints.data(), sizeof(Hitpoint) * r->hitpoints.size()); return 0; }
int ora_get_photons(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->photons.data(), sizeof(Photon) * r->photons.size()); return 0; }
int ora_get_photons_as_emitted(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->photonsAsEmitted.data(), sizeof(Photon) * r->photonsAsEmitted.size()); return 0; }
int ora_get_photon_keys(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->photonKeys.data(), 4 * r->photonKeys.size()); return 0; }
int ora_get_offsets(void* h, void* dst, unsigned n) { Renderer* r = (Renderer*)h; memcpy(dst, r->offsets.data(), 4 * std::min<size_t>(n, r->offsets.size())); return 0; }
int ora_get_grid(void* h, float* cellSize, unsigned* gridSize3, float* origo3, unsigned* numValid)
{ Renderer* r = (Renderer*)h; *cellSize = r->grid.cellSize; gridSize3[0] = r->grid.gridSize.x; gridSize3[1] = r->grid.gridSize.y; gridSize3[2] = r->grid.gridSize.z;
  origo3[0] = r->grid.origo.x; origo3[1] = r->grid.origo.y; origo3[2] = r->grid.origo.z; *numValid = r->grid.numValid; return 0; }
int ora_get_indirect(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->indirect.data(), sizeof(float3) * r->indirect.size()); return 0; }
int ora_get_direct(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->direct.data(), sizeof(float3) * r->direct.size()); return 0; }
int ora_get_random_states(void* h, void* dst, unsigned long long first, unsigned long long count)
{ Renderer* r = (Renderer*)h; memcpy(dst, r->randomStates.data() + first, sizeof(RandomState) * count); return 0; }
unsigned long long ora_num_random_states(void* h) { return ((Renderer*)h)->randomStates.size(); }
int ora_get_ray_counts(void* h, unsigned long long* closest, unsigned long long* any)
{ Renderer* r = (Renderer*)h; *closest = r->stats.closest.load(); *any = r->stats.any.load(); return 0; }
int ora_get_photon_path_lengths(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->photonPathLength.data(), 4 * r->photonPathLength.size()); return 0; }
int ora_get_direct_sample_mask(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->directSampleMask.data(), r->directSampleMask.size()); return 0; }
int ora_get_pt_signature(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->ptSignature.data(), 4 * r->ptSignature.size()); return 0; }
int ora_get_path_signatures(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->pathSignature.data(), 4 * r->pathSignature.size()); return 0; }
int ora_get_connection_counts(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->connectionCount.data(), 4 * r->connectionCount.size()); return 0; }
int ora_get_light_vertex_counts(void* h, void* dst) { Renderer* r = (Renderer*)h; memcpy(dst, r->lightSubpathVertexCount.data(), 4 * r->lightSubpathVertexCount.size()); return 0; }

// stand-alone pieces -----------------------------------------------------------------------------------
// RNG: states = 6 uint32 each {d, v0..v4}
void ora_rng_init(unsigned* states6, unsigned seed, unsigned first, unsigned count)
{ for (unsigned i = 0; i < count; i++) rng_init((RandomState*)(states6 + 6 * (size_t)i), seed, first + i); }
void ora_rng_uniform(unsigned* state6, float* out, unsigned n)
{ for (unsigned i = 0; i < n; i++) out[i] = getRandomUniformFloat((RandomState*)state6); }

// photon grid on caller-supplied photons (AoS 36 B, reference layout); photons are sorted in place.
// keys: n entries; offsets: capacity PHOTON_GRID_MAX_SIZE+2.  Returns 2 if the grid would exceed 10^6 cells.
int ora_build_photon_grid(float* photonsAoS, unsigned n, unsigned* keys, unsigned* offsets,
                          float* cellSize, unsigned* gridSize3, float* origo3, unsigned* numValid)
{
    std::vector<uint> k, o; PhotonGrid g;
    bool ok = build_photon_grid((Photon*)photonsAoS, n, k, o, g);
    *cellSize = g.cellSize; gridSize3[0] = g.gridSize.x; gridSize3[1] = g.gridSize.y; gridSize3[2] = g.gridSize.z;
    origo3[0] = g.origo.x; origo3[1] = g.origo.y; origo3[2] = g.origo.z; *numValid = g.numValid;
    if (!ok) return 2;
    memcpy(keys, k.data(), 4 * (size_t)n);
    memcpy(offsets, o.data(), 4 * o.size());
    return 0;
}

// IndirectRadianceEstimation::kernel on caller-supplied hitpoints (64 B each) + sorted photons (36 B each) + offset table
int ora_gather(unsigned numPixels, const float* hitpointsAoS, unsigned numPhotons, const float* photonsAoS, const unsigned* offsets,
               float cellSize, const unsigned* gridSize3, const float* origo3, float ppmRadius, float emittedPhotons, float* indirectOut)
{
    Renderer r;
    r.width = numPixels; r.height = 1;
    r.hitpoints.assign((const Hitpoint*)hitpointsAoS, (const Hitpoint*)hitpointsAoS + numPixels);
    r.photons.assign((const Photon*)photonsAoS, (const Photon*)photonsAoS + numPhotons);
    r.grid.cellSize = cellSize; r.grid.gridSize = {gridSize3[0], gridSize3[1], gridSize3[2]}; r.grid.origo = f3(origo3);
    r.grid.numCells = gridSize3[0] * gridSize3[1] * gridSize3[2];
    r.offsets.assign(offsets, offsets + r.grid.numCells + 1);
    r.indirect.assign(numPixels, make_float3(0.f));
    r.direct.assign(numPixels, make_float3(0.f)); r.output.assign(numPixels, make_float3(0.f));
    r.ppmRadius = ppmRadius; r.ppmRadiusSquared = ppmRadius * ppmRadius; r.emittedPhotonsPerIterationFloat = emittedPhotons;
    r.rsW = 1; r.rsH = 1; r.randomStates.resize(1);
    pass_begin(r);
    #pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)numPixels; i++) ppm_indirect(r, (uint)i, 0);
    memcpy(indirectOut, r.indirect.data(), sizeof(float3) * numPixels);
    return 0;
}

// closest-hit / any-hit queries against the loaded scene.  rays: n x 8 floats (o.xyz,tmin,d.xyz,tmax);
// out hits: n x {t (float), prim-kind (int), material (int), geometric normal xyz}
int ora_trace_closest(void* h, const float* rays8, unsigned n, float* outT, int* outMaterial, float* outNg3)
{
    Renderer* r = (Renderer*)h;
    #pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; i++) {
        const float* p = rays8 + 8 * i;
        Ray ray{{p[0], p[1], p[2]}, {p[4], p[5], p[6]}, 0, p[3], p[7]};
        Hit hit;
        if (r->scene.closestHit(ray, hit)) {
            outT[i] = hit.t; outMaterial[i] = hit.material;
            outNg3[3 * i] = hit.geometricNormal.x; outNg3[3 * i + 1] = hit.geometricNormal.y; outNg3[3 * i + 2] = hit.geometricNormal.z;
        } else { outT[i] = -1.f; outMaterial[i] = -1; outNg3[3 * i] = outNg3[3 * i + 1] = outNg3[3 * i + 2] = 0.f; }
    }
    return 0;
}
int ora_trace_any(void* h, const float* rays8, unsigned n, int* outOccluded)
{
    Renderer* r = (Renderer*)h;
    #pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; i++) {
        const float* p = rays8 + 8 * i;
        Ray ray{{p[0], p[1], p[2]}, {p[4], p[5], p[6]}, 7, p[3], p[7]};
        outOccluded[i] = r->scene.anyHit(ray) ? 1 : 0;
    }
    return 0;
}

// host helpers exposed so that the Python mirror of Camera/Light/IScene can be checked against them
void ora_camera_setup(opr_camera* c) { camera_setup(*c); }
void ora_camera_set_aspect_ratio(opr_camera* c, float ratio) { camera_set_aspect_ratio(*c, ratio); }
void ora_make_area_light(opr_light* out, const float* power, const float* position, const float* v1, const float* v2)
{ *out = make_area_light(f3(power), f3(position), f3(v1), f3(v2)); }
void ora_make_point_light(opr_light* out, const float* power, const float* position)
{ *out = make_point_light(f3(power), f3(position)); }
float ora_initial_ppm_radius(const float* mn, const float* mx) { return initial_ppm_radius(f3(mn), f3(mx)); }
void ora_bounding_sphere(const float* mn, const float* mx, float* out4)
{ Sphere s = aabb_bounding_sphere(f3(mn), f3(mx)); out4[0] = s.center.x; out4[1] = s.center.y; out4[2] = s.center.z; out4[3] = s.radius; }
int ora_max_threads() { return omp_get_max_threads(); }

} // extern "C"
