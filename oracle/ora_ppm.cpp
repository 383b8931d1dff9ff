// ORACLE — test infrastructure only (see ora_math.h).  Progressive-photon-mapping device programs of
// the reference restated on the CPU: ray generators (renderer/ppm/*.cu), the material closest-hit
// programs for RADIANCE and PHOTON rays (material/*.cu) and the helpers they call
// (renderer/helpers/{samplers,light,camera}.h).  Recursion mirrors the reference's rtTrace recursion.
#include "ora_render.h"

namespace ora {

// ---- helpers/helpers.h:119-165 -----------------------------------------------------------------------
static inline void createCoordinateSystem(const float3& N, float3& U, float3& V)
{
    if (fabsf(N.x) > fabsf(N.y)) {
        float invLength = 1.f / sqrtf(N.x * N.x + N.z * N.z);
        U = make_float3(-N.z * invLength, 0.f, N.x * invLength);
    } else {
        float invLength = 1.f / sqrtf(N.y * N.y + N.z * N.z);
        U = make_float3(0.f, N.z * invLength, -N.y * invLength);
    }
    V = cross(N, U);
}
static inline float maxf(float a, float b) { return a > b ? a : b; }
static inline int intmin(int a, int b) { return a < b ? a : b; }
static inline float favgf(const float3& v) { return (v.x + v.y + v.z) * 0.3333333333f; }
static inline bool hitFromOutside(const float3& rayDirection, const float3& normal) { return dot(normal, rayDirection) < 0; }

// ---- helpers/samplers.h:21-95 ------------------------------------------------------------------------
float3 sampleUnitHemisphereCos(const float3& normal, const float2& sample, float* oPdfW, float* oCosTheta, bool biasSmallCosine)
{
    float theta = acosf(sqrtf(sample.x));
    float phi = 2.0f * M_PIf_ * sample.y;
    float xs = sinf(theta) * cosf(phi);
    float ys = cosf(theta);
    float zs = sinf(theta) * sinf(phi);
    float3 U, V;
    createCoordinateSystem(normal, U, V);
    if (biasSmallCosine && ys < EPS_COSINE_) ys = EPS_COSINE_;
    if (oPdfW) *oPdfW = ys * M_1_PIf_;
    if (oCosTheta) *oCosTheta = ys;
    return normalize(xs * U + ys * normal + zs * V);
}

static float3 sampleUnitHemisphere(const float3& normal, const float2& sample)
{
    float3 U, V;
    createCoordinateSystem(normal, U, V);
    float phi = 2.0f * M_PIf_ * sample.x;
    float r = sqrtf(sample.y);
    float x = r * cosf(phi);
    float y = r * sinf(phi);
    float z = 1.0f - x * x - y * y;
    z = z > 0.0f ? sqrtf(z) : 0.0f;
    return normalize(U * x + V * y + normal * z);
}

float3 sampleUnitSphere(const float2& sample, float* oPdfW)
{
    float3 v;
    v.z = 1.f - 2.f * sample.x;
    float phi = 2 * M_PIf_ * sample.y;
    float r = sqrtf(1.f - v.z * v.z);
    v.x = r * cosf(phi);
    v.y = r * sinf(phi);
    if (oPdfW) *oPdfW = 0.25f * M_1_PIf_;
    return v;
}

static float2 sampleUnitDisc(const float2& sample)
{
    float r = sqrtf(sample.x);
    float theta = 2.f * M_PIf_ * sample.y;
    return make_float2(r * cosf(theta), r * sinf(theta));
}

static float3 sampleDisc(const float2& sample, const float3& center, const float radius, const float3& normal)
{
    float3 U, V;
    createCoordinateSystem(normal, U, V);
    float2 unitDisc = sampleUnitDisc(sample);
    return center + radius * (U * unitDisc.x + V * unitDisc.y);
}

float4 tex2D(const Texture& t, float u, float v);   // ora_tex.cpp
void pass_begin(Renderer&) {}                         // hook used by the oracle/_ref build only

// ---- payloads ------------------------------------------------------------------------------------------
struct RadiancePRD {   // renderer/RadiancePRD.h:11-29
    float3 attenuation, radiance; uint depth; float3 position, normal; uint flags;
    RandomState randomState; float lastTHit;
    float3 randomNewDirection;   // written only under PRD_PATH_TRACING
};
struct PhotonPRD {     // renderer/ppm/PhotonPRD.h:9-17
    float3 power; float weight; uint pm_index, numStoredPhotons, depth; RandomState randomState;
};

static void traceRadiance(Renderer& r, const Ray& ray, RadiancePRD& prd);
static void tracePhoton(Renderer& r, const Ray& ray, PhotonPRD& prd);

// material/Texture.cu:72-80
static float3 getNormalMappedNormal(const float3& normal, const float3& tangent, const float3& bitangent, const float4& normalMap)
{
    float4 nMap = 2 * normalMap - 1;
    float3 N;
    N.x = nMap.x * tangent.x + nMap.y * bitangent.x + nMap.z * normal.x;
    N.y = nMap.x * tangent.y + nMap.y * bitangent.y + nMap.z * normal.y;
    N.z = nMap.x * tangent.z + nMap.y * bitangent.z + nMap.z * normal.z;
    return normalize(N);
}

// material/Glass.cu:69-88
static inline float reflectionFactor(float cosI, float cosT, float n1, float n2)
{
    float rp = (n2 * cosI - n1 * cosT) / (n2 * cosI + n1 * cosT);
    float rs = (n1 * cosI - n2 * cosT) / (n1 * cosI + n2 * cosT);
    return (rp * rp + rs * rs) / 2.f;
}

// ---- RADIANCE closest-hit programs ----------------------------------------------------------------------
static void closestHitRadiance(Renderer& r, const Ray& ray, const Hit& hit, RadiancePRD& prd)
{
    const opr_material& m = r.scene.materials[hit.material];
    const float tHit = hit.t;
    // rtTransformNormal(RT_OBJECT_TO_WORLD, n) is the identity: no transforms are instantiated.
    switch (m.kind) {
    case OPR_MAT_DIFFUSE:
    case OPR_MAT_GLOSSY: {   // material/Diffuse.cu:71-87, material/Glossy.cu:74-90
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        float3 hitPoint = ray.origin + tHit * ray.direction;
        prd.flags |= PRD_HIT_NON_SPECULAR;
        prd.attenuation *= f3(m.Kd);
        prd.normal = worldShadingNormal;
        prd.position = hitPoint;
        prd.lastTHit = tHit;
        prd.depth++;
        if (prd.flags & PRD_PATH_TRACING) {   // Diffuse.cu:82-86, Glossy.cu:85-89
            float2 sample = getRandomUniformFloat2(&prd.randomState);
            prd.randomNewDirection = sampleUnitHemisphereCos(worldShadingNormal, sample, nullptr, nullptr, false);
        }
        break;
    }
    case OPR_MAT_TEXTURE: {  // material/Texture.cu:83-110
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        float3 hitPoint = ray.origin + tHit * ray.direction;
        float3 normal = worldShadingNormal;
        // without tangents the reference reads unwritten attributes (undefined): the map is ignored here, as in the product
        if (m.normalTexture >= 0 && hit.hasTangents) {
            float3 worldTangent = normalize(hit.tangent);
            float3 worldBitangent = normalize(hit.bitangent);
            normal = getNormalMappedNormal(worldShadingNormal, worldTangent, worldBitangent,
                                           tex2D(r.scene.textures[m.normalTexture], hit.textureCoordinate.x, hit.textureCoordinate.y));
        }
        prd.flags |= PRD_HIT_NON_SPECULAR;
        prd.normal = normal;
        prd.position = hitPoint;
        prd.lastTHit = tHit;
        if (prd.flags & PRD_PATH_TRACING)     // Texture.cu:102-105: the un-mapped shading normal
            prd.randomNewDirection = sampleUnitHemisphereCos(worldShadingNormal, getRandomUniformFloat2(&prd.randomState), nullptr, nullptr, false);
        float4 value = tex2D(r.scene.textures[m.diffuseTexture], hit.textureCoordinate.x, hit.textureCoordinate.y);
        prd.attenuation *= make_float3(value.x, value.y, value.z);
        break;
    }
    case OPR_MAT_DIFFUSE_EMITTER: {   // material/DiffuseEmitter.cu:40-51
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        prd.flags |= PRD_HIT_EMITTER;
        if (dot(worldShadingNormal, -ray.direction) < 0.f) return;
        float3 powerPerArea = f3(m.power); powerPerArea.x *= m.inverseArea; powerPerArea.y *= m.inverseArea; powerPerArea.z *= m.inverseArea;
        float3 Le = powerPerArea / M_PIf_;
        prd.radiance += prd.attenuation * Le;
        prd.lastTHit = tHit;
        break;
    }
    case OPR_MAT_MIRROR: {   // material/Mirror.cu:50-63
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        float3 hitPoint = ray.origin + tHit * ray.direction;
        prd.depth++;
        if (prd.depth <= MAX_RADIANCE_TRACE_DEPTH_) {
            prd.attenuation *= f3(m.Kr);
            float3 newRayDirection = reflect(ray.direction, worldShadingNormal);
            Ray newRay{hitPoint, newRayDirection, 0, 0.0001f, RT_DEFAULT_MAX_};
            traceRadiance(r, newRay, prd);
        }
        prd.lastTHit = tHit;
        break;
    }
    case OPR_MAT_GLASS: {    // material/Glass.cu:90-143
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        bool isHitFromOutside = hitFromOutside(ray.direction, worldShadingNormal);
        float3 N = isHitFromOutside ? worldShadingNormal : -worldShadingNormal;
        float3 hitPoint = ray.origin + tHit * ray.direction;
        float n1 = isHitFromOutside ? 1.f : m.indexOfRefraction;
        float n2 = isHitFromOutside ? m.indexOfRefraction : 1.f;
        float3 refractionDirection = make_float3(0.f);
        bool validRefraction = refract(refractionDirection, ray.direction, N, n2 / n1);
        float cosThetaI = -dot(ray.direction, N);
        float cosThetaT = -dot(refractionDirection, N);
        float reflFactor = validRefraction ? reflectionFactor(cosThetaI, cosThetaT, n1, n2) : 1.f;
        float sample = getRandomUniformFloat(&prd.randomState);
        bool isReflected = (sample <= reflFactor);
        float3 newRayDirection;
        if (isReflected) {
            newRayDirection = reflect(ray.direction, N);
        } else {
            newRayDirection = refractionDirection;
            prd.attenuation *= (n2 * n2) / (n1 * n1);
        }
        prd.flags |= PRD_HIT_SPECULAR;
        prd.flags &= ~PRD_HIT_NON_SPECULAR;
        prd.depth++;
        Ray newRay{hitPoint, newRayDirection, 0, 0.0001f, RT_DEFAULT_MAX_};
        if (prd.depth <= MAX_RADIANCE_TRACE_DEPTH_) traceRadiance(r, newRay, prd);
        else prd.attenuation *= 0;
        prd.lastTHit = tHit;
        break;
    }
    }
}

static void traceRadiance(Renderer& r, const Ray& ray, RadiancePRD& prd)
{
    Hit hit;
    r.stats.closest.fetch_add(1, std::memory_order_relaxed);
    if (r.scene.closestHit(ray, hit)) closestHitRadiance(r, ray, hit, prd);
    else {   // miss, ppm/RayGeneratorPPM.cu:72-77
        prd.flags = PRD_MISS;
        prd.attenuation = make_float3(0.f);
        prd.radiance = 0 * make_float3(5.f);
    }
}

// ppm/RayGeneratorPPM.cu:31-66 with helpers/camera.h:11-27
void ppm_generateRay(Renderer& r, uint x, uint y)
{
    RadiancePRD prd;
    prd.attenuation = make_float3(1.0f);
    prd.radiance = make_float3(0.f);
    prd.depth = 0;
    prd.flags = 0;
    prd.position = prd.normal = make_float3(0.f);   // uninitialised in the reference; 0 here for determinism
    prd.lastTHit = 0.f;
    prd.randomState = r.rs(x, y);

    const Camera& camera = r.camera;
    float2 screen = make_float2((float)r.width, (float)r.height);
    float2 sample = getRandomUniformFloat2(&prd.randomState);
    float2 d = (make_float2((float)x, (float)y) + sample) / screen * 2.0f - 1.0f;
    float3 rayOrigin = f3(camera.eye);
    float3 rayDirection = normalize(d.x * f3(camera.camera_u) + d.y * f3(camera.camera_v) + f3(camera.lookdir));

    if (camera.aperture > 0) {
        float3 focalPlaneCenterPoint = f3(camera.eye) + f3(camera.lookdir);
        float3 camLookDir = normalize(f3(camera.lookdir));
        float focalPlaneT = (dot(camLookDir, focalPlaneCenterPoint) - dot(camLookDir, f3(camera.eye))) / dot(camLookDir, rayDirection);
        float3 lookAt = rayOrigin + focalPlaneT * rayDirection;
        float2 dofSample = getRandomUniformFloat2(&prd.randomState);
        float2 disc = sampleUnitDisc(dofSample);
        rayOrigin += disc.x * f3(camera.camera_u) * camera.aperture + disc.y * f3(camera.camera_v) * camera.aperture;
        rayDirection = normalize(lookAt - rayOrigin);
    }

    Ray ray{rayOrigin, rayDirection, 0, 0.001f, RT_DEFAULT_MAX_};
    traceRadiance(r, ray, prd);

    Hitpoint& rec = r.hitpoints[(size_t)y * r.width + x];
    rec.position = prd.position;
    rec.normal = prd.normal;
    rec.attenuation = prd.attenuation;
    rec.radiance = prd.radiance;
    rec.flags = prd.flags;
    rec.volumetricRadiance = make_float3(0.f);
    r.rs(x, y) = prd.randomState;
}

// ---- PHOTON closest-hit programs ------------------------------------------------------------------------
static inline void storePhoton(Renderer& r, PhotonPRD& prd, const float3& power, const float3& pos, const float3& dir)
{   // helpers/store_photon.h:12-15
    Photon& p = r.photons[prd.pm_index + prd.numStoredPhotons];
    p.power = power; p.position = pos; p.rayDirection = dir;
    prd.numStoredPhotons++;
}

static void closestHitPhoton(Renderer& r, const Ray& ray, const Hit& hit, PhotonPRD& prd)
{
    const opr_material& m = r.scene.materials[hit.material];
    const float tHit = hit.t;
    switch (m.kind) {
    case OPR_MAT_DIFFUSE:
    case OPR_MAT_GLOSSY:
    case OPR_MAT_TEXTURE: {   // material/Diffuse.cu:92-135, Glossy.cu:94-137, Texture.cu:116-175
        const bool tex = m.kind == OPR_MAT_TEXTURE;
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        float3 hitPoint = ray.origin + tHit * ray.direction;
        // Texture.cu stores without the `< maxPhotonDepositsPerEmitted` test (:131); the early return below
        // guarantees numStoredPhotons < 4 whenever a hit program runs, so both forms store the same photons.
        if (prd.depth >= 1 && (tex || prd.numStoredPhotons < MAX_PHOTON_DEPOSITS))
            storePhoton(r, prd, prd.power, hitPoint, ray.direction);
        float3 albedo;
        if (tex) {
            float4 value = tex2D(r.scene.textures[m.diffuseTexture], hit.textureCoordinate.x, hit.textureCoordinate.y);
            albedo = make_float3(value.x, value.y, value.z);
        } else albedo = f3(m.Kd);
        prd.power *= albedo;
        prd.weight *= fmaxf3(albedo);
        if (prd.depth >= PHOTON_TRACING_RR_START_DEPTH_) {
            float probContinue = favgf(albedo);
            float probSample = getRandomUniformFloat(&prd.randomState);
            if (probSample >= probContinue) return;
            prd.power /= probContinue;
        }
        prd.depth++;
        if (prd.depth >= MAX_PHOTON_TRACE_DEPTH_ || (double)prd.weight < (tex ? 0.01 : 0.001)) return;
        if (prd.numStoredPhotons >= MAX_PHOTON_DEPOSITS) return;
        float3 newPhotonDirection = sampleUnitHemisphereCos(worldShadingNormal, getRandomUniformFloat2(&prd.randomState), nullptr, nullptr, false);
        Ray newRay{hitPoint, newPhotonDirection, 1, tex ? 0.01f : 0.0001f, RT_DEFAULT_MAX_};
        tracePhoton(r, newRay, prd);
        break;
    }
    case OPR_MAT_DIFFUSE_EMITTER:   // material/DiffuseEmitter.cu:56-59
        prd.depth++;
        break;
    case OPR_MAT_MIRROR: {   // material/Mirror.cu:65-77
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        float3 hitPoint = ray.origin + tHit * ray.direction;
        prd.depth++;
        if (prd.depth <= MAX_PHOTON_TRACE_DEPTH_) {
            prd.power *= f3(m.Kr);
            float3 newPhotonDirection = reflect(ray.direction, worldShadingNormal);
            Ray newPhoton{hitPoint, newPhotonDirection, 1, 0.0001f, RT_DEFAULT_MAX_};
            tracePhoton(r, newPhoton, prd);
        }
        break;
    }
    case OPR_MAT_GLASS: {    // material/Glass.cu:164-205
        float3 worldShadingNormal = normalize(hit.shadingNormal);
        bool isHitFromOutside = hitFromOutside(ray.direction, worldShadingNormal);
        float3 N = isHitFromOutside ? worldShadingNormal : -worldShadingNormal;
        float3 hitPoint = ray.origin + tHit * ray.direction;
        float n1 = isHitFromOutside ? 1.f : m.indexOfRefraction;
        float n2 = isHitFromOutside ? m.indexOfRefraction : 1.f;
        float3 refractionDirection = make_float3(0.f);
        bool validRefraction = refract(refractionDirection, ray.direction, N, n2 / n1);
        float cosThetaI = -dot(ray.direction, N);
        float cosThetaT = -dot(refractionDirection, N);
        float reflFactor = validRefraction ? reflectionFactor(cosThetaI, cosThetaT, n1, n2) : 1.f;
        float sample = getRandomUniformFloat(&prd.randomState);
        bool isReflected = (sample <= reflFactor);
        float3 newRayDirection = isReflected ? reflect(ray.direction, N) : refractionDirection;
        prd.depth++;
        if (prd.depth <= MAX_PHOTON_TRACE_DEPTH_) {
            Ray newRay{hitPoint, newRayDirection, 1, 0.0001f, RT_DEFAULT_MAX_};
            tracePhoton(r, newRay, prd);
        }
        break;
    }
    }
}

static void tracePhoton(Renderer& r, const Ray& ray, PhotonPRD& prd)
{
    Hit hit;
    r.stats.closest.fetch_add(1, std::memory_order_relaxed);
    if (r.scene.closestHit(ray, hit)) closestHitPhoton(r, ray, hit, prd);
    // miss: no-op (ppm/PhotonGenerator.cu:132-135)
}

// ppm/PhotonGenerator.cu:40-128
void ppm_photonGenerator(Renderer& r, uint x, uint y)
{
    PhotonPRD prd;
    prd.pm_index = (y * r.photonLaunchWidth + x) * MAX_PHOTON_DEPOSITS;
    prd.numStoredPhotons = 0;
    prd.depth = 0;
    prd.weight = 1.0f;
    prd.randomState = r.rs(x, y);

    const std::vector<Light>& lights = r.scene.lights;
    int lightIndex = 0;
    if (lights.size() > 1) {
        float sample = getRandomUniformFloat(&prd.randomState);
        lightIndex = intmin((int)(sample * lights.size()), int(lights.size() - 1));
    }
    const Light& light = lights[lightIndex];
    float powerScale = (float)lights.size();
    prd.power = f3(light.power) * powerScale;

    float3 rayOrigin, rayDirection = make_float3(0.f);
    float photonPowerFactor = 1.f;
    {   // generatePhotonOriginAndDirection :40-78
        const Sphere& boundingSphere = r.scene.boundingSphere;
        rayOrigin = f3(light.position);
        float2 sample1 = getRandomUniformFloat2(&prd.randomState);
        if (light.lightType == OPR_LIGHT_AREA) {
            float2 sample2 = getRandomUniformFloat2(&prd.randomState);
            rayOrigin += sample1.x * f3(light.v1) + sample1.y * f3(light.v2);
            rayDirection = sampleUnitHemisphere(f3(light.normal), sample2);
        } else if (light.lightType == OPR_LIGHT_POINT) {
            float3 sceneCenterToLight = f3(light.position) - boundingSphere.center;
            float lightDistance = length(sceneCenterToLight);
            sceneCenterToLight /= lightDistance;
            bool lightWellOutsideSphere = ((double)lightDistance > 1.5 * (double)boundingSphere.radius);
            if (lightWellOutsideSphere) {
                float3 pointOnDisc = sampleDisc(sample1, boundingSphere.center, boundingSphere.radius, sceneCenterToLight);
                rayDirection = normalize(pointOnDisc - rayOrigin);
                photonPowerFactor = (1 - lightDistance * (1.0f / sqrtf(boundingSphere.radius * boundingSphere.radius + lightDistance * lightDistance))) / 2.f;
            } else {
                rayDirection = sampleUnitSphere(sample1, nullptr);
            }
        } else if (light.lightType == OPR_LIGHT_SPOT) {
            float3 direction = f3(light.normal);
            float3 pointOnDisc = sampleDisc(sample1, rayOrigin + direction, sinf(light.area_or_angle / 2), direction);
            rayDirection = normalize(pointOnDisc - rayOrigin);
        }
    }
    prd.power *= photonPowerFactor;

    Ray photon{rayOrigin, rayDirection, 1, 0.0001f, RT_DEFAULT_MAX_};
    for (uint i = 0; i < MAX_PHOTON_DEPOSITS; ++i) {
        r.photons[prd.pm_index + i].position = make_float3(0.0f);
        r.photons[prd.pm_index + i].power = make_float3(0.0f);
    }
    tracePhoton(r, photon, prd);
    r.rs(x, y) = prd.randomState;
    r.photonPathLength[(size_t)y * r.photonLaunchWidth + x] = prd.depth;
}

// ---- ppm/IndirectRadianceEstimation.cu:54-227 (uniform-grid branch) -------------------------------------
void ppm_indirect(Renderer& r, uint x, uint y)
{
    const Hitpoint rec = r.hitpoints[(size_t)y * r.width + x];
    float3 indirectAccumulatedPower = make_float3(0.0f, 0.0f, 0.0f);
    if (rec.flags & PRD_HIT_NON_SPECULAR) {
        float radius2 = r.ppmRadiusSquared;
        float radius = r.ppmRadius;
        const uint3 gs = r.grid.gridSize;
        float invCellSize = 1.f / r.grid.cellSize;
        float3 normalizedPosition = rec.position - r.grid.origo;
        uint x_lo = (uint)std::max(0, f2i_sat((normalizedPosition.x - radius) * invCellSize));
        uint y_lo = (uint)std::max(0, f2i_sat((normalizedPosition.y - radius) * invCellSize));
        uint z_lo = (uint)std::max(0, f2i_sat((normalizedPosition.z - radius) * invCellSize));
        uint x_hi = std::min(gs.x - 1, f2u_sat((normalizedPosition.x + radius) * invCellSize));
        uint y_hi = std::min(gs.y - 1, f2u_sat((normalizedPosition.y + radius) * invCellSize));
        uint z_hi = std::min(gs.z - 1, f2u_sat((normalizedPosition.z + radius) * invCellSize));
        if (x_lo <= x_hi) {
            for (uint z = z_lo; z <= z_hi; z++) {
                for (uint yy = y_lo; yy <= y_hi; yy++) {
                    uint from = x_lo + yy * gs.x + z * gs.x * gs.y;
                    uint to = from + (x_hi - x_lo);
                    uint offset = r.offsets[from];
                    uint offsetTo = r.offsets[to + 1];
                    uint numPhotons = offsetTo - offset;
                    for (uint i = offset; i < offset + numPhotons; i++) {
                        const Photon& photon = r.photons[i];
                        float3 diff = rec.position - photon.position;
                        float distance2 = dot(diff, diff);
                        if (distance2 <= radius2 && dot(-photon.rayDirection, rec.normal) >= 0) {
                            // photonPower :59-67 — Jensen's gaussian filter
                            const float alpha = 1.818f, beta = 1.953f, expNegativeBeta = 0.141847f;
                            float weight = alpha * (1 - (1 - expf(-beta * distance2 / (2 * radius2))) / (1 - expNegativeBeta));
                            indirectAccumulatedPower += photon.power * weight;
                        }
                    }
                }
            }
        }
    }
    float3 indirectRadiance = indirectAccumulatedPower * rec.attenuation * (1.0f / (M_PIf_ * r.ppmRadiusSquared)) * (1.0f / r.emittedPhotonsPerIterationFloat);
    r.indirect[(size_t)y * r.width + x] = indirectRadiance;
}

// ---- helpers/light.h:29-87 -----------------------------------------------------------------------------
static float3 getLightContribution(Renderer& r, const Light& light, const float3& rec_position, const float3& rec_normal, RandomState& randomState)
{
    float lightFactor = 1;
    float lightDistance = 0;
    float3 pointOnLight = make_float3(0.f);
    if (light.lightType == OPR_LIGHT_AREA) {
        float2 sample = getRandomUniformFloat2(&randomState);
        pointOnLight = f3(light.position) + sample.x * f3(light.v1) + sample.y * f3(light.v2);
    } else if (light.lightType == OPR_LIGHT_POINT) {
        pointOnLight = f3(light.position);
        lightFactor *= 1.f / 4.f;
    } else if (light.lightType == OPR_LIGHT_SPOT) {
        lightFactor = 0;
    }
    float3 towardsLight = pointOnLight - rec_position;
    lightDistance = length(towardsLight);
    towardsLight = towardsLight / lightDistance;
    float n_dot_l = maxf(0, dot(rec_normal, towardsLight));
    lightFactor *= n_dot_l / (M_PIf_ * lightDistance * lightDistance);
    if (light.lightType == OPR_LIGHT_AREA)
        lightFactor *= maxf(0, dot(-towardsLight, f3(light.normal)));
    if (lightFactor > 0.0f) {
        float attenuation = 1.0f;
        Ray shadow_ray{rec_position, towardsLight, 7, 0.0001f, (float)((double)lightDistance - 0.0001)};
        r.stats.any.fetch_add(1, std::memory_order_relaxed);
        // gatherAnyHitOnNonEmitter (ppm/DirectRadianceEstimation.cu:79-83) is bound to every material
        // (material/Material.cpp:18-26 overrides the emitter's own any-hit, DiffuseEmitter.cpp:35,38)
        if (r.scene.anyHit(shadow_ray)) attenuation = 0.0f;
        lightFactor *= attenuation;
        return f3(light.power) * lightFactor;
    }
    return make_float3(0.f);
}

// ppm/DirectRadianceEstimation.cu:29-77
void ppm_direct(Renderer& r, uint x, uint y)
{
    const size_t idx = (size_t)y * r.width + x;
    const Hitpoint rec = r.hitpoints[idx];
    if (!(rec.flags & PRD_HIT_NON_SPECULAR)) {
        if ((rec.flags & PRD_HIT_EMITTER) && !(rec.flags & PRD_HIT_SPECULAR))
            r.direct[idx] = fminf3(rec.radiance, make_float3(1.f));
        else
            r.direct[idx] = rec.radiance;
        return;
    }
    const std::vector<Light>& lights = r.scene.lights;
    int numLights = (int)lights.size();
    const int numShadowSamples = 4;
    float3 avgLightRadiance = make_float3(0.f);
    unsigned sampleMask = 0;
    RandomState& st = r.rs(x, y);
    for (int shadowSample = 0; shadowSample < numShadowSamples; shadowSample++) {
        float sample = getRandomUniformFloat(&st);
        int randomLightIndex = intmin(int(sample * numLights), int(lights.size() - 1));
        const Light& light = lights[randomLightIndex];
        float scale = (float)numLights;
        float3 lightContrib = getLightContribution(r, light, rec.position, rec.normal, st);
        avgLightRadiance += scale * lightContrib;
        if (fmaxf(fmaxf(lightContrib.x, lightContrib.y), lightContrib.z) > 0.f) sampleMask |= 1u << shadowSample;
    }
    if (idx < r.directSampleMask.size()) r.directSampleMask[idx] = (unsigned char)sampleMask;
    r.direct[idx] = rec.attenuation * avgLightRadiance / (float)numShadowSamples;
}

// pt/RayGeneratorPT.cu:46-131 (RenderMethod::PATH_TRACING, OptixRenderer.cpp:559-567 with ptDirectLightSampling = 1).
// Two RNG streams, as in the reference: the payload copy (jitter, material samples, point on light) and the buffer state
// randomStates[launchIndex] itself (light pick, Russian roulette), which the final write-back of the payload copy discards.
void pt_generateRay(Renderer& r, uint x, uint y)
{
    RadiancePRD prd;
    prd.attenuation = make_float3(1.0f);
    prd.radiance = make_float3(0.f);
    prd.depth = 0u;
    prd.flags = 0;
    prd.position = prd.normal = prd.randomNewDirection = make_float3(0.f);
    prd.lastTHit = 0.f;
    prd.randomState = r.rs(x, y);
    RandomState& bufferState = r.rs(x, y);

    const Camera& camera = r.camera;
    float2 screen = make_float2((float)r.width, (float)r.height);
    float2 sample = getRandomUniformFloat2(&prd.randomState);
    float2 d = (make_float2((float)x, (float)y) + sample) / screen * 2.0f - 1.0f;
    float3 rayOrigin = f3(camera.eye);
    float3 rayDirection = normalize(d.x * f3(camera.camera_u) + d.y * f3(camera.camera_v) + f3(camera.lookdir));
    if (camera.aperture > 0) {   // modifyRayForDepthOfField, helpers/camera.h:11-27
        float3 focalPlaneCenterPoint = f3(camera.eye) + f3(camera.lookdir);
        float3 camLookDir = normalize(f3(camera.lookdir));
        float focalPlaneT = (dot(camLookDir, focalPlaneCenterPoint) - dot(camLookDir, f3(camera.eye))) / dot(camLookDir, rayDirection);
        float3 lookAt = rayOrigin + focalPlaneT * rayDirection;
        float2 dofSample = getRandomUniformFloat2(&prd.randomState);
        float2 disc = sampleUnitDisc(dofSample);
        rayOrigin += disc.x * f3(camera.camera_u) * camera.aperture + disc.y * f3(camera.camera_v) * camera.aperture;
        rayDirection = normalize(lookAt - rayOrigin);
    }
    Ray ray{rayOrigin, rayDirection, 0, 0.001f, RT_DEFAULT_MAX_};
    float3 finalRadiance = make_float3(0.f);
    const int ptDirectLightSampling = 1;
    const int numShadowSamples = ptDirectLightSampling ? 1 : 0;
    const int numPaths = ptDirectLightSampling ? 5 : 10;
    const std::vector<Light>& lights = r.scene.lights;
    // parity diagnostics, not part of the reference: per segment (hit flags, light sample contributed, ended by Russian roulette)
    uint sig = 0;
    auto sign = [&](uint code) { sig = sig * 131u + code + 1u; };
    for (int i = 0; i < numPaths; i++) {
        prd.flags = PRD_PATH_TRACING;
        traceRadiance(r, ray, prd);
        uint code = (prd.flags >> 27) & 0x1fu;
        if (prd.flags & PRD_HIT_EMITTER) {
            if ((prd.flags & PRD_HIT_SPECULAR) || i == 0) finalRadiance = prd.radiance;
            sign(code);
            break;
        } else if (prd.flags & PRD_HIT_NON_SPECULAR) {
            int numLights = (int)lights.size();
            float3 accumLightRadiancePreBrdf = make_float3(0.f);
            for (int shadowSample = 0; shadowSample < numShadowSamples; shadowSample++) {
                int randomLightIndex = int(getRandomUniformFloat(&bufferState) * numLights);
                const Light& light = lights[randomLightIndex];
                float scale = (float)numLights;
                float3 lightContrib = scale * getLightContribution(r, light, prd.position, prd.normal, prd.randomState);
                accumLightRadiancePreBrdf += lightContrib;
                if (lightContrib.x != 0.f || lightContrib.y != 0.f || lightContrib.z != 0.f) code |= 32u;
            }
            float3 directRadiance = prd.attenuation * accumLightRadiancePreBrdf / (float)numShadowSamples;
            finalRadiance += directRadiance;
            ray.origin = prd.position;
            ray.direction = prd.randomNewDirection;
        } else {
            sign(code);
            break;
        }
        if (i >= 3) {   // PATH_TRACING_RR_START_DEPTH, config.h:37
            float s = getRandomUniformFloat(&bufferState);
            float probabilityContinue = fmaxf(fmaxf(prd.attenuation.x, prd.attenuation.y), prd.attenuation.z);
            if (s > probabilityContinue) { sign(code | 64u); break; }
            prd.attenuation /= probabilityContinue;
        }
        sign(code);
    }
    const size_t idx = (size_t)y * r.width + x;
    if (idx < r.ptSignature.size()) r.ptSignature[idx] = sig;
    if (idx < r.hitpoints.size()) {     // diagnostics only: the last vertex of the path, where the product's hit-point buffer holds it too
        Hitpoint& h = r.hitpoints[idx];
        h.position = prd.position; h.normal = prd.normal; h.attenuation = prd.attenuation; h.radiance = prd.radiance; h.flags = prd.flags;
    }
    if (!(finalRadiance.x != finalRadiance.x || finalRadiance.y != finalRadiance.y || finalRadiance.z != finalRadiance.z))
        r.output[idx] = r.localIterationNumber == 0 ? finalRadiance : r.output[idx] + finalRadiance;
    r.rs(x, y) = prd.randomState;
}

// ppm/Output.cu:32-37
void ppm_output(Renderer& r, uint x, uint y)
{
    const size_t idx = (size_t)y * r.width + x;
    float3 finalRadiance = r.direct[idx] + r.indirect[idx];
    r.output[idx] = r.localIterationNumber == 0 ? finalRadiance : r.output[idx] + finalRadiance;
}

} // namespace ora
