// ppm.cu — Progressive Photon Mapping passes as wavefront kernels: camera-ray generation and the
// RADIANCE material programs, photon emission and the PHOTON material programs, and the fused
// direct-lighting + output pass.  Each kernel cites the reference device program it replaces.
//
// The reference's recursion (rtTrace inside closest-hit programs) becomes an iterative state machine:
// per-path state lives in HBM (hitpoint SoA for camera paths, phPower/phState for photons) and a shading
// kernel either terminates the path or pushes ONE continuation ray into the next queue (warp-aggregated).
// Shading runs bucketed by material kind: blockIdx.y selects the material queue filled by the trace kernel.
#include "renderer.h"
#include <cstdlib>

namespace opr {

constexpr int SHADE_THREADS = 128;
#define SHADE_BLOCKS_X (numSMs() * 8)    // per material kind; blocks of empty kinds exit at once (2 per SM left the dominant kind at 12 % occupancy: profiles/r1f)

// ---- helpers -------------------------------------------------------------------------------------------
OPR_DEV unsigned rng_index(unsigned path, unsigned launchW, unsigned rsW)
{   // randomStates is a 2-D buffer indexed by launchIndex (x,y): OptixRenderer.cpp:368-373
    return (path / launchW) * rsW + (path % launchW);
}

// material/Glass.cu:69-88
OPR_DEV float reflectionFactor(float cosI, float cosT, float n1, float n2)
{
    float rp = (n2 * cosI - n1 * cosT) / (n2 * cosI + n1 * cosT);
    float rs = (n1 * cosI - n2 * cosT) / (n1 * cosI + n2 * cosT);
    return (rp * rp + rs * rs) / 2.f;
}

// Fresnel reflect/refract choice shared by Glass closestHitRadiance (:90-143) and closestHitPhoton (:164-205)
OPR_DEV float3 glass_scatter(const float3& dir, const float3& worldShadingNormal, float ior, Rng& rng, bool& isReflected, float& n1, float& n2)
{
    bool isHitFromOutside = dot(worldShadingNormal, dir) < 0;      // helpers.h:146-149
    float3 N = isHitFromOutside ? worldShadingNormal : -worldShadingNormal;
    n1 = isHitFromOutside ? 1.f : ior;
    n2 = isHitFromOutside ? ior : 1.f;
    float3 refractionDirection = f3(0.f);
    bool validRefraction = refract(refractionDirection, dir, N, n2 / n1);
    float cosThetaI = -dot(dir, N);
    float cosThetaT = -dot(refractionDirection, N);
    float reflFactor = validRefraction ? reflectionFactor(cosThetaI, cosThetaT, n1, n2) : 1.f;
    float sample = rng_uniform(rng);
    isReflected = (sample <= reflFactor);
    return isReflected ? reflect(dir, N) : refractionDirection;
}

// material/Texture.cu:72-80
OPR_DEV float3 getNormalMappedNormal(const float3& normal, const float3& tangent, const float3& bitangent, const float4& normalMap)
{
    float4 nMap = make_float4(2 * normalMap.x - 1, 2 * normalMap.y - 1, 2 * normalMap.z - 1, 2 * normalMap.w - 1);
    float3 N;
    N.x = nMap.x * tangent.x + nMap.y * bitangent.x + nMap.z * normal.x;
    N.y = nMap.x * tangent.y + nMap.y * bitangent.y + nMap.z * normal.y;
    N.z = nMap.x * tangent.z + nMap.y * bitangent.z + nMap.z * normal.z;
    return normalize(N);
}

// Per-bounce queue counters (renderer.h CTR_*): the generation kernel of a pass zeroes them and seeds bounce 0 with the
// launch size, so the bounce loop needs no memset.  The trace kernel of bounce 0 is stream-ordered after this kernel.
OPR_DEV void reset_wave_counters(unsigned* __restrict__ counters, unsigned n)
{
    if (blockIdx.x == 0)
        for (unsigned i = CTR_RAYCOUNT + threadIdx.x; i < CTR_TOTAL; i += blockDim.x) counters[i] = (i == CTR_RAYCOUNT) ? n : 0u;
}

// tail-kernel shape (measurement knobs): rays per warp = 32 / spread; freeRun = lanes traverse without votes
static unsigned tail_blocks() { static int v = -1; if (v < 0) { const char* e = getenv("OPR_TAIL_BLOCKS_PER_SM"); v = e ? atoi(e) : 4; if (v < 1) v = 1; if (v > 16) v = 16; } return (unsigned)v; }
static unsigned tail_refill_idle() { static int v = -1; if (v < 0) { const char* e = getenv("OPR_TAIL_REFILL"); v = e ? atoi(e) : 8; if (v < 1) v = 1; if (v > 32) v = 32; } return (unsigned)v; }
static unsigned tail_free() { static int v = -1; if (v < 0) { const char* e = getenv("OPR_TAIL_FREE"); v = e ? atoi(e) : 1; } return (unsigned)v; }

// 32-bit per-iteration ray counters -> 64-bit totals (a long render would wrap 32 bits after a few hundred iterations)
__global__ void stats_fold_kernel(unsigned* __restrict__ c32, unsigned long long* __restrict__ t64)
{
    const unsigned i = threadIdx.x;
    if (i < STAT_NUM) { t64[i] += c32[i]; c32[i] = 0u; }
}
void launch_stats_fold(unsigned* counters32, unsigned long long* totals64, LaunchCtx lc)
{
    stats_fold_kernel<<<1, 32, 0, lc.st>>>(counters32, totals64);
}

// ---- RNG state init: initRandomStateBuffer (OptixRenderer_SpatialHash.cu:310-317) ---------------------
__global__ void rng_init_kernel(RngBuf rng, unsigned seed, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        Rng s; rng_seed(s, seed, (unsigned)i);
        rng_store(rng, i, s);
    }
}
void launch_rng_init(const RngBuf& rng, unsigned seed, size_t n, LaunchCtx lc)
{
    rng_init_kernel<<<numSMs() * 8, 256, 0, lc.st>>>(rng, seed, n);
    ++*lc.launches;
}

// ---- camera pass: RayGeneratorPPM::generateRay (ppm/RayGeneratorPPM.cu:31-66) + DOF (helpers/camera.h:11-27)
__global__ void __launch_bounds__(SHADE_THREADS) ppm_camera_gen_kernel(opr_camera cam, unsigned w, unsigned h, unsigned rsW, RngBuf rngBuf,
                                                                       HitpointBuf hp, float4* __restrict__ rayO, float4* __restrict__ rayD,
                                                                       unsigned* __restrict__ rayPath, unsigned* __restrict__ counters,
                                                                       unsigned char* __restrict__ camPending)
{
    const unsigned n = w * h;
    reset_wave_counters(counters, n);
    if (blockIdx.x == 0 && threadIdx.x == 0) { counters[CTR_TAIL_COUNT] = 0u; counters[CTR_DEFERRED_COUNT] = 0u; counters[CTR_TAIL_FETCH] = 0u; counters[CTR_LATE_FETCH] = 0u; }
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const unsigned p = tiled_pixel(s, w, h);
        camPending[p] = 0;
        const unsigned x = p % w, y = p / w;
        const size_t ri = (size_t)y * rsW + x;
        Rng rng = rng_load(rngBuf, ri);
        const float3 eye = f3(cam.eye[0], cam.eye[1], cam.eye[2]);
        const float3 cu = f3(cam.camera_u[0], cam.camera_u[1], cam.camera_u[2]);
        const float3 cv = f3(cam.camera_v[0], cam.camera_v[1], cam.camera_v[2]);
        const float3 lookdir = f3(cam.lookdir[0], cam.lookdir[1], cam.lookdir[2]);
        float2 sample = rng_uniform2(rng);
        float dx = ((float)x + sample.x) / (float)w * 2.0f - 1.0f;
        float dy = ((float)y + sample.y) / (float)h * 2.0f - 1.0f;
        float3 rayOrigin = eye;
        float3 rayDirection = normalize(dx * cu + dy * cv + lookdir);
        if (cam.aperture > 0) {
            float3 focalPlaneCenterPoint = eye + lookdir;
            float3 camLookDir = normalize(lookdir);
            float focalPlaneT = (dot(camLookDir, focalPlaneCenterPoint) - dot(camLookDir, eye)) / dot(camLookDir, rayDirection);
            float3 lookAt = rayOrigin + focalPlaneT * rayDirection;
            float2 dofSample = rng_uniform2(rng);
            float2 disc = sampleUnitDisc(dofSample);
            rayOrigin += disc.x * cu * cam.aperture + disc.y * cv * cam.aperture;
            rayDirection = normalize(lookAt - rayOrigin);
        }
        rng_store(rngBuf, ri, rng);
        // RadiancePRD initial state (:33-37); position/normal are uninitialised in the reference, 0 here
        hp.pos[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
        hp.nrm[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
        hp.att[p] = make_float4(1.f, 1.f, 1.f, 0.f);
        hp.rad[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        rayO[s] = make_float4(rayOrigin.x, rayOrigin.y, rayOrigin.z, 0.001f);
        rayD[s] = make_float4(rayDirection.x, rayDirection.y, rayDirection.z, RT_DEFAULT_MAX);
        rayPath[s] = p;
    }
}

void launch_ppm_camera_gen(const opr_camera& cam, unsigned w, unsigned h, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp,
                           const RayQueues& q, LaunchCtx lc)
{
    unsigned n = w * h;
    int blocks = (int)std::min<unsigned>((n + SHADE_THREADS - 1) / SHADE_THREADS, numSMs() * 16);
    ppm_camera_gen_kernel<<<blocks, SHADE_THREADS, 0, lc.st>>>(cam, w, h, rsW, rng, hp, q.rayO[0], q.rayD[0], q.rayPath[0], q.counters, q.camPending);
    ++*lc.launches;
}

// RADIANCE closest-hit programs (Diffuse.cu:71-87, Glossy.cu:74-90, Texture.cu:83-110, DiffuseEmitter.cu:40-51,
// Mirror.cu:50-63, Glass.cu:90-143, miss RayGeneratorPPM.cu:72-77) for ONE hit of camera path p.  Path state (the
// RadiancePRD fields) lives in the hitpoint SoA.  Returns true when the program recursed (rtTrace): the continuation
// ray is (newO, newD) with tmin 1e-4.
// ptNormal (PATH_TRACING only): receives the world shading normal of a Diffuse / Glossy / Texture hit, from which the
// programs draw randomNewDirection (Diffuse.cu:82-86, Glossy.cu:85-89, Texture.cu:102-105 — the un-mapped normal).
OPR_DEV bool shade_camera_hit(const SceneView& sc, int kind, unsigned p, const float3& o, const float3& d, const HitRec& hit, unsigned w, unsigned rsW,
                              const RngBuf& rngBuf, const HitpointBuf& hp, float3& newO, float3& newD, float3* ptNormal = nullptr)
{
    if (kind == QUEUE_MISS) {
        hp.pos[p].w = __uint_as_float(PRD_MISS);
        hp.att[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        hp.rad[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        return false;
    }
    bool push = false;
    const Surface sf = eval_surface(sc, hit, o, d);
    const DMaterial& m = sc.materials[sf.material];
    const float tHit = hit.t;
    unsigned flags = __float_as_uint(hp.pos[p].w);
    unsigned depth = __float_as_uint(hp.nrm[p].w);
    float3 att = f3(hp.att[p]);
    const float3 worldShadingNormal = normalize(sf.shadingNormal);
    const float3 hitPoint = o + tHit * d;
    if (ptNormal && (kind == OPR_MAT_DIFFUSE || kind == OPR_MAT_GLOSSY || kind == OPR_MAT_TEXTURE)) *ptNormal = worldShadingNormal;
    if (kind == OPR_MAT_DIFFUSE || kind == OPR_MAT_GLOSSY) {
        flags |= PRD_HIT_NON_SPECULAR;
        att *= m.Kd;
        depth++;
        hp.pos[p] = make_float4(hitPoint.x, hitPoint.y, hitPoint.z, __uint_as_float(flags));
        hp.nrm[p] = make_float4(worldShadingNormal.x, worldShadingNormal.y, worldShadingNormal.z, __uint_as_float(depth));
        hp.att[p] = make_float4(att.x, att.y, att.z, 0.f);
    } else if (kind == OPR_MAT_TEXTURE) {
        float3 normal = worldShadingNormal;
        // A normal map on a mesh without tangents / bitangents: the reference reads OptiX attributes its intersection program
        // never wrote (TriangleMesh.cu:61-72 vs Texture.cu:89-95, undefined); here the map is ignored and the interpolated
        // shading normal is used.  assimp's CalcTangentSpace (scene/Scene.cpp:96) gives every imported mesh
        // with normals and uvs a tangent frame, so reference scenes never get here.
        if (m.hasNormalTex && sf.hasTangents) {
            float3 worldTangent = normalize(sf.tangent), worldBitangent = normalize(sf.bitangent);
            normal = getNormalMappedNormal(worldShadingNormal, worldTangent, worldBitangent, tex2D<float4>(m.normalTex, sf.uv.x, sf.uv.y));
        }
        flags |= PRD_HIT_NON_SPECULAR;
        float4 value = tex2D<float4>(m.diffuseTex, sf.uv.x, sf.uv.y);
        att *= f3(value.x, value.y, value.z);
        hp.pos[p] = make_float4(hitPoint.x, hitPoint.y, hitPoint.z, __uint_as_float(flags));
        hp.nrm[p] = make_float4(normal.x, normal.y, normal.z, __uint_as_float(depth));   // no depth++ (Texture.cu:83-110)
        hp.att[p] = make_float4(att.x, att.y, att.z, 0.f);
    } else if (kind == OPR_MAT_DIFFUSE_EMITTER) {
        flags |= PRD_HIT_EMITTER;
        hp.pos[p].w = __uint_as_float(flags);
        if (!(dot(worldShadingNormal, -d) < 0.f)) {
            float3 Le = m.powerPerArea / PI_F;
            float3 rad = f3(hp.rad[p]);
            rad += att * Le;
            hp.rad[p] = make_float4(rad.x, rad.y, rad.z, 0.f);
        }
    } else if (kind == OPR_MAT_MIRROR) {
        depth++;
        hp.nrm[p].w = __uint_as_float(depth);
        if (depth <= MAX_RADIANCE_TRACE_DEPTH) {
            att *= m.Kr;
            hp.att[p] = make_float4(att.x, att.y, att.z, 0.f);
            newO = hitPoint; newD = reflect(d, worldShadingNormal);
            push = true;
        }
    } else {   // OPR_MAT_GLASS
        const size_t ri = rng_index(p, w, rsW);
        Rng rng = rng_load(rngBuf, ri);
        bool isReflected; float n1, n2;
        newD = glass_scatter(d, worldShadingNormal, m.ior, rng, isReflected, n1, n2);
        rng_store(rngBuf, ri, rng);
        if (!isReflected) att *= (n2 * n2) / (n1 * n1);
        flags |= PRD_HIT_SPECULAR;
        flags &= ~PRD_HIT_NON_SPECULAR;
        depth++;
        newO = hitPoint;
        if (depth <= MAX_RADIANCE_TRACE_DEPTH) push = true;
        else att *= 0.f;
        hp.pos[p].w = __uint_as_float(flags);
        hp.nrm[p].w = __uint_as_float(depth);
        hp.att[p] = make_float4(att.x, att.y, att.z, 0.f);
    }
    return push;
}

// Wavefront shading of one bounce, one material kind per blockIdx.y (the queues the trace kernel filled).
__global__ void __launch_bounds__(SHADE_THREADS) ppm_camera_shade_kernel(SceneView sc, unsigned w, unsigned rsW, RngBuf rngBuf, HitpointBuf hp,
                                                                         const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                         const unsigned* __restrict__ rayPath, const float4* __restrict__ hits,
                                                                         const unsigned* __restrict__ matQueue, const unsigned* __restrict__ matCount,
                                                                         unsigned capacity, float4* __restrict__ nextO, float4* __restrict__ nextD,
                                                                         unsigned* __restrict__ nextPath, unsigned* __restrict__ nextCount,
                                                                         unsigned char* __restrict__ pending)
{
    const int kind = blockIdx.y;
    const unsigned n = matCount[kind];
    const unsigned* queue = matQueue + (size_t)kind * capacity;
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned qi = base + threadIdx.x;
        const bool active = qi < n;
        bool push = false;
        float3 newO = f3(0.f), newD = f3(0.f);
        unsigned p = 0;
        if (active) {
            const unsigned slot = queue[qi];
            p = rayPath[slot];
            HitRec hit; hit.t = 0.f; hit.prim = -1; hit.u = hit.v = 0.f;
            float3 o = f3(0.f), d = f3(0.f);
            if (kind != QUEUE_MISS) {
                const float4 o4 = rayO[slot], d4 = rayD[slot], h4 = hits[slot];
                o = f3(o4); d = f3(d4);
                hit.t = h4.x; hit.prim = __float_as_int(h4.y); hit.u = h4.z; hit.v = h4.w;
            }
            push = shade_camera_hit(sc, kind, p, o, d, hit, w, rsW, rngBuf, hp, newO, newD);
        }
        // ray regeneration: one atomicAdd per warp
        const unsigned dst = warp_append(nextCount, push);
        if (push) {
            nextO[dst] = make_float4(newO.x, newO.y, newO.z, 0.0001f);
            nextD[dst] = make_float4(newD.x, newD.y, newD.z, RT_DEFAULT_MAX);
            nextPath[dst] = p;
            if (pending) pending[p] = 1;     // this pixel's path (and its RNG state) is still in flight: see ppm_photon_gen_kernel
        }
    }
}

// Tail of the camera pass: the few specular continuation rays left after the wavefront bounces are finished here, one
// lane per path looping trace -> material program until the path ends (the reference's recursion, unrolled in place).
// A near-empty wavefront bounce costs a full trace + shade launch pair whose duration is one ray's dependent-load
// latency; this kernel pays that latency once per path instead of once per bounce.
//
// Path regeneration: warps are persistent and pull paths from the queue through a shared cursor; a lane whose path ended
// takes the next path as soon as `refillIdle` lanes of its warp are idle, instead of idling until the warp's longest path is done
// (chains of 1 to 9 specular bounces here, 1 to 4 more Russian-roulette bounces in the photon tail: the fixed assignment left warps
// running with a quarter of their lanes).
template <bool QN>
__global__ void __launch_bounds__(SHADE_THREADS) ppm_camera_tail_kernel(SceneView sc, unsigned w, unsigned rsW, RngBuf rngBuf, HitpointBuf hp,
                                                                        const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                        const unsigned* __restrict__ rayPath, const unsigned* __restrict__ count,
                                                                        unsigned* __restrict__ stat, unsigned* __restrict__ cursor, unsigned refillIdle, unsigned freeRun)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n = *count;
    unsigned traced = 0, p = 0;
    bool alive = false, exhausted = false;
    float3 o = f3(0.f), d = f3(0.f, 0.f, 1.f);
    float tmin = 0.0001f;
    for (;;) {
        unsigned i = 0;
        if (tail_refill(n, cursor, alive, exhausted, refillIdle, i)) {
            const float4 o4 = rayO[i], d4 = rayD[i];
            p = rayPath[i]; o = f3(o4); d = f3(d4); tmin = o4.w; alive = true;
        }
        if (!__any_sync(FULL, alive)) break;
        HitRec hit;
        if (freeRun) { if (alive) traverse_free<false, QN>(sc, o, d, tmin, RT_DEFAULT_MAX, hit); }
        else traverse_warp<false, QN>(sc, FULL, alive, o, d, tmin, RT_DEFAULT_MAX, hit);
        if (alive) {
            traced++;
            const int kind = hit.prim < 0 ? QUEUE_MISS : sc.materials[prim_material(sc, hit.prim)].kind;
            float3 newO = f3(0.f), newD = f3(0.f);
            alive = shade_camera_hit(sc, kind, p, o, d, hit, w, rsW, rngBuf, hp, newO, newD);
            o = newO; d = newD; tmin = 0.0001f;
        }
    }
    for (int s = 16; s > 0; s >>= 1) traced += __shfl_xor_sync(FULL, traced, s);
    if ((threadIdx.x & 31) == 0 && traced) atomicAdd(stat, traced);
}

// lastWave: the continuation rays go to the dedicated tail queue (and their pixels are flagged), so that the tail kernel can run
// on its own stream while the photon pass re-uses the ping-pong queues.
void launch_ppm_camera_shade(const SceneView& sc, unsigned w, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp,
                             const RayQueues& q, int bounce, bool lastWave, LaunchCtx lc)
{
    dim3 grid(SHADE_BLOCKS_X, NUM_QUEUES);
    const int cur = bounce & 1, nxt = cur ^ 1;
    if (lastWave)
        ppm_camera_shade_kernel<<<grid, SHADE_THREADS, 0, lc.st>>>(sc, w, rsW, rng, hp, q.rayO[cur], q.rayD[cur], q.rayPath[cur], q.hits, q.matQueue,
                                                                   q.counters + CTR_MATCOUNT + 8 * bounce, q.capacity, q.tailO, q.tailD, q.tailPath,
                                                                   q.counters + CTR_TAIL_COUNT, q.camPending);
    else
        ppm_camera_shade_kernel<<<grid, SHADE_THREADS, 0, lc.st>>>(sc, w, rsW, rng, hp, q.rayO[cur], q.rayD[cur], q.rayPath[cur], q.hits, q.matQueue,
                                                                   q.counters + CTR_MATCOUNT + 8 * bounce, q.capacity, q.rayO[nxt], q.rayD[nxt], q.rayPath[nxt],
                                                                   q.counters + CTR_RAYCOUNT + bounce + 1, nullptr);
    ++*lc.launches;
}

void launch_ppm_camera_tail(const SceneView& sc, unsigned w, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp, const RayQueues& q, LaunchCtx lc, bool quantized)
{
    auto k = quantized ? ppm_camera_tail_kernel<true> : ppm_camera_tail_kernel<false>;
    k<<<numSMs() * tail_blocks(), SHADE_THREADS, 0, lc.st>>>(sc, w, rsW, rng, hp, q.tailO, q.tailD, q.tailPath,
                                                                 q.counters + CTR_TAIL_COUNT, q.counters + 16 + STAT_CLOSEST_RAYS, q.counters + CTR_TAIL_FETCH,
                                                                 tail_refill_idle(), tail_free());
    ++*lc.launches;
}

// ---- photon pass: PhotonGenerator::generator (ppm/PhotonGenerator.cu:40-128) ----------------------------
// Emission of photon thread t: draws from its RNG state (1 light pick if several lights, 2 + 2 for an area light), resets
// the thread's PhotonPRD (phPower / phState) and returns the first ray.
OPR_DEV void photon_emit(const SceneView& sc, unsigned t, unsigned pw, unsigned rsW, const RngBuf& rngBuf, float4* __restrict__ phPower,
                         unsigned* __restrict__ phState, float3& origin, float3& direction)
{
    const unsigned x = t % pw, y = t / pw;
    const size_t ri = (size_t)y * rsW + x;
    Rng rng = rng_load(rngBuf, ri);
    int lightIndex = 0;
    if (sc.numLights > 1) {
        float sample = rng_uniform(rng);
        lightIndex = min((int)(sample * sc.numLights), sc.numLights - 1);
    }
    const opr_light light = sc.lights[lightIndex];
    float powerScale = (float)sc.numLights;
    float3 power = f3(light.power[0], light.power[1], light.power[2]) * powerScale;
    origin = f3(light.position[0], light.position[1], light.position[2]);
    direction = f3(0.f);
    float photonPowerFactor = 1.f;
    float2 sample1 = rng_uniform2(rng);
    if (light.lightType == OPR_LIGHT_AREA) {
        float2 sample2 = rng_uniform2(rng);
        origin += sample1.x * f3(light.v1[0], light.v1[1], light.v1[2]) + sample1.y * f3(light.v2[0], light.v2[1], light.v2[2]);
        direction = sampleUnitHemisphere(f3(light.normal[0], light.normal[1], light.normal[2]), sample2);
    } else if (light.lightType == OPR_LIGHT_POINT) {
        float3 sceneCenterToLight = origin - sc.boundingSphere.center;
        float lightDistance = length(sceneCenterToLight);
        sceneCenterToLight /= lightDistance;
        bool lightWellOutsideSphere = ((double)lightDistance > 1.5 * (double)sc.boundingSphere.radius);
        if (lightWellOutsideSphere) {
            float3 pointOnDisc = sampleDisc(sample1, sc.boundingSphere.center, sc.boundingSphere.radius, sceneCenterToLight);
            direction = normalize(pointOnDisc - origin);
            float r = sc.boundingSphere.radius;
            photonPowerFactor = (1 - lightDistance * (1.0f / sqrtf(r * r + lightDistance * lightDistance))) / 2.f;
        } else {
            direction = sampleUnitSphere(sample1);
        }
    } else {
        float3 ldir = f3(light.normal[0], light.normal[1], light.normal[2]);
        float3 pointOnDisc = sampleDisc(sample1, origin + ldir, sinf(light.area_or_angle / 2), ldir);
        direction = normalize(pointOnDisc - origin);
    }
    power *= photonPowerFactor;
    rng_store(rngBuf, ri, rng);
    phPower[t] = make_float4(power.x, power.y, power.z, 1.0f);
    phState[t] = 0u;    // depth 0, numStoredPhotons 0 (also "clears" the 4 slots this thread owns)
}

// Photon thread (x, y) shares RNG state (x, y) with pixel (x, y) (OptixRenderer.cpp:368-373) and must start from the state the
// camera pass leaves behind.  The few pixels whose camera path is still bouncing between specular surfaces (camPending, set by
// the last wavefront shade of the camera pass) are finished by the camera tail kernel on another stream WHILE this pass runs;
// their photon threads are deferred to ppm_photon_late_kernel, which follows the tail kernel on that stream.  Every other
// thread emits here; the rays are appended compactly (one atomic per warp).
__global__ void __launch_bounds__(SHADE_THREADS) ppm_photon_gen_kernel(SceneView sc, unsigned pw, unsigned ph, unsigned rsW, RngBuf rngBuf,
                                                                       float4* __restrict__ phPower, unsigned* __restrict__ phState,
                                                                       float4* __restrict__ rayO, float4* __restrict__ rayD, unsigned* __restrict__ rayPath,
                                                                       unsigned* __restrict__ counters, const unsigned char* __restrict__ camPending,
                                                                       unsigned w, unsigned h, unsigned* __restrict__ deferred,
                                                                       unsigned char* __restrict__ photonPending, unsigned char pendingValue)
{
    const unsigned n = pw * ph;     // the per-bounce counters were zeroed by the launcher: this kernel appends to bounce 0
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned t = base + threadIdx.x;
        const bool valid = t < n;
        const unsigned x = valid ? t % pw : 0u, y = valid ? t / pw : 0u;
        const bool deferredThread = valid && x < w && y < h && camPending[(size_t)y * w + x];
        const bool emit = valid && !deferredThread;
        float3 origin = f3(0.f), direction = f3(0.f);
        if (emit) photon_emit(sc, t, pw, rsW, rngBuf, phPower, phState, origin, direction);
        if (valid) photonPending[t] = emit ? pendingValue : (unsigned char)0;   // (a deferred thread is covered by its pixel's camPending)
        const unsigned dst = warp_append(counters + CTR_RAYCOUNT, emit);
        if (emit) {
            rayO[dst] = make_float4(origin.x, origin.y, origin.z, 0.0001f);
            rayD[dst] = make_float4(direction.x, direction.y, direction.z, RT_DEFAULT_MAX);
            rayPath[dst] = t;
        }
        const unsigned dd = warp_append(counters + CTR_DEFERRED_COUNT, deferredThread);
        if (deferredThread) deferred[dd] = t;
    }
}

void launch_ppm_photon_gen(const SceneView& sc, unsigned pw, unsigned ph, unsigned rsW, const RngBuf& rng, float4* phPower,
                           unsigned* phState, const RayQueues& q, unsigned w, unsigned h, LaunchCtx lc, bool feedsTail)
{
    unsigned n = pw * ph;
    // keep the camera pass's queue lengths for the occupancy report, then zero the per-bounce counters of this pass
    OPR_CUDA_CHECK(cudaMemcpyAsync(q.counters + CTR_SNAPSHOT, q.counters + CTR_RAYCOUNT, 16 * sizeof(unsigned), cudaMemcpyDeviceToDevice, lc.st));
    OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + CTR_RAYCOUNT, 0, (CTR_TOTAL - CTR_RAYCOUNT) * sizeof(unsigned), lc.st));
    int blocks = (int)std::min<unsigned>((n + SHADE_THREADS - 1) / SHADE_THREADS, numSMs() * 16);
    ppm_photon_gen_kernel<<<blocks, SHADE_THREADS, 0, lc.st>>>(sc, pw, ph, rsW, rng, phPower, phState, q.rayO[0], q.rayD[0], q.rayPath[0], q.counters,
                                                               q.camPending, w, h, q.deferred, q.photonPending, feedsTail ? 1 : 0);
    ++*lc.launches;
}

// PHOTON closest-hit programs (Diffuse.cu:92-135, Glossy.cu:94-137, Texture.cu:116-175, Mirror.cu:65-77,
// Glass.cu:164-205, DiffuseEmitter.cu:56-59; the miss program is a no-op, PhotonGenerator.cu:132-135) for ONE hit of
// photon thread t.  PhotonPRD lives in phPower / phState.  Returns true when the program recursed.
OPR_DEV bool shade_photon_hit(const SceneView& sc, int kind, unsigned t, const float3& o, const float3& d, const HitRec& hit, unsigned pw, unsigned rsW,
                              const RngBuf& rngBuf, float4* __restrict__ phPower, unsigned* __restrict__ phState, const PhotonBuf& photons,
                              float3& newO, float3& newD, float& newTmin)
{
    bool push = false;
    newTmin = 0.0001f;
    unsigned state = phState[t];
    unsigned depth = state & 0xffu, numStored = (state >> 8) & 0xffu;
    if (kind == OPR_MAT_DIFFUSE_EMITTER) {
        depth++;
    } else {
        const Surface sf = eval_surface(sc, hit, o, d);
        const DMaterial& m = sc.materials[sf.material];
        const float3 worldShadingNormal = normalize(sf.shadingNormal);
        const float3 hitPoint = o + hit.t * d;
        float4 pw4 = phPower[t];
        float3 power = f3(pw4);
        float weight = pw4.w;
        if (kind == OPR_MAT_MIRROR) {
            depth++;
            if (depth <= MAX_PHOTON_TRACE_DEPTH) {
                power *= m.Kr;
                newO = hitPoint; newD = reflect(d, worldShadingNormal); push = true;
            }
        } else if (kind == OPR_MAT_GLASS) {
            const size_t ri = rng_index(t, pw, rsW);
            Rng rng = rng_load(rngBuf, ri);
            bool isReflected; float n1, n2;
            newD = glass_scatter(d, worldShadingNormal, m.ior, rng, isReflected, n1, n2);
            rng_store(rngBuf, ri, rng);
            depth++;
            newO = hitPoint;
            if (depth <= MAX_PHOTON_TRACE_DEPTH) push = true;
        } else {   // Diffuse / Glossy / Texture
            const bool tex = (kind == OPR_MAT_TEXTURE);
            if (depth >= 1 && numStored < MAX_PHOTON_DEPOSITS) {
                // STORE_PHOTON (helpers/store_photon.h:12-15): slot pm_index + numStoredPhotons
                const size_t ps = (size_t)t * MAX_PHOTON_DEPOSITS + numStored;
                photons.a[ps] = make_float4(hitPoint.x, hitPoint.y, hitPoint.z, d.x);
                photons.b[ps] = make_float4(power.x, power.y, power.z, d.y);
                photons.c[ps] = d.z;
                numStored++;
            }
            float3 albedo;
            if (tex) { float4 v = tex2D<float4>(m.diffuseTex, sf.uv.x, sf.uv.y); albedo = f3(v.x, v.y, v.z); }
            else albedo = m.Kd;
            power *= albedo;
            weight *= fmax3(albedo);
            bool alive = true;
            Rng rng; bool rngLoaded = false;
            const size_t ri = rng_index(t, pw, rsW);
            if (depth >= PHOTON_TRACING_RR_START_DEPTH) {
                rng = rng_load(rngBuf, ri); rngLoaded = true;
                float probContinue = favgf(albedo);
                float probSample = rng_uniform(rng);
                if (probSample >= probContinue) alive = false;
                else power /= probContinue;
            }
            if (alive) {
                depth++;
                if (depth >= MAX_PHOTON_TRACE_DEPTH || (double)weight < (tex ? 0.01 : 0.001)) alive = false;
            }
            if (alive && numStored >= MAX_PHOTON_DEPOSITS) alive = false;
            if (alive) {
                if (!rngLoaded) { rng = rng_load(rngBuf, ri); rngLoaded = true; }
                float2 s = rng_uniform2(rng);
                newD = sampleUnitHemisphereCos(worldShadingNormal, s);
                newO = hitPoint;
                newTmin = tex ? 0.01f : 0.0001f;
                push = true;
            }
            if (rngLoaded) rng_store(rngBuf, ri, rng);
        }
        phPower[t] = make_float4(power.x, power.y, power.z, weight);
    }
    phState[t] = depth | (numStored << 8);
    return push;
}

__global__ void __launch_bounds__(SHADE_THREADS) ppm_photon_shade_kernel(SceneView sc, unsigned pw, unsigned rsW, RngBuf rngBuf,
                                                                         float4* __restrict__ phPower, unsigned* __restrict__ phState, PhotonBuf photons,
                                                                         const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                         const unsigned* __restrict__ rayPath, const float4* __restrict__ hits,
                                                                         const unsigned* __restrict__ matQueue, const unsigned* __restrict__ matCount,
                                                                         unsigned capacity, float4* __restrict__ nextO, float4* __restrict__ nextD,
                                                                         unsigned* __restrict__ nextPath, unsigned* __restrict__ nextCount,
                                                                         unsigned char* __restrict__ pendingOut /* last wavefront bounce only */)
{
    const int kind = blockIdx.y;
    if (kind == QUEUE_MISS) return;
    const unsigned n = matCount[kind];
    const unsigned* queue = matQueue + (size_t)kind * capacity;
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned qi = base + threadIdx.x;
        const bool active = qi < n;
        bool push = false;
        float3 newO = f3(0.f), newD = f3(0.f);
        float newTmin = 0.0001f;
        unsigned t = 0;
        if (active) {
            const unsigned slot = queue[qi];
            t = rayPath[slot];
            HitRec hit; hit.t = 0.f; hit.prim = -1; hit.u = hit.v = 0.f;
            float3 o = f3(0.f), d = f3(0.f);
            if (kind != OPR_MAT_DIFFUSE_EMITTER) {
                const float4 o4 = rayO[slot], d4 = rayD[slot], h4 = hits[slot];
                o = f3(o4); d = f3(d4);
                hit.t = h4.x; hit.prim = __float_as_int(h4.y); hit.u = h4.z; hit.v = h4.w;
            }
            push = shade_photon_hit(sc, kind, t, o, d, hit, pw, rsW, rngBuf, phPower, phState, photons, newO, newD, newTmin);
        }
        const unsigned dst = warp_append(nextCount, push);
        if (push) {
            nextO[dst] = make_float4(newO.x, newO.y, newO.z, newTmin);
            nextD[dst] = make_float4(newD.x, newD.y, newD.z, RT_DEFAULT_MAX);
            nextPath[dst] = t;
            if (pendingOut) pendingOut[t] = 1;      // this photon goes on in the tail kernel: its RNG state is not final yet
        }
    }
}

// Tail of the photon pass (see ppm_camera_tail_kernel): photons still alive after the wavefront bounces finish their
// random walk here, one thread per photon.
template <bool QN>
__global__ void __launch_bounds__(SHADE_THREADS) ppm_photon_tail_kernel(SceneView sc, unsigned pw, unsigned rsW, RngBuf rngBuf, float4* __restrict__ phPower,
                                                                        unsigned* __restrict__ phState, PhotonBuf photons, const float4* __restrict__ rayO,
                                                                        const float4* __restrict__ rayD, const unsigned* __restrict__ rayPath,
                                                                        const unsigned* __restrict__ count, unsigned* __restrict__ stat, unsigned* __restrict__ cursor,
                                                                        unsigned refillIdle, unsigned freeRun)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n = *count;
    unsigned traced = 0, t = 0;
    bool alive = false, exhausted = false;
    float3 o = f3(0.f), d = f3(0.f, 0.f, 1.f);
    float tmin = 0.0001f;
    for (;;) {
        unsigned i = 0;
        if (tail_refill(n, cursor, alive, exhausted, refillIdle, i)) {     // path regeneration: see ppm_camera_tail_kernel
            const float4 o4 = rayO[i], d4 = rayD[i];
            t = rayPath[i]; o = f3(o4); d = f3(d4); tmin = o4.w; alive = true;
        }
        if (!__any_sync(FULL, alive)) break;
        HitRec hit;
        if (freeRun) { if (alive) traverse_free<false, QN>(sc, o, d, tmin, RT_DEFAULT_MAX, hit); }
        else traverse_warp<false, QN>(sc, FULL, alive, o, d, tmin, RT_DEFAULT_MAX, hit);
        if (alive) {
            traced++;
            if (hit.prim < 0) alive = false;   // miss: no-op
            else {
                const int kind = sc.materials[prim_material(sc, hit.prim)].kind;
                float3 newO = f3(0.f), newD = f3(0.f);
                float newTmin;
                alive = shade_photon_hit(sc, kind, t, o, d, hit, pw, rsW, rngBuf, phPower, phState, photons, newO, newD, newTmin);
                o = newO; d = newD; tmin = newTmin;
            }
        }
    }
    for (int s = 16; s > 0; s >>= 1) traced += __shfl_xor_sync(FULL, traced, s);
    if ((threadIdx.x & 31) == 0 && traced) atomicAdd(stat, traced);
}

void launch_ppm_photon_shade(const SceneView& sc, unsigned pw, unsigned rsW, const RngBuf& rng, float4* phPower, unsigned* phState,
                             const PhotonBuf& photons, const RayQueues& q, int bounce, LaunchCtx lc, bool lastWave)
{
    dim3 grid(SHADE_BLOCKS_X, NUM_QUEUES);
    const int cur = bounce & 1, nxt = cur ^ 1;
    ppm_photon_shade_kernel<<<grid, SHADE_THREADS, 0, lc.st>>>(sc, pw, rsW, rng, phPower, phState, photons, q.rayO[cur], q.rayD[cur], q.rayPath[cur],
                                                               q.hits, q.matQueue, q.counters + CTR_MATCOUNT + 8 * bounce, q.capacity, q.rayO[nxt], q.rayD[nxt],
                                                               q.rayPath[nxt], q.counters + CTR_RAYCOUNT + bounce + 1, lastWave ? q.photonPending : nullptr);
    ++*lc.launches;
}

void launch_ppm_photon_tail(const SceneView& sc, unsigned pw, unsigned rsW, const RngBuf& rng, float4* phPower, unsigned* phState,
                            const PhotonBuf& photons, const RayQueues& q, int bounce, LaunchCtx lc, bool quantized)
{
    const int cur = bounce & 1;
    auto k = quantized ? ppm_photon_tail_kernel<true> : ppm_photon_tail_kernel<false>;
    k<<<numSMs() * tail_blocks(), SHADE_THREADS, 0, lc.st>>>(sc, pw, rsW, rng, phPower, phState, photons, q.rayO[cur], q.rayD[cur], q.rayPath[cur],
                                                                 q.counters + CTR_RAYCOUNT + bounce, q.counters + 16 + STAT_CLOSEST_RAYS, q.counters + CTR_FETCH + bounce,
                                                                 tail_refill_idle(), tail_free());
    ++*lc.launches;
}

// The deferred photon threads (see ppm_photon_gen_kernel): emission and the whole random walk, one thread per photon.
template <bool QN>
__global__ void __launch_bounds__(SHADE_THREADS) ppm_photon_late_kernel(SceneView sc, unsigned pw, unsigned rsW, RngBuf rngBuf, float4* __restrict__ phPower,
                                                                        unsigned* __restrict__ phState, PhotonBuf photons, const unsigned* __restrict__ deferred,
                                                                        const unsigned* __restrict__ count, unsigned* __restrict__ stat, unsigned* __restrict__ cursor,
                                                                        unsigned refillIdle, unsigned freeRun)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n = *count;
    unsigned traced = 0, t = 0;
    bool alive = false, exhausted = false;
    float3 o = f3(0.f), d = f3(0.f, 0.f, 1.f);
    float tmin = 0.0001f;
    for (;;) {
        unsigned i = 0;
        if (tail_refill(n, cursor, alive, exhausted, refillIdle, i)) {     // path regeneration: see ppm_camera_tail_kernel
            t = deferred[i];
            photon_emit(sc, t, pw, rsW, rngBuf, phPower, phState, o, d);
            tmin = 0.0001f; alive = true;
        }
        if (!__any_sync(FULL, alive)) break;
        HitRec hit;
        if (freeRun) { if (alive) traverse_free<false, QN>(sc, o, d, tmin, RT_DEFAULT_MAX, hit); }
        else traverse_warp<false, QN>(sc, FULL, alive, o, d, tmin, RT_DEFAULT_MAX, hit);
        if (alive) {
            traced++;
            if (hit.prim < 0) alive = false;   // miss: no-op
            else {
                const int kind = sc.materials[prim_material(sc, hit.prim)].kind;
                float3 newO = f3(0.f), newD = f3(0.f);
                float newTmin;
                alive = shade_photon_hit(sc, kind, t, o, d, hit, pw, rsW, rngBuf, phPower, phState, photons, newO, newD, newTmin);
                o = newO; d = newD; tmin = newTmin;
            }
        }
    }
    for (int s = 16; s > 0; s >>= 1) traced += __shfl_xor_sync(FULL, traced, s);
    if ((threadIdx.x & 31) == 0 && traced) atomicAdd(stat, traced);
}

void launch_ppm_photon_late(const SceneView& sc, unsigned pw, unsigned rsW, const RngBuf& rng, float4* phPower, unsigned* phState, const PhotonBuf& photons,
                            const RayQueues& q, LaunchCtx lc, bool quantized)
{
    auto k = quantized ? ppm_photon_late_kernel<true> : ppm_photon_late_kernel<false>;
    k<<<numSMs() * tail_blocks(), SHADE_THREADS, 0, lc.st>>>(sc, pw, rsW, rng, phPower, phState, photons, q.deferred, q.counters + CTR_DEFERRED_COUNT,
                                                                 q.counters + 16 + STAT_CLOSEST_RAYS, q.counters + CTR_LATE_FETCH, tail_refill_idle(), tail_free());
    ++*lc.launches;
}

// ---- direct lighting + output: DirectRadianceEstimation::kernel (ppm/DirectRadianceEstimation.cu:29-77),
// getLightContribution (helpers/light.h:29-87), gatherAnyHitOnNonEmitter (:79-83), Output::kernel (ppm/Output.cu:32-37).
//
// Three steps instead of one thread walking four shadow rays: (1) ppm_direct_gen_kernel draws the four light samples of
// every diffuse pixel in the reference's RNG order (DirectRadianceEstimation.cu:62,66), stores each sample's light index
// and un-occluded factor and appends the rays that must be cast (factor > 0: 39 % of the 4*W*H slots on the bench scene)
// to a compacted queue; (2) trace_shadow_queue_kernel (trace.cu) clears the factor of occluded samples; (3)
// ppm_resolve_output_kernel sums the samples exactly as the reference does and accumulates direct + indirect.
// The one-thread-per-pixel version kept 61 % of its lanes idle in lockstep and ran the four traversals back to back.
// The four light samples of pixel p (all lanes of the warp call this together; lanes with inImage = false ride along).
OPR_DEV void direct_samples(const SceneView& sc, unsigned p, bool inImage, unsigned w, unsigned rsW, const RngBuf& rngBuf, const HitpointBuf& hp,
                            float4* __restrict__ sampleFactor, uint2* __restrict__ sampleLight, float4* __restrict__ shadowD,
                            unsigned* __restrict__ shadowIdx, unsigned* __restrict__ shadowCount)
{
    const float4 pos4 = inImage ? hp.pos[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool diffuse = inImage && (__float_as_uint(pos4.w) & PRD_HIT_NON_SPECULAR);
    float factor[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned lights01 = 0u, lights23 = 0u;      // 16 bits per sample
    float3 rec_position = f3(pos4), rec_normal = f3(0.f);
    size_t ri = 0;
    Rng rng{};
    if (diffuse) {
        rec_normal = f3(hp.nrm[p]);
        ri = (size_t)(p / w) * rsW + (p % w);
        rng = rng_load(rngBuf, ri);
    }
    const int numLights = sc.numLights;
    #pragma unroll
    for (int shadowSample = 0; shadowSample < 4; shadowSample++) {
        float lightFactor = 0.f, lightDistance = 1.f;
        float3 towardsLight = f3(0.f, 0.f, 1.f);
        if (diffuse) {
            float sample = rng_uniform(rng);
            int randomLightIndex = min(int(sample * numLights), numLights - 1);
            const opr_light& light = sc.lights[randomLightIndex];
            if (shadowSample < 2) lights01 |= (unsigned)randomLightIndex << (16 * shadowSample); else lights23 |= (unsigned)randomLightIndex << (16 * (shadowSample - 2));
            // getLightContribution
            lightFactor = 1;
            float3 pointOnLight = f3(0.f);
            const int lightType = light.lightType;
            if (lightType == OPR_LIGHT_AREA) {
                float2 s = rng_uniform2(rng);
                pointOnLight = f3(light.position[0], light.position[1], light.position[2]) + s.x * f3(light.v1[0], light.v1[1], light.v1[2]) +
                               s.y * f3(light.v2[0], light.v2[1], light.v2[2]);
            } else if (lightType == OPR_LIGHT_POINT) {
                pointOnLight = f3(light.position[0], light.position[1], light.position[2]);
                lightFactor *= 1.f / 4.f;
            } else {
                lightFactor = 0;
            }
            towardsLight = pointOnLight - rec_position;
            lightDistance = length(towardsLight);
            towardsLight = towardsLight / lightDistance;
            float n_dot_l = maxf_(0, dot(rec_normal, towardsLight));
            lightFactor *= n_dot_l / (PI_F * lightDistance * lightDistance);
            if (lightType == OPR_LIGHT_AREA)
                lightFactor *= maxf_(0, dot(-towardsLight, f3(light.normal[0], light.normal[1], light.normal[2])));
        }
        const bool cast = diffuse && lightFactor > 0.0f;
        factor[shadowSample] = cast ? lightFactor : 0.f;
        const unsigned dst = warp_append(shadowCount, cast);
        if (cast) {
            shadowD[dst] = make_float4(towardsLight.x, towardsLight.y, towardsLight.z, (float)((double)lightDistance - 0.0001));
            shadowIdx[dst] = 4u * p + shadowSample;
        }
    }
    if (diffuse) rng_store(rngBuf, ri, rng);
    if (inImage) { sampleFactor[p] = make_float4(factor[0], factor[1], factor[2], factor[3]); sampleLight[p] = make_uint2(lights01, lights23); }
}

// Early part: every pixel whose hit point and RNG state are final once the photon wavefronts are done.  A pixel waits (is
// appended to `deferred`) when its camera path is still in the camera tail (camPending) or when the photon thread that shares
// its RNG state is still walking (photonPending): the tails run on another stream meanwhile.
__global__ void __launch_bounds__(SHADE_THREADS) ppm_direct_gen_kernel(SceneView sc, unsigned w, unsigned h, unsigned rsW, RngBuf rngBuf, HitpointBuf hp,
                                                                       float4* __restrict__ sampleFactor, uint2* __restrict__ sampleLight,
                                                                       float4* __restrict__ shadowD, unsigned* __restrict__ shadowIdx,
                                                                       unsigned* __restrict__ counters, const unsigned char* __restrict__ camPending,
                                                                       const unsigned char* __restrict__ photonPending, unsigned pw, unsigned ph,
                                                                       unsigned* __restrict__ deferred)
{
    const unsigned n = w * h;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CTR_FETCH_SHADOW] = 0u;     // consumed by the trace kernel that follows
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const bool inRange = base + threadIdx.x < n;
        const unsigned p = inRange ? tiled_pixel(base + threadIdx.x, w, h) : n;
        bool inImage = inRange;
        bool wait = false;
        if (inImage) {
            const unsigned x = p % w, y = p / w;
            wait = camPending && (camPending[p] || (x < pw && y < ph && photonPending[(size_t)y * pw + x]));   // null: nothing is pending
        }
        const unsigned dd = warp_append(counters + CTR_DIRECT_DEFERRED, wait);
        if (wait) { deferred[dd] = p; inImage = false; }
        direct_samples(sc, p, inImage, w, rsW, rngBuf, hp, sampleFactor, sampleLight, shadowD, shadowIdx, counters + CTR_SHADOW_COUNT);
    }
}

// Late part: the pixels that waited, after the tail kernels.
__global__ void __launch_bounds__(SHADE_THREADS) ppm_direct_late_kernel(SceneView sc, unsigned w, unsigned rsW, RngBuf rngBuf, HitpointBuf hp,
                                                                        float4* __restrict__ sampleFactor, uint2* __restrict__ sampleLight,
                                                                        float4* __restrict__ shadowD, unsigned* __restrict__ shadowIdx,
                                                                        unsigned* __restrict__ counters, const unsigned* __restrict__ deferred)
{
    const unsigned n = counters[CTR_DIRECT_DEFERRED];
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CTR_FETCH_SHADOW2] = 0u;
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned i = base + threadIdx.x;
        const bool valid = i < n;
        const unsigned p = valid ? deferred[i] : 0u;
        direct_samples(sc, p, valid, w, rsW, rngBuf, hp, sampleFactor, sampleLight, shadowD, shadowIdx, counters + CTR_SHADOW2_COUNT);
    }
}

__global__ void __launch_bounds__(256) ppm_resolve_output_kernel(SceneView sc, unsigned n, HitpointBuf hp, const float4* __restrict__ sampleFactor,
                                                                 const uint2* __restrict__ sampleLight, const float* __restrict__ indirect,
                                                                 float* __restrict__ direct, float* __restrict__ output, unsigned localIteration)
{
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const unsigned flags = __float_as_uint(hp.pos[p].w);
        float3 directRadiance;
        if (!(flags & PRD_HIT_NON_SPECULAR)) {
            float3 rad = f3(hp.rad[p]);
            if ((flags & PRD_HIT_EMITTER) && !(flags & PRD_HIT_SPECULAR)) directRadiance = fmin3(rad, f3(1.f));
            else directRadiance = rad;
        } else {
            const float3 attenuation = f3(hp.att[p]);
            const float4 f4 = sampleFactor[p];
            const float factor[4] = {f4.x, f4.y, f4.z, f4.w};
            const uint2 lights = sampleLight[p];
            const float scale = (float)sc.numLights;
            float3 avgLightRadiance = f3(0.f);
            #pragma unroll
            for (int s = 0; s < 4; s++) {
                const opr_light& light = sc.lights[((s < 2 ? lights.x : lights.y) >> (16 * (s & 1))) & 0xffffu];
                const float3 lightContrib = f3(light.power[0], light.power[1], light.power[2]) * factor[s];
                avgLightRadiance += scale * lightContrib;
            }
            directRadiance = attenuation * avgLightRadiance / 4.0f;
        }
        const float3 ind = f3(indirect[3 * (size_t)p], indirect[3 * (size_t)p + 1], indirect[3 * (size_t)p + 2]);
        if (direct) { direct[3 * (size_t)p] = directRadiance.x; direct[3 * (size_t)p + 1] = directRadiance.y; direct[3 * (size_t)p + 2] = directRadiance.z; }
        const float3 finalRadiance = directRadiance + ind;
        float3 out = finalRadiance;
        if (localIteration != 0) out = f3(output[3 * (size_t)p], output[3 * (size_t)p + 1], output[3 * (size_t)p + 2]) + finalRadiance;
        output[3 * (size_t)p] = out.x; output[3 * (size_t)p + 1] = out.y; output[3 * (size_t)p + 2] = out.z;
    }
}

void launch_ppm_direct_gen(const SceneView& sc, unsigned w, unsigned h, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp, float4* sampleFactor,
                           uint2* sampleLight, float4* shadowD, unsigned* shadowIdx, const RayQueues& q, unsigned pw, unsigned ph, int part, LaunchCtx lc)
{   // part 0: every pixel (nothing pending); 1: early part, pixels behind the tails are deferred; 2: the deferred pixels
    unsigned n = w * h;
    const bool late = part == 2;
    if (!late) {
        OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + CTR_SHADOW_COUNT, 0, sizeof(unsigned), lc.st));
        OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + CTR_SHADOW2_COUNT, 0, 3 * sizeof(unsigned), lc.st));   // late queue length, cursor, deferred pixels
        int blocks = (int)std::min<unsigned>((n + SHADE_THREADS - 1) / SHADE_THREADS, numSMs() * 16);
        ppm_direct_gen_kernel<<<blocks, SHADE_THREADS, 0, lc.st>>>(sc, w, h, rsW, rng, hp, sampleFactor, sampleLight, shadowD, shadowIdx, q.counters,
                                                                   part == 1 ? q.camPending : nullptr, q.photonPending, pw, ph, q.directDeferred);
    } else {
        ppm_direct_late_kernel<<<numSMs() * 4, SHADE_THREADS, 0, lc.st>>>(sc, w, rsW, rng, hp, sampleFactor, sampleLight, shadowD, shadowIdx, q.counters, q.directDeferred);
    }
    ++*lc.launches;
}

void launch_ppm_resolve_output(const SceneView& sc, unsigned n, const HitpointBuf& hp, const float4* sampleFactor, const uint2* sampleLight,
                               const float* indirect, float* direct, float* output, unsigned localIteration, LaunchCtx lc)
{
    int blocks = (int)std::min<unsigned>((n + 255) / 256, numSMs() * 16);
    ppm_resolve_output_kernel<<<blocks, 256, 0, lc.st>>>(sc, n, hp, sampleFactor, sampleLight, indirect, direct, output, localIteration);
    ++*lc.launches;
}

// ---- PATH_TRACING: pt/RayGeneratorPT.cu:46-131 (OptixRenderer.cpp:559-567, ptDirectLightSampling = 1) ----------------------
// One thread per pixel, warp-lockstep traversal.  Up to five path segments; each segment is a closest-hit with the same
// RADIANCE material programs as the PPM camera pass (specular chains unrolled in place), one light sample with a shadow ray
// at every diffuse vertex, Russian roulette from the fourth segment.  The reference draws from TWO states that start
// equal: the payload copy (pixel jitter, Fresnel choice, bounce direction, point on light; written back at the end) and the
// buffer state itself (light pick, Russian roulette; overwritten by that write-back).  Here the buffer in HBM holds the
// payload stream and `bufferState` is the register copy of the other one.
constexpr unsigned PRD_PATH_TRACING = 1u << 26;
constexpr int PATH_TRACING_RR_START_DEPTH = 3;   // config.h:37

__global__ void __launch_bounds__(SHADE_THREADS) pt_generate_ray_kernel(SceneView sc, opr_camera cam, unsigned w, unsigned h, unsigned rsW, RngBuf rngBuf,
                                                                        HitpointBuf hp, float* __restrict__ output, unsigned localIteration,
                                                                        unsigned* __restrict__ stats, unsigned* __restrict__ signature)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n = w * h;
    unsigned closestRays = 0, anyRays = 0;
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const bool inImage = base + threadIdx.x < n;
        const unsigned p = inImage ? tiled_pixel(base + threadIdx.x, w, h) : n;
        const unsigned x = inImage ? p % w : 0u, y = inImage ? p / w : 0u;
        const size_t ri = (size_t)y * rsW + x;
        float3 rayOrigin = f3(0.f), rayDirection = f3(0.f, 0.f, 1.f);
        Rng bufferState{};
        if (inImage) {
            Rng prdState = rng_load(rngBuf, ri);
            bufferState = prdState;
            const float3 eye = f3(cam.eye[0], cam.eye[1], cam.eye[2]);
            const float3 cu = f3(cam.camera_u[0], cam.camera_u[1], cam.camera_u[2]);
            const float3 cv = f3(cam.camera_v[0], cam.camera_v[1], cam.camera_v[2]);
            const float3 lookdir = f3(cam.lookdir[0], cam.lookdir[1], cam.lookdir[2]);
            float2 sample = rng_uniform2(prdState);
            float dx = ((float)x + sample.x) / (float)w * 2.0f - 1.0f;
            float dy = ((float)y + sample.y) / (float)h * 2.0f - 1.0f;
            rayOrigin = eye;
            rayDirection = normalize(dx * cu + dy * cv + lookdir);
            if (cam.aperture > 0) {
                float3 focalPlaneCenterPoint = eye + lookdir;
                float3 camLookDir = normalize(lookdir);
                float focalPlaneT = (dot(camLookDir, focalPlaneCenterPoint) - dot(camLookDir, eye)) / dot(camLookDir, rayDirection);
                float3 lookAt = rayOrigin + focalPlaneT * rayDirection;
                float2 dofSample = rng_uniform2(prdState);
                float2 disc = sampleUnitDisc(dofSample);
                rayOrigin += disc.x * cu * cam.aperture + disc.y * cv * cam.aperture;
                rayDirection = normalize(lookAt - rayOrigin);
            }
            rng_store(rngBuf, ri, prdState);
            hp.pos[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
            hp.nrm[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
            hp.att[p] = make_float4(1.f, 1.f, 1.f, 0.f);
            hp.rad[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float3 finalRadiance = f3(0.f);
        bool alive = inImage;
        const int numLights = sc.numLights;
        // parity diagnostics (null otherwise): per path segment, the hit flags, whether the light sample contributed and whether
        // Russian roulette ended the path — the discrete decisions of a path, none of which shows in the final RNG state
        unsigned sig = 0;
        for (int i = 0; i < 5 && __any_sync(FULL, alive); i++) {
            const bool aliveAtEntry = alive;
            unsigned code = 0;
            if (alive) hp.pos[p].w = __uint_as_float(PRD_PATH_TRACING);
            // ---- rtTrace(sceneRootObject, ray, radiancePrd): the chain of recursive closest-hit programs
            bool chain = alive;
            float3 o = rayOrigin, d = rayDirection, ptNormal = f3(0.f, 0.f, 1.f);
            float tmin = 0.001f;     // the Ray object keeps the tmin it was constructed with (:63)
            bool sampled = false;
            while (__any_sync(FULL, chain)) {
                HitRec hit;
                traverse_warp<false>(sc, FULL, chain, o, d, tmin, RT_DEFAULT_MAX, hit);
                if (chain) {
                    closestRays++;
                    const int kind = hit.prim < 0 ? QUEUE_MISS : sc.materials[prim_material(sc, hit.prim)].kind;
                    float3 newO = f3(0.f), newD = f3(0.f);
                    chain = shade_camera_hit(sc, kind, p, o, d, hit, w, rsW, rngBuf, hp, newO, newD, &ptNormal);
                    sampled = (kind == OPR_MAT_DIFFUSE || kind == OPR_MAT_GLOSSY || kind == OPR_MAT_TEXTURE);
                    o = newO; d = newD; tmin = 0.0001f;
                }
            }
            float lightFactor = 0.f, lightDistance = 1.f;
            float3 towardsLight = f3(0.f, 0.f, 1.f), power = f3(0.f), position = f3(0.f), attenuation = f3(0.f);
            bool diffuse = false;
            if (alive) {
                const float4 pos4 = hp.pos[p];
                const unsigned flags = __float_as_uint(pos4.w);
                code = (flags >> 27) & 0x1fu;
                if (flags & PRD_HIT_EMITTER) {
                    if ((flags & PRD_HIT_SPECULAR) || i == 0) finalRadiance = f3(hp.rad[p]);
                    alive = false;
                } else if (flags & PRD_HIT_NON_SPECULAR) {
                    diffuse = true;
                    position = f3(pos4); attenuation = f3(hp.att[p]);
                    const float3 normal = f3(hp.nrm[p]);
                    Rng prdState = rng_load(rngBuf, ri);
                    if (sampled) {     // randomNewDirection, drawn inside the closest-hit program
                        float2 s = rng_uniform2(prdState);
                        rayDirection = sampleUnitHemisphereCos(ptNormal, s);
                    }
                    rayOrigin = position;
                    const int randomLightIndex = int(rng_uniform(bufferState) * numLights);
                    const opr_light& light = sc.lights[randomLightIndex];
                    // getLightContribution, helpers/light.h:29-87
                    lightFactor = 1;
                    float3 pointOnLight = f3(0.f);
                    const int lightType = light.lightType;
                    if (lightType == OPR_LIGHT_AREA) {
                        float2 s = rng_uniform2(prdState);
                        pointOnLight = f3(light.position[0], light.position[1], light.position[2]) + s.x * f3(light.v1[0], light.v1[1], light.v1[2]) +
                                       s.y * f3(light.v2[0], light.v2[1], light.v2[2]);
                    } else if (lightType == OPR_LIGHT_POINT) {
                        pointOnLight = f3(light.position[0], light.position[1], light.position[2]);
                        lightFactor *= 1.f / 4.f;
                    } else {
                        lightFactor = 0;
                    }
                    rng_store(rngBuf, ri, prdState);
                    towardsLight = pointOnLight - position;
                    lightDistance = length(towardsLight);
                    towardsLight = towardsLight / lightDistance;
                    float n_dot_l = maxf_(0, dot(normal, towardsLight));
                    lightFactor *= n_dot_l / (PI_F * lightDistance * lightDistance);
                    if (lightType == OPR_LIGHT_AREA)
                        lightFactor *= maxf_(0, dot(-towardsLight, f3(light.normal[0], light.normal[1], light.normal[2])));
                    power = f3(light.power[0], light.power[1], light.power[2]);
                } else {
                    alive = false;
                }
            }
            const bool cast = diffuse && lightFactor > 0.0f;
            HitRec shadowHit;
            const bool occluded = traverse_warp<true>(sc, FULL, cast, position, towardsLight, 0.0001f, (float)((double)lightDistance - 0.0001), shadowHit);
            if (diffuse) {
                float3 lightContrib = f3(0.f);
                if (cast) { anyRays++; lightFactor *= occluded ? 0.0f : 1.0f; lightContrib = power * lightFactor; }
                if (lightContrib.x != 0.f || lightContrib.y != 0.f || lightContrib.z != 0.f) code |= 32u;
                const float scale = (float)numLights;
                const float3 accumLightRadiancePreBrdf = f3(0.f) + scale * lightContrib;
                const float3 directRadiance = attenuation * accumLightRadiancePreBrdf / 1.0f;
                finalRadiance += directRadiance;
            }
            if (alive && i >= PATH_TRACING_RR_START_DEPTH) {
                const float sample = rng_uniform(bufferState);
                const float probabilityContinue = fmax3(attenuation);
                if (sample > probabilityContinue) { alive = false; code |= 64u; }
                else { attenuation /= probabilityContinue; hp.att[p] = make_float4(attenuation.x, attenuation.y, attenuation.z, 0.f); }
            }
            if (aliveAtEntry) sig = sig * 131u + code + 1u;
        }
        if (signature && inImage) signature[p] = sig;
        if (inImage && !(finalRadiance.x != finalRadiance.x || finalRadiance.y != finalRadiance.y || finalRadiance.z != finalRadiance.z)) {
            float3 out = finalRadiance;
            if (localIteration != 0) out = f3(output[3 * (size_t)p], output[3 * (size_t)p + 1], output[3 * (size_t)p + 2]) + finalRadiance;
            output[3 * (size_t)p] = out.x; output[3 * (size_t)p + 1] = out.y; output[3 * (size_t)p + 2] = out.z;
        }
    }
    for (int s = 16; s > 0; s >>= 1) { closestRays += __shfl_xor_sync(FULL, closestRays, s); anyRays += __shfl_xor_sync(FULL, anyRays, s); }
    if ((threadIdx.x & 31) == 0) { if (closestRays) atomicAdd(stats + STAT_CLOSEST_RAYS, closestRays); if (anyRays) atomicAdd(stats + STAT_ANY_RAYS, anyRays); }
}

void launch_pt_generate_ray(const SceneView& sc, const opr_camera& cam, unsigned w, unsigned h, unsigned rsW, const RngBuf& rng, const HitpointBuf& hp,
                            float* output, unsigned localIteration, unsigned* stats, LaunchCtx lc, unsigned* signature)
{
    unsigned n = w * h;
    int blocks = (int)std::min<unsigned>((n + SHADE_THREADS - 1) / SHADE_THREADS, numSMs() * 16);
    pt_generate_ray_kernel<<<blocks, SHADE_THREADS, 0, lc.st>>>(sc, cam, w, h, rsW, rng, hp, output, localIteration, stats, signature);
    ++*lc.launches;
}

} // namespace opr
