// vcm.cu — vertex-connection bidirectional path tracing with recursive MIS (the reference's renderer/vcm) as
// wavefront kernels.  Replaces VCMLightPass::lightPass / VCMCameraPass::cameraPass (vcm/VCMLightPass.cu:60-162,
// vcm/VCMCameraPass.cu:52-146), the material vcmClosestHitLight / vcmClosestHitCamera programs (material/*.cu),
// lightHit / cameraHit / connectCameraT1 / connectVertices / connectLightSourceS0,S1 / sampleScattering
// (vcm/vcm.h:43-628), the MIS updates (vcm/mis.h:35-178) and lightEmit / lightIlluminate (helpers/light.h:92-213).
//
// Light vertices: the reference appends 184-byte LightVertex records through one global atomic cursor plus a
// per-path index table (vcm.h:281-287).  Here path p owns slot k*numPaths + p for its k-th stored vertex, so there
// is no cursor, no index table and no overflow; a vertex is six float4 (96 B): hit point + dVCM, throughput + dVC,
// BSDF normal + dVM, local incident direction + material kind, albedo + exponent, Ks.
#include "renderer.h"
#include "bsdf.cuh"

namespace opr {

constexpr int VCM_THREADS = 128;
#define VCM_BLOCKS_X (numSMs() * 6)

struct VcmState {
    float4* state = nullptr;   // per path, one 32-byte record: throughput.xyz, depth (uint bits) | dVCM, dVC, dVM, -
    float4* color = nullptr;   // camera-path colour accumulator
    unsigned* lvCount = nullptr;
    float4* lv = nullptr;      // stored light vertices, six float4 (96 B, contiguous) per slot
    // deferred connections of one bounce: shadow segment + the contribution it carries if unoccluded
    float4* connO = nullptr;   // origin.xyz, tmax
    float4* connD = nullptr;   // direction.xyz, target index (uint bits)
    float4* connC = nullptr;   // contribution.xyz
    float4* conn2[3] = {nullptr, nullptr, nullptr};   // second set: bounce b+1 fills it while stream B still tests the segments of bounce b
    unsigned* dbgConn = nullptr;   // parity diagnostics (opr_debug_set_connection_counts): contributions that passed isOccluded, per target pixel
    unsigned* dbgSig = nullptr;    // same switch: hash of the material kinds each subpath met (light paths [0, paths), camera paths [paths, 2 paths))
    size_t paths = 0; unsigned maxVerts = 0;
};

struct VcmParams {
    opr_camera camera;
    unsigned w, h, rsW;
    unsigned maxPathLen, estimatePass, lightSubpathCount;
    float misVcWeightFactor, misVmWeightFactor;
    float2 pixelSizeFactor;
};

struct VcmBuffers { float4* state; float4* color; unsigned* lvCount; float4* lv; unsigned paths;
                    float4* connO; float4* connD; float4* connC; unsigned* connCount; unsigned* dbgConn; unsigned* dbgSig; };
// path state record: one 256-bit load / store (the shading kernels reach paths in queue order: one sector instead of two each way)
OPR_DEV void path_load(const VcmBuffers& B, unsigned p, float4& t4, float4& m4)
{
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w), "=f"(m4.x), "=f"(m4.y), "=f"(m4.z), "=f"(m4.w) : "l"(B.state + 2 * (size_t)p) : "memory");
}
OPR_DEV void path_store(const VcmBuffers& B, unsigned p, const float4& t4, const float4& m4)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(B.state + 2 * (size_t)p), "f"(t4.x), "f"(t4.y), "f"(t4.z), "f"(t4.w), "f"(m4.x), "f"(m4.y), "f"(m4.z), "f"(m4.w) : "memory");
}
constexpr int CTR_CONN_COUNT = 9;     // RayQueues::counters word: length of the connection queue of the current bounce
constexpr int CTR_CONN_FETCH = 11;    // its fetch cursor
constexpr int VCM_MAX_CONNECTIONS = 16;   // per camera vertex: one light sample + at most maxPathLength - 1 light vertices

void vcm_free(Renderer& r)
{
    if (!r.vcm) return;
    VcmState* s = r.vcm;
    devFree(s->state); devFree(s->color); devFree(s->lvCount);
    devFree(s->connO); devFree(s->connD); devFree(s->connC);
    for (int i = 0; i < 3; i++) devFree(s->conn2[i]);
    devFree(s->lv); devFree(s->dbgConn); devFree(s->dbgSig);
    delete s; r.vcm = nullptr;
}

static void vcm_alloc(Renderer& r)
{
    const size_t paths = (size_t)r.width * r.height;
    const unsigned maxVerts = r.vcmMaxPathLength - 1;
    if (r.vcm && r.vcm->paths == paths && r.vcm->maxVerts == maxVerts) return;
    vcm_free(r);
    VcmState* s = new VcmState(); r.vcm = s;
    s->paths = paths; s->maxVerts = maxVerts;
    devAlloc((void**)&s->state, paths * 32);
    devAlloc((void**)&s->color, paths * 16); devAlloc((void**)&s->lvCount, paths * 4);
    devAlloc((void**)&s->lv, paths * maxVerts * 96);
    if (maxVerts + 1 > VCM_MAX_CONNECTIONS) throw std::runtime_error("vcmMaxPathLength too large for the connection queue");
    const size_t cap = paths * (maxVerts + 1);
    devAlloc((void**)&s->connO, cap * 16); devAlloc((void**)&s->connD, cap * 16);
    devAlloc((void**)&s->connC, cap * 16);
    if (r.overlap) for (int i = 0; i < 3; i++) devAlloc((void**)&s->conn2[i], cap * 16);
}

// ---- shared device pieces ----------------------------------------------------------------------------------------
OPR_DEV float pdfWtoA(const float aPdfW, const float aDist, const float aCosThere) { return aPdfW * fabsf(aCosThere) / sqr(aDist); }

// isOccluded (vcm.h:43-59) is deferred: a shading thread records the shadow segment and the contribution it would add, the
// records of a warp are appended to the connection queue with one atomic, and trace_connect_queue_kernel (persistent any-hit
// traversal, full warps) adds the contributions of the unoccluded segments.  Testing occlusion inside the shading threads
// ran the traversal with 5 of 32 lanes (profiles/r1i).  Segments shorter than 3 EPS_RAY are unoccluded by definition (:48).
struct Connection { float3 dir; float tmax; float3 contrib; };
// isOccluded tested in place (tail kernels: a handful of paths, nothing to compact)
OPR_DEV bool vcm_occluded_inline(const SceneView& sc, const float3& p, const float3& d, float tmax, unsigned& shadowRays)
{
    if (tmax < 3.f * EPS_RAY) return false;
    HitRec hit;
    shadowRays++;
    return traverse_free<true>(sc, p, d, EPS_RAY, tmax - 2.f * EPS_RAY, hit);
}
OPR_DEV void warp_flush_connections(const VcmBuffers& B, const float3& origin, unsigned target, const Connection* c, unsigned n)
{   // must be reached by all 32 lanes
    const unsigned lane = threadIdx.x & 31;
    unsigned incl = n;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += y; }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    unsigned base = 0;
    if (lane == 31) base = atomicAdd(B.connCount, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n;
    for (unsigned k = 0; k < n; k++) {
        B.connO[base + k] = make_float4(origin.x, origin.y, origin.z, c[k].tmax);
        B.connD[base + k] = make_float4(c[k].dir.x, c[k].dir.y, c[k].dir.z, __uint_as_float(target));
        B.connC[base + k] = make_float4(c[k].contrib.x, c[k].contrib.y, c[k].contrib.z, 0.f);
    }
}

// updateMisTermsOnHit, mis.h:109-121
OPR_DEV void mis_on_hit(float& dVCM, float& dVC, float& dVM, float cosThetaIn, float rayLen)
{
    dVCM *= sqr(rayLen);
    dVCM /= cosThetaIn; dVC /= cosThetaIn; dVM /= cosThetaIn;
}

// updateMisTermsOnScatter, mis.h:133-178
OPR_DEV void mis_on_scatter(float& dVCM, float& dVC, float& dVM, float cosThetaOut, float bsdfDirPdfW, float bsdfRevPdfW,
                            float misVcWeightFactor, float misVmWeightFactor, int sampledEvent)
{
    if (sampledEvent & B_Specular) { dVCM = 0.f; dVC *= cosThetaOut; dVM *= cosThetaOut; return; }
    const float ndVC = (cosThetaOut / bsdfDirPdfW) * (dVC * bsdfRevPdfW + dVCM + misVmWeightFactor);
    const float ndVM = (cosThetaOut / bsdfDirPdfW) * (dVM * bsdfRevPdfW + dVCM * misVcWeightFactor * 1.f + 1.f);
    dVC = ndVC; dVM = ndVM;
    dVCM = 1.f / bsdfDirPdfW;
}

// The material VCM closest-hit programs: build the VcmBSDF from the GEOMETRIC normal (Diffuse.cu:174-243,
// Glossy.cu:173-252, Texture.cu:210-286, Mirror.cu:118-189, Glass.cu:258-367).  Returns false when the program ends
// the path without a BSDF; *hitNormal is the normal handed to lightHit / cameraHit.
OPR_DEV bool vcm_build_bsdf(const DMaterial& m, int kind, const Surface& sf, const float3& rayDir, bool isLight, VcmBSDF& bsdf, float3& hitNormal, float3& albedo)
{
    const float3 worldGeometricNormal = normalize(sf.geometricNormal);
    hitNormal = worldGeometricNormal;
    albedo = m.Kd;
    if (kind == OPR_MAT_DIFFUSE) {
        bsdf.init(worldGeometricNormal, -rayDir, isLight);
        bsdf.add(bx_lambertian(m.Kd));
        return true;
    }
    if (kind == OPR_MAT_GLOSSY) {
        bsdf.init(worldGeometricNormal, -rayDir, isLight);
        bsdf.add(bx_lambertian(m.Kd));
        bsdf.add(bx_phong(m.Ks, m.exponent));
        return true;
    }
    if (kind == OPR_MAT_TEXTURE) {
        const float4 v = tex2D<float4>(m.diffuseTex, sf.uv.x, sf.uv.y);
        albedo = f3(v.x, v.y, v.z);
        // the light program hands the attribute geometricNormal itself to the BSDF (Texture.cu:223)
        bsdf.init(isLight ? sf.geometricNormal : worldGeometricNormal, -rayDir, isLight);
        bsdf.add(bx_lambertian(albedo));
        return true;
    }
    if (kind == OPR_MAT_MIRROR) {
        if (isZero(m.Kr)) return false;
        bsdf.init(worldGeometricNormal, -rayDir, isLight);
        bsdf.add(bx_spec_refl(m.Kr, 0, 1.f, 1.f));
        return true;
    }
    // Glass
    const bool isHitFromOutside = dot(worldGeometricNormal, rayDir) < 0;
    const float3 N = isHitFromOutside ? worldGeometricNormal : -worldGeometricNormal;
    const float n1 = isHitFromOutside ? 1.f : m.ior, n2 = isHitFromOutside ? m.ior : 1.f;
    const bool krBlack = isZero(m.Kr), ktBlack = isZero(m.Kt);
    if (krBlack && ktBlack) return false;
    hitNormal = N;
    bsdf.init(N, -rayDir, isLight);
    if (!ktBlack) bsdf.add(bx_spec_trans(m.Kt, n1, n2));
    if (!krBlack) bsdf.add(bx_spec_refl(m.Kr, 1, n1, n2));
    return true;
}

// sampleScattering, vcm.h:166-204.  Returns true when a continuation ray must be traced.
OPR_DEV bool vcm_sample_scattering(const VcmBSDF& bsdf, Rng& rng, const VcmParams& P, float3& throughput, float& dVCM, float& dVC, float& dVM,
                                   const float3& hitPoint, float3& rayOrigin, float3& rayDir)
{
    const float contProb = bsdf.contProb;
    const float rrSample = rng_uniform(rng);
    if (contProb < rrSample) return false;
    float bsdfDirPdfW = 0.f, cosThetaOut = 0.f;
    const float3 bsdfSample = rng_uniform3(rng);
    int sampledEvent = B_Null;
    float3 newDir = rayDir; bool dirWritten = false;
    const float3 bsdfFactor = bsdf.vcmSampleF(&newDir, bsdfSample, &bsdfDirPdfW, &cosThetaOut, &sampledEvent, &dirWritten);
    if (dirWritten) rayDir = newDir;
    if (isZero(bsdfFactor)) return true;   // sic (vcm.h:188-189): not `done` — the previous origin is traced again
    float bsdfRevPdfW = bsdfDirPdfW;
    const bool isSpecularEvent = bx_match(sampledEvent, B_Specular);
    if (!isSpecularEvent) bsdfRevPdfW = bsdf.pdf(rayDir, B_All & ~B_Specular, true);
    bsdfDirPdfW *= contProb;
    bsdfRevPdfW *= contProb;
    mis_on_scatter(dVCM, dVC, dVM, cosThetaOut, bsdfDirPdfW, bsdfRevPdfW, P.misVcWeightFactor, P.misVmWeightFactor, sampledEvent);
    throughput *= bsdfFactor * (cosThetaOut / bsdfDirPdfW);
    rayOrigin = hitPoint;
    return true;
}

// ---- light pass ------------------------------------------------------------------------------------------------------
// lightPass prologue + initLightPayload (VCMLightPass.cu:60-72, :117-162), lightEmit (helpers/light.h:92-142),
// initLightMisTerms (mis.h:35-79)
__global__ void __launch_bounds__(VCM_THREADS) vcm_light_gen_kernel(SceneView sc, VcmParams P, RngBuf rngBuf, VcmBuffers B,
                                                                    float4* __restrict__ rayO, float4* __restrict__ rayD, unsigned* __restrict__ rayPath)
{
    const unsigned n = P.w * P.h;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        B.lvCount[p] = 0u;
        const size_t ri = (size_t)(p / P.w) * P.rsW + (p % P.w);
        Rng rng = rng_load(rngBuf, ri);
        int lightIndex = 0;
        if (1 < sc.numLights) {
            float sample = rng_uniform(rng);
            lightIndex = min((int)(sample * sc.numLights), sc.numLights - 1);
        }
        const opr_light light = sc.lights[lightIndex];
        const float lightPickPdf = 1.f / sc.numLights;
        float emissionPdfW = 0.f, directPdfW = 0.f, cosAtLight = 0.f;
        float3 origin = f3(0.f), direction = f3(0.f), radiance = f3(0.f);
        const float2 dirRnd = rng_uniform2(rng);
        const float3 lpos = f3(light.position[0], light.position[1], light.position[2]);
        if (light.lightType == OPR_LIGHT_AREA) {
            const float2 posRnd = rng_uniform2(rng);
            origin = lpos + posRnd.x * f3(light.v1[0], light.v1[1], light.v1[2]) + posRnd.y * f3(light.v2[0], light.v2[1], light.v2[2]);
            direction = sampleUnitHemisphereCos(f3(light.normal[0], light.normal[1], light.normal[2]), dirRnd, &emissionPdfW, &cosAtLight, true);
            emissionPdfW *= light.inverseArea;
            directPdfW = light.inverseArea;
            radiance = f3(light.Lemit[0], light.Lemit[1], light.Lemit[2]) * cosAtLight;
        } else if (light.lightType == OPR_LIGHT_POINT) {
            origin = lpos;
            float3 lightToSceneCenter = sc.boundingSphere.center - lpos;
            float lightSceneCenterDistance = length(lightToSceneCenter);
            lightToSceneCenter /= lightSceneCenterDistance;
            if (sc.boundingSphere.radius < lightSceneCenterDistance) {
                // sampleCone, helpers/samplers.h:132-155
                float theta = asinf(sc.boundingSphere.radius / lightSceneCenterDistance);
                float3 U, V; createCoordinateSystem(lightToSceneCenter, U, V);
                float cosTheta = cosf(theta);
                float z = cosTheta + (1.f - cosTheta) * dirRnd.x;
                float phi = 2 * PI_F * dirRnd.y;
                float r = sqrtf(1.f - z * z);
                float sp, cp; sincosf(phi, &sp, &cp);
                emissionPdfW = 1.f / (2.f * PI_F * (1.f - cosTheta));
                direction = normalize((r * cp) * U + z * lightToSceneCenter + (r * sp) * V);
            } else {
                direction = sampleUnitSphere(dirRnd, &emissionPdfW);
            }
            directPdfW = 1.f; cosAtLight = 1.f;
            radiance = f3(light.Lemit[0], light.Lemit[1], light.Lemit[2]);   // intensity
        }
        emissionPdfW *= lightPickPdf;
        directPdfW *= lightPickPdf;
        float3 throughput = radiance; throughput /= emissionPdfW;
        float dVCM = directPdfW / emissionPdfW, dVC = 0.f;
        if (!light.isDelta) { const float usedCosLight = light.isFinite ? cosAtLight : 1.f; dVC = usedCosLight / emissionPdfW; }
        const float dVM = dVC * P.misVcWeightFactor;
        rng_store(rngBuf, ri, rng);
        path_store(B, p, make_float4(throughput.x, throughput.y, throughput.z, __uint_as_float(0u)), make_float4(dVCM, dVC, dVM, 0.f));
        rayO[p] = make_float4(origin.x, origin.y, origin.z, RAY_LEN_MIN);
        rayD[p] = make_float4(direction.x, direction.y, direction.z, RT_DEFAULT_MAX);
        rayPath[p] = p;
    }
}

// vcmClosestHitLight + lightHit (vcm.h:210-309) + connectCameraT1 (vcm.h:65-161) + sampleScattering for ONE hit of light path p.
// DEFER = true (wavefront kernels): the camera connection is recorded in `conn`; false (tail kernel): tested in place.
// Returns true when a continuation ray (newO, newD) must be traced.
template <bool DEFER>
OPR_DEV bool vcm_light_hit(const SceneView& sc, const VcmParams& P, const RngBuf& rngBuf, const VcmBuffers& B, float* __restrict__ output, int kind, unsigned p,
                           const float3& o, const float3& d, const HitRec& hit, float3& newO, float3& newD, Connection& conn, unsigned& numConn,
                           unsigned& connTarget, float3& connOrigin, unsigned& shadowRays)
{
    bool push = false;
    if (B.dbgSig) B.dbgSig[p] = B.dbgSig[p] * 131u + (unsigned)kind + 1u;     // parity diagnostics only (null otherwise)
    const Surface sf = eval_surface(sc, hit, o, d);
    const DMaterial& m = sc.materials[sf.material];
    VcmBSDF bsdf; float3 hitNormal, albedo;
    if (vcm_build_bsdf(m, kind, sf, d, true, bsdf, hitNormal, albedo)) {
        const float3 hitPoint = o + hit.t * d;
        float4 t4, m4;
        path_load(B, p, t4, m4);
        float3 throughput = f3(t4);
        unsigned depth = __float_as_uint(t4.w);
        float dVCM = m4.x, dVC = m4.y, dVM = m4.z;
        depth++;
        const float cosThetaIn = dot(hitNormal, -d);
        bool done = false;
        if (cosThetaIn < EPS_COSINE) done = true;
        if (!done) {
            mis_on_hit(dVCM, dVC, dVM, cosThetaIn, hit.t);
            const bool isBsdfSpecular = bsdf.isSpecular();
            if (!isBsdfSpecular) {
                const unsigned k = B.lvCount[p];
                B.lvCount[p] = k + 1;
                if (!P.estimatePass) {
                    const size_t vs = (size_t)k * B.paths + p;
                    float4* V = B.lv + 6 * vs;      // one contiguous 96-byte record: three sectors for the camera pass instead of six
                    V[0] = make_float4(hitPoint.x, hitPoint.y, hitPoint.z, dVCM);
                    V[1] = make_float4(throughput.x, throughput.y, throughput.z, dVC);
                    V[2] = make_float4(bsdf.geometricNormal.x, bsdf.geometricNormal.y, bsdf.geometricNormal.z, dVM);
                    V[3] = make_float4(bsdf.localDirFix.x, bsdf.localDirFix.y, bsdf.localDirFix.z, __int_as_float(kind));
                    V[4] = make_float4(albedo.x, albedo.y, albedo.z, m.exponent);
                    V[5] = make_float4(m.Ks.x, m.Ks.y, m.Ks.z, __uint_as_float(depth));
                    // ---- connectCameraT1
                    const float3 eye = f3(P.camera.eye[0], P.camera.eye[1], P.camera.eye[2]);
                    const float3 lookdir = f3(P.camera.lookdir[0], P.camera.lookdir[1], P.camera.lookdir[2]);
                    float3 dirToCamera = eye - hitPoint;
                    if (!(dot(lookdir, -dirToCamera) <= 0.f)) {
                        const float distance = length(dirToCamera);
                        dirToCamera /= distance;
                        const float cosAtCamera = dot(normalize(lookdir), -dirToCamera);
                        const float imagePointToCameraDist = length(lookdir) / cosAtCamera;
                        const float3 imagePlanePointWorld = eye + (-dirToCamera * imagePointToCameraDist);
                        const float3 imagePlaneCenterWorld = eye + lookdir;
                        const float3 c2p = imagePlanePointWorld - imagePlaneCenterWorld;
                        const float3 unitCameraU = normalize(f3(P.camera.camera_u[0], P.camera.camera_u[1], P.camera.camera_u[2]));
                        const float3 unitCameraV = normalize(f3(P.camera.camera_v[0], P.camera.camera_v[1], P.camera.camera_v[2]));
                        const float pu = dot(c2p, unitCameraU), pv = dot(c2p, unitCameraV);
                        const float ipx = P.camera.imagePlaneSize[0], ipy = P.camera.imagePlaneSize[1];
                        if (2.f * fabsf(pu) < ipx && 2.f * fabsf(pv) < ipy) {
                            const float pcx = (pu + 0.5f * ipx) / ipx, pcy = (pv + 0.5f * ipy) / ipy;
                            unsigned px = (unsigned)(pcx * P.w), py = (unsigned)(pcy * P.h);
                            px = min(px, P.w - 1); py = min(py, P.h - 1);
                            float cosToCamera = 0.f, bsdfDirPdfW, bsdfRevPdfW;
                            const float3 bsdfFactor = bsdf.vcmF(dirToCamera, cosToCamera, &bsdfDirPdfW, &bsdfRevPdfW);
                            if (!isZero(bsdfFactor)) {
                                bsdfRevPdfW *= bsdf.contProb;
                                const float imageToSolidAngleFactor = sqr(imagePointToCameraDist) / cosAtCamera;
                                const float pixelArea = P.pixelSizeFactor.x * ipx * P.pixelSizeFactor.x * ipy;   // sic: .x twice (vcm.h:126)
                                const float imageSamplePdfA = 1.f / pixelArea;
                                const float cameraPdfW = imageSamplePdfA * imageToSolidAngleFactor;
                                const float cameraPdfA = cameraPdfW * fabsf(cosToCamera) / sqr(distance);
                                const float wLight = (cameraPdfA / P.lightSubpathCount) * (P.misVmWeightFactor + dVCM + dVC * bsdfRevPdfW);
                                const float misWeight = 1.f / (wLight + 1.f);
                                const float cameraPdfAConvertedToImagePdfA = 1.f / cameraPdfA;
                                const float3 contrib = misWeight * throughput * bsdfFactor / (P.lightSubpathCount * cameraPdfAConvertedToImagePdfA);
                                if (distance < 3.f * EPS_RAY || (!DEFER && !vcm_occluded_inline(sc, hitPoint, dirToCamera, distance, shadowRays))) {
                                    // the reference's splat is a plain, racy += (vcm.h:159); atomics here
                                    float* op = output + 3 * ((size_t)py * P.w + px);
                                    if (B.dbgConn) atomicAdd(B.dbgConn + ((size_t)py * P.w + px), 1u);
                                    atomicAdd(op, contrib.x); atomicAdd(op + 1, contrib.y); atomicAdd(op + 2, contrib.z);
                                } else if (DEFER) {
                                    conn.dir = dirToCamera; conn.tmax = distance - 2.f * EPS_RAY; conn.contrib = contrib;
                                    numConn = 1; connTarget = py * P.w + px; connOrigin = hitPoint;
                                }
                            }
                        }
                    }
                }
            }
            if (P.maxPathLen < depth + 2) done = true;
            if (!done) {
                const size_t ri = (size_t)(p / P.w) * P.rsW + (p % P.w);
                Rng rng = rng_load(rngBuf, ri);
                newO = o; newD = d;
                push = vcm_sample_scattering(bsdf, rng, P, throughput, dVCM, dVC, dVM, hitPoint, newO, newD);
                rng_store(rngBuf, ri, rng);
            }
        }
        path_store(B, p, make_float4(throughput.x, throughput.y, throughput.z, __uint_as_float(depth)), make_float4(dVCM, dVC, dVM, 0.f));
    }
    return push;
}

// wavefront shading of one light-pass bounce, one material kind per blockIdx.y
__global__ void __launch_bounds__(VCM_THREADS, 6) vcm_light_shade_kernel(SceneView sc, VcmParams P, RngBuf rngBuf, VcmBuffers B, float* __restrict__ output,
                                                                      const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                      const unsigned* __restrict__ rayPath, const float4* __restrict__ hits,
                                                                      const unsigned* __restrict__ matQueue, const unsigned* __restrict__ matCount, unsigned capacity,
                                                                      float4* __restrict__ nextO, float4* __restrict__ nextD, unsigned* __restrict__ nextPath,
                                                                      unsigned* __restrict__ nextCount, unsigned* __restrict__ stat)
{
    const int kind = blockIdx.y;
    if (kind == QUEUE_MISS || kind == OPR_MAT_DIFFUSE_EMITTER) return;   // miss / emitter: done (VCMLightPass.cu:95-98, DiffuseEmitter.cu:76-79)
    const unsigned n = matCount[kind];
    const unsigned* queue = matQueue + (size_t)kind * capacity;
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned qi = base + threadIdx.x;
        bool push = false;
        float3 newO = f3(0.f), newD = f3(0.f);
        unsigned p = 0;
        Connection conn; unsigned numConn = 0, connTarget = 0, shadowRays = 0; float3 connOrigin = f3(0.f);
        if (qi < n) {
            const unsigned slot = queue[qi];
            p = rayPath[slot];
            const float4 o4 = rayO[slot], d4 = rayD[slot], h4 = hits[slot];
            const float3 o = f3(o4), d = f3(d4);
            HitRec hit; hit.t = h4.x; hit.prim = __float_as_int(h4.y); hit.u = h4.z; hit.v = h4.w;
            push = vcm_light_hit<true>(sc, P, rngBuf, B, output, kind, p, o, d, hit, newO, newD, conn, numConn, connTarget, connOrigin, shadowRays);
        }
        warp_flush_connections(B, connOrigin, connTarget, &conn, numConn);
        const unsigned dst = warp_append(nextCount, push);
        if (push) {
            nextO[dst] = make_float4(newO.x, newO.y, newO.z, RAY_LEN_MIN);
            nextD[dst] = make_float4(newD.x, newD.y, newD.z, RT_DEFAULT_MAX);
            nextPath[dst] = p;
        }
    }
    (void)stat;
}

// ---- camera pass -----------------------------------------------------------------------------------------------------
// cameraPass prologue + initCameraPayload (VCMCameraPass.cu:52-60, :108-146), initCameraMisTerms (mis.h:83-103)
__global__ void __launch_bounds__(VCM_THREADS) vcm_camera_gen_kernel(VcmParams P, RngBuf rngBuf, VcmBuffers B,
                                                                     float4* __restrict__ rayO, float4* __restrict__ rayD, unsigned* __restrict__ rayPath)
{
    const unsigned n = P.w * P.h;
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const unsigned p = tiled_pixel(s, P.w, P.h);      // the camera path keeps its index (= pixel); only its queue slot moves
        const unsigned x = p % P.w, y = p / P.w;
        const size_t ri = (size_t)y * P.rsW + x;
        Rng rng = rng_load(rngBuf, ri);
        const float2 sample = rng_uniform2(rng);
        rng_store(rngBuf, ri, rng);
        const float dx = ((float)x + sample.x) / (float)P.w * 2.0f - 1.0f;
        const float dy = ((float)y + sample.y) / (float)P.h * 2.0f - 1.0f;
        const float3 eye = f3(P.camera.eye[0], P.camera.eye[1], P.camera.eye[2]);
        const float3 lookdir = f3(P.camera.lookdir[0], P.camera.lookdir[1], P.camera.lookdir[2]);
        const float3 dir = normalize(dx * f3(P.camera.camera_u[0], P.camera.camera_u[1], P.camera.camera_u[2]) +
                                     dy * f3(P.camera.camera_v[0], P.camera.camera_v[1], P.camera.camera_v[2]) + lookdir);
        const float cosAtCamera = dot(normalize(lookdir), dir);
        const float imagePointToCameraDist = length(lookdir) / cosAtCamera;
        const float imageToSolidAngleFactor = sqr(imagePointToCameraDist) / cosAtCamera;
        const float pixelArea = P.pixelSizeFactor.x * P.camera.imagePlaneSize[0] * P.pixelSizeFactor.x * P.camera.imagePlaneSize[1];   // sic (:138)
        const float cameraPdfW = (1.f / pixelArea) * imageToSolidAngleFactor;
        path_store(B, p, make_float4(1.f, 1.f, 1.f, __uint_as_float(0u)), make_float4(P.lightSubpathCount / cameraPdfW, 0.f, 0.f, 0.f));
        B.color[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        rayO[s] = make_float4(eye.x, eye.y, eye.z, RAY_LEN_MIN);
        rayD[s] = make_float4(dir.x, dir.y, dir.z, RT_DEFAULT_MAX);
        rayPath[s] = p;
    }
}

// vcmClosestHitCamera + cameraHit (vcm.h:527-628): connectLightSourceS1 (:406-488, lightIlluminate helpers/light.h:147-213),
// connectVertices against the same-index light path (:315-400, :607-615), connectLightSourceS0 on emitters
// (:493-522, DiffuseEmitter.cu:95-120), sampleScattering.
// for ONE hit of camera path p.  DEFER as in vcm_light_hit.  Returns true when a continuation ray must be traced.
template <bool DEFER>
OPR_DEV bool vcm_camera_hit(const SceneView& sc, const VcmParams& P, const RngBuf& rngBuf, const VcmBuffers& B, int kind, unsigned p, const float3& o,
                            const float3& d, const HitRec& hit, float3& newO, float3& newD, Connection* conn, unsigned& numConn, float3& connOrigin,
                            unsigned& shadowRays)
{
    bool push = false;
    if (B.dbgSig) B.dbgSig[B.paths + p] = B.dbgSig[B.paths + p] * 131u + (unsigned)kind + 1u;     // parity diagnostics only (null otherwise)
    const Surface sf = eval_surface(sc, hit, o, d);
    const DMaterial& m = sc.materials[sf.material];
    float4 t4, m4;
    path_load(B, p, t4, m4);
    float3 throughput = f3(t4);
    unsigned depth = __float_as_uint(t4.w);
    float dVCM = m4.x, dVC = m4.y, dVM = m4.z;
    float3 color = f3(B.color[p]);
    if (kind == OPR_MAT_DIFFUSE_EMITTER) {
        depth++;
        if (!isZero(m.Lemit)) {
            const float3 worldGeometricNormal = normalize(sf.geometricNormal);
            if (!(dot(worldGeometricNormal, -d) < 0.f)) {
                const float lightPickProb = 1.f / sc.numLights;
                float directPdfA = m.inverseArea;
                float emissionPdfW = maxf_(0.f, dot(worldGeometricNormal, -d)) * INV_PI_F * m.inverseArea;
                if (depth == 1) color += throughput * m.Lemit;
                else {
                    directPdfA *= lightPickProb; emissionPdfW *= lightPickProb;
                    const float wCamera = directPdfA * dVCM + emissionPdfW * dVC;
                    const float misWeight = 1.f / (1.f + wCamera);
                    color += throughput * misWeight * m.Lemit;
                }
            }
        }
    } else {
        VcmBSDF bsdf; float3 hitNormal, albedo;
        if (vcm_build_bsdf(m, kind, sf, d, false, bsdf, hitNormal, albedo)) {
            const float3 hitPoint = o + hit.t * d;
            connOrigin = hitPoint;
            depth++;
            const float cosThetaIn = dot(hitNormal, -d);
            bool done = cosThetaIn < EPS_COSINE;
            if (!done) {
                mis_on_hit(dVCM, dVC, dVM, cosThetaIn, hit.t);
                const bool isBsdfSpecular = bsdf.isSpecular();
                const size_t ri = (size_t)(p / P.w) * P.rsW + (p % P.w);
                Rng rng = rng_load(rngBuf, ri);
                if (!isBsdfSpecular) {
                    // ---- connectLightSourceS1
                    int lightIndex = 0;
                    if (1 < sc.numLights) {
                        float sample = rng_uniform(rng);
                        lightIndex = min((int)(sample * sc.numLights), sc.numLights - 1);
                    }
                    const opr_light light = sc.lights[lightIndex];
                    const float lightPickProb = 1.f / sc.numLights;
                    float emissionPdfW = 0.f, directPdfW = 0.f, cosAtLight = 0.f, distance = 0.f;
                    float3 dirToLight = f3(0.f), radiance = f3(0.f);
                    const float3 lpos = f3(light.position[0], light.position[1], light.position[2]);
                    if (light.lightType == OPR_LIGHT_AREA) {
                        const float2 posRnd = rng_uniform2(rng);
                        const float3 pointOnLight = lpos + posRnd.x * f3(light.v1[0], light.v1[1], light.v1[2]) + posRnd.y * f3(light.v2[0], light.v2[1], light.v2[2]);
                        dirToLight = pointOnLight - hitPoint;
                        distance = length(dirToLight);
                        dirToLight /= distance;
                        const float cosThetaLight = dot(f3(light.normal[0], light.normal[1], light.normal[2]), -dirToLight);
                        if (!(cosThetaLight < EPS_COSINE)) {
                            directPdfW = light.inverseArea * sqr(distance) / cosThetaLight;
                            cosAtLight = cosThetaLight;
                            emissionPdfW = light.inverseArea * cosThetaLight * INV_PI_F;
                            radiance = f3(light.Lemit[0], light.Lemit[1], light.Lemit[2]);
                        }
                    } else if (light.lightType == OPR_LIGHT_POINT) {
                        dirToLight = lpos - hitPoint;
                        distance = length(dirToLight);
                        dirToLight /= distance;
                        const float lightSceneCenterDistance = length(sc.boundingSphere.center - lpos);
                        if (sc.boundingSphere.radius < lightSceneCenterDistance) {
                            const float theta = asinf(sc.boundingSphere.radius / lightSceneCenterDistance);
                            emissionPdfW = 1.f / (2.f * PI_F * (1.f - cosf(theta)));
                        } else emissionPdfW = 0.25f * INV_PI_F;
                        directPdfW = sqr(distance);
                        cosAtLight = 1.f;
                        radiance = f3(light.Lemit[0], light.Lemit[1], light.Lemit[2]);
                    }
                    if (!isZero(radiance)) {
                        float bsdfDirPdfW, bsdfRevPdfW, cosToLight = 0.f;
                        const float3 bsdfFactor = bsdf.vcmF(dirToLight, cosToLight, &bsdfDirPdfW, &bsdfRevPdfW);
                        if (!isZero(bsdfFactor)) {
                            const float contiueProb = bsdf.contProb;
                            bsdfDirPdfW *= light.isDelta ? 0.f : contiueProb;
                            bsdfRevPdfW *= contiueProb;
                            const float wLight = bsdfDirPdfW / (lightPickProb * directPdfW);
                            const float wCamera = (emissionPdfW * cosToLight / (directPdfW * cosAtLight)) * (P.misVmWeightFactor + dVCM + dVC * bsdfRevPdfW);
                            const float misWeight = 1.f / (wLight + 1.f + wCamera);
                            const float3 contrib = (misWeight * cosToLight / (lightPickProb * directPdfW)) * (radiance * bsdfFactor);
                            if (!isZero(contrib)) {
                                if (distance < 3.f * EPS_RAY || (!DEFER && !vcm_occluded_inline(sc, hitPoint, dirToLight, distance, shadowRays))) { color += contrib * throughput; if (B.dbgConn) atomicAdd(B.dbgConn + p, 1u); }
                                else if (DEFER) { conn[numConn].dir = dirToLight; conn[numConn].tmax = distance - 2.f * EPS_RAY; conn[numConn].contrib = contrib * throughput; numConn++; }
                            }
                        }
                    }
                    // ---- connectVertices with the light path of the same index
                    const unsigned pathIndex = p % P.lightSubpathCount;
                    const unsigned lightSubpathVertexCount = B.lvCount[pathIndex];
                    for (unsigned i = 0; i < lightSubpathVertexCount; ++i) {
                        const size_t vs = (size_t)i * B.paths + pathIndex;
                        const float4* V = B.lv + 6 * vs;
                        const Float8 v01 = ldg256(V);
                        const float4 a0 = v01.lo, a1 = v01.hi;
                        float3 direction = f3(a0) - hitPoint;
                        const float dist2 = dot(direction, direction);
                        const float distance2 = sqrtf(dist2);
                        direction /= distance2;
                        float cameraCosTheta = 0.f, cameraBsdfDirPdfW, cameraBsdfRevPdfW;
                        const float3 cameraBsdfFactor = bsdf.vcmF(direction, cameraCosTheta, &cameraBsdfDirPdfW, &cameraBsdfRevPdfW);
                        if (isZero(cameraBsdfFactor)) continue;
                        const float cameraCont = bsdf.contProb;
                        cameraBsdfDirPdfW *= cameraCont; cameraBsdfRevPdfW *= cameraCont;
                        // rebuild the light vertex' BSDF from its 96-byte record
                        const Float8 v23 = ldg256(V + 2), v45 = ldg256(V + 4);
                        const float4 a2 = v23.lo, a3 = v23.hi, a4 = v45.lo, a5 = v45.hi;
                        VcmBSDF lb;
                        lb.initLocal(f3(a2), f3(a3), true);
                        const int lkind = __float_as_int(a3.w);
                        lb.add(bx_lambertian(f3(a4)));
                        if (lkind == OPR_MAT_GLOSSY) lb.add(bx_phong(f3(a5), a4.w));
                        float lightCosTheta = 0.f, lightBsdfDirPdfW, lightBsdfRevPdfW;
                        const float3 lightBsdfFactor = lb.vcmF(-direction, lightCosTheta, &lightBsdfDirPdfW, &lightBsdfRevPdfW);
                        if (isZero(lightBsdfFactor)) continue;
                        const float lightCont = lb.contProb;
                        lightBsdfDirPdfW *= lightCont; lightBsdfRevPdfW *= lightCont;
                        const float geometryTerm = lightCosTheta * cameraCosTheta / dist2;
                        if (geometryTerm < 0.f) continue;
                        const float cameraBsdfDirPdfA = pdfWtoA(cameraBsdfDirPdfW, distance2, cameraCosTheta);
                        const float lightBsdfDirPdfA = pdfWtoA(lightBsdfDirPdfW, distance2, lightCosTheta);
                        const float wLight = cameraBsdfDirPdfA * (P.misVmWeightFactor * 1.f + a0.w + a1.w * lightBsdfRevPdfW);
                        const float wCamera = lightBsdfDirPdfA * (P.misVmWeightFactor * 1.f + dVCM + dVC * cameraBsdfRevPdfW);
                        const float misWeight = 1.f / (wLight + 1.f + wCamera);
                        float3 contrib = geometryTerm * cameraBsdfFactor * lightBsdfFactor * 1.f;
                        contrib *= misWeight * throughput * f3(a1);
                        if (distance2 < 3.f * EPS_RAY || (!DEFER && !vcm_occluded_inline(sc, hitPoint, direction, distance2, shadowRays))) { color += contrib; if (B.dbgConn) atomicAdd(B.dbgConn + p, 1u); }
                        else if (DEFER && numConn < VCM_MAX_CONNECTIONS) { conn[numConn].dir = direction; conn[numConn].tmax = distance2 - 2.f * EPS_RAY; conn[numConn].contrib = contrib; numConn++; }
                    }
                }
                if (P.maxPathLen <= depth) done = true;
                if (!done) {
                    newO = o; newD = d;
                    push = vcm_sample_scattering(bsdf, rng, P, throughput, dVCM, dVC, dVM, hitPoint, newO, newD);
                }
                rng_store(rngBuf, ri, rng);
            }
        }
    }
    path_store(B, p, make_float4(throughput.x, throughput.y, throughput.z, __uint_as_float(depth)), make_float4(dVCM, dVC, dVM, 0.f));
    B.color[p] = make_float4(color.x, color.y, color.z, 0.f);
    return push;
}

// wavefront shading of one camera-pass bounce, one material kind per blockIdx.y
#ifndef OPR_VCM_CAMERA_MINBLOCKS
#define OPR_VCM_CAMERA_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(VCM_THREADS, OPR_VCM_CAMERA_MINBLOCKS) vcm_camera_shade_kernel(SceneView sc, VcmParams P, RngBuf rngBuf, VcmBuffers B,
                                                                       const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                       const unsigned* __restrict__ rayPath, const float4* __restrict__ hits,
                                                                       const unsigned* __restrict__ matQueue, const unsigned* __restrict__ matCount, unsigned capacity,
                                                                       float4* __restrict__ nextO, float4* __restrict__ nextD, unsigned* __restrict__ nextPath,
                                                                       unsigned* __restrict__ nextCount, unsigned* __restrict__ stat)
{
    const int kind = blockIdx.y;
    if (kind == QUEUE_MISS) return;   // VCMCameraPass.cu:87-90
    const unsigned n = matCount[kind];
    const unsigned* queue = matQueue + (size_t)kind * capacity;
    for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const unsigned qi = base + threadIdx.x;
        bool push = false;
        float3 newO = f3(0.f), newD = f3(0.f);
        unsigned p = 0;
        Connection conn[VCM_MAX_CONNECTIONS]; unsigned numConn = 0, shadowRays = 0; float3 connOrigin = f3(0.f);
        if (qi < n) {
            const unsigned slot = queue[qi];
            p = rayPath[slot];
            const float4 o4 = rayO[slot], d4 = rayD[slot], h4 = hits[slot];
            const float3 o = f3(o4), d = f3(d4);
            HitRec hit; hit.t = h4.x; hit.prim = __float_as_int(h4.y); hit.u = h4.z; hit.v = h4.w;
            push = vcm_camera_hit<true>(sc, P, rngBuf, B, kind, p, o, d, hit, newO, newD, conn, numConn, connOrigin, shadowRays);
        }
        warp_flush_connections(B, connOrigin, p, conn, numConn);
        const unsigned dst = warp_append(nextCount, push);
        if (push) {
            nextO[dst] = make_float4(newO.x, newO.y, newO.z, RAY_LEN_MIN);
            nextD[dst] = make_float4(newD.x, newD.y, newD.z, RT_DEFAULT_MAX);
            nextPath[dst] = p;
        }
    }
    (void)stat;
}

// ---- tails of the two passes: the paths still alive after the wavefront bounces are finished one thread per path (trace ->
// material program -> connections tested in place), as in the PPM passes (ppm_camera_tail_kernel): late bounces hold a few per
// cent of the paths, and a wavefront bounce over them costs three launches whose duration is one ray's dependent-load latency.
template <bool QN>
__global__ void __launch_bounds__(VCM_THREADS) vcm_light_tail_kernel(SceneView sc, VcmParams P, RngBuf rngBuf, VcmBuffers B, float* __restrict__ output,
                                                                     const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                     const unsigned* __restrict__ rayPath, const unsigned* __restrict__ count,
                                                                     unsigned* __restrict__ stats, unsigned* __restrict__ cursor, unsigned refillIdle)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n = *count;
    unsigned closest = 0, any = 0, p = 0;
    bool alive = false, exhausted = false;
    float3 o = f3(0.f), d = f3(0.f, 0.f, 1.f);
    for (;;) {
        unsigned i = 0;
        if (tail_refill(n, cursor, alive, exhausted, refillIdle, i)) {     // path regeneration (common.cuh)
            p = rayPath[i]; o = f3(rayO[i]); d = f3(rayD[i]); alive = true;
        }
        if (!__any_sync(FULL, alive)) break;
        if (alive) {
            HitRec hit;
            traverse_free<false, QN>(sc, o, d, RAY_LEN_MIN, RT_DEFAULT_MAX, hit);
            closest++;
            if (hit.prim < 0) alive = false;                                      // VCMLightPass.cu:95-98
            else {
                const int kind = sc.materials[prim_material(sc, hit.prim)].kind;
                if (kind == OPR_MAT_DIFFUSE_EMITTER) alive = false;               // DiffuseEmitter.cu:76-79
                else {
                    Connection conn; unsigned numConn = 0, connTarget = 0; float3 connOrigin = f3(0.f), newO = f3(0.f), newD = f3(0.f);
                    alive = vcm_light_hit<false>(sc, P, rngBuf, B, output, kind, p, o, d, hit, newO, newD, conn, numConn, connTarget, connOrigin, any);
                    o = newO; d = newD;
                }
            }
        }
    }
    for (int s = 16; s > 0; s >>= 1) { closest += __shfl_xor_sync(FULL, closest, s); any += __shfl_xor_sync(FULL, any, s); }
    if ((threadIdx.x & 31) == 0) { if (closest) atomicAdd(stats + STAT_CLOSEST_RAYS, closest); if (any) atomicAdd(stats + STAT_ANY_RAYS, any); }
}

template <bool QN>
__global__ void __launch_bounds__(VCM_THREADS) vcm_camera_tail_kernel(SceneView sc, VcmParams P, RngBuf rngBuf, VcmBuffers B,
                                                                      const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                                      const unsigned* __restrict__ rayPath, const unsigned* __restrict__ count,
                                                                      unsigned* __restrict__ stats, unsigned* __restrict__ cursor, unsigned refillIdle)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n = *count;
    unsigned closest = 0, any = 0, p = 0;
    bool alive = false, exhausted = false;
    float3 o = f3(0.f), d = f3(0.f, 0.f, 1.f);
    for (;;) {
        unsigned i = 0;
        if (tail_refill(n, cursor, alive, exhausted, refillIdle, i)) {     // path regeneration (common.cuh)
            p = rayPath[i]; o = f3(rayO[i]); d = f3(rayD[i]); alive = true;
        }
        if (!__any_sync(FULL, alive)) break;
        if (alive) {
            HitRec hit;
            traverse_free<false, QN>(sc, o, d, RAY_LEN_MIN, RT_DEFAULT_MAX, hit);
            closest++;
            if (hit.prim < 0) alive = false;                                      // VCMCameraPass.cu:87-90
            else {
                const int kind = sc.materials[prim_material(sc, hit.prim)].kind;
                unsigned numConn = 0; float3 connOrigin = f3(0.f), newO = f3(0.f), newD = f3(0.f);
                alive = vcm_camera_hit<false>(sc, P, rngBuf, B, kind, p, o, d, hit, newO, newD, nullptr, numConn, connOrigin, any);
                o = newO; d = newD;
            }
        }
    }
    for (int s = 16; s > 0; s >>= 1) { closest += __shfl_xor_sync(FULL, closest, s); any += __shfl_xor_sync(FULL, any, s); }
    if ((threadIdx.x & 31) == 0) { if (closest) atomicAdd(stats + STAT_CLOSEST_RAYS, closest); if (any) atomicAdd(stats + STAT_ANY_RAYS, any); }
}

// cameraPass epilogue: outputBuffer[launchIndex] += cameraPrd.color (VCMCameraPass.cu:75-78)
__global__ void vcm_accumulate_kernel(VcmBuffers B, unsigned n, float* __restrict__ output)
{
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float4 c = B.color[p];
        float* o = output + 3 * (size_t)p;
        o[0] = o[0] + c.x; o[1] = o[1] + c.y; o[2] = o[2] + c.z;
    }
}

int Renderer::debugConnectionCountsGet(unsigned* dstHost)
{
    if (!vcm || !vcm->dbgConn) return 1;
    OPR_CUDA_CHECK(cudaMemcpyAsync(dstHost, vcm->dbgConn, vcm->paths * 4, cudaMemcpyDeviceToHost, stream));
    OPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

int Renderer::debugPathSignaturesGet(unsigned* dstHost)
{
    if (!vcm || !vcm->dbgSig) return 1;
    OPR_CUDA_CHECK(cudaMemcpyAsync(dstHost, vcm->dbgSig, vcm->paths * 8, cudaMemcpyDeviceToHost, stream));
    OPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

int Renderer::debugLightVertexCounts(unsigned* dstHost)
{
    if (!vcm) return 1;
    OPR_CUDA_CHECK(cudaMemcpyAsync(dstHost, vcm->lvCount, vcm->paths * 4, cudaMemcpyDeviceToHost, stream));
    OPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

// Parity diagnostics: (owning path, hit point) of every stored light vertex of the last light pass, 16 bytes each, path-major.
// Returns the number of vertices (<= capacityVerts are written).
long long Renderer::debugLightVertexPositions(float* dstHost, size_t capacityVerts)
{
    if (!vcm) return -1;
    const size_t paths = vcm->paths;
    std::vector<unsigned> cnt(paths);
    OPR_CUDA_CHECK(cudaMemcpyAsync(cnt.data(), vcm->lvCount, paths * 4, cudaMemcpyDeviceToHost, stream));
    OPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    std::vector<float4> lv((size_t)paths * vcm->maxVerts * 6);
    OPR_CUDA_CHECK(cudaMemcpy(lv.data(), vcm->lv, lv.size() * sizeof(float4), cudaMemcpyDeviceToHost));
    size_t n = 0;
    for (size_t p = 0; p < paths; p++)
        for (unsigned k = 0; k < cnt[p] && k < vcm->maxVerts; k++, n++) {
            if (n >= capacityVerts) continue;
            const float4 v = lv[6 * ((size_t)k * paths + p)];
            const unsigned path = (unsigned)p;
            memcpy(dstHost + 4 * n, &path, 4);
            dstHost[4 * n + 1] = v.x; dstHost[4 * n + 2] = v.y; dstHost[4 * n + 3] = v.z;
        }
    return (long long)n;
}

// ---- host orchestration: OptixRenderer.cpp:675-797 ---------------------------------------------------------------------
void Renderer::renderVCM(unsigned long long, unsigned localIterationNumber, float PPMRadius, const opr_request_details& details)
{
    LaunchCtx lc{stream, &launches};
    const SceneView& sc = scene.view;
    vcm_alloc(*this);
    VcmState& S = *vcm;
    if (debugConnectionCounts && !S.dbgConn) { devAlloc((void**)&S.dbgConn, S.paths * 4); devAlloc((void**)&S.dbgSig, S.paths * 8); }
    VcmBuffers B{S.state, S.color, S.lvCount, S.lv, (unsigned)S.paths,
                 S.connO, S.connD, S.connC, q.counters + CTR_CONN_COUNT, debugConnectionCounts ? S.dbgConn : nullptr, debugConnectionCounts ? S.dbgSig : nullptr};
    const unsigned lightSubPathCount = width * height;
    const float ppmRadiusSquared = PPMRadius * PPMRadius;
    const float etaVCM = (float(lightSubPathCount) / 1u) * PI_F * ppmRadiusSquared;   // nVM / nVC * pi * r^2 (:687)
    VcmParams P;
    P.camera = details.camera; P.w = width; P.h = height; P.rsW = rsW;
    P.maxPathLen = vcmMaxPathLength; P.estimatePass = 0; P.lightSubpathCount = lightSubPathCount;
    P.misVmWeightFactor = 0.f;                  // m_vcmUseVM = false (:301)
    P.misVcWeightFactor = 1.f / etaVCM;         // m_vcmUseVC = true
    P.pixelSizeFactor = make_float2(1.0f / width, 1.0f / height);   // resizeBuffers :846

    // Connection queues: with the stream pipeline on, bounce b's shadow segments are tested on stream B while `stream` traces and
    // shades bounce b+1 into the other queue.  The camera pass then adds the unoccluded contributions to the output image (float
    // atomics, like the light pass' splats) instead of the per-path colour that the next bounce's shading reads and rewrites.
    const bool ov = overlap && !timing && S.conn2[0] != nullptr;
    VcmBuffers Bq[2] = {B, B};
    Bq[1].connO = S.conn2[0]; Bq[1].connD = S.conn2[1]; Bq[1].connC = S.conn2[2]; Bq[1].connCount = q.counters + CTR_CONN2_COUNT;
    unsigned* const connFetch[2] = {q.counters + CTR_CONN_FETCH, q.counters + CTR_CONN2_FETCH};
    int connSet = 0;
    bool connBusy[2] = {false, false};
    if (ov) joinStreams();
    auto beginConnections = [&]() -> const VcmBuffers& {      // before the shading kernel that fills the queue
        if (ov) { connSet ^= 1; if (connBusy[connSet]) { OPR_CUDA_CHECK(cudaStreamWaitEvent(stream, evVcmConn[connSet], 0)); connBusy[connSet] = false; } }
        OPR_CUDA_CHECK(cudaMemsetAsync(Bq[connSet].connCount, 0, sizeof(unsigned) * 3, stream));   // count, (legacy fetch), connection fetch
        return Bq[connSet];
    };
    auto traceConnections = [&](float* target, unsigned stride) {
        const VcmBuffers& b = Bq[connSet];
        if (!ov) {
            launch_trace_connect_queue(sc, b.connO, b.connD, b.connC, b.connCount, target, stride, connFetch[connSet], q.counters + CTR_STATS + STAT_ANY_RAYS, lc, 0, b.dbgConn,
                                       (qnodeMask & QN_VCM_CONNECT) != 0);
            return;
        }
        OPR_CUDA_CHECK(cudaEventRecord(evVcmShade, stream));
        OPR_CUDA_CHECK(cudaStreamWaitEvent(streamB, evVcmShade, 0));
        launch_trace_connect_queue(sc, b.connO, b.connD, b.connC, b.connCount, target, stride, connFetch[connSet], q.counters + CTR_STATS + STAT_ANY_RAYS,
                                   LaunchCtx{streamB, &launches}, vcmConnectBlocksPerSM, b.dbgConn, (qnodeMask & QN_VCM_CONNECT) != 0);
        OPR_CUDA_CHECK(cudaEventRecord(evVcmConn[connSet], streamB));
        connBusy[connSet] = true;
    };
    auto joinConnections = [&]() {
        for (int s = 0; s < 2; s++) if (connBusy[s]) { OPR_CUDA_CHECK(cudaStreamWaitEvent(stream, evVcmConn[s], 0)); connBusy[s] = false; }
    };

    auto lightPass = [&](const char* name) {
        NvtxRange nv("OptixEntryPoint::VCM_LIGHT_PASS");
        ScopedTimer t(*this, name);
        unsigned c[2 + NUM_QUEUES] = {0};
        c[0] = width * height;
        OPR_CUDA_CHECK(cudaMemcpyAsync(q.counters, c, sizeof c, cudaMemcpyHostToDevice, stream));
        int blocks = (int)std::min<unsigned>((c[0] + VCM_THREADS - 1) / VCM_THREADS, numSMs() * 16);
        vcm_light_gen_kernel<<<blocks, VCM_THREADS, 0, stream>>>(sc, P, rng, B, q.rayO[0], q.rayD[0], q.rayPath[0]);
        ++launches;
        int cur = 0;
        const unsigned waves = std::min(vcmLightWaveBounces, P.maxPathLen + 1);
        for (unsigned bounce = 0; bounce < waves; bounce++) {
            OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + 2, 0, sizeof(unsigned) * NUM_QUEUES, stream));
            OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + (cur ^ 1), 0, sizeof(unsigned), stream));
            launch_trace_closest(sc, q, cur, lc, bounce > 0 && (qnodeMask & QN_VCM_BOUNCES) != 0);
            const VcmBuffers& Bc = beginConnections();
            dim3 grid(VCM_BLOCKS_X, NUM_QUEUES);
            vcm_light_shade_kernel<<<grid, VCM_THREADS, 0, stream>>>(sc, P, rng, Bc, output, q.rayO[cur], q.rayD[cur], q.rayPath[cur], q.hits, q.matQueue,
                                                                     q.counters + 2, q.capacity, q.rayO[cur ^ 1], q.rayD[cur ^ 1], q.rayPath[cur ^ 1],
                                                                     q.counters + (cur ^ 1), q.counters + 16 + STAT_ANY_RAYS);
            ++launches;
            if (!P.estimatePass)   // connectCameraT1 splats of the unoccluded segments (vcm.h:159)
                traceConnections(output, 3);
            cur ^= 1;
        }
        if (waves <= P.maxPathLen) {
            auto kt = (qnodeMask & QN_VCM_TAILS) ? vcm_light_tail_kernel<true> : vcm_light_tail_kernel<false>;
            OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + CTR_TAIL_FETCH, 0, sizeof(unsigned), stream));
            kt<<<numSMs() * 4, VCM_THREADS, 0, stream>>>(sc, P, rng, B, output, q.rayO[cur], q.rayD[cur], q.rayPath[cur], q.counters + cur,
                                                                       q.counters + CTR_STATS, q.counters + CTR_TAIL_FETCH, 8u);
            ++launches;
        }
    };

    if (!lightVertexCountEstimated) {
        // The reference sizes its light-vertex cache from an extra light pass (:699-763).  Slots are per path here, so
        // no sizing is needed, but the pass is kept: it consumes the RNG states exactly as the reference does.
        P.maxPathLen = vcmMaxPathLength; P.estimatePass = 1;
        lightPass("vcm.estimate");
        P.estimatePass = 0;
        lightVertexCountEstimated = true;
    }
    if (B.dbgConn) { OPR_CUDA_CHECK(cudaMemsetAsync(B.dbgConn, 0, S.paths * 4, stream)); OPR_CUDA_CHECK(cudaMemsetAsync(B.dbgSig, 0, S.paths * 8, stream)); }
    lightPass("vcm.light");
    {
        NvtxRange nv("OptixEntryPoint::VCM_CAMERA_PASS");
        ScopedTimer t(*this, "vcm.camera");
        unsigned c[2 + NUM_QUEUES] = {0};
        c[0] = width * height;
        OPR_CUDA_CHECK(cudaMemcpyAsync(q.counters, c, sizeof c, cudaMemcpyHostToDevice, stream));
        int blocks = (int)std::min<unsigned>((c[0] + VCM_THREADS - 1) / VCM_THREADS, numSMs() * 16);
        vcm_camera_gen_kernel<<<blocks, VCM_THREADS, 0, stream>>>(P, rng, B, q.rayO[0], q.rayD[0], q.rayPath[0]);
        ++launches;
        int cur = 0;
        const unsigned waves = std::min(vcmCameraWaveBounces, P.maxPathLen + 1);
        for (unsigned bounce = 0; bounce < waves; bounce++) {
            OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + 2, 0, sizeof(unsigned) * NUM_QUEUES, stream));
            OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + (cur ^ 1), 0, sizeof(unsigned), stream));
            launch_trace_closest(sc, q, cur, lc, bounce > 0 && (qnodeMask & QN_VCM_BOUNCES) != 0);
            const VcmBuffers& Bc = beginConnections();
            dim3 grid(VCM_BLOCKS_X, NUM_QUEUES);
            vcm_camera_shade_kernel<<<grid, VCM_THREADS, 0, stream>>>(sc, P, rng, Bc, q.rayO[cur], q.rayD[cur], q.rayPath[cur], q.hits, q.matQueue,
                                                                      q.counters + 2, q.capacity, q.rayO[cur ^ 1], q.rayD[cur ^ 1], q.rayPath[cur ^ 1],
                                                                      q.counters + (cur ^ 1), q.counters + 16 + STAT_ANY_RAYS);
            ++launches;
            // light-source and vertex connections of this bounce: add the unoccluded ones to the paths' colour accumulators
            if (ov) traceConnections(output, 3); else traceConnections((float*)S.color, 4);
            cur ^= 1;
        }
        if (waves <= P.maxPathLen) {
            auto kt = (qnodeMask & QN_VCM_TAILS) ? vcm_camera_tail_kernel<true> : vcm_camera_tail_kernel<false>;
            OPR_CUDA_CHECK(cudaMemsetAsync(q.counters + CTR_TAIL_FETCH, 0, sizeof(unsigned), stream));
            kt<<<numSMs() * 4, VCM_THREADS, 0, stream>>>(sc, P, rng, B, q.rayO[cur], q.rayD[cur], q.rayPath[cur], q.counters + cur, q.counters + CTR_STATS,
                                                                        q.counters + CTR_TAIL_FETCH, 8u);
            ++launches;
        }
        joinConnections();
        vcm_accumulate_kernel<<<blocks, VCM_THREADS, 0, stream>>>(B, width * height, output);
        ++launches;
    }
    (void)localIterationNumber;
}

} // namespace opr
