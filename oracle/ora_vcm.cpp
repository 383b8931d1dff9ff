// ORACLE — test infrastructure only (see ora_math.h).  Vertex-connection BDPT of the reference restated on
// the CPU: renderer/vcm/{VCMLightPass.cu,VCMCameraPass.cu,vcm.h,mis.h}, renderer/{BSDF,BxDF,reflection}.h,
// math/DifferentialGeometry.h, helpers/light.h:92-213 and the VCM closest-hit programs of material/*.cu.
#include "ora_render.h"
#include <cstring>

namespace ora {

float3 sampleUnitHemisphereCos(const float3& normal, const float2& sample, float* oPdfW, float* oCosTheta, bool biasSmallCosine);   // ora_ppm.cpp
float3 sampleUnitSphere(const float2& sample, float* oPdfW);
float4 tex2D(const Texture& t, float u, float v);

static inline float sqr(float a) { return a * a; }
static inline float maxf(float a, float b) { return a > b ? a : b; }
static inline bool isZero(const float3& v) { return v.x == 0.f && v.y == 0.f && v.z == 0.f; }
static inline float vcmMis(float a) { return a; }   // vcm/vcm_shared.h:13-17 (balance heuristic)
static inline int intmin(int a, int b) { return a < b ? a : b; }

// ---- helpers/samplers.h:96-171 ------------------------------------------------------------------------
static void createCoordinateSystem(const float3& N, float3& U, float3& V)
{
    if (fabsf(N.x) > fabsf(N.y)) { float invLength = 1.f / sqrtf(N.x * N.x + N.z * N.z); U = make_float3(-N.z * invLength, 0.f, N.x * invLength); }
    else { float invLength = 1.f / sqrtf(N.y * N.y + N.z * N.z); U = make_float3(0.f, N.z * invLength, -N.y * invLength); }
    V = cross(N, U);
}
static float CosHemispherePdfW(const float3& aNormal, const float3& aDirection) { return maxf(0.f, dot(aNormal, aDirection)) * M_1_PIf_; }
static float powerCosHemispherePdfW(const float3& aNormal, const float3& aDirection, const float aPower)
{
    const float cosTheta = fmaxf(0.f, dot(aNormal, aDirection));
    return (aPower + 1.f) * powf(cosTheta, aPower) * (M_1_PIf_ * 0.5f);
}
static float3 samplePowerCosHemisphereW(const float2& aSamples, const float aPower)
{
    const float phi = 2.f * M_PIf_ * aSamples.x;
    const float z = powf(aSamples.y, 1.f / (aPower + 1.f));
    const float r = sqrtf(1.f - z * z);
    return make_float3(cosf(phi) * r, sinf(phi) * r, z);
}
static float3 sampleCone(const float2& aSample, const float aThetaRad, const float3& aNormal, float* oPdfW)
{
    float3 U, V;
    createCoordinateSystem(aNormal, U, V);
    float cosTheta = cosf(aThetaRad);
    float zSampleRange = 1.f - cosTheta;
    float z = cosTheta + zSampleRange * aSample.x;
    float phi = 2 * M_PIf_ * aSample.y;
    float r = sqrtf(1.f - z * z);
    float x = r * cosf(phi);
    float y = r * sinf(phi);
    if (oPdfW) { float coneSolidAngle = 2.f * M_PIf_ * (1.f - cosTheta); *oPdfW = 1.f / coneSolidAngle; }
    return normalize(x * U + z * aNormal + y * V);
}
static float sampleConePdfW(const float aTheta) { float coneSolidAngle = 2.f * M_PIf_ * (1.f - cosf(aTheta)); return 1.f / coneSolidAngle; }
static float pdfWtoA(const float aPdfW, const float aDist, const float aCosThere) { return aPdfW * fabsf(aCosThere) / sqr(aDist); }

// ---- reflection.h:132-173 -------------------------------------------------------------------------------
static float fresnelDielectric(float cosi, float eta_i, float eta_t)
{
    cosi = cosi < -1.0f ? -1.0f : (cosi > 1.0f ? 1.0f : cosi);
    bool entering = cosi > 0.0f;
    float ei = eta_i, et = eta_t;
    if (!entering) std::swap(ei, et);
    float sint = ei / et * sqrtf(fmaxf(0.0f, 1.0f - cosi * cosi));
    if (sint >= 1.0f) return 1.0f;
    cosi = fabsf(cosi);
    float cost = sqrtf(fmaxf(0.0f, 1.0f - sint * sint));
    float Rparl = ((et * cosi) - (ei * cost)) / ((et * cosi) + (ei * cost));
    float Rperp = ((ei * cosi) - (et * cost)) / ((ei * cosi) + (et * cost));
    return (Rparl * Rparl + Rperp * Rperp) * 0.5f;
}

// ---- BxDF.h:60-97 flags ---------------------------------------------------------------------------------
enum BxDFType {
    B_Null = 0, B_Reflection = 1 << 0, B_Transmission = 1 << 1, B_Diffuse = 1 << 2, B_Glossy = 1 << 3, B_Specular = 1 << 4,
    B_AllType = B_Diffuse | B_Glossy | B_Specular, B_AllReflection = B_Reflection | B_AllType, B_AllTransmission = B_Transmission | B_AllType,
    B_All = B_AllReflection | B_AllTransmission,
    B_Lambertian = 1 << 5, B_SpecularReflection = 1 << 6, B_SpecularTransmission = 1 << 7, B_Phong = 1 << 8
};
static inline bool matchFlags(int flagsToCheck, int flagsToCheckFor) { return (flagsToCheck & B_All & flagsToCheckFor) == (flagsToCheck & B_All); }

// One BxDF of {Lambertian, Phong, SpecularReflection(NoOp|Dielectric), SpecularTransmission} (BxDF.h:171-577)
struct BxDF {
    int type;
    float3 color;        // _reflectance / _transmittance
    float exponent;      // Phong
    int fresnelDielectricFlag;   // SpecularReflection: 0 = FresnelNoOp, 1 = FresnelDielectric
    float eta_i, eta_t;

    float fresnelEval(float cosi) const { return fresnelDielectricFlag ? fresnelDielectric(cosi, eta_i, eta_t) : 1.0f; }

    float continuationProb(const float3& wo) const
    {
        if (type & B_SpecularReflection) return fresnelEval(wo.z) * fmaxf(color.x, maxf(color.y, color.z));
        if (type & B_SpecularTransmission) return 1.f - fresnelDielectric(wo.z, eta_i, eta_t);
        return fmaxf(color.x, maxf(color.y, color.z));
    }
    float albedo(const float3& wo) const
    {
        if (type & B_SpecularReflection) return fresnelEval(wo.z) * luminanceCIE(color);
        if (type & B_SpecularTransmission) return (1.f - fresnelDielectric(wo.z, eta_i, eta_t)) * luminanceCIE(color);
        return luminanceCIE(color);
    }
    float pdf(const float3& wo, const float3& wi, bool evalReverse) const
    {
        if (type & B_Lambertian) {
            if (wo.z * wi.z > 0.0f) return (!evalReverse ? fabsf(wi.z) : fabsf(wo.z)) * M_1_PIf_;
            return 0.f;
        }
        if (type & B_Phong) {
            const float3 r = make_float3(-wo.x, -wo.y, wo.z);
            const float d = dot(r, wi);
            if (d <= EPS_PHONG_) return 0.f;
            return powerCosHemispherePdfW(r, wi, exponent);
        }
        return 0.f;
    }
    // f(wo, wi, oPdfW): the reference dereferences a NULL oPdfW (BxDF.h:187); treated as `if (oPdfW)`
    float3 f(const float3& wo, const float3& wi, float* oPdfW) const
    {
        if (type & B_Lambertian) { if (oPdfW) *oPdfW = pdf(wo, wi, false); return color * M_1_PIf_; }
        if (type & B_Phong) {
            if (wo.z < EPS_COSINE_ || wi.z < EPS_COSINE_) { if (oPdfW) *oPdfW = 0.f; return make_float3(0.f); }
            const float3 r = make_float3(-wo.x, -wo.y, wo.z);
            const float d = dot(r, wi);
            if (d <= EPS_PHONG_) { if (oPdfW) *oPdfW = 0.f; return make_float3(0.f); }
            if (oPdfW) *oPdfW = powerCosHemispherePdfW(r, wi, exponent);
            float3 rho = color * (exponent + 2.f) * 0.5f * M_1_PIf_;
            return rho * powf(d, exponent);
        }
        return make_float3(0.f);
    }
    float3 sampleF(const float3& wo, float3* oWi, const float2& s, float* oPdfW, bool computeAdjoint) const
    {
        if (type & B_Lambertian) {
            if (wo.z < EPS_COSINE_) { *oPdfW = 0.f; return make_float3(0.f); }
            cosine_sample_hemisphere(s.x, s.y, *oWi);
            *oPdfW = pdf(wo, *oWi, false);
            return f(wo, *oWi, nullptr);
        }
        if (type & B_Phong) {
            *oWi = samplePowerCosHemisphereW(s, exponent);
            if (wo.z < EPS_COSINE_ || oWi->z < EPS_COSINE_) { if (oPdfW) *oPdfW = 0.f; return make_float3(0.f); }
            const float3 r = make_float3(-wo.x, -wo.y, wo.z);
            {   // DifferentialGeometry::SetFromNormal + ToWorld (DifferentialGeometry.h:50-64)
                float3 normal = normalize(r);
                float3 tmpBiTan = (fabsf(normal.x) > 0.99f) ? make_float3(0.f, 1.f, 0.f) : make_float3(1.f, 0.f, 0.f);
                float3 tangent = normalize(cross(normal, tmpBiTan));
                float3 bitangent = cross(tangent, normal);
                *oWi = bitangent * oWi->x + tangent * oWi->y + normal * oWi->z;
            }
            const float d = dot(r, *oWi);
            if (d <= EPS_PHONG_) { if (oPdfW) *oPdfW = 0.f; return make_float3(0.f); }
            if (oPdfW) *oPdfW = pdf(wo, *oWi, false);
            float3 rho = color * (exponent + 2.f) * 0.5f * M_1_PIf_;
            return rho * powf(d, exponent);
        }
        if (type & B_SpecularReflection) {
            *oWi = make_float3(-wo.x, -wo.y, wo.z);
            *oPdfW = 1.0f;
            float R = fresnelEval(wo.z);
            return R * color / fabsf(oWi->z);
        }
        {   // SpecularTransmission (BxDF.h:526-564)
            bool entering = wo.z > 0.0f;
            float ei = eta_i, et = eta_t;
            if (!entering) std::swap(ei, et);
            float sini2 = 1.0f - wo.z * wo.z;
            float eta = ei / et;
            float sint2 = eta * eta * sini2;
            if (sint2 >= 1.f) return make_float3(0.f);
            float cost = sqrtf(fmaxf(0.f, 1.f - sint2));
            if (entering) cost = -cost;
            float sintOverSini = eta;
            *oWi = make_float3(sintOverSini * -wo.x, sintOverSini * -wo.y, cost);
            *oPdfW = 1.f;
            float R = fresnelDielectric(wo.z, eta_i, eta_t);
            float T = 1.f - R;
            if (computeAdjoint) return T * color / fabsf(oWi->z);
            return T * color * sqr(sintOverSini) / fabsf(oWi->z);
        }
    }
    float3 vcmF(const float3& wo, const float3& wi, float* oDirectPdfW, float* oReversePdfW) const
    {
        if (type & B_Lambertian) {
            if (wo.z < EPS_COSINE_ || wi.z < EPS_COSINE_) return make_float3(0.f);
            if (oDirectPdfW) *oDirectPdfW = fmaxf(0.f, wi.z * M_1_PIf_);
            if (oReversePdfW) *oReversePdfW = fmaxf(0.f, wo.z * M_1_PIf_);
            return color * M_1_PIf_;
        }
        if (type & B_Phong) {
            float p = 0.f;
            float3 r = f(wo, wi, &p);
            if (oDirectPdfW) *oDirectPdfW = p;
            if (oReversePdfW) *oReversePdfW = p;
            return r;
        }
        if (oDirectPdfW) *oDirectPdfW = 0.f;
        if (oReversePdfW) *oReversePdfW = 0.f;
        return make_float3(0.f);
    }
};

// VcmBSDF (BSDF.h:305-641) over the frame of DifferentialGeometry::SetFromNormal
struct VcmBSDF {
    float3 bitangent, tangent, normal, geometricNormal;
    unsigned nBxDFs;
    BxDF bxdf[2];
    float pickProb[2];
    float contProb;
    bool dirFixIsLight;
    float3 localDirFix;

    float3 ToWorld(const float3& a) const { return bitangent * a.x + tangent * a.y + normal * a.z; }
    float3 ToLocal(const float3& a) const { return make_float3(dot(bitangent, a), dot(tangent, a), dot(normal, a)); }

    VcmBSDF() {}
    VcmBSDF(const float3& aWorldGeometricNormal, const float3& aIncidentDir, bool aIncidentDirIsLight)
    {
        normal = normalize(aWorldGeometricNormal);
        float3 tmpBiTan = (fabsf(normal.x) > 0.99f) ? make_float3(0.f, 1.f, 0.f) : make_float3(1.f, 0.f, 0.f);
        tangent = normalize(cross(normal, tmpBiTan));
        bitangent = cross(tangent, normal);
        geometricNormal = aWorldGeometricNormal;
        nBxDFs = 0;
        dirFixIsLight = aIncidentDirIsLight;
        localDirFix = ToLocal(aIncidentDir);
        pickProb[0] = pickProb[1] = 0.f;
        contProb = 0.f;
    }
    void AddBxDF(const BxDF& b)
    {   // BSDF.h:347-370
        if (nBxDFs == 2) return;
        bxdf[nBxDFs] = b;
        float rr = 0.f; rr += b.continuationProb(localDirFix);
        contProb = fminf(1.f, contProb + rr);
        float al = 0.f; al += b.albedo(localDirFix);
        pickProb[nBxDFs] = al;
        nBxDFs++;
    }
    unsigned countMatching(int type) const { unsigned c = 0; for (unsigned i = 0; i < nBxDFs; i++) if (matchFlags(bxdf[i].type, type)) ++c; return c; }
    bool isSpecular() const { return countMatching(B_All & ~B_Specular) == 0; }
    float sumPickProb(int type) const { float s = 0.f; for (unsigned i = 0; i < nBxDFs; i++) if (matchFlags(bxdf[i].type, type)) s += pickProb[i]; return s; }

    float pdf(const float3& worldDirGen, int sampleType, bool evalRev) const
    {   // BSDF.h:414-436
        float3 wi = ToLocal(worldDirGen);
        float sum = sumPickProb(sampleType);
        if (sum == 0.f) return 0.f;
        float p = 0.f;
        for (unsigned i = 0; i < nBxDFs; i++)
            if (matchFlags(bxdf[i].type, sampleType)) {
                float comp = 0.f; comp += bxdf[i].pdf(localDirFix, wi, evalRev);
                p += comp * pickProb[i] / sum;
            }
        return p;
    }

    // vcmSampleF -> sampleF (BSDF.h:440-571)
    float3 vcmSampleF(float3* oWorldDirGen, const float3& aSample, float* oPdfW, float* oCosThetaOut, int aSampleType, int* oSampledType) const
    {
        const float3 aWorldWo = ToWorld(localDirFix);
        unsigned index = 0;
        float matchedSum = 0.f;
        unsigned nMatched = countMatching(aSampleType);
        if (nMatched != 0) {   // sampleBxDF :373-396
            matchedSum = sumPickProb(aSampleType);
            float prev = 0.f;
            for (unsigned i = 0; i < nBxDFs; ++i)
                if (matchFlags(bxdf[i].type, aSampleType)) {
                    float c = pickProb[i] / matchedSum;
                    if (aSample.x < prev + c) { index = i; break; }
                    prev += c;
                }
        }
        const BxDF& b = bxdf[index];
        if (matchedSum == 0.f) { *oPdfW = 0.0f; return make_float3(0.0f); }
        float3 wo = ToLocal(aWorldWo);
        float3 wi = make_float3(0.f);
        float2 s = make_float2(aSample.y, aSample.z);
        float3 f = b.sampleF(wo, &wi, s, oPdfW, dirFixIsLight);
        *oPdfW *= pickProb[index] / matchedSum;
        if (*oPdfW == 0.0f) { if (oSampledType) *oSampledType = B_Null; return make_float3(0.0f); }
        if (oSampledType) *oSampledType = b.type;
        *oWorldDirGen = ToWorld(wi);
        if (oCosThetaOut) *oCosThetaOut = fabsf(wi.z);
        if (!(b.type & B_Specular) && nMatched > 1) {
            for (unsigned i = 0; i < nBxDFs; i++) {
                if (i == index) continue;
                if (matchFlags(bxdf[i].type, aSampleType)) {
                    float comp = 0.f; comp += bxdf[i].pdf(wo, wi, false);
                    *oPdfW += comp * (pickProb[index] / matchedSum);   // sic: the sampled component's pick probability (:541)
                }
            }
        }
        if (!(b.type & B_Specular)) {
            if (dot(geometricNormal, *oWorldDirGen) * dot(geometricNormal, aWorldWo) >= 0.0f) aSampleType = aSampleType & ~B_Transmission;
            else aSampleType = aSampleType & ~B_Reflection;
            for (unsigned i = 0; i < nBxDFs; ++i) {
                if (i == index) continue;
                if (matchFlags(bxdf[i].type, aSampleType)) f += bxdf[i].f(wo, wi, nullptr);
            }
        }
        return f;
    }

    // vcmF (BSDF.h:577-639)
    float3 vcmF(const float3& aWorldDirGen, float& oCosThetaGen, float* oDirectPdfW, float* oReversePdfW) const
    {
        int aSampleType = B_All;
        float3 localDirGen = ToLocal(aWorldDirGen);
        float3 worldDirFix = ToWorld(localDirFix);
        if (oDirectPdfW) *oDirectPdfW = 0.f;
        if (oReversePdfW) *oReversePdfW = 0.f;
        if (localDirFix.z < EPS_COSINE_ || localDirGen.z < EPS_COSINE_) return make_float3(0.f);
        oCosThetaGen = localDirGen.z;
        if (dot(geometricNormal, aWorldDirGen) * dot(geometricNormal, worldDirFix) >= 0.0f) aSampleType = aSampleType & ~B_Transmission;
        else aSampleType = aSampleType & ~B_Reflection;
        float sum = sumPickProb(aSampleType);
        if (sum == 0.f) return make_float3(0.0f);
        float3 f = make_float3(0.0f);
        for (unsigned i = 0; i < nBxDFs; ++i)
            if (matchFlags(bxdf[i].type, aSampleType)) {
                float d = 0.f, r = 0.f;
                f += bxdf[i].vcmF(localDirFix, localDirGen, &d, &r);
                d *= pickProb[i] / sum; r *= pickProb[i] / sum;
                if (oDirectPdfW) *oDirectPdfW += d;
                if (oReversePdfW) *oReversePdfW += r;
            }
        return f;
    }
};

static BxDF makeLambertian(const float3& kd) { BxDF b{}; b.type = B_Lambertian | B_Reflection | B_Diffuse; b.color = kd; return b; }
static BxDF makePhong(const float3& ks, float e) { BxDF b{}; b.type = B_Phong | B_Reflection | B_Glossy; b.color = ks; b.exponent = e; return b; }
static BxDF makeSpecRefl(const float3& kr, bool dielectric, float ei, float et)
{ BxDF b{}; b.type = B_SpecularReflection | B_Reflection | B_Specular; b.color = kr; b.fresnelDielectricFlag = dielectric; b.eta_i = ei; b.eta_t = et; return b; }
static BxDF makeSpecTrans(const float3& kt, float ei, float et)
{ BxDF b{}; b.type = B_SpecularTransmission | B_Transmission | B_Specular; b.color = kt; b.eta_i = ei; b.eta_t = et; return b; }

struct LightVertex {   // vcm/LightVertex.h:14-27
    float3 hitPoint, throughput; VcmBSDF bsdf; uint2 launchIndex; uint pathLen; float dVCM, dVC, dVM;
};

struct SubpathPRD {    // vcm/SubpathPRD.h:15-34
    float3 origin, direction, throughput, color; uint2 launchIndex; uint launchIndex1D; uint depth; RandomState randomState;
    float dVCM, dVC, dVM; bool done, isGenByFiniteLight, isSpecularPath;
};

void vcm_alloc_vertices(Renderer& r, uint n)
{
    if (!r.lightVertexBuffer) r.lightVertexBuffer = new std::vector<LightVertex>();
    r.lightVertexBuffer->assign(std::max<uint>(n, 1u), LightVertex{});
    r.lightVertexBufferSize = n;
}
void vcm_free(Renderer& r) { delete r.lightVertexBuffer; r.lightVertexBuffer = nullptr; }

// parity diagnostics: (owning path, hit point) of every light vertex stored by the last light pass, 16 bytes each
unsigned vcm_light_vertex_positions(Renderer& r, float* dst4, unsigned capacity)
{
    const uint n = std::min<uint>(r.lightVertexBufferIndex.load(), r.lightVertexBufferSize);
    for (uint i = 0; i < n && i < capacity; i++) {
        const LightVertex& v = (*r.lightVertexBuffer)[i];
        const uint path = v.launchIndex.y * r.width + v.launchIndex.x;
        memcpy(dst4 + 4 * (size_t)i, &path, 4);
        dst4[4 * (size_t)i + 1] = v.hitPoint.x; dst4[4 * (size_t)i + 2] = v.hitPoint.y; dst4[4 * (size_t)i + 3] = v.hitPoint.z;
    }
    return n;
}

// ---- vcm.h:43-59 -----------------------------------------------------------------------------------------
static int isOccluded(Renderer& r, const float3& aPoint, const float3& aDirection, const float aTMax)
{
    if (aTMax < 3.f * EPS_RAY_) return false;
    Ray occlusionRay{aPoint, aDirection, 7, EPS_RAY_, aTMax - 2.f * EPS_RAY_};
    r.stats.any.fetch_add(1, std::memory_order_relaxed);
    return r.scene.anyHit(occlusionRay) ? 1 : 0;
}

static inline void atomicAddFloat(float* p, float v)
{   // the reference's splat is a plain, racy `+=` (vcm.h:159); the oracle makes it atomic so that runs are reproducible
    unsigned* u = reinterpret_cast<unsigned*>(p);
    unsigned old = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        float f; memcpy(&f, &old, 4);
        f += v;
        unsigned nw; memcpy(&nw, &f, 4);
        if (__atomic_compare_exchange_n(u, &old, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
    }
}

// ---- vcm.h:65-161 ----------------------------------------------------------------------------------------
static void connectCameraT1(Renderer& r, SubpathPRD& aLightPrd, const VcmBSDF& aLightBsdf, const float3& aLightHitpoint)
{
    const Camera& aCamera = r.camera;
    const float3 eye = f3(aCamera.eye), lookdir = f3(aCamera.lookdir);
    float3 dirToCamera = eye - aLightHitpoint;
    if (dot(lookdir, -dirToCamera) <= 0.f) return;
    const float distance = length(dirToCamera);
    dirToCamera /= distance;
    const float cosAtCamera = dot(normalize(lookdir), -dirToCamera);
    const float imagePointToCameraDist = length(lookdir) / cosAtCamera;
    const float3 imagePlanePointWorld = eye + (-dirToCamera * imagePointToCameraDist);
    const float3 imagePlaneCenterWorld = eye + lookdir;
    const float3 imagePlaneCenterToPointWorld = imagePlanePointWorld - imagePlaneCenterWorld;
    const float3 unitCameraU = normalize(f3(aCamera.camera_u));
    const float3 unitCameraV = normalize(f3(aCamera.camera_v));
    const float proj_u_len = dot(imagePlaneCenterToPointWorld, unitCameraU);
    const float proj_v_len = dot(imagePlaneCenterToPointWorld, unitCameraV);
    float2 posOnPlane = make_float2(proj_u_len, proj_v_len);
    const float2 ips = make_float2(aCamera.imagePlaneSize[0], aCamera.imagePlaneSize[1]);
    int isOnImage = 2.f * fabsf(posOnPlane.x) < ips.x && 2.f * fabsf(proj_v_len) < ips.y;
    if (!isOnImage) return;
    float2 pixelCoord = (posOnPlane + 0.5f * ips) / ips;
    uint2 pixelIndex = {f2u_sat(pixelCoord.x * r.width), f2u_sat(pixelCoord.y * r.height)};
    pixelIndex.x = std::min(std::max(pixelIndex.x, 0u), r.width - 1);
    pixelIndex.y = std::min(std::max(pixelIndex.y, 0u), r.height - 1);

    float cosToCamera = 0.f, bsdfDirPdfW, bsdfRevPdfW;
    const float3 bsdfFactor = aLightBsdf.vcmF(dirToCamera, cosToCamera, &bsdfDirPdfW, &bsdfRevPdfW);
    if (isZero(bsdfFactor)) return;
    bsdfRevPdfW *= aLightBsdf.contProb;
    const float imageToSolidAngleFactor = sqr(imagePointToCameraDist) / cosAtCamera;
    // `aPixelSizeFactor.x` is used for both axes in the reference (vcm.h:126)
    const float pixelArea = r.pixelSizeFactor.x * ips.x * r.pixelSizeFactor.x * ips.y;
    float imageSamplePdfA = 1.f / pixelArea;
    const float cameraPdfW = imageSamplePdfA * imageToSolidAngleFactor;
    const float cameraPdfA = cameraPdfW * fabsf(cosToCamera) / sqr(distance);
    const float lightHitpointRevPdfA = cameraPdfA;
    const float wLight = vcmMis(lightHitpointRevPdfA / r.lightSubpathCount) * (r.misVmWeightFactor + aLightPrd.dVCM + aLightPrd.dVC * vcmMis(bsdfRevPdfW));
    const float misWeight = 1.f / (wLight + 1.f);
    const float cameraPdfAConvertedToImagePdfA = 1.f / cameraPdfA;
    float3 contrib = misWeight * aLightPrd.throughput * bsdfFactor / (r.lightSubpathCount * cameraPdfAConvertedToImagePdfA);
    if (!isOccluded(r, aLightHitpoint, dirToCamera, distance)) {
        __atomic_fetch_add(&r.connectionCount[(size_t)pixelIndex.y * r.width + pixelIndex.x], 1u, __ATOMIC_RELAXED);
        float* o = &r.output[(size_t)pixelIndex.y * r.width + pixelIndex.x].x;
        atomicAddFloat(o, contrib.x); atomicAddFloat(o + 1, contrib.y); atomicAddFloat(o + 2, contrib.z);
    }
}

// ---- mis.h:109-178 ---------------------------------------------------------------------------------------
static void updateMisTermsOnHit(SubpathPRD& p, const float aCosThetaIn, const float aRayLen)
{
    p.dVCM *= vcmMis(sqr(aRayLen));
    p.dVCM /= vcmMis(aCosThetaIn);
    p.dVC /= vcmMis(aCosThetaIn);
    p.dVM /= vcmMis(aCosThetaIn);
}
static void updateMisTermsOnScatter(SubpathPRD& p, const float aCosThetaOut, const float aBsdfDirPdfW, const float aBsdfRevPdfW,
                                    const float aMisVcWeightFactor, const float aMisVmWeightFactor, int aSampledEvent)
{
    float vertPickPdf = 1.f;
    if (aSampledEvent & B_Specular) {
        p.dVCM = 0.f;
        p.dVC *= vcmMis(aCosThetaOut);
        p.dVM *= vcmMis(aCosThetaOut);
        return;
    }
    p.dVC = vcmMis(aCosThetaOut / aBsdfDirPdfW) * (p.dVC * vcmMis(aBsdfRevPdfW) + p.dVCM + aMisVmWeightFactor);
    p.dVM = vcmMis(aCosThetaOut / aBsdfDirPdfW) * (p.dVM * vcmMis(aBsdfRevPdfW) + p.dVCM * aMisVcWeightFactor * vertPickPdf + 1.f);
    p.dVCM = vcmMis(1.f / aBsdfDirPdfW);
}

// ---- vcm.h:166-204 ---------------------------------------------------------------------------------------
static void sampleScattering(Renderer& r, SubpathPRD& p, const float3& aHitPoint, const VcmBSDF& aBsdf)
{
    float contProb = aBsdf.contProb;
    float rrSample = getRandomUniformFloat(&p.randomState);
    if (contProb < rrSample) { p.done = true; return; }
    float bsdfDirPdfW = 0.f, cosThetaOut = 0.f;
    float3 bsdfSample = getRandomUniformFloat3(&p.randomState);
    int sampledEvent = B_Null;
    float3 bsdfFactor = aBsdf.vcmSampleF(&p.direction, bsdfSample, &bsdfDirPdfW, &cosThetaOut, B_All, &sampledEvent);
    if (isZero(bsdfFactor)) return;   // sic: no `done` (vcm.h:188-189) — the same origin is traced again
    float bsdfRevPdfW = bsdfDirPdfW;
    bool isSpecularEvent = matchFlags(sampledEvent, B_Specular);
    if (!isSpecularEvent) bsdfRevPdfW = aBsdf.pdf(p.direction, B_All & ~B_Specular, true);
    p.isSpecularPath = (p.isSpecularPath && isSpecularEvent);
    bsdfDirPdfW *= contProb;
    bsdfRevPdfW *= contProb;
    updateMisTermsOnScatter(p, cosThetaOut, bsdfDirPdfW, bsdfRevPdfW, r.misVcWeightFactor, r.misVmWeightFactor, sampledEvent);
    p.throughput *= bsdfFactor * (cosThetaOut / bsdfDirPdfW);
    p.origin = aHitPoint;
}

// ---- vcm.h:210-309 ---------------------------------------------------------------------------------------
static void lightHit(Renderer& r, SubpathPRD& aLightPrd, const float3& aHitPoint, const float3& aWorldNormal, const VcmBSDF& aLightBsdf,
                     const float3 aRayWorldDir, const float aRayTHit)
{
    aLightPrd.depth++;
    float cosThetaIn = dot(aWorldNormal, -aRayWorldDir);
    if (cosThetaIn < EPS_COSINE_) { aLightPrd.done = true; return; }
    updateMisTermsOnHit(aLightPrd, cosThetaIn, aRayTHit);
    bool isBsdfSpecular = aLightBsdf.isSpecular();
    if (!isBsdfSpecular) {
        uint currPathVertIdx = r.lightSubpathVertexCount[aLightPrd.launchIndex1D]++;
        if (!r.lightVertexCountEstimatePass) {
            LightVertex lv;
            lv.launchIndex = aLightPrd.launchIndex; lv.hitPoint = aHitPoint; lv.throughput = aLightPrd.throughput; lv.pathLen = aLightPrd.depth;
            lv.dVCM = aLightPrd.dVCM; lv.dVC = aLightPrd.dVC; lv.dVM = aLightPrd.dVM; lv.bsdf = aLightBsdf;
            uint vertIdx = r.lightVertexBufferIndex.fetch_add(1u);
            // the reference has no overflow guard (vcm.h:281-282); out-of-range vertices are dropped here and connect to nothing
            if (vertIdx < r.lightVertexBufferSize) (*r.lightVertexBuffer)[vertIdx] = lv;
            r.lightSubpathVertexIndex[(size_t)aLightPrd.launchIndex1D * (r.vcmMaxPathLength - 1) + currPathVertIdx] = vertIdx;
        }
    }
    if (!r.lightVertexCountEstimatePass && !isBsdfSpecular) connectCameraT1(r, aLightPrd, aLightBsdf, aHitPoint);
    if (r.maxPathLen < aLightPrd.depth + 2) { aLightPrd.done = true; return; }
    sampleScattering(r, aLightPrd, aHitPoint, aLightBsdf);
}

// ---- vcm.h:315-400 ---------------------------------------------------------------------------------------
static void connectVertices(Renderer& r, const LightVertex& aLightVertex, SubpathPRD& aCameraPrd, const VcmBSDF& aCameraBsdf, const float3& aCameraHitpoint)
{
    float3 direction = aLightVertex.hitPoint - aCameraHitpoint;
    float dist2 = dot(direction, direction);
    float distance = sqrtf(dist2);
    direction /= distance;
    float cameraCosTheta = 0.f, cameraBsdfDirPdfW, cameraBsdfRevPdfW;
    const float3 cameraBsdfFactor = aCameraBsdf.vcmF(direction, cameraCosTheta, &cameraBsdfDirPdfW, &cameraBsdfRevPdfW);
    if (isZero(cameraBsdfFactor)) return;
    const float cameraCont = aCameraBsdf.contProb;
    cameraBsdfDirPdfW *= cameraCont; cameraBsdfRevPdfW *= cameraCont;
    float lightCosTheta = 0.f, lightBsdfDirPdfW, lightBsdfRevPdfW;
    const float3 lightBsdfFactor = aLightVertex.bsdf.vcmF(-direction, lightCosTheta, &lightBsdfDirPdfW, &lightBsdfRevPdfW);
    if (isZero(lightBsdfFactor)) return;
    const float lightCont = aLightVertex.bsdf.contProb;
    lightBsdfDirPdfW *= lightCont; lightBsdfRevPdfW *= lightCont;
    const float geometryTerm = lightCosTheta * cameraCosTheta / dist2;
    if (geometryTerm < 0.f) return;
    const float cameraBsdfDirPdfA = pdfWtoA(cameraBsdfDirPdfW, distance, cameraCosTheta);
    const float lightBsdfDirPdfA = pdfWtoA(lightBsdfDirPdfW, distance, lightCosTheta);
    const float invVertPickPdf = 1.f;
    const float wLight = vcmMis(cameraBsdfDirPdfA) * (r.misVmWeightFactor * invVertPickPdf + aLightVertex.dVCM + aLightVertex.dVC * vcmMis(lightBsdfRevPdfW));
    const float wCamera = vcmMis(lightBsdfDirPdfA) * (r.misVmWeightFactor * invVertPickPdf + aCameraPrd.dVCM + aCameraPrd.dVC * vcmMis(cameraBsdfRevPdfW));
    const float misWeight = 1.f / (wLight + 1.f + wCamera);
    float3 contrib = geometryTerm * cameraBsdfFactor * lightBsdfFactor * invVertPickPdf;
    contrib *= misWeight * aCameraPrd.throughput * aLightVertex.throughput;
    if (isOccluded(r, aCameraHitpoint, direction, distance)) return;
    __atomic_fetch_add(&r.connectionCount[aCameraPrd.launchIndex1D], 1u, __ATOMIC_RELAXED);
    aCameraPrd.color += contrib;
}

// ---- helpers/light.h:92-213 ------------------------------------------------------------------------------
static float3 lightEmit(const Sphere& bs, const Light& aLight, RandomState& rs, float3& oPosition, float3& oDirection, float& oEmissionPdfW,
                        float& oDirectPdfA, float& oCosThetaLight)
{
    float3 radiance = make_float3(0.f);
    float2 dirRnd = getRandomUniformFloat2(&rs);
    if (aLight.lightType == OPR_LIGHT_AREA) {
        float2 posRnd = getRandomUniformFloat2(&rs);
        oPosition = f3(aLight.position) + posRnd.x * f3(aLight.v1) + posRnd.y * f3(aLight.v2);
        oDirection = sampleUnitHemisphereCos(f3(aLight.normal), dirRnd, &oEmissionPdfW, &oCosThetaLight, true);
        oEmissionPdfW *= aLight.inverseArea;
        oDirectPdfA = aLight.inverseArea;
        radiance = f3(aLight.Lemit) * oCosThetaLight;
    } else if (aLight.lightType == OPR_LIGHT_POINT) {
        oPosition = f3(aLight.position);
        float3 lightToSceneCenter = bs.center - f3(aLight.position);
        float lightSceneCenterDistance = length(lightToSceneCenter);
        lightToSceneCenter /= lightSceneCenterDistance;
        bool lightOutsideSphere = bs.radius < lightSceneCenterDistance;
        if (lightOutsideSphere) {
            float theta = asinf(bs.radius / lightSceneCenterDistance);
            oDirection = sampleCone(dirRnd, theta, lightToSceneCenter, &oEmissionPdfW);
        } else {
            oDirection = sampleUnitSphere(dirRnd, &oEmissionPdfW);
        }
        oDirectPdfA = 1.f;
        oCosThetaLight = 1.f;
        radiance = f3(aLight.Lemit);   // intensity
    }
    return radiance;
}

static float3 lightIlluminate(const Sphere& bs, const Light& aLight, RandomState& rs, const float3& aReceivePosition, float3& oDirectionToLight,
                              float& oDistance, float& oDirectPdfW, float* oEmissionPdfW, float* oCosThetaLight)
{
    float3 radiance = make_float3(0.f);
    if (aLight.lightType == OPR_LIGHT_AREA) {
        float2 posRnd = getRandomUniformFloat2(&rs);
        float3 pointOnLight = f3(aLight.position) + posRnd.x * f3(aLight.v1) + posRnd.y * f3(aLight.v2);
        oDirectionToLight = pointOnLight - aReceivePosition;
        oDistance = length(oDirectionToLight);
        oDirectionToLight /= oDistance;
        float cosThetaLight = dot(f3(aLight.normal), -oDirectionToLight);
        if (cosThetaLight < EPS_COSINE_) return radiance;
        oDirectPdfW = aLight.inverseArea * sqr(oDistance) / cosThetaLight;
        if (oCosThetaLight) *oCosThetaLight = cosThetaLight;
        if (oEmissionPdfW) *oEmissionPdfW = aLight.inverseArea * cosThetaLight * M_1_PIf_;
        return f3(aLight.Lemit);
    } else if (aLight.lightType == OPR_LIGHT_POINT) {
        oDirectionToLight = f3(aLight.position) - aReceivePosition;
        oDistance = length(oDirectionToLight);
        oDirectionToLight /= oDistance;
        float3 lightToSceneCenter = bs.center - f3(aLight.position);
        float lightSceneCenterDistance = length(lightToSceneCenter);
        bool lightOutsideSphere = bs.radius < lightSceneCenterDistance;
        if (lightOutsideSphere) {
            float theta = asinf(bs.radius / lightSceneCenterDistance);
            if (oEmissionPdfW) *oEmissionPdfW = sampleConePdfW(theta);
        } else {
            if (oEmissionPdfW) *oEmissionPdfW = 0.25f * M_1_PIf_;
        }
        oDirectPdfW = sqr(oDistance);
        if (oCosThetaLight) *oCosThetaLight = 1.f;
        radiance = f3(aLight.Lemit);
    }
    return radiance;
}

// ---- vcm.h:406-488 ---------------------------------------------------------------------------------------
static void connectLightSourceS1(Renderer& r, SubpathPRD& aCameraPrd, const VcmBSDF& aCameraBsdf, const float3& aCameraHitpoint)
{
    const std::vector<Light>& lights = r.scene.lights;
    int lightIndex = 0;
    if (1 < lights.size()) {
        float sample = getRandomUniformFloat(&aCameraPrd.randomState);
        lightIndex = intmin((int)(sample * lights.size()), int(lights.size() - 1));
    }
    const Light light = lights[lightIndex];
    const float lightPickProb = 1.f / lights.size();
    float emissionPdfW = 0.f, directPdfW = 0.f, cosAtLight = 0.f, distance = 0.f;
    float3 dirToLight = make_float3(0.f);
    float3 radiance = lightIlluminate(r.scene.boundingSphere, light, aCameraPrd.randomState, aCameraHitpoint, dirToLight, distance, directPdfW, &emissionPdfW, &cosAtLight);
    if (isZero(radiance)) return;
    float bsdfDirPdfW, bsdfRevPdfW, cosToLight = 0.f;
    float3 bsdfFactor = aCameraBsdf.vcmF(dirToLight, cosToLight, &bsdfDirPdfW, &bsdfRevPdfW);
    if (isZero(bsdfFactor)) return;
    const float contiueProb = aCameraBsdf.contProb;
    bsdfDirPdfW *= light.isDelta ? 0.f : contiueProb;
    bsdfRevPdfW *= contiueProb;
    const float wLight = vcmMis(bsdfDirPdfW / (lightPickProb * directPdfW));
    const float wCamera = vcmMis(emissionPdfW * cosToLight / (directPdfW * cosAtLight)) * (r.misVmWeightFactor + aCameraPrd.dVCM + aCameraPrd.dVC * vcmMis(bsdfRevPdfW));
    const float misWeight = 1.f / (wLight + 1.f + wCamera);
    float3 contrib = (misWeight * cosToLight / (lightPickProb * directPdfW)) * (radiance * bsdfFactor);
    if (isZero(contrib)) return;
    if (isOccluded(r, aCameraHitpoint, dirToLight, distance)) return;
    __atomic_fetch_add(&r.connectionCount[aCameraPrd.launchIndex1D], 1u, __ATOMIC_RELAXED);
    aCameraPrd.color += contrib * aCameraPrd.throughput;
}

// ---- vcm.h:493-522 ---------------------------------------------------------------------------------------
static void connectLightSourceS0(SubpathPRD& aCameraPrd, const float3& aRadiance, float aDirectPdfA, float aEmissionPdfW, float aLightPickProb)
{
    if (aCameraPrd.depth == 1) { aCameraPrd.color += aCameraPrd.throughput * aRadiance; return; }
    aDirectPdfA *= aLightPickProb;
    aEmissionPdfW *= aLightPickProb;
    const float wCamera = vcmMis(aDirectPdfA) * aCameraPrd.dVCM + vcmMis(aEmissionPdfW * aCameraPrd.dVC);
    const float misWeight = 1.f / (1.f + wCamera);
    aCameraPrd.color += aCameraPrd.throughput * misWeight * aRadiance;
}

// ---- vcm.h:527-628 ---------------------------------------------------------------------------------------
static void cameraHit(Renderer& r, SubpathPRD& aCameraPrd, const float3& aHitPoint, const float3& aWorldNormal, const VcmBSDF& aCameraBsdf,
                      const float3 aRayWorldDir, const float aRayTHit)
{
    aCameraPrd.depth++;
    float cosThetaIn = dot(aWorldNormal, -aRayWorldDir);
    if (cosThetaIn < EPS_COSINE_) { aCameraPrd.done = true; return; }
    updateMisTermsOnHit(aCameraPrd, cosThetaIn, aRayTHit);
    bool isBsdfSpecular = aCameraBsdf.isSpecular();
    if (!isBsdfSpecular) connectLightSourceS1(r, aCameraPrd, aCameraBsdf, aHitPoint);
    if (!isBsdfSpecular) {
        uint pathIndex = aCameraPrd.launchIndex1D % r.lightSubpathCount;
        uint lightSubpathVertexCount = r.lightSubpathVertexCount[pathIndex];
        for (uint i = 0; i < lightSubpathVertexCount; ++i) {
            uint vertIdx = r.lightSubpathVertexIndex[(size_t)pathIndex * (r.vcmMaxPathLength - 1) + i];
            if (vertIdx >= r.lightVertexBufferSize) continue;   // dropped on overflow (see lightHit)
            const LightVertex& lightVertex = (*r.lightVertexBuffer)[vertIdx];
            connectVertices(r, lightVertex, aCameraPrd, aCameraBsdf, aHitPoint);
        }
    }
    if (r.maxPathLen <= aCameraPrd.depth) { aCameraPrd.done = true; return; }
    sampleScattering(r, aCameraPrd, aHitPoint, aCameraBsdf);
}

// ---- material VCM closest-hit programs -------------------------------------------------------------------
// Returns false for a program that returns without touching the payload (Mirror with Kr == 0, Mirror.cu:134-135)
static bool buildBsdf(Renderer& r, const Ray& ray, const Hit& hit, bool isLight, VcmBSDF& bsdf, float3& normalForHit, bool& terminate)
{
    const opr_material& m = r.scene.materials[hit.material];
    terminate = false;
    float3 worldGeometricNormal = normalize(hit.geometricNormal);
    normalForHit = worldGeometricNormal;
    switch (m.kind) {
    case OPR_MAT_DIFFUSE:   // Diffuse.cu:174-243
        bsdf = VcmBSDF(worldGeometricNormal, -ray.direction, isLight);
        bsdf.AddBxDF(makeLambertian(f3(m.Kd)));
        return true;
    case OPR_MAT_GLOSSY:    // Glossy.cu:173-252
        bsdf = VcmBSDF(worldGeometricNormal, -ray.direction, isLight);
        bsdf.AddBxDF(makeLambertian(f3(m.Kd)));
        bsdf.AddBxDF(makePhong(f3(m.Ks), m.exponent));
        return true;
    case OPR_MAT_TEXTURE: { // Texture.cu:210-286: the light program passes the un-normalised object-space geometricNormal (:223)
        float4 value = tex2D(r.scene.textures[m.diffuseTexture], hit.textureCoordinate.x, hit.textureCoordinate.y);
        float3 n = isLight ? hit.geometricNormal : worldGeometricNormal;
        bsdf = VcmBSDF(n, -ray.direction, isLight);
        bsdf.AddBxDF(makeLambertian(make_float3(value.x, value.y, value.z)));
        return true;
    }
    case OPR_MAT_MIRROR:    // Mirror.cu:118-189
        if (isZero(f3(m.Kr))) return false;
        bsdf = VcmBSDF(worldGeometricNormal, -ray.direction, isLight);
        bsdf.AddBxDF(makeSpecRefl(f3(m.Kr), false, 0.f, 0.f));
        return true;
    case OPR_MAT_GLASS: {   // Glass.cu:258-367
        bool isHitFromOutside = dot(worldGeometricNormal, ray.direction) < 0;
        float3 N = isHitFromOutside ? worldGeometricNormal : -worldGeometricNormal;
        float n1 = isHitFromOutside ? 1.f : m.indexOfRefraction, n2 = isHitFromOutside ? m.indexOfRefraction : 1.f;
        bool krBlack = isZero(f3(m.Kr)), ktBlack = isZero(f3(m.Kt));
        if (krBlack && ktBlack) { terminate = true; return false; }
        normalForHit = N;
        bsdf = VcmBSDF(N, -ray.direction, isLight);
        if (!ktBlack) bsdf.AddBxDF(makeSpecTrans(f3(m.Kt), n1, n2));
        if (!krBlack) bsdf.AddBxDF(makeSpecRefl(f3(m.Kr), true, n1, n2));
        return true;
    }
    }
    return false;
}

static void traceLight(Renderer& r, const Ray& ray, SubpathPRD& prd)
{
    Hit hit;
    r.stats.closest.fetch_add(1, std::memory_order_relaxed);
    if (!r.scene.closestHit(ray, hit)) { prd.done = true; return; }   // miss, VCMLightPass.cu:95-98
    const opr_material& m = r.scene.materials[hit.material];
    if (m.kind == OPR_MAT_DIFFUSE_EMITTER) { prd.done = true; return; }   // DiffuseEmitter.cu:76-79
    if (prd.launchIndex1D < r.pathSignature.size()) { uint& sg = r.pathSignature[prd.launchIndex1D]; sg = sg * 131u + (uint)m.kind + 1u; }   // diagnostics only
    VcmBSDF bsdf; float3 n; bool terminate;
    if (!buildBsdf(r, ray, hit, true, bsdf, n, terminate)) { if (terminate) prd.done = true; return; }
    float3 hitPoint = ray.origin + hit.t * ray.direction;
    lightHit(r, prd, hitPoint, n, bsdf, ray.direction, hit.t);
}

static void traceCamera(Renderer& r, const Ray& ray, SubpathPRD& prd)
{
    Hit hit;
    r.stats.closest.fetch_add(1, std::memory_order_relaxed);
    if (!r.scene.closestHit(ray, hit)) { prd.done = true; return; }   // miss, VCMCameraPass.cu:87-90
    const opr_material& m = r.scene.materials[hit.material];
    { const size_t si = (size_t)r.width * r.height + prd.launchIndex1D; if (si < r.pathSignature.size()) { uint& sg = r.pathSignature[si]; sg = sg * 131u + (uint)m.kind + 1u; } }   // diagnostics only
    if (m.kind == OPR_MAT_DIFFUSE_EMITTER) {   // DiffuseEmitter.cu:95-120
        prd.depth++;
        prd.done = true;
        float3 Lemit = make_float3(m.power[0] * m.inverseArea * M_1_PIf_, m.power[1] * m.inverseArea * M_1_PIf_, m.power[2] * m.inverseArea * M_1_PIf_);
        if (isZero(Lemit)) return;
        float3 worldGeometricNormal = normalize(hit.geometricNormal);
        if (dot(worldGeometricNormal, -ray.direction) < 0.f) return;
        float lightPickProb = 1.f / r.scene.lights.size();
        float directPdfA = m.inverseArea;
        float emissionPdfW = CosHemispherePdfW(worldGeometricNormal, -ray.direction) * m.inverseArea;
        connectLightSourceS0(prd, Lemit, directPdfA, emissionPdfW, lightPickProb);
        return;
    }
    VcmBSDF bsdf; float3 n; bool terminate;
    if (!buildBsdf(r, ray, hit, false, bsdf, n, terminate)) { if (terminate) prd.done = true; return; }
    float3 hitPoint = ray.origin + hit.t * ray.direction;
    cameraHit(r, prd, hitPoint, n, bsdf, ray.direction, hit.t);
}

// ---- VCMLightPass.cu:60-162 with mis.h:35-79 -------------------------------------------------------------
void vcm_lightPass(Renderer& r, uint x, uint y)
{
    const uint idx1D = x + y * r.width;
    r.lightSubpathVertexCount[idx1D] = 0u;
    SubpathPRD p;
    p.launchIndex = {x, y}; p.launchIndex1D = idx1D;
    p.throughput = make_float3(1.f); p.color = make_float3(0.f);
    p.depth = 0; p.dVC = p.dVM = p.dVCM = 0.f; p.done = false; p.isSpecularPath = true;
    p.randomState = r.rs(x, y);
    const std::vector<Light>& lights = r.scene.lights;
    int lightIndex = 0;
    if (1 < lights.size()) {
        float sample = getRandomUniformFloat(&p.randomState);
        lightIndex = intmin((int)(sample * lights.size()), int(lights.size() - 1));
    }
    const Light light = lights[lightIndex];
    const float lightPickPdf = 1.f / lights.size();
    float emissionPdfW = 0.f, directPdfW = 0.f, cosAtLight = 0.f;
    p.origin = p.direction = make_float3(0.f);
    p.throughput = lightEmit(r.scene.boundingSphere, light, p.randomState, p.origin, p.direction, emissionPdfW, directPdfW, cosAtLight);
    emissionPdfW *= lightPickPdf;
    directPdfW *= lightPickPdf;
    p.throughput /= emissionPdfW;
    p.isGenByFiniteLight = light.isFinite;
    // initLightMisTerms
    p.dVCM = vcmMis(directPdfW / emissionPdfW);
    if (!light.isDelta) {
        const float usedCosLight = light.isFinite ? cosAtLight : 1.f;
        p.dVC = vcmMis(usedCosLight / emissionPdfW);
    } else p.dVC = 0.f;
    p.dVM = p.dVC * r.misVcWeightFactor;

    Ray lightRay{p.origin, p.direction, 3, RAY_LEN_MIN_, RT_DEFAULT_MAX_};
    for (int i = 0;; i++) {
        traceLight(r, lightRay, p);
        if (p.done) break;
        lightRay.origin = p.origin;
        lightRay.direction = p.direction;
        if (i > 100000) break;   // a Kr == 0 mirror never sets `done` in the reference (Mirror.cu:134-135); bounded here
    }
    r.rs(x, y) = p.randomState;
}

// ---- VCMCameraPass.cu:52-146 with mis.h:83-103 -----------------------------------------------------------
void vcm_cameraPass(Renderer& r, uint x, uint y)
{
    SubpathPRD p;
    p.launchIndex = {x, y}; p.launchIndex1D = x + y * r.width;
    p.randomState = r.rs(x, y);
    p.throughput = make_float3(1.0f); p.color = make_float3(0.0f);
    p.depth = 0; p.dVC = p.dVM = p.dVCM = 0; p.isSpecularPath = true; p.done = false; p.isGenByFiniteLight = false;
    const Camera& camera = r.camera;
    float2 screen = make_float2((float)r.width, (float)r.height);
    float2 sample = getRandomUniformFloat2(&p.randomState);
    float2 d = (make_float2((float)x, (float)y) + sample) / screen * 2.0f - 1.0f;
    p.origin = f3(camera.eye);
    p.direction = normalize(d.x * f3(camera.camera_u) + d.y * f3(camera.camera_v) + f3(camera.lookdir));
    float cosAtCamera = dot(normalize(f3(camera.lookdir)), p.direction);
    float distToImagePlane = length(f3(camera.lookdir));
    float imagePointToCameraDist = distToImagePlane / cosAtCamera;
    float imageToSolidAngleFactor = sqr(imagePointToCameraDist) / cosAtCamera;
    float pixelArea = r.pixelSizeFactor.x * camera.imagePlaneSize[0] * r.pixelSizeFactor.x * camera.imagePlaneSize[1];   // sic (.x twice, :138)
    float areaSamplePdf = 1.f / pixelArea;
    float cameraPdfW = areaSamplePdf * imageToSolidAngleFactor;
    p.dVC = .0f; p.dVM = .0f;
    p.dVCM = vcmMis(r.lightSubpathCount / cameraPdfW);

    Ray cameraRay{p.origin, p.direction, 2, RAY_LEN_MIN_, RT_DEFAULT_MAX_};
    for (int i = 0;; i++) {
        traceCamera(r, cameraRay, p);
        if (p.done) break;
        cameraRay.origin = p.origin;
        cameraRay.direction = p.direction;
        if (i > 100000) break;
    }
    float* o = &r.output[(size_t)y * r.width + x].x;
    atomicAddFloat(o, p.color.x); atomicAddFloat(o + 1, p.color.y); atomicAddFloat(o + 2, p.color.z);
    r.rs(x, y) = p.randomState;
}

} // namespace ora

// parity diagnostics (restatement only; the oracle/_ref build has no such entry)
extern "C" unsigned ora_get_light_vertex_positions(void* h, float* dst4, unsigned capacity)
{ return ora::vcm_light_vertex_positions(*(ora::Renderer*)h, dst4, capacity); }
