// bsdf.cuh — device-side BSDF container for the vertex-connection (BDPT) passes.
//
// Replaces the reference's VcmBSDF / BxDF class hierarchy with macro "fake virtual" dispatch
// (RenderEngine/renderer/BSDF.h:305-641, BxDF.h:60-577, reflection.h:47-175, math/DifferentialGeometry.h:50-72)
// by a flat two-component record whose component set is fixed by the material kind:
//   Diffuse / Texture : Lambertian            Glossy : Lambertian + Phong
//   Mirror            : SpecularReflection(FresnelNoOp)
//   Glass             : SpecularTransmission + SpecularReflection(FresnelDielectric)
// Results follow the reference's arithmetic, including its quirks (pick-probability weighting in sampleF :541).
#pragma once
#include "common.cuh"

namespace opr {

enum BxFlags {
    B_Null = 0, B_Reflection = 1 << 0, B_Transmission = 1 << 1, B_Diffuse = 1 << 2, B_Glossy = 1 << 3, B_Specular = 1 << 4,
    B_AllType = B_Diffuse | B_Glossy | B_Specular, B_AllReflection = B_Reflection | B_AllType, B_AllTransmission = B_Transmission | B_AllType,
    B_All = B_AllReflection | B_AllTransmission,
    B_Lambertian = 1 << 5, B_SpecularReflection = 1 << 6, B_SpecularTransmission = 1 << 7, B_Phong = 1 << 8
};
OPR_DEV bool bx_match(int flagsToCheck, int flagsToCheckFor) { return (flagsToCheck & B_All & flagsToCheckFor) == (flagsToCheck & B_All); }

// reflection.h:132-173
OPR_DEV float fresnel_dielectric(float cosi, float eta_i, float eta_t)
{
    cosi = fminf(fmaxf(cosi, -1.0f), 1.0f);
    bool entering = cosi > 0.0f;
    float ei = eta_i, et = eta_t;
    if (!entering) { float t = ei; ei = et; et = t; }
    float sint = ei / et * sqrtf(fmaxf(0.0f, 1.0f - cosi * cosi));
    if (sint >= 1.0f) return 1.0f;
    cosi = fabsf(cosi);
    float cost = sqrtf(fmaxf(0.0f, 1.0f - sint * sint));
    float Rparl = ((et * cosi) - (ei * cost)) / ((et * cosi) + (ei * cost));
    float Rperp = ((ei * cosi) - (et * cost)) / ((ei * cosi) + (et * cost));
    return (Rparl * Rparl + Rperp * Rperp) * 0.5f;
}

OPR_DEV float powerCosHemispherePdfW(const float3& aNormal, const float3& aDirection, const float aPower)
{   // helpers/samplers.h:103-109
    const float cosTheta = fmaxf(0.f, dot(aNormal, aDirection));
    return (aPower + 1.f) * powf(cosTheta, aPower) * (INV_PI_F * 0.5f);
}

struct BxDF {
    int type;
    float3 color;      // reflectance / transmittance
    float exponent;    // Phong
    int dielectric;    // SpecularReflection: FresnelDielectric (1) or FresnelNoOp (0)
    float eta_i, eta_t;

    OPR_DEV float fresnel(float cosi) const { return dielectric ? fresnel_dielectric(cosi, eta_i, eta_t) : 1.0f; }
    OPR_DEV float continuationProb(const float3& wo) const
    {
        if (type & B_SpecularReflection) return fresnel(wo.z) * fmaxf(color.x, maxf_(color.y, color.z));
        if (type & B_SpecularTransmission) return 1.f - fresnel_dielectric(wo.z, eta_i, eta_t);
        return fmaxf(color.x, maxf_(color.y, color.z));
    }
    OPR_DEV float albedo(const float3& wo) const
    {
        if (type & B_SpecularReflection) return fresnel(wo.z) * luminanceCIE(color);
        if (type & B_SpecularTransmission) return (1.f - fresnel_dielectric(wo.z, eta_i, eta_t)) * luminanceCIE(color);
        return luminanceCIE(color);
    }
    OPR_DEV float pdf(const float3& wo, const float3& wi, bool evalReverse) const
    {
        if (type & B_Lambertian) {
            if (wo.z * wi.z > 0.0f) return (!evalReverse ? fabsf(wi.z) : fabsf(wo.z)) * INV_PI_F;
            return 0.f;
        }
        if (type & B_Phong) {
            const float3 r = f3(-wo.x, -wo.y, wo.z);
            if (dot(r, wi) <= EPS_PHONG) return 0.f;
            return powerCosHemispherePdfW(r, wi, exponent);
        }
        return 0.f;
    }
    OPR_DEV float3 f(const float3& wo, const float3& wi, float* oPdfW) const
    {
        if (type & B_Lambertian) { if (oPdfW) *oPdfW = pdf(wo, wi, false); return color * INV_PI_F; }
        if (type & B_Phong) {
            if (wo.z < EPS_COSINE || wi.z < EPS_COSINE) { if (oPdfW) *oPdfW = 0.f; return f3(0.f); }
            const float3 r = f3(-wo.x, -wo.y, wo.z);
            const float d = dot(r, wi);
            if (d <= EPS_PHONG) { if (oPdfW) *oPdfW = 0.f; return f3(0.f); }
            if (oPdfW) *oPdfW = powerCosHemispherePdfW(r, wi, exponent);
            float3 rho = color * (exponent + 2.f) * 0.5f * INV_PI_F;
            return rho * powf(d, exponent);
        }
        return f3(0.f);
    }
    OPR_DEV float3 sampleF(const float3& wo, float3* oWi, const float2& s, float* oPdfW, bool computeAdjoint) const
    {
        if (type & B_Lambertian) {   // BxDF.h:206-224; cosine_sample_hemisphere is the OptiX SDK function
            if (wo.z < EPS_COSINE) { *oPdfW = 0.f; return f3(0.f); }
            const float r = sqrtf(s.x);
            const float phi = 2.0f * PI_F * s.y;
            float sp, cp; sincosf(phi, &sp, &cp);
            float3 p; p.x = r * cp; p.y = r * sp; p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
            *oWi = p;
            *oPdfW = pdf(wo, p, false);
            return f(wo, p, nullptr);
        }
        if (type & B_Phong) {        // BxDF.h:325-363
            {   // samplePowerCosHemisphereW, helpers/samplers.h:112-127
                const float phi = 2.f * PI_F * s.x;
                const float z = powf(s.y, 1.f / (exponent + 1.f));
                const float r = sqrtf(1.f - z * z);
                float sp, cp; sincosf(phi, &sp, &cp);
                *oWi = f3(cp * r, sp * r, z);
            }
            if (wo.z < EPS_COSINE || oWi->z < EPS_COSINE) { *oPdfW = 0.f; return f3(0.f); }
            const float3 r = f3(-wo.x, -wo.y, wo.z);
            {   // DifferentialGeometry::SetFromNormal + ToWorld
                float3 normal = normalize(r);
                float3 tmpBiTan = (fabsf(normal.x) > 0.99f) ? f3(0.f, 1.f, 0.f) : f3(1.f, 0.f, 0.f);
                float3 tangent = normalize(cross(normal, tmpBiTan));
                float3 bitangent = cross(tangent, normal);
                *oWi = bitangent * oWi->x + tangent * oWi->y + normal * oWi->z;
            }
            const float d = dot(r, *oWi);
            if (d <= EPS_PHONG) { *oPdfW = 0.f; return f3(0.f); }
            *oPdfW = pdf(wo, *oWi, false);
            float3 rho = color * (exponent + 2.f) * 0.5f * INV_PI_F;
            return rho * powf(d, exponent);
        }
        if (type & B_SpecularReflection) {   // BxDF.h:452-466
            *oWi = f3(-wo.x, -wo.y, wo.z);
            *oPdfW = 1.0f;
            float R = fresnel(wo.z);
            return R * color / fabsf(oWi->z);
        }
        {   // SpecularTransmission, BxDF.h:526-564
            bool entering = wo.z > 0.0f;
            float ei = eta_i, et = eta_t;
            if (!entering) { float t = ei; ei = et; et = t; }
            float sini2 = 1.0f - wo.z * wo.z;
            float eta = ei / et;
            float sint2 = eta * eta * sini2;
            if (sint2 >= 1.f) return f3(0.f);
            float cost = sqrtf(fmaxf(0.f, 1.f - sint2));
            if (entering) cost = -cost;
            float sintOverSini = eta;
            *oWi = f3(sintOverSini * -wo.x, sintOverSini * -wo.y, cost);
            *oPdfW = 1.f;
            float R = fresnel_dielectric(wo.z, eta_i, eta_t);
            float T = 1.f - R;
            if (computeAdjoint) return T * color / fabsf(oWi->z);
            return T * color * sqr(sintOverSini) / fabsf(oWi->z);
        }
    }
    OPR_DEV float3 vcmF(const float3& wo, const float3& wi, float* oDirectPdfW, float* oReversePdfW) const
    {
        if (type & B_Lambertian) {
            if (wo.z < EPS_COSINE || wi.z < EPS_COSINE) return f3(0.f);
            *oDirectPdfW = fmaxf(0.f, wi.z * INV_PI_F);
            *oReversePdfW = fmaxf(0.f, wo.z * INV_PI_F);
            return color * INV_PI_F;
        }
        if (type & B_Phong) {
            float p = 0.f;
            float3 r = f(wo, wi, &p);
            *oDirectPdfW = p; *oReversePdfW = p;
            return r;
        }
        *oDirectPdfW = 0.f; *oReversePdfW = 0.f;
        return f3(0.f);
    }
};

struct VcmBSDF {
    float3 bitangent, tangent, normal, geometricNormal;
    int nBxDFs;
    BxDF bxdf[2];
    float pickProb[2];
    float contProb;
    bool dirFixIsLight;
    float3 localDirFix;

    OPR_DEV float3 toWorld(const float3& a) const { return bitangent * a.x + tangent * a.y + normal * a.z; }
    OPR_DEV float3 toLocal(const float3& a) const { return f3(dot(bitangent, a), dot(tangent, a), dot(normal, a)); }

    // BSDF(aWorldNormal) + VcmBSDF ctor (BSDF.h:103-110, :327-335) with DifferentialGeometry::SetFromNormal
    OPR_DEV void init(const float3& aWorldGeometricNormal, const float3& aIncidentDir, bool aIncidentDirIsLight)
    {
        normal = normalize(aWorldGeometricNormal);
        float3 tmpBiTan = (fabsf(normal.x) > 0.99f) ? f3(0.f, 1.f, 0.f) : f3(1.f, 0.f, 0.f);
        tangent = normalize(cross(normal, tmpBiTan));
        bitangent = cross(tangent, normal);
        geometricNormal = aWorldGeometricNormal;
        nBxDFs = 0; dirFixIsLight = aIncidentDirIsLight;
        localDirFix = toLocal(aIncidentDir);
        pickProb[0] = pickProb[1] = 0.f; contProb = 0.f;
    }
    // same, but from an already-local incident direction (light vertices are stored with their local direction)
    OPR_DEV void initLocal(const float3& aWorldGeometricNormal, const float3& aLocalDirFix, bool aIncidentDirIsLight)
    {
        init(aWorldGeometricNormal, f3(0.f, 0.f, 1.f), aIncidentDirIsLight);
        localDirFix = aLocalDirFix;
    }
    OPR_DEV void add(const BxDF& b)   // AddBxDF, BSDF.h:347-370
    {
        bxdf[nBxDFs] = b;
        float rr = 0.f; rr += b.continuationProb(localDirFix);
        contProb = fminf(1.f, contProb + rr);
        float al = 0.f; al += b.albedo(localDirFix);
        pickProb[nBxDFs] = al;
        nBxDFs++;
    }
    OPR_DEV int countMatching(int type) const { int c = 0; for (int i = 0; i < nBxDFs; i++) if (bx_match(bxdf[i].type, type)) ++c; return c; }
    OPR_DEV bool isSpecular() const { return countMatching(B_All & ~B_Specular) == 0; }
    OPR_DEV float sumPickProb(int type) const { float s = 0.f; for (int i = 0; i < nBxDFs; i++) if (bx_match(bxdf[i].type, type)) s += pickProb[i]; return s; }

    OPR_DEV float pdf(const float3& worldDirGen, int sampleType, bool evalRev) const   // BSDF.h:414-436
    {
        float3 wi = toLocal(worldDirGen);
        float sum = sumPickProb(sampleType);
        if (sum == 0.f) return 0.f;
        float p = 0.f;
        for (int i = 0; i < nBxDFs; i++)
            if (bx_match(bxdf[i].type, sampleType)) {
                float comp = 0.f; comp += bxdf[i].pdf(localDirFix, wi, evalRev);
                p += comp * pickProb[i] / sum;
            }
        return p;
    }

    // vcmSampleF -> sampleF (BSDF.h:440-571).  *dirWritten tells whether *oWorldDirGen was assigned.
    OPR_DEV float3 vcmSampleF(float3* oWorldDirGen, const float3& aSample, float* oPdfW, float* oCosThetaOut, int* oSampledType, bool* dirWritten) const
    {
        int aSampleType = B_All;
        *dirWritten = false;
        const float3 aWorldWo = toWorld(localDirFix);
        int index = 0;
        float matchedSum = 0.f;
        const int nMatched = countMatching(aSampleType);
        if (nMatched != 0) {   // sampleBxDF :373-396
            matchedSum = sumPickProb(aSampleType);
            float prev = 0.f;
            for (int i = 0; i < nBxDFs; ++i)
                if (bx_match(bxdf[i].type, aSampleType)) {
                    float c = pickProb[i] / matchedSum;
                    if (aSample.x < prev + c) { index = i; break; }
                    prev += c;
                }
        }
        if (matchedSum == 0.f) { *oPdfW = 0.0f; return f3(0.0f); }
        float3 wo = toLocal(aWorldWo);
        float3 wi = f3(0.f);
        float3 f = bxdf[index].sampleF(wo, &wi, make_float2(aSample.y, aSample.z), oPdfW, dirFixIsLight);
        *oPdfW *= pickProb[index] / matchedSum;
        if (*oPdfW == 0.0f) { *oSampledType = B_Null; return f3(0.0f); }
        *oSampledType = bxdf[index].type;
        *oWorldDirGen = toWorld(wi); *dirWritten = true;
        *oCosThetaOut = fabsf(wi.z);
        const bool specular = (bxdf[index].type & B_Specular) != 0;
        if (!specular && nMatched > 1) {
            for (int i = 0; i < nBxDFs; i++) {
                if (i == index) continue;
                if (bx_match(bxdf[i].type, aSampleType)) {
                    float comp = 0.f; comp += bxdf[i].pdf(wo, wi, false);
                    *oPdfW += comp * (pickProb[index] / matchedSum);   // sic (BSDF.h:541)
                }
            }
        }
        if (!specular) {
            if (dot(geometricNormal, *oWorldDirGen) * dot(geometricNormal, aWorldWo) >= 0.0f) aSampleType = aSampleType & ~B_Transmission;
            else aSampleType = aSampleType & ~B_Reflection;
            for (int i = 0; i < nBxDFs; ++i) {
                if (i == index) continue;
                if (bx_match(bxdf[i].type, aSampleType)) f += bxdf[i].f(wo, wi, nullptr);
            }
        }
        return f;
    }

    // vcmF (BSDF.h:577-639)
    OPR_DEV float3 vcmF(const float3& aWorldDirGen, float& oCosThetaGen, float* oDirectPdfW, float* oReversePdfW) const
    {
        int aSampleType = B_All;
        float3 localDirGen = toLocal(aWorldDirGen);
        float3 worldDirFix = toWorld(localDirFix);
        *oDirectPdfW = 0.f; *oReversePdfW = 0.f;
        if (localDirFix.z < EPS_COSINE || localDirGen.z < EPS_COSINE) return f3(0.f);
        oCosThetaGen = localDirGen.z;
        if (dot(geometricNormal, aWorldDirGen) * dot(geometricNormal, worldDirFix) >= 0.0f) aSampleType = aSampleType & ~B_Transmission;
        else aSampleType = aSampleType & ~B_Reflection;
        float sum = sumPickProb(aSampleType);
        if (sum == 0.f) return f3(0.0f);
        float3 f = f3(0.0f);
        for (int i = 0; i < nBxDFs; ++i)
            if (bx_match(bxdf[i].type, aSampleType)) {
                float d = 0.f, r = 0.f;
                f += bxdf[i].vcmF(localDirFix, localDirGen, &d, &r);
                d *= pickProb[i] / sum; r *= pickProb[i] / sum;
                *oDirectPdfW += d; *oReversePdfW += r;
            }
        return f;
    }
};

OPR_DEV BxDF bx_lambertian(const float3& kd) { BxDF b; b.type = B_Lambertian | B_Reflection | B_Diffuse; b.color = kd; b.exponent = 0.f; b.dielectric = 0; b.eta_i = b.eta_t = 1.f; return b; }
OPR_DEV BxDF bx_phong(const float3& ks, float e) { BxDF b; b.type = B_Phong | B_Reflection | B_Glossy; b.color = ks; b.exponent = e; b.dielectric = 0; b.eta_i = b.eta_t = 1.f; return b; }
OPR_DEV BxDF bx_spec_refl(const float3& kr, int dielectric, float ei, float et)
{ BxDF b; b.type = B_SpecularReflection | B_Reflection | B_Specular; b.color = kr; b.exponent = 0.f; b.dielectric = dielectric; b.eta_i = ei; b.eta_t = et; return b; }
OPR_DEV BxDF bx_spec_trans(const float3& kt, float ei, float et)
{ BxDF b; b.type = B_SpecularTransmission | B_Transmission | B_Specular; b.color = kt; b.exponent = 0.f; b.dielectric = 1; b.eta_i = ei; b.eta_t = et; return b; }

} // namespace opr
