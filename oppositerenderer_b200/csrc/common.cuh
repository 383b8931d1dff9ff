// common.cuh — device math, constants and RNG shared by all sm_100a kernels of the light-transport path.
//
// Vector helpers follow the semantics of the OptiX SDK math the reference's device programs rely on
// (normalize = v * (1/sqrtf(dot)), float3/float = v * (1/s)) so that results stay
// within ulps of the reference's; constants come from RenderEngine/config.h:33-43.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define OPR_DEV __device__ __forceinline__
#define OPR_HD __host__ __device__ __forceinline__

namespace opr {

// ---- RenderEngine/config.h:33-43, BxDF.h:266, OptixRenderer.cpp:39 ---------------------------------
constexpr unsigned MAX_PHOTON_TRACE_DEPTH = 7;
constexpr unsigned MAX_RADIANCE_TRACE_DEPTH = 9;
constexpr unsigned PHOTON_TRACING_RR_START_DEPTH = 3;
constexpr float RAY_LEN_MIN = 0.0001f;
constexpr float EPS_COSINE = 1e-6f;
constexpr float EPS_RAY = 1e-3f;
constexpr float EPS_PHONG = 1e-3f;
constexpr unsigned MAX_PHOTON_DEPOSITS = 4;             // config.h:26
constexpr unsigned PHOTON_GRID_MAX_SIZE = 100 * 100 * 100;
constexpr float RT_DEFAULT_MAX = 1.e27f;
constexpr float PI_F = 3.14159265358979323846f;
constexpr float INV_PI_F = 0.318309886183790671538f;

// RadiancePRD.h:31-36
constexpr unsigned PRD_HIT_EMITTER = 1u << 31;
constexpr unsigned PRD_ERROR = 1u << 30;
constexpr unsigned PRD_MISS = 1u << 29;
constexpr unsigned PRD_HIT_SPECULAR = 1u << 28;
constexpr unsigned PRD_HIT_NON_SPECULAR = 1u << 27;

// ---- float3 helpers ------------------------------------------------------------------------------------
OPR_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
OPR_HD float3 f3(float s) { return make_float3(s, s, s); }
OPR_HD float3 f3(const float4& a) { return make_float3(a.x, a.y, a.z); }
OPR_HD float3 operator-(const float3& a) { return f3(-a.x, -a.y, -a.z); }
OPR_HD float3 operator+(const float3& a, const float3& b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
OPR_HD float3 operator-(const float3& a, const float3& b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
OPR_HD float3 operator-(const float3& a, float b) { return f3(a.x - b, a.y - b, a.z - b); }
OPR_HD float3 operator+(const float3& a, float b) { return f3(a.x + b, a.y + b, a.z + b); }
OPR_HD float3 operator*(const float3& a, const float3& b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
OPR_HD float3 operator*(const float3& a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
OPR_HD float3 operator*(float s, const float3& a) { return f3(a.x * s, a.y * s, a.z * s); }
OPR_HD float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
OPR_HD float3 operator/(const float3& a, const float3& b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
OPR_HD float3& operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
OPR_HD float3& operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
OPR_HD float3& operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
OPR_HD float3& operator/=(float3& a, float s) { float inv = 1.0f / s; a *= inv; return a; }
OPR_HD float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OPR_HD float3 cross(const float3& a, const float3& b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
OPR_HD float length(const float3& v) { return sqrtf(dot(v, v)); }
OPR_HD float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
OPR_HD float fmax3(const float3& a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
OPR_HD float3 fmin3(const float3& a, const float3& b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
OPR_HD float3 fmax3(const float3& a, const float3& b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
OPR_HD float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
OPR_HD bool isZero(const float3& v) { return v.x == 0.f && v.y == 0.f && v.z == 0.f; }
OPR_HD float sqr(float a) { return a * a; }
OPR_HD float maxf_(float a, float b) { return a > b ? a : b; }      // helpers.h:139-142
OPR_HD float favgf(const float3& v) { return (v.x + v.y + v.z) * 0.3333333333f; }
OPR_HD float luminanceCIE(const float3& rgb) { return dot(rgb, f3(0.212671f, 0.715160f, 0.072169f)); }

// OptiX SDK refract (optixu_math_namespace.h), restated
OPR_HD bool refract(float3& r, const float3& i, const float3& n, const float ior)
{
    float3 nn = n;
    float negNdotV = dot(i, nn);
    float eta;
    if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; }
    else { eta = 1.f / ior; }
    const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
    if (k < 0.0f) return false;
    r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
    return true;
}

// helpers/helpers.h:119-135
OPR_HD void createCoordinateSystem(const float3& N, float3& U, float3& V)
{
    if (fabsf(N.x) > fabsf(N.y)) {
        float invLength = 1.f / sqrtf(N.x * N.x + N.z * N.z);
        U = f3(-N.z * invLength, 0.f, N.x * invLength);
    } else {
        float invLength = 1.f / sqrtf(N.y * N.y + N.z * N.z);
        U = f3(0.f, N.z * invLength, -N.y * invLength);
    }
    V = cross(N, U);
}

// ---- XORWOW (curand_kernel.h: _curand_init_inplace with subsequence = offset = 0, curand()) ------------
// helpers/random.h:59-69.  State is 6 words {d, v0..v4}; kept SoA as three uint2 arrays in HBM.
struct Rng { unsigned d, v0, v1, v2, v3, v4; };

OPR_HD void rng_seed(Rng& s, unsigned seed, unsigned index)
{
    unsigned sd = seed + index;                 // `seed+index` is an unsigned int sum (random.h:61)
    unsigned s0 = sd ^ 0xaad26b49u;
    unsigned s1 = 0u ^ 0xf7dcefddu;             // high word of the 64-bit seed is zero
    unsigned t0 = 1099087573u * s0;
    unsigned t1 = 2591861531u * s1;
    s.d = 6615241u + t1 + t0;
    s.v0 = 123456789u + t0;
    s.v1 = 362436069u ^ t0;
    s.v2 = 521288629u + t1;
    s.v3 = 88675123u ^ t1;
    s.v4 = 5783321u + t0;
}

OPR_HD unsigned rng_next(Rng& s)
{
    unsigned t = (s.v0 ^ (s.v0 >> 2));
    s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return s.v4 + s.d;
}

// getRandomUniformFloat = max(curand_uniform - FLT_EPSILON, 0); curand_uniform = x*2^-32 + 2^-33 as one FMA
OPR_HD float rng_uniform(Rng& s)
{
    const float k = 2.3283064e-10f;
    float u = fmaf((float)rng_next(s), k, k / 2.0f);
    float r = u - FLT_EPSILON;
    return r > 0.0f ? r : 0.0f;
}
OPR_HD float2 rng_uniform2(Rng& s) { float2 r; r.x = rng_uniform(s); r.y = rng_uniform(s); return r; }
OPR_HD float3 rng_uniform3(Rng& s) { float3 r; r.x = rng_uniform(s); r.y = rng_uniform(s); r.z = rng_uniform(s); return r; }

// One 32-byte record per state, (d, v0, v1, v2) (v3, v4, -, -), read and written with one 256-bit access: the shading kernels
// reach the states through path indices in queue order, where three separate arrays cost three sectors per read and three per
// write; the generation kernels walk the states in order and stay fully coalesced.
struct RngBuf { uint4* s; };
OPR_DEV Rng rng_load(const RngBuf& buf, size_t i)
{
    Rng s; [[maybe_unused]] unsigned p0, p1;
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s.d), "=r"(s.v0), "=r"(s.v1), "=r"(s.v2), "=r"(s.v3), "=r"(s.v4), "=r"(p0), "=r"(p1) : "l"(buf.s + 2 * i) : "memory");
    return s;
}
OPR_DEV void rng_store(const RngBuf& buf, size_t i, const Rng& s)
{
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(buf.s + 2 * i), "r"(s.d), "r"(s.v0), "r"(s.v1), "r"(s.v2), "r"(s.v3), "r"(s.v4), "r"(0u), "r"(0u) : "memory");
}

// ---- samplers (helpers/samplers.h:21-95) ----------------------------------------------------------------
OPR_DEV float3 sampleUnitHemisphereCos(const float3& normal, const float2& sample, float* oPdfW = nullptr,
                                       float* oCosTheta = nullptr, bool biasSmallCosine = false)
{
    float theta = acosf(sqrtf(sample.x));
    float phi = 2.0f * PI_F * sample.y;
    float st, ct, sp, cp;
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
    float xs = st * cp;
    float ys = ct;
    float zs = st * sp;
    float3 U, V;
    createCoordinateSystem(normal, U, V);
    if (biasSmallCosine && ys < EPS_COSINE) ys = EPS_COSINE;
    if (oPdfW) *oPdfW = ys * INV_PI_F;
    if (oCosTheta) *oCosTheta = ys;
    return normalize(xs * U + ys * normal + zs * V);
}

OPR_DEV float3 sampleUnitHemisphere(const float3& normal, const float2& sample)
{
    float3 U, V;
    createCoordinateSystem(normal, U, V);
    float phi = 2.0f * PI_F * sample.x;
    float r = sqrtf(sample.y);
    float sp, cp; sincosf(phi, &sp, &cp);
    float x = r * cp;
    float y = r * sp;
    float z = 1.0f - x * x - y * y;
    z = z > 0.0f ? sqrtf(z) : 0.0f;
    return normalize(U * x + V * y + normal * z);
}

OPR_DEV float3 sampleUnitSphere(const float2& sample, float* oPdfW = nullptr)
{
    float3 v;
    v.z = 1.f - 2.f * sample.x;
    float phi = 2 * PI_F * sample.y;
    float r = sqrtf(1.f - v.z * v.z);
    float sp, cp; sincosf(phi, &sp, &cp);
    v.x = r * cp;
    v.y = r * sp;
    if (oPdfW) *oPdfW = 0.25f * INV_PI_F;
    return v;
}

OPR_DEV float2 sampleUnitDisc(const float2& sample)
{
    float r = sqrtf(sample.x);
    float theta = 2.f * PI_F * sample.y;
    float st, ct; sincosf(theta, &st, &ct);
    return make_float2(r * ct, r * st);
}

OPR_DEV float3 sampleDisc(const float2& sample, const float3& center, const float radius, const float3& normal)
{
    float3 U, V;
    createCoordinateSystem(normal, U, V);
    float2 unitDisc = sampleUnitDisc(sample);
    return center + radius * (U * unitDisc.x + V * unitDisc.y);
}

// ---- warp-aggregated queue append --------------------------------------------------------------------
// One atomicAdd per warp per queue: lanes that push elect a leader (ballot), the leader reserves a
// contiguous range and every pusher gets base + its rank.  Used for ray regeneration / material queues.
OPR_DEV unsigned warp_append(unsigned* counter, bool push)
{
    unsigned mask = __ballot_sync(0xffffffffu, push);
    if (!push) return 0xffffffffu;
    unsigned lane = threadIdx.x & 31;
    unsigned leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1));
}

// Path regeneration of the per-path tail kernels (ppm.cu, vcm.cu): persistent warps pull paths from a queue through a shared cursor;
// a lane whose path ended takes the next one as soon as `refillIdle` lanes of its warp are idle.  Returns true for a lane that must
// load path `index`; all 32 lanes call this together.
OPR_DEV bool tail_refill(unsigned n, unsigned* __restrict__ cursor, bool alive, bool& exhausted, unsigned refillIdle, unsigned& index)
{   // returns true for a lane that must load path `index`; all 32 lanes call this together
    const unsigned FULL = 0xffffffffu;
    const unsigned idle = __ballot_sync(FULL, !alive);
    bool take = false;
    if (!exhausted && (idle == FULL || (unsigned)__popc(idle) >= refillIdle)) {
        const unsigned lane = threadIdx.x & 31, want = __popc(idle);
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(cursor, want);
        base = __shfl_sync(FULL, base, 0);
        if (!alive) { index = base + __popc(idle & ((1u << lane) - 1)); take = index < n; }
        if (base + want >= n) exhausted = true;
    }
    return take;
}

// Pixel of queue slot s: a warp's 32 consecutive slots cover an 8 x 4 pixel tile instead of a 32 x 1 strip (when the image divides into
// tiles), so the lanes of a primary-ray or shadow-ray warp start closer together and walk more of the tree side by side.  Per-pixel
// results do not depend on the slot order.
OPR_HD unsigned tiled_pixel(unsigned s, unsigned w, unsigned h)
{
    if ((w & 7u) | (h & 3u)) return s;
    const unsigned tile = s >> 5, l = s & 31u, tilesX = w >> 3;
    return ((tile / tilesX) * 4u + (l >> 3)) * w + (tile % tilesX) * 8u + (l & 7u);
}

} // namespace opr

namespace opr {
// devmem.cu: cudaMalloc / cudaFree with optional guard bands (opr_debug_set_guards)
void devAlloc(void** out, size_t bytes);
void devFree(void* p);
void devSetGuards(bool on);
int devCheckGuards(int* liveBlocks);
// multiprocessor count of the current device (cudaDevAttrMultiProcessorCount, cached per device): every grid is sized from it
int numSMs();
}

#define OPR_CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)
