// ORACLE — test infrastructure only. Nothing under oracle/ is part of the product; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may build or call it.
// PARITY PINNED: against oracle/_ref — the reference's own device programs host-compiled from /root/reference by oracle/refshim
// (tests/test_oracle_vs_ref.py; the same outputs are committed as tests/golden/*.npz) — and, for the XORWOW generator, against
// NVIDIA's cuRAND host generator (tests/test_oracle_cpu.py).
// Pinned inputs: every built-in Cornell / CornellSmall case, BASELINE configs 1-4 at their stated sizes, config 5's workload at 960x540,
// the spot-light / depth-of-field room and the imported COLLADA room of the GPU parity tests (integer results identical, images 1e-5).
// One case has no pin because the reference is undefined there: a SPOT light in the BDPT passes (helpers/light.h:136-139,208-211).
//
// ora_math.h — vector types plus the OptiX SDK 3.5/3.6 `optixu_math_namespace.h` functions the hot
// path uses.  The OptiX SDK is NOT vendored in /root/reference (SDKs.props:16-17), so these restate
// the SDK's published formulas (SURVEY.md §8c "Arithmetic living in third-party code"):
//   normalize = v * (1/sqrtf(dot)), float3/float = v * (1/s), reflect, refract,
//   intersect_triangle_branchless, cosine_sample_hemisphere, luminanceCIE.
// All arithmetic is IEEE fp32, evaluated left to right, no FMA contraction (-ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <cfloat>
#include <algorithm>

namespace ora {

typedef unsigned int uint;

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint2 { uint x, y; };
struct uint3 { uint x, y, z; };
struct int3 { int x, y, z; };

static const float M_PIf_ = 3.14159265358979323846f;
static const float M_1_PIf_ = 0.318309886183790671538f;

inline float2 make_float2(float x, float y) { return {x, y}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float3 make_float3(float s) { return {s, s, s}; }
inline float3 make_float3(const float4& a) { return {a.x, a.y, a.z}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline float4 make_float4(const float3& a, float w) { return {a.x, a.y, a.z, w}; }

// float2
inline float2 operator+(const float2& a, const float2& b) { return {a.x + b.x, a.y + b.y}; }
inline float2 operator-(const float2& a, float b) { return {a.x - b, a.y - b}; }
inline float2 operator*(const float2& a, float s) { return {a.x * s, a.y * s}; }
inline float2 operator*(float s, const float2& a) { return {a.x * s, a.y * s}; }
inline float2 operator/(const float2& a, const float2& b) { return {a.x / b.x, a.y / b.y}; }

// float3
inline float3 operator-(const float3& a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator+(const float3& a, const float3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator+(const float3& a, float b) { return {a.x + b, a.y + b, a.z + b}; }
inline float3 operator-(const float3& a, const float3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(const float3& a, float b) { return {a.x - b, a.y - b, a.z - b}; }
inline float3 operator*(const float3& a, const float3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(const float3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, const float3& a) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator/(const float3& a, const float3& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
// SDK: float3 / float multiplies by the reciprocal.
inline float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
inline float3 operator/(float s, const float3& a) { return {s / a.x, s / a.y, s / a.z}; }
inline float3& operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline float3& operator-=(float3& a, const float3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline float3& operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
inline float3& operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
inline float3& operator/=(float3& a, float s) { float inv = 1.0f / s; a *= inv; return a; }

inline float4 operator*(float s, const float4& a) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline float4 operator-(const float4& a, float b) { return {a.x - b, a.y - b, a.z - b, a.w - b}; }

inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
inline float fmaxf3(const float3& a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
inline float3 fminf3(const float3& a, const float3& b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline float3 fmaxf3(const float3& a, const float3& b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
inline float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }

inline bool refract(float3& r, const float3& i, const float3& n, const float ior)
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

inline float luminanceCIE(const float3& rgb)
{
    const float3 ntsc = {0.212671f, 0.715160f, 0.072169f};
    return dot(rgb, ntsc);
}

// SDK cosine_sample_hemisphere(u1,u2,p): r=sqrt(u1), phi=2*pi*u2, p=(r cos, r sin, sqrt(max(0,1-x^2-y^2)))
inline void cosine_sample_hemisphere(float u1, float u2, float3& p)
{
    const float r = sqrtf(u1);
    const float phi = 2.0f * M_PIf_ * u2;
    p.x = r * cosf(phi);
    p.y = r * sinf(phi);
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}

struct Ray {
    float3 origin; float3 direction; unsigned ray_type; float tmin; float tmax;
};
static const float RT_DEFAULT_MAX_ = 1.e27f;

// SDK intersect_triangle_branchless
inline bool intersect_triangle(const Ray& ray, const float3& p0, const float3& p1, const float3& p2,
                               float3& n, float& t, float& beta, float& gamma)
{
    const float3 e0 = p1 - p0;
    const float3 e1 = p0 - p2;
    n = cross(e1, e0);
    const float3 e2 = (1.0f / dot(n, ray.direction)) * (p0 - ray.origin);
    const float3 i = cross(ray.direction, e2);
    beta = dot(i, e1);
    gamma = dot(i, e0);
    t = dot(n, e2);
    return ((t < ray.tmax) & (t > ray.tmin) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1));
}

// CUDA float->uint conversion semantics (cvt.rzi.u32.f32 saturates; NaN -> 0).  The reference relies on
// them implicitly (PhotonGrid.h:12-18, IndirectRadianceEstimation.cu:89-95).
inline uint f2u_sat(float f)
{
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint)f;
}
inline int f2i_sat(float f)
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 0x7fffffff;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}

} // namespace ora
