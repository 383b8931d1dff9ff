// ORACLE/_ref — test infrastructure only.  A minimal host-side stand-in for the OptiX 3.x device API so that the
// reference's own device programs (RenderEngine/renderer/ppm/*.cu, renderer/vcm/*, material/*.cu,
// geometry_instance/*.cu and the headers they include) compile UNCHANGED as host C++ where they lie under
// /root/reference.  Nothing of the reference is copied: this header only supplies what the OptiX SDK, CUDA and
// curand headers would (SURVEY.md §8c lists the symbols).  The OptiX SDK math is restated from its published
// formulas, exactly as in oracle/ora_math.h.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <cstdint>
#include <ctime>
#include <algorithm>
#include <string>
#include <typeinfo>
#include <vector>
#include <stdint.h>
#include <float.h>
#include <stdio.h>

#ifndef __CUDACC__
#define __CUDACC__ 1          /* the reference selects its device-side declarations with this */
#endif
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __inline__ inline
#define RT_PROGRAM
#define RT_HOSTDEVICE
#define RT_DEFAULT_MAX 1.e27f
#define RENDER_ENGINE_EXPORT_API
#define RENDER_ENGINE_EXPORT_API_QT
#define RT_OBJECT_TO_WORLD 0
#define RT_WORLD_TO_OBJECT 1
#ifndef M_PIf
#define M_PIf 3.14159265358979323846f
#endif
#ifndef M_1_PIf
#define M_1_PIf 0.318309886183790671538f
#endif

// ---- CUDA vector types (vector_types.h) ------------------------------------------------------------------
typedef unsigned int uint;
// defined inside namespace optix so that argument-dependent lookup finds the SDK operators even where the
// reference declares its own operators in an enclosing scope (math/Vector3.h:62-76)
namespace optix {
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
struct char4 { signed char x, y, z, w; };
struct size_t2 { size_t x, y; };
struct size_t3 { size_t x, y, z; };
}
using optix::float2; using optix::float3; using optix::float4; using optix::int2; using optix::int3; using optix::uint2; using optix::uint3;
using optix::uchar4; using optix::char4; using optix::size_t2; using optix::size_t3;
enum cudaTextureReadMode { cudaReadModeElementType = 0, cudaReadModeNormalizedFloat = 1 };

inline float __int_as_float(int a) { float f; memcpy(&f, &a, 4); return f; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
// CUDA's global min/max overloads (math_functions.h)
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// ---- curand XORWOW (curand_kernel.h) -----------------------------------------------------------------------
struct curandStateXORWOW { unsigned int d, v[5]; int boxmuller_flag; int boxmuller_flag_double; float boxmuller_extra; double boxmuller_extra_double; };
typedef curandStateXORWOW curandState;
inline void curand_init(unsigned long long seed, unsigned long long, unsigned long long, curandState* state)
{
    unsigned int s0 = ((unsigned int)seed) ^ 0xaad26b49UL;
    unsigned int s1 = (unsigned int)(seed >> 32) ^ 0xf7dcefddUL;
    unsigned int t0 = 1099087573UL * s0;
    unsigned int t1 = 2591861531UL * s1;
    state->d = 6615241 + t1 + t0;
    state->v[0] = 123456789UL + t0; state->v[1] = 362436069UL ^ t0; state->v[2] = 521288629UL + t1;
    state->v[3] = 88675123UL ^ t1; state->v[4] = 5783321UL + t0;
    state->boxmuller_flag = 0; state->boxmuller_flag_double = 0; state->boxmuller_extra = 0.f; state->boxmuller_extra_double = 0.;
}
inline unsigned int curand(curandState* state)
{
    unsigned int t = (state->v[0] ^ (state->v[0] >> 2));
    state->v[0] = state->v[1]; state->v[1] = state->v[2]; state->v[2] = state->v[3]; state->v[3] = state->v[4];
    state->v[4] = (state->v[4] ^ (state->v[4] << 4)) ^ (t ^ (t << 1));
    state->d += 362437;
    return state->v[4] + state->d;
}
inline float curand_uniform(curandState* state) { const float k = 2.3283064e-10f; return fmaf((float)curand(state), k, k / 2.0f); }

// ---- OptiX SDK math (optixu_math_namespace.h), restated ----------------------------------------------------
namespace optix {
typedef unsigned int uint;

inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
inline float2 make_float2(float s) { float2 r = {s, s}; return r; }
inline float2 make_float2(const uint2& a) { float2 r = {float(a.x), float(a.y)}; return r; }
inline float2 make_float2(const size_t2& a) { float2 r = {float(a.x), float(a.y)}; return r; }
inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
inline float3 make_float3(float s) { float3 r = {s, s, s}; return r; }
inline float3 make_float3(const float4& a) { float3 r = {a.x, a.y, a.z}; return r; }
inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
inline float4 make_float4(const float3& a, float w) { float4 r = {a.x, a.y, a.z, w}; return r; }
inline float4 make_float4(const float3& a) { float4 r = {a.x, a.y, a.z, 0.f}; return r; }
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
inline uint3 make_uint3(unsigned x, unsigned y, unsigned z) { uint3 r = {x, y, z}; return r; }
inline int3 make_int3(int x, int y, int z) { int3 r = {x, y, z}; return r; }

inline float2 operator+(const float2& a, const float2& b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 operator-(const float2& a, const float2& b) { return make_float2(a.x - b.x, a.y - b.y); }
inline float2 operator-(const float2& a, float b) { return make_float2(a.x - b, a.y - b); }
inline float2 operator*(const float2& a, float s) { return make_float2(a.x * s, a.y * s); }
inline float2 operator*(float s, const float2& a) { return make_float2(a.x * s, a.y * s); }
inline float2 operator*(const float2& a, const float2& b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 operator/(const float2& a, const float2& b) { return make_float2(a.x / b.x, a.y / b.y); }
inline float2 operator/(const float2& a, float s) { float inv = 1.0f / s; return a * inv; }

inline float3 operator-(const float3& a) { return make_float3(-a.x, -a.y, -a.z); }
inline float3 operator+(const float3& a, const float3& b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator+(const float3& a, float b) { return make_float3(a.x + b, a.y + b, a.z + b); }
inline float3 operator-(const float3& a, const float3& b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(const float3& a, float b) { return make_float3(a.x - b, a.y - b, a.z - b); }
inline float3 operator*(const float3& a, const float3& b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(const float3& a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, const float3& a) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator/(const float3& a, const float3& b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
inline float3 operator/(float s, const float3& a) { return make_float3(s / a.x, s / a.y, s / a.z); }
inline void operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline void operator-=(float3& a, const float3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
inline void operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
inline void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
inline void operator/=(float3& a, float s) { float inv = 1.0f / s; a *= inv; }
inline float4 operator*(float s, const float4& a) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float4 operator*(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float4 operator-(const float4& a, float b) { return make_float4(a.x - b, a.y - b, a.z - b, a.w - b); }

inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
using ::fminf; using ::fmaxf; using ::min; using ::max;
inline float fmaxf(const float3& a) { return ::fmaxf(::fmaxf(a.x, a.y), a.z); }
inline float fminf(const float3& a) { return ::fminf(::fminf(a.x, a.y), a.z); }
inline float3 fminf(const float3& a, const float3& b) { return make_float3(::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z)); }
inline float3 fmaxf(const float3& a, const float3& b) { return make_float3(::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z)); }
inline float clamp(float f, float a, float b) { return ::fmaxf(a, ::fminf(f, b)); }
inline unsigned clamp(unsigned f, unsigned a, unsigned b) { return std::max(a, std::min(f, b)); }
inline int clamp(int f, int a, int b) { return std::max(a, std::min(f, b)); }
inline float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
inline bool refract(float3& r, const float3& i, const float3& n, const float ior)
{
    float3 nn = n;
    float negNdotV = dot(i, nn);
    float eta;
    if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; } else { eta = 1.f / ior; }
    const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
    if (k < 0.0f) return false;
    r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
    return true;
}
inline float luminanceCIE(const float3& rgb) { const float3 ntsc = {0.212671f, 0.715160f, 0.072169f}; return dot(rgb, ntsc); }
inline void cosine_sample_hemisphere(const float u1, const float u2, float3& p)
{
    const float r = sqrtf(u1);
    const float phi = 2.0f * M_PIf * u2;
    p.x = r * cosf(phi);
    p.y = r * sinf(phi);
    p.z = sqrtf(::fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}

struct Ray {
    float3 origin; float3 direction; unsigned int ray_type; float tmin; float tmax;
    Ray() {}
    Ray(float3 o, float3 d, unsigned int type, float tmin_, float tmax_ = RT_DEFAULT_MAX) : origin(o), direction(d), ray_type(type), tmin(tmin_), tmax(tmax_) {}
};

inline bool intersect_triangle(const Ray& ray, const float3& p0, const float3& p1, const float3& p2, float3& n, float& t, float& beta, float& gamma)
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

struct Aabb {
    float3 m_min, m_max;
    void invalidate() { m_min = make_float3(1e37f); m_max = make_float3(-1e37f); }
};
struct Matrix4x4 { float m[16]; };
} // namespace optix
using namespace optix;

// ---- OptiX device API -----------------------------------------------------------------------------------------
struct rtObject { int id; };

namespace shim {

struct BufferDesc { void* data; size_t elemSize; size_t dim[3]; };
BufferDesc* bufferById(int id);

template <class T, int N> struct BufSize { typedef size_t type; static type get(const size_t* d) { return d[0]; } };
template <class T> struct BufSize<T, 2> { typedef size_t2 type; static type get(const size_t* d) { size_t2 r = {d[0], d[1]}; return r; } };
template <class T> struct BufSize<T, 3> { typedef size_t3 type; static type get(const size_t* d) { size_t3 r = {d[0], d[1], d[2]}; return r; } };

template <class T, int N = 1>
struct Buffer {
    T* data; size_t dim[3];
    Buffer() : data(nullptr) { dim[0] = dim[1] = dim[2] = 0; }
    explicit Buffer(int id) { BufferDesc* b = bufferById(id); data = (T*)b->data; dim[0] = b->dim[0]; dim[1] = b->dim[1]; dim[2] = b->dim[2]; }
    void bind(const BufferDesc& b) { data = (T*)b.data; dim[0] = b.dim[0]; dim[1] = b.dim[1]; dim[2] = b.dim[2]; }
    typename BufSize<T, N>::type size() const { return BufSize<T, N>::get(dim); }
    T& operator[](size_t i) const { return data[i]; }
    T& operator[](const uint2& i) const { return data[(size_t)i.y * dim[0] + i.x]; }
    T& operator[](const size_t2& i) const { return data[i.y * dim[0] + i.x]; }
    T& operator[](const uint3& i) const { return data[((size_t)i.z * dim[1] + i.y) * dim[0] + i.x]; }
};

struct TextureData { int width, height; const unsigned char* rgba; };
template <class T, int D, int M>
struct TexSampler { const TextureData* tex; TexSampler() : tex(nullptr) {} };
float4 sampleTexture(const TextureData* t, float u, float v);

// variable registry of one translation unit (= one reference .cu)
struct Var { const char* name; const char* semantic; size_t size; const std::type_info* type; void* (*get)(); };
struct TU {
    const char* name;
    std::vector<Var> vars;
    explicit TU(const char* n);
};
struct VarReg { VarReg(TU& tu, const char* name, const char* semantic, size_t size, const std::type_info* type, void* (*get)()) { Var v = {name, semantic, size, type, get}; tu.vars.push_back(v); } };

// per-thread trace state used by the intersection / any-hit API
bool potentialIntersection(float t);
bool reportIntersection(unsigned material);
void terminateRay();
void ignoreIntersection();
void trace(const optix::Ray& ray, void* payload, const std::type_info& payloadType);

} // namespace shim

#define rtBuffer thread_local shim::Buffer
#define rtBufferId shim::Buffer
#define rtTextureSampler thread_local shim::TexSampler
#define SHIM_CAT2(a, b) a##b
#define SHIM_CAT(a, b) SHIM_CAT2(a, b)
// semantic / annotation may be empty or two tokens ("attribute geometricNormal"): stringified as written
#define rtDeclareVariable(type, name, semantic, annotation) \
    extern thread_local type name; \
    static shim::VarReg SHIM_CAT(shim_reg_, name)(shim_tu, #name, #semantic, sizeof(type), &typeid(type), []() -> void* { return (void*)&name; }); \
    thread_local type name

template <class PRD> inline void rtTrace(const rtObject&, const optix::Ray& ray, PRD& prd) { shim::trace(ray, &prd, typeid(PRD)); }
inline bool rtPotentialIntersection(float t) { return shim::potentialIntersection(t); }
inline bool rtReportIntersection(unsigned m) { return shim::reportIntersection(m); }
inline void rtTerminateRay() { shim::terminateRay(); }
inline void rtIgnoreIntersection() { shim::ignoreIntersection(); }
inline float3 rtTransformNormal(int, const float3& n) { return n; }   // no transforms are instantiated
inline float3 rtTransformPoint(int, const float3& p) { return p; }
inline float3 rtTransformVector(int, const float3& v) { return v; }
inline void rtPrintExceptionDetails() {}
#define rtPrintf(...) ((void)0)
template <class T, int D, int M> inline float4 tex2D(const shim::TexSampler<T, D, M>& s, float u, float v) { return shim::sampleTexture(s.tex, u, v); }
