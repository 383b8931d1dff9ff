// ORACLE — test infrastructure only (see ora_math.h).  Software restatement of the texture fetch the reference
// performs on the hardware texture unit.
#include "ora_core.h"

namespace ora {

// ---- texture fetch: rtTextureSampler<uchar4,2,cudaReadModeNormalizedFloat>, wrap=repeat, linear,
// normalised coordinates (material/Texture.cpp:109-121).  The reference uses the hardware texture unit;
// this follows the CUDA programming guide's "Texture Fetching" appendix: xB = N*frac(x) - 0.5,
// i = floor(xB), alpha = frac(xB) kept in 1.8 fixed point, tex = (1-a)(1-b)T[i,j] + ...
float4 tex2D_raw(const uint8_t* rgba, int width, int height, float u, float v)
{
    if (width <= 0 || height <= 0 || !rgba) return make_float4(0, 0, 0, 0);
    float x = (u - floorf(u)) * width - 0.5f;
    float y = (v - floorf(v)) * height - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = floorf((x - fx) * 256.0f + 0.5f) * (1.0f / 256.0f);
    float b = floorf((y - fy) * 256.0f + 0.5f) * (1.0f / 256.0f);
    int i0 = (int)fx, j0 = (int)fy;
    auto wrap = [](int i, int n) { i %= n; return i < 0 ? i + n : i; };
    int i1 = wrap(i0 + 1, width), j1 = wrap(j0 + 1, height);
    i0 = wrap(i0, width); j0 = wrap(j0, height);
    auto T = [&](int i, int j, int c) { return rgba[((size_t)j * width + i) * 4 + c] * (1.0f / 255.0f); };
    float4 r;
    float* rp = &r.x;
    for (int c = 0; c < 4; c++)
        rp[c] = (1 - a) * (1 - b) * T(i0, j0, c) + a * (1 - b) * T(i1, j0, c) + (1 - a) * b * T(i0, j1, c) + a * b * T(i1, j1, c);
    return r;
}
float4 tex2D(const Texture& t, float u, float v) { return tex2D_raw(t.rgba.data(), t.width, t.height, u, v); }


} // namespace ora
