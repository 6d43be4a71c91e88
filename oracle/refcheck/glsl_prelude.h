// oracle/refcheck/glsl_prelude.h -- TEST INFRASTRUCTURE: a GLSL-450 environment on top of the
// reference's own vendored glm (Vendor/glm 0.9.9.9), so that the reference's fragment shaders
// (Sources/Shaders/lib/Light.frag, LightAmbient.frag, LightPoint.frag, LightSpot.frag,
// LightReflection.frag) compile as C++ FROM WHERE THEY LIE and run on the host.  Built by
// oracle/refcheck/build_shaders.py into oracle/_ref/libvxshader.so; nothing in the product or in
// the oracle includes this file.
//
// What is emulated here is only what the Vulkan implementation supplies and the shader source does
// not: bindless resource arrays, texel fetches with the fixed-point decode rules of the formats the
// reference creates (Sources/Graphics/Graphics.h:51-60), out-of-range texelFetch -> 0, nearest
// sampling at pixel centres, and GLSL's implicit int -> float conversions that glm's templates do
// not perform.  All arithmetic of the shaders themselves is executed by their own source text
// with glm's implementations of the GLSL built-ins (float literals are single precision:
// the translation unit is compiled with -fsingle-precision-constant -ffp-contract=off).
#pragma once
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace vxref {

struct FetchLog { long long count; int x, y, z; };
extern thread_local FetchLog g_fetch;

// one logged call of raycastShadowVolumeSparse / SuperSparse (48 bytes)
struct RayRecord {
    float ox, oy, oz, dx, dy, dz, dist;   // arguments as the shader computed them
    float result;                          // returned d
    int variant;                           // 0 Sparse, 1 SuperSparse
    int fetches;                           // texelFetch calls on the shadow volume inside the call
    int lx, ly;                            // x, y of the last fetch (lz below)
};
struct RayLog { int n; RayRecord r[4]; int lz[4]; };
extern thread_local RayLog g_rays;
extern thread_local bool g_discarded;

}  // namespace vxref

using namespace glm;

// ---- resources --------------------------------------------------------------------------------------
struct usampler3D { const uint8_t* data; int sx, sy, sz; const uint8_t* mip1; const uint8_t* mip2; };   // mips: VoxAsset volumes (3 levels, sizes halve)
enum TexFmt { TEX_NONE = 0, TEX_D24, TEX_RGBA8_SNORM, TEX_RGBA8_UNORM, TEX_RGBA32F, TEX_RG32F };   // the float formats hold the light / motion planes (values before the RGBA16F / RG16F attachment conversion)
struct sampler2D { const uint32_t* data; int w, h; TexFmt fmt; };
struct samplerCube { float r, g, b; };                                     // a uniform sky (the cube map itself is outside the path)

inline vec4 vxref_decode(const sampler2D& t, int x, int y) {
    if (!t.data || x < 0 || y < 0 || x >= t.w || y >= t.h) return vec4(0.0f);
    if (t.fmt == TEX_RGBA32F) { const float* f = (const float*)t.data + ((size_t)y * t.w + x) * 4; return vec4(f[0], f[1], f[2], f[3]); }
    if (t.fmt == TEX_RG32F) { const float* f = (const float*)t.data + ((size_t)y * t.w + x) * 2; return vec4(f[0], f[1], 0.0f, 1.0f); }
    const uint32_t v = t.data[(size_t)y * t.w + x];
    switch (t.fmt) {
        case TEX_D24: return vec4((float)(v & 0xFFFFFFu) / 16777215.0f, 0.0f, 0.0f, 1.0f);      // D24_UNORM
        case TEX_RGBA8_SNORM: {
            auto s = [](uint32_t c) { return fmaxf((float)(int8_t)(c & 0xFFu) / 127.0f, -1.0f); };   // Vulkan spec: SNORM decode
            return vec4(s(v), s(v >> 8), s(v >> 16), s(v >> 24));
        }
        case TEX_RGBA8_UNORM: {
            auto u = [](uint32_t c) { return (float)(c & 0xFFu) / 255.0f; };
            return vec4(u(v), u(v >> 8), u(v >> 16), u(v >> 24));
        }
        default: return vec4(0.0f);
    }
}
inline uvec4 texelFetch(const usampler3D& t, ivec3 p, int lod) {
    vxref::g_fetch.count++; vxref::g_fetch.x = p.x; vxref::g_fetch.y = p.y; vxref::g_fetch.z = p.z;
    const uint8_t* d = lod == 0 ? t.data : (lod == 1 ? t.mip1 : (lod == 2 ? t.mip2 : nullptr));
    int sx = t.sx, sy = t.sy, sz = t.sz;
    for (int l = 0; l < lod; ++l) { sx = sx > 1 ? sx / 2 : 1; sy = sy > 1 ? sy / 2 : 1; sz = sz > 1 ? sz / 2 : 1; }
    if (!d || p.x < 0 || p.y < 0 || p.z < 0 || p.x >= sx || p.y >= sy || p.z >= sz) return uvec4(0u);
    return uvec4((uint)d[(size_t)p.x + (size_t)p.y * sx + (size_t)p.z * sx * sy], 0u, 0u, 1u);
}
inline vec4 texelFetch(const sampler2D& t, ivec2 p, int) { return vxref_decode(t, p.x, p.y); }
inline ivec3 textureSize(const usampler3D& t, int) { return ivec3(t.sx, t.sy, t.sz); }
inline ivec2 textureSize(const sampler2D& t, int) { return ivec2(t.w, t.h); }
// nearest filtering (the reference's samplers, Vendor/evk/evk.cpp:277-293), unnormalised coordinate = uv * size, floor
inline vec4 texture(const sampler2D& t, vec2 uv) { return vxref_decode(t, (int)std::floor(uv.x * (float)t.w), (int)std::floor(uv.y * (float)t.h)); }
inline vec4 texture(const samplerCube& c, vec3) { return vec4(c.r, c.g, c.b, 1.0f); }

// ---- GLSL implicit conversions glm's templates do not perform (int -> float promotes first) -------------
namespace glm {
inline vec3 operator/(ivec3 const& a, float b) { return vec3(a) / b; }
inline vec3 operator/(vec3 const& a, int b) { return a / (float)b; }
inline vec2 operator+(ivec2 const& a, vec2 const& b) { return vec2(a) + b; }
inline vec2 operator/(float a, ivec2 const& b) { return a / vec2(b); }
inline float max(float a, int b) { return glm::max(a, (float)b); }
inline float mod(int a, int b) { return glm::mod((float)a, (float)b); }
inline float mod(float a, int b) { return glm::mod(a, (float)b); }
inline float step(uint edge, int x) { return glm::step((float)edge, (float)x); }
inline vec2 clamp(vec2 const& v, int lo, int hi) { return glm::clamp(v, (float)lo, (float)hi); }
inline float clamp(float v, int lo, int hi) { return glm::clamp(v, (float)lo, (float)hi); }
}  // namespace glm
// GLSL cos/sin are specified by accuracy only; the oracle pins them as the correctly rounded value
// (SURVEY hard part 1; glibc's cosf differs from it for 6 of the 256 possible theta = 6.283 * k/255).
inline float vxref_cos(float x) { return (float)std::cos((double)x); }
inline float vxref_sin(float x) { return (float)std::sin((double)x); }
#define cos(x) vxref_cos(x)
#define sin(x) vxref_sin(x)

// ---- keywords -----------------------------------------------------------------------------------------
#define layout(...)
#define uniform extern
#define readonly extern
#define buffer struct
#define discard vxref::g_discarded = true
// multi-component swizzles are member functions under GLM_FORCE_SWIZZLE
#define xyz xyz()
#define zxy zxy()
#define yzx yzx()
#define xy xy()
#define rgb rgb()
