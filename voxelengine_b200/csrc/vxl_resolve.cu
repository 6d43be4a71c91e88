// vxl_resolve.cu -- the colour the light passes add to the light buffer.
//
// What the reference's fragment shaders compute AFTER the shadow / AO march, as float32 RGBA before the
// RGBA16F attachment conversion and the additive blend:
//   LightAmbient.frag:134-214  out_Color = vec4(ambient + Lo, 0)   (+ calculateOcclusion :89-109, screenspaceOcclusion :54-79)
//   LightPoint.frag:131-152    out_Color = vec4(Lo, 0)  per light  (LightSpot.frag:118-138 with the cone term)
//   lib/PBR.frag:48-69         PBRDirectLight (the specular term is commented out in the reference)
// Elementwise over the lit pixels (HBM-bound: ~60 B per pixel), one thread per pixel, same thread -> pixel mapping
// as the march kernels.  Every operation is in the shaders' order with no contraction; pow() is powf, which GLSL
// specifies by accuracy only, so these planes are compared with a tolerance (the march planes stay bit-exact).
#include "vxl_internal.h"
#include "vxl_math.cuh"
#include "vxl_pixel.cuh"

#ifndef VXL_RESOLVE_BLOCKS
#define VXL_RESOLVE_BLOCKS 3         // resident 512-thread blocks per SM the resolve kernels' registers are capped for
#endif

namespace vxl {

constexpr float NEAR_ = 0.1f;                 // Common.frag:12
constexpr float PI_ = 3.14159265359f;         // PBR.frag:1

// pow(x, 5.0) for x >= 0 by four multiplications (within 2 ulp of the correctly rounded power; GLSL bounds pow only
// through exp2(y * log2(x)), tens of ulp): the generic powf is ~100 instructions, and the resolve evaluates it 3x per pixel
__device__ __forceinline__ float pow5(float x) { const float x2 = x * x; return (x2 * x2) * x; }
__device__ __forceinline__ float3 splat(float v) { return make_float3(v, v, v); }
__device__ __forceinline__ float3 unorm8x3(uint32_t c) { return make_float3(unorm8(c), unorm8(c >> 8), unorm8(c >> 16)); }
__device__ __forceinline__ float3 div3(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }

// PBR.frag:48-69
__device__ __forceinline__ float3 pbr_direct_light(float3 radiance, float3 albedo, float3 V, float3 N, float3 L, float metallic) {
    const float3 F0 = mix3(splat(0.04f), splat(1.0f), metallic);
    const float p5 = pow5(fmaxf(1.0f - fmaxf(dot3(N, V), 0.0f), 0.0f));
    const float3 F = F0 + (splat(1.0f) - F0) * p5;
    float3 kD = splat(1.0f) - F;
    kD = kD * (1.0f - metallic);
    const float NdotL = fmaxf(dot3(N, L), 0.0f);
    return (div3(kD * albedo, PI_) * radiance) * NdotL;
}

// whole-frame depth, nearest filter, out of range reads 0 (LightAmbient.frag:66 texture(DEPTH_TEXTURE, uv))
__device__ __forceinline__ float depth_at_uv(const uint32_t* __restrict__ depth_full, int W, int H, float u, float v) {
    const int x = (int)floorf(u * (float)W), y = (int)floorf(v * (float)H);
    if (x < 0 || y < 0 || x >= W || y >= H) return 0.0f;
    return unorm24(__ldg(depth_full + (size_t)y * W + x));
}

__device__ __forceinline__ float screenspace_occlusion(const ViewK& K, const uint32_t* __restrict__ depth_full, int W, int H, float3 pos, float3 dir, float dist) {
    const float3 mid = pos + dir * dist;
    const float4 mp = mat_mul(K.Proj, make_float4(mid.x, mid.y, mid.z, 1.0f));
    const float sampleDepth = (mp.w - NEAR_) / (FAR_ - NEAR_);
    const float u = ((mp.x / mp.w) * 1.0f) * 0.5f + 0.5f;
    const float v = ((mp.y / mp.w) * -1.0f) * 0.5f + 0.5f;
    const float minDepth = depth_at_uv(depth_full, W, H, u, v);
    const float maxDepth = minDepth + 0.2f / FAR_;
    if (gclamp(u, 0.0f, 1.0f) != u || gclamp(v, 0.0f, 1.0f) != v) return 0.0f;
    if (minDepth < sampleDepth && sampleDepth < maxDepth) return gsmoothstep(maxDepth, minDepth, sampleDepth) * dist;
    return 0.0f;
}

__global__ void __launch_bounds__(BLOCK_THREADS, VXL_RESOLVE_BLOCKS) k_resolve_ambient(FrameView F, ViewK K, const float* __restrict__ g_lut,
                                                                   const uint32_t* __restrict__ albedo, const uint32_t* __restrict__ depth_full,
                                                                   const float* __restrict__ shadow, const float* __restrict__ ao,
                                                                   float4* __restrict__ out) {
    __shared__ __align__(16) float s_lut[LUT_FLOATS];
    __shared__ float s_dec[512];             // unorm8[256], snorm8[256]: the decoders' own divisions, done once per block
    load_luts(s_lut, g_lut);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_dec[i] = unorm8((uint32_t)i); s_dec[256 + i] = snorm8((uint32_t)i); }
    __syncthreads();
    const PixelCtx p = pixel_ctx(F, K);
    if (!p.valid) return;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const float depth = unorm24(__ldg(F.depth24 + p.idx));
    if (depth < 0.999f) {                                                                     // :138
        const uint32_t ab = __ldg(albedo + p.idx);
        const float3 alb = make_float3(s_dec[ab & 0xFFu], s_dec[(ab >> 8) & 0xFFu], s_dec[(ab >> 16) & 0xFFu]);                                    // :139
        const uint32_t m = __ldg(F.material + p.idx);
        const float roughness = s_dec[m & 0xFFu], metallic = s_dec[(m >> 8) & 0xFFu], emit = s_dec[(m >> 16) & 0xFFu];  // :182-184
        const float3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));                          // :141
        const uint32_t nb = __ldg(F.normal + p.idx);
        const float3 normal = make_float3(s_dec[256 + (nb & 0xFFu)], s_dec[256 + ((nb >> 8) & 0xFFu)], s_dec[256 + ((nb >> 16) & 0xFFu)]);
        const float3 SUN = normalize3(make_float3(0.3f, 0.4f, 0.5f));
        const float3 sunDir = xyz(mat_mul(K.View, make_float4(SUN.x, SUN.y, SUN.z, 0.0f)));     // :179
        const float3 F0 = mix3(splat(0.04f), alb, metallic);                                   // :187-188
        const float3 Vv = normalize3(pos) * -1.0f;                                             // :190
        const float3 N = xyz(mat_mul(K.View, make_float4(normal.x, normal.y, normal.z, 0.0f)));   // :191
        const float3 SUN_COLOR = make_float3(0.9f, 0.9f, 0.8f) * 0.5f;                         // :16
        const float3 radiance = (SUN_COLOR * 1.0f) * __ldg(shadow + p.idx);                    // :198
        const float3 Lo = pbr_direct_light(radiance, alb, Vv, N, sunDir, metallic);            // :199
        const float c = fmaxf(dot3(N, Vv), 0.0f);
        const float3 rr = splat(1.0f - roughness);
        const float3 mx = make_float3(fmaxf(rr.x, F0.x), fmaxf(rr.y, F0.y), fmaxf(rr.z, F0.z));
        const float3 Fr = F0 + (mx - F0) * pow5(fmaxf(1.0f - c, 0.0f));                   // :204, PBR.frag:8-10
        const float3 kD = splat(1.0f) - Fr;                                                    // :206
        const float ey = (fmaxf(0.0f, 0.0f) * 0.8f + 0.2f) * 0.8f;                             // :128-131 getSkyColor(vec3(1,0,0))
        const float3 irradiance = make_float3((1.0f - ey) * (1.0f - ey), 1.0f - ey, 0.6f + (1.0f - ey) * 0.4f) * 1.1f;
        const float3 diffuse = irradiance * alb;                                               // :208
        // calculateOcclusion(N) :89-109
        const float3 tangent = normalize3(fabsf(N.z) > 0.5f ? make_float3(0.0f, -N.z, N.y) : make_float3(-N.y, N.x, 0.0f));
        const float3 bitangent = normalize3(cross3(N, tangent));
        const float3 opos = p.farvec * (depth * (1.0f + 1.0f / FAR_) + NEAR_ / FAR_);          // :94
        const float sizeMultiplier = 0.2f * (1.0f + depth * 0.0f);                             // :97
        float occlusion = 0.0f;
        for (int i = 0; i < 4; ++i) {
            const uint32_t n = get_noise(F, K, p, i);
            const float3 rv = cosine_sample_hemisphere(s_lut, n, n >> 8);
            const float3 dir = tangent * rv.x + bitangent * rv.y + N * rv.z;
            occlusion += screenspace_occlusion(K, depth_full, F.width, F.height, opos, normalize3(dir) * sizeMultiplier, s_dec[(n >> 16) & 0xFFu]);
        }
        occlusion /= 4.0f;
        occlusion *= 3.0f;
        const float occ = gclamp(1.0f - occlusion, 0.0f, 1.0f);
        const float3 ambient = diffuse * (splat(emit * 10.0f) + (kD * splat(__ldg(ao + p.idx))) * occ);   // :210
        const float3 c3 = ambient + Lo;                                                        // :213
        o = make_float4(c3.x, c3.y, c3.z, 0.0f);
    }
    out[p.idx] = o;
}

template <bool SPOT>
__global__ void __launch_bounds__(BLOCK_THREADS, VXL_RESOLVE_BLOCKS) k_resolve_local(FrameView F, ViewK K, const uint32_t* __restrict__ albedo,
                                                                 const float* __restrict__ lights, int n_lights,
                                                                 const float* __restrict__ shadow, size_t plane_stride,
                                                                 float4* __restrict__ inout) {
    constexpr int STRIDE = SPOT ? 16 : 8;
    __shared__ float s_light[VXL_MAX_LIGHTS * STRIDE];
    for (int i = threadIdx.x; i < n_lights * STRIDE; i += blockDim.x) s_light[i] = lights[i];
    __syncthreads();
    const PixelCtx p = pixel_ctx(F, K);
    if (!p.valid) return;
    const float depth = unorm24(__ldg(F.depth24 + p.idx));
    const float metallic = unorm8(__ldg(F.material + p.idx) >> 8);
    const float3 alb = unorm8x3(__ldg(albedo + p.idx));
    const float3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));
    const float3 normal = decode_normal(__ldg(F.normal + p.idx));
    const float3 worldPos = xyz(mat_mul(K.InvView, make_float4(pos.x, pos.y, pos.z, 1.0f)));
    const float3 Vv = normalize3(pos) * -1.0f;
    const float3 N = xyz(mat_mul(K.View, make_float4(normal.x, normal.y, normal.z, 0.0f)));
    float4 o = inout[p.idx];
    for (int li = 0; li < n_lights; ++li) {
        const float* lt = s_light + li * STRIDE;
        const float3 lpos = make_float3(lt[0], lt[1], lt[2]);
        const float range = lt[3];
        const float3 color = make_float3(lt[4], lt[5], lt[6]);
        const float atten = lt[7];
        const float3 lightPos = xyz(mat_mul(K.View, make_float4(lpos.x, lpos.y, lpos.z, 1.0f)));
        const float3 lightDir = lpos - worldPos;
        const float lightDistance = length3(lightDir);
        if (lightDistance > range) continue;                                                   // discard
        const float3 Lv = xyz(mat_mul(K.View, make_float4(lightDir.x, lightDir.y, lightDir.z, 0.0f)));
        const float dist = length3(lightPos - pos);
        float attenuation;
        if (!SPOT) attenuation = gclamp(range - lightDistance, 0.0f, 1.0f) / powf(dist, atten);
        else {
            const float3 sdir = make_float3(lt[8], lt[9], lt[10]);
            const float angle = lt[11], angleAtten = lt[12];
            const float angleDist = fmaxf(dot3(normalize3(lightDir), sdir) - (1.0f - angle), 0.0f) / angle;
            attenuation = (powf(angleDist, angleAtten) * gclamp(range - lightDistance, 0.0f, 1.0f)) / powf(dist, atten);
        }
        const float3 radiance = (color * attenuation) * __ldg(shadow + (size_t)li * plane_stride + p.idx);
        const float3 Lo = pbr_direct_light(radiance, alb, Vv, N, Lv, metallic);
        o.x += Lo.x; o.y += Lo.y; o.z += Lo.z; o.w += 0.0f;
    }
    inout[p.idx] = o;
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_resolve_ambient(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_resolve* r,
                        const float* shadow, const float* ao, float* out_rgba) {
    if (!ctx || !view || !frame || !r || !r->albedo || !shadow || !ao || !out_rgba) { set_error("vxl_resolve_ambient: bad argument"); return VXL_ERR_INVALID; }
    FrameView F;
    if (int e = frame_view(frame, &F)) return e;
    if (!F.material) { set_error("vxl_resolve_ambient: frame.material is NULL"); return VXL_ERR_INVALID; }
    const bool whole = F.n_tiles == 1 && F.tile_w == F.width && F.tile_h == F.height && F.tile_first == 0;
    if (!r->depth_full && !whole) { set_error("vxl_resolve_ambient: a tile-sharded frame needs vxl_resolve.depth_full (the screen-space occlusion samples other pixels)"); return VXL_ERR_INVALID; }
    if (F.n_tiles == 0) return VXL_OK;
    k_resolve_ambient<<<grid_for(F), BLOCK_THREADS, 0, ctx->stream>>>(F, make_viewk(view), ctx->d_luts, r->albedo, r->depth_full ? r->depth_full : F.depth24,
                                                                      shadow, ao, (float4*)out_rgba);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

static int resolve_local(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_resolve* r, const void* lights, int n_lights,
                         size_t light_bytes, bool spot, const float* shadow, float* inout_rgba) {
    if (!ctx || !view || !frame || !r || !r->albedo || n_lights < 0 || (n_lights > 0 && (!lights || !shadow)) || !inout_rgba) {
        set_error("vxl_resolve_point/spot: bad argument");
        return VXL_ERR_INVALID;
    }
    if (n_lights > VXL_MAX_LIGHTS) { set_error("more than VXL_MAX_LIGHTS lights"); return VXL_ERR_LIMIT; }
    FrameView F;
    if (int e = frame_view(frame, &F)) return e;
    if (!F.material) { set_error("vxl_resolve_point/spot: frame.material is NULL"); return VXL_ERR_INVALID; }
    if (n_lights == 0 || F.n_tiles == 0) return VXL_OK;
    VXL_CUDA(cudaMemcpyAsync(ctx->d_lights, lights, (size_t)n_lights * light_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const size_t plane = frame_pixels(frame);
    if (spot) k_resolve_local<true><<<grid_for(F), BLOCK_THREADS, 0, ctx->stream>>>(F, make_viewk(view), r->albedo, (const float*)ctx->d_lights, n_lights, shadow, plane, (float4*)inout_rgba);
    else k_resolve_local<false><<<grid_for(F), BLOCK_THREADS, 0, ctx->stream>>>(F, make_viewk(view), r->albedo, (const float*)ctx->d_lights, n_lights, shadow, plane, (float4*)inout_rgba);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

int vxl_resolve_point(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_resolve* r,
                      const vxl_point_light* lights, int n_lights, const float* shadow, float* inout_rgba) {
    return resolve_local(ctx, view, frame, r, lights, n_lights, sizeof(vxl_point_light), false, shadow, inout_rgba);
}

int vxl_resolve_spot(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_resolve* r,
                     const vxl_spot_light* lights, int n_lights, const float* shadow, float* inout_rgba) {
    return resolve_local(ctx, view, frame, r, lights, n_lights, sizeof(vxl_spot_light), true, shadow, inout_rgba);
}

}  // extern "C"
