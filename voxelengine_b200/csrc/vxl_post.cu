// vxl_post.cu -- the steps after the light passes.
//
//   k_light_taa           Sources/Shaders/LightTAA.frag:37-141: temporal + spatial accumulation of the light buffer
//   k_resolve_reflection  Sources/Shaders/LightReflection.frag:60-139: the colour around the specular-occlusion march
//
// Both sample OTHER pixels (a 12-tap golden-angle spiral or a 19x19 window; the reflected ray's end point), so their
// inputs are whole-frame row-major planes (vxl_full_planes); the output pixels and their tile-compact layout come from
// the vxl_frame like in every other pass.  Sampling is the nearest filter of the reference's samplers
// (Vendor/evk/evk.cpp:277-293): texel = floor(uv * size), out of range reads 0.  Light / motion planes are float32 (the
// values before the RGBA16F / RG16F attachment conversion, as in vxl_resolve.cu).
//
// LightTAA's only transcendentals are cos / sin of the spiral angle, which takes 256 x 12 values (the noise byte and the
// tap index; the angle itself is a float recurrence): they are tabulated on the host in double and rounded once, which is
// this project's pinned definition of GLSL's cos / sin.  Everything else is IEEE arithmetic in the shader's order (--fmad=false),
// glm's ternary min / max / clamp included (they differ from fminf / fmaxf on NaN, which a zero weight sum produces), so
// the TAA plane is bit-exact against the CPU restatement of the shader.  The reflection colour carries pow() and is compared with a tolerance.
#include "vxl_internal.h"
#include "vxl_math.cuh"
#include "vxl_pixel.cuh"

#include <cmath>

namespace vxl {

constexpr float NEAR_P = 0.1f;                // Common.frag:12
#ifndef VXL_TAA_BLOCKS
#define VXL_TAA_BLOCKS 2
#endif
#ifndef VXL_TAA_UNROLL
#define VXL_TAA_UNROLL 3
#endif
constexpr int TAA_UNROLL = VXL_TAA_UNROLL;   // taps in flight per thread (more: spills at 64 registers)
constexpr int TAA_TAPS = 12;                  // LightTAA.frag:96: radius runs 2..13 while radius <= size (12)

struct FullView {
    const float* __restrict__ depthf;         // unorm24 of the depth plane, decoded once per frame by k_decode_depth (13 taps per pixel read it)
    const uint32_t* __restrict__ depth24; const uint32_t* __restrict__ normal; const uint32_t* __restrict__ material; const uint32_t* __restrict__ albedo;
    const float2* __restrict__ motion; const float4* __restrict__ light; const float4* __restrict__ last_light;
};

__device__ __forceinline__ float tmin(float a, float b) { return (b < a) ? b : a; }          // glm::min (func_common.inl)
__device__ __forceinline__ float tmax(float a, float b) { return (a < b) ? b : a; }          // glm::max
__device__ __forceinline__ float tclamp(float x, float lo, float hi) { return tmin(tmax(x, lo), hi); }
__device__ __forceinline__ int tex_index(int W, int H, float u, float v) {
    const int x = (int)floorf(u * (float)W), y = (int)floorf(v * (float)H);
    if (x < 0 || y < 0 || x >= W || y >= H) return -1;
    return y * W + x;
}
__device__ __forceinline__ float3 splat3(float v) { return make_float3(v, v, v); }

struct TaaTexel { float3 color, normal, light; float4 material; float depth, mx, my, la; };
// lut: unorm8[256] then snorm8[256], filled per block by the same divisions the decoders perform (13 texels x 10 channels per
// pixel would otherwise be 130 IEEE divisions)
__device__ __forceinline__ TaaTexel taa_fetch(const FullView& P, const float* __restrict__ lut, int W, int H, float u, float v) {
    TaaTexel t;
    const int i = tex_index(W, H, u, v);
    if (i < 0) {
        t.color = t.normal = t.light = splat3(0.0f); t.material = make_float4(0.f, 0.f, 0.f, 0.f); t.depth = t.mx = t.my = t.la = 0.0f;
        return t;
    }
    const uint32_t a = __ldg(P.albedo + i), m = __ldg(P.material + i);
    const uint32_t nn = __ldg(P.normal + i);
    t.color = make_float3(lut[a & 0xFFu], lut[(a >> 8) & 0xFFu], lut[(a >> 16) & 0xFFu]);
    t.normal = make_float3(lut[256 + (nn & 0xFFu)], lut[256 + ((nn >> 8) & 0xFFu)], lut[256 + ((nn >> 16) & 0xFFu)]);
    t.material = make_float4(lut[m & 0xFFu], lut[(m >> 8) & 0xFFu], lut[(m >> 16) & 0xFFu], lut[m >> 24]);
    t.depth = __ldg(P.depthf + i);
    const float2 mo = __ldg(P.motion + i);
    t.mx = mo.x; t.my = mo.y;
    const float4 l = __ldg(P.light + i);
    t.light = make_float3(l.x, l.y, l.z); t.la = l.w;
    return t;
}
__device__ __forceinline__ float material_distance(float4 a, float4 b) {                     // length(vec4): glm compute_dot<vec4>
    const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
    return sqrtf((x * x + y * y) + (z * z + w * w));
}
__device__ __forceinline__ float length2(float x, float y) { return sqrtf(x * x + y * y); }

// The shader compares three square roots with constants.  sqrtf is correctly rounded, hence monotone, and so are the float
// operations applied to it, so each comparison is a comparison of the radicand with ONE float found by bisection over the bit
// patterns (tools/exp/taa_thresholds.py prints them and checks the neighbours):
//   length(a) > 0.1            <=>  dot(a, a) > T_MOTION           (LightTAA.frag:76, :113)
//   step(0.8, 1.0 - length(a)) == 0  <=>  dot(a, a) > T_MATERIAL   (:72, :109)
//   clamp(length(a) * 10000.0, 0.0, 1.0) == 1.0  <=>  dot(a, a) >= T_COLOR   (:112)
// A NaN radicand fails every comparison here exactly like the NaN root fails the shader's.
constexpr float T_MOTION = 0x1.47ae16p-7f, T_MATERIAL = 0x1.47ae16p-5f, T_COLOR = 0x1.5798ecp-27f;
__device__ __forceinline__ float material_dot(float4 a, float4 b) {
    const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
    return (x * x + y * y) + (z * z + w * w);
}
// 1.212 - radius / size for radius = 2 .. 13, size = 12 (:108): the quotients are constants of the loop
__constant__ float c_taa_rs[TAA_TAPS] = {2.0f / 12.0f, 3.0f / 12.0f, 4.0f / 12.0f, 5.0f / 12.0f, 6.0f / 12.0f, 7.0f / 12.0f,
                                         8.0f / 12.0f, 9.0f / 12.0f, 10.0f / 12.0f, 11.0f / 12.0f, 12.0f / 12.0f, 13.0f / 12.0f};

__global__ void __launch_bounds__(256) k_decode_depth(const uint32_t* __restrict__ d24, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = unorm24(__ldg(d24 + i));
}

__global__ void __launch_bounds__(BLOCK_THREADS, VXL_TAA_BLOCKS) k_light_taa(FrameView F, ViewK K, FullView P, const float2* __restrict__ g_cs /* [256][12] */,
                                                             float4* __restrict__ out) {
    __shared__ float lut[512];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { lut[i] = unorm8((uint32_t)i); lut[256 + i] = snorm8((uint32_t)i); }
    __syncthreads();
    const PixelCtx p = pixel_ctx(F, K);
    if (!p.valid) return;
    const int W = F.width, H = F.height;
    const float iRx = 1.0f / (float)W, iRy = 1.0f / (float)H;                                 // :38
    const TaaTexel c = taa_fetch(P, lut, W, H, p.u, p.v);                                     // :40-47
    if (c.depth == 1.0f) { out[p.idx] = make_float4(c.light.x, c.light.y, c.light.z, c.la); return; }   // :49-52
    const float oldU = p.u + c.mx, oldV = p.v + c.my;                                         // :41
    const int li = tex_index(W, H, oldU, oldV);
    float3 lastLight = splat3(0.0f);
    float lastVariance = 0.0f;
    if (li >= 0) { const float4 l = __ldg(P.last_light + li); lastLight = make_float3(l.x, l.y, l.z); lastVariance = l.w; }   // :53-54
    if (tclamp(oldU, 0.0f, 1.0f) != oldU || tclamp(oldV, 0.0f, 1.0f) != oldV) {               // :57 history outside the frame
        float count = 0.0f;
        float3 neigh = splat3(0.0f);
        for (int x = -9; x <= 9; ++x)
            for (int y = -9; y <= 9; ++y) {
                const float u = tclamp(p.u + (float)x * iRx, 0.001f, 0.999f), v = tclamp(p.v + (float)y * iRy, 0.001f, 0.999f);   // :62-63
                const TaaTexel n = taa_fetch(P, lut, W, H, u, v);
                float factor = tmax(dot3(c.normal, n.normal), 0.0f);                          // :71
                factor *= material_dot(c.material, n.material) > T_MATERIAL ? 0.0f : 1.0f;     // :72
                factor *= 1.0f - tclamp(fabsf(c.depth - n.depth) * FAR_, 0.0f, 1.0f);          // :73
                factor *= 1.0f - tclamp(length3(c.color - n.color), 0.0f, 1.0f);              // :74
                { const float dx = c.mx - n.mx, dy = c.my - n.my; if (dx * dx + dy * dy > T_MOTION) factor = 0.0f; }   // :76
                neigh = neigh + n.light * factor;                                             // :79
                count += factor;
            }
        out[p.idx] = make_float4(neigh.x / count, neigh.y / count, neigh.z / count, 1.0f);    // :83
        return;
    }
    float3 nmin = splat3(10000.0f), nmax = splat3(0.0f), neigh = splat3(0.0f);                // :89-91
    float diffSum = 1.0f, count = 0.0f, radius = 1.0f;
    const uint32_t nz = get_noise(F, K, p, -1);
    const float2* cs = g_cs + (nz & 0xFFu) * TAA_TAPS;
    const float k2 = tclamp(0.1f, 0.5f, 1.0f / c.depth);                                      // :99 (sic: clamp(x = 0.1, 0.5, 1/depth))
#pragma unroll (TAA_UNROLL)
    for (int k = 0; k < TAA_TAPS; ++k) {                                                      // :96 radius <= size
        radius += 1.0f;
        const float2 a = __ldg(cs + k);
        const float k1 = radius * (lastVariance + 1.0f);
        const float ox = ((a.x * iRx) * k1) * k2, oy = ((a.y * iRy) * k1) * k2;
        const float u = tclamp(p.u + ox, 0.001f, 0.999f), v = tclamp(p.v + oy, 0.001f, 0.999f);
        const TaaTexel n = taa_fetch(P, lut, W, H, u, v);
        float factor = 1.212f - c_taa_rs[k];                                                  // :108
        factor *= material_dot(c.material, n.material) > T_MATERIAL ? 0.0f : 1.0f;
        factor *= tmax(dot3(c.normal, n.normal), 0.0f);
        factor *= 1.0f - tclamp(fabsf(c.depth - n.depth) * FAR_, 0.0f, 1.0f);
        {
            const float3 dc = c.color - n.color;
            const float s2 = dot3(dc, dc);
            float cl = 1.0f;                                                                  // albedo bytes differ by >= 1/255: always here
            if (s2 == 0.0f) cl = 0.0f;
            else if (!(s2 >= T_COLOR)) cl = tclamp(sqrtf(s2) * 10000.0f, 0.0f, 1.0f);
            factor *= 1.0f - cl;
        }
        { const float dx = c.mx - n.mx, dy = c.my - n.my; if (dx * dx + dy * dy > T_MOTION) factor = 0.0f; }
        nmin = make_float3(tmin(nmin.x, n.light.x), tmin(nmin.y, n.light.y), tmin(nmin.z, n.light.z));   // :116
        nmax = make_float3(tmax(nmax.x, n.light.x), tmax(nmax.y, n.light.y), tmax(nmax.z, n.light.z));
        neigh = neigh + n.light * factor;
        count += factor;
        diffSum += length3(n.light - c.light) * factor;                                       // :121
    }
    float3 cur = c.light + neigh;                                                             // :123
    const float cd = count + 1.0f;
    cur = make_float3(cur.x / cd, cur.y / cd, cur.z / cd);                                    // :124
    diffSum /= radius;                                                                        // :125
    lastLight = make_float3(tclamp(lastLight.x, nmin.x, nmax.x), tclamp(lastLight.y, nmin.y, nmax.y), tclamp(lastLight.z, nmin.z, nmax.z));   // :129
    const float3 ad = make_float3(fabsf(cur.x - lastLight.x), fabsf(cur.y - lastLight.y), fabsf(cur.z - lastLight.z));
    float variance = dot3(ad, make_float3(0.2125f, 0.7154f, 0.0721f));                        // :132, :30-35
    variance = tclamp(variance * 5.5f, 0.0f, 1.0f);
    lastVariance += variance;
    lastVariance -= diffSum * 0.08f;
    lastVariance += length2(c.mx, c.my) * 20.0f;                                              // :136
    const float3 m = mix3(cur, lastLight, tclamp(1.0f - lastVariance, 0.3f, 0.9f));           // :141
    out[p.idx] = make_float4(m.x, m.y, m.z, tclamp(lastVariance * 0.7f, 0.0f, 1.0f));
}

__device__ __forceinline__ float pow5p(float x) { const float x2 = x * x; return (x2 * x2) * x; }   // pow(x, 5.0), x >= 0 (see vxl_resolve.cu)

// spec_t: the plane of vxl_pass_reflection (tile-compact).  depth_full / light_full: whole-frame row-major planes.
__global__ void __launch_bounds__(BLOCK_THREADS) k_resolve_reflection(FrameView F, ViewK K, const float* __restrict__ g_lut, const float* __restrict__ spec_t,
                                                                      const uint32_t* __restrict__ depth_full, const float4* __restrict__ light_full,
                                                                      float3 sky, float4* __restrict__ out) {
    __shared__ __align__(16) float s_lut[LUT_FLOATS];
    load_luts(s_lut, g_lut);
    __syncthreads();
    const PixelCtx p = pixel_ctx(F, K);
    if (!p.valid) return;
    const uint32_t mt = __ldg(F.material + p.idx);
    const float roughness = unorm8(mt), metallic = unorm8(mt >> 8);                           // :68-69
    const float depth = unorm24(__ldg(F.depth24 + p.idx));
    const float3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));                             // :64
    const float3 normal = decode_normal(__ldg(F.normal + p.idx));
    const float3 F0 = mix3(splat3(0.04f), splat3(1.0f), metallic);                            // :74-77
    const float3 V = normalize3(pos) * -1.0f;                                                 // :79
    const float3 N = xyz(mat_mul(K.View, make_float4(normal.x, normal.y, normal.z, 0.0f)));   // :80
    const float3 I = V * -1.0f;
    const float3 R = I - N * dot3(N, I) * 2.0f;                                               // :81
    float3 ambient = splat3(0.0f);
    const float cth = fmaxf(dot3(N, V), 0.0f);
    const float3 rr = splat3(1.0f - roughness);
    const float3 mx = make_float3(fmaxf(rr.x, F0.x), fmaxf(rr.y, F0.y), fmaxf(rr.z, F0.z));
    const float3 Fr = (F0 + (mx - F0) * pow5p(fmaxf(1.0f - cth, 0.0f))) * 5.0f + splat3(0.0f);   // :86
    if (depth < 0.999f) {
        float3 wd = normalize3(xyz(mat_mul(K.InvView, make_float4(R.x, R.y, R.z, 0.0f))));    // :92
        float3 wcp = xyz(mat_mul(K.InvView, make_float4(pos.x, pos.y, pos.z, 1.0f))) * 10.0f; // :93
        const uint32_t n = get_noise(F, K, p, -1);
        float3 rv = cosine_sample_hemisphere(s_lut, n, n >> 8);                               // :94
        rv.z *= gsign(unorm8(n >> 16) - 0.5f);                                                // :95
        wd = mix3(wd, rv, roughness * 0.1f);                                                  // :96
        const float nw = unorm8(n >> 24);
        wcp = wcp + normal * nw;                                                              // :97
        wd = wd * (1.0f + nw * 0.5f);                                                         // :98
        const float t = __ldg(spec_t + p.idx);                                                // :113
        const float3 e = (wcp + wd * t) * 0.1f;
        const float3 hp = xyz(mat_mul(K.View, make_float4(e.x, e.y, e.z, 1.0f)));             // :115
        const float4 pp = mat_mul(K.Proj, make_float4(hp.x, hp.y, hp.z, 1.0f));               // :116
        const float linearDepth = (pp.w - NEAR_P) / (FAR_ - NEAR_P);                          // :117
        const float u = ((pp.x / pp.w) * 1.0f) * 0.5f + 0.5f, v = ((pp.y / pp.w) * -1.0f) * 0.5f + 0.5f;   // :118
        const int ti = tex_index(F.width, F.height, u, v);
        const float d2 = ti < 0 ? 0.0f : unorm24(__ldg(depth_full + ti));                     // :119
        if (t == 256.0f) ambient = ambient + sky;                                             // :121-122
        else if (linearDepth > d2 - 0.001f) {                                                 // :124
            if (linearDepth < d2 + 0.001f && ti >= 0 && light_full) {                         // :125-126
                const float4 l = __ldg(light_full + ti);
                ambient = ambient + make_float3(l.x, l.y, l.z);
            }
        }
    }
    const float3 c3 = (ambient * Fr) * (1.0f - roughness);                                    // :139
    out[p.idx] = make_float4(c3.x, c3.y, c3.z, Fr.x);
}

}  // namespace vxl

using namespace vxl;

namespace {
// cos / sin of the 256 x 12 spiral angles of LightTAA.frag:96 (angle_0 = noise.x * 3.1415 * GOLDEN_RATIO, += 2.39 per tap, in
// float like the shader), evaluated in double and rounded once
int ensure_taa_lut(vxl_ctx* ctx) {
    if (ctx->d_taa_lut) return VXL_OK;
    std::vector<float> lut(256 * TAA_TAPS * 2);
    for (int b = 0; b < 256; ++b) {
        volatile float angle = ((float)b / 255.0f) * 3.1415f;
        angle = angle * GOLDEN_RATIO;
        for (int k = 0; k < TAA_TAPS; ++k) {
            const float a = angle;
            lut[(b * TAA_TAPS + k) * 2] = (float)std::cos((double)a);
            lut[(b * TAA_TAPS + k) * 2 + 1] = (float)std::sin((double)a);
            angle = a + 2.39f;
        }
    }
    VXL_CUDA(cudaMalloc((void**)&ctx->d_taa_lut, lut.size() * sizeof(float)));
    VXL_CUDA(cudaMemcpy(ctx->d_taa_lut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice));
    return VXL_OK;
}
bool whole_frame(const FrameView& F) { return F.n_tiles == 1 && F.tile_w == F.width && F.tile_h == F.height && F.tile_first == 0; }
}  // namespace

extern "C" {

int vxl_light_taa(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_full_planes* full, float* out_rgba) {
    if (!ctx || !view || !frame || !full || !out_rgba || !full->depth24 || !full->normal || !full->material || !full->albedo || !full->motion ||
        !full->light || !full->last_light) { set_error("vxl_light_taa: bad argument (every plane of vxl_full_planes is required)"); return VXL_ERR_INVALID; }
    FrameView F;
    if (int e = frame_view(frame, &F)) return e;
    if (F.n_tiles == 0) return VXL_OK;
    VXL_CUDA(cudaSetDevice(ctx->device));
    if (int e = ensure_taa_lut(ctx)) return e;
    const size_t npx = (size_t)F.width * (size_t)F.height;
    if (ctx->taa_depth_cap < npx) {
        if (ctx->d_taa_depth) cudaFree(ctx->d_taa_depth);
        ctx->d_taa_depth = nullptr; ctx->taa_depth_cap = 0;
        VXL_CUDA(cudaMalloc((void**)&ctx->d_taa_depth, npx * sizeof(float)));
        ctx->taa_depth_cap = npx;
    }
    k_decode_depth<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(full->depth24, ctx->d_taa_depth, npx);
    ctx->launches++;
    FullView P{ctx->d_taa_depth, full->depth24, full->normal, full->material, full->albedo, (const float2*)full->motion, (const float4*)full->light, (const float4*)full->last_light};
    k_light_taa<<<grid_for(F), BLOCK_THREADS, 0, ctx->stream>>>(F, make_viewk(view), P, (const float2*)ctx->d_taa_lut, (float4*)out_rgba);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

int vxl_resolve_reflection(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const float* spec_t, const uint32_t* depth_full,
                           const float* light_full, const float* sky_rgb, float* out_rgba) {
    if (!ctx || !view || !frame || !spec_t || !out_rgba) { set_error("vxl_resolve_reflection: bad argument"); return VXL_ERR_INVALID; }
    FrameView F;
    if (int e = frame_view(frame, &F)) return e;
    if (!F.material) { set_error("vxl_resolve_reflection: frame.material is NULL"); return VXL_ERR_INVALID; }
    if (!depth_full && !whole_frame(F)) { set_error("vxl_resolve_reflection: a tile-sharded frame needs depth_full (the reflected ray's end point is another pixel)"); return VXL_ERR_INVALID; }
    if (F.n_tiles == 0) return VXL_OK;
    VXL_CUDA(cudaSetDevice(ctx->device));
    const float3 sky = sky_rgb ? make_float3(sky_rgb[0], sky_rgb[1], sky_rgb[2]) : make_float3(0.f, 0.f, 0.f);
    k_resolve_reflection<<<grid_for(F), BLOCK_THREADS, 0, ctx->stream>>>(F, make_viewk(view), ctx->d_luts, spec_t, depth_full ? depth_full : F.depth24,
                                                                         (const float4*)light_full, sky, (float4*)out_rgba);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

}  // extern "C"
