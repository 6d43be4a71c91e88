// vxl_model.cu -- the G-buffer producer's traversal of one model volume (core).
//
//   VoxAsset::Upload mip rule     Sources/Asset/VoxAsset.cpp:16-53   -> k_model_mip
//   clipToAABB                    Sources/Shaders/GeometryVoxel.frag:49-61
//   intersectVolume               GeometryVoxel.frag:64-125           -> k_trace_model_rays
// The reference's hierarchical-mip DDA: start at mip 2, descend where a coarse voxel is hit (unless the LOD rule
// accepts it at that level), three-level R8 volume, glass (palette index < 16) dithered by a frame-alternating
// checkerboard.  One thread per ray; every operation in the shader's order, no contraction, IEEE division, so the
// records are bit-identical to the reference's own function run on the host (tests/test_model_traversal.py).
#include "vxl_internal.h"
#include "vxl_math.cuh"

#include <vector>

namespace vxl {

// one mip level from its parent: first non-zero of the 2x2x2 children, x fastest (the reference's 9th iteration revisits child 0)
__global__ void __launch_bounds__(256) k_model_mip(const uint8_t* __restrict__ parent, int psx, int psy, int psz, int sx, int sy, int sz,
                                                   uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)sx * sy * sz) return;
    const int x = (int)(i % sx), y = (int)((i / sx) % sy), z = (int)(i / ((long long)sx * sy));
    uint8_t vox = 0;
    for (int vi = 0; vi < 8 && !vox; ++vi) {
        const int cx = 2 * x + (vi & 1), cy = 2 * y + ((vi >> 1) & 1), cz = 2 * z + ((vi >> 2) & 1);
        if (cx >= psx || cy >= psy || cz >= psz) continue;
        vox = parent[(size_t)cx + (size_t)cy * psx + (size_t)cz * psx * psy];
    }
    out[i] = vox;
}

struct ModelMips { const uint8_t* d[3]; int sx[3], sy[3], sz[3]; };

__device__ __forceinline__ unsigned model_fetch(const ModelMips& M, int x, int y, int z, int mip) {
    if (x < 0 || y < 0 || z < 0 || x >= M.sx[mip] || y >= M.sy[mip] || z >= M.sz[mip]) return 0u;
    return (unsigned)__ldg(M.d[mip] + (size_t)x + (size_t)y * M.sx[mip] + (size_t)z * M.sx[mip] * M.sy[mip]);
}

// clipToAABB (:49-61) + intersectVolume (:64-125) for one fragment: cam = In.localCameraPos, dir = In.localDirection, (u, v) = UV
__device__ __forceinline__ vxl_model_hit traverse_model(const ModelMips& M, float3 cam, float3 dir, float u, float v, int frame, float res_x, float res_y) {
    vxl_model_hit h;
    h.hit = 0; h.material = 0u; h.fetches = 0; h.steps = 0;
    h.pos[0] = h.pos[1] = h.pos[2] = 0.f; h.normal[0] = h.normal[1] = h.normal[2] = 0.f;
    const float3 vsize = make_float3((float)M.sx[0], (float)M.sy[0], (float)M.sz[0]);
    const float3 direction = normalize3(dir);                                                    // GeometryVoxel.frag:145
    // clipToAABB :49-61
    float3 origin = cam;
    if (!(gclamp(cam.x, 0.0f, vsize.x) == cam.x && gclamp(cam.y, 0.0f, vsize.y) == cam.y && gclamp(cam.z, 0.0f, vsize.z) == cam.z)) {
        const float3 invDir = make_float3(1.0f, 1.0f, 1.0f) / direction;
        const float3 sgn = make_float3(gstep(direction.x, 0.0f), gstep(direction.y, 0.0f), gstep(direction.z, 0.0f));
        const float3 t = (sgn * vsize - cam) * invDir;
        const float tmin = fmaxf(fmaxf(t.x, t.y), t.z);
        origin = cam + direction * (tmin - 0.001f);
    }
    // intersectVolume :64-125
    const float3 stepSign = make_float3(gsign(direction.x), gsign(direction.y), gsign(direction.z));
    const float3 t_delta = make_float3(1.0f, 1.0f, 1.0f) / (direction * stepSign);
    int mip = 2, i = 0, nt = 0, fetches = 0;
    bool done = false;
    do {
        const float mipSize = (float)(1 << mip);
        origin = make_float3(origin.x / mipSize, origin.y / mipSize, origin.z / mipSize);
        int cx = f2i(floorf(origin.x)), cy = f2i(floorf(origin.y)), cz = f2i(floorf(origin.z));
        const float3 next_bounds = make_float3((float)cx, (float)cy, (float)cz) + (stepSign * 0.5f + make_float3(0.5f, 0.5f, 0.5f));
        float3 t_max = (next_bounds - origin) / direction;
        const int hx = f2i((float)M.sx[0] / mipSize) + 1, hy = f2i((float)M.sy[0] / mipSize) + 1, hz = f2i((float)M.sz[0] / mipSize) + 1;
        int nn = 0;
        do {
            const float3 select = make_float3(gstep(t_max.x, t_max.z) * gstep(t_max.x, t_max.y), gstep(t_max.y, t_max.x) * gstep(t_max.y, t_max.z),
                                              gstep(t_max.z, t_max.y) * gstep(t_max.z, t_max.x));
            const float3 adv = select * stepSign;
            cx += f2i(adv.x); cy += f2i(adv.y); cz += f2i(adv.z);
            if (cx < -1 || cy < -1 || cz < -1 || cx > hx || cy > hy || cz > hz) { done = true; break; }     // :84-86
            const unsigned voxel = model_fetch(M, cx, cy, cz, mip);
            ++fetches;
            if (voxel != 0u) {
                const float best_t = dot3(t_max, select);
                const float3 at = (origin + direction * best_t) * mipSize;
                if (mip == 0 || (float)mip < 0.001f * length3(at - cam)) {                       // :92-96 LOD early accept
                    const float cxr = roundf(u * res_x * 0.5f), cyr = roundf(v * res_y * 0.5f);   // :99
                    const bool glass = voxel < 16u && gmod(cyr + cxr, 2.0f) == (float)(frame % 2);   // :100
                    if (!glass) {
                        h.hit = 1; h.material = voxel;
                        const float3 nrm = (stepSign * -1.0f) * select;
                        h.normal[0] = nrm.x; h.normal[1] = nrm.y; h.normal[2] = nrm.z;
                        h.pos[0] = at.x; h.pos[1] = at.y; h.pos[2] = at.z;
                        done = true;
                        break;
                    }
                } else {
                    mip--;                                                                       // :110-112
                    origin = origin + make_float3((direction.x * best_t) / mipSize, (direction.y * best_t) / mipSize, (direction.z * best_t) / mipSize);
                    break;
                }
            }
            t_max = t_max + t_delta * select;
            nt++;
        } while (++nn < 512);
        if (done) break;
        origin = origin * mipSize;                                                               // :121
    } while (++i < 4);
    h.fetches = fetches; h.steps = nt;
    return h;
}

__global__ void __launch_bounds__(256) k_trace_model_rays(ModelMips M, const vxl_model_ray* __restrict__ rays, long long n, int frame,
                                                          float res_x, float res_y, vxl_model_hit* __restrict__ out) {
    const long long ri = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= n) return;
    const vxl_model_ray r = rays[ri];
    out[ri] = traverse_model(M, make_float3(r.cam[0], r.cam[1], r.cam[2]), make_float3(r.dir[0], r.dir[1], r.dir[2]), r.uv[0], r.uv[1], frame, res_x, res_y);
}

// ---- the geometry pass over a draw list (GeometryVoxelPipeline::Use, Pipelines/GeometryVoxelPipeline.h:49-71) ----------------
// Per pixel, in list order: every model whose box the pixel's view ray enters from outside (the pipeline culls back faces,
// evk.cpp:470) runs GeometryVoxel.frag's main() (:127-182); depth test LESS on the D24 value (evk.cpp:481); outputs converted
// to the attachment formats (Graphics.h:51-60).  The fragment stage's interpolated inputs are evaluated per pixel centre
// (In.localDirection = the pixel's view ray in model space), the declared definition that In.FarVec has in the light passes.
struct DrawDev {                     // per draw, derived on the host like GeometryVoxel.vert:62-66 does per vertex
    float inv[16], mvp[16], world[16], last_world[16];
    float cam[3], size[3];
    int rid, palette;
    ModelMips M;
};
constexpr int GEOM_MAX_DRAWS = 4096;   // GeometryVoxelPipeline MAX_INSTANCES
#ifndef VXL_GEOM_BLOCKS
#define VXL_GEOM_BLOCKS 4            // resident 256-thread blocks per SM k_gbuffer_models' registers are capped for
#endif
struct GeomK { float InvView[16], InvProj[16], PV[16], PVlast[16]; float cam[3], jit[2], ires[2], res[2]; int frame; };

__device__ __forceinline__ uint32_t pack_unorm8x4(float a, float b, float c, float d) {
    return (uint32_t)rintf(gclamp(a, 0.0f, 1.0f) * 255.0f) | ((uint32_t)rintf(gclamp(b, 0.0f, 1.0f) * 255.0f) << 8) |
           ((uint32_t)rintf(gclamp(c, 0.0f, 1.0f) * 255.0f) << 16) | ((uint32_t)rintf(gclamp(d, 0.0f, 1.0f) * 255.0f) << 24);
}
__device__ __forceinline__ uint32_t pack_snorm8x4(float a, float b, float c, float d) {
    return ((uint32_t)(int)rintf(gclamp(a, -1.0f, 1.0f) * 127.0f) & 0xFFu) | (((uint32_t)(int)rintf(gclamp(b, -1.0f, 1.0f) * 127.0f) & 0xFFu) << 8) |
           (((uint32_t)(int)rintf(gclamp(c, -1.0f, 1.0f) * 127.0f) & 0xFFu) << 16) | (((uint32_t)(int)rintf(gclamp(d, -1.0f, 1.0f) * 127.0f) & 0xFFu) << 24);
}

__global__ void __launch_bounds__(256, VXL_GEOM_BLOCKS) k_gbuffer_models(FrameView F, GeomK K, const DrawDev* __restrict__ draws, int n_draws,
                                                        const uint32_t* __restrict__ pal_color, const uint32_t* __restrict__ pal_material,
                                                        uint32_t* __restrict__ depth24, uint32_t* __restrict__ normal, uint32_t* __restrict__ material,
                                                        uint32_t* __restrict__ albedo, float2* __restrict__ motion) {
    // same thread -> pixel mapping as k_gbuffer_primary (32x8 block, 8x4 warps)
    const int bpt_x = (F.tile_w + 31) / 32, bpt_y = (F.tile_h + 7) / 8;
    const int bpt = bpt_x * bpt_y;
    const int lt = blockIdx.x / bpt, b = blockIdx.x - lt * bpt;
    const int by = b / bpt_x, bx = b - by * bpt_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = bx * 32 + (warp & 3) * 8 + (lane & 7), ly = by * 8 + (warp >> 2) * 4 + (lane >> 3);
    const int gt = F.tile_first + lt * F.tile_stride;
    const int tyy = gt / F.tiles_x, txx = gt - tyy * F.tiles_x;
    const int px = txx * F.tile_w + lx, py = tyy * F.tile_h + ly;
    const bool valid = lx < F.tile_w && ly < F.tile_h && px < F.width && py < F.height && lt < F.n_tiles;
    const size_t idx = ((size_t)lt * F.tile_h + ly) * F.tile_w + lx;

    // Draw culling per block: a draw stays on the block's list unless the screen bounding box of its model box (the 8 corners
    // through In.MVPMatrix, +-2 pixels for rounding and the sub-pixel jitter) misses the block's 32x8 pixels.  Boxes with a corner
    // at or behind the camera plane always stay.  The list is a bit mask, walked in draw order (ties keep the first draw); the
    // exact per-pixel coverage test below still decides, so culling never changes a result.
    __shared__ uint32_t s_draws[GEOM_MAX_DRAWS / 32];
    const int n_words = (n_draws + 31) >> 5;
    for (int w = threadIdx.x; w < n_words; w += blockDim.x) s_draws[w] = 0u;
    __syncthreads();
    {
        const float bx0 = (float)(txx * F.tile_w + bx * 32) - 2.0f, bx1 = bx0 + 32.0f + 4.0f;
        const float by0 = (float)(tyy * F.tile_h + by * 8) - 2.0f, by1 = by0 + 8.0f + 4.0f;
        for (int c = threadIdx.x; c < n_draws; c += blockDim.x) {
            const DrawDev& d = draws[c];
            float mnx = 3.0e38f, mny = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f;
            bool keep = false;
            int behind = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 cp = mat_mul(d.mvp, make_float4((k & 1) ? d.size[0] * 0.1f : 0.0f, (k & 2) ? d.size[1] * 0.1f : 0.0f, (k & 4) ? d.size[2] * 0.1f : 0.0f, 1.0f));
                if (!(cp.w > 1.0e-3f)) keep = true;                                             // also catches NaN
                behind += cp.w < -1.0e-2f ? 1 : 0;
                const float sx = (cp.x / cp.w * 0.5f + 0.5f) * (float)F.width - 0.5f, sy = (0.5f - cp.y / cp.w * 0.5f) * (float)F.height - 0.5f;
                mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx); mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
            }
            // clip w is the view depth, linear over the box: with all 8 corners clearly behind the camera plane the whole box is, and
            // a pixel's ray (w >= 0 along it) cannot enter it -- half of a scene's draws when the camera stands among them
            if (behind == 8) continue;
            if (keep || !(mxx < bx0 || mnx > bx1 || mxy < by0 || mny > by1)) atomicOr(&s_draws[c >> 5], 1u << (c & 31));
        }
    }
    __syncthreads();
    if (!valid) return;

    uint32_t best = 0xFFFFFFu, onrm = 0u, omat = 0u, oalb = 0u;
    float omx = 0.0f, omy = 0.0f;
    const float u = ((float)px + 0.5f) / (float)F.width, v = ((float)py + 0.5f) / (float)F.height;
    const float ndcx = (2.0f * u - 1.0f) - K.jit[0] * K.ires[0] * 2.0f, ndcy = (1.0f - 2.0f * v) - K.jit[1] * K.ires[1] * 2.0f;
    const float4 fp = mat_mul(K.InvProj, make_float4(ndcx, ndcy, 1.0f, 1.0f));
    const float3 farv = make_float3(fp.x / fp.w, fp.y / fp.w, fp.z / fp.w);
    const float3 camW = make_float3(K.cam[0], K.cam[1], K.cam[2]);
    const float3 dirW = xyz(mat_mul(K.InvView, make_float4(farv.x, farv.y, farv.z, 1.0f))) - camW;
    for (int w = 0; w < n_words; ++w)
    for (uint32_t live = s_draws[w]; live; live &= live - 1u) {
        const int c = (w << 5) + __ffs((int)live) - 1;
        const DrawDev& d = draws[c];
        const float3 ld = xyz(mat_mul(d.inv, make_float4(dirW.x, dirW.y, dirW.z, 0.0f))) * 10.0f;
        const float lo[3] = {d.cam[0], d.cam[1], d.cam[2]}, dd[3] = {ld.x, ld.y, ld.z};
        if (lo[0] >= 0.0f && lo[0] <= d.size[0] && lo[1] >= 0.0f && lo[1] <= d.size[1] && lo[2] >= 0.0f && lo[2] <= d.size[2]) continue;
        float t0 = 0.0f, t1 = 3.0e38f;
        bool miss = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (dd[a] == 0.0f) { if (lo[a] < 0.0f || lo[a] > d.size[a]) miss = true; continue; }
            const float ta = (0.0f - lo[a]) / dd[a], tb = (d.size[a] - lo[a]) / dd[a];
            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
        }
        if (miss || !(t0 <= t1)) continue;
        // GeometryVoxel.frag main() :127-182
        const float4 sd = mat_mul(d.mvp, make_float4(ld.x, ld.y, ld.z, 0.0f));                       // :140
        float fu = sd.x / sd.w, fv = sd.y / sd.w;                                                 // :141
        fu += K.jit[0] * K.ires[0]; fv += K.jit[1] * K.ires[1];                                      // :142
        const vxl_model_hit h = traverse_model(d.M, make_float3(lo[0], lo[1], lo[2]), ld, fu, fv, K.frame, K.res[0], K.res[1]);
        if (!h.hit) continue;                                                                    // discard
        const float4 hp = make_float4(h.pos[0] * 0.1f, h.pos[1] * 0.1f, h.pos[2] * 0.1f, 1.0f);
        const float4 cur = mat_mul(K.PV, mat_mul(d.world, hp));                                     // :159-161
        const float depth = (1.0f - 0.0000001f * (float)d.rid) * (cur.w - 0.1f) / (4096.0f - 0.1f);   // :169-170
        const uint32_t d24 = (uint32_t)rintf(gclamp(depth, 0.0f, 1.0f) * 16777215.0f);
        if (d24 < best) {                                                                        // CompareOp::eLess
            best = d24;
            const uint32_t pc = __ldg(pal_color + (size_t)d.palette * 256 + h.material), pm = __ldg(pal_material + (size_t)d.palette * 256 + h.material);
            oalb = pack_unorm8x4(unorm8(pc), unorm8(pc >> 8), unorm8(pc >> 16), gstep((float)h.material, 16.0f));   // :154-155
            omat = pack_unorm8x4(unorm8(pm), unorm8(pm >> 8), unorm8(pm >> 16), unorm8(pm >> 24));       // :156
            const float4 nw = mat_mul(d.world, make_float4(h.normal[0], h.normal[1], h.normal[2], 0.0f));
            const float inv = 1.0f / sqrtf((nw.x * nw.x + nw.y * nw.y) + (nw.z * nw.z + nw.w * nw.w));   // :157 normalize(vec4)
            onrm = pack_snorm8x4(nw.x * inv, nw.y * inv, nw.z * inv, nw.w * inv);
            const float4 last = mat_mul(K.PVlast, mat_mul(d.last_world, hp));                       // :160,:162
            omx = -0.5f * (cur.x / cur.w - last.x / last.w); omy = 0.5f * (cur.y / cur.w - last.y / last.w);   // :166
        }
    }
    depth24[idx] = best; normal[idx] = onrm; material[idx] = omat; albedo[idx] = oalb;
    if (motion) motion[idx] = make_float2(omx, omy);
}

}  // namespace vxl

using namespace vxl;

namespace {
// mat4 * mat4 in glm's order (type_mat4x4.inl:630-646): column j = ((A0 B[j][0] + A1 B[j][1]) + A2 B[j][2]) + A3 B[j][3]
void mat_mat(const float* A, const float* B, float* R) {
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            R[j * 4 + r] = ((A[0 + r] * B[j * 4 + 0] + A[4 + r] * B[j * 4 + 1]) + A[8 + r] * B[j * 4 + 2]) + A[12 + r] * B[j * 4 + 3];
}
// inverse of an affine mat4 (last row 0 0 0 1; TransformSystem.cpp:124-135 builds T*R*S): adjugate / determinant of the 3x3 block,
// then -A^-1 t.  Stands in for the vertex shader's inverse(cmd.WorldMatrix) (GeometryVoxel.vert:66).
void affine_inverse(const float* M, float* R) {
    const float a = M[0], b = M[4], c = M[8], d = M[1], e = M[5], f = M[9], g = M[2], h = M[6], i = M[10];
    const float A = e * i - f * h, B = f * g - d * i, C = d * h - e * g;
    const float det = (a * A + b * B) + c * C;
    const float id = 1.0f / det;
    const float r00 = A * id, r01 = (c * h - b * i) * id, r02 = (b * f - c * e) * id;
    const float r10 = B * id, r11 = (a * i - c * g) * id, r12 = (c * d - a * f) * id;
    const float r20 = C * id, r21 = (b * g - a * h) * id, r22 = (a * e - b * d) * id;
    const float tx = M[12], ty = M[13], tz = M[14];
    R[0] = r00; R[4] = r01; R[8] = r02; R[12] = -((r00 * tx + r01 * ty) + r02 * tz);
    R[1] = r10; R[5] = r11; R[9] = r12; R[13] = -((r10 * tx + r11 * ty) + r12 * tz);
    R[2] = r20; R[6] = r21; R[10] = r22; R[14] = -((r20 * tx + r21 * ty) + r22 * tz);
    R[3] = 0.0f; R[7] = 0.0f; R[11] = 0.0f; R[15] = 1.0f;
}
void mat_vec(const float* m, const float* v, float* r) {           // (M0 v0 + M1 v1) + (M2 v2 + M3 v3), type_mat4x4.inl:561-572
    for (int k = 0; k < 4; ++k) r[k] = (m[k] * v[0] + m[4 + k] * v[1]) + (m[8 + k] * v[2] + m[12 + k] * v[3]);
}

int model_mips(vxl_ctx* ctx, int model_id, ModelMips* out) {
    ModelDev& m = ctx->models[model_id];
    ModelMips M;
    M.d[0] = m.voxels; M.sx[0] = m.sx; M.sy[0] = m.sy; M.sz[0] = m.sz;
    for (int l = 1; l < 3; ++l) { M.sx[l] = M.sx[l - 1] > 1 ? M.sx[l - 1] / 2 : 1; M.sy[l] = M.sy[l - 1] > 1 ? M.sy[l - 1] / 2 : 1; M.sz[l] = M.sz[l - 1] > 1 ? M.sz[l - 1] / 2 : 1; }
    if (!m.mip1) {                                                          // VoxAsset::Upload's chain, once per model
        uint8_t* p[3] = {const_cast<uint8_t*>(m.voxels), nullptr, nullptr};
        for (int l = 1; l < 3; ++l) {
            const size_t cnt = (size_t)M.sx[l] * M.sy[l] * M.sz[l];
            VXL_CUDA(cudaMalloc(&p[l], cnt));
            k_model_mip<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(p[l - 1], M.sx[l - 1], M.sy[l - 1], M.sz[l - 1], M.sx[l], M.sy[l], M.sz[l], p[l]);
            VXL_LAUNCH_CHECK(ctx);
        }
        m.mip1 = p[1]; m.mip2 = p[2];
    }
    M.d[1] = m.mip1; M.d[2] = m.mip2;
    *out = M;
    return VXL_OK;
}
}  // namespace

extern "C" {

int vxl_gbuffer_models(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_vox_cmd* cmds, int n_cmds,
                       const uint32_t* pal_color, const uint32_t* pal_material, const vxl_gbuffer_out* out) {
    if (!ctx || !view || !frame || n_cmds < 0 || (n_cmds > 0 && (!cmds || !pal_color || !pal_material)) || !out || !out->depth24 || !out->normal ||
        !out->material || !out->albedo) { set_error("vxl_gbuffer_models: bad argument"); return VXL_ERR_INVALID; }
    if (n_cmds > 4096) { set_error("vxl_gbuffer_models: more than 4096 draws (GeometryVoxelPipeline MAX_INSTANCES)"); return VXL_ERR_LIMIT; }
    vxl_frame f = *frame;
    f.depth24 = out->depth24; f.normal = out->normal; f.material = out->material; f.noise = out->depth24;   // the pass reads none of them
    FrameView F;
    if (int e = frame_view(&f, &F)) return e;
    if (F.n_tiles == 0) return VXL_OK;
    VXL_CUDA(cudaSetDevice(ctx->device));
    GeomK K;
    for (int i = 0; i < 16; ++i) { K.InvView[i] = view->InverseViewMatrix[i]; K.InvProj[i] = view->InverseProjectionMatrix[i]; }
    mat_mat(view->ProjectionMatrix, view->ViewMatrix, K.PV);
    mat_mat(view->ProjectionMatrix, view->LastViewMatrix, K.PVlast);
    for (int i = 0; i < 3; ++i) K.cam[i] = view->CameraPosition[i];
    for (int i = 0; i < 2; ++i) { K.jit[i] = view->Jitter[i]; K.ires[i] = view->iRes[i]; K.res[i] = view->Res[i]; }
    K.frame = view->Frame;
    std::vector<DrawDev> draws((size_t)n_cmds);
    for (int c = 0; c < n_cmds; ++c) {
        if (cmds[c].model < 0 || cmds[c].model >= (int)ctx->models.size()) { set_error("vxl_gbuffer_models: unknown model id"); return VXL_ERR_INVALID; }
        DrawDev& d = draws[c];
        affine_inverse(cmds[c].WorldMatrix, d.inv);
        mat_mat(K.PV, cmds[c].WorldMatrix, d.mvp);                                               // GeometryVoxel.vert:62
        for (int i = 0; i < 16; ++i) { d.world[i] = cmds[c].WorldMatrix[i]; d.last_world[i] = cmds[c].LastWorldMatrix[i]; }
        const float cw[4] = {K.cam[0], K.cam[1], K.cam[2], 1.0f};
        float lc[4];
        mat_vec(d.inv, cw, lc);
        for (int i = 0; i < 3; ++i) d.cam[i] = lc[i] * 10.0f;                                    // :66
        if (int e = model_mips(ctx, cmds[c].model, &d.M)) return e;
        d.size[0] = (float)d.M.sx[0]; d.size[1] = (float)d.M.sy[0]; d.size[2] = (float)d.M.sz[0];
        d.rid = cmds[c].VolumeRID; d.palette = cmds[c].PalleteIndex;
    }
    DrawDev* d_draws = nullptr;
    if (n_cmds > 0) {
        if (ctx->draws_cap < n_cmds) {                                       // persistent, grown on demand
            if (ctx->d_draws) VXL_CUDA(cudaFree(ctx->d_draws));
            ctx->d_draws = nullptr; ctx->draws_cap = 0;
            VXL_CUDA(cudaMalloc(&ctx->d_draws, sizeof(DrawDev) * (size_t)n_cmds));
            ctx->draws_cap = n_cmds;
        }
        d_draws = (DrawDev*)ctx->d_draws;
        // pageable source: the call returns once `draws` has been copied to the driver's staging memory
        VXL_CUDA(cudaMemcpyAsync(d_draws, draws.data(), sizeof(DrawDev) * (size_t)n_cmds, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int bpt = ((F.tile_w + 31) / 32) * ((F.tile_h + 7) / 8);
    k_gbuffer_models<<<(unsigned)(bpt * F.n_tiles), 256, 0, ctx->stream>>>(F, K, d_draws, n_cmds, pal_color, pal_material, out->depth24, out->normal,
                                                                         out->material, out->albedo, (float2*)out->motion);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

int vxl_trace_model_rays(vxl_ctx* ctx, int model_id, const vxl_model_ray* rays, int64_t n, int frame, float res_x, float res_y,
                         vxl_model_hit* out) {
    if (!ctx || n < 0 || (n > 0 && (!rays || !out))) { set_error("vxl_trace_model_rays: bad argument"); return VXL_ERR_INVALID; }
    if (model_id < 0 || model_id >= (int)ctx->models.size()) { set_error("vxl_trace_model_rays: unknown model id"); return VXL_ERR_INVALID; }
    if (n == 0) return VXL_OK;
    VXL_CUDA(cudaSetDevice(ctx->device));
    ModelMips M;
    if (int e = model_mips(ctx, model_id, &M)) return e;
    k_trace_model_rays<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(M, rays, (long long)n, frame, res_x, res_y, out);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

}  // extern "C"
