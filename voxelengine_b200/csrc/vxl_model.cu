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

__global__ void __launch_bounds__(256) k_trace_model_rays(ModelMips M, const vxl_model_ray* __restrict__ rays, long long n, int frame,
                                                          float res_x, float res_y, vxl_model_hit* __restrict__ out) {
    const long long ri = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= n) return;
    const vxl_model_ray r = rays[ri];
    vxl_model_hit h;
    h.hit = 0; h.material = 0u; h.fetches = 0; h.steps = 0;
    h.pos[0] = h.pos[1] = h.pos[2] = 0.f; h.normal[0] = h.normal[1] = h.normal[2] = 0.f;
    const float3 vsize = make_float3((float)M.sx[0], (float)M.sy[0], (float)M.sz[0]);
    const float3 cam = make_float3(r.cam[0], r.cam[1], r.cam[2]);
    const float3 direction = normalize3(make_float3(r.dir[0], r.dir[1], r.dir[2]));              // GeometryVoxel.frag:145
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
                    const float cxr = roundf(r.uv[0] * res_x * 0.5f), cyr = roundf(r.uv[1] * res_y * 0.5f);   // :99
                    const bool glass = voxel < 16u && gmod(cyr + cxr, 2.0f) == (float)(frame % 2);            // :100
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
    out[ri] = h;
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_trace_model_rays(vxl_ctx* ctx, int model_id, const vxl_model_ray* rays, int64_t n, int frame, float res_x, float res_y,
                         vxl_model_hit* out) {
    if (!ctx || n < 0 || (n > 0 && (!rays || !out))) { set_error("vxl_trace_model_rays: bad argument"); return VXL_ERR_INVALID; }
    if (model_id < 0 || model_id >= (int)ctx->models.size()) { set_error("vxl_trace_model_rays: unknown model id"); return VXL_ERR_INVALID; }
    if (n == 0) return VXL_OK;
    VXL_CUDA(cudaSetDevice(ctx->device));
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
    k_trace_model_rays<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(M, rays, (long long)n, frame, res_x, res_y, out);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

}  // extern "C"
