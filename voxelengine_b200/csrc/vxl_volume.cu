// vxl_volume.cu -- the world occupancy volume: allocation, region upload, device voxeliser with the
// reference's sequential semantics, and the synthetic-input generators (terrain, primary G-buffer).
//
// Replaces Sources/World/Systems/ShadowVoxSystem.{h,cpp} (host triple loops + staged
// vkCmdCopyBufferToImage) with kernels that write the packed bytes in HBM directly.
#include <algorithm>
#include <cstring>

#include "vxl_internal.h"
#include "vxl_math.cuh"
#include "vxl_trace.cuh"

namespace vxl {

constexpr unsigned long long HEMPTY = ~0ull;

struct VoxDims { int sx, sy, sz; };

// ---- voxeliser ------------------------------------------------------------------------------------
// Reference semantics (ShadowVoxSystem.cpp:116-191): entities are visited in order; each clears every
// solid voxel at its previous transform, then sets every solid voxel at its current transform; a
// later command overwrites an earlier one bit by bit.  On the GPU every (entity, phase, voxel)
// write is tagged with its position in that sequence (seq = 2*entity + phase), the latest tag per
// touched world voxel is kept with atomicMax in an open-addressing table, and a second kernel
// applies bit = seq & 1.  This is order-independent and reproduces the sequential result exactly.

struct Basis { float3 o, dx, dy, dz; };

// glm::translate(m, -pivot) then o/dx/dy/dz (ShadowVoxSystem.cpp:134-140; Vendor/glm/ext/matrix_transform.inl:10-15)
VXL_DI Basis basis_from(const float* __restrict__ m, const float* __restrict__ pivot, bool use_pivot) {
    Basis b;
    if (use_pivot) {
        const float v0 = -pivot[0], v1 = -pivot[1], v2 = -pivot[2];
        b.o.x = ((m[0] * v0 + m[4] * v1) + m[8] * v2) + m[12];
        b.o.y = ((m[1] * v0 + m[5] * v1) + m[9] * v2) + m[13];
        b.o.z = ((m[2] * v0 + m[6] * v1) + m[10] * v2) + m[14];
    } else {
        b.o = make_float3(m[12], m[13], m[14]);
    }
    b.dx = make_float3(m[0] * 0.1f, m[1] * 0.1f, m[2] * 0.1f);
    b.dy = make_float3(m[4] * 0.1f, m[5] * 0.1f, m[6] * 0.1f);
    b.dz = make_float3(m[8] * 0.1f, m[9] * 0.1f, m[10] * 0.1f);
    return b;
}

VXL_DI unsigned long long mix64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__global__ void k_vox_init(VoxDims D, int n, int* __restrict__ aabb, int* __restrict__ tight) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    tight[e * 6 + 0] = tight[e * 6 + 1] = tight[e * 6 + 2] = 0x7fffffff;                     // the box of what the command really wrote
    tight[e * 6 + 3] = tight[e * 6 + 4] = tight[e * 6 + 5] = -0x7fffffff - 1;
    aabb[e * 6 + 0] = D.sx - 1; aabb[e * 6 + 1] = D.sy - 1; aabb[e * 6 + 2] = D.sz - 1;   // startmin (:128, texel units, sic)
    aabb[e * 6 + 3] = 0; aabb[e * 6 + 4] = 0; aabb[e * 6 + 5] = 0;                        // startmax
}

__global__ void __launch_bounds__(256) k_vox_emit(VoxDims D, const ModelDev* __restrict__ models, const vxl_entity* __restrict__ ents,
                                                  unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                                                  unsigned long long mask, int* __restrict__ aabb, int* __restrict__ tight) {
    const int e = blockIdx.y >> 1, phase = blockIdx.y & 1;
    const vxl_entity& en = ents[e];
    const bool destroy = (en.flags & VXL_ENT_DESTROY) != 0;
    if (destroy && phase == 1) return;
    const ModelDev M = models[en.model];
    const Basis b = destroy ? basis_from(en.cur, en.pivot, false) : basis_from(phase ? en.cur : en.prev, en.pivot, true);
    const unsigned seq1 = (unsigned)(2 * e + phase) + 1u;
    const int total = M.sx * M.sy * M.sz;
    const int VX = D.sx * 2, VY = D.sy * 2, VZ = D.sz * 2;
    int mnx = 0x7fffffff, mny = 0x7fffffff, mnz = 0x7fffffff, mxx = -0x7fffffff - 1, mxy = mxx, mxz = mxx;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        if (M.voxels[idx] < 16) continue;                                    // :145 palette < 16 = glass / empty
        const int x = idx % M.sx, yz = idx / M.sx, y = yz % M.sy, z = yz / M.sy;
        const float3 wp = b.o + b.dx * (float)x + b.dy * (float)y + b.dz * (float)z;   // :146
        const int fx = f2i(wp.x * 10.0f), fy = f2i(wp.y * 10.0f), fz = f2i(wp.z * 10.0f);   // :147
        mnx = min(mnx, fx); mny = min(mny, fy); mnz = min(mnz, fz);
        mxx = max(mxx, fx); mxy = max(mxy, fy); mxz = max(mxz, fz);
        if (fx < 0 || fy < 0 || fz < 0 || fx >= VX || fy >= VY || fz >= VZ) continue;       // :83
        const unsigned long long key = (unsigned long long)fx + (unsigned long long)VX * ((unsigned long long)fy + (unsigned long long)VY * (unsigned long long)fz);
        unsigned long long slot = mix64(key) & mask;
        while (true) {
            const unsigned long long prev = atomicCAS(&keys[slot], HEMPTY, key);
            if (prev == HEMPTY || prev == key) { atomicMax(&vals[slot], seq1); break; }
            slot = (slot + 1) & mask;
        }
    }
    // per-entity AABB over every solid voxel's (possibly out-of-volume) coordinate (:148-149)
    mnx = __reduce_min_sync(0xFFFFFFFFu, mnx); mny = __reduce_min_sync(0xFFFFFFFFu, mny); mnz = __reduce_min_sync(0xFFFFFFFFu, mnz);
    mxx = __reduce_max_sync(0xFFFFFFFFu, mxx); mxy = __reduce_max_sync(0xFFFFFFFFu, mxy); mxz = __reduce_max_sync(0xFFFFFFFFu, mxz);
    if ((threadIdx.x & 31) == 0 && mnx != 0x7fffffff) {
        atomicMin(&aabb[e * 6 + 0], mnx); atomicMin(&aabb[e * 6 + 1], mny); atomicMin(&aabb[e * 6 + 2], mnz);
        atomicMax(&aabb[e * 6 + 3], mxx); atomicMax(&aabb[e * 6 + 4], mxy); atomicMax(&aabb[e * 6 + 5], mxz);
        // the same box without the reference's start values (:128 seeds the minimum with the TEXEL extent): what the occupancy
        // levels have to be rebuilt for (vxl_occupancy.cu)
        atomicMin(&tight[e * 6 + 0], mnx); atomicMin(&tight[e * 6 + 1], mny); atomicMin(&tight[e * 6 + 2], mnz);
        atomicMax(&tight[e * 6 + 3], mxx); atomicMax(&tight[e * 6 + 4], mxy); atomicMax(&tight[e * 6 + 5], mxz);
    }
}

__global__ void __launch_bounds__(256) k_vox_resolve(VoxDims D, const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals,
                                                     unsigned long long cap, unsigned* __restrict__ words) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned long long key = keys[i];
    if (key == HEMPTY) return;
    const unsigned value = (vals[i] - 1u) & 1u;
    const unsigned long long VX = (unsigned long long)D.sx * 2, VY = (unsigned long long)D.sy * 2;
    const int x = (int)(key % VX), y = (int)((key / VX) % VY), z = (int)(key / (VX * VY));
    const int bit = (x & 1) | ((y & 1) << 1) | ((z & 1) << 2);                            // :85
    const size_t off = (size_t)(x >> 1) + (size_t)(y >> 1) * D.sx + (size_t)(z >> 1) * ((size_t)D.sx * D.sy);
    const unsigned m = 1u << (unsigned)((off & 3) * 8 + bit);
    if (value) atomicOr(&words[off >> 2], m); else atomicAnd(&words[off >> 2], ~m);       // :93
}

__global__ void k_vox_regions(VoxDims D, int n, const int* __restrict__ aabb, vxl_region* __restrict__ regions, int* __restrict__ valid) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int mn[3] = {aabb[e * 6], aabb[e * 6 + 1], aabb[e * 6 + 2]}, mx[3] = {aabb[e * 6 + 3], aabb[e * 6 + 4], aabb[e * 6 + 5]};
    const int hi[3] = {D.sx - 1, D.sy - 1, D.sz - 1};
    vxl_region r; r.x = r.y = r.z = 0; r.w = r.h = r.d = 0; r.mip = 0;
    int ok = 0;
    if (mx[0] != 0 || mx[1] != 0 || mx[2] != 0) {                                         // :181
        for (int a = 0; a < 3; ++a) { mn[a] /= 2; mx[a] /= 2; mn[a] = max(mn[a], 0); mx[a] = min(mx[a], hi[a]); }   // :182-186
        r.x = mn[0]; r.y = mn[1]; r.z = mn[2];
        r.w = (uint32_t)(mx[0] - mn[0] + 1); r.h = (uint32_t)(mx[1] - mn[1] + 1); r.d = (uint32_t)(mx[2] - mn[2] + 1);
        ok = 1;
    }
    regions[e] = r;
    valid[e] = ok;
}

// ---- synthetic terrain ------------------------------------------------------------------------------
// Gradient noise after the published FastNoise 0.4 Perlin algorithm (Vendor/FastNoise/FastNoise.cpp
// :826-880 3-D, :950-985 2-D, quintic interpolation, frequency 0.01), combined as
// Noise::GetTerrainNoise (Sources/Util/Noise.cpp:93-135).  Input synthesis, restated not copied.
__constant__ float c_GX[12] = {1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
__constant__ float c_GY[12] = {1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
__constant__ float c_GZ[12] = {0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};

VXL_DI int fast_floor(float f) { return f >= 0 ? (int)f : (int)f - 1; }
VXL_DI float quintic(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
VXL_DI float lerpf(float a, float b, float t) { return a + t * (b - a); }
VXL_DI float grad2(const uint8_t* P, int x, int y, float xd, float yd) {
    const int i = P[512 + (x & 0xff) + P[(y & 0xff)]];
    return xd * c_GX[i] + yd * c_GY[i];
}
VXL_DI float grad3(const uint8_t* P, int x, int y, int z, float xd, float yd, float zd) {
    const int i = P[512 + (x & 0xff) + P[(y & 0xff) + P[(z & 0xff)]]];
    return xd * c_GX[i] + yd * c_GY[i] + zd * c_GZ[i];
}
VXL_DI float perlin2(const uint8_t* P, float x, float y) {
    const int x0 = fast_floor(x), y0 = fast_floor(y), x1 = x0 + 1, y1 = y0 + 1;
    const float xs = quintic(x - (float)x0), ys = quintic(y - (float)y0);
    const float xd0 = x - (float)x0, yd0 = y - (float)y0, xd1 = xd0 - 1.0f, yd1 = yd0 - 1.0f;
    const float xf0 = lerpf(grad2(P, x0, y0, xd0, yd0), grad2(P, x1, y0, xd1, yd0), xs);
    const float xf1 = lerpf(grad2(P, x0, y1, xd0, yd1), grad2(P, x1, y1, xd1, yd1), xs);
    return lerpf(xf0, xf1, ys);
}
VXL_DI float perlin3(const uint8_t* P, float x, float y, float z) {
    const int x0 = fast_floor(x), y0 = fast_floor(y), z0 = fast_floor(z), x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    const float xs = quintic(x - (float)x0), ys = quintic(y - (float)y0), zs = quintic(z - (float)z0);
    const float xd0 = x - (float)x0, yd0 = y - (float)y0, zd0 = z - (float)z0;
    const float xd1 = xd0 - 1.0f, yd1 = yd0 - 1.0f, zd1 = zd0 - 1.0f;
    const float xf00 = lerpf(grad3(P, x0, y0, z0, xd0, yd0, zd0), grad3(P, x1, y0, z0, xd1, yd0, zd0), xs);
    const float xf10 = lerpf(grad3(P, x0, y1, z0, xd0, yd1, zd0), grad3(P, x1, y1, z0, xd1, yd1, zd0), xs);
    const float xf01 = lerpf(grad3(P, x0, y0, z1, xd0, yd0, zd1), grad3(P, x1, y0, z1, xd1, yd0, zd1), xs);
    const float xf11 = lerpf(grad3(P, x0, y1, z1, xd0, yd1, zd1), grad3(P, x1, y1, z1, xd1, yd1, zd1), xs);
    const float yf0 = lerpf(xf00, xf10, ys), yf1 = lerpf(xf01, xf11, ys);
    return lerpf(yf0, yf1, zs);
}
constexpr float NOISE_FREQ = 0.01f;
VXL_DI float octave2(const uint8_t* P, float x, float y, int octaves) {
    float total = 0.0f, frequency = 1.0f, amplitude = 1.0f, maxValue = 0.0f;
    for (int i = 0; i < octaves; i++) {
        total += perlin2(P, (x * frequency) * NOISE_FREQ, (y * frequency) * NOISE_FREQ) * amplitude;
        maxValue += amplitude; amplitude *= 0.5f; frequency *= 2.0f;
    }
    return total / maxValue;
}
VXL_DI float octave3(const uint8_t* P, float x, float y, float z, int octaves) {
    float total = 0.0f, frequency = 1.0f, amplitude = 1.0f, maxValue = 0.0f;
    for (int i = 0; i < octaves; i++) {
        total += perlin3(P, (x * frequency) * NOISE_FREQ, (y * frequency) * NOISE_FREQ, (z * frequency) * NOISE_FREQ) * amplitude;
        maxValue += amplitude; amplitude *= 0.5f; frequency *= 2.0f;
    }
    return total / maxValue;
}

__global__ void __launch_bounds__(256) k_terrain2d(VoxDims D, const uint8_t* __restrict__ g_perm, float* __restrict__ col2) {
    __shared__ uint8_t P[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) P[i] = g_perm[i];
    __syncthreads();
    const int VX = D.sx * 2, VZ = D.sz * 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)VX * VZ) return;
    const int vx = (int)(i % VX), vz = (int)(i / VX);
    col2[i] = octave2(P, (float)vx * 1.0f, (float)vz * 1.0f, 4) + 0.0f;
}

__global__ void __launch_bounds__(256) k_terrain(VoxDims D, const uint8_t* __restrict__ g_perm, const float* __restrict__ col2, uint8_t* __restrict__ bytes) {
    __shared__ uint8_t P[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) P[i] = g_perm[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)D.sx * D.sy * D.sz;
    if (i >= total) return;
    const int tx = (int)(i % D.sx), ty = (int)((i / D.sx) % D.sy), tz = (int)(i / ((long long)D.sx * D.sy));
    const float NY = (float)(2 * D.sy);
    const int VX = D.sx * 2;
    unsigned byte = 0;
#pragma unroll 1
    for (int bit = 0; bit < 8; ++bit) {
        const int vx = 2 * tx + (bit & 1), vy = 2 * ty + ((bit >> 1) & 1), vz = 2 * tz + (bit >> 2);
        float v = col2[(size_t)vz * VX + vx];
        v += octave3(P, (float)vx * 2.0f, (float)vy * 2.0f, (float)vz * 2.0f, 3) + 0.0f;
        const float thr = ((float)vy / NY - 0.5f) * 2.0f;
        if (v > thr) byte |= 1u << bit;
    }
    bytes[i] = (uint8_t)byte;
}

// ---- synthetic primary-visibility G-buffer -----------------------------------------------------------
struct PrimK { float InvView[16], View[16], InvProj[16]; };

__global__ void __launch_bounds__(256) k_gbuffer_primary(VolView V, FrameView F, PrimK K, uint32_t* __restrict__ depth24,
                                                         uint32_t* __restrict__ normal, uint32_t* __restrict__ material) {
    // same thread->pixel mapping as the passes (32x8 block, 8x4 warps)
    const int bpt_x = (F.tile_w + 31) / 32, bpt_y = (F.tile_h + 7) / 8;
    const int bpt = bpt_x * bpt_y;
    const int lt = blockIdx.x / bpt, b = blockIdx.x - lt * bpt;
    const int by = b / bpt_x, bx = b - by * bpt_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = bx * 32 + (warp & 3) * 8 + (lane & 7), ly = by * 8 + (warp >> 2) * 4 + (lane >> 3);
    const int gt = F.tile_first + lt * F.tile_stride;
    const int tyy = gt / F.tiles_x, txx = gt - tyy * F.tiles_x;
    const int px = txx * F.tile_w + lx, py = tyy * F.tile_h + ly;
    if (!(lx < F.tile_w && ly < F.tile_h && px < F.width && py < F.height && lt < F.n_tiles)) return;
    const size_t idx = ((size_t)lt * F.tile_h + ly) * F.tile_w + lx;

    const float u = ((float)px + 0.5f) / (float)F.width, v = ((float)py + 0.5f) / (float)F.height;
    const float ndcx = 2.0f * u - 1.0f, ndcy = 1.0f - 2.0f * v;
    const float4 f = mat_mul(K.InvProj, make_float4(ndcx, ndcy, 1.0f, 1.0f));
    const float3 farvec = make_float3(f.x / f.w, f.y / f.w, f.z / f.w);
    const float3 dir = normalize3(xyz(mat_mul(K.InvView, make_float4(farvec.x, farvec.y, farvec.z, 0.0f))));
    const float3 org = make_float3(K.InvView[12], K.InvView[13], K.InvView[14]) * 10.0f;
    const float3 box = make_float3((float)(V.sx * 2), (float)(V.sy * 2), (float)(V.sz * 2));
    const int max_steps = 2 * (V.sx + V.sy + V.sz) * 2 + 8;
    uint32_t od = 0xFFFFFFu, on = 0u, om = 0u;
    float t0 = 0.0f;
    const bool inside = org.x >= 0.0f && org.y >= 0.0f && org.z >= 0.0f && org.x < box.x && org.y < box.y && org.z < box.z;
    bool ok = true;
    if (!inside) {
        float tmin = 0.0f, tmax = 3.0e38f;
        const float o[3] = {org.x, org.y, org.z}, d[3] = {dir.x, dir.y, dir.z}, bb[3] = {box.x, box.y, box.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (d[a] == 0.0f) { if (o[a] < 0.0f || o[a] >= bb[a]) ok = false; continue; }
            const float ta = (0.0f - o[a]) / d[a], tb = (bb[a] - o[a]) / d[a];
            const float lo = fminf(ta, tb), hi = fmaxf(ta, tb);
            tmin = fmaxf(tmin, lo); tmax = fminf(tmax, hi);
        }
        if (!(tmin <= tmax)) ok = false;
        t0 = tmin + 0.001f;
    }
    if (ok) {
        const float3 o = org + dir * t0;
        const float3 stepSign = make_float3(gsign(dir.x), gsign(dir.y), gsign(dir.z));
        const float3 t_delta = make_float3(1.0f, 1.0f, 1.0f) / (dir * stepSign);
        int cx = f2i(floorf(o.x)), cy = f2i(floorf(o.y)), cz = f2i(floorf(o.z));
        const float3 nb = make_float3((float)cx, (float)cy, (float)cz) + (stepSign * 0.5f + make_float3(0.5f, 0.5f, 0.5f));
        float3 t_max = (nb - o) / dir;
        float t_enter = 0.0f;
        float3 nrm = make_float3(0.0f, 1.0f, 0.0f);
        for (int n = 0; n < max_steps; ++n) {
            if (cx < 0 || cy < 0 || cz < 0 || cx >= V.sx * 2 || cy >= V.sy * 2 || cz >= V.sz * 2) break;
            if (get_volume_at(V, cx, cy, cz, 0)) {
                const float3 hit = o + dir * t_enter;
                const float3 hw = hit * 0.1f;
                const float4 pv = mat_mul(K.View, make_float4(hw.x, hw.y, hw.z, 1.0f));
                const float w = -pv.z;
                float dlin = (w - 0.1f) / (4096.0f - 0.1f);                    // GeometryVoxel.frag:169
                dlin = gclamp(dlin, 0.0f, 1.0f);
                od = (uint32_t)f2i(floorf(dlin * 16777215.0f + 0.5f));
                const int qx = f2i(floorf(nrm.x * 127.0f + 0.5f)), qy = f2i(floorf(nrm.y * 127.0f + 0.5f)), qz = f2i(floorf(nrm.z * 127.0f + 0.5f));
                on = ((uint32_t)(qx & 0xFF)) | ((uint32_t)(qy & 0xFF) << 8) | ((uint32_t)(qz & 0xFF) << 16);
                const uint32_t rough = (uint32_t)((cx * 7 + cy * 13 + cz * 29) & 255);
                om = rough | (255u << 24);
                break;
            }
            const float3 select = make_float3(gstep(t_max.x, t_max.z) * gstep(t_max.x, t_max.y),
                                              gstep(t_max.y, t_max.x) * gstep(t_max.y, t_max.z),
                                              gstep(t_max.z, t_max.y) * gstep(t_max.z, t_max.x));
            t_enter = dot3(t_max, select);
            if (t_enter != t_enter) break;
            nrm = (stepSign * -1.0f) * select;
            const float3 adv = select * stepSign;
            cx += f2i(adv.x); cy += f2i(adv.y); cz += f2i(adv.z);
            t_max = t_max + t_delta * select;
        }
    }
    depth24[idx] = od; normal[idx] = on; material[idx] = om;
}

static size_t padded_bytes(const vxl_volume* v) { return (((size_t)v->sx * v->sy * v->sz) + 255) & ~(size_t)255; }

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_volume_create(vxl_ctx* ctx, int sx, int sy, int sz, vxl_volume** out) {
    if (!ctx || !out || sx <= 0 || sy <= 0 || sz <= 0 || sx > 16384 || sy > 16384 || sz > 16384) { set_error("vxl_volume_create: bad argument"); return VXL_ERR_INVALID; }
    *out = nullptr;
    VXL_CUDA(cudaSetDevice(ctx->device));
    vxl_volume* v = new vxl_volume();
    v->ctx = ctx; v->sx = sx; v->sy = sy; v->sz = sz;
    cudaError_t e = cudaMalloc(&v->d_bytes, padded_bytes(v));
    if (e != cudaSuccess) { delete v; return cuda_fail(e, "cudaMalloc(volume)"); }
    VXL_CUDA(cudaMemsetAsync(v->d_bytes, 0, padded_bytes(v), ctx->stream));   // ShadowVoxSystem.cpp:66-70
    v->dirty = true; v->dirty_partial = false;
    *out = v;
    return VXL_OK;
}

int vxl_volume_destroy(vxl_volume* v) {
    if (!v) return VXL_OK;
    cudaStreamSynchronize(v->ctx->stream);
    cudaFree(v->d_bytes);
    cudaFree(v->tex.d_words);
    for (auto& L : v->occ) cudaFree(L.d_words);
    for (auto& L : v->dil) cudaFree(L.d_words);
    delete v;
    return VXL_OK;
}

int vxl_volume_dims(const vxl_volume* v, int* sx, int* sy, int* sz) {
    if (!v) { set_error("vxl_volume_dims: vol is NULL"); return VXL_ERR_INVALID; }
    if (sx) *sx = v->sx; if (sy) *sy = v->sy; if (sz) *sz = v->sz;
    return VXL_OK;
}

// Vendor/evk/evk.cpp:759-780: bufferOffset = x + y*W + z*W*H, bufferRowLength = W, bufferImageHeight = H
int vxl_volume_upload_regions(vxl_volume* v, const uint8_t* host, const vxl_region* regions, int n) {
    if (!v || n < 0 || (n > 0 && (!host || !regions))) { set_error("vxl_volume_upload_regions: bad argument"); return VXL_ERR_INVALID; }
    for (int i = 0; i < n; ++i) {
        const vxl_region& r = regions[i];
        if (r.mip != 0) { set_error("vxl_volume_upload_regions: the world volume has one mip level"); return VXL_ERR_INVALID; }
        const int x0 = std::max(r.x, 0), y0 = std::max(r.y, 0), z0 = std::max(r.z, 0);
        const long long x1 = std::min<long long>((long long)r.x + r.w, v->sx), y1 = std::min<long long>((long long)r.y + r.h, v->sy), z1 = std::min<long long>((long long)r.z + r.d, v->sz);
        if (x1 <= x0 || y1 <= y0 || z1 <= z0) continue;
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof p);
        p.srcPtr = make_cudaPitchedPtr((void*)host, (size_t)v->sx, (size_t)v->sx, (size_t)v->sy);
        p.dstPtr = make_cudaPitchedPtr((void*)v->d_bytes, (size_t)v->sx, (size_t)v->sx, (size_t)v->sy);
        p.srcPos = make_cudaPos((size_t)x0, (size_t)y0, (size_t)z0);
        p.dstPos = p.srcPos;
        p.extent = make_cudaExtent((size_t)(x1 - x0), (size_t)(y1 - y0), (size_t)(z1 - z0));
        p.kind = cudaMemcpyHostToDevice;
        VXL_CUDA(cudaMemcpy3DAsync(&p, v->ctx->stream));
    }
    if (n > 0) { v->dirty = true; v->dirty_partial = false; }
    return VXL_OK;
}

int vxl_volume_upload(vxl_volume* v, const uint8_t* host) {
    if (!v || !host) { set_error("vxl_volume_upload: bad argument"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemcpyAsync(v->d_bytes, host, (size_t)v->sx * v->sy * v->sz, cudaMemcpyHostToDevice, v->ctx->stream));
    v->dirty = true; v->dirty_partial = false;
    return VXL_OK;
}

int vxl_volume_download(vxl_volume* v, uint8_t* host) {
    if (!v || !host) { set_error("vxl_volume_download: bad argument"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemcpyAsync(host, v->d_bytes, (size_t)v->sx * v->sy * v->sz, cudaMemcpyDeviceToHost, v->ctx->stream));
    VXL_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return VXL_OK;
}

int vxl_volume_clear(vxl_volume* v) {
    if (!v) { set_error("vxl_volume_clear: vol is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemsetAsync(v->d_bytes, 0, padded_bytes(v), v->ctx->stream));
    v->dirty = true; v->dirty_partial = false;
    return VXL_OK;
}

int vxl_volume_device_ptr(vxl_volume* v, uint8_t** out) {
    if (!v || !out) { set_error("vxl_volume_device_ptr: bad argument"); return VXL_ERR_INVALID; }
    *out = v->d_bytes;
    return VXL_OK;
}

int vxl_volume_mark_dirty(vxl_volume* v) {
    if (!v) { set_error("vxl_volume_mark_dirty: vol is NULL"); return VXL_ERR_INVALID; }
    v->dirty = true; v->dirty_partial = false;
    return VXL_OK;
}

int vxl_model_create(vxl_ctx* ctx, const uint8_t* voxels, int sx, int sy, int sz, int* out_id) {
    if (!ctx || !voxels || !out_id || sx <= 0 || sy <= 0 || sz <= 0 || (long long)sx * sy * sz > (1ll << 30)) { set_error("vxl_model_create: bad argument"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)sx * sy * sz;
    ModelDev m;
    uint8_t* d = nullptr;
    VXL_CUDA(cudaMalloc(&d, n));
    VXL_CUDA(cudaMemcpyAsync(d, voxels, n, cudaMemcpyHostToDevice, ctx->stream));
    VXL_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned solid = 0;
    for (size_t i = 0; i < n; ++i) solid += voxels[i] >= 16;
    m.voxels = d; m.sx = sx; m.sy = sy; m.sz = sz; m.solid = solid; m.mip1 = nullptr; m.mip2 = nullptr;
    ctx->models.push_back(m);
    *out_id = (int)ctx->models.size() - 1;
    return VXL_OK;
}

int vxl_volume_voxelize(vxl_volume* v, const vxl_entity* ents, int n, vxl_region* out_regions, int32_t* out_valid) {
    if (!v || n < 0 || (n > 0 && !ents)) { set_error("vxl_volume_voxelize: bad argument"); return VXL_ERR_INVALID; }
    if (n == 0) return VXL_OK;
    if (n > (1 << 29)) { set_error("vxl_volume_voxelize: too many commands"); return VXL_ERR_LIMIT; }
    vxl_ctx* c = v->ctx;
    VXL_CUDA(cudaSetDevice(c->device));
    // upper bound of table entries and of the per-model launch width
    unsigned long long ops = 0;
    int max_vox = 1;
    for (int i = 0; i < n; ++i) {
        if (ents[i].model < 0 || ents[i].model >= (int)c->models.size()) { set_error("vxl_volume_voxelize: unknown model id"); return VXL_ERR_INVALID; }
        const ModelDev& m = c->models[ents[i].model];
        ops += (unsigned long long)m.solid * ((ents[i].flags & VXL_ENT_DESTROY) ? 1u : 2u);
        max_vox = std::max(max_vox, m.sx * m.sy * m.sz);
    }
    size_t cap = 4096;
    while (cap * 2 < ops * 3) cap <<= 1;                 // load factor <= 2/3 if every write hit a different voxel (clear + set mostly overlap)
    if (cap > c->hcap) {
        if (c->d_hkeys) { VXL_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->d_hkeys); cudaFree(c->d_hvals); c->d_hkeys = nullptr; c->d_hvals = nullptr; c->hcap = 0; }
        VXL_CUDA(cudaMalloc(&c->d_hkeys, cap * sizeof(unsigned long long)));
        VXL_CUDA(cudaMalloc(&c->d_hvals, cap * sizeof(unsigned)));
        c->hcap = cap;
    }
    if (n > c->ents_cap) {
        if (c->d_ents) { VXL_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->d_ents); cudaFree(c->d_aabb); c->d_ents = nullptr; c->d_aabb = nullptr; }
        // region/valid outputs live behind the aabb block: [n][6] ints, [n] regions, [n] valid
        VXL_CUDA(cudaMalloc(&c->d_ents, (size_t)n * sizeof(vxl_entity)));
        VXL_CUDA(cudaMalloc(&c->d_aabb, (size_t)n * (12 * sizeof(int) + sizeof(vxl_region) + sizeof(int))));
        c->ents_cap = n;
    }
    if ((int)c->models.size() > c->d_models_cap) {
        if (c->d_models) { VXL_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->d_models); c->d_models = nullptr; }
        c->d_models_cap = (int)c->models.size() + 16;
        VXL_CUDA(cudaMalloc(&c->d_models, (size_t)c->d_models_cap * sizeof(ModelDev)));
    }
    VXL_CUDA(cudaMemcpyAsync(c->d_models, c->models.data(), c->models.size() * sizeof(ModelDev), cudaMemcpyHostToDevice, c->stream));
    VXL_CUDA(cudaMemcpyAsync(c->d_ents, ents, (size_t)n * sizeof(vxl_entity), cudaMemcpyHostToDevice, c->stream));
    VXL_CUDA(cudaMemsetAsync(c->d_hkeys, 0xFF, cap * sizeof(unsigned long long), c->stream));
    VXL_CUDA(cudaMemsetAsync(c->d_hvals, 0, cap * sizeof(unsigned), c->stream));
    const VoxDims D{v->sx, v->sy, v->sz};
    int* d_aabb = c->d_aabb;
    vxl_region* d_regions = (vxl_region*)(d_aabb + (size_t)c->ents_cap * 6);
    int* d_valid = (int*)(d_regions + c->ents_cap);
    int* d_tight = d_valid + c->ents_cap;
    c->aabb_gen++;
    k_vox_init<<<(n + 255) / 256, 256, 0, c->stream>>>(D, n, d_aabb, d_tight);
    VXL_LAUNCH_CHECK(c);
    const int bx = std::min(std::max((max_vox + 256 * 8 - 1) / (256 * 8), 1), 1024);
    // gridDim.y is limited to 65535: issue the command list in slabs (the sequence tag uses the global index)
    for (int e0 = 0; e0 < n; e0 += 32767) {
        const int ne = std::min(32767, n - e0);
        k_vox_emit<<<dim3((unsigned)bx, (unsigned)(ne * 2)), 256, 0, c->stream>>>(D, c->d_models, c->d_ents + e0, c->d_hkeys, c->d_hvals, (unsigned long long)cap - 1, d_aabb + (size_t)e0 * 6, d_tight + (size_t)e0 * 6);
        VXL_LAUNCH_CHECK(c);
        if (n > 32767) {
            // sequence tags restart per slab, so resolve each slab before the next one is emitted
            k_vox_resolve<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(D, c->d_hkeys, c->d_hvals, cap, (unsigned*)v->d_bytes);
            VXL_LAUNCH_CHECK(c);
            VXL_CUDA(cudaMemsetAsync(c->d_hkeys, 0xFF, cap * sizeof(unsigned long long), c->stream));
            VXL_CUDA(cudaMemsetAsync(c->d_hvals, 0, cap * sizeof(unsigned), c->stream));
        }
    }
    if (n <= 32767) {
        k_vox_resolve<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(D, c->d_hkeys, c->d_hvals, cap, (unsigned*)v->d_bytes);
        VXL_LAUNCH_CHECK(c);
    }
    // levels that were up to date before this call only need the commands' boxes rebuilt (a second call before the rebuild, or a
    // command list too long for one block row per box, falls back to the full rebuild)
    v->dirty_partial = !v->dirty && v->occ[0].d_words != nullptr && n <= 16384;
    v->dirty_boxes = d_tight; v->n_dirty_boxes = n; v->dirty_gen = c->aabb_gen;
    v->dirty = true;
    if (out_regions || out_valid) {
        k_vox_regions<<<(n + 255) / 256, 256, 0, c->stream>>>(D, n, d_aabb, d_regions, d_valid);
        VXL_LAUNCH_CHECK(c);
        if (out_regions) VXL_CUDA(cudaMemcpyAsync(out_regions, d_regions, (size_t)n * sizeof(vxl_region), cudaMemcpyDeviceToHost, c->stream));
        if (out_valid) VXL_CUDA(cudaMemcpyAsync(out_valid, d_valid, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        VXL_CUDA(cudaStreamSynchronize(c->stream));
    }
    return VXL_OK;
}

int vxl_volume_gen_terrain(vxl_volume* v) {
    if (!v) { set_error("vxl_volume_gen_terrain: vol is NULL"); return VXL_ERR_INVALID; }
    vxl_ctx* c = v->ctx;
    VXL_CUDA(cudaSetDevice(c->device));
    const VoxDims D{v->sx, v->sy, v->sz};
    const size_t ncol = (size_t)v->sx * 2 * v->sz * 2;
    float* col2 = nullptr;
    VXL_CUDA(cudaMalloc(&col2, ncol * sizeof(float)));
    k_terrain2d<<<(unsigned)((ncol + 255) / 256), 256, 0, c->stream>>>(D, c->d_perm, col2);
    VXL_LAUNCH_CHECK(c);
    const size_t total = (size_t)v->sx * v->sy * v->sz;
    k_terrain<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(D, c->d_perm, col2, v->d_bytes);
    VXL_LAUNCH_CHECK(c);
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    VXL_CUDA(cudaFree(col2));
    v->dirty = true; v->dirty_partial = false;
    return VXL_OK;
}

int vxl_gbuffer_primary(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame) {
    if (!ctx || !vol || !view || !frame) { set_error("vxl_gbuffer_primary: bad argument"); return VXL_ERR_INVALID; }
    FrameView F;
    vxl_frame f = *frame;
    if (!f.noise) f.noise = f.depth24;   // the generator does not read noise
    if (int e = frame_view(&f, &F)) return e;
    if (!F.material) { set_error("vxl_gbuffer_primary: frame.material is NULL"); return VXL_ERR_INVALID; }
    if (F.n_tiles == 0) return VXL_OK;
    PrimK K;
    for (int i = 0; i < 16; ++i) { K.InvView[i] = view->InverseViewMatrix[i]; K.View[i] = view->ViewMatrix[i]; K.InvProj[i] = view->InverseProjectionMatrix[i]; }
    const int bpt = ((F.tile_w + 31) / 32) * ((F.tile_h + 7) / 8);
    k_gbuffer_primary<<<(unsigned)(bpt * F.n_tiles), 256, 0, ctx->stream>>>(vol_view(vol), F, K, (uint32_t*)F.depth24, (uint32_t*)F.normal, (uint32_t*)F.material);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

}  // extern "C"
