// vxl_pixel.cuh -- thread -> pixel mapping, G-buffer decode helpers and the noise / hemisphere sampling shared by the
// light-pass kernels (vxl_passes.cu) and the light-buffer resolve kernels (vxl_resolve.cu).
#pragma once
#include "vxl_internal.h"
#include "vxl_math.cuh"

#include <cstddef>

namespace vxl {

constexpr float FAR_ = 4096.0f;                     // Sources/Shaders/lib/Common.frag:13
constexpr float GOLDEN_RATIO = 2.118033988749895f;  // Common.frag:9 (sic)

constexpr int BLOCK_W = 32, BLOCK_H = 16;           // pixels per thread block (16 warps of 8x4)
constexpr int BLOCK_THREADS = BLOCK_W * BLOCK_H;
#ifndef VXL_PASS_BLOCKS
#define VXL_PASS_BLOCKS 3         // resident 512-thread blocks per SM the pass kernels' registers are capped for (<= 42 regs);
#endif                            // measured: 3 blocks 5.91 ms vs 2 blocks 6.51 ms for k_ambient on config 3 (profiles/r1f)
#ifndef VXL_AMBIENT_BLOCKS
#define VXL_AMBIENT_BLOCKS 2      // k_ambient with the pooled AO resolve: 93 KB of shared memory per block, <= 64 registers
#endif

constexpr int NOISE_OFFS = 17;                      // getNoise() and getNoise(0..15): the frame-dependent texture offsets, per launch not per ray
struct ViewK { float InvView[16], View[16], InvProj[16], Proj[16]; int Frame; float nfx[NOISE_OFFS], nfy[NOISE_OFFS]; };

struct PixelCtx {
    bool valid;
    int px, py;       // frame coordinates
    size_t idx;       // index into the tile-compact planes
    float u, v;       // In.UV
    float3 farvec;    // LightAmbient.vert:32-36, evaluated per pixel
};

// thread -> pixel of the shard.  blockIdx.x enumerates (tile, block-in-tile).
__device__ __forceinline__ PixelCtx pixel_ctx(const FrameView& F, const ViewK& K) {
    PixelCtx p;
    const int bpt_x = (F.tile_w + BLOCK_W - 1) / BLOCK_W, bpt_y = (F.rows + BLOCK_H - 1) / BLOCK_H;
    const int bpt = bpt_x * bpt_y;
    const int lt = blockIdx.x / bpt, b = blockIdx.x - lt * bpt;
    const int by = b / bpt_x, bx = b - by * bpt_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 16 warps as 4 (x) by 4 (y); each warp 8 (x) by 4 (y)
    const int lx = bx * BLOCK_W + (warp & 3) * 8 + (lane & 7);
    const int lb = by * BLOCK_H + (warp >> 2) * 4 + (lane >> 3);      // row within the band
    const int ly = F.row0 + lb;
    const int gt = F.tile_first + lt * F.tile_stride;
    const int ty = gt / F.tiles_x, tx = gt - ty * F.tiles_x;
    p.px = tx * F.tile_w + lx;
    p.py = ty * F.tile_h + ly;
    p.valid = lx < F.tile_w && lb < F.rows && ly < F.tile_h && p.px < F.width && p.py < F.height && lt < F.n_tiles;
    p.idx = ((size_t)lt * F.tile_h + ly) * F.tile_w + lx;
    p.u = ((float)p.px + 0.5f) / (float)F.width;
    p.v = ((float)p.py + 0.5f) / (float)F.height;
    const float ndcx = 2.0f * p.u - 1.0f, ndcy = 1.0f - 2.0f * p.v;
    const float4 f = mat_mul(K.InvProj, make_float4(ndcx, ndcy, 1.0f, 1.0f));
    p.farvec = make_float3(f.x / f.w, f.y / f.w, f.z / f.w);
    return p;
}

// ---- light passes: a thread block covers a REGION of the shard and its warps draw 8x4-pixel work items from it ----
// The block stages ONE occupancy tile for the whole region (vxl_passes.cu: block_prologue), so the staging, the LUTs and the tile
// placement are paid once per 2048 pixels, and a warp that finishes its item takes the next one instead of waiting at the block's
// final barrier for the slowest warp.
// Region sizes per kernel: the AO pass wants its tile tight around few pixels (32x16: one item per warp; larger regions push more
// pixels out of the window than the shared staging saves, measured r2i), the one-ray-per-pixel passes amortise the staging over
// 64x32 pixels (point 0.52 -> 0.42 ms, reflection 1.46 -> 1.30 ms on config 3).
constexpr int ITEM_W = 8, ITEM_H = 4;
template <int RW, int RH>
struct Region {
    static constexpr int W = RW, H = RH, ITEMS_X = RW / ITEM_W, ITEMS = (RW / ITEM_W) * (RH / ITEM_H);
};
#ifndef VXL_AMB_REGION_W
#define VXL_AMB_REGION_W 32
#endif
#ifndef VXL_AMB_REGION_H
#define VXL_AMB_REGION_H 16
#endif
#ifndef VXL_PASS_REGION_W
#define VXL_PASS_REGION_W 64
#endif
#ifndef VXL_PASS_REGION_H
#define VXL_PASS_REGION_H 32
#endif
typedef Region<VXL_AMB_REGION_W, VXL_AMB_REGION_H> AmbientRegion;
typedef Region<VXL_PASS_REGION_W, VXL_PASS_REGION_H> PassRegion;

struct RegionCtx { int lt, x0, y0; };               // tile of the shard, pixel offset of the region inside the tile / the row band

template <typename RG>
__device__ __forceinline__ RegionCtx region_ctx(const FrameView& F) {
    const int rx = (F.tile_w + RG::W - 1) / RG::W, ry = (F.rows + RG::H - 1) / RG::H;
    const int rpt = rx * ry;
    RegionCtx R;
    R.lt = blockIdx.x / rpt;
    const int r = blockIdx.x - R.lt * rpt, iy = r / rx;
    R.x0 = (r - iy * rx) * RG::W; R.y0 = iy * RG::H;
    return R;
}
// pixel (x, y) of the region
__device__ __forceinline__ PixelCtx region_pixel(const FrameView& F, const ViewK& K, const RegionCtx& R, int x, int y) {
    PixelCtx p;
    const int lx = R.x0 + x, lb = R.y0 + y, ly = F.row0 + lb;
    const int gt = F.tile_first + R.lt * F.tile_stride;
    const int ty = gt / F.tiles_x, tx = gt - ty * F.tiles_x;
    p.px = tx * F.tile_w + lx;
    p.py = ty * F.tile_h + ly;
    p.valid = lx < F.tile_w && lb < F.rows && ly < F.tile_h && p.px < F.width && p.py < F.height && R.lt < F.n_tiles;
    p.idx = ((size_t)R.lt * F.tile_h + ly) * F.tile_w + lx;
    p.u = ((float)p.px + 0.5f) / (float)F.width;
    p.v = ((float)p.py + 0.5f) / (float)F.height;
    const float ndcx = 2.0f * p.u - 1.0f, ndcy = 1.0f - 2.0f * p.v;
    const float4 f = mat_mul(K.InvProj, make_float4(ndcx, ndcy, 1.0f, 1.0f));
    p.farvec = make_float3(f.x / f.w, f.y / f.w, f.z / f.w);
    return p;
}
// this lane's pixel of work item `item` (a warp covers 8x4 pixels: coherent rays)
template <typename RG>
__device__ __forceinline__ PixelCtx item_pixel(const FrameView& F, const ViewK& K, const RegionCtx& R, int item) {
    const int lane = threadIdx.x & 31, iy = item / RG::ITEMS_X, ix = item - iy * RG::ITEMS_X;
    return region_pixel(F, K, R, ix * ITEM_W + (lane & 7), iy * ITEM_H + (lane >> 3));
}
template <typename RG>
static inline unsigned grid_regions(const FrameView& F) {
    const int rx = (F.tile_w + RG::W - 1) / RG::W, ry = (F.rows + RG::H - 1) / RG::H;
    return (unsigned)(rx * ry * F.n_tiles);
}
// Several GPUs: the region's finished output rows, re-read from the rank's own planes (L2) after a block barrier, are repeated into
// the same slot of every peer's copy of the gathered stack (peer-to-peer stores over NVLink; vxl_ctx_set_output_mirrors).  A warp
// takes whole rows of the region: 128-byte stores -- an 8x4 item on its own would be four 32-byte packets per plane.
template <typename RG>
__device__ __forceinline__ void mirror_region(const FrameView& F, const RegionCtx& R, float* planes, int n_planes, size_t plane_stride) {
    if (!planes) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int gt = F.tile_first + R.lt * F.tile_stride;
    const int ty = gt / F.tiles_x, tx = gt - ty * F.tiles_x;
    for (int y = warp; y < RG::H; y += nwarps)
        for (int x = lane; x < RG::W; x += 32) {
            const int lx = R.x0 + x, lb = R.y0 + y, ly = F.row0 + lb;
            if (!(lx < F.tile_w && lb < F.rows && ly < F.tile_h && tx * F.tile_w + lx < F.width && ty * F.tile_h + ly < F.height && R.lt < F.n_tiles)) continue;
            const size_t idx = ((size_t)R.lt * F.tile_h + ly) * F.tile_w + lx;
            for (int k = 0; k < n_planes; ++k) {
                float* const q = planes + (size_t)k * plane_stride + idx;
                const float v = __ldcg(q);                       // written by this block before the barrier; L2 is the point of coherence
                for (int i = 0; i < F.n_mirror; ++i) *reinterpret_cast<float*>(reinterpret_cast<char*>(q) + F.mirror[i]) = v;
            }
        }
}

// LightAmbient.frag:44-52 getNoise() (s < 0) / getNoise(int s)
__device__ __forceinline__ uint32_t get_noise(const FrameView& F, const ViewK& K, const PixelCtx& p, int s) {
    float fx, fy;
    if (s + 1 < NOISE_OFFS) { fx = K.nfx[s + 1]; fy = K.nfy[s + 1]; }      // make_viewk: the same expressions, evaluated once on the host
    else {
        fx = GOLDEN_RATIO * gmod((float)(K.Frame + s * 5), 64.0f);
        fy = GOLDEN_RATIO * gmod((float)(K.Frame + s * 7 + 1), 64.0f);
    }
    // p.u, p.v > 0 and fx, fy >= 0 (GOLDEN_RATIO times a GLSL mod by 64), so the products are non-negative and % 512 is a mask
    const int ix = f2i((p.u + fx) * (float)F.width), iy = f2i((p.v + fy) * (float)F.height);
    __builtin_assume(ix >= 0);
    __builtin_assume(iy >= 0);
    const int cx = ix % 512, cy = iy % 512;
    return __ldg(F.noise + cy * 512 + cx);
}

// LightAmbient.frag:81-87 with the cos/sin of theta = 6.283*(k/255) tabulated (host, double, rounded once)
// and r = sqrt(u), z = sqrt(max(0, 1-u)) tabulated per block over the 256 possible u = k/255 (same
// device sqrtf on the same input as the per-ray evaluation, so the values are identical).
constexpr int LUT_FLOATS = 1024;   // cos[256] sin[256] r[256] z[256]
__device__ __forceinline__ float3 cosine_sample_hemisphere(const float* __restrict__ lut, uint32_t nx, uint32_t ny) {
    const float r = lut[512 + (nx & 0xFFu)];
    const float x = r * lut[ny & 0xFFu];
    const float y = r * lut[256 + (ny & 0xFFu)];
    return make_float3(x, y, lut[768 + (nx & 0xFFu)]);
}

// (no barrier: the caller's block prologue synchronises before the first use)
__device__ __forceinline__ void load_luts(float* s_lut, const float* __restrict__ g_lut) {
    for (int i = threadIdx.x; i < LUT_FLOATS / 4; i += blockDim.x) reinterpret_cast<float4*>(s_lut)[i] = __ldg(reinterpret_cast<const float4*>(g_lut) + i);
}
// r = sqrt(u), z = sqrt(max(0, 1 - u)) over the 256 possible u = k / 255, written behind the host's cos / sin tables when the context is
// created: the same device sqrtf on the same input as a per-ray evaluation
static __global__ void k_fill_sqrt_luts(float* __restrict__ g_lut) {
    const int i = threadIdx.x;
    if (i < 256) {
        const float u = unorm8((uint32_t)i);
        g_lut[512 + i] = sqrtf(u);
        g_lut[768 + i] = sqrtf(fmaxf(0.0f, 1.0f - u));
    }
}

static inline ViewK make_viewk(const vxl_view* v) {
    ViewK k;
    for (int i = 0; i < 16; ++i) { k.InvView[i] = v->InverseViewMatrix[i]; k.View[i] = v->ViewMatrix[i]; k.InvProj[i] = v->InverseProjectionMatrix[i]; k.Proj[i] = v->ProjectionMatrix[i]; }
    k.Frame = v->Frame;
    // LightAmbient.frag:44-52: index 0 = getNoise(), 1 + s = getNoise(s); IEEE division / floor / multiply / subtract, identical on host and device
    k.nfx[0] = GOLDEN_RATIO * gmod((float)k.Frame, 16.0f);
    k.nfy[0] = GOLDEN_RATIO * gmod((float)(k.Frame + 1), 16.0f);
    for (int s = 0; s + 1 < NOISE_OFFS; ++s) {
        k.nfx[s + 1] = GOLDEN_RATIO * gmod((float)(k.Frame + s * 5), 64.0f);
        k.nfy[s + 1] = GOLDEN_RATIO * gmod((float)(k.Frame + s * 7 + 1), 64.0f);
    }
    return k;
}

// vxl_lighting_host runs the passes band by band (rows of every tile) so that copies overlap them
static inline void apply_band(const vxl_ctx* ctx, FrameView& F) {
    if (ctx->band_rows > 0) { F.row0 = ctx->band_row0; F.rows = ctx->band_rows < F.tile_h - F.row0 ? ctx->band_rows : F.tile_h - F.row0; }
}

static inline void apply_mirrors(const vxl_ctx* ctx, FrameView& F) {
    F.n_mirror = ctx->n_mirror;
    for (int i = 0; i < 15; ++i) F.mirror[i] = i < ctx->n_mirror ? ctx->mirror[i] : 0;
}

// one output value of a light pass: the rank's own plane and, on several GPUs, the same slot of every peer's gathered stack
// (peer-to-peer stores over NVLink, issued while the block's other warps are still marching: the transfer hides under the pass)
__device__ __forceinline__ void store_out(const FrameView& F, float* p, float v) {
    *p = v;
    for (int i = 0; i < F.n_mirror; ++i) *reinterpret_cast<float*>(reinterpret_cast<char*>(p) + F.mirror[i]) = v;
}

// Multi-GPU write-out.  A warp covers an 8x4 pixel patch (coherent rays), i.e. four 32-byte pieces per plane: fine for local
// HBM, poor for NVLink, where every piece is a packet.  With mirrors on, the block first parks its values in shared memory and
// writes them back row-major: one warp = one 32-pixel row = one 128-byte store per plane and per copy of the stack.
// block_pixel: plane index of the pixel at (x, y) inside this block's 32x16 rectangle (false = outside the shard).
__device__ __forceinline__ bool block_pixel(const FrameView& F, int x, int y, size_t& idx) {
    const int bpt_x = (F.tile_w + BLOCK_W - 1) / BLOCK_W, bpt_y = (F.rows + BLOCK_H - 1) / BLOCK_H;
    const int bpt = bpt_x * bpt_y;
    const int lt = blockIdx.x / bpt, b = blockIdx.x - lt * bpt;
    const int by = b / bpt_x, bx = b - by * bpt_x;
    const int lx = bx * BLOCK_W + x, lb = by * BLOCK_H + y, ly = F.row0 + lb;
    const int gt = F.tile_first + lt * F.tile_stride;
    const int ty = gt / F.tiles_x, tx = gt - ty * F.tiles_x;
    idx = ((size_t)lt * F.tile_h + ly) * F.tile_w + lx;
    return lx < F.tile_w && lb < F.rows && ly < F.tile_h && tx * F.tile_w + lx < F.width && ty * F.tile_h + ly < F.height && lt < F.n_tiles;
}
// slot of this thread's own pixel in a 16x32 row-major staging plane (the inverse of pixel_ctx's warp layout)
__device__ __forceinline__ int own_slot() {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return ((warp >> 2) * 4 + (lane >> 3)) * BLOCK_W + (warp & 3) * 8 + (lane & 7);
}
// NPL <= 2 planes through `stage` (>= NPL * 512 floats of shared memory no warp still reads: the caller's barrier comes first)
template <int NPL>
__device__ __forceinline__ void store_rows(const FrameView& F, float* stage, float* const (&planes)[NPL], const float (&vals)[NPL]) {
    const int slot = own_slot();
#pragma unroll
    for (int k = 0; k < NPL; ++k) stage[k * BLOCK_THREADS + slot] = vals[k];
    __syncthreads();
    size_t idx;
    if (block_pixel(F, (int)(threadIdx.x & 31), (int)(threadIdx.x >> 5), idx)) {
#pragma unroll
        for (int k = 0; k < NPL; ++k)
            if (planes[k]) store_out(F, planes[k] + idx, stage[k * BLOCK_THREADS + threadIdx.x]);
    }
}

static inline unsigned grid_for(const FrameView& F) {
    const int bpt_x = (F.tile_w + BLOCK_W - 1) / BLOCK_W, bpt_y = (F.rows + BLOCK_H - 1) / BLOCK_H;
    return (unsigned)(bpt_x * bpt_y * F.n_tiles);
}

}  // namespace vxl
