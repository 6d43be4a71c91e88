// vxl_api.cu -- context, error reporting, memory helpers, counters and the whole-frame host
// drop-in of libvxl.so.  No CPU fallback anywhere: every entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "vxl_internal.h"
#include "vxl_math.cuh"
#include "vxl_pixel.cuh"

namespace vxl {

static thread_local std::string g_err = "";

void set_error(const std::string& s) { g_err = s; }

int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return e == cudaErrorMemoryAllocation ? VXL_ERR_OOM : VXL_ERR_CUDA;
}

int frame_view(const vxl_frame* f, FrameView* out) {
    if (f->width <= 0 || f->height <= 0 || f->tile_w <= 0 || f->tile_h <= 0 || f->n_tiles < 0 || f->tile_first < 0 ||
        f->tile_stride <= 0 || !f->depth24 || !f->normal || !f->noise) {
        set_error("invalid vxl_frame (sizes must be positive; depth24/normal/noise non-NULL)");
        return VXL_ERR_INVALID;
    }
    const int tiles_x = (f->width + f->tile_w - 1) / f->tile_w, tiles_y = (f->height + f->tile_h - 1) / f->tile_h;
    if (f->n_tiles > 0 && (long long)f->tile_first + (long long)(f->n_tiles - 1) * f->tile_stride >= (long long)tiles_x * tiles_y) {
        set_error("vxl_frame: tile range exceeds the frame's tile grid");
        return VXL_ERR_INVALID;
    }
    out->width = f->width; out->height = f->height; out->tile_w = f->tile_w; out->tile_h = f->tile_h;
    out->tile_first = f->tile_first; out->tile_stride = f->tile_stride; out->n_tiles = f->n_tiles; out->tiles_x = tiles_x;
    out->row0 = 0; out->rows = f->tile_h;
    out->n_mirror = 0;
    for (int i = 0; i < 15; ++i) out->mirror[i] = 0;
    out->depth24 = f->depth24; out->normal = f->normal; out->material = f->material; out->noise = f->noise;
    return VXL_OK;
}

size_t frame_pixels(const vxl_frame* f) { return (size_t)f->n_tiles * (size_t)f->tile_w * (size_t)f->tile_h; }

// Perm tables of the terrain generator: the published FastNoise 0.4 seeding scheme
// (Vendor/FastNoise/FastNoise.cpp:197-215), restated.
static void build_perm(int seed, uint8_t* p, uint8_t* p12) {
    std::mt19937_64 gen((unsigned long long)seed);
    for (int i = 0; i < 256; i++) p[i] = (uint8_t)i;
    for (int j = 0; j < 256; j++) {
        const int k = (int)(gen() % (uint64_t)(256 - j)) + j;
        const uint8_t l = p[j];
        p[j] = p[j + 256] = p[k];
        p[k] = l;
        p12[j] = p12[j + 256] = (uint8_t)(p[j] % 12);
    }
}

// vxl_debug_read_bandwidth: grid-stride 16-byte loads, 8 independent loads in flight per thread
__global__ void __launch_bounds__(512) k_read_bw(const uint4* __restrict__ buf, size_t n_vec, int reps, unsigned* __restrict__ sink) {
    unsigned acc = 0u;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7 * stride < n_vec; i += 8 * stride) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(buf + i + (size_t)u * stride);      // .cg: L2 only, so the figure is not an L1 figure
#pragma unroll
            for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
        for (; i < n_vec; i += stride) { const uint4 v = __ldcg(buf + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    }
    if (acc == 0x9E3779B9u) *sink = acc;          // never true for the fill pattern; keeps the loads alive
}

// spec_t takes 180 values (Light.frag:131-173 with dist = 256: d = 0.5 (k + 1) for k < 31, 16 + j for j < 148, or 256 on a miss):
// code 0..30 -> 0.5 (code + 1); 31..178 -> 16 + (code - 31); 255 -> 256.  254 = a value outside that set (never produced; the decoder rejects it)
__device__ __forceinline__ unsigned spec_code_of(float t) {
    if (t == 256.0f) return 255u;
    if (t >= 16.0f) { const float j = t - 16.0f; return (j < 148.0f && j == floorf(j)) ? 31u + (unsigned)j : 254u; }
    const float k = t * 2.0f - 1.0f;
    return (k >= 0.0f && k < 31.0f && k == floorf(k)) ? (unsigned)k : 254u;
}

// Pack rows [row0, row0 + rows) of every tile: the 0/1 shadow planes (sun, point lights, spot lights; plane_stride pixels apart)
// into mask_bytes bytes per pixel, bit p = plane p is lit (1.0f), and spec_t into its one-byte code.
__global__ void __launch_bounds__(256) k_pack_planes(const float* __restrict__ shadow, const float* __restrict__ point, int n_point,
                                                     const float* __restrict__ spot, int n_spot, size_t plane_stride,
                                                     const float* __restrict__ spec, uint8_t* __restrict__ mask, int mask_bytes,
                                                     uint8_t* __restrict__ code, int tile_w, int tile_h, int row0, int rows, int n_tiles,
                                                     int width, int height, int tile_first, int tile_stride) {
    const size_t band = (size_t)rows * tile_w;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= band * (size_t)n_tiles) return;
    const size_t t = i / band, r = i - t * band;
    const size_t idx = t * ((size_t)tile_w * tile_h) + (size_t)row0 * tile_w + r;
    // pixels of an edge tile that lie beyond the frame are not written by the passes: they pack as "lit" / "miss"
    const int tiles_x = (width + tile_w - 1) / tile_w, gt = tile_first + (int)t * tile_stride;
    const int py = (gt / tiles_x) * tile_h + row0 + (int)(r / (size_t)tile_w), pxl = (gt % tiles_x) * tile_w + (int)(r % (size_t)tile_w);
    if (pxl >= width || py >= height) {
        if (mask) for (int byte = 0; byte < mask_bytes; ++byte) mask[idx * (size_t)mask_bytes + byte] = 0;
        if (code) code[idx] = 255;
        return;
    }
    if (mask) {
        const int n_planes = 1 + n_point + n_spot;
        for (int byte = 0; byte < mask_bytes; ++byte) {
            unsigned m = 0u;
            for (int bit = 0; bit < 8; ++bit) {
                const int pl = byte * 8 + bit;
                if (pl >= n_planes) break;
                float v;
                if (pl == 0) v = shadow ? shadow[idx] : 1.0f;
                else if (pl <= n_point) v = point[(size_t)(pl - 1) * plane_stride + idx];
                else v = spot[(size_t)(pl - 1 - n_point) * plane_stride + idx];
                if (v != 0.0f) m |= 1u << bit;
            }
            mask[idx * (size_t)mask_bytes + byte] = (uint8_t)m;
        }
    }
    if (code) code[idx] = (uint8_t)spec_code_of(spec[idx]);
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_abi_version(void) { return VXL_ABI_VERSION; }

int vxl_debug_read_bandwidth(vxl_ctx* c, size_t bytes, int reps, double* out_gbs) {
    if (!c || !out_gbs || bytes < (1u << 20) || reps < 1) { set_error("vxl_debug_read_bandwidth: bad argument"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(c->device));
    uint4* buf = nullptr;
    unsigned* sink = nullptr;
    const size_t n_vec = bytes / sizeof(uint4);
    VXL_CUDA(cudaMalloc(&buf, n_vec * sizeof(uint4) + 16));
    sink = reinterpret_cast<unsigned*>(buf + n_vec);
    cudaEvent_t a = nullptr, b = nullptr;
    int rc = VXL_OK, sms = 0;
    cudaError_t e = cudaMemsetAsync(buf, 0x5A, n_vec * sizeof(uint4) + 16, c->stream);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    if (e == cudaSuccess) e = cudaEventCreate(&a);
    if (e == cudaSuccess) e = cudaEventCreate(&b);
    if (e == cudaSuccess) {
        const unsigned grid = (unsigned)sms * 4u;
        k_read_bw<<<grid, 512, 0, c->stream>>>(buf, n_vec, 1, sink);             // warm-up: brings a small buffer into L2
        cudaEventRecord(a, c->stream);
        k_read_bw<<<grid, 512, 0, c->stream>>>(buf, n_vec, reps, sink);
        cudaEventRecord(b, c->stream);
        c->launches += 2;
        e = cudaEventSynchronize(b);
        float ms = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
        if (e == cudaSuccess) *out_gbs = (double)(n_vec * sizeof(uint4)) * reps / ((double)ms * 1e6);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "vxl_debug_read_bandwidth");
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    cudaFree(buf);
    return rc;
}
const char* vxl_last_error_string(void) { return g_err.c_str(); }

int vxl_ctx_create(int device, vxl_ctx** out) {
    if (!out) { set_error("vxl_ctx_create: out is NULL"); return VXL_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
        return VXL_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("vxl_ctx_create: device index out of range"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(device));
    vxl_ctx* c = new vxl_ctx();
    c->device = device;
    if (const char* ev = getenv("VXL_VARIANT")) c->variant = atoi(ev);
    VXL_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    VXL_CUDA(cudaMalloc(&c->d_stats, STAT_SLOTS * 4 * sizeof(unsigned long long)));
    VXL_CUDA(cudaMemset(c->d_stats, 0, STAT_SLOTS * 4 * sizeof(unsigned long long)));
    // cos/sin of theta = 6.283 * (k/255) in double, rounded once (LightAmbient.frag:81-87)
    float lut[512];
    for (int k = 0; k < 256; ++k) {
        const float v = (float)k / 255.0f;
        const float theta = 6.283f * v;
        lut[k] = (float)cos((double)theta);
        lut[256 + k] = (float)sin((double)theta);
    }
    VXL_CUDA(cudaMalloc(&c->d_luts, LUT_FLOATS * sizeof(float)));
    VXL_CUDA(cudaMemcpy(c->d_luts, lut, sizeof lut, cudaMemcpyHostToDevice));
    k_fill_sqrt_luts<<<1, 256, 0, c->stream>>>(c->d_luts);      // r[256] z[256]: the device's own sqrtf, once per context instead of once per block
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    VXL_CUDA(cudaMalloc(&c->d_lights, VXL_MAX_LIGHTS * sizeof(vxl_spot_light)));
    uint8_t perm[1024];
    build_perm(1337, perm, perm + 512);
    VXL_CUDA(cudaMalloc(&c->d_perm, sizeof perm));
    VXL_CUDA(cudaMemcpy(c->d_perm, perm, sizeof perm, cudaMemcpyHostToDevice));
    *out = c;
    return VXL_OK;
}

int vxl_ctx_destroy(vxl_ctx* c) {
    if (!c) return VXL_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (void* p : c->ipc_open) cudaIpcCloseMemHandle(p);
    cudaFree(c->d_stats); cudaFree(c->d_luts); cudaFree(c->d_taa_lut); cudaFree(c->d_taa_depth); cudaFree(c->d_lights); cudaFree(c->d_perm);
    for (auto& m : c->models) { cudaFree((void*)m.voxels); cudaFree((void*)m.mip1); cudaFree((void*)m.mip2); }
    cudaFree(c->d_models); cudaFree(c->d_draws); cudaFree(c->d_hkeys); cudaFree(c->d_hvals); cudaFree(c->d_ents); cudaFree(c->d_aabb);
    cudaFree(c->h_planes); cudaFree(c->h_out); cudaFree(c->h_noise);
    if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_d2h); }
    if (c->s_side[0]) { cudaStreamDestroy(c->s_side[0]); cudaStreamDestroy(c->s_side[1]); }
    for (auto& e : c->ev) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return VXL_OK;
}

int vxl_ctx_set_stream(vxl_ctx* c, void* s) {
    if (!c) { set_error("vxl_ctx_set_stream: ctx is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)s;
    return VXL_OK;
}

// ---- peer memory: one process per GPU on one node ---------------------------------------------------------------
int vxl_ipc_export(vxl_ctx* c, void* dev, vxl_ipc_handle* out) {
    if (!c || !dev || !out) { set_error("vxl_ipc_export: bad argument"); return VXL_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(vxl_ipc_handle), "vxl_ipc_handle must hold a cudaIpcMemHandle_t");
    VXL_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    VXL_CUDA(cudaIpcGetMemHandle(&h, dev));
    memcpy(out->bytes, &h, sizeof h);
    return VXL_OK;
}

int vxl_ipc_open(vxl_ctx* c, const vxl_ipc_handle* handle, void** out_dev) {
    if (!c || !handle || !out_dev) { set_error("vxl_ipc_open: bad argument"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof h);
    void* p = nullptr;
    VXL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_open.push_back(p);
    *out_dev = p;
    return VXL_OK;
}

int vxl_ipc_close(vxl_ctx* c, void* dev) {
    if (!c || !dev) { set_error("vxl_ipc_close: bad argument"); return VXL_ERR_INVALID; }
    for (size_t i = 0; i < c->ipc_open.size(); ++i)
        if (c->ipc_open[i] == dev) {
            c->ipc_open.erase(c->ipc_open.begin() + (long)i);
            VXL_CUDA(cudaStreamSynchronize(c->stream));
            VXL_CUDA(cudaIpcCloseMemHandle(dev));
            return VXL_OK;
        }
    set_error("vxl_ipc_close: not a mapping opened by this context");
    return VXL_ERR_INVALID;
}

int vxl_ctx_set_output_mirrors(vxl_ctx* c, int n, const int64_t* byte_deltas) {
    if (!c || n < 0 || n > VXL_MAX_MIRRORS || (n > 0 && !byte_deltas)) { set_error("vxl_ctx_set_output_mirrors: bad argument (at most VXL_MAX_MIRRORS mirrors)"); return VXL_ERR_INVALID; }
    for (int i = 0; i < n; ++i)
        if (byte_deltas[i] % 4 != 0) { set_error("vxl_ctx_set_output_mirrors: deltas must be multiples of 4 bytes"); return VXL_ERR_INVALID; }
    c->n_mirror = n;
    for (int i = 0; i < n; ++i) c->mirror[i] = (long long)byte_deltas[i];
    return VXL_OK;
}

int vxl_ctx_set_light_plane_stride(vxl_ctx* c, uint64_t pixels) {
    if (!c) { set_error("vxl_ctx_set_light_plane_stride: ctx is NULL"); return VXL_ERR_INVALID; }
    c->light_plane_stride = (size_t)pixels;
    return VXL_OK;
}

int vxl_sync(vxl_ctx* c) {
    if (!c) { set_error("vxl_sync: ctx is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    return VXL_OK;
}

int vxl_stats_reset(vxl_ctx* c) {
    if (!c) { set_error("vxl_stats_reset: ctx is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemsetAsync(c->d_stats, 0, STAT_SLOTS * 4 * sizeof(unsigned long long), c->stream));
    return VXL_OK;
}

int vxl_stats_read(vxl_ctx* c, vxl_stats* out) {
    if (!c || !out) { set_error("vxl_stats_read: bad argument"); return VXL_ERR_INVALID; }
    unsigned long long h[STAT_SLOTS * 4];
    VXL_CUDA(cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    out->rays = out->steps = out->pixels = 0;
    for (int i = 0; i < STAT_SLOTS; ++i) { out->rays += h[i * 4]; out->steps += h[i * 4 + 1]; out->pixels += h[i * 4 + 2]; }
    return VXL_OK;
}

int vxl_debug_set_variant(vxl_ctx* c, int variant) {
    if (!c || variant < 0 || variant > 2) { set_error("vxl_debug_set_variant: bad argument"); return VXL_ERR_INVALID; }
    c->variant = variant;
    return VXL_OK;
}

int vxl_debug_fetched_probes(vxl_ctx* c, uint64_t* out) {
    if (!c || !out) { set_error("vxl_debug_fetched_probes: bad argument"); return VXL_ERR_INVALID; }
    unsigned long long h[STAT_SLOTS * 4];
    VXL_CUDA(cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    *out = 0;
    for (int i = 0; i < STAT_SLOTS; ++i) *out += h[i * 4 + 3];
    return VXL_OK;
}

int vxl_launch_count(vxl_ctx* c, uint64_t* out) {
    if (!c || !out) { set_error("vxl_launch_count: bad argument"); return VXL_ERR_INVALID; }
    *out = c->launches;
    return VXL_OK;
}

int vxl_malloc(vxl_ctx* c, size_t bytes, void** out) {
    if (!c || !out) { set_error("vxl_malloc: bad argument"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(c->device));
    VXL_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return VXL_OK;
}
int vxl_free(vxl_ctx* c, void* dev) {
    if (!c) { set_error("vxl_free: ctx is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    VXL_CUDA(cudaFree(dev));
    return VXL_OK;
}
int vxl_host_alloc(size_t bytes, void** out) {
    if (!out) { set_error("vxl_host_alloc: out is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return VXL_OK;
}
int vxl_host_free(void* p) {
    VXL_CUDA(cudaFreeHost(p));
    return VXL_OK;
}
int vxl_memcpy_h2d(vxl_ctx* c, void* dev, const void* host, size_t bytes) {
    if (!c) { set_error("vxl_memcpy_h2d: ctx is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return VXL_OK;
}
int vxl_memcpy_d2h(vxl_ctx* c, void* host, const void* dev, size_t bytes) {
    if (!c) { set_error("vxl_memcpy_d2h: ctx is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    return VXL_OK;
}
int vxl_memset(vxl_ctx* c, void* dev, int value, size_t bytes) {
    if (!c) { set_error("vxl_memset: ctx is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemsetAsync(dev, value, bytes, c->stream));
    return VXL_OK;
}

// -------------------------------------------------------------------------------------------------
// Whole-frame host drop-in: H2D of the shard's G-buffer, the four passes, D2H of the planes.
// Replaces the "Lights" and "Reflection" blocks of WorldRenderer::DrawWorld
// (Sources/Graphics/Renderer/WorldRenderer.cpp:239-260,269-274) for a caller whose G-buffer lives
// on the host side of the boundary.
// -------------------------------------------------------------------------------------------------
static int ensure(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return VXL_OK;
    if (*p) { VXL_CUDA(cudaFree(*p)); *p = nullptr; *cap = 0; }
    VXL_CUDA(cudaMalloc(p, need));
    *cap = need;
    return VXL_OK;
}

// All requested passes of one frame with DEVICE planes, the three kernels side by side: the ambient pass on the context's stream,
// the local-light passes and the reflection pass on two side streams forked from it and joined back before the call returns
// (stream-ordered; nothing blocks).  The passes are independent (same inputs, disjoint output planes), so a kernel's last,
// partially filled wave of blocks overlaps the next kernel's first -- which matters when a rank's share of a frame is only a few
// waves (8 GPUs: 4.6 waves per kernel).  Mirrors and the light plane stride apply as in the single passes.
int vxl_lighting(vxl_ctx* c, vxl_volume* vol, const vxl_lighting_host_args* a) {
    if (!c || !vol || !a || !a->view) { set_error("vxl_lighting: bad argument"); return VXL_ERR_INVALID; }
    if (a->n_point < 0 || a->n_spot < 0 || a->n_point > VXL_MAX_LIGHTS || a->n_spot > VXL_MAX_LIGHTS) { set_error("vxl_lighting: light count out of range"); return VXL_ERR_LIMIT; }
    const bool want_amb = a->out_shadow || a->out_ao;
    const bool want_pt = a->n_point > 0 && a->out_point_shadow;
    const bool want_sp = a->n_spot > 0 && a->out_spot_shadow;
    const bool want_rf = a->out_spec_t != nullptr;
    VXL_CUDA(cudaSetDevice(c->device));
    if (vol->dirty) { if (int e = vxl_volume_build_occupancy(vol)) return e; }       // once, before the fork
    if (!c->s_side[0]) {
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_side[0], cudaStreamNonBlocking));
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_side[1], cudaStreamNonBlocking));
    }
    while (c->ev.size() < 3) { cudaEvent_t e; VXL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev.push_back(e); }
    cudaEvent_t* ev = c->ev.data();
    cudaStream_t main = c->stream, side[2] = {c->s_side[0], c->s_side[1]};
    struct Restore { vxl_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{c, main};
    VXL_CUDA(cudaEventRecord(ev[0], main));
    int rc = VXL_OK;
    if (want_amb) rc = vxl_pass_ambient(c, vol, a->view, &a->frame, a->n_ao, a->out_shadow, a->out_ao);   // first in the queue: its blocks fill the machine first
    if (rc == VXL_OK && (want_pt || want_sp)) {
        VXL_CUDA(cudaStreamWaitEvent(side[0], ev[0], 0));
        c->stream = side[0];                                                        // point then spot: they share the light staging buffer
        if (want_pt) rc = vxl_pass_point(c, vol, a->view, &a->frame, a->point, a->n_point, a->out_point_shadow);
        if (rc == VXL_OK && want_sp) rc = vxl_pass_spot(c, vol, a->view, &a->frame, a->spot, a->n_spot, a->out_spot_shadow);
        c->stream = main;
        VXL_CUDA(cudaEventRecord(ev[1], side[0]));
        VXL_CUDA(cudaStreamWaitEvent(main, ev[1], 0));
    }
    if (rc == VXL_OK && want_rf) {
        VXL_CUDA(cudaStreamWaitEvent(side[1], ev[0], 0));
        c->stream = side[1];
        rc = vxl_pass_reflection(c, vol, a->view, &a->frame, a->out_spec_t);
        c->stream = main;
        VXL_CUDA(cudaEventRecord(ev[2], side[1]));
        VXL_CUDA(cudaStreamWaitEvent(main, ev[2], 0));
    }
    return rc;
}

// Both host drop-ins.  packed == nullptr: float planes out (vxl_lighting_host); else the packed planes (vxl_lighting_host_packed).
static int lighting_host_impl(vxl_ctx* c, vxl_volume* vol, const vxl_lighting_host_args* a, const vxl_packed_planes* packed) {
    const char* const who = packed ? "vxl_lighting_host_packed" : "vxl_lighting_host";
    if (!c || !vol || !a || !a->view) { set_error(std::string(who) + ": bad argument"); return VXL_ERR_INVALID; }
    if (a->n_point < 0 || a->n_spot < 0 || a->n_point > VXL_MAX_LIGHTS || a->n_spot > VXL_MAX_LIGHTS) { set_error(std::string(who) + ": light count out of range"); return VXL_ERR_LIMIT; }
    FrameView Fh;
    if (int e = frame_view(&a->frame, &Fh)) return e;
    const size_t px = frame_pixels(&a->frame);
    if (px == 0) return VXL_OK;
    const bool want_mask = packed && packed->shadow_mask;
    const bool want_sun = packed ? want_mask : a->out_shadow != nullptr;
    const bool want_ao = packed ? packed->ao != nullptr : a->out_ao != nullptr;
    const bool want_amb = want_sun || want_ao;
    const bool want_pt = a->n_point > 0 && (packed ? want_mask : a->out_point_shadow != nullptr);
    const bool want_sp = a->n_spot > 0 && (packed ? want_mask : a->out_spot_shadow != nullptr);
    const bool want_rf = packed ? packed->spec_code != nullptr : a->out_spec_t != nullptr;
    if (want_rf && !a->frame.material) { set_error(std::string(who) + ": spec pass needs frame.material"); return VXL_ERR_INVALID; }
    if ((want_pt && !a->point) || (want_sp && !a->spot)) { set_error(std::string(who) + ": light list is NULL"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(c->device));
    if (!c->s_h2d) {
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    }
    if (!c->s_side[0]) {
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_side[0], cudaStreamNonBlocking));
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_side[1], cudaStreamNonBlocking));
    }
    // The host drop-in writes to its own staging planes (no mirrored stores, default plane stride) and walks the frame in row
    // bands.  Whatever path leaves this function -- including a CUDA error in the middle of the band loop -- the context gets its
    // mirror / stride / band state back and the two copy streams are joined to the context's stream again.
    struct Restore {
        vxl_ctx* c; int n; size_t st; cudaStream_t main; cudaEvent_t join[4] = {nullptr, nullptr, nullptr, nullptr};
        ~Restore() {
            c->n_mirror = n; c->light_plane_stride = st; c->band_row0 = 0; c->band_rows = 0; c->stream = main;
            cudaStream_t side[4] = {c->s_h2d, c->s_d2h, c->s_side[0], c->s_side[1]};
            for (int i = 0; i < 4; ++i)
                if (join[i] && cudaEventRecord(join[i], side[i]) == cudaSuccess) cudaStreamWaitEvent(main, join[i], 0);
        }
    } restore{c, c->n_mirror, c->light_plane_stride, c->stream};
    c->n_mirror = 0; c->light_plane_stride = 0;
    if (int e = ensure((void**)&c->h_planes, &c->h_planes_bytes, px * 4 * 3)) return e;
    const size_t n_out = 3 + (size_t)a->n_point + (size_t)a->n_spot;
    const int n_planes = 1 + a->n_point + a->n_spot, mask_bytes = (n_planes + 7) / 8;
    // float planes, then (packed mode) the mask and code planes
    if (int e = ensure((void**)&c->h_out, &c->h_out_bytes, px * 4 * n_out + (packed ? px * (size_t)(mask_bytes + 1) : 0))) return e;
    if (!c->h_noise) VXL_CUDA(cudaMalloc(&c->h_noise, 512 * 512 * 4));
    uint32_t* d_depth = c->h_planes; uint32_t* d_normal = d_depth + px; uint32_t* d_mat = d_normal + px;
    // Streams: uploads, read-backs, and three for the passes (the context's stream for the ambient pass, two side streams for the
    // local-light and reflection passes: the kernels are independent, so one's tail overlaps another's head).  The frame goes
    // through in NB row bands (rows of every tile of the shard).  PCIe is full duplex and the copy engines run beside the SMs,
    // so band b+1 uploads and band b-1 reads back while band b is in the passes.
    // A whole frame on one GPU is ONE tile: its bands are row ranges.  A rank's shard of a sharded frame is many small tiles: its
    // bands are ranges of tiles (contiguous in the tile-compact planes).
    const int nt = a->frame.n_tiles;
    const bool by_tiles = a->frame.tile_h < 256 && nt >= 8;
    // bands of a whole frame: one per 800 k pixels, at most 16 -- 3 at 1080p, 10 at 4K
    // (measured at 4K with equal bands, float / packed planes: 4 bands 7.76 / 7.15 ms, 8: 7.03 / 6.74, 12: 6.86 / 6.71, 16: 6.84 / 6.79, 24: 7.22 / 7.55)
    int NB = by_tiles ? 4 : (a->frame.tile_h >= 256 ? (int)std::max<size_t>(3, std::min<size_t>(16, px / 800000)) : 1);
    if (const char* nb = getenv("VXL_HOST_BANDS")) { if (*nb) NB = std::max(1, std::min(64, atoi(nb))); }   // tuning knob
    if (by_tiles) NB = std::min(NB, nt);
    const int band_h = ((a->frame.tile_h + NB - 1) / NB + 15) / 16 * 16;          // whole 16-row thread blocks
    const int band_t = (nt + NB - 1) / NB;
    // Row bands of a whole frame: `starts[b]` .. `starts[b + 1]`, heights 1 : 2 : 4 : ... : 4 : 2 : 1 from six bands on (measured at 4K,
    // packed / float planes: 12 equal bands 6.60 / 6.77 ms, ten ramped bands 6.46 / 6.72 ms; five or six bands with one large middle
    // 6.49-6.52 / 7.0-7.5 ms).  VXL_HOST_SCHED="w0,w1,..." overrides the relative heights (tuning knob).
    std::vector<int> starts;
    if (!by_tiles) {
        std::vector<double> w;
        if (const char* sc = getenv("VXL_HOST_SCHED")) {
            for (const char* q = sc; *q;) { char* e = nullptr; const double v = strtod(q, &e); if (e == q) break; if (v > 0) w.push_back(v); q = (*e == ',') ? e + 1 : e; }
        }
        if (w.empty() && NB >= 6) {                 // 1, 2, 4, ..., 4, 2, 1: a short first band starts the passes early, a short last one ends the read-back early
            w.assign((size_t)NB, 4.0);
            w[0] = w[NB - 1] = 1.0; w[1] = w[NB - 2] = 2.0;
        }
        if (!w.empty()) {
            double tot = 0; for (double v : w) tot += v;
            const int units = (a->frame.tile_h + 15) / 16;
            double acc = 0; starts.push_back(0);
            for (size_t i = 0; i + 1 < w.size(); ++i) { acc += w[i]; const int u = std::min(units, std::max(starts.back() / 16 + 1, (int)(acc / tot * units + 0.5))); starts.push_back(u * 16); }
            starts.push_back(a->frame.tile_h);
            NB = (int)starts.size() - 1;
        }
    }
    const size_t n_ev = 6 + (size_t)NB * 6;
    while (c->ev.size() < n_ev) { cudaEvent_t e; VXL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev.push_back(e); }
    cudaEvent_t* ev = c->ev.data();
    const size_t tile_px = (size_t)a->frame.tile_w * a->frame.tile_h;
    int r0 = 0, rows = a->frame.tile_h, t0 = 0, tn = nt;                          // the current band
    // the current band of a tile-compact plane of `el`-byte pixels: rows [r0, r0 + rows) of every tile (nt pieces, tile_px apart), or
    // the tiles [t0, t0 + tn) (one piece)
    auto copy_band = [&](void* dst, const void* src, size_t el, cudaMemcpyKind kind, cudaStream_t st) -> cudaError_t {
        if (by_tiles) { const size_t off = (size_t)t0 * tile_px * el; return cudaMemcpyAsync((char*)dst + off, (const char*)src + off, (size_t)tn * tile_px * el, kind, st); }
        const size_t off = (size_t)r0 * a->frame.tile_w * el;
        if (nt == 1) return cudaMemcpyAsync((char*)dst + off, (const char*)src + off, (size_t)rows * a->frame.tile_w * el, kind, st);
        return cudaMemcpy2DAsync((char*)dst + off, tile_px * el, (const char*)src + off, tile_px * el, (size_t)rows * a->frame.tile_w * el, (size_t)nt, kind, st);
    };
    VXL_CUDA(cudaEventRecord(ev[0], c->stream));                          // order behind whatever the caller queued (voxelise, ...)
    VXL_CUDA(cudaStreamWaitEvent(c->s_h2d, ev[0], 0));
    VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, ev[0], 0));
    VXL_CUDA(cudaStreamWaitEvent(c->s_side[0], ev[0], 0));
    VXL_CUDA(cudaStreamWaitEvent(c->s_side[1], ev[0], 0));
    restore.join[0] = ev[2]; restore.join[1] = ev[3]; restore.join[2] = ev[4]; restore.join[3] = ev[5];   // from here on the side streams carry work
    cudaStream_t const main = c->stream, sl = c->s_side[0], sr = c->s_side[1];
    VXL_CUDA(cudaMemcpyAsync(c->h_noise, a->frame.noise, 512 * 512 * 4, cudaMemcpyHostToDevice, c->s_h2d));
    float* o_shadow = c->h_out; float* o_ao = o_shadow + px; float* o_spec = o_ao + px;
    float* o_pt = o_spec + px; float* o_sp = o_pt + px * (size_t)a->n_point;
    uint8_t* o_mask = reinterpret_cast<uint8_t*>(o_sp + px * (size_t)a->n_spot); uint8_t* o_code = o_mask + px * (size_t)mask_bytes;
    if (by_tiles) c->light_plane_stride = px;                             // a band's light planes sit inside the shard's
    for (int b = 0; b < NB; ++b) {
        if (by_tiles) { t0 = b * band_t; tn = std::min(band_t, nt - t0); if (tn <= 0) break; }
        else if (!starts.empty()) { r0 = starts[b]; rows = std::min(starts[b + 1], a->frame.tile_h) - r0; if (rows <= 0) continue; }
        else { r0 = b * band_h; rows = std::min(band_h, a->frame.tile_h - r0); if (rows <= 0) break; }
        const size_t bo = by_tiles ? (size_t)t0 * tile_px : 0;            // where the band starts in every plane
        vxl_frame fd = a->frame;
        fd.depth24 = d_depth + bo; fd.normal = d_normal + bo; fd.material = d_mat + bo; fd.noise = c->h_noise;
        if (by_tiles) { fd.tile_first = a->frame.tile_first + t0 * a->frame.tile_stride; fd.n_tiles = tn; }
        cudaEvent_t* eb = ev + 6 + (size_t)b * 6;
        VXL_CUDA(copy_band(d_depth, a->frame.depth24, 4, cudaMemcpyHostToDevice, c->s_h2d));
        VXL_CUDA(copy_band(d_normal, a->frame.normal, 4, cudaMemcpyHostToDevice, c->s_h2d));
        if (want_rf) VXL_CUDA(copy_band(d_mat, a->frame.material, 4, cudaMemcpyHostToDevice, c->s_h2d));
        VXL_CUDA(cudaEventRecord(eb[0], c->s_h2d));
        VXL_CUDA(cudaStreamWaitEvent(main, eb[0], 0));
        if (!by_tiles) { c->band_row0 = r0; c->band_rows = rows; }
        if (want_amb) {                                                   // first in the queue: its blocks fill the machine first
            if (int e = vxl_pass_ambient(c, vol, a->view, &fd, a->n_ao, want_sun ? o_shadow + bo : nullptr, want_ao ? o_ao + bo : nullptr)) return e;
            VXL_CUDA(cudaEventRecord(eb[3], main)); VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, eb[3], 0));
            if (!packed && a->out_shadow) VXL_CUDA(copy_band(a->out_shadow, o_shadow, 4, cudaMemcpyDeviceToHost, c->s_d2h));
            if (want_ao) VXL_CUDA(copy_band(packed ? packed->ao : a->out_ao, o_ao, 4, cudaMemcpyDeviceToHost, c->s_d2h));
        }
        if (want_pt || want_sp) {
            VXL_CUDA(cudaStreamWaitEvent(sl, eb[0], 0));
            c->stream = sl;                                               // point then spot: they share the light staging buffer
            if (want_pt) {
                if (int e = vxl_pass_point(c, vol, a->view, &fd, a->point, a->n_point, o_pt + bo)) return e;
                if (!packed) {
                    VXL_CUDA(cudaEventRecord(eb[1], sl)); VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, eb[1], 0));
                    for (int l = 0; l < a->n_point; ++l) VXL_CUDA(copy_band(a->out_point_shadow + px * (size_t)l, o_pt + px * (size_t)l, 4, cudaMemcpyDeviceToHost, c->s_d2h));
                }
            }
            if (want_sp) {
                if (int e = vxl_pass_spot(c, vol, a->view, &fd, a->spot, a->n_spot, o_sp + bo)) return e;
                if (!packed) {
                    VXL_CUDA(cudaEventRecord(eb[2], sl)); VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, eb[2], 0));
                    for (int l = 0; l < a->n_spot; ++l) VXL_CUDA(copy_band(a->out_spot_shadow + px * (size_t)l, o_sp + px * (size_t)l, 4, cudaMemcpyDeviceToHost, c->s_d2h));
                }
            }
            c->stream = main;
            if (packed) { VXL_CUDA(cudaEventRecord(eb[1], sl)); VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, eb[1], 0)); }
        }
        if (want_rf) {
            VXL_CUDA(cudaStreamWaitEvent(sr, eb[0], 0));
            c->stream = sr;
            const int e = vxl_pass_reflection(c, vol, a->view, &fd, o_spec + bo);
            c->stream = main;
            if (e) return e;
            if (!packed) {
                VXL_CUDA(cudaEventRecord(eb[4], sr)); VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, eb[4], 0));
                VXL_CUDA(copy_band(a->out_spec_t, o_spec, 4, cudaMemcpyDeviceToHost, c->s_d2h));
            } else { VXL_CUDA(cudaEventRecord(eb[4], sr)); VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, eb[4], 0)); }
        }
        if (packed && (want_mask || want_rf)) {                           // on the read-back stream, behind all three pass streams (joined into it above):
                                                                          // the next band's passes do not wait for this band's packing
            const int prow0 = by_tiles ? 0 : r0, prows = by_tiles ? a->frame.tile_h : rows, pnt = by_tiles ? tn : nt;
            const size_t n = (size_t)prows * a->frame.tile_w * (size_t)pnt;
            k_pack_planes<<<(unsigned)((n + 255) / 256), 256, 0, c->s_d2h>>>(want_sun ? o_shadow + bo : nullptr, o_pt + bo, want_pt ? a->n_point : 0, o_sp + bo, want_sp ? a->n_spot : 0, px,
                                                                         o_spec + bo, want_mask ? o_mask + bo * (size_t)mask_bytes : nullptr, mask_bytes, want_rf ? o_code + bo : nullptr,
                                                                         a->frame.tile_w, a->frame.tile_h, prow0, prows, pnt, a->frame.width, a->frame.height, fd.tile_first, fd.tile_stride);
            VXL_LAUNCH_CHECK(c);
            if (want_mask) VXL_CUDA(copy_band(packed->shadow_mask, o_mask, (size_t)mask_bytes, cudaMemcpyDeviceToHost, c->s_d2h));
            if (want_rf) VXL_CUDA(copy_band(packed->spec_code, o_code, 1, cudaMemcpyDeviceToHost, c->s_d2h));
        }
    }
    c->band_row0 = 0; c->band_rows = 0;
    VXL_CUDA(cudaEventRecord(ev[1], c->s_d2h));
    VXL_CUDA(cudaStreamWaitEvent(main, ev[1], 0));                        // the context's stream stays the single point of order
    VXL_CUDA(cudaStreamSynchronize(c->s_h2d));
    VXL_CUDA(cudaStreamSynchronize(sl));
    VXL_CUDA(cudaStreamSynchronize(sr));
    VXL_CUDA(cudaStreamSynchronize(main));
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "vxl_lighting_host copies");
    return VXL_OK;
}

int vxl_lighting_host(vxl_ctx* c, vxl_volume* vol, const vxl_lighting_host_args* a) { return lighting_host_impl(c, vol, a, nullptr); }

int vxl_lighting_host_packed(vxl_ctx* c, vxl_volume* vol, const vxl_lighting_host_args* a, const vxl_packed_planes* out) {
    if (!out) { set_error("vxl_lighting_host_packed: out is NULL"); return VXL_ERR_INVALID; }
    return lighting_host_impl(c, vol, a, out);
}

}  // extern "C"
