// vxl_api.cu -- context, error reporting, memory helpers, counters and the whole-frame host
// drop-in of libvxl.so.  No CPU fallback anywhere: every entry point needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "vxl_internal.h"

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

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_abi_version(void) { return VXL_ABI_VERSION; }
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
    VXL_CUDA(cudaMalloc(&c->d_luts, sizeof lut));
    VXL_CUDA(cudaMemcpy(c->d_luts, lut, sizeof lut, cudaMemcpyHostToDevice));
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
    cudaFree(c->d_stats); cudaFree(c->d_luts); cudaFree(c->d_lights); cudaFree(c->d_perm);
    for (auto& m : c->models) cudaFree((void*)m.voxels);
    cudaFree(c->d_models); cudaFree(c->d_hkeys); cudaFree(c->d_hvals); cudaFree(c->d_ents); cudaFree(c->d_aabb);
    cudaFree(c->h_planes); cudaFree(c->h_out); cudaFree(c->h_noise);
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

int vxl_lighting_host(vxl_ctx* c, vxl_volume* vol, const vxl_lighting_host_args* a) {
    if (!c || !vol || !a || !a->view) { set_error("vxl_lighting_host: bad argument"); return VXL_ERR_INVALID; }
    if (a->n_point < 0 || a->n_spot < 0 || a->n_point > VXL_MAX_LIGHTS || a->n_spot > VXL_MAX_LIGHTS) { set_error("vxl_lighting_host: light count out of range"); return VXL_ERR_LIMIT; }
    FrameView Fh;
    if (int e = frame_view(&a->frame, &Fh)) return e;
    const size_t px = frame_pixels(&a->frame);
    if (px == 0) return VXL_OK;
    const bool want_amb = a->out_shadow || a->out_ao;
    const bool want_pt = a->n_point > 0 && a->out_point_shadow;
    const bool want_sp = a->n_spot > 0 && a->out_spot_shadow;
    const bool want_rf = a->out_spec_t != nullptr;
    if (want_rf && !a->frame.material) { set_error("vxl_lighting_host: spec pass needs frame.material"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(c->device));
    if (int e = ensure((void**)&c->h_planes, &c->h_planes_bytes, px * 4 * 3)) return e;
    const size_t n_out = 3 + (size_t)a->n_point + (size_t)a->n_spot;
    if (int e = ensure((void**)&c->h_out, &c->h_out_bytes, px * 4 * n_out)) return e;
    if (!c->h_noise) VXL_CUDA(cudaMalloc(&c->h_noise, 512 * 512 * 4));
    uint32_t* d_depth = c->h_planes; uint32_t* d_normal = d_depth + px; uint32_t* d_mat = d_normal + px;
    VXL_CUDA(cudaMemcpyAsync(d_depth, a->frame.depth24, px * 4, cudaMemcpyHostToDevice, c->stream));
    VXL_CUDA(cudaMemcpyAsync(d_normal, a->frame.normal, px * 4, cudaMemcpyHostToDevice, c->stream));
    if (want_rf) VXL_CUDA(cudaMemcpyAsync(d_mat, a->frame.material, px * 4, cudaMemcpyHostToDevice, c->stream));
    VXL_CUDA(cudaMemcpyAsync(c->h_noise, a->frame.noise, 512 * 512 * 4, cudaMemcpyHostToDevice, c->stream));
    vxl_frame fd = a->frame;
    fd.depth24 = d_depth; fd.normal = d_normal; fd.material = d_mat; fd.noise = c->h_noise;
    float* o_shadow = c->h_out; float* o_ao = o_shadow + px; float* o_spec = o_ao + px;
    float* o_pt = o_spec + px; float* o_sp = o_pt + px * (size_t)a->n_point;
    if (want_amb) {
        if (int e = vxl_pass_ambient(c, vol, a->view, &fd, a->n_ao, a->out_shadow ? o_shadow : nullptr, a->out_ao ? o_ao : nullptr)) return e;
        if (a->out_shadow) VXL_CUDA(cudaMemcpyAsync(a->out_shadow, o_shadow, px * 4, cudaMemcpyDeviceToHost, c->stream));
        if (a->out_ao) VXL_CUDA(cudaMemcpyAsync(a->out_ao, o_ao, px * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (want_pt) {
        if (int e = vxl_pass_point(c, vol, a->view, &fd, a->point, a->n_point, o_pt)) return e;
        VXL_CUDA(cudaMemcpyAsync(a->out_point_shadow, o_pt, px * 4 * (size_t)a->n_point, cudaMemcpyDeviceToHost, c->stream));
    }
    if (want_sp) {
        if (int e = vxl_pass_spot(c, vol, a->view, &fd, a->spot, a->n_spot, o_sp)) return e;
        VXL_CUDA(cudaMemcpyAsync(a->out_spot_shadow, o_sp, px * 4 * (size_t)a->n_spot, cudaMemcpyDeviceToHost, c->stream));
    }
    if (want_rf) {
        if (int e = vxl_pass_reflection(c, vol, a->view, &fd, o_spec)) return e;
        VXL_CUDA(cudaMemcpyAsync(a->out_spec_t, o_spec, px * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    return VXL_OK;
}

}  // extern "C"
