// vxl_api.cu -- context, error reporting, memory helpers, counters and the whole-frame host
// drop-in of libvxl.so.  No CPU fallback anywhere: every entry point needs a CUDA device.
#include <algorithm>
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
    for (void* p : c->ipc_open) cudaIpcCloseMemHandle(p);
    cudaFree(c->d_stats); cudaFree(c->d_luts); cudaFree(c->d_taa_lut); cudaFree(c->d_lights); cudaFree(c->d_perm);
    for (auto& m : c->models) { cudaFree((void*)m.voxels); cudaFree((void*)m.mip1); cudaFree((void*)m.mip2); }
    cudaFree(c->d_models); cudaFree(c->d_draws); cudaFree(c->d_hkeys); cudaFree(c->d_hvals); cudaFree(c->d_ents); cudaFree(c->d_aabb);
    cudaFree(c->h_planes); cudaFree(c->h_out); cudaFree(c->h_noise);
    if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_d2h); }
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
    if (!c->s_h2d) {
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    }
    while (c->ev.size() < 3) { cudaEvent_t e; VXL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev.push_back(e); }
    cudaEvent_t* ev = c->ev.data();
    cudaStream_t main = c->stream, side[2] = {c->s_h2d, c->s_d2h};
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
    // the host drop-in writes to its own staging planes: no mirrored stores, default plane stride (restored on every exit path)
    struct Restore { vxl_ctx* c; int n; size_t st; ~Restore() { c->n_mirror = n; c->light_plane_stride = st; } } restore{c, c->n_mirror, c->light_plane_stride};
    c->n_mirror = 0; c->light_plane_stride = 0;
    VXL_CUDA(cudaSetDevice(c->device));
    if (int e = ensure((void**)&c->h_planes, &c->h_planes_bytes, px * 4 * 3)) return e;
    const size_t n_out = 3 + (size_t)a->n_point + (size_t)a->n_spot;
    if (int e = ensure((void**)&c->h_out, &c->h_out_bytes, px * 4 * n_out)) return e;
    if (!c->h_noise) VXL_CUDA(cudaMalloc(&c->h_noise, 512 * 512 * 4));
    uint32_t* d_depth = c->h_planes; uint32_t* d_normal = d_depth + px; uint32_t* d_mat = d_normal + px;
    // Three streams: uploads, passes (the context's stream), read-backs; the frame goes through in NB row bands (rows of
    // every tile of the shard).  PCIe is full duplex and the copy engines run beside the SMs, so band b+1 uploads and band
    // b-1 reads back while band b is in the passes; within a band the local-light planes (the largest output) go first.
    int NB = a->frame.tile_h >= 256 ? 4 : 1;
    if (const char* nb = getenv("VXL_HOST_BANDS")) NB = std::max(1, std::min(64, atoi(nb)));   // tuning knob
    const int band_h = ((a->frame.tile_h + NB - 1) / NB + 15) / 16 * 16;          // whole 16-row thread blocks
    const size_t n_ev = 2 + (size_t)NB * 5;
    if (!c->s_h2d) {
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        VXL_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    }
    while (c->ev.size() < n_ev) { cudaEvent_t e; VXL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev.push_back(e); }
    cudaEvent_t* ev = c->ev.data();
    const size_t tile_px = (size_t)a->frame.tile_w * a->frame.tile_h;
    const int nt = a->frame.n_tiles;
    // rows [r0, r0 + rows) of every tile of a tile-compact plane: nt chunks of rows * tile_w pixels, tile_px apart
    auto copy_band = [&](void* dst, const void* src, int r0, int rows, cudaMemcpyKind kind, cudaStream_t st) -> cudaError_t {
        const size_t off = (size_t)r0 * a->frame.tile_w * 4;
        if (nt == 1) return cudaMemcpyAsync((char*)dst + off, (const char*)src + off, (size_t)rows * a->frame.tile_w * 4, kind, st);
        return cudaMemcpy2DAsync((char*)dst + off, tile_px * 4, (const char*)src + off, tile_px * 4, (size_t)rows * a->frame.tile_w * 4, (size_t)nt, kind, st);
    };
    VXL_CUDA(cudaEventRecord(ev[0], c->stream));                          // order behind whatever the caller queued (voxelise, ...)
    VXL_CUDA(cudaStreamWaitEvent(c->s_h2d, ev[0], 0));
    VXL_CUDA(cudaStreamWaitEvent(c->s_d2h, ev[0], 0));
    VXL_CUDA(cudaMemcpyAsync(c->h_noise, a->frame.noise, 512 * 512 * 4, cudaMemcpyHostToDevice, c->s_h2d));
    vxl_frame fd = a->frame;
    fd.depth24 = d_depth; fd.normal = d_normal; fd.material = d_mat; fd.noise = c->h_noise;
    float* o_shadow = c->h_out; float* o_ao = o_shadow + px; float* o_spec = o_ao + px;
    float* o_pt = o_spec + px; float* o_sp = o_pt + px * (size_t)a->n_point;
    int rc = VXL_OK;
    for (int b = 0; b < NB && rc == VXL_OK; ++b) {
        const int r0 = b * band_h, rows = std::min(band_h, a->frame.tile_h - r0);
        if (rows <= 0) break;
        cudaEvent_t* eb = ev + 2 + (size_t)b * 5;
        VXL_CUDA(copy_band(d_depth, a->frame.depth24, r0, rows, cudaMemcpyHostToDevice, c->s_h2d));
        VXL_CUDA(copy_band(d_normal, a->frame.normal, r0, rows, cudaMemcpyHostToDevice, c->s_h2d));
        if (want_rf) VXL_CUDA(copy_band(d_mat, a->frame.material, r0, rows, cudaMemcpyHostToDevice, c->s_h2d));
        VXL_CUDA(cudaEventRecord(eb[0], c->s_h2d));
        VXL_CUDA(cudaStreamWaitEvent(c->stream, eb[0], 0));
        c->band_row0 = r0; c->band_rows = rows;
        if (want_pt && rc == VXL_OK) {
            rc = vxl_pass_point(c, vol, a->view, &fd, a->point, a->n_point, o_pt);
            if (rc == VXL_OK) {
                cudaEventRecord(eb[1], c->stream); cudaStreamWaitEvent(c->s_d2h, eb[1], 0);
                for (int l = 0; l < a->n_point; ++l) copy_band(a->out_point_shadow + px * (size_t)l, o_pt + px * (size_t)l, r0, rows, cudaMemcpyDeviceToHost, c->s_d2h);
            }
        }
        if (want_sp && rc == VXL_OK) {
            rc = vxl_pass_spot(c, vol, a->view, &fd, a->spot, a->n_spot, o_sp);
            if (rc == VXL_OK) {
                cudaEventRecord(eb[2], c->stream); cudaStreamWaitEvent(c->s_d2h, eb[2], 0);
                for (int l = 0; l < a->n_spot; ++l) copy_band(a->out_spot_shadow + px * (size_t)l, o_sp + px * (size_t)l, r0, rows, cudaMemcpyDeviceToHost, c->s_d2h);
            }
        }
        if (want_amb && rc == VXL_OK) {
            rc = vxl_pass_ambient(c, vol, a->view, &fd, a->n_ao, a->out_shadow ? o_shadow : nullptr, a->out_ao ? o_ao : nullptr);
            if (rc == VXL_OK) {
                cudaEventRecord(eb[3], c->stream); cudaStreamWaitEvent(c->s_d2h, eb[3], 0);
                if (a->out_shadow) copy_band(a->out_shadow, o_shadow, r0, rows, cudaMemcpyDeviceToHost, c->s_d2h);
                if (a->out_ao) copy_band(a->out_ao, o_ao, r0, rows, cudaMemcpyDeviceToHost, c->s_d2h);
            }
        }
        if (want_rf && rc == VXL_OK) {
            rc = vxl_pass_reflection(c, vol, a->view, &fd, o_spec);
            if (rc == VXL_OK) {
                cudaEventRecord(eb[4], c->stream); cudaStreamWaitEvent(c->s_d2h, eb[4], 0);
                copy_band(a->out_spec_t, o_spec, r0, rows, cudaMemcpyDeviceToHost, c->s_d2h);
            }
        }
    }
    c->band_row0 = 0; c->band_rows = 0;
    VXL_CUDA(cudaEventRecord(ev[1], c->s_d2h));
    VXL_CUDA(cudaStreamWaitEvent(c->stream, ev[1], 0));                   // the context's stream stays the single point of order
    VXL_CUDA(cudaStreamSynchronize(c->s_h2d));
    VXL_CUDA(cudaStreamSynchronize(c->stream));
    if (rc != VXL_OK) return rc;
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "vxl_lighting_host copies");
    return VXL_OK;
}

}  // extern "C"
