// vxl_internal.h -- private definitions shared by the .cu files of libvxl.so.
#pragma once
#include <cuda.h>             // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/vxl.h"

namespace vxl {

// One occupancy bitmask level (vxl_occupancy.cu): bit (x & 31) of word x >> 5, cell = 2^shift voxels.
struct BitLevel {
    uint32_t* d_words = nullptr;
    int cx = 0, cy = 0, cz = 0;          // array extent in cells, borders included
    int xw = 0;                          // words per (y, z) row: ceil(cx/32) + 1 spare zero word
    int cyp = 0;                         // cy rounded up to a multiple of 4: the words are ordered [az][word][ay], y innermost
    int shift = 0;                       // log2(voxels per cell edge)
    int border = 0;                      // array index = cell index + border (zero cells around the level proper)
    int copies = 1;                      // 2: a second copy of the array follows, shifted by 16 cells along x (k_occ_shift16)
    // TMA descriptors of this array as a 3-D tensor of 32-bit words (cyp, xw, cz), one per box shape a kernel stages
    // (vxl_passes.cu); built on first use by level_tensor_map (vxl_occupancy.cu); ty = key of the box shape
    struct BoxMap { int ty = 0; CUtensorMap map; };
    BoxMap box[2];
};
struct BitView { const uint32_t* __restrict__ words; int cx, cy, cz, xw, cyp, border, copies; };

// Device-side view of the occupancy volume handed to kernels by value.
struct VolView {
    const uint8_t* __restrict__ bytes;   // canonical packed bytes, x fastest (reference layout)
    int sx, sy, sz;                      // texels
    // derived, acceleration only (never changes a result); see vxl_occupancy.cu
    BitView tex;                         // texel level: 1 bit per packed byte (2-voxel cells); bit = (byte != 0)
    BitView occ[3];                      // plain levels: 4-, 8-, 16-voxel cells
    BitView dil[2];                      // dilated levels: 8-, 16-voxel cells
};

struct FrameView {
    int width, height, tile_w, tile_h, tile_first, tile_stride, n_tiles, tiles_x;
    int row0, rows;                      // the light passes cover rows [row0, row0 + rows) of every tile (a band; default: all)
    const uint32_t* __restrict__ depth24;
    const uint32_t* __restrict__ normal;
    const uint32_t* __restrict__ material;
    const uint32_t* __restrict__ noise;
    // multi-GPU: every output store of the light-pass kernels is repeated at (address + mirror[i]) -- the same offset inside each
    // peer's copy of the gathered tile stack, mapped through CUDA IPC (vxl_ctx_set_output_mirrors); 0 on one GPU
    int n_mirror;
    long long mirror[15];
};

struct ModelDev { const uint8_t* voxels; int sx, sy, sz; unsigned solid; const uint8_t* mip1; const uint8_t* mip2; };   // mips: VoxAsset::Upload's chain, built on first use

constexpr int STAT_SLOTS = 64;   // striped counters: slot = blockIdx & 63, 4 x u64 each

}  // namespace vxl

struct vxl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    int variant = 1;                         // 0: plain march on the volume bytes; 1: occupancy-bit tile in shared memory (default)
    unsigned long long* d_stats = nullptr;   // [STAT_SLOTS][4]
    float* d_luts = nullptr;                 // cos[256] sin[256]
    float* d_taa_lut = nullptr;              // (cos, sin)[256][12] of LightTAA's spiral angles (vxl_post.cu), built on first use
    float* d_taa_depth = nullptr;            // the frame's depth plane decoded to float (vxl_light_taa), grown on demand
    size_t taa_depth_cap = 0;
    void* d_lights = nullptr;                // VXL_MAX_LIGHTS * 64 B
    uint8_t* d_perm = nullptr;               // perm[512] perm12[512] (terrain generator)
    std::vector<vxl::ModelDev> models;
    vxl::ModelDev* d_models = nullptr;
    int d_models_cap = 0;
    void* d_draws = nullptr;                 // draw list of vxl_gbuffer_models (vxl_model.cu), grown on demand
    int draws_cap = 0;
    // voxeliser scratch
    unsigned long long* d_hkeys = nullptr;
    unsigned* d_hvals = nullptr;
    size_t hcap = 0;
    vxl_entity* d_ents = nullptr;
    int* d_aabb = nullptr;                   // [n][6]
    int ents_cap = 0;
    unsigned long long aabb_gen = 0;         // bumped by every voxelise call (they share the scratch block)
    // host drop-in scratch
    uint32_t* h_planes = nullptr;            // device buffers for vxl_lighting_host
    size_t h_planes_bytes = 0;
    float* h_out = nullptr;
    size_t h_out_bytes = 0;
    uint32_t* h_noise = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of vxl_lighting_host (uploads / read-backs overlap the passes)
    cudaStream_t s_side[2] = {nullptr, nullptr};     // side streams of vxl_lighting / vxl_lighting_host: the local-light and reflection passes beside the ambient pass
    std::vector<cudaEvent_t> ev;
    int band_row0 = 0, band_rows = 0;                // > 0 rows: the pass entry points cover this row band of every tile only
    int n_mirror = 0;                                // vxl_ctx_set_output_mirrors
    long long mirror[15] = {0};
    size_t light_plane_stride = 0;                   // vxl_ctx_set_light_plane_stride (pixels; 0 = the shard's own size)
    std::vector<void*> ipc_open;                     // peer mappings to close with the context
};

struct vxl_volume {
    vxl_ctx* ctx = nullptr;
    int sx = 0, sy = 0, sz = 0;
    uint8_t* d_bytes = nullptr;
    bool dirty = true;
    // dirty, the levels exist, and everything that changed since they were built lies inside the boxes of the last voxelise call
    // (device, [n][6] voxel coordinates in the context's scratch block, valid while the context's aabb_gen still equals dirty_gen)
    bool dirty_partial = false;
    const int* dirty_boxes = nullptr;
    int n_dirty_boxes = 0;
    unsigned long long dirty_gen = 0;
    vxl::BitLevel tex;                   // occupancy bitmask at texel (2-voxel) cells
    vxl::BitLevel occ[3];                // occupancy bitmasks at 4-, 8-, 16-voxel cells
    vxl::BitLevel dil[2];                // 3x3x3-dilated bitmasks at 8-, 16-voxel cells
};

namespace vxl {
void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what);
#define VXL_CUDA(call)                                             \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return vxl::cuda_fail(_e, #call);   \
    } while (0)
#define VXL_LAUNCH_CHECK(ctx)                                                  \
    do {                                                                       \
        (ctx)->launches++;                                                     \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) return vxl::cuda_fail(_e, "kernel launch");     \
    } while (0)

inline VolView vol_view(const vxl_volume* v) {
    VolView r;
    r.bytes = v->d_bytes; r.sx = v->sx; r.sy = v->sy; r.sz = v->sz;
    auto view = [](const BitLevel& L) { return BitView{L.d_words, L.cx, L.cy, L.cz, L.xw, L.cyp, L.border, L.copies}; };
    r.tex = view(v->tex);
    for (int i = 0; i < 3; ++i) r.occ[i] = view(v->occ[i]);
    for (int i = 0; i < 2; ++i) r.dil[i] = view(v->dil[i]);
    return r;
}
// TMA descriptor for boxes of ty cells (y, a multiple of 4) x tw words (x) x tz slices of an occupancy level (cached in the level)
int level_tensor_map(BitLevel& L, int tw, int ty, int tz, const CUtensorMap** out);
int frame_view(const vxl_frame* f, FrameView* out);
size_t frame_pixels(const vxl_frame* f);   // n_tiles * tile_w * tile_h
}  // namespace vxl
