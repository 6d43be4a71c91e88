// vxl_occupancy.cu -- derived occupancy levels of the world volume: packed cell bitmasks.
//
// Pure acceleration (SURVEY.md 7 step 5 "k_occupancy_mips"): nothing here exists in the reference
// (its volume has one mip level, ShadowVoxSystem.cpp:61) and nothing here may change a result.
//
// Texel level (L = 1, cell = one packed byte = 2 voxels per axis): bit = (byte != 0).  This is the reference's own
//   coarse test (Light.frag:163 getVolumeAt(pos, 1)), so inside a staged window of this level a phase-2 probe is
//   answered exactly, and a phase-1 probe whose bit is clear cannot hit.
// Plain level L (cell = 2^L voxels per axis; L = 2, 3, 4):
//   occ_L[c] = 1 iff any packed byte of the canonical volume inside cell c is non-zero.
//   A march probe (Light.frag:140 fine bit test, :163 coarse byte test) whose cell bit is 0 reads a zero
//   texel and cannot hit, so the light passes test the bit (staged per thread block in shared memory,
//   vxl_bitmarch.cuh) and touch the volume only when it is set.
// Dilated level L (L = 3, 4):
//   dil_L[c] = OR of occ_L over the 3x3x3 neighbourhood of c.  dil_L[c] = 0 means every point within 2^L voxels
//   (per axis) of any point of c lies in an empty cell, so one test clears a whole run of small-step probes.
//   Stored with a 1-cell border so that cells just outside the volume still see their occupied neighbours.
// Cells outside a level's array are empty (texelFetch out of range reads 0).
// Storage: 32 cells per word along x, pitch = words holding cells + 1 spare zero word.
#include "vxl_internal.h"

namespace vxl {

// texel level: one thread per output word = 32 consecutive bytes of a volume row; HBM-bound (reads the volume once)
__device__ __forceinline__ uint32_t nonzero_bytes4(uint32_t v) {        // bit i = (byte i of v != 0)
    const uint32_t m = (v | ((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
    return ((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u);
}
__global__ void __launch_bounds__(256) k_occ_texels(const uint8_t* __restrict__ bytes, int sx, int sy, int sz,
                                                    int pitch, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)pitch * sy * sz;
    if (i >= total) return;
    const int w = (int)(i % pitch), y = (int)((i / pitch) % sy), z = (int)(i / ((long long)pitch * sy));
    uint32_t word = 0;
    const int x0 = w * 32;
    if (x0 < sx) {
        const uint8_t* row = bytes + (size_t)y * sx + (size_t)z * ((size_t)sx * sy) + x0;
        if (x0 + 32 <= sx && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(row)), b = __ldg(reinterpret_cast<const uint4*>(row) + 1);
            word = nonzero_bytes4(a.x) | (nonzero_bytes4(a.y) << 4) | (nonzero_bytes4(a.z) << 8) | (nonzero_bytes4(a.w) << 12) |
                   (nonzero_bytes4(b.x) << 16) | (nonzero_bytes4(b.y) << 20) | (nonzero_bytes4(b.z) << 24) | (nonzero_bytes4(b.w) << 28);
        } else {
            const int n = min(32, sx - x0);
            for (int k = 0; k < n; ++k)
                if (row[k]) word |= 1u << k;
        }
    }
    out[i] = word;
}

// coarser plain level from a finer one: bit c = OR of the 2x2x2 child cells
__global__ void __launch_bounds__(256) k_occ_coarsen(const uint32_t* __restrict__ fine, int fcy, int fcz, int fpitch,
                                                     int cy, int cz, int pitch, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)pitch * cy * cz;
    if (i >= total) return;
    const int w = (int)(i % pitch), y = (int)((i / pitch) % cy), z = (int)(i / ((long long)pitch * cy));
    uint32_t word = 0;
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy) {
            const int fy = 2 * y + dy, fz = 2 * z + dz;
            if (fy >= fcy || fz >= fcz) continue;
            const uint32_t* row = fine + ((size_t)fz * fcy + fy) * fpitch;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int fw = 2 * w + h;
                if (fw >= fpitch) continue;
                uint32_t v = row[fw];
                v = (v | (v >> 1)) & 0x55555555u;                  // OR of bit pairs at even positions
                v = (v | (v >> 1)) & 0x33333333u; v = (v | (v >> 2)) & 0x0F0F0F0Fu;
                v = (v | (v >> 4)) & 0x00FF00FFu; v = (v | (v >> 8)) & 0x0000FFFFu;   // compact to 16 bits
                word |= v << (16 * h);
            }
        }
    out[i] = word;
}

// dilated level (array index = cell + 1): OR over the 3x3x3 neighbourhood of the plain level
__global__ void __launch_bounds__(256) k_occ_dilate(const uint32_t* __restrict__ in, int icy, int icz, int ipitch,
                                                    int cy, int cz, int pitch, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)pitch * cy * cz;
    if (i >= total) return;
    const int w = (int)(i % pitch), y = (int)((i / pitch) % cy) - 1, z = (int)(i / ((long long)pitch * cy)) - 1;   // cell coords
    // output bit b of word w is cell x = 32*w + b - 1; gather input bits x-1, x, x+1 = input positions 32*w + b - 2 .. 32*w + b
    uint32_t word = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = y + dy, iz = z + dz;
            if (iy < 0 || iz < 0 || iy >= icy || iz >= icz) continue;
            const uint32_t* row = in + ((size_t)iz * icy + iy) * ipitch;
            const uint32_t cur = w < ipitch ? row[w] : 0u;
            const uint32_t prev = (w >= 1 && w - 1 < ipitch) ? row[w - 1] : 0u;
            // input position p = 32*w + b - s for s = 0, 1, 2  ->  (cur << s) | (prev >> (32 - s))
            word |= cur | (cur << 1) | (prev >> 31) | (cur << 2) | (prev >> 30);
        }
    out[i] = word;
}

static int alloc_level(BitLevel& L, int shift, int cx, int cy, int cz, int border) {
    L.shift = shift; L.border = border;
    L.cx = cx + 2 * border; L.cy = cy + 2 * border; L.cz = cz + 2 * border;
    L.pitch = (L.cx + 31) / 32 + 1;                           // one spare (all-zero) word per row for the staging funnel shift
    VXL_CUDA(cudaMalloc(&L.d_words, (size_t)L.pitch * L.cy * L.cz * 4));
    return VXL_OK;
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_volume_build_occupancy(vxl_volume* v) {
    if (!v) { set_error("vxl_volume_build_occupancy: vol is NULL"); return VXL_ERR_INVALID; }
    vxl_ctx* c = v->ctx;
    VXL_CUDA(cudaSetDevice(c->device));
    if (!v->occ[0].d_words) {
        if (int e = alloc_level(v->tex, 1, v->sx, v->sy, v->sz, 0)) return e;
        for (int li = 0; li < 3; ++li) {
            const int tpc = 2 << li;                          // texels per cell edge: 2, 4, 8
            if (int e = alloc_level(v->occ[li], 2 + li, (v->sx + tpc - 1) / tpc, (v->sy + tpc - 1) / tpc, (v->sz + tpc - 1) / tpc, 0)) return e;
        }
        for (int li = 0; li < 2; ++li) {
            const BitLevel& P = v->occ[1 + li];
            if (int e = alloc_level(v->dil[li], P.shift, P.cx, P.cy, P.cz, 1)) return e;
        }
    }
    auto grid_of = [](const BitLevel& L) { return (unsigned)(((long long)L.pitch * L.cy * L.cz + 255) / 256); };
    BitLevel& L1 = v->tex;
    k_occ_texels<<<grid_of(L1), 256, 0, c->stream>>>(v->d_bytes, v->sx, v->sy, v->sz, L1.pitch, L1.d_words);
    VXL_LAUNCH_CHECK(c);
    for (int li = 0; li < 3; ++li) {
        const BitLevel& F = li ? v->occ[li - 1] : v->tex;
        BitLevel& L = v->occ[li];
        k_occ_coarsen<<<grid_of(L), 256, 0, c->stream>>>(F.d_words, F.cy, F.cz, F.pitch, L.cy, L.cz, L.pitch, L.d_words);
        VXL_LAUNCH_CHECK(c);
    }
    for (int li = 0; li < 2; ++li) {
        const BitLevel& P = v->occ[1 + li];
        BitLevel& D = v->dil[li];
        k_occ_dilate<<<grid_of(D), 256, 0, c->stream>>>(P.d_words, P.cy, P.cz, P.pitch, D.cy, D.cz, D.pitch, D.d_words);
        VXL_LAUNCH_CHECK(c);
    }
    v->dirty = false;
    return VXL_OK;
}

/* diagnostics: download one occupancy level unpacked to 0/1 bytes [cz][cy][cx]; level = 1, 2, 3, 4 (plain, cell =
 * 2^level voxels) or 13, 14 (dilated levels 3, 4, including their 1-cell border); out_dims = {cx, cy, cz} */
int vxl_volume_debug_occupancy(vxl_volume* v, int level, uint8_t* host_out, int* out_dims) {
    if (!v || !((level >= 1 && level <= 4) || level == 13 || level == 14)) { set_error("vxl_volume_debug_occupancy: bad argument"); return VXL_ERR_INVALID; }
    if (v->dirty || !v->occ[0].d_words) { if (int e = vxl_volume_build_occupancy(v)) return e; }
    const BitLevel& L = level == 1 ? v->tex : (level < 10 ? v->occ[level - 2] : v->dil[level - 13]);
    if (out_dims) { out_dims[0] = L.cx; out_dims[1] = L.cy; out_dims[2] = L.cz; }
    if (!host_out) return VXL_OK;
    const size_t words = (size_t)L.pitch * L.cy * L.cz;
    std::vector<uint32_t> h(words);
    VXL_CUDA(cudaMemcpyAsync(h.data(), L.d_words, words * 4, cudaMemcpyDeviceToHost, v->ctx->stream));
    VXL_CUDA(cudaStreamSynchronize(v->ctx->stream));
    for (int z = 0; z < L.cz; ++z)
        for (int y = 0; y < L.cy; ++y)
            for (int x = 0; x < L.cx; ++x)
                host_out[((size_t)z * L.cy + y) * L.cx + x] = (uint8_t)((h[((size_t)z * L.cy + y) * L.pitch + (x >> 5)] >> (x & 31)) & 1u);
    return VXL_OK;
}

}  // extern "C"
