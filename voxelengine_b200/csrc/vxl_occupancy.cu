// vxl_occupancy.cu -- derived occupancy levels of the world volume: packed cell bitmasks.
//
// Pure acceleration (SURVEY.md 7 step 5 "k_occupancy_mips"): nothing here exists in the reference
// (its volume has one mip level, ShadowVoxSystem.cpp:61) and nothing here may change a result.
//
// For a cell of 2^L voxels per axis (L = 2: 2x2x2 texels; L = 3: 4x4x4 texels)
//   occ_L[c] = 1 iff any packed byte of the canonical volume inside cell c is non-zero,
// stored 32 cells per word along x.  A march probe (Light.frag:140 fine bit test, :163 coarse byte
// test) whose cell bit is 0 reads a zero texel and cannot hit, so the light passes test the bit
// (staged per thread block in shared memory, vxl_bitmarch.cuh) and touch the volume only when it is
// set.  Cells outside the volume are empty (texelFetch out of range reads 0).
#include "vxl_internal.h"

namespace vxl {

// one thread per output word: 32 cells along x, each TPC^3 texels
template <int TPC>
__global__ void __launch_bounds__(256) k_occ_bits(const uint8_t* __restrict__ bytes, int sx, int sy, int sz,
                                                  int cy, int cz, int pitch, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)pitch * cy * cz;
    if (i >= total) return;
    const int w = (int)(i % pitch), y = (int)((i / pitch) % cy), z = (int)(i / ((long long)pitch * cy));
    uint32_t word = 0;
    const int x0 = w * 32 * TPC;
    if (x0 < sx) {
        for (int dz = 0; dz < TPC; ++dz) {
            const int tz = z * TPC + dz;
            if (tz >= sz) break;
            for (int dy = 0; dy < TPC; ++dy) {
                const int ty = y * TPC + dy;
                if (ty >= sy) break;
                const uint8_t* row = bytes + (size_t)ty * sx + (size_t)tz * ((size_t)sx * sy);
                const int n = min(32 * TPC, sx - x0);
#pragma unroll 4
                for (int k = 0; k < n; ++k)
                    if (row[x0 + k]) word |= 1u << (k / TPC);
            }
        }
    }
    out[i] = word;
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_volume_build_occupancy(vxl_volume* v) {
    if (!v) { set_error("vxl_volume_build_occupancy: vol is NULL"); return VXL_ERR_INVALID; }
    vxl_ctx* c = v->ctx;
    VXL_CUDA(cudaSetDevice(c->device));
    for (int li = 0; li < 2; ++li) {
        BitLevel& L = v->occ[li];
        const int tpc = 2 << li;                              // texels per cell edge: 2, 4
        if (!L.d_words) {
            L.shift = 2 + li;
            L.cx = (v->sx + tpc - 1) / tpc; L.cy = (v->sy + tpc - 1) / tpc; L.cz = (v->sz + tpc - 1) / tpc;
            L.pitch = (L.cx + 31) / 32 + 1;                   // one spare (all-zero) word per row for the staging funnel shift
            VXL_CUDA(cudaMalloc(&L.d_words, (size_t)L.pitch * L.cy * L.cz * 4));
        }
        const long long total = (long long)L.pitch * L.cy * L.cz;
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (li == 0) k_occ_bits<2><<<grid, 256, 0, c->stream>>>(v->d_bytes, v->sx, v->sy, v->sz, L.cy, L.cz, L.pitch, L.d_words);
        else k_occ_bits<4><<<grid, 256, 0, c->stream>>>(v->d_bytes, v->sx, v->sy, v->sz, L.cy, L.cz, L.pitch, L.d_words);
        VXL_LAUNCH_CHECK(c);
    }
    v->dirty = false;
    return VXL_OK;
}

/* diagnostics: download one occupancy level (shift 2: 4-voxel cells, 3: 8-voxel cells) unpacked to bytes
 * [cz][cy][cx] of 0/1; out_dims = {cx, cy, cz} */
int vxl_volume_debug_occupancy(vxl_volume* v, int shift, uint8_t* host_out, int* out_dims) {
    if (!v || (shift != 2 && shift != 3)) { set_error("vxl_volume_debug_occupancy: bad argument"); return VXL_ERR_INVALID; }
    if (v->dirty || !v->occ[0].d_words) { if (int e = vxl_volume_build_occupancy(v)) return e; }
    const BitLevel& L = v->occ[shift - 2];
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
