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
// Storage (one array per level): bit (ax & 31) of word (ax >> 5) of row (ay, az), array index a = cell index + border, the words of
//   the array ordered [az][word][ay] -- y innermost.  A window of the array is then a 3-D box whose innermost extent runs along y in
//   single cells, which is what lets the Tensor Memory Accelerator fetch it (cp.async.bulk.tensor wants the box to start on a
//   16-byte boundary of the innermost dimension: 4 cells along y here, against 128 cells if x words were innermost).  Rows are
//   padded to cyp = a multiple of 4 words; each (ay, az) row has one spare all-zero word at the end (staging funnel shift).
//   The plain levels carry a zero border of 192 voxels per side (96 / 48 / 24 / 12 cells: each level's border is twice the next
//   coarser one's, so 2x2x2 children stay aligned), so that a ray box that pokes out of the volume can still be looked up
//   without bounds checks; the dilated levels add one more cell.
#include "vxl_internal.h"

namespace vxl {

__device__ __forceinline__ size_t level_index(int xw, int cyp, int w, int ay, int az) { return ((size_t)az * xw + w) * cyp + ay; }

// texel level: one thread per output word = 32 consecutive bytes of a volume row; HBM-bound (reads the volume once)
__device__ __forceinline__ uint32_t nonzero_bytes4(uint32_t v) {        // bit i = (byte i of v != 0)
    const uint32_t m = (v | ((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
    return ((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u);
}
__device__ __forceinline__ uint32_t texel_word(const uint8_t* __restrict__ bytes, int sx, int sy, int sz, int border, int w, int ay, int az) {
    // Texel -1 of every axis repeats texel 0: the reference truncates toward zero (ivec3(pos / 2), ivec3(pos) / 2), so a probe at a
    // coordinate in (-2, 0) reads texel 0.  With floor(pos / 2) = -1 there, the repeated texel gives the same answer, and the
    // coarser levels built from this one inherit it (cell -1 = texel -1 | texel -2 = texel 0).  Everything further out is empty.
    int y = ay - border, z = az - border;
    const int x0 = w * 32 - border;                                        // border is a multiple of 32: x0 stays 32-aligned
    if (y == -1) y = 0;
    if (z == -1) z = 0;
    uint32_t word = 0;
    if (y >= 0 && y < sy && z >= 0 && z < sz) {
        const uint8_t* row0 = bytes + (size_t)y * sx + (size_t)z * ((size_t)sx * sy);
        if (x0 == -32) word = row0[0] ? 0x80000000u : 0u;                  // texel -1 along x
        else if (x0 >= 0 && x0 < sx) {
            const uint8_t* row = row0 + x0;
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
    }
    return word;
}
__global__ void __launch_bounds__(256) k_occ_texels(const uint8_t* __restrict__ bytes, int sx, int sy, int sz,
                                                    int xw, int cyp, int cz, int border, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)xw * cyp * cz;
    if (i >= total) return;
    const int ay = (int)(i % cyp), w = (int)((i / cyp) % xw), az = (int)(i / ((long long)cyp * xw));
    out[i] = texel_word(bytes, sx, sy, sz, border, w, ay, az);
}

// coarser plain level from a finer one: bit c = OR of the 2x2x2 child cells.  border_fine = 2 * border_coarse, so array index
// a of the coarse level has the children 2a, 2a + 1 of the fine array on every axis.
__device__ __forceinline__ uint32_t coarsen_word(const uint32_t* __restrict__ fine, int fcy, int fcz, int fxw, int fcyp, int w, int ay, int az) {
    uint32_t word = 0;
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy) {
            const int fy = 2 * ay + dy, fz = 2 * az + dz;
            if (fy >= fcy || fz >= fcz) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int fw = 2 * w + h;
                if (fw >= fxw) continue;
                uint32_t v = fine[level_index(fxw, fcyp, fw, fy, fz)];
                v = (v | (v >> 1)) & 0x55555555u;                  // OR of bit pairs at even positions
                v = (v | (v >> 1)) & 0x33333333u; v = (v | (v >> 2)) & 0x0F0F0F0Fu;
                v = (v | (v >> 4)) & 0x00FF00FFu; v = (v | (v >> 8)) & 0x0000FFFFu;   // compact to 16 bits
                word |= v << (16 * h);
            }
        }
    return word;
}
__global__ void __launch_bounds__(256) k_occ_coarsen(const uint32_t* __restrict__ fine, int fcy, int fcz, int fxw, int fcyp,
                                                     int xw, int cyp, int cz, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)xw * cyp * cz;
    if (i >= total) return;
    const int ay = (int)(i % cyp), w = (int)((i / cyp) % xw), az = (int)(i / ((long long)cyp * xw));
    out[i] = coarsen_word(fine, fcy, fcz, fxw, fcyp, w, ay, az);
}

// dilated level (array index = input array index + 1): OR over the 3x3x3 neighbourhood of the plain level
__device__ __forceinline__ uint32_t dilate_word(const uint32_t* __restrict__ in, int icy, int icz, int ixw, int icyp, int w, int oy, int oz) {
    const int y = oy - 1, z = oz - 1;                                      // input array coords
    // output bit b of word w is input x = 32*w + b - 1; gather input bits x-1, x, x+1 = input positions 32*w + b - 2 .. 32*w + b
    uint32_t word = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = y + dy, iz = z + dz;
            if (iy < 0 || iz < 0 || iy >= icy || iz >= icz) continue;
            const uint32_t cur = w < ixw ? in[level_index(ixw, icyp, w, iy, iz)] : 0u;
            const uint32_t prev = (w >= 1 && w - 1 < ixw) ? in[level_index(ixw, icyp, w - 1, iy, iz)] : 0u;
            // input position p = 32*w + b - s for s = 0, 1, 2  ->  (cur << s) | (prev >> (32 - s))
            word |= cur | (cur << 1) | (prev >> 31) | (cur << 2) | (prev >> 30);
        }
    return word;
}
__global__ void __launch_bounds__(256) k_occ_dilate(const uint32_t* __restrict__ in, int icy, int icz, int ixw, int icyp,
                                                    int xw, int cyp, int cz, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)xw * cyp * cz;
    if (i >= total) return;
    out[i] = dilate_word(in, icy, icz, ixw, icyp, (int)((i / cyp) % xw), (int)(i % cyp), (int)(i / ((long long)cyp * xw)));
}

// second copy of a level, shifted by 16 cells along x: bit (ax + 16) & 31 of word (ax + 16) >> 5.  A TMA box starts on a word
// boundary, so with the two copies a staged window can start every 16 cells instead of every 32.
__global__ void __launch_bounds__(256) k_occ_shift16(const uint32_t* __restrict__ in, int xw, int cyp, int cz, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)xw * cyp * cz;
    if (i >= total) return;
    const int w = (int)((i / cyp) % xw);
    out[i] = (in[i] << 16) | (w >= 1 ? in[i - cyp] >> 16 : 0u);          // the word before along x is cyp words back
}

// ---- the same levels, rebuilt only where the volume changed -------------------------------------------------------------------
// After a voxelise call the volume differs from the levels inside the entities' voxel boxes only (vxl_volume.cu keeps a tight box
// per command next to the reference's dirty regions).  One block row per box: the words of the level whose cells meet the box
// are recomputed by the very functions above, level after level, so the arrays end up identical to a full rebuild.
//   sh:    cell of the level = 2^sh texels (0 texel level, 1..3 plain levels; dilated levels: that of their plain level)
//   KIND:  0 texel level, 1 coarsen, 2 the shifted copy, 3 dilate (array index + 1, neighbourhood +-1)
struct BoxLevel { int sx, sy, sz; int sh; int xw, cyp, cy, cz; int fcy, fcz, fxw, fcyp; };
template <int KIND>
__global__ void __launch_bounds__(256) k_occ_boxes(const int* __restrict__ boxes, BoxLevel P, const void* __restrict__ src, uint32_t* __restrict__ out) {
    const int* b = boxes + (size_t)blockIdx.y * 6;
    int lo[3], hi[3];
    const int lim[3] = {P.sx, P.sy, P.sz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int mn = b[a], mx = b[3 + a];                                 // voxels, inclusive; an untouched command keeps mn > mx
        if (mn > mx || mx < 0 || mn >= 2 * lim[a]) return;
        const int t0 = max(mn >> 1, 0) - 1 + 96, t1 = min(mx >> 1, lim[a] - 1) + 96;    // texel-array indices; -1: the repeated texel
        lo[a] = (t0 >> P.sh) + (KIND == 3 ? 0 : 0);
        hi[a] = (t1 >> P.sh) + (KIND == 3 ? 2 : 0);
    }
    const int w0 = lo[0] >> 5, w1 = min((hi[0] >> 5) + (KIND == 2 ? 1 : 0), P.xw - 1);
    const int y0 = lo[1], y1 = min(hi[1], P.cyp - 1), z0 = lo[2], z1 = min(hi[2], P.cz - 1);
    const int nW = w1 - w0 + 1, nY = y1 - y0 + 1, nZ = z1 - z0 + 1;
    if (nW <= 0 || nY <= 0 || nZ <= 0) return;
    const int total = nW * nY * nZ;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ay = y0 + i % nY, w = w0 + (i / nY) % nW, az = z0 + i / (nY * nW);
        const size_t o = level_index(P.xw, P.cyp, w, ay, az);
        if (KIND == 0) out[o] = texel_word((const uint8_t*)src, P.sx, P.sy, P.sz, 96, w, ay, az);
        else if (KIND == 1) out[o] = coarsen_word((const uint32_t*)src, P.fcy, P.fcz, P.fxw, P.fcyp, w, ay, az);
        else if (KIND == 2) { const uint32_t* in = (const uint32_t*)src; out[o] = (in[o] << 16) | (w >= 1 ? in[o - P.cyp] >> 16 : 0u); }
        else out[o] = dilate_word((const uint32_t*)src, P.fcy, P.fcz, P.fxw, P.fcyp, w, ay, az);
    }
}

// ncx, ncy, ncz: cells of the level proper; border: zero cells added on every side; copies: 1, or 2 = also the copy shifted by 16 cells
static int alloc_level(BitLevel& L, int shift, int ncx, int ncy, int ncz, int border, int copies = 1) {
    L.shift = shift; L.border = border; L.copies = copies;
    L.cx = ncx + 2 * border; L.cy = ncy + 2 * border; L.cz = ncz + 2 * border;
    L.xw = (L.cx + 31) / 32 + 1;                              // one spare (all-zero) word per row: the staging funnel shift / the shifted copy
    L.cyp = (L.cy + 3) & ~3;                                  // rows of whole 16 bytes (TMA strides)
    VXL_CUDA(cudaMalloc(&L.d_words, (size_t)L.xw * L.cyp * L.cz * 4 * copies));
    for (auto& b : L.box) b.ty = 0;
    return VXL_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libvxl.so does not link libcuda, so it still loads on a box without a driver)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int level_tensor_map(BitLevel& L, int tw, int ty, int tz, const CUtensorMap** out) {
    const int key = (tw << 20) | (ty << 10) | tz;
    for (auto& b : L.box)
        if (b.ty == key) { *out = &b.map; return VXL_OK; }
    BitLevel::BoxMap* slot = nullptr;
    for (auto& b : L.box)
        if (b.ty == 0) { slot = &b; break; }
    if (!slot || !L.d_words || tw < 1 || ty < 4 || tz < 1 || tw > 256 || ty > 256 || tz > 256 || (ty & 3)) { set_error("level_tensor_map: no free slot / bad box"); return VXL_ERR_INVALID; }
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        VXL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return VXL_ERR_CUDA; }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    // the level (words [copy][az][word][ay] in memory) as a tensor with dimensions (y, z, x word, copy): a box is ty cells along y
    // (a multiple of 4) x tz slices x tw words of one copy and lands densely in shared memory in that order, [x word][z][y];
    // everything outside the array reads as zero = empty (texelFetch out of range)
    const cuuint64_t plane = (cuuint64_t)L.cyp * (cuuint64_t)L.xw * (cuuint64_t)L.cz * 4u;
    const cuuint64_t dims[4] = {(cuuint64_t)L.cyp, (cuuint64_t)L.cz, (cuuint64_t)L.xw, (cuuint64_t)L.copies};
    const cuuint64_t strides[3] = {(cuuint64_t)L.cyp * (cuuint64_t)L.xw * 4u, (cuuint64_t)L.cyp * 4u, plane};
    const cuuint32_t box[4] = {(cuuint32_t)ty, (cuuint32_t)tz, (cuuint32_t)tw, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = encode(&slot->map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, L.d_words, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult " + std::to_string((int)r)); return VXL_ERR_CUDA; }
    slot->ty = key;
    *out = &slot->map;
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
        // borders: 192 voxels of zero cells around every plain level (a multiple of 32 texels, so that the texel level's words
        // stay aligned with 32-byte pieces of the volume rows); each level's border is half the next finer one's
        if (int e = alloc_level(v->tex, 1, v->sx, v->sy, v->sz, 96)) return e;
        for (int li = 0; li < 3; ++li) {
            const int tpc = 2 << li;                          // texels per cell edge: 2, 4, 8
            if (int e = alloc_level(v->occ[li], 2 + li, (v->sx + tpc - 1) / tpc, (v->sy + tpc - 1) / tpc, (v->sz + tpc - 1) / tpc, 48 >> li, li == 0 ? 2 : 1)) return e;
        }
        for (int li = 0; li < 2; ++li) {
            const BitLevel& P = v->occ[1 + li];
            if (int e = alloc_level(v->dil[li], P.shift, P.cx - 2 * P.border, P.cy - 2 * P.border, P.cz - 2 * P.border, P.border + 1)) return e;
        }
    }
    if (v->dirty_partial && v->n_dirty_boxes > 0 && v->occ[0].d_words && v->dirty_gen == c->aabb_gen) {
        // the levels exist and only the boxes of the last voxelise call changed
        const dim3 grid(4u, (unsigned)v->n_dirty_boxes);
        auto params = [&](const BitLevel& L, int sh, const BitLevel* F) {
            BoxLevel P{v->sx, v->sy, v->sz, sh, L.xw, L.cyp, L.cy, L.cz, F ? F->cy : 0, F ? F->cz : 0, F ? F->xw : 0, F ? F->cyp : 0};
            return P;
        };
        k_occ_boxes<0><<<grid, 256, 0, c->stream>>>(v->dirty_boxes, params(v->tex, 0, nullptr), v->d_bytes, v->tex.d_words);
        VXL_LAUNCH_CHECK(c);
        for (int li = 0; li < 3; ++li) {
            const BitLevel& F = li ? v->occ[li - 1] : v->tex;
            BitLevel& L = v->occ[li];
            k_occ_boxes<1><<<grid, 256, 0, c->stream>>>(v->dirty_boxes, params(L, li + 1, &F), F.d_words, L.d_words);
            VXL_LAUNCH_CHECK(c);
            if (L.copies == 2) {
                k_occ_boxes<2><<<grid, 256, 0, c->stream>>>(v->dirty_boxes, params(L, li + 1, nullptr), L.d_words, L.d_words + (size_t)L.xw * L.cyp * L.cz);
                VXL_LAUNCH_CHECK(c);
            }
        }
        for (int li = 0; li < 2; ++li) {
            const BitLevel& Pl = v->occ[1 + li];
            BitLevel& D = v->dil[li];
            k_occ_boxes<3><<<grid, 256, 0, c->stream>>>(v->dirty_boxes, params(D, li + 2, &Pl), Pl.d_words, D.d_words);
            VXL_LAUNCH_CHECK(c);
        }
        v->dirty = false; v->dirty_partial = false;
        return VXL_OK;
    }
    auto grid_of = [](const BitLevel& L) { return (unsigned)(((long long)L.xw * L.cyp * L.cz + 255) / 256); };
    BitLevel& L1 = v->tex;
    k_occ_texels<<<grid_of(L1), 256, 0, c->stream>>>(v->d_bytes, v->sx, v->sy, v->sz, L1.xw, L1.cyp, L1.cz, L1.border, L1.d_words);
    VXL_LAUNCH_CHECK(c);
    for (int li = 0; li < 3; ++li) {
        const BitLevel& F = li ? v->occ[li - 1] : v->tex;
        BitLevel& L = v->occ[li];
        k_occ_coarsen<<<grid_of(L), 256, 0, c->stream>>>(F.d_words, F.cy, F.cz, F.xw, F.cyp, L.xw, L.cyp, L.cz, L.d_words);
        VXL_LAUNCH_CHECK(c);
        if (L.copies == 2) {
            k_occ_shift16<<<grid_of(L), 256, 0, c->stream>>>(L.d_words, L.xw, L.cyp, L.cz, L.d_words + (size_t)L.xw * L.cyp * L.cz);
            VXL_LAUNCH_CHECK(c);
        }
    }
    for (int li = 0; li < 2; ++li) {
        const BitLevel& P = v->occ[1 + li];
        BitLevel& D = v->dil[li];
        k_occ_dilate<<<grid_of(D), 256, 0, c->stream>>>(P.d_words, P.cy, P.cz, P.xw, P.cyp, D.xw, D.cyp, D.cz, D.d_words);
        VXL_LAUNCH_CHECK(c);
    }
    v->dirty = false; v->dirty_partial = false;
    return VXL_OK;
}

/* diagnostics: download one occupancy level unpacked to 0/1 bytes [cz][cy][cx]; level = 1, 2, 3, 4 (plain, cell =
 * 2^level voxels) or 13, 14 (dilated levels 3, 4, including their 1-cell border); out_dims = {cx, cy, cz} */
int vxl_volume_debug_occupancy(vxl_volume* v, int level, uint8_t* host_out, int* out_dims) {
    const bool shifted = level == 22;                        // level 2 read back from its copy shifted by 16 cells (what TMA boxes at odd origins see)
    if (shifted) level = 2;
    if (!v || !((level >= 1 && level <= 4) || level == 13 || level == 14)) { set_error("vxl_volume_debug_occupancy: bad argument"); return VXL_ERR_INVALID; }
    if (v->dirty || !v->occ[0].d_words) { if (int e = vxl_volume_build_occupancy(v)) return e; }
    const BitLevel& L = level == 1 ? v->tex : (level < 10 ? v->occ[level - 2] : v->dil[level - 13]);
    // the cells of the level proper; the dilated levels with one cell of border (cells -1 .. n)
    const int keep = level < 10 ? 0 : 1, skip = L.border - keep;
    const int nx = L.cx - 2 * skip, ny = L.cy - 2 * skip, nz = L.cz - 2 * skip;
    if (out_dims) { out_dims[0] = nx; out_dims[1] = ny; out_dims[2] = nz; }
    if (!host_out) return VXL_OK;
    const size_t words = (size_t)L.xw * L.cyp * L.cz;
    std::vector<uint32_t> h(words);
    if (shifted && L.copies != 2) { set_error("vxl_volume_debug_occupancy: level 2 has no shifted copy"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemcpyAsync(h.data(), L.d_words + (shifted ? words : 0), words * 4, cudaMemcpyDeviceToHost, v->ctx->stream));
    VXL_CUDA(cudaStreamSynchronize(v->ctx->stream));
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                const int ax = x + skip + (shifted ? 16 : 0), ay = y + skip, az = z + skip;
                host_out[((size_t)z * ny + y) * nx + x] = (uint8_t)((h[((size_t)az * L.xw + (ax >> 5)) * L.cyp + ay] >> (ax & 31)) & 1u);
            }
    return VXL_OK;
}

}  // extern "C"
