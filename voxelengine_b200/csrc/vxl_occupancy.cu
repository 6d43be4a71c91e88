// vxl_occupancy.cu -- derived occupancy levels of the world volume: Chebyshev clearance maps.
//
// Pure acceleration (SURVEY.md 7 step 5 "k_occupancy_mips"): nothing here exists in the reference
// (its volume has one mip level, ShadowVoxSystem.cpp:61) and nothing here may change a result.
//
// For a cell size of 2^L voxels (L = 2: 4 voxels = 2x2x2 texels; L = 4: 16 voxels = 8x8x8 texels)
//   base_L[c]  = 1 iff any packed byte of the canonical volume inside cell c is non-zero
//   R_L[c]     = min(cap, Chebyshev distance in cells from c to the nearest cell with base_L = 1)
// so R_L[c] >= r  =>  every texel of every cell within Chebyshev distance r-1 of c is 0, hence every
// occupancy probe (Light.frag:140 fine bit test, :163 coarse byte test) landing there misses.
// Out-of-volume cells count as empty (texelFetch out of range reads 0).
// R is stored as nibbles, 8 cells per 32-bit word along x ("clearance map", CM_L).
//
// The L-infinity distance transform is separable: three passes of
//   out[p] = min_{|k| <= cap} max(in[p + k*e_axis], |k|)          (out-of-range taps = cap)
// starting from in0 = base ? 0 : cap.
#include "vxl_internal.h"

namespace vxl {

// base cells of level 2 from the packed bytes: one thread per cell
__global__ void __launch_bounds__(256) k_occ_base4(const uint8_t* __restrict__ bytes, int sx, int sy, int sz,
                                                   int cx, int cy, int cz, int border, uint8_t cap, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)cx * cy * cz;
    if (i >= total) return;
    const int x = (int)(i % cx) - border, y = (int)((i / cx) % cy) - border, z = (int)(i / ((long long)cx * cy)) - border;
    unsigned any = 0;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int tx = 2 * x + dx, ty = 2 * y + dy, tz = 2 * z + dz;
                if (tx >= 0 && ty >= 0 && tz >= 0 && tx < sx && ty < sy && tz < sz) any |= bytes[(size_t)tx + (size_t)ty * sx + (size_t)tz * ((size_t)sx * sy)];
            }
    out[i] = any ? (uint8_t)0 : cap;
}

// base cells of a coarser level: OR over F^3 children of the finer level's initial array (0 = occupied)
// (fb / cb = border of the fine / coarse array; cell coordinates are array index - border)
__global__ void __launch_bounds__(256) k_occ_coarsen(const uint8_t* __restrict__ fine, int fx, int fy, int fz, int fb, int F,
                                                     int cx, int cy, int cz, int cb, uint8_t cap, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)cx * cy * cz;
    if (i >= total) return;
    const int x = (int)(i % cx) - cb, y = (int)((i / cx) % cy) - cb, z = (int)(i / ((long long)cx * cy)) - cb;
    bool any = false;
    for (int dz = 0; dz < F && !any; ++dz)
        for (int dy = 0; dy < F && !any; ++dy)
            for (int dx = 0; dx < F; ++dx) {
                const int ax = x * F + dx + fb, ay = y * F + dy + fb, az = z * F + dz + fb;
                if (ax >= 0 && ay >= 0 && az >= 0 && ax < fx && ay < fy && az < fz && fine[(size_t)ax + (size_t)ay * fx + (size_t)az * ((size_t)fx * fy)] == 0) { any = true; break; }
            }
    out[i] = any ? (uint8_t)0 : cap;
}

// one separable pass of the capped Chebyshev distance transform along `axis`
__global__ void __launch_bounds__(256) k_occ_dt(const uint8_t* __restrict__ in, int cx, int cy, int cz, int axis, int cap,
                                                uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)cx * cy * cz;
    if (i >= total) return;
    const int x = (int)(i % cx), y = (int)((i / cx) % cy), z = (int)(i / ((long long)cx * cy));
    const int p = axis == 0 ? x : (axis == 1 ? y : z);
    const int n = axis == 0 ? cx : (axis == 1 ? cy : cz);
    const long long stride = axis == 0 ? 1 : (axis == 1 ? cx : (long long)cx * cy);
    int best = in[i];
    for (int k = 1; k < best; ++k) {          // taps farther than the current best cannot improve it
        if (p - k >= 0) best = min(best, max((int)in[i - k * stride], k));
        if (p + k < n) best = min(best, max((int)in[i + k * stride], k));
    }
    out[i] = (uint8_t)min(best, cap);
}

// bytes -> nibbles, 8 cells per word along x; cells past the padded array read `fill` = min(15, border + 1)
__global__ void __launch_bounds__(256) k_occ_pack(const uint8_t* __restrict__ in, int cx, int cy, int cz, int pitch, uint32_t fill,
                                                  uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)pitch * cy * cz;
    if (i >= total) return;
    const int w = (int)(i % pitch);
    const long long row = i / pitch;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int x = w * 8 + k;
        const uint32_t r = x < cx ? (uint32_t)in[row * cx + x] : fill;
        v |= (r & 15u) << (4 * k);
    }
    out[i] = v;
}

static int build_level(vxl_volume* v, vxl::ClearLevel& L, const uint8_t* init, uint8_t* a, uint8_t* b) {
    vxl_ctx* c = v->ctx;
    const long long total = (long long)L.cx * L.cy * L.cz;
    const unsigned grid = (unsigned)((total + 255) / 256);
    // init is in `a`
    (void)init;
    k_occ_dt<<<grid, 256, 0, c->stream>>>(a, L.cx, L.cy, L.cz, 0, L.cap, b);
    VXL_LAUNCH_CHECK(c);
    k_occ_dt<<<grid, 256, 0, c->stream>>>(b, L.cx, L.cy, L.cz, 1, L.cap, a);
    VXL_LAUNCH_CHECK(c);
    k_occ_dt<<<grid, 256, 0, c->stream>>>(a, L.cx, L.cy, L.cz, 2, L.cap, b);
    VXL_LAUNCH_CHECK(c);
    const long long words = (long long)L.pitch * L.cy * L.cz;
    k_occ_pack<<<(unsigned)((words + 255) / 256), 256, 0, c->stream>>>(b, L.cx, L.cy, L.cz, L.pitch, (uint32_t)(L.cap + 1 > 15 ? 15 : L.cap + 1), L.d_words);
    VXL_LAUNCH_CHECK(c);
    return VXL_OK;
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_volume_build_occupancy(vxl_volume* v) {
    if (!v) { set_error("vxl_volume_build_occupancy: vol is NULL"); return VXL_ERR_INVALID; }
    vxl_ctx* c = v->ctx;
    VXL_CUDA(cudaSetDevice(c->device));
    ClearLevel& A = v->cm4;
    ClearLevel& B = v->cm16;
    if (!A.d_words) {
        A.shift = 2; A.cap = 8;
        A.cx = (v->sx + 1) / 2 + 2 * A.cap; A.cy = (v->sy + 1) / 2 + 2 * A.cap; A.cz = (v->sz + 1) / 2 + 2 * A.cap;
        A.pitch = (A.cx + 7) / 8 + 1;
        B.shift = 4; B.cap = 15;
        B.cx = (v->sx + 7) / 8 + 2 * B.cap; B.cy = (v->sy + 7) / 8 + 2 * B.cap; B.cz = (v->sz + 7) / 8 + 2 * B.cap;
        B.pitch = (B.cx + 7) / 8 + 1;
        const size_t cells = (size_t)A.cx * A.cy * A.cz;
        VXL_CUDA(cudaMalloc(&A.d_words, (size_t)A.pitch * A.cy * A.cz * 4));
        VXL_CUDA(cudaMalloc(&B.d_words, (size_t)B.pitch * B.cy * B.cz * 4));
        const size_t cellsB0 = (size_t)B.cx * B.cy * B.cz;
        VXL_CUDA(cudaMalloc(&v->d_scratch, cells * 2 + cellsB0 * 2));   // ping-pong pairs of both levels
    }
    const size_t cellsA = (size_t)A.cx * A.cy * A.cz, cellsB = (size_t)B.cx * B.cy * B.cz;
    uint8_t* a = v->d_scratch;
    uint8_t* b = a + cellsA;
    uint8_t* a2 = b + cellsA;
    uint8_t* b2 = a2 + cellsB;
    k_occ_base4<<<(unsigned)((cellsA + 255) / 256), 256, 0, c->stream>>>(v->d_bytes, v->sx, v->sy, v->sz, A.cx, A.cy, A.cz, A.cap, (uint8_t)A.cap, a);
    VXL_LAUNCH_CHECK(c);
    k_occ_coarsen<<<(unsigned)((cellsB + 255) / 256), 256, 0, c->stream>>>(a, A.cx, A.cy, A.cz, A.cap, 4, B.cx, B.cy, B.cz, B.cap, (uint8_t)B.cap, a2);
    VXL_LAUNCH_CHECK(c);
    if (int e = build_level(v, A, a, a, b)) return e;
    if (int e = build_level(v, B, a2, a2, b2)) return e;
    v->dirty = false;
    return VXL_OK;
}

/* diagnostics: download one clearance level (padded array incl. border) unpacked to bytes [cz][cy][cx];
 * out_dims = {cx, cy, cz, border} */
int vxl_volume_debug_clearance(vxl_volume* v, int level, uint8_t* host_out, int* out_dims) {
    if (!v || (level != 2 && level != 4)) { set_error("vxl_volume_debug_clearance: bad argument"); return VXL_ERR_INVALID; }
    if (v->dirty || !v->cm4.d_words) { if (int e = vxl_volume_build_occupancy(v)) return e; }
    const ClearLevel& L = level == 2 ? v->cm4 : v->cm16;
    if (out_dims) { out_dims[0] = L.cx; out_dims[1] = L.cy; out_dims[2] = L.cz; out_dims[3] = L.cap; }
    if (!host_out) return VXL_OK;
    const size_t words = (size_t)L.pitch * L.cy * L.cz;
    std::vector<uint32_t> h(words);
    VXL_CUDA(cudaMemcpyAsync(h.data(), L.d_words, words * 4, cudaMemcpyDeviceToHost, v->ctx->stream));
    VXL_CUDA(cudaStreamSynchronize(v->ctx->stream));
    for (int z = 0; z < L.cz; ++z)
        for (int y = 0; y < L.cy; ++y)
            for (int x = 0; x < L.cx; ++x) {
                const uint32_t w = h[((size_t)z * L.cy + y) * L.pitch + (x >> 3)];
                host_out[((size_t)z * L.cy + y) * L.cx + x] = (uint8_t)((w >> (4 * (x & 7))) & 15u);
            }
    return VXL_OK;
}

}  // extern "C"
