// vxl_bitmarch.cuh -- the two-phase fixed-step march against an occupancy-bit tile in shared memory.
//
// The arithmetic is the reference's, unchanged (Light.frag:131-173 / :175-217): the same float
// recurrence pos += stepDir, the same probe positions, the same texel tests on the same bytes.  The
// only difference is WHEN the volume is read: each probe first tests one bit of a per-block tile of
// the cell occupancy mask (vxl_occupancy.cu; cell = 2^SHIFT voxels).  A clear bit means every texel
// of that cell is zero, so the probe's texelFetch would return 0 and both the fine bit test (:149) and
// the coarse byte test (:163) fail -- the fetch is skipped.  A set bit falls through to the
// reference's fetch + test on the canonical bytes.  Results are bit-identical by construction.
//
// On config 3 93 % of AO probes and 98 % of sun-shadow probes land in cells whose bit is
// clear, so the volume bytes are touched ~1.5 times per AO ray instead of ~22 and the per-probe work
// is 3 FADD (recurrence) + 3 FADD.RZ (cell index) + address + one LDS + a bit test.
//
// A ray uses the tile only if its whole extent lies inside it at non-negative coordinates (there
// floor == the reference's truncation) and all its values are finite; any other ray runs the plain
// march of vxl_trace.cuh.  The tile is TW*32 x TY x TY cells around the centre of the block's ray
// origins; neighbouring pixels start within a few voxels of each other (median spread 2 voxels at 4K).
#pragma once
#include "vxl_internal.h"
#include "vxl_math.cuh"
#include "vxl_trace.cuh"

namespace vxl {

constexpr float BM_MARGIN = 0.125f;          // > accumulated rounding drift of <= 179 additions at |coords| < 8192
constexpr float BM_MAXCOORD = 8191.0f;

struct BitTile {
    const uint32_t* w;     // [TY][TY][TW] words; bit (x & 31) of word x >> 5
    int ox, oy, oz;        // tile origin, cells
    bool enabled;
};

// floor(p / 2^SHIFT) for 0 <= p < 2^(23+SHIFT)
template <int SHIFT>
VXL_DI int cell_floor(float p) {
#ifdef __CUDA_ARCH__
    // p + 2^(23+SHIFT), rounded toward zero, lies on the grid of spacing 2^SHIFT: its mantissa is the quotient
    constexpr int MB = 0x4B000000 + (SHIFT << 23);
    return __float_as_int(__fadd_rz(p, __int_as_float(MB))) - MB;
#else
    return (int)floorf(p * (1.0f / (float)(1 << SHIFT)));
#endif
}

VXL_DI float3 fma3(float3 s, float k, float3 o) {
#ifdef __CUDA_ARCH__
    return make_float3(__fmaf_rn(s.x, k, o.x), __fmaf_rn(s.y, k, o.y), __fmaf_rn(s.z, k, o.z));
#else
    return make_float3(fmaf(s.x, k, o.x), fmaf(s.y, k, o.y), fmaf(s.z, k, o.z));
#endif
}

// Tile lookup with the constant parts of the address folded once per ray:
//   raw_a = bits of (p_a + 2^(23+SHIFT)) rounded toward zero = MB + floor(p_a / cell)
//   word  = ((raw_z - MB - oz) * TY + (raw_y - MB - oy)) * TW + ((raw_x - MB - ox) >> 5)
//         = raw_z * (TY*TW) + raw_y * TW + ((raw_x - kx) >> 5) + c0          (32-bit wrap-around arithmetic)
template <int SHIFT, int TY, int TW>
struct TileAddr {
    const uint32_t* w;
    int kx, c0;
    VXL_DI TileAddr(const BitTile& T) {
        constexpr int MB = 0x4B000000 + (SHIFT << 23);
        w = T.w;
        kx = MB + T.ox;
        c0 = (int)(0u - (unsigned)(MB + T.oz) * (unsigned)(TY * TW) - (unsigned)(MB + T.oy) * (unsigned)TW);
    }
    // non-zero iff the occupancy bit of the cell containing p is set
    VXL_DI unsigned test(float3 p) const {
#ifdef __CUDA_ARCH__
        constexpr float M = (float)(1 << 23) * (float)(1 << SHIFT);
        const int bx = __float_as_int(__fadd_rz(p.x, M)), by = __float_as_int(__fadd_rz(p.y, M)), bz = __float_as_int(__fadd_rz(p.z, M));
#else
        constexpr int MB = 0x4B000000 + (SHIFT << 23);
        const int bx = MB + (int)floorf(p.x * (1.0f / (float)(1 << SHIFT))), by = MB + (int)floorf(p.y * (1.0f / (float)(1 << SHIFT))),
                  bz = MB + (int)floorf(p.z * (1.0f / (float)(1 << SHIFT)));
#endif
        const int rx = bx - kx;
        const unsigned idx = (unsigned)bz * (unsigned)(TY * TW) + (unsigned)by * (unsigned)TW + (unsigned)(rx >> 5) + (unsigned)c0;
        return w[idx] & (1u << (rx & 31));
    }
};

template <bool SUPER>
VXL_DI int phase2_count(float lim) {
    constexpr float d0 = SUPER ? 17.5f : 16.0f;
    if (!(lim > d0)) return 0;
    if (!SUPER) return (int)ceilf(lim - d0);               // lim - 16 is exact
    int n = (int)ceilf((lim - d0) / 5.0f);                  // the division rounds: fix up against the exact sequence
    while (n > 0 && d0 + 5.0f * (float)(n - 1) >= lim) --n;
    while (d0 + 5.0f * (float)n < lim) ++n;
    return n;
}

// SUPER = false: raycastShadowVolumeSparse (step 0.5 then 1); true: ...SuperSparse (step 2.5 then 5).
// `fetched` counts probes that had to read the volume (diagnostics).
template <bool SUPER, bool RECORD, int SHIFT, int TY, int TW>
VXL_DI float march_bits(const VolView& V, const BitTile& T, float3 origin, float3 dir, float dist, int& steps_out,
                        MarchResult* rec, unsigned& fetched) {
    constexpr float step0 = SUPER ? 2.5f : 0.5f;
    constexpr float step2 = SUPER ? 5.0f : 1.0f;
    constexpr int n1 = SUPER ? 6 : 31;                      // probes with d = step0*(k+1) < 16 (the d sequence is exact)
    constexpr float d0 = SUPER ? 17.5f : 16.0f;             // d entering phase 2
    const float lim = fminf(dist, 164.0f);                  // lod1MaxT (:157)

    // ---- eligibility: whole ray inside the tile, at non-negative coordinates, finite ----
    const float reach = fmaxf(lim, 16.0f) + 1.0f;
    const float3 end = fma3(dir, reach, origin);
    const float3 lo = make_float3(fminf(origin.x, end.x), fminf(origin.y, end.y), fminf(origin.z, end.z));
    const float3 hi = make_float3(fmaxf(origin.x, end.x), fmaxf(origin.y, end.y), fmaxf(origin.z, end.z));
    const float cell = (float)(1 << SHIFT);
    const float3 tlo = make_float3((float)T.ox * cell, (float)T.oy * cell, (float)T.oz * cell);
    bool fast = T.enabled;
    fast = fast && (lo.x >= fmaxf(tlo.x, 0.0f) + BM_MARGIN) && (lo.y >= fmaxf(tlo.y, 0.0f) + BM_MARGIN) && (lo.z >= fmaxf(tlo.z, 0.0f) + BM_MARGIN);
    fast = fast && (hi.x <= fminf(tlo.x + (float)(TW * 32) * cell, BM_MAXCOORD) - BM_MARGIN) &&
           (hi.y <= fminf(tlo.y + (float)TY * cell, BM_MAXCOORD) - BM_MARGIN) && (hi.z <= fminf(tlo.z + (float)TY * cell, BM_MAXCOORD) - BM_MARGIN);
    // fminf/fmaxf drop a NaN operand, so test the inputs themselves too
    fast = fast && (origin.x == origin.x) && (origin.y == origin.y) && (origin.z == origin.z) && (end.x == end.x) && (end.y == end.y) && (end.z == end.z);
    if (!fast) return march<RECORD>(V, origin, dir, dist, step0, steps_out, rec);

    const TileAddr<SHIFT, TY, TW> A(T);
    float3 stepDir = dir * step0;
    float3 pos = origin;

    // ---- phase 1 (:138-154): fine steps, position-hashed bit of the texel ----
    int k = 0;
    while (true) {
        // tight scan: advance while the cell's occupancy bit is clear
        while (k < n1 && !A.test(pos)) { pos = pos + stepDir; ++k; }
        if (k >= n1) break;
        {
            ++fetched;
            const int tx = f2i(pos.x / 2.0f), ty = f2i(pos.y / 2.0f), tz = f2i(pos.z / 2.0f);
            const unsigned v = fetch_texel(V, tx, ty, tz);
            unsigned bit = 0u;
            bit += gmod(pos.x, 0.5f) > 0.25f ? 1u : 0u;
            bit += gmod(pos.y, 0.5f) > 0.25f ? 2u : 0u;
            bit += gmod(pos.z, 0.5f) > 0.25f ? 4u : 0u;
            if ((v >> bit) & 1u) {
                steps_out += k + 1;
                const float d = step0 * (float)(k + 1);
                if (RECORD) {
                    rec->d = d; rec->steps = k + 1; rec->status = 1;
                    rec->vx = tx * 2 + (int)(bit & 1u); rec->vy = ty * 2 + (int)((bit >> 1) & 1u); rec->vz = tz * 2 + (int)((bit >> 2) & 1u);
                    rec->pos = pos;
                }
                return d;
            }
        }
        pos = pos + stepDir; ++k;
    }

    // ---- phase 2 (:156-170): doubled steps, byte != 0 ----
    stepDir = stepDir * 2.0f;
    const int n2 = phase2_count<SUPER>(lim);
    int j = 0;
    while (true) {
        while (j < n2 && !A.test(pos)) { pos = pos + stepDir; ++j; }
        if (j >= n2) break;
        {
            ++fetched;
            const int px = f2i(pos.x), py = f2i(pos.y), pz = f2i(pos.z);
            if (fetch_texel(V, px / 2, py / 2, pz / 2) != 0u) {   // getVolumeAt(ivec3(pos), 1)
                steps_out += n1 + j + 1;
                const float d = d0 + step2 * (float)j;
                if (RECORD) {
                    rec->d = d; rec->steps = n1 + j + 1; rec->status = 2;
                    rec->vx = px; rec->vy = py; rec->vz = pz;
                    rec->pos = pos;
                }
                return d;
            }
        }
        pos = pos + stepDir; ++j;
    }
    steps_out += n1 + n2;
    if (RECORD) { rec->d = dist; rec->steps = n1 + n2; rec->status = 0; rec->vx = rec->vy = rec->vz = 0; rec->pos = make_float3(0.f, 0.f, 0.f); }
    return dist;
}

// Stage the TW*32 x TY x TY-cell window of an occupancy level whose origin cell is (ox, oy, oz).
// Everything outside the level's array is empty.
#ifdef __CUDACC__
template <int TY, int TW>
__device__ __forceinline__ void stage_bits(uint32_t* __restrict__ dst, const BitView& M, int ox, int oy, int oz) {
    const int w0 = ox >> 5;                                 // arithmetic shift: floor for negatives
    const int sh = ox & 31;
    const int words_x = M.pitch - 1;                        // the spare word of each row is zero
    for (int i = threadIdx.x; i < TY * TY * TW; i += blockDim.x) {
        const int xw = i % TW, y = (i / TW) % TY, z = i / (TW * TY);
        const int ay = oy + y, az = oz + z;
        uint32_t a = 0u, b = 0u;
        if ((unsigned)ay < (unsigned)M.cy && (unsigned)az < (unsigned)M.cz) {
            const uint32_t* row = M.words + ((size_t)az * M.cy + ay) * M.pitch;
            const int wa = w0 + xw, wb = wa + 1;
            if ((unsigned)wa < (unsigned)words_x) a = __ldg(row + wa);
            if ((unsigned)wb < (unsigned)words_x) b = __ldg(row + wb);
        }
        dst[i] = __funnelshift_r(a, b, sh);                 // sh == 0 returns a
    }
}
#endif

}  // namespace vxl
