// vxl_fastmarch.cuh -- bit-exact accelerated form of the two-phase fixed-step march.
//
// The reference march (Light.frag:131-173 / :175-217) visits probe k at
//     pos_k = (...((origin + s) + s) ... + s)            (k float additions, phase 1: s = dir*step0,
//                                                         phase 2: s = 2*dir*step0)
// and returns at the first probe whose texel test succeeds.  Its result depends on pos_k only for
// probes that can hit.  This header proves most probes cannot:
//
//   * clearance maps (vxl_occupancy.cu) give, for the cell containing a point p, a radius R such that
//     every texel of every cell within Chebyshev distance R-1 is zero;
//   * a probe's position is known to within DRIFT of the closed form  p~_k = origin + k*s  (one FMA),
//     because each of the <= 179 additions rounds by at most ulp/2 (coordinates are checked < 8192);
//   * so one lookup at p~_k with R >= 2 clears probe k and the next floor(((R-1)*cell - MARGIN)/max|s_a|)
//     probes without touching the volume -- and without computing their exact positions.
//
// Only probes that cannot be cleared are executed exactly: the float recurrence is replayed from the
// last exact state to that probe (3 FADD per step, no memory access) and the reference's texel test
// is applied to the exact position.  Returned distance, probe count and hit record are therefore
// bit-identical to the plain march; "steps" still counts every probe the reference performs.
//
// Per thread block the relevant part of both clearance maps is staged in shared memory (32^3 cells
// each, 16 KB): level 2 (4-voxel cells, +-64 voxels around the block's ray origins) for the near
// field, level 4 (16-voxel cells, +-256 voxels) for the far field.  Rays that leave the staged
// region, start at negative coordinates (where ivec3() truncation differs from floor) or carry
// non-finite values take the plain march.
#pragma once
#include "vxl_internal.h"
#include "vxl_math.cuh"
#include "vxl_trace.cuh"

namespace vxl {

constexpr int CT = 32;                       // clearance tile edge, cells
constexpr int CTW = CT / 8;                  // words per tile row
constexpr int CT_WORDS = CT * CT * CTW;      // 4096 words = 16 KB
constexpr float FM_MARGIN = 0.125f;          // > accumulated rounding drift for |coords| < 8192, <= 179 steps
constexpr float FM_MAXCOORD = 8191.0f;

struct TileRef {
    const uint32_t* w;     // [CT][CT][CTW], nibble (x & 7) of word x >> 3
    int ox, oy, oz;        // tile origin, cells
};
struct FastCtx {
    TileRef t4, t16;
    bool enabled;
};

// floor(p / 2^SHIFT) for 0 <= p < 2^(23+SHIFT)
template <int SHIFT>
VXL_DI int cell_floor(float p) {
#ifdef __CUDA_ARCH__
    // p + 2^(23+SHIFT) rounded toward zero lands on the grid of spacing 2^SHIFT: the mantissa is the quotient
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

// clearance radius at p (0 when p lies outside the staged tile)
template <int SHIFT>
VXL_DI int tile_clearance(const TileRef& T, float3 p, bool& inside) {
    const int rx = cell_floor<SHIFT>(p.x) - T.ox, ry = cell_floor<SHIFT>(p.y) - T.oy, rz = cell_floor<SHIFT>(p.z) - T.oz;
    inside = ((unsigned)rx < (unsigned)CT) & ((unsigned)ry < (unsigned)CT) & ((unsigned)rz < (unsigned)CT);
    if (!inside) return 0;
    const uint32_t w = T.w[(rz * CT + ry) * CTW + (rx >> 3)];
    return (int)((w >> ((rx & 7) * 4)) & 15u);
}

// number of probes after the queried one that are provably empty (R >= 2)
VXL_DI int clear_run(int R, float cell, float invm) {
    const float a = ((float)(R - 1) * cell - FM_MARGIN) * invm;
    return (int)fminf(a, 4096.0f);
}

// Light.frag:139-149: fine-phase texel + position-hashed bit select (sic)
VXL_DI bool fine_probe(const VolView& V, float3 pos, int& tx, int& ty, int& tz, unsigned& bit) {
    tx = f2i(pos.x / 2.0f); ty = f2i(pos.y / 2.0f); tz = f2i(pos.z / 2.0f);
    const unsigned v = fetch_texel(V, tx, ty, tz);
    bit = 0u;
    bit += gmod(pos.x, 0.5f) > 0.25f ? 1u : 0u;
    bit += gmod(pos.y, 0.5f) > 0.25f ? 2u : 0u;
    bit += gmod(pos.z, 0.5f) > 0.25f ? 4u : 0u;
    return ((v >> bit) & 1u) != 0u;
}

template <bool SUPER>
VXL_DI int phase2_count(float lim) {
    constexpr float d0 = SUPER ? 17.5f : 16.0f;
    if (!(lim > d0)) return 0;
    if (!SUPER) return (int)ceilf(lim - d0);               // lim - 16 is exact
    int n = (int)ceilf((lim - d0) / 5.0f);                  // division rounds: fix up against the exact sequence
    while (n > 0 && d0 + 5.0f * (float)(n - 1) >= lim) --n;
    while (d0 + 5.0f * (float)n < lim) ++n;
    return n;
}

// SUPER = false: raycastShadowVolumeSparse (step 0.5 / 1); true: ...SuperSparse (step 2.5 / 5).
// `exact` counts probes executed exactly (diagnostics).
template <bool SUPER, bool RECORD>
VXL_DI float march_fast(const VolView& V, const FastCtx& C, float3 origin, float3 dir, float dist, int& steps_out,
                        MarchResult* rec, unsigned& exact) {
    constexpr float step0 = SUPER ? 2.5f : 0.5f;
    constexpr float step2 = SUPER ? 5.0f : 1.0f;
    constexpr int n1 = SUPER ? 6 : 31;                      // probes with d = step0*(k+1) < 16
    constexpr float d0 = SUPER ? 17.5f : 16.0f;             // d entering phase 2
    const float lim = fminf(dist, 164.0f);                  // lod1MaxT
    const int n2 = phase2_count<SUPER>(lim);

    // ---- eligibility: the whole ray inside the far tile, at non-negative coordinates, finite ----
    const float reach = fmaxf(lim, 16.0f) + 1.0f;
    const float3 end = fma3(dir, reach, origin);
    const float3 lo = make_float3(fminf(origin.x, end.x), fminf(origin.y, end.y), fminf(origin.z, end.z));
    const float3 hi = make_float3(fmaxf(origin.x, end.x), fmaxf(origin.y, end.y), fmaxf(origin.z, end.z));
    const float3 tlo = make_float3((float)(C.t16.ox * 16), (float)(C.t16.oy * 16), (float)(C.t16.oz * 16));
    bool fast = C.enabled;
    fast = fast && (lo.x >= fmaxf(tlo.x, 0.0f) + FM_MARGIN) && (lo.y >= fmaxf(tlo.y, 0.0f) + FM_MARGIN) && (lo.z >= fmaxf(tlo.z, 0.0f) + FM_MARGIN);
    fast = fast && (hi.x <= fminf(tlo.x + (float)(CT * 16), FM_MAXCOORD) - FM_MARGIN) &&
           (hi.y <= fminf(tlo.y + (float)(CT * 16), FM_MAXCOORD) - FM_MARGIN) &&
           (hi.z <= fminf(tlo.z + (float)(CT * 16), FM_MAXCOORD) - FM_MARGIN);
    // fminf/fmaxf drop a NaN operand, so test the inputs themselves too
    fast = fast && (origin.x == origin.x) && (origin.y == origin.y) && (origin.z == origin.z) && (end.x == end.x) && (end.y == end.y) && (end.z == end.z);
    if (!fast) return march<RECORD>(V, origin, dir, dist, step0, steps_out, rec);

    const float3 s1 = dir * step0;                          // stepDir, phase 1
    const float3 s2 = s1 * 2.0f;                            // stepDir, phase 2 (exact doubling)
    const float m1 = fmaxf(fmaxf(fabsf(s1.x), fabsf(s1.y)), fabsf(s1.z));
    const float invm1 = 0.999f / m1;                        // +inf for a zero direction: the ray never moves
    const float invm2 = invm1 * 0.5f;

    float3 pe = origin;                                     // exact position of probe ke
    int ke = 0;
    bool use4 = true;                                       // still inside the near tile

    // ---- phase 1: fine steps, bit test ----
    int k = 0;
    while (k < n1) {
        if (use4) {
            const int R = tile_clearance<2>(C.t4, fma3(s1, (float)k, origin), use4);
            if (R >= 2) { k += 1 + clear_run(R, 4.0f, invm1); continue; }
        }
        while (ke < k) { pe = pe + s1; ++ke; }
        ++exact;
        int tx, ty, tz; unsigned bit;
        if (fine_probe(V, pe, tx, ty, tz, bit)) {
            steps_out += k + 1;
            const float d = step0 * (float)(k + 1);
            if (RECORD) {
                rec->d = d; rec->steps = k + 1; rec->status = 1;
                rec->vx = tx * 2 + (int)(bit & 1u); rec->vy = ty * 2 + (int)((bit >> 1) & 1u); rec->vz = tz * 2 + (int)((bit >> 2) & 1u);
                rec->pos = pe;
            }
            return d;
        }
        ++k;
    }

    // ---- phase 2: coarse steps, byte test ----
    const float3 P0 = fma3(s1, (float)n1, origin);          // ~ position of phase-2 probe 0
    int j = 0;
    while (j < n2) {
        const float3 pa = fma3(s2, (float)j, P0);
        bool near_unclear = false;
        if (use4) {
            const int R = tile_clearance<2>(C.t4, pa, use4);
            if (R >= 2) { j += 1 + clear_run(R, 4.0f, invm2); continue; }
            near_unclear = use4;     // a level-2 cell with R < 2 implies level-4 R < 2 as well
        }
        if (!near_unclear) {
            bool in16;
            const int R = tile_clearance<4>(C.t16, pa, in16);
            if (R >= 2) { j += 1 + clear_run(R, 16.0f, invm2); continue; }
        }
        const int kk = n1 + j;
        while (ke < kk) { pe = pe + (ke < n1 ? s1 : s2); ++ke; }
        ++exact;
        const int px = f2i(pe.x), py = f2i(pe.y), pz = f2i(pe.z);
        if (fetch_texel(V, px / 2, py / 2, pz / 2) != 0u) {   // getVolumeAt(ivec3(pos), 1)
            steps_out += kk + 1;
            const float d = d0 + step2 * (float)j;
            if (RECORD) {
                rec->d = d; rec->steps = kk + 1; rec->status = 2;
                rec->vx = px; rec->vy = py; rec->vz = pz;
                rec->pos = pe;
            }
            return d;
        }
        ++j;
    }
    steps_out += n1 + n2;
    if (RECORD) { rec->d = dist; rec->steps = n1 + n2; rec->status = 0; rec->vx = rec->vy = rec->vz = 0; rec->pos = make_float3(0.f, 0.f, 0.f); }
    return dist;
}

// Stage the CT^3-cell window of a clearance map whose origin cell is (ox, oy, oz) into `dst`.
// Cells outside the padded global array are more than `border` cells from the volume: clearance
// min(15, border + 1).
#ifdef __CUDACC__
__device__ __forceinline__ void stage_tile(uint32_t* __restrict__ dst, const ClearView& M, int ox, int oy, int oz) {
    const int ax0 = ox + M.border;                          // array x index of the tile's first cell
    const int w0 = ax0 >> 3;                                // arithmetic shift: floor for negatives
    const int sh = (ax0 & 7) * 4;
    const int words_x = (M.cx + 7) >> 3;                    // words that hold real cells (pitch has one spare)
    const uint32_t fill = 0x11111111u * (uint32_t)min(15, M.border + 1);
    for (int i = threadIdx.x; i < CT_WORDS; i += blockDim.x) {
        const int xw = i & (CTW - 1), y = (i / CTW) & (CT - 1), z = i / (CTW * CT);
        const int ay = oy + y + M.border, az = oz + z + M.border;
        uint32_t a = fill, b = fill;
        if ((unsigned)ay < (unsigned)M.cy && (unsigned)az < (unsigned)M.cz) {
            const uint32_t* row = M.words + ((size_t)az * M.cy + ay) * M.pitch;
            const int wa = w0 + xw, wb = wa + 1;
            if ((unsigned)wa < (unsigned)words_x) a = __ldg(row + wa);
            if ((unsigned)wb < (unsigned)words_x) b = __ldg(row + wb);
        }
        dst[i] = sh ? __funnelshift_r(a, b, sh) : a;
    }
}
#endif

}  // namespace vxl
