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
#include <cmath>
#include <cstring>
#include "vxl_math.cuh"
#include "vxl_trace.cuh"

namespace vxl {

#ifndef VXL_P1_UNROLL
#define VXL_P1_UNROLL 2
#endif
constexpr int P1_UNROLL = VXL_P1_UNROLL;
constexpr float BM_MARGIN = 0.125f;          // > accumulated rounding drift of <= 179 additions at |coords| < 8192
constexpr float BM_MAXCOORD = 8191.0f;

struct BitTile {
    const uint32_t* w;     // [TW][TY (z)][TY (y)] words, y innermost (the order the TMA box of the level array lands in); bit (x & 31) of word x >> 5
    int ox, oy, oz;        // tile origin, cells
    bool enabled;
    bool direct;           // texels of occupied cells can be fetched without bounds checks and with 32-bit offsets:
                           // volume dims are multiples of the cell's texel edge and the volume is < 4 GiB
    unsigned koff;         // folded constant of the direct texel offset (see TileAddr::texel_offset)
    const uint32_t* wd;    // dilated level (cell = 2^(SHIFT+1) voxels): [DW][DT][DT] words
    int dx, dy, dz;        // its origin, cells
    const uint32_t* wn;    // near tile: texel level (cell = 2 voxels), [NEAR_T][NEAR_T] words of 32 texels; nullptr = none
    int nx, ny, nz;        // its origin, texels
};

VXL_DI float3 fma3(float3 s, float k, float3 o) {
#ifdef __CUDA_ARCH__
    return make_float3(__fmaf_rn(s.x, k, o.x), __fmaf_rn(s.y, k, o.y), __fmaf_rn(s.z, k, o.z));
#else
    return make_float3(fmaf(s.x, k, o.x), fmaf(s.y, k, o.y), fmaf(s.z, k, o.z));
#endif
}

// bits of (p + m) rounded toward zero.  With m = 2^(23+S) - o*2^S (exactly representable) and
// o*2^S <= p < (o + 2^23) * 2^S the exact sum lies in [2^(23+S), 2^(24+S)), whose floats are spaced 2^S apart,
// so the result is 2^(23+S) + 2^S * (floor(p / 2^S) - o): the mantissa field holds the cell index relative to o.
VXL_DI unsigned magic_floor_bits(float p, float m, int S, int o) {
#ifdef __CUDA_ARCH__
    (void)S; (void)o;
    return __float_as_uint(__fadd_rz(p, m));
#else
    (void)m;
    return (unsigned)(0x4B000000 + (S << 23)) + (unsigned)((int)floorf(p * (1.0f / (float)(1 << S))) - o);
#endif
}

constexpr int NEAR_T = 32;                   // near tile: 32^3 texels = 64^3 voxels = 4 KB

// Lookup in the near tile: word = z * 32 + y, bit = x, all relative to the tile origin.
struct NearAddr {
    static constexpr unsigned MB = 0x4B000000u + (1u << 23);
    const uint32_t* w;
    unsigned sbase;
    float mx, my, mz;
    int ox, oy, oz;
    VXL_DI NearAddr(const BitTile& T) {
        const float M = 16777216.0f;
        w = T.wn; ox = T.nx; oy = T.ny; oz = T.nz;
        mx = M - (float)ox * 2.0f; my = M - (float)oy * 2.0f; mz = M - (float)oz * 2.0f;
#ifdef __CUDA_ARCH__
        sbase = (unsigned)__cvta_generic_to_shared(T.wn);
#else
        sbase = 0;
#endif
    }
    // the texel bit of the texel containing p, in bit 0 (p inside the near window)
    VXL_DI unsigned bit(float3 p) const {
        const unsigned bx = magic_floor_bits(p.x, mx, 1, ox), by = magic_floor_bits(p.y, my, 1, oy), bz = magic_floor_bits(p.z, mz, 1, oz);
        const unsigned idx = (bz * (unsigned)NEAR_T + by) & (unsigned)(NEAR_T * NEAR_T - 1);      // MB * 33 = 0 (mod 1024)
#ifdef __CUDA_ARCH__
        unsigned word;
        asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(sbase + 4u * idx));
        return __funnelshift_r(word, 0u, bx);
#else
        return w[idx] >> (bx & 31);
#endif
    }
};

// Tile lookup with everything constant folded once per ray.  Tile words are ordered [x word][z][y]: a row of the tile along y
// sits in consecutive banks and the next z is TY = 4 (mod 32) banks further, so the probes of neighbouring rays -- a few cells
// apart in y and z -- hit different banks.
//   b_a  = MB + rel_a                      (MB = 0x4B000000 + (SHIFT << 23), a multiple of 32)
//   word = (rel_x >> 5) * (TY*TY) + rel_z * TY + rel_y = (b_x >> 5) * (TY*TY) + b_z * TY + b_y - CC      (mod 2^32)
//   bit  = rel_x & 31 = b_x & 31
template <int SHIFT, int TY, int TW>
struct TileAddr {
    static constexpr unsigned MB = 0x4B000000u + ((unsigned)SHIFT << 23);
    static constexpr unsigned CC0 = (MB >> 5) * (unsigned)(TY * TY) + MB * (unsigned)TY + MB;
    // Bias the per-axis origins by (32 JX, KY, KZ) cells so that the folded constant vanishes mod 2^30 (word index ->
    // byte address drops two more bits): ((MB >> 5) + JX) TY TY + (MB + KZ) TY + MB + KY = 0, which saves the add of
    // the constant per lookup.  The biased index must stay inside the binade of the magic sum (32 JX + 32 TW < 2^23); a
    // tile that does not admit a solution keeps the constant.
    static constexpr unsigned RR = (0u - CC0) & 0x3FFFFFFFu;
    static constexpr bool FOLD0 = (RR / (unsigned)(TY * TY)) * 32u < (1u << 23) - 4096u;
    static constexpr int JX = FOLD0 ? (int)(RR / (unsigned)(TY * TY)) : 0;
    static constexpr int KZ = FOLD0 ? (int)((RR % (unsigned)(TY * TY)) / (unsigned)TY) : 0;
    static constexpr int KY = FOLD0 ? (int)(RR % (unsigned)TY) : 0;
    static constexpr unsigned CC = CC0 + (unsigned)JX * (unsigned)(TY * TY) + (unsigned)KZ * (unsigned)TY + (unsigned)KY;   // 4 * CC == 0 (mod 2^32) when FOLD0
    static constexpr unsigned MB1 = 0x4B000000u + (1u << 23);       // texel grid (2 voxels)
    const uint32_t* w;
    unsigned sbase;        // device: shared-window byte address of w[-CC] (wraps mod 2^32)
    float mx, my, mz;      // 2^(23+SHIFT) - o_a * 2^SHIFT
    int ox, oy, oz;        // biased origin, cells
    VXL_DI TileAddr(const BitTile& T) { init(T.w, T.ox, T.oy, T.oz); }
    VXL_DI TileAddr(const uint32_t* words, int ox_, int oy_, int oz_) { init(words, ox_, oy_, oz_); }
    VXL_DI void init(const uint32_t* words, int ox_, int oy_, int oz_) {
        const float M = (float)(1 << 23) * (float)(1 << SHIFT), cell = (float)(1 << SHIFT);
        w = words; ox = ox_ - 32 * JX; oy = oy_ - KY; oz = oz_ - KZ;
        mx = M - (float)ox * cell; my = M - (float)oy * cell; mz = M - (float)oz * cell;
#ifdef __CUDA_ARCH__
        sbase = (unsigned)__cvta_generic_to_shared(words) - 4u * CC;
        if (!FOLD0) asm volatile("" : "+r"(sbase));      // keep the folded base in one register (one LEA per lookup)
#else
        sbase = 0;
#endif
    }
    // non-zero iff the occupancy bit of the cell containing p is set (p inside the tile)
    VXL_DI unsigned test(float3 p) const {
        const unsigned bx = magic_floor_bits(p.x, mx, SHIFT, ox), by = magic_floor_bits(p.y, my, SHIFT, oy), bz = magic_floor_bits(p.z, mz, SHIFT, oz);
#ifdef __CUDA_ARCH__
        const unsigned idx = (bx >> 5) * (unsigned)(TY * TY) + bz * (unsigned)TY + by;
        unsigned word, mask;
        asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(sbase + 4u * idx));
        asm("shf.l.wrap.b32 %0, 0, 1, %1;" : "=r"(mask) : "r"(bx));       // high word of {1:0} << (bx & 31) = 1 << (bx & 31); opaque so it stays a mask test
        return word & mask;
#else
        const unsigned idx = ((bx >> 5) * (unsigned)(TY * TY) + bz * (unsigned)TY + by - CC) & 0x3FFFFFFFu;
        return w[idx] & (1u << (bx & 31));
#endif
    }
    // the occupancy bit of the cell containing p, in bit 0 (upper bits: the neighbouring cells of the word)
    VXL_DI unsigned bit(float3 p) const {
        const unsigned bx = magic_floor_bits(p.x, mx, SHIFT, ox), by = magic_floor_bits(p.y, my, SHIFT, oy), bz = magic_floor_bits(p.z, mz, SHIFT, oz);
#ifdef __CUDA_ARCH__
        const unsigned idx = (bx >> 5) * (unsigned)(TY * TY) + bz * (unsigned)TY + by;
        unsigned word;
        asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(sbase + 4u * idx));
        return __funnelshift_r(word, 0u, bx);                             // word >> (bx & 31)
#else
        const unsigned idx = ((bx >> 5) * (unsigned)(TY * TY) + bz * (unsigned)TY + by - CC) & 0x3FFFFFFFu;
        return w[idx] >> (bx & 31);
#endif
    }
    // byte offset of the texel containing p (p >= 0, inside the volume): floor(p/2) per axis, x fastest
    static VXL_DI unsigned texel_offset(const VolView& V, unsigned koff, float3 p) {
        const float M1 = 16777216.0f;                               // 2^24: grid of 2 voxels
        const unsigned tx = magic_floor_bits(p.x, M1, 1, 0), ty = magic_floor_bits(p.y, M1, 1, 0), tz = magic_floor_bits(p.z, M1, 1, 0);
        return tz * (unsigned)(V.sx * V.sy) + ty * (unsigned)V.sx + tx - koff;
    }
    static VXL_DI unsigned texel_koff(const VolView& V) { return MB1 * (unsigned)(V.sx * V.sy) + MB1 * (unsigned)V.sx + MB1; }
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

VXL_DI bool warp_any(bool p) {
#ifdef __CUDA_ARCH__
    return __any_sync(__activemask(), p);
#else
    return p;
#endif
}

// SUPER = false: raycastShadowVolumeSparse (step 0.5 then 1); true: ...SuperSparse (step 2.5 then 5).
// `fetched` counts probes that had to read the volume (diagnostics).
// LOCKSTEP: a lane whose ray has hit keeps walking with its tests masked off (live = 0) instead of leaving the
// loop, so the loop has one trip count for the whole warp (a compile-time constant when `dist` is a literal) and
// no data-dependent exit; the other lanes of the warp would have kept the issue slots busy anyway.  Every 8
// probes the warp leaves early if no lane is live.  Use it when `dist` is warp-uniform.
// COUNT: maintain `fetched` (diagnostic kernels only).
// GU: unroll factor of the per-probe loop of a group that is not clear (3 divides both group sizes; measured per kernel).
// GH > 0 (Sparse only): phase 2 runs in groups of 2*GH+1 probes; one test of the dilated level (cell = 2^(SHIFT+1)
// voxels, tile [DT][DT][DW]) at the group's middle probe clears the whole group when it reads 0, because every probe
// of the group is then within one dilated cell of the middle one (GH * max|stepDir_a| <= cell, checked per ray).
// Skipped probes only advance the float recurrence.  Pays off for rays that are coherent across a warp (sun,
// reflection, point-light shadows); a group that is not clear runs the per-probe loop.
template <bool SUPER, bool RECORD, bool LOCKSTEP, bool COUNT, int SHIFT, int TY, int TW, int DT = 1, int DW = 1, int GH = 0, int GU = 1>
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

    // A set occupancy bit drops into the reference's fetch + test right there, which is short when the texel can
    // be addressed directly.  The loops only record the hit index; everything derived from it comes after the loop.
    unsigned live = 0xFFFFFFFFu;
    float3 hpos = pos;
    // ---- phase 1 (:138-154): fine steps, position-hashed bit of the texel ----
    int hit1 = -1;
    unsigned hbit = 0u;
#pragma unroll (P1_UNROLL)
    for (int k = 0; k < n1; ++k) {
        if (A.test(pos) & live) {
            if (COUNT) ++fetched;
            unsigned v;
            if (T.direct) v = V.bytes[TileAddr<SHIFT, TY, TW>::texel_offset(V, T.koff, pos)];
            else v = fetch_texel(V, f2i(pos.x / 2.0f), f2i(pos.y / 2.0f), f2i(pos.z / 2.0f));
            if (v != 0u) {
                unsigned bit = 0u;
                bit += gmod(pos.x, 0.5f) > 0.25f ? 1u : 0u;
                bit += gmod(pos.y, 0.5f) > 0.25f ? 2u : 0u;
                bit += gmod(pos.z, 0.5f) > 0.25f ? 4u : 0u;
                if ((v >> bit) & 1u) {
                    hit1 = k; hbit = bit;
                    if (RECORD) hpos = pos;
                    if (LOCKSTEP) live = 0u; else break;
                }
            }
        }
        pos = pos + stepDir;
        if (LOCKSTEP && !SUPER && (k & 7) == 7 && !warp_any(live != 0u)) break;
    }
    if (hit1 >= 0) {
        steps_out += hit1 + 1;
        const float d = step0 * (float)(hit1 + 1);
        if (RECORD) {
            const int tx = f2i(hpos.x / 2.0f), ty = f2i(hpos.y / 2.0f), tz = f2i(hpos.z / 2.0f);
            rec->d = d; rec->steps = hit1 + 1; rec->status = 1;
            rec->vx = tx * 2 + (int)(hbit & 1u); rec->vy = ty * 2 + (int)((hbit >> 1) & 1u); rec->vz = tz * 2 + (int)((hbit >> 2) & 1u);
            rec->pos = hpos;
        }
        return d;
    }

    // ---- phase 2 (:156-170): doubled steps, byte != 0 ----
    stepDir = stepDir * 2.0f;
    const int n2 = phase2_count<SUPER>(lim);
    int hit2 = -1;
    int j = 0;
    // one per-probe test of phase 2; returns true when the ray ends at probe jj
    auto probe2 = [&](int jj) -> bool {
        if (A.test(pos) & live) {
            if (COUNT) ++fetched;
            unsigned v;
            if (T.direct) v = V.bytes[TileAddr<SHIFT, TY, TW>::texel_offset(V, T.koff, pos)];
            else { const int px = f2i(pos.x), py = f2i(pos.y), pz = f2i(pos.z); v = fetch_texel(V, px / 2, py / 2, pz / 2); }   // getVolumeAt(ivec3(pos), 1)
            if (v != 0u) {
                hit2 = jj;
                if (RECORD) hpos = pos;
                return true;
            }
        }
        return false;
    };
    if (!SUPER && GH > 0) {
        constexpr int G = 2 * GH + 1;
        const TileAddr<SHIFT + 1, DT, DW> D(T.wd, T.dx, T.dy, T.dz);
        const float m2 = fmaxf(fmaxf(fabsf(stepDir.x), fabsf(stepDir.y)), fabsf(stepDir.z));
        const bool can_group = (float)GH * m2 <= (float)(2 << SHIFT) - 2.0f * BM_MARGIN;
        bool done = false;
        while (!done && j + G <= n2) {
            const float3 p0 = pos;
#pragma unroll
            for (int i = 0; i < GH; ++i) pos = pos + stepDir;            // middle probe of the group
            if (can_group && !D.test(pos)) {
#pragma unroll
                for (int i = 0; i < GH + 1; ++i) pos = pos + stepDir;    // first probe of the next group
                j += G;
                continue;
            }
            pos = p0;
#pragma unroll (GU)
            for (int i = 0; i < G; ++i) {
                if (probe2(j + i)) { done = true; break; }
                pos = pos + stepDir;
            }
            j += G;
        }
        if (done) j = n2;
    }
#pragma unroll 2
    for (; j < n2; ++j) {
        if (probe2(j)) { if (LOCKSTEP) live = 0u; else break; }
        pos = pos + stepDir;
        if (LOCKSTEP && (j & 7) == 7 && !warp_any(live != 0u)) break;
    }
    // single exit: everything below is arithmetic on the recorded hit index
    const bool h2 = hit2 >= 0;
    const int nsteps = h2 ? n1 + hit2 + 1 : n1 + n2;
    const float d = h2 ? d0 + step2 * (float)hit2 : dist;
    steps_out += nsteps;
    if (RECORD) {
        rec->d = d; rec->steps = nsteps; rec->status = h2 ? 2 : 0;
        rec->vx = h2 ? f2i(hpos.x) : 0; rec->vy = h2 ? f2i(hpos.y) : 0; rec->vz = h2 ? f2i(hpos.z) : 0;
        rec->pos = h2 ? hpos : make_float3(0.f, 0.f, 0.f);
    }
    return d;
}

// ---- scan + resolve march (SuperSparse with a compile-time probe count: the AO rays) --------------------------
// The per-probe loop above leaves the warp whenever ONE lane meets a set occupancy bit (8 % of the probes, i.e. 80 %
// of the warp iterations at 19 live lanes), and every such excursion is a dependent L2 round trip.  Here a ray first
// SCANS all N = 6 + N2 probe positions with the reference's float recurrence, branch-free and fully unrolled, packing
// the occupancy bits of its probes into one word (13 instructions per probe: 3 FADD, 3 FADD.RZ, 4 address, LDS, 2
// funnel shifts).  Then it RESOLVES the set bits in probe order with the reference's own texel test, stopping at the
// first hit.  The probes after a hit are wasted scan work, bounded by N.
//
// Resolve needs the probe position.  Phase-1 candidates (k < 6) replay the recurrence (<= 5 additions).  A phase-2
// candidate only needs its TEXEL, floor(pos / 2): q = fma(stepDir, 2k - 6, origin) differs from the recurrence's pos_k
// by at most (k + 2) roundings of values below hi + 1 (<= 31 * 2^-24 * (hi + 1) < eps = (hi + 1) * 2^-18, origin + (2k - 6) stepDir
// being what the recurrence sums without rounding: stepDir * 2 is exact); when q - eps and q + eps fall in the same
// texel on every axis, so does pos_k.  Otherwise (about 1 % of candidates) the recurrence is replayed.
template <int SHIFT, int TY, int TW>
VXL_DI bool tile_eligible(const BitTile& T, float3 origin, float3 dir, float reach, float& hi_max) {
    const float3 end = fma3(dir, reach, origin);
    const float3 lo = make_float3(fminf(origin.x, end.x), fminf(origin.y, end.y), fminf(origin.z, end.z));
    const float3 hi = make_float3(fmaxf(origin.x, end.x), fmaxf(origin.y, end.y), fmaxf(origin.z, end.z));
    const float cell = (float)(1 << SHIFT);
    const float3 tlo = make_float3((float)T.ox * cell, (float)T.oy * cell, (float)T.oz * cell);
    bool fast = T.enabled;
    fast = fast && (lo.x >= fmaxf(tlo.x, 0.0f) + BM_MARGIN) && (lo.y >= fmaxf(tlo.y, 0.0f) + BM_MARGIN) && (lo.z >= fmaxf(tlo.z, 0.0f) + BM_MARGIN);
    fast = fast && (hi.x <= fminf(tlo.x + (float)(TW * 32) * cell, BM_MAXCOORD) - BM_MARGIN) &&
           (hi.y <= fminf(tlo.y + (float)TY * cell, BM_MAXCOORD) - BM_MARGIN) && (hi.z <= fminf(tlo.z + (float)TY * cell, BM_MAXCOORD) - BM_MARGIN);
    fast = fast && (origin.x == origin.x) && (origin.y == origin.y) && (origin.z == origin.z) && (end.x == end.x) && (end.y == end.y) && (end.z == end.z);
    hi_max = fmaxf(fmaxf(hi.x, hi.y), hi.z);
    return fast;
}

VXL_DI unsigned funnel_r(unsigned lo, unsigned hi, unsigned sh) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}

// N2 = phase2_count<true>(min(dist, 164)), known at compile time (23 for the AO rays' dist = 128); 6 + N2 <= 32.
// NEAR: the block also staged the near tile (texel bits of the 64^3 voxels around the ray origins).  A ray whose first
// KN = 8 probes (d <= 25) stay inside it tests those probes against the texel bits instead of the 4-voxel cells: a
// phase-2 probe is then answered exactly (the bit IS getVolumeAt(pos, 1)), and a phase-1 probe is a candidate only when
// its texel is non-zero (a quarter of the cell-level candidates, measured on config 3).
// ScanPre: what the caller already proved for every ray of a bundle that shares `origin` and whose directions obey
// |dir_a| <= bound_a (the 16 AO rays of a pixel): the box origin +- reach * bound lies inside the tile / the near tile.
struct ScanPre { bool ok, near_ok; float hi_max; };

// blo / bhi: every direction of the bundle obeys -blo_a <= dir_a <= bhi_a (all >= 0)
// low: the lowest coordinate the caller's lookups and texel tests can handle: 0 (there floor == the reference's truncation), or
// -BM_LOWCOORD for a caller whose candidate test reproduces the truncation at negative coordinates (test_super_cand); the occupancy
// levels repeat texel 0 at texel -1 for exactly this (vxl_occupancy.cu) and are empty further out.  The near tile's bits are
// taken as the reference's answer, so it is only used at non-negative coordinates.
constexpr float BM_LOWCOORD = 150.0f;
template <int SHIFT, int TY, int TW>
VXL_DI ScanPre scan_precheck(const BitTile& T, float3 origin, float3 blo, float3 bhi, float reach, float near_reach, float low = 0.0f) {
    ScanPre P;
    const float cell = (float)(1 << SHIFT);
    const float3 lo = origin - blo * (reach * 1.00002f), hi = origin + bhi * (reach * 1.00002f);
    const float3 tlo = make_float3((float)T.ox * cell, (float)T.oy * cell, (float)T.oz * cell);
    bool ok = T.enabled && T.direct;
    ok = ok && (lo.x >= fmaxf(tlo.x, low) + BM_MARGIN) && (lo.y >= fmaxf(tlo.y, low) + BM_MARGIN) && (lo.z >= fmaxf(tlo.z, low) + BM_MARGIN);
    ok = ok && (hi.x <= fminf(tlo.x + (float)(TW * 32) * cell, BM_MAXCOORD) - BM_MARGIN) &&
         (hi.y <= fminf(tlo.y + (float)TY * cell, BM_MAXCOORD) - BM_MARGIN) && (hi.z <= fminf(tlo.z + (float)TY * cell, BM_MAXCOORD) - BM_MARGIN);
    P.ok = ok;                                                             // NaN anywhere makes a comparison fail
    P.hi_max = fmaxf(fmaxf(hi.x, hi.y), hi.z);
    const float3 nl = blo * (near_reach * 1.00002f), nh = bhi * (near_reach * 1.00002f);
    const float lx = (float)T.nx * 2.0f + BM_MARGIN, ly = (float)T.ny * 2.0f + BM_MARGIN, lz = (float)T.nz * 2.0f + BM_MARGIN;
    const float w = (float)(2 * NEAR_T) - 2.0f * BM_MARGIN;
    P.near_ok = ok && T.wn != nullptr && origin.x - nl.x >= fmaxf(lx, BM_MARGIN) && origin.x + nh.x <= lx + w && origin.y - nl.y >= fmaxf(ly, BM_MARGIN) &&
                origin.y + nh.y <= ly + w && origin.z - nl.z >= fmaxf(lz, BM_MARGIN) && origin.z + nh.z <= lz + w;
    return P;
}
template <int SHIFT, int TY, int TW>
VXL_DI ScanPre scan_precheck(const BitTile& T, float3 origin, float3 bound, float reach, float near_reach) {
    return scan_precheck<SHIFT, TY, TW>(T, origin, bound, bound, reach, near_reach);
}

// The scan of march_scan_super on its own: candidate word of an eligible ray (probe k in bit k).
template <bool NEAR, int SHIFT, int TY, int TW, int N2>
VXL_DI unsigned scan_super_cand(const BitTile& T, float3 origin, float3 dir, bool near_ok) {
    constexpr int N1 = 6, N = N1 + N2, KN = 8;
    static_assert(N <= 32 && KN <= N, "candidate mask is one word");
    typedef TileAddr<SHIFT, TY, TW> TA;
    const TA A(T);
    const float3 s1 = dir * 2.5f;
    const float3 s2 = s1 * 2.0f;
    unsigned cand = 0u;
    float3 pos = origin;
    if (NEAR && near_ok) {
        const NearAddr B(T);
#pragma unroll
        for (int k = 0; k < KN; ++k) {
            cand = funnel_r(cand, B.bit(pos), 1u);
            pos = pos + (k + 1 <= N1 ? s1 : s2);
        }
    } else {
#pragma unroll
        for (int k = 0; k < KN; ++k) {
            cand = funnel_r(cand, A.bit(pos), 1u);
            pos = pos + (k + 1 <= N1 ? s1 : s2);                            // the addition after probe 5 still uses s1 (:151)
        }
    }
#pragma unroll
    for (int k = KN; k < N; ++k) {
        cand = funnel_r(cand, A.bit(pos), 1u);
        if (k + 1 < N) pos = pos + s2;
    }
    return cand >> (32 - N);
}

// ---- two rays per lane, packed (sm_100a add.f32x2) ------------------------------------------------------------------------
// The scan is bound by issue slots (13 per probe: 3 FADD recurrence + 3 FADD.RZ cell index + 4 address + LDS + 2 funnel shifts), not
// by FP throughput.  Blackwell's FADD2 adds two independent binary32 pairs in one issue slot with the same IEEE rounding per half
// (rn for the recurrence, rz for the magic floor), so two rays that share an origin (two AO rays of a pixel) scan side by side with
// 6 packed additions per probe PAIR instead of 12: 10 issue slots per probe.  Same values, bit for bit.
struct F2 {
#ifdef __CUDA_ARCH__
    unsigned long long v;
#else
    float lo, hi;
#endif
};
VXL_DI F2 f2_pack(float lo, float hi) {
    F2 r;
#ifdef __CUDA_ARCH__
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
#else
    r.lo = lo; r.hi = hi;
#endif
    return r;
}
VXL_DI F2 f2_add(F2 a, F2 b) {
    F2 r;
#ifdef __CUDA_ARCH__
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
#else
    r.lo = a.lo + b.lo; r.hi = a.hi + b.hi;
#endif
    return r;
}
#ifndef __CUDA_ARCH__
// host emulation of add.rz.f32: the exact sum in double (53 bits hold any sum of two binary32 of the magnitudes used here), truncated
VXL_DI unsigned host_add_rz_bits(float a, float b) {
    const double e = (double)a + (double)b;
    float r = (float)e;                                   // round to nearest
    if (fabs((double)r) > fabs(e)) r = nextafterf(r, 0.0f);
    unsigned u;
    memcpy(&u, &r, 4);
    return u;
}
#endif
// both halves of (p + m) rounded toward zero, as raw bits (magic_floor_bits on two values)
VXL_DI void f2_floor_bits(F2 p, F2 m, int S, int o, unsigned& lo, unsigned& hi) {
#ifdef __CUDA_ARCH__
    (void)S; (void)o;
    unsigned long long r;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p.v), "l"(m.v));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(r));
#else
    (void)S; (void)o;
    lo = host_add_rz_bits(p.lo, m.lo); hi = host_add_rz_bits(p.hi, m.hi);
#endif
}

// Where the probes of one stretch of a scan look: a bit array (32 cells per word along x, y innermost) described by run-time values,
// so that ONE unrolled scan serves the near tile, the block's main tile and the level array in global memory.
//   b_a  = MB + rel_a  (float bits of p_a + m_a rounded toward zero; MB = 0x4B000000 + (S << 23), a multiple of 32)
//   word = (rel_x >> 5) * sy + rel_z * sz + rel_y = (b_x >> 5) * sy + b_z * sz + b_y - CC,  CC = (MB >> 5) * sy + MB * (sz + 1)   (mod 2^32)
// SY / SZ > 0: the strides are compile-time constants (immediate operands); 0: taken from sy / sz.
template <bool GLOBAL, unsigned SY = 0, unsigned SZ = 0>
struct ScanLook {
    float mx, my, mz;          // 2^(23+S) - o_a * 2^S
    unsigned sy, sz;           // words from one x word to the next / from one z to the next
    unsigned sbase;            // shared memory: byte address of word 0 minus 4 * CC (wraps); global memory: CC
    const uint32_t* w;         // word 0 (global memory; host emulation)
    VXL_DI static ScanLook make(const uint32_t* words, int S, int ox, int oy, int oz, unsigned sy_, unsigned sz_) {
        ScanLook L;
        const float M = (float)(1 << 23) * (float)(1 << S), cell = (float)(1 << S);
        L.mx = M - (float)ox * cell; L.my = M - (float)oy * cell; L.mz = M - (float)oz * cell;
        L.sy = SY ? SY : sy_; L.sz = SZ ? SZ : sz_; L.w = words;
        const unsigned MB = 0x4B000000u + ((unsigned)S << 23);
        const unsigned CC = MB * (L.sz + 1u) + (MB >> 5) * L.sy;
#ifdef __CUDA_ARCH__
        L.sbase = GLOBAL ? CC : (unsigned)__cvta_generic_to_shared(words) - 4u * CC;
#else
        L.sbase = CC;
#endif
        return L;
    }
    VXL_DI unsigned word(unsigned bx, unsigned by, unsigned bz) const {
        const unsigned idx = (bx >> 5) * (SY ? SY : sy) + bz * (SZ ? SZ : sz) + by;
#ifdef __CUDA_ARCH__
        if (GLOBAL) return __ldg(w + (idx - sbase));
        unsigned v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + 4u * idx));
        return v;
#else
        return w[idx - sbase];
#endif
    }
};

// scan_super_cand for two rays from the same origin (two AO rays of a pixel): probe k of ray a / b ends up in bit k of cand_a /
// cand_b.  The first KN = 8 probes look at `first` (the near tile when the pixel's rays stay inside it, else the same as `rest`).
template <int N2, typename LookA, typename LookB>
VXL_DI void scan_super_pair(const LookA& first, const LookB& rest, float3 origin, float3 dir_a, float3 dir_b,
                            unsigned& cand_a, unsigned& cand_b) {
    constexpr int N1 = 6, N = N1 + N2, KN = 8;
    static_assert(N <= 32 && KN <= N, "candidate mask is one word");
    const float3 s1a = dir_a * 2.5f, s1b = dir_b * 2.5f;
    const float3 s2a = s1a * 2.0f, s2b = s1b * 2.0f;
    const F2 s1x = f2_pack(s1a.x, s1b.x), s1y = f2_pack(s1a.y, s1b.y), s1z = f2_pack(s1a.z, s1b.z);
    const F2 s2x = f2_pack(s2a.x, s2b.x), s2y = f2_pack(s2a.y, s2b.y), s2z = f2_pack(s2a.z, s2b.z);
    F2 px = f2_pack(origin.x, origin.x), py = f2_pack(origin.y, origin.y), pz = f2_pack(origin.z, origin.z);
    unsigned ca = 0u, cb = 0u;
    auto look = [&](const auto& L) {
        unsigned xa, xb, ya, yb, za, zb;
        f2_floor_bits(px, f2_pack(L.mx, L.mx), 0, 0, xa, xb);
        f2_floor_bits(py, f2_pack(L.my, L.my), 0, 0, ya, yb);
        f2_floor_bits(pz, f2_pack(L.mz, L.mz), 0, 0, za, zb);
        ca = funnel_r(ca, funnel_r(L.word(xa, ya, za), 0u, xa), 1u);
        cb = funnel_r(cb, funnel_r(L.word(xb, yb, zb), 0u, xb), 1u);
    };
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (k < KN) look(first); else look(rest);
        if (k + 1 <= N1) { px = f2_add(px, s1x); py = f2_add(py, s1y); pz = f2_add(pz, s1z); }      // the addition after probe 5 still uses s1 (:151)
        else if (k + 1 < N) { px = f2_add(px, s2x); py = f2_add(py, s2y); pz = f2_add(pz, s2z); }
    }
    cand_a = ca >> (32 - N);
    cand_b = cb >> (32 - N);
}

// The reference's own test of candidate probe k (see above): true = the march returns at this probe.
// PHASE: 0 = k decides; 1 / 2 = the caller knows k < 6 / k >= 6 (only that test is compiled in).
template <bool COUNT, bool NEAR, int SHIFT, int TY, int TW, int N2, int PHASE = 0>
VXL_DI bool resolve_super_cand(const VolView& V, const BitTile& T, float3 origin, float3 s1, int k, bool near_ok, float eps, unsigned& fetched, unsigned& hbit) {
    constexpr int N1 = 6, KN = 8;
    typedef TileAddr<SHIFT, TY, TW> TA;
    if (PHASE == 1 || (PHASE == 0 && k < N1)) {
        if (COUNT) ++fetched;
        float3 p = origin;
#pragma unroll
        for (int i = 0; i < N1 - 1; ++i)
            if (i < k) p = p + s1;
        const unsigned v = V.bytes[TA::texel_offset(V, T.koff, p)];
        if (v != 0u) {
            unsigned bit = 0u;
            bit += gmod(p.x, 0.5f) > 0.25f ? 1u : 0u;
            bit += gmod(p.y, 0.5f) > 0.25f ? 2u : 0u;
            bit += gmod(p.z, 0.5f) > 0.25f ? 4u : 0u;
            if ((v >> bit) & 1u) { hbit = bit; return true; }
        }
        return false;
    }
    if (NEAR && near_ok && k < KN) return true;                           // texel bit set = byte != 0 = the reference's test
    if (COUNT) ++fetched;
    const float3 s2 = s1 * 2.0f;
    const float3 q = fma3(s1, (float)(2 * k - N1), origin);
    const float M1 = 16777216.0f;
    const unsigned ax = magic_floor_bits(q.x - eps, M1, 1, 0), bx = magic_floor_bits(q.x + eps, M1, 1, 0);
    const unsigned ay = magic_floor_bits(q.y - eps, M1, 1, 0), by = magic_floor_bits(q.y + eps, M1, 1, 0);
    const unsigned az = magic_floor_bits(q.z - eps, M1, 1, 0), bz = magic_floor_bits(q.z + eps, M1, 1, 0);
    unsigned off = bz * (unsigned)(V.sx * V.sy) + by * (unsigned)V.sx + bx - T.koff;
    if ((ax ^ bx) | (ay ^ by) | (az ^ bz)) {                             // within eps of a texel face: exact position
        float3 p = origin;
        for (int i = 0; i < N1; ++i) p = p + s1;
        for (int i = N1; i < k; ++i) p = p + s2;
        off = TA::texel_offset(V, T.koff, p);
    }
    return V.bytes[off] != 0u;
}

// Light.frag:142-146 for one coordinate 0 <= x < 2^21:  mod(x, 0.5) > 0.25.  The shader's float evaluation is exact (2x, floor, the
// product with 0.5 and the final difference are all representable), so the test is frac(2x) > 0.5, i.e. m = floor(4x) is odd and
// 4x != m.  x + 2^21 rounded toward zero has m in its low mantissa bits, and the addition was inexact iff 4x was not an integer.
VXL_DI unsigned hash_bit1(float x) {
#ifdef __CUDA_ARCH__
    const float r = __fadd_rz(x, 2097152.0f);
    return (__float_as_uint(r) & 1u) & (unsigned)(__fadd_rn(r, -2097152.0f) != x);
#else
    return gmod(x, 0.5f) > 0.25f ? 1u : 0u;
#endif
}

// The reference's test of candidate probe k of a scanned SuperSparse ray, as one converged piece of code for a warp that holds a mix
// of phase-1 (k < 6) and phase-2 candidates (ao_pooled's resolve passes).  Both phases fetch ONE texel byte; they differ in the position
// it is fetched at -- phase 1 replays the recurrence (<= 5 predicated additions: its bit hash needs the exact position), phase 2
// uses q = fma(s1, 2k - 6, origin) with the eps argument of resolve_super_cand -- and in the final test of the byte.
// near_ok: the ray's first 8 probes were scanned against the near tile, where a set bit of probe 6 / 7 IS the reference's test.
template <bool COUNT>
VXL_DI bool test_super_cand(const VolView& V, unsigned koff, float3 origin, float3 s1, int k, bool near_ok, float eps, unsigned& fetched) {
    constexpr int N1 = 6, KN = 8;
    if (near_ok && k >= N1 && k < KN) return true;
    const bool ph1 = k < N1;
    float3 p = origin;
#pragma unroll
    for (int i = 0; i < N1 - 1; ++i)
        if (i < k) p = p + s1;                                         // exact for k < 6; unused otherwise
    const float3 q = fma3(s1, (float)(2 * k - N1), origin);
    const float3 pos = ph1 ? p : q;
    const float e = ph1 ? 0.0f : eps;
    const float M1 = 16777216.0f;
    const unsigned ax = magic_floor_bits(pos.x - e, M1, 1, 0), bx = magic_floor_bits(pos.x + e, M1, 1, 0);
    const unsigned ay = magic_floor_bits(pos.y - e, M1, 1, 0), by = magic_floor_bits(pos.y + e, M1, 1, 0);
    const unsigned az = magic_floor_bits(pos.z - e, M1, 1, 0), bz = magic_floor_bits(pos.z + e, M1, 1, 0);
    const unsigned off = bz * (unsigned)(V.sx * V.sy) + by * (unsigned)V.sx + bx - koff;
    if (COUNT) ++fetched;
    unsigned v;
    const bool low = fminf(fminf(pos.x, pos.y), pos.z) - e < 0.0f;    // a coordinate below zero: the reference truncates toward zero there
    if (((ax ^ bx) | (ay ^ by) | (az ^ bz)) != 0u || low) {            // phase 2 within eps of a texel face (about 1 %), or a negative coordinate
        bool decided = false;
        v = 0u;
        if (!low && !ph1 && (int)(ax != bx) + (int)(ay != by) + (int)(az != bz) == 1) {
            // pos_k lies within eps of q on every axis, so its texel is one of the two that share the face (a: from q - eps, b: from
            // q + eps); the phase-2 test only asks whether the byte is zero, and neighbouring texels mostly agree on that
            const unsigned MB1 = 0x4B000000u + (1u << 23);
            const unsigned va = fetch_texel(V, (int)(ax - MB1), (int)(ay - MB1), (int)(az - MB1));
            const unsigned vb = fetch_texel(V, (int)(bx - MB1), (int)(by - MB1), (int)(bz - MB1));
            decided = (va != 0u) == (vb != 0u);
            v = vb;
        }
        if (!decided) {                                                  // the exact position: replay the recurrence
            const float3 s2 = s1 * 2.0f;
            float3 r = origin;
#pragma unroll 1
            for (int i = 0; i < k; ++i) r = r + (i < N1 ? s1 : s2);
            // Light.frag:140 ivec3(pos / 2) in phase 1, :163 ivec3(pos) / 2 in phase 2; out of range reads 0
            if (ph1) v = fetch_texel(V, f2i(r.x / 2.0f), f2i(r.y / 2.0f), f2i(r.z / 2.0f));
            else v = fetch_texel(V, f2i(r.x) / 2, f2i(r.y) / 2, f2i(r.z) / 2);
        }
    } else v = ldg(V.bytes + off);
    if (v == 0u) return false;
    if (!ph1) return true;
    unsigned bit;
    if (low) bit = (gmod(pos.x, 0.5f) > 0.25f ? 1u : 0u) | (gmod(pos.y, 0.5f) > 0.25f ? 2u : 0u) | (gmod(pos.z, 0.5f) > 0.25f ? 4u : 0u);    // :142-146
    else bit = hash_bit1(pos.x) | (hash_bit1(pos.y) << 1) | (hash_bit1(pos.z) << 2);
    return ((v >> bit) & 1u) != 0u;
}

VXL_DI float super_hit_distance(int hit) { return hit < 6 ? 2.5f * (float)(hit + 1) : 17.5f + 5.0f * (float)(hit - 6); }

template <bool RECORD, bool COUNT, bool NEAR, int SHIFT, int TY, int TW, int N2>
VXL_DI float march_scan_super(const VolView& V, const BitTile& T, float3 origin, float3 dir, float dist, int& steps_out,
                              MarchResult* rec, unsigned& fetched, ScanPre pre = ScanPre{false, false, 0.0f}) {
    constexpr int N1 = 6, N = N1 + N2, KN = 8;
    const float lim = fminf(dist, 164.0f);
    float hi_max = pre.hi_max;
    if (!pre.ok && (!T.direct || !tile_eligible<SHIFT, TY, TW>(T, origin, dir, fmaxf(lim, 16.0f) + 1.0f, hi_max)))
        return march<RECORD>(V, origin, dir, dist, 2.5f, steps_out, rec);

    const float3 s1 = dir * 2.5f;
    const float3 s2 = s1 * 2.0f;
    bool near_ok = pre.near_ok;
    if (NEAR && !pre.ok) {
        const float3 e = fma3(s1, (float)(2 * (KN - 1) - N1), origin);       // probe KN - 1, up to rounding (<< margin)
        const float lx = (float)T.nx * 2.0f + BM_MARGIN, ly = (float)T.ny * 2.0f + BM_MARGIN, lz = (float)T.nz * 2.0f + BM_MARGIN;
        const float w = (float)(2 * NEAR_T) - 2.0f * BM_MARGIN;
        near_ok = T.wn != nullptr && fminf(origin.x, e.x) >= lx && fmaxf(origin.x, e.x) <= lx + w && fminf(origin.y, e.y) >= ly &&
                  fmaxf(origin.y, e.y) <= ly + w && fminf(origin.z, e.z) >= lz && fmaxf(origin.z, e.z) <= lz + w;
    }
    // ---- scan: probe k ends up in bit k ----
    unsigned cand = scan_super_cand<NEAR, SHIFT, TY, TW, N2>(T, origin, dir, near_ok);
    // ---- resolve ----
    const float eps = (hi_max + 1.0f) * (1.0f / 262144.0f);
    int hit = -1;
    unsigned hbit = 0u;
    while (cand) {
#ifdef __CUDA_ARCH__
        const int k = __ffs((int)cand) - 1;
#else
        const int k = __builtin_ctz(cand);
#endif
        cand &= cand - 1u;
        if (resolve_super_cand<COUNT, NEAR, SHIFT, TY, TW, N2>(V, T, origin, s1, k, near_ok, eps, fetched, hbit)) { hit = k; break; }
    }
    const bool h = hit >= 0;
    const int nsteps = h ? hit + 1 : N;
    const float d = !h ? dist : super_hit_distance(hit);
    steps_out += nsteps;
    if (RECORD) {
        float3 p = origin;
        for (int i = 0; i < hit; ++i) p = p + (i < N1 ? s1 : s2);
        rec->d = d; rec->steps = nsteps; rec->status = !h ? 0 : (hit < N1 ? 1 : 2);
        if (h && hit < N1) {
            const int tx = f2i(p.x / 2.0f), ty = f2i(p.y / 2.0f), tz = f2i(p.z / 2.0f);
            rec->vx = tx * 2 + (int)(hbit & 1u); rec->vy = ty * 2 + (int)((hbit >> 1) & 1u); rec->vz = tz * 2 + (int)((hbit >> 2) & 1u);
        } else {
            rec->vx = h ? f2i(p.x) : 0; rec->vy = h ? f2i(p.y) : 0; rec->vz = h ? f2i(p.z) : 0;
        }
        rec->pos = h ? p : make_float3(0.f, 0.f, 0.f);
    }
    return d;
}

// Stage the TW*32 x TY x TY-cell window of an occupancy level whose origin cell is (ox, oy, oz), words ordered [x word][z][y].  The window may start at any bit of the level's words (funnel shift).  Everything outside the level's
// array is empty.  Consecutive threads take consecutive y: coalesced reads, conflict-free writes.
#ifdef __CUDACC__
template <int TY, int TW>
__device__ __forceinline__ void stage_bits(uint32_t* __restrict__ dst, const BitView& M, int ox, int oy, int oz) {
    const int ax0 = ox + M.border;                          // array index = cell index + border
    const int w0 = ax0 >> 5;                                // arithmetic shift: floor for negatives
    const int sh = ax0 & 31;
#pragma unroll 2
    for (int r = threadIdx.x; r < TY * TY; r += blockDim.x) {   // one tile row (TW words) per iteration
        const int z = r / TY, y = r - z * TY;
        const int ay = oy + y + M.border, az = oz + z + M.border;
        uint32_t g[TW + 1];
#pragma unroll
        for (int k = 0; k <= TW; ++k) g[k] = 0u;
        if ((unsigned)ay < (unsigned)M.cy && (unsigned)az < (unsigned)M.cz) {
            const uint32_t* row = M.words + (size_t)az * M.xw * M.cyp + ay;
#pragma unroll
            for (int k = 0; k <= TW; ++k)
                if ((unsigned)(w0 + k) < (unsigned)M.xw) g[k] = __ldg(row + (size_t)(w0 + k) * M.cyp);   // the last word of each row is zero
        }
#pragma unroll
        for (int k = 0; k < TW; ++k) dst[(k * TY + z) * TY + y] = __funnelshift_r(g[k], g[k + 1], sh);   // sh == 0 returns g[k]
    }
}
#endif

}  // namespace vxl
