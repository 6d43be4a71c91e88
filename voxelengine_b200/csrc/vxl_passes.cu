// vxl_passes.cu -- the four light passes and the ray-level entry as sm_100a kernels.
//
// A thread block covers a region of a tile-compact frame shard (32x16 pixels for the ambient pass, 64x32 for the others); its 16 warps
// draw 8x4-pixel work items from it.
// Ray generation follows the reference fragment shaders line by line (citations inline).  The
// rays of a block start within a few voxels of each other, so the block stages the occupancy-bit
// tile around them in shared memory once and every probe tests that tile before touching the
// volume (vxl_bitmarch.cuh); kernel variant 0 is the plain march on the bytes (vxl_trace.cuh).
#include "vxl_internal.h"
#include "vxl_math.cuh"
#include "vxl_trace.cuh"
#include "vxl_bitmarch.cuh"
#include "vxl_pixel.cuh"
#include "vxl_tma.cuh"

#include <cstddef>

namespace vxl {

// Shared-memory state of one thread block (dynamic: the tile exceeds the 48 KB static limit).
template <typename G>
struct BlockShared {
    float lut[LUT_FLOATS];
    int bb[12];                               // boxes of the block's rays: [0..5] whole rays, [6..11] their first stretch (near tile)
    unsigned acc[4];
    uint64_t bar;                             // mbarrier the TMA copy of `tile` completes
    int next_item;                            // work items of the region handed out so far
    unsigned zero;                            // 0, read back through a volatile load: an addend the compiler cannot see through (ao_pooled)
    alignas(128) uint32_t tile[G::TY * G::TY * G::TW];     // from here on: kernel variant 0 allocates only the header
    uint32_t dtile[G::DW * G::DT * G::DT];
    uint32_t ntile[G::NEAR ? NEAR_T * NEAR_T : 1];
    // pooled AO resolve (ao_pooled): per warp a stack of pending candidate tests and the per-pixel AO sums
    uint4 q_ent[G::QCAP ? BLOCK_THREADS / 32 : 1][G::QCAP ? G::QCAP : 1];       // (s1.xyz, remaining candidate bits)
    uint8_t q_own[G::QCAP ? BLOCK_THREADS / 32 : 1][G::QCAP ? G::QCAP : 4];     // owner lane | near_ok << 5
    unsigned ao_acc[G::QCAP ? BLOCK_THREADS : 1];                               // sum of (256 d / 128)^2 over the rays resolved from the stack
};
template <typename G, bool FAST>
constexpr size_t smem_bytes() {
    typedef BlockShared<G> BS;
    return FAST ? sizeof(BS) : offsetof(BS, tile);
}
// Tile geometry per pass.  Plain tile: cells of 2^SHIFT voxels, TW*32 x TY x TY cells.  Dilated tile: cells of
// 2^(SHIFT+1) voxels, DW*32 x DT x DT cells, covering at least the plain tile.  GH: half width of a probe group of the
// Sparse march (GH * max|stepDir_a| must stay <= the dilated cell: 7 * 1.0 <= 8, 10 * 1.5 <= 16).
// GU: unroll factor of the walk through a probe group the dilated level could not clear (reflection 1.23 -> 1.14 ms, point 0.36 -> 0.35 ms
// with 3; the ambient kernel, whose sun ray is a fraction of its code, gets slower: 4.30 -> 4.37 ms).
// NEAR: also stage the near tile (texel bits of the 64^3 voxels around the ray origins, 4 KB; vxl_bitmarch.cuh).
// The plain tile is one TMA box of the level array (vxl_occupancy.cu; it lands ordered [x word][z][y]): TY cells along y starting
// on a multiple of 4 cells, TY slices, and TW = 3 words (96 cells) along x starting on a word boundary of the array or of its copy
// shifted by 16 cells -- so the window reaches at least 40 cells (160 voxels) either side of the block's centre along x.
// Ambient: 67.7 KB + 11.9 KB + 4 KB near tile + 4 KB LUTs + 19 KB pooled-resolve state = 107 KB, two 512-thread blocks per SM.
// Local lights / reflection: three blocks per SM with TY = 68 (54.2 + 9.6 + 4 KB), or two with TY = 80 (VXL_PASS_BLOCKS = 2).
#ifndef VXL_AO_PROMOTE
#define VXL_AO_PROMOTE 0
#endif
#ifndef VXL_AO_QCAP
#define VXL_AO_QCAP 64            // pending candidate tests per warp (>= 63: 31 left over + 32 new); 0 = per-lane resolve (round-1 kernel)
#endif
#ifndef VXL_AMB_STATIC
#define VXL_AMB_STATIC 1
#endif
// The one-ray passes place their tile from a SAMPLE of the region's pixels: every VXL_PREPASS_STRIDE-th work item (the boxes carry
// +-20 voxels of slack and every ray still checks its own eligibility, so a pixel the sample missed at worst takes the plain march).
#ifndef VXL_PREPASS_STRIDE
#define VXL_PREPASS_STRIDE 4
#endif
#ifndef VXL_AO_POOL_MIN
#define VXL_AO_POOL_MIN 2            // fewer AO rays per pixel than this: per-lane scan + resolve (the pool needs rays to fill its passes)
#endif
#ifndef VXL_AMB_GH
#define VXL_AMB_GH 7
#endif
#ifndef VXL_LOCAL_GH
#define VXL_LOCAL_GH 7
#endif
#ifndef VXL_REFL_GH
#define VXL_REFL_GH 10
#endif
#ifndef VXL_LOCAL_GU
#define VXL_LOCAL_GU 3
#endif
#ifndef VXL_REFL_GU
#define VXL_REFL_GU 3
#endif
#ifndef VXL_PASS_TY
#define VXL_PASS_TY (VXL_PASS_BLOCKS >= 3 ? 68 : 80)
#endif
struct AmbientGeom { static constexpr int SHIFT = 2, TY = 76, TW = 3, DT = 39, DW = 2, GH = VXL_AMB_GH, GU = 1, QCAP = VXL_AO_QCAP; static constexpr bool NEAR = true; };     // +-152 voxels in y and z (AO 128, sun 128)
struct LocalGeom   { static constexpr int SHIFT = 2, TY = VXL_PASS_TY, TW = 3, DT = VXL_PASS_TY / 2 + 1, DW = 2, GH = VXL_LOCAL_GH, GU = VXL_LOCAL_GU, QCAP = 0; static constexpr bool NEAR = false; };    // point/spot rays that leave the window take the plain march
struct ReflGeom    { static constexpr int SHIFT = 3, TY = VXL_PASS_TY, TW = 3, DT = VXL_PASS_TY / 2 + 1, DW = 2, GH = VXL_REFL_GH, GU = VXL_REFL_GU, QCAP = 0; static constexpr bool NEAR = false; };   // 8-voxel cells (164 steps * |wd| <= 1.5)

// Bounding box of the block's rays -> tile placement -> stage the occupancy tile.
// [flo, fhi] is (close to) the box the thread's rays stay in, [nlo, nhi] the box of their first 20 voxels, in voxel units; threads
// without rays pass valid = false.  The tiles are centred on the union of the boxes, not on the ray origins: the AO rays of a pixel
// fill the hemisphere around its normal, so a block of terrain pixels needs 129 voxels upward and next to nothing downward.
// The plain tile arrives by TMA (one box copy issued by thread 0, vxl_tma.cuh) while all threads stage the two small tiles, whose
// rows start at an arbitrary bit of the level's words (funnel shift).  Ends with a block barrier (which also publishes the LUTs).
template <bool FAST, typename G>
__device__ __forceinline__ BitTile block_prologue(const VolView& V, BlockShared<G>& S, const CUtensorMap* tm_tile, bool valid, float3 flo, float3 fhi,
                                                  float3 nlo, float3 nhi) {
    BitTile T;
    T.w = S.tile; T.ox = T.oy = T.oz = 0; T.enabled = false;
    T.wd = S.dtile; T.dx = T.dy = T.dz = 0;
    T.wn = nullptr; T.nx = T.ny = T.nz = 0;
    constexpr int TPC = 1 << (G::SHIFT - 1);                // texels per cell edge
    T.direct = (V.sx % TPC == 0) && (V.sy % TPC == 0) && (V.sz % TPC == 0) && ((unsigned long long)V.sx * V.sy * V.sz < (1ull << 32));
    T.koff = TileAddr<G::SHIFT, G::TY, G::TW>::texel_koff(V);
    constexpr int NBB = G::NEAR ? 12 : 6;
    if (threadIdx.x < NBB) S.bb[threadIdx.x] = (threadIdx.x % 6) < 3 ? 0x7fffffff : -0x7fffffff - 1;
    if (threadIdx.x < 4) S.acc[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { S.next_item = 0; S.zero = 0u; }
    if (FAST && threadIdx.x == 0) mbar_init(&S.bar, 1u);
    __syncthreads();
    if (!FAST) return T;
    // bounding box of the block's ray boxes (and of their first stretch, for the near tile): warp reductions, then shared atomics
    const int big = 1 << 24;
    auto clampi = [&](float v) { return max(-big, min(big, f2i(floorf(v)))); };
    {
        const float lo[6] = {flo.x, flo.y, flo.z, nlo.x, nlo.y, nlo.z}, hi[6] = {fhi.x, fhi.y, fhi.z, nhi.x, nhi.y, nhi.z};
#pragma unroll
        for (int a = 0; a < NBB / 2; ++a) {
            const int g = a / 3, c = a % 3;
            const int mn = __reduce_min_sync(0xFFFFFFFFu, valid ? clampi(lo[a]) : 0x7fffffff);
            const int mx = __reduce_max_sync(0xFFFFFFFFu, valid ? clampi(hi[a]) : -0x7fffffff - 1);
            if ((threadIdx.x & 31) == 0 && mn != 0x7fffffff) { atomicMin(&S.bb[g * 6 + c], mn); atomicMax(&S.bb[g * 6 + 3 + c], mx); }
        }
    }
    __syncthreads();
    if (S.bb[0] == 0x7fffffff) return T;                    // no ray in this block (uniform)
    const int cx = (S.bb[0] + S.bb[3]) >> 1, cy = (S.bb[1] + S.bb[4]) >> 1, cz = (S.bb[2] + S.bb[5]) >> 1;
    const BitView& L = V.occ[G::SHIFT - 2];
    // x origin: a word boundary of the level array -- of its copy shifted by 16 cells where there is one and that comes closer to
    // (centre - half a window); y origin: a multiple of 4 cells (the borders are)
    const int ax_want = (cx >> G::SHIFT) - G::TW * 16 + L.border;
    const int ax = L.copies == 2 ? ((ax_want + 8) >> 4) << 4 : ((ax_want + 16) >> 5) << 5;
    const int copy = (ax & 16) ? 1 : 0;                     // in the shifted copy array cell a sits at bit a + 16
    T.ox = ax - L.border;
    T.oy = (((cy >> G::SHIFT) - G::TY / 2 + 2) >> 2) << 2; T.oz = (cz >> G::SHIFT) - G::TY / 2;
    T.dx = T.ox >> 1; T.dy = T.oy >> 1; T.dz = T.oz >> 1;   // floor: the dilated tile starts at or before the plain one
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&S.bar, (unsigned)sizeof(S.tile));
        tma_load_4d(S.tile, tm_tile, T.oy + L.border, T.oz + L.border, (ax + 16 * copy) >> 5, copy, &S.bar);
    }
    stage_bits<G::DT, G::DW>(S.dtile, V.dil[G::SHIFT - 2], T.dx, T.dy, T.dz);
    if (G::NEAR) {
        const int ncx = (S.bb[6] + S.bb[9]) >> 1, ncy = (S.bb[7] + S.bb[10]) >> 1, ncz = (S.bb[8] + S.bb[11]) >> 1;
        T.nx = (ncx >> 1) - NEAR_T / 2; T.ny = (ncy >> 1) - NEAR_T / 2; T.nz = (ncz >> 1) - NEAR_T / 2;
        stage_bits<NEAR_T, 1>(S.ntile, V.tex, T.nx, T.ny, T.nz);
        T.wn = S.ntile;
    }
    __syncthreads();
    mbar_wait(&S.bar, 0u);                                  // the box has landed (and is visible to every waiting thread)
    T.enabled = true;
    return T;
}

constexpr int NWARPS = BLOCK_THREADS / 32;

// j-th item of the tile-placement sample of a region: every S-th item of each row of items, staggered from row to row (S = 1, or a
// region too small to sample: every item)
template <typename RG>
struct PlaceSample {
    static constexpr int S = (RG::ITEMS >= 4 * NWARPS && RG::ITEMS_X % VXL_PREPASS_STRIDE == 0) ? VXL_PREPASS_STRIDE : 1;
    static constexpr int COUNT = RG::ITEMS / S, PER_ROW = RG::ITEMS_X / S;
    __device__ __forceinline__ static int item(int j) { const int row = j / PER_ROW; return row * RG::ITEMS_X + (j - row * PER_ROW) * S + (row % S); }
};
// the next work item of the region for this warp (>= REGION_ITEMS: none left)
template <typename BS>
__device__ __forceinline__ int next_item(BS& S) {
    int item = 0;
    if ((threadIdx.x & 31) == 0) item = atomicAdd(&S.next_item, 1);
    return __shfl_sync(0xFFFFFFFFu, item, 0);
}
struct Box3 {
    float3 lo, hi;
    __device__ __forceinline__ Box3() : lo(make_float3(3e38f, 3e38f, 3e38f)), hi(make_float3(-3e38f, -3e38f, -3e38f)) {}
    __device__ __forceinline__ void add(float3 p) {
        lo = make_float3(fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z));
        hi = make_float3(fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z));
    }
};

// UNIFORM: `dist` is the same for every lane of the warp, which allows the masked-lane lockstep loops of
// march_bits.  Measured slower than per-lane exits on config 3 (profiles/r1e vs r1d), so the passes use false.
// MODE: 0 plain march on the bytes, 1 tile march, 2 tile march that also counts the probes that read the volume
template <int MODE, bool SUPER, bool UNIFORM, typename G>
__device__ __forceinline__ float ray_march(const VolView& V, const BitTile& T, float3 origin, float3 dir, float dist, int& steps, unsigned& fetched) {
    if (MODE > 0) return march_bits<SUPER, false, UNIFORM, MODE == 2, G::SHIFT, G::TY, G::TW, G::DT, G::DW, G::GH, G::GU>(V, T, origin, dir, dist, steps, nullptr, fetched);
    return march<false>(V, origin, dir, dist, SUPER ? 2.5f : 0.5f, steps, nullptr);
}

// block-level accumulation of (rays, steps, pixels) into striped global counters
// (slot 3: probes that read the volume in the tile march -- diagnostics).  S.acc is zeroed by block_prologue.
template <typename BS>
__device__ __forceinline__ void flush_stats(BS& S, unsigned long long* __restrict__ g_stats, unsigned rays, unsigned steps,
                                            unsigned pixels, unsigned exact) {
    rays = __reduce_add_sync(0xFFFFFFFFu, rays);
    steps = __reduce_add_sync(0xFFFFFFFFu, steps);
    pixels = __reduce_add_sync(0xFFFFFFFFu, pixels);
    exact = __reduce_add_sync(0xFFFFFFFFu, exact);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&S.acc[0], rays); atomicAdd(&S.acc[1], steps); atomicAdd(&S.acc[2], pixels); atomicAdd(&S.acc[3], exact); }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned v = S.acc[threadIdx.x];
        if (v) atomicAdd(&g_stats[(blockIdx.x & (STAT_SLOTS - 1)) * 4 + threadIdx.x], (unsigned long long)v);
    }
}

static_assert(smem_bytes<AmbientGeom, true>() <= (size_t)(233472 / VXL_AMBIENT_BLOCKS - 1024), "k_ambient: shared memory per block exceeds the SM's share");
static_assert(smem_bytes<LocalGeom, true>() + 1024 <= (size_t)(233472 / VXL_PASS_BLOCKS - 1024), "k_local_lights: shared memory per block exceeds the SM's share");
static_assert(smem_bytes<ReflGeom, true>() <= (size_t)(233472 / VXL_PASS_BLOCKS - 1024), "k_reflection: shared memory per block exceeds the SM's share");

constexpr int AO_N2 = 23;          // phase-2 probes of a SuperSparse ray with dist = 128: d = 17.5, 22.5, ..., 127.5 (:121)


// (256 * d / 128)^2 of a SuperSparse ray that returns at probe k: d = 2.5 (k + 1) in phase 1, 17.5 + 5 (k - 6) = 2.5 (2k - 5) in
// phase 2 (Light.frag:181-211), so 256 d / 128 = 5 (k + 1) or 5 (2k - 5); a miss returns dist = 128 -> 256.
__device__ __forceinline__ unsigned ao_term(int k) {
    const unsigned n = 5u * (unsigned)(k < 6 ? k + 1 : 2 * k - 5);
    return n * n;
}
constexpr unsigned AO_MISS = 65536u;
constexpr int AO_POOL_MAX_RAYS = 256;     // the sums below stay < 2^24, where float and integer accumulation agree exactly

// One AO ray the slow way (a pixel whose rays leave the volume, or whose coordinates are too large for the scan's folded
// addressing): the per-ray tile march / plain march of round 1.  Cold code, one copy.
template <int MODE, typename G>
__device__ __noinline__ float ao_ray_slow(const VolView& V, const BitTile& C, float3 origin, float3 dir, int& steps, unsigned& exact) {
    const float d = march_scan_super<false, MODE == 2, G::NEAR, G::SHIFT, G::TY, G::TW, AO_N2>(V, C, origin, dir, 128.0f, steps, nullptr, exact) / 128.0f;   // :121
    return d * d;
}

// Where the AO rays of a pixel can go (LightAmbient.frag:112-119).  Every direction is tangent * x + bitangent * y + normal * z with
// (x, y, z) the LUT's hemisphere sample: x^2 + y^2 + z^2 = u (cos^2 + sin^2) + (1 - u) = 1 up to a few ulp and z >= 0.  On that half
// sphere the component along axis a, T_a x + B_a y + N_a z, is at most |(T_a, B_a, N_a)| when N_a >= 0 (Cauchy-Schwarz, attained at
// z >= 0) and at most |(T_a, B_a)| when N_a < 0 (the N_a z term only subtracts; the rest is bounded on the rim z = 0); the minimum
// mirrors that.  Result: -blo_a <= dir_a <= bhi_a; the factor absorbs the roundings.
__device__ __forceinline__ void ao_bounds(float3 normal, float3& blo, float3& bhi) {
    const float3 t = fabsf(normal.z) > 0.5f ? make_float3(0.0f, -normal.z, normal.y) : make_float3(-normal.y, normal.x, 0.0f);    // :112
    const float3 b = cross3(normal, t);                                                                                           // :113
    const float n[3] = {normal.x, normal.y, normal.z};
    const float r2[3] = {t.x * t.x + b.x * b.x, t.y * t.y + b.y * b.y, t.z * t.z + b.z * b.z};
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float full = sqrtf(r2[a] + n[a] * n[a]) * 1.0001f, rim = sqrtf(r2[a]) * 1.0001f;
        hi[a] = n[a] >= 0.0f ? full : rim;
        lo[a] = n[a] <= 0.0f ? full : rim;
    }
    blo = make_float3(lo[0], lo[1], lo[2]); bhi = make_float3(hi[0], hi[1], hi[2]);
}

// The AO rays of a warp's 32 pixels with a pooled resolve (LightAmbient.frag:111-126, n_ao rays per pixel).
//
// Every lane scans its own rays, two at a time (scan_super_pair: all 29 probes of both, branch-free, packed f32x2 additions).
// What the scan cannot decide -- a probe in an occupied cell needs the reference's texel test on the volume bytes -- used to be
// resolved by the lane itself, a loop that ran at 4-8 live lanes (34 % of the round-1 kernel's instructions).  Here the undecided
// rays go on a per-warp stack in shared memory, (2.5 dir, candidate bits, owner lane), and whenever 32 are pending the warp tests
// the first candidate of each in one converged pass: a hit (or the last candidate failing) adds the ray's term to the owner's sum
// with a shared-memory atomic, a failure with candidates left goes back on the stack.  The reference returns at the FIRST probe
// whose test passes, and candidates are tested in probe order per ray, so the result is the same ray by ray.
//
// Where a pixel's rays look: the block's tiles in shared memory when the box around its rays lies inside them (scan_precheck);
// otherwise -- far terrain, silhouettes: the block's ray origins spread over more voxels than the tile has slack -- the same scan
// reads the 4-voxel level from global memory (L2-resident), as long as the box lies inside the volume.  Both feed the same stack.
//
// acc = sum over rays of (d / 128)^2 is exact in binary32 whatever the order: every term is a multiple of 2^-16 that is <= 1, so for
// n_ao <= 256 all partial sums are multiples of 2^-16 below 2^8.  It is therefore kept as the integer sum of (256 d / 128)^2.
template <int MODE, typename G>
__device__ __forceinline__ float ao_pooled(const VolView& V, const BitTile& C, BlockShared<G>& S, const FrameView& F, const ViewK& K, const PixelCtx& p,
                                           bool lit, float3 origin, float3 normal, float3 blo, float3 bhi, uint32_t n0, int n_ao, int& steps, unsigned& exact) {
    constexpr int N = 6 + AO_N2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    uint4* const qe = S.q_ent[warp];
    uint8_t* const qo = S.q_own[warp];
    unsigned* const wacc = S.ao_acc + warp * 32;
    wacc[lane] = 0u;
    float3 tangent = make_float3(0.f, 0.f, 0.f), bitangent = tangent;
    // mode of this pixel: 0 = no rays, 1 = scan in shared memory, 2 = scan in global memory, 3 = slow path
    int mode = 0;
    bool near_ok = false;
    if (lit) {
        tangent = fabsf(normal.z) > 0.5f ? make_float3(0.0f, -normal.z, normal.y) : make_float3(-normal.y, normal.x, 0.0f);    // :112
        bitangent = cross3(normal, tangent);                                                                                   // :113
        // one eligibility test per pixel covers all its rays: -blo_a <= dir_a <= bhi_a (ao_bounds); reach: 128 + 1 voxels of march, 20 for
        // the near tile's 8 probes
        const ScanPre pre = scan_precheck<G::SHIFT, G::TY, G::TW>(C, origin, blo, bhi, 129.0f, 20.0f, -BM_LOWCOORD);
        near_ok = pre.near_ok;
        // the box around the rays no more than 150 voxels below zero and 160 voxels past the volume's far faces (the level array
        // has 192 voxels of border around it: texel -1 repeats texel 0, the rest is empty) and below the magic floor's
        // coordinate limit: the level array can be indexed directly
        const float3 rl = blo * (129.0f * 1.00002f), rh = bhi * (129.0f * 1.00002f);
        const float3 hi = make_float3(fminf((float)(2 * V.sx + 160), BM_MAXCOORD), fminf((float)(2 * V.sy + 160), BM_MAXCOORD), fminf((float)(2 * V.sz + 160), BM_MAXCOORD));
        const float lowc = BM_MARGIN - BM_LOWCOORD;
        const bool inside = C.direct && origin.x - rl.x >= lowc && origin.y - rl.y >= lowc && origin.z - rl.z >= lowc &&
                            origin.x + rh.x <= hi.x - BM_MARGIN && origin.y + rh.y <= hi.y - BM_MARGIN && origin.z + rh.z <= hi.z - BM_MARGIN;
        mode = pre.ok ? 1 : (inside ? 2 : 3);                                       // (NaN anywhere fails the comparisons)
#if VXL_AO_PROMOTE
        // a warp whose pixels disagree would run the scan twice, once per source: when some pixel has to look in global memory,
        // the others of the warp do too (the same bits, read from L2)
        if (__any_sync(__activemask(), mode == 2) && mode == 1 && inside) { mode = 2; near_ok = false; }
#endif
#ifdef VXL_EXP_SKIP        // timing experiments only (results are wrong): drop the pixels of mode >= VXL_EXP_SKIP
        if (mode >= VXL_EXP_SKIP) mode = 0;
#endif
#ifdef VXL_EXP_MODECNT     // diagnostics: `exact` (vxl_debug_fetched_probes) counts the pixels of mode VXL_EXP_MODECNT
        exact += mode == VXL_EXP_MODECNT ? 1u : 0u;
#endif
    }
    // one eps for everybody (test_super_cand): no coordinate of a scanned ray is further than 200 voxels from the volume
    const float eps = (fminf((float)(2 * max(V.sx, max(V.sy, V.sz)) + 200), BM_MAXCOORD) + 1.0f) * (1.0f / 262144.0f);
    typedef ScanLook<false, (unsigned)(G::TY * G::TY), (unsigned)G::TY> TileLook;
    TileLook look_tile = TileLook::make(C.w, G::SHIFT, C.ox, C.oy, C.oz, 0u, 0u);
    // The folded base (window address minus a compile-time constant) has to sit in ONE register so that a lookup's address is one
    // LEA; left alone the compiler keeps the window address in a uniform register and adds the constant per lookup.
    look_tile.sbase += *(volatile unsigned*)&S.zero;
    const ScanLook<false> look_near = (G::NEAR && near_ok) ? ScanLook<false>::make(C.wn, 1, C.nx, C.ny, C.nz, (unsigned)NEAR_T, (unsigned)NEAR_T)
                                                           : ScanLook<false>::make(C.w, G::SHIFT, C.ox, C.oy, C.oz, (unsigned)(G::TY * G::TY), (unsigned)G::TY);
    int qn = 0;                       // entries on the stack (warp-uniform)
    unsigned own = 0u;                // terms of the rays this lane decided itself
    float accf = 0.0f;                // mode 3: sequential float sum like the reference
    __syncwarp();

    // what a lane can decide itself about a scanned ray; returns the candidate word that has to go on the stack (0 = decided)
    auto decide = [&](unsigned cand) -> unsigned {
#ifdef VXL_EXP_ABLATE             // timing experiments only (results are wrong): 1 = no resolve, 2 = no scan either
        own += cand; cand = 0u;
#endif
        if (cand == 0u) { own += AO_MISS; steps += N; return 0u; }
        if (G::NEAR && near_ok) {
            // inside the near tile bits 6 and 7 ARE the reference's test (texel != 0): a first candidate there is the hit
            const int k = __ffs((int)cand) - 1;
            if (k == 6 || k == 7) { own += ao_term(k); steps += k + 1; return 0u; }
        }
        return cand;
    };
    // the noise texels of a ray pair are fetched one trip ahead, so that their L2 latency hides under the scan of the current pair
    auto noise_of = [&](int i) -> uint32_t { return (i == 0) ? n0 : ((mode != 0 && i < n_ao) ? get_noise(F, K, p, i) : 0u); };
    auto ray_dir = [&](uint32_t ni) -> float3 {
        const float3 rv = cosine_sample_hemisphere(S.lut, ni, ni >> 8);                           // :118
        return tangent * rv.x + bitangent * rv.y + normal * rv.z;                                 // :119
    };
    uint32_t na = noise_of(0), nb = noise_of(1);
    const int n_pairs = (n_ao + 1) >> 1;
    for (int it = 0; it <= n_pairs; ++it) {             // the last trip only drains the stack
        unsigned ca = 0u, cb = 0u;
        float3 sa = make_float3(0.f, 0.f, 0.f), sb = sa;
        if (it < n_pairs && mode != 0) {
            const int i = 2 * it;
            const bool two = i + 1 < n_ao;
            const float3 da = ray_dir(na), db = two ? ray_dir(nb) : da;                           // odd n_ao: the last ray scans twice, counts once
            na = noise_of(i + 2); nb = noise_of(i + 3);
            if (mode == 1) {
#if defined(VXL_EXP_ABLATE) && VXL_EXP_ABLATE == 2
                ca = __float_as_uint(da.x + da.y + da.z); cb = __float_as_uint(db.x + db.y + db.z);
#else
                scan_super_pair<AO_N2>(look_near, look_tile, origin, da, db, ca, cb);
#endif
            } else if (mode == 2) {
                const BitView& L4 = V.occ[G::SHIFT - 2];
                const ScanLook<true> look_glob = ScanLook<true>::make(L4.words, G::SHIFT, -L4.border, -L4.border, -L4.border, (unsigned)L4.cyp, (unsigned)(L4.xw * L4.cyp));   // [az][word][ay]
                scan_super_pair<AO_N2>(look_glob, look_glob, origin, da, db, ca, cb);
            }
            if (mode != 3) {
                sa = da * 2.5f; sb = db * 2.5f;
                ca = decide(ca);
                cb = two ? decide(cb) : 0u;
            } else {
                accf += ao_ray_slow<MODE, G>(V, C, origin, da, steps, exact);
                if (two) accf += ao_ray_slow<MODE, G>(V, C, origin, db, steps, exact);
            }
        }
        // push the undecided rays, one ray of the pair per round (a round adds at most 32 entries to at most 31 left over),
        // each followed by the converged passes over the top of the stack: ONE copy of that code for both rounds and for the
        // final trip that drains whatever is left
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
        const unsigned ch = h ? cb : ca;
        const unsigned mh = __ballot_sync(0xFFFFFFFFu, ch != 0u);
        if (mh) {
            if (ch != 0u) {
                const float3 sh = h ? sb : sa;
                const int idx = qn + __popc(mh & lt);
                qe[idx] = make_uint4(__float_as_uint(sh.x), __float_as_uint(sh.y), __float_as_uint(sh.z), ch);
                qo[idx] = (uint8_t)((unsigned)lane | (near_ok ? 32u : 0u));
            }
            qn += __popc(mh);
            __syncwarp();
        }
        // converged passes over the top of the stack: full ones while rays are still being scanned, whatever is left at the end
        const int need = it < n_pairs ? 32 : 1;
        while (qn >= need) {
            const int nb = qn < 32 ? qn : 32;
            const int base = qn - nb;
            const bool act = lane < nb;
            uint4 e = make_uint4(0u, 0u, 0u, 1u);
            unsigned ow = (unsigned)lane;
            if (act) { e = qe[base + lane]; ow = qo[base + lane]; }
            const int owner = (int)(ow & 31u);
            const float3 o = make_float3(__shfl_sync(0xFFFFFFFFu, origin.x, owner), __shfl_sync(0xFFFFFFFFu, origin.y, owner), __shfl_sync(0xFFFFFFFFu, origin.z, owner));
            const int k = __ffs((int)e.w) - 1;
            const unsigned rest = e.w & (e.w - 1u);
            const float3 s1 = make_float3(__uint_as_float(e.x), __uint_as_float(e.y), __uint_as_float(e.z));
            const bool hit = act && test_super_cand<MODE == 2>(V, C.koff, o, s1, k, (ow >> 5) != 0u, eps, exact);
            const bool done = act && (hit || rest == 0u);                 // the ray returns at probe k, or no candidate is left: a miss
            if (done) {
                atomicAdd(&wacc[owner], hit ? ao_term(k) : AO_MISS);
                steps += hit ? k + 1 : N;
            }
            const bool again = act && !done;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, again);
            if (again) {
                const int idx = base + __popc(m & lt);
                qe[idx] = make_uint4(e.x, e.y, e.z, rest);
                qo[idx] = (uint8_t)ow;
            }
            qn = base + __popc(m);
            __syncwarp();
        }
        }
    }
    const float acc = mode == 3 ? accf : (float)(own + wacc[lane]) * (1.0f / 65536.0f);
    return (acc / (float)n_ao) * 0.05f;                                                           // :125
}

// -------------------------------------------------------------------------------------------------
// LightAmbient.frag:134-175 + calculateAmbientIrradiance :111-126
// -------------------------------------------------------------------------------------------------
// what the ambient pass derives from the G-buffer for one pixel
struct AmbPixel { bool lit; float3 normal, wcp0; float bias; };
__device__ __forceinline__ AmbPixel ambient_pixel(const FrameView& F, const ViewK& K, const PixelCtx& p) {
    AmbPixel a;
    a.lit = false; a.normal = make_float3(0.f, 0.f, 0.f); a.wcp0 = a.normal; a.bias = 0.0f;
    if (p.valid) {
        const float depth = unorm24(__ldg(F.depth24 + p.idx));
        if (depth < 0.999f) {                                                         // :138
            a.lit = true;
            const float3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));              // :141
            a.normal = decode_normal(__ldg(F.normal + p.idx));                         // :142
            a.wcp0 = xyz(mat_mul(K.InvView, make_float4(pos.x, pos.y, pos.z, 1.0f))) * 10.0f;   // :150
            a.bias = gsmoothstep(0.0f, 0.2f, depth) * 50.0f + 1.5f;                    // :158
        }
    }
    return a;
}

template <int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS, VXL_AMBIENT_BLOCKS) k_ambient(const __grid_constant__ VolView V, const __grid_constant__ CUtensorMap tm_tile, const __grid_constant__ FrameView F, const __grid_constant__ ViewK K, const float* __restrict__ g_lut, int n_ao,
                                                 float* __restrict__ out_shadow, float* __restrict__ out_ao,
                                                 unsigned long long* __restrict__ g_stats) {
    typedef AmbientGeom G;
    typedef AmbientRegion RG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockShared<G>& S = *reinterpret_cast<BlockShared<G>*>(smem_raw);
    load_luts(S.lut, g_lut);
    const RegionCtx R = region_ctx<RG>(F);
    const int warp = threadIdx.x >> 5;
    const bool want_ao = out_ao && n_ao > 0;
    // ---- the boxes the region's rays stay in, for the placement of the block's tiles (approximate: the exact origins come later) ----
    // STATIC: one item per warp (the AO pass's 32x16 regions) -- the pixels decoded here are the ones the warp works on below
    constexpr bool STATIC = VXL_AMB_STATIC && RG::ITEMS == NWARPS;
    PixelCtx p0 = item_pixel<RG>(F, K, R, warp);
    AmbPixel a0 = ambient_pixel(F, K, p0);
    Box3 far, near;
    bool any = false;
    for (int item = warp; item < RG::ITEMS; item += NWARPS) {
        const AmbPixel a = item == warp ? a0 : ambient_pixel(F, K, item_pixel<RG>(F, K, R, item));
        if (!a.lit) continue;
        any = true;
        const float3 o = a.wcp0 + a.normal * a.bias;
        far.add(o); near.add(o);
        if (want_ao) {
            float3 blo, bhi;
            ao_bounds(a.normal, blo, bhi);
            far.add(o - blo * 130.0f); far.add(o + bhi * 130.0f);
            near.add(o - blo * 21.0f); near.add(o + bhi * 21.0f);
        }
        if (out_shadow) {          // 128 voxels along normalize(mix(SUN_DIR, jitter, 0.5)): (54, 72, 91) +- 14 of jitter, from an origin jittered by 1.25
            far.add(o - make_float3(16.0f, 16.0f, 16.0f)); far.add(o + make_float3(70.0f, 88.0f, 107.0f));
        }
    }
    const BitTile C = block_prologue<(MODE > 0), G>(V, S, &tm_tile, any, far.lo, far.hi, near.lo, near.hi);
    unsigned rays = 0, pixels = 0, exact = 0;
    int steps = 0;
    // the AO rays of tile-march launches go through the warp-pooled resolve (all 32 lanes take part, lit or not)
    const bool POOL = MODE > 0 && G::QCAP > 0 && n_ao <= AO_POOL_MAX_RAYS && n_ao >= VXL_AO_POOL_MIN;
    // ---- the region's 8x4-pixel work items, handed to whichever warp is free ----
    for (int item = STATIC ? warp : next_item(S); item < RG::ITEMS; item = STATIC ? RG::ITEMS : next_item(S)) {
        const PixelCtx p = STATIC ? p0 : item_pixel<RG>(F, K, R, item);
        const AmbPixel a = STATIC ? a0 : ambient_pixel(F, K, p);
        const float3 normal = a.normal;
        float shadow = 1.0f, ao = 0.0f;
        float3 ao_origin = make_float3(0.f, 0.f, 0.f), blo = ao_origin, bhi = ao_origin;
        uint32_t ao_noise = 0u;
        if (a.lit) {
            float3 wd = normalize3(make_float3(0.3f, 0.4f, 0.5f));                     // SUN_DIR :15,:149
            float3 wcp = a.wcp0;
            const uint32_t n = get_noise(F, K, p, -1);
            float3 randomVec = cosine_sample_hemisphere(S.lut, n, n >> 8) * 0.1f;      // :151
            randomVec.z *= gsign(unorm8(n >> 16) - 0.5f);                             // :152
            wd = mix3(wd, randomVec, 0.5f);                                            // :153
            wd = normalize3(wd);                                                       // :154
            wcp = wcp + wd * (unorm8(n >> 24) * 1.0f);                                 // :155
            wcp = wcp + randomVec * 2.5f;                                              // :156
            const float3 origin = wcp + normal * a.bias;
            if (out_shadow) {
                if (ray_march<MODE, false, false, G>(V, C, origin, wd, 128.0f, steps, exact) != 128.0f) shadow = 0.0f;   // :167-169
                rays += 1;
            }
            if (want_ao) ao_bounds(normal, blo, bhi);
            if (want_ao && !POOL) {
                const float3 tangent = fabsf(normal.z) > 0.5f ? make_float3(0.0f, -normal.z, normal.y)
                                                             : make_float3(-normal.y, normal.x, 0.0f);    // :112
                const float3 bitangent = cross3(normal, tangent);                                         // :113
                // one eligibility test per pixel covers all its rays (ao_bounds; reach: 128 + 1 voxels of march, 20 for the near tile's 8 probes)
                ScanPre pre = ScanPre{false, false, 0.0f};
                if (MODE > 0) pre = scan_precheck<G::SHIFT, G::TY, G::TW>(C, origin, blo, bhi, 129.0f, 20.0f);
                float acc = 0.0f;
                for (int i = 0; i < n_ao; ++i) {
                    const uint32_t ni = (i == 0) ? n : get_noise(F, K, p, i);
                    const float3 rv = cosine_sample_hemisphere(S.lut, ni, ni >> 8);                       // :118
                    const float3 dir = tangent * rv.x + bitangent * rv.y + normal * rv.z;                 // :119
                    const float d = (MODE > 0 ? march_scan_super<false, MODE == 2, G::NEAR, G::SHIFT, G::TY, G::TW, AO_N2>(V, C, origin, dir, 128.0f, steps, nullptr, exact, pre)
                                              : march<false>(V, origin, dir, 128.0f, 2.5f, steps, nullptr)) / 128.0f;            // :121
                    acc += d * d;
                }
                ao = (acc / (float)n_ao) * 0.05f;                                                         // :125
            }
            if (want_ao) rays += (unsigned)n_ao;
            ao_origin = origin; ao_noise = n;
            pixels += 1;
        }
        if (G::QCAP > 0 && POOL && want_ao) {
            const float v = ao_pooled<MODE, G>(V, C, S, F, K, p, a.lit, ao_origin, normal, blo, bhi, ao_noise, n_ao, steps, exact);
            if (a.lit) ao = v;
        }
        if (p.valid) {
            if (out_shadow) out_shadow[p.idx] = shadow;
            if (out_ao) out_ao[p.idx] = ao;
        }
    }
    if (F.n_mirror) {                                   // several GPUs: the region's rows into every copy of the stack (vxl_pixel.cuh)
        __syncthreads();
        mirror_region<RG>(F, R, out_shadow, 1, 0);
        mirror_region<RG>(F, R, out_ao, 1, 0);
    }
    flush_stats(S, g_stats, rays, (unsigned)steps, pixels, exact);
}

// -------------------------------------------------------------------------------------------------
// LightPoint.frag:85-129 / LightSpot.frag:73-117 -- all lights of the list in one launch; the
// G-buffer, noise and world position are read / derived once per pixel instead of once per light.
// -------------------------------------------------------------------------------------------------
template <bool SPOT, int MODE, typename RG>
__global__ void __launch_bounds__(BLOCK_THREADS, VXL_PASS_BLOCKS) k_local_lights(const __grid_constant__ VolView V, const __grid_constant__ CUtensorMap tm_tile, const __grid_constant__ FrameView F, const __grid_constant__ ViewK K, const float* __restrict__ g_lut,
                                                      const float* __restrict__ lights, int n_lights,
                                                      float* __restrict__ out_shadow, size_t plane_stride,
                                                      unsigned long long* __restrict__ g_stats) {
    typedef LocalGeom G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockShared<G>& S = *reinterpret_cast<BlockShared<G>*>(smem_raw);
    __shared__ float s_light[VXL_MAX_LIGHTS * 4];   // position.xyz, range
    load_luts(S.lut, g_lut);
    constexpr int STRIDE = SPOT ? 16 : 8;
    for (int i = threadIdx.x; i < n_lights * 4; i += blockDim.x) s_light[i] = lights[(i >> 2) * STRIDE + (i & 3)];
    __syncthreads();
    const RegionCtx R = region_ctx<RG>(F);
    const int warp = threadIdx.x >> 5;
    // Tile placement considers the pixels that will cast a ray from the scene: not sky (which the reference shades
    // like any other pixel -- no depth test here -- but whose world position is ~40000 voxels away) and inside
    // some light's range.  A block without such pixels stages nothing.
    // The box of a pixel's shadow rays: from the surface point towards every light in range, as far as the march goes
    // (min(hitDist, 164) + 1 voxels along a unit direction; hitDist = 10.5 |L| world units = 1.05 x the way to the light).
    Box3 box;
    bool any = false;
    for (int j = warp; j < PlaceSample<RG>::COUNT; j += NWARPS) {
        const PixelCtx p = item_pixel<RG>(F, K, R, PlaceSample<RG>::item(j));
        if (!p.valid) continue;
        const float depth = unorm24(__ldg(F.depth24 + p.idx));
        if (!(depth < 0.999f)) continue;
        const float3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));                        // LightPoint.frag:89
        const float3 worldPos = xyz(mat_mul(K.InvView, make_float4(pos.x, pos.y, pos.z, 1.0f)));          // :95
        const float3 o = worldPos * 10.0f;
        for (int li = 0; li < n_lights; ++li) {
            const float3 L = make_float3(s_light[li * 4], s_light[li * 4 + 1], s_light[li * 4 + 2]) - worldPos;
            const float len = length3(L);
            if (len > s_light[li * 4 + 3]) continue;
            if (!any) { box.add(o - make_float3(3.f, 3.f, 3.f)); box.add(o + make_float3(3.f, 3.f, 3.f)); }     // origin jitter: normal * 0.5 + wd * n.w + rv * 2.5
            any = true;
            box.add(o + L * (fminf(len * 10.5f, 166.0f) / fmaxf(len, 1e-6f)));
        }
    }
    const BitTile C = block_prologue<(MODE > 0), G>(V, S, &tm_tile, any, box.lo, box.hi, box.lo, box.hi);
    unsigned rays = 0, pixels = 0, exact = 0;
    int steps = 0;
    for (int item = next_item(S); item < RG::ITEMS; item = next_item(S)) {
        const PixelCtx p = item_pixel<RG>(F, K, R, item);
        if (p.valid) {
        const float depth = unorm24(__ldg(F.depth24 + p.idx));
        const float3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));                        // LightPoint.frag:89
        const float3 normal = decode_normal(__ldg(F.normal + p.idx));                        // :90
        const float3 worldPos = xyz(mat_mul(K.InvView, make_float4(pos.x, pos.y, pos.z, 1.0f)));          // :95
        const uint32_t n = get_noise(F, K, p, -1);
        float3 rv0 = cosine_sample_hemisphere(S.lut, n, n >> 8) * 0.1f;                      // :111
        rv0.z *= gsign(unorm8(n >> 16) - 0.5f);                                             // :112
        const float nw = unorm8(n >> 24) * 1.0f;
        bool counted = false;
        for (int li = 0; li < n_lights; ++li) {
            const float3 lpos = make_float3(s_light[li * 4], s_light[li * 4 + 1], s_light[li * 4 + 2]);
            const float range = s_light[li * 4 + 3];
            const float3 lightDir = lpos - worldPos;                                         // :97
            const float lightDistance = length3(lightDir);                                   // :98
            float shadow = 1.0f;
            if (!(lightDistance > range)) {                                                  // :100-103
                float3 wd; float hitDist;
                if (!SPOT) { wd = lightDir; hitDist = lightDistance * 10.5f; }               // :108-109
                else { wd = normalize3(lightDir) * 10.0f; hitDist = lightDistance * 10.0f; } // LightSpot.frag:96-97
                float3 wcp = worldPos * 10.0f;                                               // :110
                wd = mix3(wd, rv0, 0.5f);                                                    // :113
                wd = normalize3(wd);                                                         // :114
                wcp = wcp + wd * nw;                                                         // :115
                wcp = wcp + rv0 * 2.5f;                                                      // :116
                if (ray_march<MODE, SPOT, false, G>(V, C, wcp + normal * 0.5f, wd, hitDist, steps, exact) < hitDist) shadow = 0.0f;   // :125
                rays += 1;
                counted = true;
            }
            out_shadow[(size_t)li * plane_stride + p.idx] = shadow;
        }
        if (counted) pixels += 1;
        }
    }
    if (F.n_mirror) {                                   // several GPUs: the region's rows of every light plane into every copy of the stack
        __syncthreads();
        mirror_region<RG>(F, R, out_shadow, n_lights, plane_stride);
    }
    flush_stats(S, g_stats, rays, (unsigned)steps, pixels, exact);
}

// -------------------------------------------------------------------------------------------------
// LightReflection.frag:60-113
// -------------------------------------------------------------------------------------------------
struct ReflPixel { bool lit; float3 pos, normal, wcp0, wd; };      // wd: the mirror direction in world space (:79-92)
__device__ __forceinline__ ReflPixel reflection_pixel(const FrameView& F, const ViewK& K, const PixelCtx& p) {
    ReflPixel a;
    a.lit = false; a.pos = make_float3(0.f, 0.f, 0.f); a.normal = a.pos; a.wcp0 = a.pos; a.wd = a.pos;
    if (p.valid) {
        const float depth = unorm24(__ldg(F.depth24 + p.idx));
        if (depth < 0.999f) {                                                               // :88
            a.lit = true;
            a.pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));                               // :64
            a.normal = decode_normal(__ldg(F.normal + p.idx));                               // :65
            a.wcp0 = xyz(mat_mul(K.InvView, make_float4(a.pos.x, a.pos.y, a.pos.z, 1.0f))) * 10.0f;  // :93
            const float3 Vv = normalize3(a.pos) * -1.0f;                                     // :79
            const float3 N = xyz(mat_mul(K.View, make_float4(a.normal.x, a.normal.y, a.normal.z, 0.0f)));   // :80
            const float3 I = Vv * -1.0f;
            const float3 Rr = I - N * dot3(N, I) * 2.0f;                                     // :81
            a.wd = normalize3(xyz(mat_mul(K.InvView, make_float4(Rr.x, Rr.y, Rr.z, 0.0f))));          // :92
        }
    }
    return a;
}

template <int MODE, typename RG>
__global__ void __launch_bounds__(BLOCK_THREADS, VXL_PASS_BLOCKS) k_reflection(const __grid_constant__ VolView V, const __grid_constant__ CUtensorMap tm_tile, const __grid_constant__ FrameView F, const __grid_constant__ ViewK K, const float* __restrict__ g_lut,
                                                    float* __restrict__ out_t, unsigned long long* __restrict__ g_stats) {
    typedef ReflGeom G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockShared<G>& S = *reinterpret_cast<BlockShared<G>*>(smem_raw);
    load_luts(S.lut, g_lut);
    const RegionCtx R = region_ctx<RG>(F);
    const int warp = threadIdx.x >> 5;
    // the box of a reflection ray: from the surface point along the mirror direction, 165 steps of up to 1.5 voxels; the roughness
    // jitter (:96) bends it by at most a tenth
    Box3 box;
    bool any = false;
    for (int j = warp; j < PlaceSample<RG>::COUNT; j += NWARPS) {
        const ReflPixel a = reflection_pixel(F, K, item_pixel<RG>(F, K, R, PlaceSample<RG>::item(j)));
        if (!a.lit) continue;
        any = true;
        const float3 o = a.wcp0 + a.normal, e = o + a.wd * 210.0f;
        box.add(o - make_float3(20.f, 20.f, 20.f)); box.add(o + make_float3(20.f, 20.f, 20.f));
        box.add(e - make_float3(20.f, 20.f, 20.f)); box.add(e + make_float3(20.f, 20.f, 20.f));
    }
    const BitTile C = block_prologue<(MODE > 0), G>(V, S, &tm_tile, any, box.lo, box.hi, box.lo, box.hi);
    unsigned rays = 0, pixels = 0, exact = 0;
    int steps = 0;
    for (int item = next_item(S); item < RG::ITEMS; item = next_item(S)) {
        const PixelCtx p = item_pixel<RG>(F, K, R, item);
        const ReflPixel a = reflection_pixel(F, K, p);
        float t = 256.0f;
        if (a.lit) {
            const float roughness = unorm8(__ldg(F.material + p.idx));                       // :68
            float3 wd = a.wd;
            float3 wcp = a.wcp0;
            const uint32_t n = get_noise(F, K, p, -1);
            float3 rv = cosine_sample_hemisphere(S.lut, n, n >> 8);                          // :94
            rv.z *= gsign(unorm8(n >> 16) - 0.5f);                                          // :95
            wd = mix3(wd, rv, roughness * 0.1f);                                             // :96
            const float nw = unorm8(n >> 24);
            wcp = wcp + a.normal * nw;                                                       // :97
            wd = wd * (1.0f + nw * 0.5f);                                                    // :98
            t = ray_march<MODE, false, false, G>(V, C, wcp + a.normal, wd, 256.0f, steps, exact);        // :113
            rays += 1; pixels += 1;
        }
        if (p.valid) out_t[p.idx] = t;
    }
    if (F.n_mirror) {
        __syncthreads();
        mirror_region<RG>(F, R, out_t, 1, 0);
    }
    flush_stats(S, g_stats, rays, (unsigned)steps, pixels, exact);
}

// -------------------------------------------------------------------------------------------------
// Ray-level entry (level-1 parity): explicit rays in, 48-byte records out.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_trace_rays(VolView V, const vxl_ray* __restrict__ rays, long long n, int variant,
                                                    vxl_hit* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const vxl_ray r = rays[i];
    vxl_hit h;
    h.t = 0.f; h.steps = 0; h.vx = h.vy = h.vz = 0; h.status = 0; h.px = h.py = h.pz = 0.f; h.nx = h.ny = h.nz = 0.f;
    const float3 o = make_float3(r.ox, r.oy, r.oz), d = make_float3(r.dx, r.dy, r.dz);
    if (variant == VXL_TRACE_DDA) {
        DdaResult R;
        dda(V, o, d, r.dist, R);
        h.t = R.t; h.steps = R.nt; h.status = R.status; h.vx = R.vx; h.vy = R.vy; h.vz = R.vz;
        if (R.status == 1) { h.px = R.hit.x; h.py = R.hit.y; h.pz = R.hit.z; h.nx = R.normal.x; h.ny = R.normal.y; h.nz = R.normal.z; }
    } else {
        MarchResult M;
        int steps = 0;
        march<true>(V, o, d, r.dist, variant == VXL_TRACE_SPARSE ? 0.5f : 2.5f, steps, &M);
        h.t = M.d; h.steps = M.steps; h.status = M.status; h.vx = M.vx; h.vy = M.vy; h.vz = M.vz;
        h.px = M.pos.x; h.py = M.pos.y; h.pz = M.pos.z;
    }
    out[i] = h;
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_pass_ambient(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame, int n_ao,
                     float* out_shadow, float* out_ao) {
    if (!ctx || !vol || !view || !frame || n_ao < 0) { set_error("vxl_pass_ambient: bad argument"); return VXL_ERR_INVALID; }
    FrameView F;
    if (int e = frame_view(frame, &F)) return e;
    apply_band(ctx, F);
    apply_mirrors(ctx, F);
    if (!out_shadow && !out_ao) return VXL_OK;
    if (F.n_tiles == 0) return VXL_OK;
    if (vol->dirty) { if (int e = vxl_volume_build_occupancy(vol)) return e; }
    const CUtensorMap* tm = nullptr;
    if (int e = level_tensor_map(vol->occ[AmbientGeom::SHIFT - 2], AmbientGeom::TW, AmbientGeom::TY, AmbientGeom::TY, &tm)) return e;
#define VXL_AMB(MODE_)                                                                                                                  \
    do {                                                                                                                            \
        if (MODE_) VXL_CUDA(cudaFuncSetAttribute(k_ambient<MODE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<AmbientGeom, (MODE_ > 0)>())); \
        k_ambient<MODE_><<<grid_regions<AmbientRegion>(F), BLOCK_THREADS, smem_bytes<AmbientGeom, (MODE_ > 0)>(), ctx->stream>>>(                       \
            vol_view(vol), *tm, F, make_viewk(view), ctx->d_luts, n_ao, out_shadow, out_ao, ctx->d_stats);                               \
    } while (0)
    if (ctx->variant == 0) VXL_AMB(0); else if (ctx->variant == 1) VXL_AMB(1); else VXL_AMB(2);
#undef VXL_AMB
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

static int local_lights(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame, const void* lights,
                        int n_lights, size_t light_bytes, bool spot, float* out_shadow) {
    if (!ctx || !vol || !view || !frame || n_lights < 0 || (n_lights > 0 && (!lights || !out_shadow))) {
        set_error("vxl_pass_point/spot: bad argument");
        return VXL_ERR_INVALID;
    }
    if (n_lights > VXL_MAX_LIGHTS) { set_error("more than VXL_MAX_LIGHTS lights"); return VXL_ERR_LIMIT; }
    FrameView F;
    if (int e = frame_view(frame, &F)) return e;
    apply_band(ctx, F);
    apply_mirrors(ctx, F);
    if (n_lights == 0 || F.n_tiles == 0) return VXL_OK;
    if (vol->dirty) { if (int e = vxl_volume_build_occupancy(vol)) return e; }
    VXL_CUDA(cudaMemcpyAsync(ctx->d_lights, lights, (size_t)n_lights * light_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const size_t plane = ctx->light_plane_stride ? ctx->light_plane_stride : frame_pixels(frame);
    const CUtensorMap* tm = nullptr;
    if (int e = level_tensor_map(vol->occ[LocalGeom::SHIFT - 2], LocalGeom::TW, LocalGeom::TY, LocalGeom::TY, &tm)) return e;
    // 64x32-pixel regions amortise the tile staging, but a rank's share of a sharded frame may then be a single wave of blocks:
    // below 8 blocks per SM the 32x16 regions balance better (8 GPUs, config 3: reflection 0.29 -> 0.2 ms per rank)
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    const bool big = grid_regions<PassRegion>(F) >= 8u * (unsigned)n_sm;
#define VXL_LL2(SPOT_, MODE_, RG_)                                                                                                        \
    do {                                                                                                                              \
        if (MODE_) VXL_CUDA(cudaFuncSetAttribute(k_local_lights<SPOT_, MODE_, RG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<LocalGeom, (MODE_ > 0)>())); \
        k_local_lights<SPOT_, MODE_, RG_><<<grid_regions<RG_>(F), BLOCK_THREADS, smem_bytes<LocalGeom, (MODE_ > 0)>(), ctx->stream>>>(              \
            vol_view(vol), *tm, F, make_viewk(view), ctx->d_luts, (const float*)ctx->d_lights, n_lights, out_shadow, plane, ctx->d_stats); \
    } while (0)
#define VXL_LL(SPOT_, MODE_) do { if (big) VXL_LL2(SPOT_, MODE_, PassRegion); else VXL_LL2(SPOT_, MODE_, AmbientRegion); } while (0)
    if (spot) { if (ctx->variant == 0) VXL_LL(true, 0); else if (ctx->variant == 1) VXL_LL(true, 1); else VXL_LL(true, 2); }
    else { if (ctx->variant == 0) VXL_LL(false, 0); else if (ctx->variant == 1) VXL_LL(false, 1); else VXL_LL(false, 2); }
#undef VXL_LL
#undef VXL_LL2
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

int vxl_pass_point(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame,
                   const vxl_point_light* lights, int n_lights, float* out_shadow) {
    return local_lights(ctx, vol, view, frame, lights, n_lights, sizeof(vxl_point_light), false, out_shadow);
}

int vxl_pass_spot(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame,
                  const vxl_spot_light* lights, int n_lights, float* out_shadow) {
    return local_lights(ctx, vol, view, frame, lights, n_lights, sizeof(vxl_spot_light), true, out_shadow);
}

int vxl_pass_reflection(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame, float* out_spec_t) {
    if (!ctx || !vol || !view || !frame) { set_error("vxl_pass_reflection: bad argument"); return VXL_ERR_INVALID; }
    if (!out_spec_t) return VXL_OK;
    FrameView F;
    if (int e = frame_view(frame, &F)) return e;
    apply_band(ctx, F);
    apply_mirrors(ctx, F);
    if (!F.material) { set_error("vxl_pass_reflection: frame.material is NULL"); return VXL_ERR_INVALID; }
    if (F.n_tiles == 0) return VXL_OK;
    if (vol->dirty) { if (int e = vxl_volume_build_occupancy(vol)) return e; }
    const CUtensorMap* tm = nullptr;
    if (int e = level_tensor_map(vol->occ[ReflGeom::SHIFT - 2], ReflGeom::TW, ReflGeom::TY, ReflGeom::TY, &tm)) return e;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    const bool big = grid_regions<PassRegion>(F) >= 8u * (unsigned)n_sm;                  // see local_lights
#define VXL_RF2(MODE_, RG_)                                                                                                            \
    do {                                                                                                                            \
        if (MODE_) VXL_CUDA(cudaFuncSetAttribute(k_reflection<MODE_, RG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<ReflGeom, (MODE_ > 0)>())); \
        k_reflection<MODE_, RG_><<<grid_regions<RG_>(F), BLOCK_THREADS, smem_bytes<ReflGeom, (MODE_ > 0)>(), ctx->stream>>>(                       \
            vol_view(vol), *tm, F, make_viewk(view), ctx->d_luts, out_spec_t, ctx->d_stats);                                             \
    } while (0)
#define VXL_RF(MODE_) do { if (big) VXL_RF2(MODE_, PassRegion); else VXL_RF2(MODE_, AmbientRegion); } while (0)
    if (ctx->variant == 0) VXL_RF(0); else if (ctx->variant == 1) VXL_RF(1); else VXL_RF(2);
#undef VXL_RF
#undef VXL_RF2
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

int vxl_trace_rays(vxl_ctx* ctx, vxl_volume* vol, const vxl_ray* rays, int64_t n, int variant, vxl_hit* out) {
    if (!ctx || !vol || n < 0 || (n > 0 && (!rays || !out)) || variant < 0 || variant > 2) { set_error("vxl_trace_rays: bad argument"); return VXL_ERR_INVALID; }
    if (n == 0) return VXL_OK;
    if (vol->dirty) { if (int e = vxl_volume_build_occupancy(vol)) return e; }
    const unsigned grid = (unsigned)((n + 255) / 256);
    k_trace_rays<<<grid, 256, 0, ctx->stream>>>(vol_view(vol), rays, (long long)n, variant, out);
    VXL_LAUNCH_CHECK(ctx);
    return VXL_OK;
}

}  // extern "C"
