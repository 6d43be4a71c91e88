// tests/emul/emul.cu -- TEST INFRASTRUCTURE: host instantiation of the device traversal code.
//
// The tile march (voxelengine_b200/csrc/vxl_bitmarch.cuh) is __host__ __device__; this file compiles
// it for the host (nvcc, no GPU needed) so that its logic -- eligibility, folded tile addressing,
// scan loops, fall-through to the reference test -- can be checked bit-for-bit against the CPU
// oracle in the `-m "not gpu"` suite.  The occupancy levels and tiles are rebuilt here by an
// independent straightforward method, not by the product's kernels.  Nothing in the product loads this.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../voxelengine_b200/csrc/vxl_bitmarch.cuh"

using namespace vxl;

namespace {

struct HostOcc {
    int shift = 0, cx = 0, cy = 0, cz = 0;
    std::vector<uint8_t> occ;    // 0/1 per cell
    bool at(int x, int y, int z) const {
        if (x < 0 || y < 0 || z < 0 || x >= cx || y >= cy || z >= cz) return false;
        return occ[(size_t)x + (size_t)y * cx + (size_t)z * cx * cy] != 0;
    }
};

HostOcc build_occ(const uint8_t* vol, int sx, int sy, int sz, int shift) {
    HostOcc L;
    L.shift = shift;
    const int tpc = 1 << (shift - 1);                        // texels per cell edge
    L.cx = (sx + tpc - 1) / tpc; L.cy = (sy + tpc - 1) / tpc; L.cz = (sz + tpc - 1) / tpc;
    L.occ.assign((size_t)L.cx * L.cy * L.cz, 0);
    for (int z = 0; z < sz; ++z)
        for (int y = 0; y < sy; ++y)
            for (int x = 0; x < sx; ++x)
                if (vol[(size_t)x + (size_t)y * sx + (size_t)z * sx * sy])
                    L.occ[(size_t)(x / tpc) + (size_t)(y / tpc) * L.cx + (size_t)(z / tpc) * L.cx * L.cy] = 1;
    return L;
}

// dilated = true: OR over the 3x3x3 neighbourhood (computed per tile cell straight from the plain level)
template <int TY, int TW>
void build_tile(const HostOcc& L, int ox, int oy, int oz, bool dilated, std::vector<uint32_t>& w) {
    w.assign((size_t)TY * TY * TW, 0);
    for (int z = 0; z < TY; ++z)
        for (int y = 0; y < TY; ++y)
            for (int x = 0; x < TW * 32; ++x) {
                bool b = false;
                if (!dilated) b = L.at(ox + x, oy + y, oz + z);
                else
                    for (int dz = -1; dz <= 1 && !b; ++dz)
                        for (int dy = -1; dy <= 1 && !b; ++dy)
                            for (int dx = -1; dx <= 1 && !b; ++dx) b = L.at(ox + x + dx, oy + y + dy, oz + z + dz);
                if (b) w[(size_t)((x >> 5) * TY + z) * TY + y] |= 1u << (x & 31);      // words ordered [x word][z][y] like the product's tiles
            }
}

struct Emul {
    std::vector<uint8_t> vol;
    int sx, sy, sz;
    HostOcc lv[6];               // index = shift (1..5)
};

template <int SHIFT, int TY, int TW, int DT, int DW, int GH, bool SCAN = false>
void trace_geom(Emul* e, const float* rays, long long n, int variant, const int* center, int fast, int direct, int lockstep, vxl_hit* out,
                unsigned long long* counters, bool near = true, bool pre = false) {
    VolView V;
    V.bytes = e->vol.data(); V.sx = e->sx; V.sy = e->sy; V.sz = e->sz;
    std::vector<uint32_t> w, wd;
    BitTile T;
    // same placement as block_prologue (vxl_passes.cu)
    T.ox = (center[0] >> SHIFT) - TW * 16; T.oy = (center[1] >> SHIFT) - TY / 2; T.oz = (center[2] >> SHIFT) - TY / 2;
    T.dx = T.ox >> 1; T.dy = T.oy >> 1; T.dz = T.oz >> 1;
    build_tile<TY, TW>(e->lv[SHIFT], T.ox, T.oy, T.oz, false, w);
    build_tile<DT, DW>(e->lv[SHIFT + 1], T.dx, T.dy, T.dz, true, wd);
    T.w = w.data(); T.wd = wd.data();
    std::vector<uint32_t> wn;
    T.wn = nullptr; T.nx = (center[0] >> 1) - NEAR_T / 2; T.ny = (center[1] >> 1) - NEAR_T / 2; T.nz = (center[2] >> 1) - NEAR_T / 2;
    if (SCAN) { build_tile<NEAR_T, 1>(e->lv[1], T.nx, T.ny, T.nz, false, wn); T.wn = wn.data(); }
    T.enabled = fast != 0;
    constexpr int TPC = 1 << (SHIFT - 1);
    T.direct = direct != 0 && (V.sx % TPC == 0) && (V.sy % TPC == 0) && (V.sz % TPC == 0);
    T.koff = TileAddr<SHIFT, TY, TW>::texel_koff(V);
    unsigned long long fetched_total = 0, steps_total = 0;
    for (long long i = 0; i < n; ++i) {
        const float* r = rays + i * 8;
        MarchResult M;
        int steps = 0;
        unsigned fetched = 0;
        const float3 o = make_float3(r[0], r[1], r[2]), d = make_float3(r[3], r[4], r[5]);
        bool done = false;
        if constexpr (SCAN) { done = true;
        if (variant == 1 && r[6] == 128.0f) {
            if (near && pre) {          // bundle precheck with the ray's own |dir| as the bound
                const ScanPre P = scan_precheck<SHIFT, TY, TW>(T, o, make_float3(fabsf(d.x), fabsf(d.y), fabsf(d.z)), 129.0f, 20.0f);
                march_scan_super<true, true, true, SHIFT, TY, TW, 23>(V, T, o, d, r[6], steps, &M, fetched, P);
            } else if (near) march_scan_super<true, true, true, SHIFT, TY, TW, 23>(V, T, o, d, r[6], steps, &M, fetched);
            else march_scan_super<true, true, false, SHIFT, TY, TW, 23>(V, T, o, d, r[6], steps, &M, fetched);
        } else done = false; }
        if (done) {
        } else if (lockstep) {
            if (variant == 0) march_bits<false, true, true, true, SHIFT, TY, TW, DT, DW, GH>(V, T, o, d, r[6], steps, &M, fetched);
            else march_bits<true, true, true, true, SHIFT, TY, TW, DT, DW, GH>(V, T, o, d, r[6], steps, &M, fetched);
        } else {
            if (variant == 0) march_bits<false, true, false, true, SHIFT, TY, TW, DT, DW, GH>(V, T, o, d, r[6], steps, &M, fetched);
            else march_bits<true, true, false, true, SHIFT, TY, TW, DT, DW, GH>(V, T, o, d, r[6], steps, &M, fetched);
        }
        vxl_hit hh;
        memset(&hh, 0, sizeof hh);
        hh.t = M.d; hh.steps = M.steps; hh.status = M.status; hh.vx = M.vx; hh.vy = M.vy; hh.vz = M.vz;
        hh.px = M.pos.x; hh.py = M.pos.y; hh.pz = M.pos.z;
        out[i] = hh;
        fetched_total += fetched; steps_total += (unsigned long long)steps;
    }
    if (counters) { counters[0] = fetched_total; counters[1] = steps_total; }
}

}  // namespace

extern "C" {

void* emul_create(const uint8_t* vol, int sx, int sy, int sz) {
    Emul* e = new Emul();
    e->vol.assign(vol, vol + (size_t)sx * sy * sz);
    e->sx = sx; e->sy = sy; e->sz = sz;
    for (int sh = 1; sh <= 5; ++sh) e->lv[sh] = build_occ(vol, sx, sy, sz, sh);
    return e;
}
void emul_destroy(void* h) { delete (Emul*)h; }

// occupancy level as 0/1 bytes [cz][cy][cx] for comparison with vxl_volume_debug_occupancy / numpy
void emul_level(void* h, int shift, uint8_t* out, int* dims) {
    Emul* e = (Emul*)h;
    const HostOcc& L = e->lv[shift];
    dims[0] = L.cx; dims[1] = L.cy; dims[2] = L.cz;
    if (out) memcpy(out, L.occ.data(), L.occ.size());
}

// rays: 8 floats each (origin, dir, dist, pad); variant 0 Sparse / 1 SuperSparse; geom 0 ambient / local lights,
// 1 the same without probe groups, 2 reflection (the tile geometries of vxl_passes.cu); the tile is placed around `center` (voxels) exactly as
// block_prologue does.  fast = 0 runs the plain march; direct = 0 forces the bounds-checked fetch.  counters: probes that read the volume, total probes.
void emul_trace(void* h, const float* rays, long long n, int variant, const int* center, int fast, int geom, int direct, int lockstep,
                vxl_hit* out, unsigned long long* counters) {
    Emul* e = (Emul*)h;
    // the geometries of vxl_passes.cu (AmbientGeom == LocalGeom, ReflGeom) and a GH = 0 twin without probe groups
    if (geom == 0) trace_geom<2, 76, 3, 39, 2, 7>(e, rays, n, variant, center, fast, direct, lockstep, out, counters);
    else if (geom == 3) trace_geom<2, 76, 3, 39, 2, 7, true>(e, rays, n, variant, center, fast, direct, lockstep, out, counters);   // AO rays (SuperSparse, dist 128) by scan + resolve, near tile
    else if (geom == 4) trace_geom<2, 76, 3, 39, 2, 7, true>(e, rays, n, variant, center, fast, direct, lockstep, out, counters, false);
    else if (geom == 5) trace_geom<2, 76, 3, 39, 2, 7, true>(e, rays, n, variant, center, fast, direct, lockstep, out, counters, true, true);   // with the per-bundle precheck   // the same without the near tile
    else if (geom == 1) trace_geom<2, 68, 3, 35, 2, 0>(e, rays, n, variant, center, fast, direct, lockstep, out, counters);
    else trace_geom<3, 68, 3, 35, 2, 10>(e, rays, n, variant, center, fast, direct, lockstep, out, counters);
}

// experiment helper: classify every probe of the plain march by what a bit-occupancy hierarchy would know.
// out[phase(2)][near(2)][8]: [0]=probes, [1]=texel byte zero, [2]=4-voxel cell empty, [3]=8-voxel cell empty, [4]=16-voxel cell empty
void emul_classify(void* h, const float* rays, long long n, int variant, const int* center, float near_half, unsigned long long* out) {
    Emul* e = (Emul*)h;
    memset(out, 0, sizeof(unsigned long long) * 2 * 2 * 8);
    const float step0 = variant ? 2.5f : 0.5f;
    auto occ = [&](int shift, float3 p) -> bool {
        const float c = (float)(1 << shift);
        return e->lv[shift].at((int)floorf(p.x / c), (int)floorf(p.y / c), (int)floorf(p.z / c));
    };
    VolView V; V.bytes = e->vol.data(); V.sx = e->sx; V.sy = e->sy; V.sz = e->sz;
    for (long long i = 0; i < n; ++i) {
        const float* r = rays + i * 8;
        float3 pos = make_float3(r[0], r[1], r[2]);
        float3 sd = make_float3(r[3], r[4], r[5]) * step0;
        float d = step0, sf = step0;
        auto tally = [&](int phase) {
            const bool nearp = fabsf(pos.x - center[0]) < near_half && fabsf(pos.y - center[1]) < near_half && fabsf(pos.z - center[2]) < near_half;
            unsigned long long* o = out + (phase * 2 + (nearp ? 1 : 0)) * 8;
            o[0]++;
            for (int sh = 1; sh <= 4; ++sh) if (!occ(sh, pos)) o[sh]++;
        };
        bool hit = false;
        while (d < 16.0f) {
            tally(0);
            const int tx = f2i(pos.x / 2.0f), ty = f2i(pos.y / 2.0f), tz = f2i(pos.z / 2.0f);
            const unsigned v = fetch_texel(V, tx, ty, tz);
            unsigned bit = 0u;
            bit += gmod(pos.x, 0.5f) > 0.25f ? 1u : 0u;
            bit += gmod(pos.y, 0.5f) > 0.25f ? 2u : 0u;
            bit += gmod(pos.z, 0.5f) > 0.25f ? 4u : 0u;
            if ((v >> bit) & 1u) { hit = true; break; }
            pos = pos + sd; d += sf;
        }
        if (hit) continue;
        sf *= 2.0f; sd = sd * 2.0f;
        const float lim = fminf(r[6], 164.0f);
        while (d < lim) {
            tally(1);
            if (fetch_texel(V, f2i(pos.x) / 2, f2i(pos.y) / 2, f2i(pos.z) / 2) != 0u) break;
            pos = pos + sd; d += sf;
        }
    }
}

}  // extern "C"
