// tests/emul/emul.cu -- TEST INFRASTRUCTURE: host instantiation of the device traversal code.
//
// The accelerated march (voxelengine_b200/csrc/vxl_fastmarch.cuh) is __host__ __device__; this file
// compiles it for the host (nvcc, no GPU needed) so that its logic -- clearance lookups, skip
// counts, exact replay, eligibility -- can be checked bit-for-bit against the CPU oracle in the
// `-m "not gpu"` suite.  The clearance maps and tiles are rebuilt here by an independent brute-force
// method (iterated 3x3x3 dilation), not by the product's kernels.  Nothing in the product loads this.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../voxelengine_b200/csrc/vxl_fastmarch.cuh"

using namespace vxl;

namespace {

struct HostLevel {
    int shift, cap, border;
    int cx, cy, cz;              // padded dims
    std::vector<uint8_t> r;      // clearance per padded cell
};

// brute force: base occupancy then `cap` rounds of 3x3x3 dilation
HostLevel build_level(const uint8_t* vol, int sx, int sy, int sz, int shift, int cap) {
    HostLevel L;
    L.shift = shift; L.cap = cap; L.border = cap;
    const int tpc = 1 << (shift - 1);                        // texels per cell edge
    const int nx = (sx + tpc - 1) / tpc, ny = (sy + tpc - 1) / tpc, nz = (sz + tpc - 1) / tpc;
    L.cx = nx + 2 * cap; L.cy = ny + 2 * cap; L.cz = nz + 2 * cap;
    const size_t n = (size_t)L.cx * L.cy * L.cz;
    std::vector<uint8_t> occ(n, 0);
    for (int z = 0; z < sz; ++z)
        for (int y = 0; y < sy; ++y)
            for (int x = 0; x < sx; ++x)
                if (vol[(size_t)x + (size_t)y * sx + (size_t)z * sx * sy])
                    occ[(size_t)(x / tpc + cap) + (size_t)(y / tpc + cap) * L.cx + (size_t)(z / tpc + cap) * L.cx * L.cy] = 1;
    L.r.assign(n, (uint8_t)cap);
    std::vector<uint8_t> cur = occ, nxt(n);
    for (size_t i = 0; i < n; ++i) if (occ[i]) L.r[i] = 0;
    for (int round = 1; round < cap; ++round) {
        for (int z = 0; z < L.cz; ++z)
            for (int y = 0; y < L.cy; ++y)
                for (int x = 0; x < L.cx; ++x) {
                    uint8_t v = 0;
                    for (int dz = -1; dz <= 1 && !v; ++dz)
                        for (int dy = -1; dy <= 1 && !v; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int ax = x + dx, ay = y + dy, az = z + dz;
                                if (ax < 0 || ay < 0 || az < 0 || ax >= L.cx || ay >= L.cy || az >= L.cz) continue;
                                if (cur[(size_t)ax + (size_t)ay * L.cx + (size_t)az * L.cx * L.cy]) { v = 1; break; }
                            }
                    nxt[(size_t)x + (size_t)y * L.cx + (size_t)z * L.cx * L.cy] = v;
                }
        for (size_t i = 0; i < n; ++i) if (nxt[i] && !cur[i]) L.r[i] = (uint8_t)round;
        cur.swap(nxt);
    }
    return L;
}

void build_tile(const HostLevel& L, int ox, int oy, int oz, std::vector<uint32_t>& w) {
    w.assign(CT_WORDS, 0);
    const int fill = std::min(15, L.border + 1);
    for (int z = 0; z < CT; ++z)
        for (int y = 0; y < CT; ++y)
            for (int x = 0; x < CT; ++x) {
                const int ax = ox + x + L.border, ay = oy + y + L.border, az = oz + z + L.border;
                int r = fill;
                if (ax >= 0 && ay >= 0 && az >= 0 && ax < L.cx && ay < L.cy && az < L.cz) r = L.r[(size_t)ax + (size_t)ay * L.cx + (size_t)az * L.cx * L.cy];
                w[(z * CT + y) * CTW + (x >> 3)] |= (uint32_t)r << (4 * (x & 7));
            }
}

struct Emul {
    std::vector<uint8_t> vol;
    int sx, sy, sz;
    HostLevel l2, l4;
};

}  // namespace

extern "C" {

void* emul_create(const uint8_t* vol, int sx, int sy, int sz) {
    Emul* e = new Emul();
    e->vol.assign(vol, vol + (size_t)sx * sy * sz);
    e->sx = sx; e->sy = sy; e->sz = sz;
    e->l2 = build_level(vol, sx, sy, sz, 2, 8);
    e->l4 = build_level(vol, sx, sy, sz, 4, 15);
    return e;
}
void emul_destroy(void* h) { delete (Emul*)h; }

// clearance arrays (padded) for comparison with vxl_volume_debug_clearance / numpy
void emul_level(void* h, int level, uint8_t* out, int* dims) {
    Emul* e = (Emul*)h;
    const HostLevel& L = level == 2 ? e->l2 : e->l4;
    dims[0] = L.cx; dims[1] = L.cy; dims[2] = L.cz; dims[3] = L.border;
    if (out) memcpy(out, L.r.data(), L.r.size());
}

// rays: 8 floats each (origin, dir, dist, pad); variant 0 Sparse / 1 SuperSparse; the clearance tiles are
// placed around `center` (voxels) exactly as block_prologue does.  fast = 0 runs the plain march.
void emul_trace(void* h, const float* rays, long long n, int variant, const int* center, int fast, vxl_hit* out,
                unsigned long long* exact_total, unsigned long long* steps_total) {
    Emul* e = (Emul*)h;
    VolView V;
    V.bytes = e->vol.data(); V.sx = e->sx; V.sy = e->sy; V.sz = e->sz;
    V.cm4 = ClearView{nullptr, 0, 0, 0, 0, 0}; V.cm16 = V.cm4;
    std::vector<uint32_t> w4, w16;
    FastCtx C;
    C.t4.ox = (center[0] >> 2) - CT / 2; C.t4.oy = (center[1] >> 2) - CT / 2; C.t4.oz = (center[2] >> 2) - CT / 2;
    C.t16.ox = (center[0] >> 4) - CT / 2; C.t16.oy = (center[1] >> 4) - CT / 2; C.t16.oz = (center[2] >> 4) - CT / 2;
    build_tile(e->l2, C.t4.ox, C.t4.oy, C.t4.oz, w4);
    build_tile(e->l4, C.t16.ox, C.t16.oy, C.t16.oz, w16);
    C.t4.w = w4.data(); C.t16.w = w16.data();
    C.enabled = fast != 0;
    unsigned long long ex = 0, st = 0;
    for (long long i = 0; i < n; ++i) {
        const float* r = rays + i * 8;
        MarchResult M;
        int steps = 0;
        unsigned exact = 0;
        const float3 o = make_float3(r[0], r[1], r[2]), d = make_float3(r[3], r[4], r[5]);
        if (variant == 0) march_fast<false, true>(V, C, o, d, r[6], steps, &M, exact);
        else march_fast<true, true>(V, C, o, d, r[6], steps, &M, exact);
        vxl_hit hh;
        memset(&hh, 0, sizeof hh);
        hh.t = M.d; hh.steps = M.steps; hh.status = M.status; hh.vx = M.vx; hh.vy = M.vy; hh.vz = M.vz;
        hh.px = M.pos.x; hh.py = M.pos.y; hh.pz = M.pos.z;
        out[i] = hh;
        ex += exact; st += (unsigned long long)steps;
    }
    if (exact_total) *exact_total = ex;
    if (steps_total) *steps_total = st;
}

}  // extern "C"
