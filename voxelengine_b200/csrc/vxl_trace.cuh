// vxl_trace.cuh -- device traversal of the packed world occupancy volume.
//
// Re-implements, for sm_100a, the three tracers of Sources/Shaders/lib/Light.frag:
//   getVolumeAt                     :14-27
//   raycastShadowVolume (DDA)       :29-81
//   raycastShadowVolumeSparse       :131-173   (step0 = 0.5)
//   raycastShadowVolumeSuperSparse  :175-217   (step0 = 2.5)
// Results are bit-identical to the reference arithmetic as pinned in vxl_math.cuh; out-of-range
// texel fetches return 0 and integer /2 truncates toward zero.
#pragma once
#include "vxl_internal.h"
#include "vxl_math.cuh"

namespace vxl {

// texelFetch(SHADOW_VOX_TEXTURE, (x,y,z), 0).r with out-of-range -> 0
VXL_DI unsigned fetch_texel(const VolView& V, int x, int y, int z) {
    if ((unsigned)x >= (unsigned)V.sx || (unsigned)y >= (unsigned)V.sy || (unsigned)z >= (unsigned)V.sz) return 0u;
    size_t off = (size_t)x + (size_t)y * (size_t)V.sx + (size_t)z * ((size_t)V.sx * (size_t)V.sy);
    return (unsigned)ldg(V.bytes + off);
}

// Light.frag:14-27
VXL_DI bool get_volume_at(const VolView& V, int x, int y, int z, int mip) {
    int bit = (x & 1) | ((y & 1) << 1) | ((z & 1) << 2);
    unsigned v = fetch_texel(V, x / 2, y / 2, z / 2);
    if (mip % 2 == 1) return v != 0u;
    return (v & (1u << bit)) != 0u;
}

struct MarchResult {
    float d;        // returned distance (== dist on a miss)
    int steps;      // occupancy probes performed
    int status;     // 0 miss, 1 fine-bit hit (phase 1), 2 coarse-byte hit (phase 2)
    int vx, vy, vz; // probed voxel on a hit
    float3 pos;     // pos at the hit probe
};

// Two-phase fixed-step march (Light.frag:131-173 / :175-217). RECORD fills the hit record fields.
template <bool RECORD>
VXL_DI float march(const VolView& V, float3 origin, float3 dir, float dist, float step0, int& steps_out, MarchResult* rec) {
    float stepFactor = step0;
    float3 stepDir = dir * stepFactor;
    float3 pos = origin;
    float d = stepFactor;
    int steps = 0;

    while (d < 16.0f) {
        int tx = f2i(pos.x / 2.0f), ty = f2i(pos.y / 2.0f), tz = f2i(pos.z / 2.0f);
        unsigned v = fetch_texel(V, tx, ty, tz);
        unsigned bit = 0u;
        bit += gmod(pos.x, 0.5f) > 0.25f ? 1u : 0u;
        bit += gmod(pos.y, 0.5f) > 0.25f ? 2u : 0u;
        bit += gmod(pos.z, 0.5f) > 0.25f ? 4u : 0u;
        ++steps;
        if ((v >> bit) & 1u) {
            steps_out += steps;
            if (RECORD) {
                rec->d = d; rec->steps = steps; rec->status = 1;
                rec->vx = tx * 2 + (int)(bit & 1u); rec->vy = ty * 2 + (int)((bit >> 1) & 1u); rec->vz = tz * 2 + (int)((bit >> 2) & 1u);
                rec->pos = pos;
            }
            return d;
        }
        pos = pos + stepDir;
        d += stepFactor;
    }

    stepFactor *= 2.0f;
    stepDir = stepDir * 2.0f;
    const float lod1MaxT = fminf(dist, 164.0f);
    while (d < lod1MaxT) {
        int px = f2i(pos.x), py = f2i(pos.y), pz = f2i(pos.z);
        ++steps;
        if (fetch_texel(V, px / 2, py / 2, pz / 2) != 0u) {   // getVolumeAt(ivec3(pos), 1)
            steps_out += steps;
            if (RECORD) {
                rec->d = d; rec->steps = steps; rec->status = 2;
                rec->vx = px; rec->vy = py; rec->vz = pz;
                rec->pos = pos;
            }
            return d;
        }
        pos = pos + stepDir;
        d += stepFactor;
    }
    steps_out += steps;
    if (RECORD) { rec->d = dist; rec->steps = steps; rec->status = 0; rec->vx = rec->vy = rec->vz = 0; rec->pos = make_float3(0.f, 0.f, 0.f); }
    return dist;
}

struct DdaResult {
    float t; int nt; int status; int vx, vy, vz; float3 hit; float3 normal; int probes;
};

// Light.frag:29-81. mip is 0 throughout (it can only decrement from 0 on a hit at mip > 0), so
// mipSize == 1 and the outer do/while replays the identical walk from the identical origin up to
// four times; the walk is deterministic, so it is executed once and its counters are scaled.
VXL_DI bool dda(const VolView& V, float3 origin, float3 direction, float maxt, DdaResult& R) {
    const float3 stepSign = make_float3(gsign(direction.x), gsign(direction.y), gsign(direction.z));
    const float3 t_delta = make_float3(1.0f, 1.0f, 1.0f) / (direction * stepSign);
    const int hx = V.sx * 2 + 1, hy = V.sy * 2 + 1, hz = V.sz * 2 + 1;
    int cx = f2i(floorf(origin.x)), cy = f2i(floorf(origin.y)), cz = f2i(floorf(origin.z));
    float3 next_bounds = make_float3((float)cx, (float)cy, (float)cz) + (stepSign * 0.5f + make_float3(0.5f, 0.5f, 0.5f));
    float3 t_max = (next_bounds - origin) / direction;
    int n = 0, nt = 0, probes = 0;
    float totalt = 0.0f, best_t = 0.0f;
    R.hit = make_float3(0.f, 0.f, 0.f); R.normal = make_float3(0.f, 0.f, 0.f);
    do {
        float3 select = make_float3(gstep(t_max.x, t_max.z) * gstep(t_max.x, t_max.y),
                                    gstep(t_max.y, t_max.x) * gstep(t_max.y, t_max.z),
                                    gstep(t_max.z, t_max.y) * gstep(t_max.z, t_max.x));
        if (cx < 0 || cy < 0 || cz < 0 || cx > hx || cy > hy || cz > hz) {
            R.t = best_t; R.nt = nt; R.status = 3; R.vx = cx; R.vy = cy; R.vz = cz; R.probes = probes;
            return false;
        }
        bool voxel = get_volume_at(V, cx, cy, cz, 0);
        ++probes;
        best_t = dot3(t_max, select);
        if (voxel) {
            R.normal = (stepSign * -1.0f) * select;
            R.hit = (origin + direction * best_t) * 1.0f;
            R.t = best_t; R.nt = nt; R.status = 1; R.vx = cx; R.vy = cy; R.vz = cz; R.probes = probes;
            return true;
        }
        float3 adv = select * stepSign;
        cx += f2i(adv.x); cy += f2i(adv.y); cz += f2i(adv.z);
        t_max = t_max + t_delta * select;
        totalt = best_t;
        nt++;
    } while (++n < 256 && totalt < maxt);
    // three more identical replays of the same miss (outer do/while, ++i < 4)
    R.t = best_t; R.nt = nt * 4; R.status = 0; R.vx = R.vy = R.vz = 0; R.probes = probes * 4;
    return false;
}

}  // namespace vxl
