// oracle/vxo.cpp -- CPU ORACLE (test infrastructure only; see the header of vxo.h).
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp   (oracle/Makefile)
// Every `a*b+c` below is two roundings -- no FMA contraction -- on purpose: the CUDA product is
// compiled with --fmad=false and must agree with this file bit-for-bit.
//
// Pinned choices for GLSL-undefined behaviour:
//   * out-of-range texelFetch            -> 0
//   * ivec3(float)                       -> truncation toward zero, saturating to INT_MIN/MAX,
//                                           NaN -> 0 (the semantics of PTX cvt.rzi.s32.f32)
//   * integer / 2 on negatives           -> C truncation
//   * min(NaN, x)                        -> x (fminf)
//   * sign(0) = 0, 1/0 = inf, 0*inf = NaN propagate per IEEE-754
// Operation orders follow glm 0.9.9.9:
//   dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z         Vendor/glm/detail/func_geometric.inl:48-55
//   normalize(v) = v * (1/sqrt(dot(v,v)))            func_geometric.inl:82-90
//   mix(x,y,a) = x*(1-a) + y*a                       func_common.inl:81-89
//   mod(x,y) = x - y*floor(x/y)                      func_common.inl:212-219
//   M*v = (M0*v0 + M1*v1) + (M2*v2 + M3*v3)          type_mat4x4.inl:561-572
#include "vxo.h"

#include <cmath>
#include <cstring>
#include <climits>
#include <random>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct I3 { int x, y, z; };

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }

inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 normalize3(V3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return v * inv; }
inline float length3(V3 v) { return sqrtf(dot3(v, v)); }
inline V3 mix3(V3 x, V3 y, float a) { float ia = 1.0f - a; return x * ia + y * a; }
inline V3 cross3(V3 a, V3 b) {  // func_geometric.inl:61-72
    return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline float gmod(float x, float y) { return x - y * floorf(x / y); }
inline float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float gstep(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float gclamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float gsmoothstep(float e0, float e1, float x) {
    float t = gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// cvt.rzi.s32.f32 semantics
inline int f2i(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT_MAX;
    if (x <= -2147483648.0f) return INT_MIN;
    return (int)x;
}
// column-major mat4 (glm) times vec4
inline V4 mat_mul(const float* m, V4 v) {
    V4 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * v.w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * v.w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w);
    r.w = (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w);
    return r;
}
inline V3 xyz(V4 v) { return V3{v.x, v.y, v.z}; }

const float NEAR_ = 0.1f;      // Sources/Shaders/lib/Common.frag:12
const float FAR_ = 4096.0f;    // Common.frag:13
const float GOLDEN_RATIO = 2.118033988749895f;  // Common.frag:9 (sic)

// ---------------------------------------------------------------------------------------------
// A1  texelFetch(SHADOW_VOX_TEXTURE, p, 0).r  with out-of-range -> 0
// ---------------------------------------------------------------------------------------------
inline unsigned fetch(const vxo_volume& v, int x, int y, int z) {
    if ((unsigned)x >= (unsigned)v.sx || (unsigned)y >= (unsigned)v.sy || (unsigned)z >= (unsigned)v.sz) return 0u;
    return v.data[(size_t)x + (size_t)y * (size_t)v.sx + (size_t)z * (size_t)v.sx * (size_t)v.sy];
}

// Sources/Shaders/lib/Light.frag:14-27  getVolumeAt(pos, mip)   (image has one level: realMip==0)
inline bool get_volume_at(const vxo_volume& v, I3 p, int mip) {
    int bit = (p.x & 1) | ((p.y & 1) << 1) | ((p.z & 1) << 2);
    int mask = 1 << bit;
    p.x /= 2; p.y /= 2; p.z /= 2;
    unsigned voxel = fetch(v, p.x, p.y, p.z);
    if (mip % 2 == 1) return voxel != 0u;
    return (voxel & (unsigned)mask) != 0u;
}

// ---------------------------------------------------------------------------------------------
// A2/A3  Light.frag:131-173 raycastShadowVolumeSparse (step0 = 0.5)
//        Light.frag:175-217 raycastShadowVolumeSuperSparse (step0 = 2.5)
// The two functions differ only in the initial stepFactor.
// ---------------------------------------------------------------------------------------------
float march(const vxo_volume& vol, V3 origin, V3 dir, float dist, float step0, uint64_t& nsteps, vxo_hit* rec) {
    float stepFactor = step0;
    V3 stepDir = dir * stepFactor;
    V3 pos = origin;
    float d = stepFactor;
    int steps = 0;

    while (d < 16.0f) {
        I3 t{f2i(pos.x / 2.0f), f2i(pos.y / 2.0f), f2i(pos.z / 2.0f)};
        unsigned v = fetch(vol, t.x, t.y, t.z);
        unsigned bit = 0u;
        bit += gmod(pos.x, 0.5f) > 0.25f ? 1u : 0u;
        bit += gmod(pos.y, 0.5f) > 0.25f ? 2u : 0u;
        bit += gmod(pos.z, 0.5f) > 0.25f ? 4u : 0u;
        unsigned mask = 1u << bit;
        ++steps;
        if ((mask & v) != 0u) {
            nsteps += (uint64_t)steps;
            if (rec) {
                rec->t = d; rec->steps = steps; rec->status = 1;
                rec->vx = t.x * 2 + (int)(bit & 1u); rec->vy = t.y * 2 + (int)((bit >> 1) & 1u); rec->vz = t.z * 2 + (int)((bit >> 2) & 1u);
                rec->px = pos.x; rec->py = pos.y; rec->pz = pos.z;
            }
            return d;
        }
        pos = pos + stepDir;
        d += stepFactor;
    }

    stepFactor *= 2.0f;
    stepDir = stepDir * 2.0f;
    float lod1MaxT = fminf(dist, 164.0f);
    while (d < lod1MaxT) {
        I3 p{f2i(pos.x), f2i(pos.y), f2i(pos.z)};
        ++steps;
        if (get_volume_at(vol, p, 1)) {
            nsteps += (uint64_t)steps;
            if (rec) {
                rec->t = d; rec->steps = steps; rec->status = 2;
                rec->vx = p.x; rec->vy = p.y; rec->vz = p.z;
                rec->px = pos.x; rec->py = pos.y; rec->pz = pos.z;
            }
            return d;
        }
        pos = pos + stepDir;
        d += stepFactor;
    }
    nsteps += (uint64_t)steps;
    if (rec) { rec->t = dist; rec->steps = steps; rec->status = 0; }
    return dist;
}

// ---------------------------------------------------------------------------------------------
// A4  Light.frag:29-81 raycastShadowVolume -- Amanatides-Woo DDA at mip 0.
// `mip` starts at 0 and can only decrement on a hit at mip>0, so mipSize == 1 throughout and the
// outer do/while restarts the same walk from the same origin up to 4 times.
// ---------------------------------------------------------------------------------------------
bool dda(const vxo_volume& vol, V3 origin, V3 direction, float maxt, vxo_hit* rec, uint64_t& nsteps) {
    I3 dim{vol.sx, vol.sy, vol.sz};
    V3 stepSign{gsign(direction.x), gsign(direction.y), gsign(direction.z)};
    V3 t_delta = v3(1.0f, 1.0f, 1.0f) / (direction * stepSign);
    int mip = 0;
    int i = 0, nt = 0, probes = 0;
    float totalt = 0.0f;
    float best_t = 0.0f;
    do {
        float mipSize = (float)(1 << mip);
        origin = origin / v3(mipSize, mipSize, mipSize);
        I3 cur{f2i(floorf(origin.x)), f2i(floorf(origin.y)), f2i(floorf(origin.z))};
        V3 next_bounds = v3((float)cur.x, (float)cur.y, (float)cur.z) + (stepSign * 0.5f + v3(0.5f, 0.5f, 0.5f));
        V3 t_max = (next_bounds - origin) / direction;
        int n = 0;
        do {
            V3 select{gstep(t_max.x, t_max.z) * gstep(t_max.x, t_max.y),
                      gstep(t_max.y, t_max.x) * gstep(t_max.y, t_max.z),
                      gstep(t_max.z, t_max.y) * gstep(t_max.z, t_max.x)};
            // clamp(current_voxel, ivec3(0), ivec3(dim/mipSize)*2+1) != current_voxel -> miss
            I3 hi{f2i((float)dim.x / mipSize) * 2 + 1, f2i((float)dim.y / mipSize) * 2 + 1, f2i((float)dim.z / mipSize) * 2 + 1};
            if (cur.x < 0 || cur.y < 0 || cur.z < 0 || cur.x > hi.x || cur.y > hi.y || cur.z > hi.z) {
                nsteps += (uint64_t)probes;
                if (rec) { rec->t = best_t; rec->steps = nt; rec->status = 3; rec->vx = cur.x; rec->vy = cur.y; rec->vz = cur.z; }
                return false;
            }
            bool voxel = get_volume_at(vol, cur, mip);
            ++probes;
            best_t = dot3(t_max, select);
            if (voxel) {
                // mip == 0 always
                nsteps += (uint64_t)probes;
                if (rec) {
                    V3 nrm = (stepSign * -1.0f) * select;
                    V3 hit = (origin + direction * best_t) * mipSize;
                    rec->t = best_t; rec->steps = nt; rec->status = 1;
                    rec->vx = cur.x; rec->vy = cur.y; rec->vz = cur.z;
                    rec->px = hit.x; rec->py = hit.y; rec->pz = hit.z;
                    rec->nx = nrm.x; rec->ny = nrm.y; rec->nz = nrm.z;
                }
                return true;
            }
            V3 adv = select * stepSign;
            cur.x += f2i(adv.x); cur.y += f2i(adv.y); cur.z += f2i(adv.z);
            t_max = t_max + t_delta * select;
            totalt = best_t;
            nt++;
        } while (++n < 256 && totalt < maxt);
        origin = origin * mipSize;
    } while (++i < 4);
    nsteps += (uint64_t)probes;
    if (rec) { rec->t = best_t; rec->steps = nt; rec->status = 0; }
    return false;
}

// ---------------------------------------------------------------------------------------------
// G-buffer decoders (Vulkan fixed-point conversion rules; formats from Sources/Graphics/Graphics.h:53-60)
// ---------------------------------------------------------------------------------------------
inline float unorm24(uint32_t d) { return (float)(d & 0xFFFFFFu) / 16777215.0f; }
inline float unorm8(uint32_t c) { return (float)(c & 0xFFu) / 255.0f; }
inline float snorm8(uint32_t c) { return fmaxf((float)(int8_t)(c & 0xFFu) / 127.0f, -1.0f); }

struct Luts { float cosT[256], sinT[256]; };
// SURVEY hard part 1: cos/sin of theta = 6.283*v (v = k/255) evaluated in double, rounded once.
const Luts& luts() {
    static Luts L;
    static bool init = false;
    if (!init) {
        for (int k = 0; k < 256; ++k) {
            float v = (float)k / 255.0f;
            float theta = 6.283f * v;
            L.cosT[k] = (float)cos((double)theta);
            L.sinT[k] = (float)sin((double)theta);
        }
        init = true;
    }
    return L;
}

// LightAmbient.frag:81-87 cosineSampleHemisphere(rand) with rand = noise.xy (bytes nx, ny)
inline V3 cosine_sample_hemisphere(const Luts& L, uint32_t nx, uint32_t ny) {
    float u = unorm8(nx);
    float r = sqrtf(u);
    float x = r * L.cosT[ny & 0xFFu];
    float y = r * L.sinT[ny & 0xFFu];
    return V3{x, y, sqrtf(fmaxf(0.0f, 1.0f - u))};
}

struct Pixel {
    V3 farvec;   // LightAmbient.vert:32-36 evaluated per pixel
    float u, v;  // In.UV
};

inline Pixel pixel_setup(const vxo_view& view, int W, int H, int px, int py) {
    Pixel p;
    p.u = ((float)px + 0.5f) / (float)W;
    p.v = ((float)py + 0.5f) / (float)H;
    float ndcx = 2.0f * p.u - 1.0f;
    float ndcy = 1.0f - 2.0f * p.v;
    V4 f = mat_mul(view.InverseProjectionMatrix, V4{ndcx, ndcy, 1.0f, 1.0f});
    p.farvec = V3{f.x / f.w, f.y / f.w, f.z / f.w};
    return p;
}

// LightAmbient.frag:44-52 getNoise() / getNoise(int s); s < 0 selects the argument-less overload.
inline uint32_t get_noise(const vxo_gbuffer& gb, const vxo_view& view, const Pixel& p, int s) {
    float fx, fy;
    if (s < 0) {
        fx = GOLDEN_RATIO * gmod((float)view.Frame, 16.0f);
        fy = GOLDEN_RATIO * gmod((float)(view.Frame + 1), 16.0f);
    } else {
        fx = GOLDEN_RATIO * gmod((float)(view.Frame + s * 5), 64.0f);
        fy = GOLDEN_RATIO * gmod((float)(view.Frame + s * 7 + 1), 64.0f);
    }
    float resx = (float)gb.width, resy = (float)gb.height;
    int cx = f2i((p.u + fx) * resx) % 512;
    int cy = f2i((p.v + fy) * resy) % 512;
    return gb.noise[cy * 512 + cx];
}

inline V3 decode_normal(uint32_t n) { return V3{snorm8(n), snorm8(n >> 8), snorm8(n >> 16)}; }

inline V3 sun_dir() { return normalize3(v3(0.3f, 0.4f, 0.5f)); }  // LightAmbient.frag:15

inline int nthreads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

struct Acc { uint64_t rays = 0, steps = 0, pixels = 0; };

}  // namespace

extern "C" {

int vxo_num_threads(void) { return nthreads(); }
void vxo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void vxo_trace_rays(const vxo_volume* vol, const vxo_ray* rays, int64_t n, int variant, vxo_hit* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const vxo_ray& r = rays[i];
        vxo_hit h;
        memset(&h, 0, sizeof h);
        uint64_t s = 0;
        V3 o{r.ox, r.oy, r.oz}, d{r.dx, r.dy, r.dz};
        if (variant == VXO_SPARSE) march(*vol, o, d, r.dist, 0.5f, s, &h);
        else if (variant == VXO_SUPERSPARSE) march(*vol, o, d, r.dist, 2.5f, s, &h);
        else dda(*vol, o, d, r.dist, &h, s);
        out[i] = h;
    }
}

// ---------------------------------------------------------------------------------------------
// A5  LightAmbient.frag:134-175 (main, shadow block) + :111-126 calculateAmbientIrradiance.
// Outputs of the new pass contract: shadow in {0,1}, ao = mean_i((d_i/128)^2) * 0.05.
// Sky pixels (depth >= 0.999) generate no rays: shadow = 1, ao = 0.
// n_ao > 1 is the build's multi-sample extension: ray 0 uses getNoise(), ray i >= 1
// uses getNoise(i).
// ---------------------------------------------------------------------------------------------
void vxo_pass_ambient(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb, int n_ao,
                      vxo_rows rows, float* out_shadow, float* out_ao, vxo_stats* stats) {
    const Luts& L = luts();
    const int W = gb->width, H = gb->height;
    const V3 SUN = sun_dir();
    uint64_t trays = 0, tsteps = 0, tpix = 0;
    if (rows.step < 1) rows.step = 1;
    const int nrows = rows.end > rows.begin ? (rows.end - rows.begin + rows.step - 1) / rows.step : 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : trays, tsteps, tpix)
    for (int ri = 0; ri < nrows; ++ri) {
        const int py = rows.begin + ri * rows.step;
        if (py < 0 || py >= H) continue;
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            float depth = unorm24(gb->depth24[idx]);
            if (!(depth < 0.999f)) { out_shadow[idx] = 1.0f; out_ao[idx] = 0.0f; continue; }
            Pixel p = pixel_setup(*view, W, H, px, py);
            V3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));       // :141
            V3 normal = decode_normal(gb->normal[idx]);               // :142
            V3 wd = SUN;                                              // :149
            V3 wcp = xyz(mat_mul(view->InverseViewMatrix, V4{pos.x, pos.y, pos.z, 1.0f})) * 10.0f;  // :150
            uint32_t n = get_noise(*gb, *view, p, -1);
            V3 randomVec = cosine_sample_hemisphere(L, n, n >> 8) * 0.1f;   // :151
            randomVec.z *= gsign(unorm8(n >> 16) - 0.5f);                  // :152
            wd = mix3(wd, randomVec, 0.5f);                                // :153
            wd = normalize3(wd);                                           // :154
            wcp = wcp + wd * (unorm8(n >> 24) * 1.0f);                     // :155
            wcp = wcp + randomVec * 2.5f;                                  // :156
            float bias = gsmoothstep(0.0f, 0.2f, depth) * 50.0f + 1.5f;    // :158
            V3 origin = wcp + normal * bias;
            uint64_t s = 0;
            float shadow = 1.0f;
            if (march(*vol, origin, wd, 128.0f, 0.5f, s, nullptr) != 128.0f) shadow = 0.0f;  // :167-169
            // calculateAmbientIrradiance(origin, normal)  :111-126
            V3 tangent = fabsf(normal.z) > 0.5f ? v3(0.0f, -normal.z, normal.y) : v3(-normal.y, normal.x, 0.0f);
            V3 bitangent = cross3(normal, tangent);
            float acc = 0.0f;
            for (int i = 0; i < n_ao; ++i) {
                uint32_t ni = (i == 0) ? n : get_noise(*gb, *view, p, i);
                V3 rv = cosine_sample_hemisphere(L, ni, ni >> 8);
                V3 dir = tangent * rv.x + bitangent * rv.y + normal * rv.z;
                float d = march(*vol, origin, dir, 128.0f, 2.5f, s, nullptr) / 128.0f;
                acc += d * d;
            }
            float ao = n_ao > 0 ? (acc / (float)n_ao) * 0.05f : 0.0f;   // AMBIENT_LIGHT_FACTOR :17
            out_shadow[idx] = shadow;
            out_ao[idx] = ao;
            trays += 1 + (uint64_t)n_ao; tsteps += s; tpix += 1;
        }
    }
    if (stats) { stats->rays = trays; stats->steps = tsteps; stats->pixels = tpix; }
}

// ---------------------------------------------------------------------------------------------
// A6  LightPoint.frag:85-129.  One plane per light; range-culled pixels (discard) -> shadow = 1.
// Note the reference applies no sky test in this pass.
// ---------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
template <bool SPOT>
void pass_local_light(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb,
                      const float* light_base, int stride_floats, int n_lights, vxo_rows rows,
                      float* out_shadow, vxo_stats* stats) {
    const Luts& L = luts();
    const int W = gb->width, H = gb->height;
    uint64_t trays = 0, tsteps = 0, tpix = 0;
    if (rows.step < 1) rows.step = 1;
    const int nrows = rows.end > rows.begin ? (rows.end - rows.begin + rows.step - 1) / rows.step : 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : trays, tsteps, tpix)
    for (int ri = 0; ri < nrows; ++ri) {
        const int py = rows.begin + ri * rows.step;
        if (py < 0 || py >= H) continue;
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            float depth = unorm24(gb->depth24[idx]);
            Pixel p = pixel_setup(*view, W, H, px, py);
            V3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));          // LightPoint.frag:89
            V3 normal = decode_normal(gb->normal[idx]);                  // :90
            V3 worldPos = xyz(mat_mul(view->InverseViewMatrix, V4{pos.x, pos.y, pos.z, 1.0f}));  // :95
            uint32_t n = get_noise(*gb, *view, p, -1);
            V3 rv0 = cosine_sample_hemisphere(L, n, n >> 8) * 0.1f;      // :111
            rv0.z *= gsign(unorm8(n >> 16) - 0.5f);                      // :112
            bool any = false;
            for (int li = 0; li < n_lights; ++li) {
                const float* lt = light_base + (size_t)li * stride_floats;
                V3 lpos{lt[0], lt[1], lt[2]};
                float range = lt[3];
                V3 lightDir = lpos - worldPos;                           // :97
                float lightDistance = length3(lightDir);                 // :98
                float* outp = out_shadow + (size_t)li * W * H + idx;
                if (lightDistance > range) { *outp = 1.0f; continue; }   // :100-103 discard
                V3 wd; float hitDist; float step0;
                if (!SPOT) { wd = lightDir; hitDist = lightDistance * 10.5f; step0 = 0.5f; }              // :108-109
                else { wd = normalize3(lightDir) * 10.0f; hitDist = lightDistance * 10.0f; step0 = 2.5f; } // LightSpot.frag:96-97
                V3 wcp = worldPos * 10.0f;                               // :110
                wd = mix3(wd, rv0, 0.5f);                                // :113
                wd = normalize3(wd);                                     // :114
                wcp = wcp + wd * (unorm8(n >> 24) * 1.0f);               // :115
                wcp = wcp + rv0 * 2.5f;                                  // :116
                uint64_t s = 0;
                float shadow = 1.0f;
                if (march(*vol, wcp + normal * 0.5f, wd, hitDist, step0, s, nullptr) < hitDist) shadow = 0.0f;  // :125
                *outp = shadow;
                trays += 1; tsteps += s; any = true;
            }
            if (any) tpix += 1;
        }
    }
    if (stats) { stats->rays = trays; stats->steps = tsteps; stats->pixels = tpix; }
}
}  // namespace
extern "C" {

void vxo_pass_point(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb,
                    const vxo_point_light* lights, int n_lights, vxo_rows rows, float* out_shadow, vxo_stats* stats) {
    pass_local_light<false>(vol, view, gb, (const float*)lights, (int)(sizeof(vxo_point_light) / 4), n_lights, rows, out_shadow, stats);
}
// LightSpot.frag:73-117 (same ray-gen; SuperSparse march, hitDist = 10*|L|)
void vxo_pass_spot(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb,
                   const vxo_spot_light* lights, int n_lights, vxo_rows rows, float* out_shadow, vxo_stats* stats) {
    pass_local_light<true>(vol, view, gb, (const float*)lights, (int)(sizeof(vxo_spot_light) / 4), n_lights, rows, out_shadow, stats);
}

// ---------------------------------------------------------------------------------------------
// A6  LightReflection.frag:60-113.  Output: t (256 on a miss / sky pixel).
// ---------------------------------------------------------------------------------------------
void vxo_pass_reflection(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb,
                         vxo_rows rows, float* out_t, vxo_stats* stats) {
    const Luts& L = luts();
    const int W = gb->width, H = gb->height;
    uint64_t trays = 0, tsteps = 0, tpix = 0;
    if (rows.step < 1) rows.step = 1;
    const int nrows = rows.end > rows.begin ? (rows.end - rows.begin + rows.step - 1) / rows.step : 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : trays, tsteps, tpix)
    for (int ri = 0; ri < nrows; ++ri) {
        const int py = rows.begin + ri * rows.step;
        if (py < 0 || py >= H) continue;
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            float depth = unorm24(gb->depth24[idx]);
            if (!(depth < 0.999f)) { out_t[idx] = 256.0f; continue; }    // :88
            Pixel p = pixel_setup(*view, W, H, px, py);
            V3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));          // :64
            V3 normal = decode_normal(gb->normal[idx]);                  // :65
            float roughness = unorm8(gb->material[idx]);                 // :68
            V3 V = normalize3(pos) * -1.0f;                              // :79
            V4 n4 = mat_mul(view->ViewMatrix, V4{normal.x, normal.y, normal.z, 0.0f});
            V3 N = xyz(n4);                                              // :80
            V3 I = V * -1.0f;
            V3 R = I - N * dot3(N, I) * 2.0f;                            // :81 reflect(-V, N)
            V3 wd = normalize3(xyz(mat_mul(view->InverseViewMatrix, V4{R.x, R.y, R.z, 0.0f})));   // :92
            V3 wcp = xyz(mat_mul(view->InverseViewMatrix, V4{pos.x, pos.y, pos.z, 1.0f})) * 10.0f; // :93
            uint32_t n = get_noise(*gb, *view, p, -1);
            V3 rv = cosine_sample_hemisphere(L, n, n >> 8);              // :94
            rv.z *= gsign(unorm8(n >> 16) - 0.5f);                       // :95
            wd = mix3(wd, rv, roughness * 0.1f);                         // :96
            float nw = unorm8(n >> 24);
            wcp = wcp + normal * nw;                                     // :97
            wd = wd * (1.0f + nw * 0.5f);                                // :98
            uint64_t s = 0;
            float t = march(*vol, wcp + normal, wd, 256.0f, 0.5f, s, nullptr);   // :113
            out_t[idx] = t;
            trays += 1; tsteps += s; tpix += 1;
        }
    }
    if (stats) { stats->rays = trays; stats->steps = tsteps; stats->pixels = tpix; }
}

// ---------------------------------------------------------------------------------------------
// A7  Sources/World/Systems/ShadowVoxSystem.cpp
// ---------------------------------------------------------------------------------------------
// :82-94 SetVolumeAt
// ---------------------------------------------------------------------------------------------
// the colour the light passes write to the RGBA16F light buffer (additive blend),
// i.e. what the reference's main() computes AFTER the shadow / AO march, as float32 before the
// attachment conversion.  Inputs: the same G-buffer plus COLOR_TEXTURE (albedo, RGBA8 UNORM) and the
// shadow / ao planes of the march passes.  pow() is powf (GLSL specifies pow by accuracy only): parity
// for these planes is a tolerance, not bit equality.  Sky pixels (which sample the sky-box cube map,
// outside the path) are written as 0.
// ---------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
const float PI_ = 3.14159265359f;                           // PBR.frag:1

inline V3 splat(float v) { return V3{v, v, v}; }
inline V3 max3(V3 a, V3 b) { return V3{fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
inline V3 div3(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 unorm8x3(uint32_t c) { return V3{unorm8(c), unorm8(c >> 8), unorm8(c >> 16)}; }

// PBR.frag:48-69.  NDF and G are evaluated by the shader but do not reach Lo (the specular term is commented out).
inline V3 pbr_direct_light(V3 radiance, V3 albedo, V3 V, V3 N, V3 L, float metallic) {
    const V3 F0 = mix3(splat(0.04f), splat(1.0f), metallic);                                   // :51
    const float p5 = powf(fmaxf(1.0f - fmaxf(dot3(N, V), 0.0f), 0.0f), 5.0f);                   // :3-5
    const V3 F = F0 + (splat(1.0f) - F0) * p5;
    V3 kD = splat(1.0f) - F;                                                                   // :58
    kD = kD * (1.0f - metallic);                                                               // :59
    const float NdotL = fmaxf(dot3(N, L), 0.0f);                                               // :65
    return (div3(kD * albedo, PI_) * radiance) * NdotL;                                        // :66
}

// LightAmbient.frag:54-79 screenspaceOcclusion (depth sampled with the nearest filter of evk's samplers; out of
// range reads 0, as the harness that runs the reference shader defines it)
inline float depth_at_uv(const vxo_gbuffer& gb, float u, float v) {
    const int x = (int)floorf(u * (float)gb.width), y = (int)floorf(v * (float)gb.height);
    if (x < 0 || y < 0 || x >= gb.width || y >= gb.height) return 0.0f;
    return unorm24(gb.depth24[(size_t)y * gb.width + x]);
}
inline float screenspace_occlusion(const vxo_view& view, const vxo_gbuffer& gb, V3 pos, V3 dir, float dist) {
    const V3 mid = pos + dir * dist;                                                           // :60
    const V4 midProj = mat_mul(view.ProjectionMatrix, V4{mid.x, mid.y, mid.z, 1.0f});          // :62
    const float sampleDepth = (midProj.w - NEAR_) / (FAR_ - NEAR_);                            // :63
    const float u = ((midProj.x / midProj.w) * 1.0f) * 0.5f + 0.5f;                            // :54-57 worldToUV
    const float v = ((midProj.y / midProj.w) * -1.0f) * 0.5f + 0.5f;
    const float minDepth = depth_at_uv(gb, u, v);                                              // :67
    const float maxDepth = minDepth + 0.2f / FAR_;                                             // :68 OCCLUSION_TICKNESS
    if (gclamp(u, 0.0f, 1.0f) != u || gclamp(v, 0.0f, 1.0f) != v) return 0.0f;                 // :70-72
    if (minDepth < sampleDepth && sampleDepth < maxDepth) return gsmoothstep(maxDepth, minDepth, sampleDepth) * dist;   // :75-77
    return 0.0f;
}

// LightAmbient.frag:89-109 calculateOcclusion(N) -- N is the VIEW-space normal
inline float calculate_occlusion(const vxo_view& view, const vxo_gbuffer& gb, const Luts& L, const Pixel& p, float depth, V3 N) {
    const V3 tangent = normalize3(fabsf(N.z) > 0.5f ? v3(0.0f, -N.z, N.y) : v3(-N.y, N.x, 0.0f));
    const V3 bitangent = normalize3(cross3(N, tangent));
    const V3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_) + NEAR_ / FAR_);                  // :94
    float occlusion = 0.0f;
    const float sizeMultiplier = 0.2f * (1.0f + depth * 0.0f);                                 // :97
    for (int i = 0; i < 4; ++i) {                                                              // SAMPLES
        const uint32_t n = get_noise(gb, view, p, i);
        const V3 rv = cosine_sample_hemisphere(L, n, n >> 8);
        const V3 dir = tangent * rv.x + bitangent * rv.y + N * rv.z;
        occlusion += screenspace_occlusion(view, gb, pos, normalize3(dir) * sizeMultiplier, unorm8(n >> 16));
    }
    occlusion /= 4.0f;
    occlusion *= 3.0f;                                                                         // OCCLUSION_STRENGTH
    return gclamp(1.0f - occlusion, 0.0f, 1.0f);
}
}  // namespace
extern "C" {

// LightAmbient.frag:134-214 after the march: out_Color = vec4(ambient + Lo, 0).  `ao` is the n_ao = 1 plane
// (d*d * AMBIENT_LIGHT_FACTOR); with the multi-sample extension it is the mean, used the same way.
void vxo_resolve_ambient(const vxo_view* view, const vxo_gbuffer* gb, const uint32_t* albedo_rgba8, const float* shadow,
                         const float* ao, vxo_rows rows, float* out_rgba) {
    const Luts& L = luts();
    const int W = gb->width, H = gb->height;
    const V3 SUN = sun_dir();
    const V3 SUN_COLOR = v3(0.9f, 0.9f, 0.8f) * 0.5f;                                          // :16
    if (rows.step < 1) rows.step = 1;
    const int nrows = rows.end > rows.begin ? (rows.end - rows.begin + rows.step - 1) / rows.step : 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ri = 0; ri < nrows; ++ri) {
        const int py = rows.begin + ri * rows.step;
        if (py < 0 || py >= H) continue;
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            float* o = out_rgba + idx * 4;
            o[0] = o[1] = o[2] = o[3] = 0.0f;
            const float depth = unorm24(gb->depth24[idx]);
            if (!(depth < 0.999f)) continue;                                                   // :138 (sky box look-up: outside the path)
            const Pixel p = pixel_setup(*view, W, H, px, py);
            const V3 alb = unorm8x3(albedo_rgba8[idx]);                                        // :139
            const uint32_t m = gb->material[idx];
            const float roughness = unorm8(m), metallic = unorm8(m >> 8), emit = unorm8(m >> 16);   // :182-184
            const V3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));                          // :141
            const V3 normal = decode_normal(gb->normal[idx]);
            const V3 sunDir = xyz(mat_mul(view->ViewMatrix, V4{SUN.x, SUN.y, SUN.z, 0.0f}));    // :179
            V3 F0 = mix3(splat(0.04f), alb, metallic);                                         // :187-188
            const V3 Vv = normalize3(pos) * -1.0f;                                             // :190
            const V3 N = xyz(mat_mul(view->ViewMatrix, V4{normal.x, normal.y, normal.z, 0.0f}));   // :191
            const V3 radiance = (SUN_COLOR * 1.0f) * shadow[idx];                              // :198
            const V3 Lo = pbr_direct_light(radiance, alb, Vv, N, sunDir, metallic);            // :199
            const float c = fmaxf(dot3(N, Vv), 0.0f);
            const V3 F = F0 + (max3(splat(1.0f - roughness), F0) - F0) * powf(fmaxf(1.0f - c, 0.0f), 5.0f);   // :204, PBR.frag:8-10
            const V3 kD = splat(1.0f) - F;                                                     // :206
            const float ey = (fmaxf(0.0f, 0.0f) * 0.8f + 0.2f) * 0.8f;                         // :128-131 getSkyColor(vec3(1,0,0))
            const V3 irradiance = v3(powf(1.0f - ey, 2.0f), 1.0f - ey, 0.6f + (1.0f - ey) * 0.4f) * 1.1f;
            const V3 diffuse = irradiance * alb;                                               // :208
            const V3 ambientIrradiance = splat(ao[idx]);                                       // :125
            const float occ = calculate_occlusion(*view, *gb, L, p, depth, N);
            const V3 ambient = diffuse * (splat(emit * 10.0f) + (kD * ambientIrradiance) * occ);   // :210
            const V3 c3 = ambient + Lo;                                                        // :213
            o[0] = c3.x; o[1] = c3.y; o[2] = c3.z; o[3] = 0.0f;
        }
    }
}

// LightPoint.frag:131-152 / LightSpot.frag:118-138 after the march: inout_rgba += sum over the lights, in list order,
// of vec4(Lo, 0) (the reference draws one additive full-screen pass per light).  Range-culled pixels add nothing.
// light stride: 8 floats (vxo_point_light) or 16 (vxo_spot_light).  shadow: [n_lights][H][W].
void vxo_resolve_local(const vxo_view* view, const vxo_gbuffer* gb, const uint32_t* albedo_rgba8, const float* lights,
                       int n_lights, int spot, const float* shadow, vxo_rows rows, float* inout_rgba) {
    const int W = gb->width, H = gb->height;
    const int stride = spot ? 16 : 8;
    if (rows.step < 1) rows.step = 1;
    const int nrows = rows.end > rows.begin ? (rows.end - rows.begin + rows.step - 1) / rows.step : 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ri = 0; ri < nrows; ++ri) {
        const int py = rows.begin + ri * rows.step;
        if (py < 0 || py >= H) continue;
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            const float depth = unorm24(gb->depth24[idx]);
            const Pixel p = pixel_setup(*view, W, H, px, py);
            const uint32_t m = gb->material[idx];
            const float metallic = unorm8(m >> 8);
            const V3 alb = unorm8x3(albedo_rgba8[idx]);
            const V3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));
            const V3 normal = decode_normal(gb->normal[idx]);
            const V3 worldPos = xyz(mat_mul(view->InverseViewMatrix, V4{pos.x, pos.y, pos.z, 1.0f}));
            const V3 Vv = normalize3(pos) * -1.0f;                                             // :139
            const V3 N = xyz(mat_mul(view->ViewMatrix, V4{normal.x, normal.y, normal.z, 0.0f}));
            float* o = inout_rgba + idx * 4;
            for (int li = 0; li < n_lights; ++li) {
                const float* lt = lights + (size_t)li * stride;
                const V3 lpos{lt[0], lt[1], lt[2]};
                const float range = lt[3];
                const V3 color{lt[4], lt[5], lt[6]};
                const float atten = lt[7];
                const V3 lightPos = xyz(mat_mul(view->ViewMatrix, V4{lpos.x, lpos.y, lpos.z, 1.0f}));   // :96
                const V3 lightDir = lpos - worldPos;
                const float lightDistance = length3(lightDir);
                if (lightDistance > range) continue;                                           // discard
                const V3 Lv = xyz(mat_mul(view->ViewMatrix, V4{lightDir.x, lightDir.y, lightDir.z, 0.0f}));   // :141
                const float dist = length3(lightPos - pos);                                    // :143 distance()
                float attenuation;
                if (!spot) attenuation = gclamp(range - lightDistance, 0.0f, 1.0f) / powf(dist, atten);   // :144
                else {
                    const V3 sdir{lt[8], lt[9], lt[10]};
                    const float angle = lt[11], angleAtten = lt[12];
                    const float angleDist = fmaxf(dot3(normalize3(lightDir), sdir) - (1.0f - angle), 0.0f) / angle;   // LightSpot.frag:132
                    attenuation = (powf(angleDist, angleAtten) * gclamp(range - lightDistance, 0.0f, 1.0f)) / powf(dist, atten);   // :133
                }
                const V3 radiance = (color * attenuation) * shadow[(size_t)li * W * H + idx];  // :146
                const V3 Lo = pbr_direct_light(radiance, alb, Vv, N, Lv, metallic);
                o[0] += Lo.x; o[1] += Lo.y; o[2] += Lo.z; o[3] += 0.0f;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the steps after the light passes.
//   vxo_light_taa           Sources/Shaders/LightTAA.frag:37-141 (temporal + spatial accumulation of the light buffer)
//   vxo_resolve_reflection  Sources/Shaders/LightReflection.frag:60-139 (the colour around the specular-occlusion march)
// Textures are sampled with the nearest filter of evk's samplers (Vendor/evk/evk.cpp:277-293): texel = floor(uv * size),
// out of range reads 0.  Light / motion planes are float32 (the reference's attachments are RGBA16F / RG16F; as in row f2
// the values are those before the attachment conversion).  glm's min / max / clamp are the ternaries of
// func_common.inl (they differ from fminf / fmaxf on NaN, which a zero weight sum can produce here).
// ---------------------------------------------------------------------------------------------
}  // extern "C"
namespace {
inline float tmin(float a, float b) { return (b < a) ? b : a; }            // glm::min
inline float tmax(float a, float b) { return (a < b) ? b : a; }            // glm::max
inline float tclamp(float x, float lo, float hi) { return tmin(tmax(x, lo), hi); }
inline int tex_index(int W, int H, float u, float v) {                     // nearest; -1 = outside
    const int x = (int)floorf(u * (float)W), y = (int)floorf(v * (float)H);
    if (x < 0 || y < 0 || x >= W || y >= H) return -1;
    return y * W + x;
}
struct TaaTexel { V3 color, normal, light; V4 material; float depth, mx, my; };
inline TaaTexel taa_fetch(const vxo_gbuffer& gb, const uint32_t* albedo, const float* motion, const float* light, float u, float v) {
    TaaTexel t;
    const int i = tex_index(gb.width, gb.height, u, v);
    if (i < 0) { t.color = t.normal = t.light = v3(0, 0, 0); t.material = V4{0, 0, 0, 0}; t.depth = t.mx = t.my = 0.0f; return t; }
    t.color = unorm8x3(albedo[i]);
    t.normal = decode_normal(gb.normal[i]);
    const uint32_t m = gb.material[i];
    t.material = V4{unorm8(m), unorm8(m >> 8), unorm8(m >> 16), unorm8(m >> 24)};
    t.depth = unorm24(gb.depth24[i]);
    t.mx = motion[(size_t)i * 2]; t.my = motion[(size_t)i * 2 + 1];
    t.light = v3(light[(size_t)i * 4], light[(size_t)i * 4 + 1], light[(size_t)i * 4 + 2]);
    return t;
}
inline float length4(V4 a) { return sqrtf((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w)); }   // glm compute_dot<vec4>
inline float length2(float x, float y) { return sqrtf(x * x + y * y); }
}  // namespace
extern "C" {

void vxo_light_taa(const vxo_view* view, const vxo_gbuffer* gb, const uint32_t* albedo_rgba8, const float* motion /* [H][W][2] */,
                   const float* light /* [H][W][4] */, const float* last_light /* [H][W][4] */, vxo_rows rows, float* out_rgba) {
    const int W = gb->width, H = gb->height;
    if (rows.step < 1) rows.step = 1;
    const int nrows = rows.end > rows.begin ? (rows.end - rows.begin + rows.step - 1) / rows.step : 0;
    const float iRx = 1.0f / (float)W, iRy = 1.0f / (float)H;                                   // :38
#pragma omp parallel for schedule(dynamic, 1)
    for (int ri = 0; ri < nrows; ++ri) {
        const int py = rows.begin + ri * rows.step;
        if (py < 0 || py >= H) continue;
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            float* o = out_rgba + idx * 4;
            const Pixel p = pixel_setup(*view, W, H, px, py);
            const TaaTexel c = taa_fetch(*gb, albedo_rgba8, motion, light, p.u, p.v);           // :40-47
            const float cla = light[idx * 4 + 3];
            const float oldU = p.u + c.mx, oldV = p.v + c.my;                                   // :41
            if (c.depth == 1.0f) { o[0] = c.light.x; o[1] = c.light.y; o[2] = c.light.z; o[3] = cla; continue; }   // :49-52
            const int li = tex_index(W, H, oldU, oldV);                                         // :53
            V3 lastLight = li < 0 ? v3(0, 0, 0) : v3(last_light[(size_t)li * 4], last_light[(size_t)li * 4 + 1], last_light[(size_t)li * 4 + 2]);
            float lastVariance = li < 0 ? 0.0f : last_light[(size_t)li * 4 + 3];               // :54
            if (tclamp(oldU, 0.0f, 1.0f) != oldU || tclamp(oldV, 0.0f, 1.0f) != oldV) {         // :57 out of bounds: current frame only
                float count = 0.0f;
                V3 neighColors = v3(0, 0, 0);
                for (int x = -9; x <= 9; ++x)
                    for (int y = -9; y <= 9; ++y) {
                        const float ox = (float)x * iRx, oy = (float)y * iRy;                   // :62
                        const float u = tclamp(p.u + ox, 0.001f, 0.999f), v = tclamp(p.v + oy, 0.001f, 0.999f);
                        const TaaTexel n = taa_fetch(*gb, albedo_rgba8, motion, light, u, v);
                        float factor = tmax(dot3(c.normal, n.normal), 0.0f);                    // :71
                        factor *= gstep(0.8f, 1.0f - length4(V4{c.material.x - n.material.x, c.material.y - n.material.y, c.material.z - n.material.z, c.material.w - n.material.w}));
                        factor *= 1.0f - tclamp(fabsf(c.depth - n.depth) * FAR_, 0.0f, 1.0f);   // :73
                        factor *= 1.0f - tclamp(length3(c.color - n.color), 0.0f, 1.0f);        // :74
                        if (length2(c.mx - n.mx, c.my - n.my) > 0.1f) factor = 0.0f;            // :76
                        neighColors = neighColors + n.light * factor;                           // :79
                        count += factor;
                    }
                o[0] = neighColors.x / count; o[1] = neighColors.y / count; o[2] = neighColors.z / count; o[3] = 1.0f;   // :83
                continue;
            }
            V3 nmin = splat(10000.0f), nmax = splat(0.0f), neighColors = splat(0.0f);           // :89-91
            float diffSum = 1.0f, count = 0.0f, radius = 1.0f;
            const float size = 12.0f;
            const uint32_t nz = get_noise(*gb, *view, p, -1);
            V3 cur = c.light;
            for (float angle = (unorm8(nz) * 3.1415f) * GOLDEN_RATIO; radius <= size; angle += 2.39f) {   // :96
                radius += 1.0f;
                const float cs = (float)cos((double)angle), sn = (float)sin((double)angle);     // correctly rounded (oracle definition)
                const float k1 = radius * (lastVariance + 1.0f), k2 = tclamp(0.1f, 0.5f, 1.0f / c.depth);   // :99 (sic: clamp(x = 0.1, 0.5, 1/depth))
                const float ox = ((cs * iRx) * k1) * k2, oy = ((sn * iRy) * k1) * k2;
                const float u = tclamp(p.u + ox, 0.001f, 0.999f), v = tclamp(p.v + oy, 0.001f, 0.999f);
                const TaaTexel n = taa_fetch(*gb, albedo_rgba8, motion, light, u, v);
                float factor = 1.212f - radius / size;                                          // :108
                factor *= gstep(0.8f, 1.0f - length4(V4{c.material.x - n.material.x, c.material.y - n.material.y, c.material.z - n.material.z, c.material.w - n.material.w}));
                factor *= tmax(dot3(c.normal, n.normal), 0.0f);
                factor *= 1.0f - tclamp(fabsf(c.depth - n.depth) * FAR_, 0.0f, 1.0f);
                factor *= 1.0f - tclamp(length3(c.color - n.color) * 10000.0f, 0.0f, 1.0f);
                if (length2(c.mx - n.mx, c.my - n.my) > 0.1f) factor = 0.0f;
                nmin = v3(tmin(nmin.x, n.light.x), tmin(nmin.y, n.light.y), tmin(nmin.z, n.light.z));   // :116
                nmax = v3(tmax(nmax.x, n.light.x), tmax(nmax.y, n.light.y), tmax(nmax.z, n.light.z));
                neighColors = neighColors + n.light * factor;
                count += factor;
                diffSum += length3(n.light - c.light) * factor;                                 // :121
            }
            cur = cur + neighColors;                                                            // :123
            cur = div3(cur, count + 1.0f);                                                      // :124
            diffSum /= radius;                                                                  // :125
            lastLight = v3(tclamp(lastLight.x, nmin.x, nmax.x), tclamp(lastLight.y, nmin.y, nmax.y), tclamp(lastLight.z, nmin.z, nmax.z));   // :129
            const V3 ad = v3(fabsf(cur.x - lastLight.x), fabsf(cur.y - lastLight.y), fabsf(cur.z - lastLight.z));
            float variance = dot3(ad, v3(0.2125f, 0.7154f, 0.0721f));                           // :132, :30-35
            variance = tclamp(variance * 5.5f, 0.0f, 1.0f);
            lastVariance += variance;
            lastVariance -= diffSum * 0.08f;
            lastVariance += length2(c.mx, c.my) * 20.0f;                                        // :136
            const V3 m = mix3(cur, lastLight, tclamp(1.0f - lastVariance, 0.3f, 0.9f));         // :141
            o[0] = m.x; o[1] = m.y; o[2] = m.z; o[3] = tclamp(lastVariance * 0.7f, 0.0f, 1.0f);
        }
    }
}

// LightReflection.frag:60-139: out_Color = vec4(ambient * F * (1 - roughness), F.x), ambient = the (TAA) light buffer where the
// reflected ray's end point is the visible surface, the sky-box colour on a miss (`sky_rgb`: a uniform sky stands in for the
// cube map, which is outside the path).  `t` is the plane of vxo_pass_reflection.  Tolerance parity (pow).
void vxo_resolve_reflection(const vxo_view* view, const vxo_gbuffer* gb, const float* t_plane, const float* light /* [H][W][4] or NULL */,
                            const float* sky_rgb, vxo_rows rows, float* out_rgba) {
    const Luts& L = luts();
    const int W = gb->width, H = gb->height;
    if (rows.step < 1) rows.step = 1;
    const int nrows = rows.end > rows.begin ? (rows.end - rows.begin + rows.step - 1) / rows.step : 0;
    const V3 sky = v3(sky_rgb[0], sky_rgb[1], sky_rgb[2]);
#pragma omp parallel for schedule(dynamic, 1)
    for (int ri = 0; ri < nrows; ++ri) {
        const int py = rows.begin + ri * rows.step;
        if (py < 0 || py >= H) continue;
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            float* o = out_rgba + idx * 4;
            const Pixel p = pixel_setup(*view, W, H, px, py);
            const uint32_t mt = gb->material[idx];
            const float roughness = unorm8(mt), metallic = unorm8(mt >> 8);                     // :68-69
            const float depth = unorm24(gb->depth24[idx]);
            const V3 pos = p.farvec * (depth * (1.0f + 1.0f / FAR_));                           // :64
            const V3 normal = decode_normal(gb->normal[idx]);
            const V3 F0 = mix3(splat(0.04f), splat(1.0f), metallic);                            // :74-77
            const V3 V = normalize3(pos) * -1.0f;                                               // :79
            const V3 N = xyz(mat_mul(view->ViewMatrix, V4{normal.x, normal.y, normal.z, 0.0f}));
            const V3 I = V * -1.0f;
            const V3 R = I - N * dot3(N, I) * 2.0f;                                             // :81
            V3 ambient = splat(0.0f);
            const float cth = fmaxf(dot3(N, V), 0.0f);
            const V3 F = (F0 + (max3(splat(1.0f - roughness), F0) - F0) * powf(fmaxf(1.0f - cth, 0.0f), 5.0f)) * 5.0f + splat(0.0f);   // :86
            if (depth < 0.999f) {
                V3 wd = normalize3(xyz(mat_mul(view->InverseViewMatrix, V4{R.x, R.y, R.z, 0.0f})));
                V3 wcp = xyz(mat_mul(view->InverseViewMatrix, V4{pos.x, pos.y, pos.z, 1.0f})) * 10.0f;
                const uint32_t n = get_noise(*gb, *view, p, -1);
                V3 rv = cosine_sample_hemisphere(L, n, n >> 8);
                rv.z *= gsign(unorm8(n >> 16) - 0.5f);
                wd = mix3(wd, rv, roughness * 0.1f);
                const float nw = unorm8(n >> 24);
                wcp = wcp + normal * nw;
                wd = wd * (1.0f + nw * 0.5f);
                const float t = t_plane[idx];                                                   // :113
                const V3 e = (wcp + wd * t) * 0.1f;
                const V3 hp = xyz(mat_mul(view->ViewMatrix, V4{e.x, e.y, e.z, 1.0f}));           // :115
                const V4 pp = mat_mul(view->ProjectionMatrix, V4{hp.x, hp.y, hp.z, 1.0f});       // :116
                const float linearDepth = (pp.w - NEAR_) / (FAR_ - NEAR_);                      // :117
                const float u = ((pp.x / pp.w) * 1.0f) * 0.5f + 0.5f, v = ((pp.y / pp.w) * -1.0f) * 0.5f + 0.5f;   // :118
                const int ti = tex_index(W, H, u, v);
                const float d2 = ti < 0 ? 0.0f : unorm24(gb->depth24[ti]);                      // :119
                if (t == 256.0f) ambient = ambient + sky;                                       // :121-122
                else if (linearDepth > d2 - 0.001f) {                                           // :124
                    if (linearDepth < d2 + 0.001f && ti >= 0 && light)                          // :125-126
                        ambient = ambient + v3(light[(size_t)ti * 4], light[(size_t)ti * 4 + 1], light[(size_t)ti * 4 + 2]);
                }
            }
            const V3 c3 = (ambient * F) * (1.0f - roughness);                                   // :139
            o[0] = c3.x; o[1] = c3.y; o[2] = c3.z; o[3] = F.x;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the G-buffer producer's traversal of ONE model volume -- the reference's
// hierarchical-mip DDA.  VoxAsset::Upload's mip rule (Sources/Asset/VoxAsset.cpp:3-64), clipToAABB
// (Sources/Shaders/GeometryVoxel.frag:49-61) and intersectVolume (:64-125).  Ray-level entry: what one
// fragment of GeometryVoxel.frag does between its inputs (In.localCameraPos, In.localDirection, UV) and
// (hit, hitPos, hitNormal, hitMat); the rasterisation around it is not restated.
// ---------------------------------------------------------------------------------------------

// One mip level from its parent: voxel = first non-zero of the 2x2x2 children in the order vi = 0..7 (x fastest, then
// y, then z); the reference's loop runs to vi = 8, which revisits child 0.  Sizes halve while > 1 (VoxAsset.cpp:19-21).
void vxo_model_mip(const uint8_t* parent, int psx, int psy, int psz, uint8_t* out) {
    const int sx = psx > 1 ? psx / 2 : 1, sy = psy > 1 ? psy / 2 : 1, sz = psz > 1 ? psz / 2 : 1;
    for (int z = 0; z < sz; ++z)
        for (int y = 0; y < sy; ++y)
            for (int x = 0; x < sx; ++x) {
                uint8_t vox = 0;
                for (int vi = 0; vi < 8; ++vi) {
                    const int cx = 2 * x + (vi & 1), cy = 2 * y + ((vi >> 1) & 1), cz = 2 * z + ((vi >> 2) & 1);
                    if (cx >= psx || cy >= psy || cz >= psz) continue;   // (sizes are multiples of 4, VoxAsset.h:27-29: never taken for 3 mips)
                    const uint8_t up = parent[(size_t)cx + (size_t)cy * psx + (size_t)cz * psx * psy];
                    if (up) { vox = up; break; }
                }
                out[(size_t)x + (size_t)y * sx + (size_t)z * sx * sy] = vox;
            }
}

}  // extern "C"
namespace {
struct ModelMips { const uint8_t* d[3]; int sx[3], sy[3], sz[3]; };
inline unsigned model_fetch(const ModelMips& M, int x, int y, int z, int mip) {   // texelFetch(VOLUME_TEXTURE, p, mip).r, out of range -> 0
    if (mip < 0 || mip > 2 || x < 0 || y < 0 || z < 0 || x >= M.sx[mip] || y >= M.sy[mip] || z >= M.sz[mip]) return 0u;
    return M.d[mip][(size_t)x + (size_t)y * M.sx[mip] + (size_t)z * M.sx[mip] * M.sy[mip]];
}
inline float ground(float x) { return roundf(x); }       // glm::round -> std::round (func_common.inl:203-209)
}  // namespace
extern "C" {

}  // extern "C"
namespace {
inline ModelMips make_mips(const uint8_t* mip0, const uint8_t* mip1, const uint8_t* mip2, int sx, int sy, int sz) {
    ModelMips M;
    M.d[0] = mip0; M.d[1] = mip1; M.d[2] = mip2;
    M.sx[0] = sx; M.sy[0] = sy; M.sz[0] = sz;
    for (int m = 1; m < 3; ++m) { M.sx[m] = M.sx[m - 1] > 1 ? M.sx[m - 1] / 2 : 1; M.sy[m] = M.sy[m - 1] > 1 ? M.sy[m - 1] / 2 : 1; M.sz[m] = M.sz[m - 1] > 1 ? M.sz[m - 1] / 2 : 1; }
    return M;
}

// clipToAABB (:49-61) + intersectVolume (:64-125) for one fragment: cam = In.localCameraPos, dir = In.localDirection, uv = UV
inline vxo_model_hit traverse_model(const ModelMips& M, V3 cam, V3 dir, float u, float v, int frame, float res_x, float res_y) {
    vxo_model_hit h;
    memset(&h, 0, sizeof h);
    const int sx = M.sx[0], sy = M.sy[0], sz = M.sz[0];
    const V3 vsize = v3((float)sx, (float)sy, (float)sz);
    const V3 direction = normalize3(dir);                                                     // GeometryVoxel.frag:145
    V3 origin = cam;
    if (!(gclamp(cam.x, 0.0f, vsize.x) == cam.x && gclamp(cam.y, 0.0f, vsize.y) == cam.y && gclamp(cam.z, 0.0f, vsize.z) == cam.z)) {
        const V3 invDir = v3(1.0f, 1.0f, 1.0f) / direction;
        const V3 sgn = v3(gstep(direction.x, 0.0f), gstep(direction.y, 0.0f), gstep(direction.z, 0.0f));
        const V3 t = (sgn * vsize - cam) * invDir;
        const float tmin = fmaxf(fmaxf(t.x, t.y), t.z);
        origin = cam + direction * (tmin - 0.001f);
    }
    const V3 stepSign = v3(gsign(direction.x), gsign(direction.y), gsign(direction.z));
    const V3 t_delta = v3(1.0f, 1.0f, 1.0f) / (direction * stepSign);
    int mip = 2, i = 0, nt = 0, fetches = 0;
    bool done = false;
    do {
        const float mipSize = (float)(1 << mip);
        origin = v3(origin.x / mipSize, origin.y / mipSize, origin.z / mipSize);
        int cx = f2i(floorf(origin.x)), cy = f2i(floorf(origin.y)), cz = f2i(floorf(origin.z));
        const V3 next_bounds = v3((float)cx, (float)cy, (float)cz) + (stepSign * 0.5f + v3(0.5f, 0.5f, 0.5f));
        V3 t_max = (next_bounds - origin) / direction;
        const int hx = f2i((float)sx / mipSize) + 1, hy = f2i((float)sy / mipSize) + 1, hz = f2i((float)sz / mipSize) + 1;   // ivec3(volumeDimension / mipSize) + 1
        int nn = 0;
        do {
            const V3 select = v3(gstep(t_max.x, t_max.z) * gstep(t_max.x, t_max.y), gstep(t_max.y, t_max.x) * gstep(t_max.y, t_max.z),
                                 gstep(t_max.z, t_max.y) * gstep(t_max.z, t_max.x));
            const V3 adv = select * stepSign;
            cx += f2i(adv.x); cy += f2i(adv.y); cz += f2i(adv.z);
            if (cx < -1 || cy < -1 || cz < -1 || cx > hx || cy > hy || cz > hz) { done = true; break; }   // :84-86 return false
            const unsigned voxel = model_fetch(M, cx, cy, cz, mip);
            ++fetches;
            if (voxel != 0u) {
                const float best_t = dot3(t_max, select);
                const V3 at = (origin + direction * best_t) * mipSize;
                if (mip == 0 || (float)mip < 0.001f * length3(at - cam)) {                       // :92-96 LOD early accept
                    const float cxr = ground(u * res_x * 0.5f), cyr = ground(v * res_y * 0.5f);   // :99
                    const bool glass = voxel < 16u && gmod(cyr + cxr, 2.0f) == (float)(frame % 2);   // :100
                    if (!glass) {
                        h.hit = 1; h.material = voxel;
                        const V3 nrm = (stepSign * -1.0f) * select;
                        h.normal[0] = nrm.x; h.normal[1] = nrm.y; h.normal[2] = nrm.z;
                        h.pos[0] = at.x; h.pos[1] = at.y; h.pos[2] = at.z;
                        done = true;
                        break;
                    }
                } else {
                    mip--;                                                                       // :110-112
                    origin = origin + v3((direction.x * best_t) / mipSize, (direction.y * best_t) / mipSize, (direction.z * best_t) / mipSize);
                    break;
                }
            }
            t_max = t_max + t_delta * select;
            nt++;
        } while (++nn < 512);
        if (done) break;
        origin = origin * mipSize;                                                               // :121
    } while (++i < 4);
    h.fetches = fetches; h.steps = nt;
    return h;
}

// mat4 * mat4 in glm's order (type_mat4x4.inl operator*): column j of the product = A0*B[j][0] + A1*B[j][1] + A2*B[j][2] + A3*B[j][3],
// summed left to right
inline void mat_mat(const float* A, const float* B, float* R) {
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            R[j * 4 + r] = ((A[0 + r] * B[j * 4 + 0] + A[4 + r] * B[j * 4 + 1]) + A[8 + r] * B[j * 4 + 2]) + A[12 + r] * B[j * 4 + 3];
}
inline V4 normalize4(V4 v) {
    const float inv = 1.0f / sqrtf((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));      // glm dot(vec4): (x + y) + (z + w), func_geometric.inl:58-64
    return V4{v.x * inv, v.y * inv, v.z * inv, v.w * inv};
}
}  // namespace
extern "C" {

// rays: vxo_model_ray {cam[3] = In.localCameraPos, dir[3] = In.localDirection (not normalised), uv[2]}.
// out: vxo_model_hit {hit, material, fetches (texelFetch calls), steps (nt), pos[3], normal[3]}.
void vxo_trace_model_rays(const uint8_t* mip0, const uint8_t* mip1, const uint8_t* mip2, int sx, int sy, int sz,
                          const vxo_model_ray* rays, int64_t n, int frame, float res_x, float res_y, vxo_model_hit* out) {
    const ModelMips M = make_mips(mip0, mip1, mip2, sx, sy, sz);
#pragma omp parallel for schedule(static)
    for (int64_t ri = 0; ri < n; ++ri) {
        const vxo_model_ray& r = rays[ri];
        out[ri] = traverse_model(M, v3(r.cam[0], r.cam[1], r.cam[2]), v3(r.dir[0], r.dir[1], r.dir[2]), r.uv[0], r.uv[1], frame, res_x, res_y);
    }
}

// One fragment of GeometryVoxel.frag's main() (:127-182) from its interpolated inputs: UV (:140-142), traversal, and on a
// hit the G-buffer outputs -- palette colour with alpha = step(hitMat, 16), palette material, world normal, motion
// vector, gl_FragDepth with the anti-z-fight factor (1 - 1e-7 * VolumeRID).  Floats, before the attachment conversions.
void vxo_geometry_fragment(const vxo_model* model_mips /* [3] */, const vxo_view* view, const vxo_vox_cmd* cmd,
                           const uint32_t* pal_color, const uint32_t* pal_material, const vxo_frag_in* in, int64_t n, vxo_frag_out* out) {
    const ModelMips M = make_mips(model_mips[0].voxels, model_mips[1].voxels, model_mips[2].voxels, model_mips[0].sx, model_mips[0].sy, model_mips[0].sz);
    float PV[16];
    mat_mat(view->ProjectionMatrix, view->ViewMatrix, PV);
    float PVlast[16];
    mat_mat(view->ProjectionMatrix, view->LastViewMatrix, PVlast);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const vxo_frag_in& f = in[i];
        vxo_frag_out o;
        memset(&o, 0, sizeof o);
        const V4 sd = mat_mul(f.mvp, V4{f.dir[0], f.dir[1], f.dir[2], 0.0f});                              // :140
        float u = sd.x / sd.w, v = sd.y / sd.w;                                                        // :141
        u += view->Jitter[0] * view->iRes[0]; v += view->Jitter[1] * view->iRes[1];                       // :142
        const vxo_model_hit h = traverse_model(M, v3(f.cam[0], f.cam[1], f.cam[2]), v3(f.dir[0], f.dir[1], f.dir[2]), u, v, view->Frame, view->Res[0], view->Res[1]);
        o.hit = h.hit; o.material_index = h.material; o.fetches = h.fetches;
        if (h.hit) {
            const uint32_t pc = pal_color[(size_t)cmd->PalleteIndex * 256 + h.material], pm = pal_material[(size_t)cmd->PalleteIndex * 256 + h.material];
            o.color[0] = unorm8(pc); o.color[1] = unorm8(pc >> 8); o.color[2] = unorm8(pc >> 16);
            o.color[3] = gstep((float)h.material, 16.0f);                                              // :155 step(hitMat, 16)
            o.material[0] = unorm8(pm); o.material[1] = unorm8(pm >> 8); o.material[2] = unorm8(pm >> 16); o.material[3] = unorm8(pm >> 24);
            const V4 nw = normalize4(mat_mul(cmd->WorldMatrix, V4{h.normal[0], h.normal[1], h.normal[2], 0.0f}));   // :157
            o.normal[0] = nw.x; o.normal[1] = nw.y; o.normal[2] = nw.z; o.normal[3] = nw.w;
            const V4 hp = V4{h.pos[0] * 0.1f, h.pos[1] * 0.1f, h.pos[2] * 0.1f, 1.0f};
            const V4 worldHit = mat_mul(cmd->WorldMatrix, hp), lastWorldHit = mat_mul(cmd->LastWorldMatrix, hp);   // :159-160
            const V4 cur = mat_mul(PV, worldHit), last = mat_mul(PVlast, lastWorldHit);                    // :161-162
            o.motion[0] = -0.5f * (cur.x / cur.w - last.x / last.w);                                    // :166
            o.motion[1] = 0.5f * (cur.y / cur.w - last.y / last.w);
            const float linearDepth = cur.w;                                                           // :169
            o.depth = (1.0f - 0.0000001f * (float)cmd->VolumeRID) * (linearDepth - 0.1f) / (FAR_ - 0.1f);   // :170
        }
        out[i] = o;
    }
}

}  // extern "C"
namespace {
// inverse of an affine mat4 (last row 0 0 0 1; TransformSystem.cpp:124-135 builds T*R*S): adjugate / determinant of the 3x3
// block, then -A^-1 t.  Stands in for the vertex shader's inverse(cmd.WorldMatrix) (GeometryVoxel.vert:66).
inline void affine_inverse(const float* M, float* R) {
    const float a = M[0], b = M[4], c = M[8], d = M[1], e = M[5], f = M[9], g = M[2], h = M[6], i = M[10];
    const float A = e * i - f * h, B = f * g - d * i, C = d * h - e * g;
    const float det = (a * A + b * B) + c * C;
    const float id = 1.0f / det;
    const float r00 = A * id, r01 = (c * h - b * i) * id, r02 = (b * f - c * e) * id;
    const float r10 = B * id, r11 = (a * i - c * g) * id, r12 = (c * d - a * f) * id;
    const float r20 = C * id, r21 = (b * g - a * h) * id, r22 = (a * e - b * d) * id;
    const float tx = M[12], ty = M[13], tz = M[14];
    R[0] = r00; R[4] = r01; R[8] = r02; R[12] = -((r00 * tx + r01 * ty) + r02 * tz);
    R[1] = r10; R[5] = r11; R[9] = r12; R[13] = -((r10 * tx + r11 * ty) + r12 * tz);
    R[2] = r20; R[6] = r21; R[10] = r22; R[14] = -((r20 * tx + r21 * ty) + r22 * tz);
    R[3] = 0.0f; R[7] = 0.0f; R[11] = 0.0f; R[15] = 1.0f;
}
inline uint32_t pack_unorm8x4(const float* v) {
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)rintf(gclamp(v[k], 0.0f, 1.0f) * 255.0f) << (8 * k);
    return r;
}
inline uint32_t pack_snorm8x4(const float* v) {
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= ((uint32_t)(int)rintf(gclamp(v[k], -1.0f, 1.0f) * 127.0f) & 0xFFu) << (8 * k);
    return r;
}
}  // namespace
extern "C" {

// The geometry pass over a list of draws (GeometryVoxelPipeline::Use, Pipelines/GeometryVoxelPipeline.h:49-71): per pixel, in
// list order, every model whose box the pixel's view ray enters from outside (the pipeline culls back faces, evk.cpp:470, so a
// camera inside a box sees nothing of it) runs the fragment above; depth test LESS on the D24 value (evk.cpp:481), outputs
// converted to the attachment formats (Graphics.h:51-60).  Declared definition: the interpolated fragment inputs are
// evaluated per pixel centre (In.localDirection = the pixel's view ray in model space), like In.FarVec for the light passes.
void vxo_gbuffer_models(const vxo_view* view, int W, int H, const vxo_vox_cmd* cmds, int n_cmds, const vxo_model* mips /* [n_models][3] */,
                        const uint32_t* pal_color, const uint32_t* pal_material, uint32_t* depth24, uint32_t* normal, uint32_t* material,
                        uint32_t* albedo, float* motion /* [H][W][2] or NULL */) {
    struct Derived { float inv[16], mvp[16]; V3 cam; V3 size; };
    std::vector<Derived> D((size_t)n_cmds);
    float PV[16];
    mat_mat(view->ProjectionMatrix, view->ViewMatrix, PV);
    const V3 camW = v3(view->CameraPosition[0], view->CameraPosition[1], view->CameraPosition[2]);
    for (int c = 0; c < n_cmds; ++c) {
        affine_inverse(cmds[c].WorldMatrix, D[c].inv);
        mat_mat(PV, cmds[c].WorldMatrix, D[c].mvp);                                             // GeometryVoxel.vert:62
        D[c].cam = xyz(mat_mul(D[c].inv, V4{camW.x, camW.y, camW.z, 1.0f})) * 10.0f;            // :66
        const vxo_model& m0 = mips[(size_t)cmds[c]._pad[0] * 3];
        D[c].size = v3((float)m0.sx, (float)m0.sy, (float)m0.sz);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            uint32_t best = 0xFFFFFFu, onrm = 0u, omat = 0u, oalb = 0u;
            float omx = 0.0f, omy = 0.0f;
            // the pixel's view ray in world space (un-jittered: the vertex shader shifts the geometry by +jitter, :70-71)
            const float u = ((float)px + 0.5f) / (float)W, v = ((float)py + 0.5f) / (float)H;
            const float ndcx = (2.0f * u - 1.0f) - view->Jitter[0] * view->iRes[0] * 2.0f, ndcy = (1.0f - 2.0f * v) - view->Jitter[1] * view->iRes[1] * 2.0f;
            const V4 fp = mat_mul(view->InverseProjectionMatrix, V4{ndcx, ndcy, 1.0f, 1.0f});
            const V3 farv = v3(fp.x / fp.w, fp.y / fp.w, fp.z / fp.w);
            const V3 farW = xyz(mat_mul(view->InverseViewMatrix, V4{farv.x, farv.y, farv.z, 1.0f}));
            const V3 dirW = farW - camW;
            for (int c = 0; c < n_cmds; ++c) {
                const Derived& d = D[c];
                const V3 ld = xyz(mat_mul(d.inv, V4{dirW.x, dirW.y, dirW.z, 0.0f})) * 10.0f;
                // coverage: camera outside the box and the ray enters it in front of the camera (slab test)
                const V3 lc = d.cam;
                if (lc.x >= 0.0f && lc.x <= d.size.x && lc.y >= 0.0f && lc.y <= d.size.y && lc.z >= 0.0f && lc.z <= d.size.z) continue;
                float t0 = 0.0f, t1 = 3.0e38f;
                bool miss = false;
                const float lo[3] = {lc.x, lc.y, lc.z}, dd[3] = {ld.x, ld.y, ld.z}, sz3[3] = {d.size.x, d.size.y, d.size.z};
                for (int a = 0; a < 3; ++a) {
                    if (dd[a] == 0.0f) { if (lo[a] < 0.0f || lo[a] > sz3[a]) miss = true; continue; }
                    const float ta = (0.0f - lo[a]) / dd[a], tb = (sz3[a] - lo[a]) / dd[a];
                    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                }
                if (miss || !(t0 <= t1)) continue;
                vxo_frag_in fi;
                fi.cam[0] = lc.x; fi.cam[1] = lc.y; fi.cam[2] = lc.z; fi.dir[0] = ld.x; fi.dir[1] = ld.y; fi.dir[2] = ld.z;
                memcpy(fi.mvp, d.mvp, 64);
                vxo_frag_out fo;
                vxo_geometry_fragment(&mips[(size_t)cmds[c]._pad[0] * 3], view, &cmds[c], pal_color, pal_material, &fi, 1, &fo);
                if (!fo.hit) continue;                                                          // discard
                const uint32_t d24 = (uint32_t)rintf(gclamp(fo.depth, 0.0f, 1.0f) * 16777215.0f);
                if (d24 < best) {                                                               // CompareOp::eLess
                    best = d24; onrm = pack_snorm8x4(fo.normal); omat = pack_unorm8x4(fo.material); oalb = pack_unorm8x4(fo.color);
                    omx = fo.motion[0]; omy = fo.motion[1];
                }
            }
            depth24[idx] = best; normal[idx] = onrm; material[idx] = omat; albedo[idx] = oalb;
            if (motion) { motion[idx * 2] = omx; motion[idx * 2 + 1] = omy; }
        }
}

void vxo_set_volume_at(uint8_t* data, int sx, int sy, int sz, int x, int y, int z, int value) {
    if (x < 0 || y < 0 || z < 0 || x >= sx * 2 || y >= sy * 2 || z >= sz * 2) return;
    int bit = (x & 1) | ((y & 1) << 1) | ((z & 1) << 2);
    int mask = 1 << bit;
    x /= 2; y /= 2; z /= 2;
    uint8_t* p = data + ((size_t)x + (size_t)y * (size_t)sx + (size_t)z * (size_t)sx * (size_t)sy);
    *p = (uint8_t)((*p & ~mask) | (value << bit));
}

int vxo_get_volume_at(const vxo_volume* vol, int x, int y, int z, int mip) {
    return get_volume_at(*vol, I3{x, y, z}, mip) ? 1 : 0;
}

namespace {
struct Basis { V3 o, dx, dy, dz; };
// glm::translate(m, -pivot) (Vendor/glm/ext/matrix_transform.inl:10-15) then the o/dx/dy/dz
// extraction of ShadowVoxSystem.cpp:134-140.
inline Basis basis_from(const float* m, const float* pivot, bool use_pivot) {
    Basis b;
    if (use_pivot) {
        float v0 = -pivot[0], v1 = -pivot[1], v2 = -pivot[2];
        b.o.x = ((m[0] * v0 + m[4] * v1) + m[8] * v2) + m[12];
        b.o.y = ((m[1] * v0 + m[5] * v1) + m[9] * v2) + m[13];
        b.o.z = ((m[2] * v0 + m[6] * v1) + m[10] * v2) + m[14];
    } else {
        b.o = V3{m[12], m[13], m[14]};
    }
    b.dx = V3{m[0] * 0.1f, m[1] * 0.1f, m[2] * 0.1f};
    b.dy = V3{m[4] * 0.1f, m[5] * 0.1f, m[6] * 0.1f};
    b.dz = V3{m[8] * 0.1f, m[9] * 0.1f, m[10] * 0.1f};
    return b;
}
inline void stamp(uint8_t* data, int sx, int sy, int sz, const vxo_model& mdl, const Basis& b, int value, I3& mn, I3& mx) {
    for (int z = 0; z < mdl.sz; z++)
        for (int y = 0; y < mdl.sy; y++)
            for (int x = 0; x < mdl.sx; x++) {
                if (mdl.voxels[(size_t)x + (size_t)y * mdl.sx + (size_t)z * mdl.sx * mdl.sy] >= 16) {   // :145
                    V3 wp = b.o + b.dx * (float)x + b.dy * (float)y + b.dz * (float)z;                  // :146
                    I3 f{f2i(wp.x * 10.0f), f2i(wp.y * 10.0f), f2i(wp.z * 10.0f)};                      // :147
                    mn.x = std::min(mn.x, f.x); mn.y = std::min(mn.y, f.y); mn.z = std::min(mn.z, f.z);
                    mx.x = std::max(mx.x, f.x); mx.y = std::max(mx.y, f.y); mx.z = std::max(mx.z, f.z);
                    vxo_set_volume_at(data, sx, sy, sz, f.x, f.y, f.z, value);
                }
            }
}
}  // namespace

// :116-191 OnUpdate (per visited entity) and :7-53 OnVoxDestroyed, sequential, in array order.
void vxo_voxelize(uint8_t* data, int sx, int sy, int sz, const vxo_model* models,
                  const vxo_entity* ents, int n, vxo_region* out_regions, int32_t* out_valid) {
    for (int e = 0; e < n; ++e) {
        const vxo_entity& en = ents[e];
        const vxo_model& mdl = models[en.model];
        I3 startmin{sx - 1, sy - 1, sz - 1};   // :128 (texel units, mixed with voxel units below: sic)
        I3 startmax{0, 0, 0};
        I3 mn = startmin, mx = startmax;
        if (en.flags & VXO_ENT_DESTROY) {
            Basis b = basis_from(en.cur, en.pivot, false);   // :22-25 pivot ignored
            stamp(data, sx, sy, sz, mdl, b, 0, mn, mx);
        } else {
            Basis bp = basis_from(en.prev, en.pivot, true);
            stamp(data, sx, sy, sz, mdl, bp, 0, mn, mx);
            Basis bc = basis_from(en.cur, en.pivot, true);
            stamp(data, sx, sy, sz, mdl, bc, 1, mn, mx);
        }
        int valid = 0;
        vxo_region r;
        memset(&r, 0, sizeof r);
        if (mx.x != startmax.x || mx.y != startmax.y || mx.z != startmax.z) {   // :181
            mn.x /= 2; mn.y /= 2; mn.z /= 2;
            mx.x /= 2; mx.y /= 2; mx.z /= 2;
            mn.x = std::max(mn.x, 0); mn.y = std::max(mn.y, 0); mn.z = std::max(mn.z, 0);
            mx.x = std::min(mx.x, startmin.x); mx.y = std::min(mx.y, startmin.y); mx.z = std::min(mx.z, startmin.z);
            r.x = mn.x; r.y = mn.y; r.z = mn.z;
            r.w = (uint32_t)(mx.x - mn.x + 1); r.h = (uint32_t)(mx.y - mn.y + 1); r.d = (uint32_t)(mx.z - mn.z + 1);
            r.mip = 0;
            valid = 1;
        }
        if (out_regions) out_regions[e] = r;
        if (out_valid) out_valid[e] = valid;
    }
}

// Vendor/evk/evk.cpp:759-780: bufferOffset = x + y*W + z*W*H, row length W, image height H.
void vxo_upload_regions(uint8_t* image, const uint8_t* staging, int sx, int sy, int sz,
                        const vxo_region* regions, int n) {
    for (int i = 0; i < n; ++i) {
        const vxo_region& r = regions[i];
        for (uint32_t z = 0; z < r.d; ++z)
            for (uint32_t y = 0; y < r.h; ++y) {
                int zz = r.z + (int)z, yy = r.y + (int)y;
                if (zz < 0 || zz >= sz || yy < 0 || yy >= sy) continue;
                int x0 = std::max(r.x, 0), x1 = std::min(r.x + (int)r.w, sx);
                if (x1 <= x0) continue;
                size_t off = (size_t)x0 + (size_t)yy * sx + (size_t)zz * sx * sy;
                memcpy(image + off, staging + off, (size_t)(x1 - x0));
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic inputs.  Gradient noise after the published FastNoise 0.4 "Perlin"
// algorithm (Vendor/FastNoise/FastNoise.cpp:197-215 seed table, :826-880 3-D, :950-985 2-D),
// combined as Noise::GetTerrainNoise (Sources/Util/Noise.cpp:93-135).  Restated, not copied;
// cross-checked against the vendored library by oracle/refcheck (oracle/_ref).
// ---------------------------------------------------------------------------------------------
namespace {
struct Perm { uint8_t p[512], p12[512]; bool init = false; };
Perm g_perm;
const float GX[12] = {1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
const float GY[12] = {1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
const float GZ[12] = {0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};

void build_perm(int seed, uint8_t* p, uint8_t* p12) {
    std::mt19937_64 gen((unsigned long long)seed);
    for (int i = 0; i < 256; i++) p[i] = (uint8_t)i;
    for (int j = 0; j < 256; j++) {
        int rng = (int)(gen() % (uint64_t)(256 - j));
        int k = rng + j;
        uint8_t l = p[j];
        p[j] = p[j + 256] = p[k];
        p[k] = l;
        p12[j] = p12[j + 256] = (uint8_t)(p[j] % 12);
    }
}
inline const Perm& perm() {
    if (!g_perm.init) { build_perm(1337, g_perm.p, g_perm.p12); g_perm.init = true; }
    return g_perm;
}
inline int fast_floor(float f) { return f >= 0 ? (int)f : (int)f - 1; }
inline float quintic(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float grad2(const Perm& P, int x, int y, float xd, float yd) {
    uint8_t i = P.p12[(x & 0xff) + P.p[(y & 0xff) + 0]];
    return xd * GX[i] + yd * GY[i];
}
inline float grad3(const Perm& P, int x, int y, int z, float xd, float yd, float zd) {
    uint8_t i = P.p12[(x & 0xff) + P.p[(y & 0xff) + P.p[(z & 0xff) + 0]]];
    return xd * GX[i] + yd * GY[i] + zd * GZ[i];
}
float perlin2(const Perm& P, float x, float y) {
    int x0 = fast_floor(x), y0 = fast_floor(y);
    int x1 = x0 + 1, y1 = y0 + 1;
    float xs = quintic(x - (float)x0), ys = quintic(y - (float)y0);
    float xd0 = x - (float)x0, yd0 = y - (float)y0;
    float xd1 = xd0 - 1.0f, yd1 = yd0 - 1.0f;
    float xf0 = lerp(grad2(P, x0, y0, xd0, yd0), grad2(P, x1, y0, xd1, yd0), xs);
    float xf1 = lerp(grad2(P, x0, y1, xd0, yd1), grad2(P, x1, y1, xd1, yd1), xs);
    return lerp(xf0, xf1, ys);
}
float perlin3(const Perm& P, float x, float y, float z) {
    int x0 = fast_floor(x), y0 = fast_floor(y), z0 = fast_floor(z);
    int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    float xs = quintic(x - (float)x0), ys = quintic(y - (float)y0), zs = quintic(z - (float)z0);
    float xd0 = x - (float)x0, yd0 = y - (float)y0, zd0 = z - (float)z0;
    float xd1 = xd0 - 1.0f, yd1 = yd0 - 1.0f, zd1 = zd0 - 1.0f;
    float xf00 = lerp(grad3(P, x0, y0, z0, xd0, yd0, zd0), grad3(P, x1, y0, z0, xd1, yd0, zd0), xs);
    float xf10 = lerp(grad3(P, x0, y1, z0, xd0, yd1, zd0), grad3(P, x1, y1, z0, xd1, yd1, zd0), xs);
    float xf01 = lerp(grad3(P, x0, y0, z1, xd0, yd0, zd1), grad3(P, x1, y0, z1, xd1, yd0, zd1), xs);
    float xf11 = lerp(grad3(P, x0, y1, z1, xd0, yd1, zd1), grad3(P, x1, y1, z1, xd1, yd1, zd1), xs);
    float yf0 = lerp(xf00, xf10, ys);
    float yf1 = lerp(xf01, xf11, ys);
    return lerp(yf0, yf1, zs);
}
const float FREQ = 0.01f;   // FastNoise.h:221
// Noise.cpp:93-121 GetOctave
float octave2(const Perm& P, float x, float y, int octaves) {
    float total = 0.0f, frequency = 1.0f, amplitude = 1.0f, maxValue = 0.0f;
    for (int i = 0; i < octaves; i++) {
        total += perlin2(P, (x * frequency) * FREQ, (y * frequency) * FREQ) * amplitude;
        maxValue += amplitude; amplitude *= 0.5f; frequency *= 2.0f;
    }
    return total / maxValue;
}
float octave3(const Perm& P, float x, float y, float z, int octaves) {
    float total = 0.0f, frequency = 1.0f, amplitude = 1.0f, maxValue = 0.0f;
    for (int i = 0; i < octaves; i++) {
        total += perlin3(P, (x * frequency) * FREQ, (y * frequency) * FREQ, (z * frequency) * FREQ) * amplitude;
        maxValue += amplitude; amplitude *= 0.5f; frequency *= 2.0f;
    }
    return total / maxValue;
}
// Noise.cpp:131-135 with info = {Bias2D 0, Frequency2D 1, Octaves2D 4, Bias3D 0, Frequency3D 2, Octaves3D 3}
inline float terrain2d(const Perm& P, float x, float z) { return octave2(P, x * 1.0f, z * 1.0f, 4) + 0.0f; }
inline float terrain3d(const Perm& P, float x, float y, float z) { return octave3(P, x * 2.0f, y * 2.0f, z * 2.0f, 3) + 0.0f; }
}  // namespace

void vxo_perm_table(int seed, uint8_t* perm512, uint8_t* perm12_512) { build_perm(seed, perm512, perm12_512); }

float vxo_terrain_noise(float x, float y, float z) {
    const Perm& P = perm();
    float v = terrain2d(P, x, z);
    v += terrain3d(P, x, y, z);
    return v;
}

// voxel (x,y,z) solid iff GetTerrainNoise(x,y,z) > (y/NY - 0.5)*2, NY = 2*sy voxels.
void vxo_gen_terrain(uint8_t* data, int sx, int sy, int sz) {
    const Perm& P = perm();
    const float NY = (float)(2 * sy);
#pragma omp parallel for schedule(dynamic, 1)
    for (int tz = 0; tz < sz; ++tz) {
        std::vector<float> col2((size_t)4 * sx);   // 2-D term for the 2x2 voxel columns of each texel
        for (int tx = 0; tx < sx; ++tx)
            for (int b = 0; b < 4; ++b)
                col2[(size_t)tx * 4 + b] = terrain2d(P, (float)(2 * tx + (b & 1)), (float)(2 * tz + (b >> 1)));
        for (int ty = 0; ty < sy; ++ty)
            for (int tx = 0; tx < sx; ++tx) {
                unsigned byte = 0;
                for (int bit = 0; bit < 8; ++bit) {
                    int vx = 2 * tx + (bit & 1), vy = 2 * ty + ((bit >> 1) & 1), vz = 2 * tz + (bit >> 2);
                    float v = col2[(size_t)tx * 4 + ((bit & 1) | ((bit >> 2) << 1))];
                    v += terrain3d(P, (float)vx, (float)vy, (float)vz);
                    float thr = ((float)vy / NY - 0.5f) * 2.0f;
                    if (v > thr) byte |= 1u << bit;
                }
                data[(size_t)tx + (size_t)ty * sx + (size_t)tz * sx * sy] = (uint8_t)byte;
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic G-buffer (SURVEY 8d "G-buffer"): primary rays through the packed volume with the A4
// DDA arithmetic (unbounded step count, clipped to the volume box first).  Input synthesis, not a
// reference pass; the encodings are the reference's (GeometryVoxel.frag:156,169; GeometrySky.frag:30).
// ---------------------------------------------------------------------------------------------
void vxo_gbuffer_primary(const vxo_volume* vol, const vxo_view* view, int W, int H,
                         uint32_t* depth24, uint32_t* normal, uint32_t* material) {
    const V3 box{(float)(vol->sx * 2), (float)(vol->sy * 2), (float)(vol->sz * 2)};
    const int max_steps = 2 * (vol->sx + vol->sy + vol->sz) * 2 + 8;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t idx = (size_t)py * W + px;
            Pixel p = pixel_setup(*view, W, H, px, py);
            V3 dir = normalize3(xyz(mat_mul(view->InverseViewMatrix, V4{p.farvec.x, p.farvec.y, p.farvec.z, 0.0f})));
            V3 org = v3(view->InverseViewMatrix[12], view->InverseViewMatrix[13], view->InverseViewMatrix[14]) * 10.0f;
            uint32_t od = 0xFFFFFFu, on = 0u, om = 0u;
            // slab clip when the eye is outside the box
            float t0 = 0.0f;
            bool inside = org.x >= 0.0f && org.y >= 0.0f && org.z >= 0.0f && org.x < box.x && org.y < box.y && org.z < box.z;
            bool ok = true;
            if (!inside) {
                float tmin = 0.0f, tmax = 3.0e38f;
                const float o[3] = {org.x, org.y, org.z}, d[3] = {dir.x, dir.y, dir.z}, b[3] = {box.x, box.y, box.z};
                for (int a = 0; a < 3; ++a) {
                    if (d[a] == 0.0f) { if (o[a] < 0.0f || o[a] >= b[a]) ok = false; continue; }
                    float ta = (0.0f - o[a]) / d[a], tb = (b[a] - o[a]) / d[a];
                    float lo = fminf(ta, tb), hi = fmaxf(ta, tb);
                    tmin = fmaxf(tmin, lo); tmax = fminf(tmax, hi);
                }
                if (!(tmin <= tmax)) ok = false;
                t0 = tmin + 0.001f;
            }
            if (ok) {
                V3 o = org + dir * t0;
                V3 stepSign{gsign(dir.x), gsign(dir.y), gsign(dir.z)};
                V3 t_delta = v3(1.0f, 1.0f, 1.0f) / (dir * stepSign);
                I3 cur{f2i(floorf(o.x)), f2i(floorf(o.y)), f2i(floorf(o.z))};
                V3 nb = v3((float)cur.x, (float)cur.y, (float)cur.z) + (stepSign * 0.5f + v3(0.5f, 0.5f, 0.5f));
                V3 t_max = (nb - o) / dir;
                float t_enter = 0.0f;
                V3 nrm{0.0f, 1.0f, 0.0f};
                for (int n = 0; n < max_steps; ++n) {
                    if (cur.x < 0 || cur.y < 0 || cur.z < 0 || cur.x >= vol->sx * 2 || cur.y >= vol->sy * 2 || cur.z >= vol->sz * 2) break;
                    if (get_volume_at(*vol, cur, 0)) {
                        V3 hit = o + dir * t_enter;                 // voxel units
                        V3 hw = hit * 0.1f;                         // world units
                        V4 pv = mat_mul(view->ViewMatrix, V4{hw.x, hw.y, hw.z, 1.0f});
                        float w = -pv.z;                            // GL perspective: w_clip = -z_view
                        float dlin = (w - NEAR_) / (FAR_ - NEAR_);  // GeometryVoxel.frag:169
                        dlin = gclamp(dlin, 0.0f, 1.0f);
                        od = (uint32_t)f2i(floorf(dlin * 16777215.0f + 0.5f));
                        int qx = f2i(floorf(nrm.x * 127.0f + 0.5f)), qy = f2i(floorf(nrm.y * 127.0f + 0.5f)), qz = f2i(floorf(nrm.z * 127.0f + 0.5f));
                        on = ((uint32_t)(uint8_t)(int8_t)qx) | ((uint32_t)(uint8_t)(int8_t)qy << 8) | ((uint32_t)(uint8_t)(int8_t)qz << 16);
                        uint32_t rough = (uint32_t)((cur.x * 7 + cur.y * 13 + cur.z * 29) & 255);
                        om = rough | (0u << 8) | (0u << 16) | (255u << 24);
                        break;
                    }
                    V3 select{gstep(t_max.x, t_max.z) * gstep(t_max.x, t_max.y),
                              gstep(t_max.y, t_max.x) * gstep(t_max.y, t_max.z),
                              gstep(t_max.z, t_max.y) * gstep(t_max.z, t_max.x)};
                    t_enter = dot3(t_max, select);
                    if (t_enter != t_enter) break;   // NaN: axis-parallel ray (0*inf); treat as a miss
                    nrm = (stepSign * -1.0f) * select;
                    V3 adv = select * stepSign;
                    cur.x += f2i(adv.x); cur.y += f2i(adv.y); cur.z += f2i(adv.z);
                    t_max = t_max + t_delta * select;
                }
            }
            depth24[idx] = od; normal[idx] = on; material[idx] = om;
        }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Debug exports so tests can compare the pinned operation orders with glm (oracle/refcheck).
// ---------------------------------------------------------------------------------------------
extern "C" {
void vxo_dbg_normalize(const float* v, float* out) { V3 r = normalize3(V3{v[0], v[1], v[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void vxo_dbg_mix(const float* a, const float* b, float t, float* out) { V3 r = mix3(V3{a[0], a[1], a[2]}, V3{b[0], b[1], b[2]}, t); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void vxo_dbg_cross(const float* a, const float* b, float* out) { V3 r = cross3(V3{a[0], a[1], a[2]}, V3{b[0], b[1], b[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
float vxo_dbg_dot(const float* a, const float* b) { return dot3(V3{a[0], a[1], a[2]}, V3{b[0], b[1], b[2]}); }
void vxo_dbg_matvec(const float* m, const float* v, float* out) { V4 r = mat_mul(m, V4{v[0], v[1], v[2], v[3]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
void vxo_dbg_reflect(const float* i, const float* n, float* out) {
    V3 I{i[0], i[1], i[2]}, N{n[0], n[1], n[2]};
    V3 r = I - N * dot3(N, I) * 2.0f;
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float vxo_dbg_mod(float x, float y) { return gmod(x, y); }
float vxo_dbg_smoothstep(float e0, float e1, float x) { return gsmoothstep(e0, e1, x); }
// o, dx, dy, dz of ShadowVoxSystem.cpp:134-140 -> out[12]
void vxo_dbg_basis(const float* m, const float* pivot, float* out) {
    Basis b = basis_from(m, pivot, true);
    const V3 v[4] = {b.o, b.dx, b.dy, b.dz};
    for (int i = 0; i < 4; ++i) { out[i * 3] = v[i].x; out[i * 3 + 1] = v[i].y; out[i * 3 + 2] = v[i].z; }
}
void vxo_dbg_hemisphere(uint32_t nx, uint32_t ny, float* out) { V3 r = cosine_sample_hemisphere(luts(), nx, ny); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void vxo_dbg_luts(float* cos256, float* sin256) { memcpy(cos256, luts().cosT, 1024); memcpy(sin256, luts().sinT, 1024); }
}
