// oracle/refcheck/shader_log.h -- TEST INFRASTRUCTURE: call log of the reference's march functions (see build_shaders.py).
#pragma once
namespace vxref {
thread_local FetchLog g_fetch = {0, 0, 0, 0};
thread_local RayLog g_rays = {0, {}, {}};
thread_local bool g_discarded = false;

template <typename F>
inline float logged(int variant, glm::vec3 o, glm::vec3 d, float dist, F fn) {
    const long long before = g_fetch.count;
    const float r = fn(o, d, dist);
    if (g_rays.n < 4) {
        RayRecord& R = g_rays.r[g_rays.n];
        R.ox = o.x; R.oy = o.y; R.oz = o.z; R.dx = d.x; R.dy = d.y; R.dz = d.z; R.dist = dist; R.result = r;
        R.variant = variant; R.fetches = (int)(g_fetch.count - before); R.lx = g_fetch.x; R.ly = g_fetch.y;
        g_rays.lz[g_rays.n] = g_fetch.z;
    }
    g_rays.n++;
    return r;
}
}  // namespace vxref
