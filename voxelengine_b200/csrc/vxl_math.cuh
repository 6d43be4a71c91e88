// vxl_math.cuh -- device arithmetic with pinned operation orders.
//
// The whole library is compiled with --fmad=false (no FMA contraction), default -prec-div /
// -prec-sqrt (IEEE division and square root) and no flush-to-zero, so that every value below is
// the IEEE-754 single-precision result of the operation sequence the reference shaders spell out,
// in the association order of glm 0.9.9.9 (the reference's host math library):
//   dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z          Vendor/glm/detail/func_geometric.inl:48-55
//   normalize(v) = v * (1/sqrt(dot(v,v)))             func_geometric.inl:82-90
//   mix(x,y,a) = x*(1-a) + y*a                        func_common.inl:81-89
//   mod(x,y) = x - y*floor(x/y)                       func_common.inl:212-219
//   M*v = (M0*v0 + M1*v1) + (M2*v2 + M3*v3)           type_mat4x4.inl:561-572
// float -> int is cvt.rzi.s32.f32 (truncate, saturate, NaN -> 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vxl {

// __host__ too: tests/emul compiles the traversal for the host so its logic can be checked
// without a GPU (test infrastructure; the product only ever runs the device instantiation).
#define VXL_DI __host__ __device__ __forceinline__

VXL_DI float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
VXL_DI float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
VXL_DI float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
VXL_DI float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
VXL_DI float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }

VXL_DI float dot3(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
VXL_DI float3 normalize3(float3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return v * inv; }
VXL_DI float length3(float3 v) { return sqrtf(dot3(v, v)); }
VXL_DI float3 mix3(float3 x, float3 y, float a) { float ia = 1.0f - a; return x * ia + y * a; }
VXL_DI float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
VXL_DI float gmod(float x, float y) { return x - y * floorf(x / y); }
VXL_DI float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
VXL_DI float gstep(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
VXL_DI float gclamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
VXL_DI float gsmoothstep(float e0, float e1, float x) {
    float t = gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
VXL_DI int f2i(float x) {
#ifdef __CUDA_ARCH__
    return __float2int_rz(x);
#else   // cvt.rzi.s32.f32 semantics on the host
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return -2147483647 - 1;
    return (int)x;
#endif
}
template <typename T> VXL_DI T ldg(const T* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// column-major mat4 * vec4
VXL_DI float4 mat_mul(const float* __restrict__ m, float4 v) {
    float4 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * v.w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * v.w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w);
    r.w = (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w);
    return r;
}
VXL_DI float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }

// fixed-point decoders (Vulkan conversion rules)
VXL_DI float unorm24(uint32_t d) { return (float)(d & 0xFFFFFFu) / 16777215.0f; }
VXL_DI float unorm8(uint32_t c) { return (float)(c & 0xFFu) / 255.0f; }
VXL_DI float snorm8(uint32_t c) { return fmaxf((float)(int)(signed char)(c & 0xFFu) / 127.0f, -1.0f); }
VXL_DI float3 decode_normal(uint32_t n) { return make_float3(snorm8(n), snorm8(n >> 8), snorm8(n >> 16)); }

}  // namespace vxl
