// oracle/refcheck/ref_wrap.cpp -- thin C wrapper around the REFERENCE's own vendored sources
// (glm 0.9.9.9, FastNoise, Sources/Util/Noise.cpp), compiled from where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libvxref.so.  Test infrastructure only:
// it lets tests/test_oracle_refcheck.py check the oracle's pinned operation orders and its
// restated terrain noise against the real library code.  No reference source is copied here.
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <FastNoise/FastNoise.h>
#include "Noise.h"
#include <cstring>

extern "C" {
void ref_normalize(const float* v, float* out) { glm::vec3 r = glm::normalize(glm::vec3(v[0], v[1], v[2])); memcpy(out, &r, 12); }
void ref_mix(const float* a, const float* b, float t, float* out) { glm::vec3 r = glm::mix(glm::vec3(a[0], a[1], a[2]), glm::vec3(b[0], b[1], b[2]), t); memcpy(out, &r, 12); }
void ref_cross(const float* a, const float* b, float* out) { glm::vec3 r = glm::cross(glm::vec3(a[0], a[1], a[2]), glm::vec3(b[0], b[1], b[2])); memcpy(out, &r, 12); }
float ref_dot(const float* a, const float* b) { return glm::dot(glm::vec3(a[0], a[1], a[2]), glm::vec3(b[0], b[1], b[2])); }
void ref_matvec(const float* m, const float* v, float* out) {
    glm::mat4 M; memcpy(&M, m, 64);
    glm::vec4 r = M * glm::vec4(v[0], v[1], v[2], v[3]); memcpy(out, &r, 16);
}
void ref_reflect(const float* i, const float* n, float* out) { glm::vec3 r = glm::reflect(glm::vec3(i[0], i[1], i[2]), glm::vec3(n[0], n[1], n[2])); memcpy(out, &r, 12); }
float ref_mod(float x, float y) { return glm::mod(x, y); }
float ref_smoothstep(float e0, float e1, float x) { return glm::smoothstep(e0, e1, x); }
// the o/dx/dy/dz extraction of Sources/World/Systems/ShadowVoxSystem.cpp:134-140, using glm
void ref_basis(const float* m, const float* pivot, float* out) {
    glm::mat4 M; memcpy(&M, m, 64);
    glm::vec3 P(pivot[0], pivot[1], pivot[2]);
    M = glm::translate(M, -P);
    glm::vec3 o = M[3];
    glm::vec3 dx = M[0] * 0.1f;
    glm::vec3 dy = M[1] * 0.1f;
    glm::vec3 dz = M[2] * 0.1f;
    memcpy(out, &o, 12); memcpy(out + 3, &dx, 12); memcpy(out + 6, &dy, 12); memcpy(out + 9, &dz, 12);
}
// voxel coordinate of model cell (x,y,z): ShadowVoxSystem.cpp:146-147
void ref_voxel_of(const float* basis12, int x, int y, int z, int* out) {
    glm::vec3 o(basis12[0], basis12[1], basis12[2]), dx(basis12[3], basis12[4], basis12[5]), dy(basis12[6], basis12[7], basis12[8]), dz(basis12[9], basis12[10], basis12[11]);
    glm::vec3 wp = o + (float)(x)*dx + (float)(y)*dy + (float)(z)*dz;
    glm::ivec3 fwp = glm::ivec3(wp * 10.0f);
    out[0] = fwp.x; out[1] = fwp.y; out[2] = fwp.z;
}
void ref_perspective(float fov, float aspect, float n, float f, float* out) { glm::mat4 P = glm::perspective(fov, aspect, n, f); memcpy(out, &P, 64); }
void ref_inverse(const float* m, float* out) { glm::mat4 M; memcpy(&M, m, 64); glm::mat4 I = glm::inverse(M); memcpy(out, &I, 64); }
// camera matrix of Sources/Editor/EditorCamera.cpp:62-67
void ref_camera(const float* pos, float yaw, float pitch, float* out) {
    glm::mat4 M = glm::identity<glm::mat4>();
    M = glm::translate(M, glm::vec3(pos[0], pos[1], pos[2]));
    M = glm::rotate(M, yaw, glm::vec3(0, 1, 0));
    M = glm::rotate(M, pitch, glm::vec3(1, 0, 0));
    memcpy(out, &M, 64);
}
// TransformSystem::RealculateMatrix (Sources/World/Systems/TransformSystem.cpp:124-135) with glm itself: matrix = T * Rz * Ry * Rx * S,
// world = parent * matrix
void ref_transform(const float* pos, const float* rot, const float* scl, const float* parent, float* matrix, float* world) {
    glm::mat4 M = glm::identity<glm::mat4>();
    M = glm::translate(M, glm::vec3(pos[0], pos[1], pos[2]));
    M = glm::rotate(M, rot[2], glm::vec3(0, 0, 1));
    M = glm::rotate(M, rot[1], glm::vec3(0, 1, 0));
    M = glm::rotate(M, rot[0], glm::vec3(1, 0, 0));
    M = glm::scale(M, glm::vec3(scl[0], scl[1], scl[2]));
    glm::mat4 P; memcpy(&P, parent, 64);
    glm::mat4 W = P * M;
    memcpy(matrix, &M, 64); memcpy(world, &W, 64);
}
float ref_terrain_noise(float x, float y, float z) {
    TerrainNoiseInfo info;
    info.Bias2D = 0; info.Scale2D = 1; info.Frequency2D = 1; info.Octaves2D = 4;
    info.Bias3D = 0; info.Scale3D = 1; info.Frequency3D = 2; info.Octaves3D = 3;
    return Noise::GetTerrainNoise(x, y, z, info);
}
float ref_perlin3(float x, float y, float z) { static FastNoise N; return N.GetPerlin(x, y, z); }
float ref_perlin2(float x, float y) { static FastNoise N; return N.GetPerlin(x, y); }
}
