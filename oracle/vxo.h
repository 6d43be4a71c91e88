/* oracle/vxo.h -- CPU ORACLE for the voxel-lighting hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain C++ restatement of the reference's GLSL traversal / lighting shaders and of
 * ShadowVoxSystem (carloshgsilva/VoxelEngine @ 08d3b81).  It exists to CHECK the CUDA product
 * (voxelengine_b200/csrc) and to be timed as the CPU baseline.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.  The product
 * never links, includes or calls anything in this directory.
 *
 * PARITY PINNED AGAINST THE REFERENCE ITSELF, RUN HERE.  The reference has no tests, golden vectors
 * or fixtures for this path and its Vulkan build cannot run in this image, but its sources for the
 * path compile on the host from where they lie under /root/reference (recipes under oracle/refcheck,
 * outputs only under oracle/_ref/):
 *   - Sources/Shaders/lib/Light.frag and Light{Ambient,Point,Spot,Reflection}.frag as C++ on the
 *     reference's vendored glm (refcheck/build_shaders.py -> _ref/libvxshader.so);
 *   - Sources/World/Systems/ShadowVoxSystem.cpp with the reference's vendored entt + glm
 *     (refcheck/shadowvox_wrap.cpp -> _ref/libvxshadowvox.so);
 *   - glm / FastNoise / Sources/Util/Noise.cpp (refcheck/ref_wrap.cpp -> _ref/libvxref.so).
 * tests/test_oracle_shaders.py and tests/test_oracle_refcheck.py check this oracle against them bit for
 * bit (distance, probe count, hit texel, DDA hit/normal, every plane of the four passes, voxelised bytes,
 * dirty regions), and tests/golden/ref_shaders.npz carries their outputs to machines without the
 * reference tree.  What stays a declared definition (no reference source exists for it): the
 * hardware's fixed-point texel decode, In.FarVec evaluated per pixel instead of interpolated, and
 * cos/sin as correctly rounded values.
 *
 * All paths cited are relative to /root/reference/.
 */
#ifndef VXO_H
#define VXO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-ray record of the level-1 (traversal-given-rays) parity interface. 48 bytes. */
typedef struct vxo_hit {
    float   t;        /* march: returned distance d (== dist on a miss). DDA: best_t at exit   */
    int32_t steps;    /* number of occupancy probes performed (DDA: nt + probes of the hit)    */
    int32_t vx, vy, vz; /* probed voxel on a hit (phase 1: texel*2+bit; phase 2: ivec3(pos);
                           DDA: current_voxel); 0 on a miss                                    */
    int32_t status;   /* 0 miss; 1 hit fine bit (phase 1) / DDA hit; 2 hit coarse byte (phase 2);
                         3 DDA left the inclusive bounds [0, 2*dim+1]                          */
    float   px, py, pz; /* DDA: hit position (origin + dir*best_t); march: pos at the hit probe */
    float   nx, ny, nz; /* DDA: face normal -stepSign*select; march: 0                          */
} vxo_hit;

/* 8 floats per ray: origin xyz, dir xyz, dist (maxt for the DDA), unused. */
typedef struct vxo_ray { float ox, oy, oz, dx, dy, dz, dist, pad; } vxo_ray;

enum { VXO_SPARSE = 0, VXO_SUPERSPARSE = 1, VXO_DDA = 2 };

/* Sources/Graphics/Renderer/View.h:16-30 -- byte-identical (380 B). Column-major mat4. */
typedef struct vxo_view {
    float LastViewMatrix[16], ViewMatrix[16], InverseViewMatrix[16];
    float ProjectionMatrix[16], InverseProjectionMatrix[16];
    float Res[2], iRes[2];
    float CameraPosition[3]; int32_t _pad0;
    float Jitter[2];
    int32_t Frame;
    int32_t ColorTextureRID, DepthTextureRID, PalleteColorRID, PalleteMaterialRID;
} vxo_view;

/* Sources/Graphics/Pipelines/LightPointPipeline.h:20-25 (32 B) */
typedef struct vxo_point_light { float Position[3], Range, Color[3], Attenuation; } vxo_point_light;
/* Sources/Graphics/Pipelines/LightSpotPipeline.h:19-28 (64 B) */
typedef struct vxo_spot_light {
    float Position[3], Range, Color[3], Attenuation, Direction[3], Angle, AngleAttenuation, _pad[3];
} vxo_spot_light;

/* G-buffer planes, one element per pixel, row 0 at the top of the image.
 *  depth24 : D24 unorm in the low 24 bits (Sources/Graphics/Graphics.h:59)
 *  normal  : R8G8B8A8_SNORM, world space (Graphics.h:56)
 *  material: UNORM8 x4, .r roughness .g metallic .b emit (GeometryVoxel.frag:154)
 *  noise   : 512x512 RGBA8 blue noise                                                    */
typedef struct vxo_gbuffer {
    int32_t width, height;
    const uint32_t* depth24;
    const uint32_t* normal;    /* bytes x,y,z,w little-endian packed */
    const uint32_t* material;  /* bytes r,g,b,a */
    const uint32_t* noise;     /* 512*512 */
} vxo_gbuffer;

typedef struct vxo_volume { const uint8_t* data; int32_t sx, sy, sz; } vxo_volume;

/* rays generated, occupancy probes performed, lit (ray-generating) pixels */
typedef struct vxo_stats { uint64_t rays, steps, pixels; } vxo_stats;

/* rows processed: row_begin, row_begin+row_step, ... < row_end */
typedef struct vxo_rows { int32_t begin, end, step; } vxo_rows;

typedef struct vxo_region { int32_t x, y, z; uint32_t w, h, d; int32_t mip; } vxo_region;

/* One voxelisation command == one visited entity of ShadowVoxSystem::OnUpdate (flags=0) or one
 * OnVoxDestroyed callback (flags=VXO_ENT_DESTROY: clear at current matrix, pivot ignored). */
typedef struct vxo_entity {
    int32_t model;          /* index into the model table */
    int32_t flags;
    float   prev[16];       /* Transform::PreviousWorldMatrix */
    float   cur[16];        /* Transform::WorldMatrix */
    float   pivot[3];       /* VoxRenderer::Pivot */
    int32_t _pad;
} vxo_entity;
enum { VXO_ENT_DESTROY = 1 };

typedef struct vxo_model { const uint8_t* voxels; int32_t sx, sy, sz; } vxo_model;

int  vxo_num_threads(void);
void vxo_set_num_threads(int n);

void vxo_trace_rays(const vxo_volume* vol, const vxo_ray* rays, int64_t n, int variant, vxo_hit* out);

void vxo_pass_ambient(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb, int n_ao,
                      vxo_rows rows, float* out_shadow, float* out_ao, vxo_stats* stats);
void vxo_pass_point(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb,
                    const vxo_point_light* lights, int n_lights, vxo_rows rows,
                    float* out_shadow /* [n_lights][H][W] */, vxo_stats* stats);
void vxo_pass_spot(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb,
                   const vxo_spot_light* lights, int n_lights, vxo_rows rows,
                   float* out_shadow, vxo_stats* stats);
void vxo_pass_reflection(const vxo_volume* vol, const vxo_view* view, const vxo_gbuffer* gb,
                         vxo_rows rows, float* out_t, vxo_stats* stats);

/* the colour the passes add to the light buffer (float32 RGBA, before the RGBA16F attachment
 * conversion), from the planes above + COLOR_TEXTURE (albedo RGBA8 UNORM).  Tolerance parity (pow). */
void vxo_resolve_ambient(const vxo_view* view, const vxo_gbuffer* gb, const uint32_t* albedo_rgba8, const float* shadow,
                         const float* ao, vxo_rows rows, float* out_rgba /* [H][W][4] */);
void vxo_resolve_local(const vxo_view* view, const vxo_gbuffer* gb, const uint32_t* albedo_rgba8, const float* lights,
                       int n_lights, int spot, const float* shadow /* [n][H][W] */, vxo_rows rows, float* inout_rgba);

/* LightTAA.frag (temporal + spatial accumulation of the light buffer) and the colour LightReflection.frag
 * writes around its march.  Full-frame row-major planes; light / motion planes are float32 (values before the RGBA16F / RG16F
 * attachment conversion); nearest sampling, out of range reads 0.  vxo_light_taa is bit-exact arithmetic (cos / sin correctly
 * rounded); vxo_resolve_reflection carries pow (tolerance parity). */
void vxo_light_taa(const vxo_view* view, const vxo_gbuffer* gb, const uint32_t* albedo_rgba8, const float* motion /* [H][W][2] */,
                   const float* light /* [H][W][4] */, const float* last_light /* [H][W][4] */, vxo_rows rows, float* out_rgba);
void vxo_resolve_reflection(const vxo_view* view, const vxo_gbuffer* gb, const float* t_plane, const float* light /* [H][W][4] or NULL */,
                            const float* sky_rgb /* [3] */, vxo_rows rows, float* out_rgba);

/* VoxAsset's mip rule and the hierarchical-mip DDA of GeometryVoxel.frag on one model volume. */
typedef struct vxo_model_ray { float cam[3], dir[3], uv[2]; } vxo_model_ray;
typedef struct vxo_model_hit { int32_t hit; uint32_t material; int32_t fetches, steps; float pos[3], normal[3]; } vxo_model_hit;
void vxo_model_mip(const uint8_t* parent, int psx, int psy, int psz, uint8_t* out);
void vxo_trace_model_rays(const uint8_t* mip0, const uint8_t* mip1, const uint8_t* mip2, int sx, int sy, int sz,
                          const vxo_model_ray* rays, int64_t n, int frame, float res_x, float res_y, vxo_model_hit* out);

/* One fragment of GeometryVoxel.frag's main(): GeometryVoxelPipeline::Cmd (Pipelines/GeometryVoxelPipeline.h:30-36) + the
 * interpolated inputs of the fragment stage in, G-buffer outputs (floats, before attachment conversion) out. */
typedef struct vxo_vox_cmd { float WorldMatrix[16], LastWorldMatrix[16]; int32_t VolumeRID, PalleteIndex, _pad[2]; } vxo_vox_cmd;
typedef struct vxo_frag_in { float cam[3], dir[3], mvp[16]; } vxo_frag_in;          /* In.localCameraPos, In.localDirection, In.MVPMatrix */
typedef struct vxo_frag_out { int32_t hit; uint32_t material_index; int32_t fetches, _pad; float color[4], normal[4], material[4], motion[2], depth, _pad2; } vxo_frag_out;
void vxo_geometry_fragment(const vxo_model* model_mips /* [3]: levels 0, 1, 2 */, const vxo_view* view, const vxo_vox_cmd* cmd,
                           const uint32_t* pal_color /* [palettes][256] RGBA8 */, const uint32_t* pal_material, const vxo_frag_in* in,
                           int64_t n, vxo_frag_out* out);

/* The geometry pass over a draw list: per pixel the nearest model hit (depth test LESS on D24), outputs in the attachment formats.
 * cmds[c]._pad[0] = index of the draw's model in `mips` ([n_models][3]). */
void vxo_gbuffer_models(const vxo_view* view, int W, int H, const vxo_vox_cmd* cmds, int n_cmds, const vxo_model* mips,
                        const uint32_t* pal_color, const uint32_t* pal_material, uint32_t* depth24, uint32_t* normal, uint32_t* material,
                        uint32_t* albedo, float* motion);

/* ShadowVoxSystem::SetVolumeAt / OnUpdate / OnVoxDestroyed on a host staging buffer. */
void vxo_set_volume_at(uint8_t* data, int sx, int sy, int sz, int x, int y, int z, int value);
int  vxo_get_volume_at(const vxo_volume* vol, int x, int y, int z, int mip);
void vxo_voxelize(uint8_t* data, int sx, int sy, int sz, const vxo_model* models,
                  const vxo_entity* ents, int n, vxo_region* out_regions, int32_t* out_valid);
/* CmdBuffer::copy(buffer,image,regions) addressing (Vendor/evk/evk.cpp:759-780) */
void vxo_upload_regions(uint8_t* image, const uint8_t* staging, int sx, int sy, int sz,
                        const vxo_region* regions, int n);

/* ---- synthetic inputs: not reference path arithmetic, but shared test inputs ---- */
void vxo_perm_table(int seed, uint8_t* perm512, uint8_t* perm12_512);
float vxo_terrain_noise(float x, float y, float z);
void vxo_gen_terrain(uint8_t* data, int sx, int sy, int sz);
void vxo_gbuffer_primary(const vxo_volume* vol, const vxo_view* view, int width, int height,
                         uint32_t* depth24, uint32_t* normal, uint32_t* material);

#ifdef __cplusplus
}
#endif
#endif
