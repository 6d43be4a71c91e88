/* include/vxl.h -- C ABI of the B200-native voxel-lighting pass (libvxl.so).
 *
 * Drop-in boundary for ONE path of carloshgsilva/VoxelEngine: the world occupancy volume built by
 * ShadowVoxSystem and the four light passes that ray-march it (sun shadow + ambient occlusion,
 * point-light shadows, spot-light shadows, specular occlusion).  Each entry point names the
 * reference interface it replaces; paths are relative to the reference tree.
 *
 *   reference (Vulkan, fragment shaders)                         this library (CUDA, sm_100a)
 *   ------------------------------------------------------------ -------------------------------
 *   ShadowVoxSystem::ShadowVoxSystem()                           vxl_volume_create
 *     Sources/World/Systems/ShadowVoxSystem.cpp:55-79
 *   CmdBuffer::copy(buffer, image, regions)                      vxl_volume_upload_regions
 *     Vendor/evk/evk.cpp:759-780 (called at ShadowVoxSystem.cpp:196)
 *   ShadowVoxSystem::OnUpdate / OnVoxDestroyed                   vxl_volume_voxelize
 *     ShadowVoxSystem.cpp:116-201, :7-53
 *   VoxAsset::Upload (model voxels -> GPU image)                 vxl_model_create
 *     Sources/Asset/VoxAsset.cpp:3-14
 *   LightAmbientPipeline::Use                                    vxl_pass_ambient
 *     Sources/Graphics/Pipelines/LightAmbientPipeline.h:35-52  -> Shaders/LightAmbient.frag:134-175
 *   LightPointPipeline::Use + DrawLight                          vxl_pass_point
 *     Pipelines/LightPointPipeline.h:58-100                    -> Shaders/LightPoint.frag:85-129
 *   LightSpotPipeline::Use + DrawLight                           vxl_pass_spot
 *     Pipelines/LightSpotPipeline.h:60-104                     -> Shaders/LightSpot.frag:73-117
 *   LightReflectionPipeline::Use                                 vxl_pass_reflection
 *     Pipelines/LightReflectionPipeline.h:34-51                -> Shaders/LightReflection.frag:60-113
 *   raycastShadowVolume{,Sparse,SuperSparse}                     vxl_trace_rays  (ray-level entry)
 *     Sources/Shaders/lib/Light.frag:29-81,131-173,175-217
 *   Graphics::Frame / Graphics::Transfer  (record, submit, WAIT) vxl_sync
 *     Sources/Graphics/Graphics.h:23-32
 *   whole "Lights" + "Reflection" timestamp blocks, host buffers vxl_lighting_host
 *     Sources/Graphics/Renderer/WorldRenderer.cpp:239-260,269-274
 *
 * Conventions
 *   - every function returns 0 on success, a negative vxl_status otherwise; nothing throws;
 *     vxl_last_error_string() describes the last failure on the calling thread.
 *   - unless a parameter is documented as HOST, data pointers are DEVICE pointers on the context's
 *     GPU and the call is asynchronous on the context's stream (stream-ordered); vxl_sync blocks.
 *   - one context per GPU, used from one host thread at a time.
 *   - there is no CPU fallback: without a CUDA device vxl_ctx_create fails with VXL_ERR_CUDA.
 *   - the new pass contract writes separate float planes where the reference folds shader locals
 *     into one RGBA16F light target:
 *        shadow   1 = lit, 0 = occluded                 (LightAmbient.frag:167-169 etc.)
 *        ao       mean_i((d_i/128)^2) * 0.05            (LightAmbient.frag:121-125)
 *        spec_t   reflection-ray distance t, 256 = miss (LightReflection.frag:113)
 *     pixels that generate no ray (sky, range-culled light) hold shadow = 1, ao = 0, spec_t = 256.
 */
#ifndef VXL_H
#define VXL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXL_ABI_VERSION 1
#define VXL_MAX_LIGHTS 64   /* LightPointPipeline.h:15, LightSpotPipeline.h:14 */

typedef enum vxl_status {
    VXL_OK = 0,
    VXL_ERR_INVALID = -1,   /* bad argument */
    VXL_ERR_CUDA = -2,      /* CUDA runtime / launch failure (incl. no device) */
    VXL_ERR_OOM = -3,
    VXL_ERR_LIMIT = -4      /* more than VXL_MAX_LIGHTS, etc. */
} vxl_status;

typedef struct vxl_ctx vxl_ctx;         /* one GPU + one stream */
typedef struct vxl_volume vxl_volume;   /* packed world occupancy volume + derived occupancy levels */

/* Sources/Graphics/Renderer/View.h:16-30 == GLSL ViewBuffer (lib/Common.frag:45-61). 380 bytes. */
typedef struct vxl_view {
    float LastViewMatrix[16], ViewMatrix[16], InverseViewMatrix[16];
    float ProjectionMatrix[16], InverseProjectionMatrix[16];   /* column-major (glm) */
    float Res[2], iRes[2];
    float CameraPosition[3]; int32_t _pad0;
    float Jitter[2];
    int32_t Frame;
    int32_t ColorTextureRID, DepthTextureRID, PalleteColorRID, PalleteMaterialRID;
} vxl_view;

/* LightPointPipeline.h:20-25 (32 B) and LightSpotPipeline.h:19-28 (64 B) */
typedef struct vxl_point_light { float Position[3], Range, Color[3], Attenuation; } vxl_point_light;
typedef struct vxl_spot_light {
    float Position[3], Range, Color[3], Attenuation, Direction[3], Angle, AngleAttenuation, _pad[3];
} vxl_spot_light;

/* evk ImageRegion as pushed by ShadowVoxSystem.cpp:189 (texel units) */
typedef struct vxl_region { int32_t x, y, z; uint32_t w, h, d; int32_t mip; } vxl_region;

/* One visited entity of ShadowVoxSystem::OnUpdate (flags = 0: clear at prev, set at cur, both
 * translated by -pivot) or one OnVoxDestroyed callback (VXL_ENT_DESTROY: clear at cur, pivot
 * ignored -- sic, ShadowVoxSystem.cpp:22-25).  Commands apply in array order, last writer wins. */
typedef struct vxl_entity {
    int32_t model;      /* id from vxl_model_create */
    int32_t flags;
    float   prev[16];   /* Transform::PreviousWorldMatrix (identity on an entity's first frame) */
    float   cur[16];    /* Transform::WorldMatrix */
    float   pivot[3];   /* VoxRenderer::Pivot */
    int32_t _pad;
} vxl_entity;
enum { VXL_ENT_DESTROY = 1 };

/* G-buffer + noise inputs of a (shard of a) frame.  Planes are "tile-compact": local tile i
 * (i < n_tiles) is global tile tile_first + i*tile_stride of the tile_w x tile_h grid laid
 * row-major over the width x height frame, stored as [n_tiles][tile_h][tile_w].  A whole frame on
 * one GPU is the single tile tile_w = width, tile_h = height (plain row-major, row 0 = top).
 *   depth24  D24 unorm in the low 24 bits, linear (w-NEAR)/(FAR-NEAR); sky = 0xFFFFFF
 *            (Graphics.h:59; GeometryVoxel.frag:169; GeometrySky.frag:30)
 *   normal   R8G8B8A8_SNORM world-space normal (Graphics.h:56; GeometryVoxel.frag:156)
 *   material UNORM8 .r roughness .g metallic .b emit (Graphics.h:57); only the spec pass reads it
 *   noise    512 x 512 RGBA8 blue noise (Assets/.../LDR_RGBA_0.png), NOT tiled                  */
typedef struct vxl_frame {
    int32_t width, height;
    int32_t tile_w, tile_h;
    int32_t tile_first, tile_stride, n_tiles;
    int32_t _pad;
    const uint32_t* depth24;
    const uint32_t* normal;
    const uint32_t* material;
    const uint32_t* noise;
} vxl_frame;

/* rays generated / occupancy probes performed by the reference algorithm / ray-generating pixels */
typedef struct vxl_stats { uint64_t rays, steps, pixels; } vxl_stats;

/* ray-level interface (level-1 parity): 32-byte rays in, 48-byte records out */
typedef struct vxl_ray { float ox, oy, oz, dx, dy, dz, dist, pad; } vxl_ray;
typedef struct vxl_hit {
    float t; int32_t steps; int32_t vx, vy, vz; int32_t status;
    float px, py, pz; float nx, ny, nz;
} vxl_hit;
enum { VXL_TRACE_SPARSE = 0, VXL_TRACE_SUPERSPARSE = 1, VXL_TRACE_DDA = 2 };

/* ---- library / context ------------------------------------------------------------------------ */
int         vxl_abi_version(void);
const char* vxl_last_error_string(void);
int vxl_ctx_create(int device, vxl_ctx** out);
int vxl_ctx_destroy(vxl_ctx* ctx);
/* use an existing cudaStream_t (e.g. torch's current stream) instead of the context's own */
int vxl_ctx_set_stream(vxl_ctx* ctx, void* cuda_stream);
int vxl_sync(vxl_ctx* ctx);
/* device counters accumulated by every pass since the last reset (vxl_stats_read synchronises) */
int vxl_stats_reset(vxl_ctx* ctx);
int vxl_stats_read(vxl_ctx* ctx, vxl_stats* out /* HOST */);
/* number of kernels this library launched on the context since creation (bench gpu_launches) */
int vxl_launch_count(vxl_ctx* ctx, uint64_t* out /* HOST */);

/* ---- peer memory: the fused output-tile gather (one process per GPU on one node) -------------------------------
 * The light-pass kernels can repeat every output store at (address + delta[i]): with each peer's copy of the gathered tile stack
 * mapped into this process through CUDA IPC, delta[i] = peer_i_base - own_base, and a rank's tiles land in every rank's stack by
 * peer-to-peer stores over NVLink while the pass is still running -- no separate collective moves the planes; a barrier closes
 * the frame.  vxl_ipc_export / _open / _close wrap cudaIpcGetMemHandle / OpenMemHandle / CloseMemHandle for buffers from
 * vxl_malloc (peer access is enabled lazily).  vxl_ctx_set_output_mirrors(ctx, 0, NULL) turns mirroring off; vxl_lighting_host
 * ignores it.  vxl_ctx_set_light_plane_stride: distance in pixels between consecutive light planes of vxl_pass_point / _spot
 * (0 = the shard's own n_tiles * tile_h * tile_w), so a rank can write straight into a stack padded to the largest shard. */
#define VXL_MAX_MIRRORS 15
typedef struct vxl_ipc_handle { unsigned char bytes[64]; } vxl_ipc_handle;
int vxl_ipc_export(vxl_ctx* ctx, void* dev, vxl_ipc_handle* out /* HOST */);
int vxl_ipc_open(vxl_ctx* ctx, const vxl_ipc_handle* handle /* HOST */, void** out_dev);
int vxl_ipc_close(vxl_ctx* ctx, void* dev);
int vxl_ctx_set_output_mirrors(vxl_ctx* ctx, int n, const int64_t* byte_deltas /* HOST */);
int vxl_ctx_set_light_plane_stride(vxl_ctx* ctx, uint64_t pixels);

/* ---- the frame sharded over the GPUs of one box, driven from C (SURVEY.md 8b "vxl_ctx_create(ndev) / vxl_gather", 8e) --------
 * One process (or thread) per GPU holds one vxl_group member; there is no reference counterpart (one GPU, one process), the call
 * sites served are WorldRenderer.cpp:239-274.  Each member owns one allocation -- n_stacks (1 or 2) copies of the gathered tile
 * stack of stack_bytes each plus a page of arrival flags -- that every other member maps:
 *   vxl_group_create            allocate this member (zeroed)
 *   vxl_group_handle            64-byte handle of the allocation, to be shipped to every peer by whatever the host has (pipe, file,
 *                               MPI, shared memory); vxl_group_connect takes all n_ranks handles (own slot ignored) and maps the peers.
 *                               Members inside ONE process exchange vxl_group_base pointers with vxl_group_connect_pointers instead
 *                               (CUDA IPC cannot open a handle in the process that exported it)
 *   vxl_group_begin_frame       point the context's output mirrors at the peers: stack (frame % n_stacks) of every copy; returns
 *                               this member's own stack of the frame, where its pass outputs must be placed (its slot of the
 *                               caller's [rank][plane][tile] layout).  The pass kernels then store every value into all copies
 *   vxl_group_fence             stream-ordered: completes when every member's kernels queued before ITS fence of this frame -- and
 *                               with them their peer stores -- have completed.  One small kernel per member: it writes the frame
 *                               number into its slot of each peer's flag page (st.release.sys) and waits for all slots of its own
 *                               (ld.acquire.sys).  No collective launch, no host round trip.  A peer that never arrives is reported
 *                               by vxl_group_status after 5 s instead of hanging the GPU
 *   vxl_group_end_frame         mirrors off
 * With n_stacks = 2 a member may consume frame N (stream-ordered behind its fence) while the peers already store frame N + 1 into
 * the other stack; every consumer of frame N must be queued on the context's stream before the passes of frame N + 1.
 * vxl_group_destroy closes the peer mappings and frees the allocation: the caller makes sure (barrier on its side) that no peer
 * still stores into it and that every peer has closed its mapping. */
typedef struct vxl_group vxl_group;
int vxl_group_create(vxl_ctx* ctx, int rank, int n_ranks, size_t stack_bytes, int n_stacks, vxl_group** out);
int vxl_group_handle(vxl_group* g, vxl_ipc_handle* out /* HOST */);
int vxl_group_connect(vxl_group* g, const vxl_ipc_handle* handles /* HOST [n_ranks] */);
int vxl_group_base(vxl_group* g, void** out_dev);
int vxl_group_connect_pointers(vxl_group* g, void* const* bases /* HOST [n_ranks], device pointers */);
int vxl_group_stack(vxl_group* g, int which, void** out_dev);
int vxl_group_begin_frame(vxl_group* g, uint64_t frame, void** out_stack_dev /* may be NULL */);
int vxl_group_fence(vxl_group* g);
int vxl_group_end_frame(vxl_group* g);
int vxl_group_status(vxl_group* g, int* out_rank_plus_one /* HOST: 0 = ok */);
int vxl_group_destroy(vxl_group* g);

/* diagnostics (no reference counterpart): kernel variant 0 = plain march on the volume bytes, 1 = march
 * against the per-block occupancy-bit tile in shared memory (default; also env VXL_VARIANT), 2 = variant 1
 * that also counts the probes that had to read the volume; all produce identical results.
 * vxl_debug_fetched_probes: that count since the last vxl_stats_reset (variant 2 only; the other probes
 * were answered by a clear occupancy bit). */
int vxl_debug_set_variant(vxl_ctx* ctx, int variant);
int vxl_debug_fetched_probes(vxl_ctx* ctx, uint64_t* out /* HOST */);
/* download one occupancy level unpacked to 0/1 bytes [cz][cy][cx]: level 1 = texels, 2, 3, 4 = plain (cell = 2^level
 * voxels), 13, 14 = the 3x3x3-dilated levels 3, 4 including their 1-cell border, 22 = level 2 decoded from its copy shifted
 * by 16 cells (what a TMA box at an odd origin reads); out_dims = {cx, cy, cz}; host_out may be NULL.
 * The levels are rebuilt lazily by the next pass (or vxl_volume_build_occupancy): in full after an upload, inside the
 * commands' voxel boxes only after a vxl_volume_voxelize call on levels that were up to date. */
int vxl_volume_debug_occupancy(vxl_volume* vol, int level, uint8_t* host_out /* HOST */, int* out_dims /* HOST[3] */);
/* measurement helper (bench.py's roofline denominators, SURVEY 8d: "L2 read bandwidth ... measured by a microbench in the bench
 * harness"): read a `bytes`-sized device buffer `reps` times with 16-byte loads from every SM (one warm-up pass first) and report
 * bytes * reps / CUDA-event time.  A buffer well below the 126 MB L2 measures L2 read bandwidth, one of several GB HBM. */
int vxl_debug_read_bandwidth(vxl_ctx* ctx, size_t bytes, int reps, double* out_gbs /* HOST */);

/* raw memory helpers so a C caller needs no CUDA headers */
int vxl_malloc(vxl_ctx* ctx, size_t bytes, void** out_dev);
int vxl_free(vxl_ctx* ctx, void* dev);
int vxl_host_alloc(size_t bytes, void** out_pinned);
int vxl_host_free(void* pinned);
int vxl_memcpy_h2d(vxl_ctx* ctx, void* dev, const void* host, size_t bytes);   /* async on stream */
int vxl_memcpy_d2h(vxl_ctx* ctx, void* host, const void* dev, size_t bytes);   /* async on stream */
int vxl_memset(vxl_ctx* ctx, void* dev, int value, size_t bytes);

/* ---- world occupancy volume (ShadowVoxSystem) -------------------------------------------------- */
/* sx,sy,sz in TEXELS (bytes); each byte packs 2x2x2 voxels, bit = (x&1)|(y&1)<<1|(z&1)<<2
 * (ShadowVoxSystem.cpp:82-94).  Created zero-filled like the reference constructor. */
int vxl_volume_create(vxl_ctx* ctx, int sx, int sy, int sz, vxl_volume** out);
int vxl_volume_destroy(vxl_volume* vol);
int vxl_volume_dims(const vxl_volume* vol, int* sx, int* sy, int* sz);
/* copy `n` regions from a HOST staging buffer of the full volume size, with the addressing of
 * CmdBuffer::copy: byte (x,y,z) at x + y*sx + z*sx*sy in both buffers */
int vxl_volume_upload_regions(vxl_volume* vol, const uint8_t* host_staging, const vxl_region* regions /* HOST */, int n);
int vxl_volume_upload(vxl_volume* vol, const uint8_t* host_bytes);   /* whole volume */
int vxl_volume_download(vxl_volume* vol, uint8_t* host_bytes);       /* synchronises */
int vxl_volume_clear(vxl_volume* vol);
/* device pointer of the canonical packed bytes (x fastest); writing through it requires a
 * following vxl_volume_mark_dirty */
int vxl_volume_device_ptr(vxl_volume* vol, uint8_t** out_dev);
int vxl_volume_mark_dirty(vxl_volume* vol);
/* (re)build the derived occupancy levels used to skip empty space.  Pure acceleration: never
 * changes a result.  Called implicitly by the passes when the volume is dirty. */
int vxl_volume_build_occupancy(vxl_volume* vol);

/* register a model: palette indices, x fastest (VoxAsset.h:52-56); 0 empty, 1..15 glass (not an
 * occluder), >= 16 solid (ShadowVoxSystem.cpp:145).  voxels is a HOST pointer. */
int vxl_model_create(vxl_ctx* ctx, const uint8_t* voxels, int sx, int sy, int sz, int* out_id);
/* apply `n` commands with the reference's sequential semantics.  ents, out_regions (n entries)
 * and out_valid (n entries; 0 = the reference would push no region) are HOST pointers; the two
 * outputs may be NULL.  Synchronises only when an output is requested. */
int vxl_volume_voxelize(vxl_volume* vol, const vxl_entity* ents, int n, vxl_region* out_regions, int32_t* out_valid);

/* ---- light passes ------------------------------------------------------------------------------ */
/* view/lights are HOST pointers (copied at call time); frame planes and outputs are DEVICE
 * pointers in the frame's tile-compact layout; outputs may be NULL to skip a plane.
 * n_ao = AO rays per pixel: 1 reproduces the reference; ray 0 uses getNoise(), ray i >= 1 uses
 * getNoise(i) (LightAmbient.frag:44-47). */
int vxl_pass_ambient(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame,
                     int n_ao, float* out_shadow, float* out_ao);
/* out_shadow: n_lights consecutive planes */
int vxl_pass_point(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame,
                   const vxl_point_light* lights, int n_lights, float* out_shadow);
int vxl_pass_spot(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame,
                  const vxl_spot_light* lights, int n_lights, float* out_shadow);
int vxl_pass_reflection(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame,
                        float* out_spec_t);
/* ---- light-buffer resolve -------------------------------------------------- */
/* The colour the reference's light passes add to the RGBA16F light buffer, i.e. what the fragment shaders compute
 * after the march (LightAmbient.frag:178-214 with calculateOcclusion :89-109; LightPoint.frag:131-152;
 * LightSpot.frag:118-138; lib/PBR.frag:48-69), as float32 RGBA (tile-compact, 16 B per pixel) BEFORE the attachment
 * conversion and blend.  Inputs: the frame's planes, COLOR_TEXTURE (albedo), and the shadow / ao planes of the march
 * passes above.  vxl_resolve_ambient writes (sky pixels, whose colour is a sky-box look-up outside this path, get 0);
 * vxl_resolve_point / _spot ADD the lights' colours in list order, like the reference's one additive draw per light.
 * pow() is specified by accuracy only, so these planes carry a tolerance (1e-5 relative), unlike the march planes. */
typedef struct vxl_resolve {
    const uint32_t* albedo;      /* DEVICE, tile-compact like the frame's planes: R8G8B8A8_UNORM colour attachment (Graphics.h:55) */
    const uint32_t* depth_full;  /* DEVICE, [height][width] D24 of the WHOLE frame: screenspaceOcclusion (LightAmbient.frag:66) samples
                                    other pixels.  NULL = the frame is one whole-frame tile and frame.depth24 is used */
} vxl_resolve;
int vxl_resolve_ambient(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_resolve* r,
                        const float* shadow, const float* ao, float* out_rgba);
int vxl_resolve_point(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_resolve* r,
                      const vxl_point_light* lights /* HOST */, int n_lights, const float* shadow /* [n_lights] planes */, float* inout_rgba);
int vxl_resolve_spot(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_resolve* r,
                     const vxl_spot_light* lights /* HOST */, int n_lights, const float* shadow, float* inout_rgba);

/* ---- after the light passes -------------------------------------------------- */
/* LightTAAPipeline::Use (Pipelines/LightTAAPipeline.h:34-53 -> Sources/Shaders/LightTAA.frag:37-141): temporal + spatial
 * accumulation of the light buffer; and the colour LightReflection.frag writes around its march (:60-139; the march itself is
 * vxl_pass_reflection).  Both sample OTHER pixels, so their inputs are whole-frame row-major planes; the output pixels and their
 * tile-compact layout come from `frame` as in every other pass (frame.depth24 / normal / material / noise: the shard's own
 * planes).  Nearest sampling (evk.cpp:277-293), out of range reads 0.  Light / motion planes are float32: the values before the
 * RGBA16F / RG16F attachment conversion.  vxl_light_taa is bit-exact against the oracle (cos / sin tabulated as correctly
 * rounded values); vxl_resolve_reflection carries pow() (1e-5 relative).  The sky box is a uniform colour `sky_rgb` (the cube
 * map is outside the path; NULL = black).  All plane pointers: DEVICE. */
typedef struct vxl_full_planes {
    const uint32_t* depth24;     /* [height][width] D24 */
    const uint32_t* normal;      /* R8G8B8A8_SNORM */
    const uint32_t* material;    /* RGBA8 UNORM */
    const uint32_t* albedo;      /* RGBA8 UNORM colour attachment */
    const float* motion;         /* [height][width][2] (GeometryVoxel.frag:166 out_Motion) */
    const float* light;          /* [height][width][4] the current light buffer (vxl_resolve_ambient + _point + _spot) */
    const float* last_light;     /* [height][width][4] the previous frame's vxl_light_taa output (alpha = variance) */
} vxl_full_planes;
int vxl_light_taa(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const vxl_full_planes* full, float* out_rgba /* tile-compact [4] */);
int vxl_resolve_reflection(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame, const float* spec_t /* tile-compact */,
                           const uint32_t* depth_full /* NULL: one whole-frame tile */, const float* light_full /* [h][w][4] TAA light, may be NULL */,
                           const float* sky_rgb /* HOST [3] or NULL */, float* out_rgba /* tile-compact [4] */);

/* ---- model traversal (core) -------------------------------------------------- */
/* The G-buffer producer's traversal of one model volume: VoxAsset::Upload's mip chain (Sources/Asset/VoxAsset.cpp:3-64,
 * built on the device on first use) and GeometryVoxel.frag's clipToAABB (:49-61) + intersectVolume (:64-125) -- the
 * reference's hierarchical-mip DDA with its LOD early accept and glass checkerboard.  Ray-level entry: one record per
 * fragment, in = (In.localCameraPos, In.localDirection, UV), out = (hit, hitMat, texel fetches, DDA steps, hitPos,
 * hitNormal).  frame / res_x / res_y: GetFrame() and GetRes() of the view.  rays / out are DEVICE pointers.
 * float -> int of a NaN (only reachable with an exactly zero direction component) follows cvt.rzi (0); GLSL leaves it undefined. */
typedef struct vxl_model_ray { float cam[3], dir[3], uv[2]; } vxl_model_ray;
typedef struct vxl_model_hit { int32_t hit; uint32_t material; int32_t fetches, steps; float pos[3], normal[3]; } vxl_model_hit;
int vxl_trace_model_rays(vxl_ctx* ctx, int model_id, const vxl_model_ray* rays, int64_t n, int frame, float res_x, float res_y,
                         vxl_model_hit* out);

/* The geometry pass over a draw list (GeometryVoxelPipeline::Use, Pipelines/GeometryVoxelPipeline.h:49-71; row f1): per pixel, in
 * list order, every model whose box the pixel's view ray enters from outside runs GeometryVoxel.frag's main(); depth test LESS on
 * D24; planes written in the attachment formats of Graphics.h:51-60 (sky: depth 0xFFFFFF, the rest 0).  vxl_vox_cmd is
 * GeometryVoxelPipeline::Cmd (:30-36) with PADDING[0] carrying the vxl model id; VolumeRID only feeds the anti-z-fight depth factor.
 * The fragment stage's interpolated inputs are evaluated at the pixel centre (In.localDirection = the view ray in model space).
 * cmds: HOST.  pal_color / pal_material: DEVICE [n_palettes][256] RGBA8 (PalleteAsset images).  out planes: DEVICE, tile-compact. */
typedef struct vxl_vox_cmd { float WorldMatrix[16], LastWorldMatrix[16]; int32_t VolumeRID, PalleteIndex, model, _pad; } vxl_vox_cmd;
typedef struct vxl_gbuffer_out { uint32_t* depth24; uint32_t* normal; uint32_t* material; uint32_t* albedo; float* motion /* [2] per pixel, may be NULL */; } vxl_gbuffer_out;
int vxl_gbuffer_models(vxl_ctx* ctx, const vxl_view* view, const vxl_frame* frame /* geometry only: sizes and tiles */, const vxl_vox_cmd* cmds,
                       int n_cmds, const uint32_t* pal_color, const uint32_t* pal_material, const vxl_gbuffer_out* out);

/* ---- on-disk formats --------------------------------------------------------- */
/* The files the reference's scenes are made of, read by host code (vxl_assets.cu); all paths are file-system paths, all buffers HOST.
 *   vxl_asset_guid         Assets::Hash (Sources/Asset/Assets.h:207-210): FNV-1a 64 of the asset path relative to Mods/
 *                          ("default/ModernHouse/0.v" -> 0x44B7A418296B6797, the GUID the shipped ModernHouse.pf stores)
 *   vxl_vox_file_read      VoxAsset::Serialize (Sources/Asset/VoxAsset.h:42-50): int32 dims[3] + dims[0]*dims[1]*dims[2] palette indices,
 *                          x fastest; out == NULL only fills dims
 *   vxl_model_load_v       the same, registered as a model (vxl_model_create)
 *   vxl_pallete_file_read  PalleteAsset::Serialize (Sources/Asset/PalleteAsset.h:63-69) + PalleteCache::UploadPallete
 *                          (Sources/Vox/PalleteCache.cpp:5-25): the 256 colour texels (r, g, b, 255) and material texels (roughness,
 *                          metallic, emit, 0) of one palette row, the layout vxl_gbuffer_models reads
 *   vxl_prefab_file_read   PrefabAsset::Spawn (Sources/Asset/PrefabAsset.cpp:30-141) + TransformSystem::RealculateMatrix
 *                          (Sources/World/Systems/TransformSystem.cpp:124-135): the entities of a .pf scene in file order with their local
 *                          and world matrices (T * Rz * Ry * Rx * S, glm arithmetic); cap == 0 only counts.  Nested prefab instances: error.
 *   vxl_scene_load         the same for `prefab_path` relative to a Mods directory, with nested instances expanded in place (the nested
 *                          root takes the instancing entity's Id, Name, Parent and components, PrefabAsset.cpp:47-56,87-139); GUIDs
 *                          resolve to the files under `mods_dir` by hashing their relative paths, like ModLoader */
enum { VXL_PF_TRANSFORM = 1, VXL_PF_VOX = 2, VXL_PF_LIGHT = 4, VXL_PF_INSTANCE = 8 };
typedef struct vxl_prefab_entity {
    int32_t  id, parent;             /* "Id"; parent = index into the returned array, -1 for a root */
    uint32_t has;                    /* VXL_PF_* */
    int32_t  light_type;             /* Light::Type: 0 point, 1 spot, 2 directional, 3 ambient (World/Components.h:39-44) */
    float    position[3], rotation[3], scale[3];
    float    pivot[3];
    float    matrix[16], world[16];  /* Transform::Matrix, Transform::WorldMatrix (column-major) */
    uint64_t vox_guid, pallete_guid;
    uint64_t instance_guid;          /* VXL_PF_INSTANCE: this entity is the root of that nested prefab (Components: Instance) */
    float    intensity, color[3], attenuation, range, angle, angle_attenuation;
    char     name[64];
} vxl_prefab_entity;
int vxl_asset_guid(const char* path, uint64_t* out);
int vxl_vox_file_read(const char* path, int32_t dims[3], uint8_t* out, uint64_t cap);
int vxl_model_load_v(vxl_ctx* ctx, const char* path, int* out_model_id);
int vxl_pallete_file_read(const char* path, uint32_t* color256, uint32_t* material256);
int vxl_prefab_file_read(const char* path, vxl_prefab_entity* out, int cap, int* n_out);
int vxl_scene_load(const char* mods_dir, const char* prefab_path, vxl_prefab_entity* out, int cap, int* n_out);

/* The MagicaVoxel .vox importer (Sources/Editor/Importer/VoxImporter.cpp:284-520, host code in vxl_voximport.cu): what the reference's
 * editor runs on a dropped .vox file to produce the .v / .p / .pf files above.
 *   vxl_vox_import / _memory  VoxImportContext::Import (:284-394) + CreateEntity (:397-476): the chunk loop, then the node tree as entities in
 *                             creation order (a group before its children) with Transform.Position in world units (voxel * 0.1, z-up -> y-up)
 *                             and one model per shape node, voxels re-oriented by the node's `_r` byte into a volume whose sizes are rounded
 *                             up to a multiple of 4 (VoxAsset.h:26-30).  Unnamed shapes are named "0", "1", ... in creation order.
 *   vxl_vox_scene_model       the .v contents of model `model` (dims + palette indices, x fastest); out == NULL only fills dims / name
 *   vxl_vox_scene_pallete     the .p contents: 256 x {r, g, b, a, roughness, metallic, emit}; `a` is uninitialised in the reference, 0 here
 *   vxl_vox_scene_write       VoxImporter::Import (:478-520) + Assets::CreateAsset + PrefabAsset::FromWorld: writes
 *                             <mods_dir>/<path>/<file_name>/<shape>.v, <mods_dir>/<path>/<file_name>/<file_name>.p and <mods_dir>/<path>/<file_name>.pf
 *                             (json11's dump format, fmt's "{}" floats, GUIDs = vxl_asset_guid of the paths relative to mods_dir); the root
 *                             entity is named <file_name>.
 * Pinned against the reference's own imports: Assets/Mods/default ships FarmHouse / ModernHouse / Player .vox next to the files its importer
 * wrote from them.  Malformed input returns VXL_ERR_INVALID. */
typedef struct vxl_vox_scene vxl_vox_scene;
typedef struct vxl_vox_import_entity {
    int32_t parent;                  /* index into the entity array, -1 for the root */
    int32_t model;                   /* index for vxl_vox_scene_model, -1 for a group */
    float   position[3];             /* Transform.Position; Rotation = 0, Scale = 1, Pivot = 0 */
    char    name[64];                /* the shape's name ("" for groups); the written .pf names the root <file_name> */
} vxl_vox_import_entity;
int vxl_vox_import(const char* vox_path, vxl_vox_scene** out);
int vxl_vox_import_memory(const void* data, uint64_t size, vxl_vox_scene** out);
int vxl_vox_scene_counts(const vxl_vox_scene* scene, int* n_entities, int* n_models);
int vxl_vox_scene_entities(const vxl_vox_scene* scene, vxl_vox_import_entity* out, int cap);
int vxl_vox_scene_model(const vxl_vox_scene* scene, int model, int32_t dims[3], char name[64], uint8_t* out, uint64_t cap);
int vxl_vox_scene_pallete(const vxl_vox_scene* scene, uint8_t records[1792]);
int vxl_vox_scene_write(const vxl_vox_scene* scene, const char* mods_dir, const char* path, const char* file_name);
int vxl_vox_scene_free(vxl_vox_scene* scene);

/* rays/out are DEVICE pointers */
int vxl_trace_rays(vxl_ctx* ctx, vxl_volume* vol, const vxl_ray* rays, int64_t n, int variant, vxl_hit* out);

/* ---- whole-frame drop-in with HOST buffers ----------------------------------------------------- */
/* What a renderer that keeps its G-buffer on the host side of the boundary calls once per frame:
 * H2D of the frame shard, ambient + point + spot + reflection passes, D2H of the output planes.
 * All pointers HOST (pinned memory recommended: vxl_host_alloc).  Planes use the frame's
 * tile-compact layout.  Any output pointer may be NULL (that pass is skipped when all of its
 * outputs are NULL).  Blocks until the outputs are in host memory. */
typedef struct vxl_lighting_host_args {
    vxl_frame frame;                 /* plane pointers are HOST pointers here */
    const vxl_view* view;
    int32_t n_ao;
    int32_t n_point, n_spot;
    const vxl_point_light* point;
    const vxl_spot_light* spot;
    float* out_shadow;               /* [tiles] */
    float* out_ao;                   /* [tiles] */
    float* out_point_shadow;         /* [n_point][tiles] */
    float* out_spot_shadow;          /* [n_spot][tiles] */
    float* out_spec_t;               /* [tiles] */
} vxl_lighting_host_args;
int vxl_lighting_host(vxl_ctx* ctx, vxl_volume* vol, const vxl_lighting_host_args* args);
/* The same call with PACKED output planes (HOST pointers, tile-compact like the float planes; any may be NULL = not wanted).
 * Five of the seven float planes of a frame are 0 / 1 shadow flags (LightAmbient.frag:167-169, LightPoint.frag:125) and spec_t
 * (LightReflection.frag:113) takes 180 values, so a frame that crosses PCIe as float32 is mostly air:
 *   shadow_mask  uint8[px][mask_bytes], mask_bytes = (1 + n_point + n_spot + 7) / 8; bit p of the little-endian mask is plane p:
 *                0 = sun shadow, 1 .. n_point = point lights, then the spot lights; set = the float plane holds 1.0f, clear = 0.0f
 *   spec_code    uint8[px]: 0..30 -> 0.5 * (code + 1); 31..178 -> 16 + (code - 31); 255 -> 256.0f (miss / unlit).  Exactly invertible.
 *   ao           float32[px], unchanged
 * The out_* members of `args` are ignored.  Which passes run: ambient if shadow_mask or ao, local lights if shadow_mask and
 * n_point / n_spot > 0, reflection if spec_code.  At 3840x2160 with 4 point lights the read-back is 50 MB instead of 232 MB. */
typedef struct vxl_packed_planes {
    uint8_t* shadow_mask;
    uint8_t* spec_code;
    float* ao;
} vxl_packed_planes;
int vxl_lighting_host_packed(vxl_ctx* ctx, vxl_volume* vol, const vxl_lighting_host_args* args, const vxl_packed_planes* out);
/* The same frame with DEVICE pointers throughout (frame planes, outputs; view and lights stay HOST): the ambient pass on the context's
 * stream, the local-light passes and the reflection pass on two side streams forked from it and joined back before the call returns
 * (stream-ordered, asynchronous).  The passes are independent, so the tail of one kernel overlaps the head of the next.  Same
 * results as the single vxl_pass_* calls; output mirrors and the light plane stride apply. */
int vxl_lighting(vxl_ctx* ctx, vxl_volume* vol, const vxl_lighting_host_args* args);

/* ---- synthetic inputs (not reference passes) ------------------------------------- */
/* FastNoise-Perlin terrain: voxel solid iff GetTerrainNoise(x,y,z) > (y/NY - 0.5)*2
 * (Sources/Util/Noise.cpp:131-135, FastNoise seed 1337 freq 0.01). Overwrites the volume. */
int vxl_volume_gen_terrain(vxl_volume* vol);
/* primary-visibility G-buffer through the packed volume (A4 DDA arithmetic, reference encodings).
 * Writes the frame's depth24/normal/material planes (cast away const: DEVICE, writable). */
int vxl_gbuffer_primary(vxl_ctx* ctx, vxl_volume* vol, const vxl_view* view, const vxl_frame* frame);

#ifdef __cplusplus
}
#endif
#endif /* VXL_H */
