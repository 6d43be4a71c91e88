// include/vxl_pipelines.hpp -- the reference's pass objects, in C++, over the C ABI of include/vxl.h.
//
// The reference is compiled C++; this header is the host side a renderer written against it keeps calling: the same class names,
// the same Get() / Use() / DrawLight() shape, the same argument meaning and the same error behaviour (the reference's CHECK logs
// and throws, Sources/Core/Core.h:83-86; here every non-zero vxl_status throws vxl::Error carrying vxl_last_error_string()).
// Header-only, C++17, no CUDA headers: link with -lvxl.  What changes against the reference is what a Vulkan -> CUDA move has to
// change: evk images / buffers (bindless RIDs) become device pointers grouped in GeometryFramebuffer, and the command buffer becomes
// the context's stream.
//
//   reference                                                              here
//   ---------------------------------------------------------------------  --------------------------------------
//   ShadowVoxSystem (World/Systems/ShadowVoxSystem.h:9-32)                  vxl::ShadowVoxSystem
//   LightAmbientPipeline::Use    (Pipelines/LightAmbientPipeline.h:35-52)   vxl::LightAmbientPipeline::Get().Use
//   LightPointPipeline::Use + DrawLight (LightPointPipeline.h:58-100)       vxl::LightPointPipeline::Get().Use(..., cb)
//   LightSpotPipeline::Use + DrawLight  (LightSpotPipeline.h:60-104)        vxl::LightSpotPipeline::Get().Use(..., cb)
//   LightReflectionPipeline::Use (LightReflectionPipeline.h:34-51)          vxl::LightReflectionPipeline::Get().Use
//   LightTAAPipeline::Use        (LightTAAPipeline.h:34-53)                 vxl::LightTAAPipeline::Get().Use
//   GeometryVoxelPipeline::Use + Draw (GeometryVoxelPipeline.h:30-71)       vxl::GeometryVoxelPipeline::Get().Use(..., cb)
//   VoxImporter::Import          (Editor/Importer/VoxImporter.cpp:478-520)  vxl::VoxImporter::Import
#pragma once
#include "vxl.h"

#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace vxl {

struct Error : std::runtime_error {
    int status;
    Error(int s, const char* what_) : std::runtime_error(std::string(what_) + ": " + vxl_last_error_string()), status(s) {}
};
inline void Check(int status, const char* what) {                 // CHECK (Core.h:83-86): log + throw
    if (status != VXL_OK) throw Error(status, what);
}

// One GPU + one stream: stands where the reference has Graphics::Frame's command buffer (Graphics.h:23-32).
class Context {
    vxl_ctx* _Ctx = nullptr;
public:
    explicit Context(int device = 0) { Check(vxl_ctx_create(device, &_Ctx), "vxl_ctx_create"); }
    ~Context() { if (_Ctx) vxl_ctx_destroy(_Ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    operator vxl_ctx*() const { return _Ctx; }
    void Wait() { Check(vxl_sync(_Ctx), "vxl_sync"); }             // the fence wait of Graphics::Frame
    template <typename T> T* Alloc(size_t count) { void* p = nullptr; Check(vxl_malloc(_Ctx, count * sizeof(T), &p), "vxl_malloc"); return (T*)p; }
    void Free(void* p) { vxl_free(_Ctx, p); }
    void Upload(void* dev, const void* host, size_t bytes) { Check(vxl_memcpy_h2d(_Ctx, dev, host, bytes), "vxl_memcpy_h2d"); }
    void Download(void* host, const void* dev, size_t bytes) { Check(vxl_memcpy_d2h(_Ctx, host, dev, bytes), "vxl_memcpy_d2h"); Wait(); }
};

// The attachments of Passes::Geometry the light passes read (Graphics.h:51-60) + the blue-noise image, as device planes.
// A whole frame on one GPU is one tile; SetShard selects the screen tiles of one rank (tile-compact planes).
struct GeometryFramebuffer {
    vxl_frame Frame{};
    const uint32_t* Color = nullptr;                                // COLOR_TEXTURE (albedo), only the colour resolve reads it
    GeometryFramebuffer(int width, int height, const uint32_t* depth24, const uint32_t* normal, const uint32_t* material, const uint32_t* blueNoise) {
        Frame.width = width; Frame.height = height; Frame.tile_w = width; Frame.tile_h = height;
        Frame.tile_first = 0; Frame.tile_stride = 1; Frame.n_tiles = 1;
        Frame.depth24 = depth24; Frame.normal = normal; Frame.material = material; Frame.noise = blueNoise;
    }
    void SetShard(int tileW, int tileH, int rank, int world) {
        const int tx = (Frame.width + tileW - 1) / tileW, ty = (Frame.height + tileH - 1) / tileH, total = tx * ty;
        Frame.tile_w = tileW; Frame.tile_h = tileH; Frame.tile_first = rank; Frame.tile_stride = world;
        Frame.n_tiles = rank < total ? (total - rank + world - 1) / world : 0;
    }
    size_t Pixels() const { return (size_t)Frame.n_tiles * Frame.tile_w * Frame.tile_h; }
};

// ShadowVoxSystem (ShadowVoxSystem.h:9-32): owns the packed world volume; OnUpdate voxelises the changed entities.
class ShadowVoxSystem {
    vxl_ctx* _Ctx;
    vxl_volume* _Volume = nullptr;
public:
    ShadowVoxSystem(Context& ctx, int texelsX = 524, int texelsY = 188, int texelsZ = 524) : _Ctx(ctx) {   // ShadowVoxSystem.cpp:55-79 sizes
        Check(vxl_volume_create(_Ctx, texelsX, texelsY, texelsZ, &_Volume), "vxl_volume_create");
    }
    ~ShadowVoxSystem() { if (_Volume) vxl_volume_destroy(_Volume); }
    ShadowVoxSystem(const ShadowVoxSystem&) = delete;
    ShadowVoxSystem& operator=(const ShadowVoxSystem&) = delete;
    vxl_volume* GetVolumeImage() const { return _Volume; }
    // VoxAsset::Upload: palette indices, x fastest -> model id for vxl_entity::model
    int AddModel(const uint8_t* voxels, int sx, int sy, int sz) { int id = -1; Check(vxl_model_create(_Ctx, voxels, sx, sy, sz, &id), "vxl_model_create"); return id; }
    // OnUpdate (:116-201) over the entities that carry Changed; OnVoxDestroyed entries have flags = VXL_ENT_DESTROY.
    // regions (optional) receives the dirty ImageRegions the reference would copy.
    void OnUpdate(const std::vector<vxl_entity>& changed, std::vector<vxl_region>* regions = nullptr) {
        if (changed.empty()) return;
        std::vector<vxl_region> r(regions ? changed.size() : 0);
        std::vector<int32_t> valid(regions ? changed.size() : 0);
        Check(vxl_volume_voxelize(_Volume, changed.data(), (int)changed.size(), regions ? r.data() : nullptr, regions ? valid.data() : nullptr), "vxl_volume_voxelize");
        if (regions) { regions->clear(); for (size_t i = 0; i < r.size(); ++i) if (valid[i]) regions->push_back(r[i]); }
    }
    void Upload(const uint8_t* hostBytes) { Check(vxl_volume_upload(_Volume, hostBytes), "vxl_volume_upload"); }
    void UploadRegions(const uint8_t* staging, const std::vector<vxl_region>& regions) {                   // CmdBuffer::copy, evk.cpp:759-780
        Check(vxl_volume_upload_regions(_Volume, staging, regions.data(), (int)regions.size()), "vxl_volume_upload_regions");
    }
    void Download(uint8_t* hostBytes) { Check(vxl_volume_download(_Volume, hostBytes), "vxl_volume_download"); }
};

// The frame sharded over the GPUs of one box (no reference counterpart: one GPU there): one ShardGroup per process / GPU.  The
// passes of a member store their output tiles into EVERY member's copy of the gathered stack (peer stores over NVLink), Fence()
// closes the frame with peer-written arrival flags -- no collective.  Handles travel between the processes by whatever the host
// has (a pipe, a file, MPI): Handle() out, Connect(all handles) in.
class ShardGroup {
    vxl_group* _Group = nullptr;
public:
    ShardGroup(Context& ctx, int rank, int ranks, size_t stackBytes, int stacks = 2) {
        Check(vxl_group_create(ctx, rank, ranks, stackBytes, stacks, &_Group), "vxl_group_create");
    }
    ~ShardGroup() { if (_Group) vxl_group_destroy(_Group); }
    ShardGroup(const ShardGroup&) = delete;
    ShardGroup& operator=(const ShardGroup&) = delete;
    vxl_ipc_handle Handle() { vxl_ipc_handle h; Check(vxl_group_handle(_Group, &h), "vxl_group_handle"); return h; }
    void Connect(const std::vector<vxl_ipc_handle>& handles) { Check(vxl_group_connect(_Group, handles.data()), "vxl_group_connect"); }
    void* Base() { void* p = nullptr; Check(vxl_group_base(_Group, &p), "vxl_group_base"); return p; }
    void ConnectPointers(const std::vector<void*>& bases) { Check(vxl_group_connect_pointers(_Group, bases.data()), "vxl_group_connect_pointers"); }
    // mirrors on; returns this member's copy of the frame's stack (its pass outputs go to its own slot in there)
    float* BeginFrame(uint64_t frame) { void* p = nullptr; Check(vxl_group_begin_frame(_Group, frame, &p), "vxl_group_begin_frame"); return (float*)p; }
    void EndFrame() { Check(vxl_group_end_frame(_Group), "vxl_group_end_frame"); }
    void Fence() { Check(vxl_group_fence(_Group), "vxl_group_fence"); }
    void CheckArrived() { int r = 0; Check(vxl_group_status(_Group, &r), "vxl_group_status"); }
    void Release() { if (_Group) { vxl_group_destroy(_Group); _Group = nullptr; } }
};

class LightAmbientPipeline {
public:
    // Use (LightAmbientPipeline.h:35-52): sun shadow + AO planes.  aoRays = 1 is the reference's one-sample estimate.
    void Use(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, ShadowVoxSystem& shadowVox, int aoRays,
             float* outShadow, float* outAO) {
        Check(vxl_pass_ambient(cmd, shadowVox.GetVolumeImage(), &view, &geometryFB.Frame, aoRays, outShadow, outAO), "vxl_pass_ambient");
    }
    // the colour the pass adds to the light buffer (LightAmbient.frag:178-214)
    void Colour(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, const float* shadow, const float* ao, float* outRGBA,
                const uint32_t* depthFull = nullptr) {
        vxl_resolve r{geometryFB.Color, depthFull};
        Check(vxl_resolve_ambient(cmd, &view, &geometryFB.Frame, &r, shadow, ao, outRGBA), "vxl_resolve_ambient");
    }
    static LightAmbientPipeline& Get() { static LightAmbientPipeline Instance; return Instance; }
};

class LightPointPipeline {
    static constexpr int MAX_POINT_LIGHTS = VXL_MAX_LIGHTS;          // LightPointPipeline.h:15
    vxl_point_light _Lights[MAX_POINT_LIGHTS];
    int _CurrentLightIndex = 0;
public:
    std::function<void(const char*)> Warn;                           // Log::warn stand-in; default: silent
    void DrawLight(const float position[3], float range, const float color[3], float attenuation) {         // :61-75
        if (_CurrentLightIndex >= MAX_POINT_LIGHTS) { if (Warn) Warn("Max number of Point Lights reached!"); return; }
        vxl_point_light& l = _Lights[_CurrentLightIndex++];
        std::memcpy(l.Position, position, 12); l.Range = range; std::memcpy(l.Color, color, 12); l.Attenuation = attenuation;
    }
    int LightCount() const { return _CurrentLightIndex; }
    const vxl_point_light* Lights() const { return _Lights; }
    // Use (:77-100): cb draws the lights; outShadow receives one plane per drawn light, in draw order.
    void Use(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, ShadowVoxSystem& shadowVox,
             const std::function<void(LightPointPipeline& P)>& cb, float* outShadow) {
        _CurrentLightIndex = 0;
        cb(*this);
        Check(vxl_pass_point(cmd, shadowVox.GetVolumeImage(), &view, &geometryFB.Frame, _Lights, _CurrentLightIndex, outShadow), "vxl_pass_point");
    }
    // the additive colour of the lights drawn by the last Use (LightPoint.frag:131-152)
    void Colour(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, const float* shadow, float* inoutRGBA) {
        vxl_resolve r{geometryFB.Color, nullptr};
        Check(vxl_resolve_point(cmd, &view, &geometryFB.Frame, &r, _Lights, _CurrentLightIndex, shadow, inoutRGBA), "vxl_resolve_point");
    }
    static LightPointPipeline& Get() { static LightPointPipeline Instance; return Instance; }
};

class LightSpotPipeline {
    static constexpr int MAX_SPOT_LIGHTS = VXL_MAX_LIGHTS;           // LightSpotPipeline.h:14
    vxl_spot_light _Lights[MAX_SPOT_LIGHTS];
    int _CurrentLightIndex = 0;
public:
    std::function<void(const char*)> Warn;
    void DrawLight(const float position[3], float range, const float color[3], float attenuation, const float direction[3], float angle,
                   float angleAttenuation) {                          // LightSpotPipeline.h:63-79
        if (_CurrentLightIndex >= MAX_SPOT_LIGHTS) { if (Warn) Warn("Max number of Spot Lights reached!"); return; }
        vxl_spot_light& l = _Lights[_CurrentLightIndex++];
        std::memset(&l, 0, sizeof l);
        std::memcpy(l.Position, position, 12); l.Range = range; std::memcpy(l.Color, color, 12); l.Attenuation = attenuation;
        std::memcpy(l.Direction, direction, 12); l.Angle = angle; l.AngleAttenuation = angleAttenuation;
    }
    int LightCount() const { return _CurrentLightIndex; }
    const vxl_spot_light* Lights() const { return _Lights; }
    void Use(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, ShadowVoxSystem& shadowVox,
             const std::function<void(LightSpotPipeline& P)>& cb, float* outShadow) {
        _CurrentLightIndex = 0;
        cb(*this);
        Check(vxl_pass_spot(cmd, shadowVox.GetVolumeImage(), &view, &geometryFB.Frame, _Lights, _CurrentLightIndex, outShadow), "vxl_pass_spot");
    }
    void Colour(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, const float* shadow, float* inoutRGBA) {
        vxl_resolve r{geometryFB.Color, nullptr};
        Check(vxl_resolve_spot(cmd, &view, &geometryFB.Frame, &r, _Lights, _CurrentLightIndex, shadow, inoutRGBA), "vxl_resolve_spot");
    }
    static LightSpotPipeline& Get() { static LightSpotPipeline Instance; return Instance; }
};

class LightReflectionPipeline {
public:
    // Use (LightReflectionPipeline.h:34-51): the march; outSpecT = reflection-ray distance, 256 = miss
    void Use(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, ShadowVoxSystem& shadowVox, float* outSpecT) {
        Check(vxl_pass_reflection(cmd, shadowVox.GetVolumeImage(), &view, &geometryFB.Frame, outSpecT), "vxl_pass_reflection");
    }
    // the colour around the march (LightReflection.frag:115-139): lightFull = the TAA light buffer of the whole frame
    void Colour(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, const float* specT, const float* lightFull,
                const float skyRGB[3], float* outRGBA, const uint32_t* depthFull = nullptr) {
        Check(vxl_resolve_reflection(cmd, &view, &geometryFB.Frame, specT, depthFull, lightFull, skyRGB, outRGBA), "vxl_resolve_reflection");
    }
    static LightReflectionPipeline& Get() { static LightReflectionPipeline Instance; return Instance; }
};

class LightTAAPipeline {
public:
    // Use (LightTAAPipeline.h:34-53): whole-frame input planes, output for the framebuffer's pixels
    void Use(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, const vxl_full_planes& full, float* outRGBA) {
        Check(vxl_light_taa(cmd, &view, &geometryFB.Frame, &full, outRGBA), "vxl_light_taa");
    }
    static LightTAAPipeline& Get() { static LightTAAPipeline Instance; return Instance; }
};

class GeometryVoxelPipeline {
    static constexpr int MAX_INSTANCES = 4096;                        // GeometryVoxelPipeline.h
    std::vector<vxl_vox_cmd> _Cmds;
public:
    // Draw (:38-47): one instance; model = the vxl model id of the VoxAsset (ShadowVoxSystem::AddModel / vxl_model_load_v)
    void Draw(int model, int volumeRID, int palleteIndex, const float worldMatrix[16], const float lastWorldMatrix[16]) {
        if ((int)_Cmds.size() >= MAX_INSTANCES) return;
        vxl_vox_cmd c{};
        std::memcpy(c.WorldMatrix, worldMatrix, 64); std::memcpy(c.LastWorldMatrix, lastWorldMatrix, 64);
        c.VolumeRID = volumeRID; c.PalleteIndex = palleteIndex; c.model = model;
        _Cmds.push_back(c);
    }
    // Use (:49-71): cb records the draws; the G-buffer planes are written for the framebuffer's pixels
    void Use(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, const uint32_t* palleteColor, const uint32_t* palleteMaterial,
             const std::function<void(GeometryVoxelPipeline& P)>& cb, const vxl_gbuffer_out& out) {
        _Cmds.clear();
        cb(*this);
        Check(vxl_gbuffer_models(cmd, &view, &geometryFB.Frame, _Cmds.data(), (int)_Cmds.size(), palleteColor, palleteMaterial, &out), "vxl_gbuffer_models");
    }
    static GeometryVoxelPipeline& Get() { static GeometryVoxelPipeline Instance; return Instance; }
};

// The "Lights" + "Reflection" blocks of WorldRenderer::DrawWorld (WorldRenderer.cpp:239-274) as ONE call on device planes: the lights
// recorded by the two callbacks, then vxl_lighting -- the three pass kernels on concurrent streams, joined on the context's stream.
// Any output may be null (that pass is skipped).  outPointShadow / outSpotShadow: one plane per drawn light.
inline void DrawLights(Context& cmd, const vxl_view& view, const GeometryFramebuffer& geometryFB, ShadowVoxSystem& shadowVox, int aoRays,
                       const std::function<void(LightPointPipeline& P)>& pointCb, const std::function<void(LightSpotPipeline& P)>& spotCb,
                       float* outShadow, float* outAO, float* outPointShadow, float* outSpotShadow, float* outSpecT) {
    struct Collect : LightPointPipeline { using LightPointPipeline::LightPointPipeline; } point;
    struct CollectSpot : LightSpotPipeline { using LightSpotPipeline::LightSpotPipeline; } spot;
    if (pointCb) pointCb(point);
    if (spotCb) spotCb(spot);
    vxl_lighting_host_args a{};
    a.frame = geometryFB.Frame; a.view = &view; a.n_ao = aoRays;
    a.n_point = point.LightCount(); a.point = point.Lights();
    a.n_spot = spot.LightCount(); a.spot = spot.Lights();
    a.out_shadow = outShadow; a.out_ao = outAO; a.out_point_shadow = outPointShadow; a.out_spot_shadow = outSpotShadow; a.out_spec_t = outSpecT;
    Check(vxl_lighting(cmd, shadowVox.GetVolumeImage(), &a), "vxl_lighting");
}

// VoxImporter (Editor/Importer/VoxImporter.cpp): a dropped .vox file -> <mods>/<path>/<file>/<shape>.v, <file>.p, <path>/<file>.pf
struct VoxImporter {
    static void Import(const std::string& voxFile, const std::string& modsDir, const std::string& path, const std::string& fileName) {
        vxl_vox_scene* sc = nullptr;
        Check(vxl_vox_import(voxFile.c_str(), &sc), "vxl_vox_import");
        const int rc = vxl_vox_scene_write(sc, modsDir.c_str(), path.c_str(), fileName.c_str());
        vxl_vox_scene_free(sc);
        Check(rc, "vxl_vox_scene_write");
    }
};

// PrefabAsset::Spawn + TransformSystem over a Mods directory: the scene as a flat entity list with world matrices
inline std::vector<vxl_prefab_entity> LoadScene(const std::string& modsDir, const std::string& prefabPath) {
    int n = 0;
    Check(vxl_scene_load(modsDir.c_str(), prefabPath.c_str(), nullptr, 0, &n), "vxl_scene_load");
    std::vector<vxl_prefab_entity> ents((size_t)n);
    if (n) Check(vxl_scene_load(modsDir.c_str(), prefabPath.c_str(), ents.data(), n, &n), "vxl_scene_load");
    return ents;
}

}  // namespace vxl
