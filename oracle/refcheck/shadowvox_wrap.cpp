// oracle/refcheck/shadowvox_wrap.cpp -- TEST INFRASTRUCTURE: drives the REFERENCE's own
// Sources/World/Systems/ShadowVoxSystem.cpp (compiled from where it lies, with the reference's vendored
// entt 3.6 and glm; engine/Vulkan headers replaced by the interface stand-ins under stubs/) so that the
// oracle's vxo_voxelize can be pinned to it: staging bytes, dirty regions, and entt's visiting order.
// Built by oracle/Makefile into oracle/_ref/libvxshadowvox.so.
#include <cstring>
#include REF_SHADOWVOX_CPP

namespace {
std::vector<ImageRegion> g_copied;      // regions of every cmd.copy since the last reset
Buffer* g_staging = nullptr;
std::vector<std::shared_ptr<std::vector<uint8_t>>> g_buffers;
}
Buffer Buffer::Create(uint64_t size) { Buffer b; b._d = std::make_shared<std::vector<uint8_t>>((size_t)size, (uint8_t)0xCD); g_buffers.push_back(b._d); return b; }
void CmdBuffer::copy(Buffer&, Image&, const std::vector<ImageRegion>& regions) { g_copied.insert(g_copied.end(), regions.begin(), regions.end()); }

// vxo_entity layout (oracle/vxo.h)
struct EntPOD { int32_t model, flags; float prev[16], cur[16], pivot[3]; int32_t pad; };
struct RegPOD { int32_t x, y, z; uint32_t w, h, d; int32_t mip; };

extern "C" {

// Runs: construct the system (524 x 188 x 524 texels, ShadowVoxSystem.cpp:55-79), OnCreate, create the entities
// in index order with (Transform, VoxRenderer, Changed), OnUpdate(0); then, for entities flagged destroy = 1,
// registry.remove<VoxRenderer> (-> OnVoxDestroyed) in index order.
// Outputs: staging bytes (sx*sy*sz), the regions passed to cmd.copy after construction (OnUpdate's batch, then one per
// destroyed entity is pushed but only copied by a further OnUpdate, which is called at the end), the order in which
// entt's view visited the entities.  Returns the number of regions.
int vxref_shadowvox_run(const uint8_t* const* model_data, const int32_t* model_dims /* [n_models][3] */, int n_models,
                        const EntPOD* ents, int n, const int32_t* destroy /* [n] or null */,
                        uint8_t* out_bytes, int32_t* out_dims, RegPOD* out_regions, int max_regions, int32_t* out_order) {
    g_copied.clear(); g_buffers.clear();
    std::vector<VoxAsset> assets((size_t)n_models);
    for (int i = 0; i < n_models; ++i) {
        VoxAsset& a = assets[i];
        a.SizeX = model_dims[i * 3]; a.SizeY = model_dims[i * 3 + 1]; a.SizeZ = model_dims[i * 3 + 2];
        a.Data.assign(model_data[i], model_data[i] + (size_t)a.SizeX * a.SizeY * a.SizeZ);
        a._Image = Image::Create(Image::Info(Format::R8Uint, {a.SizeX, a.SizeY, a.SizeZ}));
    }
    entt::registry reg;
    ShadowVoxSystem sys;
    sys.R = &reg; sys.W = nullptr;
    sys.OnCreate();
    g_copied.clear();                                   // drop the constructor's (empty) copy
    std::vector<entt::entity> es((size_t)n);
    for (int i = 0; i < n; ++i) {
        es[i] = reg.create();
        Transform t; memcpy(&t.PreviousWorldMatrix, ents[i].prev, 64); memcpy(&t.WorldMatrix, ents[i].cur, 64);
        reg.emplace<Transform>(es[i], t);
        VoxRenderer vr; vr.Vox = AssetSlot<VoxAsset>(&assets[ents[i].model]); vr.Pivot = glm::vec3(ents[i].pivot[0], ents[i].pivot[1], ents[i].pivot[2]);
        reg.emplace<VoxRenderer>(es[i], vr);
        reg.emplace<Changed>(es[i]);
    }
    int k = 0;
    reg.view<Transform, VoxRenderer, Changed>().each([&](const entt::entity e, Transform&, VoxRenderer&) {
        for (int i = 0; i < n; ++i) if (es[i] == e) out_order[k++] = i;
    });
    sys.OnUpdate(0.0f);
    reg.clear<Changed>();
    if (destroy) {
        for (int i = 0; i < n; ++i) if (destroy[i]) reg.remove<VoxRenderer>(es[i]);
        sys.OnUpdate(0.0f);                             // flushes the regions pushed by OnVoxDestroyed
    }
    Extent e = sys.GetVolumeImage().getExtent();
    out_dims[0] = (int)e.width; out_dims[1] = (int)e.height; out_dims[2] = (int)e.depth;
    // the staging buffer is the first (and only) Buffer the system created
    if (out_bytes && !g_buffers.empty()) memcpy(out_bytes, g_buffers[0]->data(), g_buffers[0]->size());
    int nr = 0;
    for (const ImageRegion& r : g_copied) { if (nr < max_regions) out_regions[nr] = RegPOD{r.x, r.y, r.z, r.width, r.height, r.depth, (int32_t)r.mip}; ++nr; }
    return nr;
}

}  // extern "C"
