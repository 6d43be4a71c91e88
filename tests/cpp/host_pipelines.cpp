// tests/cpp/host_pipelines.cpp -- a C++ renderer-side caller of include/vxl_pipelines.hpp (the reference's pass objects over the C ABI).
// Reads a scene file written by tests/test_cpp_host.py, runs the four light passes the way WorldRenderer::DrawWorld does
// (Sources/Graphics/Renderer/WorldRenderer.cpp:239-274), writes the output planes; the test compares them bit for bit with the
// oracle.  usage: host_pipelines <scene.bin> <out.bin> [device]      exit 3 = vxl::Error (message on stderr)
#include "vxl_pipelines.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Header { int32_t magic, sx, sy, sz, W, H, n_ao, n_point, n_spot; };

template <typename T> static bool rd(FILE* f, T* p, size_t n) { return fread(p, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: host_pipelines <scene.bin> <out.bin> [device]\n"); return 2; }
    try {
        vxl::Context ctx(argc > 3 ? atoi(argv[3]) : 0);
        FILE* f = fopen(argv[1], "rb");
        Header h;
        if (!f || !rd(f, &h, 1) || h.magic != 0x4C5856) { fprintf(stderr, "bad scene file\n"); return 2; }
        vxl_view view;
        std::vector<uint8_t> volume((size_t)h.sx * h.sy * h.sz);
        std::vector<uint32_t> noise(512 * 512);
        std::vector<vxl_point_light> pl((size_t)h.n_point);
        std::vector<vxl_spot_light> sl((size_t)h.n_spot);
        if (!rd(f, &view, 1) || !rd(f, volume.data(), volume.size()) || !rd(f, noise.data(), noise.size()) || !rd(f, pl.data(), pl.size()) ||
            !rd(f, sl.data(), sl.size())) { fprintf(stderr, "short scene file\n"); return 2; }
        fclose(f);

        vxl::ShadowVoxSystem shadowVox(ctx, h.sx, h.sy, h.sz);
        shadowVox.Upload(volume.data());
        const size_t px = (size_t)h.W * h.H;
        uint32_t* depth = ctx.Alloc<uint32_t>(px); uint32_t* normal = ctx.Alloc<uint32_t>(px); uint32_t* material = ctx.Alloc<uint32_t>(px);
        uint32_t* d_noise = ctx.Alloc<uint32_t>(noise.size());
        ctx.Upload(d_noise, noise.data(), noise.size() * 4);
        vxl::GeometryFramebuffer fb(h.W, h.H, depth, normal, material, d_noise);
        vxl::Check(vxl_gbuffer_primary(ctx, shadowVox.GetVolumeImage(), &view, &fb.Frame), "vxl_gbuffer_primary");   // synthetic geometry pass

        const size_t planes = 3 + (size_t)h.n_point + (size_t)h.n_spot;
        float* out = ctx.Alloc<float>(planes * px);
        float *shadow = out, *ao = out + px, *spec = out + 2 * px, *point = out + 3 * px, *spot = point + (size_t)h.n_point * px;
        vxl::Check(vxl_stats_reset(ctx), "vxl_stats_reset");
        vxl::LightAmbientPipeline::Get().Use(ctx, view, fb, shadowVox, h.n_ao, shadow, ao);
        int warned = 0;
        vxl::LightPointPipeline::Get().Warn = [&](const char*) { ++warned; };
        vxl::LightPointPipeline::Get().Use(ctx, view, fb, shadowVox, [&](vxl::LightPointPipeline& P) {
            for (auto& l : pl) P.DrawLight(l.Position, l.Range, l.Color, l.Attenuation);
            for (int i = 0; i < 70 && h.n_point == VXL_MAX_LIGHTS; ++i) P.DrawLight(pl[0].Position, 1.0f, pl[0].Color, 1.0f);   // beyond the limit: dropped with a warning
        }, point);
        vxl::LightSpotPipeline::Get().Use(ctx, view, fb, shadowVox, [&](vxl::LightSpotPipeline& P) {
            for (auto& l : sl) P.DrawLight(l.Position, l.Range, l.Color, l.Attenuation, l.Direction, l.Angle, l.AngleAttenuation);
        }, spot);
        vxl::LightReflectionPipeline::Get().Use(ctx, view, fb, shadowVox, spec);
        vxl_stats st;
        vxl::Check(vxl_stats_read(ctx, &st), "vxl_stats_read");
        // the same frame through the one-call form (vxl_lighting: concurrent pass kernels) must give the same planes
        float* out2 = ctx.Alloc<float>(planes * px);
        vxl::DrawLights(ctx, view, fb, shadowVox, h.n_ao,
                        [&](vxl::LightPointPipeline& P) { for (auto& l : pl) P.DrawLight(l.Position, l.Range, l.Color, l.Attenuation); },
                        [&](vxl::LightSpotPipeline& P) { for (auto& l : sl) P.DrawLight(l.Position, l.Range, l.Color, l.Attenuation, l.Direction, l.Angle, l.AngleAttenuation); },
                        out2, out2 + px, h.n_point ? out2 + 3 * px : nullptr, h.n_spot ? out2 + (3 + (size_t)h.n_point) * px : nullptr, out2 + 2 * px);
        {
            std::vector<float> a(planes * px), b(planes * px);
            ctx.Download(a.data(), out, a.size() * 4); ctx.Download(b.data(), out2, b.size() * 4);
            if (memcmp(a.data(), b.data(), a.size() * 4) != 0) { fprintf(stderr, "DrawLights planes differ from the single passes\n"); return 6; }
        }
        ctx.Free(out2);

        std::vector<float> host(planes * px);
        std::vector<uint32_t> gb(3 * px);
        ctx.Download(host.data(), out, host.size() * 4);
        ctx.Download(gb.data(), depth, px * 4); ctx.Download(gb.data() + px, normal, px * 4); ctx.Download(gb.data() + 2 * px, material, px * 4);
        FILE* o = fopen(argv[2], "wb");
        if (!o) { fprintf(stderr, "cannot write %s\n", argv[2]); return 2; }
        const uint64_t tail[3] = {st.rays, st.steps, (uint64_t)warned};
        fwrite(gb.data(), 4, gb.size(), o); fwrite(host.data(), 4, host.size(), o); fwrite(tail, 8, 3, o);
        fclose(o);
        // a failing call throws like the reference's CHECK: more lights than a frame can hold through the C entry point
        try { vxl::Check(vxl_pass_point(ctx, shadowVox.GetVolumeImage(), &view, &fb.Frame, pl.data(), VXL_MAX_LIGHTS + 1, point), "vxl_pass_point"); return 4; }
        catch (const vxl::Error& e) { if (e.status != VXL_ERR_LIMIT && e.status != VXL_ERR_INVALID) return 5; }
        ctx.Free(out); ctx.Free(depth); ctx.Free(normal); ctx.Free(material); ctx.Free(d_noise);
        printf("ok rays=%llu probes=%llu\n", (unsigned long long)st.rays, (unsigned long long)st.steps);
        return 0;
    } catch (const vxl::Error& e) {
        fprintf(stderr, "vxl::Error(%d): %s\n", e.status, e.what());
        return 3;
    }
}
