// tests/cpp/host_group.cpp -- TEST PROGRAM: the tile-sharded frame driven from C++ through include/vxl_pipelines.hpp, two PROCESSES
// (forked before any CUDA call), one vxl::ShardGroup member each, on the GPUs the box has (both on device 0 if there is one).
// Each process generates the same synthetic scene (terrain volume + primary-visibility G-buffer, vxl_volume_gen_terrain /
// vxl_gbuffer_primary), runs the ambient and reflection passes over ITS round-robin tile shard into its slot of the frame's stack
// -- the mirrors store the same values into the peer's copy -- and closes the frame with the flag fence.  The handles travel
// through pipes.  After two frames (the two stacks) each process reports a checksum of its whole copy and of each rank's slot: the
// copies must be identical and both slots non-empty.  Prints "OK" and exits 0, or says what differed.
//   usage: host_group <texels> <width> <height> <n_devices>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vxl_pipelines.hpp"

namespace {

struct Report { unsigned long long whole[2], slot[2][2]; int ok; };

unsigned long long fnv(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

bool write_all(int fd, const void* p, size_t n) { return write(fd, p, n) == (ssize_t)n; }
bool read_all(int fd, void* p, size_t n) {
    char* c = (char*)p;
    while (n) { ssize_t k = read(fd, c, n); if (k <= 0) return false; c += k; n -= (size_t)k; }
    return true;
}

int member(int rank, int ranks, int device, int texels, int W, int H, int to_peer, int from_peer, int to_parent) {
    Report rep{};
    try {
        vxl::Context ctx(device);
        vxl::ShadowVoxSystem vox(ctx, texels, texels, texels);
        vxl::Check(vxl_volume_gen_terrain(vox.GetVolumeImage()), "vxl_volume_gen_terrain");
        const int TW = 64, TH = 32;
        vxl::GeometryFramebuffer fb(W, H, nullptr, nullptr, nullptr, nullptr);
        fb.SetShard(TW, TH, rank, ranks);
        const size_t px = fb.Pixels();
        uint32_t* planes = ctx.Alloc<uint32_t>(px * 3 + 512 * 512);
        fb.Frame.depth24 = planes; fb.Frame.normal = planes + px; fb.Frame.material = planes + 2 * px; fb.Frame.noise = planes + 3 * px;
        std::vector<uint32_t> noise(512 * 512);
        unsigned s = 12345u;
        for (auto& v : noise) { s = s * 1664525u + 1013904223u; v = s; }
        ctx.Upload(planes + 3 * px, noise.data(), noise.size() * 4);
        vxl_view view;
        std::memset(&view, 0, sizeof view);
        // camera at the volume centre, 0.75 of its height, looking down a little: built by the library's own helper is not
        // exported, so an axis-aligned one is spelled out: View = T(-eye), Proj = perspective(0.8, W/H, 0.1, 4096) column-major
        const float ext = 2.0f * (float)texels * 0.1f, ex = ext * 0.5f, ey = ext * 0.75f, ez = ext * 0.5f;
        const float f = 1.0f / 0.4227932f /* tan(0.4) */, asp = (float)W / (float)H, n = 0.1f, fr = 4096.0f;
        float V[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -ex, -ey, -ez, 1}, IV[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, ex, ey, ez, 1};
        float P[16] = {f / asp, 0, 0, 0, 0, f, 0, 0, 0, 0, (fr + n) / (n - fr), -1, 0, 0, 2 * fr * n / (n - fr), 0};
        float IP[16] = {asp / f, 0, 0, 0, 0, 1 / f, 0, 0, 0, 0, 0, (n - fr) / (2 * fr * n), 0, 0, -1, (fr + n) / (2 * fr * n)};
        std::memcpy(view.LastViewMatrix, V, 64); std::memcpy(view.ViewMatrix, V, 64); std::memcpy(view.InverseViewMatrix, IV, 64);
        std::memcpy(view.ProjectionMatrix, P, 64); std::memcpy(view.InverseProjectionMatrix, IP, 64);
        view.Res[0] = (float)W; view.Res[1] = (float)H; view.iRes[0] = 1.0f / W; view.iRes[1] = 1.0f / H;
        view.CameraPosition[0] = ex; view.CameraPosition[1] = ey; view.CameraPosition[2] = ez;
        vxl::Check(vxl_gbuffer_primary(ctx, vox.GetVolumeImage(), &view, &fb.Frame), "vxl_gbuffer_primary");

        const int tx = (W + TW - 1) / TW, ty = (H + TH - 1) / TH, padded = (tx * ty + ranks - 1) / ranks;
        const size_t plane = (size_t)padded * TW * TH, slot = 3 * plane, stack = slot * (size_t)ranks;
        vxl::ShardGroup group(ctx, rank, ranks, stack * sizeof(float));
        const vxl_ipc_handle mine = group.Handle();
        vxl_ipc_handle theirs;
        if (!write_all(to_peer, &mine, sizeof mine) || !read_all(from_peer, &theirs, sizeof theirs)) throw vxl::Error(VXL_ERR_INVALID, "handle exchange");
        std::vector<vxl_ipc_handle> all(2);
        all[rank] = mine; all[1 - rank] = theirs;
        group.Connect(all);
        char go = 1, got = 0;                                     // both members connected (and their stacks zeroed) before anyone stores
        if (!write_all(to_peer, &go, 1) || !read_all(from_peer, &got, 1)) throw vxl::Error(VXL_ERR_INVALID, "barrier");

        std::vector<float> host(stack);
        for (int frame = 0; frame < 2; ++frame) {
            view.Frame = frame;
            float* st = group.BeginFrame((uint64_t)frame);
            float* own = st + slot * (size_t)rank;
            vxl::LightAmbientPipeline::Get().Use(ctx, view, fb, vox, 2, own, own + plane);
            vxl::LightReflectionPipeline::Get().Use(ctx, view, fb, vox, own + 2 * plane);
            group.EndFrame();
            group.Fence();
            ctx.Download(host.data(), st, stack * sizeof(float));
            group.CheckArrived();
            rep.whole[frame] = fnv(host.data(), stack * sizeof(float));
            for (int r = 0; r < 2; ++r) {
                bool any = false;
                for (size_t i = 0; i < slot && !any; ++i) any = host[slot * r + i] != 0.0f;
                rep.slot[frame][r] = any ? fnv(host.data() + slot * r, slot * sizeof(float)) : 0ull;
            }
        }
        // nobody frees while the peer may still store or has the mapping open
        if (!write_all(to_peer, &go, 1) || !read_all(from_peer, &got, 1)) throw vxl::Error(VXL_ERR_INVALID, "barrier");
        group.Release();
        if (!write_all(to_peer, &go, 1) || !read_all(from_peer, &got, 1)) throw vxl::Error(VXL_ERR_INVALID, "barrier");
        ctx.Free(planes);
        rep.ok = 1;
    } catch (const vxl::Error& e) {
        std::fprintf(stderr, "rank %d: vxl::Error(%d): %s\n", rank, e.status, e.what());
        rep.ok = 0;
    }
    write_all(to_parent, &rep, sizeof rep);
    return rep.ok ? 0 : 3;
}

}  // namespace

int main(int argc, char** argv) {
    const int texels = argc > 1 ? std::atoi(argv[1]) : 64, W = argc > 2 ? std::atoi(argv[2]) : 256, H = argc > 3 ? std::atoi(argv[3]) : 128;
    const int ndev = argc > 4 ? std::atoi(argv[4]) : 1;
    int ab[2], ba[2], up[2][2];
    if (pipe(ab) || pipe(ba) || pipe(up[0]) || pipe(up[1])) { std::perror("pipe"); return 2; }
    pid_t pid[2];
    for (int r = 0; r < 2; ++r) {
        pid[r] = fork();                                          // before any CUDA call: each child initialises its own driver state
        if (pid[r] == 0) return member(r, 2, ndev > 1 ? r : 0, texels, W, H, r == 0 ? ab[1] : ba[1], r == 0 ? ba[0] : ab[0], up[r][1]);
    }
    Report rep[2];
    bool ok = true;
    for (int r = 0; r < 2; ++r) {
        ok = read_all(up[r][0], &rep[r], sizeof(Report)) && ok;
        int st = 0;
        waitpid(pid[r], &st, 0);
        ok = ok && WIFEXITED(st) && WEXITSTATUS(st) == 0 && rep[r].ok;
    }
    if (!ok) { std::fprintf(stderr, "a member failed\n"); return 3; }
    for (int f = 0; f < 2; ++f) {
        if (rep[0].whole[f] != rep[1].whole[f]) { std::fprintf(stderr, "frame %d: the two copies of the stack differ\n", f); return 4; }
        for (int r = 0; r < 2; ++r)
            if (rep[0].slot[f][r] == 0ull) { std::fprintf(stderr, "frame %d: slot of rank %d is empty\n", f, r); return 5; }
    }
    if (rep[0].whole[0] == rep[0].whole[1]) { std::fprintf(stderr, "the two frames are identical (noise frame index ignored?)\n"); return 6; }
    std::printf("OK %016llx %016llx\n", rep[0].whole[0], rep[0].whole[1]);
    return 0;
}
