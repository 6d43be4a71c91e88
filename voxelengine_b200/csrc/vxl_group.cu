// vxl_group.cu -- the frame sharded over the GPUs of one box, driven from C ("vxl_ctx_create(ndev) / vxl_gather", 8e).
//
// One process (or thread) per GPU holds one vxl_group member.  Every member owns ONE allocation: n_stacks copies of the gathered
// tile stack followed by a page of arrival flags.  The allocations are mapped into every member (CUDA IPC between processes, plain
// peer pointers inside one process), so
//   * the light-pass kernels store each output value into every member's stack (vxl_ctx_set_output_mirrors; the gather of
//     SURVEY 8e happens inside the passes, over NVLink peer stores), and
//   * the frame is closed without a collective: vxl_group_fence launches one tiny kernel that writes this member's frame
//     number into its slot of every peer's flag page (st.release.sys, stream-ordered behind the passes, so the peer stores of the
//     frame are complete when the flag lands) and then waits until all slots of its own page carry the frame number
//     (ld.acquire.sys).  No NCCL launch, no host round trip on a 1 ms frame.
// With n_stacks = 2 consecutive frames alternate between the stacks: a member that passes the fence of frame N+1 knows every peer
// has queued -- behind its consumer of frame N, in stream order -- the passes of frame N+1, so stack N & 1 is free again by the
// time frame N+2 is stored (the read-after-write hazard of a single stack).
// The reference has no counterpart (one GPU, one process); the call sites this serves are WorldRenderer.cpp:239-274.
#include "vxl_internal.h"

#include <cstring>

struct vxl_group {
    vxl_ctx* ctx = nullptr;
    int rank = 0, n = 1, n_stacks = 1;
    size_t stack_bytes = 0, flag_off = 0, total = 0;
    char* base = nullptr;                       // this member's allocation
    char* peer[VXL_MAX_MIRRORS + 1] = {nullptr};    // every member's allocation as seen from here (peer[rank] == base)
    bool opened[VXL_MAX_MIRRORS + 1] = {false};     // mapped through vxl_ipc_open (to be closed)
    bool connected = false;
    unsigned long long seq = 0;                 // fences issued so far
    int* d_status = nullptr;                    // device word: 0 ok, else the rank a fence timed out on + 1
};

namespace vxl {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct FencePeers { unsigned long long* flags[VXL_MAX_MIRRORS + 1]; };   // every member's flag page

// thread r: tell member r that `rank` has finished frame `seq`, then wait for member r's word in our own page
__global__ void k_group_fence(FencePeers P, int rank, int n, unsigned long long seq, unsigned long long timeout_ns, int* status) {
    const int r = (int)threadIdx.x;
    if (r >= n) return;
    if (r != rank) st_release_sys(P.flags[r] + rank, seq);
    const unsigned long long* mine = P.flags[rank] + r;
    if (r == rank) return;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(mine) < seq) {
        if (global_timer_ns() - t0 > timeout_ns) { atomicCAS(status, 0, r + 1); return; }     // a peer never arrived: report, do not hang the GPU
        __nanosleep(64);
    }
}

}  // namespace vxl

using namespace vxl;

extern "C" {

int vxl_group_create(vxl_ctx* c, int rank, int n_ranks, size_t stack_bytes, int n_stacks, vxl_group** out) {
    if (!c || !out || n_ranks < 1 || n_ranks > VXL_MAX_MIRRORS + 1 || rank < 0 || rank >= n_ranks || stack_bytes == 0 || (n_stacks != 1 && n_stacks != 2)) {
        set_error("vxl_group_create: bad argument (1 <= n_ranks <= 16, 0 <= rank < n_ranks, n_stacks 1 or 2)");
        return VXL_ERR_INVALID;
    }
    VXL_CUDA(cudaSetDevice(c->device));
    vxl_group* g = new (std::nothrow) vxl_group();
    if (!g) return VXL_ERR_OOM;
    g->ctx = c; g->rank = rank; g->n = n_ranks; g->n_stacks = n_stacks;
    g->stack_bytes = (stack_bytes + 255) & ~(size_t)255;
    g->flag_off = g->stack_bytes * (size_t)n_stacks;
    g->total = g->flag_off + 256;
    cudaError_t e = cudaMalloc(&g->base, g->total);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_status, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(g->base, 0, g->total, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_status, 0, sizeof(int), c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cudaFree(g->base); cudaFree(g->d_status); delete g; return cuda_fail(e, "vxl_group_create"); }
    g->peer[rank] = g->base;
    *out = g;
    return VXL_OK;
}

int vxl_group_handle(vxl_group* g, vxl_ipc_handle* out) {
    if (!g || !out) { set_error("vxl_group_handle: bad argument"); return VXL_ERR_INVALID; }
    return vxl_ipc_export(g->ctx, g->base, out);
}

static int finish_connect(vxl_group* g) {
    g->connected = true;
    return VXL_OK;
}

int vxl_group_connect(vxl_group* g, const vxl_ipc_handle* handles) {
    if (!g || (g->n > 1 && !handles)) { set_error("vxl_group_connect: bad argument"); return VXL_ERR_INVALID; }
    if (g->connected) { set_error("vxl_group_connect: already connected"); return VXL_ERR_INVALID; }
    for (int r = 0; r < g->n; ++r) {
        if (r == g->rank) continue;
        void* p = nullptr;
        if (int e = vxl_ipc_open(g->ctx, &handles[r], &p)) {
            for (int q = 0; q < r; ++q)
                if (g->opened[q]) { vxl_ipc_close(g->ctx, g->peer[q]); g->opened[q] = false; g->peer[q] = nullptr; }
            return e;
        }
        g->peer[r] = (char*)p; g->opened[r] = true;
    }
    return finish_connect(g);
}

int vxl_group_connect_pointers(vxl_group* g, void* const* bases) {
    if (!g || (g->n > 1 && !bases)) { set_error("vxl_group_connect_pointers: bad argument"); return VXL_ERR_INVALID; }
    if (g->connected) { set_error("vxl_group_connect_pointers: already connected"); return VXL_ERR_INVALID; }
    for (int r = 0; r < g->n; ++r) {
        if (r == g->rank) continue;
        if (!bases[r]) { set_error("vxl_group_connect_pointers: NULL peer"); return VXL_ERR_INVALID; }
        g->peer[r] = (char*)bases[r];
    }
    return finish_connect(g);
}

int vxl_group_base(vxl_group* g, void** out_dev) {
    if (!g || !out_dev) { set_error("vxl_group_base: bad argument"); return VXL_ERR_INVALID; }
    *out_dev = g->base;
    return VXL_OK;
}

int vxl_group_stack(vxl_group* g, int which, void** out_dev) {
    if (!g || !out_dev || which < 0 || which >= g->n_stacks) { set_error("vxl_group_stack: bad argument"); return VXL_ERR_INVALID; }
    *out_dev = g->base + g->stack_bytes * (size_t)which;
    return VXL_OK;
}

int vxl_group_begin_frame(vxl_group* g, uint64_t frame, void** out_stack_dev) {
    if (!g) { set_error("vxl_group_begin_frame: group is NULL"); return VXL_ERR_INVALID; }
    if (!g->connected) { set_error("vxl_group_begin_frame: not connected"); return VXL_ERR_INVALID; }
    int64_t deltas[VXL_MAX_MIRRORS];
    int k = 0;
    for (int r = 0; r < g->n; ++r)
        if (r != g->rank) deltas[k++] = (int64_t)(g->peer[r] - g->base);       // the same stack of the same frame sits at the same offset in every copy
    if (out_stack_dev) *out_stack_dev = g->base + g->stack_bytes * (size_t)(frame % (uint64_t)g->n_stacks);
    return vxl_ctx_set_output_mirrors(g->ctx, k, deltas);
}

int vxl_group_end_frame(vxl_group* g) {
    if (!g) { set_error("vxl_group_end_frame: group is NULL"); return VXL_ERR_INVALID; }
    return vxl_ctx_set_output_mirrors(g->ctx, 0, nullptr);
}

int vxl_group_fence(vxl_group* g) {
    if (!g) { set_error("vxl_group_fence: group is NULL"); return VXL_ERR_INVALID; }
    if (!g->connected) { set_error("vxl_group_fence: not connected"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaSetDevice(g->ctx->device));
    ++g->seq;
    if (g->n == 1) return VXL_OK;
    FencePeers P;
    for (int r = 0; r <= VXL_MAX_MIRRORS; ++r) P.flags[r] = r < g->n ? reinterpret_cast<unsigned long long*>(g->peer[r] + g->flag_off) : nullptr;
    k_group_fence<<<1, 32, 0, g->ctx->stream>>>(P, g->rank, g->n, g->seq, 5ull * 1000ull * 1000ull * 1000ull, g->d_status);
    VXL_LAUNCH_CHECK(g->ctx);
    return VXL_OK;
}

int vxl_group_status(vxl_group* g, int* out_rank_plus_one) {
    if (!g || !out_rank_plus_one) { set_error("vxl_group_status: bad argument"); return VXL_ERR_INVALID; }
    VXL_CUDA(cudaMemcpyAsync(out_rank_plus_one, g->d_status, sizeof(int), cudaMemcpyDeviceToHost, g->ctx->stream));
    VXL_CUDA(cudaStreamSynchronize(g->ctx->stream));
    if (*out_rank_plus_one != 0) { set_error("vxl_group_fence: timed out waiting for rank " + std::to_string(*out_rank_plus_one - 1)); return VXL_ERR_CUDA; }
    return VXL_OK;
}

int vxl_group_destroy(vxl_group* g) {
    if (!g) return VXL_OK;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    vxl_ctx_set_output_mirrors(g->ctx, 0, nullptr);
    for (int r = 0; r < g->n; ++r)
        if (g->opened[r]) vxl_ipc_close(g->ctx, g->peer[r]);
    cudaFree(g->d_status);
    cudaFree(g->base);
    delete g;
    return VXL_OK;
}

}  // extern "C"
