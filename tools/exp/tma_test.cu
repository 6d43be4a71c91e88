// experiment: stage a 3-D box of 32-bit words with cp.async.bulk.tensor and compare with a plain gather
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../../voxelengine_b200/csrc/vxl_tma.cuh"
using namespace vxl;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int TY>
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int z, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(raw);
    uint32_t* tile = reinterpret_cast<uint32_t*>(raw + 128);
    if (threadIdx.x == 0) mbar_init(bar, 1u);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, 4u * TY * TY * 4u); tma_load_3d(tile, &tm, x, y, z, bar); }
    mbar_wait(bar, 0u);
    for (int i = threadIdx.x; i < 4 * TY * TY; i += blockDim.x) out[i] = tile[i];
}
int main() {
    const int pitch = 12, cy = 256, cz = 256, TY = 69;
    std::vector<uint32_t> h((size_t)pitch * cy * cz);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)(i * 2654435761u);
    uint32_t *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 4 * TY * TY * 4);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry %p q %d\n", fn, (int)q);
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)cy, (cuuint64_t)cz};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4u, (cuuint64_t)pitch * cy * 4u};
    const cuuint32_t box[3] = {4u, (cuuint32_t)TY, (cuuint32_t)TY};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    const int smem = 128 + 4 * TY * TY * 4;
    cudaFuncSetAttribute(k<TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int t = 0; t < 3; ++t) {
        const int x = t == 0 ? 4 : (t == 1 ? -8 : 8), y = t == 0 ? 10 : (t == 1 ? -5 : 230), z = t == 0 ? 20 : (t == 1 ? -3 : 250);   // x (innermost): multiples of 4 words
        k<TY><<<1, 256, smem>>>(map, x, y, z, o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("run %d: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint32_t> g(4 * TY * TY);
        cudaMemcpy(g.data(), o, g.size() * 4, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int zz = 0; zz < TY; ++zz) for (int yy = 0; yy < TY; ++yy) for (int xx = 0; xx < 4; ++xx) {
            const int gx = x + xx, gy = y + yy, gz = z + zz;
            uint32_t want = 0;
            if (gx >= 0 && gx < pitch && gy >= 0 && gy < cy && gz >= 0 && gz < cz) want = h[((size_t)gz * cy + gy) * pitch + gx];
            if (g[(zz * TY + yy) * 4 + xx] != want) ++bad;
        }
        printf("  mismatches %zu\n", bad);
    }
    return 0;
}
