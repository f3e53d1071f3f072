// tma_probe.cu -- developer probe: which way of handing a CUtensorMap to cp.async.bulk.tensor works on this box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
struct Maps { CUtensorMap m[4]; };
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ void do_load(uint8_t* dst, uint64_t* bar, const CUtensorMap* map, int rank, int c0, int c1, int c2, uint32_t bytes)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(dst)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(dst)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s32(bar)), "r"(0u) : "memory");
    } while (!done);
}
__global__ void k_single(const __grid_constant__ CUtensorMap map, int rank, int c0, int c1, int c2, uint32_t bytes, uint32_t* out)
{
    __shared__ __align__(128) uint8_t buf[4096]; __shared__ uint64_t bar;
    do_load(buf, &bar, &map, rank, c0, c1, c2, bytes);
    uint32_t s = 0; for (int i = threadIdx.x; i < (int)bytes; i += 32) s += buf[i];
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (threadIdx.x == 0) *out = s;
}
__global__ void k_struct(const __grid_constant__ Maps maps, int idx, int rank, int c0, int c1, int c2, uint32_t bytes, uint32_t* out)
{
    __shared__ __align__(128) uint8_t buf[4096]; __shared__ uint64_t bar;
    do_load(buf, &bar, &maps.m[idx], rank, c0, c1, c2, bytes);
    uint32_t s = 0; for (int i = threadIdx.x; i < (int)bytes; i += 32) s += buf[i];
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (threadIdx.x == 0) *out = s;
}
__global__ void k_global(const CUtensorMap* map, int rank, int c0, int c1, int c2, uint32_t bytes, uint32_t* out)
{
    __shared__ __align__(128) uint8_t buf[4096]; __shared__ uint64_t bar;
    do_load(buf, &bar, map, rank, c0, c1, c2, bytes);
    uint32_t s = 0; for (int i = threadIdx.x; i < (int)bytes; i += 32) s += buf[i];
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (threadIdx.x == 0) *out = s;
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int report(const char* name, uint32_t* d_out, uint32_t expect)
{
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t h = 0;
    if (e == cudaSuccess) cudaMemcpy(&h, d_out, 4, cudaMemcpyDeviceToHost);
    printf("%-40s %s  sum=%u expect=%u %s\n", name, e == cudaSuccess ? "OK " : cudaGetErrorName(e), h, expect, (e == cudaSuccess && h == expect) ? "MATCH" : "");
    fflush(stdout);
    return e == cudaSuccess;
}
int main(int argc, char** argv)
{
    int which = argc > 1 ? atoi(argv[1]) : 0;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN enc = (PFN)fn;
    const int pitch = 1344, hp = 440, nimg = 4;
    size_t plane = (size_t)pitch * hp;
    uint8_t* d; cudaMalloc(&d, plane * nimg);
    uint8_t* h = (uint8_t*)malloc(plane * nimg);
    for (size_t i = 0; i < plane * nimg; i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    cudaMemcpy(d, h, plane * nimg, cudaMemcpyHostToDevice);
    uint32_t* d_out; cudaMalloc(&d_out, 4);
    auto expect = [&](int x, int y, int z, int bw, int bh) { uint32_t s = 0; for (int r = 0; r < bh; r++) for (int c = 0; c < bw; c++) s += h[(size_t)z * plane + (size_t)(y + r) * pitch + x + c]; return s; };
    CUtensorMap m2, m3; 
    { cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)hp * nimg}; cuuint64_t str[1] = {(cuuint64_t)pitch}; cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
      CUresult r = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); printf("encode 2d: %d\n", (int)r); }
    { cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)hp, (cuuint64_t)nimg}; cuuint64_t str[2] = {(cuuint64_t)pitch, (cuuint64_t)plane}; cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
      CUresult r = enc(&m3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); printf("encode 3d: %d\n", (int)r); }
    const int x = 37, y = 53, z = 1;
    if (which == 0) { k_single<<<1, 32>>>(m2, 2, x, y, 0, 1024, d_out); report("2d single grid_constant", d_out, expect(x, y, 0, 32, 32)); }
    if (which == 1) { k_single<<<1, 32>>>(m3, 3, x, y, z, 1024, d_out); report("3d single grid_constant", d_out, expect(x, y, z, 32, 32)); }
    if (which == 2) { Maps ms; memset(&ms, 0, sizeof(ms)); ms.m[2] = m3; k_struct<<<1, 32>>>(ms, 2, 3, x, y, z, 1024, d_out); report("3d struct[dyn idx] grid_constant", d_out, expect(x, y, z, 32, 32)); }
    if (which == 3) { CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &m3, sizeof(m3), cudaMemcpyHostToDevice); k_global<<<1, 32>>>(dm, 3, x, y, z, 1024, d_out); report("3d descriptor in global memory", d_out, expect(x, y, z, 32, 32)); }
    if (which == 4) { k_single<<<1, 32>>>(m3, 3, x + 1, y, z, 1024, d_out); report("3d unaligned x (odd)", d_out, expect(x + 1, y, z, 32, 32)); }
    if (which == 5) { k_single<<<1, 32>>>(m3, 3, 48, y, z, 1024, d_out); report("3d x multiple of 16", d_out, expect(48, y, z, 32, 32)); }
    return 0;
}
