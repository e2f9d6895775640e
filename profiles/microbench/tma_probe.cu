// Standalone check of the TMA tensor-map path used by pb_tiled.cuh.
//   tma_probe <elem_bytes> <box_elems> <box_rows> <x> <y> [W H]
// Encodes a 3-D map over a W x H x 3-byte image, loads one box at element coordinates (x, y)
// into shared memory and checks every byte (zero fill outside the image); then stores a 96 x 64
// byte box through a u8 map and checks it.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../photonbend_b200/csrc/pb_ptx.cuh"
using namespace pb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Args { CUtensorMap src, dst; int x, y, box_bytes; const CUtensorMap* gsrc; const CUtensorMap* gdst; unsigned* dbg; };

__global__ void probe(const __grid_constant__ Args a, unsigned char* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        ptx::mbarrier_init(&bar, 1);
        ptx::fence_mbarrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.dbg[0] = ptx::smem_addr(smem);
        a.dbg[1] = ptx::smem_addr(&bar);
        ptx::mbarrier_arrive_expect_tx(&bar, a.box_bytes);
        ptx::tma_load_3d(smem, a.gsrc ? (const void*)a.gsrc : (const void*)&a.src, a.x, a.y, 0, &bar);
    }
    ptx::mbarrier_wait(&bar, 0);
    for (int i = threadIdx.x; i < a.box_bytes; i += blockDim.x) out[i] = smem[i];
    __syncthreads();
    for (int i = threadIdx.x; i < 96 * 64; i += blockDim.x) smem[i] = (unsigned char)(i * 7 + 3);
    ptx::fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) { ptx::tma_store_3d(a.gdst ? (const void*)a.gdst : (const void*)&a.dst, 48, 10, 0, smem); ptx::bulk_commit(); ptx::bulk_wait_read0(); }
}

int main(int argc, char** argv) {
    const int eb = argc > 1 ? atoi(argv[1]) : 2, be = argc > 2 ? atoi(argv[2]) : 144, br = argc > 3 ? atoi(argv[3]) : 16;
    const int x = argc > 4 ? atoi(argv[4]) : 1, y = argc > 5 ? atoi(argv[5]) : 2;
    const int W = argc > 6 ? atoi(argv[6]) : 48, H = argc > 7 ? atoi(argv[7]) : 52;
    const long long pitch = W * 3;
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    unsigned char *src, *dst, *out;
    const int box_bytes = be * eb * br;
    cudaMalloc(&src, pitch * H); cudaMalloc(&dst, 288 * 100); cudaMalloc(&out, box_bytes);
    std::vector<unsigned char> h(pitch * H);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 13 + (i >> 8) + 1);
    cudaMemcpy(src, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMemset(dst, 0, 288 * 100);
    Args a; memset(&a, 0, sizeof(a)); a.x = x; a.y = y; a.box_bytes = box_bytes;
    {
        cuuint64_t dims[3] = {(cuuint64_t)pitch / eb, (cuuint64_t)H, 1}, strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(pitch * H)};
        cuuint32_t box[3] = {(cuuint32_t)be, (cuuint32_t)br, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&a.src, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : eb == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                         3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode src failed: %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[3] = {288, 100, 1}, strides[2] = {288, 28800};
        cuuint32_t box[3] = {96, 64, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&a.dst, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode dst failed: %d\n", (int)r); return 1; }
    }
    cudaMalloc(&a.dbg, 64);
    if (getenv("TMAP_GLOBAL")) {
        CUtensorMap* g; cudaMalloc(&g, 2 * sizeof(CUtensorMap));
        cudaMemcpy(g, &a.src, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        cudaMemcpy(g + 1, &a.dst, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        a.gsrc = g; a.gdst = g + 1;
    }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<<<1, 128, 65536>>>(a, out);
    printf("launch: %s; ", cudaGetErrorString(cudaGetLastError()));
    cudaError_t e = cudaDeviceSynchronize();
    printf("elem %d B, box %d x %d (%d B wide), at (%d,%d), image %dx%d: %s", eb, be, br, be * eb, x, y, W, H, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 0; }
    std::vector<unsigned char> ho(box_bytes); cudaMemcpy(ho.data(), out, box_bytes, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int r = 0; r < br; ++r) for (int b = 0; b < be * eb; ++b) {
        const long gx = (long)x * eb + b, gy = y + r;
        const unsigned char want = (gx < pitch && gy < H && gx >= 0 && gy >= 0) ? h[gy * pitch + gx] : 0;
        bad += ho[r * be * eb + b] != want;
    }
    std::vector<unsigned char> hd(288 * 100); cudaMemcpy(hd.data(), dst, hd.size(), cudaMemcpyDeviceToHost);
    long sbad = 0;
    for (int r = 0; r < 100; ++r) for (int b = 0; b < 288; ++b) {
        const int tr = r - 10, tb = b - 48;
        const unsigned char want = (tr >= 0 && tr < 64 && tb >= 0 && tb < 96) ? (unsigned char)((tr * 96 + tb) * 7 + 3) : 0;
        sbad += hd[r * 288 + b] != want;
    }
    unsigned dbg[2]; cudaMemcpy(dbg, a.dbg, 8, cudaMemcpyDeviceToHost);
    printf("; load mismatches %ld of %d, store mismatches %ld; smem@%u bar@%u first bytes %d %d %d %d\n", bad, box_bytes, sbad, dbg[0], dbg[1], ho[0], ho[1], ho[2], ho[3]);
    return 0;
}
