// Microbenchmark: sustained shared -> global rate of TMA tensor-map box STORES shaped like the tiled
// remap kernel's tile stores (u8 map over a batch of 7680 x 3840 x 3 frames, box = `bw` bytes x `bh`
// rows), one CTA per tile walking the frames with two tiles in flight, no producer work at all --
// against plain coalesced 16-byte global stores of the same bytes.  Answers: is the 96-byte x
// 64-row tile store itself below the DRAM write rate (5.1 TB/s measured for the stores of the
// one-lens class alone), and would a wider box do better?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu ; ./store_bw
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../photonbend_b200/csrc/pb_ptx.cuh"
using namespace pb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Args {
    CUtensorMap map;
    int bw, bh, tiles_x, frames, band;
};

// one CTA per tile (raster order in bands of `band` tile rows, like the product), `frames` stores each
__global__ void __launch_bounds__(256) store_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tile_bytes = a.bw * a.bh;
    for (int i = threadIdx.x; i < 2 * tile_bytes / 4; i += blockDim.x) reinterpret_cast<unsigned*>(smem)[i] = i * 2654435761u;
    ptx::fence_async_smem();
    __syncthreads();
    int t = blockIdx.x, tx, ty;
    if (a.band > 0) {
        const int per_band = a.band * a.tiles_x;
        const int b = t / per_band, r = t - b * per_band;
        tx = r / a.band;
        ty = b * a.band + (r - tx * a.band);
    } else {
        ty = t / a.tiles_x;
        tx = t - ty * a.tiles_x;
    }
    if (threadIdx.x == 0) {
        for (int f = 0; f < a.frames; ++f) {
            ptx::bulk_wait_read1();
            ptx::tma_store_3d_hint(&a.map, tx * a.bw, ty * a.bh, f, smem + (f & 1) * tile_bytes, ptx::policy_evict_first());
            ptx::bulk_commit();
        }
        ptx::bulk_wait_read0();
    }
}

__global__ void __launch_bounds__(256) plain_kernel(uint4* dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        dst[i] = make_uint4((unsigned)i, 1u, 2u, 3u);
}

int main() {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const int W = 7680, H = 3840, frames = 16;
    const size_t pitch = (size_t)W * 3, frame_bytes = pitch * H;
    unsigned char* dst;
    cudaMalloc(&dst, frame_bytes * frames);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const double bytes = (double)frame_bytes * frames;
    float ms = 0;
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        plain_kernel<<<148 * 8, 256>>>(reinterpret_cast<uint4*>(dst), frame_bytes * frames / 16);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("plain 16-byte stores, grid-stride:            %8.1f GB/s\n", bytes / ms / 1e6);
    const int cfgs[][3] = {{96, 64, 16}, {96, 64, 0}, {96, 64, 4}, {192, 32, 16}, {384, 16, 16}, {384, 16, 0}, {768, 8, 16}, {128, 48, 16}, {256, 24, 16}};
    for (auto& c : cfgs) {
        Args a;
        memset(&a, 0, sizeof(a));
        a.bw = c[0]; a.bh = c[1]; a.band = c[2]; a.frames = frames;
        a.tiles_x = (int)(pitch / a.bw);
        const int tiles_y = H / a.bh;
        if (a.band > 0 && tiles_y % a.band) a.band = 0;
        cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)H, (cuuint64_t)frames};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_bytes};
        cuuint32_t box[3] = {(cuuint32_t)a.bw, (cuuint32_t)a.bh, 1}, es[3] = {1, 1, 1};
        if (enc(&a.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
            printf("encode failed for %d x %d\n", a.bw, a.bh);
            continue;
        }
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            store_kernel<<<a.tiles_x * tiles_y, 256, 2 * a.bw * a.bh>>>(a);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("TMA box %4d B x %2d rows, raster band %2d:       %8.1f GB/s   %s\n", a.bw, a.bh, a.band, bytes / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
