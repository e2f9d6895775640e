// Microbenchmark: sustained global -> shared ingest rate of TMA tensor-map box loads shaped like
// the tiled remap kernel's stage loads (u16 map over a 3-byte-per-pixel image, box = pitch x 16
// rows, `boxes` boxes per item), with `nbuf` items in flight per CTA and no consumer at all.
// Answers: what is the ceiling of the staging path itself, from L2-resident and from
// DRAM-resident sources, as a function of bytes in flight per SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_pipe tma_pipe.cu ; ./tma_pipe
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
    int pitch, boxes, nbuf, items, W, H, frames, warps;
};

__global__ void __launch_bounds__(256) pipe_kernel(const __grid_constant__ Args a, unsigned* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[16];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < a.nbuf; ++i) ptx::mbarrier_init(&bar[i], 1);
        ptx::fence_mbarrier_init();
    }
    __syncthreads();
    const int buf_bytes = a.boxes * 16 * a.pitch;
    // warps = 1: thread 0 issues every box; warps = k: warp w issues boxes w, w+k, ...
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0 && warp < a.warps) {
        unsigned h = blockIdx.x * 2654435761u + 12345u;
        for (int i = 0; i < a.items; ++i) {
            const int b = i % a.nbuf;
            if (i >= a.nbuf) ptx::mbarrier_wait(&bar[b], ((i / a.nbuf) - 1) & 1);
            h = h * 1664525u + 1013904223u;
            const int y0 = (h >> 8) % (a.H - a.boxes * 16);
            const int x0 = (((h >> 4) * 7u) % ((a.W * 3 - a.pitch) / 16)) * 8;  // u16 elements, 16-byte aligned
            const int f = i % a.frames;
            if (warp == 0) ptx::mbarrier_arrive_expect_tx(&bar[b], buf_bytes);
            for (int k = warp; k < a.boxes; k += a.warps)
                ptx::tma_load_3d(smem + b * buf_bytes + k * 16 * a.pitch, &a.map, x0, y0 + k * 16, f, &bar[b]);
        }
        if (warp == 0)
            for (int i = a.items; i < a.items + a.nbuf; ++i)
                if (i >= a.nbuf) ptx::mbarrier_wait(&bar[i % a.nbuf], ((i / a.nbuf) - 1) & 1);
    }
    __syncthreads();
    if (smem[tid] == 0x7b && sink) *sink = 1;
}

int main() {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const int W = 3840, H = 3840, frames_max = 16;
    const size_t frame_bytes = (size_t)W * 3 * H;
    unsigned char* src;
    cudaMalloc(&src, frame_bytes * frames_max);
    cudaMemset(src, 1, frame_bytes * frames_max);
    unsigned* sink;
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("pitch boxes nbuf ctas/SM frames warps |  KB in flight/SM   GB/s   us/item\n");
    const int cfgs[][6] = {
        // pitch, boxes, nbuf, ctas per SM, frames (1 = L2 resident, 16 = DRAM), issuing warps
        {208, 4, 2, 5, 1, 1},  {208, 4, 2, 5, 16, 1},  {208, 4, 4, 3, 1, 1},  {208, 4, 4, 3, 16, 1},
        {208, 4, 8, 1, 1, 1},  {208, 4, 8, 1, 16, 1},  {208, 4, 8, 2, 1, 1},  {208, 4, 8, 2, 16, 1},
        {208, 4, 8, 2, 16, 4}, {208, 4, 12, 1, 16, 4},
        {272, 5, 2, 3, 1, 1},  {272, 5, 2, 3, 16, 1},  {272, 5, 4, 2, 1, 1},  {272, 5, 4, 2, 16, 1},
        {272, 5, 8, 1, 16, 1}, {272, 5, 8, 1, 16, 5},  {272, 5, 4, 2, 16, 5},
        {144, 4, 8, 2, 16, 1}, {496, 4, 4, 2, 16, 1},  {496, 2, 8, 2, 16, 1}, {112, 8, 8, 2, 16, 1},
    };
    for (auto& c : cfgs) {
        Args a;
        memset(&a, 0, sizeof(a));
        a.pitch = c[0]; a.boxes = c[1]; a.nbuf = c[2]; a.frames = c[4]; a.warps = c[5];
        a.W = W; a.H = H; a.items = 400;
        const int ctas = c[3];
        cuuint64_t dims[3] = {(cuuint64_t)W * 3 / 2, (cuuint64_t)H, (cuuint64_t)frames_max};
        cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)frame_bytes};
        cuuint32_t box[3] = {(cuuint32_t)a.pitch / 2, 16, 1}, es[3] = {1, 1, 1};
        if (enc(&a.map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
            printf("encode failed\n");
            return 1;
        }
        const int smem = a.nbuf * a.boxes * 16 * a.pitch;
        float ms = 0;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            pipe_kernel<<<148 * ctas, 256, smem>>>(a, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double bytes = (double)148 * ctas * a.items * a.boxes * 16 * a.pitch;
        printf("%5d %5d %4d %7d %6d %5d | %10.1f %10.1f %8.3f   %s\n", a.pitch, a.boxes, a.nbuf, ctas, a.frames, a.warps,
               ctas * smem / 1024.0, bytes / ms / 1e6, ms * 1e3 / a.items * 1.0, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
