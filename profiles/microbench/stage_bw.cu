// Microbenchmark: how fast can one SM pull many small row segments from L2-resident global
// memory into shared memory?  (a) one TMA bulk copy (cp.async.bulk) per row segment,
// (b) cp.async 16-byte (LDGSTS) per thread, (c) LDG.128 + STS.128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stage_bw stage_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../photonbend_b200/csrc/pb_ptx.cuh"

using namespace pb;

constexpr int kPitch = 11520;   // bytes per source row (3840 px * 3)
constexpr int kRows = 3840;

template <int MODE>
__global__ void __launch_bounds__(256) stage_kernel(const unsigned char* __restrict__ src, int rows_per_tile,
                                                    int row_bytes, int tiles_per_cta, unsigned* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        ptx::mbarrier_init(&bar, 1);
        ptx::fence_mbarrier_init();
    }
    __syncthreads();
    unsigned acc = 0;
    for (int t = 0; t < tiles_per_cta; ++t) {
        const int tile = blockIdx.x * tiles_per_cta + t;
        const int y0 = (tile * 37) % (kRows - rows_per_tile);
        const int xb = ((tile * 53) % ((kPitch - row_bytes) / 16)) * 16;
        if (MODE == 0) {
            if (tid == 0) ptx::mbarrier_arrive_expect_tx(&bar, rows_per_tile * row_bytes);
            if (tid < rows_per_tile)
                ptx::bulk_g2s(smem + tid * row_bytes, src + (size_t)(y0 + tid) * kPitch + xb, row_bytes, &bar);
            ptx::mbarrier_wait(&bar, t & 1);
        } else {
            const int vecs = row_bytes / 16;
            const int tx = tid & 15, ty = tid >> 4;
            for (int r = ty; r < rows_per_tile; r += 16)
                for (int v = tx; v < vecs; v += 16) {
                    const unsigned char* g = src + (size_t)(y0 + r) * kPitch + xb + v * 16;
                    unsigned char* s = smem + r * row_bytes + v * 16;
                    if (MODE == 1) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_addr(s)), "l"(g) : "memory");
                    } else {
                        *reinterpret_cast<int4*>(s) = __ldg(reinterpret_cast<const int4*>(g));
                    }
                }
            if (MODE == 1) asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
        }
        acc += reinterpret_cast<unsigned*>(smem)[(tid * 7) % (rows_per_tile * row_bytes / 4)];
        __syncthreads();
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    unsigned char* src;
    unsigned* sink;
    cudaMalloc(&src, (size_t)kPitch * kRows);
    cudaMemset(src, 1, (size_t)kPitch * kRows);
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int tiles_per_cta = 64, grid = 148 * 4;
    const int cfgs[][2] = {{64, 192}, {64, 256}, {32, 384}, {96, 128}, {16, 1024}};
    for (auto& c : cfgs) {
        const int rows = c[0], rb = c[1];
        const int smem = rows * rb;
        for (int mode = 0; mode < 3; ++mode) {
            auto k = mode == 0 ? stage_kernel<0> : mode == 1 ? stage_kernel<1> : stage_kernel<2>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                k<<<grid, 256, smem>>>(src, rows, rb, tiles_per_cta, sink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)grid * tiles_per_cta * rows * rb;
            const double ops = (double)grid * tiles_per_cta * rows;
            printf("rows %3d x %4d B  mode %d (%s): %.3f ms  %.1f GB/s  %.2f us per tile-load per SM-slot, %.1f ns/row-op/SM\n",
                   rows, rb, mode, mode == 0 ? "TMA bulk/row" : mode == 1 ? "cp.async16 " : "LDG+STS    ", ms,
                   bytes / ms / 1e6, ms * 1e3 / (tiles_per_cta * 4.0), ms * 1e6 / (ops / 148.0));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
