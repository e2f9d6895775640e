// host_copy_ceiling.cu -- what the host of this box can feed: pinned host <-> device copies of
// 8K frames (88,473,600 bytes, one cfg5 frame) on N GPUs at once, each way alone and both ways
// together.  This is the ceiling of bench.py's `e2e` figure (every frame crosses PCIe twice); the
// remap kernel is not involved.
//
//   nvcc -O2 -o host_copy_ceiling host_copy_ceiling.cu && ./host_copy_ceiling [n_gpus] [seconds] [h2d_fraction] [rects]
//
// h2d_fraction < 1: every upload carries only that share of a frame (what an upload restricted to
// the source pixels a plan reads would move); rects > 0: as that many 2-D copies (row bands of a
// narrower rectangle) instead of one contiguous block.
//
// One JSON line per (n_gpus, mode): aggregate GB/s each way and the Gpix/s of an 8K RGB stream that
// bandwidth carries (one frame in + one frame out per 29.49 Mpix).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));    \
            std::exit(1);                                                                      \
        }                                                                                      \
    } while (0)

constexpr size_t kFrame = 3840ull * 7680ull * 3ull;
constexpr int kDepth = 3;  // frames in flight per GPU and direction (the pipeline's depth)
static double g_h2d_fraction = 1.0;
static int g_rects = 0;

struct Gpu {
    cudaStream_t up[kDepth], down[kDepth];
    unsigned char* h_in[kDepth];
    unsigned char* h_out[kDepth];
    unsigned char* d_in[kDepth];
    unsigned char* d_out[kDepth];
};

// rects = -1: the upload as a KERNEL that reads the pinned host frame in place (mapped memory),
// row by row, only the first h2d_fraction of every row -- spans of the rows instead of a block
__global__ void __launch_bounds__(256) upload_rows_kernel(const uint4* __restrict__ host, uint4* __restrict__ dev, int pitch16,
                                                          int width16, int rows) {
    const long long n = (long long)width16 * rows;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / width16), c = (int)(i - (long long)r * width16);
        dev[(long long)r * pitch16 + c] = host[(long long)r * pitch16 + c];
    }
}

static double run(std::vector<Gpu>& gpus, bool h2d, bool d2h, double seconds, long long* frames_each_way) {
    const int n = (int)gpus.size();
    long long copies = 0;
    for (int g = 0; g < n; ++g) {
        CK(cudaSetDevice(g));
        CK(cudaDeviceSynchronize());
    }
    const auto t0 = std::chrono::steady_clock::now();
    double dt = 0;
    do {
        for (int round = 0; round < 4; ++round)
            for (int g = 0; g < n; ++g) {
                CK(cudaSetDevice(g));
                for (int k = 0; k < kDepth; ++k) {
                    if (h2d && g_rects < 0) {
                        const int pitch16 = 7680 * 3 / 16, width16 = (int)(pitch16 * g_h2d_fraction);
                        upload_rows_kernel<<<-g_rects * 148, 256, 0, gpus[g].up[k]>>>(
                            reinterpret_cast<const uint4*>(gpus[g].h_in[k]), reinterpret_cast<uint4*>(gpus[g].d_in[k]), pitch16,
                            width16, 3840);
                    } else if (h2d && g_rects == 0) {
                        const size_t bytes = (size_t)((double)kFrame * g_h2d_fraction) & ~(size_t)255;
                        CK(cudaMemcpyAsync(gpus[g].d_in[k], gpus[g].h_in[k], bytes, cudaMemcpyHostToDevice, gpus[g].up[k]));
                    } else if (h2d) {
                        const size_t pitch = 7680ull * 3ull, rows = 3840 / g_rects;
                        const size_t width = (size_t)((double)pitch * g_h2d_fraction) & ~(size_t)15;
                        for (int r = 0; r < g_rects; ++r)
                            CK(cudaMemcpy2DAsync(gpus[g].d_in[k] + r * rows * pitch + 16 * r, pitch,
                                                 gpus[g].h_in[k] + r * rows * pitch + 16 * r, pitch, width, rows,
                                                 cudaMemcpyHostToDevice, gpus[g].up[k]));
                    }
                    if (d2h) CK(cudaMemcpyAsync(gpus[g].h_out[k], gpus[g].d_out[k], kFrame, cudaMemcpyDeviceToHost, gpus[g].down[k]));
                }
            }
        copies += 4 * kDepth;
        for (int g = 0; g < n; ++g) {
            CK(cudaSetDevice(g));
            CK(cudaDeviceSynchronize());
        }
        dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } while (dt < seconds);
    *frames_each_way = copies * n;
    return dt;
}

int main(int argc, char** argv) {
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    int n = argc > 1 ? std::atoi(argv[1]) : n_dev;
    if (n > n_dev) n = n_dev;
    const double seconds = argc > 2 ? std::atof(argv[2]) : 2.0;
    if (argc > 3) g_h2d_fraction = std::atof(argv[3]);
    if (argc > 4) g_rects = std::atoi(argv[4]);
    std::vector<Gpu> gpus(n);
    for (int g = 0; g < n; ++g) {
        CK(cudaSetDevice(g));
        for (int k = 0; k < kDepth; ++k) {
            CK(cudaStreamCreateWithFlags(&gpus[g].up[k], cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&gpus[g].down[k], cudaStreamNonBlocking));
            CK(cudaHostAlloc((void**)&gpus[g].h_in[k], kFrame, cudaHostAllocDefault));
            CK(cudaHostAlloc((void**)&gpus[g].h_out[k], kFrame, cudaHostAllocDefault));
            CK(cudaMalloc((void**)&gpus[g].d_in[k], kFrame));
            CK(cudaMalloc((void**)&gpus[g].d_out[k], kFrame));
            for (size_t b = 0; b < kFrame; b += 4096) gpus[g].h_in[k][b] = (unsigned char)b;  // touch the pages
        }
    }
    const char* names[3] = {"h2d", "d2h", "both"};
    for (int mode = 0; mode < 3; ++mode) {
        const bool h2d = mode != 1, d2h = mode != 0;
        long long frames = 0;
        run(gpus, h2d, d2h, 0.3, &frames);  // warm-up
        const double dt = run(gpus, h2d, d2h, seconds, &frames);
        const double gbs = (double)frames * (double)kFrame / dt / 1e9;
        // an e2e stream needs one frame up and one frame down per output frame
        const double gpix = (double)frames * 3840.0 * 7680.0 / dt / 1e9;
        std::printf("{\"n_gpus\": %d, \"h2d_fraction\": %.2f, \"rects\": %d, \"mode\": \"%s\", \"GBps_each_way\": %.2f, \"frames_per_s_each_way\": %.1f, "
                    "\"gpix_per_s_if_stream\": %.2f, \"seconds\": %.2f}\n",
                    n, g_h2d_fraction, g_rects, names[mode], gbs, frames / dt, mode == 2 ? gpix : 0.0, dt);
        std::fflush(stdout);
    }
    return 0;
}
