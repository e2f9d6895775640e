// pb_io.cpp -- libpbio.so: JPEG <-> device uint8 HWC tensors through nvJPEG (see include/pb_io.h).
//
// Build (photonbend_b200/build.py):  g++ -O2 -fPIC -shared -I<cuda>/include -Iinclude pb_io.cpp
//                                        -L<cuda>/lib64 -lnvjpeg -lcudart -o libpbio.so
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nvjpeg.h>

#include "pb_io.h"

namespace {

thread_local std::string g_error;
std::atomic<long long> g_single_state_decodes{0};  // decodes that went through nvjpegDecode on the device's one state

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

const char* status_name(nvjpegStatus_t s) {
    switch (s) {
        case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
        case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
        case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
        case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "jpeg not supported";
        case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
        case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
        case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
        case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
        case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "implementation not supported";
        default: return "unknown status";
    }
}

int codec_fail(nvjpegStatus_t s, const char* what) {
    return fail(s == NVJPEG_STATUS_JPEG_NOT_SUPPORTED || s == NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED
                    ? PB_IO_ERR_UNSUPPORTED
                    : PB_IO_ERR_CODEC,
                std::string(what) + ": nvjpeg " + status_name(s));
}

// one handle, one decoder state and one encoder state per DEVICE (created on first use with that
// device current), each serialised by its own mutex: a caller decodes / encodes one image at a
// time per GPU; the host threads that drive different GPUs do not wait for each other
struct Codec {
    std::mutex lock;      // start-up and the decoder state
    std::mutex enc_lock;  // the encoder state: a host thread may encode while another one decodes
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t dec_state = nullptr;
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
    bool tried = false;
    nvjpegStatus_t init_status = NVJPEG_STATUS_NOT_INITIALIZED;
    std::mutex pool_lock;                      // the decoupled decoders nobody is using (see ThreadDecoder)
    std::vector<struct ThreadDecoder*> idle;

    nvjpegStatus_t ensure() {
        if (tried) return init_status;
        tried = true;
        // PB_IO_BACKEND: hybrid (Huffman stages on the host), gpu (Huffman stages on the device: large
        // images), hardware (the JPEG engines, baseline single-scan images); default: the library's choice
        nvjpegBackend_t backend = NVJPEG_BACKEND_DEFAULT;
        if (const char* e = std::getenv("PB_IO_BACKEND")) {
            if (!std::strcmp(e, "hybrid")) backend = NVJPEG_BACKEND_HYBRID;
            else if (!std::strcmp(e, "gpu")) backend = NVJPEG_BACKEND_GPU_HYBRID;
            else if (!std::strcmp(e, "hardware")) backend = NVJPEG_BACKEND_HARDWARE;
        }
        init_status = backend == NVJPEG_BACKEND_DEFAULT ? nvjpegCreateSimple(&handle)
                                                         : nvjpegCreateEx(backend, nullptr, nullptr, 0, &handle);
        if (init_status == NVJPEG_STATUS_SUCCESS) init_status = nvjpegJpegStateCreate(handle, &dec_state);
        return init_status;
    }
    nvjpegStatus_t ensure_encoder(cudaStream_t st) {
        if (enc_state) return NVJPEG_STATUS_SUCCESS;
        nvjpegStatus_t s = nvjpegEncoderStateCreate(handle, &enc_state, st);
        if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncoderParamsCreate(handle, &enc_params, st);
        return s;
    }
};

constexpr int kMaxDevices = 64;

// The decoupled decoder (nvjpegDecodeJpegHost / TransferToDevice / Device) with its own state and
// buffers per CONCURRENT CALL on a device (a pool: a call takes an idle one or makes one, and puts
// it back): several host threads decode at once on one GPU, and with the GPU_HYBRID implementation
// the Huffman stage runs on the device instead of one host core.
struct ThreadDecoder {
    nvjpegJpegDecoder_t decoder = nullptr;
    nvjpegJpegState_t state = nullptr;
    nvjpegBufferPinned_t pinned = nullptr;
    nvjpegBufferDevice_t device = nullptr;
    nvjpegJpegStream_t stream = nullptr;
    nvjpegDecodeParams_t params = nullptr;
    cudaEvent_t done = nullptr;  // the last decode through this state has left its pinned buffer
    bool tried = false;
    nvjpegStatus_t status = NVJPEG_STATUS_NOT_INITIALIZED;

    nvjpegStatus_t ensure(nvjpegHandle_t h, nvjpegBackend_t backend) {
        if (tried) return status;
        tried = true;
        status = nvjpegDecoderCreate(h, backend, &decoder);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegDecoderStateCreate(h, decoder, &state);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegBufferPinnedCreate(h, nullptr, &pinned);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegBufferDeviceCreate(h, nullptr, &device);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegJpegStreamCreate(h, &stream);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegDecodeParamsCreate(h, &params);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegDecodeParamsSetOutputFormat(params, NVJPEG_OUTPUT_RGBI);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegStateAttachPinnedBuffer(state, pinned);
        if (status == NVJPEG_STATUS_SUCCESS) status = nvjpegStateAttachDeviceBuffer(state, device);
        // (blocking: a thread that waits for its decode sleeps instead of spinning on a core the other
        // decode threads -- and the other ranks of a multi-GPU box -- need)
        if (status == NVJPEG_STATUS_SUCCESS &&
            cudaEventCreateWithFlags(&done, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess)
            status = NVJPEG_STATUS_ALLOCATOR_FAILURE;
        return status;
    }
};

// PB_IO_DECODER: "gpu" (default) = decoupled decoder, Huffman stage on the device (one 8K frame: 8 ms
// against 15 ms, and four host threads decode 9.5 Gpix/s on one GPU against 2.0); "threads" =
// decoupled decoder, Huffman stage on the calling host thread; "single" = nvjpegDecode with the one
// state of the device, calls serialised.  Images the decoupled decoder refuses (progressive, four
// components ...) take the single-state path whatever the mode.
int decoder_mode() {
    static const int mode = [] {
        const char* e = std::getenv("PB_IO_DECODER");
        if (e && !std::strcmp(e, "single")) return 0;
        if (e && !std::strcmp(e, "threads")) return 2;
        return 1;
    }();
    return mode;
}

Codec& codec() {
    static Codec* c = new Codec[kMaxDevices];  // never destroyed: the CUDA context may already be gone at exit
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return c[dev];
}

}  // namespace

extern "C" {

int pb_io_version(void) { return 1; }

long long pb_io_single_state_decodes(void) { return g_single_state_decodes.load(std::memory_order_relaxed); }

const char* pb_io_last_error(void) { return g_error.c_str(); }

int pb_io_jpeg_info(const uint8_t* jpeg, size_t jpeg_bytes, int32_t* width, int32_t* height, int32_t* components) {
    if (!jpeg || jpeg_bytes == 0 || !width || !height || !components)
        return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_info: null pointer or empty buffer");
    Codec& c = codec();
    std::lock_guard<std::mutex> guard(c.lock);
    nvjpegStatus_t s = c.ensure();
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_info (nvjpeg start-up)");
    int n = 0, widths[NVJPEG_MAX_COMPONENT] = {0}, heights[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t css;
    s = nvjpegGetImageInfo(c.handle, jpeg, jpeg_bytes, &n, &css, widths, heights);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_info");
    *width = widths[0];
    *height = heights[0];
    *components = n;
    return PB_IO_OK;
}

int pb_io_jpeg_decode_rgb_u8(const uint8_t* jpeg, size_t jpeg_bytes, uint8_t* dst, int32_t width, int32_t height,
                             void* stream) {
    if (!jpeg || jpeg_bytes == 0 || !dst)
        return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_decode_rgb_u8: null pointer or empty buffer");
    if (width < 1 || height < 1) return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_decode_rgb_u8: bad size");
    Codec& c = codec();
    std::unique_lock<std::mutex> guard(c.lock);
    nvjpegStatus_t s = c.ensure();
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_decode_rgb_u8 (nvjpeg start-up)");
    int n = 0, widths[NVJPEG_MAX_COMPONENT] = {0}, heights[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t css;
    s = nvjpegGetImageInfo(c.handle, jpeg, jpeg_bytes, &n, &css, widths, heights);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_decode_rgb_u8");
    if (widths[0] != width || heights[0] != height)
        return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_decode_rgb_u8: destination size differs from the image's");
    nvjpegImage_t img;
    std::memset(&img, 0, sizeof(img));
    img.channel[0] = dst;
    img.pitch[0] = (size_t)width * 3;
    if (decoder_mode() != 0) {
        guard.unlock();  // the state below belongs to this call alone
        struct Lease {
            Codec& c;
            ThreadDecoder* d = nullptr;
            explicit Lease(Codec& codec_) : c(codec_) {
                std::lock_guard<std::mutex> g(c.pool_lock);
                if (!c.idle.empty()) {
                    d = c.idle.back();
                    c.idle.pop_back();
                }
                if (!d) d = new ThreadDecoder();
            }
            ~Lease() {
                std::lock_guard<std::mutex> g(c.pool_lock);
                c.idle.push_back(d);
            }
        } lease(c);
        ThreadDecoder& d = *lease.d;
        s = d.ensure(c.handle, decoder_mode() == 1 ? NVJPEG_BACKEND_GPU_HYBRID : NVJPEG_BACKEND_HYBRID);
        if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_decode_rgb_u8 (decoupled decoder)");
        if (cudaEventSynchronize(d.done) != cudaSuccess) return fail(PB_IO_ERR_CUDA, "pb_io_jpeg_decode_rgb_u8: event");
        s = nvjpegJpegStreamParse(c.handle, jpeg, jpeg_bytes, 0, 0, d.stream);
        int unsupported = 1;
        if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegDecoderJpegSupported(d.decoder, d.stream, d.params, &unsupported);
        if (s == NVJPEG_STATUS_SUCCESS && unsupported == 0) {
            cudaStream_t st = (cudaStream_t)stream;
            s = nvjpegDecodeJpegHost(c.handle, d.decoder, d.state, d.params, d.stream);
            if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegDecodeJpegTransferToDevice(c.handle, d.decoder, d.state, d.stream, st);
            if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegDecodeJpegDevice(c.handle, d.decoder, d.state, &img, st);
            if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_decode_rgb_u8 (decoupled decoder)");
            // wait here, asleep, for the device stages: the caller's own synchronize then returns at once
            static const bool spin = std::getenv("PB_IO_SPIN") && std::atoi(std::getenv("PB_IO_SPIN")) != 0;  // A/B runs
            if (cudaEventRecord(d.done, st) != cudaSuccess || (!spin && cudaEventSynchronize(d.done) != cudaSuccess))
                return fail(PB_IO_ERR_CUDA, "pb_io_jpeg_decode_rgb_u8: waiting for the decode");
            return PB_IO_OK;
        }
        guard.lock();  // (progressive, 4-component ... images: the single-state decoder takes them)
    }
    g_single_state_decodes.fetch_add(1, std::memory_order_relaxed);
    s = nvjpegDecode(c.handle, c.dec_state, jpeg, jpeg_bytes, NVJPEG_OUTPUT_RGBI, &img, (cudaStream_t)stream);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_decode_rgb_u8");
    // the state's device buffers are in use until the stream has run the decode: another host thread
    // (another stream) must not start the next image on this state before that
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
        return fail(PB_IO_ERR_CUDA, "pb_io_jpeg_decode_rgb_u8: stream synchronize");
    return PB_IO_OK;
}

int pb_io_jpeg_encode_rgb_u8(const uint8_t* src, int32_t width, int32_t height, int32_t quality, int32_t subsampling,
                             void* stream, uint8_t* out, size_t* out_bytes) {
    if (!src || !out_bytes) return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_encode_rgb_u8: null pointer");
    if (width < 1 || height < 1 || width > 65535 || height > 65535)
        return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_encode_rgb_u8: a JPEG is 1..65535 pixels on a side");
    if (quality < 1 || quality > 100) return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_encode_rgb_u8: quality must be 1..100");
    nvjpegChromaSubsampling_t css;
    switch (subsampling) {
        case PB_IO_CSS_444: css = NVJPEG_CSS_444; break;
        case PB_IO_CSS_422: css = NVJPEG_CSS_422; break;
        case PB_IO_CSS_420: css = NVJPEG_CSS_420; break;
        default: return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_encode_rgb_u8: unknown subsampling");
    }
    cudaStream_t st = (cudaStream_t)stream;
    Codec& c = codec();
    nvjpegStatus_t s;
    {
        std::lock_guard<std::mutex> guard(c.lock);
        s = c.ensure();
    }
    std::lock_guard<std::mutex> enc_guard(c.enc_lock);
    if (s == NVJPEG_STATUS_SUCCESS) s = c.ensure_encoder(st);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_encode_rgb_u8 (nvjpeg start-up)");
    s = nvjpegEncoderParamsSetQuality(c.enc_params, quality, st);
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncoderParamsSetSamplingFactors(c.enc_params, css, st);
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncoderParamsSetOptimizedHuffman(c.enc_params, 0, st);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_encode_rgb_u8 (parameters)");
    nvjpegImage_t img;
    std::memset(&img, 0, sizeof(img));
    img.channel[0] = const_cast<uint8_t*>(src);
    img.pitch[0] = (size_t)width * 3;
    s = nvjpegEncodeImage(c.handle, c.enc_state, c.enc_params, &img, NVJPEG_INPUT_RGBI, width, height, st);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_encode_rgb_u8");
    cudaError_t e = cudaStreamSynchronize(st);  // the bitstream is assembled on the host from device results
    if (e != cudaSuccess) return fail(PB_IO_ERR_CUDA, std::string("pb_io_jpeg_encode_rgb_u8: ") + cudaGetErrorString(e));
    size_t need = 0;
    s = nvjpegEncodeRetrieveBitstream(c.handle, c.enc_state, nullptr, &need, st);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_encode_rgb_u8 (size)");
    if (!out || *out_bytes < need) {
        *out_bytes = need;
        return fail(PB_IO_ERR_INVALID_ARGUMENT, "pb_io_jpeg_encode_rgb_u8: output buffer too small");
    }
    s = nvjpegEncodeRetrieveBitstream(c.handle, c.enc_state, out, &need, st);
    if (s != NVJPEG_STATUS_SUCCESS) return codec_fail(s, "pb_io_jpeg_encode_rgb_u8 (bitstream)");
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(PB_IO_ERR_CUDA, std::string("pb_io_jpeg_encode_rgb_u8: ") + cudaGetErrorString(e));
    *out_bytes = need;
    return PB_IO_OK;
}

}  // extern "C"
