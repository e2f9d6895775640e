/*
 * pb_io.h -- C ABI of libpbio.so: JPEG <-> device uint8 HWC tensors (nvJPEG), the image I/O
 * plumbing either side of the remap path.
 *
 * The reference decodes and encodes on the host with Pillow
 * (photonbend/scripts/commands/__init__.py:135-143 `_open_image`: PIL.Image.open -> np.asarray;
 * make_pano.py:132-139, alter_photo.py:155-162, make_photo.py:134-141: Image.fromarray(out).save(path)).
 * These entry points are what a binding of those two spots calls when the pixels should never
 * visit host memory: the decoded image lands in device memory in the layout pb_remap_u8 reads
 * (uint8, HWC, RGB, tightly packed) and the remapped image is encoded from device memory.
 *
 * nvJPEG is a library codec, not bit-identical to libjpeg-turbo (IDCT rounding, chroma
 * upsampling): decoded pixels differ from Pillow's by a few LSB.  The Python host side therefore
 * keeps Pillow as the default and takes this path only when asked to (PHOTONBEND_B200_CODEC=nvjpeg).
 *
 * Conventions as in pb_remap.h: device pointers on the CURRENT device, caller owns every buffer,
 * work is enqueued on `stream`, int return codes (PB_IO_OK = 0) + pb_io_last_error(), nothing
 * throws across the ABI.  The library keeps one nvJPEG handle per device (created on first use with that device current).
 */
#ifndef PB_IO_H
#define PB_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PB_IO_OK = 0, PB_IO_ERR_INVALID_ARGUMENT = 1, PB_IO_ERR_UNSUPPORTED = 2, PB_IO_ERR_CODEC = 3, PB_IO_ERR_CUDA = 4 };

/* chroma subsampling of an encoded image */
enum { PB_IO_CSS_444 = 0, PB_IO_CSS_422 = 1, PB_IO_CSS_420 = 2 };

int pb_io_version(void);

/* Diagnostics: how many decodes of this process took nvjpegDecode on the device's single decoder
 * state (serialised; PB_IO_DECODER=single, or an image the decoupled decoder refused). */
long long pb_io_single_state_decodes(void);
const char *pb_io_last_error(void);

/* Size and component count (1 = grey, 3 = colour) of a JPEG held in host memory. */
int pb_io_jpeg_info(const uint8_t *jpeg, size_t jpeg_bytes, int32_t *width, int32_t *height,
                    int32_t *components);

/* Decode a host JPEG into dst (device, height x width x 3, RGB interleaved, tightly packed).
 * Grey images are expanded to RGB.  Runs on `stream` and returns when the image is there (the
 * calling thread sleeps meanwhile: several decode threads per GPU do not cost a core each).  Callable from several host threads at once on one device: every thread has
 * its own decoder state and buffers (nvJPEG's decoupled decoder with the Huffman stage on the
 * device; PB_IO_DECODER=single|threads selects the others, see pb_io.cpp). */
int pb_io_jpeg_decode_rgb_u8(const uint8_t *jpeg, size_t jpeg_bytes, uint8_t *dst, int32_t width,
                             int32_t height, void *stream);

/* Encode src (device, height x width x 3, RGB interleaved) as a baseline JPEG into out (host,
 * capacity *out_bytes); on return *out_bytes is the size of the bitstream.  quality 1..100,
 * subsampling PB_IO_CSS_* (Pillow's defaults, which the reference's save() uses, are 75 and 4:2:0).
 * Synchronises `stream` (the bitstream has to reach the host).  If the capacity is too small the
 * call fails with PB_IO_ERR_INVALID_ARGUMENT and *out_bytes holds the size needed. */
int pb_io_jpeg_encode_rgb_u8(const uint8_t *src, int32_t width, int32_t height, int32_t quality,
                             int32_t subsampling, void *stream, uint8_t *out, size_t *out_bytes);

#ifdef __cplusplus
}
#endif

#endif /* PB_IO_H */
