/*
 * pb_remap.h -- C ABI of libpbremap.so, the B200 (sm_100a) implementation of photonbend's
 * per-pixel remap path.
 *
 * The reference (wmoreirae/photonbend) is pure Python + NumPy and has no FFI of its own; the
 * boundary these entry points replace is its ProjectionImage protocol
 * (photonbend/core/__init__.py:51-92, photonbend/core/projection.py:40-66):
 *
 *     map = dst.get_coordinate_map()                      projection.py:147, 341, 487
 *     map = Rotation(p, y, r).rotate_coordinate_map(map)  rotation.py:102-176   (0..n times)
 *     out = src.process_coordinate_map(map)               projection.py:197, 408, 515
 *
 * pb_remap_u8 fuses the three calls into one kernel launch; the other three entry points are
 * the same protocol with the float64 coordinate map materialised in device memory, for
 * callers that inspect or edit the map between the calls.
 *
 * Conventions
 *  - every pointer marked "device" is a CUDA device pointer on the CURRENT device; the caller
 *    allocates and owns every buffer; the library allocates nothing that outlives a call, except
 *    the small tables owned by a pb_plan (see below);
 *  - source images may have at most 32767 rows and 65535 columns;
 *  - images are uint8, HWC, tightly packed (row pitch = width * channels);
 *  - coordinate maps are float64 (H, W, 3) = (latitude, longitude, invalid != 0), tightly packed
 *    (photonbend/core/__init__.py:42-49);
 *  - all work is enqueued on `stream` (a cudaStream_t, NULL = the default stream) and the call
 *    returns without synchronising;
 *  - every function returns PB_OK (0) or a PB_ERR_* code; pb_last_error() gives the message of
 *    the last failure on the calling thread.  Nothing throws or aborts across the ABI;
 *  - host-side parameter derivation stays with the caller exactly as the reference does it:
 *    f_distance = magnitude / lens_forward(fov / 2) (projection.py:123-144, 318-339) and the
 *    row-major 3x3 matrices Rotation.rotation_matrix (rotation.py:27-62, 100), so that these
 *    constants are bit-identical to the reference's.
 */
#ifndef PB_REMAP_H
#define PB_REMAP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 2

/* image formats: projection.py CameraImage :69, DoubleCameraImage :277, PanoramaImage :465 */
enum { PB_KIND_CAMERA = 0, PB_KIND_DOUBLE = 1, PB_KIND_EQUIRECT = 2 };

/* lens models: lens.py:341-401 */
enum {
    PB_LENS_EQUIDISTANT = 0,
    PB_LENS_EQUISOLID = 1,
    PB_LENS_ORTHOGRAPHIC = 2,
    PB_LENS_STEREOGRAPHIC = 3,
    PB_LENS_RECTILINEAR = 4,
    PB_LENS_THOBY = 5,
    /* a user-defined lens (lens.py:48-64 lets Lens wrap any pair of callables): the function the
     * kernel needs is handed over as a table of samples, see pb_image_desc.lens_table */
    PB_LENS_TABLE = 6
};

enum {
    PB_OK = 0,
    PB_ERR_INVALID_ARGUMENT = 1,
    PB_ERR_UNSUPPORTED = 2,
    PB_ERR_CUDA = 3,
    PB_ERR_TOO_MANY_ROTATIONS = 4
};

#define PB_MAX_ROTATIONS 16

/* One side (output or source) of a remap. */
typedef struct pb_image_desc {
    int32_t kind;      /* PB_KIND_* */
    int32_t lens;      /* PB_LENS_*; ignored for PB_KIND_EQUIRECT */
    int32_t height;    /* pixels */
    int32_t width;     /* pixels (a double image holds two width/2 halves side by side) */
    double fov;        /* radians: full field of view (camera) or per-sensor fov (double) */
    double f_distance; /* pixels per focal unit = magnitude / lens_forward(fov / 2) */
    /* PB_LENS_TABLE only (otherwise NULL / 0): lens_table_n >= 2 float64 samples in HOST memory of
     * the one lens function this side of the remap evaluates, at x_k = k * lens_table_max /
     * (lens_table_n - 1): as `out`, the lens's reverse function (radius in focal units ->
     * latitude); as `src`, its forward function (latitude -> radius in focal units).  The kernel
     * interpolates linearly and clamps x to [0, lens_table_max].  Copied during the call (plans
     * keep a device copy): the caller may free it afterwards.  Results follow the user's callables
     * to the interpolation error, not bit for bit -- the only lens kind for which that holds. */
    const double *lens_table;
    int32_t lens_table_n;
    int32_t reserved_;
    double lens_table_max;
} pb_image_desc;

/* A whole remap: output geometry, rotations applied in order, source geometry. */
typedef struct pb_remap_desc {
    pb_image_desc out;
    pb_image_desc src;
    int32_t channels;    /* 1..4 interleaved uint8 channels */
    int32_t n_rotations; /* 0..PB_MAX_ROTATIONS */
    double rotations[PB_MAX_ROTATIONS][9]; /* row-major Rotation.rotation_matrix, in order */
} pb_remap_desc;

int pb_version(void);
const char *pb_last_error(void);

/* Kernels this library has launched in this process so far (every grid counts once; a batch
 * through a double-fisheye source is two grids, one per tile class).  For benchmarks that have to
 * state how many of their own kernels ran inside a timed region. */
int64_t pb_kernel_launches(void);

/* Width of the image a geometry produces: 2*(width/2) for PB_KIND_DOUBLE (projection.py:389-397),
 * width otherwise. */
int32_t pb_output_width(const pb_image_desc *out);

/*
 * Fused remap of n_frames frames that share one geometry:
 *   dst[k] = src_format.process_coordinate_map(rotate*(out_format.get_coordinate_map()))  on src[k]
 * src: device, n_frames images of src.height x src.width x channels, frame k at src + k*src_frame_stride
 * dst: device, n_frames images of out.height x pb_output_width(out) x channels, likewise.
 */
int pb_remap_u8(const pb_remap_desc *desc, const uint8_t *src, int64_t src_frame_stride,
                uint8_t *dst, int64_t dst_frame_stride, int32_t n_frames, void *stream);

/*
 * Plans.  A plan is one validated geometry together with everything derived from it: the host
 * constants and, for an un-rotated equirect output (the video case, BASELINE config 5), small
 * separable device tables (cos/sin of each column's longitude, lens radius of each row's
 * latitude; 16*W + 32*H bytes) that the kernel would otherwise rebuild per call.  pb_remap_u8 is
 * pb_plan_create + pb_plan_remap_u8 + pb_plan_destroy with the tables in a transient
 * stream-ordered allocation.  The tables are the only device memory the library ever owns; they
 * live on the device that was current at pb_plan_create and die with pb_plan_destroy.
 * pb_plan_create enqueues the table kernel on `stream`; use the plan on that stream or after it.
 */
typedef struct pb_plan pb_plan;
int pb_plan_create(const pb_remap_desc *desc, void *stream, pb_plan **plan);
int pb_plan_remap_u8(const pb_plan *plan, const uint8_t *src, int64_t src_frame_stride, uint8_t *dst,
                     int64_t dst_frame_stride, int32_t n_frames, void *stream);
/*
 * Output rows [row_begin, row_end) of ONE frame: dst_band (device) holds (row_end - row_begin) x
 * pb_output_width(out) x channels bytes, i.e. the band alone.  This is how a single large frame is
 * sharded over the GPUs of a box by output-row bands: every GPU holds the whole source and
 * produces its own band, no exchange (the reference's protocol works on any row slice of the
 * coordinate map the same way: photonbend/core/projection.py:197-245 takes any (h, w, 3) map).
 * Bands that start on a multiple of 64 rows take the TMA-staged kernels.
 */
int pb_plan_remap_rows_u8(const pb_plan *plan, const uint8_t *src, uint8_t *dst_band,
                          int32_t row_begin, int32_t row_end, void *stream);
void pb_plan_destroy(pb_plan *plan);

/*
 * get_coordinate_map() followed by desc->n_rotations rotate_coordinate_map() calls, written to
 * map (device, float64 out.height x pb_output_width(out) x 3).  desc->src is ignored.
 */
int pb_materialize_map_f64(const pb_remap_desc *desc, double *map, void *stream);

/*
 * Rotation.rotate_coordinate_map on an explicit map of n_pixels entries (rotation.py:102-176):
 * map_out = rotated map; like the reference, the invalid entries of map_in are zeroed in place.
 * map_in and map_out are device pointers and may not alias.
 */
int pb_rotate_map_f64(const double matrix[9], double *map_in, double *map_out, int64_t n_pixels,
                      void *stream);

/*
 * process_coordinate_map on an explicit map (device, float64 map_height x map_width x 3):
 * dst (device, map_height x map_width x channels) = src image sampled through the map.
 * Like the reference's PanoramaImage.process_coordinate_map (projection.py:533-536), an
 * equirect source zeroes the invalid entries of map in place.
 */
int pb_gather_from_map_u8(const pb_image_desc *src_desc, int32_t channels, double *map,
                          int32_t map_height, int32_t map_width, const uint8_t *src, uint8_t *dst,
                          void *stream);

/*
 * map_projection (projection.py:550-599) on an explicit map (device, float64 map_height x
 * map_width x 3): dst (device, map_height x map_width x 3 uint8) = (latitude stretched over the
 * range it takes on the valid pixels, longitude * 255 / 2 pi, 255 where invalid), each rounded half
 * to even and cast like numpy's astype(uint8).  Like the reference it zeroes (lat, lon) of the
 * invalid entries of map in place.  A map without a valid pixel makes the reference raise
 * (numpy.min of an empty array); callers check that themselves -- here the red channel is then 0.
 */
int pb_map_projection_u8(double *map, int32_t map_height, int32_t map_width, uint8_t *dst, void *stream);

/*
 * Diagnostics of the FP32-first tier of the per-pixel chain (csrc/pb_fast32.cuh), over every output
 * pixel of a geometry, synchronous: stats[0..1] = largest |float - double| / (2^-24 * error shape)
 * of a source coordinate along x / y (the calibration of the tier's error bound K), stats[2] =
 * pixels, stats[3] = pixels the tier left undecided (they take the float64 tiers), stats[4] = pixels
 * it decided DIFFERENTLY from the float64 tiers (must be 0), stats[5] = pixels whose fov / no-pixel
 * status float and double evaluations disagree on (must be 0).  No reference counterpart: the
 * reference computes in float64 throughout (projection.py:37,46,57).
 */
int pb_debug_fast32_stats(const pb_remap_desc *desc, double stats[6], void *stream);

/*
 * The FP32-first tier as calibrated for one plan: pb_plan_create measures the largest
 * |float - double| / (2^-24 * error shape) over the plan's own output pixels (out[1]; -1 when the
 * plan resolves through tables and needs no tier) and bounds the tier's error with
 * K = max(1.5, 1.25 * that + 0.25) (out[0]; 0 when the tier is off) instead of the K = 16 that holds for
 * every geometry.  No reference counterpart.
 */
int pb_debug_plan_fast32(const pb_plan *plan, double out[2]);

#ifdef __cplusplus
}
#endif
#endif /* PB_REMAP_H */
