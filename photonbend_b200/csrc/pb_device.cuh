// pb_device.cuh -- per-pixel float64 device math of the remap path (sm_100a).
//
// One output pixel goes through: output ray (a1/a3/a4 of SURVEY.md section 8) -> 0..n rotations
// (a7) -> source lookup (a9/a10/a11).  Everything here is float64 without FMA contraction
// (compiled with -fmad=false; the explicit __dmul_rn/__dadd_rn calls document where the
// reference rounds a product before it adds), because the truncated source index has to match
// the reference's NumPy float64 arithmetic bit for bit.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "pb_remap.h"

namespace pb {

constexpr double kPi = 3.141592653589793;  // == numpy.pi

// ---------------------------------------------------------------------------------- parameters

// Output geometry with the constants the reference derives from it, computed once on the host
// in the reference's own order of operations (see derive_out in pb_remap.cu).
struct OutGeom {
    int kind, lens;
    int H, W;      // W = effective output width (2*(width/2) for a double image)
    int half_w;    // double: width of one half
    double f;      // pixels per focal unit
    double half_fov;      // fov / 2: rays beyond it are invalid            projection.py:160, 357
    double right_lat_min; // double, right half: pi - fov/2                 projection.py:358-360
    // pixel-grid axes as numpy.linspace(start, stop, n): value(i) = fl(fl(i*step) + start), last == stop
    double x_start, x_stop, x_step;  // camera/double: x of a column        projection.py:177, 390-392
    double y_start, y_stop, y_step;  // camera/double: y of a row           projection.py:178-180, 399-401
                                     // equirect: x = longitude, y = latitude projection.py:502-505
};

struct SrcGeom {
    int kind, lens;
    int H, W, C;
    int wl, wr;           // double: widths of the left / right halves      projection.py:413, 429-431
    double f;
    double cy, cx;        // camera image centre (H/2-.5, W/2-.5)            projection.py:262-274
    double cxl, cxr;      // double: centres of the two halves
    double rect_limit;    // to_radians(89), rectilinear forward domain      lens.py:97-98
    double seg_h, seg_w, half_w;  // equirect: pi/H, pi/(W/2), W/2           projection.py:538-543
    double mrg_lo, mrg_hi, mrg_hi_safe, mrg_span;  // double blend band      projection.py:414-418
};

struct Rotations {
    int n;
    double m[PB_MAX_ROTATIONS][9];
};

// ---------------------------------------------------------------------------------- helpers

__device__ __forceinline__ double linspace_at(double start, double stop, double step, int n, int i) {
    if (i == n - 1 && n > 1) return stop;
    return __dadd_rn(__dmul_rn((double)i, step), start);
}

// _shared.py:25-55 + np.log(z).imag: both parts are broadcast as x + y*0, y + x*0 first (this
// turns a lone -0.0 into +0.0 and spreads NaN), then the angle is atan2.
__device__ __forceinline__ double angle_of(double x, double y) {
    double fx = __dadd_rn(x, __dmul_rn(y, 0.0));
    double fy = __dadd_rn(y, __dmul_rn(x, 0.0));
    return atan2(fy, fx);
}

// ndarray.astype(int) on x86 (cvttsd2si): truncation toward zero; NaN and out-of-range give
// INT64_MIN.
__device__ __forceinline__ long long trunc_i64(double v) {
    if (!(fabs(v) < 9223372036854775808.0)) return (long long)0x8000000000000000ULL;
    return __double2ll_rz(v);
}

// numpy's % on int64 (sign of the divisor)
__device__ __forceinline__ long long floor_mod(long long a, long long m) {
    long long r = a % m;
    if (r != 0 && ((r < 0) != (m < 0))) r += m;
    return r;
}

// ---------------------------------------------------------------------------------- lenses

// lens.py:75-103, 126-144, 168-187, 224-243, 266-286, 313-335 (array branch)
__device__ __forceinline__ double lens_forward(int lens, double theta, double rect_limit) {
    switch (lens) {
        case PB_LENS_EQUIDISTANT: return theta;
        case PB_LENS_EQUISOLID: return 2.0 * sin(theta / 2.0);
        case PB_LENS_ORTHOGRAPHIC: return sin(theta);
        case PB_LENS_STEREOGRAPHIC: return 2.0 * tan(theta / 2.0);
        case PB_LENS_RECTILINEAR:
            if (theta < 0.0 || theta > rect_limit) return __longlong_as_double(0x7ff8000000000000LL);
            return tan(theta);
        default: return __dmul_rn(1.47, sin(__dmul_rn(0.713, theta)));
    }
}

// lens.py:68-72, 106-124, 147-165, 190-220, 246-262, 289-309
__device__ __forceinline__ double lens_inverse(int lens, double d) {
    switch (lens) {
        case PB_LENS_EQUIDISTANT: return d;
        case PB_LENS_EQUISOLID: {
            double t = 2.0 * asin(d / 2.0);
            return (t != t) ? 0.0 : t;  // out-of-domain radius -> latitude 0 (and valid)
        }
        case PB_LENS_ORTHOGRAPHIC: return asin(d);
        case PB_LENS_STEREOGRAPHIC: return 2.0 * atan(d / 2.0);
        case PB_LENS_RECTILINEAR: return atan(d);
        default: return __ddiv_rn(asin(__ddiv_rn(d, 1.47)), 0.713);
    }
}

// ---------------------------------------------------------------------------------- rays

struct Ray {
    double lat, lon;
    bool invalid;
};

// a1 projection.py:487-513, a3 projection.py:147-194, a4 projection.py:341-406
template <int OUT_KIND>
__device__ __forceinline__ Ray output_ray(const OutGeom& g, int i, int j) {
    Ray r;
    if (OUT_KIND == PB_KIND_EQUIRECT) {
        r.lon = linspace_at(g.x_start, g.x_stop, g.x_step, g.W, j);
        r.lat = linspace_at(g.y_start, g.y_stop, g.y_step, g.H, i);
        r.invalid = false;
    } else {
        const bool right = (OUT_KIND == PB_KIND_DOUBLE) && j >= g.half_w;
        const int n_cols = (OUT_KIND == PB_KIND_DOUBLE) ? g.half_w : g.W;
        double x = linspace_at(g.x_start, g.x_stop, g.x_step, n_cols, right ? j - g.half_w : j);
        if (right) x = __dmul_rn(x, -1.0);
        const double y = linspace_at(g.y_start, g.y_stop, g.y_step, g.H, i);
        const double d = __ddiv_rn(sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y))), g.f);
        double lat = lens_inverse(g.lens, d);
        if (right) {
            lat = __dadd_rn(__dmul_rn(lat, -1.0), kPi);
            r.invalid = lat < g.right_lat_min;
        } else {
            r.invalid = lat > g.half_fov;
        }
        r.lat = lat;
        r.lon = angle_of(x, y);
    }
    return r;
}

// a7 rotation.py:102-176 for one ray; m is row-major 3x3.
__device__ __forceinline__ Ray rotate_ray(Ray in, const double* __restrict__ m) {
    Ray out;
    out.invalid = in.invalid;
    if (in.invalid) {
        out.lat = 0.0;
        out.lon = 0.0;
        return out;
    }
    double sl, cl, so, co;
    sincos(in.lat, &sl, &cl);
    sincos(in.lon, &so, &co);
    const double vx = __dmul_rn(co, sl);
    const double vy = cl;
    const double vz = __dmul_rn(so, sl);
    const double nx = __dadd_rn(__dadd_rn(__dmul_rn(m[0], vx), __dmul_rn(m[1], vy)), __dmul_rn(m[2], vz));
    const double ny = __dadd_rn(__dadd_rn(__dmul_rn(m[3], vx), __dmul_rn(m[4], vy)), __dmul_rn(m[5], vz));
    const double nz = __dadd_rn(__dadd_rn(__dmul_rn(m[6], vx), __dmul_rn(m[7], vy)), __dmul_rn(m[8], vz));
    out.lat = acos(ny);
    out.lon = angle_of(nx, nz);
    return out;
}

// ---------------------------------------------------------------------------------- source lookup

// What one output pixel reads: up to two source pixels (linear pixel offsets, -1 = black) and
// their float64 weights.  Camera / equirect sources only use slot 0 with weight 1.
struct Lookup {
    int off0, off1;
    double w0, w1;
    bool blend;  // double source: out = wrap_u8(p0*w0 + p1*w1)
};

// a9 projection.py:247-274 + 223-231: truncation toward zero BEFORE the bounds test, so a
// coordinate in (-1, 0) lands on index 0 and is valid; NaN/inf are "problem positions".
__device__ __forceinline__ int camera_offset(int lens, double f, double rect_limit, int h, int w,
                                             double cy, double cx, double lat, double lon,
                                             int row_pitch_px, int col0, bool flip) {
    const double dist = __dmul_rn(lens_forward(lens, lat, rect_limit), f);
    double s, c;
    sincos(lon, &s, &c);
    const double fx = __dadd_rn(__dmul_rn(c, dist), cx);
    const double fy = __dadd_rn(__dmul_rn(__dmul_rn(s, dist), -1.0), cy);
    if (!(fabs(fx) < 2147483648.0) || !(fabs(fy) < 2147483648.0)) return -1;
    const int px = __double2int_rz(fx);
    const int py = __double2int_rz(fy);
    if (px < 0 || px >= w || py < 0 || py >= h) return -1;
    return py * row_pitch_px + col0 + (flip ? (w - 1 - px) : px);
}

// a10 projection.py:439-456
__device__ __forceinline__ double merge_weight(const SrcGeom& s, double lat) {
    if (lat >= s.mrg_lo && lat <= s.mrg_hi_safe) return __dmul_rn(__ddiv_rn(__dadd_rn(lat, -s.mrg_hi), s.mrg_span), -1.0);
    return 1.0;
}

template <int SRC_KIND>
__device__ __forceinline__ Lookup source_lookup(const SrcGeom& s, Ray r) {
    Lookup L;
    L.off0 = L.off1 = -1;
    L.w0 = L.w1 = 1.0;
    L.blend = false;
    if (r.invalid) return L;
    if (SRC_KIND == PB_KIND_CAMERA) {
        L.off0 = camera_offset(s.lens, s.f, s.rect_limit, s.H, s.W, s.cy, s.cx, r.lat, r.lon, s.W, 0, false);
    } else if (SRC_KIND == PB_KIND_EQUIRECT) {
        // a11 projection.py:515-547: true division, truncation, Python-sign modulo
        const double frow = __ddiv_rn(r.lat, s.seg_h);
        const double fcol = __dadd_rn(__ddiv_rn(r.lon, s.seg_w), s.half_w);
        int row, col;
        if (frow >= 0.0 && frow < 2147483648.0) {
            row = __double2int_rz(frow);
            if (row >= s.H) row = (row < 2 * s.H) ? row - s.H : row % s.H;
        } else {
            row = (int)floor_mod(trunc_i64(frow), (long long)s.H);
        }
        if (fcol >= 0.0 && fcol < 2147483648.0) {
            col = __double2int_rz(fcol);
            if (col >= s.W) col = (col < 2 * s.W) ? col - s.W : col % s.W;
        } else {
            col = (int)floor_mod(trunc_i64(fcol), (long long)s.W);
        }
        L.off0 = row * s.W + col;
    } else {
        // a10 projection.py:408-462: both halves are sampled as plain cameras of magnitude H/2
        const double lat_l = r.lat;
        const double lat_r = __dadd_rn(__dmul_rn(r.lat, -1.0), kPi);
        L.off0 = camera_offset(s.lens, s.f, s.rect_limit, s.H, s.wl, s.cy, s.cxl, lat_l, r.lon, s.W, 0, false);
        L.off1 = camera_offset(s.lens, s.f, s.rect_limit, s.H, s.wr, s.cy, s.cxr, lat_r, r.lon, s.W, s.wl, true);
        L.w0 = merge_weight(s, lat_l);
        L.w1 = merge_weight(s, lat_r);
        L.blend = true;
    }
    return L;
}

// (left*wl + right*wr).astype(np.uint8): truncate, keep the low byte  (projection.py:459)
__device__ __forceinline__ unsigned char blend_u8(unsigned a, double wa, unsigned b, double wb) {
    const double v = __dadd_rn(__dmul_rn((double)a, wa), __dmul_rn((double)b, wb));
    return (unsigned char)(trunc_i64(v) & 0xFF);
}

}  // namespace pb
