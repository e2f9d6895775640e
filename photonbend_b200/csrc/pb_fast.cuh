// pb_fast.cuh -- guarded short cut through the per-pixel float64 chain (sm_100a).
//
// The exact chain (pb_device.cuh: output_ray -> rotate_ray* -> source_lookup) follows the
// reference call by call: every rotation goes latitude/longitude -> unit vector -> matrix ->
// acos/atan2 -> latitude/longitude again (rotation.py:129-164), and every projection goes through
// its angle (projection.py:189-193, 251-252).  That is seven libm calls per pixel for a rotated
// camera -> camera remap and ~760 instructions; the FP64 pipe, not HBM, bounds such a launch.
//
// The short cut keeps the ray as a unit vector from the output pixel to the source projection:
//   * lens inverses that are algebraic in the radius give (sin lat, cos lat) without asin/atan
//     (equisolid, orthographic, stereographic, rectilinear), and cos/sin of the pixel's longitude
//     are x/r, y/r;
//   * rotations are plain matrix products of the vector;
//   * lens forwards that are algebraic in cos(lat) need no angle either; an equidistant source
//     needs one acos, a panorama source one acos + one atan2.
// It evaluates the same real function as the exact chain, but rounds differently (~1e-15
// relative), so its truncated source index can differ where a coordinate sits within ~1e-9 px of
// an integer.  Therefore every decision the short cut takes -- each truncation, the fov test, the
// lens domains, the blend band of a double source -- is only accepted when the value is further
// than a guard band from the decision boundary (2^-19 px, 1e-9 relative for angles; the two
// chains differ by < 1e-8 px wherever the conditioning guards below let the short cut run).
// Otherwise the pixel is "undecided" and the caller runs the exact chain for it (~4e-6 of the
// pixels).  Results are therefore those of the exact chain, bit for bit
// (tests/test_gpu_parity.py::test_fast_path_equals_exact_chain).
// Caveats: (1) the matrices must be orthonormal (derive_fast checks, else the exact chain runs);
// (2) the conditioning guards look at the FINAL vector only.  With two or more rotations an
// INTERMEDIATE vector of the exact chain that lands within ~1e-7 rad of a pole has its acos error
// amplified past the guard band; the short cut (which never forms that intermediate angle) then
// differs from the exact chain.  That needs a pixel whose intermediate ray hits a cone of 1e-7 rad:
// probability ~1e-14 per pixel and rotation, i.e. not observable, but not zero.
#pragma once

#include "pb_device.cuh"

namespace pb {

// T = float: what the kernels use; T = double: the same constants unrounded, for the calibration
// reference (so that the rounding of the constants themselves is part of the measured error)
template <typename T>
struct Fast32GeomT {
    int enabled;
    int has_rot;
    T rot[9];
    // output side, camera / double: x = col + x0 (col counted within the half), y = y0 - row
    T x0, y0;
    T inv_f, quarter_inv_f2;  // 1 / f, 0.25 / f^2
    T r2_valid, r2_invalid, r2_domain;
    T r2_nan;  // r^2 >= r2_nan: beyond the lens inverse's domain, the reference's ray is NaN (not flagged invalid)
    // output side, equirect: lon = col * lon_step + lon0, lat = row * lat_step
    T lon0, lon_step, lat_step;
    // source side
    T src_f;
    T inv_seg_h, inv_seg_w;    // equirect source: rows / columns per radian
    T ny_rect_in, ny_rect_out;
    T ny_band_lo, ny_band_hi;
    T k_eps;                   // K * 2^-24
};
using Fast32Geom = Fast32GeomT<float>;

inline Fast32Geom fast32_to_float(const Fast32GeomT<double>& d) {
    Fast32Geom f;
    f.enabled = d.enabled;
    f.has_rot = d.has_rot;
    for (int e = 0; e < 9; ++e) f.rot[e] = (float)d.rot[e];
    f.x0 = (float)d.x0;
    f.y0 = (float)d.y0;
    f.inv_f = (float)d.inv_f;
    f.quarter_inv_f2 = (float)d.quarter_inv_f2;
    f.r2_valid = (float)d.r2_valid;
    f.r2_invalid = (float)d.r2_invalid;
    f.r2_domain = (float)d.r2_domain;
    f.r2_nan = (float)d.r2_nan;
    f.lon0 = (float)d.lon0;
    f.lon_step = (float)d.lon_step;
    f.lat_step = (float)d.lat_step;
    f.src_f = (float)d.src_f;
    f.inv_seg_h = (float)d.inv_seg_h;
    f.inv_seg_w = (float)d.inv_seg_w;
    f.ny_rect_in = (float)d.ny_rect_in;
    f.ny_rect_out = (float)d.ny_rect_out;
    f.ny_band_lo = (float)d.ny_band_lo;
    f.ny_band_hi = (float)d.ny_band_hi;
    f.k_eps = (float)d.k_eps;
    return f;
}

struct FastGeom {
    int enabled;
    int n_rot;
    int has_rot;
    double rot[9];                 // all rotations composed, row-major: R_n ... R_1
    // output side (camera / double)
    double inv_f;                  // 1 / out.f
    double r2_valid, r2_invalid;   // r^2 < r2_valid: inside the fov; r^2 > r2_invalid: outside; else undecided
    double r2_domain;              // r^2 >= r2_domain: lens inverse near the edge of its domain -> undecided
    // source side
    double src_f;
    double inv_seg_h, inv_seg_w;   // equirect source: rows / columns per radian
    double ny_rect_in, ny_rect_out;  // rectilinear source lens: cos(lat) > in: tan defined; < out: NaN (no pixel)
    double ny_band_lo, ny_band_hi;   // double source: cos(lat) in [lo, hi] may be blended -> undecided
    Fast32Geom f32;                  // the FP32-first tier in front of this one (pb_fast32.cuh)
};

constexpr double kTwo52 = 4503599627370496.0;

// atan(t) / t as a polynomial in t^2 on |t| <= tan(pi/8): Chebyshev interpolant of degree 8,
// |t * P(t^2) - atan(t)| < 1e-14 (tests/analysis/atan_fit.py).  In constant memory so that the
// coefficients are operands of the DFMAs instead of 64-bit immediates built by two UMOVs each.
__constant__ double kAtanPoly[9] = {0.9999999999999734,  -0.33333333330806103, 0.199999996052076,
                                    -0.1428569043658466, 0.11110385295384498,  -0.0907839471343116,
                                    0.07563718994323626, -0.058745523239813566, 0.030663008090113325};

// 1 / d for a normal, finite d: hardware seed (MUFU.RCP64H) + two Newton steps, ~1 ulp
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// atan2(y, x) with an absolute error < 5e-14 for finite (x, y) with x^2 + y^2 well above zero (the
// callers guarantee > 1e-8): one reduction to |t| <= tan(pi/8) by a rotation of pi/4 BEFORE the
// single division, a degree-8 polynomial, three reflections.  The short cut only needs angles to
// ~1e-10 (its decisions carry a 2^-19 px guard band); libm's correctly-rounded-ish atan2 / acos
// cost three times as many issue slots.
// acos(c) for a unit vector's component c with s = sqrt(1 - c^2) > 0 at hand: atan2(s, c) without
// the reflections of the general case (s >= 0)
__device__ __forceinline__ double fast_acos_sc(double s, double c) {
    const double ac = fabs(c);
    const double mx = fmax(ac, s), mn = fmin(ac, s);
    const bool big = mn > 0.41421356237309503 * mx;
    const double t = (big ? mn - mx : mn) * fast_rcp(big ? mn + mx : mx);
    const double u = t * t;
    double p = kAtanPoly[8];
#pragma unroll
    for (int k = 7; k >= 0; --k) p = fma(p, u, kAtanPoly[k]);
    double r = fma(t, p, big ? 0.78539816339744831 : 0.0);  // atan(mn / mx)
    if (s > ac) r = 1.5707963267948966 - r;                  // atan(s / |c|)
    return (c < 0.0) ? kPi - r : r;
}

__device__ __forceinline__ double fast_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    const bool big = mn > 0.41421356237309503 * mx;
    const double num = big ? mn - mx : mn;
    const double den = big ? mn + mx : mx;
    const double t = num * fast_rcp(den);
    const double u = t * t;
    double p = kAtanPoly[8];
#pragma unroll
    for (int k = 7; k >= 0; --k) p = fma(p, u, kAtanPoly[k]);
    double r = fma(t, p, big ? 0.78539816339744831 : 0.0);
    if (ay > ax) r = 1.5707963267948966 - r;
    if (x < 0.0) r = kPi - r;
    return (y < 0.0) ? -r : r;
}

// Index of source coordinate v along an axis of n pixels, under the reference's rule (truncate
// toward zero, then 0 <= index < n; projection.py:223-231, 254-259).
// One FP64 add: |v| + 1.5 * 2^32 lies in [1.5 * 2^32, 2^33) for |v| < 2^31, where one ulp is 2^-20,
// so the mantissa of the sum is |v| + 2^31 in fixed point with 20 fraction bits.
struct FastIndex {
    int idx;       // trunc(|v|)
    bool decided;  // v is finite, |v| < 2^31 and further than 2^-19 from every integer
    bool inside;   // 0 <= trunc(v) < n
};
__device__ __forceinline__ FastIndex fast_index(double v, int n) {
    const double g = __dadd_rn(fabs(v), 6442450944.0);
    const unsigned lo = (unsigned)__double2loint(g), hi = (unsigned)__double2hiint(g);
    FastIndex r;
    r.idx = (int)(__funnelshift_r(lo, hi, 20) ^ 0x80000000u);
    // hi in [0x41F80000, 0x42000000): |v| < 2^31, not NaN; fraction (rounded to nearest 2^-20) in [2, 2^20 - 2]
    r.decided = ((hi & 0xFFF80000u) == 0x41F80000u) & (((lo & 0xFFFFFu) - 2u) < (0x100000u - 3u));
    // (-1, 0) truncates to 0: a negative v is inside only with idx == 0
    r.inside = (unsigned)r.idx < ((__double2hiint(v) < 0) ? 1u : (unsigned)n);
    return r;
}

// Unit vector of the ray of output pixel (i, j) before any rotation:
// (cos lon sin lat, cos lat, sin lon sin lat), rotation.py:129-132.
// 0 = ray inside the fov, 1 = outside (black pixel), 2 = undecided.
template <int OUT_KIND>
__device__ __forceinline__ int fast_out_vector(const OutGeom& g, const FastGeom& fg, int i, int j, double& vx,
                                               double& vy, double& vz) {
    if (OUT_KIND == PB_KIND_EQUIRECT) {
        const double lon = linspace_at(g.x_start, g.x_stop, g.x_step, g.W, j);
        const double lat = linspace_at(g.y_start, g.y_stop, g.y_step, g.H, i);
        double sl, cl, so, co;
        sincos(lat, &sl, &cl);
        sincos(lon, &so, &co);
        vx = co * sl;
        vy = cl;
        vz = so * sl;
        return 0;
    }
    const bool right = (OUT_KIND == PB_KIND_DOUBLE) && j >= g.half_w;
    const int n_cols = (OUT_KIND == PB_KIND_DOUBLE) ? g.half_w : g.W;
    // pixel-centre coordinates: the linspace steps of projection.py:177-183 are exactly +-1
    (void)n_cols;
    double x = (double)(right ? j - g.half_w : j) + g.x_start;
    if (right) x = -x;
    const double y = g.y_start - (double)i;
    const double r2 = fma(x, x, y * y);
    if (!(r2 < fg.r2_domain)) return 2;
    if (!(r2 < fg.r2_valid)) return (r2 > fg.r2_invalid) ? 1 : 2;
    const double inv_f = fg.inv_f;
    double k;  // sin(lat) / r
    switch (g.lens) {
        case PB_LENS_EQUISOLID: {  // sin(lat/2) = d/2
            const double u2 = r2 * (0.25 * inv_f * inv_f);
            k = sqrt(1.0 - u2) * inv_f;
            vy = fma(-2.0, u2, 1.0);
            break;
        }
        case PB_LENS_ORTHOGRAPHIC: {  // sin(lat) = d
            k = inv_f;
            vy = sqrt(fma(-r2, inv_f * inv_f, 1.0));
            break;
        }
        case PB_LENS_STEREOGRAPHIC: {  // tan(lat/2) = d/2
            const double t2 = r2 * (0.25 * inv_f * inv_f);
            const double w = 1.0 / (1.0 + t2);
            k = inv_f * w;
            vy = (1.0 - t2) * w;
            break;
        }
        case PB_LENS_RECTILINEAR: {  // tan(lat) = d
            const double s = rsqrt(fma(r2, inv_f * inv_f, 1.0));
            k = inv_f * s;
            vy = s;
            break;
        }
        default: {  // equidistant, thoby: the latitude itself is needed
            if (!(r2 > 0.0)) return 2;
            const double inv_r = rsqrt(r2);
            const double d = r2 * inv_r * inv_f;
            double lat = d;
            if (g.lens != PB_LENS_EQUIDISTANT) lat = asin(d / 1.47) / 0.713;  // (a branch: `?:` would evaluate asin for both)
            double sl, cl;
            sincos(lat, &sl, &cl);
            k = sl * inv_r;
            vy = cl;
            break;
        }
    }
    vx = x * k;
    vz = y * k;
    if (right) vy = -vy;  // lat = pi - lat (projection.py:381-382)
    return 0;
}

// dist / sin(lat) of a camera source lens for a ray with cos(lat) = ny, sin(lat) = sqrt(h2),
// dist = lens_forward(lat) * f (projection.py:251).  0 = ok, 1 = no pixel (NaN radius), 2 = undecided.
__device__ __forceinline__ int fast_lens_q(int lens, const FastGeom& fg, double ny, double inv_h, double theta, double& q) {
    const double f = fg.src_f;
    switch (lens) {
        case PB_LENS_EQUIDISTANT: q = theta * f * inv_h; return 0;
        case PB_LENS_EQUISOLID:  // 2 sin(t/2) / sin t = 1 / cos(t/2)
            if (!(1.0 + ny > 1e-8)) return 2;
            q = f * rsqrt(0.5 * (1.0 + ny));
            return 0;
        case PB_LENS_ORTHOGRAPHIC: q = f; return 0;
        case PB_LENS_STEREOGRAPHIC:  // 2 tan(t/2) / sin t = 2 / (1 + cos t)
            if (!(1.0 + ny > 1e-8)) return 2;
            q = 2.0 * f / (1.0 + ny);
            return 0;
        case PB_LENS_RECTILINEAR:  // tan t / sin t = 1 / cos t, defined for t <= 89 deg (lens.py:97-98)
            if (ny > fg.ny_rect_in) {
                q = f / ny;
                return 0;
            }
            return (ny < fg.ny_rect_out) ? 1 : 2;
        default: q = 1.47 * sin(0.713 * theta) * f * inv_h; return 0;
    }
}

__device__ __forceinline__ bool lens_needs_angle(int lens) {
    return lens == PB_LENS_EQUIDISTANT || lens == PB_LENS_THOBY;
}

// One camera sample from (cos lon, sin lon) * dist = (nx, nz) * q: false = undecided; xy = packed pixel or none.
__device__ __forceinline__ bool fast_camera_xy(double nx, double nz, double q, int h, int w, double cy, double cx,
                                               int col0, bool flip, int& xy) {
    const FastIndex ix = fast_index(fma(nx, q, cx), w), iy = fast_index(fma(-nz, q, cy), h);
    const int col = col0 + (flip ? (w - 1 - ix.idx) : ix.idx);
    xy = (ix.inside & iy.inside) ? ((iy.idx << 16) | col) : kNoPixel;
    return ix.decided & iy.decided;
}

// Source lookup of a rotated unit vector; false = undecided.
template <int SRC_KIND>
__device__ __forceinline__ bool fast_src_lookup(const SrcGeom& s, const FastGeom& fg, double nx, double ny, double nz,
                                                Lookup& L) {
    L.xy0 = L.xy1 = kNoPixel;
    L.w0 = L.w1 = 1.0;
    const double h2 = fma(nx, nx, nz * nz);
    if (!(h2 > 1e-8)) return false;  // within 1e-4 rad of a pole: acos / atan2 are ill-conditioned there
    if (SRC_KIND == PB_KIND_EQUIRECT) {
        const double lat = fast_acos_sc(h2 * rsqrt(h2), ny);
        const double lon = fast_atan2(nz, nx);
        const FastIndex row = fast_index(lat * fg.inv_seg_h, s.H);
        const FastIndex col = fast_index(fma(lon, fg.inv_seg_w, s.half_w), s.W);
        L.xy0 = (row.idx << 16) | col.idx;
        // (a coordinate outside [0, n) wraps around in the reference: exact chain)
        return row.decided & col.decided & row.inside & col.inside;
    }
    const double inv_h = rsqrt(h2);
    if (SRC_KIND == PB_KIND_CAMERA) {
        double theta = 0.0;
        if (lens_needs_angle(s.lens)) theta = fast_acos_sc(h2 * inv_h, ny);  // (a branch: `?:` would evaluate both)
        double q;
        const int st = fast_lens_q(s.lens, fg, ny, inv_h, theta, q);
        if (st == 2) return false;
        if (st == 1) return true;
        return fast_camera_xy(nx, nz, q, s.H, s.W, s.cy, s.cx, 0, false, L.xy0);
    }
    // double source (projection.py:408-462): unit weights only, the blend band takes the exact chain
    if (!((ny > fg.ny_band_hi) || (ny < fg.ny_band_lo))) return false;
    const bool ang = lens_needs_angle(s.lens);
    double theta = 0.0;
    if (ang) theta = fast_acos_sc(h2 * inv_h, ny);
    double ql, qr;
    const int sl = fast_lens_q(s.lens, fg, ny, inv_h, theta, ql);
    const int sr = fast_lens_q(s.lens, fg, -ny, inv_h, ang ? kPi - theta : 0.0, qr);
    if (sl == 2 || sr == 2) return false;
    if (sl == 0 && !fast_camera_xy(nx, nz, ql, s.H, s.wl, s.cy, s.cxl, 0, false, L.xy0)) return false;
    if (sr == 0 && !fast_camera_xy(nx, nz, qr, s.H, s.wr, s.cy, s.cxr, s.wl, true, L.xy1)) return false;
    return true;
}

// The whole short cut for output pixel (i, j); false = undecided (run the exact chain).
template <int OUT_KIND, int SRC_KIND>
__device__ __forceinline__ bool fast_lookup(const OutGeom& out, const FastGeom& fg, const Rotations& rot,
                                            const SrcGeom& src, int i, int j, Lookup& L) {
    double vx, vy, vz;
    const int st = fast_out_vector<OUT_KIND>(out, fg, i, j, vx, vy, vz);
    if (st == 2) return false;
    if (st == 1) {
        L.xy0 = L.xy1 = kNoPixel;
        L.w0 = L.w1 = 1.0;
        return true;
    }
    if (fg.has_rot) {
        const double* __restrict__ m = fg.rot;
        const double tx = fma(m[2], vz, fma(m[1], vy, m[0] * vx));
        const double ty = fma(m[5], vz, fma(m[4], vy, m[3] * vx));
        const double tz = fma(m[8], vz, fma(m[7], vy, m[6] * vx));
        vx = tx;
        vy = ty;
        vz = tz;
    }
    (void)rot;
    return fast_src_lookup<SRC_KIND>(src, fg, vx, vy, vz, L);
}

// The exact chain, out of line: the rare fall-back of the short cut.
template <int OUT_KIND, int SRC_KIND>
__device__ __noinline__ Lookup exact_lookup(const OutGeom& out, const Rotations& rot, const SrcGeom& src, int i, int j) {
    Ray r = output_ray<OUT_KIND>(out, i, j);
    for (int n = 0; n < rot.n; ++n) r = rotate_ray(r, rot.m[n]);
    return source_lookup<SRC_KIND>(src, r);
}

}  // namespace pb
