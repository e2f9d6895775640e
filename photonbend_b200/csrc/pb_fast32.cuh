// pb_fast32.cuh -- FP32-first evaluation of the per-pixel chain, with an error-aware guard band.
//
// BASELINE.json's north_star asks for FP32 trig; SURVEY.md section 0.3 measured that plain FP32 misses
// the parity bar (3e-4 .. 1e-3 of the pixels land on the wrong side of an integer boundary) and
// admits FP32 only as "FP32 + guard band + FP64 recompute inside the band and near the acos/atan2
// singularities".  This is that scheme, as the first of three tiers:
//
//   tier 1  (this file)   the whole chain in float -- unit vector of the output pixel (algebraic
//                         lens inverses), ONE composed rotation matrix, algebraic lens forwards,
//                         a degree-4 polynomial atan2 where an angle is needed -- ~70 instructions,
//                         together with a bound E on its own error in source pixels;
//   tier 2  pb_fast.cuh   the same in float64 with a 2^-19 px guard (for the pixels tier 1 cannot
//                         decide: their coordinate lies within E of an integer, ~1 % of the pixels);
//   tier 3  pb_device.cuh the reference's exact chain (for what tier 2 cannot decide, ~4e-6).
//
// A tier-1 result is used only if EVERY decision taken on the way is further than its error bound
// from the decision boundary: each truncation to a source index (the coordinate is further than E
// from every integer), the fov test of the output lens, lens domains, the blend band of a double
// source, the poles of acos / atan2.  The output is therefore that of the exact chain bit for bit.
//
// Error bound.  With eps = 2^-24 (half an ulp) the unit vector after the rotation carries an
// absolute error of a few eps per component (dn); propagated to the source coordinate
// fx = cos(lon) * dist + cx = nx * q + cx (q = dist / sin(lat)) that gives
//     |error(fx)| <= dn * (amp + 2 q)
// where amp = sin(lat) * dq/dn is the lens-specific amplification (1.4 f + q for an equidistant
// lens, q tan(lat/2) for a stereographic one, ...; derivation in DESIGN.md), and for a panorama
// source dn * 1.5 / seg_h for the row, dn / (sin(lat) seg_w) for the column.  E = K eps * (that
// shape); K is calibrated on the GPU against the float64 evaluation of the SAME formulas
// (pb_debug_fast32_stats: the largest |float - double| / shape over every pixel of a geometry),
// with a factor of >= 2 on the largest ratio seen over the whole case matrix
// (tests/test_gpu_parity.py::test_fp32_tier_error_bound).
#pragma once

#include "pb_fast.cuh"

namespace pb {

// ------------------------------------------------------------------------------------ scalar helpers, float and double

template <typename T> struct F32Ops;
template <> struct F32Ops<float> {
    static __device__ __forceinline__ float rsqrt(float x) {
        float r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float rcp(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ void sincos(float a, float* s, float* c) { sincosf(a, s, c); }
    static __device__ __forceinline__ float asin(float a) { return asinf(a); }
    static __device__ __forceinline__ float sin(float a) { return sinf(a); }
};
template <> struct F32Ops<double> {
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ void sincos(double a, double* s, double* c) { ::sincos(a, s, c); }
    static __device__ __forceinline__ double asin(double a) { return ::asin(a); }
    static __device__ __forceinline__ double sin(double a) { return ::sin(a); }
};

// atan2(y, x) for (x, y) well away from the origin: reduction to |t| <= tan(pi/8), then
// t * P(t^2) with a degree-4 P (|error| < 6e-8 evaluated in float; tests/analysis/atan_fit.py).
// In double (the calibration reference) libm's atan2 stands in.
__device__ __forceinline__ float atan2_32(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const bool big = mn > 0.41421357f * mx;
    const float t = (big ? mn - mx : mn) * F32Ops<float>::rcp(big ? mn + mx : mx);
    const float u = t * t;
    float p = 0.07976041734218597f;
    p = fmaf(p, u, -0.1384841501712799f);
    p = fmaf(p, u, 0.19974075257778168f);
    p = fmaf(p, u, -0.33332785964012146f);
    p = fmaf(p, u, 1.0f);
    float r = fmaf(t, p, big ? 0.78539816339744831f : 0.0f);
    if (ay > ax) r = 1.5707963267948966f - r;
    if (x < 0.0f) r = 3.14159265358979324f - r;
    return (y < 0.0f) ? -r : r;
}
__device__ __forceinline__ double atan2_32(double y, double x) { return atan2(y, x); }

// ------------------------------------------------------------------------------------ the chain, generic in the float type

// What tier 1 knows about one output pixel.  Coordinates are CENTRED: the source coordinate along
// an axis is centre + v, with centre = W/2 - 0.5 (camera), W/2 (panorama columns), H/2 (panorama
// rows) added in integer arithmetic afterwards, so that float keeps ~2^-13 px of resolution.
template <typename T>
struct Coords32 {
    int status;     // 0 = coordinates valid, 1 = black pixel (outside the fov / no source pixel), 2 = undecided
    T vx, vy;       // slot 0 (left lens / only)
    T ex, ey;       // error shapes of vx, vy (multiply by K eps)
    T wx, wy, fx_, fy_;  // slot 1 (right lens of a double source) and its error shapes
    int slot1;      // 0 = coordinates valid, 1 = no pixel
};

template <typename T, int OUT_KIND>
__device__ __forceinline__ int out_vector32(const OutGeom& g, const Fast32GeomT<T>& fg, int i, int j, T& vx, T& vy, T& vz) {
    using O = F32Ops<T>;
    if (OUT_KIND == PB_KIND_EQUIRECT) {
        const T lon = O::fma((T)j, fg.lon_step, fg.lon0);
        const T lat = (T)i * fg.lat_step;
        T sl, cl, so, co;
        O::sincos(lat, &sl, &cl);
        O::sincos(lon, &so, &co);
        vx = co * sl;
        vy = cl;
        vz = so * sl;
        return 0;
    }
    const bool right = (OUT_KIND == PB_KIND_DOUBLE) && j >= g.half_w;
    T x = (T)(right ? j - g.half_w : j) + fg.x0;
    if (right) x = -x;
    const T y = fg.y0 - (T)i;
    const T r2 = O::fma(x, x, y * y);
    // outside the fov (the reference flags the ray invalid: black) -- but only where the lens inverse is
    // still defined: beyond its domain the reference's latitude is NaN, which no comparison flags
    if (!(r2 < fg.r2_valid)) return (r2 > fg.r2_invalid && r2 < fg.r2_nan) ? 1 : 2;
    if (!(r2 < fg.r2_domain)) return 2;
    const T inv_f = fg.inv_f;
    const int lens = g.lens;  // (uniform: an if-chain costs two instructions per test, a jump table seven)
    T k;  // sin(lat) / r
    if (lens == PB_LENS_EQUISOLID) {
        const T u2 = r2 * fg.quarter_inv_f2;
        const T w = (T)1 - u2;
        k = w * O::rsqrt(w) * inv_f;
        vy = O::fma((T)-2, u2, (T)1);
    } else if (lens == PB_LENS_RECTILINEAR) {
        const T s = O::rsqrt(O::fma(r2, inv_f * inv_f, (T)1));
        k = inv_f * s;
        vy = s;
    } else if (lens == PB_LENS_STEREOGRAPHIC) {
        const T t2 = r2 * fg.quarter_inv_f2;
        const T w = O::rcp((T)1 + t2);
        k = inv_f * w;
        vy = ((T)1 - t2) * w;
    } else if (lens == PB_LENS_ORTHOGRAPHIC) {
        k = inv_f;
        const T w = O::fma(-r2, inv_f * inv_f, (T)1);
        vy = w * O::rsqrt(w);
    } else {
        if (!(r2 > (T)0.25)) return 2;  // the centre pixel: longitude undefined
        const T inv_r = O::rsqrt(r2);
        const T d = r2 * inv_r * inv_f;
        T lat = d;
        if (lens != PB_LENS_EQUIDISTANT) lat = O::asin(d * (T)(1.0 / 1.47)) * (T)(1.0 / 0.713);
        T sl, cl;
        O::sincos(lat, &sl, &cl);
        k = sl * inv_r;
        vy = cl;
    }
    vx = x * k;
    vz = y * k;
    if (right) vy = -vy;
    return 0;
}

// q = dist / sin(lat) of a camera lens and amp = sin(lat) * dq/dn.  0 = ok, 1 = no pixel, 2 = undecided.
template <typename T>
__device__ __forceinline__ int lens_q32(int lens, const Fast32GeomT<T>& fg, T ny, T h, T inv_h, T& q, T& amp) {
    using O = F32Ops<T>;
    const T f = fg.src_f;
    if (lens == PB_LENS_EQUIDISTANT) {
        const T theta = atan2_32(h, ny);
        q = theta * f * inv_h;
        amp = O::fma((T)1.4, f, q);
        return 0;
    }
    if (lens == PB_LENS_EQUISOLID) {
        const T w = (T)1 + ny;
        if (!(w > (T)1e-3)) return 2;
        q = f * O::rsqrt((T)0.5 * w);
        amp = (T)0.5 * q * h * O::rcp(w);
        return 0;
    }
    if (lens == PB_LENS_STEREOGRAPHIC) {
        const T w = (T)1 + ny;
        if (!(w > (T)1e-3)) return 2;
        const T iw = O::rcp(w);
        q = (T)2 * f * iw;
        amp = q * h * iw;
        return 0;
    }
    if (lens == PB_LENS_RECTILINEAR) {
        if (ny > fg.ny_rect_in) {
            const T iy = O::rcp(ny);
            q = f * iy;
            amp = q * h * iy;
            return 0;
        }
        return (ny < fg.ny_rect_out) ? 1 : 2;
    }
    if (lens == PB_LENS_ORTHOGRAPHIC) {
        q = f;
        amp = (T)0;
        return 0;
    }
    const T theta = atan2_32(h, ny);
    q = (T)1.47 * O::sin((T)0.713 * theta) * f * inv_h;
    amp = O::fma((T)1.5, f, q);
    return 0;
}

template <typename T, int OUT_KIND, int SRC_KIND>
__device__ __forceinline__ Coords32<T> coords32(const OutGeom& out, const Fast32GeomT<T>& fg, const SrcGeom& src, int i, int j) {
    using O = F32Ops<T>;
    Coords32<T> c;
    c.slot1 = 1;
    c.vx = c.vy = c.ex = c.ey = c.wx = c.wy = c.fx_ = c.fy_ = (T)0;
    T vx, vy, vz;
    c.status = out_vector32<T, OUT_KIND>(out, fg, i, j, vx, vy, vz);
    if (c.status != 0) return c;
    T nx = vx, ny = vy, nz = vz;
    if (fg.has_rot) {
        nx = O::fma(fg.rot[2], vz, O::fma(fg.rot[1], vy, fg.rot[0] * vx));
        ny = O::fma(fg.rot[5], vz, O::fma(fg.rot[4], vy, fg.rot[3] * vx));
        nz = O::fma(fg.rot[8], vz, O::fma(fg.rot[7], vy, fg.rot[6] * vx));
    }
    const T h2 = O::fma(nx, nx, nz * nz);
    if (!(h2 > (T)1e-6)) {  // within 1e-3 rad of a pole
        c.status = 2;
        return c;
    }
    const T inv_h = O::rsqrt(h2);
    const T h = h2 * inv_h;
    if (SRC_KIND == PB_KIND_EQUIRECT) {
        // rows: (lat - pi/2) / seg_h, lat - pi/2 = atan2(-ny, h); columns: lon / seg_w
        c.vy = atan2_32(-ny, h) * fg.inv_seg_h;
        c.vx = atan2_32(nz, nx) * fg.inv_seg_w;
        c.ey = (T)1.5 * fg.inv_seg_h;
        c.ex = O::fma(inv_h, (T)1, (T)0.5) * fg.inv_seg_w;
        return c;
    }
    if (SRC_KIND == PB_KIND_CAMERA) {
        T q, amp;
        const int st = lens_q32<T>(src.lens, fg, ny, h, inv_h, q, amp);
        if (st != 0) {
            c.status = st;
            return c;
        }
        c.vx = nx * q;
        c.vy = -nz * q;
        c.ex = c.ey = O::fma((T)2, q, amp);
        return c;
    }
    // double source: unit weights only (the blend band goes to the float64 tiers)
    if (!((ny > fg.ny_band_hi) || (ny < fg.ny_band_lo))) {
        c.status = 2;
        return c;
    }
    T ql, al, qr, ar;
    const int sl = lens_q32<T>(src.lens, fg, ny, h, inv_h, ql, al);
    const int sr = lens_q32<T>(src.lens, fg, -ny, h, inv_h, qr, ar);
    if (sl == 2 || sr == 2) {
        c.status = 2;
        return c;
    }
    c.status = 0;
    if (sl == 0) {
        c.vx = nx * ql;
        c.vy = -nz * ql;
        c.ex = c.ey = O::fma((T)2, ql, al);
    } else {
        c.status = 1;  // slot 0 has no pixel
    }
    if (sr == 0) {
        c.slot1 = 0;
        c.wx = nx * qr;
        c.wy = -nz * qr;
        c.fx_ = c.fy_ = O::fma((T)2, qr, ar);
    }
    return c;
}

// ------------------------------------------------------------------------------------ decisions (float only)

// Index along an axis of n pixels for the centred coordinate v, centre = c_int + c_half * 0.5:
// coordinate = c_int + (v + 0.5 c_half); the reference truncates toward zero and then tests
// 0 <= index < n (projection.py:223-231, 254-259), so a coordinate in (-1, 0) is index 0.
// decided: further than e from every integer.  One conversion each way (F2I.FLOOR, I2F): a value
// too large for an int saturates and a NaN converts to 0 -- either way the "fraction" fails the
// test below and the pixel goes to the float64 tiers.
struct Index32 {
    int idx;
    bool decided, inside;
};
__device__ __forceinline__ Index32 index32(float v, float e, int c_int, int c_half, int n) {
    const float t = v + 0.5f * (float)c_half;
    const int fr = __float2int_rd(t);          // floor of the centred coordinate
    const float d = t - (float)fr;             // its fraction, exact
    Index32 r;
    r.decided = fminf(d, 1.0f - d) > e;
    int fl = c_int + fr;                       // floor of the coordinate
    if (fl == -1) fl = 0;                      // (-1, 0) truncates to 0
    r.idx = fl;
    r.inside = (unsigned)fl < (unsigned)n;
    return r;
}

// Tier 1 for output pixel (i, j): true = decided, L filled in (weights 1).
template <int OUT_KIND, int SRC_KIND>
__device__ __forceinline__ bool fast32_lookup(const OutGeom& out, const Fast32Geom& fg, const SrcGeom& src, int i, int j,
                                              Lookup& L) {
    L.xy0 = L.xy1 = kNoPixel;
    L.w0 = L.w1 = 1.0;
    const Coords32<float> c = coords32<float, OUT_KIND, SRC_KIND>(out, fg, src, i, j);
    if (c.status == 2) return false;
    if (SRC_KIND == PB_KIND_EQUIRECT) {
        if (c.status == 1) return true;
        const Index32 row = index32(c.vy, fg.k_eps * c.ey, src.H >> 1, src.H & 1, src.H);
        const Index32 col = index32(c.vx, fg.k_eps * c.ex, src.W >> 1, src.W & 1, src.W);
        L.xy0 = (row.idx << 16) | col.idx;
        // (a coordinate outside [0, n) wraps around in the reference: float64 tiers)
        return row.decided & col.decided & row.inside & col.inside;
    }
    if (SRC_KIND == PB_KIND_CAMERA) {
        if (c.status == 1) return true;
        // centre = W/2 - 0.5: integer part (W - 1) >> 1, half flag = W even
        const Index32 ix = index32(c.vx, fg.k_eps * c.ex, (src.W - 1) >> 1, (src.W & 1) ^ 1, src.W);
        const Index32 iy = index32(c.vy, fg.k_eps * c.ey, (src.H - 1) >> 1, (src.H & 1) ^ 1, src.H);
        if (ix.inside & iy.inside) L.xy0 = (iy.idx << 16) | ix.idx;
        return ix.decided & iy.decided;
    }
    bool ok = true;
    if (c.status == 0) {
        const Index32 ix = index32(c.vx, fg.k_eps * c.ex, (src.wl - 1) >> 1, (src.wl & 1) ^ 1, src.wl);
        const Index32 iy = index32(c.vy, fg.k_eps * c.ey, (src.H - 1) >> 1, (src.H & 1) ^ 1, src.H);
        if (ix.inside & iy.inside) L.xy0 = (iy.idx << 16) | ix.idx;
        ok = ix.decided & iy.decided;
    }
    if (c.slot1 == 0) {
        const Index32 ix = index32(c.wx, fg.k_eps * c.fx_, (src.wr - 1) >> 1, (src.wr & 1) ^ 1, src.wr);
        const Index32 iy = index32(c.wy, fg.k_eps * c.fy_, (src.H - 1) >> 1, (src.H & 1) ^ 1, src.H);
        if (ix.inside & iy.inside) L.xy1 = (iy.idx << 16) | (src.wl + (src.wr - 1 - ix.idx));
        ok = ok & ix.decided & iy.decided;
    }
    return ok;
}

// The three tiers for output pixel (i, j).
template <int OUT_KIND, int SRC_KIND>
__device__ __forceinline__ Lookup resolve_lookup(const OutGeom& out, const FastGeom& fg, const Rotations& rot,
                                                 const SrcGeom& src, int i, int j) {
    Lookup L;
    if (fg.f32.enabled && fast32_lookup<OUT_KIND, SRC_KIND>(out, fg.f32, src, i, j, L)) return L;
    if (fg.enabled && fast_lookup<OUT_KIND, SRC_KIND>(out, fg, rot, src, i, j, L)) return L;
    return exact_lookup<OUT_KIND, SRC_KIND>(out, rot, src, i, j);
}

// Tiers 2 and 3 only (what tier 1 could not decide).
template <int OUT_KIND, int SRC_KIND>
__device__ __forceinline__ Lookup resolve_lookup64(const OutGeom& out, const FastGeom& fg, const Rotations& rot,
                                                   const SrcGeom& src, int i, int j) {
    Lookup L;
    if (fg.enabled && fast_lookup<OUT_KIND, SRC_KIND>(out, fg, rot, src, i, j, L)) return L;
    return exact_lookup<OUT_KIND, SRC_KIND>(out, rot, src, i, j);
}

}  // namespace pb
