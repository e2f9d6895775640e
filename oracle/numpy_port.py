"""NumPy restatement of photonbend's per-pixel remap path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  ``photonbend_b200`` never does.

It restates, in whole-array float64 NumPy, the algorithm of the reference
(`/root/reference/photonbend`, pure Python + NumPy).  Every transcendental the
reference evaluates lives in NumPy (pinned ``numpy 1.23.3`` in the reference's
``poetry.lock:145-146``; this image has NumPy 2.3.5) -- the restatement calls the
*same* ufuncs in the *same* order (``np.exp`` of a complex array for cos/sin of a
longitude, ``np.log(complex).imag`` for atan2, ``np.matmul`` for the rotation, and
``.astype(int)`` for truncation), so on the same NumPy build it is bit-identical to
the reference: coordinate maps compare equal as float64, images compare equal as
uint8.  Parity is pinned by ``tests/golden/`` (vectors generated from the live
reference by ``tests/golden/make_golden.py``) and checked by
``tests/test_oracle_golden.py``.

Geometry is described by plain dicts so that the oracle shares no code with the
product:

    {"kind": "camera",   "height": H, "width": W, "lens": "equidistant", "fov": rad, "magnitude": M|None}
    {"kind": "double",   "height": H, "width": W, "lens": "equidistant", "fov": sensor_fov_rad}
    {"kind": "equirect", "height": H, "width": W}

A coordinate map is float64 ``(H, W, 3)`` = (latitude, longitude, invalid != 0)
(reference ``photonbend/core/__init__.py:42-49``).
"""

from __future__ import annotations

import warnings

import numpy as np

LENS_NAMES = (
    "equidistant",
    "equisolid",
    "orthographic",
    "stereographic",
    "rectilinear",
    "thoby",
)

# thoby constants, reference lens.py:303-304, 331-332
_THOBY_K1 = 1.47
_THOBY_K2 = 0.713


def deg2rad(deg):
    """reference photonbend/utils/__init__.py:27-37 -- ``deg / 180 * pi`` in that order."""
    return deg / 180 * np.pi


# --------------------------------------------------------------------------- lenses


def lens_forward(lens: str, theta):
    """angle of incidence -> radius in focal units.  reference lens.py:75-103 (rectilinear),
    126-144 (stereographic), 168-187 (equidistant), 224-243 (equisolid), 266-286
    (orthographic), 313-335 (thoby)."""
    if lens == "equidistant":
        return theta
    if lens == "equisolid":
        return 2 * np.sin(theta / 2.0)
    if lens == "orthographic":
        return np.sin(theta)
    if lens == "stereographic":
        return 2.0 * np.tan(theta / 2.0)
    if lens == "thoby":
        return _THOBY_K1 * np.sin(_THOBY_K2 * theta)
    if lens == "rectilinear":
        limit = deg2rad(89)
        if isinstance(theta, float):  # np.float64 is a float too (lens.py:88)
            if theta < 0:
                raise ValueError("The angle theta cannot be negative")
            if theta > limit:
                raise ValueError(
                    "The Rectilinear lens can't handle FoV larger than 179 degrees"
                )
            return np.tan(theta)
        out_of_domain = np.logical_or(theta < 0, theta > limit)
        r = np.tan(theta)
        r[out_of_domain] = np.nan
        return r
    raise KeyError(lens)


def lens_inverse(lens: str, d):
    """radius in focal units -> angle of incidence (arrays only on the hot path).
    reference lens.py:68-72, 106-124, 147-165, 190-220, 246-262, 289-309."""
    if lens == "equidistant":
        return d
    if lens == "equisolid":
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            theta = 2.0 * np.arcsin(d / 2.0)
        theta[np.isnan(theta)] = 0.0  # out-of-domain radii become latitude 0 (valid!)
        return theta
    if lens == "orthographic":
        return np.arcsin(d)
    if lens == "stereographic":
        return 2.0 * np.arctan(d / 2.0)
    if lens == "rectilinear":
        return np.arctan(d)
    if lens == "thoby":
        return np.arcsin(d / _THOBY_K1) / _THOBY_K2
    raise KeyError(lens)


def focal_distance(geom: dict) -> float:
    """pixels per focal unit.  reference projection.py:118-144 (camera), 315-339 (double)."""
    if geom["kind"] == "double":
        magnitude = geom["height"] / 2.0
    else:
        magnitude = geom.get("magnitude")
        if magnitude is None:
            magnitude = geom["height"] / 2.0
    return magnitude / lens_forward(geom["lens"], geom["fov"] / 2)


# --------------------------------------------------------------------------- helpers


def _pair_to_complex(x, y):
    """reference _shared.py:25-55: broadcast then reinterpret (x, y) as complex128."""
    fx = x + y * 0
    fy = y + x * 0
    z = np.empty(fx.shape, dtype=np.complex128)
    z.real = fx
    z.imag = fy
    return z


def _angle_of(x, y):
    """atan2(y, x) the way the reference gets it: imaginary part of a complex log
    (projection.py:193, 383; rotation.py:159-164)."""
    with np.errstate(all="ignore"):
        return np.log(_pair_to_complex(x, y)).imag


def _stack_map(lat, lon, invalid):
    h, w = lat.shape
    out = np.empty((h, w, 3), dtype=np.float64)
    out[:, :, 0] = lat
    out[:, :, 1] = lon
    out[:, :, 2] = invalid
    return out


# --------------------------------------------------------------------------- output rays


def _band(axis, rows):
    """Row-band restriction: slicing the 1-D row axis before the mesh is built gives the
    same values as slicing the finished map."""
    return axis if rows is None else axis[rows[0] : rows[1]]


def equirect_rays(geom: dict, rows=None):
    """reference projection.py:487-513."""
    h, w = geom["height"], geom["width"]
    half_px = np.pi / w / 2
    lon = np.linspace(-np.pi + half_px, np.pi - half_px, num=w)
    lat = _band(np.linspace(0, np.pi, num=h), rows)
    lat2d, lon2d = np.meshgrid(lat, lon, sparse=False, indexing="ij")
    return _stack_map(lat2d, lon2d, np.zeros(lat2d.shape))


def camera_rays(geom: dict, rows=None):
    """reference projection.py:147-194."""
    h, w = geom["height"], geom["width"]
    f = focal_distance(geom)
    xs = np.linspace(-w / 2 + 0.5, w / 2 - 0.5, num=w)
    ys = _band(np.linspace(h / 2 - 0.5, -h / 2 + 0.5, num=h), rows)
    gy, gx = np.meshgrid(ys, xs, sparse=True, indexing="ij")
    with np.errstate(all="ignore"):
        d = np.sqrt(gx**2 + gy**2) / f
        lat = lens_inverse(geom["lens"], d)
        lon = _angle_of(gx, gy)
        invalid = lat > geom["fov"] / 2
    return _stack_map(lat, lon, invalid.astype(np.float64))


def double_rays(geom: dict, rows=None):
    """reference projection.py:341-406."""
    h, w = geom["height"], geom["width"]
    hw = w // 2
    f = focal_distance(geom)
    half_xs = np.linspace(-hw / 2 + 0.5, hw / 2 - 0.5, num=hw)
    xs = np.concatenate([half_xs, half_xs * (-1)], 0)  # right lens is mirrored
    ys = _band(np.linspace(h / 2 - 0.5, -h / 2 + 0.5, num=h), rows)
    gy, gx = np.meshgrid(ys, xs, sparse=True, indexing="ij")
    with np.errstate(all="ignore"):
        d = np.sqrt(gx**2 + gy**2) / f
        lat = lens_inverse(geom["lens"], d)
        lat[:, hw:] *= -1
        lat[:, hw:] += np.pi
        lon = _angle_of(gx, gy)
        invalid = lat > geom["fov"] / 2.0
        invalid[:, hw:] = lat[:, hw:] < np.pi - (geom["fov"] / 2.0)
    return _stack_map(lat, lon, invalid.astype(np.float64))


def output_rays(geom: dict, rows=None):
    kind = geom["kind"]
    if kind == "equirect":
        return equirect_rays(geom, rows)
    if kind == "camera":
        return camera_rays(geom, rows)
    if kind == "double":
        return double_rays(geom, rows)
    raise KeyError(kind)


# --------------------------------------------------------------------------- rotation


def rotation_matrix(pitch: float, yaw: float, roll: float):
    """``Rotation(pitch, yaw, roll).rotation_matrix``: reference rotation.py:27-62 with the
    sign flip of rotation.py:100."""
    p, y, r = -pitch, -yaw, -roll
    cp, sp = np.cos(p), np.sin(p)
    cy, sy = np.cos(y), np.sin(y)
    cr, sr = np.cos(r), np.sin(r)
    m_pitch = np.array((1, 0, 0, 0, cp, sp, 0, -sp, cp)).reshape((3, 3))
    m_yaw = np.array((cy, 0, -sy, 0, 1, 0, sy, 0, cy)).reshape((3, 3))
    m_roll = np.array((cr, sr, 0, -sr, cr, 0, 0, 0, 1)).reshape((3, 3))
    return m_pitch @ m_yaw @ m_roll


def rotate_map(cmap, matrix):
    """reference rotation.py:102-176.  Zeroes invalid entries of ``cmap`` in place, like
    the reference does."""
    bad = cmap[:, :, 2] != 0.0
    polar = cmap[:, :, :2]
    polar[bad] = 0
    lat = polar[:, :, 0]
    lon = polar[:, :, 1]
    with np.errstate(all="ignore"):
        vy = np.cos(lat)
        vxz = np.exp(lon * 1j) * np.sin(lat)
        vec = np.empty(lat.shape + (3, 1), dtype=np.float64)
        vec[:, :, 0, 0] = vxz.real
        vec[:, :, 1, 0] = vy
        vec[:, :, 2, 0] = vxz.imag
        turned = np.matmul(matrix, vec)[..., 0]
        new_lat = np.arccos(turned[:, :, 1])
        new_lon = _angle_of(turned[:, :, 0], turned[:, :, 2])
    new_lat[bad] = 0
    new_lon[bad] = 0
    return _stack_map(new_lat, new_lon, bad)


# --------------------------------------------------------------------------- sampling


def _camera_pixel_positions(lat, lon, lens, f, h, w):
    """reference projection.py:247-274."""
    cy = h / 2 - 0.5
    cx = w / 2 - 0.5
    with np.errstate(all="ignore"):
        dist = lens_forward(lens, lat) * f
        p = np.exp(lon * 1j) * dist
        py = ((p.imag * (-1)) + cy).astype(int)
        px = (p.real + cx).astype(int)
    return px, py


def _sample_camera_arrays(image, lens, f, lat, lon, invalid):
    """reference projection.py:197-245 (no source-fov test, truncation before bounds)."""
    h, w = image.shape[:2]
    px, py = _camera_pixel_positions(lat, lon, lens, f, h, w)
    bad_y = np.logical_or(py >= h, py < 0)
    py[bad_y] = 0
    bad_x = np.logical_or(px >= w, px < 0)
    px[bad_x] = 0
    out = image[py, px]
    out[np.logical_or(bad_y, bad_x)] = 0
    out[invalid] = 0
    return out


def sample_camera(geom: dict, image, cmap):
    invalid = cmap[:, :, 2] != 0.0
    return _sample_camera_arrays(
        image, geom["lens"], focal_distance(geom), cmap[:, :, 0], cmap[:, :, 1], invalid
    )


def sample_double(geom: dict, image, cmap):
    """reference projection.py:408-462."""
    hs, ws = image.shape[:2]
    wl = ws // 2
    fov = geom["fov"]
    ref = (fov / 2) - (np.pi / 2)
    lo = np.pi / 2 - ref
    hi = np.pi / 2 + ref
    span = 2.0 * ref
    safety = deg2rad(0.5)

    invalid = cmap[:, :, 2] != 0.0
    lon = cmap[:, :, 1]
    lat_l = cmap[:, :, 0]
    lat_r = np.copy(lat_l)
    lat_r *= -1
    lat_r += np.pi

    img_l = image[:, :wl]
    img_r = np.copy(image[:, wl:])[:, ::-1]
    # both halves are plain cameras with the default magnitude (their height / 2)
    sub = {"kind": "camera", "lens": geom["lens"], "fov": fov, "magnitude": None}
    f_l = focal_distance(dict(sub, height=img_l.shape[0], width=img_l.shape[1]))
    f_r = focal_distance(dict(sub, height=img_r.shape[0], width=img_r.shape[1]))
    pix_l = _sample_camera_arrays(img_l, geom["lens"], f_l, lat_l, lon, invalid)
    pix_r = _sample_camera_arrays(img_r, geom["lens"], f_r, lat_r, lon, invalid)

    def weight(lat):
        with np.errstate(all="ignore"):
            band = np.logical_and(lat >= lo, lat <= (hi + safety))
            wgt = (lat - hi) / span * -1
        wgt[np.logical_not(band)] = 1.0
        return np.expand_dims(wgt, 2)

    with np.errstate(all="ignore"):
        blend = pix_l.astype(np.float64) * weight(lat_l) + pix_r.astype(
            np.float64
        ) * weight(lat_r)
        out = blend.astype(np.uint8)  # trunc, then wraps mod 256
    out[invalid] = 0
    return out


def sample_equirect(geom: dict, image, cmap):
    """reference projection.py:515-547.  Zeroes invalid entries of ``cmap`` in place."""
    invalid = cmap[:, :, 2] != 0.0
    polar = cmap[:, :, :2]
    polar[invalid] = 0
    h, w = image.shape[:2]
    seg_w = np.pi / (w / 2)
    seg_h = np.pi / h
    with np.errstate(all="ignore"):
        row = polar[:, :, 0] / seg_h
        col = polar[:, :, 1] / seg_w + (w / 2)
        out = image[row.astype(int) % h, col.astype(int) % w]
    out[invalid] = 0
    return out


def sample(geom: dict, image, cmap):
    kind = geom["kind"]
    if kind == "camera":
        return sample_camera(geom, image, cmap)
    if kind == "double":
        return sample_double(geom, image, cmap)
    if kind == "equirect":
        return sample_equirect(geom, image, cmap)
    raise KeyError(kind)


# --------------------------------------------------------------------------- whole path


def coordinate_map(out_geom: dict, rotations=()):
    """rays of ``out_geom`` after applying ``rotations`` (an iterable of (pitch, yaw, roll)
    in radians, applied in order -- make_pano.py:126-129)."""
    cmap = output_rays(out_geom)
    for pyr in rotations:
        cmap = rotate_map(cmap, rotation_matrix(*pyr))
    return cmap


def remap(out_geom: dict, rotations, src_geom: dict, image, rows=None):
    """The reference's three-call protocol (core/__init__.py:66-92) end to end.

    ``rows=(r0, r1)`` evaluates only that band of output rows (the protocol works on any
    (h, w, 3) slice of the map; banded output == whole output bit for bit), used by the
    multi-process CPU baseline in bench.py.
    """
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cmap = output_rays(out_geom, rows)
        for pyr in rotations:
            cmap = rotate_map(cmap, rotation_matrix(*pyr))
        return sample(src_geom, image, cmap)


def map_projection(cmap):
    """projection.py:550-599: coordinate map -> RGB visualisation (zeroes the invalid (lat, lon) of
    ``cmap`` in place through the view, :563-566)."""
    rgb_range = 255.0
    invalid = cmap[:, :, 2] != 0.0
    valid = np.logical_not(invalid)
    polar = cmap[:, :, :2]
    polar[invalid] = 0
    distance = polar[:, :, 0]
    lo = np.min(distance[valid])
    hi = np.max(distance[valid])
    factor = rgb_range / (hi - lo)
    red = distance.copy()
    red[valid] -= lo
    red[valid] *= factor
    red8 = np.round(red).astype(np.uint8)
    green8 = np.round(rgb_range / (np.pi * 2) * polar[:, :, 1]).astype(np.uint8)
    blue8 = (invalid.astype(np.uint8) * 255).astype(np.uint8)
    return np.concatenate([red8[:, :, None], green8[:, :, None], blue8[:, :, None]], axis=2)
