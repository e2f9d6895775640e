"""Host-side mirror of the reference's API: parameter derivation, laziness, error behaviour.
CPU only -- nothing here launches a kernel."""

import math

import numpy as np
import pytest

import case_matrix
import helpers
from oracle import numpy_port
from photonbend_b200 import _native, engine
from photonbend_b200.core import (
    CameraImage,
    CoordinateMap,
    DoubleCameraImage,
    Lens,
    PanoramaImage,
    Rotation,
    lens as pb_lens,
)
from photonbend_b200.utils import calculate_size_panorama_to_photo, to_degrees, to_radians


def test_to_radians_is_the_reference_expression():
    for deg in (0.5, 10, 89, 140, 180, 195, 360):
        assert to_radians(deg) == numpy_port.deg2rad(deg)
    assert to_radians(360) == 2 * math.pi
    assert to_degrees(to_radians(37.0)) == pytest.approx(37.0)


@pytest.mark.parametrize("name", numpy_port.LENS_NAMES)
def test_lens_functions_match_oracle(name):
    lens = getattr(pb_lens, name)()
    theta = np.linspace(0.0, 1.5, 50)
    radius = np.linspace(0.0, 2.5, 50)
    with np.errstate(all="ignore"):
        assert np.array_equal(lens.forward_function(theta.copy()), numpy_port.lens_forward(name, theta.copy()), equal_nan=True)
        assert np.array_equal(lens.reverse_function(radius.copy()), numpy_port.lens_inverse(name, radius.copy()), equal_nan=True)
    assert lens.forward_function(0.7) == numpy_port.lens_forward(name, 0.7)
    assert pb_lens.lens_id(lens.forward_function, lens.reverse_function) == numpy_port.LENS_NAMES.index(name)


def test_rectilinear_raises_like_reference():
    with pytest.raises(ValueError):
        CameraImage(np.zeros((8, 8, 3), np.uint8), to_radians(179), pb_lens.rectilinear())
    with pytest.raises(ValueError):
        pb_lens.rectilinear().forward_function(-0.1)
    big = pb_lens.rectilinear().forward_function(np.array([0.1, to_radians(89.5), -0.2]))
    assert np.isnan(big[1]) and np.isnan(big[2]) and not np.isnan(big[0])


def test_custom_lens_is_sampled_not_refused():
    custom = Lens(lambda t: t * 1.01, lambda r: r / 1.01)
    cam = CameraImage(np.zeros((8, 8, 3), np.uint8), 2.0, custom)  # construction: f is host math
    assert cam.f_distance == pytest.approx(4.0 / 1.01)
    rays = cam.get_coordinate_map().rays  # lazy: no kernel, but the reverse function is sampled
    assert rays.out.lens == 6 and rays.out.table is not None
    assert rays.out.table[-1] == pytest.approx(rays.out.table_max / 1.01)


def test_focal_distance_and_magnitude_defaults():
    for name, geom in case_matrix.output_geometries() + case_matrix.source_geometries():
        if geom["kind"] == "equirect":
            continue
        img = helpers.product_image(geom, np.zeros((geom["height"], geom["width"], 3), np.uint8))
        assert img.f_distance == numpy_port.focal_distance(geom), name
        if geom["kind"] == "double" or geom.get("magnitude") is None:
            assert img.magnitude == geom["height"] / 2.0
    dbl = DoubleCameraImage(np.zeros((10, 20, 3), np.uint8), to_radians(190), pb_lens.equidistant(), magnitude=123.0)
    assert dbl.magnitude == 5.0  # the keyword is accepted and ignored (projection.py:296-316)


def test_rotation_matrix_bit_identical_to_oracle():
    for pyr in [(0.1, 0.2, 0.3), (to_radians(-90), 0.0, to_radians(195)), (0.0, 0.0, 0.0), (3.0, -2.0, 1.0)]:
        assert np.array_equal(Rotation(*pyr).rotation_matrix, numpy_port.rotation_matrix(*pyr))


def test_coordinate_map_is_lazy_and_shaped_like_the_array():
    pano = PanoramaImage(np.zeros((20, 40, 3), np.uint8))
    cmap = pano.get_coordinate_map()
    assert isinstance(cmap, CoordinateMap) and cmap.is_lazy
    assert cmap.shape == (20, 40, 3) and cmap.dtype == np.float64 and cmap.ndim == 3 and len(cmap) == 20
    rotated = Rotation(0.1, 0.2, 0.3).rotate_coordinate_map(cmap)
    assert rotated is not cmap and rotated.is_lazy and len(rotated.rays.rotations) == 1
    assert rotated.rays.rotations[0] == tuple(numpy_port.rotation_matrix(0.1, 0.2, 0.3).reshape(9))
    twice = Rotation(0.0, 1.0, 0.0).rotate_coordinate_map(rotated)
    assert len(twice.rays.rotations) == 2 and len(rotated.rays.rotations) == 1
    dbl = DoubleCameraImage(np.zeros((11, 23, 3), np.uint8), to_radians(190), pb_lens.equidistant())
    assert dbl.get_coordinate_map().shape == (11, 22, 3)  # 2 * (23 // 2) columns, like the reference


def test_descriptor_carries_reference_constants():
    geom = dict(case_matrix.source_geometries())["cam-equisolid-360"]
    cam = helpers.product_image(geom, np.zeros((geom["height"], geom["width"], 3), np.uint8))
    g = cam._source_geometry()
    assert (g.kind, g.height, g.width, g.lens) == (_native.KIND_CAMERA, 52, 48, _native.LENS_EQUISOLID)
    assert g.fov == geom["fov"] and g.f_distance == numpy_port.focal_distance(geom)
    desc = engine._remap_desc(engine.RayPlan(g).rotated(np.eye(3)), g, 3)
    assert desc.n_rotations == 1 and desc.channels == 3
    assert [desc.rotations[0][k] for k in range(9)] == [1, 0, 0, 0, 1, 0, 0, 0, 1]
    with pytest.raises(_native.NativeError):
        rays = engine.RayPlan(g)
        for _ in range(_native.PB_MAX_ROTATIONS + 1):
            rays = rays.rotated(np.eye(3))
        engine._remap_desc(rays, g, 3)


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    pano = PanoramaImage(np.zeros((8, 16, 3), np.uint8))
    cam = CameraImage(np.zeros((16, 16, 3), np.uint8), to_radians(360), pb_lens.equidistant())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cam.process_coordinate_map(pano.get_coordinate_map())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        np.asarray(pano.get_coordinate_map())


def test_non_uint8_sources_are_refused():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs the CPU-only error path")
    cam = CameraImage(np.zeros((16, 16, 3), np.float32), 2.0, pb_lens.equidistant())
    with pytest.raises((NotImplementedError, RuntimeError)):
        cam.process_coordinate_map(cam.get_coordinate_map())


def test_size_helper():
    assert calculate_size_panorama_to_photo((6144, 3072), pb_lens.equidistant().forward_function) == (3912, 3912)
    with pytest.raises(AssertionError):
        calculate_size_panorama_to_photo((100, 60), pb_lens.equidistant().forward_function)
    side = calculate_size_panorama_to_photo((2048, 1024), pb_lens.equisolid().forward_function, True)
    assert side[0] == side[1] and side[0] >= 922


def test_user_defined_lens_becomes_a_table():
    """A Lens of user callables (reference lens.py:48-64) is not refused: it is sampled into a
    table per role (forward on [0, pi] as a source, reverse on [0, largest pixel radius] as an
    output), the descriptor carries the host pointers, and the plan-cache key carries the tables'
    content instead of their addresses."""
    import ctypes

    import numpy as np

    from photonbend_b200 import _native, engine
    from photonbend_b200.core.lens import LENS_TABLE_SAMPLES, Lens, equisolid, lens_id
    from photonbend_b200.core.projection import CameraImage

    builtin = equisolid()
    assert lens_id(builtin.forward_function, builtin.reverse_function) == _native.LENS_EQUISOLID
    custom = Lens(lambda t: 2 * np.sin(t / 2), lambda r: 2 * np.arcsin(r / 2))
    assert lens_id(custom.forward_function, custom.reverse_function) == _native.LENS_TABLE
    cam = CameraImage(np.zeros((64, 80, 3), np.uint8), np.pi * 0.9, custom, magnitude=31.5)
    src, out = cam._source_geometry(), cam._output_geometry()
    assert src.table.shape == (LENS_TABLE_SAMPLES,) and src.table_max == np.pi
    assert np.allclose(src.table[[0, -1]], [0.0, 2.0])
    assert np.isclose(out.table_max, np.hypot(39.5, 31.5) / cam.f_distance)
    assert cam._source_geometry().table is src.table  # cached on the image object
    desc = engine._remap_desc(cam.get_coordinate_map().rays, src, 3)
    assert desc.src.lens == _native.LENS_TABLE and desc.src.lens_table_n == LENS_TABLE_SAMPLES
    assert ctypes.addressof(desc.src.lens_table.contents) == src.table.ctypes.data
    other = CameraImage(np.zeros((64, 80, 3), np.uint8), np.pi * 0.9, Lens(custom.forward_function, custom.reverse_function),
                        magnitude=31.5)
    desc2 = engine._remap_desc(other.get_coordinate_map().rays, other._source_geometry(), 3)
    # same geometry, same table CONTENT (another array at another address): one descriptor, one plan
    assert other._source_geometry().table is not src.table
    assert engine._desc_key(desc) == engine._desc_key(desc2)
    third = CameraImage(np.zeros((64, 80, 3), np.uint8), np.pi * 0.9,
                        Lens(lambda t: 2.0001 * np.sin(t / 2), custom.reverse_function), magnitude=31.5 * 2.0001 / 2)
    desc3 = engine._remap_desc(third.get_coordinate_map().rays, third._source_geometry(), 3)
    assert engine._desc_key(desc3) != engine._desc_key(desc)  # another forward function: another plan
    # C-ABI validation without a GPU: a table lens without a table is refused
    lib = _native.load()
    bad = _native.RemapDesc.from_buffer_copy(bytes(desc))
    bad.src.lens_table = None
    fake = ctypes.c_void_p(256)
    assert lib.pb_remap_u8(ctypes.byref(bad), fake, 0, fake, 0, 1, None) == _native.PB_ERR_INVALID_ARGUMENT
    assert b"PB_LENS_TABLE" in lib.pb_last_error()


def test_decode_threads_leave_room_for_the_other_ranks(monkeypatch):
    """stream._default_decode_threads: per GPU a producer, a consumer and the decode threads; never more
    than one per frame of a batch, never more than eight, and with N GPUs on one host only half of
    what is left of a GPU's share of the cores (spinning or not, 2 ranks x 7 threads on 16 cores
    measured 7 Gpix/s for both GPUs together against 22 with 2 x 4)."""
    import os

    from photonbend_b200 import stream

    for cores, want in ((16, {1: 8, 2: 4, 4: 2, 8: 1}), (32, {1: 8, 2: 8, 4: 4, 8: 2}), (2, {1: 1, 8: 1})):
        monkeypatch.setattr(os, "sched_getaffinity", lambda pid, c=cores: set(range(c)), raising=False)
        for n, threads in want.items():
            assert stream._default_decode_threads(8, n) == threads, (cores, n)
        assert stream._default_decode_threads(1, 1) == 1 and stream._default_decode_threads(3, 1) <= 3
