"""The three commands keep the reference's flags and exit behaviour (CPU: parsing and errors;
GPU: a real conversion checked against the oracle)."""

import os

import numpy as np
import pytest
from click.testing import CliRunner
from PIL import Image

from conftest import GOLDEN
from photonbend_b200.scripts.main import main


def test_commands_and_flags_exist():
    runner = CliRunner()
    top = runner.invoke(main, ["--help"])
    assert top.exit_code == 0
    for name in ("make-pano", "alter-photo", "make-photo"):
        assert name in top.output
    pano = runner.invoke(main, ["make-pano", "--help"]).output
    for flag in ("--type", "--lens", "--fov", "--rotation", "--size"):
        assert flag in pano
    alter = runner.invoke(main, ["alter-photo", "--help"]).output
    for flag in ("--itype", "--otype", "--ilens", "--olens", "--ifov", "--ofov", "--rotation", "--size"):
        assert flag in alter
    photo = runner.invoke(main, ["make-photo", "--help"]).output
    for flag in ("--type", "--lens", "--fov", "--rotation", "--size"):
        assert flag in photo


def test_bad_output_suffix_exits_1(tmp_path):
    src = os.path.join(GOLDEN, "equidistant.jpg")
    res = CliRunner().invoke(main, ["make-pano", "--type", "inscribed", "--lens", "equidistant", "--fov", "360",
                                    src, str(tmp_path / "out.bmp")])
    assert res.exit_code == 1 and "JPG or PNG" in res.output


def test_existing_output_prompts_and_n_exits_0(tmp_path):
    src = os.path.join(GOLDEN, "equidistant.jpg")
    out = tmp_path / "out.png"
    out.write_bytes(b"x")
    res = CliRunner().invoke(main, ["make-pano", "--type", "inscribed", "--lens", "equidistant", "--fov", "360",
                                    src, str(out)], input="n\n")
    assert res.exit_code == 0 and "Overwrite" in res.output
    assert out.read_bytes() == b"x"


def test_fov_rules_raise_value_error(tmp_path):
    src = os.path.join(GOLDEN, "equidistant.jpg")
    res = CliRunner().invoke(main, ["make-pano", "--type", "double", "--lens", "equidistant", "--fov", "170",
                                    src, str(tmp_path / "o.png")])
    assert isinstance(res.exception, ValueError)
    res = CliRunner().invoke(main, ["make-photo", "--type", "inscribed", "--lens", "equidistant", "--fov", "400",
                                    src, str(tmp_path / "o.png")])
    assert isinstance(res.exception, ValueError)


@pytest.mark.gpu
def test_make_pano_and_make_photo_match_oracle(tmp_path):
    from oracle import numpy_port
    from photonbend_b200.workloads import to_radians

    src_path = os.path.join(GOLDEN, "equidistant.jpg")
    with Image.open(src_path) as im:
        pixels = np.asarray(im)
    out = tmp_path / "pano.png"
    res = CliRunner().invoke(main, ["make-pano", "--type", "inscribed", "--lens", "equidistant", "--fov", "360",
                                    "-r", "10", "20", "30", "-s", "512", src_path, str(out)])
    assert res.exit_code == 0, res.output
    with Image.open(out) as im:
        got = np.asarray(im)
    og = {"kind": "equirect", "height": 512, "width": 1024}
    sg = {"kind": "camera", "height": 3072, "width": 3072, "lens": "equidistant", "fov": to_radians(360),
          "magnitude": 3072 / 2 - 0.5}
    want = numpy_port.remap(og, [(to_radians(10), to_radians(20), to_radians(30))], sg, pixels)
    assert np.array_equal(got, want)

    # panorama -> full-frame rectilinear photo (the CLI only makes square photos)
    photo = tmp_path / "photo.png"
    res = CliRunner().invoke(main, ["make-photo", "--type", "full", "--lens", "rectilinear", "--fov", "140",
                                    "-r", "-90", "0", "195", "-s", "400", str(out), str(photo)])
    assert res.exit_code == 0, res.output
    with Image.open(photo) as im:
        got2 = np.asarray(im)
    og2 = {"kind": "camera", "height": 400, "width": 400, "lens": "rectilinear", "fov": to_radians(140),
           "magnitude": float(np.sqrt(199.5**2 + 199.5**2))}
    want2 = numpy_port.remap(og2, [(to_radians(-90), 0.0, to_radians(195))], {"kind": "equirect", "height": 512, "width": 1024}, got)
    assert np.array_equal(got2, want2)
