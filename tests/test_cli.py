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


def test_directory_input_needs_a_directory_output(tmp_path):
    frames = tmp_path / "frames"
    frames.mkdir()
    Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(frames / "a.png")
    res = CliRunner().invoke(main, ["make-pano", "--type", "inscribed", "--lens", "equidistant", "--fov", "360",
                                    str(frames), str(tmp_path / "out.png")])
    assert res.exit_code == 1 and "must be a directory" in res.output
    empty = tmp_path / "empty"
    empty.mkdir()
    res = CliRunner().invoke(main, ["make-pano", "--type", "inscribed", "--lens", "equidistant", "--fov", "360",
                                    str(empty), str(tmp_path / "out")])
    assert res.exit_code == 1 and "no .jpg" in res.output


@pytest.mark.gpu
def test_directory_of_frames_matches_oracle(tmp_path, monkeypatch):
    """A directory as INPUT (photonbend_b200/stream.py: FramePipeline per GPU, frames sharded
    k mod G): every frame equals the oracle's remap of that frame; the compressed-stream variant
    (PHOTONBEND_B200_CODEC=nvjpeg: decode, remap and encode on the device) equals what the
    single-file command gives with the same codec."""
    from oracle import numpy_port
    from photonbend_b200.workloads import to_radians

    rng = np.random.default_rng(3)
    frames = tmp_path / "frames"
    frames.mkdir()
    sg = {"kind": "double", "height": 192, "width": 384, "lens": "equidistant", "fov": to_radians(195)}
    og = {"kind": "equirect", "height": 160, "width": 320}
    images = []
    for k in range(7):
        img = rng.integers(0, 256, (192, 384, 3), dtype=np.uint8)
        images.append(img)
        Image.fromarray(img).save(frames / f"f{k:03d}.png")
    out = tmp_path / "out"
    res = CliRunner().invoke(main, ["make-pano", "--type", "double", "--lens", "equidistant", "--fov", "195", "-s", "160",
                                    "--frames-per-launch", "3", str(frames), str(out)])
    assert res.exit_code == 0, res.output
    assert "7 frames" in res.output
    for k in range(7):
        with Image.open(out / f"f{k:03d}.png") as im:
            got = np.asarray(im)
        assert np.array_equal(got, numpy_port.remap(og, (), sg, images[k])), k
    # rotated, one frame per launch, existing outputs: prompt once
    res = CliRunner().invoke(main, ["make-pano", "--type", "double", "--lens", "equidistant", "--fov", "195", "-s", "160",
                                    "-r", "10", "20", "30", "--frames-per-launch", "1", str(frames), str(out)], input="y\n")
    assert res.exit_code == 0 and "Overwrite" in res.output, res.output
    rot = [(to_radians(10), to_radians(20), to_radians(30))]
    for k in (0, 6):
        with Image.open(out / f"f{k:03d}.png") as im:
            got = np.asarray(im)
        want = numpy_port.remap(og, rot, sg, images[k])
        diff = (got != want).any(axis=2)
        # a rotated double-fisheye source: the 1-LSB blend-truncation class only, within the parity budget
        assert diff.sum() <= 1e-4 * diff.size and np.abs(got.astype(int) - want.astype(int)).max() <= 1, (k, int(diff.sum()))

    # compressed stream: JPEG frames, nvJPEG on the device
    monkeypatch.setenv("PHOTONBEND_B200_CODEC", "nvjpeg")
    jframes = tmp_path / "jframes"
    jframes.mkdir()
    smooth = np.stack(np.meshgrid(np.arange(384), np.arange(192)), axis=2).sum(axis=2)
    for k in range(3):
        img = np.stack([(smooth * (k + 1)) % 256, (smooth // 2) % 256, (smooth * 3) % 256], axis=2).astype(np.uint8)
        Image.fromarray(img).save(jframes / f"j{k}.jpg", quality=92)
    jout = tmp_path / "jout"
    res = CliRunner().invoke(main, ["make-pano", "--type", "double", "--lens", "equidistant", "--fov", "195", "-s", "160",
                                    "--frames-per-launch", "2", str(jframes), str(jout)])
    assert res.exit_code == 0 and "codec nvjpeg" in res.output, res.output
    for k in range(3):
        single = tmp_path / f"single{k}.jpg"
        res = CliRunner().invoke(main, ["make-pano", "--type", "double", "--lens", "equidistant", "--fov", "195", "-s", "160",
                                        str(jframes / f"j{k}.jpg"), str(single)])
        assert res.exit_code == 0, res.output
        assert (jout / f"j{k}.jpg").read_bytes() == single.read_bytes(), k
